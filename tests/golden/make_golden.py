"""Convert the reference's golden regression files into small .npz fixtures.

Run once in the build container (where /root/reference is mounted):
    python tests/golden/make_golden.py
The GPU box has no /root/reference, so the fixtures are committed.  Inputs:
/root/reference/tests/references/*.hdf5 (produced by the reference's CI build, see
/root/reference/tests/README.md).  h5py is not installed here, hence oracle/hdf5_min.py.

To keep the fixtures small the 32^3 f64 grids of test_2lpt_baryons_generic are stored once
(DM_*) and the BA_* copies - bit-identical for the Eisenstein&Hu transfer function, which has
delta_bc = theta_bc = 0 (src/plugins/transfer_eisenstein.cc:240) - are replaced by a flag that
records the verified bit-equality.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
from hdf5_min import H5File  # noqa: E402

REF = "/root/reference/tests/references"


def main():
    for name in ["test_1lpt_sc_generic", "test_2lpt_baryons_generic", "test_2lpt_sc_gadget",
                 "test_2lpt_baryons_vrel_gadget", "test_3lpt_bcc_swift"]:
        h = H5File(os.path.join(REF, name + ".hdf5"))
        out = {}
        for p, kind in h.walk():
            if kind == "dataset":
                out[p.strip("/").replace("/", "__")] = h.dataset(p)
        for grp in ["/Header"]:
            for k, v in h.attrs(grp).items():
                if isinstance(v, (np.ndarray, np.generic)):
                    out["attr__Header__" + k] = np.asarray(v)
        if name == "test_2lpt_baryons_generic":
            same = True
            for c in ["dx", "dy", "dz", "vx", "vy", "vz", "delta"]:
                same &= bool(np.array_equal(out["BA_" + c].view(np.uint64), out["DM_" + c].view(np.uint64)))
                del out["BA_" + c]
            out["flag__BA_equals_DM_bitwise"] = np.array(same)
            # mass grids are constant: keep min/max/first only
            for c in ["BA_", "DM_"]:
                a = out.pop(c)
                out["massgrid__" + c] = np.array([a.min(), a.max(), a.ravel()[0]])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: (v.dtype.str, v.shape) for k, v in out.items() if not k.startswith("attr")})


if __name__ == "__main__":
    main()
