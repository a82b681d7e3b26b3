"""Digest of the CPU oracle for bench.py's 256^3 parity leg (run in the build container; the digest travels, the oracle run
does not have to): the bench's synthetic 3LPT workload - same per-plane noise seeds, amplitude, growth factors and units - run
through oracle/lpt_oracle.py; the FP64 displacement and velocity fields are sampled at bench.sample_cells(n).

  python tests/golden/make_bench_digest.py [n=256] [order=3]   ->  tests/golden/bench_digest_oracle_<n>.npz  (order 3)
                                                                     tests/golden/bench_digest_oracle_<n>_<order>lpt.npz (else)
The 2LPT 512^3 digest takes about 25 GB of host memory.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import bench as B  # noqa: E402
import lpt_oracle as O  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    order, L = (int(sys.argv[2]) if len(sys.argv) > 2 else 3), B.LBOX
    t0 = time.time()
    x = np.empty((n, n, n))
    for i in range(n):
        x[i] = B.noise_plane(n, i)
    fft = None
    try:
        fft = O.TorchFFTBackend(os.cpu_count() or 1)
    except Exception:
        pass
    g = O.GridSpec(n, L, fft) if fft is not None else O.GridSpec(n, L)
    W = g.fwd(x)                                  # what Grid_FFT::FourierTransformForward leaves (symmetric norm)
    Wc = O.condition_noise(g, W)
    lunit, vunit = L, L / np.sqrt(1 / 51.0)       # bench.Bench.cfg
    kw = {"fft": fft} if fft is not None else {}
    ref = O.run_pipeline(n, L, Wc, B.synthetic_amplitude(), order, B.DPLUS, B.VFAC, "sc", lunit, vunit, False, False, **kw)
    cells = B.sample_cells(n)
    names = ["dx", "dy", "dz", "vx", "vy", "vz"]
    fields = np.stack([ref["fields"][k][cells[:, 0], cells[:, 1], cells[:, 2]] for k in names])
    maxabs = np.array([np.abs(ref["fields"][k]).max() for k in names])
    out = os.path.join(ROOT, "tests", "golden", f"bench_digest_oracle_{n}.npz" if order == 3 else f"bench_digest_oracle_{n}_{order}lpt.npz")
    np.savez_compressed(out, n=n, order=order, fields=fields, maxabs=maxabs, cells=cells.astype(np.int32))
    print(f"wrote {out}: {fields.shape}, max|Psi| = {maxabs[:3].max():.4g}, max|v| = {maxabs[3:].max():.4g}, {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
