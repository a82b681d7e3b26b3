"""GPU diagnostics: golden comparison statistics (run on the box)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import golden_util as G
import parity_cases as PC
import monofonic_b200 as M
M.init(0)

def stats(name, a, b):
    err = np.abs(a.astype(np.float64) - b.astype(np.float64))
    tol = G.ATOL + G.RTOL * np.abs(b)
    bad = err > tol
    print(f"{name:28s} max|err|={err.max():.3e} max|ref|={np.abs(b).max():.3e} rel={err.max()/np.abs(b).max():.2e} "
          f"n_bad={int(bad.sum())}/{bad.size} worst_ratio={(err/tol).max():.2f} min|ref| among bad="
          f"{(np.abs(b)[bad].min() if bad.any() else 0):.2e}")

gold = G.load("test_3lpt_bcc_swift")
pl = PC._golden_pipeline(M, 3, "bcc", "swift")
p = pl.fill_particles(0).download()
stats("swift pos", p["pos"], gold["PartType1__Coordinates"])
stats("swift vel", p["vel"], gold["PartType1__Velocities"])
for name, order in [("test_1lpt_sc_generic", 1), ("test_2lpt_baryons_generic", 2)]:
    gold = G.load(name)
    pl = PC._golden_pipeline(M, order, "sc", "generic")
    for d, c in enumerate("xyz"):
        stats(f"{name[5:15]} d{c}", pl.displacement_field(d).get_real(), gold["DM_d" + c])
    for d, c in enumerate("xyz"):
        stats(f"{name[5:15]} v{c}", pl.velocity_field(d).get_real(), gold["DM_v" + c])
    stats(f"{name[5:15]} delta", pl.density_field().get_real(), gold["DM_delta"])
