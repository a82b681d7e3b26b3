"""Shared helpers to use the reference's five golden files without GSL / ranlxd1 / E&H
(SURVEY.md Appendix D): all five test configs share N=32, L=100, seed, redshift and cosmology,
hence the same first-order potential.  The generic golden stores delta = bwd[k^2 * g1*phi], so
phi_s = g1*phi is recovered exactly; LPT is homogeneous in phi, so running the reference
formulas on phi_s with D+ := -1 yields the growth-scaled higher-order terms of the real run.
vfac is the element-wise ratio DM_vx/DM_dx of the 1LPT generic golden."""
import os

import numpy as np

import lpt_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N, L = 32, 100.0
ASTART = 1.0 / 51.0
HUBBLE = 0.67742
RTOL, ATOL = 1e-9, 1e-14  # tests/scripts/compare_hdf5.py:259-262


def load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def bootstrap():
    """-> (GridSpec, phi_s in k-space [kx][ky][kz], vfac)"""
    g = O.GridSpec(N, L)
    g1 = load("test_1lpt_sc_generic")
    dk = g.fwd(g1["DM_delta"])
    k2 = g.k2()
    with np.errstate(divide="ignore", invalid="ignore"):
        phi_s = dk / k2
    phi_s[0, 0, 0] = 0.0
    m = np.abs(g1["DM_dx"]) > 1e-4 * np.abs(g1["DM_dx"]).max()
    vfac = float(np.median(g1["DM_vx"][m] / g1["DM_dx"][m]))
    return g, phi_s, vfac


def units(fmt):
    """(lunit, vunit, real64) per output plugin: output_generic.cc:70-78,
    output_gadget_hdf5.cc:83-84, output_swift.cc:73-74"""
    if fmt == "generic":
        return 1.0, 1.0, True
    if fmt == "gadget":
        return L, L / np.sqrt(ASTART), False
    if fmt == "swift":
        return L / HUBBLE, L, True
    raise ValueError(fmt)
