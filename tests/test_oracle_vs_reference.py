"""Pins the CPU oracle against the REFERENCE'S OWN code for the quantities the reference's regression files do not cover.

* always: oracle vs tests/golden/ref_intermediates_32.npz (made by tests/golden/make_ref_fixtures.py from the reference's
  Grid_FFT / OrszagConvolver / op:: sources, see oracle/Makefile and oracle/refshim/);
* where /root/reference is mounted (the build container): oracle/_ref/ref_driver is rebuilt from the reference sources and
  run live on a second, different pair of fields."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")]
import lpt_oracle as O  # noqa: E402

TOL = 1e-12


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def compare(R, phi, psi, n, L, D):
    g = O.GridSpec(n, L)
    cv = O.OrszagConvolver(g)
    x = g.bwd(phi)
    assert rel(x, R["phi_real"]) < TOL
    assert abs(x.mean() - R["phi_real_mean_std"][0]) < 1e-14 * np.abs(x).max()
    assert abs(x.std() - R["phi_real_mean_std"][1]) < 1e-13 * x.std()
    xs = O.shift_field_r(g, x, (0.5, 0.25, 0.5))
    assert rel(xs, R["phi_real_shifted"]) < TOL
    assert rel(O.inverse_laplacian(g, g.fwd(xs)), R["invlap_shifted_phi"]) < TOL
    assert rel(O.fourier_interpolate_copy(g, O.GridSpec(3 * n // 2, L), psi), R["psi_interp_3half"]) < TOL
    assert rel(cv.convolve_Hessians(phi, (0, 1), psi, (1, 2)), R["hess_01_12"]) < TOL
    assert rel(cv.convolve_SumOfHessians(phi, (0, 0), psi, (1, 1), (2, 2)), R["sumhess_00_11_22"]) < TOL
    assert rel(cv.convolve_DifferenceOfHessians(phi, (1, 2), psi, (1, 1), (2, 2)), R["diffhess_12_11_22"]) < TOL
    assert rel(cv.convolve_Hessians3(phi, (0, 0), psi, (0, 2), phi, (2, 2)), R["hess3_00_02_22"]) < TOL
    assert rel(cv.convolve_Gradients(phi, 0, psi, 2), R["grad_0_2"]) < TOL
    assert rel(cv.convolve_Gradient_and_Hessian(phi, 1, psi, (0, 2)), R["gradhess_1_02"]) < TOL
    m = cv.multiply_field(phi, psi)
    assert rel(m, R["multiply"]) < TOL
    assert rel(1.5 * m - cv.multiply_field(phi, phi) + cv.multiply_field(psi, psi), R["multiply_ops"]) < TOL
    pots = O.lpt_potentials(g, phi, 3, D)
    assert rel(pots["phi"], R["phi_scaled"]) < TOL and rel(pots["phi2"], R["phi2"]) < TOL
    assert rel(pots["phi3"], R["phi3"]) < 10 * TOL
    for d, nm in enumerate(("A3x", "A3y", "A3z")):
        assert rel(pots["A3"][d], R[nm]) < 100 * TOL


def compare_naive_and_complex(R, a, b, c, n, L):
    g = O.GridSpec(n, L)
    nv = O.NaiveConvolver(g)
    if "naive_hess_01_12" in R:
        assert rel(nv.convolve_Hessians(a, (0, 1), b, (1, 2)), R["naive_hess_01_12"]) < TOL
        assert rel(nv.convolve_SumOfHessians(a, (0, 0), b, (1, 1), (2, 2)), R["naive_sumhess_00_11_22"]) < TOL
        assert rel(nv.convolve_Gradients(a, 0, b, 2), R["naive_grad_0_2"]) < TOL
    assert rel(nv.convolve_Hessians3(a, (0, 0), b, (0, 2), a, (2, 2)), R["naive_hess3_00_02_22"]) < TOL
    assert rel(nv.convolve_Gradient_and_Hessian(a, 1, b, (0, 2)), R["naive_gradhess_1_02"]) < TOL
    assert rel(nv.multiply_field(a, b), R["naive_multiply"]) < TOL
    assert rel(np.fft.fftn(c) * g.norm, R["cplx_k"]) < TOL        # c2c forward of Grid_FFT<ccomplex_t>: sign and norm


def compare_mpi_nyquist(R, a, b, n, L, D, prefix=""):
    """the reference's USE_MPI code path on fields with populated Nyquist planes: the oracle restates exactly these
    semantics (pad / unpad through FourierInterpolateCopyTo, Nyquist planes dropped)"""
    g = O.GridSpec(n, L)
    cv = O.OrszagConvolver(g)
    xs = O.shift_field_r(g, g.bwd(a), (0.5, 0.25, 0.5))
    assert rel(xs, R[prefix + "phi_real_shifted"]) < TOL
    assert rel(O.inverse_laplacian(g, g.fwd(xs)), R[prefix + "invlap_shifted_phi"]) < TOL
    assert rel(cv.convolve_Hessians3(a, (0, 0), b, (0, 2), a, (2, 2)), R[prefix + "hess3_00_02_22"]) < TOL
    assert rel(cv.convolve_Gradient_and_Hessian(a, 1, b, (0, 2)), R[prefix + "gradhess_1_02"]) < TOL
    m = cv.multiply_field(a, b)
    assert rel(1.5 * m - cv.multiply_field(a, a) + cv.multiply_field(b, b), R[prefix + "multiply_ops"]) < TOL
    pots = O.lpt_potentials(g, a, 3, D)
    assert rel(pots["phi3"], R[prefix + "phi3"]) < 10 * TOL
    assert rel(pots["A3"][0], R[prefix + "A3x"]) < 100 * TOL


def oracle_particles(g, phi, psi, lt, li, zero=False):
    x, v = g.bwd(phi), g.bwd(psi)
    if zero:
        x, v = np.zeros_like(x), np.zeros_like(v)
    return {"ids": O.particle_ids(g.n, lt, li, True),
            "pos": np.stack([O.set_positions(g, x, d, lt, li, g.L, True) for d in range(3)], 1),
            "vel": np.stack([O.set_velocities(g, v, lt, li, True) for d in range(3)], 1),
            "mass": O.set_masses(g, x, lt, li, 2.5, True)}


def test_oracle_particles_match_committed_reference_fixture():
    """particle::lattice_generator of the reference (MPI build): ids and bare-lattice positions bit-exact (digests),
    displaced positions / velocities / masses on every 101st particle, all four Bravais lattices, both lattice indices"""
    import hashlib

    R = np.load(os.path.join(ROOT, "tests", "golden", "ref_intermediates_32.npz"))
    n, L = int(R["meta"][0]), float(R["meta"][1])
    g = O.GridSpec(n, L)
    for lt, nm in enumerate(("sc", "bcc", "fcc", "rsc")):
        for li in (0, 1):
            Z = oracle_particles(g, R["phi_in"], R["psi_in"], lt, li, zero=True)
            assert hashlib.sha256(Z["ids"].tobytes()).digest() == R[f"part_{nm}{li}_ids_sha256"].tobytes()
            assert hashlib.sha256(Z["pos"].tobytes()).digest() == R[f"part_{nm}{li}_lattice_sha256"].tobytes()
            P = oracle_particles(g, R["phi_in"], R["psi_in"], lt, li)
            assert np.abs(P["pos"][::101] - R[f"part_{nm}{li}_pos_sub"]).max() < 1e-13 * L
            assert rel(P["vel"][::101], R[f"part_{nm}{li}_vel_sub"]) < TOL
            assert rel(P["mass"][::101], R[f"part_{nm}{li}_mass_sub"]) < TOL


def oracle_masked(g, phi, psi, do_b, mt, li):
    x, v = g.bwd(phi), g.bwd(psi)
    mask = O.particle_type_mask(do_b, mt)
    return {"ids": O.masked_ids(g.n, mask, li, li, True),
            "pos": np.stack([O.masked_positions(g, x, d, mask, li, g.L, True) for d in range(3)], 1),
            "vel": np.stack([O.masked_velocities(g, v, mask, li, True) for d in range(3)], 1),
            "mass": O.masked_masses(g, x, mask, li, 2.5, True)}


def compare_cic(g, phi, pts, out):
    """oracle restatements of Grid_FFT::get_cic_kspace, get_k<real_t>, grid_interpolate::get_cic_at and compensation_kernel
    against what the reference's own headers return (tests/golden/make_ref_fixtures.py::run_reference_cic)"""
    n = g.n
    lgr = O.PLTGradient(n, n, g.L, phi.real, phi.real, phi.real, phi.imag)      # any tables: cic_kspace is linear in them
    x, y, zk = pts[:, 0], pts[:, 1], pts[:, 2] * float(n // 2) / float(n)
    assert np.abs(lgr.cic_kspace(phi.real, x, y, zk) - out[:, 0]).max() < 1e-13 * np.abs(out[:, 0]).max()
    assert np.abs(lgr.cic_kspace(phi.imag, x, y, zk) - out[:, 1]).max() < 1e-13 * np.abs(out[:, 1]).max()
    h = float(n // 2)
    for c, w in enumerate((x, y, zk)):
        assert np.abs((w - (w > h) * n) * g.kfac - out[:, 2 + c]).max() < 1e-13 * np.abs(out[:, 2 + c]).max()
    fr = g.bwd(phi)
    assert np.abs(O.cic_at(fr, pts) - out[:, 5]).max() < 1e-13 * np.abs(fr).max()
    comp = O.cic_compensation_kernel(g)
    i, j, k = (pts[:, c].astype(int) for c in range(3))
    k = k % (n // 2 + 1)
    assert np.abs(np.broadcast_to(comp, (n, n, n // 2 + 1))[i, j, k] - out[:, 6]).max() < 1e-13 * np.abs(out[:, 6]).max()


def test_oracle_f4_matches_committed_reference_fixture():
    """the masked sc lattice of the reference's particle::lattice_generator (ids bit-exact, positions /
    velocities / masses with get_mean_mask_value) and the CIC building blocks of the PLT gradient and of glass loads"""
    import hashlib
    import make_ref_fixtures as F

    R = np.load(os.path.join(ROOT, "tests", "golden", "ref_f4_32.npz"))
    M0 = np.load(os.path.join(ROOT, "tests", "golden", "ref_intermediates_32.npz"))
    n, L = int(R["meta"][0]), float(R["meta"][1])
    g = O.GridSpec(n, L)
    for do_b, mt, li in F.MASKED_CASES:
        key = f"masked_{int(do_b)}{mt}{li}"
        Q = oracle_masked(g, M0["phi_in"], M0["psi_in"], do_b, mt, li)
        assert len(Q["ids"]) == int(R[key + "_np"][0])
        assert hashlib.sha256(Q["ids"].tobytes()).digest() == R[key + "_ids_sha256"].tobytes()
        assert np.abs(Q["pos"][::37] - R[key + "_pos_sub"]).max() < 1e-13 * L
        assert rel(Q["vel"][::37], R[key + "_vel_sub"]) < TOL
        assert rel(Q["mass"][::37], R[key + "_mass_sub"]) < TOL
    compare_cic(g, M0["phi_in"], R["cic_points"], R["cic_out"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/include") or shutil.which("g++") is None,
                    reason="the reference sources are only mounted in the build container")
def test_oracle_f4_matches_reference_built_here():
    import make_ref_fixtures as F

    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "all"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    n = 16
    g = O.GridSpec(n, F.L)
    amp = O.power_law_amplitude(ns=0.93, kcut=0.8)
    phi = O.phi1_from_noise(g, O.condition_noise(g, O.make_white_noise_k(n, 17, True)), amp)
    psi = O.phi1_from_noise(g, O.condition_noise(g, O.make_white_noise_k(n, 18, True)), amp)
    for do_b, mt, li in ((True, 3, 1), (True, 0, 0), (False, 3, 0)):
        P = F.run_reference_particles(phi, psi, -2, li, n=n, do_baryons=do_b, mask_type=mt)
        Q = oracle_masked(g, phi, psi, do_b, mt, li)
        assert np.array_equal(P["ids"], Q["ids"])
        assert np.abs(P["pos"] - Q["pos"]).max() < 1e-13 * F.L
        assert rel(Q["vel"], P["vel"]) < TOL and rel(Q["mass"], P["mass"]) < TOL
    pts = F.cic_points(n, 200, 9)
    compare_cic(g, phi, pts, F.run_reference_cic(phi, pts, n))


def test_oracle_matches_committed_reference_fixture():
    R = np.load(os.path.join(ROOT, "tests", "golden", "ref_intermediates_32.npz"))
    compare(R, R["phi_in"], R["psi_in"], int(R["meta"][0]), float(R["meta"][1]), float(R["meta"][2]))
    compare_naive_and_complex(R, R["naive_a_in"], R["naive_b_in"], R["cplx_in"], int(R["meta"][0]), float(R["meta"][1]))
    compare_mpi_nyquist(R, R["naive_a_in"], R["naive_b_in"], int(R["meta"][0]), float(R["meta"][1]), float(R["meta"][2]), "mpinyq_")


@pytest.mark.skipif(not os.path.isdir("/root/reference/include") or shutil.which("g++") is None,
                    reason="the reference sources are only mounted in the build container")
def test_oracle_matches_reference_built_here():
    import make_ref_fixtures as F

    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "all"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    n = 32
    g = O.GridSpec(n, F.L)
    amp = O.power_law_amplitude(ns=0.93, kcut=0.8)
    phi = O.phi1_from_noise(g, O.condition_noise(g, O.make_white_noise_k(n, 7, True)), amp)
    psi = O.phi1_from_noise(g, O.condition_noise(g, O.make_white_noise_k(n, 8, True)), amp)
    a, b, c = F.nyquist_inputs(n, (31, 41, 59))
    R = F.run_reference(phi, psi, n, extra=(a, b, c))
    compare(R, phi, psi, n, F.L, F.DPLUS)
    compare_naive_and_complex(R, a, b, c, n, F.L)
    assert rel(R["cplx_roundtrip"], c) < TOL
    # the USE_MPI build on one rank: same results on the Nyquist-free pair, and the Nyquist-carrying pair
    Rm = F.run_reference(phi, psi, n, extra=(a, b, c), mpi=True)
    compare(Rm, phi, psi, n, F.L, F.DPLUS)
    compare_naive_and_complex(Rm, a, b, c, n, F.L)
    Rn = F.run_reference(a, b, n, mpi=True)
    compare(Rn, a, b, n, F.L, F.DPLUS)
    compare_mpi_nyquist(Rn, a, b, n, F.L, F.DPLUS)
    # the reference's particle::lattice_generator, live, on these fresh fields: two of the eight lattice configurations
    for lt, li in ((2, 1), (3, 0)):
        P = F.run_reference_particles(phi, psi, lt, li, n=n)
        Q = oracle_particles(g, phi, psi, lt, li)
        assert np.array_equal(P["ids"], Q["ids"])
        assert np.abs(P["pos"] - Q["pos"]).max() < 1e-13 * F.L
        assert rel(Q["vel"], P["vel"]) < TOL and rel(Q["mass"], P["mass"]) < TOL
        Z = F.run_reference_particles(phi, psi, lt, li, zero_field=True, n=n)
        assert np.array_equal(Z["pos"], oracle_particles(g, phi, psi, lt, li, zero=True)["pos"])


def test_fast_fft_backend(tmp_path):
    """oracle/_ref/ref_bench - the reference's Grid_FFT / OrszagConvolver sources behind the FAST FFT shim
    (oracle/refshim/fftw_fast.cc), the binary bench.py times as its CPU arm - computes what the oracle computes: the shim's
    mixed-radix transforms (2^a 3^b lengths, two-for-one r2c / c2r) are not an exact DFT, so they get a check of their own"""
    if not os.path.isdir("/root/reference/include"):
        pytest.skip("reference sources not mounted: ref_bench cannot be (re)built here")
    r = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref/ref_bench"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    n, L = 48, 100.0           # 48 and the padded 72 exercise the radix-3 and radix-4/2 paths
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_bench"), str(n), "3", "1", "0", str(tmp_path)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "seconds_per_step" in r.stdout, r.stdout + r.stderr
    ld = lambda nm: np.fromfile(os.path.join(tmp_path, nm + ".bin"), dtype=np.complex128).reshape(n, n, n // 2 + 1)
    phi = ld("phi")
    g = O.GridSpec(n, L)
    pots = O.lpt_potentials(g, phi, 3, -1.0)     # D+ = -1: g1 = 1, g2 = -3/7, g3 = -1/3, g3c = -1/7
    assert rel(ld("phi2"), pots["phi2"] / (-3.0 / 7.0)) < 1e-11
    assert rel(ld("phi3"), pots["phi3"] / (-1.0 / 3.0)) < 1e-10
    assert rel(ld("A3x"), pots["A3"][0] / (-1.0 / 7.0)) < 1e-9
