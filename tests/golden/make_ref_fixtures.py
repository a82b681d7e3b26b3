"""Golden vectors produced by the REFERENCE'S OWN code for the quantities its regression files do not cover.

Run in the build container (where /root/reference is mounted):
    make -C oracle && python tests/golden/make_ref_fixtures.py
`oracle/_ref/ref_driver` is the reference's Grid_FFT / OrszagConvolver / op:: sources (include/grid_fft.hh,
include/convolution.hh, include/operators.hh, src/grid_fft.cc) compiled where they lie against the shims of oracle/refshim/
(an exact-DFT stand-in for FFTW3, an inert hdf5.h, a hand-written cmake_config.hh); oracle/refshim/ref_driver.cc feeds it a
first-order potential and dumps every named product, the phi2 / phi3 / A3 source stage of src/ic_generator.cc:475-574, the
Poisson solve, shift_field and FourierInterpolateCopyTo.  The result, tests/golden/ref_intermediates_32.npz, travels to the
GPU box; tests/test_oracle_vs_reference.py pins the oracle against it and tests/test_gpu_parity.py the CUDA path.
Inputs are Nyquist-free (like NGenIC noise), where the non-MPI pad path of this build and the MPI path the goldens come
from coincide."""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lpt_oracle as O  # noqa: E402

N, L, DPLUS = 32, 100.0, 0.0257
KNAMES = ["invlap_shifted_phi", "hess_01_12", "sumhess_00_11_22", "diffhess_12_11_22", "hess3_00_02_22", "grad_0_2",
          "gradhess_1_02", "multiply", "multiply_ops", "phi_scaled", "phi2", "phi3", "A3x", "A3y", "A3z"]
RNAMES = ["phi_real", "phi_real_shifted"]
NAIVE = ["naive_hess_01_12", "naive_sumhess_00_11_22", "naive_hess3_00_02_22", "naive_grad_0_2", "naive_gradhess_1_02",
         "naive_multiply"]
NAIVE_STORED = ["naive_hess3_00_02_22", "naive_gradhess_1_02", "naive_multiply"]   # the committed fixture keeps three
# from the USE_MPI build run on the Nyquist-CARRYING pair (stored with the prefix "mpinyq_")
MPI_NYQ_STORED = ["phi_real_shifted", "invlap_shifted_phi", "hess3_00_02_22", "gradhess_1_02", "multiply_ops", "phi3", "A3x"]


def inputs(n=N):
    g = O.GridSpec(n, L)
    amp = O.power_law_amplitude()
    phi = O.phi1_from_noise(g, O.condition_noise(g, O.make_white_noise_k(n, 101, True)), amp)
    psi = O.phi1_from_noise(g, O.condition_noise(g, O.make_white_noise_k(n, 202, True)), amp)
    return g, phi, psi


def nyquist_inputs(n=N, seeds=(303, 404, 505)):
    """two half-spectra whose Nyquist planes are populated (NaiveConvolver) and a complex real-space field (c2c)"""
    g = O.GridSpec(n, L)
    amp = O.power_law_amplitude()
    a = O.phi1_from_noise(g, O.condition_noise(g, O.make_white_noise_k(n, seeds[0], False)), amp)
    b = O.phi1_from_noise(g, O.condition_noise(g, O.make_white_noise_k(n, seeds[1], False)), amp)
    rng = np.random.default_rng(seeds[2])
    c = rng.standard_normal((n, n, n)) + 1j * rng.standard_normal((n, n, n))
    return a, b, c


def run_reference(phi, psi, n=N, extra=None, mpi=False):
    """mpi=True: the USE_MPI build of the reference on one rank (oracle/_ref/ref_driver_mpi)"""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_driver_mpi" if mpi else "ref_driver")
    with tempfile.TemporaryDirectory() as d:
        np.ascontiguousarray(phi, np.complex128).tofile(os.path.join(d, "phi.bin"))
        np.ascontiguousarray(psi, np.complex128).tofile(os.path.join(d, "psi.bin"))
        cmd = [exe, str(n), repr(L), repr(DPLUS), os.path.join(d, "phi.bin"), os.path.join(d, "psi.bin"), d]
        if extra is not None:
            for nm, arr in zip(("a.bin", "b.bin", "c.bin"), extra):
                np.ascontiguousarray(arr, np.complex128).tofile(os.path.join(d, nm))
                cmd.append(os.path.join(d, nm))
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0 and "ref_driver ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
        out = {}
        nzc, p = n // 2 + 1, 3 * n // 2
        for k in KNAMES:
            out[k] = np.fromfile(os.path.join(d, k + ".bin"), np.complex128).reshape(n, n, nzc)
        for k in RNAMES:
            out[k] = np.fromfile(os.path.join(d, k + ".bin"), np.float64).reshape(n, n, n)
        out["psi_interp_3half"] = np.fromfile(os.path.join(d, "psi_interp_3half.bin"), np.complex128).reshape(p, p, p // 2 + 1)
        if extra is not None:
            for k in NAIVE:
                out[k] = np.fromfile(os.path.join(d, k + ".bin"), np.complex128).reshape(n, n, nzc)
            out["cplx_k"] = np.fromfile(os.path.join(d, "cplx_k.bin"), np.complex128).reshape(n, n, n)
            out["cplx_roundtrip"] = np.fromfile(os.path.join(d, "cplx_roundtrip.bin"), np.complex128).reshape(n, n, n)
        for ln in r.stdout.splitlines():
            if ln.startswith("mean"):
                f = ln.split()
                out["phi_real_mean_std"] = np.array([float(f[1]), float(f[3])])
    return out


def run_reference_particles(phi, psi, lattice_type, lattice_index, zero_field=False, n=N, do_baryons=None, mask_type=3):
    """the reference's particle::lattice_generator (MPI build, one rank): positions from phi, velocities from psi, masses
    from phi (munit 2.5), 64-bit reals and ids, IDoffset = lattice_index.  -> dict(pos, vel, mass, ids)"""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_driver_mpi")
    with tempfile.TemporaryDirectory() as d:
        np.ascontiguousarray(phi, np.complex128).tofile(os.path.join(d, "phi.bin"))
        np.ascontiguousarray(psi, np.complex128).tofile(os.path.join(d, "psi.bin"))
        r = subprocess.run([exe, "particles", str(n), repr(L), os.path.join(d, "phi.bin"), os.path.join(d, "psi.bin"), d,
                            str(lattice_type), str(lattice_index), "1" if zero_field else "0"] +
                           ([("yes" if do_baryons else "no"), str(mask_type)] if do_baryons is not None else []),
                           capture_output=True, text=True)
        assert r.returncode == 0 and "particles ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
        raw = np.fromfile(os.path.join(d, "particles.bin"), np.uint8)
    npart = int(raw[:8].view(np.uint64)[0])
    b = raw[8:]
    return {"pos": b[:24 * npart].view(np.float64).reshape(npart, 3).copy(),
            "vel": b[24 * npart:48 * npart].view(np.float64).reshape(npart, 3).copy(),
            "mass": b[48 * npart:56 * npart].view(np.float64).copy(),
            "ids": b[56 * npart:64 * npart].view(np.uint64).copy()}


def run_reference_cic(phi, points, n=N):
    """the interpolation building blocks of the PLT gradient and of glass loads from the reference's own headers (non-MPI
    build): per point (index units) [Re, Im get_cic_kspace(x, y, z/2), get_k<real_t>(x, y, z/2) x3, get_cic_at(x, y, z) on the
    backward-transformed field, compensation_kernel(get_k(floor x, floor y, floor z % (n/2+1)))]"""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    with tempfile.TemporaryDirectory() as d:
        np.ascontiguousarray(phi, np.complex128).tofile(os.path.join(d, "phi.bin"))
        np.ascontiguousarray(points, np.float64).tofile(os.path.join(d, "pts.bin"))
        r = subprocess.run([exe, "cic", str(n), repr(L), os.path.join(d, "phi.bin"), os.path.join(d, "pts.bin"), str(len(points)),
                            os.path.join(d, "out.bin")], capture_output=True, text=True)
        assert r.returncode == 0 and "cic ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
        return np.fromfile(os.path.join(d, "out.bin"), np.float64).reshape(len(points), 7)


MASKED_CASES = ((True, 3, 0), (True, 3, 1), (True, 1, 0), (True, 2, 1), (False, 3, 0))


def cic_points(n=N, npts=400, seed=5):
    rng = np.random.default_rng(seed)
    pts = rng.random((npts, 3)) * n
    pts[0] = (n - 0.25, n - 0.5, n - 0.125)     # wraps on every axis
    pts[1] = (0.0, 0.0, 0.0)
    pts[2] = (n / 2 + 0.5, n / 2, n / 2 - 0.5)  # around the Nyquist index
    return pts


def main_f4():
    """tests/golden/ref_f4_32.npz: SURVEY 8 f-4 quantities from the reference's own code - the masked sc lattice of
    particle::lattice_generator (MPI build, one rank) and the CIC building blocks of the PLT gradient / glass loads"""
    import hashlib
    g, phi, psi = inputs()
    out = {"meta": np.array([N, L])}
    for do_b, mt, li in MASKED_CASES:
        P = run_reference_particles(phi, psi, -2, li, do_baryons=do_b, mask_type=mt)
        key = f"masked_{int(do_b)}{mt}{li}"
        out[key + "_np"] = np.array([len(P["ids"])])
        out[key + "_ids_sha256"] = np.frombuffer(hashlib.sha256(P["ids"].tobytes()).digest(), np.uint8)
        for k in ("pos", "vel", "mass"):
            out[f"{key}_{k}_sub"] = P[k][::37].copy()
    pts = cic_points()
    out["cic_points"] = pts
    out["cic_out"] = run_reference_cic(phi, pts)
    path = os.path.join(HERE, "ref_f4_32.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KB,", len(out), "arrays")


def main():
    g, phi, psi = inputs()
    a, b, c = nyquist_inputs()
    out = run_reference(phi, psi, extra=(a, b, c))
    assert np.abs(out.pop("cplx_roundtrip") - c).max() < 1e-13     # the reference's own c2c round trip; not stored
    for k in NAIVE:
        if k not in NAIVE_STORED:
            out.pop(k)
    out["phi_in"], out["psi_in"] = phi, psi
    out["naive_a_in"], out["naive_b_in"], out["cplx_in"] = a, b, c
    # the MPI code path of the reference (one rank) on the Nyquist-carrying pair: pad / unpad drop the Nyquist planes
    Rm = run_reference(a, b, mpi=True)
    for k in MPI_NYQ_STORED:
        out["mpinyq_" + k] = Rm[k]
    # and the MPI build must reproduce the non-MPI build on the Nyquist-free pair
    Rf = run_reference(phi, psi, mpi=True)
    for k in KNAMES + RNAMES:
        assert np.abs(Rf[k] - out[k]).max() <= 1e-13 * np.abs(out[k]).max(), k
    out["meta"] = np.array([N, L, DPLUS])
    # particle::lattice_generator of the reference: ids and bare-lattice positions as SHA-256 digests (bit-exact quantities),
    # displaced positions / velocities / masses on every 101st particle
    import hashlib
    for lt, nm in enumerate(("sc", "bcc", "fcc", "rsc")):
        for li in (0, 1):
            Z = run_reference_particles(phi, psi, lt, li, zero_field=True)
            out[f"part_{nm}{li}_ids_sha256"] = np.frombuffer(hashlib.sha256(Z["ids"].tobytes()).digest(), np.uint8)
            out[f"part_{nm}{li}_lattice_sha256"] = np.frombuffer(hashlib.sha256(Z["pos"].tobytes()).digest(), np.uint8)
            P = run_reference_particles(phi, psi, lt, li)
            assert np.array_equal(P["ids"], Z["ids"])
            for k in ("pos", "vel", "mass"):
                out[f"part_{nm}{li}_{k}_sub"] = P[k][::101].copy()
    path = os.path.join(HERE, "ref_intermediates_32.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KB,", len(out), "arrays")


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "f4":
        main_f4()
    else:
        main()
