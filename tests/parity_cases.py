"""Backend-agnostic parity checks: the CUDA path (through the C ABI host mirror `M`) against the
CPU oracle.  Used twice: tests/test_emu_parity.py runs them on the thread-emulation build at tiny
sizes (no GPU needed), tests/test_gpu_parity.py on the real sm_100a build on a B200.

Tolerances: FP64 fields  max|err| <= 1e-12 * max|ref|  (far inside north_star's 1e-9*max|Psi|);
goldens at the reference's own np.allclose(rtol=1e-9, atol=1e-14); integer IDs exact."""
import numpy as np

import golden_util as G
import lpt_oracle as O

TOL = 1e-12


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def noise_and_phi(n, L=100.0, seed=7, nyquist_zero=True):
    g = O.GridSpec(n, L)
    W = O.make_white_noise_k(n, seed, nyquist_zero)
    Wc = O.condition_noise(g, W)
    phi = O.phi1_from_noise(g, Wc, O.power_law_amplitude())
    return g, W, Wc, phi


# ------------------------------------------------------------------ FFT
def check_fft(M, n, n1=None, n2=None):
    n1, n2 = n1 or n, n2 or n
    rng = np.random.default_rng(n + 3 * n1 + 7 * n2)
    x = rng.standard_normal((n, n1, n2))
    G_ = M.Grid_FFT((n, n1, n2), (100.0, 100.0, 100.0))
    G_.set_real(x)
    G_.FourierTransformForward()
    ref = np.fft.rfftn(x) / np.sqrt(float(n * n1 * n2))
    assert relerr(G_.get_kspace(), ref) < TOL
    G_.FourierTransformBackward()
    assert relerr(G_.get_real(), x) < TOL
    # do_transform=False only flips the flag (src/grid_fft.cc:218,230-235)
    G_.FourierTransformForward(False)
    assert G_.space_ == M.kspace_id
    G_.FourierTransformBackward(False)
    assert relerr(G_.get_real(), x) < TOL


def check_c2r_nonhermitian(M, n):
    """FFTW c2r ignores Im of the self-conjugate kz=0 / kz=n/2 entries: same here."""
    rng = np.random.default_rng(5)
    X = rng.standard_normal((n, n, n // 2 + 1)) + 1j * rng.standard_normal((n, n, n // 2 + 1))
    G_ = M.Grid_FFT((n,) * 3, (1.0,) * 3)
    G_.set_kspace(X)
    G_.FourierTransformBackward()
    ref = np.fft.irfftn(X, s=(n, n, n), axes=(0, 1, 2)) * np.sqrt(float(n) ** 3)
    assert relerr(G_.get_real(), ref) < TOL


# ------------------------------------------------------------------ k-space operators
def check_kops(M, n):
    L = 100.0
    g, W, Wc, phi = noise_and_phi(n, L)
    Gw = M.Grid_FFT((n,) * 3, (L,) * 3)
    Gw.set_kspace(W)
    lib = M.grid._lib()
    import ctypes as C
    volfac = (L / n / 2.0 / np.pi) ** 1.5
    assert lib.mfb_noise_condition(Gw._h, volfac, 0, 0, 0) == 0
    assert relerr(Gw.get_kspace(), Wc) < TOL
    # fixing + inversion + corner removal
    Gw.set_kspace(W)
    assert lib.mfb_noise_condition(Gw._h, volfac, 1, 1, 1) == 0
    assert relerr(Gw.get_kspace(), O.condition_noise(g, W, True, True, True)) < TOL
    # phi1 through the radial table
    Gw.set_kspace(Wc)
    Gp = M.Grid_FFT((n,) * 3, (L,) * 3)
    tab = O.radial_table(g, O.power_law_amplitude())
    assert lib.mfb_apply_radial_table(Gp._h, Gw._h, tab.ctypes.data_as(C.c_void_p), tab.size, 1) == 0
    assert relerr(Gp.get_kspace(), phi) < TOL
    # Poisson / Laplacians / DC
    Gp.apply_InverseLaplacian()
    assert relerr(Gp.get_kspace(), O.inverse_laplacian(g, phi)) < TOL
    Gp.set_kspace(phi)
    Gp.apply_Laplacian()
    ref = -g.k2() * phi
    ref[0, 0, 0] = 0
    assert relerr(Gp.get_kspace(), ref) < TOL
    # scale, copy, axpy
    Gp.set_kspace(phi)
    Gp *= 2.5
    Gq = M.Grid_FFT((n,) * 3, (L,) * 3)
    Gq.copy_from(Gp)
    assert Gq.space_ == M.kspace_id
    Gq.axpy(2.0, Gp, -1.0)
    assert relerr(Gq.get_kspace(), 2.5 * phi) < TOL
    # shift_field incl. Nyquist planes (include/grid_fft.hh:832-842)
    _, _, _, phin = noise_and_phi(n, L, 11, nyquist_zero=False)
    x = g.bwd(phin)
    Gq.set_real(x)
    Gq.shift_field((0.5, 0.25, 0.5))
    assert relerr(Gq.get_real(), O.shift_field_r(g, x, (0.5, 0.25, 0.5))) < TOL
    # reductions
    assert abs(Gq.mean() - Gq.get_real().mean()) < 1e-13 * max(1.0, np.abs(x).max())
    assert abs(Gq.std() - Gq.get_real().std()) < 1e-12 * x.std()
    xr = Gq.get_real()
    d = max(np.abs(np.roll(xr, -1, a) - xr)[: n - 1].max() for a in range(3))
    assert abs(Gq.max_neighbour_diff() - d) <= 1e-15 * d
    Gq.zero_DC_mode()
    assert abs(Gq.get_real().mean()) < 1e-14 * np.abs(x).max()
    # resample N -> 3N/2 -> N (FourierInterpolateCopyTo, Nyquist dropped)
    n15 = 3 * n // 2
    Gb = M.Grid_FFT((n15,) * 3, (L,) * 3)
    Gp.set_kspace(phin)
    Gp.FourierInterpolateCopyTo(Gb)
    assert relerr(Gb.get_kspace(), O.fourier_interpolate_copy(g, O.GridSpec(n15, L), phin)) < TOL
    Gb.FourierInterpolateCopyTo(Gq)
    ref = phin.copy()
    ref[n // 2], ref[:, n // 2], ref[:, :, n // 2] = 0, 0, 0
    assert relerr(Gq.get_kspace(), ref) < TOL


# ------------------------------------------------------------------ convolver
def check_convolver(M, n, nyquist_zero=True):
    L = 100.0
    g, W, Wc, phi = noise_and_phi(n, L, 9, nyquist_zero)
    _, _, _, psi = noise_and_phi(n, L, 10, nyquist_zero)
    cv = O.OrszagConvolver(g)
    P = M.Grid_FFT((n,) * 3, (L,) * 3)
    Q = M.Grid_FFT((n,) * 3, (L,) * 3)
    out = M.Grid_FFT((n,) * 3, (L,) * 3, True, M.kspace_id)
    P.set_kspace(phi)
    Q.set_kspace(psi)
    C = M.OrszagConvolver((n,) * 3, (L,) * 3)
    op = M.op
    ref = cv.convolve_Hessians(phi, (0, 1), psi, (1, 2))
    C.convolve_Hessians(P, (0, 1), Q, (1, 2), op.assign_to(out))
    assert relerr(out.get_kspace(), ref) < TOL
    got = out.get_kspace()
    h = n // 2
    assert not got[h].any() and not got[:, h].any() and not got[:, :, h].any()  # Nyquist planes are exactly 0
    ref = ref + cv.convolve_SumOfHessians(phi, (0, 0), psi, (1, 1), (2, 2))
    C.convolve_SumOfHessians(P, (0, 0), Q, (1, 1), (2, 2), op.add_to(out))
    assert relerr(out.get_kspace(), ref) < TOL
    ref = ref - cv.convolve_DifferenceOfHessians(phi, (1, 2), psi, (1, 1), (2, 2))
    C.convolve_DifferenceOfHessians(P, (1, 2), Q, (1, 1), (2, 2), op.subtract_from(out))
    assert relerr(out.get_kspace(), ref) < TOL
    ref = ref + (-5.0 / 7.0) * cv.convolve_Hessians3(phi, (0, 0), psi, (0, 2), phi, (2, 2))
    C.convolve_Hessians(P, (0, 0), Q, (0, 2), P, (2, 2), op.multiply_add_to(out, -5.0 / 7.0))
    assert relerr(out.get_kspace(), ref) < TOL
    ref = cv.convolve_Gradients(phi, 0, psi, 2)
    C.convolve_Gradients(P, (0,), Q, (2,), op.assign_to(out))
    assert relerr(out.get_kspace(), ref) < TOL
    ref = cv.convolve_Gradient_and_Hessian(phi, 1, psi, (0, 2))
    C.convolve_Gradient_and_Hessian(P, (1,), Q, (0, 2), op.assign_to(out))
    assert relerr(out.get_kspace(), ref) < TOL
    ref = cv.multiply_field(phi, psi)
    C.multiply_field(P, Q, op.assign_to(out))
    assert relerr(out.get_kspace(), ref) < TOL
    # convolve2 / convolve3 with caller-formed spectra (mfb_conv_generic, mfb_conv_generic3; convolution.hh:512-549)
    C.convolve2(P, Q, op.assign_to(out))
    assert relerr(out.get_kspace(), ref) < TOL
    R = M.Grid_FFT((n,) * 3, (L,) * 3)
    _, _, _, chi = noise_and_phi(n, L, 11, nyquist_zero)
    R.set_kspace(chi)
    ref3 = cv.convolve3(phi, psi, chi)
    C.convolve3(P, Q, R, op.assign_to(out))
    assert relerr(out.get_kspace(), ref3) < TOL
    C.convolve3(P, Q, R, op.multiply_add_to(out, -0.25))
    assert relerr(out.get_kspace(), 0.75 * ref3) < TOL
    P.FourierTransformBackward()
    try:
        C.convolve2(P, Q, op.assign_to(out))     # spectra are k-space grids by contract
        raise AssertionError("convolve2 accepted a real-space operand")
    except M.MfbError:
        pass
    # operands handed over in real space are transformed first (convolution.hh:142-143)
    C.multiply_field(P, Q, op.assign_to(out))
    assert relerr(out.get_kspace(), ref) < 10 * TOL


def check_naive_convolver(M, n, nyquist_zero=False):
    """NaiveConvolver (include/convolution.hh:296-426): every named product against the oracle, Nyquist planes
    populated (they take part in an un-padded product), three-field product in one go"""
    L = 100.0
    g, W, Wc, phi = noise_and_phi(n, L, 9, nyquist_zero)
    _, _, _, psi = noise_and_phi(n, L, 10, nyquist_zero)
    cv = O.NaiveConvolver(g)
    P = M.Grid_FFT((n,) * 3, (L,) * 3)
    Q = M.Grid_FFT((n,) * 3, (L,) * 3)
    out = M.Grid_FFT((n,) * 3, (L,) * 3, True, M.kspace_id)
    P.set_kspace(phi)
    Q.set_kspace(psi)
    C = M.NaiveConvolver((n,) * 3, (L,) * 3)
    assert C.workspace_bytes() == 2 * 8 * n * n * (n + 2)
    op = M.op
    ref = cv.convolve_Hessians(phi, (0, 1), psi, (1, 2))
    C.convolve_Hessians(P, (0, 1), Q, (1, 2), op.assign_to(out))
    assert relerr(out.get_kspace(), ref) < TOL
    ref = ref + cv.convolve_SumOfHessians(phi, (0, 0), psi, (1, 1), (2, 2))
    C.convolve_SumOfHessians(P, (0, 0), Q, (1, 1), (2, 2), op.add_to(out))
    assert relerr(out.get_kspace(), ref) < TOL
    ref = ref - cv.convolve_DifferenceOfHessians(phi, (1, 2), psi, (1, 1), (2, 2))
    C.convolve_DifferenceOfHessians(P, (1, 2), Q, (1, 1), (2, 2), op.subtract_from(out))
    assert relerr(out.get_kspace(), ref) < TOL
    ref = ref + (-5.0 / 7.0) * cv.convolve_Hessians3(phi, (0, 0), psi, (0, 2), phi, (2, 2))
    C.convolve_Hessians(P, (0, 0), Q, (0, 2), P, (2, 2), op.multiply_add_to(out, -5.0 / 7.0))
    assert relerr(out.get_kspace(), ref) < TOL
    ref = cv.convolve_Gradients(phi, 0, psi, 2)
    C.convolve_Gradients(P, (0,), Q, (2,), op.assign_to(out))
    assert relerr(out.get_kspace(), ref) < TOL
    ref = cv.convolve_Gradient_and_Hessian(phi, 1, psi, (0, 2))      # get_k: Nyquist wave number is +N/2, not 0
    C.convolve_Gradient_and_Hessian(P, (1,), Q, (0, 2), op.assign_to(out))
    assert relerr(out.get_kspace(), ref) < TOL
    ref = cv.multiply_field(phi, psi)
    C.multiply_field(P, Q, op.assign_to(out))
    assert relerr(out.get_kspace(), ref) < TOL
    if not nyquist_zero:
        assert np.abs(out.get_kspace()[n // 2]).max() > 0       # no Nyquist zeroing in the naive convolver
    # the aliased product differs from the dealiased one: the two convolvers must NOT agree
    Cz = M.OrszagConvolver((n,) * 3, (L,) * 3)
    o2 = M.Grid_FFT((n,) * 3, (L,) * 3, True, M.kspace_id)
    Cz.multiply_field(P, Q, op.assign_to(o2))
    dealiased = o2.get_kspace()
    assert relerr(dealiased, ref) > 1e-10                        # far above round-off (the spectra are red: aliasing is small)
    assert not dealiased[n // 2].any() and np.abs(ref[n // 2]).max() > 0   # Orszag zeroes the Nyquist planes, Naive fills them


def check_naive_pipeline(M, n=32):
    """the whole 3LPT source stage with CONVOLVER_TYPE=NAIVE against the oracle run with its NaiveConvolver"""
    L, Dp = 100.0, 0.0257
    g, W, Wc, phi = noise_and_phi(n, L, 17)
    amp = O.power_law_amplitude()
    ref = O.lpt_potentials(g, phi, 3, Dp, convolver=O.NaiveConvolver(g))
    pl = M.LPTPipeline(M.lpt_config(GridRes=n, BoxLength=L, LPTorder=3, Dplus0=Dp, ConvolverType="naive"))
    pl.set_noise_k(W)
    pl.condition_noise()
    pl.compute_phi1(O.radial_table(g, amp))
    pl.compute_sources()
    assert relerr(pl.phi2.get_kspace(), ref["phi2"]) < TOL
    assert relerr(pl.phi3.get_kspace(), ref["phi3"]) < 10 * TOL
    for d in range(3):
        assert relerr(pl.A3[d].get_kspace(), ref["A3"][d]) < 10 * TOL


# ------------------------------------------------------------------ fixtures produced by the reference's own code
def load_reference_fixture():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_intermediates_32.npz"))


def check_reference_fixture(M):
    """the CUDA path against what the REFERENCE'S OWN Grid_FFT / OrszagConvolver / op:: sources produce
    (tests/golden/ref_intermediates_32.npz, made by tests/golden/make_ref_fixtures.py from oracle/_ref/ref_driver):
    transforms, shift_field, Poisson solve, FourierInterpolateCopyTo, every named product with every output operator, and
    the phi2 / phi3 / A3 source stage of src/ic_generator.cc:475-574"""
    R = load_reference_fixture()
    n, L, D = int(R["meta"][0]), float(R["meta"][1]), float(R["meta"][2])
    phi, psi = R["phi_in"], R["psi_in"]
    N3, L3 = (n,) * 3, (L,) * 3
    tol = 1e-11      # relative to the field's maximum; the bar is 1e-9
    P, Q, T = M.Grid_FFT(N3, L3), M.Grid_FFT(N3, L3), M.Grid_FFT(N3, L3)
    out = M.Grid_FFT(N3, L3, True, M.kspace_id)
    P.set_kspace(phi)
    Q.set_kspace(psi)
    T.set_kspace(phi)
    T.FourierTransformBackward()
    assert relerr(T.get_real(), R["phi_real"]) < tol
    assert abs(T.mean() - R["phi_real_mean_std"][0]) < 1e-13 * np.abs(R["phi_real"]).max()
    assert abs(T.std() - R["phi_real_mean_std"][1]) < 1e-12 * R["phi_real_mean_std"][1]
    T.shift_field((0.5, 0.25, 0.5))
    assert relerr(T.get_real(), R["phi_real_shifted"]) < tol
    T.FourierTransformForward()
    T.apply_InverseLaplacian()
    assert relerr(T.get_kspace(), R["invlap_shifted_phi"]) < tol
    big = M.Grid_FFT((3 * n // 2,) * 3, L3)
    Q.FourierInterpolateCopyTo(big)
    assert relerr(big.get_kspace(), R["psi_interp_3half"]) < tol
    C, op = M.OrszagConvolver(N3, L3), M.op
    C.convolve_Hessians(P, (0, 1), Q, (1, 2), op.assign_to(out))
    assert relerr(out.get_kspace(), R["hess_01_12"]) < tol
    C.convolve_SumOfHessians(P, (0, 0), Q, (1, 1), (2, 2), op.assign_to(out))
    assert relerr(out.get_kspace(), R["sumhess_00_11_22"]) < tol
    C.convolve_DifferenceOfHessians(P, (1, 2), Q, (1, 1), (2, 2), op.assign_to(out))
    assert relerr(out.get_kspace(), R["diffhess_12_11_22"]) < tol
    C.convolve_Hessians(P, (0, 0), Q, (0, 2), P, (2, 2), op.assign_to(out))
    assert relerr(out.get_kspace(), R["hess3_00_02_22"]) < tol
    C.convolve_Gradients(P, (0,), Q, (2,), op.assign_to(out))
    assert relerr(out.get_kspace(), R["grad_0_2"]) < tol
    C.convolve_Gradient_and_Hessian(P, (1,), Q, (0, 2), op.assign_to(out))
    assert relerr(out.get_kspace(), R["gradhess_1_02"]) < tol
    C.multiply_field(P, Q, op.assign_to(out))
    assert relerr(out.get_kspace(), R["multiply"]) < tol
    C.multiply_field(P, Q, op.multiply_add_to(out, 0.5))
    C.multiply_field(P, P, op.subtract_from(out))
    C.multiply_field(Q, Q, op.add_to(out))
    assert relerr(out.get_kspace(), R["multiply_ops"]) < tol
    # NaiveConvolver on Nyquist-carrying fields, and the c2c transform of the complex grids
    A, B = M.Grid_FFT(N3, L3), M.Grid_FFT(N3, L3)
    A.set_kspace(R["naive_a_in"])
    B.set_kspace(R["naive_b_in"])
    NC = M.NaiveConvolver(N3, L3)
    NC.convolve_Hessians(A, (0, 0), B, (0, 2), A, (2, 2), op.assign_to(out))
    assert relerr(out.get_kspace(), R["naive_hess3_00_02_22"]) < tol
    NC.convolve_Gradient_and_Hessian(A, (1,), B, (0, 2), op.assign_to(out))
    assert relerr(out.get_kspace(), R["naive_gradhess_1_02"]) < tol
    NC.multiply_field(A, B, op.assign_to(out))
    assert relerr(out.get_kspace(), R["naive_multiply"]) < tol
    # Nyquist-CARRYING fields through the dealiased products: the reference's USE_MPI code path (pad / unpad drop the
    # Nyquist planes) is what the CUDA path implements
    A.shift_field((0.5, 0.25, 0.5))
    assert relerr(A.get_real(), R["mpinyq_phi_real_shifted"]) < tol
    A.FourierTransformForward()
    A.apply_InverseLaplacian()
    assert relerr(A.get_kspace(), R["mpinyq_invlap_shifted_phi"]) < tol
    A.set_kspace(R["naive_a_in"])
    C.convolve_Hessians(A, (0, 0), B, (0, 2), A, (2, 2), op.assign_to(out))
    assert relerr(out.get_kspace(), R["mpinyq_hess3_00_02_22"]) < tol
    C.convolve_Gradient_and_Hessian(A, (1,), B, (0, 2), op.assign_to(out))
    assert relerr(out.get_kspace(), R["mpinyq_gradhess_1_02"]) < tol
    C.multiply_field(A, B, op.assign_to(out))
    C.multiply_field(A, B, op.multiply_add_to(out, 0.5))
    C.multiply_field(A, A, op.subtract_from(out))
    C.multiply_field(B, B, op.add_to(out))
    assert relerr(out.get_kspace(), R["mpinyq_multiply_ops"]) < tol
    pn = M.LPTPipeline(M.lpt_config(GridRes=n, BoxLength=L, LPTorder=3, Dplus0=D))
    pn.phi.set_kspace(R["naive_a_in"])
    pn.compute_sources()
    assert relerr(pn.phi3.get_kspace(), R["mpinyq_phi3"]) < 10 * tol
    assert relerr(pn.A3[0].get_kspace(), R["mpinyq_A3x"]) < 10 * tol
    Z = M.Grid_FFT_complex(N3, L3)
    Z.set(R["cplx_in"])
    Z.FourierTransformForward()
    assert relerr(Z.get(), R["cplx_k"]) < tol
    # the whole source stage, grouped (default) and product by product
    lib = M.grid._lib()
    try:
        for grouping in (1, 0):
            assert lib.mfb_set_option(b"conv_grouping", grouping) == 0
            pl = M.LPTPipeline(M.lpt_config(GridRes=n, BoxLength=L, LPTorder=3, Dplus0=D))
            pl.phi.set_kspace(phi)
            pl.compute_sources()
            assert relerr(pl.phi.get_kspace(), R["phi_scaled"]) < tol
            assert relerr(pl.phi2.get_kspace(), R["phi2"]) < tol
            assert relerr(pl.phi3.get_kspace(), R["phi3"]) < 10 * tol
            for d, nm in enumerate(("A3x", "A3y", "A3z")):
                assert relerr(pl.A3[d].get_kspace(), R[nm]) < 10 * tol
    finally:
        lib.mfb_set_option(b"conv_grouping", 1)


def check_reference_particles(M):
    """the particle read-out against the reference's own particle::lattice_generator (fixture digests and subsamples):
    ids and bare-lattice positions bit-exact, displaced positions / velocities / masses to 1e-12, every Bravais lattice
    and both lattice indices"""
    import hashlib

    R = load_reference_fixture()
    n, L = int(R["meta"][0]), float(R["meta"][1])
    N3, L3 = (n,) * 3, (L,) * 3
    F = M.Grid_FFT(N3, L3)
    for lt, nm in enumerate(("sc", "bcc", "fcc", "rsc")):
        for li in (0, 1):
            parts = M.Particles(nm, li, True, True, True, li, F)
            for zero in (True, False):
                for d in range(3):
                    F.set_kspace(R["phi_in"])
                    if zero:
                        F.zero()
                    F.FourierTransformBackward()
                    parts.set_positions(d, L, F)
                if zero:
                    got = parts.download(vel=False, mass=False)
                    assert hashlib.sha256(got["ids"].tobytes()).digest() == R[f"part_{nm}{li}_ids_sha256"].tobytes(), (nm, li)
                    assert hashlib.sha256(got["pos"].tobytes()).digest() == R[f"part_{nm}{li}_lattice_sha256"].tobytes(), (nm, li)
            for d in range(3):
                F.set_kspace(R["psi_in"])
                F.FourierTransformBackward()
                parts.set_velocities(d, F)
            F.set_kspace(R["phi_in"])
            F.FourierTransformBackward()
            parts.set_masses(2.5, F)
            got = parts.download()
            assert np.abs(got["pos"][::101] - R[f"part_{nm}{li}_pos_sub"]).max() < 1e-12 * L, (nm, li)
            assert relerr(got["vel"][::101], R[f"part_{nm}{li}_vel_sub"]) < 1e-12, (nm, li)
            assert relerr(got["mass"][::101], R[f"part_{nm}{li}_mass_sub"]) < 1e-12, (nm, li)


def check_reference_f4(M):
    """SURVEY 8 f-4 against the reference's OWN code (tests/golden/ref_f4_32.npz, made by make_ref_fixtures.py f4): the
    masked sc lattice of particle::lattice_generator - ids bit-exact, positions / velocities / masses - and the CIC read-out
    of a glass load against grid_interpolate::get_cic_at"""
    import hashlib
    import os

    R0 = load_reference_fixture()
    R = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_f4_32.npz"))
    n, L = int(R["meta"][0]), float(R["meta"][1])
    F = M.Grid_FFT((n,) * 3, (L,) * 3)
    for do_b, mt, li in ((True, 3, 0), (True, 3, 1), (True, 1, 0), (True, 2, 1), (False, 3, 0)):
        key = f"masked_{int(do_b)}{mt}{li}"
        parts = M.Particles("masked", li, True, True, True, li, F, mask=M.particle_type_mask(do_b, mt))
        assert parts.get_local_num_particles() == int(R[key + "_np"][0])
        for d in range(3):
            F.set_kspace(R0["phi_in"])
            F.FourierTransformBackward()
            parts.set_positions(d, L, F)
            F.set_kspace(R0["psi_in"])
            F.FourierTransformBackward()
            parts.set_velocities(d, F)
        F.set_kspace(R0["phi_in"])
        F.FourierTransformBackward()
        parts.set_masses(2.5, F)
        got = parts.download()
        assert hashlib.sha256(got["ids"].tobytes()).digest() == R[key + "_ids_sha256"].tobytes(), key
        assert np.abs(got["pos"][::37] - R[key + "_pos_sub"]).max() < 1e-12 * L, key
        assert relerr(got["vel"][::37], R[key + "_vel_sub"]) < 1e-12, key
        assert relerr(got["mass"][::37], R[key + "_mass_sub"]) < 1e-12, key
    pts = R["cic_points"]
    parts = M.Particles("glass", 0, True, True, False, 0, F, glass=(pts, 0))
    F.set_kspace(R0["phi_in"])
    F.FourierTransformBackward()
    parts.set_velocities(1, F)
    got = parts.download(pos=False)
    assert np.abs(got["vel"][:, 1] - R["cic_out"][:, 5]).max() < 1e-12 * np.abs(R["cic_out"][:, 5]).max()


# ------------------------------------------------------------------ two-fluid terms, tides, density, masses
def check_two_fluid_terms(M, n=32, lattice="bcc"):
    """the operators a CDM+baryon run adds (config C4) with NON-degenerate two-fluid amplitudes (the goldens use
    Eisenstein&Hu, where delta_bc == 0): external tides in the 2LPT source (:488-497), the relative-velocity term of the
    velocity assembly (:918-921), the mass field (1 + C rho) Omega munit (:663-675) read out as particle masses on the
    secondary lattice (:292-350), and the SPT density delta = k^2 phi (:962-973)"""
    import ctypes as C

    L, Dp, vfac = 100.0, 0.0257, 396.69
    lam = (0.12, -0.07, -0.05)
    g, W, Wc, phi = noise_and_phi(n, L, 41)
    amp = O.power_law_amplitude()
    amp_theta = lambda k: 0.37 * amp(k) / (1.0 + k)           # stand-ins for get_amplitude_theta_bc / _delta_bc
    amp_delta = lambda k: 2e-3 * amp(k) * k * k
    Csp, Omega, munit, vunit = -0.1567, 0.3, 2.5, L / np.sqrt(1 / 51.0)
    pots = O.lpt_potentials(g, phi, 2, Dp, tides=lam)
    lt = O.LATTICE_TYPES[lattice]
    cfg = M.lpt_config(GridRes=n, BoxLength=L, LPTorder=2, ParticleLoad=lattice, Dplus0=Dp, vfac=vfac, lunit=L, vunit=vunit,
                       tides=lam)
    pl = M.LPTPipeline(cfg)
    pl.set_noise_k(W)
    pl.condition_noise()
    pl.compute_phi1(O.radial_table(g, amp))
    pl.compute_sources()
    assert relerr(pl.phi2.get_kspace(), pots["phi2"]) < TOL
    # velocities with the baryon-CDM relative velocity term
    vtab = O.radial_table(g, amp_theta)
    for d in range(3):
        ref = g.bwd(O.velocity_k(g, pots, 2, d, vfac, vunit / L, vbc=(Csp, amp_theta, Wc), tides=lam, Dplus0=Dp))
        got = pl.velocity_field(d, vbc=(vtab, vfac * Csp)).get_real()
        assert np.abs(got - ref).max() < 1e-11 * np.abs(ref).max(), d
        # the anisotropic velocity growth factor, stated independently of the oracle: with external tides every velocity
        # mode is the tide-free assembly times (1 - D+ lambda_d)^2 (src/ic_generator.cc:238-242, 931-934)
        plain = g.bwd(O.velocity_k(g, pots, 2, d, vfac, vunit / L, vbc=(Csp, amp_theta, Wc)))
        alpha2 = (1.0 - Dp * lam[d]) * (1.0 - Dp * lam[d])
        assert abs(alpha2 - 1.0) > 1e-3
        assert np.abs(got - alpha2 * plain).max() < 1e-11 * np.abs(plain).max(), d
    # SPT density
    ref = g.bwd(O.density_k(g, pots))
    assert np.abs(pl.density_field().get_real() - ref).max() < 1e-11 * np.abs(ref).max()
    # mass field -> particle masses on the secondary lattice
    rho_k = O.zero_dc(Wc * np.where(np.isfinite(amp_delta(np.sqrt(g.k2()))), amp_delta(np.sqrt(g.k2())), 0.0))
    rho_r = g.bwd(rho_k)
    mass_field = (1.0 + Csp * rho_r) * Omega * munit
    rho = M.Grid_FFT((n,) * 3, (L,) * 3)
    dtab = O.radial_table(g, amp_delta)
    lib = M.grid._lib()
    assert lib.mfb_apply_radial_table(rho._h, pl.wnoise._h, dtab.ctypes.data_as(C.c_void_p), dtab.size, 0) == 0
    rho.zero_DC_mode()
    rho.FourierTransformBackward()
    rho.affine(Omega * munit, Csp * Omega * munit)
    assert np.abs(rho.get_real() - mass_field).max() < 1e-12 * np.abs(mass_field).max()
    parts = M.Particles(lattice, 1, True, True, True, 1, rho)
    parts.set_masses(1.0, rho)
    got = parts.download(pos=False, vel=False)
    ref_m = O.set_masses(g, mass_field, lt, 1, 1.0, True)
    assert np.abs(got["mass"] - ref_m).max() < 1e-12 * np.abs(ref_m).max()
    assert np.array_equal(got["ids"], O.particle_ids(n, lt, 1, True))


# ------------------------------------------------------------------ noise ingest and device noise fill
def check_noise_ingest(M, n=32):
    """how white noise reaches the device path, the three ways the reference's RNG plugins deliver it:
    (1) NGenIC: a k-space fill (src/plugins/random_ngenic.cc:84-154) - here generated on the device with the plugin's layout
        rules and a counter-based generator, against the numpy restatement;
    (2) MUSIC1: a real-space host fill (src/plugins/random_music.cc:92-116) streamed in on the ingest stream, then r2c;
    (3) Panphasia-HO: a 2x-oversampled grid with populated Nyquist planes resampled down with FourierInterpolateCopyTo
        (src/plugins/random_panphasia.cc:136-186, src/grid_fft.cc:273-492).
    Each ends in a pipeline run that agrees with the oracle fed the same noise."""
    L, Dp, vfac = 100.0, 0.0257, 396.69
    g = O.GridSpec(n, L)
    amp = O.power_law_amplitude()
    tab = O.radial_table(g, amp)
    seed = 0x1234567887654321

    def run(pl):
        pl.condition_noise()
        return pl.run_device(tab, particles=False)

    def ref_fields(W):
        pots = O.lpt_potentials(g, O.phi1_from_noise(g, O.condition_noise(g, W), amp), 2, Dp)
        return O.fields(g, pots, 2, vfac)
    cfg = M.lpt_config(GridRes=n, BoxLength=L, LPTorder=2, Dplus0=Dp, vfac=vfac)
    pl = M.LPTPipeline(cfg)
    # (1) device fill, NGenIC layout
    Wref = O.ngenic_layout_noise(n, seed)
    pl.wnoise.fill_noise_ngenic_layout(seed)
    assert pl.wnoise.space_ == M.kspace_id
    got = pl.wnoise.get_kspace()
    assert np.abs(got - Wref).max() < 1e-13, np.abs(got - Wref).max()
    h = n // 2
    assert not got[h].any() and not got[:, h].any() and not got[:, :, h].any() and got[0, 0, 0] == 0
    assert np.abs(got[1:, 1:, 0] - np.conj(got[1:, 1:, 0][::-1, ::-1])).max() == 0.0     # Hermitian k_z = 0 plane, bit for bit
    inner = got[1:h, 1:h, 1:h]
    assert abs(np.mean(np.abs(inner) ** 2) - 1.0) < 5.0 / np.sqrt(inner.size)            # unit variance per complex mode
    pl.wnoise.fill_noise_ngenic_layout(seed + 1)
    assert np.abs(pl.wnoise.get_kspace() - Wref).max() > 1.0                             # another seed, another field
    # ... and the reference's own NGenIC noise (RANLUX column streams, src/plugins/random_ngenic.cc:42-121), generated on the
    # device: equal to the plugin's restatement to the last bits of log / cos / sin; Hermitian plane bit for bit
    Wx = O.ngenic_noise(n, 12345)
    pl.wnoise.fill_noise_ngenic(12345)
    assert pl.wnoise.space_ == M.kspace_id
    gx = pl.wnoise.get_kspace()
    assert np.abs(gx - Wx).max() < 1e-14, np.abs(gx - Wx).max()
    assert not gx[h].any() and not gx[:, h].any() and not gx[:, :, h].any() and gx[0, 0, 0] == 0
    assert np.abs(gx[1:, 1:, 0] - np.conj(gx[1:, 1:, 0][::-1, ::-1])).max() == 0.0
    assert np.abs(gx[0, 1:, 0] - np.conj(gx[0, 1:, 0][::-1])).max() == 0.0
    pl.wnoise.fill_noise_ngenic_layout(seed)
    out, ref = run(pl), ref_fields(Wref)
    for k in ("dx", "dy", "dz", "vx", "vy", "vz"):
        assert np.abs(out[k] - ref[k]).max() < 1e-11 * np.abs(ref[k]).max(), k
    # (2) real-space host fill, streamed in
    rng = np.random.default_rng(99)
    x = rng.standard_normal((n, n, n))
    w = pl.wnoise
    if w.local_0_size_ == n:
        raw = np.zeros(w.ntot_)
        raw.reshape(n, n, n + 2)[:, :, :n] = x
        w.FourierTransformBackward(False)
        w.stream_in(raw)
        w.stream_in_wait()
        w.FourierTransformForward()
        assert relerr(w.get_kspace(), g.fwd(x)) < TOL
        out, ref = run(pl), ref_fields(g.fwd(x))
        for k in ("dx", "vz"):
            assert np.abs(out[k] - ref[k]).max() < 1e-11 * np.abs(ref[k]).max(), k
    # (3) 2x-oversampled k-space noise with populated Nyquist planes, resampled down
    g2 = O.GridSpec(2 * n, L)
    W2 = O.make_white_noise_k(2 * n, 77, nyquist_zero=False)
    big = M.Grid_FFT((2 * n,) * 3, (L,) * 3)
    big.set_kspace(W2)
    big.FourierInterpolateCopyTo(pl.wnoise)
    Wd = O.fourier_interpolate_copy(g2, g, W2)
    assert relerr(pl.wnoise.get_kspace(), Wd) < TOL
    out, ref = run(pl), ref_fields(Wd)
    for k in ("dx", "vz"):
        assert np.abs(out[k] - ref[k]).max() < 1e-11 * np.abs(ref[k]).max(), k


# ------------------------------------------------------------------ PLT gradient, masked lattice, glass load
def check_plt_masked_glass(M, n=32, order=3):
    """(1) particle::lattice_gradient in place of op::fourier_gradient (include/particle_plt.hh:551-585, src/ic_generator.cc:
    377-381, 855-858, 911-929): component fields built with the table-driven gradient and the growth-rate correction;
    (2) the masked sc lattice (include/particle_generator.hh:171-260, 351-398, 479-497, 557-571) for every ParticleMaskType and
    both species; (3) a glass load (:262-290, 498-515, 573-590): CIC read-out at arbitrary positions, component fields with
    the CIC compensation kernel (include/grid_interpolate.hh:195-208)."""
    L, Dp, vfac = 100.0, 0.0257, 396.69
    lunit, vunit = L, L / np.sqrt(1 / 51.0)
    g, W, Wc, phi = noise_and_phi(n, L, 31)
    amp = O.power_law_amplitude()
    pots = O.lpt_potentials(g, phi, order, Dp)
    tab = O.radial_table(g, amp)
    rng = np.random.default_rng(12)
    # ---- (1) PLT: smooth, non-trivial helper-mesh tables (the real ones come from the host's eigen-decomposition)
    ng = 16
    kk = np.fft.fftfreq(ng)[:, None, None] ** 2 + np.fft.fftfreq(ng)[None, :, None] ** 2 + np.fft.rfftfreq(ng)[None, None, :] ** 2
    tabs = [1.0 - 0.8 * kk + 0.05 * rng.standard_normal(kk.shape), 0.3 * kk + 0.02 * rng.standard_normal(kk.shape),
            -0.2 * kk + 0.02 * rng.standard_normal(kk.shape), 1.0 + 0.4 * kk + 0.03 * rng.standard_normal(kk.shape)]
    olg = O.PLTGradient(ng, n, L, *tabs)
    cfg = M.lpt_config(GridRes=n, BoxLength=L, LPTorder=order, ParticleLoad="sc", Dplus0=Dp, vfac=vfac, lunit=lunit, vunit=vunit)
    pl = M.LPTPipeline(cfg)
    pl.set_noise_k(W)
    pl.condition_noise()
    pl.compute_phi1(tab)
    pl.compute_sources()
    plain = [pl.displacement_field(d).get_real() for d in range(3)]
    pl.lg = M.lattice_gradient(ng, n, L, *tabs)
    for d in range(3):
        ref = g.bwd(O.displacement_k(g, pots, order, d, lunit / L, lg=olg))
        got = pl.displacement_field(d).get_real()
        assert np.abs(got - ref).max() < 1e-11 * np.abs(ref).max(), d
        assert np.abs(got - plain[d]).max() > 1e-3 * np.abs(ref).max()      # the tables matter
        ref = g.bwd(O.velocity_k(g, pots, order, d, vfac, vunit / L, lg=olg))
        got = pl.velocity_field(d).get_real()
        assert np.abs(got - ref).max() < 1e-11 * np.abs(ref).max(), d
    # identity tables reproduce the Fourier gradient away from the Nyquist planes... only where the helper mesh resolves the
    # mode exactly (ng == n): D_r = 1, D_theta = D_phi = 0, mu = 1  ->  i k_d, except that get_k keeps the Nyquist sign
    one, zero = np.ones((n, n, n // 2 + 1)), np.zeros((n, n, n // 2 + 1))
    pl.lg = M.lattice_gradient(n, n, L, one, zero, zero, one)
    olg1 = O.PLTGradient(n, n, L, one, zero, zero, one)
    Gp, vc = olg1.all(g)
    kx, ky, kz = np.broadcast_arrays(*g.kvec())
    for d, kd in enumerate((kx, ky, kz)):
        assert np.abs(Gp[d] - kd).max() < 1e-12 * np.abs(kd).max()
    ref = g.bwd(O.displacement_k(g, pots, order, 0, lunit / L, lg=olg1))
    assert np.abs(pl.displacement_field(0).get_real() - ref).max() < 1e-11 * np.abs(ref).max()
    pl.lg = None
    with pytest_raises(ValueError):
        M.lattice_gradient(ng, n, L, tabs[0][:4], tabs[1], tabs[2], tabs[3])
    # ---- (2) masked lattice
    psi = [g.bwd(O.displacement_k(g, pots, order, d, lunit / L)) for d in range(3)]
    vel = [g.bwd(O.velocity_k(g, pots, order, d, vfac, vunit / L)) for d in range(3)]
    rho = 1.0 + 0.1 * rng.standard_normal((n, n, n))
    F = M.Grid_FFT((n,) * 3, (L,) * 3)
    for do_b, mtype in ((True, 3), (True, 1), (True, 2), (True, 0), (False, 3)):
        mask = O.particle_type_mask(do_b, mtype)
        assert mask == M.particle_type_mask(do_b, mtype)
        for lidx in ((0, 1) if do_b else (0,)):
            for r64 in (True, False):
                P = M.Particles("masked", lidx, r64, True, True, lidx, F, mask=mask)
                assert P.get_local_num_particles() == int(np.sum(np.asarray(mask) == lidx)) * n ** 3 // 8
                for d in range(3):
                    F.set_real(psi[d])
                    P.set_positions(d, lunit, F)
                    F.set_real(vel[d])
                    P.set_velocities(d, F)
                F.set_real(rho)
                P.set_masses(2.5, F)
                got = P.download()
                assert np.array_equal(got["ids"], O.masked_ids(n, mask, lidx, lidx, True))
                tol = 1e-13 if r64 else 1e-6
                for d in range(3):
                    ref = O.masked_positions(g, psi[d], d, mask, lidx, lunit, r64)
                    assert np.abs(got["pos"][:, d].astype(np.float64) - ref).max() <= tol * lunit
                    ref = O.masked_velocities(g, vel[d], mask, lidx, r64)
                    if r64:
                        assert np.array_equal(got["vel"][:, d], ref)
                    else:
                        assert np.abs(got["vel"][:, d].astype(np.float64) - ref).max() <= tol * np.abs(ref).max()
                ref = O.masked_masses(g, rho, mask, lidx, 2.5, r64)
                assert np.abs(got["mass"].astype(np.float64) - ref).max() <= (1e-13 if r64 else 1e-6) * np.abs(ref).max()
    with pytest_raises(Exception):
        M.Particles("masked", 1, True, True, False, 0, F, mask=[0] * 8)     # no cell carries species 1
    # through the pipeline: ParticleLoad = masked, CDM species of a two-species mask
    cfgm = M.lpt_config(GridRes=n, BoxLength=L, LPTorder=order, ParticleLoad="masked", Dplus0=Dp, vfac=vfac, lunit=lunit, vunit=vunit,
                        DoBaryons=True, ParticleMaskType=3, ids64=True)
    plm = M.LPTPipeline(cfgm)
    plm.set_noise_k(W)
    plm.condition_noise()
    host = plm.run_device(tab).download()
    mask = O.particle_type_mask(True, 3)
    assert np.array_equal(host["ids"], O.masked_ids(n, mask, 0, 0, True))
    for d in range(3):
        ref = O.masked_positions(g, psi[d], d, mask, 0, lunit, True)
        assert np.abs(host["pos"][:, d] - ref).max() < 1e-11 * lunit
    # ---- (3) glass load
    npart = 3 * n * n + 17
    posr = rng.random((npart, 3)) * n
    posr[0] = (n - 0.25, n - 0.5, n - 1e-9)        # wraps on every axis
    posr[1] = (0.0, 0.0, 0.0)
    cfgg = M.lpt_config(GridRes=n, BoxLength=L, LPTorder=order, ParticleLoad="glass", Dplus0=Dp, vfac=vfac, lunit=lunit, vunit=vunit,
                        glass=(posr, 1000), ids64=True)
    plg = M.LPTPipeline(cfgg)
    plg.set_noise_k(W)
    plg.condition_noise()
    host = plg.run_device(tab).download()
    assert np.array_equal(host["ids"], (7 * 0 + 1000 + np.arange(npart)).astype(np.uint64))
    for d in range(3):
        fr = g.bwd(O.displacement_k(g, pots, order, d, lunit / L, glass=True))
        ref = O.glass_positions(g, fr, d, posr, lunit)
        assert np.abs(host["pos"][:, d] - ref).max() < 1e-11 * lunit, d
        fv = g.bwd(O.velocity_k(g, pots, order, d, vfac, vunit / L, glass=True))
        ref = O.glass_velocities(g, fv, posr)
        assert np.abs(host["vel"][:, d] - ref).max() < 1e-11 * np.abs(ref).max(), d
    # the compensation kernel undoes the CIC window: particles ON the lattice get the uncompensated field back only up to the
    # window, so just check that it was applied
    assert np.abs(g.bwd(O.displacement_k(g, pots, order, 0, lunit / L, glass=True)) - psi[0]).max() > 1e-3 * np.abs(psi[0]).max()
    with pytest_raises(Exception):
        M.Particles("glass", 0, True, True, True, 0, F, glass=(posr, 0))    # no per-particle masses for glasses
    with pytest_raises(Exception):
        M.Particles("glass", 0, True, True, False, 0, F, glass=(posr + n, 0))   # positions must be in grid units


# ------------------------------------------------------------------ complex grids + PPT / Eulerian branch
def check_cfft(M, n0, n1=None, n2=None):
    """Grid_FFT<ccomplex_t>: c2c forward/backward with the symmetric norm vs numpy"""
    n1, n2 = n1 or n0, n2 or n0
    rng = np.random.default_rng(n0 + 3 * n1 + 7 * n2)
    x = rng.standard_normal((n0, n1, n2)) + 1j * rng.standard_normal((n0, n1, n2))
    norm = 1.0 / np.sqrt(float(n0) * n1 * n2)
    G = M.Grid_FFT_complex((n0, n1, n2), (1.0, 2.0, 3.0))
    G.set(x)
    G.FourierTransformForward()
    assert G.space_ == M.kspace_id
    ref = np.fft.fftn(x) * norm
    assert relerr(G.get(), ref) < TOL
    G.FourierTransformForward()            # no-op when already in k-space
    assert relerr(G.get(), ref) < TOL
    G.FourierTransformBackward()
    assert relerr(G.get(), x) < TOL
    y = rng.standard_normal((n0, n1, n2)) + 1j * rng.standard_normal((n0, n1, n2))
    G.set(y, M.kspace_id)
    G.apply_gradient(1)
    i1 = np.arange(n1)
    k1 = (i1 - n1 * (i1 > n1 // 2)) * (2 * np.pi / 2.0)       # get_k: Nyquist = +n/2
    assert relerr(G.get(), y * (1j * k1[None, :, None])) < TOL
    G.set(y, M.kspace_id)
    G.FourierTransformBackward()
    assert relerr(G.get(), np.fft.ifftn(y) / norm) < TOL


def check_ppt(M, n, order, seed=31):
    """Eulerian (semiclassical PPT) density and velocity fields against the oracle (src/ic_generator.cc:685-803)"""
    L = 100.0
    Dp, vfac = 0.0257, 396.69
    g, W, Wc, phi = noise_and_phi(n, L, seed)
    amp = O.power_law_amplitude()
    amp_bc = lambda k: 2e-3 * amp(k) * k * k          # a delta_bc-like amplitude (~k^2 phi), small enough that 1+C rho > 0
    Csp, vu = 0.8433, 1.7
    pots = O.lpt_potentials(g, phi, order, Dp)
    ref, hbar_ref = O.ppt_eulerian(g, Wc, pots, order, Dp, vfac, amp_bc, Csp, vu)
    cfg = M.lpt_config(GridRes=n, BoxLength=L, LPTorder=order, ParticleLoad="sc", Dplus0=Dp, vfac=vfac)
    pl = M.LPTPipeline(cfg)
    pl.set_noise_k(W)
    pl.condition_noise()
    pl.compute_phi1(O.radial_table(g, amp))
    pl.compute_sources()
    got, hbar = pl.eulerian_fields(O.radial_table(g, amp_bc), Csp, vu)
    assert abs(hbar - hbar_ref) < 1e-12 * hbar_ref
    assert np.abs(ref["density"]).max() > 1e-3      # the test field is not trivial
    for k in ("density", "vx", "vy", "vz"):
        assert got[k].shape == (n, n, n)
        err = np.abs(got[k] - ref[k]).max() / np.abs(ref[k]).max()
        assert err < 1e-9, (k, err)
    # error behaviour: state checks instead of silent garbage
    psi = M.Grid_FFT_complex((n,) * 3, (L,) * 3)
    lib = psi._lib
    assert lib.mfb_ppt_drift(psi._h, 1.0, 1.0) != 0 and b"k-space" in lib.mfb_last_error()
    return max(np.abs(got[k] - ref[k]).max() / np.abs(ref[k]).max() for k in got)


# ------------------------------------------------------------------ streamed particle hand-off
def check_stream_handoff(M, n=32, order=2, lattice="bcc"):
    """the attribute-by-attribute hand-off on the copy stream delivers exactly what the blocking download does, also
    when the next run starts overwriting the device buffers right after (set_* waits for the pending copy)"""
    L = 100.0
    g, W, Wc, phi = noise_and_phi(n, L, 5)
    amp = O.power_law_amplitude()
    cfg = M.lpt_config(GridRes=n, BoxLength=L, LPTorder=order, ParticleLoad=lattice, Dplus0=0.03, vfac=300.0,
                       lunit=L, vunit=L, real64=False, ids64=False)
    pl = M.LPTPipeline(cfg)
    tab = O.radial_table(g, amp)

    def run(stream_to=None):
        pl.set_noise_k(W)
        pl.condition_noise()
        return pl.run_device(tab, particles=True, stream_to=stream_to)
    ref = run().download()
    npart = pl.particles.get_local_num_particles()
    out = {"pos": np.zeros((npart, 3), np.float32), "vel": np.zeros((npart, 3), np.float32), "ids": np.zeros(npart, np.uint32)}
    run(out)
    run()                      # queued behind the streamed copies; must not corrupt them
    pl.particles.stream_wait()
    for k in ("pos", "vel", "ids"):
        assert np.array_equal(out[k], ref[k]), k
    with pytest_raises(ValueError):
        pl.particles.stream_out("pos", np.zeros(5, np.float32))
    # the ingest side: a series of runs whose noise is streamed in under the previous run (the noise grid is dead once phi1
    # exists) delivers what blocking uploads deliver
    w = pl.wnoise
    if w.local_0_size_ == n:   # single slab: the real-space field of the second run, raw layout
        rng = np.random.default_rng(77)
        noise = [rng.standard_normal((n, n, n)) for _ in range(2)]
        raw = [np.zeros(w.ntot_) for _ in range(2)]
        for r_, x in zip(raw, noise):
            r_.reshape(n, n, n + 2)[:, :, :n] = x
        res = []
        for streamed in (False, True):
            outs = []
            for k in range(2):
                if streamed and k > 0:
                    w.stream_in_wait()
                    w.FourierTransformBackward(False)
                else:
                    w.FourierTransformBackward(False)
                    w.upload_raw(raw[k])
                w.FourierTransformForward()
                pl.condition_noise()
                pl.compute_phi1(tab)
                if streamed and k == 0:
                    w.stream_in(raw[1])
                pl.compute_sources()
                outs.append(pl.fill_particles().download())
            res.append(outs)
        for k in range(2):
            for key in ("pos", "vel"):
                assert np.array_equal(res[0][k][key], res[1][k][key]), (k, key)
        assert not np.array_equal(res[0][0]["pos"], res[0][1]["pos"])
        with pytest_raises(ValueError):
            w.stream_in(np.zeros(5))


def check_fused_readout(M, n=32):
    """single-shift (sc) loads handed a k-space field: the lattice read-out as the epilogue of the c2r pass delivers bit for bit
    what the two-pass sequence (backward transform, read-out kernel) delivers - positions, velocities, f32 and f64; other
    lattices take the two-pass route whatever the option says"""
    L = 100.0
    lib = M.grid._lib()
    g, W, Wc, phi = noise_and_phi(n, L, 13)
    tab = O.radial_table(g, O.power_law_amplitude())
    for lattice, r64 in (("sc", True), ("sc", False), ("bcc", True)):
        res = []
        try:
            for fused in (1, 0):
                assert lib.mfb_set_option(b"fused_readout", fused) == 0
                cfg = M.lpt_config(GridRes=n, BoxLength=L, LPTorder=2, ParticleLoad=lattice, Dplus0=0.03, vfac=300.0, lunit=L, vunit=L,
                                   real64=r64, ids64=False)
                pl = M.LPTPipeline(cfg)
                pl.set_noise_k(W)
                pl.condition_noise()
                l0 = M.kernel_launches()
                host = pl.run_device(tab).download()
                res.append((host, M.kernel_launches() - l0))
        finally:
            lib.mfb_set_option(b"fused_readout", 1)
        for k in ("pos", "vel", "ids"):
            assert np.array_equal(res[0][0][k], res[1][0][k]), (lattice, r64, k)
        if lattice == "sc":
            assert res[0][1] == res[1][1] - 6, (res[0][1], res[1][1])     # six read-out launches fewer
        else:
            assert res[0][1] == res[1][1]


class pytest_raises:
    def __init__(self, exc):
        self.exc = exc

    def __enter__(self):
        return self

    def __exit__(self, et, ev, tb):
        assert et is not None and issubclass(et, self.exc), f"expected {self.exc}"
        return True


# ------------------------------------------------------------------ whole pipeline vs oracle
def check_pipeline(M, n, order, lattice="sc", real64=True, seed=21):
    L = 100.0
    Dp, vfac = 0.0257, 396.69
    lunit, vunit = L, L / np.sqrt(1 / 51.0)
    g, W, Wc, phi = noise_and_phi(n, L, seed)
    amp = O.power_law_amplitude()
    ref = O.run_pipeline(n, L, Wc, amp, order, Dp, vfac, lattice, lunit, vunit, real64, ids64=True)
    cfg = M.lpt_config(GridRes=n, BoxLength=L, LPTorder=order, ParticleLoad=lattice, Dplus0=Dp, vfac=vfac,
                       lunit=lunit, vunit=vunit, real64=real64, ids64=True)
    pl = M.LPTPipeline(cfg)
    pl.set_noise_k(W)
    pl.condition_noise()
    parts = pl.run_device(O.radial_table(g, amp), particles=True)
    # potentials
    scale = np.abs(ref["pots"]["phi"]).max()
    assert relerr(pl.phi.get_kspace(), ref["pots"]["phi"]) < TOL
    if order > 1:
        assert relerr(pl.phi2.get_kspace(), ref["pots"]["phi2"]) < TOL
    if order > 2:
        assert relerr(pl.phi3.get_kspace(), ref["pots"]["phi3"]) < 10 * TOL
        for d in range(3):
            assert relerr(pl.A3[d].get_kspace(), ref["pots"]["A3"][d]) < 10 * TOL
    host = parts.download()
    assert np.array_equal(host["ids"], ref["ids"])
    maxpsi = max(np.abs(ref["fields"]["d" + c]).max() for c in "xyz")
    maxv = max(np.abs(ref["fields"]["v" + c]).max() for c in "xyz")
    if real64:
        # north_star: max|err| <= 1e-9 * max|Psi|; we hold 1e-11
        assert np.abs(host["pos"] - ref["pos"]).max() < 1e-11 * maxpsi
        assert np.abs(host["vel"] - ref["vel"]).max() < 1e-11 * maxv
    else:
        assert host["pos"].dtype == np.float32
        assert (host["pos"] == ref["pos"]).mean() > 0.999
        assert np.allclose(host["pos"], ref["pos"], rtol=1e-6, atol=1e-6 * maxpsi)
        assert np.allclose(host["vel"], ref["vel"], rtol=1e-6, atol=1e-6 * maxv)
    return pl


def check_lattice_positions_bitexact(M, n, lattice):
    """zero displacement: positions are the bare lattice coordinates, bit-for-bit"""
    L = 100.0
    g = O.GridSpec(n, L)
    F = M.Grid_FFT((n,) * 3, (L,) * 3)
    lt = O.LATTICE_TYPES[lattice]
    for lattice_index in (0, 1):
        P = M.Particles(lattice, lattice_index, True, True, False, lattice_index, F)
        for d in range(3):
            F.set_real(np.zeros((n,) * 3))
            P.set_positions(d, 7.3, F)
        host = P.download(vel=False)
        for d in range(3):
            ref = O.set_positions(g, np.zeros((n,) * 3), d, lt, lattice_index, 7.3, True)
            # shift_field of an all-zero field stays all-zero -> exact lattice
            assert np.array_equal(host["pos"][:, d], ref)
        assert np.array_equal(host["ids"], O.particle_ids(n, lt, lattice_index, True))


# ------------------------------------------------------------------ goldens through the CUDA path
def _golden_pipeline(M, order, lattice, fmt):
    g, phi_s, vfac = G.bootstrap()
    lunit, vunit, r64 = G.units(fmt)
    cfg = M.lpt_config(GridRes=G.N, BoxLength=G.L, LPTorder=order, ParticleLoad=lattice, Dplus0=-1.0, vfac=vfac,
                       lunit=lunit, vunit=vunit, real64=r64, ids64=False)
    pl = M.LPTPipeline(cfg)
    pl.phi.set_kspace(phi_s)
    pl.compute_sources()
    return pl


NOISE_FLOOR = 8.0 * np.finfo(np.float64).eps  # of max|ref|


def close(a, b):
    """The reference comparator (tests/scripts/compare_hdf5.py:66-75): element-wise
    |a-b| <= atol + rtol*|b| with rtol=1e-9, atol=1e-14.  An element whose |ref| is ~1e-7 of the field
    maximum can only meet that if the *absolute* error is below one ulp of the field scale, which no
    FFT with a different summation order guarantees (on the B200: 1 of 196608 SWIFT velocities, |ref| =
    2e-7 max, |err| = 9e-16 max, ratio 1.03; profiles/r01_golden_diag_gpu.txt).  Such elements are accepted
    iff their error is inside the rounding-noise floor |err| <= 8 eps * max|ref|; at most 1e-5 of the
    elements may need that.  The north_star bound max|err| <= 1e-9 max|ref| is asserted on top."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    err = np.abs(a - b)
    scale = np.abs(b).max()
    bad = err > G.ATOL + G.RTOL * np.abs(b)
    if err.max() > 1e-9 * scale:
        return False
    if not bad.any():
        return True
    return bool(bad.sum() <= max(1, int(1e-5 * bad.size)) and err[bad].max() <= NOISE_FLOOR * scale)


def check_golden_generic(M, name, order):
    gold = G.load(name)
    pl = _golden_pipeline(M, order, "sc", "generic")
    for d, c in enumerate("xyz"):
        assert close(pl.displacement_field(d).get_real(), gold["DM_d" + c]), c
    for d, c in enumerate("xyz"):
        assert close(pl.velocity_field(d).get_real(), gold["DM_v" + c]), c
    assert close(pl.density_field().get_real(), gold["DM_delta"])


def check_golden_from_seed(M, name, order):
    """The golden files from the SEED: the white noise is generated on the device by the reference's own generator
    (mfb_noise_fill_ngenic_ranlxd, [random] seed = 12345 of the reference's regression configurations); the only thing taken
    from a golden file is the radial amplitude A(|k|) - delta_k / W_k per |k| shell of the 1LPT run, a function of |n|^2 -,
    tabulated like a transfer-function plugin would be.  Everything else (noise, phi(1), sources, fields) is the device path."""
    g = O.GridSpec(G.N, G.L)
    _, _, vfac = G.bootstrap()
    W = O.ngenic_noise(G.N, 12345)
    dk = g.fwd(G.load("test_1lpt_sc_generic")["DM_delta"])
    m = np.abs(W) > 0
    shell = np.round(g.k2() / (2.0 * np.pi / G.L) ** 2).astype(np.int64)
    h = G.N // 2
    table = np.zeros(3 * h * h + 1)
    cnt = np.zeros_like(table)
    np.add.at(table, shell[m], (dk[m] / W[m]).real)
    np.add.at(cnt, shell[m], 1.0)
    table = -np.divide(table, cnt, out=np.zeros_like(table), where=cnt > 0)      # phi(1) = -W A / k^2  (D+ := -1, as in bootstrap)
    cfg = M.lpt_config(GridRes=G.N, BoxLength=G.L, LPTorder=order, ParticleLoad="sc", Dplus0=-1.0, vfac=vfac, lunit=1.0, vunit=1.0,
                       real64=True, ids64=False)
    pl = M.LPTPipeline(cfg)
    pl.wnoise.fill_noise_ngenic(12345)
    pl.compute_phi1(table)
    pl.compute_sources()
    gold = G.load(name)
    for d, c in enumerate("xyz"):
        assert close(pl.displacement_field(d).get_real(), gold["DM_d" + c]), c
    for d, c in enumerate("xyz"):
        assert close(pl.velocity_field(d).get_real(), gold["DM_v" + c]), c
    assert close(pl.density_field().get_real(), gold["DM_delta"])


def check_golden_gadget(M):
    gold = G.load("test_2lpt_baryons_vrel_gadget")
    gold_dm = G.load("test_2lpt_sc_gadget")
    pl = _golden_pipeline(M, 2, "sc", "gadget")
    p = pl.fill_particles(0).download()
    assert close(p["pos"], gold_dm["PartType1__Coordinates"]) and close(p["vel"], gold_dm["PartType1__Velocities"])
    assert close(p["pos"], gold["PartType1__Coordinates"]) and close(p["vel"], gold["PartType1__Velocities"])
    assert np.array_equal(p["ids"], gold_dm["PartType1__ParticleIDs"])
    # baryons: secondary lattice (index 1), IDs offset by N^3, uniform masses through set_masses
    pl.particles = None
    pl.make_particles(lattice_index=1, with_masses=True, id_offset=1)
    pb = pl.fill_particles(1)
    munit = 27.7519737 * G.L ** 3
    rho = M.Grid_FFT(pl.N, pl.Lbox)
    rho.set_real(np.zeros((G.N,) * 3))
    rho.affine(0.048891054 * munit, 0.048891054 * munit * 0.5)  # (1 + C*rho)*Omega*munit with rho == 0
    pb.set_masses(1.0, rho)
    b = pb.download()
    assert close(b["pos"], gold["PartType0__Coordinates"]) and close(b["vel"], gold["PartType0__Velocities"])
    assert np.array_equal(b["ids"], gold["PartType0__ParticleIDs"])
    assert close(b["mass"], gold["PartType0__Masses"])


def check_golden_swift(M):
    gold = G.load("test_3lpt_bcc_swift")
    pl = _golden_pipeline(M, 3, "bcc", "swift")
    p = pl.fill_particles(0).download()
    assert p["pos"].shape == (65536, 3)
    assert close(p["pos"], gold["PartType1__Coordinates"])
    assert close(p["vel"], gold["PartType1__Velocities"])


# ------------------------------------------------------------------ non-cubic product (exercises every z_conv / strided plan)
def _pad_idx(n, p):
    i = np.array([k for k in range(n) if k != n // 2])
    return i, np.where(i < n // 2, i, i + p - n)


def check_multiply_noncubic(M, shape):
    """multiply_field on an (n0,n1,n2) grid against a direct numpy statement of
    unpad(fwd_P(bwd_P(pad a) * bwd_P(pad b))) with Nyquist planes dropped (MPI semantics)."""
    n = np.array(shape)
    p = 3 * n // 2
    rng = np.random.default_rng(int(n.sum()))
    try:
        import scipy.fft as sfft
        rfftn = lambda x: sfft.rfftn(x, workers=-1)
        irfftn = lambda X, s: sfft.irfftn(X, s=s, axes=(0, 1, 2), workers=-1)
    except ImportError:
        rfftn = np.fft.rfftn
        irfftn = lambda X, s: np.fft.irfftn(X, s=s, axes=(0, 1, 2))
    def spec():
        return rfftn(rng.standard_normal(shape)) / np.sqrt(float(n.prod()))
    A, B = spec(), spec()
    (s0, d0), (s1, d1) = _pad_idx(n[0], p[0]), _pad_idx(n[1], p[1])
    kz = np.arange(n[2] // 2)
    def pad(X):
        Y = np.zeros((p[0], p[1], p[2] // 2 + 1), complex)
        Y[np.ix_(d0, d1, kz)] = X[np.ix_(s0, s1, kz)] * 1.5 ** 1.5
        return Y
    sp = np.sqrt(float(p.prod()))
    a = irfftn(pad(A), tuple(p)) * sp
    b = irfftn(pad(B), tuple(p)) * sp
    Cp = rfftn(a * b) / sp
    del a, b
    ref = np.zeros_like(A)
    ref[np.ix_(s0, s1, kz)] = Cp[np.ix_(d0, d1, kz)] / (sp / np.sqrt(float(n.prod())))
    L = (100.0, 100.0, 100.0)
    P, Q = M.Grid_FFT(shape, L), M.Grid_FFT(shape, L)
    out = M.Grid_FFT(shape, L, True, M.kspace_id)
    P.set_kspace(A)
    Q.set_kspace(B)
    C = M.OrszagConvolver(shape, L)
    C.multiply_field(P, Q, M.op.assign_to(out))
    assert relerr(out.get_kspace(), ref) < TOL
    # the read-modify-write output operators take their own plan in the last pass (8-line tiles at n = 1536)
    C.multiply_field(P, Q, M.op.multiply_add_to(out, 0.5))
    assert relerr(out.get_kspace(), 1.5 * ref) < TOL


def check_sources_shared_field_pass(M, shape):
    import ctypes
    """third-order sources on an (n0,n1,n2) grid: the shared-field z pass (z_fields_kernel - every Hessian component of phi and
    phi2 transformed once per row, targets formed pointwise) against the multi-product z pass (one transform per product
    operand).  Slim grids with a long z axis reach the plans of the BASELINE sizes (n2 = 1024 -> 1536-point lines)."""
    lib = M.grid._lib()
    L = (100.0, 100.0, 100.0)
    rng = np.random.default_rng(int(np.sum(shape)))
    phi0 = np.fft.rfftn(rng.standard_normal(shape)) / np.sqrt(float(np.prod(shape)))
    res = {}
    try:
        for shared in (1, 0):
            assert lib.mfb_set_option(b"sources_shared", shared) == 0
            grids = [M.Grid_FFT(shape, L, True, M.kspace_id) for _ in range(6)]
            phi, phi2, phi3, A3 = grids[0], grids[1], grids[2], grids[3:]
            phi.set_kspace(phi0)
            conv = M.OrszagConvolver(shape, L)
            A3h = (ctypes.c_void_p * 3)(*[g._h for g in A3])
            assert lib.mfb_lpt_sources(conv._h, 3, 0.7, phi._h, phi2._h, phi3._h, A3h, None) == 0, lib.mfb_last_error()
            info = conv.sources_info()
            assert info["path"] == 2 and info["loops"] == (3 if shared else 4), info
            res[shared] = [g.get_kspace() for g in grids]
            del conv, grids, phi, phi2, phi3, A3
    finally:
        lib.mfb_set_option(b"sources_shared", 1)
    # the sources are total divergences: their low-k modes are sums that cancel to (k_min / k_max)^2 of the typical term, and
    # 1 / k^2 then makes exactly those modes the largest of phi2 - any two summation orders differ by eps * (n_max / 2)^2 there
    # (64 x 32 x 512: shared-field vs per-product 3.3e-12, multi-product vs per-product 5.1e-12, shared vs multi 5.2e-12)
    tol = max(1e-12, 1e-16 * float(max(shape)) ** 2)
    for a, b in zip(res[1], res[0]):
        assert relerr(a, b) < tol, (relerr(a, b), tol)


# ------------------------------------------------------------------ kernel generations agree
def check_product_grouping(M, n=32):
    """3LPT sources three ways - (2) shared x passes + x-plane chunks + multi-product z pass with the products of a target
    summed in real space (mfb_sources2.cuh), (1) per-product transforms with the products of a target summed in z-Fourier
    space before ONE forward transform, (0) the reference's one-forward-transform-per-product sequence: equal up to rounding"""
    L, Dp = 100.0, 0.0257
    lib = M.grid._lib()
    g, W, Wc, phi = noise_and_phi(n, L, 23)
    amp = O.power_law_amplitude()
    res = {}
    try:
        # (v2, grouping, max_slots, chunk planes): the last two make the slot workspace as tight as it may be (7 slots: one
        # target per loop, an evicted X buffer is made again) and the chunks ragged
        # "v2direct": the multi-product z pass with direct loads instead of TMA-staged rows - the same arithmetic, bit for bit
        # "v2": stages A and C as shared-field loops (z_fields_kernel: every Hessian component transformed once per row, the
        # last stage's targets overwrite consumed X planes); "v2multi": the multi-product z pass in every stage; "v2noalias":
        # the shared last stage with a slot per target (10 slots, what slabs do)
        for key, v2, grouping, slots, chunk in (("v2", 1, 1, 0, 0), ("v2multi", 1, 1, 0, 0), ("v2direct", 1, 1, 0, 0), ("v2noalias", 1, 1, 0, 3),
                                                ("v2tight", 1, 1, 7, 5), ("v2split", 1, 1, 5, 7), (1, 0, 1, 0, 0), (0, 0, 0, 0, 0)):
            assert lib.mfb_set_option(b"sources_v2", v2) == 0 and lib.mfb_set_option(b"conv_grouping", grouping) == 0
            assert lib.mfb_set_option(b"zm_prefetch", 0 if key == "v2direct" else 1) == 0
            assert lib.mfb_set_option(b"sources_shared", 0 if key in ("v2multi", "v2direct") else 1) == 0
            assert lib.mfb_set_option(b"sources_alias", 0 if key == "v2noalias" else 1) == 0
            assert lib.mfb_set_option(b"sources_max_slots", slots) == 0 and lib.mfb_set_option(b"sources_chunk", chunk) == 0
            pl = M.LPTPipeline(M.lpt_config(GridRes=n, BoxLength=L, LPTorder=3, Dplus0=Dp))
            pl.set_noise_k(W)
            pl.condition_noise()
            pl.compute_phi1(O.radial_table(g, amp))
            l0 = M.kernel_launches()
            pl.compute_sources()
            info = pl.conv.sources_info()
            res[key] = ([pl.phi2.get_kspace(), pl.phi3.get_kspace()] + [a.get_kspace() for a in pl.A3],
                        M.kernel_launches() - l0, pl.conv.workspace_bytes(), info)
            # named products after the source stage: the convolver goes back to its per-product workspace
            out = M.Grid_FFT((n,) * 3, (L,) * 3, True, M.kspace_id)
            pl.conv.convolve_Hessians(pl.phi, (0, 1), pl.phi2, (1, 2), M.op.assign_to(out))
            res[key] += (out.get_kspace(),)
    finally:
        for o_ in (b"sources_max_slots", b"sources_chunk"):
            lib.mfb_set_option(o_, 0)
        lib.mfb_set_option(b"zm_prefetch", 1)
        lib.mfb_set_option(b"sources_shared", 1)
        lib.mfb_set_option(b"sources_alias", 1)
        lib.mfb_set_option(b"conv_grouping", 1)
        lib.mfb_set_option(b"sources_v2", 1)
    print({k: (v[3], v[1]) for k, v in res.items()})
    assert res["v2"][3]["path"] == 2 and res["v2tight"][3]["path"] == 2 and res[1][3]["path"] == 1 and res[0][3]["path"] == 1
    # 10 backward x passes (3 of phi, 3 of phi2, 4 inner products) where the per-product path does 49; with 8 slots the
    # four targets of the last stage go two per loop (multi-product pass); with 7 the first stage takes two multi-product
    # loops, X(phi, 1) is made twice, and the last stage still is one shared-field loop (6 X buffers, targets in their planes)
    assert res["v2multi"][3]["x_passes"] == 10 and res["v2multi"][3]["loops"] == 4 and res["v2multi"][3]["slots"] >= 8, res["v2multi"][3]
    # shared-field loops: the five targets of the first stage and the four of the last one in one loop each
    assert res["v2"][3]["x_passes"] == 10 and res["v2"][3]["loops"] == 3 and res["v2"][3]["slots"] == 8, res["v2"][3]
    assert res["v2"][3]["y_passes"] < res["v2multi"][3]["y_passes"] and res["v2"][3]["z_launches"] < res["v2multi"][3]["z_launches"]
    assert res["v2noalias"][3]["loops"] == 3 and res["v2noalias"][3]["slots"] == 10, res["v2noalias"][3]
    assert res["v2tight"][3]["slots"] == 7 and res["v2tight"][3]["x_passes"] == 11 and res["v2tight"][3]["loops"] == 4, res["v2tight"][3]
    # equal up to rounding: the sums are taken in a different order (real space / z-Fourier space / per product)
    assert res["v2split"][3]["path"] == 2 and res["v2split"][3]["slots"] == 5 and res["v2split"][3]["loops"] > 7, res["v2split"][3]
    for a, b in zip(res["v2multi"][0], res["v2direct"][0]):
        assert np.array_equal(a, b)
    for key in ("v2", "v2multi", "v2noalias", "v2tight", "v2split", 1):
        for a, b in zip(res[key][0], res[0][0]):
            assert relerr(a, b) < 1e-12, key
        assert relerr(res[key][4], res[0][4]) < 1e-12
    # memory-pressure path: the idle particle arrays are lent to the convolver - as the accumulator of product groups on the
    # per-product path (v2 = 0: no own accumulator -> ungrouped until the loan), as slots of the workspace on the v2 path
    for v2 in (0, 1):
        try:
            assert lib.mfb_set_option(b"conv_own_accumulator", 0) == 0 and lib.mfb_set_option(b"sources_v2", v2) == 0
            pl = M.LPTPipeline(M.lpt_config(GridRes=n, BoxLength=L, LPTorder=3, Dplus0=Dp, ParticleLoad="sc", real64=False))
            assert not pl.conv.grouped()
            pl.make_particles()
            assert pl.conv.grouped()
            for rep in range(2):          # second run: the loan is taken back from live particle data + pending hand-off copies
                pl.set_noise_k(W)
                pl.condition_noise()
                npart = pl.particles.get_local_num_particles()
                out = {"pos": np.zeros((npart, 3), np.float32), "vel": np.zeros((npart, 3), np.float32)}
                pl.run_device(O.radial_table(g, amp), particles=True, stream_to=out)
            pl.particles.stream_wait()
            assert pl.conv.sources_info()["path"] == 1 + v2
            if v2:
                # the second run found hand-off copies pending on the lent slots: its first two stages stayed on the other
                # slots (more, narrower loops) and the lent ones were waited for only when the third stage took them
                assert pl.conv.sources_info()["loops"] > 4, pl.conv.sources_info()
            got = [pl.phi2.get_kspace(), pl.phi3.get_kspace()] + [a.get_kspace() for a in pl.A3]
            assert np.array_equal(out["pos"], pl.particles.download()["pos"])
            # the loan ends with the lender: destroying the particles leaves a convolver that still works on what it owns
            pl.particles = None
            assert not pl.conv.grouped()
            pl.set_noise_k(W)
            pl.condition_noise()
            pl.compute_phi1(O.radial_table(g, amp))
            pl.compute_sources()
            assert relerr(pl.phi3.get_kspace(), got[1]) < 1e-12
        finally:
            lib.mfb_set_option(b"conv_own_accumulator", 1)
            lib.mfb_set_option(b"sources_v2", 1)
        for a, b in zip(got, res[1][0]):
            assert relerr(a, b) < 1e-12
    R = 8 * n * n * (n + 2)
    if res[1][2] > 7.5 * R:      # the accumulator was granted: 32 products instead of 33, 9 forward passes instead of 33
        assert res[1][1] < res[0][1] - 50, (res[1][1], res[0][1])


def check_kernel_generations(M, n):
    """the general strided pass (strided_pass_kernel: run-time pad / truncate maps, any size) and its compile-time
    specialisations (strided2_kernel) implement the same arithmetic: a dealiased product and an FFT round trip agree to
    rounding.  (Older generations of the z passes live in scripts/kbench/kbench_legacy_kernels.cuh.)"""
    L = 100.0
    lib = M.grid._lib()
    g, W, Wc, phi = noise_and_phi(n, L, 9, False)
    _, _, _, psi = noise_and_phi(n, L, 10, False)
    res = {}
    try:
        for gen in (0, 1):
            assert lib.mfb_set_option(b"strided_v2", gen) == 0
            P, Q = M.Grid_FFT((n,) * 3, (L,) * 3), M.Grid_FFT((n,) * 3, (L,) * 3)
            out = M.Grid_FFT((n,) * 3, (L,) * 3, True, M.kspace_id)
            P.set_kspace(phi)
            Q.set_kspace(psi)
            C = M.OrszagConvolver((n,) * 3, (L,) * 3)
            C.convolve_Hessians(P, (0, 1), Q, (1, 2), M.op.assign_to(out))
            C.convolve_SumOfHessians(P, (2, 2), Q, (0, 0), (1, 1), M.op.multiply_add_to(out, -0.5))
            P.FourierTransformBackward()
            res[gen] = (out.get_kspace(), P.get_real())
    finally:
        lib.mfb_set_option(b"strided_v2", 1)
    assert relerr(res[1][0], res[0][0]) < TOL and relerr(res[1][1], res[0][1]) < TOL
    assert lib.mfb_set_option(b"no_such_option", 1) != 0
