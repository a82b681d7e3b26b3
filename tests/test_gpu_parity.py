"""GPU parity tests proper: the sm_100a library through the C ABI against the CPU oracle, the
reference goldens, and size-independent properties at full BASELINE sizes."""
import numpy as np
import pytest

import parity_cases as PC

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import monofonic_b200 as MM

    MM.init(0)
    assert MM.load().mfb_build_info() == b"cuda sm_100a"
    return MM


@pytest.mark.parametrize("n", [32, 48, 64, 96, 128, 192, 256, 384])
def test_fft_vs_numpy(M, n):
    PC.check_fft(M, n)


@pytest.mark.parametrize("shape", [(512, 32, 48), (32, 768, 64), (48, 32, 1024), (1536, 32, 32), (32, 32, 2048),
                                   (3072, 32, 32), (32, 2048, 32), (32, 32, 3072), (1024, 48, 32), (64, 1536, 32)])
def test_fft_every_plan(M, shape):
    PC.check_fft(M, *shape)


def test_c2r_nonhermitian(M):
    PC.check_c2r_nonhermitian(M, 64)


@pytest.mark.parametrize("n", [32, 64])
def test_kspace_operators(M, n):
    PC.check_kops(M, n)


@pytest.mark.parametrize("n,nyq0", [(32, True), (32, False), (64, True), (128, False)])
def test_convolver(M, n, nyq0):
    PC.check_convolver(M, n, nyq0)


@pytest.mark.parametrize("n,order,lattice,real64", [(32, 1, "sc", True), (32, 2, "sc", False), (32, 3, "bcc", True),
                                                    (64, 3, "fcc", True), (128, 3, "sc", False), (128, 2, "rsc", True)])
def test_pipeline_vs_oracle(M, n, order, lattice, real64):
    PC.check_pipeline(M, n, order, lattice, real64)


@pytest.mark.parametrize("n,nyq0", [(32, False), (64, True), (128, False)])
def test_naive_convolver(M, n, nyq0):
    PC.check_naive_convolver(M, n, nyq0)


@pytest.mark.parametrize("shape", [(32, 32, 32), (64, 48, 96), (256, 32, 128), (32, 384, 32), (32, 32, 1024), (512, 32, 32)])
def test_complex_grid_fft(M, shape):
    PC.check_cfft(M, *shape)


@pytest.mark.parametrize("n,order", [(32, 1), (64, 2), (128, 3)])
def test_ppt_eulerian(M, n, order):
    PC.check_ppt(M, n, order)


@pytest.mark.parametrize("n", [32, 128])
def test_product_grouping(M, n):
    PC.check_product_grouping(M, n)


@pytest.mark.parametrize("shape", [(32, 64, 64), (64, 64, 256), (64, 32, 512), (32, 64, 1024)])
def test_sources_shared_field_pass(M, shape):
    """every plan of the shared-field z pass up to the 1536-point lines of the 1024^3 runs, against the multi-product pass"""
    PC.check_sources_shared_field_pass(M, shape)


@pytest.mark.parametrize("n,lattice", [(32, "bcc"), (64, "sc")])
def test_two_fluid_terms(M, n, lattice):
    PC.check_two_fluid_terms(M, n, lattice)


@pytest.mark.parametrize("n,order", [(32, 3), (64, 2)])
def test_plt_masked_glass(M, n, order):
    PC.check_plt_masked_glass(M, n, order)


@pytest.mark.parametrize("n", [32, 128])
def test_noise_ingest(M, n):
    PC.check_noise_ingest(M, n)


@pytest.mark.parametrize("n", [32, 256])
def test_fused_readout(M, n):
    PC.check_fused_readout(M, n)


def test_f4_against_reference_produced_fixture(M):
    PC.check_reference_f4(M)


def test_against_reference_produced_fixture(M):
    PC.check_reference_fixture(M)


def test_particles_against_reference_produced_fixture(M):
    PC.check_reference_particles(M)


def test_naive_pipeline(M):
    PC.check_naive_pipeline(M, 64)


def test_stream_handoff(M):
    PC.check_stream_handoff(M, 64, 3, "bcc")


@pytest.mark.parametrize("lattice", ["sc", "bcc", "fcc", "rsc"])
def test_lattice_bitexact(M, lattice):
    PC.check_lattice_positions_bitexact(M, 64, lattice)


def test_golden_1lpt_generic(M):
    PC.check_golden_generic(M, "test_1lpt_sc_generic", 1)


def test_golden_2lpt_generic(M):
    PC.check_golden_generic(M, "test_2lpt_baryons_generic", 2)


@pytest.mark.parametrize("name,order", [("test_1lpt_sc_generic", 1), ("test_2lpt_baryons_generic", 2)])
def test_golden_from_seed(M, name, order):
    PC.check_golden_from_seed(M, name, order)


def test_golden_gadget(M):
    PC.check_golden_gadget(M)


def test_golden_swift(M):
    PC.check_golden_swift(M)


# ---- BASELINE sizes: size-independent properties (the oracle would take minutes here)
@pytest.mark.parametrize("n", [512, 768])
def test_fft_roundtrip_and_parseval_large(M, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((n, n, n))
    G = M.Grid_FFT((n,) * 3, (100.0,) * 3)
    G.set_real(x)
    s2 = float((x * x).sum())
    G.FourierTransformForward()
    X = G.get_kspace()
    # Parseval with the symmetric normalisation: sum|x|^2 == sum over the full spectrum |X|^2
    w = np.full(n // 2 + 1, 2.0)
    w[0] = w[-1] = 1.0
    p2 = float(((X.real ** 2 + X.imag ** 2) * w).sum())
    assert abs(p2 - s2) < 1e-12 * s2
    # a few modes against a direct DFT sum
    j = np.arange(n)
    for (a, b, c) in [(0, 0, 0), (1, 2, 3), (n - 1, n // 2, n // 2), (17, n - 5, 40)]:
        e = np.exp(-2j * np.pi * (a * j[:, None, None] + b * j[None, :, None] + c * j[None, None, :]) / n)
        ref = (x * e).sum() / np.sqrt(float(n) ** 3)
        assert abs(X[a, b, c] - ref) < 1e-11 * np.sqrt(s2 / n ** 3) * np.sqrt(float(n) ** 3) ** 0 + 1e-9
    del X
    G.FourierTransformBackward()
    assert np.abs(G.get_real() - x).max() < 1e-12 * np.abs(x).max()


@pytest.mark.parametrize("shape", [(1024, 32, 32), (32, 1024, 32), (32, 32, 1024), (2048, 32, 32), (32, 2048, 32), (512, 64, 32),
                                   (1024, 128, 512)])
def test_dealiased_product_every_long_plan(M, shape):
    """every pad / truncate / fused-z plan of the BASELINE sizes (1024 -> 1536, 2048 -> 3072 points per line) on slim grids,
    against a direct numpy statement of the dealiased product; the last shape has the megabyte position stride of a real
    x pass, which selects the 8-line plans"""
    PC.check_multiply_noncubic(M, shape)


def test_complex_grid_fft_large_noncubic(M):
    """Grid_FFT<ccomplex_t> at 384 x 256 x 512 (three different plans, 0.8 GB): c2c forward against numpy, Parseval,
    round trip"""
    shape = (384, 256, 512)
    rng = np.random.default_rng(7)
    x = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    norm = 1.0 / np.sqrt(float(np.prod(shape)))
    G = M.Grid_FFT_complex(shape, (100.0, 50.0, 200.0))
    G.set(x)
    G.FourierTransformForward()
    X = G.get()
    s2 = float((x.real ** 2 + x.imag ** 2).sum())
    assert abs(float((X.real ** 2 + X.imag ** 2).sum()) - s2) < 1e-12 * s2
    ref = np.fft.fftn(x) * norm
    assert np.abs(X - ref).max() < 1e-12 * np.abs(ref).max()
    del X, ref
    G.FourierTransformBackward()
    assert np.abs(G.get() - x).max() < 1e-12 * np.abs(x).max()


def test_ppt_eulerian_256(M):
    """PPT / Eulerian branch (config C4's sub-case) at 256^3 against the oracle (first order: the oracle needs no
    dealiased products and finishes in seconds; the kick is covered at 64^3 / 128^3)"""
    assert PC.check_ppt(M, 256, 1) < 1e-9


def test_convolver_linearity_and_product_512(M):
    """2LPT-size (512^3, padded 768^3): conv(a,b) is bilinear and equals the alias-free product
    for band-limited fields (checked on random modes by direct evaluation)."""
    n, L = 512, 100.0
    rng = np.random.default_rng(1)
    # band-limited (|n| <= 3) random real fields -> the exact product spectrum is cheap to get on the host
    def small_field():
        X = np.zeros((n, n, n // 2 + 1), complex)
        idx = [0, 1, 2, 3, n - 3, n - 2, n - 1]
        for a in idx:
            for b in idx:
                for c in range(4):
                    X[a, b, c] = rng.standard_normal() + 1j * rng.standard_normal()
        # Hermitian symmetry of the kz=0 plane
        for a in idx:
            for b in idx:
                X[(-a) % n, (-b) % n, 0] = np.conj(X[a, b, 0])
        X[0, 0, 0] = X[0, 0, 0].real
        return X
    A, B = small_field(), small_field()
    P, Q = M.Grid_FFT((n,) * 3, (L,) * 3), M.Grid_FFT((n,) * 3, (L,) * 3)
    out = M.Grid_FFT((n,) * 3, (L,) * 3, True, M.kspace_id)
    P.set_kspace(A)
    Q.set_kspace(B)
    C = M.OrszagConvolver((n,) * 3, (L,) * 3)
    C.multiply_field(P, Q, M.op.assign_to(out))
    got = out.get_kspace()
    # reference on a 16^3 grid (all modes |n|<=6 fit): same symmetric normalisation per grid size
    m = 16
    def to_small(X):
        S = np.zeros((m, m, m // 2 + 1), complex)
        for a in range(-3, 4):
            for b in range(-3, 4):
                S[a % m, b % m, :4] = X[a % n, b % n, :4]
        return S
    a_r = np.fft.irfftn(to_small(A), s=(m,) * 3, axes=(0, 1, 2)) * m ** 3
    b_r = np.fft.irfftn(to_small(B), s=(m,) * 3, axes=(0, 1, 2)) * m ** 3
    prod = np.fft.rfftn(a_r * b_r) / m ** 3 / np.sqrt(float(n) ** 3)
    ref = np.zeros_like(got)
    for a in range(-6, 7):
        for b in range(-6, 7):
            ref[a % n, b % n, :7] = prod[a % m, b % m, :7]
    assert np.abs(got - ref).max() < 1e-12 * np.abs(ref).max()
    # bilinearity: conv(2A, B) accumulated with -2*conv(A,B) is zero
    P *= 2.0
    C.multiply_field(P, Q, M.op.multiply_add_to(out, -0.5))
    assert np.abs(out.get_kspace()).max() < 1e-12 * np.abs(ref).max()


def test_kernel_generations_agree(M):
    PC.check_kernel_generations(M, 128)


@pytest.mark.parametrize("n,order,digest", [(256, 3, "bench_digest_oracle_256.npz"), (512, 2, "bench_digest_oracle_512_2lpt.npz")])
def test_pipeline_vs_oracle_digest(M, n, order, digest):
    """SURVEY 8d C2 / C3 at sizes the oracle does not finish in seconds: the bench's synthetic workload (per-plane noise
    seeds, amplitude, growth factors) through the CUDA path, its FP64 displacement and velocity fields sampled at 32768
    cells against the CPU oracle's values at the same cells (committed digest, tests/golden/make_bench_digest.py).
    Bar: max|err| <= 1e-9 * max|field| (north_star)."""
    import gc
    import os

    import torch

    import bench as B

    path = os.path.join(B.GOLD, digest)
    if not os.path.exists(path):
        pytest.skip(f"{digest} not committed")
    W = B.Bench(M, M.load(), torch, None, 0, n, order, "sc")
    W.make_noise()
    W.upload_noise()
    W.step_device()
    r = W.check_against_digest(path, "CPU oracle digest")
    del W
    gc.collect()
    assert r["checked"] and r["ok"], r
    assert r["max_rel_psi"] < 1e-9 and r["max_rel_vel"] < 1e-9
