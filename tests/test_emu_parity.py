"""CPU-only: the kernel sources, compiled for the thread emulator (tests/emu), against the oracle
and the reference goldens at 32^3 (padded 48^3).  This validates index arithmetic, pad/unpad,
Nyquist handling, operand factors and the host logic without a GPU; it is NOT a product path."""
import pytest

import emu_util
import parity_cases as PC


@pytest.fixture(scope="module")
def M():
    return emu_util.emu_backend()


@pytest.mark.parametrize("n", [32, 48, 64, 96])
def test_fft_roundtrip(M, n):
    PC.check_fft(M, n)


def test_fft_noncubic(M):
    PC.check_fft(M, 32, 48, 64)


@pytest.mark.parametrize("shape", [(128, 32, 32), (32, 192, 32), (32, 32, 256), (384, 32, 32), (32, 512, 32),
                                   (32, 32, 768), (1024, 32, 32), (32, 1536, 32), (32, 32, 2048), (3072, 32, 32),
                                   (32, 32, 1536), (32, 32, 3072), (32, 2048, 32)])
def test_fft_every_plan(M, shape):
    """each of the 14 radix plans as a strided pass and as a z pass"""
    PC.check_fft(M, *shape)


@pytest.mark.parametrize("shape", [(32, 64, 128), (256, 32, 32), (32, 32, 512), (32, 32, 1024), (32, 32, 2048), (1024, 32, 32),
                                   (32, 1024, 32), (2048, 32, 32), (32, 2048, 32), (512, 32, 32), (32, 512, 32)])
def test_multiply_noncubic_every_conv_plan(M, shape):
    PC.check_multiply_noncubic(M, shape)


def test_c2r_nonhermitian(M):
    PC.check_c2r_nonhermitian(M, 32)


def test_kspace_operators(M):
    PC.check_kops(M, 32)


@pytest.mark.parametrize("nyq0", [True, False])
def test_convolver(M, nyq0):
    PC.check_convolver(M, 32, nyq0)


@pytest.mark.parametrize("order,lattice,real64", [(1, "sc", True), (2, "sc", False), (3, "bcc", True)])
def test_pipeline_vs_oracle(M, order, lattice, real64):
    PC.check_pipeline(M, 32, order, lattice, real64)


@pytest.mark.parametrize("lattice", ["sc", "bcc", "fcc", "rsc"])
def test_lattice_bitexact(M, lattice):
    PC.check_lattice_positions_bitexact(M, 32, lattice)


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


def test_error_behaviour(M):
    import monofonic_b200 as MM
    with pytest.raises(MM.MfbError):
        MM.Grid_FFT((33, 33, 33), (1.0,) * 3)           # no plan for this size
    with pytest.raises(RuntimeError):
        MM.LPTPipeline(MM.lpt_config(GridRes=33, LPTorder=2))  # Orszag needs an even grid (ic_generator.cc:118)


def test_kernel_generations_agree(M):
    PC.check_kernel_generations(M, 32)


def test_stream_handoff(M):
    PC.check_stream_handoff(M, 32)


@pytest.mark.parametrize("nyq0", [True, False])
def test_naive_convolver(M, nyq0):
    PC.check_naive_convolver(M, 32, nyq0)


@pytest.mark.parametrize("shape", [(32, 32, 32), (32, 48, 64), (96, 32, 32)])
def test_complex_grid_fft(M, shape):
    PC.check_cfft(M, *shape)


@pytest.mark.parametrize("order", [1, 2, 3])
def test_ppt_eulerian(M, order):
    PC.check_ppt(M, 32, order)


def test_naive_pipeline(M):
    PC.check_naive_pipeline(M, 32)


def test_product_grouping(M):
    PC.check_product_grouping(M, 32)


@pytest.mark.parametrize("shape", [(32, 64, 64), (32, 32, 128), (32, 32, 256)])
def test_sources_shared_field_pass(M, shape):
    PC.check_sources_shared_field_pass(M, shape)


def test_two_fluid_terms(M):
    PC.check_two_fluid_terms(M, 32, "bcc")


def test_against_reference_produced_fixture(M):
    PC.check_reference_fixture(M)


def test_particles_against_reference_produced_fixture(M):
    PC.check_reference_particles(M)


def test_plt_masked_glass(M):
    PC.check_plt_masked_glass(M, 32, 3)


def test_f4_against_reference_produced_fixture(M):
    PC.check_reference_f4(M)


def test_noise_ingest(M):
    PC.check_noise_ingest(M, 32)


def test_fused_readout(M):
    PC.check_fused_readout(M, 32)
