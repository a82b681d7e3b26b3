"""Multi-GPU parity (needs >= 2 B200s on the box; skipped otherwise): the slab-decomposed pipeline with the
exchange fused into the FFT passes over peer memory, against the single-process CPU oracle - the analogue of
the reference's tests/scripts/test_mpi_consistency.sh (1/2/4 ranks must agree)."""
import pytest

from test_dist_gloo import run_world

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_pipeline_peer_memory(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    run_world(world, ["--grid", "64", "--lattice", "bcc"])


def test_slab_pipeline_two_exchange_buffers():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    run_world(2, ["--grid", "64", "--lattice", "sc", "--nslots", "2"])


def test_slab_pipeline_eight_slots():
    """8 slots: first third-order stage as a shared-field loop, the last one as two multi-product loops (the default block
    has 10 slots: both shared)"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    run_world(2, ["--grid", "64", "--lattice", "sc", "--nslots", "8"])


def test_slab_pipeline_tight_slots():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    run_world(2, ["--grid", "64", "--lattice", "bcc", "--nslots", "7"])


def test_slab_pipeline_split_targets():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    run_world(2, ["--grid", "64", "--lattice", "sc", "--nslots", "6"])


def test_slab_pipeline_nccl_staged():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    run_world(2, ["--grid", "64", "--lattice", "sc", "--staged"])


def test_slab_pipeline_borrowed_accumulator():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    run_world(2, ["--grid", "64", "--lattice", "sc", "--borrow"])
