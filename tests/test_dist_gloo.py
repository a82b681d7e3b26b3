"""world_size-2 and -4 runs of the slab-decomposed path on CPU: emulation build of the kernels + gloo.
(The GPU version of the same worker, with the exchange fused into the FFT passes over peer memory, is
tests/test_gpu_dist.py.)"""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run_world(world, extra, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(free_port()), os.path.join(ROOT, "tests", "dist_worker.py")] + extra
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "dist_worker ok" in r.stdout, r.stdout[-2000:]
    return r.stdout


@pytest.mark.parametrize("world,lattice", [(2, "bcc"), (4, "sc")])
def test_slab_pipeline_gloo(world, lattice):
    import emu_util

    emu_util.build_emu()
    run_world(world, ["--emu", "--grid", "32", "--lattice", lattice])


def test_slab_pipeline_gloo_two_exchange_buffers():
    """a peer-mapped block with only the ping-pong pair: the source stage takes the per-product path"""
    import emu_util

    emu_util.build_emu()
    run_world(2, ["--emu", "--grid", "32", "--lattice", "sc", "--nslots", "2"])


def test_slab_pipeline_gloo_eight_slots():
    """8 slots: the first third-order stage is a shared-field loop, the last one two multi-product loops (the default block
    of these tests has 10: both shared)"""
    import emu_util

    emu_util.build_emu()
    run_world(2, ["--emu", "--grid", "32", "--lattice", "sc", "--nslots", "8"])


def test_slab_pipeline_gloo_tight_slots():
    """7 slots (the minimum of the shared-transform source stage at 3LPT): one target per x-plane loop"""
    import emu_util

    emu_util.build_emu()
    run_world(2, ["--emu", "--grid", "32", "--lattice", "sc", "--nslots", "7"])


def test_slab_pipeline_gloo_split_targets():
    """5 slots (what 3LPT 2048^3 leaves on 8 GPUs is 6): targets whose products need more X buffers than that go through
    several loops, the later ones adding to the target"""
    import emu_util

    emu_util.build_emu()
    run_world(2, ["--emu", "--grid", "32", "--lattice", "sc", "--nslots", "5"])


def test_slab_pipeline_gloo_borrowed_accumulator():
    """the memory-pressure configuration of 3LPT 2048^3 on 8 GPUs, on two CPU ranks: product groups accumulate in the
    (idle) particle arrays of each slab"""
    import emu_util

    emu_util.build_emu()
    run_world(2, ["--emu", "--grid", "32", "--lattice", "sc", "--borrow"])
