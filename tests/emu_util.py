"""Build / load the CPU thread-emulation build of the kernels (TEST INFRASTRUCTURE ONLY).

tests/emu/libmfb_emu.so is monofonic_b200/csrc/mfb_api.cu compiled with g++ -DMFB_EMU: every CUDA
thread is an OS thread, __syncthreads() a real barrier.  It lets `pytest -m "not gpu"` check the
kernels' index arithmetic against the oracle in the GPU-less build container.  The product never
loads it (monofonic_b200._lib only knows libmfb200.so)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "libmfb_emu.so")
CSRC = os.path.join(ROOT, "monofonic_b200", "csrc")


def build_emu():
    subprocess.run(["make", "-s", "-C", CSRC, "emu"], check=True, capture_output=True)
    return EMU


_loaded = {}


def emu_backend():
    """initialise monofonic_b200 against the emulation library (tests only)"""
    import monofonic_b200 as M
    from monofonic_b200 import grid

    if "lib" not in _loaded:
        build_emu()
        lib = M.load(EMU)
        assert b"emulation" in lib.mfb_build_info()
        rc = lib.mfb_init(0)
        assert rc == 0, lib.mfb_last_error()
        _loaded["lib"] = lib
    grid._state["lib"] = _loaded["lib"]
    return M
