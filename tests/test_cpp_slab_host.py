"""A C++ host with one process per rank on slab-decomposed grids (Grid_FFT<real_t, true>, what the reference's USE_MPI build
instantiates), CPU only: tests/cpp/slab_host.cc is built twice against the drop-in headers and the thread-emulation build of
the kernels - once for the whole grid, once with GRID_FFT_DISTRIBUTED - and 2 and 4 separately started rank processes must
reproduce the whole-grid result on their planes (phi(2) source with the reference's call sequence, Poisson solve, a gradient by
host lambda in the transposed k-layout, global reductions).  The slab exchanges travel through the library's staged mode with
a host callback over a shared-memory file - the place of MPI_Alltoall in an MPI host."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "slab_host.cc")
EMU = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def binaries(tmp_path_factory):
    import emu_util

    emu_util.build_emu()
    d = tmp_path_factory.mktemp("slab_host")
    out = {}
    for name, flags in (("ref", []), ("rank", ["-DGRID_FFT_DISTRIBUTED=true"])):
        exe = str(d / ("slab_host_" + name))
        subprocess.run(["g++", "-O1", "-std=c++17"] + flags + [SRC, "-o", exe, "-L", EMU, "-lmfb_emu", "-Wl,-rpath," + EMU],
                       check=True, capture_output=True)
        out[name] = exe
    ref = str(d / "ref.bin")
    r = subprocess.run([out["ref"], "ref", ref], capture_output=True, text=True)
    assert r.returncode == 0 and "reference ok" in r.stdout, r.stdout + r.stderr
    out["ref_file"], out["dir"] = ref, str(d)
    return out


@pytest.mark.parametrize("world", [2, 4])
def test_cpp_host_on_slabs(binaries, world):
    shm = os.path.join(binaries["dir"], f"shm{world}")
    with open(shm, "wb") as f:
        f.truncate(64 + world * (64 << 20))     # sparse; every rank maps header + world exchange buffers
    procs = [subprocess.Popen([binaries["rank"], "rank", str(r), str(world), shm, binaries["ref_file"]], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=600)[0])
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"slab host rank {r} ok" in o, o[-2000:]
