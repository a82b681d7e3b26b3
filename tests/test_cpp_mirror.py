"""The header-only C++ host mirror (include/mfb/*.hh) driven like the reference drives its classes:
a mock RNG plugin fills a Grid_FFT through relem, the phi(2) source is computed with the call sequence of
src/ic_generator.cc:482-486, particles are read out.  CPU: linked against the emulation build; GPU: against
libmfb200.so."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cc")


def _build_and_run(libdir, libname, tmp_path):
    exe = str(tmp_path / "test_host_mirror")
    subprocess.run(["g++", "-O1", "-std=c++17", SRC, "-o", exe, "-L", libdir, "-l" + libname, "-Wl,-rpath," + libdir],
                   check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_cpp_mirror_emulated(tmp_path):
    import emu_util

    emu_util.build_emu()
    out = _build_and_run(os.path.join(ROOT, "tests", "emu"), "mfb_emu", tmp_path)
    assert "host mirror ok" in out


@pytest.mark.gpu
def test_cpp_mirror_gpu(tmp_path):
    out = _build_and_run(os.path.join(ROOT, "monofonic_b200"), "mfb200", tmp_path)
    assert "host mirror ok (cuda sm_100a)" in out
