"""A C++ host driver written against the drop-in headers (tests/cpp/ic_driver.cc: RNG plugin -> transfer table -> LPT
sources -> species loop -> output plugin, the shape of ic_generator::run) against the CPU oracle on the same noise.
CPU: linked against the emulation build; GPU: against libmfb200.so."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle")]
SRC = os.path.join(ROOT, "tests", "cpp", "ic_driver.cc")


def _run_driver(libdir, libname, tmp_path, n, order, lattice):
    import lpt_oracle as O

    exe = str(tmp_path / "ic_driver")
    subprocess.run(["g++", "-O1", "-std=c++17", SRC, "-o", exe, "-L", libdir, "-l" + libname, "-Wl,-rpath," + libdir],
                   check=True, capture_output=True)
    rng = np.random.default_rng(2024)
    noise = rng.standard_normal((n, n, n))
    noise -= noise.mean()
    noise_path, out_path = str(tmp_path / "noise.bin"), str(tmp_path / "out.bin")
    noise.tofile(noise_path)
    lt = O.LATTICE_TYPES[lattice]
    r = subprocess.run([exe, str(n), str(order), str(lt), noise_path, out_path], capture_output=True, text=True)
    assert r.returncode == 0 and "ic_driver ok" in r.stdout, r.stdout + r.stderr
    raw = np.fromfile(out_path, dtype=np.uint8)
    npart = int(raw[:8].view(np.uint64)[0])
    body = raw[8:]
    pos = body[:24 * npart].view(np.float64).reshape(npart, 3)
    vel = body[24 * npart:48 * npart].view(np.float64).reshape(npart, 3)
    ids = body[48 * npart:56 * npart].view(np.uint64)
    # the oracle on the same noise: the RNG plugin fills real space, ic_generator transforms it (:313-346)
    L, Dp, vfac = 100.0, 0.0257, 396.69
    g = O.GridSpec(n, L)
    Wc = O.condition_noise(g, g.fwd(noise))
    ref = O.run_pipeline(n, L, Wc, O.power_law_amplitude(), order, Dp, vfac, lattice, L, L / np.sqrt(1 / 51.0), True, True)
    assert npart == ref["ids"].size and np.array_equal(ids, ref["ids"])
    maxpsi = max(np.abs(ref["fields"]["d" + c]).max() for c in "xyz")
    maxv = max(np.abs(ref["fields"]["v" + c]).max() for c in "xyz")
    assert np.abs(pos - ref["pos"]).max() < 1e-11 * maxpsi      # north_star bar: 1e-9 max|Psi|
    assert np.abs(vel - ref["vel"]).max() < 1e-11 * maxv
    return r.stdout


@pytest.mark.parametrize("order,lattice", [(3, "bcc"), (1, "sc")])
def test_cpp_ic_driver_emulated(tmp_path, order, lattice):
    import emu_util

    emu_util.build_emu()
    _run_driver(os.path.join(ROOT, "tests", "emu"), "mfb_emu", tmp_path, 32, order, lattice)


@pytest.mark.gpu
def test_cpp_ic_driver_gpu(tmp_path):
    out = _run_driver(os.path.join(ROOT, "monofonic_b200"), "mfb200", tmp_path, 64, 3, "fcc")
    assert "cuda sm_100a" in out
