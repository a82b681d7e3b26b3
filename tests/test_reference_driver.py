"""The reference's OWN executable on the library: main.cc, ic_generator.cc (the numerical driver, host lambdas and all), the
cosmology calculator, the plugin factories and the Eisenstein&Hu / NGenIC plugins, every source compiled UNMODIFIED where it
lies under /root/reference against the drop-in headers (tests/cpp/Makefile), run on the reference's own regression
configuration (tests/configs/test_*lpt_*.conf: 32^3, L = 100, z = 50, Planck2018EE+BAO+SN, seed 12345) with the output
plugin swapped for tests/cpp/output_raw_plugin.cc (the reference's plugin API; raw files instead of HDF5, which is not
installed).  Only src/grid_fft.cc is absent: FFTs, dealiased products, Poisson solves and the particle read-out run in
libmfb200 (CPU: the thread-emulation build of the same kernels; GPU: the prebuilt executable travels to the box).

Check, as for the reference's goldens (tests/golden_util.py): phi(1) is recovered from the density field of the
1LPT run, the CPU oracle is run on it, and every displacement / velocity / density field and the particle container of the
1LPT / 2LPT / 3LPT runs must agree with the oracle to 1e-12 of the field maximum (north_star: 1e-9); ids exact.
The generator behind the NGenIC plugin is Luescher's RANLUX (tests/cpp/gslmini.cc): the noise is the reference's own, and the
five golden files themselves are reproduced (last section of this file).

The reference's USE_MPI build of the same sources runs too: one process per rank (tests/cpp/minimpi stands in for the MPI
library), every rank a slab of the library - Grid_FFT<real_t, true> in the transposed layout, MPI_Allreduce in the driver and the
particle container, the ghost-plane ring shift of grid_interpolate.hh - with the slab transposes inside the library (CPU: staged
through MPI_Alltoall; GPUs: peer memory after one MPI_Allgather of the IPC handles).  This is the reference's own MPI
consistency test (tests/scripts/test_mpi_consistency.sh:45-138: 1 vs 2 vs 4 tasks of the same configuration, rtol 1e-9) with
a stricter bar: 2 and 4 ranks must reproduce the
single-process run of the same configuration bit for bit (fields) and particle for particle (matched by id).

And the reference ITSELF is run beside it: tests/cpp/_dropin/monofonIC_ref_cpu is the same set of sources plus the reference's own
src/grid_fft.cc against its OWN headers - its Grid_FFT, its OrszagConvolver with the pad / unpad of its MPI build, its
lattice_generator - with no line of this repository's library in it.  Identical configuration files and seeds go through both
executables and every output is compared directly: north_star's "results must match the reference FFTW path on identical
configs and noise seeds" (FFTW itself is not installed: its calls are served by the FFT of oracle/refshim/fftw_fast.cc)."""
import os
import subprocess

import numpy as np
import pytest

import lpt_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
N, L = 32, 100.0
FIELDS = ["delta", "dx", "dy", "dz", "vx", "vy", "vz"]
TOL = 1e-12

CONF = """[setup]
GridRes         = {grid}
BoxLength       = 100
zstart          = 50.0
LPTorder        = {order}
DoBaryons       = {baryons}
DoBaryonVrel    = {vrel}
DoFixing        = no
DoInversion     = no
DoRemoveCornerModes = no
ParticleLoad    = {lattice}
[cosmology]
ParameterSet    = Planck2018EE+BAO+SN
ZeroRadiation   = false
transfer        = {transfer}
[random]
{random}
[execution]
NumThreads      = {threads}
[output]
format          = mfb_raw
filename        = {name}
raw_particles   = {particles}
raw_eulerian    = {eulerian}
"""


RANDOM = {"NGENIC": "generator       = NGENIC\nseed            = 12345",
          "NGENIC_GPU": "generator       = NGENIC_GPU\nseed            = 12345",   # include/mfb/plugins/random_ngenic_device.cc
          "MUSIC1": "generator       = MUSIC1\nseed[{level}]         = 12345"}   # MUSIC1: one seed per level, levelmin = log2(GridRes)


def _run(exe, tmp, name, order, lattice="sc", particles=False, baryons=False, generator="NGENIC", eulerian=False, grid=32, threads=1,
         vrel=False, transfer="eisenstein", edit=None):
    conf = os.path.join(tmp, name + ".conf")
    with open(conf, "w") as f:
        f.write((edit or (lambda t: t))(CONF.format(order=order, lattice=lattice, name=name, particles="yes" if particles else "no",
                            baryons="yes" if baryons else "no", random=RANDOM[generator].format(level=grid.bit_length() - 1),
                            eulerian="yes" if eulerian else "no", grid=grid, threads=threads, vrel="yes" if vrel else "no",
                            transfer=transfer)))
    r = subprocess.run([exe, conf], cwd=tmp, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "Have a nice day" in r.stdout
    return r.stdout


def _fields(tmp, name, species="DM", n=N):
    return {c: np.fromfile(os.path.join(tmp, f"{name}.{species}_{c}.f64")).reshape(n, n, n) for c in FIELDS}


def _check(exe, tmp):
    tmp = str(tmp)
    g = O.GridSpec(N, L)
    # ---- 1LPT: recover phi_s = g1 * phi(1) and vfac from the run's own output (golden_util.bootstrap)
    out = _run(exe, tmp, "o1", 1)
    assert "eisenstein" in out and "NGENIC" in out.upper()
    # the self-registering factories of the reference know every plugin that was linked in (printed when run without arguments)
    listing = subprocess.run([exe], capture_output=True, text=True).stdout
    for plugin in ("'CAMB_file'", "'eisenstein'", "'MUSIC1'", "'NGENIC'", "'gadget_hdf5'", "'SWIFT'", "'AREPO'", "'grafic2'", "'generic'", "'mfb_raw'"):
        assert plugin in listing, plugin
    d1 = _fields(tmp, "o1")
    with np.errstate(divide="ignore", invalid="ignore"):
        phi_s = g.fwd(d1["delta"]) / g.k2()
    phi_s[0, 0, 0] = 0.0
    m = np.abs(d1["dx"]) > 1e-4 * np.abs(d1["dx"]).max()
    vfac = float(np.median(d1["vx"][m] / d1["dx"][m]))
    assert abs(vfac / 396.6919511923062 - 1.0) < 1e-9   # the goldens' value: the cosmology calculator ran on the stand-in quadrature / splines
    assert 0.03 < d1["delta"].std() < 0.12             # sigma of the linear density at z = 50 on this grid (golden: 0.064)
    prev = None
    for order in (1, 2, 3):
        name = f"o{order}"
        if order > 1:
            _run(exe, tmp, name, order)
        d = _fields(tmp, name)
        pots = O.lpt_potentials(g, phi_s, order, -1.0)
        f = O.fields(g, pots, order, vfac)
        for c in FIELDS:
            err = np.abs(f[c] - d[c]).max() / np.abs(d[c]).max()
            assert err < TOL, (order, c, err)
        if prev is not None:   # each order really adds its term
            assert np.abs(d["dx"] - prev["dx"]).max() > 1e-5 * np.abs(d["dx"]).max()
        prev = d
    # ---- particles: 3LPT on a bcc lattice through particle::lattice_generator and the output plugin's container access
    _run(exe, tmp, "p3", 3, lattice="bcc", particles=True)
    pots = O.lpt_potentials(g, phi_s, 3, -1.0)
    lt = O.LATTICE_TYPES["bcc"]
    pos = np.fromfile(os.path.join(tmp, "p3.DM.pos.f64")).reshape(-1, 3)
    vel = np.fromfile(os.path.join(tmp, "p3.DM.vel.f64")).reshape(-1, 3)
    ids = np.fromfile(os.path.join(tmp, "p3.DM.ids.u64"), dtype=np.uint64)
    assert np.array_equal(ids, O.particle_ids(N, lt, 0, True))
    for dd in range(3):
        dr = g.bwd(O.displacement_k(g, pots, 3, dd, 1.0 / L))
        vr = g.bwd(O.velocity_k(g, pots, 3, dd, vfac, 1.0 / L))
        assert np.abs(O.set_positions(g, dr, dd, lt, 0, 1.0, True) - pos[:, dd]).max() < 1e-14
        v = O.set_velocities(g, vr, lt, 0, True)
        assert np.abs(v - vel[:, dd]).max() < TOL * np.abs(v).max()

    # ---- two species (DoBaryons): dark matter on the primary sc lattice, baryons on the shifted secondary one with id offset
    # 1 and per-particle masses (src/ic_generator.cc:636-680); E&H has no distinct baryon transfer function, so rho_bc = 0
    _run(exe, tmp, "b2", 2, lattice="sc", particles=True, baryons=True)
    pots = O.lpt_potentials(g, phi_s, 2, -1.0)
    omega_b, omega_m = 0.0488911, 0.3099      # printed by the run ("Cosmological parameters are"); masses checked to 1e-5
    for species, index, omega in (("DM", 0, omega_m - omega_b), ("BA", 1, omega_b)):
        pos = np.fromfile(os.path.join(tmp, f"b2.{species}.pos.f64")).reshape(-1, 3)
        vel = np.fromfile(os.path.join(tmp, f"b2.{species}.vel.f64")).reshape(-1, 3)
        ids = np.fromfile(os.path.join(tmp, f"b2.{species}.ids.u64"), dtype=np.uint64)
        mass = np.fromfile(os.path.join(tmp, f"b2.{species}.mass.f64"))
        assert np.array_equal(ids, O.particle_ids(N, 0, index, True)), species
        for dd in range(3):
            dr = g.bwd(O.displacement_k(g, pots, 2, dd, 1.0 / L))
            vr = g.bwd(O.velocity_k(g, pots, 2, dd, vfac, 1.0 / L))
            assert np.abs(O.set_positions(g, dr, dd, 0, index, 1.0, True) - pos[:, dd]).max() < 1e-14, species
            v = O.set_velocities(g, vr, 0, index, True)
            assert np.abs(v - vel[:, dd]).max() < TOL * np.abs(v).max(), species
        assert mass.shape == (N ** 3,) and np.allclose(mass, omega / N ** 3, rtol=1e-5), (species, mass[:3], omega / N ** 3)

    # ---- Eulerian output: the semiclassical (PPT) branch of the driver (src/ic_generator.cc:685-803) as written - host lambdas on
    # Grid_FFT<ccomplex_t> (the specialisation of include/mfb/grid_fft_complex.hh), c2c transforms on the device.  E&H has no
    # distinct baryon transfer function, so rho_bc = 0; D+ cancels out of every output (hbar ~ 1/D+).
    for order in (1, 2):
        out = _run(exe, tmp, f"e{order}", order, eulerian=True)
        assert "Semiclassical PT : hbar" in out
        e = {c: np.fromfile(os.path.join(tmp, f"e{order}.DM_{c}.f64")).reshape(N, N, N) for c in ("delta", "vx", "vy", "vz")}
        pots = O.lpt_potentials(g, phi_s, order, -1.0)
        ref, hbar = O.ppt_eulerian(g, np.zeros_like(phi_s), pots, order, 0.0251908, vfac, lambda k: 0.0 * k, 1.0)
        for c, r in (("delta", ref["density"]), ("vx", ref["vx"]), ("vy", ref["vy"]), ("vz", ref["vz"])):
            err = np.abs(e[c] - r).max() / np.abs(r).max()
            assert err < 1e-9, ("PPT", order, c, err)

    # ---- the MUSIC1 noise plugin (random_music.cc + its white-noise generator, which runs FFTs of its own): real-space noise
    # with POPULATED Nyquist planes - the case where the pad / unpad semantics of the reference's MPI build matter
    for order in (1, 2):
        _run(exe, tmp, f"m{order}", order, generator="MUSIC1")
    m1, m2 = _fields(tmp, "m1"), _fields(tmp, "m2")
    with np.errstate(divide="ignore", invalid="ignore"):
        phi_m = g.fwd(m1["delta"]) / g.k2()
    phi_m[0, 0, 0] = 0.0
    assert np.abs(phi_m[N // 2]).max() > 1e-6 * np.abs(phi_m).max()      # Nyquist planes carry power
    f = O.fields(g, O.lpt_potentials(g, phi_m, 2, -1.0), 2, vfac)
    for c in FIELDS:
        err = np.abs(f[c] - m2[c]).max() / np.abs(m2[c]).max()
        assert err < TOL, ("MUSIC1", c, err)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference sources not mounted")
def test_reference_executable_on_the_emulated_library(tmp_path):
    import emu_util

    emu_util.build_emu()
    r = subprocess.run(["make", "-s", "-C", CPP, "-j4", "emu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    _check(os.path.join(CPP, "_dropin", "monofonIC_mfb_emu"), tmp_path)


@pytest.mark.gpu
def test_reference_executable_on_the_gpu(tmp_path):
    exe = os.path.join(CPP, "_dropin", "monofonIC_mfb")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/_dropin/monofonIC_mfb was not built (needs /root/reference at build time; __graft_entry__.build())")
    _check(exe, tmp_path)


def _run_mpi(exe, tmp, name, world, order, lattice="sc", particles=False, extra_env=None, generator="NGENIC", eulerian=False):
    conf = os.path.join(tmp, name + ".conf")
    with open(conf, "w") as f:
        f.write(CONF.format(order=order, lattice=lattice, name=name, particles="yes" if particles else "no", baryons="no",
                            random=RANDOM[generator], eulerian="yes" if eulerian else "no", grid=32, threads=1, vrel="no",
                            transfer="eisenstein"))
    shm = os.path.join(tmp, name + ".shm")
    with open(shm, "wb") as f:
        f.truncate(64 + world * (64 << 20))     # sparse: header + one 64 MiB slot per rank
    procs = []
    for r in range(world):
        env = dict(os.environ, MINIMPI_RANK=str(r), MINIMPI_SIZE=str(world), MINIMPI_SHM=shm, **(extra_env or {}))
        procs.append(subprocess.Popen([exe, conf], cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=900)[0])
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, (r, o[-3000:])
    assert "Have a nice day" in outs[0] and "MPI is enabled                   : yes" in outs[0]


def _check_mpi(exe_mpi, exe_single, tmp, worlds):
    tmp = str(tmp)
    _run(exe_single, tmp, "s1", 1)
    _run(exe_single, tmp, "s3", 3)
    _run(exe_single, tmp, "sp", 3, lattice="bcc", particles=True)
    single = _fields(tmp, "s3")
    sp = {k: np.fromfile(os.path.join(tmp, "sp.DM." + k), dtype=np.uint64 if k.endswith("u64") else np.float64) for k in ("pos.f64", "vel.f64", "ids.u64")}
    order = np.argsort(sp["ids.u64"])
    for world in worlds:
        name = f"m3w{world}"
        _run_mpi(exe_mpi, tmp, name, world, 3)
        for c in FIELDS:   # slabs along axis 0, in rank order
            got = np.concatenate([np.fromfile(os.path.join(tmp, f"{name}.r{r}.DM_{c}.f64")).reshape(-1, N, N) for r in range(world)])
            assert got.shape == (N, N, N)
            assert np.abs(got - single[c]).max() <= 1e-13 * np.abs(single[c]).max(), (world, c)
        if world == 2:   # the ping-pong pair only: the source stage takes its per-product path on the slabs
            _run_mpi(exe_mpi, tmp, "m3w2s2", world, 3, extra_env={"MFB_EXCHANGE_SLOTS": "2"})
            for c in FIELDS:
                got = np.concatenate([np.fromfile(os.path.join(tmp, f"m3w2s2.r{r}.DM_{c}.f64")).reshape(-1, N, N) for r in range(world)])
                assert np.abs(got - single[c]).max() <= 1e-13 * np.abs(single[c]).max(), (world, c, "2 slots")
        if world == 2:
            # the semiclassical (PPT) branch on slabs: Grid_FFT<ccomplex_t, true> with the reference's host lambdas; hbar follows the
            # reference's MPI loop (every rank leaves out the x neighbour of its last local plane, then MPI_Allreduce MAX), so it is
            # compared with the oracle's nranks = 2 statement of that loop, not with the single-process run
            _run_mpi(exe_mpi, tmp, "e2w2", world, 2, eulerian=True)
            g = O.GridSpec(N, L)
            d1 = _fields(tmp, "s1")
            with np.errstate(divide="ignore", invalid="ignore"):
                phi_s = g.fwd(d1["delta"]) / g.k2()
            phi_s[0, 0, 0] = 0.0
            mk = np.abs(d1["dx"]) > 1e-4 * np.abs(d1["dx"]).max()
            vfac = float(np.median(d1["vx"][mk] / d1["dx"][mk]))
            ref, _ = O.ppt_eulerian(g, np.zeros_like(phi_s), O.lpt_potentials(g, phi_s, 2, -1.0), 2, 0.0251908, vfac, lambda k: 0.0 * k, 1.0, 1.0, 2)
            for c, rname in (("delta", "density"), ("vx", "vx"), ("vy", "vy"), ("vz", "vz")):
                got = np.concatenate([np.fromfile(os.path.join(tmp, f"e2w2.r{r}.DM_{c}.f64")).reshape(-1, N, N) for r in range(world)])
                assert np.abs(got - ref[rname]).max() < 1e-9 * np.abs(ref[rname]).max(), ("PPT on slabs", c)
        name = f"mpw{world}"
        _run_mpi(exe_mpi, tmp, name, world, 3, lattice="bcc", particles=True)
        ids = np.concatenate([np.fromfile(os.path.join(tmp, f"{name}.r{r}.DM.ids.u64"), dtype=np.uint64) for r in range(world)])
        assert len(ids) == 2 * N ** 3 and len(np.unique(ids)) == len(ids)
        o2 = np.argsort(ids)
        assert np.array_equal(ids[o2], sp["ids.u64"][order])
        for k in ("pos.f64", "vel.f64"):
            got = np.concatenate([np.fromfile(os.path.join(tmp, f"{name}.r{r}.DM.{k}")).reshape(-1, 3) for r in range(world)])
            ref = sp[k].reshape(-1, 3)
            assert np.abs(got[o2] - ref[order]).max() <= 1e-13 * np.abs(ref).max(), (world, k)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference sources not mounted")
def test_reference_mpi_executable_on_slabs_of_the_emulated_library(tmp_path):
    import emu_util

    emu_util.build_emu()
    r = subprocess.run(["make", "-s", "-C", CPP, "-j4", "emu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    _check_mpi(os.path.join(CPP, "_dropin", "monofonIC_mfb_mpi_emu"), os.path.join(CPP, "_dropin", "monofonIC_mfb_emu"), tmp_path, (2, 4))


@pytest.mark.gpu
def test_reference_mpi_executable_on_two_gpus(tmp_path):
    import torch

    exe = os.path.join(CPP, "_dropin", "monofonIC_mfb_mpi")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/_dropin/monofonIC_mfb_mpi was not built (needs /root/reference at build time)")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (one rank per GPU, exchange over peer memory)")
    _check_mpi(exe, os.path.join(CPP, "_dropin", "monofonIC_mfb"), tmp_path, (2,))


def _check_against_the_reference_itself(ref_exe, our_exe, tmp, grid, threads):
    tmp = str(tmp)
    worst = 0.0
    for order in (1, 2, 3):
        _run(ref_exe, tmp, f"r{order}", order, grid=grid, threads=threads)
        _run(our_exe, tmp, f"d{order}", order, grid=grid, threads=threads)
        a, b = _fields(tmp, f"r{order}", n=grid), _fields(tmp, f"d{order}", n=grid)
        for c in FIELDS:
            err = np.abs(a[c] - b[c]).max() / np.abs(a[c]).max()
            worst = max(worst, err)
            assert err < TOL, (order, c, err)
    # particles of a 3LPT bcc load: ids exact, positions to an ulp of the box, velocities to 1e-12 of their maximum
    _run(ref_exe, tmp, "rp", 3, lattice="bcc", particles=True, grid=grid, threads=threads)
    _run(our_exe, tmp, "dp", 3, lattice="bcc", particles=True, grid=grid, threads=threads)
    ids = [np.fromfile(os.path.join(tmp, f"{t}.DM.ids.u64"), dtype=np.uint64) for t in ("rp", "dp")]
    assert len(ids[0]) == 2 * grid ** 3 and np.array_equal(ids[0], ids[1])
    pos = [np.fromfile(os.path.join(tmp, f"{t}.DM.pos.f64")) for t in ("rp", "dp")]
    vel = [np.fromfile(os.path.join(tmp, f"{t}.DM.vel.f64")) for t in ("rp", "dp")]
    assert np.abs(pos[0] - pos[1]).max() < 4e-16
    assert np.abs(vel[0] - vel[1]).max() < TOL * np.abs(vel[0]).max()
    # MUSIC1 noise (populated Nyquist planes) at 2LPT
    _run(ref_exe, tmp, "rm", 2, generator="MUSIC1", grid=grid, threads=threads)
    _run(our_exe, tmp, "dm", 2, generator="MUSIC1", grid=grid, threads=threads)
    a, b = _fields(tmp, "rm", n=grid), _fields(tmp, "dm", n=grid)
    for c in FIELDS:
        err = np.abs(a[c] - b[c]).max() / np.abs(a[c]).max()
        worst = max(worst, err)
        assert err < TOL, ("MUSIC1", c, err)
    print(f"reference itself vs drop-in at {grid}^3: worst field error {worst:.2e} of the field maximum")


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference sources not mounted")
def test_the_reference_itself_and_the_dropin_agree_on_identical_configs(tmp_path):
    import emu_util

    emu_util.build_emu()
    r = subprocess.run(["make", "-s", "-C", CPP, "-j4", "emu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    _check_against_the_reference_itself(os.path.join(CPP, "_dropin", "monofonIC_ref_cpu"), os.path.join(CPP, "_dropin", "monofonIC_mfb_emu"),
                                        tmp_path, 32, 1)


@pytest.mark.gpu
def test_the_reference_itself_on_the_host_and_the_dropin_on_the_gpu_agree_at_128(tmp_path):
    ref, ours = os.path.join(CPP, "_dropin", "monofonIC_ref_cpu"), os.path.join(CPP, "_dropin", "monofonIC_mfb")
    if not (os.path.exists(ref) and os.path.exists(ours)):
        pytest.skip("tests/cpp/_dropin executables were not built (need /root/reference at build time)")
    _check_against_the_reference_itself(ref, ours, tmp_path, 128, min(16, os.cpu_count() or 1))


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference sources not mounted")
def test_dropin_with_the_reference_openmp_loops_on(tmp_path):
    """-fopenmp build of the same unmodified sources, [execution] NumThreads = 4: the driver's `#pragma omp parallel for` loops
    call kelem / relem of the drop-in grids from inside OpenMP teams (thread-safe lazy host synchronisation); results equal the
    serial build bit for bit."""
    import emu_util

    emu_util.build_emu()
    r = subprocess.run(["make", "-s", "-C", CPP, "-j4", "emu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    tmp = str(tmp_path)
    ser, omp = os.path.join(CPP, "_dropin", "monofonIC_mfb_emu"), os.path.join(CPP, "_dropin", "monofonIC_mfb_omp_emu")
    _run(ser, tmp, "s3", 3)
    out = _run(omp, tmp, "o3", 3, threads=4)
    assert "(4 used)" in out
    a, b = _fields(tmp, "s3"), _fields(tmp, "o3")
    for c in FIELDS:
        assert np.array_equal(a[c], b[c]), c
    _run(ser, tmp, "sp", 3, lattice="bcc", particles=True)
    _run(omp, tmp, "op", 3, lattice="bcc", particles=True, threads=4)
    for k, dt in (("pos.f64", np.float64), ("vel.f64", np.float64), ("ids.u64", np.uint64)):
        assert np.array_equal(np.fromfile(os.path.join(tmp, "sp.DM." + k), dtype=dt), np.fromfile(os.path.join(tmp, "op.DM." + k), dtype=dt)), k


# ---------------------------------------------------------------------------------------------------------------------------
# The reference's five regression configurations against its five committed golden files (tests/references/*.hdf5 of the
# reference, converted by tests/golden/make_golden.py).  With Luescher's RANLUX behind "ranlxd1" (tests/cpp/gslmini.cc) the NGenIC
# plugin draws the reference's own white noise, so the reference's unmodified executable on the library must reproduce the golden
# files at the reference's own tolerance (tests/scripts/compare_hdf5.py: rtol 1e-9, atol 1e-14) - the one freedom left is the
# sigma_8 normalisation, which here comes from the stand-in quadrature instead of GSL's QAG (measured: a common factor 1 + 3.5e-12;
# see _close_to_golden).
GOLDEN_RUNS = {   # reference config -> (golden file, output plugin it names, what to emulate of that plugin: units and widths)
    "test_1lpt_sc_generic": dict(fmt="generic", particles=False),
    "test_2lpt_baryons_generic": dict(fmt="generic", particles=False),
    "test_2lpt_sc_gadget": dict(fmt="gadget", particles=True),
    "test_2lpt_baryons_vrel_gadget": dict(fmt="gadget", particles=True),
    "test_3lpt_bcc_swift": dict(fmt="swift", particles=True),
}


def _golden_conf(name, tmp):
    """the reference's own configuration file, verbatim, with its [output] section pointed at the raw plugin (same units and
    widths as the plugin it names); where the reference is not mounted (GPU box) the same keys from this file's template"""
    import golden_util as G

    spec = GOLDEN_RUNS[name]
    # the raw plugin asks the plugin the configuration names - an instance made by the reference's own factory - for its units,
    # widths and species mode (output_gadget_hdf5.cc:83-84, output_swift.cc:73-74, output_generic.cc:70-78); UseLongids as in the
    # reference's test_3lpt_bcc_swift.conf
    plugin = {"generic": "generic", "gadget": "gadget_hdf5", "swift": "SWIFT"}[spec["fmt"]]
    out = (f"[output]\nformat = mfb_raw\nfilename = {name}\nraw_units_of = {plugin}\n"
           + ("UseLongids = true\n" if spec["fmt"] == "swift" else ""))
    src = os.path.join("/root/reference/tests/configs", name + ".conf")
    if os.path.exists(src):
        text = open(src).read()
        text = text[:text.index("[output]")] + out
    else:
        order = int(name.split("_")[1][0])
        text = CONF.format(order=order, lattice="bcc" if "bcc" in name else "sc", name=name, particles="no", eulerian="no",
                           baryons="yes" if "baryons" in name else "no", random=RANDOM["NGENIC"], grid=32, threads=1,
                           vrel="yes" if "vrel" in name else "no", transfer="eisenstein")
        text = text[:text.index("[output]")] + out
    path = os.path.join(tmp, name + ".conf")
    with open(path, "w") as f:
        f.write(text)
    return path


def _close_to_golden(a, b):
    """the reference's comparator (tests/scripts/compare_hdf5.py:66-75): element-wise |a - b| <= 1e-14 + 1e-9 |b|.  The one
    difference between this run and the run that made the golden file is the sigma_8 normalisation integral (GSL's 61-point QAG
    to an estimated 1e-10 there, a converged Gauss-Kronrod(7,15) bisection here): a common factor 1 + 3.5e-12 on first-order
    terms, its square on second-order ones.  An element in which the orders nearly cancel can therefore miss the RELATIVE bound
    while its error stays at 3.5e-12 of the field maximum; such elements are accepted up to 1e-11 of the maximum, at most 1e-4 of
    all elements, and the whole field must be within 1e-11 of its maximum (north_star: 1e-9)."""
    if np.asarray(b).dtype == np.float32:
        # float32 particle data (Gadget): a 3.5e-12 offset can only show as a flipped last bit - at least 99.9 % of the values
        # must be bit-identical, none further off than one float32 ulp of the field maximum
        a32, b32 = np.asarray(a, dtype=np.float32), np.asarray(b, dtype=np.float32)
        return bool((a32 == b32).mean() > 0.999 and np.abs(a32.astype(np.float64) - b32).max() <= 2.0 ** -23 * np.abs(b32).max())
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    err, scale = np.abs(a - b), np.abs(b).max()
    if err.max() > 1e-11 * scale:
        return False
    bad = err > 1e-14 + 1e-9 * np.abs(b)
    return bool(bad.sum() <= max(1, int(1e-4 * bad.size)))


def _check_goldens(exe, tmp):
    import golden_util as G

    close = _close_to_golden
    tmp = str(tmp)
    worst = 0.0
    for name, spec in GOLDEN_RUNS.items():
        conf = _golden_conf(name, tmp)
        r = subprocess.run([exe, conf], cwd=tmp, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "Have a nice day" in r.stdout, (name, (r.stdout + r.stderr)[-2000:])
        gold = G.load(name)
        if not spec["particles"]:
            for c in FIELDS:
                got = np.fromfile(os.path.join(tmp, f"{name}.DM_{c}.f64")).reshape(N, N, N)
                worst = max(worst, float(np.abs(got - gold["DM_" + c]).max() / np.abs(gold["DM_" + c]).max()))
                assert close(got, gold["DM_" + c]), (name, c)
            if "baryons" in name:   # E&H: one transfer function for both species - the golden's BA fields equal its DM fields bit for bit
                assert bool(gold["flag__BA_equals_DM_bitwise"])
                for c in FIELDS:
                    ba = np.fromfile(os.path.join(tmp, f"{name}.BA_{c}.f64")).reshape(N, N, N)
                    assert close(ba, gold["DM_" + c]), (name, "BA", c)
            continue
        rs, rt = ("f64", np.float64) if G.units(spec["fmt"])[2] else ("f32", np.float32)
        it, idt = ("u64", np.uint64) if spec["fmt"] == "swift" else ("u32", np.uint32)
        for species, ptype in (("DM", "PartType1"), ("BA", "PartType0")):
            if ptype + "__Coordinates" not in gold:
                continue
            pos = np.fromfile(os.path.join(tmp, f"{name}.{species}.pos.{rs}"), dtype=rt).reshape(-1, 3)
            vel = np.fromfile(os.path.join(tmp, f"{name}.{species}.vel.{rs}"), dtype=rt).reshape(-1, 3)
            assert close(pos, gold[ptype + "__Coordinates"]), (name, species, "pos")
            assert close(vel, gold[ptype + "__Velocities"]), (name, species, "vel")
            if rt == np.float64:
                worst = max(worst, float(np.abs(vel - gold[ptype + "__Velocities"]).max() / np.abs(gold[ptype + "__Velocities"]).max()))
            if ptype + "__Masses" in gold:
                mass = np.fromfile(os.path.join(tmp, f"{name}.{species}.mass.{rs}"), dtype=rt)
                assert close(mass, gold[ptype + "__Masses"]), (name, species, "mass")
            if ptype + "__ParticleIDs" in gold:
                ids = np.fromfile(os.path.join(tmp, f"{name}.{species}.ids.{it}"), dtype=idt)
                assert np.array_equal(ids, gold[ptype + "__ParticleIDs"]), (name, species, "ids")
    print(f"reference goldens through the reference's executable on the library: worst FP64 deviation {worst:.2e} of the field maximum")


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference sources not mounted")
def test_reference_goldens_reproduced_by_its_own_executable_on_the_emulated_library(tmp_path):
    import emu_util

    emu_util.build_emu()
    r = subprocess.run(["make", "-s", "-C", CPP, "-j4", "emu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    _check_goldens(os.path.join(CPP, "_dropin", "monofonIC_mfb_emu"), tmp_path)


@pytest.mark.gpu
def test_reference_goldens_reproduced_by_its_own_executable_on_the_gpu(tmp_path):
    exe = os.path.join(CPP, "_dropin", "monofonIC_mfb")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/_dropin/monofonIC_mfb was not built (needs /root/reference at build time)")
    _check_goldens(exe, tmp_path)


def _check_device_noise_plugin(exe, exe_mpi, tmp, tol):
    """[random] generator = NGENIC_GPU - a plugin on the reference's RNG plugin API that fills the grid on the device with the
    NGENIC plugin's own generator - gives the runs of generator = NGENIC (same seed): `tol` = 0 where host and device share
    their libm (CPU emulation), a few ulp of log / cos / sin otherwise."""
    tmp = str(tmp)
    listing = subprocess.run([exe], capture_output=True, text=True).stdout
    assert "'NGENIC_GPU'" in listing
    _run(exe, tmp, "h2", 2)
    _run(exe, tmp, "g2", 2, generator="NGENIC_GPU")
    a, b = _fields(tmp, "h2"), _fields(tmp, "g2")
    for c in FIELDS:
        assert np.abs(a[c] - b[c]).max() <= tol * np.abs(a[c]).max(), c
    if exe_mpi is not None:     # every rank fills its own k-slab
        _run_mpi(exe_mpi, tmp, "g2w2", 2, 2, generator="NGENIC_GPU")
        for c in FIELDS:
            got = np.concatenate([np.fromfile(os.path.join(tmp, f"g2w2.r{r}.DM_{c}.f64")).reshape(-1, N, N) for r in range(2)])
            assert np.abs(got - a[c]).max() <= max(tol, 1e-13) * np.abs(a[c]).max(), c


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference sources not mounted")
def test_device_noise_plugin_equals_the_ngenic_plugin(tmp_path):
    import emu_util

    emu_util.build_emu()
    r = subprocess.run(["make", "-s", "-C", CPP, "-j4", "emu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    _check_device_noise_plugin(os.path.join(CPP, "_dropin", "monofonIC_mfb_emu"), os.path.join(CPP, "_dropin", "monofonIC_mfb_mpi_emu"), tmp_path, 0.0)


@pytest.mark.gpu
def test_device_noise_plugin_equals_the_ngenic_plugin_on_the_gpu(tmp_path):
    exe = os.path.join(CPP, "_dropin", "monofonIC_mfb")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/_dropin/monofonIC_mfb was not built (needs /root/reference at build time)")
    _check_device_noise_plugin(exe, None, tmp_path, 1e-13)


def _camb_table(path):
    """a synthetic 13-column table in the format of transfer_CAMB_file.cc:66-80 (no CAMB / CLASS output is available offline):
    smooth, positive, DISTINCT cold-dark-matter and baryon transfer functions and velocity transfer functions, so that the
    two-fluid terms of the driver are all non-zero: rho_bc (masses), the theta_bc velocity term, distinct species displacements"""
    k = np.logspace(-4, 3, 512)
    q = k / 0.14
    tc = np.log(1.0 + 2.34 * q) / (2.34 * q) * (1 + 3.89 * q + (16.1 * q) ** 2 + (5.46 * q) ** 3 + (6.71 * q) ** 4) ** -0.25
    tb = tc * (1.0 - 0.02 * k ** 2 / (0.02 + k ** 2)) * (1.0 + 0.005 * np.sin(40.0 * k) * np.exp(-(k / 0.3) ** 2))   # small enough for 1 + C rho_bc > 0 in the PPT branch
    fb = 0.0488911 / 0.3099
    tm = (1 - fb) * tc + fb * tb
    vc, vb = tc * (1.0 + 0.02 * k / (1 + k)), tb * (1.0 - 0.05 * k / (1 + k))
    cols = [k, tc, tb, 0 * k, 0 * k, tm, tm, tm, tm, tm, vc, vb, vb - vc]
    np.savetxt(path, np.stack(cols, 1), fmt="%.17e")


def _check_two_fluid(ref_exe, our_exe, tmp):
    """BASELINE config C4 at the reference's regression size: CDM + baryons with two-fluid transfer functions (CAMB_file plugin,
    the reference's own) and DoBaryonVrel - the reference itself against the drop-in, identical configuration and table"""
    tmp = str(tmp)
    table = os.path.join(tmp, "camb_like.dat")
    _camb_table(table)
    tf = "CAMB_file\ntransfer_file   = " + table
    for tag, kw in (("f2", dict(order=2)), ("p2", dict(order=2, particles=True)), ("f3", dict(order=3, lattice="bcc")),
                    ("e2", dict(order=2, eulerian=True))):
        _run(ref_exe, tmp, "r" + tag, baryons=True, vrel=True, transfer=tf, **kw)
        _run(our_exe, tmp, "d" + tag, baryons=True, vrel=True, transfer=tf, **kw)
        for species in ("DM", "BA"):
            if kw.get("particles"):
                for k, dt in (("pos.f64", np.float64), ("vel.f64", np.float64), ("mass.f64", np.float64), ("ids.u64", np.uint64)):
                    a = np.fromfile(os.path.join(tmp, f"r{tag}.{species}.{k}"), dtype=dt)
                    b = np.fromfile(os.path.join(tmp, f"d{tag}.{species}.{k}"), dtype=dt)
                    if dt == np.uint64:
                        assert np.array_equal(a, b), (tag, species, k)
                    else:
                        assert a.size and np.abs(a - b).max() <= (4e-16 if k.startswith("pos") else TOL * np.abs(a).max()), (tag, species, k)
                mass = np.fromfile(os.path.join(tmp, f"r{tag}.{species}.mass.f64"))
                assert mass.std() > 1e-7 * mass.mean()          # rho_bc really is non-zero with this table
            else:
                comps = ("delta", "vx", "vy", "vz") if kw.get("eulerian") else FIELDS
                for c in comps:
                    a = np.fromfile(os.path.join(tmp, f"r{tag}.{species}_{c}.f64"))
                    b = np.fromfile(os.path.join(tmp, f"d{tag}.{species}_{c}.f64"))
                    assert np.isfinite(a).all() and np.isfinite(b).all(), (tag, species, c)
                    assert a.size == N ** 3 and np.abs(a - b).max() <= (1e-9 if kw.get("eulerian") else TOL) * np.abs(a).max(), (tag, species, c)
        if not kw.get("particles") and not kw.get("eulerian"):   # the species really differ
            a = np.fromfile(os.path.join(tmp, f"r{tag}.DM_vx.f64")); b = np.fromfile(os.path.join(tmp, f"r{tag}.BA_vx.f64"))
            assert np.abs(a - b).max() > 1e-7 * np.abs(a).max()


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference sources not mounted")
def test_two_fluid_transfer_table_reference_itself_vs_dropin(tmp_path):
    import emu_util

    emu_util.build_emu()
    r = subprocess.run(["make", "-s", "-C", CPP, "-j4", "emu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    _check_two_fluid(os.path.join(CPP, "_dropin", "monofonIC_ref_cpu"), os.path.join(CPP, "_dropin", "monofonIC_mfb_emu"), tmp_path)


@pytest.mark.gpu
def test_two_fluid_transfer_table_reference_on_the_host_vs_dropin_on_the_gpu(tmp_path):
    ref, ours = os.path.join(CPP, "_dropin", "monofonIC_ref_cpu"), os.path.join(CPP, "_dropin", "monofonIC_mfb")
    if not (os.path.exists(ref) and os.path.exists(ours)):
        pytest.skip("tests/cpp/_dropin executables were not built (need /root/reference at build time)")
    _check_two_fluid(ref, ours, tmp_path)


# options of the driver that run through host lambdas on the grids (noise conditioning, the primordial non-Gaussianity branch with
# its own dealiased products, external tides) or through other lattice branches: identical configurations through the reference
# itself and through the drop-in
OPTION_SWEEP = {
    "fixing": (2, dict(), lambda t: t.replace("DoFixing        = no", "DoFixing        = yes")),
    "inversion_corners": (2, dict(), lambda t: t.replace("DoInversion     = no", "DoInversion     = yes").replace("DoRemoveCornerModes = no", "DoRemoveCornerModes = yes")),
    "fnl": (2, dict(), lambda t: t.replace("[cosmology]\n", "[cosmology]\nfnl = 50.0\n")),                       # src/ic_generator.cc:402-455
    "fnl_running": (1, dict(), lambda t: t.replace("[cosmology]\n", "[cosmology]\nfnl = 50.0\nnf = 0.5\nk0 = 0.05\n")),
    "gnl": (2, dict(), lambda t: t.replace("[cosmology]\n", "[cosmology]\ngnl = 1.0e4\n")),
    "tides": (2, dict(), lambda t: t.replace("[cosmology]\n", "[cosmology]\nLSS_aniso_lx = 0.1\nLSS_aniso_ly = 0.05\nLSS_aniso_lz = -0.15\n")),  # :238-242, 492-496, 931-934
    "fcc_particles": (3, dict(lattice="fcc", particles=True), None),
    "rsc_particles": (2, dict(lattice="rsc", particles=True), None),
    "baryons_3lpt_bcc_particles": (3, dict(lattice="bcc", particles=True, baryons=True), None),
    # masked sc lattice (include/particle_generator.hh:171-260): both species on one lattice, selected by a 2x2x2 mask
    "masked_particles": (2, dict(lattice="masked", particles=True, baryons=True), None),
    "masked_type1_particles": (2, dict(lattice="masked", particles=True, baryons=True), lambda t: t.replace("[setup]\n", "[setup]\nParticleMaskType = 1\n")),
}


def _check_option_sweep(ref_exe, our_exe, tmp):
    tmp = str(tmp)
    for tag, (order, kw, edit) in OPTION_SWEEP.items():
        _run(ref_exe, tmp, "r_" + tag, order, edit=edit, **kw)
        _run(our_exe, tmp, "d_" + tag, order, edit=edit, **kw)
        species = ("DM", "BA") if kw.get("baryons") else ("DM",)
        for sp in species:
            if kw.get("particles"):
                for k, dt in (("pos.f64", np.float64), ("vel.f64", np.float64), ("ids.u64", np.uint64)):
                    a = np.fromfile(os.path.join(tmp, f"r_{tag}.{sp}.{k}"), dtype=dt)
                    b = np.fromfile(os.path.join(tmp, f"d_{tag}.{sp}.{k}"), dtype=dt)
                    assert a.size and a.size == b.size, (tag, sp, k)
                    if dt == np.uint64:
                        assert np.array_equal(a, b), (tag, sp, k)
                    else:
                        assert np.abs(a - b).max() <= (4e-16 if k.startswith("pos") else TOL * np.abs(a).max()), (tag, sp, k)
            else:
                a, b = _fields(tmp, "r_" + tag, sp), _fields(tmp, "d_" + tag, sp)
                for c in FIELDS:
                    assert np.isfinite(a[c]).all() and np.abs(a[c] - b[c]).max() <= TOL * np.abs(a[c]).max(), (tag, sp, c)
    # the options really change the result
    base = _fields(tmp, "r_fixing")
    for tag in ("fnl", "gnl", "tides", "inversion_corners"):
        assert np.abs(_fields(tmp, "r_" + tag)["vx"] - base["vx"]).max() > 1e-6 * np.abs(base["vx"]).max(), tag


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference sources not mounted")
def test_driver_options_reference_itself_vs_dropin(tmp_path):
    import emu_util

    emu_util.build_emu()
    r = subprocess.run(["make", "-s", "-C", CPP, "-j4", "emu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    _check_option_sweep(os.path.join(CPP, "_dropin", "monofonIC_ref_cpu"), os.path.join(CPP, "_dropin", "monofonIC_mfb_emu"), tmp_path)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference sources not mounted")
def test_naive_convolver_build_reference_itself_vs_dropin(tmp_path):
    """CONVOLVER_TYPE=NAIVE builds of both executables (include/convolution.hh:296-426: the same named products without padding,
    aliased, Nyquist planes kept): identical configurations, 2LPT and 3LPT fields and a MUSIC1 run; and the aliased result really
    differs from the dealiased one"""
    import emu_util

    emu_util.build_emu()
    r = subprocess.run(["make", "-s", "-C", CPP, "-j4", "emu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    tmp = str(tmp_path)
    ref, ours = os.path.join(CPP, "_dropin", "monofonIC_ref_cpu_naive"), os.path.join(CPP, "_dropin", "monofonIC_mfb_naive_emu")
    for tag, order, gen in (("n2", 2, "NGENIC"), ("n3", 3, "NGENIC"), ("nm", 2, "MUSIC1")):
        out = _run(ref, tmp, "r" + tag, order, generator=gen)
        assert "Convolutions : naive" in out
        out = _run(ours, tmp, "d" + tag, order, generator=gen)
        assert "Convolutions : naive" in out
        a, b = _fields(tmp, "r" + tag), _fields(tmp, "d" + tag)
        for c in FIELDS:
            assert np.abs(a[c] - b[c]).max() <= TOL * np.abs(a[c]).max(), (tag, c)
    _run(os.path.join(CPP, "_dropin", "monofonIC_mfb_emu"), tmp, "o3", 3)
    o3, n3 = _fields(tmp, "o3"), _fields(tmp, "dn3")
    assert np.abs(o3["vx"] - n3["vx"]).max() > 1e-7 * np.abs(o3["vx"]).max()
