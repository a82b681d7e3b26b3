#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native monofonIC LPT field pipeline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4] [--grid G] [--order O] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic input: conditioned white noise
(k-space, resident in HBM) -> phi(1) -> phi(2) -> phi(3a,3b) -> A(3) -> growth scaling -> the six
displacement/velocity component fields -> particle AoS (pos, vel, ids) on the device.
metric = Mcells/s = GridRes^3 / step time / 1e6.  Workloads (BASELINE.json configs):
  c3 (default): 3LPT 1024^3, 3/2-padded Orszag convolutions, sc lattice, Gadget-style output (f32, u32 ids)
  c2          : 2LPT 512^3 sc, DM only
  c4          : CDM+baryon 2LPT 1024^3 bcc, two-fluid amplitudes (synthetic delta_bc / theta_bc tables), DoBaryonVrel,
                per-particle masses on the secondary lattice for the baryons
--gpus 8 with the default workload also runs 3LPT 2048^3 (config C5) and attaches it as "north_star_2048".

  value : device-timed, inputs already in HBM (CUDA events on the library stream).
  e2e   : the same step through the public host API with HOST buffers: pinned white noise H2D,
          noise conditioning, pipeline, particle arrays D2H into pinned memory, all inside the timed region.
  roofline : for the kernel role with the largest share of the step; algorithmic bytes
          / average launch duration measured live with CUDA events around every launch.
  fft3d : stand-alone 3-D r2c + c2r FP64 transforms of the N-grid, this library next to cuFFT (torch.fft, outside the
          product path), as algorithmic GB/s = 6 R(N) / time  (BASELINE metric "3D FFT HBM GB/s vs peak").
  parity : (a) sampled displacement / velocity fields (FP64, before the f32 particle cast) against the digest of an N=1 run
          of the same workload (tests/golden/bench_digest_<grid>.npz; `--write-digest` makes it) and IDs checked against
          the closed form; (b) the same pipeline at 256^3 against a digest made by the CPU oracle
          (tests/golden/make_bench_digest.py).  Bar: 1e-9 * max|field| (north_star).
  cpu_baseline / --impl reference : the reference's CPU implementation of the path on the host cores: the reference's own
          sources compiled behind the FFT shim (oracle/_ref, kind "reference") when that binary travelled with the repo,
          else the oracle port (numpy + MKL FFT through torch, kind "port"), on a bounded sample of the workload - the
          sample's grid is named in config.workload.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT]

import numpy as np  # noqa: E402

NOISE_SEED = 12345          # one numpy Generator per x-plane: the global field does not depend on the number of GPUs
SAMPLE_SEED, NSAMPLE = 777, 32768
DPLUS, VFAC, LBOX = 0.0257, 396.69, 100.0
GOLD = os.path.join(ROOT, "tests", "golden")


def synthetic_amplitude(ns=0.9649, kcut=0.5, norm=1.0):
    """synthetic stand-in for k^{ns/2} T(k) sqrt(pnorm) (BBKS-like roll-off): what the transfer-function plugin would be
    tabulated into.  Its values do not influence the timing."""

    def amp(k):
        k = np.asarray(k, dtype=np.float64)
        q = k / kcut
        T = np.log1p(2.34 * q) / np.where(q > 0, 2.34 * q, 1.0) * (
            1 + 3.89 * q + (16.1 * q) ** 2 + (5.46 * q) ** 3 + (6.71 * q) ** 4) ** -0.25
        return norm * np.power(k, 0.5 * ns) * np.where(q > 0, T, 1.0)

    return amp


def noise_plane(n, i):
    """x-plane i of the synthetic real-space white noise"""
    return np.random.default_rng(NOISE_SEED + i).standard_normal((n, n))


def sample_cells(n):
    return np.random.default_rng(SAMPLE_SEED + n).integers(0, n, size=(NSAMPLE, 3)).astype(np.int64)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=["c3", "c2", "c4"])
    ap.add_argument("--grid", type=int, default=0, help="GridRes (default: the workload's)")
    ap.add_argument("--order", type=int, default=0, help="LPT order (default: the workload's)")
    ap.add_argument("--lattice", default="", help="sc|bcc|fcc|rsc (default: the workload's)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-grid", type=int, default=0, help="GridRes of the bounded CPU sample (default: chosen from a calibration run)")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (mfb_set_option), for A/B runs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-fft3d", action="store_true")
    ap.add_argument("--no-2048", action="store_true", help="--gpus 8: skip the 3LPT 2048^3 north-star point")
    ap.add_argument("--write-digest", action="store_true", help="N=1: write tests/golden/bench_digest_<grid>.npz from this run")
    ap.add_argument("--staged-exchange", action="store_true",
                    help="N>1: do the slab exchange with NCCL all_to_all instead of peer stores fused into the FFT passes")
    ap.add_argument("--nslots", type=int, default=0, help="N>1: exchange buffers in the peer-mapped block (0: from free memory)")
    return ap.parse_args()


def workload_defaults(args):
    w = {"c3": (1024, 3, "sc"), "c2": (512, 2, "sc"), "c4": (1024, 2, "bcc")}[args.workload]
    return args.grid or w[0], args.order or w[1], args.lattice or w[2]


# --------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_name(n, order, lattice="sc", two_fluid=False):
    ov = {"sc": 1, "bcc": 2, "fcc": 4, "rsc": 8}[lattice]
    ids = "u64" if ov * n ** 3 * (2 if two_fluid else 1) > 2 ** 32 - 1 else "u32"
    s = f"{order}LPT {n}^3 {lattice} lattice, Orszag 3/2 padded ({3 * n // 2}^3), gadget-style output (f32 pos/vel, {ids} ids)"
    if two_fluid:
        s += ", CDM + baryons (two-fluid amplitudes, DoBaryonVrel, baryon masses from delta_bc on the secondary lattice)"
    return s


# --------------------------------------------------------------------------------------------
# CPU arm.  The ONLY place where bench.py touches oracle/: the cpu_baseline leg and the --impl reference arm.
def cpu_reference_run(n, order, threads, steps=1, warmup=0):
    """the oracle port of the reference pipeline on the host cores (MKL FFT through torch)"""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lpt_oracle as O

    torch.set_num_threads(threads)
    fft = O.TorchFFTBackend(threads)
    g = O.GridSpec(n, LBOX, fft)
    W = O.condition_noise(g, O.make_white_noise_k(n, 1))
    amp = O.power_law_amplitude()
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        O.run_pipeline(n, LBOX, W, amp, order, DPLUS, VFAC, "sc", LBOX, LBOX, False, False, fft=fft)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return float(np.mean(times))


def cpu_compiled_reference_run(n, order, threads, steps=1, warmup=0):
    """the reference's OWN Grid_FFT / OrszagConvolver sources (compiled in the build container into oracle/_ref/ref_bench
    behind the FFT shim of oracle/refshim) running its source stage src/ic_generator.cc:475-574 + the six output FFTs.
    Returns seconds per step, or None if the binary did not travel / does not run here."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_bench")
    if not os.path.exists(exe):
        return None
    try:
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))
        r = subprocess.run([exe, str(n), str(order), str(steps), str(warmup)], capture_output=True, text=True, env=env, timeout=1700)
        if r.returncode != 0:
            return None
        for ln in r.stdout.splitlines():
            if ln.startswith("seconds_per_step"):
                return float(ln.split()[1])
    except Exception:
        return None
    return None


def cpu_arm(order, threads, steps, warmup, grid=0, budget_s=150.0):
    """(seconds per step, grid, kind, description) of the CPU arm on a bounded sample: the largest grid whose
    (steps + warmup) runs fit `budget_s`, judged from a calibration run at 64^3 (cost ~ N^3 log N)."""
    use_ref = cpu_compiled_reference_run(32, order, threads) is not None
    run = cpu_compiled_reference_run if use_ref else cpu_reference_run
    if not grid:
        t64 = run(64, order, threads)
        grid = 64
        for g in (96, 128, 192, 256, 384, 512):
            est = t64 * (g / 64.0) ** 3 * (np.log2(1.5 * g) / np.log2(96.0))
            if est * (steps + warmup) <= budget_s:
                grid = g
    dt = run(grid, order, threads, steps=steps, warmup=warmup)
    kind = "reference" if use_ref else "port"
    what = ("the reference's own Grid_FFT / OrszagConvolver / operators sources (oracle/_ref/ref_bench: compiled where they lie, "
            "FFTW calls served by the mixed-radix OpenMP FFT of oracle/refshim) running the source stage "
            "src/ic_generator.cc:475-574 and the six output transforms" if use_ref else
            "oracle port of the reference algorithm (numpy + torch/MKL FP64 FFT); the compiled reference oracle/_ref/ref_bench "
            "did not travel with the repo")
    return dt, grid, kind, what


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    ng, order, lattice = workload_defaults(args)
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 2))
    dt, n, kind, what = cpu_arm(order, threads, steps, warmup, args.cpu_grid)
    val = n ** 3 / dt / 1e6
    sample = f"full {order}LPT pipeline at {n}^3 (padded {3 * n // 2}^3) on {threads} host threads: {what}"
    line = {"impl": "reference", "metric": "3LPT IC Mcells/s", "value": val, "unit": "Mcells/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n, order, lattice) + f"  [bounded CPU sample of the {ng}^3 workload: "
                                   f"this arm ran {n}^3, not {ng}^3]",
                       "grid": n, "grid_of_gpu_arm": ng, "lpt_order": order, "bounded_sample": sample},
            "cpu_baseline": {"value": val, "unit": "Mcells/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
class _DevBuf:
    """raw device pointer -> torch tensor (via __cuda_array_interface__)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


class Bench:
    """one workload on this rank's slab: pipeline, noise, timing and parity helpers"""

    def __init__(self, M, lib, torch, dist, local_rank, n, order, lattice, two_fluid=False):
        self.M, self.lib, self.torch, self.dist, self.local_rank = M, lib, torch, dist, local_rank
        self.n, self.order, self.lattice, self.two_fluid = n, order, lattice, two_fluid
        self.world = dist.get_world_size() if dist is not None else 1
        ov = {"sc": 1, "bcc": 2, "fcc": 4, "rsc": 8}[lattice]
        self.ids64 = ov * n ** 3 * (2 if two_fluid else 1) > 2 ** 32 - 1   # UseLongids
        self.cfg = M.lpt_config(GridRes=n, BoxLength=LBOX, LPTorder=order, ParticleLoad=lattice, Dplus0=DPLUS, vfac=VFAC,
                                lunit=LBOX, vunit=LBOX / np.sqrt(1 / 51.0), real64=False, ids64=self.ids64)
        self.pl = M.LPTPipeline(self.cfg)
        self.pl.make_particles(with_masses=two_fluid)
        self.table = M.pipeline.radial_table(n, LBOX, synthetic_amplitude())
        if two_fluid:
            # stand-ins for get_amplitude_theta_bc / get_amplitude_delta_bc of a two-fluid transfer function
            # (include/cosmology_calculator.hh:416-431); CAMB/CLASS tables are not available offline
            amp = synthetic_amplitude()
            self.theta_table = M.pipeline.radial_table(n, LBOX, lambda k: 0.37 * amp(k) / (1.0 + k))
            self.delta_table = M.pipeline.radial_table(n, LBOX, lambda k: 2e-3 * amp(k) * k * k)
            self.rho = M.Grid_FFT((n,) * 3, (LBOX,) * 3)
        self.host_noise, self.out, self.outs, self.out_done = None, None, [], []
        self.noise_in_flight = False

    # ---- synthetic white noise: real-space Gaussian field generated on the host, one seed per x-plane
    def make_noise(self):
        n, w = self.n, self.pl.wnoise
        l0, o0 = w.local_0_size_, w.local_0_start_
        self.host_noise = self.torch.empty((w.ntot_,), dtype=self.torch.float64, pin_memory=True)
        hn = self.host_noise.numpy()
        hn[:] = 0.0
        hv = hn[:l0 * n * (n + 2)].reshape(l0, n, n + 2)
        for i in range(l0):
            hv[i, :, :n] = noise_plane(n, o0 + i)
        return hn

    def upload_noise(self):
        w = self.pl.wnoise
        w.FourierTransformBackward(False)
        w.upload_raw(self.host_noise.numpy())
        w.FourierTransformForward()
        self.pl.condition_noise()

    def make_out(self):
        """pinned host arrays the output plugin would be handed; two sets when the host has the memory, so that the
        consumer of run k can still be reading while run k+1 streams its particles out"""
        t, npart = self.torch, self.pl.particles.get_local_num_particles()

        def one():
            o = {"pos": t.empty((npart, 3), dtype=t.float32, pin_memory=True).numpy(),
                 "vel": t.empty((npart, 3), dtype=t.float32, pin_memory=True).numpy(),
                 "ids": (t.empty((npart,), dtype=t.int64, pin_memory=True).numpy().view(np.uint64) if self.ids64 else
                         t.empty((npart,), dtype=t.int32, pin_memory=True).numpy().view(np.uint32))}
            if self.two_fluid:
                o["mass"] = t.empty((npart,), dtype=t.float32, pin_memory=True).numpy()
            return o
        self.outs = [one()]
        if not self.two_fluid:
            try:
                self.outs.append(one())
            except Exception:
                pass   # host memory is short: one set, every run waits for the previous hand-off
        self.out = self.outs[0]
        self.out_done = [None] * len(self.outs)
        return self.out

    # ---- one step
    def species(self, stream_to=None):
        """the species loop of src/ic_generator.cc:619-955 for the workload"""
        pl = self.pl
        if not self.two_fluid:
            pl.fill_particles(stream_to=stream_to)
            return
        import ctypes as C

        c = self.cfg
        Om, Ob = 0.3, 0.049
        fb = Ob / Om
        p = pl.particles
        for ispec, (C_species, Omega) in enumerate(((-fb, Om - Ob), (1.0 - fb, Ob))):   # dm, baryon (:619-631)
            p.reset_species(ispec, ispec)      # baryons: secondary lattice, ID offset 1 (:636-645)
            # mass field (1 + C rho) Omega munit (:663-675) -> per-particle masses
            t = self.delta_table
            rc = self.lib.mfb_apply_radial_table(self.rho._h, pl.wnoise._h, t.ctypes.data_as(C.c_void_p), t.size, 0)
            assert rc == 0, self.lib.mfb_last_error()
            self.rho.zero_DC_mode()
            self.rho.FourierTransformBackward()
            self.rho.affine(Omega, C_species * Omega)
            p.set_masses(1.0, self.rho)
            for d in range(3):
                p.set_positions(d, c.lunit, pl.displacement_field(d))
            for d in range(3):
                p.set_velocities(d, pl.velocity_field(d, vbc=(self.theta_table, c.vfac * C_species)))
            if stream_to is not None:
                # the output plugin takes one species at a time (write_particle_data per species, :944-955)
                p.download(out=stream_to)

    def step_device(self):
        pl = self.pl
        pl.compute_phi1(self.table)
        pl.compute_sources()
        self.species()

    def step_e2e(self, k=0):
        """one run through the host API.  Two-fluid: upload, pipeline, one blocking hand-off per species.  Single fluid: a
        series of runs as a pipeline - this run's noise was streamed in under the previous run (the noise grid is dead once
        phi1 exists), the next run's is started here, and the particle hand-off of this run drains under the first kernels
        of the next one.  Every run still moves its own H2D and D2H bytes inside the timed region."""
        pl = self.pl
        if self.two_fluid:
            self.upload_noise()
            pl.compute_phi1(self.table)
            pl.compute_sources()
            self.species(stream_to=self.out)
            return
        w = pl.wnoise
        if self.noise_in_flight:
            w.stream_in_wait()                   # the compute stream waits for the ingest stream; the host does not
            w.FourierTransformBackward(False)    # what arrived is a real-space field
            w.FourierTransformForward()
            pl.condition_noise()
        else:
            self.upload_noise()
        out = self.outs[k % len(self.outs)]
        if self.out_done[k % len(self.outs)] is not None:
            self.out_done[k % len(self.outs)].synchronize()   # the consumer got run k-2's arrays before they are reused
        self.out = out
        pl.particles.stream_out("ids", out["ids"])            # the IDs do not depend on the fields
        pl._ids_streamed = True
        pl.compute_phi1(self.table)
        w.stream_in(self.host_noise.numpy())                  # next run's noise, under this run's source stage
        self.noise_in_flight = True
        pl.compute_sources()
        pl.fill_particles(stream_to=out)                      # pos / vel handed off as each completes (copy stream)
        ev = self.torch.cuda.Event()
        ev.record(self.torch.cuda.ExternalStream(self.lib.mfb_copy_stream()))
        self.out_done[k % len(self.outs)] = ev

    def drain_e2e(self):
        """end of a series of runs: every hand-off has landed, the ingest stream is idle"""
        if self.pl.particles is not None:
            self.pl.particles.stream_wait()
        if self.noise_in_flight:
            self.pl.wnoise.stream_in_wait()
            self.noise_in_flight = False
            self.pl.wnoise.FourierTransformBackward(False)
            self.pl.wnoise.FourierTransformForward()
            self.pl.condition_noise()

    # ---- parity helpers
    def field_samples(self):
        """FP64 Psi_d and v_d (the fields the particles are read out from) at the sample cells of this slab"""
        t, n, pl = self.torch, self.n, self.pl
        cells = sample_cells(n)
        l0, o0 = pl.tmp.local_0_size_, pl.tmp.local_0_start_
        mine = np.nonzero((cells[:, 0] >= o0) & (cells[:, 0] < o0 + l0))[0]
        c = cells[mine]
        idx = t.from_numpy(((c[:, 0] - o0) * n + c[:, 1]) * (n + 2) + c[:, 2]).to(f"cuda:{self.local_rank}")
        nbytes = l0 * n * (n + 2) * 8
        vals, maxabs = np.zeros((6, len(mine))), np.zeros(6)
        for d in range(3):
            for kind in range(2):
                g = pl.displacement_field(d) if kind == 0 else pl.velocity_field(d)
                self.M.synchronize()
                buf = t.as_tensor(_DevBuf(self.lib.mfb_grid_device_ptr(g._h), nbytes), device=f"cuda:{self.local_rank}").view(t.float64)
                vals[3 * kind + d] = buf[idx].cpu().numpy()
                rows, m = buf.view(l0 * n, n + 2), 0.0
                for r0 in range(0, l0 * n, 16 * n):     # plane blocks: no grid-sized temporary next to a full HBM
                    m = max(m, float(rows[r0:r0 + 16 * n, :n].abs().max().item()))
                maxabs[3 * kind + d] = m
                del rows
                del buf
                t.cuda.synchronize()
        return mine, vals, maxabs

    def allmax(self, x):
        if self.dist is None:
            return float(x)
        v = self.torch.tensor([float(x)], dtype=self.torch.float64, device=f"cuda:{self.local_rank}")
        self.dist.all_reduce(v, op=self.dist.ReduceOp.MAX)
        return float(v.item())

    def check_against_digest(self, path, label):
        if not os.path.exists(path):
            return {"checked": False, "why": f"{os.path.relpath(path, ROOT)} not found"}
        dg = np.load(path)
        if int(dg["n"]) != self.n or int(dg["order"]) != self.order:
            return {"checked": False, "why": "digest is for another workload"}
        mine, vals, maxabs = self.field_samples()
        ref = dg["fields"][:, mine]
        mp, mv = float(dg["maxabs"][:3].max()), float(dg["maxabs"][3:].max())
        ep = np.abs(vals[:3] - ref[:3]).max() / mp if len(mine) else 0.0
        ev = np.abs(vals[3:] - ref[3:]).max() / mv if len(mine) else 0.0
        ep, ev = self.allmax(ep), self.allmax(ev)
        return {"checked": True, "against": label, "samples": int(NSAMPLE), "max_rel_psi": ep, "max_rel_vel": ev,
                "tolerance": 1e-9, "ok": bool(ep <= 1e-9 and ev <= 1e-9)}

    def check_ids(self):
        """IDs of the first lattice shift against the closed form id = ov * cell + iload (include/particle_generator.hh:129-170)
        and the SHA-256 of this rank's id array (for the record)"""
        if self.out is None:
            return None
        n, p = self.n, self.pl.particles
        l0, o0 = self.pl.tmp.local_0_size_, self.pl.tmp.local_0_start_
        ov = {"sc": 1, "bcc": 2, "fcc": 4, "rsc": 8}[self.lattice]
        ids = self.out["ids"]
        ncell = l0 * n * n
        ok = True
        for s in range(ov):
            want = np.arange(o0 * n * n, (o0 + l0) * n * n, dtype=ids.dtype) * ids.dtype.type(ov) + ids.dtype.type(s)
            if self.two_fluid:     # the container holds the last species written (baryons: id offset 1)
                want = want + ids.dtype.type(ov * n ** 3)
            ok = ok and bool(np.array_equal(ids[s * ncell:(s + 1) * ncell], want))
        ok = self.allmax(0.0 if ok else 1.0) == 0.0
        return {"ids_equal": ok, "sha256_rank0": hashlib.sha256(ids.tobytes()).hexdigest()[:16]}


def fft3d_leg(M, lib, torch, n, local_rank, reps=3):
    """stand-alone 3-D FP64 r2c + c2r of an n^3 grid: this library (in place, norm fused) next to cuFFT (torch.fft, out of place).
    GB/s are algorithmic: 6 R(n) per transform (three sweeps, each reading and writing the array once)."""
    R = 8.0 * n * n * (n + 2)
    res = {"grid": n, "alg_gb_per_transform": 6 * R / 1e9}
    dev = f"cuda:{local_rank}"
    stream = torch.cuda.ExternalStream(lib.mfb_stream())
    g = M.Grid_FFT((n,) * 3, (LBOX,) * 3)
    g.zero()
    for _ in range(2):
        g.FourierTransformForward(); g.FourierTransformBackward()
    M.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for _ in range(reps):
        e[0].record(stream); g.FourierTransformForward(); e[1].record(stream); g.FourierTransformBackward(); e[2].record(stream)
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]) / reps
        tb += e[1].elapsed_time(e[2]) / reps
    del g
    res["ours"] = {"r2c_ms": tf, "c2r_ms": tb, "r2c_gbs": 6 * R / 1e9 / (tf / 1e3), "c2r_gbs": 6 * R / 1e9 / (tb / 1e3)}
    try:
        x = torch.zeros((n, n, n), dtype=torch.float64, device=dev)
        for _ in range(2):
            y = torch.fft.rfftn(x); x2 = torch.fft.irfftn(y, s=(n, n, n)); del y, x2
        torch.cuda.synchronize()
        tf = tb = 0.0
        for _ in range(reps):
            e[0].record(); y = torch.fft.rfftn(x); e[1].record(); x2 = torch.fft.irfftn(y, s=(n, n, n)); e[2].record()
            torch.cuda.synchronize()
            tf += e[0].elapsed_time(e[1]) / reps
            tb += e[1].elapsed_time(e[2]) / reps
            del y, x2
        del x
        res["cufft"] = {"r2c_ms": tf, "c2r_ms": tb, "r2c_gbs": 6 * R / 1e9 / (tf / 1e3), "c2r_gbs": 6 * R / 1e9 / (tb / 1e3),
                        "note": "torch.fft.rfftn / irfftn (cuFFT D2Z / Z2D, out of place, unnormalised forward)"}
        res["ours_over_cufft"] = (res["cufft"]["r2c_ms"] + res["cufft"]["c2r_ms"]) / (res["ours"]["r2c_ms"] + res["ours"]["c2r_ms"])
    except Exception as ex:   # cuFFT plan / memory failure must not take the bench down
        res["cufft"] = {"unavailable": str(ex)[:200]}
    torch.cuda.empty_cache()
    return res


def time_steps(B, torch, stream, steps, barrier, max_over_ranks):
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        B.step_device()
    e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    return max_over_ranks(e0.elapsed_time(e1) / steps)   # device time, max over ranks


def role_table(lib, steps):
    import ctypes as C

    roles = []
    for r in range(lib.mfb_profile_num_roles()):
        t, c, b = C.c_double(), C.c_longlong(), C.c_double()
        lib.mfb_profile_get(r, C.byref(t), C.byref(c), C.byref(b))
        if c.value:
            roles.append({"role": lib.mfb_profile_role_name(r).decode(), "launches_per_step": c.value // steps,
                          "ms_per_step": t.value / steps, "avg_ms": t.value / c.value,
                          "alg_gb_per_launch": b.value / c.value / 1e9,
                          "gbs": (b.value / 1e9) / (t.value / 1e3) if t.value > 0 and b.value > 0 else None})
    return roles


def run_b200(args):
    import torch

    import monofonic_b200 as M

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n, order, lattice = workload_defaults(args)
    two_fluid = args.workload == "c4"
    torch.cuda.set_device(local_rank)
    dist = None
    do_2048 = world == 8 and args.workload == "c3" and not args.grid and not args.no_2048
    if world > 1:
        # one process per GPU (torchrun); slabs of the same global grid on every N => strong scaling
        import torch.distributed as dist
        from monofonic_b200 import dist as MD

        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        lib = MD.init(local_rank, n, order > 1, staged=args.staged_exchange, lpt_order=order, nslots=args.nslots or None)
    else:
        lib = M.init(local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for kv in args.opt:
        k, v = kv.split("=")
        assert lib.mfb_set_option(k.encode(), int(v)) == 0, lib.mfb_last_error()
    stream = torch.cuda.ExternalStream(lib.mfb_stream())
    warm = max(3, args.warmup)

    # ---- stand-alone 3-D FFT next to cuFFT (N = 1 only, before the pipeline takes the memory)
    fft3d = None
    if world == 1 and not args.no_fft3d:
        fft3d = fft3d_leg(M, lib, torch, n, local_rank)

    B = Bench(M, lib, torch, dist, local_rank, n, order, lattice, two_fluid)
    pl = B.pl
    hn = B.make_noise()
    B.upload_noise()
    if not args.no_e2e:
        B.make_out()

    # ---- warm-up
    for _ in range(warm):
        B.step_device()
    M.synchronize()
    barrier()

    # ---- timed region: device-resident inputs, per-role CUDA events on
    lib.mfb_profile_enable(1)
    lib.mfb_profile_reset()
    clocks = ClockSampler(local_rank)
    clocks.start()
    l0c = M.kernel_launches()
    ms = time_steps(B, torch, stream, args.steps, barrier, max_over_ranks)
    launches = (M.kernel_launches() - l0c) // args.steps * world
    clk = clocks.stop()
    roles = role_table(lib, args.steps)
    lib.mfb_profile_enable(0)
    peak, peak_src = peaks()
    fftroles = [r for r in roles if r["gbs"]]
    dom = max(fftroles, key=lambda r: r["ms_per_step"])
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture of this very pipeline (per launch, next
    # to the algorithmic bytes of the captured launch); null when no capture matches kernel role and padded length
    traffic, traffic_note = None, None
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        ent = cap.get(f"{dom['role']}@{3 * n // 2 if dom['role'].startswith('conv') else n}")
        if ent and world == 1:
            # the capture is one launch; scale to this run's average launch by the algorithmic bytes
            scale = dom["alg_gb_per_launch"] * 1e9 / ent["alg_bytes_of_captured_launch"]
            traffic = ent["dram_bytes_per_launch"] * scale
            traffic_note = {"captured_launch_alg_bytes": ent["alg_bytes_of_captured_launch"],
                            "captured_launch_dram_bytes": ent["dram_bytes_per_launch"], "source": ent["source"],
                            # what bounds the kernel inside the SM (same capture): busiest pipes, % of their peak
                            "pipes_pct_of_peak": ent.get("pipes_pct_of_peak")}
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom["role"], "achieved": dom["gbs"], "peak": peak, "unit": "GB/s",
                "frac": dom["gbs"] / peak, "traffic": traffic, "traffic_capture": traffic_note, "peak_source": peak_src,
                "share_of_step": dom["ms_per_step"] / ms, "avg_launch_ms": dom["avg_ms"],
                "alg_bytes_per_launch": dom["alg_gb_per_launch"] * 1e9,
                "frac_of_8000_nominal": dom["gbs"] / 8000.0,
                "per_kernel_frac": {r["role"]: r["gbs"] / peak for r in fftroles}}
    if traffic_note and traffic_note.get("pipes_pct_of_peak"):
        # the dominant kernel is bound inside the SM, not by HBM: its busiest pipe in the committed ncu capture
        pipe, pct = max(traffic_note["pipes_pct_of_peak"].items(), key=lambda kv: kv[1])
        roofline["onchip_bound"] = {"pipe": pipe, "frac_of_pipe_peak": pct / 100.0, "source": "ncu capture named in traffic_capture"}
    total_alg = sum(r["alg_gb_per_launch"] * r["launches_per_step"] for r in fftroles)
    sinfo = pl.conv.sources_info() if pl.conv is not None else {"path": 0}

    # ---- end to end through the host API (host buffers, H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        B.step_e2e(0)   # untimed: primes the pipeline (its stream_in carries the noise of the first timed run)
        if two_fluid:
            M.synchronize()
        else:
            B.pl.particles.stream_wait()
            torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        ksteps = max(1, args.steps)
        for k in range(ksteps):
            B.step_e2e(k + 1)
        B.drain_e2e() if not two_fluid else None   # last hand-off landed; the K-th upload of the region has arrived too
        M.synchronize()
        torch.cuda.synchronize()
        barrier()
        dt_local = (time.perf_counter() - t0) / ksteps
        dt = max_over_ranks(dt_local)
        out = B.out
        nsp = 2 if two_fluid else 1
        e2e = {"value": n ** 3 / dt / 1e6, "unit": "Mcells/s", "ms_per_step": dt * 1e3,
               "h2d_bytes_per_step": int(hn.nbytes + B.table.nbytes) * world,
               "d2h_bytes_per_step": int(sum(v.nbytes for v in out.values())) * world * nsp,
               "this_rank_ms": dt_local * 1e3,
               "host_output_sets": len(B.outs),
               "includes": "a series of runs through the host API, every run with its own pinned noise H2D (streamed in on an "
                           "ingest stream under the previous run's source stage), r2c FFT + conditioning, pipeline and particle "
                           "AoS D2H into pinned host arrays (ids / pos / vel handed off on a copy stream as each completes, "
                           "draining under the next run's first kernels); all ranks" if not two_fluid else
                           "pinned noise H2D, r2c FFT + conditioning, pipeline, one blocking particle hand-off per species"}

    # ---- parity: IDs (closed form), sampled FP64 fields against the N=1 digest of this workload
    parity = None
    if not args.no_parity:
        dpath = os.path.join(GOLD, f"bench_digest_{args.workload}_{n}.npz")
        if args.write_digest and world == 1:
            mine, vals, maxabs = B.field_samples()
            fields = np.zeros((6, NSAMPLE))
            fields[:, mine] = vals
            np.savez_compressed(dpath, n=n, order=order, fields=fields, maxabs=maxabs, cells=sample_cells(n).astype(np.int32))
        parity = {"vs_n1": B.check_against_digest(dpath, f"N=1 run of the same workload ({os.path.relpath(dpath, ROOT)})")}
        if not args.no_e2e:
            parity["ids"] = B.check_ids()
        if args.workload == "c2" and n == 512 and order == 2 and lattice == "sc":
            # SURVEY C2 is small enough for the CPU oracle itself: the whole workload against its digest
            parity["vs_oracle"] = B.check_against_digest(os.path.join(GOLD, "bench_digest_oracle_512_2lpt.npz"),
                                                         "CPU oracle, 2LPT 512^3 (tests/golden/make_bench_digest.py 512 2)")

    # ---- CPU baseline: bounded sample of the same workload on the host cores
    cpu = None
    if not args.no_cpu and rank == 0 and world == 1:
        threads = os.cpu_count() or 1
        dtc, ncpu, kind, what = cpu_arm(order, threads, 1, 0, args.cpu_grid, budget_s=25.0)
        cpu = {"value": ncpu ** 3 / dtc / 1e6, "unit": "Mcells/s", "cores": threads, "kind": kind,
               "sample": f"full {order}LPT pipeline at {ncpu}^3 (padded {3 * ncpu // 2}^3) = {dtc:.1f} s on {threads} threads: {what}; "
                         f"FFT cost grows as N^3 log N and the {n}^3 working set leaves the caches, so the {n}^3 figure would be lower"}

    nex = lib.mfb_dist_exchanges() if dist is not None else 0
    nslots = lib.mfb_dist_nslots() if dist is not None else 0

    # ---- parity against the CPU oracle at 256^3 (digest made by tests/golden/make_bench_digest.py) on this many GPUs, and
    # the 2048^3 north-star point on 8: both need the memory of the main pipeline
    del B, pl
    import gc
    gc.collect()
    oracle256 = None
    if not args.no_parity and args.workload == "c3":
        try:
            B2 = Bench(M, lib, torch, dist, local_rank, 256, 3, "sc")
            B2.make_noise(); B2.upload_noise()
            B2.step_device()
            oracle256 = B2.check_against_digest(os.path.join(GOLD, "bench_digest_oracle_256.npz"),
                                                "CPU oracle, 3LPT 256^3 (tests/golden/make_bench_digest.py)")
            oracle256["source_stage_path"] = B2.pl.conv.sources_info()["path"]
            del B2
            gc.collect()
        except Exception as ex:
            oracle256 = {"checked": False, "why": str(ex)[:300]}
    ns2048 = None
    if do_2048:
        ns2048 = run_2048(M, lib, torch, dist, local_rank, stream, barrier, max_over_ranks, peak)

    if dist is not None:
        from monofonic_b200 import dist as MD

        MD.check()
        barrier()
        MD.shutdown()
        dist.destroy_process_group()
        if rank != 0:
            return
    # the reference does 33 products (each with its own forward transform) at 3LPT, 5 at 2LPT
    grouped = sinfo["path"] == 2
    nprod = {1: 0, 2: 5, 3: 32 if grouped else 33}[order]
    line = {"metric": "3LPT IC Mcells/s", "value": n ** 3 / (ms / 1e3) / 1e6, "unit": "Mcells/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n, order, lattice, two_fluid), "workload_id": args.workload, "grid": n, "lpt_order": order,
                       "l2": "inputs larger than L2 (every grid >= 1 GB; no flush needed)",
                       "products_per_step": nprod,
                       "source_stage": ({"path": "shared x passes + x-plane chunks + multi-product z pass", **sinfo}
                                        if sinfo["path"] == 2 else {"path": "one padded transform pair per product", **sinfo}),
                       "decomposition": (f"{world} x-slabs (real) / ky-slabs (k-space, transposed), exchange fused into the "
                                         f"FFT passes over peer memory, {nslots} exchange buffers, "
                                         f"{nex} exchanges in the whole run" if world > 1 else "single slab")},
            "gpu_launches": int(launches), "clocks": clk, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
            "fft3d": fft3d, "parity": parity, "parity_oracle_256": oracle256, "north_star_2048": ns2048,
            "nvlink": nvlink_table(roles, n, world, order),
            "pipeline_alg_gb_per_step": total_alg, "pipeline_gbs": total_alg / (ms / 1e3), "stages": roles}
    print(json.dumps(line))


def nvlink_table(roles, n, world, order):
    """NVLink rate of the passes that carry an exchange (their stores go straight into the peers' receive buffers): bytes this
    rank sends to its peers per step / the pass time, against 900 GB/s per direction.  x-bwd: one pruned [ky_loc][P0][N/2]
    intermediate per pass; y-fwd: the [P0_loc][N][N/2] y-truncated target chunks; fft_x: the [ky_loc][N][N/2+1] spectra of the
    six output transforms"""
    if world < 2:
        return None
    P, f = world, (world - 1) / world
    per_step = {"conv_x_bwd": lambda r: r["launches_per_step"] * (n // P) * (1.5 * n) * (n // 2) * 16.0,
                "conv_y_fwd": lambda r: {2: 1, 3: 10}.get(order, 0) * (1.5 * n / P) * n * (n // 2) * 16.0,
                "fft_x": lambda r: r["launches_per_step"] * (n // P) * n * (n // 2 + 1) * 16.0}
    out = {}
    for r in roles:
        if r["role"] in per_step and r["ms_per_step"] > 0:
            gb = per_step[r["role"]](r) * f / 1e9
            out[r["role"]] = {"gb_to_peers_per_step": gb, "ms_per_step": r["ms_per_step"], "gbs": gb / (r["ms_per_step"] / 1e3),
                              "frac_of_900": gb / (r["ms_per_step"] / 1e3) / 900.0}
    return out


def run_2048(M, lib, torch, dist, local_rank, stream, barrier, max_over_ranks, peak):
    """config C5: 3LPT 2048^3 on 8 GPUs (device-timed; the 8 x 43 GB of host particle buffers of an e2e step exceed the box)"""
    try:
        n = 2048
        # the peer-mapped block was sized for the main workload: carve it again for the 2048^3 slabs
        from monofonic_b200 import dist as MD

        M.synchronize()
        barrier()
        MD.shutdown()
        import gc
        gc.collect()
        torch.cuda.empty_cache()   # what torch cached for the 1024^3 run's checks counts as free memory for the slot workspace
        lib = MD.init(local_rank, n, True, lpt_order=3)
        B = Bench(M, lib, torch, dist, local_rank, n, 3, "sc")
        B.make_noise(); B.upload_noise()
        B.host_noise = None
        for _ in range(2):
            B.step_device()
        M.synchronize()
        lib.mfb_profile_enable(1)
        lib.mfb_profile_reset()
        ms = time_steps(B, torch, stream, 2, barrier, max_over_ranks)
        roles = role_table(lib, 2)
        lib.mfb_profile_enable(0)
        sinfo = B.pl.conv.sources_info()
        # NVLink: the passes that carry an exchange move 7/8 of their routed output to peers
        nv = {}
        for r in roles:
            if r["role"] in ("conv_x_bwd", "conv_y_fwd", "fft_x", "fft_y"):
                nv[r["role"]] = r
        res = {"value": n ** 3 / (ms / 1e3) / 1e6, "unit": "Mcells/s", "ms_per_step": ms, "steps": 2, "warmup": 2,
               "workload": workload_name(n, 3, "sc"), "source_stage": sinfo, "exchange_buffers": int(lib.mfb_dist_nslots()),
               "frac_of_hbm_ceiling_6641": n ** 3 / (ms / 1e3) / 1e6 / 6641.0,
               "stages": roles, "nvlink": nvlink_table(roles, n, dist.get_world_size(), 3),
               "per_kernel_frac": {r["role"]: r["gbs"] / peak for r in roles if r["gbs"]},
               "parity": "same code path as the 256^3 oracle check and the 1024^3 N=1 digest check of this line (no 2048^3 "
                         "single-GPU or CPU run is possible: 3LPT 2048^3 needs 8 GPUs, SURVEY App. C)"}
        del B
        return res
    except Exception as ex:
        return {"unavailable": str(ex)[:300]}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
