#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native monofonIC LPT field pipeline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--grid G] [--order O] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic input: conditioned white noise
(k-space, resident in HBM) -> phi(1) -> phi(2) -> phi(3a,3b) -> A(3) -> growth scaling -> the six
displacement/velocity component fields -> particle AoS (pos, vel, ids) on the device.
metric = Mcells/s = GridRes^3 / step time / 1e6.  Default workload at N=1:
BASELINE config "3LPT 1024^3, 3/2-padded Orszag convolutions, sc lattice, Gadget output (f32, u32 ids)".

  value : device-timed, inputs already in HBM (CUDA events on the library stream).
  e2e   : the same step through the public host API with HOST buffers: pinned white noise H2D,
          noise conditioning, pipeline, particle arrays D2H into pinned memory, all inside the timed region.
  roofline : for the kernel role with the largest share of the step; algorithmic bytes
          / average launch duration measured live with CUDA events around every launch.
  cpu_baseline / --impl reference : the CPU oracle port of the reference algorithm (oracle/lpt_oracle.py,
          MKL FFT via torch on all host threads; the reference itself needs FFTW/GSL/HDF5/MPI which are not
          installed) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT]

import numpy as np  # noqa: E402


def synthetic_amplitude(ns=0.9649, kcut=0.5):
    """synthetic stand-in for k^{ns/2} T(k) sqrt(pnorm) (BBKS-like roll-off): what the transfer-function plugin would be
    tabulated into.  Its values do not influence the timing."""

    def amp(k):
        k = np.asarray(k, dtype=np.float64)
        q = k / kcut
        T = np.log1p(2.34 * q) / np.where(q > 0, 2.34 * q, 1.0) * (
            1 + 3.89 * q + (16.1 * q) ** 2 + (5.46 * q) ** 3 + (6.71 * q) ** 4) ** -0.25
        return np.power(k, 0.5 * ns) * np.where(q > 0, T, 1.0)

    return amp


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--grid", type=int, default=0, help="GridRes (default: 1024 at N=1)")
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-grid", type=int, default=0, help="GridRes of the bounded CPU sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--staged-exchange", action="store_true",
                    help="N>1: do the slab exchange with NCCL all_to_all instead of peer stores fused into the FFT passes")
    return ap.parse_args()


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


def workload_name(n, order, lattice="sc"):
    ids = "u64" if n ** 3 > 2 ** 32 - 1 else "u32"
    return f"{order}LPT {n}^3 {lattice} lattice, Orszag 3/2 padded ({3 * n // 2}^3), gadget-style output (f32 pos/vel, {ids} ids)"


# --------------------------------------------------------------------------------------------
def cpu_reference_run(n, order, threads, steps=1, warmup=0):
    """the oracle port of the reference pipeline on the host cores (MKL FFT through torch).  The ONLY place where bench.py
    touches oracle/: the cpu_baseline leg and the --impl reference arm."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lpt_oracle as O

    torch.set_num_threads(threads)
    fft = O.TorchFFTBackend(threads)
    g = O.GridSpec(n, 100.0, fft)
    W = O.condition_noise(g, O.make_white_noise_k(n, 1))
    amp = O.power_law_amplitude()
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        O.run_pipeline(n, 100.0, W, amp, order, 0.0257, 396.69, "sc", 100.0, 100.0, False, False, fft=fft)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return float(np.mean(times))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.cpu_grid or 128
    ng = args.grid or 1024
    dt = cpu_reference_run(n, args.order, threads, steps=max(1, args.steps), warmup=min(1, args.warmup))
    val = n ** 3 / dt / 1e6
    sample = (f"full {args.order}LPT pipeline at {n}^3 (padded {3 * n // 2}^3), oracle port of the reference algorithm, "
              f"torch/MKL FP64 FFT, {threads} threads; the reference's own FFTW/GSL/HDF5/MPI build is not available")
    line = {"impl": "reference", "metric": "3LPT IC Mcells/s", "value": val, "unit": "Mcells/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(ng, args.order), "bounded_sample": sample},
            "cpu_baseline": {"value": val, "unit": "Mcells/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
def run_b200(args):
    import torch

    import monofonic_b200 as M

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.grid or 1024
    order = args.order
    Lbox = 100.0
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        # one process per GPU (torchrun); slabs of the same global grid on every N => strong scaling
        import torch.distributed as dist
        from monofonic_b200 import dist as MD

        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        lib = MD.init(local_rank, n, order > 1, staged=args.staged_exchange, lpt_order=order)
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
    stream = torch.cuda.ExternalStream(lib.mfb_stream())
    ids64 = n ** 3 > 2 ** 32 - 1       # UseLongids: 2048^3 does not fit 32-bit IDs
    cfg = M.lpt_config(GridRes=n, BoxLength=Lbox, LPTorder=order, ParticleLoad="sc", Dplus0=0.0257, vfac=396.69,
                       lunit=Lbox, vunit=Lbox / np.sqrt(1 / 51.0), real64=False, ids64=ids64)
    pl = M.LPTPipeline(cfg)
    pl.make_particles()
    table = M.pipeline.radial_table(n, Lbox, synthetic_amplitude())

    # synthetic white noise: real-space Gaussian field generated on the host (one seed per x-plane, so the global
    # field does not depend on the number of GPUs), transformed on the device.  Each rank holds its slab.
    nz2 = n + 2
    l0, o0 = pl.wnoise.local_0_size_, pl.wnoise.local_0_start_
    host_noise = torch.empty((pl.wnoise.ntot_,), dtype=torch.float64, pin_memory=True)
    hn = host_noise.numpy()
    hn[:] = 0.0
    hv = hn[:l0 * n * nz2].reshape(l0, n, nz2)
    for i in range(l0):
        hv[i, :, :n] = np.random.default_rng(12345 + o0 + i).standard_normal((n, n))

    def upload_noise():
        pl.wnoise.FourierTransformBackward(False)
        pl.wnoise.upload_raw(hn)
        pl.wnoise.FourierTransformForward()
        pl.condition_noise()

    upload_noise()
    npart = pl.particles.get_local_num_particles()
    out = None if args.no_e2e else {
           "pos": torch.empty((npart, 3), dtype=torch.float32, pin_memory=True).numpy(),
           "vel": torch.empty((npart, 3), dtype=torch.float32, pin_memory=True).numpy(),
           "ids": (torch.empty((npart,), dtype=torch.int64, pin_memory=True).numpy().view(np.uint64) if ids64 else
                   torch.empty((npart,), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32))}

    def step_device():
        pl.run_device(table, particles=True)

    def step_e2e():
        upload_noise()
        pl.run_device(table, particles=True, stream_to=out)   # ids/pos/vel handed off as each completes (copy stream)
        pl.particles.stream_wait()

    # ---- warm-up
    for _ in range(max(3, args.warmup)):
        step_device()
    M.synchronize()
    barrier()

    # ---- timed region: device-resident inputs, per-role CUDA events on
    lib.mfb_profile_enable(1)
    lib.mfb_profile_reset()
    clocks = ClockSampler(local_rank)
    clocks.start()
    l0c = M.kernel_launches()
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)   # device time, max over ranks
    launches = (M.kernel_launches() - l0c) // args.steps * world
    clk = clocks.stop()

    # ---- per-role table
    import ctypes as C
    roles = []
    for r in range(lib.mfb_profile_num_roles()):
        t, c, b = C.c_double(), C.c_longlong(), C.c_double()
        lib.mfb_profile_get(r, C.byref(t), C.byref(c), C.byref(b))
        if c.value:
            roles.append({"role": lib.mfb_profile_role_name(r).decode(), "launches_per_step": c.value // args.steps,
                          "ms_per_step": t.value / args.steps, "avg_ms": t.value / c.value,
                          "alg_gb_per_launch": b.value / c.value / 1e9,
                          "gbs": (b.value / 1e9) / (t.value / 1e3) if t.value > 0 and b.value > 0 else None})
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
            traffic = ent["dram_bytes_per_launch"]
            traffic_note = {"captured_launch_alg_bytes": ent["alg_bytes_of_captured_launch"], "source": ent["source"]}
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom["role"], "achieved": dom["gbs"], "peak": peak, "unit": "GB/s",
                "frac": dom["gbs"] / peak, "traffic": traffic, "traffic_capture": traffic_note, "peak_source": peak_src,
                "share_of_step": dom["ms_per_step"] / ms, "avg_launch_ms": dom["avg_ms"],
                "alg_bytes_per_launch": dom["alg_gb_per_launch"] * 1e9,
                "frac_of_8000_nominal": dom["gbs"] / 8000.0}
    total_alg = sum(r["alg_gb_per_launch"] * r["launches_per_step"] for r in fftroles)

    # ---- end to end through the host API (host buffers, H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        step_e2e()
        M.synchronize()
        barrier()
        t0 = time.perf_counter()
        ksteps = max(1, min(args.steps, 2))
        for _ in range(ksteps):
            step_e2e()
        M.synchronize()
        barrier()
        dt = max_over_ranks((time.perf_counter() - t0) / ksteps)
        e2e = {"value": n ** 3 / dt / 1e6, "unit": "Mcells/s", "ms_per_step": dt * 1e3,
               "h2d_bytes_per_step": int(hn.nbytes + table.nbytes) * world,
               "d2h_bytes_per_step": int(out["pos"].nbytes + out["vel"].nbytes + out["ids"].nbytes) * world,
               "includes": "pinned noise H2D, r2c FFT + conditioning, pipeline, particle AoS D2H into pinned host arrays "
                           "(ids / positions streamed on a copy stream under the remaining component FFTs; all ranks)"}

    # ---- CPU baseline: bounded sample of the same workload on the host cores
    cpu = None
    if not args.no_cpu and rank == 0 and world == 1:
        threads = os.cpu_count() or 1
        ncpu = args.cpu_grid or 128
        dtc = cpu_reference_run(ncpu, order, threads)
        cpu = {"value": ncpu ** 3 / dtc / 1e6, "unit": "Mcells/s", "cores": threads, "kind": "port",
               "sample": f"full {order}LPT pipeline at {ncpu}^3 (padded {3 * ncpu // 2}^3) = {dtc:.1f} s of oracle "
                         f"(numpy + torch/MKL FP64 FFT, {threads} threads); FFT cost grows as N^3 log N, so the "
                         f"{n}^3 figure would be lower"}

    # the reference does 33 products (each with its own forward transform) at 3LPT, 5 at 2LPT; with the accumulator
    # workspace the library sums the products of one target first
    sinfo = pl.conv.sources_info() if pl.conv is not None else {"path": 0}
    grouped = bool(pl.conv is not None and (sinfo["path"] == 2 or pl.conv.grouped()))
    nprod = {1: 0, 2: 5, 3: 32 if grouped else 33}[order]
    nfwd = {1: 0, 2: 1 if grouped else 5, 3: (10 if sinfo["path"] == 2 else 9) if grouped else 33}[order]
    if dist is not None:
        from monofonic_b200 import dist as MD

        MD.check()
        nex = lib.mfb_dist_exchanges()
        barrier()
        MD.shutdown()
        dist.destroy_process_group()
        if rank != 0:
            return
    line = {"metric": "3LPT IC Mcells/s", "value": n ** 3 / (ms / 1e3) / 1e6, "unit": "Mcells/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n, order), "grid": n, "lpt_order": order,
                       "l2": "inputs larger than L2 (every grid >= 1 GB; no flush needed)",
                       "products_per_step": nprod, "forward_transforms_per_step": nfwd,
                       "product_grouping": grouped,
                       "source_stage": ({"path": "shared x passes + x-plane chunks + multi-product z pass", **sinfo}
                                        if sinfo["path"] == 2 else {"path": "per product", **sinfo}),
                       "decomposition": (f"{world} x-slabs (real) / ky-slabs (k-space, transposed), exchange fused into the "
                                         f"FFT passes over peer memory, {nex // (max(3, args.warmup) + args.steps + 3)} exchanges/step"
                                         if world > 1 else "single slab")},
            "gpu_launches": int(launches), "clocks": clk, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
            "pipeline_alg_gb_per_step": total_alg, "pipeline_gbs": total_alg / (ms / 1e3), "stages": roles}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
