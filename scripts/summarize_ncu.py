"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.
  python scripts/summarize_ncu.py launches <launches.csv> <out.txt>
  python scripts/summarize_ncu.py full <file.ncu-rep> <out.txt>
"""
import collections
import csv
import re
import subprocess
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "lts__t_sector_hit_rate.pct"]


def launches(src, dst):
    rows = list(csv.DictReader(l for l in open(src) if l.startswith('"')))
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows:
        k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        t = float(r["Metric Value"]) / 1e6
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none ; {len(rows)} launches, {tot:.2f} ms total "
                f"(cold-cache, serialised: compare SHARES)\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:72s} n={c:4d} {t:10.3f} ms {100 * t / tot:5.1f}%  avg {t / c:.4f} ms\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none, source {src}\n")
        for r in rows[2:]:
            f.write("---\n")
            for w, i in idx:
                f.write(f"{w} = {r[i][:110]} {units[i]}\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
