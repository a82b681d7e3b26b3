"""print the handful of ncu metrics that decide what bounds an FFT pass kernel:  python scripts/ncu_key.py <file.ncu-rep> [launch index]"""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum"]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    sel = rows[2:] if len(sys.argv) < 3 else [rows[2 + int(sys.argv[2])]]
    for r in sel:
        print("---")
        for k in KEYS:
            if k in hdr:
                print(f"{k} = {r[hdr.index(k)][:100]} {units[hdr.index(k)]}")
        st = [(float(r[i]), h[len(STALL):].replace("_per_issue_active.ratio", "")) for i, h in enumerate(hdr)
              if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i]]
        print("stalls per issue:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:9]))


if __name__ == "__main__":
    main()
