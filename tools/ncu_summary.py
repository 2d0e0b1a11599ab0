#!/usr/bin/env python
"""Print the handful of ncu metrics we track from a .ncu-rep (read on the CPU box)."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio" ]
STALL = "smsp__average_warps_issue_stalled_"


def main(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        if pattern and pattern not in r[ki]:
            continue
        print("==", r[ki][:70])
        for k in KEYS:
            if k in hdr:
                print(f"   {k:70s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
        stalls = [(float(r[i] or 0), h[len(STALL):]) for i, h in enumerate(hdr)
                  if h.startswith(STALL) and h.endswith("_per_warp_active.pct")]
        for v, h in sorted(stalls, reverse=True)[:7]:
            print(f"   stall {h:64s} {v:10.2f}")


if __name__ == "__main__":
    main(*sys.argv[1:])
