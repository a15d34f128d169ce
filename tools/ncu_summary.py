#!/usr/bin/env python3
"""Summarises an .ncu-rep (ncu --set full) into the handful of metrics profiles/ keeps.
usage: tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread ",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg ", "smsp__cycles_active.avg ",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg ",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum ", "dram__bytes_write.sum ", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum ", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active", "smsp__average_warps_issue_stalled_wait_per_issue_active",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active",
    "smsp__inst_executed.sum ", "sm__inst_executed_pipe_xu.sum ",
]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
print(f"# {rep}: ncu --set full --clock-control none; one row per captured launch")
for vals in rows[2:]:
    d = dict(zip(hdr, zip(units, vals)))
    print("kernel:", d.get("Kernel Name", ("", "?"))[1][:100])
    for h in hdr:
        if any(k.strip() in h and ((h + " ").startswith(k) or h.endswith(k.strip()) or k.strip() in h.split(".TriageCompute.")[-1][:len(k.strip())]) for k in KEYS):
            u, v = d[h]
            print(f"  {h:95s} {v:>18s} {u}")
