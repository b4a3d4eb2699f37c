#!/usr/bin/env python3
"""Condense an .ncu-rep into the handful of numbers DESIGN.md / bench.py cite.  usage: ncu_summary.py rep.ncu-rep [frames_per_launch]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
frames = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
col = {h: i for i, h in enumerate(hdr)}
for r in data:
    print("=" * 100)
    vals = {}
    for k in KEYS:
        if k in col:
            vals[k] = r[col[k]]
            print(f"{k:95s} {r[col[k]]} {units[col[k]]}")
    if frames:
        inst = float(vals["smsp__inst_executed.sum"])
        print(f"{'derived: warp instructions per frame':95s} {inst / frames:.0f}")
        rd = float(vals["dram__bytes_read.sum"]) * (1e6 if units[col['dram__bytes_read.sum']] == 'Mbyte' else 1e9 if units[col['dram__bytes_read.sum']] == 'Gbyte' else 1)
        wr = float(vals["dram__bytes_write.sum"]) * (1e6 if units[col['dram__bytes_write.sum']] == 'Mbyte' else 1e9 if units[col['dram__bytes_write.sum']] == 'Gbyte' else 1)
        print(f"{'derived: DRAM bytes per frame (read+write)':95s} {(rd + wr) / frames:.0f}")
        print(f"{'derived: DRAM bytes per launch (read+write)':95s} {rd + wr:.0f}")
