#!/usr/bin/env python
"""Summarise ncu --set full reports (gpurun_out/<tag>_<kernel>.ncu-rep) into a markdown table.  Usage: tools/ncu_summary.py <tag> kernel..."""
import csv, io, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio']
tag, kernels = sys.argv[1], sys.argv[2:]
cols = {}
for k in kernels:
    import os
    raw = f'gpurun_out/{tag}_{k}_raw.csv'
    if os.path.exists(raw):
        out = open(raw).read()
    else:
        out = subprocess.run(['ncu', '-i', f'gpurun_out/{tag}_{k}.ncu-rep', '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cols[k] = {h: (rows[1][i], rows[2][i]) for i, h in enumerate(rows[0])}
print("| metric | " + " | ".join(kernels) + " |")
print("|---|" + "---|" * len(kernels))
for w in WANT:
    unit = next((cols[k][w][0] for k in kernels if w in cols[k]), "")
    print(f"| {w} ({unit}) | " + " | ".join(cols[k].get(w, ("", "n/a"))[1] for k in kernels) + " |")
