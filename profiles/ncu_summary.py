#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): python profiles/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'smsp__cycles_active.avg', 'sm__inst_executed_pipe_alu.sum', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active']

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H, U = rows[0], rows[1]
ki = H.index('Kernel Name')
for r in rows[2:]:
    print('==', r[ki][:90], 'grid', r[H.index('Grid Size')], 'block', r[H.index('Block Size')])
    for w in WANT:
        hit = [i for i, h in enumerate(H) if h == w or h.endswith('.' + w)]
        if hit:
            print(f'   {w:75s} {r[hit[0]]:>16s} {U[hit[0]]}')
