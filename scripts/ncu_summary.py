"""Summarise one kernel of an .ncu-rep capture (run here, no GPU): python scripts/ncu_summary.py REP 'title' [row] > profiles/X.txt"""
import csv
import subprocess
import sys

rep, title = sys.argv[1], sys.argv[2]
row = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
units = dict(zip(hdr, rows[1]))
d = dict(zip(hdr, rows[2 + row]))
keys = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_active.avg', 'sm__cycles_elapsed.max',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum', 'smsp__sass_inst_executed_op_shared_ld.sum', 'smsp__sass_inst_executed_op_global_ld.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed']
print(title)
print('kernel:', d.get('Kernel Name', '')[:150])
for k in keys:
    if k in d:
        print('  %-75s %s %s' % (k, d[k], units.get(k, '')))
for k in d:
    if 'stalled' in k and 'per_issue_active.ratio' in k and float(d[k] or 0) > 0.1:
        print('  stall %-69s %s' % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), d[k]))
