#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
python profiles/launch_list_summary.py gpurun_out/launches.csv "<comment line>" > profiles/x_launch_list_summary.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors='replace')) if len(r) > 10]
H = rows[0]
ki, vi, ui = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
acc = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[ki] == 'Kernel Name':
        continue
    name = re.sub(r'\(.*$', '', r[ki]).replace(',', ';')
    v = float(r[vi].replace(',', ''))
    us = v / 1e3 if r[ui] in ('ns', 'nsecond') else v
    acc[name][0] += 1
    acc[name][1] += us
tot = sum(v[1] for v in acc.values())
print('# ' + (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print('kernel,launches,total_us,share')
for k, (n, us) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f'{k},{n},{us:.1f},{us / tot:.4f}')
