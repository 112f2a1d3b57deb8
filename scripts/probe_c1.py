import sys, time, json
sys.path.insert(0, '.')
import numpy as np
from tests.helpers import synth_genome
from mimeo_b200 import _lib, genome as G, align as A
_lib.init()
t0 = time.time()
g = synth_genome(1001, 10, 500_000, 20, copies=(5, 30), fam_len=(300, 3000), sub=0.106, indel=0.005)
print('synth', time.time() - t0)
names = sorted(g)
t0 = time.time()
T = G.Genome(names, [g[n] for n in names])
Trc = T.both_strands()
_lib.sync()
print('upload', time.time() - t0)
for it in range(3):
    _lib.prof_reset(); _lib.prof_enable(True)
    t0 = time.time()
    hits, stats = A.align(T, T, G.align_params(3000), Q_aux=Trc)
    dt = time.time() - t0
    _lib.prof_enable(False)
    print('align secs', dt, 'hits', len(hits['t_id']))
    print({k: _lib.prof_get(k) for k in ('seed_table_build', 'seed_scan', 'surv_sort', 'hsp_extend', 'hsp_sort', 'chain', 'gapped')})
print(json.dumps(stats))
