# C5-shaped genome at a chosen size: stage timings and counters (no oracle; properties only)
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
from tests.helpers import synth_genome
from mimeo_b200 import _lib, genome as G, align as A, engine
_lib.init()
nscaf = int(sys.argv[1]); scaf_len = int(sys.argv[2]); nfam = int(sys.argv[3]); cmin = int(sys.argv[4]); cmax = int(sys.argv[5])
t0 = time.time()
g = synth_genome(1005, nscaf, scaf_len, nfam, copies=(cmin, cmax), fam_len=(2000, 10000), sub=0.08, indel=0.004)
print('synth', round(time.time() - t0, 1), 's; genome Mbp', nscaf * scaf_len / 1e6, flush=True)
names = sorted(g)
T = G.Genome(names, [g[n] for n in names])
Tb = T.both_strands()
_lib.sync()
for it in range(2):
    _lib.prof_reset(); _lib.prof_enable(True)
    t0 = time.time()
    inter, intra, hits, stats = engine.self_segments(T, Tb, [len(g[n]) for n in names], 80, 100, 3, 4, 3000, True)
    dt = time.time() - t0
    _lib.prof_enable(False)
    print('secs', round(dt, 3), 'Mbp/s', round(nscaf * scaf_len / 1e6 / dt, 2), 'hits', len(hits['t_id']), 'segments', len(inter[0]), len(intra[0]))
    print({k: round(_lib.prof_get(k)[0], 2) for k in ('seed_table_build', 'seed_scan', 'surv_sort', 'hsp_extend', 'hsp_sort', 'chain', 'gapped')})
print(json.dumps(stats))
