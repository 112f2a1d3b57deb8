import sys, time
sys.path.insert(0, '.')
import numpy as np
import bench
from mimeo_b200 import _lib, parallel, engine, align as A, coverage
from mimeo_b200.genome import Genome, align_params
_lib.init()
wl = bench.ShardedSelfWorkload(0)
mine = parallel.partition_targets(wl.sizes, 2)[1]
Q = Genome(wl.names, wl.seqs); Qb = Q.both_strands()
T = Genome([wl.names[i] for i in mine], [wl.seqs[i] for i in mine])
t0 = time.time()
hits, stats = A.align(T, Q, align_params(3000), Q_aux=Qb)
print('align secs', time.time() - t0, stats, flush=True)
hits['t_id'] = np.asarray(mine, dtype=np.int32)[hits['t_id']]
keep = engine.filter_hits(hits, 100, 80)
intra = (hits['t_id'] == hits['q_id']) & keep
inter = keep & ~intra
np.savez('gpurun_out/shard_hits.npz', t_id=hits['t_id'][inter], start=hits['start1'][inter], end=hits['end1'][inter], sizes=np.array(wl.sizes))
print('saved', inter.sum(), flush=True)
c, s, e = coverage.coverage_segments(hits['t_id'][inter], hits['start1'][inter], hits['end1'][inter], wl.sizes, 3, 100)
print('coverage ok', len(c))
