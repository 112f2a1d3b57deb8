# single trivial self-diagonal: isolates the sequential critical path of the hsp and gapped stages
import sys, time
sys.path.insert(0, '.')
import numpy as np
from mimeo_b200 import _lib, genome as G, align as A
_lib.init()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
rng = np.random.default_rng(5)
seq = np.frombuffer(b'ACGT', dtype=np.uint8)[rng.integers(0, 4, n)].copy()
if len(sys.argv) > 2:
    seq[n // 2] = ord('N')          # one N: the closed-form self-diagonal shortcut does not apply, the general DP runs
T = G.Genome(['s'], [seq])
for it in range(2):
    _lib.prof_reset(); _lib.prof_enable(True)
    t0 = time.time()
    hits, stats = A.align(T, T, G.align_params(3000), strands=1)
    dt = time.time() - t0
    _lib.prof_enable(False)
    print('secs', dt, {k: _lib.prof_get(k)[0] for k in ('seed_scan', 'hsp_extend', 'chain', 'gapped')}, 'gapped_cells', stats['gapped_cells'], 'hits', len(hits['t_id']))
