# Randomised parity stress: many small genomes of odd shapes (tiny and empty-ish scaffolds, N runs, low-complexity stretches,
# tandem copies, dense families) through the whole GPU alignment stage, every row compared with the CPU oracle.
import sys, time
sys.path.insert(0, '.')
import numpy as np
from tests.helpers import odd_genome
from oracle import lastz_oracle as lo
from mimeo_b200 import _lib, genome as G, align as A

_lib.init()


def oracle_rows(g, p):
    rows = set()
    names = list(g)
    for ti, t in enumerate(names):
        tix = lo.TargetIndex(lo.encode(g[t]))
        for qi, qn in enumerate(names):
            qc = lo.encode(g[qn]); m = len(qc)
            for st in (0, 1):
                q = qc if st == 0 else lo.revcomp_codes(qc)
                for (s1, e1, s2, e2, sc, nm, nc, _a, _b) in lo.align_tile(tix, q, p).tolist():
                    qs, qe = (s2 + 1, e2) if st == 0 else (m - e2 + 1, m - s2)
                    rows.add((ti, qi, st, s1 + 1, e1, qs, qe, sc, nm, nc))
    return rows


n = int(sys.argv[1]) if len(sys.argv) > 1 else 18
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
bad = 0
t0 = time.time()
for k in range(n):
    g = odd_genome(rng, k)
    hsp = int(rng.choice([3000, 2200, 5000]))
    T = G.Genome.from_dict(g)
    hits, stats = A.align(T, T, G.align_params(hsp))
    got = set(zip(*[hits[f].tolist() for f in A.HIT_FIELDS]))
    want = oracle_rows(g, lo.default_params(hsp))
    ok = got == want
    bad += 0 if ok else 1
    print('case %2d kind %d scaffolds %2d bases %7d hspthresh %d rows %4d %s' % (k, k % 6, len(g), sum(len(v) for v in g.values()), hsp, len(want), 'ok' if ok else 'MISMATCH (gpu-only %d, oracle-only %d)' % (len(got - want), len(want - got))), flush=True)
    T.close()
print('done: %d cases, %d mismatches, %.0f s' % (n, bad, time.time() - t0))
sys.exit(1 if bad else 0)
