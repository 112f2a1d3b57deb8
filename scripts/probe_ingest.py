# Throughput of the native text ingest (mb2_tab_project / mb2_fasta_read) beside the parsers it replaced
# (pandas C parser for the .tab projection, numpy masking for FASTA). Host-only: runs without a GPU.
import os, sys, time, tempfile
sys.path.insert(0, '.')
import numpy as np
import pandas as pd
from mimeo_b200 import engine, fasta

nrows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
mbp = float(sys.argv[2]) if len(sys.argv) > 2 else 200.0
d = tempfile.mkdtemp()
rng = np.random.default_rng(5)
# ---- .tab
t0 = time.time()
s = rng.integers(0, 2_000_000, nrows)
df = pd.DataFrame({'n1': np.char.add('scaf_', rng.integers(0, 50, nrows).astype(str)), 's1': '+', 'a': s, 'b': s + rng.integers(100, 20000, nrows),
                   'n2': 'scaf_q', 's2': '-', 'c': s, 'd': s + 500, 'score': 60 * (s % 5000 + 100), 'id': np.round(rng.uniform(60, 100, nrows), 1)})
tab = os.path.join(d, 'hits.tab')
with open(tab, 'w') as f:
    f.write('#name1\tstrand1\tstart1\tend1\tname2\tstrand2\tstart2+\tend2+\tscore\tidentity\n')
    df.to_csv(f, sep='\t', header=False, index=False)
mb = os.path.getsize(tab) / 1e6
print('tab: %d rows, %.0f MB written in %.1f s' % (nrows, mb, time.time() - t0))
for nt in (1, 0):
    t0 = time.perf_counter(); names, ids, a, b = engine.parse_tab_hits(tab, nthreads=nt); dt = time.perf_counter() - t0
    print('  native mb2_tab_project threads=%s: %.3f s  %.0f MB/s  %.1f M rows/s' % (nt or 'all', dt, mb / dt, nrows / dt / 1e6))
t0 = time.perf_counter()
pdf = pd.read_csv(tab, sep='\t', comment='#', header=None, usecols=[0, 2, 3], dtype={0: str, 2: np.int64, 3: np.int64}, engine='c')
dt = time.perf_counter() - t0
print('  pandas read_csv (previous path):      %.3f s  %.0f MB/s' % (dt, mb / dt))
assert np.array_equal(pdf[2].to_numpy(), a) and np.array_equal(pdf[3].to_numpy(), b) and [names[i] for i in ids[:1000]] == pdf[0][:1000].tolist()
# ---- FASTA
n = int(mbp * 1e6)
seq = np.frombuffer(b'ACGT', dtype=np.uint8)[rng.integers(0, 4, n)]
fa = os.path.join(d, 'g.fa')
t0 = time.time()
nscaf = 20
with open(fa, 'wb') as f:
    for k in range(nscaf):
        part = seq[k * (n // nscaf):(k + 1) * (n // nscaf)]
        f.write(b'>scaf_%d synthetic\n' % k)
        full = len(part) // 60
        block = np.empty((full, 61), dtype=np.uint8); block[:, :60] = part[:full * 60].reshape(full, 60); block[:, 60] = 10
        f.write(block.tobytes()); f.write(part[full * 60:].tobytes() + b'\n')
mbf = os.path.getsize(fa) / 1e6
print('fasta: %.0f Mbp, %.0f MB written in %.1f s' % (mbp, mbf, time.time() - t0))
for nt in (1, 0):
    t0 = time.perf_counter(); recs = fasta.read_fasta(fa, nthreads=nt); dt = time.perf_counter() - t0
    print('  native mb2_fasta_read threads=%s: %.3f s  %.0f MB/s' % (nt or 'all', dt, mbf / dt))
t0 = time.perf_counter()
data = open(fa, 'rb').read(); buf = np.frombuffer(data, dtype=np.uint8)
gt = np.flatnonzero(buf == ord('>')); starts = [int(p) for p in gt if p == 0 or buf[p - 1] == 10]
old = []
for k, st in enumerate(starts):
    e = starts[k + 1] if k + 1 < len(starts) else len(buf)
    nl = data.find(b'\n', st, e); body = buf[nl + 1:e]
    old.append(np.ascontiguousarray(body[(body != 10) & (body != 13) & (body != 32)]))
dt = time.perf_counter() - t0
print('  numpy masking (previous path):        %.3f s  %.0f MB/s' % (dt, mbf / dt))
assert all(np.array_equal(o, r[2]) for o, r in zip(old, recs)) and len(old) == len(recs)
print('cores', os.cpu_count())
