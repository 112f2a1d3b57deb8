import sys
sys.path.insert(0, '.')
import numpy as np
from mimeo_b200 import coverage
from oracle import annot_oracle as ao
rng = np.random.default_rng(3)
sizes = np.full(50, 2_000_000, dtype=np.int64)
for n in (5000, 3, 1, 70000):
    chrom = (rng.integers(0, 25, n) * 2 + 1).astype(np.int32)
    start = rng.integers(1, 1_990_000, n).astype(np.int32)
    end = (start + rng.integers(100, 9000, n)).astype(np.int32)
    got = coverage.coverage_segments(chrom, start, end, sizes, 1, 100)
    want = ao.coverage_segments_arrays(chrom, start, end, sizes, 1, 100)
    print(n, len(got[0]), all((g == w).all() for g, w in zip(got, want)))
