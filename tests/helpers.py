"""Shared test helpers (host-side text <-> arrays, synthetic inputs of the BASELINE shapes)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def read_golden(name, mode='r'):
    with open(os.path.join(GOLDEN, name), mode) as f:
        return f.read()


def tab_to_arrays(lines, names):
    """Columns 1,3,4 of the non-'#' lines of a .tab file -> (scaffold index, start, end) int32 arrays."""
    idx = {n: i for i, n in enumerate(names)}
    c, s, e = [], [], []
    for line in lines:
        if line.startswith('#') or not line.strip():
            continue
        f = line.split()
        c.append(idx[f[0]]); s.append(int(f[2])); e.append(int(f[3]))
    return np.array(c, np.int32), np.array(s, np.int32), np.array(e, np.int32)


def segments_to_gff_rows(seg, names, source, label, prefix):
    c, s, e = seg
    return [f'{names[int(c[i])]}\t{source}\t{label}\t{int(s[i])}\t{int(e[i])}\t.\t+\t.\tID={prefix}_{i + 1:05d}\n'
            for i in range(len(c))]


def synth_hits(seed, nchrom, chrom_size, nhits, hotspots):
    """SURVEY 8(d) config-2 generator: 70 % of hits in `hotspots` hotspots (sigma 500 bp), 30 % uniform;
    length 100 + Exp(600) capped at 20 kbp."""
    rng = np.random.default_rng(seed)
    sizes = np.full(nchrom, chrom_size, dtype=np.int64)
    nh = int(nhits * 0.7)
    hc = rng.integers(0, nchrom, hotspots)
    hp = rng.integers(0, chrom_size, hotspots)
    pick = rng.integers(0, hotspots, nh)
    chrom = np.concatenate([hc[pick], rng.integers(0, nchrom, nhits - nh)]).astype(np.int32)
    start = np.concatenate([hp[pick] + rng.normal(0, 500, nh), rng.integers(0, chrom_size, nhits - nh)])
    start = np.clip(start, 1, chrom_size - 1).astype(np.int32)
    length = np.minimum(100 + rng.exponential(600, nhits), 20000).astype(np.int32)
    end = np.minimum(start + length, chrom_size).astype(np.int32)
    return chrom, start, end, sizes
