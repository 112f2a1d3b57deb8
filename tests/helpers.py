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


# ----------------------------------------------------------------------------------------- synthetic genomes
def mutate(rng, seq, sub, indel):
    """Copy of `seq` (uint8 ASCII, ACGT) with per-base substitution prob `sub` and indel prob `indel` (length 1-3,
    half deletions, half insertions of random bases). Vectorised: substitutions first, then indels at sampled sites."""
    bases = np.frombuffer(b'ACGT', dtype=np.uint8)
    out = np.array(seq, dtype=np.uint8, copy=True)
    n = len(out)
    if n == 0:
        return out
    hit = np.flatnonzero(rng.random(n) < sub)
    if len(hit):
        code = np.searchsorted(bases, out[hit])
        out[hit] = bases[(code + rng.integers(1, 4, len(hit))) % 4]
    sites = np.flatnonzero(rng.random(n) < indel)
    if len(sites) == 0:
        return out
    lens = rng.integers(1, 4, len(sites))
    is_del = rng.random(len(sites)) < 0.5
    keep = np.ones(n, dtype=bool)
    for p, ln in zip(sites[is_del], lens[is_del]):
        keep[p:p + ln] = False
    ins_sites = sites[~is_del]
    ins_lens = lens[~is_del]
    if len(ins_sites):
        ins_pos = np.repeat(ins_sites, ins_lens)
        ins_val = bases[rng.integers(0, 4, len(ins_pos))]
        # insert before the site; deleted bases are dropped afterwards by the shifted mask
        out2 = np.insert(out, ins_pos, ins_val)
        keep2 = np.insert(keep, ins_pos, True)
        return out2[keep2]
    return out[keep]


_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b'ACGTN', b'TGCAN'):
    _COMP[_a] = _b


def revcomp_ascii(seq):
    return _COMP[seq[::-1]]


def synth_genome(seed, nscaf, scaf_len, nfam, copies=(5, 30), fam_len=(300, 3000), sub=0.106, indel=0.005, n_runs=0):
    """SURVEY 8(d) config-1 shaped generator: uniform random scaffolds with planted repeat families; each copy is the
    family consensus with per-base substitutions/indels on a random strand at a random non-overlapping place.
    Returns {name: uint8 ASCII array}."""
    rng = np.random.default_rng(seed)
    bases = np.frombuffer(b'ACGT', dtype=np.uint8)
    scafs = [bases[rng.integers(0, 4, scaf_len)].copy() for _ in range(nscaf)]
    used = [np.zeros(scaf_len // 64 + 2, dtype=bool) for _ in range(nscaf)]     # 64-base occupancy map
    for _ in range(nfam):
        L = int(rng.integers(fam_len[0], fam_len[1] + 1))
        cons = bases[rng.integers(0, 4, L)]
        for _c in range(int(rng.integers(copies[0], copies[1] + 1))):
            cp = mutate(rng, cons, sub, indel)
            if rng.random() < 0.5:
                cp = revcomp_ascii(cp)
            for _try in range(50):
                s = int(rng.integers(0, nscaf))
                p = int(rng.integers(0, scaf_len - len(cp)))
                if not used[s][p // 64:(p + len(cp)) // 64 + 1].any():
                    scafs[s][p:p + len(cp)] = cp
                    used[s][p // 64:(p + len(cp)) // 64 + 1] = True
                    break
    for _ in range(n_runs):                       # a few N runs to exercise non-ACGT handling
        s = int(rng.integers(0, nscaf)); p = int(rng.integers(0, scaf_len - 50))
        scafs[s][p:p + int(rng.integers(1, 40))] = ord('N')
    return {f'scaf{i:03d}': scafs[i] for i in range(nscaf)}


def synth_c5(seed, total_bp=1_000_000_000, nscaf=500, repeat_frac=0.40, fam_len=(2000, 10000), copies=(50, 2000),
             identity=(0.75, 0.98), sigma=0.8):
    """SURVEY 8(d) config-5 shaped generator: `nscaf` scaffolds with log-normal lengths summing to `total_bp`, uniform
    random background, repeat families (LTR-like 2-10 kbp, 50-2000 copies, per-family identity 75-98 %) planted until
    `repeat_frac` of the bases are repeats. Returns {name: uint8 ASCII array}."""
    rng = np.random.default_rng(seed)
    bases = np.frombuffer(b'ACGT', dtype=np.uint8)
    w = rng.lognormal(0.0, sigma, nscaf)
    lens = np.maximum((w / w.sum() * total_bp).astype(np.int64), 20_000)
    scafs = [bases[rng.integers(0, 4, int(n))] for n in lens]
    used = [np.zeros(int(n) // 64 + 2, dtype=bool) for n in lens]
    cum = np.cumsum(lens) / lens.sum()
    target = repeat_frac * lens.sum()
    planted = 0
    while planted < target:
        L = int(rng.integers(fam_len[0], fam_len[1] + 1))
        ncopy = int(np.exp(rng.uniform(np.log(copies[0]), np.log(copies[1]))))
        ident = rng.uniform(identity[0], identity[1])
        sub = 1.0 - np.sqrt(ident)                      # two copies each diverged from the consensus
        cons = bases[rng.integers(0, 4, L)]
        for _ in range(ncopy):
            cp = mutate(rng, cons, sub, 0.004)
            if rng.random() < 0.5:
                cp = revcomp_ascii(cp)
            for _try in range(20):
                s = int(np.searchsorted(cum, rng.random()))
                if len(cp) + 64 >= lens[s]:
                    continue
                p = int(rng.integers(0, lens[s] - len(cp)))
                if not used[s][p // 64:(p + len(cp)) // 64 + 1].any():
                    scafs[s][p:p + len(cp)] = cp
                    used[s][p // 64:(p + len(cp)) // 64 + 1] = True
                    planted += len(cp)
                    break
            if planted >= target:
                break
    return {f'scaf{i:04d}': scafs[i] for i in range(nscaf)}


def synth_c4(seed=1004, a_scaf=50, b_scaf=100, scaf_len=1_000_000, nfam=40, fam_len=(300, 3000)):
    """SURVEY 8(d) config-4 generator (`mimeo x`): genome A (50 x 1 Mbp) and genome B (100 x 1 Mbp), uniform random
    background; 40 repeat families, each present 5-30 times in B and 1-3 times in A, every copy diverged from the family
    consensus so that two copies are 80-90 % identical (per-copy substitution rate 5-10.6 %, indel rate 0.5 %), random
    strand, random non-overlapping place. Returns (A, B) as {name: uint8 ASCII array}."""
    rng = np.random.default_rng(seed)
    bases = np.frombuffer(b'ACGT', dtype=np.uint8)

    def blank(n, prefix):
        return [bases[rng.integers(0, 4, scaf_len)] for _ in range(n)], [np.zeros(scaf_len // 64 + 2, dtype=bool) for _ in range(n)], prefix
    ga, gb = blank(a_scaf, 'A'), blank(b_scaf, 'B')

    def plant(g, cp):
        scafs, used, _ = g
        for _try in range(50):
            s = int(rng.integers(0, len(scafs)))
            p = int(rng.integers(0, scaf_len - len(cp)))
            if not used[s][p // 64:(p + len(cp)) // 64 + 1].any():
                scafs[s][p:p + len(cp)] = cp
                used[s][p // 64:(p + len(cp)) // 64 + 1] = True
                return
    for _ in range(nfam):
        cons = bases[rng.integers(0, 4, int(rng.integers(fam_len[0], fam_len[1] + 1)))]
        for g, lo, hi in ((gb, 5, 30), (ga, 1, 3)):
            for _c in range(int(rng.integers(lo, hi + 1))):
                cp = mutate(rng, cons, float(rng.uniform(0.05, 0.106)), 0.005)
                plant(g, revcomp_ascii(cp) if rng.random() < 0.5 else cp)
    return ({f'A{i:03d}': ga[0][i] for i in range(a_scaf)}, {f'B{i:03d}': gb[0][i] for i in range(b_scaf)})


def synth_c3(seed=1003, nscaf=40, scaf_len=1_000_000, block=50_000, keep_frac=0.60, sub=0.08, indel=0.005, inv_frac=0.10):
    """SURVEY 8(d) config-3 generator (`mimeo map`): genome A (40 x 1 Mbp, uniform random) and genome B = A with 8 %
    substitutions and 0.5 % indels over 60 % of its length (blocks of 50 kbp; the other blocks are re-randomised), 10 % of
    the kept blocks inverted. Returns (A, B) as {name: uint8 ASCII array}."""
    rng = np.random.default_rng(seed)
    bases = np.frombuffer(b'ACGT', dtype=np.uint8)
    A, B = {}, {}
    for i in range(nscaf):
        a = bases[rng.integers(0, 4, scaf_len)]
        parts = []
        for p in range(0, scaf_len, block):
            blk = a[p:p + block]
            if rng.random() < keep_frac:
                m = mutate(rng, blk, sub, indel)
                parts.append(revcomp_ascii(m) if rng.random() < inv_frac else m)
            else:
                parts.append(bases[rng.integers(0, 4, len(blk))])
        A[f'A{i:03d}'] = a
        B[f'B{i:03d}'] = np.concatenate(parts)
    return A, B


def odd_genome(rng, k):
    """Small genomes of awkward shapes for parity stress (kind = k % 6): ordinary / tiny scaffolds / tandem array and low
    complexity / N-rich / one dense family / two long near-identical scaffolds."""
    kind = k % 6
    if kind == 0:      # ordinary
        return synth_genome(1000 + k, int(rng.integers(1, 5)), int(rng.integers(3000, 40000)), int(rng.integers(1, 5)), copies=(2, 8),
                            fam_len=(200, 2500), sub=float(rng.uniform(0.02, 0.14)), indel=float(rng.uniform(0, 0.01)), n_runs=int(rng.integers(0, 4)))
    g = {}
    acgt = np.frombuffer(b'ACGT', dtype=np.uint8)
    if kind == 1:      # tiny scaffolds next to a normal one
        g['big'] = acgt[rng.integers(0, 4, 20000)].copy()
        for i, n in enumerate((1, 5, 18, 19, 20, 40, 63, 64, 65)):
            g['t%02d' % i] = g['big'][100 * i:100 * i + n].copy()
    elif kind == 2:    # tandem array + low complexity
        unit = acgt[rng.integers(0, 4, int(rng.integers(150, 400)))]
        arr = np.concatenate([mutate(rng, unit, 0.05, 0.003) for _ in range(int(rng.integers(5, 25)))])
        g['tandem'] = np.concatenate([acgt[rng.integers(0, 4, 3000)], arr, acgt[rng.integers(0, 4, 3000)], np.tile(np.frombuffer(b'AC', dtype=np.uint8), 400)])
        g['other'] = np.concatenate([acgt[rng.integers(0, 4, 2000)], mutate(rng, arr[:3000], 0.1, 0.005), acgt[rng.integers(0, 4, 2000)]])
    elif kind == 3:    # N-rich
        s = acgt[rng.integers(0, 4, 30000)].copy()
        rep = s[2000:4500].copy()
        s[10000:12500] = mutate(rng, rep, 0.08, 0.0)[:2500]
        for _ in range(12):
            p0 = int(rng.integers(0, 29000)); s[p0:p0 + int(rng.integers(1, 300))] = ord('N')
        g['nrich'] = s
        g['nother'] = np.concatenate([mutate(rng, rep, 0.06, 0.004), np.full(500, ord('N'), np.uint8), mutate(rng, rep, 0.12, 0.004)])
    elif kind == 4:    # one dense family: many HSPs per tile (batched chain, long rings)
        return synth_genome(2000 + k, 2, 60000, 2, copies=(20, 35), fam_len=(300, 1200), sub=0.05, indel=0.004)
    else:              # near-identical long scaffolds (long gapped extensions, wide payload when > 65536 columns)
        a = acgt[rng.integers(0, 4, int(rng.integers(70000, 90000)))].copy()
        g['a'] = a
        g['b'] = mutate(rng, a, 0.01, 0.0005)
    return g


def py_tab_blocks(hits, tnames, qnames, minLen, minIdt):
    """Plain-Python statement of the filter that follows every LASTZ call (wrappers.py:1044-1056), the checker of the
    native formatter: keep length1 >= minLen and printed identity >= minIdt, 10 columns, per (t, q) block sorted by
    start1 numerically, then by the whole line."""
    out = {}
    for k in range(len(hits['t_id'])):
        s1, e1 = int(hits['start1'][k]), int(hits['end1'][k])
        if e1 - s1 + 1 < minLen:
            continue
        nm, nc = int(hits['nmatch'][k]), int(hits['ncols'][k])
        pct = '%.1f' % (100.0 * nm / nc) if nc else '0.0'
        if float(pct) < float(minIdt):
            continue
        t, q = int(hits['t_id'][k]), int(hits['q_id'][k])
        row = '\t'.join((tnames[t], '+', str(s1), str(e1), qnames[q], '-' if hits['strand'][k] else '+',
                         str(int(hits['start2'][k])), str(int(hits['end2'][k])), str(int(hits['score'][k])), pct))
        out.setdefault((t, q), []).append(row)
    for key, rows in out.items():
        rows.sort(key=lambda r: (float(r.split('\t')[2]), r.encode()))
        out[key] = [r + '\n' for r in rows]
    return out
