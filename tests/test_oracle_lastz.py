"""CPU tests of the LASTZ-restatement oracle (PARITY UNPINNED against a real LASTZ: none exists here).
Each stage is checked against an independent brute-force statement of the same published algorithm."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import lastz_oracle as lo
from oracle import annot_oracle as ao
from tests.helpers import synth_genome

SUB = np.array([[91, -114, -31, -123, -100], [-114, 100, -125, -31, -100], [-31, -125, 100, -114, -100],
                [-123, -31, -114, 91, -100], [-100] * 5])
CARE = [0, 1, 2, 4, 7, 8, 11, 13, 15, 16, 17, 18]


def rand_codes(rng, n):
    return rng.integers(0, 4, n).astype(np.uint8)


def test_seed_pattern_and_transition_rule():
    rng = np.random.default_rng(1)
    t = rand_codes(rng, 60)
    q = t.copy()
    L = lo.lib()
    at = lambda tt, qq, i, j: L.lzo_seed_at(tt.ctypes.data, len(tt), qq.ctypes.data, len(qq), i, j, 1)
    assert at(t, q, 5, 5) == 1
    for c in range(19):                       # any single change at a don't-care position keeps the hit
        q2 = q.copy(); q2[5 + c] = (q2[5 + c] + 1) % 4
        assert at(t, q2, 5, 5) == (0 if c in CARE else 1)
        q3 = q.copy(); q3[5 + c] ^= 2         # transition at a care position is tolerated once
        assert at(t, q3, 5, 5) == 1
    q4 = q.copy(); q4[5] ^= 2; q4[6] ^= 2     # two transitions: no hit
    assert at(t, q4, 5, 5) == 0
    q5 = q.copy(); q5[5 + 3] = 4              # N anywhere in the window kills the seed
    assert at(t, q5, 5, 5) == 0
    assert at(t, q, 42, 42) == 0              # window runs off the end (42+19 > 60)


def test_entropy_fixed_point_tracks_shannon():
    L = lo.lib()
    rng = np.random.default_rng(2)
    for _ in range(200):
        cnt = rng.integers(0, 2000, 4).astype(np.uint32)
        if cnt.sum() == 0:
            continue
        h = L.lzo_entropy_q24(cnt.ctypes.data) / float(1 << 24)
        p = cnt[cnt > 0] / cnt.sum()
        assert abs(h - float(-(p * np.log(p)).sum() / math.log(4))) < 1e-5
    one = np.array([50, 0, 0, 0], np.uint32)
    assert L.lzo_entropy_q24(one.ctypes.data) == 0
    flat = np.array([64, 64, 64, 64], np.uint32)
    assert L.lzo_entropy_q24(flat.ctypes.data) == 1 << 24


def brute_xdrop(t, q, i, j, X=910):
    n, m = len(t), len(q)
    run = best = 0; be = i + 19; c1, c2 = i + 19, j + 19
    while c1 < n and c2 < m:
        run += SUB[t[c1], q[c2]]; c1 += 1; c2 += 1
        if run > best: best, be = run, c1
        elif run < best - X: break
    runl = bestl = 0; bs = i + 19; c1, c2 = i + 18, j + 18
    while c1 >= 0 and c2 >= 0:
        runl += SUB[t[c1], q[c2]]
        if runl > bestl: bestl, bs = runl, c1
        elif runl < bestl - X: break
        c1 -= 1; c2 -= 1
    return bs, be, best + bestl


def test_hsps_against_bruteforce_enumeration():
    """Small tile: enumerate every (i,j), apply D1/D2 and the x-drop rule in pure Python, compare HSP sets."""
    rng = np.random.default_rng(3)
    t = rand_codes(rng, 3000)
    q = rand_codes(rng, 2500)
    q[400:1300] = t[1000:1900]                       # one planted ungapped homology ...
    flip = rng.random(900) < 0.12
    q[400:1300][flip] = (q[400:1300][flip] + rng.integers(1, 4, flip.sum())) % 4   # ... at ~88 % identity
    q[1800:2000] = 4                                  # an N run
    p = lo.default_params(3000, entropy=0)
    got = lo.hsps(lo.TargetIndex(t), q, p)
    L = lo.lib()
    want = []
    n, m = len(t), len(q)
    for d in range(-(m - 19), n - 19 + 1):
        covered = -1
        for j in range(max(0, -d), min(m - 19, n - 19 - d) + 1):
            i = j + d
            if not L.lzo_seed_at(t.ctypes.data, n, q.ctypes.data, m, i, j, 1): continue
            if L.lzo_seed_at(t.ctypes.data, n, q.ctypes.data, m, i - 1, j - 1, 1): continue
            if i + 19 <= covered: continue
            bs, be, sc = brute_xdrop(t, q, i, j)
            if sc >= 3000:
                want.append((bs, bs - d, be - bs, sc)); covered = max(covered, be)
    assert len(want) >= 1
    assert sorted(map(tuple, got.tolist())) == sorted(want)


def brute_chain(h):
    h = sorted(map(tuple, h))
    n = len(h)
    Cs = [0] * n; pred = [-1] * n
    for b in range(n):
        best, bi = 0, -1
        for a in range(n):
            if h[a][0] + h[a][2] <= h[b][0] and h[a][1] + h[a][2] <= h[b][1]:
                if Cs[a] > best or (Cs[a] == best and bi >= 0 and a < bi): best, bi = Cs[a], a
        Cs[b] = h[b][3] + best; pred[b] = bi
    end = max(range(n), key=lambda k: (Cs[k], -k))
    out = []
    while end >= 0:
        out.append(h[end]); end = pred[end]
    return sorted(out)


@pytest.mark.parametrize('seed', range(8))
def test_chain_matches_quadratic_dp(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 120))
    span = int(rng.choice([200, 2000, 20000]))
    h = np.stack([rng.integers(0, span, n), rng.integers(0, span, n), rng.integers(20, 300, n),
                  rng.choice([3000, 3500, 5000, 9000], n)], axis=1).astype(np.int32)
    if seed % 2:
        h[:, 3] = 3000                                # all ties: exercises the index tie-break
    got = sorted(map(tuple, lo.chain(h).tolist()))
    assert got == brute_chain(h.tolist())


def brute_gapped(t, q, O=400, E=30):
    """Unpruned affine extension from (0,0): best H over all cells, ties -> smallest anti-diagonal then row."""
    n, m = len(t), len(q)
    NEG = -10 ** 9
    H = np.full((n + 1, m + 1), NEG); D = np.full((n + 1, m + 1), NEG); I = np.full((n + 1, m + 1), NEG)
    H[0, 0] = 0
    for k in range(1, n + m + 1):
        for i in range(max(0, k - m), min(n, k) + 1):
            j = k - i
            if i > 0 and H[i - 1, j] > NEG: D[i, j] = max(H[i - 1, j] - O - E, D[i - 1, j] - E)
            if j > 0 and H[i, j - 1] > NEG: I[i, j] = max(H[i, j - 1] - O - E, I[i, j - 1] - E)
            mv = H[i - 1, j - 1] + SUB[t[i - 1], q[j - 1]] if i > 0 and j > 0 and H[i - 1, j - 1] > NEG else NEG
            H[i, j] = max(mv, D[i, j], I[i, j])
    best = (0, 0, 0)
    for k in range(0, n + m + 1):
        for i in range(max(0, k - m), min(n, k) + 1):
            if H[i, k - i] > best[0]: best = (int(H[i, k - i]), i, k - i)
    return best


def test_gapped_extension_equals_unpruned_dp_when_ydrop_is_inactive():
    """With sequences shorter than the y-drop horizon nothing is pruned, so the oracle must equal a full DP."""
    rng = np.random.default_rng(5)
    cons = rand_codes(rng, 70)
    t = np.concatenate([cons[:30], cons[33:]])                     # 3-base deletion in t
    q = cons.copy(); q[10] = (q[10] + 1) % 4; q[50] ^= 2
    h = np.array([[0, 0, 25, 5000]], dtype=np.int32)               # pretend HSP; anchor = its midpoint
    p = lo.default_params(0)
    out = np.zeros((4, 9), np.int32)
    st = lo.Stats()
    tt = np.ascontiguousarray(t); qq = np.ascontiguousarray(q)
    n = lo.lib().lzo_gapped(tt.ctypes.data, len(tt), qq.ctypes.data, len(qq), h.ctypes.data, 1, C.byref(p), out.ctypes.data, 4, C.byref(st))
    assert n == 1
    s1, e1, s2, e2, score, nm, nc, a1, a2 = out[0].tolist()
    assert (a1, a2) == (12, 12)
    f = brute_gapped(t[a1:], q[a2:])
    b = brute_gapped(t[:a1][::-1], q[:a2][::-1])
    assert score == f[0] + b[0]
    assert (s1, e1, s2, e2) == (a1 - b[1], a1 + f[1], a2 - b[2], a2 + f[2])
    assert (e1 - s1, e2 - s2) == (len(t), len(q)) and nc == 67 and nm == 65


def test_pipeline_recovers_planted_repeats_and_is_symmetric():
    g = synth_genome(7, 3, 40_000, 2, copies=(4, 4), fam_len=(900, 1200), sub=0.05, indel=0.004)
    enc = {k: lo.encode(v) for k, v in g.items()}
    tab, intra, gff = lo.mimeo_self(enc, minIdt=80, minLen=100, minCov=2, intraCov=2, strictSelf=True)
    rows = [l.split('\t') for l in tab.splitlines()[1:]]
    assert len(rows) >= 4
    fwd = {(r[0], r[2], r[3], r[4], r[5], r[6], r[7], r[8]) for r in rows}
    for r in rows:                                   # every A-vs-B hit has its mirror B-vs-A hit with the same score
        assert (r[4], r[6], r[7], r[0], r[5], r[2], r[3], r[8]) in fwd
    # each scaffold aligned to itself yields the trivial full-length alignment in the intra table
    for name, seq in g.items():
        assert f'{name}\t+\t1\t{len(seq)}\t{name}\t+\t1\t{len(seq)}\t' in intra
    assert gff.startswith(ao.GFF_HEADER_SELF)


def test_minus_strand_coordinates():
    rng = np.random.default_rng(11)
    t = rand_codes(rng, 5000)
    q = rand_codes(rng, 4000)
    q[1000:1600] = lo.revcomp_codes(t[2000:2600])
    p = lo.default_params(3000)
    lines = lo.lastz_general('T', lo.TargetIndex(t), 'Q', q, p)
    rows = [l.split('\t') for l in lines if not l.startswith('#')]
    assert len(rows) == 1
    r = rows[0]
    assert r[6] == '-' and abs(int(r[2]) - 2001) <= 25 and abs(int(r[3]) - 2600) <= 25
    assert abs(int(r[7]) - 1001) <= 25 and abs(int(r[8]) - 1600) <= 25 and float(r[12].strip().rstrip('%')) >= 99.0


def test_soft_masked_windows_are_never_seeds_but_score_by_base():
    """A lower-case base makes every 19-mer window that holds it a non-seed on either sequence; extensions score it as its base."""
    rng = np.random.default_rng(9)
    seq = rng.integers(0, 4, 400).astype(np.uint8)
    soft = seq.copy()
    soft[100:130] |= 8
    l = lo.lib()
    for i in range(60, 150):
        clean = not (i <= 129 and i + 19 > 100)
        assert bool(l.lzo_seed_at(soft.ctypes.data, 400, seq.ctypes.data, 400, i, i, 1)) == clean
        assert bool(l.lzo_seed_at(seq.ctypes.data, 400, soft.ctypes.data, 400, i, i, 1)) == clean
    p = lo.default_params(3000)
    a = lo.align_tile(lo.TargetIndex(seq), seq, p)
    b = lo.align_tile(lo.TargetIndex(soft), seq, p)
    assert a.tolist() == b.tolist() and len(a) == 1 and a[0][5] == 400     # seeded elsewhere, extended through the masked stretch: 400 matches
    assert lo.revcomp_codes(lo.revcomp_codes(soft)).tolist() == soft.tolist()
    assert lo.encode('acgtnACGTN').tolist() == [8, 9, 10, 11, 12, 0, 1, 2, 3, 4]
