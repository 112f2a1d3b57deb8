"""GPU parity: coverage -> threshold -> merged segments through the C ABI, bit-exact against the oracle,
the reference-generated goldens and the SURVEY 9.3 known answers."""
import json

import numpy as np
import pytest

from tests.helpers import read_golden
from oracle import annot_oracle as ao
from tests.helpers import tab_to_arrays, segments_to_gff_rows, synth_hits

pytestmark = pytest.mark.gpu
MAN = json.loads(read_golden('manifest.json'))


@pytest.fixture(scope='module')
def cov():
    from mimeo_b200 import coverage
    return coverage


def run(cov, chrom, start, end, sizes, c, ml):
    got = cov.coverage_segments(chrom, start, end, sizes, c, ml)
    want = ao.coverage_segments_arrays(np.asarray(chrom), np.asarray(start), np.asarray(end), sizes, max(c, 1), ml)
    for g, w in zip(got, want):
        assert g.dtype == np.int32 and (g == w).all(), (got, want)
    return got


def test_known_answers(cov):
    c, s, e = run(cov, [0, 0, 0], [100, 200, 250], [300, 400, 500], [1000], 2, 100)
    assert (c.tolist(), s.tolist(), e.tolist()) == ([0], [200], [400])                       # KAT 1
    c, s, e = run(cov, [0] * 4, [0, 0, 150, 150], [150, 150, 300, 300], [1000], 2, 1)
    assert (s.tolist(), e.tolist()) == ([0], [300])                                            # KAT 2 book-ended
    c, s, e = run(cov, [0] * 3, [200] * 3, [400] * 3, [250], 3, 50)
    assert (s.tolist(), e.tolist()) == ([200], [250])                                          # KAT 3 clipped
    assert len(run(cov, [0, 0], [10, 10], [110, 110], [500], 2, 100)[0]) == 1                  # KAT 4 inclusive
    assert len(run(cov, [0, 0], [10, 10], [110, 110], [500], 2, 101)[0]) == 0


def test_edge_cases(cov):
    z = np.zeros(0, np.int32)
    assert len(cov.coverage_segments(z, z, z, [100], 1, 1)[0]) == 0                           # empty input
    run(cov, [0], [10], [10], [100], 1, 0)                                                     # zero-length hit invisible
    run(cov, [0], [0], [0], [100], 1, 0)                                                       # (0,0) quirk of the restated sweep
    run(cov, [0], [100], [120], [100], 1, 0)                                                   # start beyond scaffold end
    run(cov, [0, 1], [90, 0], [500, 10], [100, 50], 1, 1)                                      # run at scaffold end + next scaffold start never join
    run(cov, [1, 1, 0], [0, 0, 99], [50, 50, 100], [100, 50], 2, 1)
    run(cov, [0], [5], [9], [8192 * 3 + 1], 1, 1)
    run(cov, [0, 0], [8191, 8190], [8193, 16385], [8192 * 3], 1, 1)                           # runs crossing tile borders
    run(cov, [0], [3], [7], [10], 0, 0)                                                        # cov <= 0 behaves as 1 (rows of depth 0 never exist)
    run(cov, [0], [3], [7], [10], -3, -5)


def test_invalid_hits_raise(cov):
    from mimeo_b200._lib import Mb2Error
    for bad in ([[0], [5], [3]], [[2], [1], [3]], [[0], [-1], [3]], [[-1], [1], [3]]):
        with pytest.raises(Mb2Error):
            cov.coverage_segments(bad[0], bad[1], bad[2], [100, 100], 1, 1)


@pytest.mark.parametrize('seed', range(12))
def test_random_small(cov, seed):
    rng = np.random.default_rng(seed)
    nchrom = int(rng.integers(1, 7))
    sizes = rng.integers(1, 40000, nchrom).astype(np.int64)
    n = int(rng.integers(1, 3000))
    chrom = rng.integers(0, nchrom, n).astype(np.int32)
    start = (rng.random(n) * (sizes[chrom] + 30)).astype(np.int32)
    end = (start + rng.integers(0, 2000, n)).astype(np.int32)
    run(cov, chrom, start, end, sizes, int(rng.integers(0, 6)), int(rng.integers(0, 300)))


def test_dense_hotspots_many_events_per_tile(cov):
    rng = np.random.default_rng(99)
    n = 200000
    start = (5000 + rng.normal(0, 300, n)).astype(np.int32).clip(0)
    end = start + rng.integers(1, 900, n).astype(np.int32)
    run(cov, np.zeros(n, np.int32), start, end, [20000], 1000, 10)
    run(cov, np.zeros(n, np.int32), start, end, [20000], 3, 100)


@pytest.mark.parametrize('case', ['cov_order', 'cov_dense', 'cov_nointra'])
def test_goldens_bit_exact_gff(cov, case):
    m = MAN[case]
    sizes = {l.split('\t')[0]: int(l.split('\t')[1]) for l in read_golden(case + '.lens').splitlines()}
    names = sorted(sizes, key=lambda s: s.encode())
    text = ao.GFF_HEADER_SELF
    blocks = [(case + '.tab', m['minCov'], m['label'])]
    if m['has_intra']:
        blocks.append((case + '.tab_intra.tab', m['intraCov'], m['label'] + '_intra'))
    for fn, c, label in blocks:
        chrom, start, end = tab_to_arrays(read_golden(fn).splitlines(), names)
        seg = cov.coverage_segments(chrom, start, end, [sizes[n] for n in names], c, m['minLen'])
        text += ''.join(segments_to_gff_rows(seg, names, 'mimeo-self', label, m['prefix']))
    assert text == read_golden(case + '.gff3')


def test_config2_shape_scaled_against_c_oracle(cov, oracle_build):
    """C2-shaped input (hotspots + uniform, 50 scaffolds) at 1/20 scale: 500 k hits over 5 Mbp."""
    import ctypes, os
    chrom, start, end, sizes = synth_hits(seed=1002, nchrom=50, chrom_size=100_000, nhits=500_000, hotspots=100)
    got = cov.coverage_segments(chrom, start, end, sizes, 21, 100)
    lib = ctypes.CDLL(os.path.join(oracle_build, 'libannot_oracle.so'))
    lib.ora_coverage_segments.restype = ctypes.c_long
    cap = len(chrom) + 8
    oc, os_, oe = (np.zeros(cap, np.int32) for _ in range(3))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    k = lib.ora_coverage_segments(P(chrom), P(start), P(end), ctypes.c_long(len(chrom)), P(sizes), ctypes.c_int(len(sizes)),
                                  ctypes.c_int(21), ctypes.c_int(100), P(oc), P(os_), P(oe), ctypes.c_long(cap))
    assert k == len(got[0]) and k > 100
    assert (got[0] == oc[:k]).all() and (got[1] == os_[:k]).all() and (got[2] == oe[:k]).all()


def test_full_size_properties(cov):
    """BASELINE config 2 at full size (10 M hits / 100 Mbp): size-independent properties."""
    chrom, start, end, sizes = synth_hits(seed=1002, nchrom=50, chrom_size=2_000_000, nhits=10_000_000, hotspots=2000)
    c1, s1, e1 = cov.coverage_segments(chrom, start, end, sizes, 21, 100)
    assert len(c1) > 1000
    key = c1.astype(np.int64) * (1 << 32) + s1
    assert (np.diff(key) > 0).all()                                  # sorted, strictly increasing
    assert ((e1 - s1) >= 100).all() and (e1 <= sizes[c1]).all()
    same = c1[1:] == c1[:-1]
    assert (s1[1:][same] > e1[:-1][same]).all()                      # merged: no overlap, no book-ends
    # permutation invariance (hit order must not matter)
    p = np.random.default_rng(1).permutation(len(chrom))
    c2, s2, e2 = cov.coverage_segments(chrom[p], start[p], end[p], sizes, 21, 100)
    assert (c1 == c2).all() and (s1 == s2).all() and (e1 == e2).all()
    # monotone in cov: covered bases shrink as the threshold rises; every cov=24 run lies inside a cov=21 run
    c4, s4, e4 = cov.coverage_segments(chrom, start, end, sizes, 24, 1)
    c3, s3, e3 = cov.coverage_segments(chrom, start, end, sizes, 21, 1)
    assert (e4 - s4).sum() <= (e3 - s3).sum()
    k3s = c3.astype(np.int64) * (1 << 32) + s3
    idx = np.searchsorted(k3s, c4.astype(np.int64) * (1 << 32) + s4, side='right') - 1
    assert (idx >= 0).all() and (c3[idx] == c4).all() and (s3[idx] <= s4).all() and (e3[idx] >= e4).all()
    # duplicating every hit doubles the depth: cov=42 on doubled input == cov=21 on the original
    cd, sd, ed = cov.coverage_segments(np.tile(chrom, 2), np.tile(start, 2), np.tile(end, 2), sizes, 42, 1)
    assert (cd == c3).all() and (sd == s3).all() and (ed == e3).all()


def test_device_resident_entry_points(cov, oracle_build):
    """Hits already in HBM (torch tensors as plain device buffers): `mb2_coverage_segments_into` writes into the caller's
    arrays; too small a capacity is reported and retried; same segments as the host entry point."""
    import ctypes as C
    import torch
    from mimeo_b200 import _lib
    chrom, start, end, sizes = synth_hits(seed=7, nchrom=6, chrom_size=300_000, nhits=200_000, hotspots=40)
    want = cov.coverage_segments(chrom, start, end, sizes, 9, 20)
    assert len(want[0]) > 10
    dev = [torch.from_numpy(a).cuda() for a in (chrom, start, end)]
    for capacity in (0, 7, len(want[0])):
        got = cov.coverage_segments_device(dev[0], dev[1], dev[2], sizes, 9, 20, capacity=capacity)
        for g, w in zip(got, want):
            assert g.is_cuda and (g.cpu().numpy() == w).all()
    # the raw call reports the needed size and writes nothing when the arrays are too small
    out = torch.full((3, 5), -7, dtype=torch.int32, device='cuda')
    n = C.c_uint64(0)
    s64 = np.ascontiguousarray(sizes, dtype=np.int64)
    rc = _lib.lib().mb2_coverage_segments_into(dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), len(chrom), s64.ctypes.data,
                                               len(s64), 9, 20, out.data_ptr(), out.data_ptr() + 20, out.data_ptr() + 40, 5, C.byref(n))
    torch.cuda.synchronize()
    assert rc == _lib.ERR_CAPACITY and int(n.value) == len(want[0]) and bool((out == -7).all())
    # very many runs (more than the single-launch run stage handles): one short hit every 16 bases
    m = 300_000
    st = (np.arange(m, dtype=np.int32) * 16)
    big = cov.coverage_segments(np.zeros(m, np.int32), st, st + 5, [16 * m + 100], 1, 1)
    assert len(big[0]) == m and (big[1] == st).all() and (big[2] == st + 5).all()
    gd = cov.coverage_segments_device(torch.zeros(m, dtype=torch.int32, device='cuda'), torch.from_numpy(st).cuda(),
                                      torch.from_numpy(st + 5).cuda(), [16 * m + 100], 1, 1)
    assert gd[0].numel() == m and (gd[1].cpu().numpy() == st).all() and (gd[2].cpu().numpy() == st + 5).all()


def test_large_sparse_genome_many_ctas_per_slot(cov, oracle_build):
    """300 Mbp in 5 scaffolds: more 8192-base tiles than one wave of CTAs may own (32 each), so the tile kernel runs in
    several waves and most tiles are empty; 400 k hits, bit-exact against the C oracle."""
    import ctypes, os
    chrom, start, end, sizes = synth_hits(seed=31, nchrom=5, chrom_size=60_000_000, nhits=400_000, hotspots=300)
    got = cov.coverage_segments(chrom, start, end, sizes, 2, 50)
    lib = ctypes.CDLL(os.path.join(oracle_build, 'libannot_oracle.so'))
    lib.ora_coverage_segments.restype = ctypes.c_long
    cap = len(chrom) + 8
    oc, os_, oe = (np.zeros(cap, np.int32) for _ in range(3))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    k = lib.ora_coverage_segments(P(chrom), P(start), P(end), ctypes.c_long(len(chrom)), P(sizes), ctypes.c_int(len(sizes)),
                                  ctypes.c_int(2), ctypes.c_int(50), P(oc), P(os_), P(oe), ctypes.c_long(cap))
    assert k == len(got[0]) and k > 300
    assert (got[0] == oc[:k]).all() and (got[1] == os_[:k]).all() and (got[2] == oe[:k]).all()
