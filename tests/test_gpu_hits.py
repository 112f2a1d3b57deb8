"""GPU parity of kernel (d) part 1 on the device hit table: `mb2_filter_sort` (identity / length predicate with LASTZ's
'%.1f' rounding, compaction, sort) and the coverage passes fed straight from it, against the plain statement of the
reference's awk filters (tests/helpers.py:py_tab_blocks, the checker that is itself held to the reference-made goldens)."""
import numpy as np
import pytest

from tests.helpers import py_tab_blocks

pytestmark = pytest.mark.gpu
FIELDS = ('t_id', 'q_id', 'strand', 'start1', 'end1', 'start2', 'end2', 'score', 'nmatch', 'ncols')


def random_hits(rng, n, nt, nq):
    h = {f: np.zeros(n, np.int32) for f in FIELDS}
    h['t_id'] = rng.integers(0, nt, n).astype(np.int32)
    h['q_id'] = rng.integers(0, nq, n).astype(np.int32)
    h['strand'] = rng.integers(0, 2, n).astype(np.int32)
    h['start1'] = rng.integers(1, 5000, n).astype(np.int32)
    h['end1'] = (h['start1'] + rng.integers(80, 400, n)).astype(np.int32)
    h['start2'] = rng.integers(1, 5000, n).astype(np.int32)
    h['end2'] = (h['start2'] + rng.integers(80, 400, n)).astype(np.int32)
    h['score'] = rng.integers(3000, 40000, n).astype(np.int32)
    h['ncols'] = rng.integers(50, 4000, n).astype(np.int32)
    h['nmatch'] = (h['ncols'] * rng.uniform(0.7, 1.0, n)).astype(np.int32)
    return h


def tie_rows():
    """nmatch / ncols pairs whose percentage sits exactly on a '%.1f' rounding boundary (x.x5): representable ties round half
    to even in printf, the others go where the correctly rounded double fell."""
    rows = []
    for nc in (8, 16, 40, 80, 200, 400, 800, 1000, 1600, 2000, 3200, 4000):
        for k in range(700, 1000):            # tenths k + 0.5  <=>  nm = (2k+1) * nc / 2000
            num = (2 * k + 1) * nc
            if num % 2000 == 0:
                rows.append((num // 2000, nc))
    return rows


def as_rows(h):
    return list(zip(*[np.asarray(h[f]).tolist() for f in FIELDS]))


@pytest.mark.parametrize('min_idt', [80, 87.3, 90])
def test_filter_sort_equals_the_awk_statement(min_idt):
    from mimeo_b200 import align as A
    rng = np.random.default_rng(3)
    nt, nq = 7, 9
    h = random_hits(rng, 20000, nt, nq)
    ties = tie_rows()
    assert len(ties) > 50
    for k, (nm, nc) in enumerate(ties):     # plant the boundary cases
        h['nmatch'][k], h['ncols'][k] = nm, nc
    # a few exact duplicates of the sort key: their relative order must be the input order (stable)
    for k in range(100, 140):
        for f in ('t_id', 'q_id', 'start1', 'end1'):
            h[f][k + 1000] = h[f][k]
    dh = A.DeviceHits.from_host(h, nt, nq)
    kept = dh.filter_sort(100, min_idt)
    got, _ = dh.download()
    dh.close()
    # the statement: per row '%.1f' text compare, then per (t, q) block by (start1, end1), stable
    pct = np.array([float('%.1f' % (100.0 * a / b)) for a, b in zip(h['nmatch'].tolist(), h['ncols'].tolist())])
    keep = ((h['end1'] - h['start1'] + 1) >= 100) & (pct >= min_idt)
    idx = np.flatnonzero(keep)
    order = sorted(idx.tolist(), key=lambda k: (h['t_id'][k], h['q_id'][k], h['start1'][k], h['end1'][k]))   # sorted() is stable
    want = [tuple(int(h[f][k]) for f in FIELDS) for k in order]
    assert kept == len(want) and as_rows(got) == want
    # and through the text formatter the blocks are what the awk | sort statement writes
    tn, qn = [f't{k}' for k in range(nt)], [f'q{k}' for k in range(nq)]
    assert A.tab_blocks(got, tn, qn, 100, min_idt) == py_tab_blocks(h, tn, qn, 100, min_idt)


def test_map_rule_and_device_coverage():
    from mimeo_b200 import align as A, coverage
    rng = np.random.default_rng(4)
    nt = 5
    h = random_hits(rng, 5000, nt, nt)
    h['end1'][:50] = h['start1'][:50] + 99          # length1 == 100: kept by awk, dropped by import_Align's end - start >= 100
    h['nmatch'][:50] = h['ncols'][:50]
    dh = A.DeviceHits.from_host(h, nt, nt)
    dh.filter_sort(100, 90, map_rule=True)
    got, _ = dh.download()
    assert ((got['end1'] - got['start1']) >= 100).all() and len(got['t_id']) > 100
    sizes = [6000] * nt
    for which, mask in ((0, np.ones(len(got['t_id']), bool)), (1, got['t_id'] != got['q_id']), (2, got['t_id'] == got['q_id'])):
        want = coverage.coverage_segments(got['t_id'][mask], got['start1'][mask], got['end1'][mask], sizes, 2, 50)
        have = dh.coverage(which, sizes, 2, 50)
        assert all(np.array_equal(a, b) for a, b in zip(have, want)) and len(want[0]) > 0
    dh.close()
