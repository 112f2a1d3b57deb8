"""The arithmetic identity behind the first-stage table of mimeo_b200/csrc/seed.cu (s1_entry / xdrop_window30) and the
three-column table of xdrop_table.cuh, restated in numpy and held against the column-by-column x-drop rule of the oracle
(oracle/lastz_oracle.c: gap-free extension: run += score; new maximum, or stop when run < best - X).

State of an extension: (best, D) with D = (best - run) + 125. Per chunk of three columns with prefix sums p1, p2, p3:
    terminate  iff  125 + X + min(p) < D        (then no column of the chunk raised best: |p_i - p_j| <= 250 < X)
    DM = max(D, 125 + max(p));  best += DM - D;  D = DM - p3
No GPU needed: this pins the claim the kernel comments make, for every x-drop the kernel accepts at its lower edge."""
import numpy as np
import pytest

HOXD70 = np.array([[91, -114, -31, -123], [-114, 100, -125, -31], [-31, -125, 100, -114], [-123, -31, -114, 91]], dtype=np.int64)
DONE = 1 << 24


def column_rule(t, q, X):
    """(best, columns consumed before the stop or None if the window ran out)"""
    run = best = 0
    for c in range(len(t)):
        run += int(HOXD70[t[c], q[c]])
        if run > best:
            best = run
        elif run < best - X:
            return best, c
    return best, None


def chunk_rule(t, q, X):
    best, D = 0, 125
    for k in range(0, len(t) - len(t) % 3, 3):
        s = HOXD70[t[k:k + 3], q[k:k + 3]]
        p = np.cumsum(s)
        mn, mx, sm = int(p.min()), int(p.max()), int(p[-1])
        # the fields as the kernel packs them (seed.cu: s1_entry) must fit their widths
        assert 0 <= mx + 125 < 512 and 0 <= mn + X + 125 < 8192 and -512 <= sm < 512
        if D >= DONE // 2:
            continue
        term = (mn + X + 125) < D
        dm = max(D, mx + 125)
        best += dm - D
        D = DONE if term else dm - sm
    return best, D >= DONE // 2


@pytest.mark.parametrize('X', [251, 300, 910, 3000, 7766])
def test_chunked_rule_equals_column_rule(X):
    rng = np.random.default_rng(X)
    n_term = 0
    for trial in range(4000):
        n = 3 * int(rng.integers(1, 31))
        t = rng.integers(0, 4, n)
        # a mix of unrelated and related sequences: related ones keep the extension alive and move `best`
        q = rng.integers(0, 4, n) if trial % 3 == 0 else np.where(rng.random(n) < rng.uniform(0.05, 0.5), rng.integers(0, 4, n), t)
        want_best, stop = column_rule(t, q, X)
        got_best, terminated = chunk_rule(t, q, X)
        assert got_best == want_best, (trial, X)
        assert terminated == (stop is not None), (trial, X)
        n_term += terminated
    if X <= 910:
        assert n_term > 100          # the terminating branch was exercised


def test_field_extraction_by_multiplies():
    """seed.cu cuts the max-prefix and sum fields out of a packed entry with multiplies: umulhi(e * 2^10, 2^9) = bits 13..21 and
    mulhi(e, 2^10) = e >> 22 (arithmetic). Checked over every entry of the table for the default x-drop."""
    X = 910
    for idx in range(4096):
        q6, t6 = idx >> 6, idx & 63
        p = np.cumsum([HOXD70[(t6 >> (2 * c)) & 3, (q6 >> (2 * c)) & 3] for c in range(3)])
        sm, mx, mn = int(p[-1]), int(p.max()), int(p.min())
        e = ((sm & 0x3FF) << 22) | ((mx + 125) << 13) | (mn + X + 125)
        assert ((e * 1024) & 0xFFFFFFFF) * 512 >> 32 == mx + 125
        signed = e - (1 << 32) if e & 0x80000000 else e
        assert (signed * 1024) >> 32 == sm
        assert e & 0x1FFF == mn + X + 125
