"""Two pieces of index arithmetic of the seed scan (mimeo_b200/csrc/seed.cu), restated with Python integers and held against the
direct definition. No GPU needed; the kernel itself is held to the oracle by tests/test_gpu_hsps.py and the parity tests.

1. SeqBlock::windows: from the 128 columns that start at the 64-column boundary below p, cut the 32-column windows at p,
   p + 12 and p + 32 (bases: two bits per column) and the flag windows (one bit per column) with two levels of word selects
   and funnel shifts.
2. The descriptor of every hit of a batch: descriptors head .. head + 31 of the ring, one OR-reduction of the offsets at
   which descriptors start inside the batch, one popcount per lane (instead of a binary search per lane)."""
import numpy as np

M32 = 0xFFFFFFFF


def funnel_r(lo, hi, s):
    return (((hi << 32) | lo) >> (s & 31)) & M32


def windows(w, n, p, c0):
    """the kernel's arithmetic: w = four 64-bit words of bases, n = four 32-bit words of flags"""
    off = p - c0
    sh = (off & 15) * 2
    b0, b1 = bool(off & 16), bool(off & 32)
    x = []
    for m in range(4):
        x += [w[m] & M32, w[m] >> 32]
    a = [x[m + 1] if b0 else x[m] for m in range(7)]
    y = [a[m + 2] if b1 else a[m] for m in range(5)]
    f0, f1 = funnel_r(y[0], y[1], sh), funnel_r(y[1], y[2], sh)
    r0, r1 = funnel_r(y[2], y[3], sh), funnel_r(y[3], y[4], sh)
    f, r = f0 | (f1 << 32), r0 | (r1 << 32)
    l = funnel_r(f0, f1, 24) | (funnel_r(f1, r0, 24) << 32)
    sn = off & 31
    z = [n[1], n[2], n[3]] if b1 else [n[0], n[1], n[2]]
    fn, rn = funnel_r(z[0], z[1], sn), funnel_r(z[1], z[2], sn)
    ln = funnel_r(fn, rn, 12)
    return f, l, r, fn, ln, rn


def test_block_windows_equal_direct_extraction():
    rng = np.random.default_rng(5)
    for _ in range(300):
        w = [int(rng.integers(0, 1 << 63)) * 2 + int(rng.integers(0, 2)) for _ in range(4)]
        n = [int(rng.integers(0, 1 << 32)) for _ in range(4)]
        big = sum(w[m] << (64 * m) for m in range(4))           # 128 columns, two bits each
        flags = sum(n[m] << (32 * m) for m in range(4))          # 128 columns, one bit each
        c0 = 64 * int(rng.integers(1, 1000))
        for off in range(64):
            f, l, r, fn, ln, rn = windows(w, n, c0 + off, c0)
            assert f == (big >> (2 * off)) & (2 ** 64 - 1)
            assert l == (big >> (2 * (off + 12))) & (2 ** 64 - 1)
            assert r == (big >> (2 * (off + 32))) & (2 ** 64 - 1)
            assert fn == (flags >> off) & M32 and rn == (flags >> (off + 32)) & M32 and ln == (flags >> (off + 12)) & M32


def test_descriptor_lookup_equals_binary_search():
    rng = np.random.default_rng(6)
    for _ in range(2000):
        # a ring of descriptors: strictly increasing first-hit numbers (every descriptor holds >= 1 hit), wrapping arithmetic
        base = int(rng.integers(0, 1 << 32))
        sizes = rng.choice([1, 1, 1, 2, 3, 7, 40, 200], size=int(rng.integers(1, 60)))
        cum = [(base + int(c)) & M32 for c in np.concatenate([[0], np.cumsum(sizes)[:-1]])]
        total = int(sizes.sum())
        # the batch starts somewhere inside descriptor `head`
        head = int(rng.integers(0, len(cum)))
        lo_hit = (cum[head] - base) & M32
        hi_hit = lo_hit + int(sizes[head]) - 1
        consumed_rel = int(rng.integers(lo_hit, hi_hit + 1))
        consumed = (base + consumed_rel) & M32
        nb = min(32, total - consumed_rel)
        tail = len(cum)
        # kernel: lane i looks at descriptor head + i
        starts = 0
        for lane in range(32):
            di = head + lane
            if di < tail and lane > 0:
                first = (cum[di] - consumed) & M32
                if first < 32:
                    starts |= 1 << first
        for lane in range(nb):
            got = head + bin(starts & (((2 << lane) - 1) & M32)).count('1')
            h_rel = consumed_rel + lane
            want = max(d for d in range(tail) if ((cum[d] - base) & M32) <= h_rel)
            assert got == want
