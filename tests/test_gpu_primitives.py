"""GPU parity: hand-written radix sort and scan against numpy (bit-exact)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def L():
    from mimeo_b200 import _lib
    _lib.init()
    return _lib


@pytest.mark.parametrize('n', [0, 1, 31, 32, 33, 4095, 4096, 4097, 100003, 3_000_000])
def test_scan(L, n):
    rng = np.random.default_rng(n)
    a = rng.integers(0, 1000, n).astype(np.uint32)
    want = np.concatenate(([0], np.cumsum(a, dtype=np.uint64)[:-1])).astype(np.uint32) if n else a.copy()
    tot = np.zeros(1, np.uint32)
    b = a.copy()
    L.check(L.lib().mb2_test_scan_u32(b.ctypes.data, n, tot.ctypes.data))
    assert (b == want).all()
    assert int(tot[0]) == int(a.sum(dtype=np.uint64) & 0xffffffff)


@pytest.mark.parametrize('n,b0,b1', [(0, 0, 32), (1, 0, 32), (1000, 0, 32), (4097, 13, 27), (123457, 0, 32),
                                     (2_000_001, 13, 31), (1_000_000, 5, 6)])
def test_sort_u32_stable_bit_range(L, n, b0, b1):
    rng = np.random.default_rng(n + b0)
    k = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    v = np.arange(n, dtype=np.uint32)
    mask = np.uint32(((1 << (b1 - b0)) - 1) << b0) if b1 - b0 < 32 else np.uint32(0xffffffff)
    order = np.argsort(k & mask, kind='stable')
    kk, vv = k.copy(), v.copy()
    L.check(L.lib().mb2_test_sort_u32(kk.ctypes.data, vv.ctypes.data, n, b0, b1))
    assert (kk == k[order]).all() and (vv == v[order]).all()
    kk = k.copy()
    L.check(L.lib().mb2_test_sort_u32(kk.ctypes.data, None, n, b0, b1))
    assert (kk == k[order]).all()


@pytest.mark.parametrize('n,b0,b1,presorted', [(1, 0, 32, False), (4096, 13, 27, True), (4097, 13, 27, False), (777777, 13, 28, True),
                                               (2_000_001, 13, 31, False), (50001, 3, 9, False)])
def test_sort_u32_pair_one_launch_set(L, n, b0, b1, presorted):
    """Two key arrays through the same launches: each must come out as its own stable sort (the sorted-input fast path of
    the histogram is exercised by the presorted cases)."""
    rng = np.random.default_rng(n + b1)
    a = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    b = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    if presorted:
        a.sort()
        b[: n // 2] = np.uint32(12345 << b0)       # one digit for half of the keys
    mask = np.uint32(((1 << (b1 - b0)) - 1) << b0)
    wa, wb = a[np.argsort(a & mask, kind='stable')], b[np.argsort(b & mask, kind='stable')]
    ga, gb = a.copy(), b.copy()
    L.check(L.lib().mb2_test_sort_u32_pair(ga.ctypes.data, gb.ctypes.data, n, b0, b1))
    assert (ga == wa).all() and (gb == wb).all()


@pytest.mark.parametrize('n,b0,b1', [(5, 0, 64), (100001, 0, 64), (1_500_000, 0, 40), (300000, 20, 64)])
def test_sort_u64(L, n, b0, b1):
    rng = np.random.default_rng(n)
    k = rng.integers(0, 1 << 63, n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n).astype(np.uint64)
    if n > 10:
        k[: n // 3] = k[n // 3: 2 * (n // 3)]      # plenty of duplicates to exercise stability
    v = np.arange(n, dtype=np.uint32)
    mask = np.uint64((((1 << (b1 - b0)) - 1) << b0) & 0xffffffffffffffff)
    order = np.argsort(k & mask, kind='stable')
    kk, vv = k.copy(), v.copy()
    L.check(L.lib().mb2_test_sort_u64(kk.ctypes.data, vv.ctypes.data, n, b0, b1))
    assert (kk == k[order]).all() and (vv == v[order]).all()
