"""The y-drop extension kernel SOURCE (mimeo_b200/csrc/ydrop_warp.cuh: forward pass + walk-back) executed on the CPU by a
32-lane warp emulator (tests/emu/) and held to the oracle's ydrop_extend (oracle/lastz_oracle.c) result by result:
score, end point, matches, aligned columns. No GPU needed; the same source is what nvcc compiles for sm_100a."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import lastz_oracle as lo

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, 'emu', '_build', 'libydrop_emu.so')
PAD = 2048


def emu_lib():
    srcs = [os.path.join(HERE, 'emu', 'ydrop_emu.cpp'), os.path.join(HERE, 'emu', 'ydrop_emu.h'),
            os.path.join(ROOT, 'mimeo_b200', 'csrc', 'ydrop_warp.cuh')]
    if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-I', os.path.join(HERE, 'emu'), '-o', SO, srcs[0]])
    l = C.CDLL(SO)
    l.emu_extend.restype = C.c_int
    l.emu_extend.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_int, C.c_void_p]
    return l


def kernel_codes(codes):
    """oracle codes (0..3, 4 = other) -> the device `codes` mirror: base | parity << 2, N = 8, pads = 12 on both sides."""
    lut = np.array([0, 5, 6, 3, 8], dtype=np.uint8)
    out = np.full(len(codes) + 2 * PAD, 12, dtype=np.uint8)
    out[PAD:PAD + len(codes)] = lut[codes]
    return out


def oracle_ext(t, a1, q, a2, d, p):
    l = lo.lib()
    l.lzo_ydrop_extend.restype = C.c_long
    l.lzo_ydrop_extend.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_void_p, C.c_long, C.c_long, C.c_int, C.POINTER(lo.Params), C.c_void_p]
    out = np.zeros(5, np.int32)
    tn, qn = (len(t) - a1, len(q) - a2) if d > 0 else (a1, a2)
    cells = l.lzo_ydrop_extend(t.ctypes.data, a1, tn, q.ctypes.data, a2, qn, d, C.byref(p), out.ctypes.data)
    return out.tolist(), cells


ALL_LAYOUTS = sum(1 << (s // 4) for s in (8, 12, 16, 20, 24, 32, 48, 64))


def emu_ext(l, tk, a1, qk, a2, d, p, max_s=32, nchunks=4096, layouts=ALL_LAYOUTS, priv=7):
    """priv: chunks of the warp slot's private part of the trace pool (the rest of the trace goes to the shared part)."""
    out = np.zeros(11, np.int32)
    l.emu_extend(tk.ctypes.data, qk.ctypes.data, PAD + a1, PAD + a2, d, p.gap_open, p.gap_extend, p.ydrop, max_s, nchunks, layouts, priv, out.ctypes.data)
    return out.tolist()


def mutate(rng, s, sub, indel, maxgap=3):
    out = []
    for b in s:
        r = rng.random()
        if r < indel / 2:
            continue
        if r < indel:
            out.extend(rng.integers(0, 4, rng.integers(1, maxgap + 1)).tolist())
        out.append(int(rng.integers(0, 4)) if rng.random() < sub else int(b))
    return np.array(out, dtype=np.uint8)


def make_case(rng, flank, core, sub, indel, maxgap=3, n_frac=0.0):
    c = rng.integers(0, 4, core).astype(np.uint8)
    t = np.concatenate([rng.integers(0, 4, flank[0]).astype(np.uint8), c, rng.integers(0, 4, flank[1]).astype(np.uint8)])
    q = np.concatenate([rng.integers(0, 4, flank[2]).astype(np.uint8), mutate(rng, c, sub, indel, maxgap), rng.integers(0, 4, flank[3]).astype(np.uint8)])
    if n_frac:
        for s in (t, q):
            for _ in range(max(1, int(n_frac * len(s) / 20))):
                x = int(rng.integers(0, len(s))); s[x:x + int(rng.integers(1, 20))] = 4
    return t, q


def check(l, t, q, a1, a2, p, max_s=32):
    tk, qk = kernel_codes(t), kernel_codes(q)
    for d in (+1, -1):
        want, cells = oracle_ext(t, a1, q, a2, d, p)
        got = emu_ext(l, tk, a1, qk, a2, d, p, max_s)
        if got[5] == 2 and max_s < 64:          # band too wide for the common kernel: the host reruns it with the wide one
            got = emu_ext(l, tk, a1, qk, a2, d, p, 64)
        assert got[5] == 0, f'status {got[5]} dir {d} (want {want})'
        assert got[:5] == want, f'dir {d}: emulated kernel {got[:5]} != oracle {want} (kbest {got[7]}, oracle cells {cells}, kernel band cells {got[6]})'


@pytest.mark.parametrize('seed', range(12))
def test_ydrop_kernel_source_matches_oracle(seed):
    l = emu_lib()
    rng = np.random.default_rng(100 + seed)
    p = lo.default_params(3000)
    core = int(rng.integers(150, 1500))
    flank = [int(rng.integers(0, 700)) for _ in range(4)]
    t, q = make_case(rng, flank, core, sub=float(rng.uniform(0.02, 0.2)), indel=float(rng.uniform(0.0, 0.02)))
    # anchor inside the homologous core: matching prefix offsets keep (a1, a2) on the true path only roughly, which is fine
    a1 = flank[0] + core // 2
    a2 = min(len(q) - 1, flank[2] + core // 2)
    check(l, t, q, a1, a2, p)


def test_ydrop_kernel_source_scaffold_ends_and_n_runs():
    l = emu_lib()
    rng = np.random.default_rng(7)
    p = lo.default_params(3000)
    for flank in ([0, 0, 0, 0], [3, 0, 0, 5], [0, 400, 300, 0]):
        t, q = make_case(rng, flank, 600, 0.08, 0.01, n_frac=0.02)
        check(l, t, q, flank[0] + 300, min(len(q) - 1, flank[2] + 300), p)
    # identical sequences: the band is widest, the extension runs to both scaffold ends
    t = rng.integers(0, 4, 900).astype(np.uint8)
    check(l, t, t.copy(), 450, 450, p)
    check(l, t, t.copy(), 0, 0, p)
    check(l, t, t.copy(), 899, 899, p)


def test_ydrop_kernel_source_long_gaps_and_frame_moves():
    l = emu_lib()
    rng = np.random.default_rng(11)
    p = lo.default_params(3000)
    # long alignment (several frame moves, layout changes 16 -> 24 -> 32) with gaps up to 120 bases
    t, q = make_case(rng, [200, 200, 200, 200], 5000, 0.05, 0.004, maxgap=120)
    check(l, t, q, 200 + 2500, min(len(q) - 1, 200 + 2500), p)


ASAN_CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r)
from tests import test_ydrop_emu as T
from oracle import lastz_oracle as lo
T.SO = %(so)r
l = T.emu_lib()
p = lo.default_params(3000)
for seed in range(8):
    rng = np.random.default_rng(900 + seed)
    core = int(rng.integers(50, 2500)); flank = [int(rng.integers(0, 900)) for _ in range(4)]
    t, q = T.make_case(rng, flank, core, float(rng.uniform(0, 0.25)), float(rng.uniform(0, 0.03)), int(rng.choice([1, 3, 30, 150])), n_frac=float(rng.choice([0, 0.01])))
    a1 = min(len(t) - 1, flank[0] + core // 2); a2 = min(len(q) - 1, flank[2] + core // 2)
    tk, qk = T.kernel_codes(t), T.kernel_codes(q)
    for d in (1, -1):
        want, _ = T.oracle_ext(t, a1, q, a2, d, p)
        got = T.emu_ext(l, tk, a1, qk, a2, d, p, 64 if seed %% 2 else 32, 2048, T.ALL_LAYOUTS, priv=int(rng.integers(0, 9)))
        if got[5] == 2:
            got = T.emu_ext(l, tk, a1, qk, a2, d, p, 64, 2048)
        assert got[:5] == want and got[5] == 0, (seed, d, want, got)
print('ASAN-SWEEP-OK')
'''


def test_ydrop_kernel_source_under_address_sanitizer():
    """Memory safety of the kernel source (trace pool chunks, re-layout scratch, sequence windows): the same sweep with the
    emulator built -fsanitize=address. The GPU pool has no compute-sanitizer; this is the bounds check the source gets."""
    import sys
    asan = subprocess.run(['g++', '-print-file-name=libasan.so'], capture_output=True, text=True).stdout.strip()
    if not asan or not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip('libasan not installed')
    so = os.path.join(HERE, 'emu', '_build', 'libydrop_emu_asan.so')
    srcs = [os.path.join(HERE, 'emu', 'ydrop_emu.cpp'), os.path.join(HERE, 'emu', 'ydrop_emu.h'),
            os.path.join(ROOT, 'mimeo_b200', 'csrc', 'ydrop_warp.cuh')]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(['g++', '-O1', '-g', '-fsanitize=address', '-fno-omit-frame-pointer', '-std=c++17', '-shared', '-fPIC',
                               '-I', os.path.join(HERE, 'emu'), '-o', so, srcs[0]])
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS='detect_leaks=0:abort_on_error=0')
    r = subprocess.run([sys.executable, '-c', ASAN_CHILD % dict(root=ROOT, so=so)], env=env, capture_output=True, text=True, timeout=900)
    assert 'ERROR: AddressSanitizer' not in r.stderr, r.stderr[-3000:]
    assert r.returncode == 0 and 'ASAN-SWEEP-OK' in r.stdout, (r.stdout[-500:], r.stderr[-2000:])
