"""Host side of the coverage stage: hit triples in, merged repeat segments out (kernel family d).

Mirrors what the reference's script does between the .tab file and the GFF3 rows
(wrappers.py:1120-1167) but the work runs in libmimeo_b200 on the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _as_i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a


def coverage_segments(chrom, start, end, chrom_sizes, min_cov: int, min_len: int):
    """HOST arrays in, HOST arrays out (copies inside). Returns (chrom_idx, start, end) int32 arrays."""
    _lib.init()
    chrom, start, end = _as_i32(chrom), _as_i32(start), _as_i32(end)
    if not (len(chrom) == len(start) == len(end)):
        raise ValueError('chrom/start/end must have equal length')
    sizes = np.ascontiguousarray(chrom_sizes, dtype=np.int64)
    seg = _lib.Segments()
    _lib.check(_lib.lib().mb2_coverage_segments(chrom.ctypes.data, start.ctypes.data, end.ctypes.data, len(chrom),
                                                sizes.ctypes.data, len(sizes), int(min_cov), int(min_len), C.byref(seg)))
    try:
        n = int(seg.n)
        if n == 0:
            z = np.zeros(0, dtype=np.int32)
            return z, z.copy(), z.copy()
        out = tuple(np.ctypeslib.as_array(p, shape=(n,)).copy() for p in (seg.chrom, seg.start, seg.end))
    finally:
        _lib.lib().mb2_free_segments(C.byref(seg))
    return out


def coverage_segments_device(d_chrom, d_start, d_end, chrom_sizes, min_cov: int, min_len: int, capacity: int = 0):
    """torch CUDA int32 tensors in (used only as device buffers); returns torch CUDA int32 tensors.
    The library writes straight into one (3, capacity) tensor allocated here (`mb2_coverage_segments_into`): nothing is
    copied afterwards; a result larger than `capacity` (default min(#hits, 2^18)) costs one retry at the reported size."""
    import torch
    _lib.init()
    for t in (d_chrom, d_start, d_end):
        assert t.is_cuda and t.dtype == torch.int32 and t.is_contiguous()
    sizes = np.ascontiguousarray(chrom_sizes, dtype=np.int64)
    nhits = d_chrom.numel()
    cap = int(capacity) if capacity > 0 else max(1, min(nhits, 1 << 18))
    n = C.c_uint64(0)
    while True:
        out = torch.empty((3, cap), dtype=torch.int32, device=d_chrom.device)
        torch.cuda.current_stream().synchronize()   # inputs (and the recycled output block) are quiescent before the library's stream touches them
        p = out.data_ptr()
        rc = _lib.lib().mb2_coverage_segments_into(d_chrom.data_ptr(), d_start.data_ptr(), d_end.data_ptr(), nhits, sizes.ctypes.data,
                                                   len(sizes), int(min_cov), int(min_len), p, p + 4 * cap, p + 8 * cap, cap, C.byref(n))
        if rc == _lib.ERR_CAPACITY and int(n.value) > cap:
            cap = int(n.value)
            continue
        _lib.check(rc)
        break
    k = int(n.value)
    return out[0, :k], out[1, :k], out[2, :k]
