"""Host side of the alignment stage: device genomes in, LASTZ-style hit rows out.

`align()` replaces every `lastz T Q ...` process of the reference's generated script
(wrappers.py:1025-1037 et al.); `tab_blocks()` replaces the sed/awk/awk/awk/sed/sort filter that follows
each of them (wrappers.py:1040-1056): keep length1 >= minLen and idPct >= minIdt, print the 10 columns,
sort each pair's rows by `sort -k 1,1 -k 3n,4n` (C locale, whole-line tie-break).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib
from .genome import Genome, align_params

HIT_FIELDS = ('t_id', 'q_id', 'strand', 'start1', 'end1', 'start2', 'end2', 'score', 'nmatch', 'ncols')
STAT_NAMES = ('survivors', 'seed_hits', 'leaders', 'stage1_cells', 'hsps', 'stage2_extensions', 'stage2_cells',
              'gapped_cells', 'alignments', 'anchors_extended')


def align(T: Genome, Q: Genome, params: Optional[_lib.AlignParams] = None, strands: int = 3,
          Q_aux: Optional[Genome] = None, t_same_q=None) -> Tuple[Dict[str, np.ndarray], Dict[str, int]]:
    """All scaffolds of T against all scaffolds of Q on the GPU. Q_aux: prebuilt Q.both_strands() (strands=3) or
    Q.revcomp() (strands=2) to avoid rebuilding it per call. t_same_q[i] = index of the query scaffold identical to target
    scaffold i (or -1): lets a rank that holds a SUBSET of a genome as T keep the closed-form trivial self-alignment.
    Returns (hit columns, stage counters)."""
    if params is None:
        params = align_params()
    h = _lib.Hits()
    same = None if t_same_q is None else np.ascontiguousarray(t_same_q, dtype=np.int32)
    if same is not None and len(same) != len(T.names):
        raise ValueError('t_same_q needs one entry per target scaffold')
    _lib.check(_lib.lib().mb2_align(T.handle, Q.handle, Q_aux.handle if Q_aux is not None else None, C.byref(params),
                                    int(strands), same.ctypes.data if same is not None else None, C.byref(h)))
    try:
        n = int(h.n)
        cols = {f: (np.ctypeslib.as_array(getattr(h, f), shape=(n,)).copy() if n else np.zeros(0, np.int32)) for f in HIT_FIELDS}
        stats = {name: int(h.stats[i]) for i, name in enumerate(STAT_NAMES)}
    finally:
        _lib.lib().mb2_free_hits(C.byref(h))
    return cols, stats


def _hits_to_numpy(h) -> Tuple[Dict[str, np.ndarray], Dict[str, int]]:
    n = int(h.n)
    cols = {f: (np.ctypeslib.as_array(getattr(h, f), shape=(n,)).copy() if n else np.zeros(0, np.int32)) for f in HIT_FIELDS}
    stats = {name: int(h.stats[i]) for i, name in enumerate(STAT_NAMES)}
    return cols, stats


class DeviceHits:
    """The hit table of one alignment job kept in HBM (`mb2_hits_dev`): filter + sort, coverage and the projection onto the
    coverage stage run on it without the rows crossing PCIe; `download()` brings the (surviving) rows to the host for the
    .tab text."""

    def __init__(self, handle):
        self.handle = handle

    @classmethod
    def from_host(cls, cols: Dict[str, np.ndarray], nt: int, nq: int) -> 'DeviceHits':
        _lib.init()
        arrs = [np.ascontiguousarray(cols[f], dtype=np.int32) for f in HIT_FIELDS]
        h = _lib.Hits()
        for f, a in zip(HIT_FIELDS, arrs):
            setattr(h, f, a.ctypes.data_as(_lib.c_i32p))
        h.n = len(arrs[0])
        out = C.c_void_p()
        _lib.check(_lib.lib().mb2_hits_dev_upload(C.byref(h), int(nt), int(nq), C.byref(out)))
        return cls(out)

    def __len__(self):
        return int(_lib.lib().mb2_hits_dev_count(self.handle))

    def filter_sort(self, minLen, minIdt, map_rule: bool = False) -> int:
        """The awk filters + per-pair sort of wrappers.py:1044-1056 on the device, in place. Returns the rows kept."""
        n = C.c_uint64(0)
        _lib.check(_lib.lib().mb2_filter_sort(self.handle, float(minLen), float(minIdt), 1 if map_rule else 0, C.byref(n)))
        return int(n.value)

    def coverage(self, which: int, sizes, cov: int, minLen: int):
        """Segments (chrom, start, end) of the rows selected by `which` (0 all, 1 t != q, 2 t == q)."""
        sz = np.ascontiguousarray(sizes, dtype=np.int64)
        seg = _lib.Segments()
        _lib.check(_lib.lib().mb2_hits_dev_coverage(self.handle, int(which), sz.ctypes.data, len(sz), int(cov), int(minLen), C.byref(seg)))
        try:
            n = int(seg.n)
            if n == 0:
                z = np.zeros(0, dtype=np.int32)
                return z, z.copy(), z.copy()
            return tuple(np.ctypeslib.as_array(p, shape=(n,)).copy() for p in (seg.chrom, seg.start, seg.end))
        finally:
            _lib.lib().mb2_free_segments(C.byref(seg))

    def download(self) -> Tuple[Dict[str, np.ndarray], Dict[str, int]]:
        h = _lib.Hits()
        _lib.check(_lib.lib().mb2_hits_dev_download(self.handle, C.byref(h)))
        try:
            return _hits_to_numpy(h)
        finally:
            _lib.lib().mb2_free_hits(C.byref(h))

    def close(self):
        if getattr(self, 'handle', None):
            _lib.lib().mb2_hits_dev_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def align_device(T: Genome, Q: Genome, params: Optional[_lib.AlignParams] = None, strands: int = 3,
                 Q_aux: Optional[Genome] = None, t_same_q=None) -> DeviceHits:
    """`align()` with the rows left in HBM."""
    if params is None:
        params = align_params()
    same = None if t_same_q is None else np.ascontiguousarray(t_same_q, dtype=np.int32)
    if same is not None and len(same) != len(T.names):
        raise ValueError('t_same_q needs one entry per target scaffold')
    out = C.c_void_p()
    _lib.check(_lib.lib().mb2_align_dev(T.handle, Q.handle, Q_aux.handle if Q_aux is not None else None, C.byref(params),
                                        int(strands), same.ctypes.data if same is not None else None, C.byref(out)))
    return DeviceHits(out)


def pct_text(nmatch: int, ncols: int) -> str:
    """LASTZ prints the identity percentage with '%.1f' of 100*n/d evaluated in double precision."""
    return '%.1f' % (100.0 * nmatch / ncols) if ncols else '0.0'


def _sort_n(s: str) -> float:
    return float(s)


def tab_blocks(hits: Dict[str, np.ndarray], tnames: List[str], qnames: List[str], minLen, minIdt) -> Dict[Tuple[int, int], List[str]]:
    """Filtered, sorted 10-column rows (with newline) per (t_id, q_id) pair, formatted natively (`mb2_format_tab`)."""
    out: Dict[Tuple[int, int], List[str]] = {}
    n = len(hits['t_id'])
    if n == 0:
        return out
    cols = [np.ascontiguousarray(hits[f], dtype=np.int32) for f in HIT_FIELDS]
    tn = (C.c_char_p * len(tnames))(*[s.encode() for s in tnames])
    qn = tn if qnames is tnames else (C.c_char_p * len(qnames))(*[s.encode() for s in qnames])
    t = _lib.TabText()
    _lib.check(_lib.lib().mb2_format_tab(*[c.ctypes.data for c in cols], n, tn, len(tnames), qn, len(qnames),
                                         float(minLen), float(minIdt), C.byref(t)))
    try:
        nb = int(t.nblocks)
        if nb:
            text = C.string_at(t.text, int(t.nbytes)).decode()
            off = np.ctypeslib.as_array(t.off, shape=(nb + 1,)).tolist()
            tid = np.ctypeslib.as_array(t.t_id, shape=(nb,)).tolist()
            qid = np.ctypeslib.as_array(t.q_id, shape=(nb,)).tolist()
            if len(text) == int(t.nbytes):                      # pure ASCII: byte offsets are character offsets
                for b in range(nb):
                    out[(tid[b], qid[b])] = text[off[b]:off[b + 1]].splitlines(keepends=True)
            else:
                raw = C.string_at(t.text, int(t.nbytes))
                for b in range(nb):
                    out[(tid[b], qid[b])] = raw[off[b]:off[b + 1]].decode().splitlines(keepends=True)
    finally:
        _lib.lib().mb2_free_tab_text(C.byref(t))
    return out
