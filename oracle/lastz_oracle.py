"""
lastz_oracle.py -- TEST INFRASTRUCTURE ONLY: ctypes front-end of oracle/lastz_oracle.c (the
"LASTZ-restatement", PARITY UNPINNED against a real LASTZ binary -- see the C file header) plus the
glue that turns its alignments into the text LASTZ would have written for mimeo's command line
(--format=general:name1,strand1,start1,end1,length1,name2,strand2,start2+,end2+,length2,score,identity
--markend; wrappers.py:1031) and the whole `mimeo self / x / map` reference pipeline on top of it
(oracle/annot_oracle.py restates the shell stages).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import annot_oracle as ao

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('hspthresh', 'xdrop', 'ydrop', 'gap_open', 'gap_extend', 'gappedthresh',
                                         'entropy', 'chain', 'gapped', 'transition')]


class Hsp(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('s1', 's2', 'len', 'score')]


class Aln(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('s1', 'e1', 's2', 'e2', 'score', 'nmatch', 'ncols', 'a1', 'a2')]


class Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ('seed_hits', 'leaders', 'extended', 'ungapped_cells', 'hsps_raw', 'hsps_kept',
                                         'chained', 'anchors_extended', 'gapped_cells')]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}

    def add(self, o):
        for n, _ in self._fields_:
            setattr(self, n, getattr(self, n) + getattr(o, n))


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, '_build', 'liblastz_oracle.so')
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(HERE, 'lastz_oracle.c')):
            subprocess.check_call(['make', '-C', HERE, '_build/liblastz_oracle.so'], stdout=subprocess.DEVNULL)
        l = C.CDLL(so)
        l.lzo_index_build.restype = C.c_void_p
        l.lzo_index_build.argtypes = [C.c_void_p, C.c_long]
        l.lzo_index_free.argtypes = [C.c_void_p]
        l.lzo_hsps_ix.restype = C.c_long
        l.lzo_hsps_ix.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.POINTER(Params), C.c_void_p,
                                  C.c_long, C.POINTER(Stats)]
        l.lzo_chain.restype = C.c_long
        l.lzo_chain.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
        l.lzo_gapped.restype = C.c_long
        l.lzo_gapped.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.POINTER(Params),
                                 C.c_void_p, C.c_long, C.POINTER(Stats)]
        l.lzo_align_tile_ix.restype = C.c_long
        l.lzo_align_tile_ix.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.POINTER(Params),
                                        C.c_void_p, C.c_long, C.POINTER(Stats)]
        l.lzo_anchor.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Hsp), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        l.lzo_seed_at.restype = C.c_int
        l.lzo_seed_at.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_long, C.c_long, C.c_int]
        l.lzo_entropy_q24.restype = C.c_uint32
        l.lzo_entropy_q24.argtypes = [C.c_void_p]
        l.lzo_default_params.argtypes = [C.POINTER(Params)]
        _LIB = l
    return _LIB


def default_params(hspthresh=3000, **kw) -> Params:
    p = Params()
    lib().lzo_default_params(C.byref(p))
    p.hspthresh = int(hspthresh)
    p.gappedthresh = int(hspthresh)
    for k, v in kw.items():
        setattr(p, k, int(v))
    return p


_ENC = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate('ACGT'):
    _ENC[ord(_c)] = _i
    _ENC[ord(_c.lower())] = _i | 8      # soft-masked: same base, bit 3 set (never seeded, still extended through)
for _c in range(ord('a'), ord('z') + 1):
    if _ENC[_c] == 4:
        _ENC[_c] = 4 | 8


def encode(seq) -> np.ndarray:
    """str / bytes / uint8 ASCII array -> codes 0..3 (ACGT), 4 (anything else)."""
    if isinstance(seq, str):
        seq = seq.encode()
    if isinstance(seq, (bytes, bytearray)):
        seq = np.frombuffer(bytes(seq), dtype=np.uint8)
    return _ENC[np.asarray(seq, dtype=np.uint8)]


def revcomp_codes(codes: np.ndarray) -> np.ndarray:
    r = codes[::-1].copy()
    m = (r & 7) < 4
    r[m] = (3 - (r[m] & 7)) | (r[m] & 8)
    return r


class TargetIndex:
    """Seed position table of one target scaffold (LASTZ rebuilds this per process; tests share it)."""

    def __init__(self, codes: np.ndarray):
        self.codes = np.ascontiguousarray(codes, dtype=np.uint8)
        self.h = lib().lzo_index_build(self.codes.ctypes.data, len(self.codes))

    def __del__(self):
        if getattr(self, 'h', None):
            lib().lzo_index_free(self.h)
            self.h = None


def hsps(t: TargetIndex, q: np.ndarray, p: Params, stats: Optional[Stats] = None) -> np.ndarray:
    """Kept ungapped HSPs of one (target, query-strand) tile as an (n,4) int32 array [s1,s2,len,score]."""
    q = np.ascontiguousarray(q, dtype=np.uint8)
    st = Stats()
    cap = 1 << 16
    while True:
        out = np.zeros((cap, 4), dtype=np.int32)
        st = Stats()
        n = lib().lzo_hsps_ix(t.h, t.codes.ctypes.data, len(t.codes), q.ctypes.data, len(q), C.byref(p), out.ctypes.data, cap, C.byref(st))
        if n >= 0:
            break
        cap *= 4
    if stats is not None:
        stats.add(st)
    return out[:n].copy()


def chain(h: np.ndarray) -> np.ndarray:
    """Best collinear chain; returns the member rows in canonical (s1,s2,len,score) order."""
    h = np.ascontiguousarray(h, dtype=np.int32).copy()
    if len(h) == 0:
        return h
    flag = np.zeros(len(h), dtype=np.uint8)
    lib().lzo_chain(h.ctypes.data, len(h), flag.ctypes.data)
    return h[flag.astype(bool)]


def align_tile(t: TargetIndex, q: np.ndarray, p: Params, stats: Optional[Stats] = None) -> np.ndarray:
    """Full pipeline for one tile-strand; (n,9) int32 rows [s1,e1,s2,e2,score,nmatch,ncols,a1,a2] (strand-local)."""
    q = np.ascontiguousarray(q, dtype=np.uint8)
    cap = 1 << 14
    while True:
        out = np.zeros((cap, 9), dtype=np.int32)
        st = Stats()
        n = lib().lzo_align_tile_ix(t.h, t.codes.ctypes.data, len(t.codes), q.ctypes.data, len(q), C.byref(p), out.ctypes.data, cap, C.byref(st))
        if n >= 0:
            break
        cap *= 4
    if stats is not None:
        stats.add(st)
    return out[:n].copy()


_FLIB = None


def faithful_lib():
    """oracle/lastz_faithful.c: the sequential, order-dependent second statement (measures the effect of deviations D1-D5)."""
    global _FLIB
    if _FLIB is None:
        so = os.path.join(HERE, '_build', 'liblastz_faithful.so')
        srcs = [os.path.join(HERE, 'lastz_faithful.c'), os.path.join(HERE, 'lastz_oracle.c')]
        if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(x) for x in srcs):
            subprocess.check_call(['make', '-C', HERE, '_build/liblastz_faithful.so'], stdout=subprocess.DEVNULL)
        l = C.CDLL(so)
        l.lzf_align_tile_ix.restype = C.c_long
        l.lzf_align_tile_ix.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.POINTER(Params), C.c_void_p, C.c_long, C.POINTER(Stats)]
        l.lzo_index_build.restype = C.c_void_p
        l.lzo_index_build.argtypes = [C.c_void_p, C.c_long]
        l.lzo_index_free.argtypes = [C.c_void_p]
        _FLIB = l
    return _FLIB


def align_tile_faithful(tcodes: np.ndarray, q: np.ndarray, p: Params, stats: Optional[Stats] = None) -> np.ndarray:
    """align_tile() through the sequential second statement (its own index handle: the two libraries do not share memory)."""
    l = faithful_lib()
    tcodes = np.ascontiguousarray(tcodes, dtype=np.uint8)
    q = np.ascontiguousarray(q, dtype=np.uint8)
    ix = l.lzo_index_build(tcodes.ctypes.data, len(tcodes))
    try:
        cap = 1 << 14
        while True:
            out = np.zeros((cap, 9), dtype=np.int32)
            st = Stats()
            n = l.lzf_align_tile_ix(ix, tcodes.ctypes.data, len(tcodes), q.ctypes.data, len(q), C.byref(p), out.ctypes.data, cap, C.byref(st))
            if n < cap:
                break
            cap *= 4
    finally:
        l.lzo_index_free(ix)
    if stats is not None:
        stats.add(st)
    return out[:n].copy()


def pct_str(nmatch: int, ncols: int) -> str:
    """LASTZ prints identity as '%.1f%%' of 100*n/d computed in double precision."""
    return '%.1f' % (100.0 * nmatch / ncols) if ncols else '0.0'


def lastz_general(tname: str, t: TargetIndex, qname: str, qcodes: np.ndarray, p: Params, stats: Optional[Stats] = None,
                  faithful: bool = False) -> List[str]:
    """The text `lastz T Q ... --format=general:... --markend --strand=both` writes (13 columns, '%' present)."""
    out = ['#name1\tstrand1\tstart1\tend1\tlength1\tname2\tstrand2\tstart2+\tend2+\tlength2\tscore\tidentity\tidPct\n']
    m = len(qcodes)
    for strand in '+-':
        q = qcodes if strand == '+' else revcomp_codes(qcodes)
        rows = align_tile_faithful(t.codes, q, p, stats) if faithful else align_tile(t, q, p, stats)
        for (s1, e1, s2, e2, score, nm, nc, _a1, _a2) in rows.tolist():
            if strand == '+':
                qs, qe = s2 + 1, e2
            else:
                qs, qe = m - e2 + 1, m - s2
            out.append('\t'.join(map(str, (tname, '+', s1 + 1, e1, e1 - s1, qname, strand, qs, qe, e2 - s2, score,
                                           f'{nm}/{nc}', pct_str(nm, nc) + '%'))) + '\n')
    out.append('# lastz end-of-file\n')
    return out


# ------------------------------------------------------------------------------------------ whole reference pipelines
def _pairs(anames: Sequence[str], bnames: Optional[Sequence[str]]):
    """get_all_pairs (utils.py:92-102) with sorted() instead of glob order."""
    return [(a, b) for a in sorted(anames) for b in sorted(bnames if bnames is not None else anames)]


def _pair_job(args):
    """One `lastz T Q ...` process of the reference's script (target table built per pair, both strands) -- pool worker."""
    a, tcodes, b, qcodes, hspthresh, kw = args
    kw = dict(kw)
    faithful = bool(kw.pop('faithful', False))
    st = Stats()
    lines = lastz_general(a, TargetIndex(tcodes), b, qcodes, default_params(hspthresh, **kw), st, faithful=faithful)
    return (a, b), lines, st.as_dict()


def general_all_pairs(agenome: Dict[str, np.ndarray], bgenome: Optional[Dict[str, np.ndarray]], hspthresh=3000, workers: int = 1,
                      stats: Optional[Stats] = None, **kw) -> Dict[Tuple[str, str], List[str]]:
    """LASTZ's text for every ordered (target, query) scaffold pair, the pairs spread over `workers` processes (the oracle
    itself stays single-threaded per pair, as LASTZ is)."""
    qg = agenome if bgenome is None else bgenome
    jobs = [(a, agenome[a], b, qg[b], hspthresh, kw) for a, b in _pairs(list(agenome), None if bgenome is None else list(bgenome))]
    if workers > 1 and len(jobs) > 1:
        from concurrent.futures import ProcessPoolExecutor
        jobs.sort(key=lambda j: -(len(j[1]) * len(j[3])))
        with ProcessPoolExecutor(max_workers=workers) as ex:
            res = list(ex.map(_pair_job, jobs, chunksize=1))
    else:
        res = [_pair_job(j) for j in jobs]
    out = {}
    for key, lines, st in res:
        out[key] = lines
        if stats is not None:
            for n, v in st.items():
                setattr(stats, n, getattr(stats, n) + v)
    return out


def mimeo_self(genome: Dict[str, np.ndarray], minIdt=60, minLen=100, minCov=3, intraCov=5, hspthresh=3000, strictSelf=False,
               label='Self_Repeat', prefix='Self_Repeat', stats: Optional[Stats] = None, workers: int = 1) -> Tuple[str, Optional[str], str]:
    """Reference pipeline of `mimeo self` (run_self.py:169-255 + wrappers.py:899-1271) on encoded scaffolds.
    Returns (tab text, intra tab text or None, gff3 text)."""
    general = general_all_pairs(genome, None, hspthresh, workers, stats)
    tab, intra = ao.TAB_HEADER, (ao.TAB_HEADER if strictSelf else None)
    for a, b in _pairs(list(genome), None):
        rows = ''.join(ao.filter_lastz_general(general[(a, b)], minLen, minIdt))
        if a == b and strictSelf:
            intra += rows
        else:
            tab += rows
    sizes = {n: len(c) for n, c in genome.items()}
    gff = ao.self_gff3(tab.splitlines(True), intra.splitlines(True) if strictSelf else None, sizes, minCov, intraCov, minLen, label, prefix)
    return tab, intra, gff


def mimeo_x(agenome, bgenome, minIdt=60, minLen=100, minCov=5, label='B_Repeat', prefix='B_Repeat', stats=None, workers: int = 1):
    """Reference pipeline of `mimeo x` (run_interspecies.py:173-258; hspthresh is always 3000 there)."""
    general = general_all_pairs(agenome, bgenome, 3000, workers, stats)
    tab = ao.TAB_HEADER
    for a, b in _pairs(list(agenome), list(bgenome)):
        tab += ''.join(ao.filter_lastz_general(general[(a, b)], minLen, minIdt))
    sizes = {n: len(c) for n, c in agenome.items()}
    return tab, ao.x_gff3(tab.splitlines(True), sizes, minCov, minLen, label, prefix)


def mimeo_map(agenome, bgenome, minIdt=90, minLen=100, hspthresh=3000, label='BHit', prefix='BHit', stats=None, workers: int = 1):
    """Reference pipeline of `mimeo map` without TRF (run_map.py:190-328)."""
    general = general_all_pairs(agenome, bgenome, hspthresh, workers, stats)
    tab = ao.TAB_HEADER
    for a, b in _pairs(list(agenome), list(bgenome)):
        tab += ''.join(ao.filter_lastz_general(general[(a, b)], minLen, minIdt))
    hits = ao.import_align_rows(tab.splitlines(True), prefix, minLen, minIdt)
    chrlens = sorted((n, str(len(c))) for n, c in agenome.items())
    return tab, ''.join(ao.write_gff_lines(hits, chrlens, label))


def general_rows(general: Dict[Tuple[str, str], List[str]], anames: Sequence[str], bnames: Sequence[str]):
    """The alignments of general_all_pairs() as the integer rows mb2_align reports:
    (t_id, q_id, strand, start1, end1, start2+, end2+, score, nmatch, ncols), ids = positions in anames / bnames."""
    ai = {n: k for k, n in enumerate(anames)}
    bi = {n: k for k, n in enumerate(bnames)}
    rows = set()
    for (a, b), lines in general.items():
        for ln in lines:
            if ln.startswith('#'):
                continue
            f = ln.rstrip('\n').split('\t')
            nm, nc = f[11].split('/')
            rows.add((ai[a], bi[b], 0 if f[6] == '+' else 1, int(f[2]), int(f[3]), int(f[7]), int(f[8]), int(f[10]), int(nm), int(nc)))
    return rows
