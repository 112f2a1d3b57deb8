"""
engine.py -- the replacement of mimeo's generated bash script (wrappers.py:899-1271, 683-896, 525-680
executed by utils.run_cmd, utils.py:213-254): the same stages, run through libmimeo_b200 on the GPU.

    align_pairs()      = every `lastz T Q ...` + sed/awk/awk/awk/sed/sort >> outtab     (wrappers.py:1015-1104)
    coverage_to_gff()  = awk BED | sort | genomecov | awk cov | sort | merge | awk GFF  (wrappers.py:1106-1177)

Text in / text out is byte-compatible with the reference (10-column .tab, GFF3); everything between
is device arrays. There is no CPU fallback: without the CUDA library these functions raise.
"""
from __future__ import annotations

import logging
import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import align as _align
from . import coverage as _coverage
from .fasta import read_fasta
from .genome import Genome, align_params

TAB_HEADER = '#name1\tstrand1\tstart1\tend1\tname2\tstrand2\tstart2+\tend2+\tscore\tidentity\n'
GFF_HEADER = '##gff-version 3\n#seqid\tsource\ttype\tstart\tend\tscore\tstrand\tphase\tattributes\n'


def c_sorted(names: Iterable[str]) -> List[str]:
    """`sort -k 1,1` order in the C locale (byte order)."""
    return sorted(names, key=lambda s: s.encode())


# ------------------------------------------------------------------------------------------ alignment -> .tab
def load_files(paths: Sequence[str]) -> Tuple[List[str], List[np.ndarray], Dict[str, int]]:
    """One scaffold per file as LASTZ sees it: name = first word of the first header. Returns names, seqs, path->index."""
    names, seqs, index = [], [], {}
    for p in paths:
        recs = read_fasta(p)
        if not recs:
            raise RuntimeError(f'no FASTA record in {p}')
        if len(recs) > 1:
            logging.warning('%s holds %d records; only the first is aligned (mimeo expects one per file)', p, len(recs))
        index[p] = len(names)
        names.append(recs[0][0])
        seqs.append(recs[0][2])
    return names, seqs, index


def align_genomes(tnames, tseqs, qnames, qseqs, hspthresh=3000, same=False, minLen=None, minIdt=None):
    """Device genomes + all-pairs alignment. Returns (hits dict, stats dict). With minLen / minIdt the rows are filtered and
    sorted on the device (`mb2_filter_sort`) and only the survivors are returned."""
    T = Genome(tnames, tseqs)
    Q = T if same else Genome(qnames, qseqs)
    try:
        if minLen is None:
            return _align.align(T, Q, align_params(hspthresh))
        dh = _align.align_device(T, Q, align_params(hspthresh))
        try:
            dh.filter_sort(minLen, minIdt)
            return dh.download()
        finally:
            dh.close()
    finally:
        if Q is not T:
            Q.close()
        T.close()


def align_pairs(pairs: Sequence[Tuple[str, str]], outtab: str, minIdt, minLen, hspthresh=3000,
                outtab_intra: Optional[str] = None) -> Dict[str, int]:
    """All (target file, query file) pairs -> filtered, per-pair sorted rows appended to outtab in pair order.
    With outtab_intra, pairs whose two paths are identical go there instead (--strictSelf, wrappers.py:1016)."""
    tpaths = list(dict.fromkeys(a for a, _ in pairs))
    qpaths = list(dict.fromkeys(b for _, b in pairs))
    same = tpaths == qpaths
    tnames, tseqs, tidx = load_files(tpaths)
    if same:
        qnames, qseqs, qidx = tnames, tseqs, tidx
    else:
        qnames, qseqs, qidx = load_files(qpaths)
    hits, stats = align_genomes(tnames, tseqs, qnames, qseqs, hspthresh, same, minLen, minIdt)     # filtered on the device
    blocks = _align.tab_blocks(hits, tnames, qnames, minLen, minIdt)                                  # text + whole-line tie-break
    with open(outtab, 'a') as ft:
        fi = open(outtab_intra, 'a') if outtab_intra else None
        try:
            for a, b in pairs:
                rows = blocks.get((tidx[a], qidx[b]))
                if rows:
                    (fi if (fi is not None and a == b) else ft).write(''.join(rows))
        finally:
            if fi is not None:
                fi.close()
    return stats


def filter_hits(hits: Dict[str, np.ndarray], minLen, minIdt) -> np.ndarray:
    """Boolean mask of the rows the reference's awk filters keep (wrappers.py:1049-1052): length1 >= minLen and the
    PRINTED identity ('%.1f') >= minIdt. Vectorised; rows within 1e-6 of a rounding tie are decided by real formatting."""
    n = len(hits['t_id'])
    if n == 0:
        return np.zeros(0, dtype=bool)
    keep = (hits['end1'].astype(np.int64) - hits['start1'] + 1) >= minLen
    nm, nc = hits['nmatch'].astype(np.float64), hits['ncols'].astype(np.float64)
    t = 1000.0 * nm / np.maximum(nc, 1.0)
    tenths = np.floor(t + 0.5)
    near_tie = np.abs((t + 0.5) - np.round(t + 0.5)) < 1e-6
    for k in np.flatnonzero(near_tie & keep):
        tenths[k] = round(float(_align.pct_text(int(hits['nmatch'][k]), int(hits['ncols'][k]))) * 10)
    return keep & (tenths >= 10.0 * float(minIdt))


def filter_hits_map(hits: Dict[str, np.ndarray], minLen, minIdt) -> np.ndarray:
    """`mimeo map`: the rows that survive BOTH the awk filter after LASTZ (filter_hits) and import_Align's own test
    (wrappers.py:76): int(end1) - int(start1) >= minLen, i.e. one base stricter than length1, and float(identity) >= minIdt."""
    keep = filter_hits(hits, minLen, minIdt)
    if len(keep):
        keep &= (hits['end1'].astype(np.int64) - hits['start1']) >= minLen
    return keep


def self_segments(T: Genome, T_both: Optional[Genome], sizes: Sequence[int], minIdt, minLen, minCov, intraCov, hspthresh=3000,
                  strictSelf=True):
    """`mimeo self` without any text: device genome in, (inter segments, intra segments or None, kept hits, stats) out.
    The hit table stays in HBM from the alignment through the filter (`mb2_filter_sort`) to both coverage passes; only the
    surviving rows come back (for the .tab text). Scaffold index order must already be the C-locale name order (it defines
    the GFF row order)."""
    dh = _align.align_device(T, T, align_params(hspthresh), Q_aux=T_both)
    try:
        dh.filter_sort(minLen, minIdt)
        inter = dh.coverage(1 if strictSelf else 0, sizes, minCov, minLen)
        intra = dh.coverage(2, sizes, intraCov, minLen) if strictSelf else None
        hits, stats = dh.download()
    finally:
        dh.close()
    return inter, intra, hits, stats


# ------------------------------------------------------------------------------------------ .tab -> GFF3
def read_lens(path: str) -> Dict[str, int]:
    sizes = {}
    with open(path) as f:
        for line in f:
            p = line.rstrip('\n').split('\t')
            if len(p) >= 2 and p[0]:
                sizes[p[0]] = int(p[1])
    return sizes


def parse_tab_hits(path: str, nthreads: int = 0):
    """Columns 1,3,4 of every non-'#' line (what awk '!/^#/ {print $1,$3,$4;}' projects, wrappers.py:1121), parsed natively
    (libmimeo_b200 `mb2_tab_project`: mmap + threads). Returns (distinct names in order of first appearance, int32 name
    index per row, start int64, end int64)."""
    import ctypes as C
    from . import _lib
    t = _lib.TabHits()
    try:
        _lib.check(_lib.lib().mb2_tab_project(os.fsencode(path), int(nthreads), C.byref(t)))
    except _lib.Mb2Error as e:
        raise RuntimeError(f'malformed alignment table: {e}') from None
    try:
        n = int(t.n)
        names = [t.names[k].decode('utf-8', 'replace') for k in range(int(t.nnames))]
        if n == 0:
            return names, np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros(0, np.int64)
        ids = np.ctypeslib.as_array(t.chrom, shape=(n,)).copy()
        start = np.ctypeslib.as_array(t.start, shape=(n,)).copy()
        end = np.ctypeslib.as_array(t.end, shape=(n,)).copy()
    finally:
        _lib.lib().mb2_free_tab_hits(C.byref(t))
    return names, ids, start, end


def coverage_rows(tab_path: str, sizes: Dict[str, int], cov, minLen, source: str, label: str, prefix) -> List[str]:
    """GFF3 feature rows of one coverage block."""
    names, ids, start, end = parse_tab_hits(tab_path)
    if not len(ids):
        return []
    names = [n.replace('%', '') for n in names]          # sed 's/%//g' on the projected columns (wrappers.py:1125)
    order = c_sorted(set(names))
    missing = [n for n in order if n not in sizes]
    if missing:
        raise RuntimeError(f'chromosome {missing[0]!r} found in {tab_path} but not in the genome length file')
    if (start < 0).any() or (start > end).any() or (end > 0x7fffffff).any():
        raise RuntimeError(f'malformed hit in {tab_path}: start must be >= 0 and <= end')
    idx = {n: i for i, n in enumerate(order)}
    chrom = np.asarray([idx[n] for n in names], dtype=np.int32)[ids]
    c, s, e = _coverage.coverage_segments(chrom, start.astype(np.int32), end.astype(np.int32), [sizes[n] for n in order],
                                          int(cov), int(minLen))
    return segment_gff_text(c, s, e, order, source, label, prefix).splitlines(keepends=True)


def segment_gff_text(chrom, start, end, names: Sequence[str], source: str, label: str, prefix, first_id: int = 1) -> str:
    """GFF3 feature rows of one coverage block as one string: the awk formatter that ends the block in the reference's script
    (wrappers.py:1166-1173), `ID=<prefix>_%05d` counted from first_id. Formatted natively (`mb2_format_gff`)."""
    import ctypes as C
    from . import _lib
    n = len(chrom)
    if n == 0:
        return ''
    c, s, e = (np.ascontiguousarray(a, dtype=np.int32) for a in (chrom, start, end))
    arr = (C.c_char_p * len(names))(*[x.encode() for x in names])
    t = _lib.Text()
    _lib.check(_lib.lib().mb2_format_gff(c.ctypes.data, s.ctypes.data, e.ctypes.data, n, arr, len(names), str(source).encode(),
                                         str(label).encode(), str(prefix).encode(), int(first_id), 0, C.byref(t)))
    try:
        return C.string_at(t.text, int(t.nbytes)).decode('utf-8', 'replace')
    finally:
        _lib.lib().mb2_free_text(C.byref(t))


def coverage_to_gff(tab_path: str, lens_path: str, outgff: str, cov, minLen, source: str, label: str, prefix,
                    write_header: bool = True) -> int:
    rows = coverage_rows(tab_path, read_lens(lens_path), cov, minLen, source, label, prefix)
    with open(outgff, 'w' if write_header else 'a') as f:
        if write_header:
            f.write(GFF_HEADER)
        f.write(''.join(rows))
    return len(rows)
