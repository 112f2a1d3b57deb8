"""
annot_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle; never on the product path).

Text-level restatement of the *annotation half* of mimeo's hot path, i.e. of the
shell commands that the reference generates and runs through bash:

  a-5   LASTZ 13-column rows -> 10-column .tab rows      wrappers.py:1044-1056 (665-675, 805-817, 1089-1101)
  a-9   BED projection + sort                            wrappers.py:1120-1128 (827-835, 1201-1220)
  a-10  bedtools genomecov -bg                           wrappers.py:1131-1138 (847-855, 1223-1231)
  a-11  awk '0+$4 >= cov'                                wrappers.py:1139-1141 (855-857, 1231-1233)
  a-12  sort + bedtools merge                            wrappers.py:1147-1150 (863-866, 1239-1250)
  a-13  minLen filter + GFF3 formatter                   wrappers.py:1153-1177 (870-894, 1253-1268)
  a-7   import_Align                                     wrappers.py:33-117
  a-8   writeGFFlines                                    wrappers.py:443-522

awk/sed/sort are restated for the C locale (byte collation), which is the only
locale in this image. bedtools (third-party, unpinned: environment.yml:8) is
restated from its published algorithm; the per-base loops live in
oracle/bedtools_oracle.c and are also mirrored here in numpy so this module works
without a compiler.

Pinning: every function here is checked in tests/ against (i) the hand-derived
known-answer vectors of SURVEY.md 9.3, (ii) golden files produced by executing
the reference's own generated script (real awk/sed/sort, bedtools = the C shim)
and the reference's own import_Align / writeGFFlines (tests/golden/make_golden.py).
"parity unpinned" remains true for the bedtools internals only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

import re
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

TAB_HEADER = '#name1\tstrand1\tstart1\tend1\tname2\tstrand2\tstart2+\tend2+\tscore\tidentity\n'
GFF_HEADER_SELF = '##gff-version 3\n#seqid\tsource\ttype\tstart\tend\tscore\tstrand\tphase\tattributes\n'

_NUM = re.compile(r'^[ \t]*([-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?))')


def awk_num(s: str) -> float:
    """awk's `0+$x`: value of the leading numeric prefix, 0 if there is none."""
    m = _NUM.match(s)
    return float(m.group(1)) if m else 0.0


def _sort_n(s: str) -> float:
    """GNU `sort -n` key value: leading blanks, optional '-', digits, optional fraction (no exponent)."""
    m = re.match(r'^[ \t]*(-?\d*\.?\d*)', s)
    t = m.group(1) if m else ''
    if t in ('', '-', '.', '-.'):
        return 0.0
    return float(t)


def _field_span(line: str, first: int, last: int) -> str:
    """Text of fields first..last (1-based, tab/blank separated the way sort -k does with default separators).
    All inputs in this pipeline are single-tab separated, so a tab split is exact."""
    f = line.split('\t')
    return '\t'.join(f[first - 1:last])


# --------------------------------------------------------------------------- a-5
def filter_lastz_general(lines13: Iterable[str], minLen, minIdt) -> List[str]:
    """LASTZ --format=general (13 columns incl. the two identity columns) -> sorted 10-column rows.

    Restates wrappers.py:1040-1056: sed 's/%//g'; drop '#' lines; keep 0+$5 >= minLen;
    keep 0+$13 >= minIdt; print $1,$2,$3,$4,$6,$7,$8,$9,$11,$13 tab-joined; sed 's/ //g';
    sort -k 1,1 -k 3n,4n. Returns rows WITH trailing newline.
    """
    out = []
    for raw in lines13:
        line = raw.rstrip('\n').replace('%', '')
        if line.startswith('#'):
            continue
        f = line.split()
        if not f:
            # awk would print an empty record; LASTZ never emits one. Ignored.
            continue
        g = lambda i: f[i - 1] if i <= len(f) else ''
        if not (awk_num(g(5)) >= float(minLen)):
            continue
        if not (awk_num(g(13)) >= float(minIdt)):
            continue
        row = '\t'.join(g(i) for i in (1, 2, 3, 4, 6, 7, 8, 9, 11, 13)).replace(' ', '')
        out.append(row)
    out.sort(key=lambda r: (_field_span(r, 1, 1).encode(), _sort_n(_field_span(r, 3, 4)), r.encode()))
    return [r + '\n' for r in out]


# --------------------------------------------------------------------------- a-9
def project_bed(tab_lines: Iterable[str]) -> List[str]:
    """awk -v OFS='\\t' '!/^#/ {print $1,$3,$4;}' | sed 's/%//g'   (wrappers.py:1120-1125)."""
    out = []
    for raw in tab_lines:
        line = raw.rstrip('\n')
        if line.startswith('#'):
            continue
        f = line.split()
        if not f:
            continue  # blank line: bedtools skips empty records
        g = lambda i: f[i - 1] if i <= len(f) else ''
        out.append(('\t'.join((g(1), g(3), g(4)))).replace('%', ''))
    return out


def sort_bed(rows: List[str]) -> List[str]:
    """sort -k 1,1 -k 2n,3n in the C locale; ties fall back to whole-line byte order (wrappers.py:1128)."""
    return sorted(rows, key=lambda r: (_field_span(r, 1, 1).encode(), _sort_n(_field_span(r, 2, 3)), r.encode()))


# --------------------------------------------------------------------------- a-10
def genomecov_bg_one(starts: np.ndarray, ends: np.ndarray, size: int) -> List[Tuple[int, int, int]]:
    """bedtools genomecov -bg for one chromosome, numpy mirror of
    ora_genomecov_bg() in bedtools_oracle.c. Returns [(start, end, depth)]."""
    if size <= 0:
        raise ValueError('chromosome size must be positive')
    st = np.zeros(size, dtype=np.int64)
    en = np.zeros(size, dtype=np.int64)
    s = np.asarray(starts, dtype=np.int64)
    e1 = np.asarray(ends, dtype=np.int64) - 1
    np.add.at(st, s[s < size], 1)
    inside = (e1 >= 0) & (e1 < size)
    np.add.at(en, e1[inside], 1)
    en[size - 1] += int((~inside).sum())
    # depth seen by the comparison at pos = cumsum(starts)[pos] - cumsum(ends)[pos-1]
    cs = np.cumsum(st)
    ce = np.concatenate(([0], np.cumsum(en)[:-1]))
    depth = (cs - ce) & 0xFFFFFFFF          # bedtools keeps depth in a uint32
    change = np.flatnonzero(np.concatenate(([True], depth[1:] != depth[:-1])))
    out = []
    for k, p in enumerate(change):
        d = int(depth[p])
        nxt = int(change[k + 1]) if k + 1 < len(change) else size
        if d > 0:
            out.append((int(p), nxt, d))
    return out


def genomecov_bg(sorted_bed: List[str], sizes: Dict[str, int]) -> List[str]:
    """Rows 'chrom\\tstart\\tend\\tdepth', chromosomes in order of first appearance."""
    out: List[str] = []
    cur = None
    ss: List[int] = []
    ee: List[int] = []

    def flush():
        if cur is None:
            return
        if cur not in sizes:
            raise ValueError(f'chromosome {cur!r} found in BED but not in genome file')
        for (a, b, d) in genomecov_bg_one(np.array(ss, dtype=np.int64), np.array(ee, dtype=np.int64), sizes[cur]):
            out.append(f'{cur}\t{a}\t{b}\t{d}')

    for n, row in enumerate(sorted_bed, 1):
        f = row.split('\t')
        if len(f) < 3 or not re.fullmatch(r'-?\d+', f[1]) or not re.fullmatch(r'-?\d+', f[2]):
            raise ValueError(f'malformed BED entry at line {n}')
        s, e = int(f[1]), int(f[2])
        if s < 0 or s > e:
            raise ValueError(f'malformed BED entry at line {n}. Start was greater than end (or negative).')
        if f[0] != cur:
            flush()
            cur, ss, ee = f[0], [], []
        ss.append(s)
        ee.append(e)
    flush()
    return out


# --------------------------------------------------------------------------- a-11 / a-12 / a-13
def threshold_rows(bg_rows: List[str], cov) -> List[str]:
    """awk '0+$4 >= cov {print ;}'  (wrappers.py:1139-1141)."""
    return [r for r in bg_rows if awk_num(r.split('\t')[3]) >= float(cov)]


def merge_bed(sorted_rows: List[str]) -> List[str]:
    """bedtools merge -i: fold while next.start <= cur.end within a chromosome."""
    out: List[str] = []
    cur = None
    cs = ce = 0
    for r in sorted_rows:
        f = r.split('\t')
        s, e = int(f[1]), int(f[2])
        if f[0] == cur and s <= ce:
            ce = max(ce, e)
        else:
            if cur is not None:
                out.append(f'{cur}\t{cs}\t{ce}')
            cur, cs, ce = f[0], s, e
    if cur is not None:
        out.append(f'{cur}\t{cs}\t{ce}')
    return out


def gff_rows(merged: List[str], minLen, source: str, label: str, prefix) -> List[str]:
    """awk '{ if($3 - $2 >= minLen) print ;}' | awk '...sprintf("%05d", i)...'  (wrappers.py:1163-1177)."""
    out = []
    i = 0
    for r in merged:
        f = r.split('\t')
        if awk_num(f[2]) - awk_num(f[1]) >= float(minLen):
            i += 1
            out.append('\t'.join((f[0], source, str(label), f[1], f[2], '.', '+', '.', f'ID={prefix}_{i:05d}')) + '\n')
    return out


def annotate_block(tab_lines: Iterable[str], sizes: Dict[str, int], cov, minLen, source, label, prefix) -> List[str]:
    """One coverage block (steps P,S,G,T,S,M,L,F of SURVEY.md 9.2) -> GFF3 feature rows."""
    bed = sort_bed(project_bed(tab_lines))
    bg = genomecov_bg(bed, sizes)
    kept = sort_bed(threshold_rows(bg, cov))
    merged = merge_bed(kept)
    return gff_rows(merged, minLen, source, label, prefix)


def self_gff3(tab_lines, intra_lines: Optional[Iterable[str]], sizes, minCov, intraCov, minLen, label, prefix) -> str:
    """Complete GFF3 text of `mimeo self` (wrappers.py:1106-1268): header, inter block, optional intra block."""
    text = GFF_HEADER_SELF + ''.join(annotate_block(tab_lines, sizes, minCov, minLen, 'mimeo-self', label, prefix))
    if intra_lines is not None:
        text += ''.join(annotate_block(intra_lines, sizes, intraCov, minLen, 'mimeo-self', str(label) + '_intra', prefix))
    return text


def x_gff3(tab_lines, sizes, minCov, minLen, label, prefix) -> str:
    """Complete GFF3 text of `mimeo x` (wrappers.py:822-894)."""
    return GFF_HEADER_SELF + ''.join(annotate_block(tab_lines, sizes, minCov, minLen, 'mimeo', label, prefix))


# --------------------------------------------------------------------------- array-level mirror
def coverage_segments_arrays(chrom: np.ndarray, start: np.ndarray, end: np.ndarray, sizes: Sequence[int],
                             cov: int, minLen: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Steps G,T,M,L on integer arrays; chromosomes by index. numpy mirror of ora_coverage_segments()."""
    chrom = np.asarray(chrom)
    oc: List[int] = []
    os_: List[int] = []
    oe: List[int] = []
    if len(chrom):
        if chrom.min() < 0 or chrom.max() >= len(sizes) or (np.asarray(start) < 0).any() or (np.asarray(start) > np.asarray(end)).any():
            raise ValueError('invalid hit')
    for c in range(len(sizes)):
        m = chrom == c
        if not m.any():
            continue
        rows = genomecov_bg_one(np.asarray(start)[m], np.asarray(end)[m], int(sizes[c]))
        cs = ce = None
        for (a, b, d) in rows:
            if d < cov:
                continue
            if cs is not None and a <= ce:
                ce = max(ce, b)
            else:
                if cs is not None and ce - cs >= minLen:
                    oc.append(c); os_.append(cs); oe.append(ce)
                cs, ce = a, b
        if cs is not None and ce - cs >= minLen:
            oc.append(c); os_.append(cs); oe.append(ce)
    return (np.array(oc, dtype=np.int32), np.array(os_, dtype=np.int32), np.array(oe, dtype=np.int32))


def coverage_segments_c(chrom: np.ndarray, start: np.ndarray, end: np.ndarray, sizes, cov: int, minLen: int):
    """The same steps by the C restatement (oracle/bedtools_oracle.c: ora_coverage_segments), for inputs of millions of hits."""
    import ctypes
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(here, '_build', 'libannot_oracle.so')
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(here, 'bedtools_oracle.c')):
        subprocess.check_call(['make', '-C', here, '_build/libannot_oracle.so'], stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(so)
    lib.ora_coverage_segments.restype = ctypes.c_long
    chrom = np.ascontiguousarray(chrom, dtype=np.int32); start = np.ascontiguousarray(start, dtype=np.int32)
    end = np.ascontiguousarray(end, dtype=np.int32); sizes = np.ascontiguousarray(sizes, dtype=np.int64)
    cap = len(chrom) + 8
    oc, os_, oe = (np.zeros(cap, np.int32) for _ in range(3))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    k = lib.ora_coverage_segments(P(chrom), P(start), P(end), ctypes.c_long(len(chrom)), P(sizes), ctypes.c_int(len(sizes)),
                                  ctypes.c_int(cov), ctypes.c_int(minLen), P(oc), P(os_), P(oe), ctypes.c_long(cap))
    if k < 0:
        raise ValueError('invalid hit')
    return oc[:k].copy(), os_[:k].copy(), oe[:k].copy()


# --------------------------------------------------------------------------- a-7 / a-8 (map)
def import_align_rows(tab_lines: Iterable[str], prefix, minLen=100, minIdt=95) -> List[Dict[str, str]]:
    """Restates import_Align (wrappers.py:66-115) without pandas: filter int(end)-int(start) >= minLen and
    float(id) >= minIdt; stable sort by (tName, tStart, tEnd, tStrand) AS STRINGS; UID = prefix_<zero-filled row>."""
    hits = []
    for line in tab_lines:
        li = line.strip()
        if not li.startswith('#'):
            f = li.split()
            if int(f[3]) - int(f[2]) >= minLen and float(f[9]) >= minIdt:
                hits.append(dict(tName=f[0], tStrand=f[1], tStart=f[2], tEnd=f[3], qName=f[4], qStrand=f[5],
                                 qStart=f[6], qEnd=f[7], score=f[8], pID=f[9], UID=None))
    if not hits:
        raise SystemExit(1)
    hits.sort(key=lambda h: (h['tName'], h['tStart'], h['tEnd'], h['tStrand']))
    fill = len(str(len(hits)))
    for i, h in enumerate(hits, 1):
        h['UID'] = (str(prefix) if prefix else 'BHit') + '_' + str(i).zfill(fill)
    return hits


def write_gff_lines(hits: List[Dict[str, str]], chrlens: Optional[List[Tuple[str, str]]], ftype='BHit') -> List[str]:
    """Restates writeGFFlines (wrappers.py:469-522)."""
    out = ['##gff-version 3\n']
    if chrlens:
        for name, maxlen in chrlens:
            out.append(' '.join(['##sequence-region', str(name), '1', str(maxlen) + '\n']))
    out.append('\t'.join(['##seqid', 'source', 'type', 'start', 'end', 'score', 'strand', 'phase', 'attributes' + '\n']))
    for h in hits:
        attributes = ';'.join(['ID=' + h['UID'], 'identity=' + str(h['pID']),
                               'B_locus=' + h['qName'] + '_' + h['qStrand'] + '_' + str(h['qStart']) + '_' + str(h['qEnd'])])
        out.append('\t'.join([h['tName'], 'mimeo-map', ftype, str(h['tStart']), str(h['tEnd']), str(h['score']),
                              h['tStrand'], '.', attributes + '\n']))
    return out
