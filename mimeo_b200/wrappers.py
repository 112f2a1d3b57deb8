"""
Host mirror of mimeo.wrappers (src/mimeo/wrappers.py) for the alignment-to-annotation path: identical function
names and signatures. The three command generators return engine operations (strings understood by
mimeo_b200.utils.run_cmd) instead of bash lines; stage order, outputs and byte formats are the reference's.
import_Align / writeGFFlines keep the pandas contract of the reference (wrappers.py:33-117, 443-522).
"""
from __future__ import annotations

import json
import logging
import os
import sys
from typing import List, Tuple

import pandas as pd

from .engine import GFF_HEADER, TAB_HEADER
from .utils import OP_PREFIX, run_cmd  # noqa: F401  (re-exported like the reference)


def _op(**kw) -> str:
    return OP_PREFIX + json.dumps(kw)


def import_Align(infile: str = None, prefix: str = None, minLen: int = 100, minIdt: float = 95) -> pd.DataFrame:
    """.tab -> DataFrame[tName,tStrand,tStart,tEnd,qName,qStrand,qStart,qEnd,score,pID,UID], all strings.
    Keeps rows with int(end)-int(start) >= minLen and float(identity) >= minIdt, sorts by the STRING values of
    (tName,tStart,tEnd,tStrand), numbers rows from 1 zero-filled to the width of the row count. Exits 1 if empty."""
    cols = ['tName', 'tStrand', 'tStart', 'tEnd', 'qName', 'qStrand', 'qStart', 'qEnd', 'score', 'pID']
    kept = []
    with open(infile) as f:
        for raw in f:
            li = raw.strip()
            if li.startswith('#'):
                continue
            p = li.split()
            if int(p[3]) - int(p[2]) >= minLen and float(p[9]) >= minIdt:
                kept.append(p[:10])
    if not kept:
        logging.warning('No alignments found in %s' % infile)
        sys.exit(1)
    df = pd.DataFrame(kept, columns=cols)
    df['UID'] = None
    df = df.sort_values(['tName', 'tStart', 'tEnd', 'tStrand'], ascending=True, kind='stable').reset_index(drop=True)
    df.index = df.index + 1
    width = len(str(len(df)))
    stem = str(prefix) if prefix else 'BHit'
    df['UID'] = [stem + '_' + str(i).zfill(width) for i in df.index]
    return df


def writeGFFlines(alnDF: pd.DataFrame = None, chrlens: List[Tuple[str, str]] = None, ftype: str = 'BHit'):
    """GFF3 text of `mimeo map`, one line per yield (wrappers.py:443-522)."""
    yield '##gff-version 3\n'
    for name, maxlen in (chrlens or []):
        yield f'##sequence-region {name} 1 {maxlen}\n'
    yield '##seqid\tsource\ttype\tstart\tend\tscore\tstrand\tphase\tattributes\n'
    for r in alnDF.itertuples(index=False):
        attrs = f'ID={r.UID};identity={r.pID};B_locus={r.qName}_{r.qStrand}_{r.qStart}_{r.qEnd}'
        yield '\t'.join([r.tName, 'mimeo-map', ftype, str(r.tStart), str(r.tEnd), str(r.score), r.tStrand, '.', attrs]) + '\n'


def map_gff_text(infile: str, prefix: str = None, minLen: int = 100, minIdt: float = 95, ftype: str = 'BHit', nthreads: int = 0):
    """(number of rows, GFF3 feature rows) of `mimeo map` straight from the .tab file: the native equivalent of
    import_Align + writeGFFlines (same filter, same string sort, same UIDs, same bytes) without building a DataFrame."""
    import ctypes as C
    from . import _lib
    t = _lib.Text()
    try:
        _lib.check(_lib.lib().mb2_map_gff(os.fsencode(infile), (str(prefix) if prefix else '').encode(), float(minLen), float(minIdt),
                                          str(ftype).encode(), int(nthreads), C.byref(t)))
    except _lib.Mb2Error as e:
        raise RuntimeError(f'malformed alignment table: {e}') from None
    try:
        return int(t.nrows), C.string_at(t.text, int(t.nbytes)).decode('utf-8', 'surrogateescape')
    finally:
        _lib.lib().mb2_free_text(C.byref(t))


def write_map_gff(infile: str, gffout: str = None, chrlens: List[Tuple[str, str]] = None, prefix: str = None, minLen: int = 100,
                  minIdt: float = 95, ftype: str = 'BHit') -> int:
    """What run_map.main does with import_Align + writeGFFlines (run_map.py:272-290), on the native path: exit 1 if no
    alignment passes the filters, else write the GFF3 (if a path is given). Returns the number of features."""
    n, body = map_gff_text(infile, prefix, minLen, minIdt, ftype)
    if n == 0:
        logging.warning('No alignments found in %s' % infile)
        sys.exit(1)
    if gffout:
        with open(gffout, 'w', encoding='utf-8', errors='surrogateescape') as f:
            f.write('##gff-version 3\n')
            for name, maxlen in (chrlens or []):
                f.write(f'##sequence-region {name} 1 {maxlen}\n')
            f.write('##seqid\tsource\ttype\tstart\tend\tscore\tstrand\tphase\tattributes\n')
            f.write(body)
    return n


def map_LZ_cmds(lzpath: str = 'lastz', pairs: List[Tuple[str, str]] = None, minIdt: float = 95, minLen: int = 100,
                hspthresh: int = 3000, outfile: str = None, verbose: bool = False,
                lastz_format: str = 'general:name1,strand1,start1,end1,length1,name2,strand2,start2+,end2+,length2,score,identity',
                step_size: int = 1, strand_mode: str = 'both', chain: bool = True, gapped: bool = True) -> List[str]:
    """Operations for `mimeo map`: header, then align+filter+sort of every pair into outfile (wrappers.py:525-680)."""
    if not pairs:
        raise ValueError('No sequence pairs provided for alignment')
    if outfile is None:
        raise ValueError('Output file path is required')
    if step_size != 1 or strand_mode != 'both' or not chain or not gapped:
        raise ValueError('the GPU engine implements the option set mimeo uses: --step=1 --strand=both --chain --gapped')
    return [_op(op='write', path=outfile, text=TAB_HEADER),
            _op(op='align', pairs=[list(p) for p in pairs], outtab=outfile, minIdt=minIdt, minLen=minLen, hspthresh=hspthresh)]


def _coverage_ops(outtab, outgff, AchrmLens, cov, minLen, source, label, prefix, write_header):
    return [_op(op='coverage', tab=outtab, lens=AchrmLens, outgff=outgff, cov=cov, minLen=minLen, source=source,
                label=str(label), prefix=str(prefix), write_header=write_header)]


def xspecies_LZ_cmds(lzpath: str = 'lastz', bdtlsPath: str = 'bedtools', Adir: str = None, Bdir: str = None,
                     pairs: List[Tuple[str, str]] = None, outtab: str = None, outgff: str = None, minIdt: float = 60,
                     minLen: int = 100, hspthresh: int = 3000, minCov: int = 5, AchrmLens: str = None, reuseTab: bool = False,
                     label: str = 'B_repeats', prefix: str = None, verbose: bool = False) -> List[str]:
    """Operations for `mimeo x` (wrappers.py:683-896): alignment block unless an existing outtab is recycled,
    then one coverage block with minCov, GFF source column `mimeo`."""
    cmds: List[str] = []
    if not reuseTab or not os.path.isfile(outtab):
        cmds.append(_op(op='write', path=outtab, text=TAB_HEADER))
        if pairs:
            cmds.append(_op(op='align', pairs=[list(p) for p in pairs], outtab=outtab, minIdt=minIdt, minLen=minLen, hspthresh=hspthresh))
    cmds.append(_op(op='echo', text='Generate non-zero coverage scores for target genome regions, filter for min coverage of x'))
    cmds += _coverage_ops(outtab, outgff, AchrmLens, minCov, minLen, 'mimeo', label, prefix, True)
    return cmds


def self_LZ_cmds(lzpath: str = 'lastz', bdtlsPath: str = 'bedtools', splitSelf: bool = False, Adir: str = None, Bdir: str = None,
                 pairs: List[Tuple[str, str]] = None, outtab: str = None, outgff: str = None, minIdt: float = 60,
                 minLen: int = 100, hspthresh: int = 3000, minCov: int = 3, intraCov: int = 5, AchrmLens: str = None,
                 reuseTab: bool = False, label: str = 'Self_repeats', prefix: str = None, verbose: bool = False) -> List[str]:
    """Operations for `mimeo self` (wrappers.py:899-1271). With splitSelf, same-file pairs go to `<outtab>_intra.tab`
    and get their own coverage block (intraCov, type `<label>_intra`, IDs restarting at 00001) appended to the GFF."""
    cmds: List[str] = []
    outtab_intra = outtab + '_intra.tab' if splitSelf else None
    if not reuseTab or not os.path.isfile(outtab):
        cmds.append(_op(op='write', path=outtab, text=TAB_HEADER))
        if splitSelf:
            cmds.append(_op(op='write', path=outtab_intra, text=TAB_HEADER))
        if pairs:
            cmds.append(_op(op='align', pairs=[list(p) for p in pairs], outtab=outtab, minIdt=minIdt, minLen=minLen, hspthresh=hspthresh,
                            outtab_intra=outtab_intra))
    cmds.append(_op(op='echo', text='Coverage filtering for BETWEEN chromosome hits (or all if not in selfSplit mode)'))
    cmds += _coverage_ops(outtab, outgff, AchrmLens, minCov, minLen, 'mimeo-self', label, prefix, True)
    if splitSelf:
        if reuseTab and not os.path.isfile(outtab_intra) and os.path.isfile(outtab):
            logging.warning("Warning: Could not find intra-chrom results file: %s \nRe-run in '--strictSelf' mode if required." % outtab_intra)
        else:
            cmds.append(_op(op='echo', text='Applying separate coverage filtering for WITHIN chromosome hits'))
            cmds += _coverage_ops(outtab_intra, outgff, AchrmLens, intraCov, minLen, 'mimeo-self', str(label) + '_intra', prefix, False)
    return cmds
