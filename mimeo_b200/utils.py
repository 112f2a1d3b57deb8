"""
Host mirror of mimeo.utils (src/mimeo/utils.py): same function names, arguments and error behaviour for the
pieces the hot path touches. The process-exec layer (`run_cmd`, utils.py:213-254) no longer writes a bash
script: it executes the engine operations emitted by mimeo_b200.wrappers in-process on the GPU.
"""
from __future__ import annotations

import json
import logging
import os
import shutil
import sys
import tempfile
from collections import Counter
from datetime import datetime, timezone
from typing import List, Optional, Tuple

from . import fasta

OP_PREFIX = 'mb2:'


def get_all_pairs(Adir: Optional[str] = None, Bdir: Optional[str] = None) -> List[Tuple[str, str]]:
    """All ordered (A file, B file) pairs; A x A when only Adir is given (utils.py:65-106).
    Files are taken in sorted order (the reference uses glob order, which is filesystem dependent)."""
    def files(d):
        return [os.path.join(d, f) for f in sorted(os.listdir(d))]
    if Adir and Bdir:
        return [(a, b) for a in files(Adir) for b in files(Bdir)]
    if Adir:
        logging.info('Compose self-genome alignment pairs.')
        fs = files(Adir)
        return [(a, b) for a in fs for b in fs]
    logging.error('Need at least one seq directory to compose alignment pairs.')
    sys.exit(1)


def import_pairs(file: str = None, Adir: str = None, Bdir: str = None) -> List[Tuple[str, str]]:
    """Pairs from a two-column file, '#' lines skipped (utils.py:31-62)."""
    pairs = []
    with open(file) as f:
        for line in f:
            li = line.strip()
            if li and not li.startswith('#'):
                a, b = li.split()[:2]
                pairs.append((os.path.join(Adir, a), os.path.join(Bdir, b)))
    return pairs


def run_cmd(cmds: List[str], verbose: bool = False, keeptemp: bool = False) -> None:
    """Execute a command list produced by mimeo_b200.wrappers (utils.py:213-254).

    Like the reference this runs inside a fresh temp directory created in the cwd (removed unless keeptemp);
    unlike the reference nothing is spawned: every command is an engine operation (`mb2:{json}`) dispatched to
    the CUDA library. A failing operation raises RuntimeError (the reference only noticed the last command's
    exit status, utils.py:194-210)."""
    from . import engine
    tmpdir = tempfile.mkdtemp(prefix='tmp.', dir=os.getcwd())
    try:
        for cmd in cmds:
            if not cmd.startswith(OP_PREFIX):
                raise RuntimeError(f'not an engine operation (this build never shells out): {cmd[:80]}')
            op = json.loads(cmd[len(OP_PREFIX):])
            kind = op.pop('op')
            logging.debug('engine op %s %s', kind, op)
            try:
                if kind == 'write':
                    with open(op['path'], 'w') as f:
                        f.write(op['text'])
                elif kind == 'align':
                    stats = engine.align_pairs([tuple(p) for p in op['pairs']], op['outtab'], op['minIdt'], op['minLen'],
                                               op['hspthresh'], op.get('outtab_intra'))
                    if verbose:
                        logging.info('alignment stage counters: %s', stats)
                elif kind == 'coverage':
                    engine.coverage_to_gff(op['tab'], op['lens'], op['outgff'], op['cov'], op['minLen'], op['source'], op['label'],
                                           op['prefix'], op['write_header'])
                elif kind == 'echo':
                    logging.debug(op['text'])
                else:
                    raise RuntimeError(f'unknown engine operation {kind!r}')
            except SystemExit:
                raise
            except Exception as e:
                print('The following operation failed:', file=sys.stderr, flush=True)
                print(cmd[:500], file=sys.stderr, flush=True)
                raise RuntimeError(f'Error running engine operation {kind}: {e}') from e
    finally:
        if not keeptemp:
            shutil.rmtree(tmpdir, ignore_errors=True)


def getTimestring() -> str:
    """UTC now as YYYYMMDDHHMMSSmmm (utils.py:257-271)."""
    now = datetime.now(timezone.utc)
    return now.strftime('%Y%m%d%H%M%S') + '%03d' % (now.microsecond // 1000)


def splitFasta(infile: str, outdir: str, unique: bool = True) -> None:
    """One `<id>.fa` per record (utils.py:274-309); exits on duplicate ids when unique (the records before the repeated
    one are written first, as in the reference). Read, wrapped at 60 columns and written natively (`mb2_fasta_split`)."""
    import ctypes as C
    from . import _lib
    n = C.c_uint64(0)
    rc = _lib.lib().mb2_fasta_split(os.fsencode(infile), os.fsencode(outdir), 1 if unique else 0, 60, 0, C.byref(n))
    if rc == _lib.ERR_DUPLICATE_ID:
        logging.error('Non-unique name in genome: %s. Quitting.' % _lib.lib().mb2_last_error().decode('utf-8', 'replace'))
        sys.exit(1)
    _lib.check(rc)


def isfile(path: str) -> str:
    path = os.path.abspath(path)
    if not os.path.isfile(path):
        logging.error('Input file not found: %s' % path)
        sys.exit(1)
    return path


def set_paths(adir=None, bdir=None, afasta=None, bfasta=None, outdir=None, outtab=None, gffout=None,
              suppresBdir: bool = False, runtrf=None):
    """Directory / file layout of a run (utils.py:339-469); returns (adir, bdir, outdir, outtab, gffout, tempdir)."""
    need_temp = (not adir) or (not bdir and not suppresBdir) or bool(runtrf)
    tempdir = None
    if need_temp:
        tempdir = os.path.join(os.getcwd(), 'temp_' + getTimestring())
        os.makedirs(tempdir)

    def prepare(d, fa, sub, which):
        if d:
            d = os.path.abspath(d)
            if not os.path.isdir(d):
                logging.info('Creating %sdir: %s' % (which, d))
                os.makedirs(d)
                if not fa:
                    logging.error('No %s-genome fasta file provided. Quitting.' % which)
                    sys.exit(1)
        else:
            d = os.path.join(tempdir, sub)
            os.makedirs(d)
        return d

    adir = prepare(adir, afasta, 'A_genome_split', 'A')
    if bdir or not suppresBdir:
        bdir = prepare(bdir, bfasta, 'B_genome_split', 'B')
    if afasta:
        if os.path.isfile(afasta):
            splitFasta(afasta, adir)
        else:
            logging.error('A-genome fasta not found at path: %s' % afasta)
    if bfasta:
        if os.path.isfile(bfasta):
            splitFasta(bfasta, bdir)
        elif not suppresBdir:
            logging.error('B-genome fasta not found at path: %s' % bfasta)
    if outdir:
        outdir = os.path.abspath(outdir)
        if not os.path.isdir(outdir):
            logging.info('Create output directory: %s' % outdir)
            os.makedirs(outdir)
    else:
        outdir = os.getcwd()
    if outtab:
        outtab = os.path.join(outdir, outtab)
        if os.path.isfile(outtab):
            logging.info('Previous alignment found: %s' % outtab)
    if gffout:
        gffout = os.path.join(outdir, gffout)
    return adir, bdir, outdir, outtab, gffout, tempdir


def checkUniqueID(records: List) -> None:
    """Exit if two records share an id (utils.py:472-499); records need an `.id` attribute or be (id, ...) tuples."""
    ids = [getattr(r, 'id', r[0] if isinstance(r, tuple) else r) for r in records]
    dup = [k for k, v in Counter(ids).items() if v > 1]
    if dup:
        logging.error(f'Input sequence IDs not unique:\n{dup}\n\nQuitting.')
        sys.exit(1)


def chromlens(seqDir: str = None, outfile: Optional[str] = None) -> List[Tuple[str, str]]:
    """[(id, str(length))] of every record in a directory, sorted by id; optionally written as `id\\tlen` lines
    (utils.py:502-557)."""
    recs = [(rid, len(seq)) for rid, _h, seq, _p in fasta.dir_records(seqDir)]
    if not recs:
        logging.error('No sequences found in %s \n Cannot calculate seq lengths.' % seqDir)
        sys.exit(1)
    checkUniqueID(recs)
    lens = sorted((rid, str(n)) for rid, n in recs)
    if outfile:
        with open(outfile, 'w') as f:
            for name, n in lens:
                f.write(name + '\t' + n + '\n')
    return lens


def missing_tool(tool_name: str) -> List[str]:
    """[] if the executable is on PATH, else [tool_name] (utils.py:560-578). --lzpath/--bedtools are accepted for
    CLI compatibility but unused: the GPU library replaces both programs."""
    return [] if shutil.which(tool_name) else [tool_name]
