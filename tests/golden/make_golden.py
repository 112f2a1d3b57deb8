#!/usr/bin/env python
"""
Generate the golden fixtures under tests/golden/ by running the REFERENCE itself in this
container (it cannot travel to the GPU box, the fixtures can).

What is executed from /root/reference (unmodified, imported with `Bio` stubbed because
Biopython is not installed and nothing on this path uses it):
  * mimeo.wrappers.self_LZ_cmds / xspecies_LZ_cmds / map_LZ_cmds  -> the literal bash commands
  * mimeo.utils.run_cmd                                           -> bash runs them with the real awk/sed/sort
  * mimeo.wrappers.import_Align / writeGFFlines                   -> map post-processing under pandas

What is NOT available and is substituted (documented in DESIGN.md, "parity unpinned" items):
  * `lastz`    -> a fake executable that copies a canned 13-column `--format=general` file to --output
                  (so the sed/awk/sort filter commands a-4/a-5/a-6 run for real on realistic text)
  * `bedtools` -> oracle/_build/bedtools, our C restatement of genomecov -bg / merge

Run:  python tests/golden/make_golden.py      (needs /root/reference; rewrites tests/golden/*)
"""
import json
import os
import shutil
import stat
import subprocess
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference/src'


def load_reference():
    bio = types.ModuleType('Bio')
    seqio = types.ModuleType('Bio.SeqIO')
    bio.SeqIO = seqio
    sys.modules['Bio'] = bio
    sys.modules['Bio.SeqIO'] = seqio
    sys.path.insert(0, REF)
    import mimeo.utils as rutils
    import mimeo.wrappers as rwrap
    return rutils, rwrap


def build_shim():
    subprocess.check_call(['make', '-C', os.path.join(ROOT, 'oracle'), '_build/bedtools'], stdout=subprocess.DEVNULL)
    return os.path.join(ROOT, 'oracle', '_build', 'bedtools')


def write(path, text):
    with open(path, 'w') as f:
        f.write(text)


def read(path):
    with open(path) as f:
        return f.read()


# ----------------------------------------------------------------------------- synthetic inputs
def random_tab(rng, chroms, sizes, nhits, hotspots=4):
    """10-column rows (no header), sorted per README (cols 1,3,4). Coordinates may overrun the chromosome end."""
    rows = []
    centres = [(rng.integers(0, len(chroms)), rng.integers(0, 1 << 30)) for _ in range(hotspots)]
    for _ in range(nhits):
        if rng.random() < 0.7:
            c, x = centres[rng.integers(0, hotspots)]
            size = sizes[c]
            s = int(x % size + rng.normal(0, 60))
        else:
            c = int(rng.integers(0, len(chroms)))
            size = sizes[c]
            s = int(rng.integers(0, size))
        s = min(max(s, 1), size)
        ln = int(20 + rng.exponential(120))
        e = s + ln
        if rng.random() < 0.9:
            e = min(e, size)
        idt = rng.integers(600, 1001) / 10.0
        q = chroms[int(rng.integers(0, len(chroms)))]
        qs = int(rng.integers(1, 5000))
        rows.append((chroms[c], '+', s, e, q, '+-'[int(rng.integers(0, 2))], qs, qs + ln, 60 * ln, f'{idt:.1f}'))
    rows.sort(key=lambda r: (r[0].encode(), r[2], r[3]))
    return ['\t'.join(str(x) for x in r) + '\n' for r in rows]


def lastz_general(rng, tname, qname, n, tsize):
    """Fake LASTZ --format=general:...,identity output (13 columns, '%' present, header + --markend trailer)."""
    out = ['#name1\tstrand1\tstart1\tend1\tlength1\tname2\tstrand2\tstart2+\tend2+\tlength2\tscore\tidentity\tidPct\n']
    for _ in range(n):
        s = int(rng.integers(1, tsize - 50))
        ln = int(rng.choice([99, 100, 101, 150, 400, 1200]))
        e = min(s + ln - 1, tsize)
        ln1 = e - s + 1
        qs = int(rng.integers(1, 4000))
        ln2 = ln1 + int(rng.integers(-3, 4))
        ncol = ln1 - int(rng.integers(0, 3))
        nm = int(ncol * rng.choice([0.6, 0.795, 0.7995, 0.8, 0.8005, 0.9, 0.95, 1.0]))
        pct = 100.0 * nm / ncol
        out.append('\t'.join(str(x) for x in (tname, '+', s, e, ln1, qname, '+-'[int(rng.integers(0, 2))], qs, qs + ln2 - 1,
                                                ln2, int(rng.integers(3000, 90000)), f'{nm}/{ncol}', f'{pct:.1f}%')) + '\n')
    out.append('# lastz end-of-file\n')
    return out


FAKE_LASTZ = r'''#!/bin/bash
# fake lastz for golden generation: copies $CANNED_DIR/<query>_onto_<target>.lz to --output=
t=$(basename "$1"); t="${t%.*}"; q=$(basename "$2"); q="${q%.*}"
for a in "$@"; do case "$a" in --output=*) out="${a#--output=}";; esac; done
cp "$CANNED_DIR/${q}_onto_${t}.lz" "$out"
'''


def main():
    rutils, rwrap = load_reference()
    shim = build_shim()
    rng = np.random.default_rng(20261018)
    manifest = {}
    work = tempfile.mkdtemp(prefix='golden.')
    cwd0 = os.getcwd()
    os.chdir(work)
    try:
        # ------------------------------------------------------------------ recycle-mode coverage cases (config-2 shape)
        cases = {
            # SURVEY 9.3 KAT 5: chromosome order S3 < s10 < s2 in byte collation; IDs restart in the intra block
            'cov_order': dict(chroms=['s2', 'S3', 's10'], sizes=[1000, 800, 1200], nhits=120, intra=60,
                              minCov=2, intraCov=3, minLen=20, label='Self_Repeat', prefix='Self_Repeat'),
            'cov_dense': dict(chroms=['chr1', 'chr2', 'chr10', 'scaffold_7'], sizes=[5000, 3000, 777, 12000], nhits=1500,
                              intra=400, minCov=3, intraCov=4, minLen=100, label='Rep', prefix='R'),
            'cov_nointra': dict(chroms=['a', 'b'], sizes=[4000, 4000], nhits=300, intra=0,
                                minCov=5, intraCov=5, minLen=50, label='Self_Repeat', prefix='P'),
        }
        for name, c in cases.items():
            tab = random_tab(rng, c['chroms'], c['sizes'], c['nhits'])
            intra = random_tab(rng, c['chroms'], c['sizes'], c['intra']) if c['intra'] else None
            outtab = os.path.join(work, name + '.tab')
            write(outtab, '#name1\tstrand1\tstart1\tend1\tname2\tstrand2\tstart2+\tend2+\tscore\tidentity\n' + ''.join(tab))
            if intra is not None:
                write(outtab + '_intra.tab', '#name1\tstrand1\tstart1\tend1\tname2\tstrand2\tstart2+\tend2+\tscore\tidentity\n' + ''.join(intra))
            lens = os.path.join(work, name + '.lens')
            write(lens, ''.join(f'{n}\t{s}\n' for n, s in sorted(zip(c['chroms'], c['sizes']))))
            gff = os.path.join(work, name + '.gff3')
            cmds = rwrap.self_LZ_cmds(lzpath='lastz', bdtlsPath=shim, splitSelf=intra is not None, Adir=None, Bdir=None,
                                      pairs=[], outtab=outtab, outgff=gff, minIdt=80, minLen=c['minLen'],
                                      hspthresh=3000, minCov=c['minCov'], intraCov=c['intraCov'], AchrmLens=lens,
                                      reuseTab=True, label=c['label'], prefix=c['prefix'])
            rutils.run_cmd(cmds)
            for ext in ('.tab', '.lens', '.gff3'):
                shutil.copy(os.path.join(work, name + ext), os.path.join(HERE, name + ext))
            if intra is not None:
                shutil.copy(outtab + '_intra.tab', os.path.join(HERE, name + '.tab_intra.tab'))
            # the same tab through `mimeo x` (source column "mimeo", single block)
            gffx = os.path.join(work, name + '.x.gff3')
            cmds = rwrap.xspecies_LZ_cmds(lzpath='lastz', bdtlsPath=shim, Adir=None, Bdir=None, pairs=[], outtab=outtab,
                                          outgff=gffx, minIdt=80, minLen=c['minLen'], minCov=c['minCov'], AchrmLens=lens,
                                          reuseTab=True, label='B_Repeat', prefix='B_Repeat')
            rutils.run_cmd(cmds)
            shutil.copy(gffx, os.path.join(HERE, name + '.x.gff3'))
            manifest[name] = {k: v for k, v in c.items() if k not in ('nhits', 'intra')}
            manifest[name]['has_intra'] = intra is not None

        # ------------------------------------------------------------------ filter stage a-4/a-5/a-6 through the real awk/sed/sort
        canned = os.path.join(work, 'canned')
        os.makedirs(canned)
        fake = os.path.join(work, 'lastz')
        write(fake, FAKE_LASTZ)
        os.chmod(fake, os.stat(fake).st_mode | stat.S_IEXEC)
        os.environ['CANNED_DIR'] = canned
        scafs = ['scafA', 'scafB', 'scaf_c']
        tsz = {'scafA': 6000, 'scafB': 9000, 'scaf_c': 3000}
        adir = os.path.join(work, 'adir')
        os.makedirs(adir)
        for s in scafs:
            write(os.path.join(adir, s + '.fa'), f'>{s}\nACGT\n')
        pairs = [(os.path.join(adir, a + '.fa'), os.path.join(adir, b + '.fa')) for a in scafs for b in scafs]
        lz_all = {}
        for a in scafs:
            for b in scafs:
                lz = lastz_general(rng, a, b, 25, tsz[a])
                write(os.path.join(canned, f'{b}_onto_{a}.lz'), ''.join(lz))
                lz_all[f'{b}_onto_{a}'] = lz
        write(os.path.join(HERE, 'filter_lastz_in.json'), json.dumps({'pairs': [[os.path.basename(a)[:-3], os.path.basename(b)[:-3]] for a, b in pairs],
                                                                      'lastz': lz_all}, indent=0))
        lens = os.path.join(work, 'filter.lens')
        write(lens, ''.join(f'{n}\t{tsz[n]}\n' for n in sorted(scafs)))
        shutil.copy(lens, os.path.join(HERE, 'filter.lens'))
        for strict in (False, True):
            tag = 'filter_strict' if strict else 'filter_plain'
            outtab = os.path.join(work, tag + '.tab')
            gff = os.path.join(work, tag + '.gff3')
            cmds = rwrap.self_LZ_cmds(lzpath=fake, bdtlsPath=shim, splitSelf=strict, Adir=adir, Bdir=None, pairs=pairs,
                                      outtab=outtab, outgff=gff, minIdt=80, minLen=100, hspthresh=3000, minCov=2,
                                      intraCov=2, AchrmLens=lens, reuseTab=False, label='Self_Repeat', prefix='Self_Repeat')
            rutils.run_cmd(cmds)
            shutil.copy(outtab, os.path.join(HERE, tag + '.tab'))
            shutil.copy(gff, os.path.join(HERE, tag + '.gff3'))
            if strict:
                shutil.copy(outtab + '_intra.tab', os.path.join(HERE, tag + '.tab_intra.tab'))
        manifest['filter'] = dict(minIdt=80, minLen=100, minCov=2, intraCov=2, label='Self_Repeat', prefix='Self_Repeat')

        # ------------------------------------------------------------------ map: map_LZ_cmds + import_Align + writeGFFlines
        outtab = os.path.join(work, 'map.tab')
        cmds = rwrap.map_LZ_cmds(lzpath=fake, pairs=pairs, minIdt=90, minLen=100, hspthresh=3000, outfile=outtab)
        rutils.run_cmd(cmds)
        shutil.copy(outtab, os.path.join(HERE, 'map.tab'))
        df = rwrap.import_Align(infile=outtab, prefix='BHit', minLen=100, minIdt=90)
        chrlens = [(n, str(tsz[n])) for n in sorted(scafs)]
        write(os.path.join(HERE, 'map.gff3'), ''.join(rwrap.writeGFFlines(alnDF=df, chrlens=chrlens, ftype='BHit')))
        # SURVEY 9.3 KAT 7: string ordering of coordinates
        kat7 = ('#h\n' + 'c\t+\t1000\t1200\tq\t+\t1\t201\t9000\t95.0\n' + 'c\t-\t200\t400\tq\t-\t1\t201\t9000\t95.0\n'
                + 'c\t+\t200\t400\tq\t+\t1\t201\t9000\t95.0\n' + 'c\t+\t99\t300\tq\t+\t1\t202\t9000\t99.5\n'
                + 'c\t+\t5\t104\tq\t+\t1\t100\t9000\t96.0\n')
        write(os.path.join(HERE, 'map_kat7.tab'), kat7)
        df = rwrap.import_Align(infile=os.path.join(HERE, 'map_kat7.tab'), prefix=None, minLen=100, minIdt=95)
        write(os.path.join(HERE, 'map_kat7.gff3'), ''.join(rwrap.writeGFFlines(alnDF=df, chrlens=None, ftype='HGT')))
        manifest['map'] = dict(minIdt=90, minLen=100, prefix='BHit', ftype='BHit', chrlens=chrlens)

        # ------------------------------------------------------------------ command-list shape (API parity of the generators)
        cmds = rwrap.self_LZ_cmds(lzpath='lastz', bdtlsPath='bedtools', splitSelf=True, Adir='A', Bdir=None,
                                  pairs=[('A/x.fa', 'A/x.fa'), ('A/x.fa', 'A/y.fa')], outtab='o.tab', outgff='o.gff3',
                                  AchrmLens='lens.txt', prefix='P')
        write(os.path.join(HERE, 'self_cmds.json'), json.dumps(cmds, indent=0))
    finally:
        os.chdir(cwd0)
        shutil.rmtree(work, ignore_errors=True)
    write(os.path.join(HERE, 'manifest.json'), json.dumps(manifest, indent=1, sort_keys=True))
    print('golden fixtures written to', HERE)


if __name__ == '__main__':
    main()
