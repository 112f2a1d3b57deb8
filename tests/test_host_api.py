"""CPU tests of the host mirror of the reference's Python interface (no GPU work is triggered here)."""
import json
import os
import sys

import numpy as np
import pandas as pd
import pytest

from mimeo_b200 import fasta, utils, wrappers
from tests.helpers import GOLDEN, read_golden

MAN = json.loads(read_golden('manifest.json'))


def test_import_align_and_gff_match_reference_pandas_output():
    m = MAN['map']
    df = wrappers.import_Align(infile=os.path.join(GOLDEN, 'map.tab'), prefix=m['prefix'], minLen=m['minLen'], minIdt=m['minIdt'])
    assert list(df.columns) == ['tName', 'tStrand', 'tStart', 'tEnd', 'qName', 'qStrand', 'qStart', 'qEnd', 'score', 'pID', 'UID']
    assert df.index[0] == 1
    got = ''.join(wrappers.writeGFFlines(alnDF=df, chrlens=[tuple(x) for x in m['chrlens']], ftype=m['ftype']))
    assert got == read_golden('map.gff3')


def test_import_align_string_sort_and_default_prefix():
    df = wrappers.import_Align(infile=os.path.join(GOLDEN, 'map_kat7.tab'), prefix=None, minLen=100, minIdt=95)
    assert ''.join(wrappers.writeGFFlines(alnDF=df, chrlens=None, ftype='HGT')) == read_golden('map_kat7.gff3')
    assert df['tStart'].tolist() == ['1000', '200', '200', '99']          # strings, lexicographic


def test_import_align_exits_when_empty(tmp_path):
    p = tmp_path / 'e.tab'
    p.write_text('#h\nc\t+\t5\t104\tq\t+\t1\t100\t9000\t96.0\n')
    with pytest.raises(SystemExit) as e:
        wrappers.import_Align(infile=str(p), prefix='x', minLen=100, minIdt=95)     # 104-5 = 99
    assert e.value.code == 1


def test_split_fasta_chromlens_pairs(tmp_path):
    fa = tmp_path / 'g.fa'
    fa.write_text('>s2 desc here\nACGTACGTAC\nGT\n>S3\n' + 'A' * 130 + '\n>s10\nNNNN\n')
    d = tmp_path / 'split'
    d.mkdir()
    utils.splitFasta(str(fa), str(d))
    assert sorted(os.listdir(d)) == ['S3.fa', 's10.fa', 's2.fa']
    assert (d / 's2.fa').read_text() == '>s2 desc here\nACGTACGTACGT\n'
    assert (d / 'S3.fa').read_text() == '>S3\n' + 'A' * 60 + '\n' + 'A' * 60 + '\n' + 'A' * 10 + '\n'
    lens = utils.chromlens(str(d), str(tmp_path / 'lens.txt'))
    assert lens == [('S3', '130'), ('s10', '4'), ('s2', '12')]
    assert (tmp_path / 'lens.txt').read_text() == 'S3\t130\ns10\t4\ns2\t12\n'
    pairs = utils.get_all_pairs(str(d))
    assert len(pairs) == 9 and pairs[0][0] == pairs[0][1]
    pairs = utils.get_all_pairs(str(d), str(d))
    assert len(pairs) == 9
    with pytest.raises(SystemExit):
        utils.get_all_pairs(None, None)


def test_split_fasta_rejects_duplicate_ids(tmp_path):
    fa = tmp_path / 'g.fa'
    fa.write_text('>a\nAC\n>a\nGT\n')
    (tmp_path / 'o').mkdir()
    with pytest.raises(SystemExit):
        utils.splitFasta(str(fa), str(tmp_path / 'o'))


def test_set_paths_layout(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    fa = tmp_path / 'g.fa'
    fa.write_text('>a\nACGT\n')
    adir, bdir, outdir, outtab, gffout, tempdir = utils.set_paths(afasta=str(fa), outtab='o.tab', gffout='o.gff3', suppresBdir=True)
    assert bdir is None and tempdir and adir == os.path.join(tempdir, 'A_genome_split') and os.path.isfile(os.path.join(adir, 'a.fa'))
    assert outdir == str(tmp_path) and outtab == str(tmp_path / 'o.tab') and gffout == str(tmp_path / 'o.gff3')
    adir2, bdir2, *_ = utils.set_paths(adir=str(tmp_path / 'A'), afasta=str(fa), bdir=str(tmp_path / 'B'), bfasta=str(fa))
    assert os.path.isfile(os.path.join(adir2, 'a.fa')) and os.path.isfile(os.path.join(bdir2, 'a.fa'))


def ops(cmds):
    return [json.loads(c[len(utils.OP_PREFIX):]) for c in cmds]


def test_self_cmds_mirror_reference_stage_structure(tmp_path):
    """Same stage sequence as the reference's command list (golden: tests/golden/self_cmds.json)."""
    ref = json.loads(read_golden('self_cmds.json'))
    n_lastz = sum(1 for c in ref if c.startswith('lastz '))
    n_cov = sum(1 for c in ref if 'genomecov' in c)
    pairs = [('A/x.fa', 'A/x.fa'), ('A/x.fa', 'A/y.fa')]
    o = ops(wrappers.self_LZ_cmds(splitSelf=True, pairs=pairs, outtab=str(tmp_path / 'o.tab'), outgff='o.gff3', AchrmLens='lens.txt', prefix='P'))
    assert [x['op'] for x in o] == ['write', 'write', 'align', 'echo', 'coverage', 'echo', 'coverage']
    assert len(o[2]['pairs']) == n_lastz and sum(1 for x in o if x['op'] == 'coverage') == n_cov
    assert o[2]['outtab_intra'] == str(tmp_path / 'o.tab') + '_intra.tab'
    assert (o[4]['cov'], o[4]['label'], o[4]['write_header']) == (3, 'Self_repeats', True)
    assert (o[6]['cov'], o[6]['label'], o[6]['write_header'], o[6]['tab']) == (5, 'Self_repeats_intra', False, o[2]['outtab_intra'])
    assert o[0]['text'] == '#name1\tstrand1\tstart1\tend1\tname2\tstrand2\tstart2+\tend2+\tscore\tidentity\n'


def test_recycle_skips_alignment_and_warns_without_intra(tmp_path, caplog):
    tab = tmp_path / 'o.tab'
    tab.write_text('#h\n')
    o = ops(wrappers.self_LZ_cmds(splitSelf=True, pairs=[('a', 'a')], outtab=str(tab), outgff='g', AchrmLens='l', reuseTab=True, prefix='P'))
    assert [x['op'] for x in o] == ['echo', 'coverage']                   # intra file missing -> warning, no intra block
    assert 'Could not find intra-chrom results file' in caplog.text
    o = ops(wrappers.xspecies_LZ_cmds(pairs=[('a', 'b')], outtab=str(tab), outgff='g', AchrmLens='l', reuseTab=True, prefix='P'))
    assert [x['op'] for x in o] == ['echo', 'coverage'] and o[1]['source'] == 'mimeo' and o[1]['cov'] == 5
    o = ops(wrappers.xspecies_LZ_cmds(pairs=[('a', 'b')], outtab=str(tmp_path / 'new.tab'), outgff='g', AchrmLens='l', reuseTab=True, prefix='P'))
    assert [x['op'] for x in o] == ['write', 'align', 'echo', 'coverage'] and o[1]['hspthresh'] == 3000


def test_map_cmds_validation():
    with pytest.raises(ValueError):
        wrappers.map_LZ_cmds(pairs=[], outfile='x')
    with pytest.raises(ValueError):
        wrappers.map_LZ_cmds(pairs=[('a', 'b')], outfile=None)
    o = ops(wrappers.map_LZ_cmds(pairs=[('a', 'b')], outfile='x.tab', minIdt=90))
    assert [x['op'] for x in o] == ['write', 'align'] and o[1]['minIdt'] == 90


def test_run_cmd_never_shells_out(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    with pytest.raises(RuntimeError):
        utils.run_cmd(['echo pwned > x'])
    assert not (tmp_path / 'x').exists()
    assert [d for d in os.listdir(tmp_path) if d.startswith('tmp.')] == []      # temp dir removed
    utils.run_cmd([wrappers._op(op='write', path=str(tmp_path / 'h.tab'), text='hi\n')], keeptemp=True)
    assert (tmp_path / 'h.tab').read_text() == 'hi\n'
    assert len([d for d in os.listdir(tmp_path) if d.startswith('tmp.')]) == 1


def test_cli_surfaces_match_reference_defaults(monkeypatch):
    from mimeo_b200 import run_interspecies, run_map, run_self
    monkeypatch.setattr(sys, 'argv', ['mimeo-self'])
    a = run_self.mainArgs()
    assert (a.minIdt, a.minLen, a.minCov, a.hspthresh, a.intraCov, a.strictSelf, a.gffout, a.outfile, a.label, a.prefix, a.lzpath, a.bedtools) == \
        (60, 100, 3, 3000, 5, False, 'mimeo-self_repeats.gff3', 'mimeo_alignment.tab', 'Self_Repeat', 'Self_Repeat', 'lastz', 'bedtools')
    monkeypatch.setattr(sys, 'argv', ['mimeo-x', '--minIdt', '80'])
    a = run_interspecies.mainArgs()
    assert (a.minIdt, a.minCov, a.gffout, a.label, a.prefix) == (80, 5, 'mimeo_B_in_A.gff3', 'B_Repeat', 'B_Repeat') and isinstance(a.minIdt, int)
    monkeypatch.setattr(sys, 'argv', ['mimeo-map'])
    a = run_map.mainArgs()
    assert (a.gffout, a.label, a.prefix, a.maxtandem, a.tmaxperiod) == (None, 'BHit', 'BHit', None, 50)


@pytest.mark.parametrize('case', ['cov_order', 'cov_dense', 'cov_nointra'])
def test_recycle_cli_host_path_against_reference_goldens_with_the_device_stage_stubbed(tmp_path, monkeypatch, case):
    """Everything AROUND the coverage kernel on the `mimeo self -r` / `mimeo x -r` path -- CLI, chromlens, native .tab
    projection, scaffold ordering, native GFF3 rows, header and block order -- against the files the reference's own script
    wrote (tests/golden), on CPU: the device stage alone is replaced by the oracle's array statement (the GPU test
    test_gpu_cli.py::test_self_recycle_is_byte_identical_to_reference_script runs the same flow through the kernels)."""
    import json
    import shutil
    import sys
    from oracle import annot_oracle as ao
    from tests.helpers import read_golden
    from mimeo_b200 import app, coverage
    golden = os.path.join(os.path.dirname(__file__), 'golden')
    m = json.loads(read_golden('manifest.json'))[case]

    def stub(chrom, start, end, sizes, cov, min_len):
        out = ao.coverage_segments_arrays(np.asarray(chrom), np.asarray(start), np.asarray(end), sizes, max(int(cov), 1), int(min_len))
        return tuple(np.asarray(a, dtype=np.int32) for a in out)
    monkeypatch.setattr(coverage, 'coverage_segments', stub)
    monkeypatch.chdir(tmp_path)
    shutil.copy(os.path.join(golden, case + '.tab'), tmp_path / 'hits.tab')
    if m['has_intra']:
        shutil.copy(os.path.join(golden, case + '.tab_intra.tab'), tmp_path / 'hits.tab_intra.tab')
    adir = tmp_path / 'A'
    adir.mkdir()
    for ln in read_golden(case + '.lens').splitlines():
        n, size = ln.split('\t')
        (adir / (n + '.fa')).write_text(f'>{n}\n' + 'A' * int(size) + '\n')

    def run(argv):
        monkeypatch.setattr(sys, 'argv', ['mimeo'] + argv)
        try:
            app.main()
        except SystemExit as e:
            assert e.code in (0, None)
    argv = ['self', '--adir', str(adir), '-r', '--outfile', 'hits.tab', '--gffout', 'out.gff3', '--minCov', str(m['minCov']),
            '--intraCov', str(m['intraCov']), '--minLen', str(m['minLen']), '--label', m['label'], '--prefix', m['prefix']]
    if m['has_intra']:
        argv.append('--strictSelf')
    run(argv)
    assert (tmp_path / 'out.gff3').read_text() == read_golden(case + '.gff3')
    assert (tmp_path / 'A_gen_lens.txt').read_text() == read_golden(case + '.lens')
    run(['x', '--adir', str(adir), '--bdir', str(adir), '-r', '--outfile', 'hits.tab', '--gffout', 'x.gff3',
         '--minCov', str(m['minCov']), '--minLen', str(m['minLen'])])
    assert (tmp_path / 'x.gff3').read_text() == read_golden(case + '.x.gff3')


def _oracle_device_stubs(monkeypatch):
    """Replace the two device stages by their oracle statements (tests only): `engine.align_genomes` by the C
    LASTZ-restatement, `coverage.coverage_segments` by the numpy sweep. Everything else on the path stays the product's."""
    from oracle import annot_oracle as ao
    from oracle import lastz_oracle as lo
    from mimeo_b200 import align, coverage, engine

    def align_stub(tnames, tseqs, qnames, qseqs, hspthresh=3000, same=False, minLen=None, minIdt=None):   # unfiltered: the formatter filters too
        p = lo.default_params(hspthresh)
        cols = {f: [] for f in align.HIT_FIELDS}
        qenc = [lo.encode(np.asarray(s)) for s in qseqs]
        for ti, ts in enumerate(tseqs):
            t = lo.TargetIndex(lo.encode(np.asarray(ts)))
            for qi, q in enumerate(qenc):
                m = len(q)
                for strand, qq in ((0, q), (1, lo.revcomp_codes(q))):
                    for (s1, e1, s2, e2, score, nm, nc, _a1, _a2) in lo.align_tile(t, qq, p).tolist():
                        qs, qe = (s2 + 1, e2) if strand == 0 else (m - e2 + 1, m - s2)
                        for f, v in zip(align.HIT_FIELDS, (ti, qi, strand, s1 + 1, e1, qs, qe, score, nm, nc)):
                            cols[f].append(v)
        return {f: np.asarray(v, dtype=np.int32) for f, v in cols.items()}, {n: 0 for n in align.STAT_NAMES}

    def cov_stub(chrom, start, end, sizes, cov, min_len):
        out = ao.coverage_segments_arrays(np.asarray(chrom), np.asarray(start), np.asarray(end), sizes, max(int(cov), 1), int(min_len))
        return tuple(np.asarray(a, dtype=np.int32) for a in out)
    monkeypatch.setattr(engine, 'align_genomes', align_stub)
    monkeypatch.setattr(coverage, 'coverage_segments', cov_stub)


def _write_genome(path, g):
    with open(path, 'w') as f:
        for n, s in g.items():
            f.write(f'>{n}\n')
            t = s.tobytes().decode()
            for k in range(0, len(t), 70):
                f.write(t[k:k + 70] + '\n')


def _cli(monkeypatch, argv):
    from mimeo_b200 import app
    monkeypatch.setattr(sys, 'argv', ['mimeo'] + argv)
    try:
        app.main()
    except SystemExit as e:
        assert e.code in (0, None)


def test_self_cli_host_path_against_the_oracle_pipeline_with_device_stages_stubbed(tmp_path, monkeypatch):
    """`mimeo self` from a FASTA file on CPU: split directory, per-pair .tab blocks, _intra.tab, GFF3 -- the product's host
    code (native splitter, .tab and GFF formatters, block order) around stubbed device stages must reproduce the oracle's
    statement of the reference script byte for byte (GPU twin: test_gpu_cli.py::test_self_from_fasta_matches_oracle_pipeline)."""
    from oracle import lastz_oracle as lo
    from tests.helpers import synth_genome
    _oracle_device_stubs(monkeypatch)
    g = synth_genome(51, 3, 12_000, 2, copies=(5, 7), fam_len=(300, 900), sub=0.08, indel=0.004)
    monkeypatch.chdir(tmp_path)
    _write_genome(tmp_path / 'g.fa', g)
    _cli(monkeypatch, ['self', '--afasta', 'g.fa', '--adir', 'split', '--strictSelf', '--minIdt', '80', '--minCov', '2',
                       '--intraCov', '2', '--outfile', 'o.tab', '--gffout', 'o.gff3'])
    enc = {k: lo.encode(v) for k, v in g.items()}
    tab, intra, gff = lo.mimeo_self(enc, minIdt=80, minLen=100, minCov=2, intraCov=2, strictSelf=True)
    assert (tmp_path / 'o.tab').read_text() == tab and tab.count('\n') > 5
    assert (tmp_path / 'o.tab_intra.tab').read_text() == intra
    assert (tmp_path / 'o.gff3').read_text() == gff and gff.count('\n') > 3
    assert sorted(os.listdir(tmp_path / 'split')) == [n + '.fa' for n in sorted(g)]


def test_x_and_map_cli_host_path_against_the_oracle_pipeline_with_device_stages_stubbed(tmp_path, monkeypatch):
    from oracle import lastz_oracle as lo
    from tests.helpers import mutate, revcomp_ascii, synth_genome
    _oracle_device_stubs(monkeypatch)
    a = synth_genome(52, 2, 12_000, 0)
    b = synth_genome(53, 3, 8_000, 0)
    rng = np.random.default_rng(1)
    fam = a['scaf000'][3000:3900].copy()
    for k, (s, p) in enumerate([('scaf000', 100), ('scaf001', 1000), ('scaf001', 4000), ('scaf002', 300), ('scaf002', 3000), ('scaf002', 6000)]):
        cp = mutate(rng, fam, 0.04, 0.003)
        if k % 2:
            cp = revcomp_ascii(cp)
        b[s][p:p + len(cp)] = cp
    monkeypatch.chdir(tmp_path)
    _write_genome(tmp_path / 'a.fa', a)
    _write_genome(tmp_path / 'b.fa', b)
    _cli(monkeypatch, ['x', '--afasta', 'a.fa', '--bfasta', 'b.fa', '--minIdt', '80', '--minCov', '5', '--outfile', 'x.tab', '--gffout', 'x.gff3'])
    ea = {k: lo.encode(v) for k, v in a.items()}
    eb = {k: lo.encode(v) for k, v in b.items()}
    tab, gff = lo.mimeo_x(ea, eb, minIdt=80, minLen=100, minCov=5)
    assert (tmp_path / 'x.tab').read_text() == tab and (tmp_path / 'x.gff3').read_text() == gff
    assert gff.count('B_Repeat_00001') == 1
    _cli(monkeypatch, ['map', '--afasta', 'a.fa', '--bfasta', 'b.fa', '--minIdt', '90', '--outfile', 'm.tab', '--gffout', 'm.gff3'])
    tabm, gffm = lo.mimeo_map(ea, eb, minIdt=90, minLen=100)
    assert (tmp_path / 'm.tab').read_text() == tabm and (tmp_path / 'm.gff3').read_text() == gffm
    assert gffm.count('mimeo-map') >= 3
