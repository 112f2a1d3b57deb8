"""GPU end-to-end: the mimeo CLI / Python API on files, against the reference-generated goldens (recycle mode) and
against the oracle pipeline (alignment modes)."""
import json
import os
import shutil
import sys

import numpy as np
import pytest

from oracle import lastz_oracle as lo
from tests.helpers import GOLDEN, read_golden, synth_genome

pytestmark = pytest.mark.gpu
MAN = json.loads(read_golden('manifest.json'))


def run_cli(monkeypatch, argv):
    from mimeo_b200 import app
    monkeypatch.setattr(sys, 'argv', ['mimeo'] + argv)
    app.main()


def write_genome(path, g):
    with open(path, 'w') as f:
        for n, s in g.items():
            f.write(f'>{n}\n')
            t = s.tobytes().decode()
            for k in range(0, len(t), 70):
                f.write(t[k:k + 70] + '\n')


@pytest.mark.parametrize('case', ['cov_order', 'cov_dense', 'cov_nointra'])
def test_self_recycle_is_byte_identical_to_reference_script(tmp_path, monkeypatch, case):
    """config-2 path: `mimeo self -r --outfile X.tab` == what the reference's own bash script wrote (golden)."""
    m = MAN[case]
    monkeypatch.chdir(tmp_path)
    shutil.copy(os.path.join(GOLDEN, case + '.tab'), tmp_path / 'hits.tab')
    if m['has_intra']:
        shutil.copy(os.path.join(GOLDEN, case + '.tab_intra.tab'), tmp_path / 'hits.tab_intra.tab')
    adir = tmp_path / 'A'
    adir.mkdir()
    for line in read_golden(case + '.lens').splitlines():
        n, ln = line.split('\t')
        (adir / (n + '.fa')).write_text(f'>{n}\n' + 'A' * int(ln) + '\n')
    argv = ['self', '--adir', str(adir), '-r', '--outfile', 'hits.tab', '--gffout', 'out.gff3', '--minCov', str(m['minCov']),
            '--intraCov', str(m['intraCov']), '--minLen', str(m['minLen']), '--label', m['label'], '--prefix', m['prefix']]
    if m['has_intra']:
        argv.append('--strictSelf')
    run_cli(monkeypatch, argv)
    assert (tmp_path / 'out.gff3').read_text() == read_golden(case + '.gff3')
    assert (tmp_path / 'A_gen_lens.txt').read_text() == read_golden(case + '.lens')
    # and the same table through `mimeo x -r`
    run_cli(monkeypatch, ['x', '--adir', str(adir), '--bdir', str(adir), '-r', '--outfile', 'hits.tab', '--gffout', 'x.gff3',
                          '--minCov', str(m['minCov']), '--minLen', str(m['minLen'])])
    assert (tmp_path / 'x.gff3').read_text() == read_golden(case + '.x.gff3')


def test_self_from_fasta_matches_oracle_pipeline(tmp_path, monkeypatch):
    g = synth_genome(51, 4, 25_000, 3, copies=(6, 9), fam_len=(300, 1200), sub=0.08, indel=0.004)
    monkeypatch.chdir(tmp_path)
    write_genome(tmp_path / 'g.fa', g)
    run_cli(monkeypatch, ['self', '--afasta', 'g.fa', '--adir', 'split', '--strictSelf', '--minIdt', '80', '--minCov', '2',
                          '--intraCov', '2', '--outfile', 'o.tab', '--gffout', 'o.gff3'])
    enc = {k: lo.encode(v) for k, v in g.items()}
    tab, intra, gff = lo.mimeo_self(enc, minIdt=80, minLen=100, minCov=2, intraCov=2, strictSelf=True)
    assert (tmp_path / 'o.tab').read_text() == tab
    assert (tmp_path / 'o.tab_intra.tab').read_text() == intra
    assert (tmp_path / 'o.gff3').read_text() == gff and gff.count('\n') > 3
    assert sorted(os.listdir(tmp_path / 'split')) == [n + '.fa' for n in sorted(g)]


def test_x_and_map_match_oracle_pipeline(tmp_path, monkeypatch):
    a = synth_genome(52, 2, 30_000, 0)
    b = synth_genome(53, 3, 20_000, 0)
    rng = np.random.default_rng(1)
    from tests.helpers import mutate, revcomp_ascii
    fam = a['scaf000'][5000:6200].copy()
    for k, (s, p) in enumerate([('scaf000', 100), ('scaf001', 3000), ('scaf001', 9000), ('scaf002', 500), ('scaf002', 7000), ('scaf002', 15000)]):
        cp = mutate(rng, fam, 0.04, 0.003)
        if k % 2:
            cp = revcomp_ascii(cp)
        b[s][p:p + len(cp)] = cp
    monkeypatch.chdir(tmp_path)
    write_genome(tmp_path / 'a.fa', a)
    write_genome(tmp_path / 'b.fa', b)
    run_cli(monkeypatch, ['x', '--afasta', 'a.fa', '--bfasta', 'b.fa', '--minIdt', '80', '--minCov', '5', '--outfile', 'x.tab', '--gffout', 'x.gff3'])
    ea = {k: lo.encode(v) for k, v in a.items()}
    eb = {k: lo.encode(v) for k, v in b.items()}
    tab, gff = lo.mimeo_x(ea, eb, minIdt=80, minLen=100, minCov=5)
    assert (tmp_path / 'x.tab').read_text() == tab and (tmp_path / 'x.gff3').read_text() == gff
    assert gff.count('B_Repeat_00001') == 1
    run_cli(monkeypatch, ['map', '--afasta', 'a.fa', '--bfasta', 'b.fa', '--minIdt', '90', '--outfile', 'm.tab', '--gffout', 'm.gff3'])
    tabm, gffm = lo.mimeo_map(ea, eb, minIdt=90, minLen=100)
    assert (tmp_path / 'm.tab').read_text() == tabm and (tmp_path / 'm.gff3').read_text() == gffm
    assert gffm.count('mimeo-map') >= 3


def test_engine_errors_are_loud(tmp_path, monkeypatch):
    from mimeo_b200 import engine
    tab = tmp_path / 't.tab'
    tab.write_text('#h\nzz\t+\t1\t50\tq\t+\t1\t50\t100\t90.0\n')
    lens = tmp_path / 'l.txt'
    lens.write_text('c\t100\n')
    with pytest.raises(RuntimeError):
        engine.coverage_to_gff(str(tab), str(lens), str(tmp_path / 'o.gff3'), 1, 1, 's', 'l', 'p')
    tab.write_text('#h\nc\t+\t60\t50\tq\t+\t1\t50\t100\t90.0\n')
    with pytest.raises(RuntimeError):
        engine.coverage_to_gff(str(tab), str(lens), str(tmp_path / 'o.gff3'), 1, 1, 's', 'l', 'p')
    tab.write_text('#only a header\n')
    assert engine.coverage_to_gff(str(tab), str(lens), str(tmp_path / 'o.gff3'), 1, 1, 's', 'l', 'p') == 0
    assert (tmp_path / 'o.gff3').read_text() == engine.GFF_HEADER
