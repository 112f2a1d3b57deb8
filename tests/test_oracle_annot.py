"""CPU tests: the annotation-half oracle against SURVEY 9.3 known answers and the reference-generated goldens."""
import ctypes
import json
import os

import numpy as np
import pytest

from tests.helpers import GOLDEN, read_golden
from oracle import annot_oracle as ao

MAN = json.loads(read_golden('manifest.json'))


def sizes_of(name):
    return {l.split('\t')[0]: int(l.split('\t')[1]) for l in read_golden(name).splitlines()}


def tab(*hits, chrom='c1'):
    return [f'{chrom}\t+\t{s}\t{e}\tq\t+\t1\t2\t100\t90.0\n' for s, e in hits]


# ------------------------------------------------------------------ SURVEY 9.3 known answers
def test_kat1_bedgraph_and_run():
    bed = ao.sort_bed(ao.project_bed(tab((100, 300), (200, 400), (250, 500))))
    bg = ao.genomecov_bg(bed, {'c1': 1000})
    assert bg == ['c1\t100\t200\t1', 'c1\t200\t250\t2', 'c1\t250\t300\t3', 'c1\t300\t400\t2', 'c1\t400\t500\t1']
    rows = ao.annotate_block(tab((100, 300), (200, 400), (250, 500)), {'c1': 1000}, 2, 100, 'mimeo-self', 'L', 'P')
    assert rows == ['c1\tmimeo-self\tL\t200\t400\t.\t+\t.\tID=P_00001\n']


def test_kat2_bookended_join():
    rows = ao.annotate_block(tab((0, 150), (0, 150), (150, 300), (150, 300)), {'c1': 1000}, 2, 1, 's', 'L', 'P')
    assert [r.split('\t')[3:5] for r in rows] == [['0', '300']]


def test_kat3_clip_to_size():
    rows = ao.annotate_block(tab((200, 400), (200, 400), (200, 400)), {'c1': 250}, 3, 50, 's', 'L', 'P')
    assert [r.split('\t')[3:5] for r in rows] == [['200', '250']]


def test_kat4_minlen_is_inclusive():
    h = tab((10, 110), (10, 110))
    assert len(ao.annotate_block(h, {'c1': 500}, 2, 100, 's', 'L', 'P')) == 1
    assert len(ao.annotate_block(h, {'c1': 500}, 2, 101, 's', 'L', 'P')) == 0


def test_kat5_chrom_order_and_id_restart():
    lines = tab((0, 100), (0, 100), chrom='s2') + tab((0, 100), (0, 100), chrom='S3') + tab((0, 100), (0, 100), chrom='s10')
    txt = ao.self_gff3(lines, lines, {'s2': 500, 'S3': 500, 's10': 500}, 2, 2, 10, 'Lab', 'P')
    rows = [r.split('\t') for r in txt.splitlines()[2:]]
    assert [r[0] for r in rows] == ['S3', 's10', 's2'] * 2
    assert [r[8] for r in rows] == ['ID=P_00001', 'ID=P_00002', 'ID=P_00003'] * 2
    assert [r[2] for r in rows] == ['Lab'] * 3 + ['Lab_intra'] * 3


def test_kat6_filter_vs_import_align():
    lz = ['#hdr\n', 't\t+\t5\t104\t100\tq\t+\t1\t100\t100\t9000\t80/100\t80.0%\n', '# lastz end-of-file\n']
    rows = ao.filter_lastz_general(lz, 100, 80)
    assert rows == ['t\t+\t5\t104\tq\t+\t1\t100\t9000\t80.0\n']
    with pytest.raises(SystemExit):
        ao.import_align_rows(rows, 'P', 100, 80)       # 104-5 = 99 < 100


def test_kat7_string_ordering_matches_reference_pandas():
    hits = ao.import_align_rows(read_golden('map_kat7.tab').splitlines(True), None, 100, 95)
    assert ''.join(ao.write_gff_lines(hits, None, 'HGT')) == read_golden('map_kat7.gff3')


def test_edge_cases_of_genomecov_restatement():
    # zero-length interval is invisible; interval past the end is clipped; start beyond the end is invisible
    assert ao.genomecov_bg_one(np.array([10]), np.array([10]), 100) == []
    assert ao.genomecov_bg_one(np.array([90]), np.array([500]), 100) == [(90, 100, 1)]
    assert ao.genomecov_bg_one(np.array([100]), np.array([120]), 100) == []
    assert ao.genomecov_bg_one(np.array([], dtype=np.int64), np.array([], dtype=np.int64), 100) == []
    with pytest.raises(ValueError):
        ao.genomecov_bg(['c\t5\t3'], {'c': 10})
    with pytest.raises(ValueError):
        ao.genomecov_bg(['zz\t1\t3'], {'c': 10})


# ------------------------------------------------------------------ goldens produced by the reference's own script
@pytest.mark.parametrize('case', ['cov_order', 'cov_dense', 'cov_nointra'])
def test_recycle_goldens_self_and_x(case):
    m = MAN[case]
    sizes = sizes_of(case + '.lens')
    lines = read_golden(case + '.tab').splitlines(True)
    intra = read_golden(case + '.tab_intra.tab').splitlines(True) if m['has_intra'] else None
    got = ao.self_gff3(lines, intra, sizes, m['minCov'], m['intraCov'], m['minLen'], m['label'], m['prefix'])
    assert got == read_golden(case + '.gff3')
    gotx = ao.x_gff3(lines, sizes, m['minCov'], m['minLen'], 'B_Repeat', 'B_Repeat')
    assert gotx == read_golden(case + '.x.gff3')


@pytest.mark.parametrize('strict', [False, True])
def test_filter_goldens(strict):
    m = MAN['filter']
    src = json.loads(read_golden('filter_lastz_in.json'))
    tabtxt = ao.TAB_HEADER
    intratxt = ao.TAB_HEADER
    for t, q in src['pairs']:
        rows = ''.join(ao.filter_lastz_general(src['lastz'][f'{q}_onto_{t}'], m['minLen'], m['minIdt']))
        if strict and t == q:
            intratxt += rows
        else:
            tabtxt += rows
    tag = 'filter_strict' if strict else 'filter_plain'
    assert tabtxt == read_golden(tag + '.tab')
    if strict:
        assert intratxt == read_golden(tag + '.tab_intra.tab')
    sizes = sizes_of('filter.lens')
    got = ao.self_gff3(tabtxt.splitlines(True), intratxt.splitlines(True) if strict else None, sizes,
                       m['minCov'], m['intraCov'], m['minLen'], m['label'], m['prefix'])
    assert got == read_golden(tag + '.gff3')


def test_map_golden():
    m = MAN['map']
    hits = ao.import_align_rows(read_golden('map.tab').splitlines(True), m['prefix'], m['minLen'], m['minIdt'])
    got = ''.join(ao.write_gff_lines(hits, [tuple(x) for x in m['chrlens']], m['ftype']))
    assert got == read_golden('map.gff3')


# ------------------------------------------------------------------ C oracle == numpy oracle
def test_c_oracle_matches_numpy(oracle_build):
    lib = ctypes.CDLL(os.path.join(oracle_build, 'libannot_oracle.so'))
    lib.ora_coverage_segments.restype = ctypes.c_long
    rng = np.random.default_rng(7)
    for trial in range(20):
        nchrom = int(rng.integers(1, 6))
        sizes = rng.integers(1, 3000, nchrom).astype(np.int64)
        n = int(rng.integers(0, 400))
        chrom = rng.integers(0, nchrom, n).astype(np.int32)
        start = (rng.random(n) * (sizes[chrom] + 20)).astype(np.int32)
        end = (start + rng.integers(0, 300, n)).astype(np.int32)
        cov, minlen = int(rng.integers(0, 5)), int(rng.integers(0, 60))
        cap = 2 * n + 8
        oc, os_, oe = (np.zeros(cap, np.int32) for _ in range(3))
        P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        k = lib.ora_coverage_segments(P(chrom), P(start), P(end), ctypes.c_long(n), P(sizes), ctypes.c_int(nchrom),
                                      ctypes.c_int(cov), ctypes.c_int(minlen), P(oc), P(os_), P(oe), ctypes.c_long(cap))
        ec, es, ee = ao.coverage_segments_arrays(chrom, start, end, sizes, cov, minlen)
        assert k == len(ec)
        assert (oc[:k] == ec).all() and (os_[:k] == es).all() and (oe[:k] == ee).all()
