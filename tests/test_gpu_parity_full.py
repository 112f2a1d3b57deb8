"""GPU parity AT THE BENCHMARKED SIZES (BASELINE configs 1 and 2): every row and every output byte of the workloads bench.py
times, against the CPU oracle run here on all host cores."""
import os

import numpy as np
import pytest

from oracle import lastz_oracle as lo

pytestmark = pytest.mark.gpu


def test_c1_full_every_row_and_output_byte(tmp_path, monkeypatch):
    """BASELINE config 1 (`mimeo self`, 5 Mbp, 10 x 500 kbp, seed 1001, minIdt 80 minLen 100 minCov 3 intraCov 4 strictSelf):
    all 100 ordered scaffold pairs through the oracle (one process per pair, like the reference's script), then
    (1) every alignment row of mb2_align equal, (2) .tab, _intra.tab and .gff3 written by the CLI byte-identical."""
    import sys
    from bench import SelfWorkload
    from mimeo_b200 import align as A, app, genome as G
    wl = SelfWorkload(0)
    enc = {n: lo.encode(s) for n, s in zip(wl.names, wl.seqs)}
    st = lo.Stats()
    general = lo.general_all_pairs(enc, None, wl.HSPTHRESH, workers=os.cpu_count() or 1, stats=st)
    want = lo.general_rows(general, wl.names, wl.names)
    T = G.Genome(wl.names, wl.seqs)
    hits, stats = A.align(T, T, G.align_params(wl.HSPTHRESH))
    T.close()
    got = set(zip(*[hits[f].tolist() for f in A.HIT_FIELDS]))
    assert len(want) > 1000
    assert got == want, f'{len(got - want)} rows only on the GPU, {len(want - got)} only in the oracle'
    assert stats['hsps'] == st.hsps_kept and stats['seed_hits'] == st.seed_hits
    # the files
    monkeypatch.chdir(tmp_path)
    with open('g.fa', 'w') as f:
        for n, s in zip(wl.names, wl.seqs):
            f.write(f'>{n}\n{s.tobytes().decode()}\n')
    monkeypatch.setattr(sys, 'argv', ['mimeo', 'self', '--afasta', 'g.fa', '--adir', 'split', '--strictSelf', '--minIdt', str(wl.MIN_IDT),
                                      '--minLen', str(wl.MIN_LEN), '--minCov', str(wl.MIN_COV), '--intraCov', str(wl.INTRA_COV),
                                      '--outfile', 'o.tab', '--gffout', 'o.gff3'])
    app.main()
    tab, intra, gff = lo.ao.TAB_HEADER, lo.ao.TAB_HEADER, None
    for a in sorted(enc):
        for b in sorted(enc):
            rows = ''.join(lo.ao.filter_lastz_general(general[(a, b)], wl.MIN_LEN, wl.MIN_IDT))
            if a == b:
                intra += rows
            else:
                tab += rows
    gff = lo.ao.self_gff3(tab.splitlines(True), intra.splitlines(True), {n: len(c) for n, c in enc.items()}, wl.MIN_COV, wl.INTRA_COV,
                          wl.MIN_LEN, 'Self_Repeat', 'Self_Repeat')
    assert (tmp_path / 'o.tab').read_text() == tab
    assert (tmp_path / 'o.tab_intra.tab').read_text() == intra
    assert (tmp_path / 'o.gff3').read_text() == gff and gff.count('\n') > 20


def test_c2_full_segments_equal_oracle():
    """BASELINE config 2 (10 M hits over 100 Mbp): the segments of the coverage stage equal the C oracle's, element for element."""
    from bench import CoverageWorkload
    from mimeo_b200 import coverage
    from oracle import annot_oracle as ao
    wl = CoverageWorkload(0)
    got = coverage.coverage_segments(wl.chrom, wl.start, wl.end, wl.sizes, wl.MIN_COV, wl.MIN_LEN)
    want = ao.coverage_segments_c(wl.chrom, wl.start, wl.end, wl.sizes, wl.MIN_COV, wl.MIN_LEN)
    assert len(want[0]) >= 50          # 70x mean depth: every scaffold is one segment, clipped at its ends
    for g, w in zip(got, want):
        assert np.array_equal(np.asarray(g), np.asarray(w))


def test_c5_shaped_genome_every_row():
    """BASELINE config 5's generator (plant-like: 40 % of the bases from repeat families, log-normal scaffold lengths) at a size
    the oracle finishes in seconds (1.5 Mbp, 6 scaffolds, 126 k HSPs, 1.1 k alignments): every row of every tile equal. The
    1 Gbp run itself (profiles/r2_c5_1Gbp_n8.json) is the same kernels on the same generator."""
    from mimeo_b200 import align as A, genome as G
    from tests.helpers import synth_c5
    g = synth_c5(1005, 1_500_000, 6, fam_len=(1000, 4000), copies=(10, 60))
    names = sorted(g, key=lambda s: s.encode())
    enc = {n: lo.encode(g[n]) for n in names}
    st = lo.Stats()
    general = lo.general_all_pairs(enc, None, 3000, workers=os.cpu_count() or 1, stats=st)
    want = lo.general_rows(general, names, names)
    T = G.Genome(names, [g[n] for n in names])
    hits, stats = A.align(T, T, G.align_params(3000))
    T.close()
    got = set(zip(*[hits[f].tolist() for f in A.HIT_FIELDS]))
    assert len(want) > 1000 and got == want, f'{len(got - want)} rows only on the GPU, {len(want - got)} only in the oracle'
    assert stats['hsps'] == st.hsps_kept and stats['seed_hits'] == st.seed_hits
