"""GPU parity: device genome, seed table/scan and gap-free x-drop HSPs, bit-exact against the LASTZ-restatement oracle."""
import numpy as np
import pytest

from oracle import lastz_oracle as lo
from tests.helpers import synth_genome

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def G():
    from mimeo_b200 import genome
    return genome


def oracle_hsps(tg, qg, p, strand='+'):
    """All tiles through the oracle; rows (tile, s1, s2, len, score) in canonical order."""
    rows = []
    qnames = list(qg)
    for ti, (tn, t) in enumerate(tg.items()):
        tix = lo.TargetIndex(lo.encode(t))
        for qi, qn in enumerate(qnames):
            q = lo.encode(qg[qn])
            if strand == '-':
                q = lo.revcomp_codes(q)
            h = lo.hsps(tix, q, p)
            for r in sorted(map(tuple, h.tolist())):
                rows.append((ti * len(qnames) + qi,) + r)
    return np.array(rows, dtype=np.int64).reshape(-1, 5)


def test_pack_and_revcomp_roundtrip(G):
    g = synth_genome(3, 3, 1000, 0, n_runs=3)
    g['odd'] = np.frombuffer(b'ACGTNNacgtRYKMacgtn' * 7 + b'A', dtype=np.uint8)       # lower case, IUPAC, length not a multiple of 32
    dev = G.Genome.from_dict(g)
    rc = dev.revcomp()
    for i, (n, s) in enumerate(g.items()):
        want = lo.encode(s) & 7                  # decode reports bases; the soft-mask bit of the oracle's codes is not a base
        assert (dev.decode(i) == want).all()
        assert (rc.decode(i) == lo.revcomp_codes(want)).all()
    assert (rc.revcomp().decode(3) == (lo.encode(g['odd']) & 7)).all()


@pytest.mark.parametrize('seed,strand', [(11, '+'), (12, '-'), (13, '+')])
def test_hsps_bit_exact_vs_oracle(G, seed, strand):
    g = synth_genome(seed, 3, 30_000, 3, copies=(3, 5), fam_len=(400, 1500), sub=0.08, indel=0.004, n_runs=2)
    T = G.Genome.from_dict(g)
    Q = T.revcomp() if strand == '-' else T
    got, stats = G.test_hsps(T, Q, G.align_params(3000))
    want = oracle_hsps(g, g, lo.default_params(3000), strand)
    assert len(want) > 10
    assert got.shape == want.shape and (got == want).all()
    assert stats[1] >= stats[2] >= stats[0] >= stats[4] == len(want)


def test_hsps_entropy_off_and_low_threshold(G):
    g = synth_genome(21, 2, 20_000, 2, copies=(4, 4), fam_len=(300, 600), sub=0.12, indel=0.0)
    low = np.frombuffer(b'AC' * 400, dtype=np.uint8)                    # low-complexity tract twice: entropy must prune it
    g['scaf000'][1000:1800] = low; g['scaf001'][5000:5800] = low
    T = G.Genome.from_dict(g)
    for kw in (dict(entropy=0), dict(entropy=1), dict(transition=0)):
        got, _ = G.test_hsps(T, T, G.align_params(2200, **kw))
        want = oracle_hsps(g, g, lo.default_params(2200, **kw))
        assert got.shape == want.shape and (got == want).all(), kw


def test_two_genomes_x_mode_shapes(G):
    a = synth_genome(31, 2, 25_000, 0)
    b = synth_genome(32, 3, 15_000, 0)
    b['scaf001'][2000:4000] = a['scaf000'][7000:9000]                  # shared segment, exact copy
    b['scaf002'][100:900] = a['scaf001'][20000:20800]
    A, B = G.Genome.from_dict(a), G.Genome.from_dict(b)
    got, _ = G.test_hsps(A, B, G.align_params(3000))
    want = oracle_hsps(a, b, lo.default_params(3000))
    assert len(want) >= 2 and got.shape == want.shape and (got == want).all()
