"""GPU parity: the whole LASTZ stage (seeds -> HSPs -> chain -> gapped) and the self/x/map outputs, against the oracle."""
import numpy as np
import pytest

from oracle import annot_oracle as ao
from oracle import lastz_oracle as lo
from tests.helpers import synth_genome

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def M():
    import mimeo_b200.align as A
    import mimeo_b200.genome as G
    return A, G


def oracle_rows(tg, qg, p):
    """Every alignment of every tile and strand as tuples in LASTZ output coordinates."""
    rows = set()
    qn = list(qg)
    for ti, (tname, t) in enumerate(tg.items()):
        tix = lo.TargetIndex(lo.encode(t))
        for qi, qname in enumerate(qn):
            qc = lo.encode(qg[qname])
            m = len(qc)
            for st in (0, 1):
                q = qc if st == 0 else lo.revcomp_codes(qc)
                for (s1, e1, s2, e2, sc, nm, nc, _a, _b) in lo.align_tile(tix, q, p).tolist():
                    qs, qe = (s2 + 1, e2) if st == 0 else (m - e2 + 1, m - s2)
                    rows.add((ti, qi, st, s1 + 1, e1, qs, qe, sc, nm, nc))
    return rows


def gpu_rows(hits):
    from mimeo_b200.align import HIT_FIELDS
    return set(zip(*[hits[f].tolist() for f in HIT_FIELDS]))


@pytest.mark.parametrize('seed', [41, 42])
def test_alignments_bit_exact_vs_oracle(M, seed):
    A, G = M
    g = synth_genome(seed, 3, 25_000, 3, copies=(3, 6), fam_len=(300, 2000), sub=0.10, indel=0.006, n_runs=2)
    T = G.Genome.from_dict(g)
    hits, stats = A.align(T, T, G.align_params(3000))
    want = oracle_rows(g, g, lo.default_params(3000))
    assert len(want) > 10
    assert gpu_rows(hits) == want
    assert stats['alignments'] == len(want) and stats['gapped_cells'] > 0


def test_flags_chain_and_gapped_off(M):
    A, G = M
    g = synth_genome(43, 2, 20_000, 2, copies=(4, 5), fam_len=(500, 900), sub=0.08, indel=0.004)
    T = G.Genome.from_dict(g)
    for kw in (dict(chain=0), dict(gapped=0), dict(chain=0, gapped=0)):
        hits, _ = A.align(T, T, G.align_params(3000, **kw))
        assert gpu_rows(hits) == oracle_rows(g, g, lo.default_params(3000, **kw)), kw


def test_self_pipeline_text_identical(M):
    """`mimeo self --strictSelf` end to end: .tab, _intra.tab and GFF3 byte-identical to the oracle pipeline."""
    A, G = M
    from mimeo_b200 import coverage
    g = synth_genome(44, 4, 30_000, 3, copies=(6, 10), fam_len=(300, 1500), sub=0.09, indel=0.005)
    enc = {k: lo.encode(v) for k, v in g.items()}
    tab_o, intra_o, gff_o = lo.mimeo_self(enc, minIdt=80, minLen=100, minCov=2, intraCov=2, strictSelf=True)
    names = sorted(g)
    T = G.Genome(names, [g[n] for n in names])
    hits, _ = A.align(T, T, G.align_params(3000))
    blocks = A.tab_blocks(hits, names, names, 100, 80)
    tab, intra = ao.TAB_HEADER, ao.TAB_HEADER
    for a in range(len(names)):
        for b in range(len(names)):
            rows = ''.join(blocks.get((a, b), []))
            if a == b:
                intra += rows
            else:
                tab += rows
    assert tab == tab_o and intra == intra_o
    assert len(tab.splitlines()) > 20
    sizes = [len(g[n]) for n in names]
    from tests.helpers import tab_to_arrays, segments_to_gff_rows
    text = ao.GFF_HEADER_SELF
    for txt, cov, label in ((tab, 2, 'Self_Repeat'), (intra, 2, 'Self_Repeat_intra')):
        c, s, e = tab_to_arrays(txt.splitlines(), names)
        text += ''.join(segments_to_gff_rows(coverage.coverage_segments(c, s, e, sizes, cov, 100), names, 'mimeo-self', label, 'Self_Repeat'))
    assert text == gff_o
    assert len(text.splitlines()) > 4


def test_sharded_driver_world1_equals_engine(M):
    """parallel.self_sharded on one rank (no process group) == engine.self_segments."""
    A, G = M
    from mimeo_b200 import engine, parallel
    g = synth_genome(45, 3, 25_000, 2, copies=(5, 7), fam_len=(400, 1200), sub=0.07, indel=0.004)
    names = sorted(g)
    seqs = [g[n] for n in names]
    T = G.Genome(names, seqs)
    inter, intra, hits, _ = engine.self_segments(T, None, [len(s) for s in seqs], 80, 100, 2, 2, 3000, True)
    h2, i2, j2 = parallel.self_sharded(names, seqs, 80, 100, 2, 2)
    keep = engine.filter_hits(h2, 100, 80)                  # self_segments returns the rows the device filter kept
    assert gpu_rows({f: v[keep] for f, v in h2.items()}) == gpu_rows(hits) and len(hits['t_id']) > 5
    assert i2.tolist() == np.stack(inter, axis=1).tolist() and j2.tolist() == np.stack(intra, axis=1).tolist()


def test_target_subset_with_identity_map_matches_full_run(M):
    """A rank holding a subset of the genome as T (row-block sharding) must produce exactly its rows of the full run,
    with or without the t_same_q hint (the hint only enables the closed-form trivial self-alignment)."""
    A, G = M
    g = synth_genome(46, 4, 20_000, 2, copies=(5, 7), fam_len=(400, 1200), sub=0.07, indel=0.004, n_runs=1)
    names = sorted(g)
    seqs = [g[n] for n in names]
    Q = G.Genome(names, seqs)
    full, _ = A.align(Q, Q, G.align_params(3000))
    full_rows = gpu_rows(full)
    sub = [1, 3]
    T = G.Genome([names[i] for i in sub], [seqs[i] for i in sub])
    for hint in (sub, None):
        part, _ = A.align(T, Q, G.align_params(3000), Q_aux=Q.both_strands(), t_same_q=hint)
        part = dict(part)
        part['t_id'] = np.asarray(sub, dtype=np.int32)[part['t_id']]
        want = {r for r in full_rows if r[0] in sub}
        assert gpu_rows(part) == want and len(want) > 4


def test_query_chunking_is_invisible(M, monkeypatch):
    """Processing the query in chunks of whole scaffolds (bounded buffers for big genomes) must not change a single row."""
    A, G = M
    g = synth_genome(47, 5, 20_000, 3, copies=(5, 8), fam_len=(400, 1500), sub=0.07, indel=0.004, n_runs=1)
    names = sorted(g)
    T = G.Genome(names, [g[n] for n in names])
    monkeypatch.delenv('MB2_CHUNK_MBP', raising=False)
    one, s1 = A.align(T, T, G.align_params(3000))
    monkeypatch.setenv('MB2_CHUNK_MBP', '0.03')          # 30 kbp chunks: the 10 strand-scaffolds fall into ~7 chunks
    many, s2 = A.align(T, T, G.align_params(3000))
    assert gpu_rows(one) == gpu_rows(many) and len(one['t_id']) > 10
    for k in ('seed_hits', 'leaders', 'survivors', 'hsps', 'alignments'):
        assert s1[k] == s2[k], k


def test_very_long_alignment_many_frame_moves_and_trace_chunks(M):
    """One alignment of ~150 000 columns (a scaffold with a single N against itself: no closed form): 300 000 anti-diagonals,
    hundreds of frame moves of the 16-bit scores and ~2 000 trace chunks walked back, equal to the oracle bit for bit."""
    A, G = M
    rng = np.random.default_rng(48)
    n = 150_000
    seq = np.frombuffer(b'ACGT', dtype=np.uint8)[rng.integers(0, 4, n)].copy()
    seq[100_000] = ord('N')
    g = {'s': seq}
    T = G.Genome.from_dict(g)
    hits, stats = A.align(T, T, G.align_params(3000), strands=1)
    p = lo.default_params(3000)
    tix = lo.TargetIndex(lo.encode(seq))
    want = {(0, 0, 0, s1 + 1, e1, s2 + 1, e2, sc, nm, nc) for (s1, e1, s2, e2, sc, nm, nc, _a, _b) in lo.align_tile(tix, lo.encode(seq), p).tolist()}
    assert gpu_rows(hits) == want
    assert max(r[4] - r[3] for r in want) > 140_000


def test_chain_batched_and_sequential_walks_agree(M, monkeypatch):
    """The batched chain DP (dense tiles) and the one-HSP-at-a-time walk (sparse tiles) are the same dynamic programme:
    forcing either on every tile must give identical alignments, equal to the oracle."""
    A, G = M
    g = synth_genome(49, 2, 60_000, 4, copies=(10, 16), fam_len=(400, 2500), sub=0.06, indel=0.004)
    T = G.Genome.from_dict(g)
    rows = {}
    for mode in ('1', '2', '0'):
        monkeypatch.setenv('MB2_CHAIN_MODE', mode)
        hits, _ = A.align(T, T, G.align_params(3000))
        rows[mode] = gpu_rows(hits)
    assert rows['1'] == rows['2'] == rows['0'] and len(rows['0']) > 50
    assert rows['0'] == oracle_rows(g, g, lo.default_params(3000))


@pytest.mark.parametrize('kind', range(6))
def test_awkward_genomes_bit_exact_vs_oracle(M, kind):
    """Tiny scaffolds (shorter than a seed), tandem arrays and low complexity, N-rich sequence, one dense family (batched chain
    walk, long descriptor rings) and long near-identical scaffolds (alignments of more than 65536 columns)."""
    A, G = M
    from tests.helpers import odd_genome
    g = odd_genome(np.random.default_rng(100 + kind), kind)
    T = G.Genome.from_dict(g)
    hits, _ = A.align(T, T, G.align_params(3000))
    want = oracle_rows(g, g, lo.default_params(3000))
    assert gpu_rows(hits) == want and len(want) >= 2


def test_soft_masked_bases_are_not_seeded_but_extended_through(M):
    """Lower-case input (soft-masking, e.g. RepeatMasker -xsmall) is excluded from seeding on both sequences and still scored
    by its base in every extension, as LASTZ does without [unmask]; GPU rows == oracle rows, and masking changes the output."""
    A, G = M
    g = synth_genome(71, 3, 30_000, 3, copies=(4, 7), fam_len=(400, 1800), sub=0.08, indel=0.004)
    rng = np.random.default_rng(5)
    masked = {}
    for name, seq in g.items():
        s = seq.copy()
        for _ in range(12):                      # lower-case stretches of 50-1500 bases, some swallow whole repeat copies
            p0 = int(rng.integers(0, len(s) - 1600)); n = int(rng.integers(50, 1500))
            s[p0:p0 + n] |= 0x20
        masked[name] = s
    T = G.Genome.from_dict(masked)
    hits, stats = A.align(T, T, G.align_params(3000))
    want = oracle_rows(masked, masked, lo.default_params(3000))
    assert gpu_rows(hits) == want and len(want) > 10
    plain = oracle_rows(g, g, lo.default_params(3000))
    assert plain != want                          # the mask matters on this input
