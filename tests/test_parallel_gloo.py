"""CPU tests of the multi-rank host logic with the gloo backend (world_size 2): target partition, rank-local
filter/coverage, gather to rank 0. The GPU stages are replaced by the CPU oracle as stand-ins (tests may do that);
the result must equal the single-rank pipeline."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from mimeo_b200 import parallel


def test_partition_is_balanced_and_complete():
    lens = [500, 100, 400, 300, 300, 50, 50]
    parts = parallel.partition_targets(lens, 3)
    assert sorted(i for p in parts for i in p) == list(range(7))
    loads = [sum(lens[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 100
    assert parallel.partition_targets(lens, 1) == [list(range(7))]
    assert parallel.partition_targets([10], 4) == [[0], [], [], []]


def _oracle_standins(names, seqs, hspthresh=3000):
    from oracle import annot_oracle as ao
    from oracle import lastz_oracle as lo
    enc = [lo.encode(s) for s in seqs]

    def align_fn(t_idx):
        cols = {f: [] for f in parallel.HIT_FIELDS}
        p = lo.default_params(hspthresh)
        for lt, ti in enumerate(t_idx):
            tix = lo.TargetIndex(enc[ti])
            for qi, q in enumerate(enc):
                m = len(q)
                for st in (0, 1):
                    qq = q if st == 0 else lo.revcomp_codes(q)
                    for (s1, e1, s2, e2, sc, nm, nc, _a, _b) in lo.align_tile(tix, qq, p).tolist():
                        qs, qe = (s2 + 1, e2) if st == 0 else (m - e2 + 1, m - s2)
                        for f, v in zip(parallel.HIT_FIELDS, (lt, qi, st, s1 + 1, e1, qs, qe, sc, nm, nc)):
                            cols[f].append(v)
        return {f: np.array(v, dtype=np.int32) for f, v in cols.items()}

    def coverage_fn(c, s, e, sizes, cov, minlen):
        return ao.coverage_segments_arrays(c, s, e, sizes, cov, minlen)

    def filter_fn(hits, minLen, minIdt):
        n = len(hits['t_id'])
        keep = np.zeros(n, dtype=bool)
        for k in range(n):
            pct = float('%.1f' % (100.0 * hits['nmatch'][k] / hits['ncols'][k]))
            keep[k] = (hits['end1'][k] - hits['start1'][k] + 1 >= minLen) and pct >= minIdt
        return keep
    return align_fn, coverage_fn, filter_fn


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tests.helpers import synth_genome
    g = synth_genome(61, 3, 20_000, 2, copies=(5, 6), fam_len=(400, 900), sub=0.06, indel=0.003)
    names = sorted(g)
    seqs = [g[n] for n in names]
    a, c, f = _oracle_standins(names, seqs)
    out = parallel.self_sharded(names, seqs, 80, 100, 2, 2, align_fn=a, coverage_fn=c, filter_fn=f)
    if rank == 0:
        hits, inter, intra = out
        q.put((sorted(zip(*[hits[k].tolist() for k in parallel.HIT_FIELDS])), inter.tolist(), intra.tolist()))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_equals_single_rank():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single rank, no process group
    from tests.helpers import synth_genome
    g = synth_genome(61, 3, 20_000, 2, copies=(5, 6), fam_len=(400, 900), sub=0.06, indel=0.003)
    names = sorted(g)
    seqs = [g[n] for n in names]
    a, c, f = _oracle_standins(names, seqs)
    hits, inter, intra = parallel.self_sharded(names, seqs, 80, 100, 2, 2, align_fn=a, coverage_fn=c, filter_fn=f)
    want = (sorted(zip(*[hits[k].tolist() for k in parallel.HIT_FIELDS])), inter.tolist(), intra.tolist())
    assert got[0] == want[0] and len(want[0]) > 5
    assert got[1] == want[1] and got[2] == want[2] and len(want[1]) + len(want[2]) > 0


# ------------------------------------------------------------------------------------------ 2-D block plan
def test_shard_plan_covers_the_pair_grid_once():
    tl, ql = [5, 9, 3, 7, 7, 2], [4, 4, 8, 1, 6]
    for world in (1, 2, 3, 4, 6):
        plan = parallel.ShardPlan(tl, ql, world)
        assert plan.gt * plan.gq == world
        seen = set()
        for r in range(world):
            t, q = plan.block(r)
            for a in t:
                for b in q:
                    assert (a, b) not in seen
                    seen.add((a, b))
        assert len(seen) == len(tl) * len(ql) and 0 < plan.balance <= 1.0
    assert parallel.ShardPlan([1] * 50, [1] * 100, 8).balance == pytest.approx(1.0)
    with pytest.raises(ValueError):
        parallel.ShardPlan([1], [1], 4)


def _block_hits(enc_t, enc_q, t_idx, q_idx, hspthresh=3000):
    """A rank's block through the oracle (stand-in for mb2_align): rows with LOCAL ids."""
    from oracle import lastz_oracle as lo
    cols = {f: [] for f in parallel.HIT_FIELDS}
    p = lo.default_params(hspthresh)
    for lt, ti in enumerate(t_idx):
        tix = lo.TargetIndex(enc_t[ti])
        for lq, qi in enumerate(q_idx):
            q = enc_q[qi]
            m = len(q)
            for st in (0, 1):
                qq = q if st == 0 else lo.revcomp_codes(q)
                for (s1, e1, s2, e2, sc, nm, nc, _a, _b) in lo.align_tile(tix, qq, p).tolist():
                    qs, qe = (s2 + 1, e2) if st == 0 else (m - e2 + 1, m - s2)
                    for f, v in zip(parallel.HIT_FIELDS, (lt, lq, st, s1 + 1, e1, qs, qe, sc, nm, nc)):
                        cols[f].append(v)
    return {f: np.array(v, dtype=np.int32) for f, v in cols.items()}


def _grid_case():
    from tests.helpers import synth_genome
    from oracle import lastz_oracle as lo
    g = synth_genome(62, 4, 15_000, 2, copies=(6, 8), fam_len=(400, 900), sub=0.06, indel=0.003)
    names = sorted(g)
    enc = [lo.encode(g[n]) for n in names]
    return names, enc


def _grid_run(world, rank):
    from mimeo_b200 import engine
    from oracle import annot_oracle as ao
    names, enc = _grid_case()
    sizes = [len(e) for e in enc]
    plan = parallel.ShardPlan(sizes, sizes, world)
    t_idx, q_idx = plan.block(rank)
    hits = _block_hits(enc, enc, t_idx, q_idx)
    return parallel.annotate_block(hits, t_idx, q_idx, plan, sizes, 80, 100, ao.coverage_segments_arrays, [('inter', 2), ('intra', 2)],
                                   engine.filter_hits, strict_self=True)


def _grid_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = _grid_run(world, rank)
    if rank == 0:
        table, segs = out
        q.put((sorted(map(tuple, table.tolist())), {k: v.tolist() for k, v in segs.items()}))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 4])
def test_block_grid_with_all_gather_equals_single_rank(world):
    """The north-star partition on CPU ranks (gloo): (target group x query group) blocks, all-gather of the filtered hit
    tables, owner-side coverage, segments gathered to rank 0 == the one-rank result."""
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_grid_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    table, segs = _grid_run(1, 0)
    want = (sorted(map(tuple, table.tolist())), {k: v.tolist() for k, v in segs.items()})
    assert got[0] == want[0] and len(want[0]) > 5
    assert got[1] == want[1] and sum(len(v) for v in want[1].values()) > 0
