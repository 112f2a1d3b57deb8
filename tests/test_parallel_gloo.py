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
