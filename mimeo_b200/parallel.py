"""
Multi-GPU execution of one `mimeo self / x / map` job: one process per GPU (torchrun). The reference's pair grid
(utils.get_all_pairs, utils.py:92-102: every target file x every query file) is cut into a gt x gq grid of
(target group x query group) blocks, one block per rank (`ShardPlan`): LASTZ's --chain and gapped scope is one
(target scaffold, query scaffold, strand) tile, so blocks are independent. A target scaffold's hits are then born on the gq
ranks of its row ("scaffolds spanning shards"): the filtered hit tables are all-gathered over NCCL / NVLink, every rank
thresholds the coverage of the target scaffolds it owns from the gathered table, and the segments are gathered to rank 0.
(Summing per-rank depth arrays instead would move 4 bytes per genome base; the hit table is 40 bytes per hit and far
smaller.) `self_sharded` is the older row-block variant (targets only) kept for `--workload c5s`.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

HIT_FIELDS = ('t_id', 'q_id', 'strand', 'start1', 'end1', 'start2', 'end2', 'score', 'nmatch', 'ncols')


def partition_targets(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time bin packing of target scaffolds onto ranks (cost ~ scaffold length, since every target
    meets the same query genome). Deterministic: ties by index. Within a rank, indices are returned in ascending order."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world
    parts: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        parts[r].append(i)
        load[r] += int(lengths[i])
    return [sorted(p) for p in parts]


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def gather_rows(table: np.ndarray, dst: Optional[int] = 0, device=None) -> Optional[np.ndarray]:
    """Gather variable-length int32 row tables [n_r, k] from every rank to `dst` (rank order; None = to every rank). Returns None elsewhere.
    Sizes first, then one padded all_gather -- the all-gather-v of hit tables SURVEY 8(e) names."""
    import torch
    dist = _dist()
    table = np.ascontiguousarray(table, dtype=np.int32)
    if dist is None or dist.get_world_size() == 1:
        return table
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device if device is not None else (torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else torch.device('cpu'))
    k = table.shape[1]
    n = torch.tensor([table.shape[0]], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    nmax = max(max(sizes), 1)
    buf = torch.zeros((nmax, k), dtype=torch.int32, device=dev)
    if table.shape[0]:
        buf[:table.shape[0]] = torch.from_numpy(table).to(dev)
    out = [torch.empty((nmax, k), dtype=torch.int32, device=dev) for _ in range(world)]
    dist.all_gather(out, buf)
    if dev.type == 'cuda':
        torch.cuda.synchronize(dev)      # the collective is complete before any rank reuses or frees its buffers
    if dst is not None and rank != dst:
        return None
    return np.concatenate([o[:s].cpu().numpy() for o, s in zip(out, sizes)], axis=0)


def self_sharded(names: Sequence[str], seqs: Sequence[np.ndarray], minIdt, minLen, minCov, intraCov, hspthresh=3000, strictSelf=True,
                 align_fn: Optional[Callable] = None, coverage_fn: Optional[Callable] = None, filter_fn: Optional[Callable] = None):
    """`mimeo self` across all ranks. names must already be in C-locale order (it defines scaffold indices and GFF order).
    Returns on rank 0: (hits dict with GLOBAL t_id, inter segments, intra segments or None); None on other ranks.
    align_fn / coverage_fn / filter_fn default to the GPU engine; CPU tests inject stand-ins."""
    dist = _dist()
    world = dist.get_world_size() if dist else 1
    rank = dist.get_rank() if dist else 0
    sizes = [len(s) for s in seqs]
    mine = partition_targets(sizes, world)[rank]
    if align_fn is None:
        from . import align as _align, coverage as _coverage, engine
        from .genome import Genome, align_params

        def align_fn(t_idx):
            Q = Genome(list(names), list(seqs))
            T = Q if len(t_idx) == len(names) else Genome([names[i] for i in t_idx], [seqs[i] for i in t_idx])
            try:
                hits, _ = _align.align(T, Q, align_params(hspthresh), t_same_q=None if T is Q else list(t_idx))
            finally:
                if T is not Q:
                    T.close()
                Q.close()
            return hits
        coverage_fn = _coverage.coverage_segments
        filter_fn = engine.filter_hits
    hits = align_fn(mine) if mine else {f: np.zeros(0, np.int32) for f in HIT_FIELDS}
    hits = dict(hits)
    hits['t_id'] = np.asarray(mine, dtype=np.int32)[hits['t_id']] if len(hits['t_id']) else hits['t_id']   # local -> global target index
    keep = filter_fn(hits, minLen, minIdt)
    intra_mask = (hits['t_id'] == hits['q_id']) & keep if strictSelf else np.zeros(len(keep), dtype=bool)
    inter_mask = keep & ~intra_mask

    def seg(mask, cov):
        if not mask.any():
            return np.zeros((0, 3), np.int32)
        c, s, e = coverage_fn(hits['t_id'][mask], hits['start1'][mask], hits['end1'][mask], sizes, cov, minLen)
        return np.stack([c, s, e], axis=1).astype(np.int32)
    inter = seg(inter_mask, minCov)
    intra = seg(intra_mask, intraCov) if strictSelf else np.zeros((0, 3), np.int32)
    table = np.stack([hits[f] for f in HIT_FIELDS], axis=1).astype(np.int32) if len(hits['t_id']) else np.zeros((0, 10), np.int32)
    g_hits, g_inter, g_intra = gather_rows(table), gather_rows(inter), gather_rows(intra)
    if rank != 0:
        return None

    def by_chrom(t):
        return t[np.lexsort((t[:, 1], t[:, 0]))] if len(t) else t    # scaffold index, then start: the single-GPU order
    hits_all = {f: g_hits[:, k] for k, f in enumerate(HIT_FIELDS)}
    return hits_all, by_chrom(g_inter), (by_chrom(g_intra) if strictSelf else None)


# ------------------------------------------------------------------------------------------ 2-D block partition
def _lpt_groups(lengths: Sequence[int], k: int) -> List[List[int]]:
    return partition_targets(lengths, k)


class ShardPlan:
    """gt x gq grid over (target scaffolds x query scaffolds), one block per rank, chosen among the factorisations of
    `world` for the smallest critical path max_r(len(T_r) * len(Q_r)); ties prefer more target groups (smaller seed
    tables per rank, better L2 locality of the lookups)."""

    def __init__(self, tlens: Sequence[int], qlens: Sequence[int], world: int):
        best = None
        for gt in range(1, world + 1):
            if world % gt:
                continue
            gq = world // gt
            if gt > len(tlens) or gq > len(qlens):
                continue
            tg, qg = _lpt_groups(tlens, gt), _lpt_groups(qlens, gq)
            crit = max(sum(tlens[i] for i in a) for a in tg) * max(sum(qlens[i] for i in b) for b in qg)
            key = (crit, -gt)
            if best is None or key < best[0]:
                best = (key, gt, gq, tg, qg)
        if best is None:
            raise ValueError(f'cannot place {world} ranks on {len(tlens)} x {len(qlens)} scaffolds')
        _, self.gt, self.gq, self.tgroups, self.qgroups = best
        self.world = world
        total = float(sum(tlens)) * float(sum(qlens))
        self.balance = total / (world * float(best[0][0])) if best[0][0] else 1.0     # 1.0 = perfectly even blocks

    def block(self, rank: int) -> Tuple[List[int], List[int]]:
        """(global target scaffold indices, global query scaffold indices) of a rank."""
        return self.tgroups[rank // self.gq], self.qgroups[rank % self.gq]

    def owner(self, t: np.ndarray) -> np.ndarray:
        """rank that thresholds the coverage of target scaffold t (round robin)."""
        return np.asarray(t) % self.world


def annotate_block(hits: Dict[str, np.ndarray], t_idx: Sequence[int], q_idx: Sequence[int], plan: ShardPlan, sizes: Sequence[int],
                   minIdt, minLen, cov_fn: Callable, covs: Sequence[Tuple[str, int]], filter_fn: Callable, strict_self: bool = False):
    """The exchange + annotation that follows a rank's alignment of its block. hits: the rank's rows with LOCAL ids.
    covs: [(name, cov)] coverage passes; with strict_self the pass named 'intra' sees the rows with t_id == q_id and every
    other pass the rest (wrappers.py:1016). Returns on rank 0 (hit table [n,10] with global ids, {name: segments [m,3]}),
    None elsewhere. Collectives: one all-gather of the filtered hit tables, one gather of segments per pass."""
    dist = _dist()
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    n = len(hits['t_id'])
    hits = dict(hits)
    if n:
        hits['t_id'] = np.asarray(t_idx, dtype=np.int32)[hits['t_id']]
        hits['q_id'] = np.asarray(q_idx, dtype=np.int32)[hits['q_id']]
    keep = filter_fn(hits, minLen, minIdt) if n else np.zeros(0, dtype=bool)
    table = np.stack([hits[f][keep] for f in HIT_FIELDS], axis=1).astype(np.int32) if n else np.zeros((0, 10), np.int32)
    full = gather_rows(table, dst=None)                      # all ranks receive the whole filtered table
    mine = plan.owner(full[:, 0]) == rank if len(full) else np.zeros(0, dtype=bool)
    out = {}
    for name, cov in covs:
        m = mine
        if strict_self and len(full):
            same = full[:, 0] == full[:, 1]
            m = mine & (same if name == 'intra' else ~same)
        if len(full) and m.any():
            c, s, e = cov_fn(full[m, 0], full[m, 3], full[m, 4], sizes, cov, minLen)
            seg = np.stack([c, s, e], axis=1).astype(np.int32)
        else:
            seg = np.zeros((0, 3), np.int32)
        g = gather_rows(seg, dst=0)
        if rank == 0:
            out[name] = g[np.lexsort((g[:, 1], g[:, 0]))] if len(g) else g
    if rank != 0:
        return None
    return full, out
