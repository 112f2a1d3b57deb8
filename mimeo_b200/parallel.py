"""
Multi-GPU execution of one `mimeo self / x / map` job: one process per GPU (torchrun), work partitioned by TARGET
scaffold (row blocks of the reference's pair grid, utils.get_all_pairs utils.py:92-102): rank r aligns its own target
scaffolds against the whole query genome, so every hit whose name1 belongs to a rank is born there and filtering and
coverage stay rank-local. The only exchange is the final gather of hit rows and segments to rank 0
(`torch.distributed`, NCCL on GPUs / gloo in CPU tests) -- there is no data-path collective.
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


def gather_rows(table: np.ndarray, dst: int = 0, device=None) -> Optional[np.ndarray]:
    """Gather variable-length int32 row tables [n_r, k] from every rank to `dst` (rank order). Returns None elsewhere.
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
    if rank != dst:
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
