"""One-off runs of BASELINE config 5 (`mimeo self` on the plant-like genome, up to 1 Gbp) outside the bench contract: a run of
minutes per step cannot afford three warm-up steps. Launch with torchrun for N > 1 ranks:
    MB2_C5_MBP=1000 python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_c5.py --steps 1 --warm 0
Prints one JSON object from rank 0 (NOT a bench.py line): per-step wall time (max over ranks), per-rank kernel times, counters."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=1)
    ap.add_argument('--warm', type=int, default=0)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bench
    from mimeo_b200 import _lib
    rank, local_rank, world = bench.env_rank()
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    torch.cuda.set_device(local_rank)
    _lib.init(local_rank)
    t0 = time.time()
    wl = bench.C5Workload(rank)
    t_gen = time.time() - t0
    print(f'[rank {rank}] genome generated in {t_gen:.1f} s ({wl.mbp:.0f} Mbp)', file=sys.stderr, flush=True)
    t0 = time.time()
    wl.to_device(torch, torch.device('cuda', local_rank))
    t_up = time.time() - t0
    for _ in range(args.warm):
        wl.step_resident()
    _lib.prof_reset(); _lib.prof_enable(True)
    times = []
    for _ in range(args.steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.time()
        wl.step_resident()
        torch.cuda.synchronize()
        dt = time.time() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        times.append(dt)
        print(f'[rank {rank}] step {dt:.1f} s', file=sys.stderr, flush=True)
    _lib.prof_enable(False)
    tags = ('seed_table_build', 'seed_scan', 'surv_sort', 'hsp_extend', 'hsp_sort', 'chain', 'gapped', 'gp_forward', 'gp_walk', 'hit_filter_sort')
    prof = {t: _lib.prof_get(t)[0] / max(args.steps, 1) for t in tags}
    rec = {'rank': rank, 'kernels_ms_per_step': prof, 'stage_counters': wl.stats, 'block': [len(wl.t_idx), len(wl.q_idx)],
           'block_mbp': [sum(wl.tsizes[i] for i in wl.t_idx) / 1e6, sum(wl.qsizes[i] for i in wl.q_idx) / 1e6],
           'scratch_reserved_gb': None}
    recs = [rec]
    if world > 1:
        recs = [None] * world
        dist.all_gather_object(recs, rec)
    if rank == 0:
        print(json.dumps({'what': 'one-off C5 run (not a bench.py contract line)', 'workload': wl.name, 'n_gpus': world, 'steps': args.steps, 'warm_steps': args.warm,
                          'seconds_per_step': times, 'mbp_per_s': wl.mbp / (sum(times) / len(times)), 'generate_s': t_gen, 'upload_s': t_up,
                          'partition': f'{wl.plan.gt} target groups x {wl.plan.gq} query groups, balance {wl.plan.balance:.3f}',
                          'rows': getattr(wl, 'nhits', None), 'segments': getattr(wl, 'nseg', None), 'ranks': recs}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
