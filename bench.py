#!/usr/bin/env python
"""
bench.py -- measures the hot path on B200 (contract: one JSON line on stdout from rank 0).

  python bench.py --gpus N --steps K --warmup W [--workload cov|self] [--impl b200|reference]

A "step" is one pass of the hot path over one batch of synthetic input.
  value : whole-job throughput with inputs already resident in HBM, timed with CUDA events on the
          library's stream, max over ranks.
  e2e   : the same metric through the host-buffer C-ABI call (pinned host memory in, host arrays out,
          H2D/D2H inside the timed region).
  roofline     : the dominant kernel's algorithmic bytes / its CUDA-event time (mb2_prof_*), against
                 MEASURED_PEAKS.json.
  cpu_baseline : the CPU oracle port timed on this box's host cores (rank 0, N=1 only).
--impl reference times the CPU restatement of the reference path (oracle/) with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ------------------------------------------------------------------------------------------------- helpers
def env_rank():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------- workloads
class CoverageWorkload:
    """BASELINE config 2: coverage/threshold stage only, 10 M-hit table over a 100 Mbp genome (50 scaffolds x 2 Mbp),
    minCov 3, minLen 100. Under N ranks every rank owns its own 50-scaffold group (weak scaling, no collective)."""
    name = 'C2: coverage/threshold stage, 10M-hit tab over 100 Mbp (50 x 2 Mbp), minCov 3, minLen 100'
    NCHROM, CHROM_SIZE, NHITS, HOTSPOTS = 50, 2_000_000, 10_000_000, 2000
    MIN_COV, MIN_LEN = 3, 100
    dtype = 'int32'

    def __init__(self, rank):
        from tests.helpers import synth_hits
        self.chrom, self.start, self.end, self.sizes = synth_hits(1002 + rank, self.NCHROM, self.CHROM_SIZE, self.NHITS, self.HOTSPOTS)
        self.mbp = self.NCHROM * self.CHROM_SIZE / 1e6

    def to_device(self, torch, dev):
        self.pinned = [torch.from_numpy(a).pin_memory() for a in (self.chrom, self.start, self.end)]
        self.dev = [t.to(dev) for t in self.pinned]
        self.h2d_bytes = sum(t.numel() * 4 for t in self.pinned)

    def step_resident(self):
        from mimeo_b200 import coverage
        out = coverage.coverage_segments_device(self.dev[0], self.dev[1], self.dev[2], self.sizes, self.MIN_COV, self.MIN_LEN)
        self.nseg = int(out[0].numel())
        return out

    def step_e2e(self):
        from mimeo_b200 import coverage
        out = coverage.coverage_segments(self.pinned[0].numpy(), self.pinned[1].numpy(), self.pinned[2].numpy(),
                                         self.sizes, self.MIN_COV, self.MIN_LEN)
        self.d2h_bytes = sum(a.nbytes for a in out)
        return out

    def kernel_bytes(self):
        """Algorithmic bytes per launch group (DESIGN.md 'roofline'): H hits, R segments."""
        H, R = self.NHITS, self.nseg
        G = int(self.sizes.sum())
        return {
            'cov_events': 12 * H + 8 * H,                       # read (chrom,start,end), write two event keys
            'cov_bin_events': 2 * 2 * (4 + 4 + 4) * H,          # 2 arrays x 2 radix passes x (hist read + scatter read + write)
            'cov_tile': 8 * H + 8 * R,                          # read both event arrays once, write flips
            '_stage_survey': 12 * H + 4 * G + 16 * H + 4 * G + 12 * R,   # dense difference-array model
        }

    def cpu_reference(self, threads):
        """CPU oracle port of the same stage (C restatement of genomecov/merge), full workload."""
        import ctypes
        subprocess.check_call(['make', '-C', os.path.join(ROOT, 'oracle'), 'all'], stdout=subprocess.DEVNULL)
        lib = ctypes.CDLL(os.path.join(ROOT, 'oracle', '_build', 'libannot_oracle.so'))
        lib.ora_coverage_segments.restype = ctypes.c_long
        cap = self.NHITS + 8
        oc, os_, oe = (np.zeros(cap, np.int32) for _ in range(3))
        P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        t0 = time.perf_counter()
        k = lib.ora_coverage_segments(P(self.chrom), P(self.start), P(self.end), ctypes.c_long(self.NHITS), P(self.sizes),
                                      ctypes.c_int(self.NCHROM), ctypes.c_int(self.MIN_COV), ctypes.c_int(self.MIN_LEN),
                                      P(oc), P(os_), P(oe), ctypes.c_long(cap))
        dt = time.perf_counter() - t0
        assert k >= 0
        return dt, 1, 'full workload (10M hits / 100 Mbp), arrays already parsed; single-threaded C port of genomecov+merge'


WORKLOADS = {'cov': CoverageWorkload}


# ------------------------------------------------------------------------------------------------- arms
def run_reference(args):
    rank, _, world = env_rank()
    if rank != 0:
        return
    wl = WORKLOADS[args.workload](0)
    times = []
    for i in range(args.warmup + args.steps):
        dt, cores, sample = wl.cpu_reference(os.cpu_count())
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = wl.mbp / (ms / 1e3)
    print(json.dumps({
        'impl': 'reference', 'metric': 'self-alignment Mbp/sec (annotated genome Mbp per second of hot-path time)',
        'value': val, 'unit': 'Mbp/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': wl.dtype, 'data': 'synthetic',
        'config': {'workload': wl.name},
        'cpu_baseline': {'value': val, 'unit': 'Mbp/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': 'Mbp/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def run_b200(args):
    import torch
    import torch.distributed as dist
    from mimeo_b200 import _lib
    rank, local_rank, world = env_rank()
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    _lib.init(local_rank)
    stream = torch.cuda.ExternalStream(_lib.stream_handle(), device=dev)

    wl = WORKLOADS[args.workload](rank)
    wl.to_device(torch, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxreduce(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident arm (value)
    for _ in range(args.warmup):
        wl.step_resident()
    sampler = ClockSampler(local_rank)
    _lib.prof_reset(); _lib.prof_enable(True)
    launches0 = _lib.launch_count()
    barrier()
    if rank == 0:
        sampler.start()
    total_ms = 0.0
    for _ in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (outside the timed interval)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            wl.step_resident()
            e1.record(stream)
        e1.synchronize()
        total_ms += e0.elapsed_time(e1)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = _lib.launch_count() - launches0
    _lib.prof_enable(False)
    ms_step = maxreduce(total_ms / args.steps)
    value = world * wl.mbp / (ms_step / 1e3)

    # ---- roofline of the dominant kernel (same timed region, CUDA events on the library stream)
    peak, peak_src = measured_peaks()
    kb = wl.kernel_bytes()
    prof = {t: _lib.prof_get(t) for t in kb if not t.startswith('_')}
    dom = max(prof, key=lambda t: prof[t][0])
    dms, dcnt = prof[dom]
    per_launch_ms = dms / max(dcnt, 1)
    achieved = kb[dom] / (per_launch_ms / 1e3) / 1e9 if per_launch_ms > 0 else 0.0
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': None, 'peak_source': peak_src, 'ms_per_launch': per_launch_ms,
                'kernels_ms_per_step': {t: prof[t][0] / args.steps for t in prof},
                'stage_survey_model': {'bytes': kb['_stage_survey'], 'achieved': kb['_stage_survey'] / (ms_step / 1e3) / 1e9,
                                       'frac': kb['_stage_survey'] / (ms_step / 1e3) / 1e9 / peak}}

    # ---- end-to-end arm through the host-buffer C ABI (H2D + D2H inside the timed region)
    for _ in range(max(1, args.warmup // 2)):
        wl.step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        wl.step_e2e()
    torch.cuda.synchronize()
    e2e_ms = maxreduce(1e3 * (time.perf_counter() - t0) / args.steps)
    barrier()
    e2e = {'value': world * wl.mbp / (e2e_ms / 1e3), 'unit': 'Mbp/s', 'ms_per_step': e2e_ms,
           'h2d_bytes_per_step': wl.h2d_bytes, 'd2h_bytes_per_step': wl.d2h_bytes}

    cpu = None
    if rank == 0 and world == 1:
        dt, cores, sample = wl.cpu_reference(1)
        cpu = {'value': wl.mbp / dt, 'unit': 'Mbp/s', 'cores': cores, 'kind': 'port', 'sample': sample, 'seconds': dt}

    if rank == 0:
        print(json.dumps({
            'metric': 'self-alignment Mbp/sec (annotated genome Mbp per second of hot-path time)',
            'value': value, 'unit': 'Mbp/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': wl.dtype, 'data': 'synthetic',
            'config': {'workload': wl.name, 'l2': 'flushed between timed steps (256 MiB memset, outside the timed interval)',
                       'sharding': 'one scaffold group per rank, no data-path collective'},
            'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks,
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='cov', choices=sorted(WORKLOADS))
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
