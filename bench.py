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


def ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` on this workload, from the committed
    `ncu --set full` capture (profiles/roofline_traffic.json; a number measured under the profiler, so it is only ever
    reported as traffic, never timed). None when no capture of this (workload, kernel) pair is committed."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json')) as f:
            rec = json.load(f)[workload][kernel]
        return rec['dram_bytes_per_launch'], rec['source']
    except Exception:
        return None, None


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
            # nvidia-smi attaches to the driver while it starts up, which stalls this process's CUDA calls for a while: let
            # it reach its steady 100 ms polling (first row printed) BEFORE the timed region begins, not inside it
            t0 = time.perf_counter()
            while not self.rows and self.proc.poll() is None and time.perf_counter() - t0 < 10.0:
                time.sleep(0.01)
            self.rows_before = len(self.rows)
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



def _oracle_pair_job(args):
    """One (target, query) scaffold pair through the CPU LASTZ-restatement (both strands) -- runs in a worker process."""
    tcodes, qcodes, hspthresh = args
    from oracle import lastz_oracle as lo
    st = lo.Stats()
    t0 = time.perf_counter()
    tix = lo.TargetIndex(tcodes)          # LASTZ rebuilds its seed table for every pair: that cost belongs to the baseline
    p = lo.default_params(hspthresh)
    n = 0
    for q in (qcodes, lo.revcomp_codes(qcodes)):
        n += len(lo.align_tile(tix, q, p, st))
    return time.perf_counter() - t0, n, st.as_dict()


class SelfWorkload:
    """BASELINE config 1: `mimeo self` on a synthetic 5 Mbp genome (10 scaffolds x 500 kbp, 20 planted repeat families of
    5-30 copies at ~80 % pairwise identity), minIdt 80, minLen 100, minCov 3, intraCov 4, --strictSelf.
    Under N ranks every rank annotates its own genome of the batch (weak scaling: one genome per GPU, no collective)."""
    name = 'C1: mimeo self, synthetic 5 Mbp genome (10 x 500 kbp, 20 repeat families at ~80% identity), minIdt 80 minLen 100 minCov 3 intraCov 4 strictSelf'
    NSCAF, SCAF_LEN, NFAM = 10, 500_000, 20
    MIN_IDT, MIN_LEN, MIN_COV, INTRA_COV, HSPTHRESH = 80, 100, 3, 4, 3000
    dtype = 'int32'

    def __init__(self, rank):
        from tests.helpers import synth_genome
        g = synth_genome(1001 + rank, self.NSCAF, self.SCAF_LEN, self.NFAM, copies=(5, 30), fam_len=(300, 3000), sub=0.106, indel=0.005)
        self.names = sorted(g, key=lambda s: s.encode())
        self.seqs = [g[n] for n in self.names]
        self.sizes = [len(x) for x in self.seqs]
        self.mbp = sum(self.sizes) / 1e6

    def to_device(self, torch, dev):
        from mimeo_b200.genome import Genome
        self.pinned = [torch.from_numpy(x.copy()).pin_memory() for x in self.seqs]
        self.T = Genome(self.names, self.seqs)
        self.Tboth = self.T.both_strands()
        self.h2d_bytes = sum(self.sizes)

    def step_resident(self):
        from mimeo_b200 import engine
        inter, intra, hits, stats = engine.self_segments(self.T, self.Tboth, self.sizes, self.MIN_IDT, self.MIN_LEN, self.MIN_COV,
                                                         self.INTRA_COV, self.HSPTHRESH, True)
        self.stats, self.nhits, self.nseg = stats, len(hits['t_id']), len(inter[0]) + len(intra[0])
        return inter, intra

    def step_e2e(self):
        """Host ASCII (pinned) in -> .tab rows and GFF3 rows out, every copy inside."""
        from mimeo_b200 import align as A, engine
        from mimeo_b200.genome import Genome
        T = Genome(self.names, [t.numpy() for t in self.pinned])
        try:
            inter, intra, hits, stats = engine.self_segments(T, None, self.sizes, self.MIN_IDT, self.MIN_LEN, self.MIN_COV,
                                                             self.INTRA_COV, self.HSPTHRESH, True)
        finally:
            T.close()
        blocks = A.tab_blocks(hits, self.names, self.names, self.MIN_LEN, self.MIN_IDT)
        ntab = sum(len(v) for v in blocks.values())
        gff = [f'{self.names[int(c)]}\tmimeo-self\tSelf_Repeat\t{int(s)}\t{int(e)}\t.\t+\t.\tID=Self_Repeat_{k + 1:05d}\n'
               for k, (c, s, e) in enumerate(zip(*inter))]
        self.d2h_bytes = 40 * len(hits['t_id']) + 12 * (len(inter[0]) + len(intra[0]))
        return ntab, len(gff)

    def kernel_bytes(self):
        """Algorithmic bytes of the HBM-bound seeding kernels, per strand-launch."""
        Q = T = sum(self.sizes)
        S = self.stats['seed_hits'] / 2.0
        S_out = self.stats['survivors'] / 2.0
        return {
            'seed_table_build': T // 4 * 2 + 4 * T + 2 * 4 * (1 << 24),
            'seed_scan': Q // 4 + Q * 13 * 4 + 4 * S + 8 * S_out,
        }

    def int_cells(self):
        return {'seed_scan': self.stats['stage1_cells'], 'hsp_extend': self.stats['stage2_cells'], 'gapped': self.stats['gapped_cells']}

    def cpu_reference(self, threads, npairs=None):
        """CPU LASTZ-restatement on a bounded, stratified sample of the 100 ordered scaffold pairs (always includes self
        pairs in proportion), extrapolated by pair count; plus nothing else (annotation stages are negligible here)."""
        from concurrent.futures import ProcessPoolExecutor
        from oracle import lastz_oracle as lo
        enc = [lo.encode(x) for x in self.seqs]
        n = len(enc)
        if npairs is None:
            npairs = max(2, min(n * n, 2 * threads))
        # stratified: one self pair per ceil(n) sampled pairs, like the full schedule (n self pairs of n*n)
        rng = np.random.default_rng(7)
        pairs = [(0, 0)] + [tuple(map(int, rng.integers(0, n, 2))) for _ in range(npairs - 1)]
        pairs = [(a, b if (k == 0 or a != b) else (b + 1) % n) for k, (a, b) in enumerate(pairs)]
        jobs = [(enc[a], enc[b], self.HSPTHRESH) for a, b in pairs]
        t0 = time.perf_counter()
        if threads > 1:
            with ProcessPoolExecutor(max_workers=threads) as ex:
                res = list(ex.map(_oracle_pair_job, jobs))
        else:
            res = [_oracle_pair_job(j) for j in jobs]
        wall = time.perf_counter() - t0
        n_self = sum(1 for a, b in pairs if a == b)
        t_self = np.mean([r[0] for (a, b), r in zip(pairs, res) if a == b])
        t_cross = np.mean([r[0] for (a, b), r in zip(pairs, res) if a != b])
        cpu_seconds_full = n * t_self + n * (n - 1) * t_cross           # all n*n ordered pairs, one core
        est = cpu_seconds_full / threads if threads > 1 else cpu_seconds_full
        sample = (f'{len(pairs)} of {n * n} ordered scaffold pairs ({n_self} self) through the C LASTZ-restatement incl. per-pair seed-table '
                  f'build, both strands; {wall:.1f} s wall on {threads} process(es); full-genome time extrapolated by pair class: '
                  f'{cpu_seconds_full:.0f} core-seconds')
        return est, threads, sample


class ShardedSelfWorkload(SelfWorkload):
    """NOT a BASELINE config: a C5-SHAPED genome scaled to 100 Mbp (50 scaffolds x 2 Mbp, 30 repeat families of 20-100
    copies, 2-10 kbp, ~92 % identity) used to exercise the multi-GPU path: ONE genome, target scaffolds row-sharded over
    the ranks (strong scaling), hits and segments gathered to rank 0 (the only collective)."""
    name = 'C5-shaped self-alignment scaled to 100 Mbp (50 x 2 Mbp, 30 families x 20-100 copies of 2-10 kbp at ~92% identity), target scaffolds sharded over ranks'
    NSCAF, SCAF_LEN, NFAM = 50, 2_000_000, 30
    scaling = 'strong'

    def __init__(self, rank):
        from tests.helpers import synth_genome
        g = synth_genome(1005, self.NSCAF, self.SCAF_LEN, self.NFAM, copies=(20, 100), fam_len=(2000, 10000), sub=0.04, indel=0.004)
        self.names = sorted(g, key=lambda s: s.encode())
        self.seqs = [g[n] for n in self.names]
        self.sizes = [len(x) for x in self.seqs]
        self.mbp = sum(self.sizes) / 1e6

    def to_device(self, torch, dev):
        from mimeo_b200 import parallel
        from mimeo_b200.genome import Genome
        import torch.distributed as dist
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
        self.mine = parallel.partition_targets(self.sizes, world)[rank]
        self.Q = Genome(self.names, self.seqs)
        self.Qb = self.Q.both_strands()
        self.T = self.Q if world == 1 else Genome([self.names[i] for i in self.mine], [self.seqs[i] for i in self.mine])
        self.pinned = []
        self.h2d_bytes = sum(self.sizes)
        self.world = world

    def _align_fn(self, t_idx):
        from mimeo_b200 import align as A
        from mimeo_b200.genome import align_params
        hits, stats = A.align(self.T, self.Q, align_params(self.HSPTHRESH), Q_aux=self.Qb, t_same_q=None if self.T is self.Q else self.mine)
        self.stats = stats
        return hits

    def step_resident(self):
        from mimeo_b200 import coverage, engine, parallel
        out = parallel.self_sharded(self.names, self.seqs, self.MIN_IDT, self.MIN_LEN, self.MIN_COV, self.INTRA_COV, self.HSPTHRESH, True,
                                    align_fn=self._align_fn, coverage_fn=coverage.coverage_segments, filter_fn=engine.filter_hits)
        if out is not None:
            self.nhits, self.nseg = len(out[0]['t_id']), len(out[1]) + len(out[2])
        return out

    def step_e2e(self):
        out = self.step_resident()
        self.d2h_bytes = 0 if out is None else 40 * self.nhits + 12 * self.nseg
        return out

    def mbp_per_rank(self, world):
        return self.mbp / world        # strong scaling: the job is one genome

    def cpu_reference(self, threads, npairs=None):
        return SelfWorkload.cpu_reference(self, threads, npairs=max(2, min(8, threads)))


class C5Workload(ShardedSelfWorkload):
    """BASELINE config 5: `mimeo self` on a 1 Gbp synthetic plant-like genome (40 % repeats, 500 scaffolds), one genome whose
    target scaffolds are row-sharded over the ranks (strong scaling). MB2_C5_MBP scales the genome down for trial runs (the
    workload name then says so)."""
    NSCAF = 500
    scaling = 'strong'

    def __init__(self, rank):
        from tests.helpers import synth_c5
        mbp = float(os.environ.get('MB2_C5_MBP', '1000'))
        nscaf = max(8, int(round(self.NSCAF * mbp / 1000.0)))
        g = synth_c5(1005, int(mbp * 1e6), nscaf)
        self.names = sorted(g, key=lambda s: s.encode())
        self.seqs = [g[n] for n in self.names]
        self.sizes = [len(x) for x in self.seqs]
        self.mbp = sum(self.sizes) / 1e6
        self.name = ('C5: mimeo self, synthetic plant-like genome %.0f Mbp, %d scaffolds (log-normal lengths), 40%% repeats (families of '
                     '2-10 kbp, 50-2000 copies, 75-98%% identity), minIdt 80 minLen 100 minCov 3 intraCov 4 strictSelf, targets sharded over ranks'
                     % (self.mbp, nscaf)) + ('' if mbp == 1000 else ' [SCALED-DOWN TRIAL of the 1 Gbp config]')

WORKLOADS = {'self': SelfWorkload, 'cov': CoverageWorkload, 'c5s': ShardedSelfWorkload, 'c5': C5Workload}


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


def quick_config2(steps=5, warmup=3):
    """BASELINE config 2 (10 M hits / 100 Mbp) measured by this same script in a fresh process (`--workload cov`): same arms,
    same timing rules as the headline; the parent holds the GPU idle meanwhile. Returns the child's line, trimmed."""
    env = dict(os.environ, MB2_BENCH_NO_C2='1')
    for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE'):
        env.pop(k, None)
    out = subprocess.run([sys.executable, os.path.abspath(__file__), '--workload', 'cov', '--steps', str(steps), '--warmup', str(warmup)],
                         env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    if out.returncode != 0:
        raise RuntimeError(f'child exited {out.returncode}: {out.stderr[-200:]}')
    d = json.loads(out.stdout.strip().splitlines()[-1])
    r = d['roofline']
    return {'workload': d['config']['workload'], 'steps': d['steps'], 'warmup': d['warmup'], 'ms_per_step': d['ms_per_step'],
            'value': d['value'], 'unit': d['unit'], 'e2e': d['e2e'], 'gpu_launches': d['gpu_launches'], 'clocks': d['clocks'],
            'cpu_baseline': d['cpu_baseline'],
            'roofline': {k: r.get(k) for k in ('kernel', 'achieved', 'peak', 'frac', 'traffic', 'ms_per_launch',
                                               'algorithmic_bytes_per_launch', 'kernels_ms_per_step', 'stage_survey_model')}}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from mimeo_b200 import _lib
    rank, local_rank, world = env_rank()
    if world > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
            os.environ['NCCL_DEBUG'] = 'WARN'          # keep NCCL's version banner out of stdout: rank 0 prints ONE JSON line
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
    if rank == 0:
        sampler.start()
    barrier()
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
    scaling = getattr(wl, 'scaling', 'weak')
    job_mbp = wl.mbp if scaling == 'strong' else world * wl.mbp
    value = job_mbp / (ms_step / 1e3)

    # ---- roofline of the dominant HBM-bound kernel (same timed region, CUDA events on the library stream)
    peak, peak_src = measured_peaks()
    kb = wl.kernel_bytes()
    tags = [t for t in kb if not t.startswith('_')]
    all_tags = tags + [t for t in ('surv_sort', 'hsp_extend', 'hsp_sort', 'chain', 'gapped', 'gp_forward', 'gp_walk', 'cov_events', 'cov_bin_events', 'cov_tile') if t not in tags]
    prof = {t: _lib.prof_get(t) for t in all_tags}
    dom = max(tags, key=lambda t: prof[t][0])
    dms, dcnt = prof[dom]
    per_launch_ms = dms / max(dcnt, 1)
    achieved = kb[dom] / (per_launch_ms / 1e3) / 1e9 if per_launch_ms > 0 else 0.0
    traffic, traffic_src = ncu_traffic(args.workload, dom)
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src, 'ms_per_launch': per_launch_ms, 'algorithmic_bytes_per_launch': kb[dom],
                'kernels_ms_per_step': {t: prof[t][0] / args.steps for t in all_tags if prof[t][1]},
                'note': 'the dominant HBM-bound launch group of this workload; kernels bound by the integer pipe (x-drop and y-drop '
                        'extension, the largest share of a self-alignment step) are reported under gcups, not against HBM'}
    if '_stage_survey' in kb:
        roofline['stage_survey_model'] = {'bytes': kb['_stage_survey'], 'achieved': kb['_stage_survey'] / (ms_step / 1e3) / 1e9,
                                          'frac': kb['_stage_survey'] / (ms_step / 1e3) / 1e9 / peak}
    extra = {}
    if hasattr(wl, 'int_cells'):
        # integer-pipe kernels: GCUPS = DP/extension cells actually evaluated / kernel seconds
        cells = wl.int_cells()
        sm = _lib.lib().mb2_sm_count()
        int_peak = sm * 128 * 1.965e9                                   # INT32 lane-ops/s at max clock
        budget = {'seed_scan': 6, 'hsp_extend': 6, 'gapped': 12}
        gc = {}
        for t, c in cells.items():
            ms = prof[t][0] / args.steps
            if ms > 0:
                g = c / (ms / 1e3) / 1e9
                gc[t] = {'gcups': g, 'cells_per_step': c, 'ms_per_step': ms, 'frac_of_int_roofline': g * 1e9 * budget[t] / int_peak}
        tot_cells = sum(cells.values())
        tot_ms = sum(prof[t][0] for t in cells) / args.steps
        extra['gcups'] = {'value': tot_cells / (tot_ms / 1e3) / 1e9 if tot_ms > 0 else 0.0, 'unit': 'GCUPS (ungapped + gapped cells / kernel seconds)',
                          'per_kernel': gc, 'int_roofline_ops_per_s': int_peak, 'ops_per_cell_budget': budget}
        extra['stage_counters'] = wl.stats

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
    e2e = {'value': job_mbp / (e2e_ms / 1e3), 'unit': 'Mbp/s', 'ms_per_step': e2e_ms,
           'h2d_bytes_per_step': wl.h2d_bytes, 'd2h_bytes_per_step': wl.d2h_bytes}

    # ---- BASELINE config 2 beside the headline (single GPU, default workload only): the coverage/threshold stage on its own
    # 10 M-hit table, same timing rules, a few steps; reported as an extra key, never mixed into `value`
    if rank == 0 and world == 1 and args.workload == 'self' and not os.environ.get('MB2_BENCH_NO_C2'):
        try:
            extra['config2_coverage_stage'] = quick_config2()
        except Exception as e:      # the headline line must not depend on this leg
            extra['config2_coverage_stage'] = {'error': f'{type(e).__name__}: {e}'[:300]}

    cpu = None
    if rank == 0 and world == 1:
        dt, cores, sample = wl.cpu_reference(1)
        cpu = {'value': wl.mbp / dt, 'unit': 'Mbp/s', 'cores': cores, 'kind': 'port', 'sample': sample, 'seconds': dt}

    if rank == 0:
        print(json.dumps({
            'metric': 'self-alignment Mbp/sec (annotated genome Mbp per second of hot-path time)',
            'value': value, 'unit': 'Mbp/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': scaling, 'vs_baseline': None,
            'dtype': wl.dtype, 'data': 'synthetic',
            'config': {'workload': wl.name, 'l2': 'flushed between timed steps (256 MiB memset, outside the timed interval)',
                       'sharding': 'one genome (scaffold group) of the batch per rank, no data-path collective'},
            'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks, **extra,
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='self', choices=sorted(WORKLOADS))
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
