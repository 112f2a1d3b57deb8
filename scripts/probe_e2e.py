import sys, time
sys.path.insert(0, '.')
import numpy as np
import bench
from mimeo_b200 import _lib, align as A, engine, coverage
from mimeo_b200.genome import Genome, align_params
_lib.init()
wl = bench.SelfWorkload(0)
import torch
wl.to_device(torch, torch.device('cuda', 0))
def T(): _lib.sync(); return time.perf_counter()
for it in range(3):
    t0 = T()
    G = Genome(wl.names, [t.numpy() for t in wl.pinned]); t1 = T()
    hits, stats = A.align(G, G, align_params(3000)); t2 = T()
    keep = engine.filter_hits(hits, 100, 80); t3 = T()
    intra = (hits['t_id'] == hits['q_id']) & keep; inter = keep & ~intra
    s1 = coverage.coverage_segments(hits['t_id'][inter], hits['start1'][inter], hits['end1'][inter], wl.sizes, 3, 100)
    s2 = coverage.coverage_segments(hits['t_id'][intra], hits['start1'][intra], hits['end1'][intra], wl.sizes, 4, 100); t4 = T()
    blocks = A.tab_blocks(hits, wl.names, wl.names, 100, 80); t5 = T()
    G.close(); t6 = T()
    print('genome %.1f  align %.1f  filter %.1f  coverage x2 %.1f  tab_blocks %.1f  close %.1f  total %.1f ms' % tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t6 - t0)))
for it in range(4):
    t0 = T(); wl.step_e2e(); t1 = T()
    print('bench step_e2e %.1f ms' % (1e3 * (t1 - t0)))
for it in range(3):
    t0 = T(); wl.step_resident(); t1 = T()
    print('bench step_resident %.1f ms' % (1e3 * (t1 - t0)))
