"""Developer timing helper (not the bench contract): per-side CUDA-event times on one config."""
import argparse
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
graph = importlib.import_module('bipartite-link-prediction_b200.graph')
synth = importlib.import_module('bipartite-link-prediction_b200.synth')

ap = argparse.ArgumentParser()
ap.add_argument('--config', default='C2')
ap.add_argument('--pairs', type=int, default=None)
ap.add_argument('--iters', type=int, default=3)
ap.add_argument('--sides', default='ub')
a = ap.parse_args()
t = time.time()
cfg, eu, eb, pu, pv = synth.make_config(a.config, n_pairs=a.pairs)
print('gen %.1fs' % (time.time() - t), flush=True)
t = time.time()
G = graph.BipartiteGraph(cfg['n_users'], cfg['n_biz'], eu, eb, device=0)
print('graph build %.2fs' % (time.time() - t), G.info(), flush=True)
du = torch.from_numpy(pu).cuda()
dv = torch.from_numpy(pv).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for side, name in ((0, 'user'), (1, 'business')):
    if name[0] not in a.sides:
        continue
    out = None
    for it in range(a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = G.score_side(side, du, dv, want_pa=(side == 0), out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print('%s side: %.3f ms  -> %.3g pairs/s  %s' % (name, ms, pu.size / ms * 1e3, G.score_stats(side)), flush=True)
    print('  cn sum', int(out['cn'].sum()), 'aa sum', float(out['adamic'].sum()))
