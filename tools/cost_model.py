"""Developer tool: where does the user-side kernel's time go -- per group, per pair or per streamed
id?  Times the C2 user side on filtered pair lists (same graph) and prints groups / pairs / ids
beside the kernel time, for a least-squares fit  t = a*groups + b*pairs + c*ids."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
graph = importlib.import_module('bipartite-link-prediction_b200.graph')
synth = importlib.import_module('bipartite-link-prediction_b200.synth')

cfg, eu, eb, pu, pv = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else 'C2')
du, db = synth.degrees(cfg['n_users'], cfg['n_biz'], eu, eb)
G = graph.BipartiteGraph(cfg['n_users'], cfg['n_biz'], eu, eb, device=0)
info = G.info()
hub_deg = info['hub_min_biz_degree']
dv = np.where(pv >= 0, db[np.maximum(pv, 0)], 0)
k = cfg['k']
pos = np.arange(pu.size) % k
first_half = np.arange(pu.size) < pu.size // 2
# light users: no hub business, <= 32 businesses, <= 512 ids walked (kLightCap) by the expansion
key = np.unique(eu.astype(np.int64) * cfg['n_biz'] + eb)
ku, kb = key // cfg['n_biz'], key % cfg['n_biz']
walked = np.bincount(ku, weights=db[kb].astype(np.float64), minlength=cfg['n_users'])
hubs_of = np.bincount(ku, weights=(db[kb] >= hub_deg).astype(np.float64), minlength=cfg['n_users'])
light_user = (hubs_of == 0) & (du <= 32) & (walked <= 512)
light = light_user[np.maximum(pu, 0)] & (pu >= 0)
cases = [
    ('light users', light),
    ('heavy users', ~light),
    ('light users, 2 pairs each', light & (pos < 2)),
    ('heavy users, 2 pairs each', (~light) & (pos < 2)),
    ('all', np.ones(pu.size, bool)),
    ('partner deg < %d' % hub_deg, dv < hub_deg),
    ('partner deg < 512', dv < 512),
    ('partner deg <= 16', dv <= 16),
    ('partner deg > 16', dv > 16),
    ('first K/2 pairs of every user', pos < k // 2),
    ('first half of the users', first_half),
    ('2 pairs per user', pos < 2),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
rows = []
for name, keep in cases:
    u, v = pu[keep], pv[keep]
    ok = (u >= 0) & (v >= 0)
    groups = np.unique(u[ok]).size
    ids = int(dv[keep].sum())
    tu, tv = torch.from_numpy(u).cuda(), torch.from_numpy(v).cuda()
    ms, out = [], None
    for it in range(6):
        flush.zero_()
        out = G.score_side(0, tu, tv, want_pa=True, out=out)
        torch.cuda.synchronize()
        ms.append(G.score_stats(0)['score_ms'])
    t = min(ms[1:])
    rows.append((groups, u.size, ids, t))
    print('%-32s groups %7d  pairs %8d  ids %11d  user-side kernel %.3f ms' % (name, groups, u.size, ids, t),
          flush=True)
A = np.array([[r[0], r[1], r[2]] for r in rows], dtype=np.float64)
y = np.array([r[3] for r in rows])
coef, *_ = np.linalg.lstsq(A, y, rcond=None)
print('fit: %.2f ns/group  %.3f ns/pair  %.4f ns/id   (probe env %s)' %
      (coef[0] * 1e6, coef[1] * 1e6, coef[2] * 1e6, os.environ.get('BLP_PROBE_MIN_DEG', 'default')))
print('residuals ms', np.round(A @ coef - y, 3))
