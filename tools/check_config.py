"""Developer tool: parity of a BASELINE.json config (sampled pairs) against the C oracle + timing."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
graph = importlib.import_module('bipartite-link-prediction_b200.graph')
synth = importlib.import_module('bipartite-link-prediction_b200.synth')
from oracle import c_oracle
name = sys.argv[1]
n_time = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
n_check = int(sys.argv[3]) if len(sys.argv) > 3 else 200_000   # 0 = timing only
t = time.time()
cfg, eu, eb, pu, pv = synth.make_config(name, n_pairs=n_time)
du_, db_ = synth.degrees(cfg['n_users'], cfg['n_biz'], eu, eb)
print(name, 'gen %.1fs' % (time.time() - t), 'max deg user/biz', du_.max(), db_.max(), 'distinct edges', du_.sum(), flush=True)
G = graph.BipartiteGraph(cfg['n_users'], cfg['n_biz'], eu, eb, device=0)
print(G.info(), flush=True)
du, dv = torch.from_numpy(pu).cuda(), torch.from_numpy(pv).cuda()
for side in (0, 1):
    for it in range(3):
        out = G.score_side(side, du, dv, want_pa=(side == 0))
        torch.cuda.synchronize()
    st = G.score_stats(side)
    print('side %d: kernel %.3f ms (warp-per-group part %.3f), grouping %.3f ms, %d pairs -> %.3g pairs/s (kernel)  ctas %d x %d passes %d' % (
        side, st['score_ms'], st['light_ms'], st['group_ms'], pu.size, pu.size / st['score_ms'] * 1e3, st['ctas'], st['threads_per_cta'], st['range_passes']), flush=True)
if n_check == 0:
    sys.exit(0)
# parity on a strided sample of the same pair list
idx = np.arange(0, pu.size, max(1, pu.size // n_check))[:n_check]
t = time.time()
want = c_oracle.score_pair_arrays(cfg['n_users'], cfg['n_biz'], eu, eb, pu[idx], pv[idx])
print('C oracle on %d pairs: %.1fs' % (idx.size, time.time() - t), flush=True)
got = G.score_pairs_host(pu[idx], pv[idx])
bad = 0
for k in ('u_cn', 'u_union', 'b_cn', 'b_union', 'pa'):
    bad += int((got[k].astype(np.int64) != want[k].astype(np.int64)).sum())
for k in ('u_jaccard', 'b_jaccard'):
    bad += int((got[k] != want[k]).sum())
rel = 0.0
for k in ('u_adamic', 'b_adamic'):
    m = want[k] != 0
    rel = max(rel, float(np.abs(got[k][m] - want[k][m]).max() / 1.0 if not m.any() else np.max(np.abs(got[k][m] - want[k][m]) / want[k][m])))
    bad += int(((got[k] == 0) != (want[k] == 0)).sum())
print('PARITY mismatches (ints, jaccard, adamic zero pattern):', bad, ' max adamic rel err %.3g' % rel, flush=True)
