"""Developer tool: A/B kernel timing of two prebuilt libraries in one process-per-library run.
usage: ab_time.py CONFIG lib_a.so lib_b.so [repeat]"""
import importlib, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get('BLP_AB_CHILD') is None:
    cfg, libs = sys.argv[1], sys.argv[2:4]
    rep = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    for r in range(rep):
        for lib in libs:
            subprocess.run([sys.executable, __file__, cfg, lib], env=dict(os.environ, BLP_AB_CHILD='1'))
    sys.exit(0)
import torch
L = importlib.import_module('bipartite-link-prediction_b200._lib')
L.LIB_PATH = os.path.abspath(sys.argv[2])
graph = importlib.import_module('bipartite-link-prediction_b200.graph')
synth = importlib.import_module('bipartite-link-prediction_b200.synth')
cfg, eu, eb, pu, pv = synth.make_config(sys.argv[1])
G = graph.BipartiteGraph(cfg['n_users'], cfg['n_biz'], eu, eb, device=0)
du, dv = torch.from_numpy(pu).cuda(), torch.from_numpy(pv).cuda()
res = []
for side in (0, 1):
    ms = []
    for it in range(6):
        G.score_side(side, du, dv, want_pa=(side == 0))
        torch.cuda.synchronize()
        ms.append(G.score_stats(side)['score_ms'])
    res.append(min(ms[1:]))
print('%-28s user %.3f ms  business %.3f ms' % (os.path.basename(sys.argv[2]), res[0], res[1]), flush=True)
