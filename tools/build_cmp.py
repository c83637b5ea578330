"""Developer tool: scoring-kernel time on a host-built vs a device-built graph (striping quality)."""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
graph = importlib.import_module('bipartite-link-prediction_b200.graph')
synth = importlib.import_module('bipartite-link-prediction_b200.synth')
name = sys.argv[1] if len(sys.argv) > 1 else 'C2'
npairs = int(sys.argv[2]) if len(sys.argv) > 2 else None
cfg, eu, eb, pu, pv = synth.make_config(name, n_pairs=npairs)
du, dv = torch.from_numpy(pu).cuda(), torch.from_numpy(pv).cuda()
for build in ('host', 'device', 'host', 'device'):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    G = graph.BipartiteGraph(cfg['n_users'], cfg['n_biz'], eu, eb, build=build)
    tb = time.perf_counter() - t0
    res = []
    for side in (0, 1):
        ms = []
        for it in range(5):
            G.score_side(side, du, dv, want_pa=(side == 0))
            torch.cuda.synchronize()
            ms.append(G.score_stats(side)['score_ms'])
        res.append(min(ms[1:]))
    print('%s build=%-6s %.3f s   user %.3f ms  business %.3f ms' % (name, build, tb, res[0], res[1]), flush=True)
    G.close()
