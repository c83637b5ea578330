"""Developer tool: business-side kernel time as a function of which outputs are written."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
graph = importlib.import_module('bipartite-link-prediction_b200.graph')
synth = importlib.import_module('bipartite-link-prediction_b200.synth')
cfg, eu, eb, pu, pv = synth.make_config('C2')
G = graph.BipartiteGraph(cfg['n_users'], cfg['n_biz'], eu, eb, device=0)
du, dv = torch.from_numpy(pu).cuda(), torch.from_numpy(pv).cuda()
for want in (('cn', 'union', 'jaccard', 'adamic'), ('cn',), ()):
    for side in (1, 0):
        ms = []
        for it in range(4):
            G.score_side(side, du, dv, want=want)
            torch.cuda.synchronize()
            ms.append(G.score_stats(side)['score_ms'])
        print('side %d outputs %-40s kernel ms %s' % (side, want, ' '.join('%.3f' % m for m in ms[1:])), flush=True)
