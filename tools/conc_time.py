"""Developer tool: sequential vs two-stream score_pairs timing."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
graph = importlib.import_module('bipartite-link-prediction_b200.graph')
synth = importlib.import_module('bipartite-link-prediction_b200.synth')
cfg, eu, eb, pu, pv = synth.make_config('C2')
G = graph.BipartiteGraph(cfg['n_users'], cfg['n_biz'], eu, eb, device=0)
du, dv = torch.from_numpy(pu).cuda(), torch.from_numpy(pv).cuda()
outs = None
for conc in (False, True, False, True):
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        outs = G.score_pairs(du, dv, out=outs, concurrent=conc)
        e1.record(); e1.synchronize()
        if it >= 3:
            print('concurrent=%s  %.3f ms  (user kernel %.3f, business kernel %.3f)' % (conc, e0.elapsed_time(e1), G.score_stats(0)['score_ms'], G.score_stats(1)['score_ms']), flush=True)
