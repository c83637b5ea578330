"""Developer tool: end-to-end host-buffer pipeline time for several chunkings."""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
graph = importlib.import_module('bipartite-link-prediction_b200.graph')
synth = importlib.import_module('bipartite-link-prediction_b200.synth')
cfg, eu, eb, pu, pv = synth.make_config('C2')
G = graph.BipartiteGraph(cfg['n_users'], cfg['n_biz'], eu, eb)
n = pu.size
sess = G.host_session(n)
hu, hb = sess.pinned_inputs(n)
hu[:] = pu; hb[:] = pv
for chunks, lead in ((4, 0), (8, 0), (8, 1), (8, 2), (8, 3), (16, 3), (16, 5), (12, 3)):
    for _ in range(2):
        sess.score_pinned(n, user_chunks=chunks, lead_chunks=lead)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(8):
        sess.score_pinned(n, user_chunks=chunks, lead_chunks=lead)
    dt = (time.perf_counter() - t0) / 8
    print('chunks %2d lead %d: %.2f ms  -> %.0f M pairs/s' % (chunks, lead, dt * 1e3, n / dt / 1e6), flush=True)
