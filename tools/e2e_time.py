"""Developer tool: end-to-end host-buffer pipeline time (blp_score_pairs_host, the seven reference
outputs, pinned buffers) for several slice plans.  usage: e2e_time.py"""
import importlib, os, sys, time
import numpy as np, torch
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
print('link', sess.measure_link(n), flush=True)
ref = {k: v.copy() for k, v in sess.score_pinned(n).items()}
for chunks, lead, biz in ((5, 1, 2), (4, 1, 2), (4, 1, 1), (3, 1, 1), (3, 1, 2), (5, 1, 1), (6, 1, 2), (8, 1, 2),
                          (5, 2, 2), (4, 2, 2), (5, 1, 2)):
    for _ in range(2):
        out = sess.score_pinned(n, user_chunks=chunks, lead_chunks=lead, biz_chunks=biz)
    same = all(np.array_equal(out[k], ref[k]) for k in ref)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        sess.score_pinned(n, user_chunks=chunks, lead_chunks=lead, biz_chunks=biz)
    dt = (time.perf_counter() - t0) / 10
    print('user slices %2d lead %d biz slices %d: %.2f ms  -> %.0f M pairs/s  %s' %
          (chunks, lead, biz, dt * 1e3, n / dt / 1e6, 'same results' if same else 'MISMATCH'), flush=True)
