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
import numpy as np
ref = {k: v.copy() for k, v in sess.score_pinned_py(n).items()}
for fn, name, growth in ((sess.score_pinned, 'native, equal slices', '1.0'), (sess.score_pinned, 'native, growth 1.5', '1.5'),
                         (sess.score_pinned, 'native, growth 2', '2.0'), (sess.score_pinned_py, 'python pipeline', '1')):
    os.environ['BLP_SLICE_GROWTH'] = growth
    for chunks, lead, biz in ((4, 1, 2), (5, 1, 2), (6, 1, 2), (6, 2, 2), (8, 2, 2), (4, 1, 1)):
        for _ in range(2):
            out = fn(n, user_chunks=chunks, lead_chunks=lead, biz_chunks=biz)
        same = all(np.array_equal(out[k], ref[k]) for k in ref)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(8):
            fn(n, user_chunks=chunks, lead_chunks=lead, biz_chunks=biz)
        dt = (time.perf_counter() - t0) / 8
        print('%-28s user slices %2d lead %d biz slices %d: %.2f ms  -> %.0f M pairs/s  %s' %
              (name, chunks, lead, biz, dt * 1e3, n / dt / 1e6, 'same results' if same else 'MISMATCH'),
              flush=True)
