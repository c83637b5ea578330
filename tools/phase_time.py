"""Developer tool: build a BLP_PHASE_TIMING variant of the library and print per-phase cycles."""
import ctypes, importlib, os, subprocess, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
L = importlib.import_module('bipartite-link-prediction_b200._lib')
out = os.path.join(ROOT, 'gpurun_out', 'libblp_phase.so')
os.makedirs(os.path.dirname(out), exist_ok=True)
subprocess.run(L.nvcc_command(out=out, extra=('-DBLP_PHASE_TIMING',)), check=True)
L.LIB_PATH = out
graph = importlib.import_module('bipartite-link-prediction_b200.graph')
synth = importlib.import_module('bipartite-link-prediction_b200.synth')
cfgname = sys.argv[1] if len(sys.argv) > 1 else 'C2'
cfg, eu, eb, pu, pv = synth.make_config(cfgname, n_pairs=int(sys.argv[2]) if len(sys.argv) > 2 else None)
G = graph.BipartiteGraph(cfg['n_users'], cfg['n_biz'], eu, eb, device=0)
lib = L.load()
du, dv = torch.from_numpy(pu).cuda(), torch.from_numpy(pv).cuda()
names = ['fetch(top barrier..item)', 'exp tile load+scan', 'clear / hub-bitmap OR', 'SET sweep', '-',
         'hop2 finalize', 'pair tile load+scan', 'TEST sweep', 'epilogue']
buf = (ctypes.c_ulonglong * 16)()
for side in (0, 1):
    G.score_side(side, du, dv, want_pa=(side == 0))
    lib.blp_debug_phase_cycles(buf, 1)
    G.score_side(side, du, dv, want_pa=(side == 0))
    lib.blp_debug_phase_cycles(buf, 1)
    st = G.score_stats(side)
    tot = sum(buf[:9])
    print('side', side, 'score_ms %.3f' % st['score_ms'], 'ctas', st['ctas'], 'sum cycles/cta %.3g' % (tot / st['ctas']))
    for i, n in enumerate(names):
        print('   %-28s %5.1f%%' % (n, 100.0 * buf[i] / tot))
