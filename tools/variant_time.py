"""Developer tool: build library variants with extra -D flags and time both sides on a config.
usage: variant_time.py CONFIG "flags1" "flags2" ...   (each flags string may also carry ENV=val)"""
import importlib, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cfgname = sys.argv[1]
variants = sys.argv[2:]
if os.environ.get('BLP_VARIANT_CHILD') is None:
    for i, v in enumerate(variants):
        env = dict(os.environ, BLP_VARIANT_CHILD=str(i))
        flags = []
        for tok in v.split():
            if tok.startswith('-'):
                flags.append(tok)
            elif '=' in tok:
                k, val = tok.split('=', 1)
                env[k] = val
        env['BLP_VARIANT_FLAGS'] = ' '.join(flags)
        print('=== variant', repr(v), flush=True)
        subprocess.run([sys.executable, __file__, cfgname, v], env=env)
    sys.exit(0)
import numpy as np, torch
L = importlib.import_module('bipartite-link-prediction_b200._lib')
out = os.path.join(ROOT, 'gpurun_out', 'libblp_var%s.so' % os.environ['BLP_VARIANT_CHILD'])
os.makedirs(os.path.dirname(out), exist_ok=True)
subprocess.run(L.nvcc_command(out=out, extra=tuple(os.environ['BLP_VARIANT_FLAGS'].split())), check=True)
L.LIB_PATH = out
graph = importlib.import_module('bipartite-link-prediction_b200.graph')
synth = importlib.import_module('bipartite-link-prediction_b200.synth')
cfg, eu, eb, pu, pv = synth.make_config(cfgname)
G = graph.BipartiteGraph(cfg['n_users'], cfg['n_biz'], eu, eb, device=0)
du, dv = torch.from_numpy(pu).cuda(), torch.from_numpy(pv).cuda()
for side in (0, 1):
    ms = []
    for it in range(4):
        out_t = G.score_side(side, du, dv, want_pa=(side == 0))
        torch.cuda.synchronize()
        ms.append(G.score_stats(side)['score_ms'])
    st = G.score_stats(side)
    print('  side %d: kernel ms %s  ctas %d x %d thr, cn sum %d' % (side, ' '.join('%.3f' % m for m in ms[1:]), st['ctas'], st['threads_per_cta'], int(out_t['cn'].sum())), flush=True)
