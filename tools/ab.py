"""Developer tool: A/B prebuilt library variants (tools/build_variants.py) on one GPU.
Every variant runs in its own process: same seeded workload, L2 flushed before every timed call
(as bench.py does), per-side kernel times from blp_score_stats, and a checksum of every output
column so that variants can be checked for bit-identical results.
usage: ab.py CONFIG[:pairs] [--reps R] [--out file.jsonl] lib_or_name ...   ('default' = the product library)"""
import hashlib, importlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parent():
    args = sys.argv[1:]
    cfg, reps, out, libs = args[0], 5, None, []
    i = 1
    while i < len(args):
        if args[i] == '--reps':
            reps = int(args[i + 1]); i += 2
        elif args[i] == '--out':
            out = args[i + 1]; i += 2
        else:
            libs.append(args[i]); i += 1
    rows = []
    for lib in libs:
        env = dict(os.environ, BLP_AB_CHILD='1')
        r = subprocess.run([sys.executable, __file__, cfg, str(reps), lib], env=env, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True)
        last = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ''
        try:
            row = json.loads(last)
        except ValueError:
            row = {'lib': lib, 'error': r.stdout[-1500:]}
        rows.append(row)
        print(json.dumps(row), flush=True)
    if out:
        os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
        with open(out, 'a') as fh:
            for row in rows:
                fh.write(json.dumps(row) + '\n')
    sums = {json.dumps(r.get('checksums'), sort_keys=True) for r in rows if 'checksums' in r}
    print('bit-identical across variants:', len(sums) <= 1, flush=True)


def child():
    import numpy as np, torch
    cfgspec, reps, lib = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    name, _, pairs = cfgspec.partition(':')
    L = importlib.import_module('bipartite-link-prediction_b200._lib')
    if lib != 'default':
        path = lib if os.path.sep in lib else os.path.join(ROOT, 'bipartite-link-prediction_b200', 'variants',
                                                            'libblp_%s.so' % lib)
        L.LIB_PATH = os.path.abspath(path)
    graph = importlib.import_module('bipartite-link-prediction_b200.graph')
    synth = importlib.import_module('bipartite-link-prediction_b200.synth')
    cfg, eu, eb, pu, pv = synth.make_config(name, n_pairs=int(pairs) if pairs else None)
    G = graph.BipartiteGraph(cfg['n_users'], cfg['n_biz'], eu, eb, device=0)
    du, dv = torch.from_numpy(pu).cuda(), torch.from_numpy(pv).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    row = {'lib': lib, 'config': cfgspec, 'pairs': int(pu.size), 'checksums': {}}
    for side, tag in ((0, 'user'), (1, 'business')):
        ms, lms, gms = [], [], []
        for it in range(reps + 2):
            flush.zero_()
            out = G.score_side(side, du, dv, want_pa=(side == 0))
            torch.cuda.synchronize()
            st = G.score_stats(side)
            if it >= 2:
                ms.append(st['score_ms']); lms.append(st['light_ms']); gms.append(st['group_ms'])
        row[tag] = {'score_ms_min': min(ms), 'score_ms_mean': sum(ms) / len(ms), 'light_ms_mean': sum(lms) / len(lms),
                    'group_ms_mean': sum(gms) / len(gms), 'ctas': st['ctas'], 'threads': st['threads_per_cta'],
                    'passes': st['range_passes'], 'groups': st['n_groups'], 'light_groups': st['light_groups']}
        for k, v in out.items():
            row['checksums'][tag + '_' + k] = hashlib.sha1(v.cpu().numpy().tobytes()).hexdigest()[:12]
    print(json.dumps(row), flush=True)


if __name__ == '__main__':
    child() if os.environ.get('BLP_AB_CHILD') else parent()
