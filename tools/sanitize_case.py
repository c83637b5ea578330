"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): hubs, multi-tile groups,
sub-warp lists, invalid pairs, both grouping modes, id-range passes."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
graph = importlib.import_module('bipartite-link-prediction_b200.graph')
from oracle import c_oracle
rng = np.random.default_rng(9)
n_users, n_biz = 3000, 400
eu = [rng.integers(0, n_users, 4000), np.arange(0, 1500), np.full(300, 7)]
eb = [rng.integers(0, n_biz, 4000), np.zeros(1500, np.int64), np.arange(0, 300)]
eu, eb = np.concatenate(eu), np.concatenate(eb)
pu = np.concatenate([np.sort(rng.integers(-1, n_users, 1500)), np.full(300, 7), rng.integers(0, n_users, 300)])
pv = np.concatenate([rng.integers(-1, n_biz, 1500), np.arange(300), np.zeros(300, np.int64)])
G = graph.BipartiteGraph(n_users, n_biz, eu, eb)
print(G.info())
want = c_oracle.score_pair_arrays(n_users, n_biz, eu, eb, pu, pv)
for env in ({}, {'BLP_RANGES': '3'}, {'BLP_GROUPING': 'sort'}):
    for k in ('BLP_RANGES', 'BLP_GROUPING'):
        os.environ.pop(k, None)
    os.environ.update(env)
    got = G.score_pairs_host(pu, pv)
    bad = sum(int((got[k].astype(np.int64) != want[k].astype(np.int64)).sum())
              for k in ('u_cn', 'u_union', 'b_cn', 'b_union', 'pa'))
    print(env, 'int mismatches', bad)
    assert bad == 0
print('ok')
