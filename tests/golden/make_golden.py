"""Generates tests/golden/cases.json with Oracle A (oracle/similarity_oracle.py).

The reference ships no golden vectors and cannot run here, so these
fixtures pin the ORACLE, not the reference: "parity unpinned".  They are produced by the
line-for-line restatement (sets + BFS) and must be reproduced by the two independent oracles
(sparse algebra, plain C) on CPU and by the CUDA path on the GPU.  Re-run:  python make_golden.py
"""
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import similarity_oracle as oa   # noqa: E402

synth = importlib.import_module('bipartite-link-prediction_b200.synth')

CASES = [
    # name, n_users, n_biz, n_reviews, n_pairs, k, seed, shift
    ('tiny', 12, 6, 30, 40, 4, 11, 1.0),
    ('small', 120, 30, 400, 300, 6, 12, 2.0),
    ('hubby', 400, 8, 900, 400, 4, 13, 1.0),        # a handful of very dense businesses
    ('wide', 40, 300, 600, 400, 20, 14, 3.0),       # more businesses than users
    ('sparse', 500, 200, 350, 300, 5, 15, 50.0),    # many degree-1 nodes (adamic's int 0 branch)
]


def main():
    out = []
    for name, nu, nb, nr, npairs, k, seed, shift in CASES:
        eu, eb = synth.make_graph(nu, nb, nr, seed=seed, shift_u=shift, shift_b=shift)
        pu, pv = synth.make_pairs(nu, nb, eu, eb, npairs, k=k, seed=seed + 1, invalid_frac=0.02)
        pu = np.concatenate([pu, eu[:10]])            # candidates that are existing edges
        pv = np.concatenate([pv, eb[:10]])
        ids_eu, ids_eb = synth.shared_ids(nu, eu, eb)
        ids_pu, ids_pv = synth.shared_ids(nu, pu, pv)
        want = oa.score_pair_arrays(ids_eu, ids_eb, ids_pu, ids_pv)
        out.append({'name': name, 'n_users': nu, 'n_biz': nb,
                    'edge_u': eu.tolist(), 'edge_b': eb.tolist(),
                    'pair_u': pu.tolist(), 'pair_b': pv.tolist(),
                    'expect': {k2: [float(x) if isinstance(x, float) else int(x) for x in v]
                               for k2, v in want.items()}})
    with open(os.path.join(HERE, 'cases.json'), 'w') as fh:
        json.dump({'generator': 'tests/golden/make_golden.py (Oracle A, parity unpinned)',
                   'cases': out}, fh)
    print('wrote', len(out), 'cases')


if __name__ == '__main__':
    main()
