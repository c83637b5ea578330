"""Generates tests/golden/cases.json and tests/golden/ref_files.json FROM THE REFERENCE ITSELF.

The reference ships no golden vectors, so they are made here by running
its own similarity.py -- oracle/_ref, built by oracle/build_ref.py from /root/reference, with the
SNAP stand-in of oracle/ref_runner.py -- on seeded inputs.  Run it where /root/reference exists:

    python tests/golden/make_golden.py

cases.json   per case: graph + pair arrays (local indices) and `expect`, one entry per pair:
             u_cn, u_jaccard, u_adamic, b_cn, b_jaccard   from the files the reference's users() /
                                                          business() write (similarity.py:61,106)
             b_adamic      the reference's adamic_adar() called as similarity.py:103 spells it
                           (the branch at :102 never fires, so the reference's FILE lacks these)
             in_graph      similarity.py:52,95
             u_union, b_union, pa   NOT reference outputs (no file carries them): from Oracle A
ref_files.json  the six score files of the reference's main() for one small case, verbatim
             (int-vs-float types and the missing in-graph b_adamic entries included).
The fixtures travel; /root/reference does not.  tests/test_oracle.py checks Oracle A, B and C
against them on CPU, tests/test_gpu_parity.py checks the CUDA path.
"""
import importlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_runner as rr          # noqa: E402
from oracle import similarity_oracle as oa   # noqa: E402

synth = importlib.import_module('bipartite-link-prediction_b200.synth')
util = importlib.import_module('bipartite-link-prediction_b200.util')

CASES = [
    # name, n_users, n_biz, n_reviews, n_pairs, k, seed, shift
    ('tiny', 12, 6, 30, 40, 4, 11, 1.0),
    ('small', 120, 30, 400, 300, 6, 12, 2.0),
    ('hubby', 400, 8, 900, 400, 4, 13, 1.0),        # a handful of very dense businesses
    ('wide', 40, 300, 600, 400, 20, 14, 3.0),       # more businesses than users
    ('sparse', 500, 200, 350, 300, 5, 15, 50.0),    # many degree-1 nodes (adamic's int 0 branch)
    ('mid', 1500, 250, 6000, 1500, 10, 16, 3.0),    # C1-like shape, scaled down
]
REF_KEYS = ('u_cn', 'u_jaccard', 'u_adamic', 'b_cn', 'b_jaccard', 'b_adamic', 'in_graph')


def _plain(x):
    return float(x) if isinstance(x, float) else int(x)


def main():
    if not rr.ensure_built():
        raise SystemExit('oracle/_ref cannot be built here (no /root/reference)')
    man = rr.manifest()
    out = []
    for name, nu, nb, nr, npairs, k, seed, shift in CASES:
        eu, eb = synth.make_graph(nu, nb, nr, seed=seed, shift_u=shift, shift_b=shift)
        pu, pv = synth.make_pairs(nu, nb, eu, eb, npairs, k=k, seed=seed + 1, invalid_frac=0.02)
        pu = np.concatenate([pu, eu[:10]])            # candidates that are existing edges
        pv = np.concatenate([pv, eb[:10]])
        ids_eu, ids_eb = synth.shared_ids(nu, eu, eb)
        ids_pu, ids_pv = synth.shared_ids(nu, pu, pv)
        ref = rr.score_pair_arrays(ids_eu, ids_eb, ids_pu, ids_pv)
        port = oa.score_pair_arrays(ids_eu, ids_eb, ids_pu, ids_pv)
        assert all(f is None or f == 0 for f in ref['b_adamic_file'])
        expect = {k2: [_plain(x) for x in ref[k2]] for k2 in REF_KEYS}
        for k2 in ('u_union', 'b_union', 'pa'):
            expect[k2] = [int(x) for x in port[k2]]
        out.append({'name': name, 'n_users': nu, 'n_biz': nb,
                    'edge_u': eu.tolist(), 'edge_b': eb.tolist(),
                    'pair_u': pu.tolist(), 'pair_b': pv.tolist(), 'expect': expect})
    with open(os.path.join(HERE, 'cases.json'), 'w') as fh:
        json.dump({'generator': 'tests/golden/make_golden.py: the reference\'s own similarity.py '
                                '(oracle/_ref) with the SNAP stand-in of oracle/ref_runner.py',
                   'reference_files': man['files'], 'from_reference': list(REF_KEYS),
                   'from_oracle_a': ['u_union', 'b_union', 'pa'], 'cases': out}, fh)
    print('wrote', len(out), 'cases')

    # file-level: the six JSON files of the reference's main(), verbatim
    nu, nb = 60, 20
    eu, eb = synth.make_graph(nu, nb, 200, seed=3, shift_u=1.0, shift_b=1.0)
    pu, pv = synth.make_pairs(nu, nb, eu, eb, 120, k=4, seed=4, invalid_frac=0.05)
    ids_eu, ids_eb = synth.shared_ids(nu, eu, eb)
    ids_pu, ids_pv = synth.shared_ids(nu, pu, pv)
    with tempfile.TemporaryDirectory() as d:
        util.write_edge_list(os.path.join(d, 'graph.txt'), ids_eu, ids_eb)
        examples = synth.examples_dict(ids_pu, ids_pv)
        util.write_json(examples, os.path.join(d, 'examples.json'))
        files = rr.run_main(os.path.join(d, 'examples.json'), os.path.join(d, 'graph.txt'), d)
        graph_lines = open(os.path.join(d, 'graph.txt')).read()
    with open(os.path.join(HERE, 'ref_files.json'), 'w') as fh:
        json.dump({'generator': 'tests/golden/make_golden.py: reference main() (similarity.py:11-18)',
                   'reference_files': man['files'], 'graph_txt': graph_lines,
                   'examples': examples, 'score_files': files}, fh)
    print('wrote ref_files.json (%d pairs)' % sum(len(v) for v in examples.values()))


if __name__ == '__main__':
    main()
