"""CPU tests of the oracles: known-answer table, golden fixtures, three-way agreement.

Three statements of the reference's algorithm -- a line-for-line restatement (A), independent
sparse algebra (B) and plain C (C) -- must agree with each other, with the hand-checked table and
with the fixtures in tests/golden/ that the reference's OWN code wrote (tests/test_reference_pin.py
holds the tests that run that code, oracle/_ref, directly against Oracle A).
"""
import json
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from conftest import pkg
from oracle import algebra_oracle as ob
from oracle import c_oracle as oc
from oracle import similarity_oracle as oa

HERE = os.path.dirname(__file__)
INT_KEYS = ('u_cn', 'u_union', 'b_cn', 'b_union', 'pa')
FLT_KEYS = ('u_jaccard', 'u_adamic', 'b_jaccard', 'b_adamic')


def same(a, b, rtol=1e-12):
    for k in INT_KEYS:
        assert np.array_equal(np.asarray(a[k], dtype=np.int64), np.asarray(b[k], dtype=np.int64)), k
    for k in ('u_jaccard', 'b_jaccard'):
        assert np.array_equal(np.asarray(a[k], dtype=np.float64), np.asarray(b[k], dtype=np.float64)), k
    for k in ('u_adamic', 'b_adamic'):
        np.testing.assert_allclose(np.asarray(a[k], dtype=np.float64),
                                   np.asarray(b[k], dtype=np.float64), rtol=rtol, atol=0, err_msg=k)


def known_answer():
    return json.load(open(os.path.join(HERE, 'golden', 'known_answer.json')))


def test_oracle_a_known_answer():
    ka = known_answer()
    lines = np.array(ka['graph_lines'])
    pu = [p['u'] for p in ka['pairs']]
    pv = [p['v'] for p in ka['pairs']]
    got = oa.score_pair_arrays(lines[:, 0], lines[:, 1], pu, pv)
    want = {k: [p[k] for p in ka['pairs']] for k in INT_KEYS + FLT_KEYS}
    same(got, want, rtol=1e-15)
    # the reference's types: int for cn, int 0 for adamic when nothing contributes
    assert isinstance(got['u_cn'][0], int) and got['u_adamic'][1] == 0 and isinstance(got['u_adamic'][1], int)
    assert got['in_graph'] == [1] * 8 + [0, 0]


def test_oracle_b_and_c_known_answer():
    ka = known_answer()
    lines = np.array(ka['graph_lines'])
    pu = np.array([p['u'] if p['u'] < 6 else -1 for p in ka['pairs']])
    pv = np.array([p['v'] - 10 if 10 <= p['v'] < 14 else -1 for p in ka['pairs']])
    want = {k: [p[k] for p in ka['pairs']] for k in INT_KEYS + FLT_KEYS}
    same(ob.score_pair_arrays(6, 4, lines[:, 0], lines[:, 1] - 10, pu, pv), want)
    same(oc.score_pair_arrays(6, 4, lines[:, 0], lines[:, 1] - 10, pu, pv), want)


def test_path3_is_only_an_upper_bound():
    """(A.A^T.A)[u,v] counts paths, the reference counts distinct nodes."""
    ka = known_answer()
    lines = np.array(ka['graph_lines'])
    A = ob.biadjacency(6, 4, lines[:, 0], lines[:, 1] - 10)
    P3 = (A @ A.T @ A).toarray()
    assert P3[0, 2] == 2 and P3[0, 0] == 4      # pair (0,12) and the existing edge (0,10)
    got = ob.score_pair_arrays(6, 4, lines[:, 0], lines[:, 1] - 10, [0, 0], [2, 0])
    assert list(got['u_cn']) == [2, 2] and list(got['b_cn']) == [2, 1]


def test_golden_cases_reproduced_by_b_and_c():
    cases = json.load(open(os.path.join(HERE, 'golden', 'cases.json')))['cases']
    assert len(cases) >= 5
    for c in cases:
        args = (c['n_users'], c['n_biz'], c['edge_u'], c['edge_b'], c['pair_u'], c['pair_b'])
        same(ob.score_pair_arrays(*args), c['expect'])
        same(oc.score_pair_arrays(*args), c['expect'])


def test_file_level_main_and_reference_bug(tmp_path):
    """similarity.py:11-18 end to end on files, including the dead b_adamic branch (:102)."""
    synth, util = pkg('synth'), pkg('util')
    eu, eb = synth.make_graph(60, 20, 200, seed=3, shift_u=1.0, shift_b=1.0)
    pu, pv = synth.make_pairs(60, 20, eu, eb, 120, k=4, seed=4, invalid_frac=0.05)
    ids_eu, ids_eb = synth.shared_ids(60, eu, eb)
    ids_pu, ids_pv = synth.shared_ids(60, pu, pv)
    util.write_edge_list(str(tmp_path / 'graph.txt'), ids_eu, ids_eb)
    util.write_json(synth.examples_dict(ids_pu, ids_pv), str(tmp_path / 'examples.json'))
    M = ['common_neighbors', 'jaccard', 'adamic_adar']
    uo = [str(tmp_path / n) for n in ('u_cn.json', 'u_jaccard.json', 'u_adamic.json')]
    bo = [str(tmp_path / n) for n in ('b_cn.json', 'b_jaccard.json', 'b_adamic.json')]
    oa.main(str(tmp_path / 'examples.json'), str(tmp_path / 'graph.txt'), M, uo, M, bo,
            reproduce_reference_bug=True)
    ex = util.load_json(str(tmp_path / 'examples.json'))
    n_pairs = sum(len(v) for v in ex.values())
    u_cn = util.load_json(uo[0])
    assert sum(len(v) for v in u_cn.values()) == n_pairs          # same key set as examples
    b_ad = util.load_json(bo[2])
    # with the bug only the literal zeros of out-of-graph pairs are present
    assert all(s == 0 for v in b_ad.values() for s in v.values())
    assert sum(len(v) for v in b_ad.values()) < n_pairs
    # faithful (list membership) and fair (set membership) modes agree
    G = oa.MiniSnapGraph.load_edge_list(str(tmp_path / 'graph.txt'))
    r1 = oa.users(ex, G, M, uo, faithful=True, write=False)
    r2 = oa.users(ex, G, M, uo, faithful=False, write=False)
    assert r1 == r2


graphs = st.tuples(st.integers(2, 25), st.integers(2, 12), st.integers(0, 2 ** 31 - 1),
                   st.integers(1, 120))


@settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
@given(graphs)
def test_three_oracles_agree_on_random_graphs(g):
    n_users, n_biz, seed, n_edges = g
    rng = np.random.default_rng(seed)
    eu = rng.integers(0, n_users, n_edges)          # duplicates, isolated ids, degree-1 nodes, hubs
    eb = rng.integers(0, n_biz, n_edges) if seed % 3 else np.minimum(rng.geometric(0.5, n_edges) - 1, n_biz - 1)
    pu = rng.integers(-1, n_users, 60)
    pv = rng.integers(-1, n_biz, 60)
    b = ob.score_pair_arrays(n_users, n_biz, eu, eb, pu, pv)
    c = oc.score_pair_arrays(n_users, n_biz, eu, eb, pu, pv)
    synth = pkg('synth')
    ids_eu, ids_eb = synth.shared_ids(n_users, eu, eb)
    ids_pu, ids_pv = synth.shared_ids(n_users, pu, pv)
    a = oa.score_pair_arrays(ids_eu, ids_eb, ids_pu, ids_pv)
    same(a, b)
    same(a, c)
    # invariants
    ok = np.asarray(a['in_graph']) == 1
    cn, uni = np.asarray(a['u_cn'])[ok], np.asarray(a['u_union'])[ok]
    assert np.all(cn >= 0) and np.all(uni >= 1) and np.all(cn <= uni)
    aa = np.asarray(a['u_adamic'], dtype=np.float64)[ok]
    assert np.all(aa <= cn / np.log(2.0) + 1e-12)


@settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck))
@given(st.integers(0, 10_000))
def test_set_identities_behind_the_probe_and_table_paths(seed):
    """The CUDA path answers some pairs without streaming N(y).  The identities it relies on,
    checked on the reference's own set semantics (Oracle A):

      probe path   cn(x,y) = |{w in hop2(x) : w in N(y)}|        (iterate the small side)
      one-hub path hop2(x) = (N(h) \\ {x})  +  S'   disjoint, S' = U_{m in N(x), m != h} N(m) \\ N(h) \\ {x}
                   |hop2(x)| = deg(h) - 1 + |S'|
                   cn(x,y)   = |S' & N(y)| + |N(h) & N(y)| - [x in N(y)]
                   aa(x,y)   likewise with weights, minus w(x) when x in N(y)
    for every middle node h of x (the kernel uses the one that is a hub)."""
    import math
    rng = np.random.default_rng(seed)
    n_users, n_biz, m = 40, 12, 140
    eu = rng.integers(0, n_users, m)
    eb = rng.integers(0, n_biz, m) + 1000            # shared id space, disjoint columns
    G = oa.MiniSnapGraph.from_edges(zip(eu.tolist(), eb.tolist()))

    def w(i):
        d = G.degree(i)
        return 1.0 / math.log(d) if d > 1 else 0.0

    users = sorted(set(eu.tolist()))
    for x in users[:12]:
        hop2 = set(G.nodes_at_hop(x, 2))
        mids = sorted(G.nodes_at_hop(x, 1))
        for y in sorted(set(eb.tolist()))[:6]:
            ny = set(G.nodes_at_hop(y, 1))
            cn = oa.common_neighbors(hop2, ny)
            assert cn == sum(1 for i in hop2 if i in ny)                       # probe path
            for h in mids:
                nh = set(G.nodes_at_hop(h, 1))
                s_prime = set()
                for mnode in mids:
                    if mnode != h:
                        s_prime |= set(G.nodes_at_hop(mnode, 1))
                s_prime -= nh
                s_prime.discard(x)
                assert hop2 == (nh - {x}) | s_prime and not ((nh - {x}) & s_prime)
                assert len(hop2) == G.degree(h) - 1 + len(s_prime)
                x_in = 1 if x in ny else 0
                assert cn == len(s_prime & ny) + len(nh & ny) - x_in
                aa = oa.adamic_adar(hop2, ny, G)
                alt = (sum(w(i) for i in s_prime & ny) + sum(w(i) for i in nh & ny) - x_in * w(x))
                assert aa == pytest.approx(alt, rel=1e-12, abs=1e-12)
