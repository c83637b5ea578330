"""Parity of the CUDA path (through the C ABI) against the CPU oracles.  Needs a B200.

Bar: cn / union / pa bit-exact; jaccard bit-exact (fp64 div.rn of the same two integers);
adamic_adar within 1e-6 relative (north_star tolerance; the fixed-point accumulation is expected
to be within 1e-11).
"""
import json
import os

import numpy as np
import pytest

from conftest import pkg

pytestmark = pytest.mark.gpu

AA_RTOL = 1e-6          # tolerance stated by BASELINE.json's north_star
AA_RTOL_EXPECTED = 1e-8   # what the Q1.31 fixed-point accumulation actually delivers

INT_KEYS = ('u_cn', 'u_union', 'b_cn', 'b_union', 'pa')
JAC_KEYS = ('u_jaccard', 'b_jaccard')
AA_KEYS = ('u_adamic', 'b_adamic')


@pytest.fixture(scope='module')
def mods(built_lib):
    import torch
    assert torch.cuda.is_available(), 'gpu tests need a CUDA device'
    return pkg('graph'), pkg('synth')


def check_against(got, want, n):
    for k in INT_KEYS:
        g = np.asarray(got[k]).astype(np.int64)
        w = np.asarray(want[k]).astype(np.int64)
        bad = np.nonzero(g != w)[0]
        assert bad.size == 0, '%s differs at %d pairs, first %s: got %s want %s' % (
            k, bad.size, bad[:5], g[bad[:5]], w[bad[:5]])
    for k in JAC_KEYS:
        g = np.asarray(got[k], dtype=np.float64)
        w = np.asarray(want[k], dtype=np.float64)
        assert np.array_equal(g, w), '%s not bit-exact (max abs diff %g)' % (k, np.abs(g - w).max())
    for k in AA_KEYS:
        g = np.asarray(got[k], dtype=np.float64)
        w = np.asarray(want[k], dtype=np.float64)
        assert np.array_equal(g == 0, w == 0), '%s zero pattern differs' % k
        np.testing.assert_allclose(g, w, rtol=AA_RTOL, atol=0, err_msg=k)
        np.testing.assert_allclose(g, w, rtol=AA_RTOL_EXPECTED, atol=0, err_msg=k + ' (expected)')
    assert len(got['pa']) == n


def test_known_answer(mods):
    graph, synth = mods
    here = os.path.dirname(__file__)
    ka = json.load(open(os.path.join(here, 'golden', 'known_answer.json')))
    lines = np.array(ka['graph_lines'], dtype=np.int64)
    G = graph.BipartiteGraph.from_id_edges(lines[:, 0], lines[:, 1])
    info = G.info()
    assert info['n_edges_in'] == 11 and info['n_edges'] == 10     # duplicate line collapsed
    for nid, d in ka['degrees'].items():
        assert G.GetNI(int(nid)).GetDeg() == d
    pu = [p['u'] for p in ka['pairs']]
    pv = [p['v'] for p in ka['pairs']]
    got = G.score_id_pairs(pu, pv)
    want = {k: [p[k] for p in ka['pairs']] for k in INT_KEYS + JAC_KEYS + AA_KEYS}
    check_against(got, want, len(pu))


@pytest.mark.parametrize('seed,n_users,n_biz,n_rev,n_pairs,k', [
    (0, 300, 80, 1500, 3000, 10),
    (1, 2000, 300, 9000, 6000, 8),
    (2, 50, 500, 1200, 2500, 50),      # more businesses than users
    (3, 5000, 40, 4000, 4000, 4),      # few, very dense businesses (hub rows)
])
def test_random_graphs_vs_oracle_a(mods, seed, n_users, n_biz, n_rev, n_pairs, k):
    from oracle import similarity_oracle as oa
    graph, synth = mods
    eu, eb = synth.make_graph(n_users, n_biz, n_rev, seed=seed, shift_u=2.0, shift_b=2.0)
    pu, pv = synth.make_pairs(n_users, n_biz, eu, eb, n_pairs, k=k, seed=seed + 100,
                              invalid_frac=0.01)
    # add candidates that are already edges (u in N(v)) -- legal, similarity.py treats them alike
    pu = np.concatenate([pu, eu[:200]])
    pv = np.concatenate([pv, eb[:200]])
    G = graph.BipartiteGraph(n_users, n_biz, eu, eb)
    got = G.score_pairs_host(pu, pv)
    ids_eu, ids_eb = synth.shared_ids(n_users, eu, eb)
    ids_pu, ids_pv = synth.shared_ids(n_users, pu, pv)
    want = oa.score_pair_arrays(ids_eu, ids_eb, ids_pu, ids_pv)
    check_against(got, want, pu.size)


def test_c1_vs_oracle_b(mods):
    """BASELINE.json configs[0] (10k x 2k, 50k edges, 100k pairs) against the algebra oracle."""
    from oracle import algebra_oracle as ob
    graph, synth = mods
    cfg, eu, eb, pu, pv = synth.make_config('C1')
    G = graph.BipartiteGraph(cfg['n_users'], cfg['n_biz'], eu, eb)
    got = G.score_pairs_host(pu, pv, want_hop2=True)
    want = ob.score_pair_arrays(cfg['n_users'], cfg['n_biz'], eu, eb, pu, pv)
    check_against(got, want, pu.size)
    # hop-3 style invariants
    ok = want['in_graph'] == 1
    assert np.all(got['u_cn'][ok] <= np.minimum(got['u_hop2'][ok], got['u_union'][ok]))
    assert np.all(got['u_union'][ok] >= 1) and np.all(got['b_union'][ok] >= 1)
    assert np.all(got['u_adamic'] <= got['u_cn'] / np.log(2.0) + 1e-9)


def test_pair_order_does_not_matter(mods):
    graph, synth = mods
    cfg, eu, eb, pu, pv = synth.make_config('C1', n_pairs=20000)
    G = graph.BipartiteGraph(cfg['n_users'], cfg['n_biz'], eu, eb)
    a = G.score_pairs_host(pu, pv)
    perm = np.random.default_rng(5).permutation(pu.size)
    b = G.score_pairs_host(pu[perm], pv[perm])
    for k in a:
        assert np.array_equal(a[k][perm], b[k]), k      # bit-identical, adamic included


def test_empty_and_all_invalid(mods):
    graph, synth = mods
    G = graph.BipartiteGraph(4, 3, [0, 1, 2], [0, 0, 1])
    got = G.score_pairs_host(np.zeros(0, np.int32), np.zeros(0, np.int32))
    assert all(v.size == 0 for v in got.values())
    got = G.score_pairs_host(np.array([-1, 3, 0, 9], np.int32), np.array([0, 0, 2, -5], np.int32))
    for k, v in got.items():
        assert not v.any(), k      # user 3 and business 2 have degree 0 -> not in graph


def test_errors(mods):
    graph, synth = mods
    with pytest.raises(ValueError):
        graph.BipartiteGraph(4, 3, [0, 5], [0, 0])          # endpoint out of range
    with pytest.raises(ValueError):
        graph.BipartiteGraph.from_id_edges([1, 2], [2, 3])  # id 2 on both sides
