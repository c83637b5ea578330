"""Graph construction on the device (blp_graph_create_device) against the host builder."""
import time

import numpy as np
import pytest

from conftest import pkg
from test_gpu_parity import check_against, mods  # noqa: F401

pytestmark = pytest.mark.gpu


def same_graph(a, b):
    ia, ib = a.info(), b.info()
    for k in ('n_users', 'n_biz', 'n_edges_in', 'n_edges', 'n_users_in_graph', 'n_biz_in_graph',
              'max_user_degree', 'max_biz_degree', 'n_hub_biz', 'n_hub_users'):
        assert ia[k] == ib[k], (k, ia[k], ib[k])
    lib = pkg('_lib')
    for side in (lib.SIDE_USER, lib.SIDE_BUSINESS):
        assert np.array_equal(a.degrees(side), b.degrees(side))


@pytest.mark.parametrize('name,n_pairs', [('C1', 40_000)])
def test_device_build_scores_match_host_build(mods, name, n_pairs):
    graph, synth = mods
    cfg, eu, eb, pu, pv = synth.make_config(name, n_pairs=n_pairs)
    Gh = graph.BipartiteGraph(cfg['n_users'], cfg['n_biz'], eu, eb, build='host')
    Gd = graph.BipartiteGraph(cfg['n_users'], cfg['n_biz'], eu, eb, build='device')
    same_graph(Gh, Gd)
    a = Gh.score_pairs_host(pu, pv, want_hop2=True)
    b = Gd.score_pairs_host(pu, pv, want_hop2=True)
    for k in a:
        assert np.array_equal(a[k], b[k]), k            # bit-identical, adamic included


def test_device_build_small_and_odd_shapes(mods):
    from oracle import c_oracle
    graph, synth = mods
    rng = np.random.default_rng(4)
    for n_users, n_biz, n_edges in ((1, 1, 1), (5, 3, 0), (7, 300, 50), (3000, 2, 5000),
                                    (70000, 9, 40000)):
        eu = rng.integers(0, n_users, n_edges)
        eb = rng.integers(0, n_biz, n_edges)                  # heavy duplication when tiny
        pu = rng.integers(-1, n_users, 500)
        pv = rng.integers(-1, n_biz, 500)
        G = graph.BipartiteGraph(n_users, n_biz, eu, eb, build='device')
        got = G.score_pairs_host(pu, pv)
        want = c_oracle.score_pair_arrays(n_users, n_biz, eu, eb, pu, pv)
        check_against(got, want, pu.size)
    for build in ('host', 'device'):
        with pytest.raises(ValueError):
            graph.BipartiteGraph(4, 3, [0, 5], [0, 0], build=build)   # endpoint out of range
    # the host builder on the same odd shapes
    for n_users, n_biz, n_edges in ((1, 1, 1), (5, 3, 0), (3000, 2, 5000)):
        eu = rng.integers(0, n_users, n_edges)
        eb = rng.integers(0, n_biz, n_edges)
        pu = rng.integers(-1, n_users, 300)
        pv = rng.integers(-1, n_biz, 300)
        G = graph.BipartiteGraph(n_users, n_biz, eu, eb, build='host')
        check_against(G.score_pairs_host(pu, pv),
                      c_oracle.score_pair_arrays(n_users, n_biz, eu, eb, pu, pv), pu.size)


def test_device_build_c2_and_timing(mods):
    """BASELINE.json configs[1] graph built on the GPU: same degrees, same scores on a sample."""
    import torch
    graph, synth = mods
    cfg, eu, eb, pu, pv = synth.make_config('C2', n_pairs=200_000)
    t0 = time.perf_counter()
    Gh = graph.BipartiteGraph(cfg['n_users'], cfg['n_biz'], eu, eb, build='host')
    t_host = time.perf_counter() - t0
    deu, deb = torch.from_numpy(eu).cuda(), torch.from_numpy(eb).cuda()
    graph.BipartiteGraph(cfg['n_users'], cfg['n_biz'], deu, deb, build='device').close()   # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    Gd = graph.BipartiteGraph(cfg['n_users'], cfg['n_biz'], deu, deb, build='device')
    t_dev = time.perf_counter() - t0
    print('graph build: host %.3f s, device %.3f s' % (t_host, t_dev))
    same_graph(Gh, Gd)
    a = Gh.score_pairs_host(pu, pv)
    b = Gd.score_pairs_host(pu, pv)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
