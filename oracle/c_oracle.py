"""ctypes loader for oracle/blp_oracle.c (TEST INFRASTRUCTURE ONLY; pinned to the reference through Oracle A, see similarity_oracle.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libblp_oracle.so')


def load():
    src = os.path.join(_HERE, 'blp_oracle.c')
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(['make', '-s', '-C', _HERE], check=True)
    lib = ctypes.CDLL(_SO)
    lib.blp_oracle_score.restype = ctypes.c_int
    lib.blp_oracle_score.argtypes = ([ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 11)
    return lib


def score_pair_arrays(n_users, n_biz, edge_u, edge_b, pair_u, pair_b):
    lib = load()
    eu = np.ascontiguousarray(edge_u, dtype=np.int32)
    eb = np.ascontiguousarray(edge_b, dtype=np.int32)
    pu = np.ascontiguousarray(pair_u, dtype=np.int32)
    pv = np.ascontiguousarray(pair_b, dtype=np.int32)
    n = pu.size
    out = {k: np.zeros(n, dtype=np.int32) for k in ('u_cn', 'u_union', 'b_cn', 'b_union')}
    for k in ('u_jaccard', 'u_adamic', 'b_jaccard', 'b_adamic'):
        out[k] = np.zeros(n, dtype=np.float64)
    out['pa'] = np.zeros(n, dtype=np.int64)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.blp_oracle_score(int(n_users), int(n_biz), int(eu.size), p(eu), p(eb), int(n), p(pu),
                              p(pv), p(out['u_cn']), p(out['u_union']), p(out['u_jaccard']),
                              p(out['u_adamic']), p(out['b_cn']), p(out['b_union']),
                              p(out['b_jaccard']), p(out['b_adamic']), p(out['pa']))
    if rc != 0:
        raise RuntimeError('blp_oracle_score failed: %d' % rc)
    return out


def score_pair_arrays_parallel(n_users, n_biz, edge_u, edge_b, pair_u, pair_b, threads=None):
    """The same over contiguous slices of the pair list on `threads` host threads (the C call
    releases the GIL; every slice rebuilds the graph, as a separate process would)."""
    from concurrent.futures import ThreadPoolExecutor
    load()
    threads = threads or len(os.sched_getaffinity(0))
    n = len(pair_u)
    threads = max(1, min(threads, n // 1000 or 1))
    bounds = [(n * t) // threads for t in range(threads + 1)]
    pu = np.ascontiguousarray(pair_u, dtype=np.int32)
    pv = np.ascontiguousarray(pair_b, dtype=np.int32)
    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(lambda lh: score_pair_arrays(n_users, n_biz, edge_u, edge_b,
                                                         pu[lh[0]:lh[1]], pv[lh[0]:lh[1]]),
                            zip(bounds[:-1], bounds[1:])))
    return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
