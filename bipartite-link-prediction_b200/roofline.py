"""Algorithmic bytes of the scoring path -- the roofline numerator.

The figure mirrors the reference's own structure -- one hop-2 set per distinct node
(similarity.py:24-33), one partner list per pair (:50-58) -- and deliberately ignores row-offset
reads and every probe of the membership structure, so an implementation cannot inflate it:

  B_side = sum over distinct grouping nodes x of in-graph pairs   4 * (deg(x) + sum_{m in N(x)} deg(m))
         + sum over in-graph pairs (x, y)                          8 + 4*deg(y) + 8*cn(x,y) + 24
         + sum over out-of-graph pairs                             8 + 24
  B_pa   = 8 * n_pairs

It needs only the edge list, the pair list and the cn column the kernels produced.
"""
import numpy as np


def dedup_graph(n_users, n_biz, edge_u, edge_b):
    key = np.unique(np.asarray(edge_u, dtype=np.int64) * n_biz + np.asarray(edge_b, dtype=np.int64))
    uu, bb = key // n_biz, key % n_biz
    du = np.bincount(uu, minlength=n_users)
    db = np.bincount(bb, minlength=n_biz)
    return uu, bb, du, db


def algorithmic_bytes(n_users, n_biz, edge_u, edge_b, pair_u, pair_b, u_cn, b_cn):
    """Returns dict(user=..., business=..., pa=..., expansion_user=..., stream_user=..., ...)."""
    uu, bb, du, db = dedup_graph(n_users, n_biz, edge_u, edge_b)
    pu = np.asarray(pair_u, dtype=np.int64)
    pv = np.asarray(pair_b, dtype=np.int64)
    ok = (pu >= 0) & (pu < n_users) & (pv >= 0) & (pv < n_biz)
    ok[ok] &= (du[pu[ok]] > 0) & (db[pv[ok]] > 0)
    n, n_ok = pu.size, int(ok.sum())
    exp_u = np.bincount(uu, weights=db[bb].astype(np.float64), minlength=n_users)
    exp_b = np.bincount(bb, weights=du[uu].astype(np.float64), minlength=n_biz)
    ux = np.unique(pu[ok])
    bx = np.unique(pv[ok])
    out = {}
    out['expansion_user'] = float(4.0 * (du[ux].sum() + exp_u[ux].sum()))
    out['expansion_business'] = float(4.0 * (db[bx].sum() + exp_b[bx].sum()))
    out['stream_user'] = float(32.0 * n_ok + 4.0 * db[pv[ok]].sum()
                               + 8.0 * np.asarray(u_cn, dtype=np.int64)[ok].sum())
    out['stream_business'] = float(32.0 * n_ok + 4.0 * du[pu[ok]].sum()
                                   + 8.0 * np.asarray(b_cn, dtype=np.int64)[ok].sum())
    out['invalid'] = float(32.0 * (n - n_ok))
    out['user'] = out['expansion_user'] + out['stream_user'] + out['invalid']
    out['business'] = out['expansion_business'] + out['stream_business'] + out['invalid']
    out['pa'] = 8.0 * n
    out['total'] = out['user'] + out['business'] + out['pa']
    out['n_pairs'] = n
    out['n_in_graph'] = n_ok
    out['groups_user'] = int(ux.size)
    out['groups_business'] = int(bx.size)
    return out
