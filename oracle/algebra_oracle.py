"""Oracle B -- independent sparse-algebra restatement (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED (see similarity_oracle.py): no reference golden vectors exist.  This oracle
shares no code with Oracle A: it never runs a BFS and never builds a Python set.  It evaluates

    u_cn = ((A.A^T > 0) - I) . A            at the pairs   (similarity.py:20-61, 113-114)
    b_cn = A . ((A^T.A > 0) - I)            at the pairs   (similarity.py:63-106)
    |hop2(u)| = nnz(row u of (A.A^T > 0) - I),   union = |hop2| + deg(partner) - cn   (:108-111)
    aa   = the same products with diag(1/ln(deg)) (0 where deg<=1) in the middle    (:116-126)
    pa   = deg(u) * deg(v)                                     ("Link prediction.R":400-415)

with A the de-duplicated users x businesses biadjacency.  Only tests/ and bench.py's CPU
baseline legs may import it.  It is also the "line-for-line NumPy re-execution" the north_star
names as the timed CPU baseline when the Python-2 `_snap.so` cannot load (it cannot).

Indices are LOCAL: users 0..n_users-1, businesses 0..n_biz-1; a pair entry of -1 (or a node of
degree 0) means "id not in graph" and yields zeros everywhere.
"""
import numpy as np
import scipy.sparse as sp


def biadjacency(n_users, n_biz, edge_u, edge_b):
    eu = np.asarray(edge_u, dtype=np.int64)
    eb = np.asarray(edge_b, dtype=np.int64)
    A = sp.csr_matrix((np.ones(eu.size, dtype=np.int64), (eu, eb)), shape=(n_users, n_biz))
    A.sum_duplicates()
    A.data[:] = 1                      # duplicate review lines collapse (SNAP TUNGraph)
    return A


def _inv_log_deg(deg):
    w = np.zeros(deg.shape, dtype=np.float64)
    m = deg > 1
    w[m] = 1.0 / np.log(deg[m].astype(np.float64))
    return w


def _side(A, rows, cols, block):
    """Scores of (rows[i], cols[i]) with hop-2 sets taken on the ROW side of A.

    Returns cn, hop2_size, aa.  Row-blocked: (A.A^T) is far too dense to hold whole.
    """
    n_r = A.shape[0]
    At = A.T.tocsr()
    deg_r = np.asarray(A.sum(axis=1)).ravel()
    w_r = sp.diags(_inv_log_deg(deg_r))
    cn = np.zeros(rows.size, dtype=np.int64)
    h2 = np.zeros(rows.size, dtype=np.int64)
    aa = np.zeros(rows.size, dtype=np.float64)
    uniq, inv = np.unique(rows, return_inverse=True)
    for s in range(0, uniq.size, block):
        blk = uniq[s:s + block]
        H = (A[blk] @ At).tocsr()                     # paths of length 2, row side
        H.data[:] = 1
        # remove the self entry (x is never at distance 2 from itself)
        self_mask = sp.csr_matrix((np.ones(blk.size, dtype=np.int64),
                                   (np.arange(blk.size), blk)), shape=H.shape)
        H = H - H.multiply(self_mask)
        H.eliminate_zeros()
        C = (H @ A).tocsr()                           # distinct-node counts
        W = (H @ w_r @ A).tocsr()
        sel = np.nonzero((inv >= s) & (inv < s + blk.size))[0]
        loc = inv[sel] - s
        cn[sel] = np.asarray(C[loc, cols[sel]]).ravel()
        aa[sel] = np.asarray(W[loc, cols[sel]]).ravel()
        h2[sel] = np.diff(H.indptr)[loc]
    return cn, h2, aa


def score_pair_arrays(n_users, n_biz, edge_u, edge_b, pair_u, pair_b, block=256):
    A = biadjacency(n_users, n_biz, edge_u, edge_b)
    pu = np.asarray(pair_u, dtype=np.int64)
    pv = np.asarray(pair_b, dtype=np.int64)
    deg_u = np.asarray(A.sum(axis=1)).ravel()
    deg_b = np.asarray(A.sum(axis=0)).ravel()
    ok = (pu >= 0) & (pv >= 0) & (pu < n_users) & (pv < n_biz)
    ok[ok] &= (deg_u[pu[ok]] > 0) & (deg_b[pv[ok]] > 0)
    idx = np.nonzero(ok)[0]
    n = pu.size
    out = {k: np.zeros(n, dtype=np.int64) for k in ('u_cn', 'u_union', 'b_cn', 'b_union', 'pa')}
    for k in ('u_jaccard', 'u_adamic', 'b_jaccard', 'b_adamic'):
        out[k] = np.zeros(n, dtype=np.float64)
    out['in_graph'] = ok.astype(np.int64)
    if idx.size == 0:
        return out
    u, v = pu[idx], pv[idx]
    cn, h2, aa = _side(A, u, v, block)
    uni = h2 + deg_b[v] - cn
    out['u_cn'][idx], out['u_union'][idx], out['u_adamic'][idx] = cn, uni, aa
    out['u_jaccard'][idx] = cn.astype(np.float64) / uni.astype(np.float64)
    At = A.T.tocsr()
    cn, h2, aa = _side(At, v, u, block)
    uni = h2 + deg_u[u] - cn
    out['b_cn'][idx], out['b_union'][idx], out['b_adamic'][idx] = cn, uni, aa
    out['b_jaccard'][idx] = cn.astype(np.float64) / uni.astype(np.float64)
    out['pa'][idx] = deg_u[u] * deg_b[v]
    return out
