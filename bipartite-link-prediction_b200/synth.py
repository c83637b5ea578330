"""Seeded synthetic Yelp-shaped inputs (the Yelp dump is not available offline).

Graphs: Chung-Lu style -- every review draws its user and its business independently from
rank-based power-law weights ``w(i) ~ (i + shift)^(-1/(alpha-1))``; duplicate draws are KEPT (the
real ``graph.txt`` has repeat reviews, dataset_maker.py:197, and the path must de-duplicate).
Node ids are shuffled so that degree is not correlated with id.

Pairs: ``K`` distinct candidate businesses for each sampled example user -- half drawn in
proportion to degree (the hop-3 candidate bias of dataset_maker.py:138-144), half uniform
(dataset_maker.py:62-64) -- and 0.1 % of the pairs get an id that is not in the graph, which
exercises the literal-0 branch (similarity.py:59-60).  Pairs come grouped by user, as
``examples.json`` stores them.

Everything is local-index based: users ``0..n_users-1``, businesses ``0..n_biz-1``, ``-1`` = id
not in the graph.  ``shared_ids`` maps to the reference's single id space (businesses offset by
``n_users``, as dataset_maker.py:173-174 keeps the two disjoint).
"""
import numpy as np

# name -> shape of the five BASELINE.json configs
CONFIGS = {
    'C1': dict(n_users=10_000, n_biz=2_000, n_reviews=50_000, n_pairs=100_000, k=10,
               alpha_u=2.3, alpha_b=2.1, shift_u=3.0, shift_b=6.0),
    'C2': dict(n_users=366_000, n_biz=61_000, n_reviews=1_500_000, n_pairs=10_000_000, k=32,
               alpha_u=2.3, alpha_b=2.1, shift_u=5.0, shift_b=20.0),
    'C3': dict(n_users=1_600_000, n_biz=190_000, n_reviews=6_700_000, n_pairs=100_000_000, k=64,
               alpha_u=2.3, alpha_b=2.1, shift_u=5.0, shift_b=20.0),
    'C4': dict(n_users=1_600_000, n_biz=190_000, n_reviews=6_700_000, n_pairs=50_000_000, k=64,
               alpha_u=2.3, alpha_b=1.8, shift_u=5.0, shift_b=8.0),
    'C5': dict(n_users=10_000_000, n_biz=1_000_000, n_reviews=100_000_000,
               n_pairs=1_000_000_000, k=100, alpha_u=2.3, alpha_b=2.1, shift_u=5.0, shift_b=20.0),
}


def powerlaw_weights(n, alpha, shift):
    w = (np.arange(n, dtype=np.float64) + shift) ** (-1.0 / (alpha - 1.0))
    return w / w.sum()


def make_graph(n_users, n_biz, n_reviews, alpha_u=2.3, alpha_b=2.1, shift_u=5.0, shift_b=20.0,
               seed=0, **_):
    """Returns (edge_u, edge_b) int32 arrays of length n_reviews, duplicates included."""
    rng = np.random.default_rng(seed)
    cu = np.cumsum(powerlaw_weights(n_users, alpha_u, shift_u))
    cb = np.cumsum(powerlaw_weights(n_biz, alpha_b, shift_b))
    ru = np.searchsorted(cu, rng.random(n_reviews) * cu[-1]).clip(0, n_users - 1)
    rb = np.searchsorted(cb, rng.random(n_reviews) * cb[-1]).clip(0, n_biz - 1)
    perm_u = rng.permutation(n_users).astype(np.int32)
    perm_b = rng.permutation(n_biz).astype(np.int32)
    return perm_u[ru], perm_b[rb]


def degrees(n_users, n_biz, edge_u, edge_b):
    """De-duplicated degrees (numpy; used to bias the candidate sampler, not for scoring)."""
    key = np.unique(edge_u.astype(np.int64) * n_biz + edge_b.astype(np.int64))
    du = np.bincount(key // n_biz, minlength=n_users)
    db = np.bincount(key % n_biz, minlength=n_biz)
    return du, db


def _distinct_rows(cand, n_biz, rng, max_rounds=64):
    """Replace within-row duplicates by uniform re-draws until every row is duplicate free."""
    for _ in range(max_rounds):
        order = np.argsort(cand, axis=1, kind='stable')
        srt = np.take_along_axis(cand, order, axis=1)
        dup_sorted = np.zeros(cand.shape, dtype=bool)
        dup_sorted[:, 1:] = srt[:, 1:] == srt[:, :-1]
        if not dup_sorted.any():
            return cand
        dup = np.zeros(cand.shape, dtype=bool)
        np.put_along_axis(dup, order, dup_sorted, axis=1)
        cand[dup] = rng.integers(0, n_biz, size=int(dup.sum()), dtype=cand.dtype)
    raise RuntimeError('could not make candidate rows distinct (K too close to n_biz?)')


def make_pairs(n_users, n_biz, edge_u, edge_b, n_pairs, k=32, seed=1, invalid_frac=0.001,
               rank=0, world=1, deg=None, **_):
    """Candidate pairs (pair_u, pair_b) as int32 local indices, grouped by user.

    With world > 1 the example users are cut into `world` contiguous slices and only slice
    `rank` is generated, so every rank can build its own shard without the others'.
    """
    du, db = deg if deg is not None else degrees(n_users, n_biz, edge_u, edge_b)
    rng = np.random.default_rng(seed)
    avail = np.nonzero(du > 0)[0]
    n_ex = -(-n_pairs // k)
    if n_ex > avail.size:                      # small graphs: use every user, widen K
        n_ex = avail.size
        k = -(-n_pairs // n_ex)
    if k > n_biz:
        raise ValueError('K=%d candidates per user exceeds n_biz=%d' % (k, n_biz))
    ex_users = np.sort(rng.choice(avail, size=n_ex, replace=False))
    lo, hi = (n_ex * rank) // world, (n_ex * (rank + 1)) // world
    ex_users = ex_users[lo:hi]
    rng = np.random.default_rng([seed, rank, world])
    n_loc = ex_users.size
    kb = k // 2
    cdf = np.cumsum(db.astype(np.float64))
    biased = np.searchsorted(cdf, rng.random((n_loc, kb)) * cdf[-1]).clip(0, n_biz - 1)
    uniform = rng.integers(0, n_biz, size=(n_loc, k - kb))
    cand = np.concatenate([biased, uniform], axis=1).astype(np.int32)
    cand = _distinct_rows(cand, n_biz, rng)
    pair_u = np.repeat(ex_users.astype(np.int32), k)
    pair_b = cand.reshape(-1)
    want = (n_pairs * (rank + 1)) // world - (n_pairs * rank) // world
    pair_u, pair_b = pair_u[:want].copy(), pair_b[:want].copy()
    n_bad = int(round(pair_u.size * invalid_frac))
    if n_bad:
        bad = rng.choice(pair_u.size, size=n_bad, replace=False)
        half = n_bad // 2
        pair_b[bad[:half]] = -1                # unknown business id
        pair_u[bad[half:]] = -1                # unknown user id
    return pair_u, pair_b


def make_config(name, seed_graph=0, seed_pairs=1, n_pairs=None, rank=0, world=1):
    """(cfg, edge_u, edge_b, pair_u, pair_b) for one of the BASELINE.json configs."""
    cfg = dict(CONFIGS[name])
    if n_pairs is not None:
        cfg['n_pairs'] = int(n_pairs)
    eu, eb = make_graph(seed=seed_graph, **cfg)
    pu, pv = make_pairs(edge_u=eu, edge_b=eb, seed=seed_pairs, rank=rank, world=world, **cfg)
    return cfg, eu, eb, pu, pv


def shared_ids(n_users, local_u, local_b, missing_base=None):
    """Local indices -> the reference's shared id space.  -1 maps to ids nobody uses."""
    lu = np.asarray(local_u, dtype=np.int64)
    lb = np.asarray(local_b, dtype=np.int64)
    if missing_base is None:
        missing_base = 1 << 40
    ids_u = np.where(lu >= 0, lu, missing_base + np.arange(lu.size))
    ids_b = np.where(lb >= 0, lb + n_users, missing_base + (1 << 36) + np.arange(lb.size))
    return ids_u, ids_b


def examples_dict(ids_u, ids_b):
    """``examples.json`` structure {"<u>": {"<b>": 0}} (labels are irrelevant to scoring)."""
    ex = {}
    for u, b in zip(ids_u.tolist(), ids_b.tolist()):
        ex.setdefault(str(u), {})[str(b)] = 0
    return ex
