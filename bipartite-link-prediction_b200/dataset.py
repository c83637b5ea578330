"""Candidate-pair generation, the step that produces the hot path's input.

Mirrors the traversal core of the reference's ``make_examples`` (dataset_maker.py:80-159): for
every sampled user the businesses at BFS distance exactly 3 are the candidates
(``snap.GetNodesAtHop(G, u, 3, ...)``, :137-139); a candidate is a positive when ``(u, b)`` is a
future edge (:141-142) and is otherwise kept with probability ``negative_sample_rate`` (:143-144).
The hop-3 sets come from the CUDA library (``blp_hop3_count`` / ``blp_hop3_fill``); the labelling
and the Bernoulli thinning are a few vectorised host operations on the resulting arrays.

Not reproduced: the Yelp-specific user filters of :97-119 (they need review.json) and Python 2's
Mersenne-Twister stream -- the negatives are drawn with ``numpy.random.default_rng(seed)``, so the
candidate SETS and the positives match the reference exactly, the sampled negatives only in
distribution.
"""
import numpy as np


def hop3_pairs(G, users):
    """(pair_u, pair_b) int32 arrays of every (user, hop-3 business) pair, grouped by user in the
    order given, businesses ascending.  `users` are LOCAL indices."""
    users = np.ascontiguousarray(users, dtype=np.int32)
    offsets, biz = G.hop3_candidates(users)
    offsets = offsets.cpu().numpy()
    pair_b = biz.cpu().numpy()
    pair_u = np.repeat(users, np.diff(offsets))
    return pair_u, pair_b, offsets


def make_examples(G, users, new_edges_u, new_edges_b, negative_sample_rate=0.01, seed=0):
    """``examples`` in the reference's structure {"<u>": {"<b>": 0|1}} (dataset_maker.py:134-159).

    users: ids of the shared id space (as ``random.sample(users, n_users)`` would return them);
    new_edges_*: the future edges (new_edges.txt) in the shared id space.
    """
    users = np.asarray(users, dtype=np.int64)
    lu = G.local_users(users)
    pair_u, pair_b, _ = hop3_pairs(G, lu)
    uid = G.user_ids if G.user_ids is not None else np.arange(G.n_users, dtype=np.int64)
    bid = G.biz_ids if G.biz_ids is not None else np.arange(G.n_biz, dtype=np.int64)
    ids_u, ids_b = uid[pair_u], bid[pair_b]
    # positives: (u, b) in new_edges
    stride = np.int64(max(int(bid.max()), int(np.max(new_edges_b, initial=0))) + 1)
    key = ids_u * stride + ids_b
    new_key = np.asarray(new_edges_u, dtype=np.int64) * stride + np.asarray(new_edges_b, dtype=np.int64)
    positive = np.isin(key, new_key)
    keep = positive | (np.random.default_rng(seed).random(key.size) < negative_sample_rate)
    examples = {}
    for u, b, y in zip(ids_u[keep].tolist(), ids_b[keep].tolist(), positive[keep].tolist()):
        examples.setdefault(str(u), {})[str(b)] = int(y)
    return examples
