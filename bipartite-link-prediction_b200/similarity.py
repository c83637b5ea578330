"""Drop-in for the reference's ``similarity.py``: same six call signatures, same file formats.

    main(example_file, graph_file, u_methods, u_outfiles, b_methods, b_outfiles)  similarity.py:11
    users(examples, G, methods, outfiles)                                          similarity.py:20
    business(examples, G, methods, outfiles)                                       similarity.py:63
    jaccard(setone, settwo)                                                        similarity.py:108
    common_neighbors(setone, settwo)                                               similarity.py:113
    adamic_adar(setone, settwo, G)                                                 similarity.py:116
    preferential_attachment(setone, settwo)      (no Python original; "Link prediction.R":400-415)

``examples`` is the reference's ``{"<user>": {"<business>": 0|1}}`` dict (dataset_maker.py:142-159),
``G`` is a ``BipartiteGraph`` (what ``LoadEdgeList`` below returns), score files are
``{"<user>": {"<business>": score}}`` with the key set of ``examples`` (similarity.py:49-61), read
unchanged by eval.py:14-18, supervised_models.py:81-86 and supervised_classifier.py:23-25.

``users`` / ``business`` flatten the examples into pair arrays and make ONE call per side into
the CUDA library (``blp_score_pairs``), which yields all three methods at once; the per-method
loop only selects which array is written where.  The three set-level helpers keep their
reference meaning on Python sets -- they are the public formulas, not the bulk path.
"""
import math
from collections import defaultdict

import numpy as np

from . import _lib, util
from .graph import BipartiteGraph

METHODS = ('common_neighbors', 'jaccard', 'adamic_adar', 'preferential_attachment')
_COLUMN = {'common_neighbors': 'cn', 'jaccard': 'jaccard', 'adamic_adar': 'adamic',
           'preferential_attachment': 'pa'}

PUNGraph = 'PUNGraph'   # placeholder for snap.PUNGraph in LoadEdgeList(PUNGraph, file, 0, 1)


def LoadEdgeList(graph_type, graph_file, src_col=0, dst_col=1, device=None):
    """``snap.LoadEdgeList(snap.PUNGraph, graph_file, 0, 1)`` (similarity.py:16)."""
    return BipartiteGraph.from_edge_list(graph_file, src_col, dst_col, device=device)


def main(example_file, graph_file, u_methods, u_outfiles, b_methods, b_outfiles, *,
         device=None, reproduce_reference_bug=False, sidecar=False):
    examples = util.load_json(example_file)
    G = LoadEdgeList(PUNGraph, graph_file, 0, 1, device=device)
    users(examples, G, u_methods, u_outfiles, sidecar=sidecar)
    business(examples, G, b_methods, b_outfiles, reproduce_reference_bug=reproduce_reference_bug,
             sidecar=sidecar)
    return G


def _flatten(examples):
    us, bs = [], []
    for u in examples:
        for v in examples[u]:
            us.append(u)
            bs.append(v)
    ids_u = np.fromiter((int(u) for u in us), dtype=np.int64, count=len(us))
    ids_b = np.fromiter((int(b) for b in bs), dtype=np.int64, count=len(bs))
    return us, bs, ids_u, ids_b


def _score_side(G, side, ids_u, ids_b):
    import torch
    lu, lb = G.local_users(ids_u), G.local_businesses(ids_b)
    du = G.degrees(_lib.SIDE_USER)
    db = G.degrees(_lib.SIDE_BUSINESS)
    in_graph = (lu >= 0) & (lb >= 0)
    in_graph[in_graph] &= (du[lu[in_graph]] > 0) & (db[lb[in_graph]] > 0)
    with torch.cuda.device(G.device):
        tu = torch.from_numpy(lu).to(G.device)
        tb = torch.from_numpy(lb).to(G.device)
        res = G.score_side(side, tu, tb, want_pa=True)
        host = {k: v.cpu().numpy() for k, v in res.items()}
    return host, in_graph


def _emit(examples, keys_u, keys_b, host, in_graph, methods, outfiles, skip_in_graph=(),
          sidecar=False):
    out = []
    for m, f in zip(methods, outfiles):
        sim = defaultdict(dict)
        col = _COLUMN.get(m)
        if col is None:
            # unknown method name: the reference writes only the literal zeros (similarity.py:53-60)
            for i, (u, v) in enumerate(zip(keys_u, keys_b)):
                if not in_graph[i]:
                    sim[u][v] = 0
        else:
            vals = host[col].tolist()
            is_float = col in ('jaccard', 'adamic')
            for i, (u, v) in enumerate(zip(keys_u, keys_b)):
                if not in_graph[i]:
                    sim[u][v] = 0                       # similarity.py:59-60, 104-105
                elif m in skip_in_graph:
                    continue
                elif is_float and not (col == 'adamic' and vals[i] == 0.0):
                    sim[u][v] = float(vals[i])
                else:
                    sim[u][v] = int(vals[i])            # cn, pa, and adamic's untouched int 0
        util.write_json(sim, f)
        if sidecar:   # columnar copy beside the JSON; same content, no parsing
            u, b, v, im = util.dict_to_columns(sim)
            np.savez(f[:-5] + '.npz' if f.endswith('.json') else f + '.npz', u=u, b=b, v=v,
                     int_mask=im)
        out.append(sim)
    return out


def users(examples, G, methods, outfiles, *, sidecar=False):
    keys_u, keys_b, ids_u, ids_b = _flatten(examples)
    host, in_graph = _score_side(G, _lib.SIDE_USER, ids_u, ids_b)
    return _emit(examples, keys_u, keys_b, host, in_graph, methods, outfiles, sidecar=sidecar)


def business(examples, G, methods, outfiles, *, reproduce_reference_bug=False, sidecar=False):
    """``reproduce_reference_bug=True`` omits in-graph ``adamic_adar`` entries exactly as the
    reference's mistyped branch does (similarity.py:102); the default writes the intended score
    (similarity.py:103)."""
    keys_u, keys_b, ids_u, ids_b = _flatten(examples)
    host, in_graph = _score_side(G, _lib.SIDE_BUSINESS, ids_u, ids_b)
    skip = ('adamic_adar',) if reproduce_reference_bug else ()
    return _emit(examples, keys_u, keys_b, host, in_graph, methods, outfiles, skip_in_graph=skip,
                 sidecar=sidecar)


# ----------------------------------------------------------------------------- set-level API
def jaccard(setone, settwo):
    inter = len(setone & settwo)
    return float(inter) / float(len(setone) + len(settwo) - inter)


def common_neighbors(setone, settwo):
    return len(setone & settwo)


def adamic_adar(setone, settwo, G):
    total = 0
    for i in setone & settwo:
        deg = G.GetNI(i).GetDeg()
        if deg > 1:
            total += (math.log(deg)) ** -1      # the reference's exact expression (similarity.py:123)
    return total


def preferential_attachment(setone, settwo):
    return len(setone) * len(settwo)


if __name__ == '__main__':   # same literal paths as similarity.py:128-142
    for split in ('train', 'test'):
        d = './data/%s/' % split
        main(d + 'examples.json', d + 'graph.txt',
             ['common_neighbors', 'jaccard', 'adamic_adar'],
             [d + 'u_cn.json', d + 'u_jaccard.json', d + 'u_adamic.json'],
             ['common_neighbors', 'jaccard', 'adamic_adar'],
             [d + 'b_cn.json', d + 'b_jaccard.json', d + 'b_adamic.json'])
