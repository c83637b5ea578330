"""Oracle A -- CPU restatement of the reference's similarity.py (TEST INFRASTRUCTURE ONLY).

PARITY PINNED TO THE REFERENCE'S OWN CODE: the reference ships no tests, fixtures or golden
outputs for this path, and its source is Python-2-only with a stripped SNAP
binding, so it cannot be imported as it lies.  oracle/build_ref.py compiles the reference's
similarity.py / util.py from /root/reference into oracle/_ref/ (three mechanical 2->3 rewrites,
bytecode only) and oracle/ref_runner.py executes it with a stand-in for SNAP's four calls.  This
restatement must agree with that code on the committed fixtures it wrote
(tests/golden/cases.json, ref_files.json), on the hand-checked table (known_answer.json) and on
hypothesis-generated graphs (tests/test_reference_pin.py); the sparse-algebra oracle
(algebra_oracle.py) and the C restatement (blp_oracle.c) must agree with it in turn.  What stays
unpinned is SNAP itself (binary absent): its documented behaviour is restated below.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.

What follows what (file:line relative to /root/reference):
  MiniSnapGraph.load_edge_list   <- snap.LoadEdgeList(snap.PUNGraph, f, 0, 1)   similarity.py:16
  MiniSnapGraph.node_ids         <- [N.GetId() for N in snap.Nodes(G)]           similarity.py:22,65
  MiniSnapGraph.nodes_at_hop     <- snap.GetNodesAtHop(G, id, hop, vec, True)    similarity.py:29,41,74,85
  MiniSnapGraph.degree           <- G.GetNI(i).GetDeg()                          similarity.py:121
  users / business / main        <- similarity.py:20-61 / 63-106 / 11-18
  jaccard / common_neighbors / adamic_adar  <- similarity.py:108-111 / 113-114 / 116-126
  preferential_attachment        <- "Link prediction.R":400-415 (deg(i)*deg(j))
  load_json / write_json         <- util.py:12-15 / 18-21

SNAP behaviour relied on (documented SNAP semantics, binary absent): TUNGraph is a simple
undirected graph (duplicate edge lines collapse), the node set is the ids seen on >=1 line,
GetNodesAtHop returns the nodes whose BFS distance from the start is EXACTLY `hop`, GetDeg is
the de-duplicated degree.
"""
import json
import math
from collections import defaultdict

SIMILARITY_METHODS = ('common_neighbors', 'jaccard', 'adamic_adar')


# --------------------------------------------------------------------------- SNAP stand-in
class MiniSnapGraph(object):
    """Just enough of snap.PUNGraph for similarity.py's four call sites."""

    def __init__(self):
        self._nbrs = {}

    def add_edge(self, a, b):
        self._nbrs.setdefault(a, set()).add(b)
        self._nbrs.setdefault(b, set()).add(a)

    @classmethod
    def load_edge_list(cls, path, src_col=0, dst_col=1):
        g = cls()
        with open(path) as fh:
            for line in fh:
                cols = line.split()
                if len(cols) <= max(src_col, dst_col):
                    continue
                g.add_edge(int(cols[src_col]), int(cols[dst_col]))
        return g

    @classmethod
    def from_edges(cls, edge_iter):
        g = cls()
        for a, b in edge_iter:
            g.add_edge(int(a), int(b))
        return g

    def node_ids(self):
        return list(self._nbrs.keys())

    def degree(self, nid):
        return len(self._nbrs[nid])

    def nodes_at_hop(self, start, hop):
        """Ids at BFS distance exactly `hop` from `start`."""
        seen = {start}
        frontier = [start]
        for _ in range(hop):
            nxt = []
            for x in frontier:
                for y in self._nbrs[x]:
                    if y not in seen:
                        seen.add(y)
                        nxt.append(y)
            frontier = nxt
        return frontier

    # the two SNAP spellings the set-level adamic_adar uses (similarity.py:121)
    def GetNI(self, nid):
        return _NodeIt(self, nid)


class _NodeIt(object):
    def __init__(self, g, nid):
        self._g, self._nid = g, nid

    def GetDeg(self):
        return self._g.degree(self._nid)


# --------------------------------------------------------------------------- set-level formulas
def jaccard(setone, settwo):
    # similarity.py:108-111 -- |a & b| / |a | b| in float
    inter = len(setone.intersection(settwo))
    union = len(setone.union(settwo))
    return float(inter) / float(union)


def common_neighbors(setone, settwo):
    # similarity.py:113-114
    return len(setone.intersection(settwo))


def adamic_adar(setone, settwo, G):
    # similarity.py:116-126 -- accumulator starts as int 0; degree-1 nodes add int 0
    total = 0
    for i in setone.intersection(settwo):
        deg = G.GetNI(i).GetDeg()
        if deg > 1:
            total += (math.log(deg)) ** -1
        else:
            total += 0
    return total


def preferential_attachment(setone, settwo):
    # "Link prediction.R":400-415: degree outer product; in set terms |N(u)| * |N(v)|
    return len(setone) * len(settwo)


def _apply(method, hop2, nbrs, G):
    if method == 'common_neighbors':
        return common_neighbors(hop2, nbrs)
    if method == 'jaccard':
        return jaccard(hop2, nbrs)
    if method == 'adamic_adar':
        return adamic_adar(hop2, nbrs, G)
    return None


# --------------------------------------------------------------------------- file-level loops
def _membership(G, faithful):
    # similarity.py:22,65 keeps the node ids in a *list* (O(|V|) per test).  faithful=True
    # reproduces that; faithful=False ("fair") uses a set -- identical arithmetic.
    ids = G.node_ids()
    return ids if faithful else set(ids)


def users(examples, G, methods, outfiles, faithful=False, write=True):
    """similarity.py:20-61.  Returns the list of per-method {u: {v: score}} dicts."""
    nodes = _membership(G, faithful)
    hop2s = {}
    for u in examples:                                   # loop A  :24-33
        nid = int(u)
        if nid in nodes:
            hop2s[nid] = set(G.nodes_at_hop(nid, 2))
    neighbors = {}
    for u in examples:                                   # loop B  :36-45
        for v in examples[u]:
            if int(v) not in neighbors and int(v) in nodes:
                neighbors[int(v)] = set(G.nodes_at_hop(int(v), 1))
    results = []
    for m, f in zip(methods, outfiles):                  # loop C  :48-61
        u_sim = defaultdict(dict)
        for u in examples:
            for v in examples[u]:
                if int(u) in nodes and int(v) in nodes:
                    s = _apply(m, hop2s[int(u)], neighbors[int(v)], G)
                    if s is not None:
                        u_sim[u][v] = s
                else:
                    u_sim[u][v] = 0
        if write:
            write_json(u_sim, f)
        results.append(u_sim)
    return results


def business(examples, G, methods, outfiles, faithful=False, write=True,
             reproduce_reference_bug=False):
    """similarity.py:63-106.

    The reference's third branch compares the method name with a log string (:102), so
    'adamic_adar' never matches and in-graph pairs are silently skipped.  The intended formula
    (:103) is the default here; reproduce_reference_bug=True skips them as the reference does.
    The per-pair print()s (:97,:100) are suppressed.
    """
    nodes = _membership(G, faithful)
    hop2s = {}
    for u in examples:                                   # loop A' :67-78
        for v in examples[u]:
            if int(v) not in hop2s:
                nid = int(v)
                if nid in nodes:
                    hop2s[nid] = set(G.nodes_at_hop(nid, 2))
    neighbors = {}
    for u in examples:                                   # loop B' :81-89
        if int(u) not in neighbors and int(u) in nodes:
            neighbors[int(u)] = set(G.nodes_at_hop(int(u), 1))
    results = []
    for m, f in zip(methods, outfiles):                  # loop C' :91-106
        b_sim = defaultdict(dict)
        for u in examples:
            for v in examples[u]:
                if int(u) in nodes and int(v) in nodes:
                    if m == 'adamic_adar' and reproduce_reference_bug:
                        continue
                    s = _apply(m, hop2s[int(v)], neighbors[int(u)], G)
                    if s is not None:
                        b_sim[u][v] = s
                else:
                    b_sim[u][v] = 0
        if write:
            write_json(b_sim, f)
        results.append(b_sim)
    return results


def main(example_file, graph_file, u_methods, u_outfiles, b_methods, b_outfiles,
         faithful=False, reproduce_reference_bug=False):
    # similarity.py:11-18
    examples = load_json(example_file)
    G = MiniSnapGraph.load_edge_list(graph_file, 0, 1)
    users(examples, G, u_methods, u_outfiles, faithful=faithful)
    business(examples, G, b_methods, b_outfiles, faithful=faithful,
             reproduce_reference_bug=reproduce_reference_bug)


def hop3_candidates(G, u):
    # dataset_maker.py:137-139 -- snap.GetNodesAtHop(G, u, 3, candidate_businesses, True):
    # the businesses at BFS distance exactly 3 of user u (test infrastructure for the device-side
    # candidate generator)
    return set(G.nodes_at_hop(u, 3))


def load_json(fname):
    with open(fname) as f:                               # util.py:12-15
        return json.loads(f.read())


def write_json(d, fname):
    with open(fname, 'w') as f:                          # util.py:18-21
        f.write(json.dumps(d))


# --------------------------------------------------------------------------- array-level view
def score_pair_arrays(edge_u, edge_b, pair_u, pair_b):
    """Same arithmetic as users()/business(), on arrays, for comparison with the CUDA path.

    edge_u/edge_b and pair_u/pair_b hold ids of ONE shared id space (users and businesses
    disjoint), duplicates allowed among the edges.  Returns a dict of python lists, one entry
    per pair: u_cn,u_union,u_jaccard,u_adamic,b_cn,b_union,b_jaccard,b_adamic,pa, in_graph.
    Pairs with an id that is not in the graph get 0 everywhere (similarity.py:59-60,104-105).
    Per-node sets are cached exactly as the reference's hop2s/neighbors dicts cache them.
    """
    G = MiniSnapGraph.from_edges(zip(edge_u, edge_b))
    nodes = set(G.node_ids())
    hop2, nbr = {}, {}

    def h2(x):
        if x not in hop2:
            hop2[x] = set(G.nodes_at_hop(x, 2))
        return hop2[x]

    def n1(x):
        if x not in nbr:
            nbr[x] = set(G.nodes_at_hop(x, 1))
        return nbr[x]

    keys = ('u_cn', 'u_union', 'u_jaccard', 'u_adamic',
            'b_cn', 'b_union', 'b_jaccard', 'b_adamic', 'pa', 'in_graph')
    out = {k: [] for k in keys}
    for u, v in zip(pair_u, pair_b):
        u, v = int(u), int(v)
        if u in nodes and v in nodes:
            hu, nv, hv, nu = h2(u), n1(v), h2(v), n1(u)
            out['u_cn'].append(common_neighbors(hu, nv))
            out['u_union'].append(len(hu.union(nv)))
            out['u_jaccard'].append(jaccard(hu, nv))
            out['u_adamic'].append(adamic_adar(hu, nv, G))
            out['b_cn'].append(common_neighbors(hv, nu))
            out['b_union'].append(len(hv.union(nu)))
            out['b_jaccard'].append(jaccard(hv, nu))
            out['b_adamic'].append(adamic_adar(hv, nu, G))
            out['pa'].append(preferential_attachment(nu, nv))
            out['in_graph'].append(1)
        else:
            for k in keys:
                out[k].append(0)
    return out
