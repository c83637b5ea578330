"""Runs the reference's OWN similarity.py (bytecode built by oracle/build_ref.py into oracle/_ref/).

TEST INFRASTRUCTURE ONLY.  This is what pins Oracle A (oracle/similarity_oracle.py, the
restatement) to the reference itself: every function executed here -- main, users, business,
jaccard, common_neighbors, adamic_adar (similarity.py:11-126) and util.load_json / write_json
(util.py:12-21) -- is the reference's code object, unmodified apart from the three mechanical
Python-2 -> 3 rewrites documented in build_ref.py.

What is NOT the reference: the `snap` module.  Its binary is stripped from the mount
(.MISSING_LARGE_BLOBS), so the four SNAP call sites of the path are served by `SnapShim`, written
from SNAP's documented behaviour:
    snap.LoadEdgeList(snap.PUNGraph, file, 0, 1)   similarity.py:16   undirected simple graph
    snap.Nodes(G) -> node iterators with GetId()   similarity.py:22,65
    snap.TIntV() / snap.GetNodesAtHop(G, id, hop, vec, True)   :28-29,40-41,73-74,84-85
                                                   nodes at BFS distance EXACTLY `hop`
    G.GetNI(i).GetDeg()                            similarity.py:121  de-duplicated degree
The graph container behind the shim is oracle.similarity_oracle.MiniSnapGraph.
"""
import contextlib
import io
import json
import marshal
import os
import sys
import tempfile
import types

from . import build_ref
from .similarity_oracle import MiniSnapGraph

REF_DIR = build_ref.OUT_DIR
METHODS = ['common_neighbors', 'jaccard', 'adamic_adar']


# --------------------------------------------------------------------------- the snap stand-in
class _TIntV(list):
    """snap.TIntV: an int vector; the reference only iterates it (similarity.py:31,43,76,87)."""

    def Len(self):
        return len(self)


class _Node(object):
    def __init__(self, nid):
        self._nid = nid

    def GetId(self):
        return self._nid


def _make_snap_module():
    snap = types.ModuleType('snap')
    snap.PUNGraph = 'PUNGraph'
    snap.TIntV = _TIntV

    def LoadEdgeList(graph_type, path, src_col=0, dst_col=1):
        if graph_type != snap.PUNGraph:
            raise ValueError('only PUNGraph is used on this path')
        return MiniSnapGraph.load_edge_list(path, src_col, dst_col)

    def Nodes(G):
        return (_Node(i) for i in G.node_ids())

    def GetNodesAtHop(G, start, hop, vec, is_dir):
        del vec[:]
        vec.extend(G.nodes_at_hop(start, hop))
        return len(vec)

    snap.LoadEdgeList, snap.Nodes, snap.GetNodesAtHop = LoadEdgeList, Nodes, GetNodesAtHop
    return snap


# --------------------------------------------------------------------------- loading the bytecode
_cache = {}


def available():
    return all(os.path.exists(os.path.join(REF_DIR, n)) for n in ('similarity.marshal', 'util.marshal'))


def ensure_built():
    """(Re)build oracle/_ref when the reference checkout is present; True when usable."""
    if os.path.isdir(build_ref.REF_ROOT):
        try:
            build_ref.build(quiet=True)
        except Exception:      # a stale but present build is still usable
            pass
    return available()


def _load_code(name):
    with open(os.path.join(REF_DIR, name + '.marshal'), 'rb') as fh:
        return marshal.load(fh)


def load():
    """The reference's `similarity` module (and its `util`), executed with the snap stand-in."""
    if 'similarity' in _cache:
        return _cache['similarity']
    if not available():
        raise RuntimeError('oracle/_ref is not built: run `python oracle/build_ref.py` where '
                           '/root/reference is present')
    util = types.ModuleType('reference_util')
    exec(_load_code('util'), util.__dict__)
    snap = _make_snap_module()
    injected = {'snap': snap, 'util': util}
    for opt in ('networkx', 'scipy'):          # imported but never used by the path (:3,:6)
        try:
            __import__(opt)
        except ImportError:
            injected[opt] = types.ModuleType(opt)
            if opt == 'scipy':
                injected['scipy'].spatial = types.ModuleType('scipy.spatial')
    saved = {k: sys.modules.get(k) for k in injected}
    sys.modules.update(injected)
    try:
        mod = types.ModuleType('reference_similarity')
        mod.__dict__['__name__'] = 'reference_similarity'     # keeps the __main__ block (:128-142) off
        with contextlib.redirect_stdout(io.StringIO()):
            exec(_load_code('similarity'), mod.__dict__)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cache['similarity'], _cache['util'], _cache['snap'] = mod, util, snap
    return mod


def manifest():
    with open(os.path.join(REF_DIR, 'MANIFEST.json')) as fh:
        return json.load(fh)


# --------------------------------------------------------------------------- running it
def run_main(example_file, graph_file, out_dir):
    """The reference's main() (similarity.py:11-18) on real files; returns the six score dicts
    read back with the reference's own util.load_json."""
    ref = load()
    util = _cache['util']
    names = ['u_cn', 'u_jaccard', 'u_adamic', 'b_cn', 'b_jaccard', 'b_adamic']
    paths = {n: os.path.join(out_dir, n + '.json') for n in names}
    with contextlib.redirect_stdout(io.StringIO()):          # :13,15,23,... and one print PER PAIR at :97,:100
        ref.main(example_file, graph_file, METHODS, [paths[n] for n in names[:3]],
                 METHODS, [paths[n] for n in names[3:]])
    return {n: util.load_json(paths[n]) for n in names}


def score_pair_arrays(edge_u, edge_b, pair_u, pair_b):
    """The reference's outputs for arrays of shared-space ids, one entry per pair.

    u_cn, u_jaccard, u_adamic, b_cn, b_jaccard come out of the reference's users() / business()
    through the files they write.  b_adamic is the reference's adamic_adar() called the way
    similarity.py:103 spells it -- that line is unreachable (the branch at :102 compares with a log
    string), so the file holds only the literal zeros; 'b_adamic_file' reports what the file has
    (None = pair absent).  in_graph as similarity.py:52,95 decide it.
    """
    ref = load()
    snap = _cache['snap']
    with tempfile.TemporaryDirectory() as d:
        gfile, efile = os.path.join(d, 'graph.txt'), os.path.join(d, 'examples.json')
        with open(gfile, 'w') as fh:
            for a, b in zip(edge_u, edge_b):
                fh.write('%d %d\n' % (int(a), int(b)))
        ex = {}
        for u, v in zip(pair_u, pair_b):
            ex.setdefault(str(int(u)), {})[str(int(v))] = 0
        _cache['util'].write_json(ex, efile)
        files = run_main(efile, gfile, d)
        G = snap.LoadEdgeList(snap.PUNGraph, gfile, 0, 1)
    nodes = set(N.GetId() for N in snap.Nodes(G))
    out = {k: [] for k in ('u_cn', 'u_jaccard', 'u_adamic', 'b_cn', 'b_jaccard', 'b_adamic',
                           'b_adamic_file', 'in_graph')}
    hop2, nbr = {}, {}

    def at_hop(x, h, memo):
        if x not in memo:
            vec = snap.TIntV()
            snap.GetNodesAtHop(G, x, h, vec, True)
            memo[x] = set(vec)
        return memo[x]

    for u, v in zip(pair_u, pair_b):
        su, sv, u, v = str(int(u)), str(int(v)), int(u), int(v)
        for k in ('u_cn', 'u_jaccard', 'u_adamic', 'b_cn', 'b_jaccard'):
            out[k].append(files[k][su][sv])
        out['b_adamic_file'].append(files['b_adamic'].get(su, {}).get(sv))
        ok = u in nodes and v in nodes
        out['in_graph'].append(1 if ok else 0)
        out['b_adamic'].append(ref.adamic_adar(at_hop(v, 2, hop2), at_hop(u, 1, nbr), G) if ok else 0)
    return out


def users_business(examples, G, write_dir):
    """users() + business() of the reference on an in-memory examples dict and a shim graph;
    used by bench.py's reference arm (files go to write_dir)."""
    ref = load()
    u_out = [os.path.join(write_dir, n) for n in ('u_cn.json', 'u_jaccard.json', 'u_adamic.json')]
    b_out = [os.path.join(write_dir, n) for n in ('b_cn.json', 'b_jaccard.json', 'b_adamic.json')]
    with contextlib.redirect_stdout(io.StringIO()):
        ref.users(examples, G, METHODS, u_out)
        ref.business(examples, G, METHODS, b_out)
