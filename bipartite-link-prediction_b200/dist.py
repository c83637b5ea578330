"""Multi-GPU layer: shard the candidate pairs, replicate the adjacency, land the results on one rank.

The reference is a single process; the only parallelism this path offers is
over independent candidate pairs.  Every rank (one process per GPU) holds the whole graph -- it is
small next to 180 GB of HBM -- and scores one contiguous slice of the pair list.  There is no
collective inside the scoring.

How the slices meet again on rank `dst`:

  * ``ResultWindow`` + ``score_sharded``  (the product path).  Rank `dst` owns ONE device buffer
    holding every result column for the whole pair list and exposes it to its peers through CUDA
    IPC (``blp_peer_alloc`` / ``blp_peer_open``).  Each rank hands addresses inside that window to
    ``blp_score_pairs`` as its output pointers, so the scoring kernels' own epilogue stores carry
    every row over NVLink / NVSwitch while the rest of the slice is still being scored: compute
    and "gather" are one kernel, no SM runs a copy, no second pass over the results.
    pa (= deg(u) * deg(v)) is a function of the pair ids and the replicated degree arrays alone:
    it never crosses the link, `dst` derives it for the peers' rows (``blp_derive_pairs``) on a
    side stream under its own scoring kernels, so nothing is left to do once the peers' rows have
    landed (40 of the 48 reference bytes per pair on the wire).  jaccard (= cn / union, the kernels'
    own IEEE division) can be derived on `dst` as well (``derived=``: 32 or 36 bytes on the wire,
    a pass over the peers' rows after they have landed); measured at 8 GPUs all three variants
    work, the default is the fastest (profiles/r02_notes.md).  Derived columns are bit-identical
    to what the scoring kernels write themselves.
  * ``gather_results``  (the baseline it is measured against, and what the gloo CPU test uses):
    score into local memory, then move every column slice with one grouped batch of
    point-to-point send / recv straight into place (NCCL over NVLink on the box).

Slices are cut at user boundaries (no hop-2 set is built on two ranks for the user side) and
balanced on an estimate of the work, not on the pair count: the streamed-list length deg(v) of
every pair plus the two-hop expansion cost of every distinct user.
One ``blp_score_pairs`` call takes fewer than 2^31 pairs, i.e. a rank's slice must stay below that
(1 B pairs over 8 ranks = 125 M per rank).
"""
import ctypes

import numpy as np

# the seven outputs the reference leaves on the host (similarity.py:61,106 + PA): 48 B per pair
REFERENCE_COLUMNS = ('u_cn', 'u_jaccard', 'u_adamic', 'b_cn', 'b_jaccard', 'b_adamic', 'pa')
# ... plus the two union sizes (not reference outputs; tests and the roofline use them): 56 B
ALL_COLUMNS = REFERENCE_COLUMNS + ('u_union', 'b_union')
# the least a peer's kernels must store into the window (32 B per pair) ...
WIRE_COLUMNS = ('u_cn', 'u_union', 'u_adamic', 'b_cn', 'b_union', 'b_adamic')
# ... what `dst` CAN derive from them and the replicated graph, and what it derives by default
DERIVED_COLUMNS = ('u_jaccard', 'b_jaccard', 'pa')
DEFAULT_DERIVED = ('pa',)
_ITEMSIZE = {'cn': 4, 'union': 4, 'jaccard': 8, 'adamic': 8, 'pa': 8}


def _kind(col):
    return col if col == 'pa' else col[2:]


def pair_costs(pair_u, pair_b, deg_u, deg_b, expansion_u=None):
    """Per-pair work estimate in ids touched: deg(v) per pair + the user's expansion spread over
    its pairs.  Pairs with an id that is not in the graph cost 1."""
    pu = np.asarray(pair_u, dtype=np.int64)
    pv = np.asarray(pair_b, dtype=np.int64)
    ok = (pu >= 0) & (pu < deg_u.size) & (pv >= 0) & (pv < deg_b.size)
    cost = np.ones(pu.size, dtype=np.float64)
    cost[ok] += deg_b[pv[ok]] + deg_u[pu[ok]]
    if expansion_u is not None and ok.any():
        users, inv, cnt = np.unique(pu[ok], return_inverse=True, return_counts=True)
        cost[ok] += expansion_u[users][inv] / cnt[inv]
    return cost


def shard_bounds(pair_u, world, cost=None):
    """Cut [0, n) into `world` contiguous slices at user boundaries, equal in summed cost.

    `pair_u` must be grouped by user (as examples.json stores pairs).  Returns world+1 offsets.
    """
    pu = np.asarray(pair_u)
    n = pu.size
    if world <= 1 or n == 0:
        return np.array([0] + [n] * max(world, 1), dtype=np.int64)
    c = np.ones(n, dtype=np.float64) if cost is None else np.asarray(cost, dtype=np.float64)
    csum = np.cumsum(c)
    # positions where a new user starts are the only legal cut points
    starts = np.concatenate([[0], np.nonzero(pu[1:] != pu[:-1])[0] + 1, [n]])
    before = np.concatenate([[0.0], csum])[starts]          # cost in front of each cut point
    bounds = [0]
    for r in range(1, world):
        target = csum[-1] * r / world
        k = int(np.argmin(np.abs(before - target)))
        bounds.append(max(int(starts[k]), bounds[-1]))
    bounds.append(n)
    return np.asarray(bounds, dtype=np.int64)


def shard_pairs(pair_u, pair_b, rank, world, cost=None):
    """This rank's contiguous slice (views, no copy) and its [lo, hi) offsets."""
    b = shard_bounds(pair_u, world, cost)
    lo, hi = int(b[rank]), int(b[rank + 1])
    return pair_u[lo:hi], pair_b[lo:hi], (lo, hi)


# ------------------------------------------------------------------------------ the fused path
class _DeviceMemory(object):
    """A raw device range as something torch.as_tensor understands."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {'shape': (int(nbytes),), 'typestr': '|u1',
                                         'data': (int(ptr), False), 'version': 2}


def window_layout(n_total, columns=REFERENCE_COLUMNS, compact=True, derived=None):
    """The memory plan of a ResultWindow (pure function; no device needed).

    Returns a dict: 'columns' (what `dst` ends up with), 'derived' (what `dst` computes instead of
    receiving), 'allocated' (the columns the window holds: the requested ones plus the cn / union
    a derived jaccard needs), 'wire' (what a peer's kernels store), 'offsets' (byte offset of every
    allocated column, struct of arrays, 256-byte aligned), 'nbytes', 'wire_bytes_per_pair'.
    """
    columns = tuple(columns)
    bad = [c for c in columns if c not in ALL_COLUMNS]
    if bad:
        raise ValueError('unknown result columns %r' % (bad,))
    # which columns `dst` derives instead of receiving.  Default ('pa',): 40 B per pair on the
    # wire and nothing left to do after the peers' rows have landed, pa being a function of
    # the pair ids alone; DERIVED_COLUMNS: 32 B on the wire, jaccard derived afterwards
    derived = tuple(DEFAULT_DERIVED if derived is None else derived) if compact else ()
    bad = [c for c in derived if c not in DERIVED_COLUMNS]
    if bad:
        raise ValueError('cannot derive %r on the destination' % (bad,))
    derived = tuple(c for c in derived if c in columns)
    alloc = list(columns)
    for side in 'ub':                            # jaccard is derived from (cn, union) on `dst`
        if side + '_jaccard' in derived:
            alloc += [c for c in (side + '_cn', side + '_union') if c not in alloc]
    offsets, at = {}, 0
    for c in alloc:                              # struct of arrays, 256-byte aligned columns
        offsets[c] = at
        at += (_ITEMSIZE[_kind(c)] * max(int(n_total), 1) + 255) // 256 * 256
    wire = tuple(c for c in alloc if c not in derived)
    return {'columns': columns, 'derived': derived, 'allocated': tuple(alloc), 'wire': wire,
            'offsets': offsets, 'nbytes': max(at, 256),
            'wire_bytes_per_pair': sum(_ITEMSIZE[_kind(c)] for c in wire)}


class ResultWindow(object):
    """The result table of a sharded scoring job: every column for all `n_total` pairs, resident
    on rank `dst`, writable by the scoring kernels of every rank (see the module docstring).

    Collective: every rank of `group` constructs it (the IPC handle is broadcast from `dst`).
    ``columns`` selects what `dst` ends up with; a column that is left out is not computed.
    With ``compact`` (default) the peers do not store the columns in ``derived`` (default
    DEFAULT_DERIVED = pa; any subset of DERIVED_COLUMNS): `dst` computes them (``derive``); the
    union columns a derived jaccard needs are kept internally even when they were not asked for.
    """

    def __init__(self, graph, n_total, columns=REFERENCE_COLUMNS, dst=0, group=None, compact=True,
                 derived=None):
        import torch.distributed as dist
        from . import _lib
        self._lib = _lib.load()
        self.device = graph.device
        self.n_total = int(n_total)
        plan = window_layout(self.n_total, columns, compact, derived)
        self.columns = plan['columns']
        self.compact = bool(compact)
        self.derived = plan['derived']
        self.graph = graph
        self._alloc_columns = plan['allocated']
        self.offsets, self.nbytes = plan['offsets'], plan['nbytes']
        self.dst = dst
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._base = ctypes.c_void_p()
        self._owner = self.rank == dst
        dev = self.device.index or 0
        handle = ctypes.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        if self._owner:
            _lib.check(self._lib.blp_peer_alloc(dev, self.nbytes, ctypes.byref(self._base), handle),
                       'blp_peer_alloc')
        if self.world > 1:
            box = [handle.raw if self._owner else None]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, dst) if group else dst,
                                       group=group)
            if not self._owner:
                _lib.check(self._lib.blp_peer_open(dev, box[0], ctypes.byref(self._base)),
                           'blp_peer_open')
        self.base = int(self._base.value)

    def _staging(self, n, kinds):
        """Local staging of a peer's business-side columns (kept until the next call)."""
        import torch
        dt = {'cn': torch.int32, 'union': torch.int32, 'jaccard': torch.float64,
              'adamic': torch.float64}
        st = getattr(self, '_stage', None)
        if st is None or any(k not in st or st[k].numel() != n for k in kinds):
            st = self._stage = {k: torch.empty(n, dtype=dt[k], device=self.device) for k in kinds}
        return {k: st[k] for k in kinds}

    def _push_stream(self):
        import torch
        if getattr(self, '_side', None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def written_columns(self):
        """The columns THIS rank's kernels store: everything on `dst` (its own rows need no
        second pass), the wire columns only on a peer of a compact window."""
        if self._owner or not self.compact:
            return self._alloc_columns
        return tuple(c for c in self._alloc_columns if c not in self.derived)

    def pointers(self, lo, columns=None):
        """Device addresses of row `lo` of the columns, valid on THIS rank's device."""
        cols = self._alloc_columns if columns is None else columns
        return {c: self.base + self.offsets[c] + _ITEMSIZE[_kind(c)] * int(lo) for c in cols}

    def bytes_per_pair(self):
        """Bytes per pair a peer sends over the link."""
        cols = tuple(c for c in self._alloc_columns if c not in self.derived)
        return sum(_ITEMSIZE[_kind(c)] for c in cols)

    def attach_pairs(self, d_all_u, d_all_b, own):
        """On `dst`: the whole pair list on its device and its own row range [lo, hi).  With the
        pairs attached, ``score_into_window`` derives pa for the peers' rows (a function of the
        pair ids alone) on a side stream under dst's own scoring kernels, and ``derive`` with no
        arguments fills what is left once the peers' rows have landed."""
        self._pairs = (d_all_u, d_all_b, (int(own[0]), int(own[1])))
        self._early = ()

    def derive(self, d_all_u=None, d_all_b=None, own=None, stream=None, which=None):
        """On `dst` of a compact window: fill DERIVED_COLUMNS (or the subset `which`) for the rows
        the PEERS wrote, i.e. all rows outside own = [lo, hi) (dst's kernels wrote every column
        of its own rows).  d_all_u / d_all_b: the whole pair list on dst's device (default: what
        attach_pairs was given; then the columns already derived early are skipped).
        Asynchronous on `stream`; the caller orders it after the peers' scoring (e.g. behind a
        collective on the same stream, ``rows_landed``)."""
        if not (self._owner and self.compact):
            return
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        skip = ()
        if d_all_u is None:
            if getattr(self, '_pairs', None) is None:
                raise ValueError('derive() needs the pair list (pass it or call attach_pairs)')
            d_all_u, d_all_b, own = self._pairs
            if which is None:
                skip = self._early
        want = [c for c in self._alloc_columns if c in self.derived and c not in skip and
                (which is None or c in which)]
        if not want:
            return
        lo, hi = int(own[0]), int(own[1])
        for a, b in ((0, lo), (hi, self.n_total)):
            if b <= a:
                continue
            p = self.pointers(a)

            def ptr(c, need):
                return ctypes.c_void_p(p[c]) if (c in p and need) else None
            uj, bj, pa = 'u_jaccard' in want, 'b_jaccard' in want, 'pa' in want
            from . import _lib
            _lib.check(self._lib.blp_derive_pairs(
                self.graph._h, ctypes.c_void_p(d_all_u.data_ptr() + 4 * a),
                ctypes.c_void_p(d_all_b.data_ptr() + 4 * a), b - a,
                ptr('u_cn', uj), ptr('u_union', uj), ptr('b_cn', bj), ptr('b_union', bj),
                ptr('u_jaccard', uj), ptr('b_jaccard', bj), ptr('pa', pa),
                ctypes.c_void_p(stream.cuda_stream)), 'blp_derive_pairs')

    def tensors(self):
        """On `dst`: dict column -> torch tensor [n_total] viewing the window.  None elsewhere
        (the peers only ever store into it from kernels)."""
        if not self._owner:
            return None
        import torch
        raw = torch.as_tensor(_DeviceMemory(self.base, self.nbytes), device=self.device)
        dt = {'cn': torch.int32, 'union': torch.int32, 'jaccard': torch.float64,
              'adamic': torch.float64, 'pa': torch.int64}
        out = {}
        for c in self.columns:
            k = _kind(c)
            out[c] = raw[self.offsets[c]:self.offsets[c] + _ITEMSIZE[k] * self.n_total].view(dt[k])
        self._keepalive = raw
        return out

    def close(self):
        if getattr(self, 'base', 0):
            dev = self.device.index or 0
            if self._owner:
                self._lib.blp_peer_free(dev, self._base)
            else:
                self._lib.blp_peer_close(dev, self._base)
            self.base = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def score_into_window(graph, d_u, d_b, window, lo, stream=None):
    """Score the device-resident pairs (d_u, d_b) -- rows [lo, lo+n) of the job -- into the window.
    Asynchronous on `stream`.

    On the window's owner both sides write their rows in place.  On a peer:
      * the USER side (the long one) runs with the window's columns as its output arrays: the
        scoring kernels' own epilogue stores carry the rows over NVLink while the kernels run;
      * the BUSINESS side runs first, into local staging: its results appear in one burst (the
        un-permute pass writes every column in a fraction of a millisecond), which seven peers
        cannot push through one GPU's NVLink ingress at once without stalling -- so its columns
        go over by copy engine (``blp_peer_push``) on a side stream, under the user-side kernels.
    """
    import torch
    from . import _lib
    if stream is None:
        stream = torch.cuda.current_stream(graph.device)
    cols = window.written_columns()
    ptr = window.pointers(lo, cols)
    up = {(_kind(c)): p for c, p in ptr.items() if c.startswith('u_') or c == 'pa'}
    bp = {(_kind(c)): p for c, p in ptr.items() if c.startswith('b_')}
    n = int(d_u.numel())
    if window._owner or window.world == 1 or not bp or n == 0:
        early = window._owner and window.world > 1 and getattr(window, '_pairs', None) is not None
        if early:
            # pa of the peers' rows depends on the pair ids alone: derived now, on the side
            # stream, under this rank's own scoring kernels
            side = window._push_stream()
            side.wait_stream(stream)
            window.derive(*window._pairs, stream=side, which=('pa',))
            window._early = ('pa',)
        if up and n:
            graph.score_side(_lib.SIDE_USER, d_u, d_b, want=(), out_ptr=up, stream=stream)
        if bp and n:
            graph.score_side(_lib.SIDE_BUSINESS, d_u, d_b, want=(), out_ptr=bp, stream=stream)
        if early:
            stream.wait_stream(side)
        return
    stage = window._staging(n, sorted(bp))
    graph.score_side(_lib.SIDE_BUSINESS, d_u, d_b, want=tuple(sorted(bp)), out=stage, stream=stream)
    side = window._push_stream()
    ev = torch.cuda.Event()
    ev.record(stream)
    side.wait_event(ev)
    dev = graph.device.index or 0
    for k in sorted(bp):
        _lib.check(window._lib.blp_peer_push(dev, ctypes.c_void_p(bp[k]), ctypes.c_void_p(stage[k].data_ptr()),
                                             stage[k].numel() * stage[k].element_size(),
                                             ctypes.c_void_p(side.cuda_stream)), 'blp_peer_push')
    if up:
        graph.score_side(_lib.SIDE_USER, d_u, d_b, want=(), out_ptr=up, stream=stream)
    stream.wait_stream(side)


def score_sharded(graph, pair_u, pair_b, cost=None, dst=0, group=None, columns=REFERENCE_COLUMNS,
                  window=None):
    """Score this rank's slice of the (replicated) pair list on `graph`; the results of every rank
    land in one window on `dst` through the kernels' own stores.

    pair_u / pair_b are host int32 arrays holding the WHOLE pair list on every rank (grouped by
    user).  Returns (dict column -> tensor over all pairs, window) on `dst`, (None, window)
    elsewhere; keep the window alive as long as the tensors are used, close() it afterwards.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    bounds = shard_bounds(pair_u, world, cost)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    if window is None:
        window = ResultWindow(graph, len(pair_u), columns=columns, dst=dst, group=group)
    with torch.cuda.device(graph.device):
        du = torch.from_numpy(np.ascontiguousarray(pair_u[lo:hi], dtype=np.int32)).to(graph.device)
        db = torch.from_numpy(np.ascontiguousarray(pair_b[lo:hi], dtype=np.int32)).to(graph.device)
        if world > 1 and rank == dst and window.compact:
            all_u = torch.from_numpy(np.ascontiguousarray(pair_u, dtype=np.int32)).to(graph.device)
            all_b = torch.from_numpy(np.ascontiguousarray(pair_b, dtype=np.int32)).to(graph.device)
            window.attach_pairs(all_u, all_b, (lo, hi))
        score_into_window(graph, du, db, window, lo)
        if world > 1:
            rows_landed(graph, group)            # stream-ordered: behind every rank's scoring
            if rank == dst and window.compact:
                window.derive()
        torch.cuda.synchronize(graph.device)
    return window.tensors(), window


def rows_landed(graph, group=None):
    """Orders the caller's CUDA stream behind the scoring kernels of EVERY rank: a one-element
    all-reduce issued on each rank's stream after its scoring.  A rank's stores into the peer
    window are complete when its scoring kernels are, i.e. before its part of the collective runs,
    so whatever `dst` enqueues after this call sees all rows.  No host synchronisation."""
    import torch
    import torch.distributed as dist
    tok = getattr(graph, '_rows_token', None)
    if tok is None:
        tok = graph._rows_token = torch.zeros(1, dtype=torch.int32, device=graph.device)
    dist.all_reduce(tok, group=group)


# ------------------------------------------------------------------------------ baseline gather
def gather_results(local, counts, dst=0, group=None, out=None):
    """Final gather of per-rank result columns to rank `dst` with point-to-point transfers.

    local  : dict name -> 1-D torch tensor of this rank's slice (same names/dtypes on every rank)
    counts : list of slice lengths per rank (known to all ranks from shard_bounds)
    Returns dict name -> tensor over all pairs on `dst` (`out` may carry preallocated ones), None
    elsewhere.  Slices are unequal, so every column slice is sent on its own, straight into its
    place -- one grouped batch of send / recv (one ncclGroup on the nccl backend; gloo in tests).
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if len(counts) != world:
        raise ValueError('counts must list one slice length per rank')
    names = sorted(local)
    for name in names:
        if local[name].numel() != counts[rank]:
            raise ValueError('%s: slice has %d entries, counts says %d' %
                             (name, local[name].numel(), counts[rank]))
    starts = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    gdst = dist.get_global_rank(group, dst) if group is not None else dst
    ops = []
    if rank == dst:
        if out is None:
            out = {k: torch.empty(int(starts[-1]), dtype=local[k].dtype, device=local[k].device)
                   for k in names}
        for k in names:
            out[k][int(starts[rank]):int(starts[rank + 1])].copy_(local[k])
            for r in range(world):
                if r != rank and counts[r] > 0:
                    src = dist.get_global_rank(group, r) if group is not None else r
                    ops.append(dist.P2POp(dist.irecv, out[k][int(starts[r]):int(starts[r + 1])], src,
                                          group=group))
    elif counts[rank] > 0:
        for k in names:
            ops.append(dist.P2POp(dist.isend, local[k].contiguous(), gdst, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out if rank == dst else None
