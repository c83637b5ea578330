"""Multi-GPU layer: shard the candidate pairs, replicate the adjacency, gather the results.

The reference is a single process; the only parallelism this path offers is
over independent candidate pairs.  Every rank holds the whole graph (it is small next to
180 GB of HBM), scores one contiguous slice of the pair list and the slices meet again in ONE
final gather -- NCCL over NVLink on the GPU box, gloo in the CPU tests.  There is no collective
inside the scoring itself.

Slices are cut at user boundaries (no hop-2 set is built on two ranks for the user side) and
balanced on an estimate of the work, not on the pair count: the streamed-list length deg(v) of
every pair plus the two-hop expansion cost of every distinct user.
"""
import numpy as np


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


def gather_results(local, counts, dst=0, group=None):
    """Final gather of per-rank result columns to rank `dst`.

    local  : dict name -> 1-D torch tensor of this rank's slice (same names/dtypes on every rank)
    counts : list of slice lengths per rank (known to all ranks from shard_bounds)
    Returns dict name -> concatenated tensor on `dst`, None elsewhere.  Slices are unequal, so
    each column is padded to the longest slice for the collective and trimmed afterwards.
    Works with the nccl backend (CUDA tensors, NVLink) and with gloo (CPU tensors, tests).
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if len(counts) != world:
        raise ValueError('counts must list one slice length per rank')
    longest = int(max(counts))
    out = {} if rank == dst else None
    for name in sorted(local):
        t = local[name]
        if t.numel() != counts[rank]:
            raise ValueError('%s: slice has %d entries, counts says %d' %
                             (name, t.numel(), counts[rank]))
        send = t
        if t.numel() != longest:
            send = torch.zeros(longest, dtype=t.dtype, device=t.device)
            send[:t.numel()] = t
        recv = None
        if rank == dst:
            recv = [torch.empty(longest, dtype=t.dtype, device=t.device) for _ in range(world)]
        dist.gather(send.contiguous(), recv, dst=dst, group=group)
        if rank == dst:
            out[name] = torch.cat([recv[r][:counts[r]] for r in range(world)])
    return out


def score_sharded(graph, pair_u, pair_b, cost=None, dst=0, group=None):
    """Score this rank's slice of the (replicated) pair list on `graph` and gather on `dst`.

    pair_u / pair_b are host int32 arrays holding the WHOLE pair list on every rank (grouped by
    user).  Returns the gathered dict of tensors on `dst` (caller order), None elsewhere.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bounds = shard_bounds(pair_u, world, cost)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    du = torch.from_numpy(np.ascontiguousarray(pair_u[lo:hi], dtype=np.int32)).to(graph.device)
    db = torch.from_numpy(np.ascontiguousarray(pair_b[lo:hi], dtype=np.int32)).to(graph.device)
    local = graph.score_pairs(du, db)
    counts = [int(bounds[r + 1] - bounds[r]) for r in range(world)]
    return gather_results(local, counts, dst=dst, group=group)


def score_and_gather_overlapped(graph, d_u, d_b, chunks=4, dst=0, group=None, out=None, recv=None,
                                reserve_sms=8):
    """Score this rank's pairs in `chunks` slices and gather every slice on `dst` while the next
    one is being scored (the gather rides a side stream; NCCL moves it over NVLink).

    d_u / d_b: int32 CUDA tensors, the SAME length on every rank (pad with -1 if needed).
    Returns (out, recv): this rank's result columns and, on `dst`, dict name -> [world, n] tensor.
    `out` / `recv` from a previous call may be passed back in to reuse the buffers.
    `reserve_sms` SMs are kept out of the persistent scoring grids for the duration of the call,
    so that NCCL's send/receive kernels can run beside them.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = d_u.numel()
    dev = graph.device
    main = torch.cuda.current_stream(dev)
    if getattr(graph, '_comm_stream', None) is None:
        graph._comm_stream = torch.cuda.Stream(device=dev)
    comm = graph._comm_stream
    dtypes = {'cn': torch.int32, 'union': torch.int32, 'jaccard': torch.float64,
              'adamic': torch.float64}
    keys = ['u_' + k for k in dtypes] + ['b_' + k for k in dtypes] + ['pa']
    if out is None:
        out = {k: torch.empty(n, dtype=dtypes.get(k[2:], torch.int64), device=dev) for k in keys}
    if recv is None and rank == dst:
        recv = {k: torch.empty((world, n), dtype=out[k].dtype, device=dev) for k in keys}
    from . import _lib
    chunks = max(1, min(int(chunks), n // 65536 or 1))
    bounds = [(n * c) // chunks for c in range(chunks + 1)]
    comm.wait_stream(main)
    graph.reserve_sms(reserve_sms)

    def gather_async(names, lo, hi, after):
        ev = torch.cuda.Event()
        ev.record(after)
        with torch.cuda.stream(comm):
            comm.wait_event(ev)
            for k in names:
                glist = [recv[k][r, lo:hi] for r in range(world)] if rank == dst else None
                dist.gather(out[k][lo:hi], glist, dst=dst, group=group)

    # business side once over all pairs (its hop-2 sets are shared by pairs of every slice), on a
    # side stream so that it fills in beside the user-side slices ...
    bkeys = [k for k in keys if k.startswith('b_')]
    ukeys = [k for k in keys if not k.startswith('b_')]
    if getattr(graph, '_side_stream', None) is None:
        graph._side_stream = torch.cuda.Stream(device=dev)
    side = graph._side_stream
    side.wait_stream(main)
    with torch.cuda.stream(side):
        graph.score_side(_lib.SIDE_BUSINESS, d_u, d_b, out={k[2:]: out[k] for k in bkeys},
                         stream=side)
    gather_async(bkeys, 0, n, side)      # queued first: it is the largest transfer
    # ... and the user side slice by slice, each slice's gather behind the next slice's scoring
    for c in range(chunks):
        lo, hi = bounds[c], bounds[c + 1]
        if hi <= lo:
            continue
        ou = {(k[2:] if k.startswith('u_') else k): out[k][lo:hi] for k in ukeys}
        graph.score_side(_lib.SIDE_USER, d_u[lo:hi], d_b[lo:hi], want_pa=True, out=ou)
        gather_async(ukeys, lo, hi, main)
    main.wait_stream(side)
    graph.reserve_sms(0)
    main.wait_stream(comm)
    return out, recv
