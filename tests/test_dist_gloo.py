"""World-size-2 gloo test of the N>1 path on CPU: shard -> score -> final gather == unsharded.

The scorer stand-in on CPU is the C oracle (tests may use it); on the GPU box the same
`shard_bounds` / `gather_results` run under NCCL -- and `score_sharded` with its peer-memory
result window -- in tests/test_gpu_multi.py (torchrun, one process per GPU) and in bench.py --gpus N.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, pkg


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import c_oracle
        synth, d = pkg('synth'), pkg('dist')
        cfg, eu, eb, pu, pv = synth.make_config('C1', n_pairs=6000)
        du, db = synth.degrees(cfg['n_users'], cfg['n_biz'], eu, eb)
        cost = d.pair_costs(pu, pv, du, db)
        bounds = d.shard_bounds(pu, world, cost)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        mine = c_oracle.score_pair_arrays(cfg['n_users'], cfg['n_biz'], eu, eb, pu[lo:hi], pv[lo:hi])
        local = {k: torch.from_numpy(v) for k, v in mine.items()}
        counts = [int(bounds[r + 1] - bounds[r]) for r in range(world)]
        got = d.gather_results(local, counts, dst=0)
        if rank == 0:
            whole = c_oracle.score_pair_arrays(cfg['n_users'], cfg['n_biz'], eu, eb, pu, pv)
            ok = all(np.array_equal(got[k].numpy(), whole[k]) for k in whole)
            q.put(('ok' if ok else 'mismatch', counts))
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_shard_score_gather_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    status, counts = q.get(timeout=10)
    assert status == 'ok'
    assert sum(counts) == 6000 and min(counts) > 0
