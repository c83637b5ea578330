"""torchrun worker of tests/test_gpu_multi.py (not collected by pytest: no test_ prefix).

Every rank: same seeded C1-shaped graph and pair list -> dist.shard_bounds -> score its slice
  (a) through dist.score_sharded: kernels store into rank 0's peer-mapped ResultWindow,
  (b) locally + dist.gather_results (grouped NCCL send/recv).
Rank 0 checks the rows of EVERY rank against the C oracle and against its own unsharded call.
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = 'bipartite-link-prediction_b200'


def main():
    import torch
    import torch.distributed as dist
    report_path = sys.argv[1]
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    graph = importlib.import_module(PKG + '.graph')
    synth = importlib.import_module(PKG + '.synth')
    d = importlib.import_module(PKG + '.dist')
    try:
        cfg, eu, eb, pu, pv = synth.make_config('C1', n_pairs=60_000)
        # some candidates that are existing edges, and an unsorted tail is NOT added: the list stays
        # grouped by user, as shard_bounds requires
        G = graph.BipartiteGraph(cfg['n_users'], cfg['n_biz'], eu, eb, device=local)
        du, db = synth.degrees(cfg['n_users'], cfg['n_biz'], eu, eb)
        cost = d.pair_costs(pu, pv, du, db)
        bounds = d.shard_bounds(pu, world, cost)
        counts = [int(bounds[r + 1] - bounds[r]) for r in range(world)]
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])

        # (a) fused: every rank's kernels write rank 0's window
        cols, window = d.score_sharded(G, pu, pv, cost=cost, columns=d.ALL_COLUMNS)
        # (a') the same with the 32-byte wire format: rank 0 derives jaccard of both sides and pa
        w32 = d.ResultWindow(G, len(pu), columns=d.ALL_COLUMNS, derived=d.DERIVED_COLUMNS)
        cols32, _ = d.score_sharded(G, pu, pv, cost=cost, window=w32)
        same32 = True
        if rank == 0:
            same32 = all(torch.equal(cols32[k], cols[k]) for k in cols) and w32.bytes_per_pair() == 32
        # (b) baseline: local scoring + grouped NCCL send/recv
        t_u = torch.from_numpy(pu[lo:hi]).to(dev)
        t_b = torch.from_numpy(pv[lo:hi]).to(dev)
        mine = G.score_pairs(t_u, t_b)
        torch.cuda.synchronize()
        gathered = d.gather_results(mine, counts, dst=0)
        torch.cuda.synchronize()
        rep = None
        if rank == 0:
            from oracle import c_oracle
            want = c_oracle.score_pair_arrays(cfg['n_users'], cfg['n_biz'], eu, eb, pu, pv)
            got = {k: v.cpu().numpy() for k, v in cols.items()}
            bad = []
            for k in ('u_cn', 'u_union', 'b_cn', 'b_union', 'pa', 'u_jaccard', 'b_jaccard'):
                if not np.array_equal(got[k].astype(want[k].dtype), want[k]):
                    bad.append(k)
            for k in ('u_adamic', 'b_adamic'):
                if not np.allclose(got[k], want[k], rtol=1e-8, atol=0):
                    bad.append(k)
            whole = G.score_pairs_host(pu, pv)
            bad_whole = [k for k in got if not np.array_equal(got[k], whole[k])]
            bad_nccl = [k for k in got if not np.array_equal(gathered[k].cpu().numpy(), got[k])]
            rep = {'world': world, 'counts': counts,
                   'window_vs_oracle': 'ok' if not bad else 'mismatch in %s' % bad,
                   'window_vs_unsharded': 'ok' if not bad_whole else 'mismatch in %s' % bad_whole,
                   'nccl_gather_vs_window': 'ok' if not bad_nccl else 'mismatch in %s' % bad_nccl,
                   'rows_checked_per_rank': counts,
                   'compact32_vs_window': 'ok' if same32 else 'mismatch',
                   'bytes_per_pair_over_nvlink': window.bytes_per_pair()}
        dist.barrier()
        window.close()
        w32.close()
        if rank == 0:
            with open(report_path, 'w') as fh:
                json.dump(rep, fh)
    finally:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
