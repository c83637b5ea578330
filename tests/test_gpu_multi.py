"""The multi-GPU path on real hardware: one process per GPU under torchrun, NCCL for the plumbing.

Runs `tests/multi_worker.py` on every visible GPU (at least 2, at most 4; skipped on a one-GPU
box -- the driver's scaling run and `gpurun --gpus N` provide more).  Every rank scores its
user-aligned, work-balanced slice of ONE pair list (`dist.shard_bounds`); the rows of ALL ranks are
compared on rank 0 with the C oracle and with the unsharded single-GPU call, for both ways of
bringing them together:
  * `dist.score_sharded`   -- the scoring kernels store their rows straight into rank 0's peer-mapped
                              window; rank 0 derives pa (default) or jaccard and pa (blp_derive_pairs)
  * `dist.gather_results`  -- local scoring, then one grouped batch of NCCL send / recv
"""
import json
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(900)
def test_sharded_scoring_all_ranks_match_oracle(built_lib, tmp_path):
    import torch
    n_gpus = torch.cuda.device_count()
    if n_gpus < 2:
        pytest.skip('needs >= 2 GPUs (one process per GPU); this box has %d' % n_gpus)
    world = min(n_gpus, 4)
    report = str(tmp_path / 'report.json')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()),
           os.path.join(ROOT, 'tests', 'multi_worker.py'), report]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=840)
    assert res.returncode == 0, res.stdout[-4000:]
    rep = json.load(open(report))
    assert rep['world'] == world and len(rep['counts']) == world and min(rep['counts']) > 0
    assert rep['window_vs_oracle'] == 'ok', rep
    assert rep['window_vs_unsharded'] == 'ok', rep
    assert rep['nccl_gather_vs_window'] == 'ok', rep
    assert rep['rows_checked_per_rank'] and min(rep['rows_checked_per_rank']) > 0
    assert rep['bytes_per_pair_over_nvlink'] == 48 and rep['compact32_vs_window'] == 'ok'
