"""Developer tool: what the host link gives.  Pinned D2H / H2D bandwidth for the e2e step's byte
counts (560 MB back, 80 MB up), one copy vs the nine result columns, plus the box's topology --
the floor under bench.py's e2e figure."""
import os
import subprocess
import sys
import time

import torch

n = 10_000_000
dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
torch.cuda.set_device(dev)
print('cpus', len(os.sched_getaffinity(0)), flush=True)
for cmd in (['nvidia-smi', 'topo', '-m'], ['lscpu'], ['nvidia-smi', '--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max', '--format=csv']):
    try:
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=20).stdout
        if cmd[0] == 'lscpu':
            out = '\n'.join(l for l in out.split('\n') if any(k in l for k in ('NUMA', 'Socket', 'Model name', 'CPU(s):')))
        print(out, flush=True)
    except Exception as e:   # noqa: BLE001
        print(cmd, 'failed:', e)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


d_big = torch.empty(56 * n, dtype=torch.uint8, device=dev)
h_big = torch.empty(56 * n, dtype=torch.uint8).pin_memory()
t = timed(lambda: h_big.copy_(d_big, non_blocking=True))
print('D2H one 560 MB copy: %.2f ms  %.1f GB/s' % (t * 1e3, 56 * n / t / 1e9))
sizes = [4, 4, 8, 8, 4, 4, 8, 8, 8]
d_cols = [torch.empty(s * n, dtype=torch.uint8, device=dev) for s in sizes]
h_cols = [torch.empty(s * n, dtype=torch.uint8).pin_memory() for s in sizes]


def cols():
    for d, h in zip(d_cols, h_cols):
        h.copy_(d, non_blocking=True)


t = timed(cols)
print('D2H nine columns      : %.2f ms  %.1f GB/s' % (t * 1e3, 56 * n / t / 1e9))
d_in = torch.empty(8 * n, dtype=torch.uint8, device=dev)
h_in = torch.empty(8 * n, dtype=torch.uint8).pin_memory()
t = timed(lambda: d_in.copy_(h_in, non_blocking=True))
print('H2D one 80 MB copy    : %.2f ms  %.1f GB/s' % (t * 1e3, 8 * n / t / 1e9))
s2 = torch.cuda.Stream()


def both():
    h_big.copy_(d_big, non_blocking=True)
    with torch.cuda.stream(s2):
        d_in.copy_(h_in, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)


t = timed(both)
print('D2H 560 MB + H2D 80 MB concurrently: %.2f ms' % (t * 1e3))
