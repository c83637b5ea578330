"""Developer tool (torchrun, one process per GPU): bandwidth of SM-issued stores into rank 0's peer
window -- every peer alone, then all peers at once -- for 4 / 8 / 16-byte stores per thread.
Explains what the fused score+gather path can expect from plain epilogue stores.
usage: torchrun --nproc-per-node N tools/store_probe.py [MB per sender]"""
import ctypes, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
local = int(os.environ.get('LOCAL_RANK', rank))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
L = importlib.import_module('bipartite-link-prediction_b200._lib')
lib = L.load()
lib.blp_debug_store_probe.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 320
nbytes = mb << 20
base = ctypes.c_void_p()
handle = ctypes.create_string_buffer(64)
if rank == 0:
    L.check(lib.blp_peer_alloc(local, nbytes * world, ctypes.byref(base), handle), 'alloc')
box = [handle.raw if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
if rank != 0:
    L.check(lib.blp_peer_open(local, box[0], ctypes.byref(base)), 'open')
mine = base.value + nbytes * rank
st = torch.cuda.current_stream().cuda_stream


def run(width, streaming, blocks, active):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if active:
        for _ in range(3):
            lib.blp_debug_store_probe(ctypes.c_void_p(mine), nbytes, width, streaming, blocks, ctypes.c_void_p(st))
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 3], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


rows = []
for width in (4, 8, 16):
    for streaming in (1, 0):
        for blocks in (148 * 4, 148 * 16):
            run(width, streaming, blocks, rank != 0)            # warm
            ms_all = run(width, streaming, blocks, rank != 0)   # every peer stores at once
            ms_one = run(width, streaming, blocks, rank == 1)   # one peer alone
            rows.append({'width': width, 'streaming': streaming, 'blocks': blocks,
                         'all_peers_ms': ms_all, 'ingress_gbs_all': nbytes * (world - 1) / ms_all / 1e6,
                         'one_peer_ms': ms_one, 'gbs_one': nbytes / ms_one / 1e6})
            if rank == 0:
                print(json.dumps(rows[-1]), flush=True)
# the copy engine for comparison: peer cudaMemcpy of the same bytes, all peers at once
src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
for which in ('all', 'one'):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if (which == 'all' and rank != 0) or (which == 'one' and rank == 1):
        for _ in range(3):
            lib.blp_peer_push(local, ctypes.c_void_p(mine), ctypes.c_void_p(src.data_ptr()), nbytes, ctypes.c_void_p(st))
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 3], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        n_send = world - 1 if which == 'all' else 1
        print(json.dumps({'copy_engine': which, 'ms': float(t.item()), 'gbs': nbytes * n_send / float(t.item()) / 1e6}), flush=True)
dist.barrier()
if rank == 0:
    lib.blp_peer_free(local, base)
else:
    lib.blp_peer_close(local, base)
dist.destroy_process_group()
