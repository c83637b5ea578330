// Peer windows: the multi-GPU result gather without a collective.
//
// The reference is one process; here every rank scores a slice of the pair
// list and the slices have to meet on one rank.  Instead of scoring into local memory and then
// moving the columns with a gather, the destination rank exposes ONE device buffer (a "window")
// to its peers through CUDA IPC; every rank passes addresses inside that window as the output
// pointers of blp_score_pairs, so the scoring kernels' own epilogue stores (st.global on a
// peer-mapped address) carry each result over NVLink / NVSwitch as it is produced -- the transfer
// rides under the computation, no SM runs a copy kernel, nothing is staged.  One process per GPU:
// an IPC handle cannot be opened by the process that exported it.
#include <algorithm>
#include <cstring>

#include "blp_internal.h"

static_assert(sizeof(cudaIpcMemHandle_t) == BLP_IPC_HANDLE_BYTES, "handle size is part of the ABI");

// The columns that are exact functions of the others and of the (replicated) graph need not be
// sent over the link: the destination rank can derive them where they are needed (pa always is;
// jaccard is an option of dist.ResultWindow).
//   jaccard = (double)cn / (double)union   (similarity.py:108-111; IEEE division, the very
//             expression the scoring kernels evaluate) -- 0.0 for a pair that is not in the graph
//   pa      = deg(u) * deg(v)              ("Link prediction.R":400-415), 0 when not in the graph
namespace blp {
namespace {
__global__ void k_derive(const int* __restrict__ pu, const int* __restrict__ pb, long long n, int n_users,
                         int n_biz, const int* __restrict__ u_deg, const int* __restrict__ b_deg,
                         const int* __restrict__ u_cn, const int* __restrict__ u_uni,
                         const int* __restrict__ b_cn, const int* __restrict__ b_uni,
                         double* __restrict__ u_jac, double* __restrict__ b_jac, long long* __restrict__ pa) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        if (u_jac) {
            const int c = __ldcs(u_cn + i), u = __ldcs(u_uni + i);
            __stcs(u_jac + i, u > 0 ? __ddiv_rn((double)c, (double)u) : 0.0);
        }
        if (b_jac) {
            const int c = __ldcs(b_cn + i), u = __ldcs(b_uni + i);
            __stcs(b_jac + i, u > 0 ? __ddiv_rn((double)c, (double)u) : 0.0);
        }
        if (pa) {
            const int x = __ldcs(pu + i), y = __ldcs(pb + i);
            long long v = 0;
            if (x >= 0 && x < n_users && y >= 0 && y < n_biz) {
                const int du = u_deg[x], db = b_deg[y];
                if (du > 0 && db > 0) v = (long long)du * (long long)db;   // both ids in the graph
            }
            __stcs(pa + i, v);
        }
    }
}
}  // namespace
}  // namespace blp

extern "C" int blp_derive_pairs(blp_graph* g, const int32_t* pair_u, const int32_t* pair_b, int64_t n,
                                const int32_t* u_cn, const int32_t* u_union, const int32_t* b_cn,
                                const int32_t* b_union, double* u_jaccard, double* b_jaccard, int64_t* pa,
                                void* stream) {
    if (!g || n < 0 || (n > 0 && ((u_jaccard && (!u_cn || !u_union)) || (b_jaccard && (!b_cn || !b_union)) ||
                                  (pa && (!pair_u || !pair_b))))) {
        blp::set_error("blp_derive_pairs: bad argument");
        return BLP_ERR_INVALID;
    }
    if (n == 0 || (!u_jaccard && !b_jaccard && !pa)) return BLP_OK;
    BLP_ON_DEVICE(g->device);
    const int blocks = (int)std::min<long long>((n + 255) / 256, (long long)g->sm_count * 16);
    blp::k_derive<<<blocks, 256, 0, (cudaStream_t)stream>>>(pair_u, pair_b, n, g->n_users, g->n_biz, g->u_deg,
                                                          g->b_deg, u_cn, u_union, b_cn, b_union, u_jaccard,
                                                          b_jaccard, (long long*)pa);
    BLP_CUDA_TRY(cudaGetLastError());
    return BLP_OK;
}

// Developer diagnostic (not part of the ABI; tools/store_probe.py): SM-issued stores of `width`
// bytes per thread, coalesced, into `dst` (typically a peer window) -- how fast can kernels push
// results to another GPU with plain stores, alone and with every peer doing the same?
namespace blp {
namespace {
template <typename T>
__global__ void k_store_probe(T* dst, long long n, int streaming) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    T v;
    memset(&v, 0x5a, sizeof(T));
    for (; i < n; i += stride) {
        if (streaming) __stcs(dst + i, v);
        else dst[i] = v;
    }
}
}  // namespace
}  // namespace blp

extern "C" int blp_debug_store_probe(void* dst, int64_t bytes, int width, int streaming, int blocks, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (width == 4) blp::k_store_probe<int><<<blocks, 256, 0, st>>>((int*)dst, bytes / 4, streaming);
    else if (width == 8) blp::k_store_probe<double><<<blocks, 256, 0, st>>>((double*)dst, bytes / 8, streaming);
    else if (width == 16) blp::k_store_probe<int4><<<blocks, 256, 0, st>>>((int4*)dst, bytes / 16, streaming);
    else return BLP_ERR_INVALID;
    BLP_CUDA_TRY(cudaGetLastError());
    return BLP_OK;
}

extern "C" int blp_peer_push(int device, void* dst, const void* src, int64_t bytes, void* stream) {
    if (bytes < 0 || (bytes > 0 && (!dst || !src))) {
        blp::set_error("blp_peer_push: bad argument");
        return BLP_ERR_INVALID;
    }
    if (bytes == 0) return BLP_OK;
    BLP_ON_DEVICE(device);
    // a copy-engine transfer: no SM is involved, so it runs beside the scoring kernels
    BLP_CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return BLP_OK;
}

extern "C" int blp_peer_alloc(int device, int64_t bytes, void** dev_ptr, unsigned char* handle_out) {
    if (!dev_ptr || !handle_out || bytes <= 0) {
        blp::set_error("blp_peer_alloc: bad argument");
        return BLP_ERR_INVALID;
    }
    *dev_ptr = nullptr;
    BLP_ON_DEVICE(device);
    void* p = nullptr;
    // plain cudaMalloc on purpose: pool / virtual-memory allocations cannot be exported this way
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        blp::set_error("blp_peer_alloc: device allocation failed");
        return BLP_ERR_OOM;
    }
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return blp::cuda_fail(e, "cudaIpcGetMemHandle", __FILE__, __LINE__);
    }
    memcpy(handle_out, &h, sizeof(h));
    *dev_ptr = p;
    return BLP_OK;
}

extern "C" int blp_peer_free(int device, void* dev_ptr) {
    if (!dev_ptr) return BLP_OK;
    BLP_ON_DEVICE(device);
    BLP_CUDA_TRY(cudaDeviceSynchronize());
    BLP_CUDA_TRY(cudaFree(dev_ptr));
    return BLP_OK;
}

extern "C" int blp_peer_open(int device, const unsigned char* handle, void** dev_ptr) {
    if (!dev_ptr || !handle) {
        blp::set_error("blp_peer_open: bad argument");
        return BLP_ERR_INVALID;
    }
    *dev_ptr = nullptr;
    BLP_ON_DEVICE(device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    // maps the owner's allocation into this process and enables peer access device -> owner
    BLP_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr = p;
    return BLP_OK;
}

extern "C" int blp_peer_close(int device, void* dev_ptr) {
    if (!dev_ptr) return BLP_OK;
    BLP_ON_DEVICE(device);
    BLP_CUDA_TRY(cudaDeviceSynchronize());   // stores of kernels still in flight land first
    BLP_CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
    return BLP_OK;
}
