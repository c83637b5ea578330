// Candidate generation on the device: for every given user the
// businesses at BFS distance exactly 3, i.e. what make_examples obtains from
// snap.GetNodesAtHop(G, u, 3, candidate_businesses, True) (dataset_maker.py:137-139):
//
//   hop2(u) = ( U_{b in N(u)} N(b) ) \ {u}          user bitmap in shared memory
//   hop3(u) = ( U_{w in hop2(u)} N(w) ) \ N(u)      business bitmap in shared memory
//
// One CTA per user, pulled from an atomic counter.  The first call counts, the second writes the
// ids in ascending order at caller-provided offsets.
#include <climits>
#include <cstdio>

#include "blp_internal.h"

namespace blp {
namespace {

constexpr unsigned kAll = 0xffffffffu;

__device__ __forceinline__ int h3_deg(unsigned long long row) { return (int)(row & 0xffffffull); }
__device__ __forceinline__ const int* h3_list(const int* adj, unsigned long long row) {
    return adj + (long long)((row >> 24) & ((1ull << BLP_ROW_FIRST4_BITS) - 1)) * 4;
}

struct Hop3Args {
    const unsigned long long* __restrict__ u_row;
    const int* __restrict__ u_adj;
    const unsigned long long* __restrict__ b_row;
    const int* __restrict__ b_adj;
    int n_users, n_biz, uw, bw;            // bitmap words
    const int* __restrict__ users;
    long long n;
    long long* counts;                     // count pass
    const long long* offsets;              // fill pass
    int* out_biz;
    int* work_counter;
};

template <bool FILL>
__global__ void __launch_bounds__(256) k_hop3(Hop3Args a) {
    extern __shared__ __align__(16) unsigned h3_smem[];
    unsigned* ubm = h3_smem;               // users at distance 2
    unsigned* bbm = h3_smem + a.uw;        // businesses at distance 3
    __shared__ int s_item;
    __shared__ int s_scan[9];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const long long it = s_item;
        if (it >= a.n) break;
        const int u = a.users[it];
        const bool ok = u >= 0 && u < a.n_users && h3_deg(a.u_row[u]) > 0;
        if (!ok) {
            if (!FILL && tid == 0) a.counts[it] = 0;
            continue;
        }
        for (int i = tid; i < a.uw + a.bw; i += 256) h3_smem[i] = 0u;
        __syncthreads();
        const unsigned long long urow = a.u_row[u];
        const int* nu = h3_list(a.u_adj, urow);
        const int du = h3_deg(urow);
        // hop 2: one warp per business of u, lanes stride its user list
        for (int k = warp; k < du; k += 8) {
            const unsigned long long brow = a.b_row[nu[k]];
            const int* lst = h3_list(a.b_adj, brow);
            for (int j = lane; j < h3_deg(brow); j += 32) {
                const int w = lst[j];
                atomicOr(&ubm[w >> 5], 1u << (w & 31));
            }
        }
        __syncthreads();
        if (tid == 0) ubm[u >> 5] &= ~(1u << (u & 31));   // distance 0, not 2
        __syncthreads();
        // hop 3: every thread walks words of the user bitmap; each set bit is a user w whose
        // (short) business list is marked
        for (int i = tid; i < a.uw; i += 256) {
            unsigned m = ubm[i];
            while (m) {
                const int w = i * 32 + __ffs(m) - 1;
                m &= m - 1;
                const unsigned long long wrow = a.u_row[w];
                const int* lst = h3_list(a.u_adj, wrow);
                for (int j = 0; j < h3_deg(wrow); ++j) {
                    const int b = lst[j];
                    atomicOr(&bbm[b >> 5], 1u << (b & 31));
                }
            }
        }
        __syncthreads();
        for (int k = tid; k < du; k += 256) {             // distance 1, not 3
            const int b = nu[k];
            atomicAnd(&bbm[b >> 5], ~(1u << (b & 31)));
        }
        __syncthreads();
        // count / ascending write: per-thread contiguous word ranges keep the order
        const int per = (a.bw + 255) / 256;
        const int w0 = min(a.bw, tid * per), w1 = min(a.bw, w0 + per);
        int mine = 0;
        for (int i = w0; i < w1; ++i) mine += __popc(bbm[i]);
        int inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(kAll, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) s_scan[warp] = inc;
        __syncthreads();
        int before = inc - mine, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            if (w < warp) before += s_scan[w];
            total += s_scan[w];
        }
        if (!FILL) {
            if (tid == 0) a.counts[it] = total;
        } else {
            int* dst = a.out_biz + a.offsets[it] + before;
            for (int i = w0; i < w1; ++i) {
                unsigned m = bbm[i];
                while (m) {
                    *dst++ = i * 32 + __ffs(m) - 1;
                    m &= m - 1;
                }
            }
        }
    }
}

int launch_hop3(blp_graph* g, Hop3Args a, bool fill, cudaStream_t st) {
    a.u_row = (const unsigned long long*)g->u_row;
    a.u_adj = g->u_adj;
    a.b_row = (const unsigned long long*)g->b_row;
    a.b_adj = g->b_adj;
    a.n_users = g->n_users;
    a.n_biz = g->n_biz;
    a.uw = (g->n_users + 31) / 32;
    a.bw = (g->n_biz + 31) / 32;
    const size_t smem = sizeof(unsigned) * ((size_t)a.uw + a.bw);
    if (smem + 1024 > (size_t)g->max_smem_optin) {
        set_error("blp_hop3: user + business bitmaps exceed one CTA's shared memory");
        return BLP_ERR_UNSUPPORTED;
    }
    int* counter = nullptr;
    BLP_CUDA_TRY(pool_alloc(g, (void**)&counter, sizeof(int), st));
    BLP_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(int), st));
    a.work_counter = counter;
    int per_sm = 0;
    if (fill) {
        BLP_CUDA_TRY(cudaFuncSetAttribute(k_hop3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BLP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_hop3<true>, 256, smem));
        k_hop3<true><<<std::max(1, per_sm) * g->sm_count, 256, smem, st>>>(a);
    } else {
        BLP_CUDA_TRY(cudaFuncSetAttribute(k_hop3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BLP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_hop3<false>, 256, smem));
        k_hop3<false><<<std::max(1, per_sm) * g->sm_count, 256, smem, st>>>(a);
    }
    BLP_CUDA_TRY(cudaGetLastError());
    cudaFreeAsync(counter, st);
    return BLP_OK;
}

}  // namespace
}  // namespace blp

extern "C" int blp_hop3_count(blp_graph* g, const int32_t* users, int64_t n, int64_t* counts,
                              void* stream) {
    if (!g || n < 0 || n >= INT_MAX || (n > 0 && (!users || !counts))) {
        blp::set_error("blp_hop3_count: bad argument");
        return BLP_ERR_INVALID;
    }
    if (n == 0) return BLP_OK;
    BLP_ON_DEVICE(g->device);
    blp::Hop3Args a{};
    a.users = users;
    a.n = n;
    a.counts = (long long*)counts;
    return blp::launch_hop3(g, a, false, (cudaStream_t)stream);
}

extern "C" int blp_hop3_fill(blp_graph* g, const int32_t* users, int64_t n, const int64_t* offsets,
                             int32_t* out_biz, void* stream) {
    if (!g || n < 0 || n >= INT_MAX || (n > 0 && (!users || !offsets || !out_biz))) {
        blp::set_error("blp_hop3_fill: bad argument");
        return BLP_ERR_INVALID;
    }
    if (n == 0) return BLP_OK;
    BLP_ON_DEVICE(g->device);
    blp::Hop3Args a{};
    a.users = users;
    a.n = n;
    a.offsets = (const long long*)offsets;
    a.out_biz = out_biz;
    return blp::launch_hop3(g, a, true, (cudaStream_t)stream);
}
