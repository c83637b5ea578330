// Evaluation on the device: what eval.run_evaluation computes from
// a score file (eval.py:17-32) -- per-user precision@k and one global ROC-AUC -- on the arrays the
// scoring path already holds in HBM.
//
//   precision@k (eval.py:21-24): per user, a STABLE descending sort by score, the mean label of
//       the first n = min(k, #pairs) entries.  One warp per user; the position of an element in
//       that order is #(larger scores) + #(equal scores in front of it), no sort needed.
//   ROC-AUC (eval.py:26, sklearn.metrics.roc_auc_score): P(s+ > s-) + P(s+ == s-)/2 over all
//       (positive, negative) pairs, from exact integer counts: the negatives' scores are radix
//       sorted once, every positive binary-searches its lower and upper bound.
#include <climits>
#include <cstdio>
#include <vector>

#include "blp_internal.h"

namespace blp {
namespace {

// SM count of the device the buffers live on (= the caller's current device for these entry points)
long long current_sm_count() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
        (void)cudaGetLastError();
        sms = 132;
    }
    return sms;
}

constexpr unsigned kEvalAll = 0xffffffffu;

// eval.py:21-24 sorts a user's pairs by score (descending, stable) and averages the labels of the
// first min(k, len).  An element's place in that order is #larger + #equal-in-front; "a before b"
// <=> score(a) > score(b) or (equal and index(a) < index(b)).  One warp per user.
//   len <= kPrecQuadratic : every element counts the elements before it (len^2 / 32 steps)
//   longer lists (hop-3 candidate sets run to thousands): the first n = min(k, len) places are
//   found by n rounds of a warp arg-best over the elements that come after the previous pick --
//   n * len / 32 steps instead of len^2 / 32
constexpr int kPrecQuadratic = 96;

__global__ void k_precision_at_k(const long long* __restrict__ off, const int* __restrict__ labels,
                                 const double* __restrict__ scores, long long n_groups, int k,
                                 double* __restrict__ prec) {
    const int lane = threadIdx.x & 31;
    long long g = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (; g < n_groups; g += n_warps) {
        const long long lo = off[g];
        const int len = (int)(off[g + 1] - lo);
        const int n = min(k, len);
        int hits = 0;
        if (len <= kPrecQuadratic) {
            for (int i = lane; i < len; i += 32) {
                const double pi = scores[lo + i];
                int pos = 0;
                for (int j = 0; j < len; ++j) {
                    const double pj = scores[lo + j];
                    pos += (pj > pi) || (pj == pi && j < i);
                }
                if (pos < n) hits += labels[lo + i];
            }
        } else {
            double prev_s = 0.0;
            int prev_i = -1;
            for (int r = 0; r < n; ++r) {
                double best_s = 0.0;
                int best_i = -1;
                for (int i = lane; i < len; i += 32) {
                    const double s = scores[lo + i];
                    const bool after = r == 0 || s < prev_s || (s == prev_s && i > prev_i);
                    if (after && (best_i < 0 || s > best_s || (s == best_s && i < best_i))) {
                        best_s = s;
                        best_i = i;
                    }
                }
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {   // butterfly: every lane ends with the same pick
                    const double os = __shfl_xor_sync(kEvalAll, best_s, d);
                    const int oi = __shfl_xor_sync(kEvalAll, best_i, d);
                    if (oi >= 0 && (best_i < 0 || os > best_s || (os == best_s && oi < best_i))) {
                        best_s = os;
                        best_i = oi;
                    }
                }
                prev_s = best_s;
                prev_i = best_i;
                if (lane == 0 && best_i >= 0) hits += labels[lo + best_i];
            }
        }
        hits = __reduce_add_sync(kEvalAll, hits);
        if (lane == 0) prec[g] = n > 0 ? (double)hits / (double)n : 0.0;
    }
}

// order-preserving map of IEEE doubles onto unsigned 64-bit integers
__device__ __forceinline__ unsigned long long sortable(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void k_split_by_label(const int* __restrict__ labels, const double* __restrict__ scores,
                                 long long n, unsigned long long* __restrict__ pos_keys,
                                 unsigned long long* __restrict__ neg_keys,
                                 unsigned long long* __restrict__ counters) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const unsigned long long key = sortable(scores[i] + 0.0);   // -0.0 == 0.0, as in Python
        if (labels[i] != 0)
            pos_keys[atomicAdd(&counters[0], 1ull)] = key;
        else
            neg_keys[atomicAdd(&counters[1], 1ull)] = key;
    }
}

__global__ void k_auc_counts(const unsigned long long* __restrict__ pos_keys, long long n_pos,
                             const unsigned long long* __restrict__ neg_sorted, long long n_neg,
                             unsigned long long* __restrict__ counters) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned long long less = 0, equal = 0;
    for (; i < n_pos; i += stride) {
        const unsigned long long key = pos_keys[i];
        long long lo = 0, hi = n_neg;           // lower bound: first negative >= key
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (neg_sorted[mid] < key) lo = mid + 1; else hi = mid;
        }
        const long long lb = lo;
        hi = n_neg;                              // upper bound: first negative > key
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (neg_sorted[mid] <= key) lo = mid + 1; else hi = mid;
        }
        less += (unsigned long long)lb;
        equal += (unsigned long long)(lo - lb);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        less += __shfl_xor_sync(kEvalAll, less, d);
        equal += __shfl_xor_sync(kEvalAll, equal, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (less) atomicAdd(&counters[2], less);      // integer sums: order independent
        if (equal) atomicAdd(&counters[3], equal);
    }
}

}  // namespace
}  // namespace blp

extern "C" int blp_eval_precision_at_k(const int64_t* offsets, const int32_t* labels,
                                       const double* scores, int64_t n_groups, int32_t k,
                                       double* precision_out, void* stream) {
    if (n_groups < 0 || k <= 0 || (n_groups > 0 && (!offsets || !labels || !scores || !precision_out))) {
        blp::set_error("blp_eval_precision_at_k: bad argument");
        return BLP_ERR_INVALID;
    }
    if (n_groups == 0) return BLP_OK;
    const long long sms = blp::current_sm_count();
    const int blocks = (int)std::min<long long>((n_groups * 32 + 255) / 256, sms * 16);
    blp::k_precision_at_k<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        (const long long*)offsets, labels, scores, n_groups, k, precision_out);
    BLP_CUDA_TRY(cudaGetLastError());
    return BLP_OK;
}

extern "C" int blp_eval_roc_auc(const int32_t* labels, const double* scores, int64_t n,
                                uint64_t* counts4_host, void* stream) {
    using namespace blp;
    if (n <= 0 || !labels || !scores || !counts4_host) {
        set_error("blp_eval_roc_auc: bad argument");
        return BLP_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *pos = nullptr, *neg = nullptr, *tmp = nullptr, *counters = nullptr;
    int rc = BLP_OK;
    auto done = [&](int code) {
        if (pos) cudaFreeAsync(pos, st);
        if (neg) cudaFreeAsync(neg, st);
        if (tmp) cudaFreeAsync(tmp, st);
        if (counters) cudaFreeAsync(counters, st);
        return code;
    };
#define BLP_TRY_E(expr)                                                              \
    do {                                                                             \
        cudaError_t e__ = (expr);                                                    \
        if (e__ != cudaSuccess) return done(cuda_fail(e__, #expr, __FILE__, __LINE__)); \
    } while (0)
    BLP_TRY_E(scratch_alloc((void**)&pos, sizeof(unsigned long long) * (size_t)n, st));
    BLP_TRY_E(scratch_alloc((void**)&neg, sizeof(unsigned long long) * (size_t)n, st));
    BLP_TRY_E(scratch_alloc((void**)&tmp, sizeof(unsigned long long) * (size_t)n, st));
    BLP_TRY_E(scratch_alloc((void**)&counters, sizeof(unsigned long long) * 4, st));
    BLP_TRY_E(cudaMemsetAsync(counters, 0, sizeof(unsigned long long) * 4, st));
    const long long sms = current_sm_count();
    const int blocks = (int)std::min<long long>((n + 255) / 256, sms * 16);
    k_split_by_label<<<blocks, 256, 0, st>>>(labels, scores, n, pos, neg, counters);
    unsigned long long h[4] = {0, 0, 0, 0};
    BLP_TRY_E(cudaMemcpyAsync(h, counters, sizeof(unsigned long long) * 2, cudaMemcpyDeviceToHost, st));
    BLP_TRY_E(cudaStreamSynchronize(st));
    if (h[0] > 0 && h[1] > 0) {
        std::vector<int> shifts;
        for (int s = 0; s < 64; s += 8) shifts.push_back(s);
        unsigned long long* sorted = neg;
        rc = radix_sort_u64(neg, tmp, (long long)h[1], shifts, st, &sorted);
        if (rc != BLP_OK) return done(rc);
        const int b2 = (int)std::min<long long>(((long long)h[0] + 255) / 256, sms * 16);
        k_auc_counts<<<b2, 256, 0, st>>>(pos, (long long)h[0], sorted, (long long)h[1], counters);
        BLP_TRY_E(cudaGetLastError());
    }
    BLP_TRY_E(cudaMemcpyAsync(h, counters, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, st));
    BLP_TRY_E(cudaStreamSynchronize(st));
#undef BLP_TRY_E
    for (int i = 0; i < 4; ++i) counts4_host[i] = h[i];
    return done(BLP_OK);
}
