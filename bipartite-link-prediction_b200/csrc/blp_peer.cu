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
#include <cstring>

#include "blp_internal.h"

static_assert(sizeof(cudaIpcMemHandle_t) == BLP_IPC_HANDLE_BYTES, "handle size is part of the ABI");

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
