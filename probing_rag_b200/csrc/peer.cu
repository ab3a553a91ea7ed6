// Device memory that the other GPUs of the box can address (CUDA IPC over NVLink / NVSwitch): the arrays of per-query
// score bounds that doc-range shards raise in each other's memory while they score (pr_index_set_peer_thetas,
// bm25_tables.cuh raise_theta).  One process per GPU: a rank allocates its array here, publishes the 64-byte handle
// (torch.distributed all_gather_object in probing_rag_b200/sharding.py) and opens every other rank's.  These are the
// only device allocations the library makes, once per sharded retriever, never on the query path.
#include <string.h>

#include "common.cuh"

static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(pr_ipc_handle_t), "pr_ipc_handle_t must hold a cudaIpcMemHandle_t");

extern "C" int pr_peer_alloc(int device, size_t bytes, void **out_dev, pr_ipc_handle_t *out_handle)
{
    if (!out_dev || !out_handle || bytes == 0) {
        pr_set_error("pr_peer_alloc: bad argument");
        return PR_EINVAL;
    }
    PR_CUDA_CHECK(cudaSetDevice(device));
    void *p = nullptr;
    PR_CUDA_CHECK(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        pr_set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return PR_ECUDA;
    }
    memcpy(out_handle, &h, sizeof(h));
    *out_dev = p;
    return PR_OK;
}

extern "C" int pr_peer_open(int device, const pr_ipc_handle_t *handle, void **out_dev)
{
    if (!handle || !out_dev) {
        pr_set_error("pr_peer_open: null argument");
        return PR_EINVAL;
    }
    PR_CUDA_CHECK(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    PR_CUDA_CHECK(cudaIpcOpenMemHandle(out_dev, h, cudaIpcMemLazyEnablePeerAccess));
    return PR_OK;
}

extern "C" int pr_peer_close(void *dev)
{
    if (dev) PR_CUDA_CHECK(cudaIpcCloseMemHandle(dev));
    return PR_OK;
}

extern "C" int pr_peer_free(void *dev)
{
    if (dev) PR_CUDA_CHECK(cudaFree(dev));
    return PR_OK;
}
