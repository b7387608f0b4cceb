// Peer-memory exchange of the joint counts between the GPUs of one box (cells shard across GPUs,
// SURVEY 8-e).  The reference has no counterpart: it is single-host multiprocessing
// (importance_scores.py:238-253).  See PeerHeader in scb_common.cuh for the protocol.
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include "scb_common.cuh"

namespace scb {

// CTA d stores this rank's counts into its slot of rank d's region as tagged words
__global__ void __launch_bounds__(512)
peer_push_kernel(const int64_t* __restrict__ counts, int64_t n_values, unsigned char* local_region,
                 void* const* __restrict__ peer_regions, int rank, int n_ranks) {
    // the counts come from the kernel before this one in the stream
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __shared__ unsigned long long s_epoch;
    if (threadIdx.x == 0) {
        PeerHeader* hdr = reinterpret_cast<PeerHeader*>(local_region);
        s_epoch = atomicAdd(&hdr->push_ticket, 1ull) / (unsigned long long)n_ranks;
    }
    __syncthreads();
    const unsigned int epoch = (unsigned int)s_epoch;
    const int parity = (int)(epoch & 1u);
    unsigned char* dst_region = static_cast<unsigned char*>(peer_regions[blockIdx.x]);
    unsigned long long* slot = reinterpret_cast<unsigned long long*>(dst_region + kPeerSlotsOffset) +
                               ((int64_t)parity * n_ranks + rank) * n_values;
    for (int64_t i = threadIdx.x; i < n_values; i += blockDim.x) st_peer(slot + i, peer_word(epoch + 1u, __ldcg(counts + i)));
}

}  // namespace scb

using namespace scb;

extern "C" int64_t scb_peer_region_bytes(int n_ranks, int64_t n_values) {
    if (n_ranks < 1 || n_ranks > SCB_MAX_PEERS || n_values < 0) return 0;
    return kPeerSlotsOffset + 2 * (int64_t)n_ranks * n_values * 8;
}

extern "C" int scb_peer_region_alloc(int64_t bytes, void** region, unsigned char* ipc_handle) {
    SCB_REQUIRE(bytes >= kPeerSlotsOffset && region && ipc_handle, "peer_region_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == SCB_IPC_HANDLE_BYTES, "IPC handle size");
    void* p = nullptr;
    SCB_CUDA(cudaMalloc(&p, (size_t)bytes));
    SCB_CUDA(cudaMemset(p, 0, (size_t)bytes));
    if (const char* e = getenv("SCB_PEER_TIMEOUT_MS")) {   // default for every region of this process (0 = 30 s)
        PeerHeader hdr = {};
        hdr.timeout_ms = (unsigned int)strtoul(e, nullptr, 10);
        SCB_CUDA(cudaMemcpy(p, &hdr, sizeof(hdr), cudaMemcpyHostToDevice));
    }
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return cuda_fail(e, "cudaIpcGetMemHandle");
    }
    memcpy(ipc_handle, &h, sizeof(h));
    SCB_CUDA(cudaDeviceSynchronize());
    *region = p;
    return SCB_OK;
}

extern "C" int scb_peer_region_open(const unsigned char* ipc_handle, void** region) {
    SCB_REQUIRE(ipc_handle && region, "peer_region_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    SCB_CUDA(cudaIpcOpenMemHandle(region, h, cudaIpcMemLazyEnablePeerAccess));
    return SCB_OK;
}

extern "C" int scb_peer_region_close(void* region) {
    if (region) SCB_CUDA(cudaIpcCloseMemHandle(region));
    return SCB_OK;
}

extern "C" int scb_peer_region_free(void* region) {
    if (region) SCB_CUDA(cudaFree(region));
    return SCB_OK;
}

extern "C" int scb_peer_push_counts(const int64_t* counts, int64_t n_values, void* local_region,
                                    void* const* peer_regions, int rank, int n_ranks, scb_stream_t stream) {
    SCB_REQUIRE(n_ranks >= 1 && n_ranks <= SCB_MAX_PEERS && rank >= 0 && rank < n_ranks, "peer_push_counts: rank %d of %d", rank, n_ranks);
    SCB_REQUIRE(n_values > 0 && n_values % 2 == 0, "peer_push_counts: n_values must be a positive even number");
    SCB_REQUIRE(counts && local_region && peer_regions, "peer_push_counts: null pointer");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n_ranks);
    cfg.blockDim = dim3(512);
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = getenv("SCB_NO_PDL") ? 0 : 1;
    SCB_CUDA(cudaLaunchKernelEx(&cfg, peer_push_kernel, counts, n_values, static_cast<unsigned char*>(local_region),
                                peer_regions, rank, n_ranks));
    return SCB_OK;
}

extern "C" int scb_peer_error(const void* local_region, int* out_error) {
    SCB_REQUIRE(local_region && out_error, "peer_error: null pointer");
    SCB_CUDA(cudaDeviceSynchronize());
    PeerHeader h;
    SCB_CUDA(cudaMemcpy(&h, local_region, sizeof(h), cudaMemcpyDeviceToHost));
    *out_error = (int)h.error;
    if (h.error) {  // reported once: the flag is cleared so that the next check speaks about the next steps
        const unsigned int zero = 0u;
        SCB_CUDA(cudaMemcpy(static_cast<unsigned char*>(const_cast<void*>(local_region)) + offsetof(PeerHeader, error), &zero,
                            sizeof(zero), cudaMemcpyHostToDevice));
    }
    return SCB_OK;
}

extern "C" int scb_peer_set_timeout(void* local_region, unsigned int timeout_ms) {
    SCB_REQUIRE(local_region, "peer_set_timeout: null pointer");
    SCB_CUDA(cudaMemcpy(static_cast<unsigned char*>(local_region) + offsetof(PeerHeader, timeout_ms), &timeout_ms,
                        sizeof(timeout_ms), cudaMemcpyHostToDevice));
    return SCB_OK;
}
