// Does the pre-wait part of a programmatic-dependent kernel A(k+1) run beside its primary B(k)?
// B: few CTAs, griddepcontrol.wait; launch_dependents; then spins tb us.  A: 148 CTAs, spins ta us, then
// launch_dependents; wait.  Chain B A B A ... timed per pair; with overlap the pair costs max(ta, tb), else ta + tb.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long now_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void kB(int us, int* sink) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const unsigned long long t0 = now_ns();
    while (now_ns() - t0 < (unsigned long long)us * 1000ull) {}
    if (sink && threadIdx.x == 9999) *sink = 1;
}
__global__ void kA(int us, int* sink) {
    extern __shared__ int sm[];
    const unsigned long long t0 = now_ns();
    while (now_ns() - t0 < (unsigned long long)us * 1000ull) {}
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (sink && threadIdx.x == 9999) *sink = sm[0];
}
static void launch(void (*k)(int, int*), int grid, int block, size_t smem, int us, cudaStream_t s, bool pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    int* sink = nullptr;
    cudaLaunchKernelEx(&cfg, k, us, sink);
}
int main() {
    cudaStream_t s; cudaStreamCreate(&s);
    const int ta = 13, tb = 20, K = 20;
    for (int kbA : {113, 114, 132, 164, 196, 200, 210, 220, 224, 226, 227}) for (int gridB : {8, 128}) {
        const size_t smemA = (size_t)kbA * 1024, smemB = 60 * 1024;
        cudaFuncSetAttribute(kA, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(kB, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(kA, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(kB, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        auto body = [&]() { for (int k = 0; k < K; ++k) { launch(kA, 148, 800, smemA, ta, s, true); launch(kB, gridB, 1024, smemB, tb, s, true); } };
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float ms = 0;
        body(); cudaStreamSynchronize(s);
        cudaEventRecord(e0, s); body(); cudaEventRecord(e1, s); cudaStreamSynchronize(s);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("smemA=%zuKB smemB=%zuKB gridB=%d: %.1f us per pair (ta=%d tb=%d) err=%s\n", smemA / 1024, smemB / 1024, gridB, ms * 1e3 / K, ta, tb, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
