// Same question as pdl_overlap.cu with REGISTER pressure: A = 800 threads x ~72 registers (cannot share an SM with
// B = 1024 threads x 64 registers), shared memory small on both sides.  Does A(k+1) run on the SMs B(k) leaves free?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long now_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
template <int R>
__device__ __forceinline__ void spin(int us, float* sink) {
    float acc[R];
#pragma unroll
    for (int i = 0; i < R; ++i) acc[i] = threadIdx.x * 0.5f + i;
    const unsigned long long t0 = now_ns();
    while (now_ns() - t0 < (unsigned long long)us * 1000ull) {
#pragma unroll
        for (int i = 0; i < R; ++i) acc[i] = acc[i] * 1.0001f + acc[(i + 1) % R];
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < R; ++i) s += acc[i];
    if (sink && s == 12345.678f) *sink = s;
}
template <int R>
__global__ void __launch_bounds__(1024, 1) kB(int us, float* sink) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    spin<R>(us, sink);
}
template <int R>
__global__ void __launch_bounds__(800, 1) kA(int us, float* sink) {
    spin<R>(us, sink);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename K>
static void launch(K k, int grid, int block, size_t smem, int us, cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    float* sink = nullptr;
    cudaLaunchKernelEx(&cfg, k, us, sink);
}
template <int RA, int RB>
static void run(cudaStream_t s, int blockB, int gridB, int blockA = 800, size_t smemA = 0, size_t smemB = 0) {
    const int ta = 13, tb = 20, K = 20;
    cudaFuncAttributes fa, fb;
    cudaFuncSetAttribute(kA<RA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(kB<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(kA<RA>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(kB<RB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncGetAttributes(&fa, kA<RA>); cudaFuncGetAttributes(&fb, kB<RB>);
    auto body = [&]() { for (int k = 0; k < K; ++k) { launch(kA<RA>, 148, blockA, smemA, ta, s); launch(kB<RB>, gridB, blockB, smemB, tb, s); } };
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    body(); cudaStreamSynchronize(s);
    cudaEventRecord(e0, s); body(); cudaEventRecord(e1, s); cudaStreamSynchronize(s);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("A %d thr x %d regs %zu KB, B %d thr x %d regs %zu KB, gridB=%d: %.1f us per pair (ta=%d tb=%d) err=%s\n", blockA, fa.numRegs, smemA / 1024, blockB, fb.numRegs, smemB / 1024, gridB, ms * 1e3 / K, ta, tb, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    cudaStream_t s; cudaStreamCreate(&s);
    run<60, 50>(s, 256, 148, 544, 142 * 1024, 63 * 1024); run<60, 50>(s, 256, 148, 544, 0, 0); run<60, 50>(s, 1024, 128, 544, 142 * 1024, 63 * 1024); run<60, 50>(s, 384, 148, 672, 142 * 1024, 63 * 1024); run<60, 50>(s, 256, 148, 672, 142 * 1024, 63 * 1024); run<60, 50>(s, 256, 148, 672, 150 * 1024, 40 * 1024);
    return 0;
}
