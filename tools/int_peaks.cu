// Register-only microbenchmark of the integer pipes that bound the scBONITA2 kernels on B200
// (SURVEY 8-d): LOP3 (ALU pipe), POPC (XU pipe), IMAD (FMA pipe) and a LOP3+POPC mix.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_peaks tools/int_peaks.cu && ./int_peaks
// Prints warp-instructions per clock per SM (clock64 inside the kernel), the SM clock the kernel
// actually ran at (clock64 cycles over %globaltimer nanoseconds of the same interval) and Gops/s
// (lane ops, from CUDA events) with every SM busy.  The three agree:
//   gops ~= warp_inst_per_clk_per_sm * 32 * sms * sm_mhz_in_kernel / 1e3
// (the event time also holds the launch ramp and tail, so gops is a few percent lower).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ILP 8
#define ITERS 32768

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) pipe_kernel(uint32_t* out, long long* cycles, long long* nanos, uint32_t seed) {
    uint32_t x[ILP], y = seed ^ threadIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = seed * (i + 1) + threadIdx.x;
    __syncthreads();
    const unsigned long long n0 = globaltimer_ns();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) {  // LOP3 chain
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y), "r"(seed));
            } else if (MODE == 1) {  // POPC chain
                asm volatile("popc.b32 %0, %0;" : "+r"(x[i]));
            } else if (MODE == 2) {  // IMAD chain
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y), "r"(seed));
            } else if (MODE == 3) {  // 4 LOP3 : 1 POPC (the rule_counts mix)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(x[i]) : "r"(y), "r"(seed));
                if (i % 2 == 0) {
                    uint32_t p;
                    asm volatile("popc.b32 %0, %1;" : "=r"(p) : "r"(x[i]));
                    y += p;
                }
            } else if (MODE == 4) {  // IADD3 chain
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y));
            }
        }
    }
    long long t1 = clock64();
    const unsigned long long n1 = globaltimer_ns();
    uint32_t acc = y;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc ^= x[i];
    if (acc == 0x12345678u) out[0] = acc;
    if (threadIdx.x == 0) {
        cycles[blockIdx.x] = t1 - t0;
        nanos[blockIdx.x] = (long long)(n1 - n0);
    }
}

template <int MODE>
static void run(const char* name, double inst_per_iter, int sms) {
    uint32_t* out;
    long long *cyc, *ns;
    int blocks = sms;  // one 32-warp CTA per SM (8 warps x ILP 8 per scheduler): every CTA runs for the whole kernel
    cudaMalloc(&out, 4);
    cudaMalloc(&cyc, blocks * sizeof(long long));
    cudaMalloc(&ns, blocks * sizeof(long long));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    pipe_kernel<MODE><<<blocks, 1024>>>(out, cyc, ns, 12345u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    pipe_kernel<MODE><<<blocks, 1024>>>(out, cyc, ns, 12345u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[1024], hn[1024];
    cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(hn, ns, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0, mean_ns = 0;
    for (int i = 0; i < blocks; ++i) {
        mean += (double)h[i];
        mean_ns += (double)hn[i];
    }
    mean /= blocks;
    mean_ns /= blocks;
    // warp-instructions per SM = 32 warps * ITERS * inst
    double warp_inst_per_sm = 32.0 * ITERS * inst_per_iter;
    double lane_ops = (double)blocks * 1024.0 * ITERS * inst_per_iter;
    const double sm_mhz = mean / mean_ns * 1e3;
    printf("{\"pipe\": \"%s\", \"warp_inst_per_clk_per_sm\": %.3f, \"lanes_per_clk_per_sm\": %.1f, \"sm_mhz_in_kernel\": %.0f, "
           "\"gops_from_clocks\": %.1f, \"gops\": %.1f, \"ms\": %.4f}\n",
           name, warp_inst_per_sm / mean, 32.0 * warp_inst_per_sm / mean, sm_mhz,
           32.0 * warp_inst_per_sm / mean * sms * sm_mhz / 1e3, lane_ops / (ms * 1e-3) / 1e9, ms);
    cudaFree(out);
    cudaFree(cyc);
    cudaFree(ns);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz_max\": %.0f}\n", p.name, sms, p.clockRate / 1000.0);
    run<0>("lop3", ILP, sms);
    run<1>("popc", ILP, sms);
    run<2>("imad", ILP, sms);
    run<4>("iadd", ILP, sms);
    run<3>("lop3x4+popc", ILP * 2 + ILP / 2, sms);
    return 0;
}
