// Shared helpers for the scBONITA2 B200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "scbonita_b200.h"

namespace scb {

// thread-local error text behind scb_last_error()
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t err, const char* what);
int sm_count();

#define SCB_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            scb::set_error(__VA_ARGS__);       \
            return SCB_ERR_INVALID_ARG;        \
        }                                      \
    } while (0)

#define SCB_CUDA(call)                                      \
    do {                                                    \
        cudaError_t _e = (call);                            \
        if (_e != cudaSuccess) return scb::cuda_fail(_e, #call); \
    } while (0)

#define SCB_LAUNCH_CHECK(name)                              \
    do {                                                    \
        cudaError_t _e = cudaGetLastError();                \
        if (_e != cudaSuccess) return scb::cuda_fail(_e, name); \
    } while (0)

static inline cudaStream_t as_stream(scb_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// lop3.b32 with a compile-time truth table: bit (a<<2|b<<1|c) of IMM
template <int IMM>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(IMM));
    return r;
}

// Runtime truth table on 32 cells at once.  `m` holds the 8 table bits expanded to full-word
// masks (m[k] = table bit k ? ~0 : 0).  Shannon expansion on C, then B, then A: 7 LOP3.
struct LutMasks {
    uint32_t m[8];
};

__host__ __device__ __forceinline__ LutMasks expand_lut(uint32_t lut) {
    LutMasks r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.m[k] = 0u - ((lut >> k) & 1u);
    return r;
}

__device__ __forceinline__ uint32_t eval_lut(const LutMasks& t, uint32_t a, uint32_t b, uint32_t c) {
    // mux(sel, x1, x0) = (sel & x1) | (~sel & x0)  == lop3 0xCA with (sel, x1, x0)
    uint32_t g0 = lop3<0xCA>(c, t.m[1], t.m[0]);  // A=0,B=0
    uint32_t g1 = lop3<0xCA>(c, t.m[3], t.m[2]);  // A=0,B=1
    uint32_t g2 = lop3<0xCA>(c, t.m[5], t.m[4]);  // A=1,B=0
    uint32_t g3 = lop3<0xCA>(c, t.m[7], t.m[6]);  // A=1,B=1
    uint32_t h0 = lop3<0xCA>(b, g1, g0);
    uint32_t h1 = lop3<0xCA>(b, g3, g2);
    return lop3<0xCA>(a, h1, h0);
}

__device__ __forceinline__ uint32_t tail_mask(int64_t word, int64_t n_cells) {
    // mask of valid cell bits in `word` (0 if the word is entirely past the end)
    int64_t first = word * 32;
    if (first >= n_cells) return 0u;
    int64_t left = n_cells - first;
    return left >= 32 ? 0xFFFFFFFFu : ((1u << (int)left) - 1u);
}

}  // namespace scb
