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

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: the high-water mark a call
// site keeps (so the enqueue path stays free of non-stream API calls, e.g. under graph capture) is
// kept per device, or a thread that has served cuda:0 would skip the call on cuda:1.
struct SmemMarks {
    size_t mark[64];  // zero-initialised (static thread_local at the call site); 0 = the 48 KB default
};
static inline int ensure_dynamic_smem(const void* kernel, size_t smem, SmemMarks& marks) {
    int dev = 0;
    SCB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) {  // no cache slot: set it every time
        SCB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        return SCB_OK;
    }
    const size_t have = marks.mark[dev] ? marks.mark[dev] : (size_t)48 * 1024;
    if (smem > have) {
        SCB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        marks.mark[dev] = smem;
    }
    return SCB_OK;
}

// Shared-memory carve-out at its maximum for a kernel that wants several large CTAs per SM (per device,
// once per call site like ensure_dynamic_smem).
static inline int ensure_max_carveout(const void* kernel, SmemMarks& marks) {
    int dev = 0;
    SCB_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && marks.mark[dev]) return SCB_OK;
    SCB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    if (dev >= 0 && dev < 64) marks.mark[dev] = 1;
    return SCB_OK;
}

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

// ---- rule-counts plan (built by scb_counts.cu, read by its kernels and by the fitness kernel) -----
constexpr int kTermsPerGroup = 11;  // multi-variable AND terms
constexpr int kTileWords = 32;      // words per row per tile (128 B)
// Built once per (groups, n_rows) by a one-CTA kernel, read by every pass:
//   header   int32[4]     {n_groups, n_rows, n_orphans, 0}
//   desc     uint4[G]     groups sorted by arity (widest first): byte offsets of rows A, B, C, T inside a
//                         stage; .x carries arity (bits 0-1) and "owns its target row" (bit 2)
//   order    uint16[G]    original group index of sorted position
//   orphans  uint16[R]    rows that are no group's target (their popcounts need items of their own)
//   target   int32[G], parents int32[3G]   copies for the final inversion
//   owner    uint32[R]    build-time temporary
struct PlanLayout {
    size_t desc, order, orphans, target, parents, owner, bytes;
};
__host__ __device__ inline size_t pad16(size_t x) { return (x + 15) & ~(size_t)15; }
__host__ __device__ inline PlanLayout plan_layout(int n_rows, int n_groups) {
    PlanLayout L;
    L.desc = 16;
    L.order = L.desc + (size_t)n_groups * 16;
    L.orphans = L.order + pad16((size_t)n_groups * 2);
    L.target = L.orphans + pad16((size_t)n_rows * 2);
    L.parents = L.target + pad16((size_t)n_groups * 4);
    L.owner = L.parents + pad16((size_t)n_groups * 12);
    L.bytes = L.owner + pad16((size_t)n_rows * 4);
    return L;
}
constexpr uint32_t kDescOwn = 4u;
// rows of a stage as the TMA boxes lay them out (a box holds at most 256 rows); the row after them is
// the all-zero row that unbound parent slots read
__host__ __device__ inline int plan_zero_row(int n_rows) {
    const int n_boxes = (n_rows + 255) / 256;
    return n_boxes * ((n_rows + n_boxes - 1) / n_boxes);
}

// Subset sum q (bit 3 = T, bit 2 = A, bit 1 = B, bit 0 = C; q = 0: every cell) of the group at sorted
// position `pos`, gathered from the accumulators acc[n_rows + 11 G] that the counting kernels fill:
// single-variable sums are row popcounts (rows from the descriptor; the all-zero row reads 0), the
// rest are the group's own AND terms.  The 16 values of a group, Moebius-inverted, are its joint counts.
// (Branch-free up to the one load: the sixteen lanes of a group take different cases, and a load inside
// each case would be issued -- and waited for -- case by case.)
__device__ __forceinline__ int subset_sum_index(const uint4 d, int g, int q, int n_rows, int zero_row) {
    // term index of the multi-variable masks (SCB_TERM_MASKS order), one nibble per mask; 15 = single row / empty
    constexpr unsigned long long kTermOfMask = 0xA784956F301F2FFFull;
    const int term = (int)((kTermOfMask >> (4 * q)) & 15ull);
    const uint32_t off = q == 8 ? d.w : (q == 4 ? d.x : (q == 2 ? d.y : d.z));
    const int row = (int)(off >> 7);
    const int idx = term != 15 ? n_rows + g * kTermsPerGroup + term : row;
    return (q == 0 || (term == 15 && row == zero_row)) ? -1 : idx;   // -1: no load (every cell / the all-zero row)
}
__device__ __forceinline__ long long subset_sum_local(const uint4 d, int g, int q, const unsigned long long* acc,
                                                      int n_rows, int zero_row, long long n_cells) {
    const int idx = subset_sum_index(d, g, q, n_rows, zero_row);
    const long long v = idx >= 0 ? (long long)__ldcg(acc + idx) : 0ll;
    return q == 0 ? n_cells : v;
}
// Moebius inversion across the 16 lanes that hold the subset sums of one group (lane & 15 = q):
// afterwards lane q holds #cells whose set of ones is exactly q (target = q >> 3, minterm = q & 7).
__device__ __forceinline__ long long moebius16(long long x, int q) {
#pragma unroll
    for (int bit = 1; bit < 16; bit <<= 1) {
        const long long other = __shfl_xor_sync(0xFFFFFFFFu, x, bit);
        if (!(q & bit)) x -= other;
    }
    return x;
}

// ---- peer exchange region (scb_peer.cu, rule_counts_kernel, population_fitness_kernel) --------
// One region per rank, mapped into every peer (CUDA IPC over NVLink).  Rank r stores its joint
// counts into slot [parity][r] of EVERY rank's region as 64-bit words {tag = epoch + 1 : count}
// (a count is < 2^31: at most 2^31-1 cells per shard).  An aligned 8-byte store is atomic, so the
// tag validates the word it travels with: no fence, no flag, one NVLink traversal (the "LL"
// protocol of collective libraries).  The reader takes the epoch from its OWN slot (written
// earlier in its stream) and spins on a word until its tag matches.  parity = epoch & 1
// double-buffers the slots; a rank can run at most one step ahead of its peers.
constexpr int kPeerSlotsOffset = 256;
struct PeerHeader {
    unsigned long long push_ticket;  // CTAs of every push draw tickets: epoch = ticket / n_ranks
    unsigned int error;              // raised by a kernel that gave up waiting for a peer (it also poisons its results)
    unsigned int timeout_ms;         // how long such a kernel waits (0 = 30 s); scb_peer_set_timeout
};
static_assert(sizeof(PeerHeader) <= kPeerSlotsOffset, "peer header must fit before the slots");

__device__ __forceinline__ unsigned long long peer_word(unsigned int tag, long long count) {
    return ((unsigned long long)tag << 32) | (unsigned long long)(unsigned int)count;
}
__device__ __forceinline__ void st_peer(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ ulonglong2 ld_peer2(const unsigned long long* p) {
    ulonglong2 v;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_peer(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ uint32_t tail_mask(int64_t word, int64_t n_cells) {
    // mask of valid cell bits in `word` (0 if the word is entirely past the end)
    int64_t first = word * 32;
    if (first >= n_cells) return 0u;
    int64_t left = n_cells - first;
    return left >= 32 ? 0xFFFFFFFFu : ((1u << (int)left) - 1u);
}

}  // namespace scb
