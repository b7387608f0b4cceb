// scb_rule_counts: joint (A,B,C,target) counts per node in ONE pass over the cell bitplanes.
//
// Replaces the inner loop of RuleDetermination.calculate_error (rule_determination.py:509-582).
// For a rule with truth table L over (A,B,C) the reference's mismatch count is
//     sum_k #{cells: (A,B,C)==k, target != L_k},
// so the 16 joint counts of a node are a sufficient statistic for EVERY candidate rule and
// every GA individual.  They are obtained as popcounts of the 15 non-empty AND-subsets of the
// four bitplane rows (4 single rows, 11 multi-variable terms) followed by a Moebius inversion.
//
// Kernel structure (B200): see rule_counts_kernel below -- one persistent CTA per SM, a ring of
// TMA-filled stages (cp.async.bulk.tensor.2d, full/empty mbarriers), 24 consumer warps drawing
// work items from a shared counter, Harley-Seal style carry-save popcounts, shared-memory
// atomics per CTA, global int64 atomics per launch, Moebius inversion in the last CTA.
#include <cuda.h>
#include <stdlib.h>

#include "scb_common.cuh"

namespace scb {

constexpr int kMaxStages = 8;
constexpr int kInFlight = 3;        // bulk copies outstanding per SM
#ifndef SCB_CONSUMER_WARPS
#define SCB_CONSUMER_WARPS 24
#endif
constexpr int kConsumerWarps = SCB_CONSUMER_WARPS;
constexpr int kCountThreads = (kConsumerWarps + 1) * 32;  // + the producer warp
constexpr int kTicketBytes = 16;
#ifndef SCB_DIRECT_RED
#define SCB_DIRECT_RED 0
#endif
// 0 (default): partial counts are kept per CTA in shared memory and flushed once at the end -- ~190 k
// global atomics in one burst, which the last CTA's finalize waits out in L2 (~3.5 k cycles);
// 1 (experiment, measured: 41 us instead of 22.5 us on H1M, 572 us instead of 143 us at 16M cells):
// every partial count goes straight to the global accumulators with a 32-bit RED.ADD -- the 2 400 hot
// addresses sustain only ~18 atomics per clock chip-wide, so spreading the atomics loses.
constexpr bool kDirectRed = SCB_DIRECT_RED != 0;
constexpr int kMaxGroupsPerLaunch = 1024;  // groups are processed in slices so the per-CTA accumulators fit in shared memory

// term order (bit3=T, bit2=A, bit1=B, bit0=C): masks of the 11 multi-variable subsets
#define SCB_TERM_MASKS {0x6 /*AB*/, 0x5 /*AC*/, 0x3 /*BC*/, 0x7 /*ABC*/, 0xC /*AT*/, 0xA /*BT*/, \
                        0x9 /*CT*/, 0xE /*ABT*/, 0xD /*ACT*/, 0xB /*BCT*/, 0xF /*ABCT*/}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// TMA 2-D tiled copy: one instruction moves a whole [box_rows x 32 words] box; rows / words
// outside the tensor are zero-filled by the hardware and still counted in the byte total.
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* map, int x, int y,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst_smem)),
        "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}

#ifdef SCB_TRACE
__device__ int g_dbg;  // experiments: bit 0 = consumers skip the arithmetic, bit 1 = producer skips the copies
__device__ long long g_trace[256][12];
__device__ unsigned long long g_pdl_a[64][4];  // globaltimer of CTA 0: entry, loop exit, after griddepcontrol.wait
__device__ unsigned int g_pdl_a_n;
#define SCB_TRACE_AT(slot, cond) do { if (cond) g_trace[blockIdx.x & 255][slot] = clock64(); } while (0)
#else
#define SCB_TRACE_AT(slot, cond) do { } while (0)
#endif

// barrier among the consumer warps only (the producer warp never joins it)
__device__ __forceinline__ void consumer_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");
}

// carry-save adder: (sum, carry) of three words, 2 LOP3
__device__ __forceinline__ void csa(uint32_t& carry, uint32_t& sum, uint32_t a, uint32_t b, uint32_t c) {
    sum = lop3<0x96>(a, b, c);
    carry = lop3<0xE8>(a, b, c);
}

// popcount of eight words: four carry-save adders take seven of them down to one word of weight
// 1, one of weight 2 and one of weight 4; the eighth is counted as it is (8 LOP3 + 4 POPC; the
// weighted sum is finished arithmetically).  POPC issues on the quarter-rate XU pipe, LOP3 on the
// half-rate ALU pipe: with the eight ANDs that form a term this split keeps the two equally busy.
#ifndef SCB_POPC8_VARIANT
#define SCB_POPC8_VARIANT 0
#endif
__device__ __forceinline__ uint32_t popc8(const uint32_t (&x)[8]) {
    uint32_t o1, t1, o2, t2, o3, t3, tt, f1;
    csa(t1, o1, x[0], x[1], x[2]);
    csa(t2, o2, x[3], x[4], x[5]);
    csa(t3, o3, o1, o2, x[6]);
#if SCB_POPC8_VARIANT == 1
    // experiment: 6 LOP3 + 5 POPC (the three weight-2 words are counted as they are)
    return __popc(o3) + __popc(x[7]) + 2 * (__popc(t1) + __popc(t2) + __popc(t3));
#else
    csa(f1, tt, t1, t2, t3);
    return __popc(o3) + __popc(x[7]) + 2 * __popc(tt) + 4 * __popc(f1);
#endif
}

// number of parent slots a group binds when they are packed to the left (A, AB, ABC); any other
// pattern takes the general three-parent path
__host__ __device__ __forceinline__ int group_arity(int ra, int rb, int rc) {
    if (rc >= 0) return 3;
    if (rb >= 0) return ra >= 0 ? 2 : 3;
    return ra >= 0 ? 1 : 0;
}

// (the plan layout -- PlanLayout, plan_layout(), kDescOwn -- lives in scb_common.cuh: the fitness kernel reads it too)

__global__ void __launch_bounds__(256)
rule_counts_plan_kernel(int n_rows, int n_groups, int zero_row, const int32_t* __restrict__ target_row,
                        const int32_t* __restrict__ parent_row, unsigned char* __restrict__ plan) {
    const PlanLayout L = plan_layout(n_rows, n_groups);
    int* header = reinterpret_cast<int*>(plan);
    uint4* desc = reinterpret_cast<uint4*>(plan + L.desc);
    unsigned short* order = reinterpret_cast<unsigned short*>(plan + L.order);
    unsigned short* orphans = reinterpret_cast<unsigned short*>(plan + L.orphans);
    int* target = reinterpret_cast<int*>(plan + L.target);
    int* parents = reinterpret_cast<int*>(plan + L.parents);
    unsigned int* owner = reinterpret_cast<unsigned int*>(plan + L.owner);
    __shared__ int s_count[4], s_fill[4], s_orphans;
    const int tid = threadIdx.x;
    if (tid < 4) s_count[tid] = s_fill[tid] = 0;
    if (tid == 0) s_orphans = 0;
    for (int r = tid; r < n_rows; r += blockDim.x) owner[r] = 0xFFFFFFFFu;
    __syncthreads();
    for (int g = tid; g < n_groups; g += blockDim.x) {
        target[g] = target_row[g];
        for (int k = 0; k < 3; ++k) parents[3 * g + k] = parent_row[3 * g + k];
        atomicAdd(&s_count[group_arity(parent_row[3 * g], parent_row[3 * g + 1], parent_row[3 * g + 2])], 1);
    }
    __syncthreads();
    // counting sort by arity, widest first; the first sorted group that targets a row owns it
    for (int g = tid; g < n_groups; g += blockDim.x) {
        const int ra = parent_row[3 * g], rb = parent_row[3 * g + 1], rc = parent_row[3 * g + 2];
        const int ar = group_arity(ra, rb, rc);
        int pos = atomicAdd(&s_fill[ar], 1);
        for (int k = 3; k > ar; --k) pos += s_count[k];
        order[pos] = (unsigned short)g;
        const uint32_t row_bytes = kTileWords * 4;
        desc[pos] = make_uint4((uint32_t)(ra < 0 ? zero_row : ra) * row_bytes | (uint32_t)ar,
                               (uint32_t)(rb < 0 ? zero_row : rb) * row_bytes,
                               (uint32_t)(rc < 0 ? zero_row : rc) * row_bytes,
                               (uint32_t)target_row[g] * row_bytes);
        atomicMin(&owner[target_row[g]], (unsigned int)pos);
    }
    __syncthreads();
    for (int pos = tid; pos < n_groups; pos += blockDim.x) {
        const int rt = (int)(desc[pos].w / (kTileWords * 4));
        if (owner[rt] == (unsigned int)pos) desc[pos].x |= kDescOwn;
    }
    for (int r = tid; r < n_rows; r += blockDim.x)
        if (owner[r] == 0xFFFFFFFFu) orphans[atomicAdd(&s_orphans, 1)] = (unsigned short)r;
    __syncthreads();
    if (tid == 0) {
        header[0] = n_groups;
        header[1] = n_rows;
        header[2] = s_orphans;
        header[3] = 0;
    }
}

__device__ __forceinline__ void load8(uint32_t (&x)[8], const uint32_t* row, int c0) {
    // two 128-bit shared loads: 16-byte chunks c0 and c0^4 of the 128-byte row
    const uint4 lo = *reinterpret_cast<const uint4*>(row + 4 * c0);
    const uint4 hi = *reinterpret_cast<const uint4*>(row + 4 * (c0 ^ 4));
    x[0] = lo.x; x[1] = lo.y; x[2] = lo.z; x[3] = lo.w;
    x[4] = hi.x; x[5] = hi.y; x[6] = hi.z; x[7] = hi.w;
}

__device__ __forceinline__ uint32_t quad_sum(uint32_t v) {
    v += __shfl_xor_sync(0xFFFFFFFFu, v, 1);
    v += __shfl_xor_sync(0xFFFFFFFFu, v, 2);
    return v;
}

// The AND-subset popcounts of one group over one 32-word tile, taken by a quad (lane l holds the
// eight words of chunks c0, c0^4), folded across the quad and added to the CTA accumulators
// dst[0..10] (term order of SCB_TERM_MASKS); the popcount of the target row itself goes to *dst_t
// when the group owns that row.  ARITY selects which terms can be non-zero.  Per-lane counts are
// <= 256, quad sums <= 1024: two ride in one register as 16-bit fields.
template <int ARITY>
__device__ __forceinline__ void group_terms(const uint32_t* __restrict__ pa, const uint32_t* __restrict__ pb,
                                            const uint32_t* __restrict__ pc, const uint32_t* __restrict__ pt,
                                            int c0, int l, bool live, bool own, uint32_t* dst, uint32_t* dst_t) {
    uint32_t a[8], t[8], x[8];
    load8(a, pa, c0);
    load8(t, pt, c0);
#define SCB_TERM(var, expr)                                      \
    _Pragma("unroll") for (int j = 0; j < 8; ++j) x[j] = (expr); \
    uint32_t var = popc8(x);
    const uint32_t tt = popc8(t);
    if (ARITY <= 1) {
        SCB_TERM(at, a[j] & t[j])
        const uint32_t p = quad_sum(at | (tt << 16));
        if (live && l == 0 && ARITY == 1) atomicAdd(dst + 4, p & 0xFFFFu);
        if (live && l == 1 && own) atomicAdd(dst_t, p >> 16);
        return;
    }
    uint32_t b[8];
    load8(b, pb, c0);
    if (ARITY == 2) {
        SCB_TERM(ab, a[j] & b[j])
        SCB_TERM(at, a[j] & t[j])
        SCB_TERM(bt, b[j] & t[j])
        SCB_TERM(abt, lop3<0x80>(a[j], b[j], t[j]))
        const uint32_t p0 = quad_sum(ab | (at << 16)), p1 = quad_sum(bt | (abt << 16)), p2 = quad_sum(tt);
        if (live) {  // lanes 0..3 own AB, AT, BT, ABT
            const uint32_t v = ((l & 2) ? p1 : p0) >> ((l & 1) * 16) & 0xFFFFu;
            const int term = l == 0 ? 0 : (l == 1 ? 4 : (l == 2 ? 5 : 7));
            atomicAdd(dst + term, v);
            if (l == 0 && own) atomicAdd(dst_t, p2);
        }
        return;
    }
    uint32_t c[8];
    load8(c, pc, c0);
    uint32_t cnt[kTermsPerGroup + 1];
#define SCB_TERM_I(idx, expr)                                    \
    _Pragma("unroll") for (int j = 0; j < 8; ++j) x[j] = (expr); \
    cnt[idx] = popc8(x);
    SCB_TERM_I(0, a[j] & b[j])
    SCB_TERM_I(1, a[j] & c[j])
    SCB_TERM_I(2, b[j] & c[j])
    SCB_TERM_I(3, lop3<0x80>(a[j], b[j], c[j]))
    SCB_TERM_I(4, a[j] & t[j])
    SCB_TERM_I(5, b[j] & t[j])
    SCB_TERM_I(6, c[j] & t[j])
    SCB_TERM_I(7, lop3<0x80>(a[j], b[j], t[j]))
    SCB_TERM_I(8, lop3<0x80>(a[j], c[j], t[j]))
    SCB_TERM_I(9, lop3<0x80>(b[j], c[j], t[j]))
    SCB_TERM_I(10, lop3<0x80>(a[j], b[j], c[j]) & t[j])
    cnt[11] = tt;
#undef SCB_TERM_I
#undef SCB_TERM
    uint32_t pk[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) pk[k] = quad_sum(cnt[2 * k] | (cnt[2 * k + 1] << 16));
    // lane l of the quad owns terms l, l+4, l+8; "term 11" is the target row's own popcount
    const bool hi_pair = (l >> 1) != 0;
    const int shift = (l & 1) * 16;
    uint32_t v0 = ((hi_pair ? pk[1] : pk[0]) >> shift) & 0xFFFFu;
    uint32_t v1 = ((hi_pair ? pk[3] : pk[2]) >> shift) & 0xFFFFu;
    uint32_t v2 = ((hi_pair ? pk[5] : pk[4]) >> shift) & 0xFFFFu;
    if (live) {
        atomicAdd(dst + l, v0);
        atomicAdd(dst + l + 4, v1);
        if (l < 3) atomicAdd(dst + l + 8, v2);
        else if (own) atomicAdd(dst_t, v2);
    }
}

// entry i of the global accumulators: uint32 when the CTAs add to them directly (a count is < 2^31),
// uint64 when they flush their shared-memory sums
__device__ __forceinline__ long long load_acc(const unsigned long long* acc, size_t i) {
    if (kDirectRed) return (long long)__ldcg(reinterpret_cast<const unsigned int*>(acc) + i);
    return (long long)__ldcg(acc + i);
}

// Moebius inversion of the subset sums, sixteen lanes per group.  Lane q holds sub[q] = #cells
// where every variable of q is 1 (T=8, A=4, B=2, C=1): one L2 load each, four butterfly steps
// through shuffles, one coalesced store (and one tagged store per peer when the cells are sharded).
// Rows and the original group index come from the shared-memory tables.
// Only the LAST CTA needs the result, at the very end, where a cold instruction cache costs ~3 k
// cycles -- so every CTA runs the routine once with dry = true (loads and arithmetic, no stores)
// while it waits for its first tile: the code is resident when it matters.  __noinline__ keeps
// the two call sites on the same instructions.
__device__ __noinline__ void moebius_pass(const uint4* __restrict__ desc, const unsigned short* __restrict__ order,
                                          const unsigned long long* __restrict__ acc, int n_rows, int rows_alloc,
                                          int n_groups, int n_do, int64_t n_cells, int64_t* __restrict__ out_counts,
                                          void* const* __restrict__ peer_regions, int n_ranks, int64_t peer_slot,
                                          unsigned int peer_tag, int tid, bool dry) {
    const int q = tid & 15;
    // term index of the multi-variable masks (SCB_TERM_MASKS), one nibble per mask
    constexpr unsigned long long kTermOfMask = 0xA784956F301F2FFFull;
    const int term = (int)((kTermOfMask >> (4 * q)) & 15ull);
    constexpr int kBatch = 4;                       // loads of kBatch groups fly together
    constexpr int kPerRound = kCountThreads / 16;   // groups per round of the CTA
    long long sink = 0;
    for (int base = 0; base < n_do; base += kBatch * kPerRound) {  // uniform trip count
        const int pos0 = base + (tid >> 4);
        long long v[kBatch];
        int g[kBatch];
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
            const int pos = pos0 + k * kPerRound;
            const int p = pos < n_groups ? pos : n_groups - 1;
            const uint4 d = desc[p];
            g[k] = pos < n_groups ? (int)order[p] : -1;
            if (q == 0) v[k] = n_cells;
            else if (term != 15) v[k] = load_acc(acc, n_rows + (size_t)order[p] * kTermsPerGroup + term);
            else {
                const uint32_t off = q == 8 ? d.w : (q == 4 ? d.x : (q == 2 ? d.y : d.z));
                const int row = (int)(off >> 7);
                v[k] = row == rows_alloc ? 0ll : load_acc(acc, row);
            }
        }
#ifdef SCB_TRACE
        if (base == 0 && tid == 0) {
            g_trace[blockIdx.x & 255][10] = clock64();
            long long t = v[0] + v[1] + v[2] + v[3];
            if (t == 0x7FFFFFFFFFFFFFFFll) sink = 1;   // wait for the loads
            g_trace[blockIdx.x & 255][11] = clock64();
        }
#endif
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
            long long x = v[k];
#pragma unroll
            for (int bit = 1; bit < 16; bit <<= 1) {
                const long long other = __shfl_xor_sync(0xFFFFFFFFu, x, bit);
                if (!(q & bit)) x -= other;  // exact[S] = #cells whose set of ones is exactly S
            }
            sink ^= x;
            // q = target << 3 | minterm  ->  counts[g][2 * minterm + target]
            if (g[k] >= 0 && !dry) {
                const size_t at = (size_t)g[k] * 16 + (((q & 7) << 1) | (q >> 3));
                out_counts[at] = x;
                for (int d = 0; d < (peer_regions ? n_ranks : 0); ++d)
                    st_peer(reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(peer_regions[d]) +
                                                                  kPeerSlotsOffset) + peer_slot + at,
                            peer_word(peer_tag, x));
            }
        }
    }
    if (dry && sink == 0x7FFFFFFFFFFFFFFFll) out_counts[0] = sink;  // keeps the dry arithmetic alive; never true
}

// One persistent CTA per SM: 24 consumer warps + 1 producer warp around a ring of TMA stages.
//   producer (one lane): waits for a stage to be released, points the stage at the CTA's next
//     tile (static round robin over the grid) and starts the 2-D bulk copy; full[stage] counts
//     the bytes;
//   consumers: every warp walks the tiles in ring order; inside a tile the work is cut into
//     ITEMS -- eight groups of one arity class (one group per quad; the group that owns its
//     target row also counts that row), or eight rows nobody targets -- and warps draw items from
//     the stage's shared counter, heaviest first, so the SM stays balanced without any CTA-wide
//     barrier; a warp that finds the counter exhausted releases the stage (empty[stage] counts
//     the 24 warps) and moves on.
// Partial counts go to the CTA's shared accumulators with shared-memory atomics, then to the global
// int64 accumulators (integer adds: order-free, exact); the last CTA to finish inverts the subset
// sums into the 16 minterm x target counts and leaves the scratch zeroed for the next call.
__global__ void __launch_bounds__(kCountThreads, 1)
rule_counts_kernel(const __grid_constant__ CUtensorMap tmap, int box_rows, int n_boxes, int n_stages,
                   int n_teams, int n_rows, int64_t n_tiles, int n_groups, const unsigned char* __restrict__ plan,
                   unsigned long long* __restrict__ acc /*[n_rows + 11*G], zero on entry, zero on exit*/,
                   unsigned int* __restrict__ cta_ticket /*zero on entry, zero on exit*/,
                   int64_t n_cells, int64_t* __restrict__ out_counts /*[G][16]*/,
                   void* const* __restrict__ peer_regions /*[n_ranks] or NULL: also push the counts to the peers*/,
                   int rank, int n_ranks,
                   int sums_only /*1: leave the subset sums in acc for scb_population_fitness_sums (no inversion, no
                                   ticket on one GPU; sharded: the last CTA pushes the sums, not the counts)*/) {
    extern __shared__ __align__(128) uint32_t smem[];
    // layout: stage[n_stages][(rows_alloc+1)][32] | desc[G] (uint4) | cta_acc[n_rows + 11*G] |
    //         full[kMaxStages], empty[kMaxStages] | s_tile[kMaxStages], s_next[kMaxStages] |
    //         order[G] (u16) | orphans[n_rows] (u16)
    // rows_alloc = n_boxes * box_rows >= n_rows (TMA zero-fills the excess); row rows_alloc is the
    // explicit all-zero row that unused parent slots read.
    const int rows_alloc = n_boxes * box_rows;
    const int stage_words = (rows_alloc + 1) * kTileWords;
    uint32_t* stages = smem;
    uint4* desc = reinterpret_cast<uint4*>(stages + (size_t)n_stages * stage_words);
    uint32_t* cta_acc = reinterpret_cast<uint32_t*>(desc + n_groups);
    const int n_acc = n_rows + kTermsPerGroup * n_groups;
    uint64_t* full = reinterpret_cast<uint64_t*>(cta_acc + ((n_acc + 1) & ~1));
    uint64_t* empty = full + kMaxStages;
    volatile int* s_tile = reinterpret_cast<volatile int*>(empty + kMaxStages);
    int* s_next = const_cast<int*>(s_tile) + kMaxStages;
    unsigned short* order = reinterpret_cast<unsigned short*>(s_next + kMaxStages);
    unsigned short* orphans = order + ((n_groups + 1) & ~1);

    const PlanLayout L = plan_layout(n_rows, n_groups);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    SCB_TRACE_AT(0, tid == 0);
#ifdef SCB_TRACE
    __shared__ unsigned int s_pdl_i;
    if (tid == 0 && blockIdx.x == 0) { s_pdl_i = atomicAdd(&g_pdl_a_n, 1u) & 63u; g_pdl_a[s_pdl_i][0] = global_timer_ns(); }
#endif

    // the plan's tables are fetched first thing (the loads fly while the barriers are set up)
    const int n_orphans = reinterpret_cast<const int*>(plan)[2];
    constexpr int kPlanPerThread = (kMaxGroupsPerLaunch + kConsumerWarps * 32 - 1) / (kConsumerWarps * 32);
    uint4 my_desc[kPlanPerThread];
    unsigned short my_order[kPlanPerThread];
    if (warp < kConsumerWarps) {
        const uint4* g_desc = reinterpret_cast<const uint4*>(plan + L.desc);
        const unsigned short* g_order = reinterpret_cast<const unsigned short*>(plan + L.order);
#pragma unroll
        for (int k = 0; k < kPlanPerThread; ++k) {
            const int i = tid + k * kConsumerWarps * 32;
            if (i < n_groups) {
                my_desc[k] = __ldg(g_desc + i);
                my_order[k] = __ldg(g_order + i);
            }
        }
    }
    const int team_warps = kConsumerWarps / n_teams;
    if (tid == kConsumerWarps * 32) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");  // hide the descriptor fetch
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], team_warps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    SCB_TRACE_AT(1, tid == 0);

    if (warp == kConsumerWarps) {
        // ---------------- producer ----------------
        if (lane == 0) {
            const uint32_t tile_bytes = (uint32_t)(rows_alloc * kTileWords * 4);
            int stage = 0;
            uint32_t phase = 1;  // parity of the "previous" empty phase: passes at once for the first lap
            int markers = 0;
            // at most kInFlight copies are outstanding: enough to cover the HBM latency at this SM's
            // share of the bandwidth; more only queue behind each other and delay the first tile
            const int in_flight = n_stages < kInFlight ? n_stages : kInFlight;  // never waits on a re-armed stage
            int land_stage = 0, seq = 0;
            uint32_t land_phase = 0;
            // Tiles are CLAIMED, not strided: the kernel may start on a few SMs while the kernel before it in the
            // stream (the previous pass's fitness kernel, programmatic dependent launch) still holds the others;
            // a CTA that starts early then simply takes more tiles.  cta_ticket[1] is the claim counter (zero on
            // entry; the last CTA to leave the claiming loop zeroes it, see below); the next claim is issued one
            // tile ahead so its round trip to L2 hides behind the copy being set up.
            // The first two tiles of a CTA are its own (no round trip before the first copies, and the burst of first
            // claims -- every CTA at once, on one address -- has two tiles' time to return); claims start at 2 * gridDim.x
            // and run three tiles ahead.
            unsigned int* claim = cta_ticket + 1;
            const int64_t claim_base = 2 * (int64_t)gridDim.x;
            int64_t tile = blockIdx.x;
            int64_t q1 = tile + gridDim.x;
            int64_t q2 = claim_base + (int64_t)atomicAdd(claim, 1u);
            int64_t q3 = claim_base + (int64_t)atomicAdd(claim, 1u);  // (the pipeline fill wants in_flight tiles at once)
            for (;; tile = q1, q1 = q2, q2 = q3, q3 = q3 < n_tiles ? claim_base + (int64_t)atomicAdd(claim, 1u) : q3, ++seq) {
                if (seq >= in_flight) {
                    mbar_wait(&full[land_stage], land_phase);
                    if (++land_stage == n_stages) {
                        land_stage = 0;
                        land_phase ^= 1;
                    }
                }
                mbar_wait(&empty[stage], phase);
                *reinterpret_cast<volatile int*>(s_next + stage) = 0;
                if (tile >= n_tiles) {  // end marker, one per team: consumers leave when they reach theirs
                    s_tile[stage] = -1;
                    mbar_arrive(&full[stage]);
                    if (++markers == n_teams) break;
                } else {
                    s_tile[stage] = (int)tile;
                    uint32_t* dst = stages + (size_t)stage * stage_words;
#ifdef SCB_TRACE
                    if (g_dbg & 2) {
                        mbar_arrive(&full[stage]);
                    } else
#endif
                    {
                    mbar_arrive_expect_tx(&full[stage], tile_bytes);
                    for (int b = 0; b < n_boxes; ++b)
                        tma_load_2d(dst + (size_t)b * box_rows * kTileWords, &tmap, (int)tile * kTileWords,
                                    b * box_rows, &full[stage]);
                    }
                }
                if (++stage == n_stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            // this CTA has claimed its last tile; the last CTA to say so zeroes the claim counter for the next launch.
            // The reset has to be in place before this CTA lets the next kernel of the stream start (the
            // griddepcontrol.launch_dependents below: a programmatic dependent -- another launch of this kernel,
            // say -- claims tiles before ITS griddepcontrol.wait), and a dependent only starts once EVERY CTA
            // has triggered it, the last claimer's included.  The consumers are still a few tiles behind the
            // producer here, so the round trip costs nothing.
            if (atomicAdd(cta_ticket + 2, 1u) == gridDim.x - 1) {
                atomicExch(cta_ticket + 1, 0u);
                atomicExch(cta_ticket + 2, 0u);
                __threadfence();
            }
        }
        __syncwarp();
    } else {
        // ---------------- consumers ----------------
        // set-up (overlaps the first copies): accumulators, all-zero rows, the plan's tables
        if (!kDirectRed)
            for (int i = tid; i < n_acc; i += kConsumerWarps * 32) cta_acc[i] = 0;
        for (int s = 0; s < n_stages; ++s)
            if (tid < kTileWords) stages[(size_t)s * stage_words + (size_t)rows_alloc * kTileWords + tid] = 0;
        {
#pragma unroll
            for (int k = 0; k < kPlanPerThread; ++k) {
                const int i = tid + k * kConsumerWarps * 32;
                if (i < n_groups) {
                    desc[i] = my_desc[k];
                    order[i] = my_order[k];
                }
            }
            const unsigned short* g_orph = reinterpret_cast<const unsigned short*>(plan + L.orphans);
            for (int i = tid; i < n_orphans; i += kConsumerWarps * 32) orphans[i] = __ldg(g_orph + i);
        }
        consumer_sync();
        SCB_TRACE_AT(2, tid == 0);

        const int quad = lane >> 2;  // 0..7: which of the item's eight groups / rows
        const int l = lane & 3;      // lane in quad
        // lane l of quad q reads 16-byte chunks c0 and c0^4: the eight lanes of every quarter-warp
        // (two quads) cover eight distinct chunks, so the 128-bit loads are conflict-free whatever
        // rows the quads point at
        const int c0 = l ^ ((quad & 1) << 2);
        const int n_gitems = (n_groups + 7) >> 3;
        const int n_items = n_gitems + ((n_orphans + 7) >> 3);

        // the warps form n_teams teams; team t takes the tiles t, t + n_teams, ... of the ring, so a
        // warp pays the per-tile bookkeeping for every n_teams-th tile only
        int stage = warp / team_warps;
        uint32_t phase = 0;
        for (;;) {
            mbar_wait(&full[stage], phase);
            SCB_TRACE_AT(3, tid == 0 && stage == 0 && phase == 0);
            if (s_tile[stage] < 0) break;
            const char* buf = reinterpret_cast<const char*>(stages + (size_t)stage * stage_words);
            const uint32_t next_addr = smem_u32(s_next + stage);
            auto claim = [&]() {
                int it = 0;
                if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(it) : "r"(next_addr) : "memory");
                return __shfl_sync(0xFFFFFFFFu, it, 0);
            };
            int item = claim();
            while (item < n_items) {
                const int next_item = claim();  // drawn one ahead: its latency hides behind the arithmetic
#ifdef SCB_TRACE
                if (g_dbg & 1) {
                    item = next_item;
                    continue;
                }
#endif
                if (item < n_gitems) {
                    const int slot = item * 8 + quad;
                    const bool live = slot < n_groups;
                    const int sl = live ? slot : n_groups - 1;
                    const uint4 d = desc[sl];
                    const int g = order[sl];
                    const int arity = __reduce_max_sync(0xFFFFFFFFu, live ? (int)(d.x & 3u) : 0);
                    const bool own = (d.x & kDescOwn) != 0;
                    const uint32_t* pa = reinterpret_cast<const uint32_t*>(buf + (d.x & ~7u));
                    const uint32_t* pb = reinterpret_cast<const uint32_t*>(buf + d.y);
                    const uint32_t* pc = reinterpret_cast<const uint32_t*>(buf + d.z);
                    const uint32_t* pt = reinterpret_cast<const uint32_t*>(buf + d.w);
                    uint32_t* tally = kDirectRed ? reinterpret_cast<uint32_t*>(acc) : cta_acc;
                    uint32_t* dst = tally + n_rows + (size_t)g * kTermsPerGroup;
                    uint32_t* dst_t = tally + (d.w >> 7);
                    if (arity >= 3) group_terms<3>(pa, pb, pc, pt, c0, l, live, own, dst, dst_t);
                    else if (arity == 2) group_terms<2>(pa, pb, pc, pt, c0, l, live, own, dst, dst_t);
                    else if (arity == 1) group_terms<1>(pa, pb, pc, pt, c0, l, live, own, dst, dst_t);
                    else group_terms<0>(pa, pb, pc, pt, c0, l, live, own, dst, dst_t);
                } else {
                    const int idx = (item - n_gitems) * 8 + quad;
                    const bool live = idx < n_orphans;
                    const int r = orphans[live ? idx : n_orphans - 1];
                    uint32_t x[8];
                    load8(x, reinterpret_cast<const uint32_t*>(buf) + (size_t)r * kTileWords, c0);
                    const uint32_t s = quad_sum(popc8(x));
                    if (l == 0 && live) atomicAdd((kDirectRed ? reinterpret_cast<uint32_t*>(acc) : cta_acc) + r, s);
                }
                item = next_item;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);  // this warp is done with the stage
            stage += n_teams;
            if (stage >= n_stages) {
                stage -= n_stages;
                phase ^= 1;
            }
        }
    }
    SCB_TRACE_AT(4, tid == 0);
    __syncthreads();
    SCB_TRACE_AT(5, tid == 0);

    // Programmatic dependent launch, both directions: the kernel that follows in the stream (e.g. the
    // fitness kernel) may be scheduled from here on, and everything above may have run while the
    // kernel BEFORE this one (a reader of out_counts from the previous pass) was still draining --
    // nothing above writes global memory.  Both are no-ops for plain launches.
#ifdef SCB_TRACE
    if (tid == 0 && blockIdx.x == 0) g_pdl_a[s_pdl_i][1] = global_timer_ns();
#endif
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef SCB_TRACE
    if (tid == 0 && blockIdx.x == 0) g_pdl_a[s_pdl_i][2] = global_timer_ns();
#endif
    // fold this CTA's sums into the global int64 accumulators (integer adds: order-free, exact)
    if (!kDirectRed)
        for (int i = tid; i < n_acc; i += blockDim.x)
            if (cta_acc[i]) atomicAdd(acc + i, (unsigned long long)cta_acc[i]);
    // sums_only on one GPU: that is all -- the consumer (the fitness kernel) is ordered behind this grid by
    // the stream, inverts the sums itself and zeroes the accumulators: no ticket, no last CTA
    if (sums_only && !peer_regions) return;
    __syncthreads();
    SCB_TRACE_AT(6, tid == 0);
    // one thread publishes the CTA: the barrier above orders every thread's adds before its
    // release, the acquire side orders the last CTA's reads after every other CTA's ticket
    __shared__ unsigned int s_ticket;
    if (tid == 0) {
        unsigned int t;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(cta_ticket) : "memory");
        s_ticket = t;
    }
    __syncthreads();
    SCB_TRACE_AT(7, tid == 0);
    if (s_ticket != gridDim.x - 1) return;
    // sharded cells: this rank's counts also go to its slot of every peer's exchange region as
    // tagged words (PeerHeader, scb_common.cuh) -- the push costs no kernel, no fence, no flag
    __shared__ unsigned int s_epoch;
    if (peer_regions && tid == 0) {
        PeerHeader* hdr = static_cast<PeerHeader*>(peer_regions[rank]);
        s_epoch = (unsigned int)(atomicAdd(&hdr->push_ticket, (unsigned long long)n_ranks) / (unsigned long long)n_ranks);
    }
    if (peer_regions) __syncthreads();
    const int64_t peer_slot = peer_regions ? ((int64_t)(s_epoch & 1u) * n_ranks + rank) * ((int64_t)n_groups * 16) : 0;
    if (sums_only) {
        // sharded cells: the 16 subset sums of every group go to every rank as they are; the fitness kernels
        // sum them over the ranks and invert
        const int q = tid & 15;
        for (int pos = tid >> 4; pos < n_groups; pos += kCountThreads / 16) {
            const int g = order[pos];
            const long long v = subset_sum_local(desc[pos], g, q, acc, n_rows, rows_alloc, n_cells);
            for (int r = 0; r < n_ranks; ++r)
                st_peer(reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(peer_regions[r]) + kPeerSlotsOffset) +
                            peer_slot + (size_t)g * 16 + q,
                        peer_word(s_epoch + 1u, v));
        }
    } else {
        // last CTA: Moebius inversion of the subset sums (moebius_pass above)
        moebius_pass(desc, order, acc, n_rows, rows_alloc, n_groups, n_groups, n_cells, out_counts, peer_regions, n_ranks,
                     peer_slot, s_epoch + 1u, tid, false);
    }
    // leave the scratch as it was found (all zero) so the next call needs no memset
    __syncthreads();

    SCB_TRACE_AT(8, tid == 0);
    for (int i = tid; i < n_acc; i += blockDim.x) acc[i] = 0ull;
    if (tid == 0) *cta_ticket = 0u;
    SCB_TRACE_AT(9, tid == 0);
}

static void box_shape(int n_rows, int* box_rows, int* n_boxes) {
    *n_boxes = (n_rows + 255) / 256;            // a TMA box is at most 256 elements per dimension
    *box_rows = (n_rows + *n_boxes - 1) / *n_boxes;
}

// shared memory besides the stages, and the bytes of one stage
static size_t counts_fixed_smem(int n_rows, int n_groups) {
    size_t n_acc = (size_t)n_rows + (size_t)kTermsPerGroup * n_groups;
    return (size_t)n_groups * 16 + ((n_acc + 1) & ~(size_t)1) * 4 + kMaxStages * (8 + 8 + 4 + 4) +
           ((size_t)n_groups * 2 + 4) + ((size_t)n_rows * 2 + 16);
}
static size_t counts_stage_bytes(int n_rows) {
    int box_rows, n_boxes;
    box_shape(n_rows, &box_rows, &n_boxes);
    return ((size_t)box_rows * n_boxes + 1) * kTileWords * 4;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

// [n_rows][wpad] uint32 bitplanes as a 2-D tensor, box = [box_rows][32 words]
static int make_tensor_map(const uint32_t* planes, int n_rows, int64_t wpad, int box_rows, CUtensorMap* out) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
            set_error("rule_counts: cuTensorMapEncodeTiled is not available from this driver");
            return SCB_ERR_CUDA;
        }
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    cuuint64_t dims[2] = {(cuuint64_t)wpad, (cuuint64_t)n_rows};
    cuuint64_t strides[1] = {(cuuint64_t)wpad * 4};
    cuuint32_t box[2] = {(cuuint32_t)kTileWords, (cuuint32_t)box_rows};
    cuuint32_t elem[2] = {1, 1};
    CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint32_t*>(planes), dims, strides, box,
                        elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("rule_counts: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
        return SCB_ERR_CUDA;
    }
    return SCB_OK;
}

}  // namespace scb

using namespace scb;

#ifdef SCB_TRACE
extern "C" int scb_debug_pdl_a(unsigned long long* out /*HOST [64][4]*/) {
    SCB_CUDA(cudaDeviceSynchronize());
    SCB_CUDA(cudaMemcpyFromSymbol(out, g_pdl_a, sizeof(unsigned long long) * 64 * 4));
    return SCB_OK;
}
extern "C" int scb_debug_trace_read(long long* out /*HOST [256][12]*/) {
    SCB_CUDA(cudaDeviceSynchronize());
    SCB_CUDA(cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * 256 * 12));
    return SCB_OK;
}
#endif


static int64_t acc_bytes(int n_rows, int n_groups) {
    int gn = n_groups < kMaxGroupsPerLaunch ? n_groups : kMaxGroupsPerLaunch;
    return (int64_t)pad16(((size_t)n_rows + (size_t)kTermsPerGroup * gn) * 8 + kTicketBytes);
}

extern "C" int64_t scb_rule_counts_plan_bytes(int n_rows, int n_groups) {
    if (n_rows <= 0 || n_groups < 0) return 0;
    int64_t total = 0;
    for (int g0 = 0; g0 < n_groups; g0 += kMaxGroupsPerLaunch) {
        int gn = n_groups - g0 < kMaxGroupsPerLaunch ? n_groups - g0 : kMaxGroupsPerLaunch;
        total += (int64_t)plan_layout(n_rows, gn).bytes;
    }
    return total;
}

extern "C" int64_t scb_rule_counts_scratch_bytes(int n_rows, int n_groups) {
    // int64 accumulators (rows + 11 per group of one launch slice) + the CTA ticket, then room for
    // the plan that scb_rule_counts builds for itself
    if (n_rows < 0 || n_groups < 0) return 0;
    return acc_bytes(n_rows, n_groups) + scb_rule_counts_plan_bytes(n_rows, n_groups);
}

extern "C" int scb_rule_counts_plan_build(int n_rows, int n_groups, const int32_t* target_row,
                                          const int32_t* parent_row, void* plan, int64_t plan_bytes,
                                          scb_stream_t stream) {
    SCB_REQUIRE(n_rows > 0 && n_rows < 65536, "rule_counts: n_rows=%d out of range", n_rows);
    SCB_REQUIRE(n_groups >= 0, "rule_counts: negative n_groups");
    if (n_groups == 0) return SCB_OK;
    SCB_REQUIRE(target_row && parent_row && plan, "rule_counts_plan_build: null pointer");
    SCB_REQUIRE(reinterpret_cast<uintptr_t>(plan) % 16 == 0, "rule_counts_plan_build: plan must be 16-byte aligned");
    SCB_REQUIRE(plan_bytes >= scb_rule_counts_plan_bytes(n_rows, n_groups), "rule_counts_plan_build: plan buffer too small (%lld < %lld bytes)", (long long)plan_bytes, (long long)scb_rule_counts_plan_bytes(n_rows, n_groups));
    int box_rows, n_boxes;
    box_shape(n_rows, &box_rows, &n_boxes);
    unsigned char* at = static_cast<unsigned char*>(plan);
    for (int g0 = 0; g0 < n_groups; g0 += kMaxGroupsPerLaunch) {
        int gn = n_groups - g0 < kMaxGroupsPerLaunch ? n_groups - g0 : kMaxGroupsPerLaunch;
        rule_counts_plan_kernel<<<1, 256, 0, as_stream(stream)>>>(n_rows, gn, box_rows * n_boxes, target_row + g0,
                                                                 parent_row + 3 * (size_t)g0, at);
        SCB_LAUNCH_CHECK("rule_counts_plan_kernel");
        at += plan_layout(n_rows, gn).bytes;
    }
    return SCB_OK;
}

static int counts_run(const uint32_t* planes, int n_rows, int64_t wpad, int64_t n_cells, int n_groups,
                      const void* plan, void* scratch, int64_t scratch_bytes, int64_t* out_counts, int flags,
                      void* const* peer_regions, int rank, int n_ranks, int sums_only, scb_stream_t stream) {
    SCB_REQUIRE(n_rows > 0 && n_rows < 65536, "rule_counts: n_rows=%d out of range", n_rows);
    SCB_REQUIRE(wpad > 0 && wpad % 4 == 0 && wpad * 32 >= n_cells && n_cells >= 0, "rule_counts: wpad must be a positive multiple of 4 covering n_cells");
    SCB_REQUIRE(n_cells < ((int64_t)1 << 31), "rule_counts: at most 2^31-1 cells per call (shard the cells)");
    SCB_REQUIRE(n_groups >= 0, "rule_counts: negative n_groups");
    SCB_REQUIRE((flags & ~(SCB_COUNTS_SCRATCH_CLEAN | SCB_COUNTS_CHAINED)) == 0, "rule_counts: unknown flags 0x%x", flags);
    if (n_groups == 0) return SCB_OK;
    SCB_REQUIRE(planes && plan && scratch && (out_counts || sums_only), "rule_counts: null pointer");
    SCB_REQUIRE(!sums_only || n_groups <= kMaxGroupsPerLaunch, "rule_sums: at most %d groups (one launch)", kMaxGroupsPerLaunch);
    SCB_REQUIRE(reinterpret_cast<uintptr_t>(planes) % 16 == 0, "rule_counts: planes must be 16-byte aligned");
    SCB_REQUIRE(reinterpret_cast<uintptr_t>(scratch) % 8 == 0, "rule_counts: scratch must be 8-byte aligned");
    SCB_REQUIRE(reinterpret_cast<uintptr_t>(plan) % 16 == 0, "rule_counts: plan must be 16-byte aligned");
    const int64_t n_tiles = (wpad + kTileWords - 1) / kTileWords;
    SCB_REQUIRE(n_tiles < ((int64_t)1 << 26), "rule_counts: too many tiles");
    const int64_t need = acc_bytes(n_rows, n_groups);
    if (need > scratch_bytes) {
        set_error("rule_counts: scratch too small (%lld < %lld bytes)", (long long)scratch_bytes, (long long)need);
        return SCB_ERR_SCRATCH;
    }
    int box_rows, n_boxes;
    box_shape(n_rows, &box_rows, &n_boxes);
    CUtensorMap tmap;
    {
        int rc = make_tensor_map(planes, n_rows, wpad, box_rows, &tmap);
        if (rc != SCB_OK) return rc;
    }
    const size_t stage_bytes = counts_stage_bytes(n_rows);
    // the kernel leaves the scratch zeroed; a caller that vouches for it skips the memset
    if (!(flags & SCB_COUNTS_SCRATCH_CLEAN)) SCB_CUDA(cudaMemsetAsync(scratch, 0, (size_t)need, as_stream(stream)));
    const unsigned char* plan_at = static_cast<const unsigned char*>(plan);
    for (int g0 = 0; g0 < n_groups; g0 += kMaxGroupsPerLaunch) {
        int gn = n_groups - g0 < kMaxGroupsPerLaunch ? n_groups - g0 : kMaxGroupsPerLaunch;
        const size_t fixed = counts_fixed_smem(n_rows, gn);
        const size_t budget = 227 * 1024 - 256;  // static shared variables (+ alignment) live in the remainder
        int n_stages = fixed < budget ? (int)((budget - fixed) / stage_bytes) : 0;
        if (n_stages > kMaxStages) n_stages = kMaxStages;
        if (const char* cap = getenv("SCB_COUNTS_STAGES")) {  // experiments
            int c = atoi(cap);
            if (c >= 2 && c < n_stages) n_stages = c;
        }
        if (n_stages < 2) {
            set_error("rule_counts: %d rows x %d groups do not fit in shared memory", n_rows, gn);
            return SCB_ERR_UNSUPPORTED;
        }
        int n_teams = n_stages >= 6 ? 3 : (n_stages >= 4 ? 2 : 1);
        if (const char* cap = getenv("SCB_COUNTS_TEAMS")) {  // experiments
            int c = atoi(cap);
            if (c >= 1 && c <= n_stages && kConsumerWarps % c == 0) n_teams = c;
        }
        const size_t smem = fixed + (size_t)n_stages * stage_bytes;
        // the attribute is raised once per thread and device to the largest size seen (keeps the enqueue
        // path free of non-stream API calls, e.g. under CUDA-graph capture)
        static thread_local SmemMarks marks;
        if (int rc = ensure_dynamic_smem((const void*)rule_counts_kernel, smem, marks)) return rc;
        int64_t grid = sm_count();
        if (grid > n_tiles) grid = n_tiles;
        const int64_t n_acc = (int64_t)n_rows + (int64_t)kTermsPerGroup * gn;
        auto* acc = static_cast<unsigned long long*>(scratch);
        auto* ticket = reinterpret_cast<unsigned int*>(acc + n_acc);
#ifdef SCB_TRACE
        if (getenv("SCB_COUNTS_DBG")) {  // (a copy node between two kernels breaks their programmatic edge)
            int dbg = atoi(getenv("SCB_COUNTS_DBG"));
            SCB_CUDA(cudaMemcpyToSymbolAsync(g_dbg, &dbg, sizeof(int), 0, cudaMemcpyHostToDevice, as_stream(stream)));
        }
#endif
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(kCountThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = as_stream(stream);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        // only the first slice of a call may overlap the previous kernel (later slices share the scratch), and
        // only on the caller's word that that kernel produced neither the bitplanes nor the plan, which
        // this kernel reads before its griddepcontrol.wait (SCB_COUNTS_CHAINED)
        cfg.numAttrs = (g0 == 0 && (flags & SCB_COUNTS_SCRATCH_CLEAN) && (flags & SCB_COUNTS_CHAINED) && !getenv("SCB_NO_PDL")) ? 1 : 0;
        SCB_CUDA(cudaLaunchKernelEx(&cfg, rule_counts_kernel, tmap, box_rows, n_boxes, n_stages, n_teams, n_rows,
                                    n_tiles, gn, plan_at, acc, ticket, n_cells, out_counts ? out_counts + (size_t)g0 * 16 : nullptr,
                                    peer_regions, rank, n_ranks, sums_only));
        plan_at += plan_layout(n_rows, gn).bytes;
    }
    return SCB_OK;
}

extern "C" int scb_rule_sums_supported(int n_rows, int n_groups, int n_untargeted_rows) {
    (void)n_untargeted_rows;
    if (n_rows <= 0 || n_groups <= 0 || n_groups > kMaxGroupsPerLaunch) return 0;  // one launch of the counting kernel
    // ... and the fitness kernel keeps its tables (200 B per group) plus a copy of the accumulators, of the
    // descriptors and of the group order in shared memory (scb_rules.cu)
    const size_t smem = (size_t)n_groups * 200 + ((size_t)n_rows + (size_t)kTermsPerGroup * n_groups + 3) * 8 +
                        (size_t)n_groups * 18 + 16;
    return smem <= 200 * 1024 ? 1 : 0;
}

extern "C" int scb_rule_sums_run(const uint32_t* planes, int n_rows, int64_t wpad, int64_t n_cells, int n_groups,
                                 int n_untargeted_rows, const void* plan, void* scratch, int64_t scratch_bytes,
                                 int flags, void* const* peer_regions, int rank, int n_ranks, scb_stream_t stream) {
    (void)n_untargeted_rows;
    SCB_REQUIRE(n_groups > 0, "rule_sums: n_groups=%d out of range", n_groups);
    SCB_REQUIRE(!peer_regions || (n_ranks >= 1 && n_ranks <= SCB_MAX_PEERS && rank >= 0 && rank < n_ranks), "rule_sums: rank %d of %d", rank, n_ranks);
    return counts_run(planes, n_rows, wpad, n_cells, n_groups, plan, scratch, scratch_bytes, nullptr, flags,
                      peer_regions, rank, peer_regions ? n_ranks : 1, 1, stream);
}

extern "C" int scb_rule_counts_run(const uint32_t* planes, int n_rows, int64_t wpad, int64_t n_cells,
                                   int n_groups, const void* plan, void* scratch, int64_t scratch_bytes,
                                   int64_t* out_counts, int flags, scb_stream_t stream) {
    return counts_run(planes, n_rows, wpad, n_cells, n_groups, plan, scratch, scratch_bytes, out_counts, flags,
                      nullptr, 0, 1, 0, stream);
}

extern "C" int scb_rule_counts_run_push(const uint32_t* planes, int n_rows, int64_t wpad, int64_t n_cells,
                                        int n_groups, const void* plan, void* scratch, int64_t scratch_bytes,
                                        int64_t* out_counts, int flags, void* const* peer_regions, int rank,
                                        int n_ranks, scb_stream_t stream) {
    SCB_REQUIRE(n_ranks >= 1 && n_ranks <= SCB_MAX_PEERS && rank >= 0 && rank < n_ranks, "rule_counts_run_push: rank %d of %d", rank, n_ranks);
    SCB_REQUIRE(peer_regions, "rule_counts_run_push: null peer table");
    SCB_REQUIRE(n_groups <= kMaxGroupsPerLaunch, "rule_counts_run_push: at most %d groups (one launch)", kMaxGroupsPerLaunch);
    return counts_run(planes, n_rows, wpad, n_cells, n_groups, plan, scratch, scratch_bytes, out_counts, flags,
                      peer_regions, rank, n_ranks, 0, stream);
}

extern "C" int scb_rule_counts_ex(const uint32_t* planes, int n_rows, int64_t wpad, int64_t n_cells,
                                  int n_groups, const int32_t* target_row, const int32_t* parent_row,
                                  void* scratch, int64_t scratch_bytes, int64_t* out_counts, int flags,
                                  scb_stream_t stream) {
    SCB_REQUIRE(n_rows > 0 && n_rows < 65536, "rule_counts: n_rows=%d out of range", n_rows);
    SCB_REQUIRE(n_groups >= 0, "rule_counts: negative n_groups");
    if (n_groups == 0) return SCB_OK;
    SCB_REQUIRE(planes && target_row && parent_row && scratch && out_counts, "rule_counts: null pointer");
    SCB_REQUIRE(reinterpret_cast<uintptr_t>(scratch) % 16 == 0, "rule_counts: scratch must be 16-byte aligned");
    const int64_t accb = acc_bytes(n_rows, n_groups), planb = scb_rule_counts_plan_bytes(n_rows, n_groups);
    if (accb + planb > scratch_bytes) {
        set_error("rule_counts: scratch too small (%lld < %lld bytes)", (long long)scratch_bytes, (long long)(accb + planb));
        return SCB_ERR_SCRATCH;
    }
    void* plan = static_cast<unsigned char*>(scratch) + accb;
    int rc = scb_rule_counts_plan_build(n_rows, n_groups, target_row, parent_row, plan, planb, stream);
    if (rc != SCB_OK) return rc;
    return scb_rule_counts_run(planes, n_rows, wpad, n_cells, n_groups, plan, scratch, accb, out_counts, flags, stream);
}

extern "C" int scb_rule_counts(const uint32_t* planes, int n_rows, int64_t wpad, int64_t n_cells,
                               int n_groups, const int32_t* target_row, const int32_t* parent_row,
                               void* scratch, int64_t scratch_bytes, int64_t* out_counts,
                               scb_stream_t stream) {
    return scb_rule_counts_ex(planes, n_rows, wpad, n_cells, n_groups, target_row, parent_row, scratch,
                              scratch_bytes, out_counts, 0, stream);
}
