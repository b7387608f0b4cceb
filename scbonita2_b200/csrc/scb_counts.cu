// scb_rule_counts: joint (A,B,C,target) counts per node in ONE pass over the cell bitplanes.
//
// Replaces the inner loop of RuleDetermination.calculate_error (rule_determination.py:509-582).
// For a rule with truth table L over (A,B,C) the reference's mismatch count is
//     sum_k #{cells: (A,B,C)==k, target != L_k},
// so the 16 joint counts of a node are a sufficient statistic for EVERY candidate rule and
// every GA individual.  They are obtained as popcounts of the 15 non-empty AND-subsets of the
// four bitplane rows (4 single rows, 11 multi-variable terms) followed by a Moebius inversion.
//
// Kernel structure (B200):
//  * persistent CTAs pull 32-word cell tiles from a global ticket; a tile (all rows x 128 B) is
//    staged in shared memory by the TMA engine: one cp.async.bulk per row, completion on an
//    mbarrier, two stages, so the copy of tile i+1 overlaps the arithmetic on tile i;
//  * a QUAD of lanes owns one node per tile: lane l reads words 4*(j^q)+l, j = 0..7 (q = quad
//    in warp), which keeps the 8 quads of a warp on disjoint banks;
//  * the eight words of a term are compressed with a Harley-Seal carry-save tree (7 CSA = 14
//    LOP3) into ones/twos/fours/eights, so a term costs 4 POPC instead of 8: POPC issues on the
//    quarter-rate XU pipe, which was the measured limiter of the plain version
//    (profiles/r1_counts_v1_ncu.txt);
//  * quad partials are folded with two shuffles on 16-bit packed pairs, accumulated per CTA in
//    shared memory, added to global int64 accumulators with atomics (integer adds: exact and
//    order-free) and inverted by the last CTA to finish.  No second kernel.
#include <cuda.h>
#include <stdlib.h>

#include "scb_common.cuh"

namespace scb {

constexpr int kTermsPerGroup = 11;  // multi-variable AND terms
constexpr int kTileWords = 32;      // words per row per tile (128 B)
constexpr int kStages = 2;
constexpr int kCountThreads = 256;
constexpr int kMaxClasses = 15;
constexpr int kTicketBytes = (kMaxClasses + 1) * 4;

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

// carry-save adder: (sum, carry) of three words, 2 LOP3
__device__ __forceinline__ void csa(uint32_t& carry, uint32_t& sum, uint32_t a, uint32_t b, uint32_t c) {
    sum = lop3<0x96>(a, b, c);
    carry = lop3<0xE8>(a, b, c);
}

// popcount of eight words through a Harley-Seal tree: 7 CSA + 4 POPC
__device__ __forceinline__ uint32_t popc8(const uint32_t (&x)[8]) {
    uint32_t ones, twos_a, twos_b, twos_c, twos_d, twos, fours_a, fours_b, fours, eights;
    csa(twos_a, ones, x[0], x[1], x[2]);
    csa(twos_b, ones, ones, x[3], x[4]);
    csa(twos_c, ones, ones, x[5], x[6]);
    twos_d = ones & x[7];
    ones ^= x[7];
    csa(fours_a, twos, twos_a, twos_b, twos_c);
    fours_b = twos & twos_d;
    twos ^= twos_d;
    eights = fours_a & fours_b;
    fours = fours_a ^ fours_b;
    return __popc(ones) + 2 * __popc(twos) + 4 * __popc(fours) + 8 * __popc(eights);
}

// number of parent slots a group binds when they are packed to the left (A, AB, ABC); any other
// pattern takes the general three-parent path
__device__ __forceinline__ int group_arity(int ra, int rb, int rc) {
    if (rc >= 0) return 3;
    if (rb >= 0) return ra >= 0 ? 2 : 3;
    return ra >= 0 ? 1 : 0;
}

// The AND-subset popcounts of one group over one 32-word tile, taken by a quad (lane l reads the
// eight words widx[]), folded across the quad and added to the CTA accumulators dst[0..10]
// (term order of SCB_TERM_MASKS).  ARITY selects which terms can be non-zero.
template <int ARITY>
__device__ __forceinline__ void group_terms(const uint32_t* __restrict__ pa, const uint32_t* __restrict__ pb,
                                            const uint32_t* __restrict__ pc, const uint32_t* __restrict__ pt,
                                            const int (&widx)[8], int l, bool live, uint32_t* dst) {
    uint32_t a[8], t[8], x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        a[j] = pa[widx[j]];
        t[j] = pt[widx[j]];
    }
#define SCB_TERM(var, expr)                                      \
    _Pragma("unroll") for (int j = 0; j < 8; ++j) x[j] = (expr); \
    uint32_t var = popc8(x);
    if (ARITY == 1) {
        SCB_TERM(at, a[j] & t[j])
        at += __shfl_xor_sync(0xFFFFFFFFu, at, 1);
        at += __shfl_xor_sync(0xFFFFFFFFu, at, 2);
        if (live && l == 0) dst[4] += at;
        return;
    }
    uint32_t b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = pb[widx[j]];
    if (ARITY == 2) {
        SCB_TERM(ab, a[j] & b[j])
        SCB_TERM(at, a[j] & t[j])
        SCB_TERM(bt, b[j] & t[j])
        SCB_TERM(abt, lop3<0x80>(a[j], b[j], t[j]))
        // counts <= 256 per lane: two per register in 16-bit fields
        uint32_t p0 = ab | (at << 16), p1 = bt | (abt << 16);
        p0 += __shfl_xor_sync(0xFFFFFFFFu, p0, 1);
        p1 += __shfl_xor_sync(0xFFFFFFFFu, p1, 1);
        p0 += __shfl_xor_sync(0xFFFFFFFFu, p0, 2);
        p1 += __shfl_xor_sync(0xFFFFFFFFu, p1, 2);
        if (live) {  // lanes 0..3 own AB, AT, BT, ABT
            const uint32_t v = ((l & 2) ? p1 : p0) >> ((l & 1) * 16) & 0xFFFFu;
            const int term = l == 0 ? 0 : (l == 1 ? 4 : (l == 2 ? 5 : 7));
            dst[term] += v;
        }
        return;
    }
    uint32_t c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) c[j] = pc[widx[j]];
    uint32_t cnt[kTermsPerGroup];
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
#undef SCB_TERM_I
#undef SCB_TERM
    uint32_t pk[6];
#pragma unroll
    for (int k = 0; k < 5; ++k) pk[k] = cnt[2 * k] | (cnt[2 * k + 1] << 16);
    pk[5] = cnt[10];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        pk[k] += __shfl_xor_sync(0xFFFFFFFFu, pk[k], 1);
        pk[k] += __shfl_xor_sync(0xFFFFFFFFu, pk[k], 2);
    }
    // lane l of the quad owns terms l, l+4, l+8 (term 11 does not exist)
    const bool hi_pair = (l >> 1) != 0;
    const int shift = (l & 1) * 16;
    uint32_t v0 = ((hi_pair ? pk[1] : pk[0]) >> shift) & 0xFFFFu;
    uint32_t v1 = ((hi_pair ? pk[3] : pk[2]) >> shift) & 0xFFFFu;
    uint32_t v2 = ((hi_pair ? pk[5] : pk[4]) >> shift) & 0xFFFFu;
    if (live) {
        dst[l] += v0;
        dst[l + 4] += v1;
        if (l < 3) dst[l + 8] += v2;
    }
}

__global__ void __launch_bounds__(kCountThreads, 3)
rule_counts_kernel(const __grid_constant__ CUtensorMap tmap, int box_rows, int n_boxes,
                   int n_rows, int64_t wpad, int64_t n_tiles,
                   int n_groups, const int32_t* __restrict__ target_row,
                   const int32_t* __restrict__ parent_row,
                   unsigned long long* __restrict__ acc /*[n_rows + 11*G], zeroed*/,
                   unsigned int* __restrict__ tickets /*[kMaxClasses+1], zeroed: tile ticket per class, CTA ticket*/,
                   int n_classes, int64_t n_cells, int64_t* __restrict__ out_counts /*[G][16]*/) {
    extern __shared__ __align__(128) uint32_t smem[];
    // layout: stage[kStages][(rows_alloc+1)][32] | cta_acc[n_rows + 11*G] | mbarriers | tile ids
    // rows_alloc = n_boxes * box_rows >= n_rows (TMA zero-fills the excess); row rows_alloc is the
    // explicit all-zero row that unused parent slots read.
    const int rows_alloc = n_boxes * box_rows;
    const int stage_words = (rows_alloc + 1) * kTileWords;
    uint32_t* stages = smem;
    uint32_t* cta_acc = stages + (size_t)kStages * stage_words;
    const int n_acc = n_rows + kTermsPerGroup * n_groups;
    uint64_t* full = reinterpret_cast<uint64_t*>(cta_acc + ((n_acc + 1) & ~1));
    int* s_tile = reinterpret_cast<int*>(full + kStages);
    unsigned short* order = reinterpret_cast<unsigned short*>(s_tile + kStages);  // groups by arity
    __shared__ int s_arity_count[4], s_arity_fill[4];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int quad_in_warp = lane >> 2;             // q: 0..7
    const int l = lane & 3;                         // lane in quad
    const int n_quads = (blockDim.x >> 5) * 8;

    // Work split: groups (and rows) can be dealt round-robin into n_classes classes; CTA b belongs
    // to class b % n_classes and walks ALL tiles for its class (n_classes = 1 by default).
    const int cls = blockIdx.x % n_classes;
    const int cta_in_cls = blockIdx.x / n_classes;
    const int ctas_in_cls = (gridDim.x - cls + n_classes - 1) / n_classes;
    // one thread starts the copy of `tile` into `stage` (a tile past the end just completes the phase)
    auto issue_tile = [&](int stage, int tile) {
        s_tile[stage] = tile;
        if (tile >= n_tiles) {
            mbar_arrive(&full[stage]);
            return;
        }
        uint32_t* dst = stages + (size_t)stage * stage_words;
        mbar_arrive_expect_tx(&full[stage], (uint32_t)(rows_alloc * kTileWords * 4));
        for (int b = 0; b < n_boxes; ++b)
            tma_load_2d(dst + (size_t)b * box_rows * kTileWords, &tmap, tile * kTileWords, b * box_rows, &full[stage]);
    };
    // the first kStages tiles of a CTA are static: get them moving before any other set-up work
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < kStages; ++s) issue_tile(s, cta_in_cls + s * ctas_in_cls);
    }
    for (int i = tid; i < n_acc; i += blockDim.x) cta_acc[i] = 0;
    if (tid < 4) s_arity_count[tid] = s_arity_fill[tid] = 0;
    __syncthreads();
    // counting sort of the groups by arity, widest first (order inside a class is irrelevant)
    for (int g = tid; g < n_groups; g += blockDim.x)
        atomicAdd(&s_arity_count[group_arity(parent_row[3 * g], parent_row[3 * g + 1], parent_row[3 * g + 2])], 1);
    __syncthreads();
    for (int g = tid; g < n_groups; g += blockDim.x) {
        const int ar = group_arity(parent_row[3 * g], parent_row[3 * g + 1], parent_row[3 * g + 2]);
        int base = 0;
        for (int k = 3; k > ar; --k) base += s_arity_count[k];
        order[base + atomicAdd(&s_arity_fill[ar], 1)] = (unsigned short)g;
    }
    for (int s = 0; s < kStages; ++s)               // the all-zero row of every stage
        for (int i = tid; i < kTileWords; i += blockDim.x)
            stages[(size_t)s * stage_words + (size_t)rows_alloc * kTileWords + i] = 0;
    __syncthreads();

    // later tiles come from the global ticket, claimed one tile ahead so that its L2 round trip
    // hides behind the arithmetic
    const int first_dynamic = kStages * ctas_in_cls;
    auto claim = [&]() {
        int t = 0;
        if (lane == 0) t = first_dynamic + (int)atomicAdd(&tickets[cls], 1u);
        return t;  // valid in lane 0
    };

    for (int it = 0;; ++it) {
        const int stage = it % kStages;
        const uint32_t parity = (it / kStages) & 1;
        mbar_wait(&full[stage], parity);
        const int tile = s_tile[stage];
        if (tile >= n_tiles) break;  // tiles are claimed in increasing order: later stages are empty too
        int next_tile = 0;
        if (warp == 0) next_tile = claim();  // consumed after this tile's arithmetic
        uint32_t* buf = stages + (size_t)stage * stage_words;
        int widx[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) widx[j] = 4 * (j ^ quad_in_warp) + l;

        // ---- single-row popcounts: one quad per row ----
        // (loops are warp-uniform: an out-of-range quad works on a clamped index and is masked out,
        //  because the quad folds below are full-warp shuffles)
        for (int i0 = warp * 8; cls + n_classes * i0 < n_rows; i0 += n_quads) {
            const int mine = cls + n_classes * (i0 + quad_in_warp);
            const bool live = mine < n_rows;
            const int r = live ? mine : n_rows - 1;
            const uint32_t* p = buf + (size_t)r * kTileWords;
            uint32_t x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = p[widx[j]];
            uint32_t s = popc8(x);
            s += __shfl_xor_sync(0xFFFFFFFFu, s, 1);
            s += __shfl_xor_sync(0xFFFFFFFFu, s, 2);
            if (l == 0 && live) cta_acc[r] += s;
        }
        // ---- the multi-variable terms: one quad per group, specialised by arity ----
        // groups are visited in arity order (3, 2, 1, 0 parents), so a warp's eight groups almost
        // always share one code path; a mixed round runs the widest path (missing parents read the
        // all-zero row, their terms come out 0).
        for (int i0 = warp * 8; cls + n_classes * i0 < n_groups; i0 += n_quads) {
            const int mine = cls + n_classes * (i0 + quad_in_warp);
            const bool live = mine < n_groups;
            const int g = live ? (int)order[mine] : 0;
            int ra = __ldg(parent_row + 3 * g + 0), rb = __ldg(parent_row + 3 * g + 1),
                rc = __ldg(parent_row + 3 * g + 2), rt = __ldg(target_row + g);
            const int arity = __reduce_max_sync(0xFFFFFFFFu, live ? group_arity(ra, rb, rc) : 0);
            const uint32_t* pa = buf + (size_t)(ra < 0 ? rows_alloc : ra) * kTileWords;
            const uint32_t* pb = buf + (size_t)(rb < 0 ? rows_alloc : rb) * kTileWords;
            const uint32_t* pc = buf + (size_t)(rc < 0 ? rows_alloc : rc) * kTileWords;
            const uint32_t* pt = buf + (size_t)rt * kTileWords;
            uint32_t* dst = cta_acc + n_rows + (size_t)g * kTermsPerGroup;
            if (arity >= 3) group_terms<3>(pa, pb, pc, pt, widx, l, live, dst);
            else if (arity == 2) group_terms<2>(pa, pb, pc, pt, widx, l, live, dst);
            else if (arity == 1) group_terms<1>(pa, pb, pc, pt, widx, l, live, dst);
        }
        __syncthreads();  // every reader is done with this stage
        if (tid == 0) issue_tile(stage, next_tile);
    }
    __syncthreads();

    // fold this CTA's sums into the global int64 accumulators (integer adds: order-free, exact)
    for (int i = tid; i < n_acc; i += blockDim.x)
        if (cta_acc[i]) atomicAdd(acc + i, (unsigned long long)cta_acc[i]);
    __threadfence();
    __syncthreads();
    __shared__ unsigned int s_ticket;
    if (tid == 0) s_ticket = atomicAdd(&tickets[kMaxClasses], 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    // last CTA: every other CTA's adds are visible; invert the subset sums per group
    __threadfence();
    constexpr int kTermMask[kTermsPerGroup] = SCB_TERM_MASKS;
    for (int g = tid; g < n_groups; g += blockDim.x) {
        int ra = parent_row[3 * g], rb = parent_row[3 * g + 1], rc = parent_row[3 * g + 2];
        int rt = target_row[g];
        long long sub[16];  // sub[S] = #cells where every variable in S is 1 (T=8, A=4, B=2, C=1)
#pragma unroll
        for (int i = 0; i < 16; ++i) sub[i] = 0;
        sub[0] = n_cells;
        if (ra >= 0) sub[0x4] = (long long)__ldcg(acc + ra);
        if (rb >= 0) sub[0x2] = (long long)__ldcg(acc + rb);
        if (rc >= 0) sub[0x1] = (long long)__ldcg(acc + rc);
        sub[0x8] = (long long)__ldcg(acc + rt);
        const unsigned long long* terms = acc + n_rows + (size_t)g * kTermsPerGroup;
#pragma unroll
        for (int k = 0; k < kTermsPerGroup; ++k) sub[kTermMask[k]] = (long long)__ldcg(terms + k);
        // Moebius inversion over the 4-variable lattice: exact[S] = #cells whose set of ones is S
#pragma unroll
        for (int bit = 1; bit < 16; bit <<= 1)
#pragma unroll
            for (int q = 0; q < 16; ++q)
                if (!(q & bit)) sub[q] -= sub[q | bit];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            out_counts[(size_t)g * 16 + 2 * k + 0] = sub[k];      // minterm k, target 0
            out_counts[(size_t)g * 16 + 2 * k + 1] = sub[8 | k];  // minterm k, target 1
        }
    }
}

static void box_shape(int n_rows, int* box_rows, int* n_boxes) {
    *n_boxes = (n_rows + 255) / 256;            // a TMA box is at most 256 elements per dimension
    *box_rows = (n_rows + *n_boxes - 1) / *n_boxes;
}

static size_t counts_smem_bytes(int n_rows, int n_groups) {
    int box_rows, n_boxes;
    box_shape(n_rows, &box_rows, &n_boxes);
    size_t n_acc = (size_t)n_rows + (size_t)kTermsPerGroup * n_groups;
    return (size_t)kStages * ((size_t)box_rows * n_boxes + 1) * kTileWords * 4 + ((n_acc + 1) & ~(size_t)1) * 4 +
           kStages * 8 + kStages * 4 + ((size_t)n_groups * 2 + 16);
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

// groups are processed in slices so the per-CTA accumulators always fit in shared memory
static const int kMaxGroupsPerLaunch = 1024;

extern "C" int64_t scb_rule_counts_scratch_bytes(int n_rows, int n_groups) {
    // int64 accumulators (rows + 11 per group of one launch slice) + the two tickets
    if (n_rows < 0 || n_groups < 0) return 0;
    int gn = n_groups < kMaxGroupsPerLaunch ? n_groups : kMaxGroupsPerLaunch;
    return ((int64_t)n_rows + (int64_t)kTermsPerGroup * gn) * 8 + kTicketBytes;
}

extern "C" int scb_rule_counts(const uint32_t* planes, int n_rows, int64_t wpad, int64_t n_cells,
                               int n_groups, const int32_t* target_row, const int32_t* parent_row,
                               void* scratch, int64_t scratch_bytes, int64_t* out_counts,
                               scb_stream_t stream) {
    SCB_REQUIRE(n_rows > 0 && n_rows < 65536, "rule_counts: n_rows=%d out of range", n_rows);
    SCB_REQUIRE(wpad > 0 && wpad % 4 == 0 && wpad * 32 >= n_cells && n_cells >= 0, "rule_counts: wpad must be a positive multiple of 4 covering n_cells");
    SCB_REQUIRE(n_cells < ((int64_t)1 << 31), "rule_counts: at most 2^31-1 cells per call (shard the cells)");
    SCB_REQUIRE(n_groups >= 0, "rule_counts: negative n_groups");
    if (n_groups == 0) return SCB_OK;
    SCB_REQUIRE(planes && target_row && parent_row && scratch && out_counts, "rule_counts: null pointer");
    SCB_REQUIRE(reinterpret_cast<uintptr_t>(planes) % 16 == 0, "rule_counts: planes must be 16-byte aligned");
    SCB_REQUIRE(reinterpret_cast<uintptr_t>(scratch) % 8 == 0, "rule_counts: scratch must be 8-byte aligned");
    const int64_t n_tiles = (wpad + kTileWords - 1) / kTileWords;
    SCB_REQUIRE(n_tiles < ((int64_t)1 << 30), "rule_counts: too many tiles");
    int box_rows, n_boxes;
    box_shape(n_rows, &box_rows, &n_boxes);
    CUtensorMap tmap;
    {
        int rc = make_tensor_map(planes, n_rows, wpad, box_rows, &tmap);
        if (rc != SCB_OK) return rc;
    }
    for (int g0 = 0; g0 < n_groups; g0 += kMaxGroupsPerLaunch) {
        int gn = n_groups - g0 < kMaxGroupsPerLaunch ? n_groups - g0 : kMaxGroupsPerLaunch;
        size_t smem = counts_smem_bytes(n_rows, gn);
        if (smem > 227 * 1024) {
            set_error("rule_counts: %d rows x %d groups do not fit in shared memory", n_rows, gn);
            return SCB_ERR_UNSUPPORTED;
        }
        int64_t n_acc = (int64_t)n_rows + (int64_t)kTermsPerGroup * gn;
        int64_t need = n_acc * 8 + kTicketBytes;
        if (need > scratch_bytes) {
            set_error("rule_counts: scratch too small (%lld < %lld bytes)", (long long)scratch_bytes, (long long)need);
            return SCB_ERR_SCRATCH;
        }
        // attribute + occupancy are queried once per shared-memory size (keeps the enqueue path
        // free of non-stream API calls, e.g. under CUDA-graph capture)
        static thread_local size_t cached_smem = 0;
        static thread_local int cached_per_sm = 0;
        if (smem != cached_smem) {
            SCB_CUDA(cudaFuncSetAttribute(rule_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int q = 1;
            SCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, rule_counts_kernel, kCountThreads, smem));
            cached_per_sm = q < 1 ? 1 : q;
            cached_smem = smem;
        }
        int per_sm = cached_per_sm;
        if (const char* cap = getenv("SCB_COUNTS_CTAS_PER_SM")) {
            int c = atoi(cap);
            if (c >= 1 && c < per_sm) per_sm = c;
        }
        // one class = at most one round of the CTA's quads (8 per warp)
        const int quads = (kCountThreads / 32) * 8;
        int n_classes = 1;
        (void)quads;
        if (const char* cap = getenv("SCB_COUNTS_CLASSES")) {
            int c = atoi(cap);
            if (c >= 1) n_classes = c;
        }
        if (n_classes > kMaxClasses) n_classes = kMaxClasses;
        if (n_classes < 1) n_classes = 1;
        int64_t grid = (int64_t)sm_count() * per_sm;
        if (grid > n_tiles * n_classes) grid = n_tiles * n_classes;
        grid -= grid % n_classes;
        if (grid < n_classes) grid = n_classes;
        SCB_CUDA(cudaMemsetAsync(scratch, 0, (size_t)need, as_stream(stream)));
        auto* acc = static_cast<unsigned long long*>(scratch);
        auto* tickets = reinterpret_cast<unsigned int*>(acc + n_acc);
        rule_counts_kernel<<<(unsigned)grid, kCountThreads, smem, as_stream(stream)>>>(
            tmap, box_rows, n_boxes, n_rows, wpad, n_tiles, gn, target_row + g0, parent_row + 3 * (size_t)g0, acc,
            tickets, n_classes, n_cells, out_counts + (size_t)g0 * 16);
        SCB_LAUNCH_CHECK("rule_counts_kernel");
    }
    return SCB_OK;
}
