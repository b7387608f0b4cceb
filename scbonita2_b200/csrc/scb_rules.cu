// Rule scoring: joint (A,B,C,target) counts per node in ONE pass over the cell bitplanes,
// then mismatch counts of arbitrary truth tables / GA individuals from those counts.
//
// Replaces RuleDetermination.calculate_error (rule_determination.py:509-582): the reference
// evaluates a rule text per column with eval and counts `predicted != target`.  For a rule
// with truth table L over (A,B,C) that count is  sum_k #{cells: (A,B,C)==k, target != L_k},
// so the 16 joint counts of a node are a sufficient statistic for EVERY candidate rule and
// every GA individual; they are pure popcounts of ANDs of the four bitplane rows.
//
// Kernel layout (rule_counts_kernel): persistent CTAs, one cell tile at a time.  The tile
// (all rows x TW words) is staged in shared memory with 128-bit loads; row popcounts are taken
// once per row, and each warp then walks its groups taking the 11 popcounts of the
// multi-variable AND terms (AB, AC, BC, ABC, AT, BT, CT, ABT, ACT, BCT, ABCT).  Lanes read
// consecutive words (bank-conflict free), partial sums are folded with redux.sync and kept in
// per-CTA shared accumulators; one partial block per CTA goes to global memory and a tiny
// finalize kernel adds the CTAs up in int64 and inverts the subset sums (Moebius) into the 16
// minterm x target counts.  No atomics, so results do not depend on scheduling.
#include "scb_common.cuh"

namespace scb {

constexpr int kTermsPerGroup = 11;   // multi-variable AND terms
constexpr int kCountThreads = 512;

struct GroupRows {
    int a, b, c, t;  // smem row indices (unused slot -> the all-zero row)
};

// term order (bit3=T, bit2=A, bit1=B, bit0=C): masks of the 11 multi-variable subsets
#define SCB_TERM_MASKS {0x6 /*AB*/, 0x5 /*AC*/, 0x3 /*BC*/, 0x7 /*ABC*/, 0xC /*AT*/, 0xA /*BT*/, \
                        0x9 /*CT*/, 0xE /*ABT*/, 0xD /*ACT*/, 0xB /*BCT*/, 0xF /*ABCT*/}

__global__ void __launch_bounds__(kCountThreads)
rule_counts_kernel(const uint32_t* __restrict__ planes, int n_rows, int64_t wpad, int64_t n_words,
                   int tile_words, int64_t n_tiles, int n_groups,
                   const int32_t* __restrict__ target_row, const int32_t* __restrict__ parent_row,
                   unsigned long long* __restrict__ acc /*[n_rows + 11*G], zeroed*/,
                   unsigned int* __restrict__ ticket /*zeroed*/, int64_t n_cells,
                   int64_t* __restrict__ out_counts /*[G][16]*/) {
    extern __shared__ __align__(16) uint32_t smem[];
    // layout: tile[(n_rows+1)][tile_words] | row_acc[n_rows] | grp_acc[G][11]
    uint32_t* tile = smem;
    uint32_t* row_acc = tile + (size_t)(n_rows + 1) * tile_words;
    uint32_t* grp_acc = row_acc + n_rows;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int n_warps = blockDim.x >> 5;

    for (int i = tid; i < n_rows + kTermsPerGroup * n_groups; i += blockDim.x) row_acc[i] = 0;
    for (int i = tid; i < tile_words; i += blockDim.x) tile[(size_t)n_rows * tile_words + i] = 0;

    const int vec_per_row = tile_words >> 2;  // uint4 per row of the tile
    for (int64_t tile_idx = blockIdx.x; tile_idx < n_tiles; tile_idx += gridDim.x) {
        const int64_t w0 = tile_idx * tile_words;
        __syncthreads();  // previous tile fully consumed (and the zero fills above are visible)
        // ---- stage the tile: 128-bit coalesced loads, zero past the end of the row ----
        for (int i = tid; i < n_rows * vec_per_row; i += blockDim.x) {
            int row = i / vec_per_row;
            int v = i - row * vec_per_row;
            int64_t w = w0 + 4 * v;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (w < wpad) val = __ldg(reinterpret_cast<const uint4*>(planes + (size_t)row * wpad + w));
            reinterpret_cast<uint4*>(tile + (size_t)row * tile_words)[v] = val;
        }
        __syncthreads();
        // ---- single-variable terms: popcount of every row over the tile ----
        for (int row = warp; row < n_rows; row += n_warps) {
            const uint32_t* r = tile + (size_t)row * tile_words;
            uint32_t s = 0;
            for (int j = lane; j < tile_words; j += 32) s += __popc(r[j]);
            s = __reduce_add_sync(0xFFFFFFFFu, s);
            if (lane == 0) row_acc[row] += s;
        }
        // ---- multi-variable terms per group ----
        for (int g = warp; g < n_groups; g += n_warps) {
            int ra = __ldg(parent_row + 3 * g + 0), rb = __ldg(parent_row + 3 * g + 1),
                rc = __ldg(parent_row + 3 * g + 2), rt = __ldg(target_row + g);
            const uint32_t* pa = tile + (size_t)(ra < 0 ? n_rows : ra) * tile_words;
            const uint32_t* pb = tile + (size_t)(rb < 0 ? n_rows : rb) * tile_words;
            const uint32_t* pc = tile + (size_t)(rc < 0 ? n_rows : rc) * tile_words;
            const uint32_t* pt = tile + (size_t)rt * tile_words;
            uint32_t s[kTermsPerGroup];
#pragma unroll
            for (int k = 0; k < kTermsPerGroup; ++k) s[k] = 0;
            for (int j = lane; j < tile_words; j += 32) {
                uint32_t a = pa[j], b = pb[j], c = pc[j], t = pt[j];
                uint32_t ab = a & b, ac = a & c, bc = b & c;
                uint32_t abc = ab & c;
                s[0] += __popc(ab);
                s[1] += __popc(ac);
                s[2] += __popc(bc);
                s[3] += __popc(abc);
                s[4] += __popc(a & t);
                s[5] += __popc(b & t);
                s[6] += __popc(c & t);
                s[7] += __popc(ab & t);
                s[8] += __popc(ac & t);
                s[9] += __popc(bc & t);
                s[10] += __popc(abc & t);
            }
#pragma unroll
            for (int k = 0; k < kTermsPerGroup; ++k) s[k] = __reduce_add_sync(0xFFFFFFFFu, s[k]);
            if (lane == 0) {
                uint32_t* dst = grp_acc + (size_t)g * kTermsPerGroup;
#pragma unroll
                for (int k = 0; k < kTermsPerGroup; ++k) dst[k] += s[k];
            }
        }
    }
    __syncthreads();
    // fold this CTA's sums into the global int64 accumulators (integer adds: order-free, exact)
    const int n_acc = n_rows + kTermsPerGroup * n_groups;
    for (int i = tid; i < n_acc; i += blockDim.x)
        if (row_acc[i]) atomicAdd(acc + i, (unsigned long long)row_acc[i]);
    __threadfence();
    __syncthreads();
    __shared__ unsigned int s_ticket;
    if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
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

__device__ __forceinline__ long long diff_from_counts(const int64_t* __restrict__ cnt, uint32_t lut) {
    // mismatches = cells of minterm k whose target differs from table bit k
    long long d = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) d += cnt[2 * k + (1 - ((lut >> k) & 1u))];
    return d;
}

__global__ void error_table_kernel(const int64_t* __restrict__ counts, int64_t n_tasks,
                                   const int32_t* __restrict__ task_group,
                                   const uint8_t* __restrict__ task_lut,
                                   int64_t* __restrict__ out_diff) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    out_diff[t] = diff_from_counts(counts + (size_t)task_group[t] * 16, task_lut[t]);
}

// One warp per individual, lanes stride over the groups.  With
//   diff(L) = sum_k c[k][1] + sum_{k: L_k = 1} (c[k][0] - c[k][1])
// the CTA stages base[g] and delta[k][g] in shared memory, laid out [k][g] so that lanes with
// consecutive g read consecutive words (no bank conflicts).
__global__ void __launch_bounds__(512)
population_fitness_kernel(const int64_t* __restrict__ counts, int n_groups, int64_t n_ind,
                          const uint8_t* __restrict__ lut, const uint8_t* __restrict__ active,
                          int64_t* __restrict__ out_diff, int64_t* __restrict__ out_diff_pg) {
    extern __shared__ __align__(16) long long s_tab[];  // base[G] | delta[8][G]
    long long* base = s_tab;
    long long* delta = s_tab + n_groups;
    for (int g = threadIdx.x; g < n_groups; g += blockDim.x) {
        const int64_t* c = counts + (size_t)g * 16;
        long long b = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            b += c[2 * k + 1];
            delta[k * n_groups + g] = c[2 * k] - c[2 * k + 1];
        }
        base[g] = b;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < n_ind; p += n_warps) {
        const uint8_t* row = lut + (size_t)p * n_groups;
        const uint8_t* act = active ? active + (size_t)p * n_groups : nullptr;
        long long total = 0;
        for (int g = lane; g < n_groups; g += 32) {
            long long d = 0;
            if (!act || act[g]) {
                const uint32_t table = row[g];
                d = base[g];
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if ((table >> k) & 1u) d += delta[k * n_groups + g];
            }
            if (out_diff_pg) out_diff_pg[(size_t)p * n_groups + g] = d;
            total += d;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, off);
        if (lane == 0) out_diff[p] = total;
    }
}

// brute force: one warp per task streams the four rows and evaluates the table per word
__global__ void __launch_bounds__(256)
error_table_direct_kernel(const uint32_t* __restrict__ planes, int64_t wpad, int64_t n_cells,
                          int64_t n_tasks,
                          const int32_t* __restrict__ target_row,
                          const int32_t* __restrict__ parent_row, const uint8_t* __restrict__ lut,
                          int64_t* __restrict__ out_diff) {
    const int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_vec = wpad >> 2;
    for (int64_t task = warp; task < n_tasks; task += n_warps) {
        int ra = parent_row[3 * task], rb = parent_row[3 * task + 1], rc = parent_row[3 * task + 2];
        const uint4* pt = reinterpret_cast<const uint4*>(planes + (size_t)target_row[task] * wpad);
        const uint4* pa = ra >= 0 ? reinterpret_cast<const uint4*>(planes + (size_t)ra * wpad) : nullptr;
        const uint4* pb = rb >= 0 ? reinterpret_cast<const uint4*>(planes + (size_t)rb * wpad) : nullptr;
        const uint4* pc = rc >= 0 ? reinterpret_cast<const uint4*>(planes + (size_t)rc * wpad) : nullptr;
        const LutMasks m = expand_lut(lut[task]);
        const uint4 zero = make_uint4(0, 0, 0, 0);
        long long diff = 0;
        for (int64_t v = lane; v < n_vec; v += 32) {
            uint4 a = pa ? __ldg(pa + v) : zero, b = pb ? __ldg(pb + v) : zero,
                  c = pc ? __ldg(pc + v) : zero, t = __ldg(pt + v);
            diff += __popc(eval_lut(m, a.x, b.x, c.x) ^ t.x) + __popc(eval_lut(m, a.y, b.y, c.y) ^ t.y) +
                    __popc(eval_lut(m, a.z, b.z, c.z) ^ t.z) + __popc(eval_lut(m, a.w, b.w, c.w) ^ t.w);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) diff += __shfl_xor_sync(0xFFFFFFFFu, diff, off);
        // padding cells (every row zero) predict table bit 0 against target 0: not real cells
        if (lane == 0) out_diff[task] = diff - (long long)(lut[task] & 1u) * (wpad * 32 - n_cells);
    }
}

struct CountsPlan {
    int grid, tile_words;
    int64_t n_tiles;
    size_t smem;
};

static int plan_counts(int n_rows, int64_t wpad, int n_groups, CountsPlan* plan) {
    // largest tile (multiple of 32 words) that leaves room for two CTAs per SM
    const size_t budget_two = 110 * 1024, budget_one = 220 * 1024;
    size_t fixed = ((size_t)n_rows + (size_t)kTermsPerGroup * n_groups) * 4;
    int tw = 0;
    for (int cand : {256, 128, 64, 32}) {
        size_t need = fixed + (size_t)(n_rows + 1) * cand * 4;
        if (need <= budget_two) { tw = cand; break; }
    }
    if (!tw) {
        for (int cand : {128, 64, 32}) {
            size_t need = fixed + (size_t)(n_rows + 1) * cand * 4;
            if (need <= budget_one) { tw = cand; break; }
        }
    }
    if (!tw) {
        set_error("rule_counts: %d rows x %d groups do not fit in shared memory", n_rows, n_groups);
        return SCB_ERR_UNSUPPORTED;
    }
    plan->tile_words = tw;
    plan->n_tiles = (wpad + tw - 1) / tw;
    plan->smem = fixed + (size_t)(n_rows + 1) * tw * 4;
    int per_sm = plan->smem <= budget_two ? 2 : 1;
    int64_t g = (int64_t)sm_count() * per_sm;
    if (g > plan->n_tiles) g = plan->n_tiles;
    if (g < 1) g = 1;
    plan->grid = (int)g;
    return SCB_OK;
}

}  // namespace scb

using namespace scb;

// groups are processed in slices so the per-CTA accumulators always fit in shared memory
static const int kMaxGroupsPerLaunch = 1024;

extern "C" int64_t scb_rule_counts_scratch_bytes(int n_rows, int n_groups) {
    // int64 accumulators (rows + 11 per group of one launch slice) + the CTA ticket
    if (n_rows < 0 || n_groups < 0) return 0;
    int gn = n_groups < kMaxGroupsPerLaunch ? n_groups : kMaxGroupsPerLaunch;
    return ((int64_t)n_rows + (int64_t)kTermsPerGroup * gn) * 8 + 16;
}

extern "C" int scb_rule_counts(const uint32_t* planes, int n_rows, int64_t wpad, int64_t n_cells,
                               int n_groups, const int32_t* target_row, const int32_t* parent_row,
                               void* scratch, int64_t scratch_bytes, int64_t* out_counts,
                               scb_stream_t stream) {
    SCB_REQUIRE(n_rows > 0 && n_rows < 65536, "rule_counts: n_rows=%d out of range", n_rows);
    SCB_REQUIRE(wpad > 0 && wpad % 4 == 0 && wpad * 32 >= n_cells && n_cells >= 0, "rule_counts: wpad must be a positive multiple of 4 covering n_cells");
    SCB_REQUIRE(n_groups >= 0, "rule_counts: negative n_groups");
    if (n_groups == 0) return SCB_OK;
    SCB_REQUIRE(planes && target_row && parent_row && scratch && out_counts, "rule_counts: null pointer");
    SCB_REQUIRE(reinterpret_cast<uintptr_t>(planes) % 16 == 0, "rule_counts: planes must be 16-byte aligned");
    int64_t n_words = (n_cells + 31) / 32;
    for (int g0 = 0; g0 < n_groups; g0 += kMaxGroupsPerLaunch) {
        int gn = n_groups - g0 < kMaxGroupsPerLaunch ? n_groups - g0 : kMaxGroupsPerLaunch;
        CountsPlan plan;
        int rc = plan_counts(n_rows, wpad, gn, &plan);
        if (rc != SCB_OK) return rc;
        int64_t n_acc = (int64_t)n_rows + (int64_t)kTermsPerGroup * gn;
        int64_t need = n_acc * 8 + 16;
        if (need > scratch_bytes) {
            set_error("rule_counts: scratch too small (%lld < %lld bytes)", (long long)scratch_bytes, (long long)need);
            return SCB_ERR_SCRATCH;
        }
        SCB_CUDA(cudaFuncSetAttribute(rule_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
        SCB_CUDA(cudaMemsetAsync(scratch, 0, (size_t)need, as_stream(stream)));
        auto* acc = static_cast<unsigned long long*>(scratch);
        auto* ticket = reinterpret_cast<unsigned int*>(acc + n_acc);
        rule_counts_kernel<<<plan.grid, kCountThreads, plan.smem, as_stream(stream)>>>(
            planes, n_rows, wpad, n_words, plan.tile_words, plan.n_tiles, gn, target_row + g0,
            parent_row + 3 * (size_t)g0, acc, ticket, n_cells, out_counts + (size_t)g0 * 16);
        SCB_LAUNCH_CHECK("rule_counts_kernel");
    }
    return SCB_OK;
}

extern "C" int scb_rule_error_table(const int64_t* counts, int n_groups, int64_t n_tasks,
                                    const int32_t* task_group, const uint8_t* task_lut,
                                    int64_t* out_diff, scb_stream_t stream) {
    SCB_REQUIRE(n_groups >= 0 && n_tasks >= 0, "rule_error_table: negative size");
    if (n_tasks == 0) return SCB_OK;
    SCB_REQUIRE(counts && task_group && task_lut && out_diff, "rule_error_table: null pointer");
    error_table_kernel<<<(unsigned)((n_tasks + 255) / 256), 256, 0, as_stream(stream)>>>(
        counts, n_tasks, task_group, task_lut, out_diff);
    SCB_LAUNCH_CHECK("error_table_kernel");
    return SCB_OK;
}

extern "C" int scb_population_fitness(const int64_t* counts, int n_groups, int64_t n_individuals,
                                      const uint8_t* lut, const uint8_t* active, int64_t* out_diff,
                                      int64_t* out_diff_pg, scb_stream_t stream) {
    SCB_REQUIRE(n_groups > 0 && n_individuals >= 0, "population_fitness: bad size");
    if (n_individuals == 0) return SCB_OK;
    SCB_REQUIRE(counts && lut && out_diff, "population_fitness: null pointer");
    size_t smem = (size_t)n_groups * 9 * sizeof(long long);
    if (smem > 200 * 1024) {
        set_error("population_fitness: %d groups exceed the shared-memory table (max %d)", n_groups, (int)(200 * 1024 / 72));
        return SCB_ERR_UNSUPPORTED;
    }
    if (smem > 48 * 1024)
        SCB_CUDA(cudaFuncSetAttribute(population_fitness_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (n_individuals + 15) / 16;  // 16 warps per CTA, one individual per warp
    int64_t cap = (int64_t)sm_count() * 2;
    if (blocks > cap) blocks = cap;
    population_fitness_kernel<<<(unsigned)blocks, 512, smem, as_stream(stream)>>>(
        counts, n_groups, n_individuals, lut, active, out_diff, out_diff_pg);
    SCB_LAUNCH_CHECK("population_fitness_kernel");
    return SCB_OK;
}

extern "C" int scb_rule_error_table_direct(const uint32_t* planes, int n_rows, int64_t wpad,
                                           int64_t n_cells, int64_t n_tasks,
                                           const int32_t* target_row, const int32_t* parent_row,
                                           const uint8_t* lut, int64_t* out_diff,
                                           scb_stream_t stream) {
    SCB_REQUIRE(n_rows > 0 && wpad > 0 && wpad % 4 == 0 && wpad * 32 >= n_cells, "rule_error_table_direct: bad plane shape");
    if (n_tasks == 0) return SCB_OK;
    SCB_REQUIRE(planes && target_row && parent_row && lut && out_diff, "rule_error_table_direct: null pointer");
    int64_t blocks = (n_tasks + 7) / 8;
    int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    error_table_direct_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
        planes, wpad, n_cells, n_tasks, target_row, parent_row, lut, out_diff);
    SCB_LAUNCH_CHECK("error_table_direct_kernel");
    return SCB_OK;
}
