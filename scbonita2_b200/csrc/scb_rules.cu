// Rule scoring, second half: mismatch counts of arbitrary truth tables / GA individuals from the
// joint counts produced by scb_rule_counts (scb_counts.cu), plus the brute-force cross-check.
//
// With counts[g][2k+t] = #cells of group g whose parents read minterm k and whose target is t, the
// reference's (difference, count) of a rule with truth table L (rule_determination.py:577-580) is
// difference = sum_k counts[g][2k + (1 - L_k)], count = n_cells.
#include "scb_common.cuh"

namespace scb {

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
// the CTA stages base[g] and, per table nibble, the 16 partial sums of the deltas:
//   diff(L) = base[g] + lo[L & 15][g] + hi[L >> 4][g]            (2 shared loads per pair)
// laid out [nibble][g] so lanes with consecutive g mostly hit distinct banks.  The nibble sums are
// kept in int32 when every count fits (always, below 2^31 cells); otherwise the CTA falls back to
// the eight int64 deltas.
__global__ void __launch_bounds__(512)
population_fitness_kernel(const int64_t* __restrict__ counts, int n_groups, int64_t n_ind,
                          const uint8_t* __restrict__ lut, const uint8_t* __restrict__ active,
                          int64_t* __restrict__ out_diff, int64_t* __restrict__ out_diff_pg) {
    extern __shared__ __align__(16) long long s_tab[];  // base[G] | delta[8][G] (int64)  or  lo/hi[32][G] (int32)
    long long* base = s_tab;
    long long* delta = s_tab + n_groups;
    int* nib = reinterpret_cast<int*>(s_tab + n_groups);  // aliases delta: only one form is built
    int wide = 0;
    for (int g = threadIdx.x; g < n_groups; g += blockDim.x) {
        const int64_t* c = counts + (size_t)g * 16;
#pragma unroll
        for (int k = 0; k < 16; ++k) wide |= (int)((unsigned long long)c[k] >> 31 != 0);
    }
    wide = __syncthreads_or(wide);
    for (int g = threadIdx.x; g < n_groups; g += blockDim.x) {
        const int64_t* c = counts + (size_t)g * 16;
        long long b = 0, d[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            b += c[2 * k + 1];
            d[k] = c[2 * k] - c[2 * k + 1];
        }
        base[g] = b;
        if (wide) {
#pragma unroll
            for (int k = 0; k < 8; ++k) delta[k * n_groups + g] = d[k];
        } else {
#pragma unroll
            for (int v = 0; v < 16; ++v) {
                long long lo = 0, hi = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if ((v >> k) & 1) {
                        lo += d[k];
                        hi += d[k + 4];
                    }
                nib[v * n_groups + g] = (int)lo;
                nib[(16 + v) * n_groups + g] = (int)hi;
            }
        }
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
                if (wide) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if ((table >> k) & 1u) d += delta[k * n_groups + g];
                } else {
                    d += nib[(table & 15u) * n_groups + g] + nib[(16u + (table >> 4)) * n_groups + g];
                }
            }
            if (out_diff_pg) out_diff_pg[(size_t)p * n_groups + g] = d;
            total += d;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, off);
        if (lane == 0) out_diff[p] = total;
    }
}

// one warp per multi-rule individual: masked sum of the per-rule mismatch table
__global__ void __launch_bounds__(256)
bitstring_fitness_kernel(const int64_t* __restrict__ task_diff, int64_t n_tasks, int64_t n_ind,
                         const uint8_t* __restrict__ bits, int64_t* __restrict__ out_diff,
                         int32_t* __restrict__ out_nsel) {
    const int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < n_ind; p += n_warps) {
        const uint8_t* row = bits + (size_t)p * n_tasks;
        long long total = 0;
        int nsel = 0;
        for (int64_t j = lane; j < n_tasks; j += 32)
            if (row[j]) {
                total += task_diff[j];
                ++nsel;
            }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            total += __shfl_xor_sync(0xFFFFFFFFu, total, off);
            nsel += __shfl_xor_sync(0xFFFFFFFFu, nsel, off);
        }
        if (lane == 0) {
            out_diff[p] = total;
            out_nsel[p] = nsel;
        }
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

}  // namespace scb

using namespace scb;

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
    size_t smem = (size_t)n_groups * 17 * sizeof(long long);  // base + max(8 int64 deltas, 32 int32 nibble sums)
    if (smem > 200 * 1024) {
        set_error("population_fitness: %d groups exceed the shared-memory table (max %d)", n_groups, (int)(200 * 1024 / 136));
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

extern "C" int scb_bitstring_fitness(const int64_t* task_diff, int64_t n_tasks, int64_t n_individuals,
                                     const uint8_t* bits, int64_t* out_diff, int32_t* out_nsel,
                                     scb_stream_t stream) {
    SCB_REQUIRE(n_tasks >= 0 && n_individuals >= 0, "bitstring_fitness: negative size");
    if (n_individuals == 0) return SCB_OK;
    SCB_REQUIRE(task_diff && bits && out_diff && out_nsel, "bitstring_fitness: null pointer");
    int64_t blocks = (n_individuals + 7) / 8;
    int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    bitstring_fitness_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(task_diff, n_tasks, n_individuals,
                                                                             bits, out_diff, out_nsel);
    SCB_LAUNCH_CHECK("bitstring_fitness_kernel");
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
