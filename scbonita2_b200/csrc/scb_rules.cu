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

// thread-block cluster helpers (population_fitness_kernel with sharded cells)
__device__ __forceinline__ uint32_t cluster_cta_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_size() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t cluster_map(const void* p, uint32_t rank) {
    uint32_t local = (uint32_t)__cvta_generic_to_shared(p), remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
    return remote;
}
__device__ __forceinline__ void st_cluster_u64(uint32_t addr, unsigned long long v) {
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void error_table_kernel(const int64_t* __restrict__ counts, int64_t n_tasks,
                                   const int32_t* __restrict__ task_group,
                                   const uint8_t* __restrict__ task_lut,
                                   int64_t* __restrict__ out_diff) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    out_diff[t] = diff_from_counts(counts + (size_t)task_group[t] * 16, task_lut[t]);
}

// Input of the fitness kernel when it consumes the subset sums of scb_rule_sums_run: mode 1 reads the
// local accumulators through the plan's descriptors (and its last CTA zeroes them for the next pass),
// mode 2 sums the ranks' slots of the exchange region, which then hold subset sums [G][16].
struct SumsInput {
    int mode;
    const unsigned char* plan;
    unsigned long long* acc;
    unsigned int* ticket;
    int n_rows, zero_row;
    long long n_cells;
};

// How long a kernel waits for a peer's tagged word before it gives up (PeerHeader.timeout_ms, 0 = 30 s).
// A kernel that gave up raises PeerHeader.error AND poisons every result it writes (INT64_MIN): a rank
// that was merely slow can never leave plausible but wrong numbers behind.
__device__ __forceinline__ unsigned long long peer_timeout_ns(const unsigned char* region) {
    const unsigned int ms = reinterpret_cast<const PeerHeader*>(region)->timeout_ms;
    return (unsigned long long)(ms ? ms : 30000u) * 1000000ull;
}

// One warp per individual.  With
//   diff(L) = sum_k c[k][1] + sum_{k: L_k = 1} (c[k][0] - c[k][1])
// the CTA stages base[g] and, per table nibble, the 16 partial sums of the deltas:
//   diff(L) = base[g] + lo[L & 15][g] + hi[L >> 4][g]            (2 shared loads per pair)
// laid out [nibble][g].  The nibble sums are kept in int32 when every count fits (always, below
// 2^31 cells); otherwise the CTA falls back to the eight int64 deltas.
// Launched as a programmatic dependent of the kernel before it in the stream (rule_counts_kernel
// in the GA step): its launch latency hides behind that kernel's tail; griddepcontrol.wait comes
// before the first global read.
#ifdef SCB_TRACE
__device__ unsigned long long g_pdl_b[64][4];  // globaltimer of CTA 0: entry, after griddepcontrol.wait, end
__device__ unsigned int g_pdl_b_n;
__device__ unsigned long long g_fit_phase[16];  // globaltimer of CTA 0 at the phase boundaries of the last launch
#define SCB_FIT_AT(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_fit_phase[i] = global_timer_ns(); } while (0)
#else
#define SCB_FIT_AT(i) do { } while (0)
#endif
__global__ void __launch_bounds__(1024)
population_fitness_kernel(const int64_t* __restrict__ counts, int n_groups, int64_t n_ind,
                          const uint8_t* __restrict__ lut, const uint8_t* __restrict__ active,
                          int64_t* __restrict__ out_diff, int64_t* __restrict__ out_diff_pg,
                          const unsigned char* region /*peer exchange region or NULL*/, int rank, int n_ranks,
                          const SumsInput sums /*sums.mode != 0: subset sums instead of counts (scb_rule_sums_run)*/,
                          int lut_off /*> 0: byte offset of the staged truth tables in shared memory*/,
                          int chunk /*individuals per CTA (contiguous) when the tables are staged*/) {
    extern __shared__ __align__(16) long long s_tab[];  // base[G] | delta[8][G] (int64) | lo/hi[32][G] (int32)
    long long* base = s_tab;
    long long* delta = s_tab + n_groups;
    int* nib = reinterpret_cast<int*>(s_tab + 9 * (size_t)n_groups);
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // Staged truth tables: the CTA owns `chunk` consecutive individuals and copies their rows into shared memory
    // BEFORE it waits for the kernel ahead of it (which never writes them) -- the copy hides behind that kernel's
    // tail, and the population loop below reads bytes from shared memory instead of waiting one L2 round trip
    // per individual.  That keeps the kernel short with FEW CTAs (four or more individuals per warp), which is
    // what lets the next pass's counting kernel start beside it on the SMs it leaves free.
    const int64_t c0 = (int64_t)blockIdx.x * chunk;
    const int64_t c1 = c0 + chunk < n_ind ? c0 + chunk : n_ind;
    uint8_t* s_lut = reinterpret_cast<uint8_t*>(s_tab) + lut_off;
    if (lut_off > 0 && c0 < c1) {
        const uint8_t* src = lut + (size_t)c0 * n_groups;
        const int64_t n_bytes = (c1 - c0) * n_groups;
        if (((reinterpret_cast<uintptr_t>(src) | (uintptr_t)n_bytes) & 15) == 0) {
            for (int64_t i = threadIdx.x; i < n_bytes / 16; i += blockDim.x)
                reinterpret_cast<uint4*>(s_lut)[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
        } else {
            for (int64_t i = threadIdx.x; i < n_bytes; i += blockDim.x) s_lut[i] = __ldg(src + i);
        }
    }
    SCB_FIT_AT(0);
#ifdef SCB_TRACE
    __shared__ unsigned int s_pdl_i;
    if (threadIdx.x == 0 && blockIdx.x == 0) { s_pdl_i = atomicAdd(&g_pdl_b_n, 1u) & 63u; g_pdl_b[s_pdl_i][0] = global_timer_ns(); }
#endif
    asm volatile("griddepcontrol.wait;" ::: "memory");       // the producers' writes are visible from here on
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    SCB_FIT_AT(1);
#ifdef SCB_TRACE
    if (threadIdx.x == 0 && blockIdx.x == 0) g_pdl_b[s_pdl_i][1] = global_timer_ns();
#endif
    // Sharded cells: the counts are the sum of every rank's slot in the local exchange region, tagged
    // words {epoch + 1 : count}.  The rank's own slot (written earlier in this stream) names the epoch.
    const unsigned long long* slots = nullptr;
    unsigned int want = 0;
    if (region) {
        __shared__ unsigned int s_want;
        if (threadIdx.x == 0) {
            const unsigned long long* own = reinterpret_cast<const unsigned long long*>(region + kPeerSlotsOffset) +
                                            (int64_t)rank * n_groups * 16;
            const unsigned int e0 = (unsigned int)(ld_peer(own) >> 32);
            const unsigned int e1 = (unsigned int)(ld_peer(own + (int64_t)n_ranks * n_groups * 16) >> 32);
            s_want = e0 > e1 ? e0 : e1;
        }
        __syncthreads();
        want = s_want;
        slots = reinterpret_cast<const unsigned long long*>(region + kPeerSlotsOffset) +
                (int64_t)((want - 1u) & 1u) * n_ranks * n_groups * 16;
    }
    // the first individual's truth tables are requested before the tables are built from the counts
    constexpr int kMaxRounds = 8;  // register-prefetched rows: up to 256 groups
    const bool fast = n_groups <= 32 * kMaxRounds && !active && !out_diff_pg;
    uint32_t tab[kMaxRounds];
    if (fast && lut_off == 0 && warp0 < n_ind) {
#pragma unroll
        for (int j = 0; j < kMaxRounds; ++j) {
            const int g = lane + 32 * j;
            tab[j] = g < n_groups ? (uint32_t)__ldg(lut + (size_t)warp0 * n_groups + g) : 0u;
        }
    }
    // one thread per (group, minterm): the two counts of the minterm in (one 16-byte load per rank),
    // the delta out; the target-1 counts are parked in the nibble area until base[] is summed
    // Sharded cells: summing the slots of R ranks costs every CTA R x 25.6 KB of L2 reads.  The CTAs
    // of a thread-block cluster split the (group, minterm) pairs among them and store each sum into
    // the shared memory of ALL of them (distributed shared memory), so a CTA reads 1/C of the slots.
    // Without a cluster attribute the cluster is this CTA alone and the same code runs.
    const uint32_t c_rank = cluster_cta_rank(), c_size = cluster_size();
    __shared__ unsigned int s_wide, s_timeout;
    if (threadIdx.x == 0) s_wide = s_timeout = 0u;
    const unsigned long long give_up_ns = region ? peer_timeout_ns(region) : 0ull;
    if (c_size > 1) cluster_sync_all();  // every CTA of the cluster is running (and s_wide is cleared) before a peer stores into it
    int wide = 0;
    unsigned int zero_ticket = 0u;  // thread 0, sums on one GPU: how many CTAs had read the accumulators before this one
    long long* ones = reinterpret_cast<long long*>(nib);
    const int sums_dbg = sums.mode >> 4;   // experiments: 1 = no ticket / zeroing, 2 = no loads
    if ((sums.mode & 15) != 0) {
        // ---- subset sums in: sixteen lanes per group gather (or sum over the ranks) the 16 sums, invert
        // them through shuffles, and lanes 0..7 write the minterm's delta and target-1 count ----
        // one GPU: the accumulators (n_rows + 11 G words, ~19 KB) are first copied into shared memory with
        // coalesced 16-byte loads -- 128 CTAs gathering single words straight from L2 make five times the
        // requests on the same 150 cache lines and take 6 us longer (measured)
        const int n_acc = sums.n_rows + kTermsPerGroup * n_groups;
        unsigned long long* s_acc = reinterpret_cast<unsigned long long*>(s_tab + ((25 * (size_t)n_groups + 1) & ~(size_t)1));  // [n_acc | 1], 16-byte aligned
        uint4* s_desc = reinterpret_cast<uint4*>(s_acc + ((n_acc + 1) & ~1));                               // [G]
        unsigned short* s_order = reinterpret_cast<unsigned short*>(s_desc + n_groups);                     // [G]
        if ((sums.mode & 15) == 1) {
            const PlanLayout L = plan_layout(sums.n_rows, n_groups);
            const uint4* g_desc = reinterpret_cast<const uint4*>(sums.plan + L.desc);
            const unsigned short* g_order = reinterpret_cast<const unsigned short*>(sums.plan + L.order);
            for (int i = threadIdx.x; i < n_groups; i += blockDim.x) {
                s_desc[i] = __ldg(g_desc + i);
                s_order[i] = __ldg(g_order + i);
            }
            if (!(sums_dbg & 2)) {
                const ulonglong2* src = reinterpret_cast<const ulonglong2*>(sums.acc);
                for (int i = threadIdx.x; i < n_acc / 2; i += blockDim.x) {
                    const ulonglong2 v2 = __ldcg(src + i);
                    s_acc[2 * i] = v2.x;
                    s_acc[2 * i + 1] = v2.y;
                }
                if ((n_acc & 1) && threadIdx.x == 0) s_acc[n_acc - 1] = __ldcg(sums.acc + n_acc - 1);
            }
        }
        __syncthreads();  // (also: s_wide / s_timeout are cleared)
        SCB_FIT_AT(2);
        // Every read of the global accumulators by this CTA is done (the values sit in shared memory): the last
        // CTA to say so zeroes them, and the ticket, for the next pass.  The ticket is a plain relaxed add whose
        // answer is only looked at after the tables are built (below): an acquire-release add here cost the CTA
        // ~2.5 us of fence and round trip in front of the inversion (tools/trace_fit.py).
        if ((sums.mode & 15) == 1 && !(sums_dbg & 1) && threadIdx.x == 0) zero_ticket = atomicAdd(sums.ticket, 1u);
        const int q = threadIdx.x & 15;
        constexpr int kBatch = 4;  // the loads of four rounds fly together (two dependent L2 trips in all)
        const int per_round = blockDim.x >> 4;
        for (int pos0 = 0; pos0 < n_groups; pos0 += kBatch * per_round) {  // uniform trip count: shuffles inside
            long long v[kBatch];
            int g[kBatch];
            bool live[kBatch];
            if ((sums.mode & 15) == 1) {
                uint4 d[kBatch];
#pragma unroll
                for (int k = 0; k < kBatch; ++k) {
                    const int pos = pos0 + k * per_round + (threadIdx.x >> 4);
                    live[k] = pos < n_groups;
                    d[k] = s_desc[live[k] ? pos : 0];
                    g[k] = (int)s_order[live[k] ? pos : 0];
                }
                int idx[kBatch];
#pragma unroll
                for (int k = 0; k < kBatch; ++k)
                    idx[k] = (live[k] && !(sums_dbg & 2)) ? subset_sum_index(d[k], g[k], q, sums.n_rows, sums.zero_row) : -1;
#pragma unroll
                for (int k = 0; k < kBatch; ++k) v[k] = idx[k] >= 0 ? (long long)s_acc[idx[k]] : 0ll;
#pragma unroll
                for (int k = 0; k < kBatch; ++k)
                    if (q == 0 && live[k]) v[k] = sums.n_cells;
                SCB_FIT_AT(8);
            } else {
                // (one round at a time: the eight ranks' loads of a word already fly together)
#pragma unroll 1
                for (int k = 0; k < kBatch; ++k) {
                    g[k] = pos0 + k * per_round + (threadIdx.x >> 4);
                    live[k] = g[k] < n_groups;
                    v[k] = 0;
                    if (!live[k]) continue;
                    unsigned long long t[SCB_MAX_PEERS];
#pragma unroll
                    for (int r = 0; r < SCB_MAX_PEERS; ++r)
                        if (r < n_ranks) t[r] = ld_peer(slots + ((int64_t)r * n_groups + g[k]) * 16 + q);
#pragma unroll
                    for (int r = 0; r < SCB_MAX_PEERS; ++r) {
                        if (r >= n_ranks) continue;
                        if ((unsigned int)(t[r] >> 32) != want) {
                            // the peer's word has not landed yet: spin, but never hang
                            const unsigned long long* w = slots + ((int64_t)r * n_groups + g[k]) * 16 + q;
                            const unsigned long long t0 = global_timer_ns();
                            do {
                                t[r] = ld_peer(w);
                                if (global_timer_ns() - t0 > give_up_ns) {
                                    reinterpret_cast<PeerHeader*>(const_cast<unsigned char*>(region))->error = 1u;
                                    s_timeout = 1u;
                                    break;
                                }
                            } while ((unsigned int)(t[r] >> 32) != want);
                        }
                        v[k] += (long long)(unsigned int)t[r];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < kBatch; ++k) {
                const long long x = moebius16(v[k], q);
                const long long y = __shfl_down_sync(0xFFFFFFFFu, x, 8);  // lane q < 8: (minterm q, target 1)
                if (live[k] && q < 8) {
                    if (((unsigned long long)(x + y) >> 28) != 0) s_wide = 1u;
                    delta[q * n_groups + g[k]] = x - y;
                    ones[q * n_groups + g[k]] = y;
                }
            }
            SCB_FIT_AT(9);
        }
    } else
    for (int e = threadIdx.x * (int)c_size + (int)c_rank; e < n_groups * 8; e += blockDim.x * (int)c_size) {
        const int g = e >> 3, k = e & 7;
        longlong2 v;  // (minterm k, target 0), (minterm k, target 1)
        if (!slots) {
            v = __ldcg(reinterpret_cast<const longlong2*>(counts + (size_t)g * 16) + k);
        } else {
            v = make_longlong2(0, 0);
            // the loads of all ranks fly together; only a word whose tag is stale is polled
            ulonglong2 t[SCB_MAX_PEERS];
#pragma unroll
            for (int r = 0; r < SCB_MAX_PEERS; ++r)
                if (r < n_ranks) t[r] = ld_peer2(slots + ((int64_t)r * n_groups + g) * 16 + 2 * k);
#pragma unroll
            for (int r = 0; r < SCB_MAX_PEERS; ++r) {
                if (r >= n_ranks) continue;
                if ((unsigned int)(t[r].x >> 32) != want || (unsigned int)(t[r].y >> 32) != want) {
                    // the peer's word has not landed yet: spin, but never hang
                    const unsigned long long* w = slots + ((int64_t)r * n_groups + g) * 16 + 2 * k;
                    const unsigned long long t0 = global_timer_ns();
                    do {
                        t[r] = ld_peer2(w);
                        if (global_timer_ns() - t0 > give_up_ns) {
                            reinterpret_cast<PeerHeader*>(const_cast<unsigned char*>(region))->error = 1u;
                            for (uint32_t cr = 0; cr < c_size; ++cr) st_cluster_u32(cluster_map(&s_timeout, cr), 1u);
                            break;
                        }
                    } while ((unsigned int)(t[r].x >> 32) != want || (unsigned int)(t[r].y >> 32) != want);
                }
                v.x += (long long)(unsigned int)t[r].x;
                v.y += (long long)(unsigned int)t[r].y;
            }
        }
        // a nibble entry sums up to four deltas and the eight minterms share the cells: with every
        // (minterm, either target) count below 2^28 no entry can leave int32
        const int w = (int)(((unsigned long long)(v.x + v.y) >> 28) != 0);
        wide |= w;
        if (c_size == 1) {
            delta[k * n_groups + g] = v.x - v.y;
            ones[k * n_groups + g] = v.y;
        } else {
            for (uint32_t r = 0; r < c_size; ++r) {
                st_cluster_u64(cluster_map(&delta[k * n_groups + g], r), (unsigned long long)(v.x - v.y));
                st_cluster_u64(cluster_map(&ones[k * n_groups + g], r), (unsigned long long)v.y);
                if (w) st_cluster_u32(cluster_map(&s_wide, r), 1u);
            }
        }
    }
    if (c_size > 1) cluster_sync_all();  // release / acquire: the peers' stores are visible from here on
    else __syncthreads();
    SCB_FIT_AT(3);
    wide |= (int)s_wide;
    const bool poisoned = s_timeout != 0u;
    for (int g = threadIdx.x; g < n_groups; g += blockDim.x) {
        long long b = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) b += ones[k * n_groups + g];
        base[g] = b;
    }
    wide = __syncthreads_or(wide);
    SCB_FIT_AT(4);
    if (!wide) {
        // warp w builds row w of the nibble tables (0..15 low nibble, 16..31 high nibble) from the
        // deltas in shared memory, lanes striding over the groups
        for (int what = threadIdx.x >> 5; what < 32; what += blockDim.x >> 5) {
            const int k0 = (what >> 4) * 4, v = what & 15;
            for (int g = lane; g < n_groups; g += 32) {
                long long sum = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if ((v >> k) & 1) sum += delta[(k0 + k) * n_groups + g];
                nib[what * n_groups + g] = (int)sum;
            }
        }
        __syncthreads();
    }
    SCB_FIT_AT(5);
    if ((sums.mode & 15) == 1 && !(sums.mode >> 4 & 1) && threadIdx.x < 32) {
        const unsigned int t = __shfl_sync(0xFFFFFFFFu, zero_ticket, 0);
        if (t == gridDim.x - 1) {
            const int n_acc_all = sums.n_rows + kTermsPerGroup * n_groups;
            for (int i = threadIdx.x; i < n_acc_all; i += 32) sums.acc[i] = 0ull;
            if (threadIdx.x == 0) *sums.ticket = 0u;
        }
    }
    if (lut_off > 0) {
        // staged tables (launch_population_fitness only stages when the fast conditions hold)
        for (int64_t p = c0 + (threadIdx.x >> 5); p < c1; p += blockDim.x >> 5) {
            const uint8_t* row = s_lut + (size_t)(p - c0) * n_groups;
            long long total = 0;
            for (int g = lane; g < n_groups; g += 32) {
                const uint32_t table = row[g];
                long long d = base[g];
                if (wide) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if ((table >> k) & 1u) d += delta[k * n_groups + g];
                } else {
                    d += nib[(table & 15u) * n_groups + g] + nib[(16u + (table >> 4)) * n_groups + g];
                }
                total += d;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, off);
            if (lane == 0) out_diff[p] = poisoned ? (long long)0x8000000000000000ull : total;
        }
        SCB_FIT_AT(6);
#ifdef SCB_TRACE
        if (threadIdx.x == 0 && blockIdx.x == 0) g_pdl_b[s_pdl_i][2] = global_timer_ns();
#endif
        return;
    }
    for (int64_t p = warp0; p < n_ind; p += n_warps) {
        long long total = 0;
        if (fast && !wide) {
            uint32_t next[kMaxRounds];
            const bool more = p + n_warps < n_ind;
#pragma unroll
            for (int j = 0; j < kMaxRounds; ++j) {
                const int g = lane + 32 * j;
                next[j] = (more && g < n_groups) ? (uint32_t)__ldg(lut + (size_t)(p + n_warps) * n_groups + g) : 0u;
            }
#pragma unroll
            for (int j = 0; j < kMaxRounds; ++j) {
                const int g = lane + 32 * j;
                if (g < n_groups)
                    total += base[g] + nib[(tab[j] & 15u) * n_groups + g] + nib[(16u + (tab[j] >> 4)) * n_groups + g];
            }
#pragma unroll
            for (int j = 0; j < kMaxRounds; ++j) tab[j] = next[j];
        } else {
            const uint8_t* row = lut + (size_t)p * n_groups;
            const uint8_t* act = active ? active + (size_t)p * n_groups : nullptr;
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
                if (out_diff_pg) out_diff_pg[(size_t)p * n_groups + g] = poisoned ? (long long)0x8000000000000000ull : d;
                total += d;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, off);
        if (lane == 0) out_diff[p] = poisoned ? (long long)0x8000000000000000ull : total;
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

#ifdef SCB_TRACE
extern "C" int scb_debug_pdl_b(unsigned long long* out /*HOST [64][4]*/) {
    SCB_CUDA(cudaDeviceSynchronize());
    SCB_CUDA(cudaMemcpyFromSymbol(out, g_pdl_b, sizeof(unsigned long long) * 64 * 4));
    return SCB_OK;
}
extern "C" int scb_debug_fit_phases(unsigned long long* out /*HOST [16]*/) {
    SCB_CUDA(cudaDeviceSynchronize());
    SCB_CUDA(cudaMemcpyFromSymbol(out, g_fit_phase, sizeof(unsigned long long) * 16));
    return SCB_OK;
}
#endif

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

static int launch_population_fitness(const int64_t* counts, int n_groups, int64_t n_individuals,
                                     const uint8_t* lut, const uint8_t* active, int64_t* out_diff,
                                     int64_t* out_diff_pg, const unsigned char* region, int rank, int n_ranks,
                                     const SumsInput& sums, scb_stream_t stream) {
    size_t smem = (size_t)n_groups * 25 * sizeof(long long);  // base + 8 int64 deltas + 32 int32 nibble sums
    if ((sums.mode & 15) == 1)   // + a copy of the accumulators, of the descriptors and of the group order
        smem += ((size_t)sums.n_rows + (size_t)kTermsPerGroup * n_groups + 3) * 8 + (size_t)n_groups * 18 + 16;
    if (smem > 200 * 1024) {
        set_error("population_fitness: %d groups exceed the shared-memory table (max %d)", n_groups, (int)(200 * 1024 / 200));
        return SCB_ERR_UNSUPPORTED;
    }
    // 32 warps per CTA and, by default, three individuals per warp: the kernel is dominated by its set-up (every
    // CTA builds the tables), so fewer CTAs cost little time and leave SMs to the kernel that follows -- in the GA
    // step the next pass's counting kernel, a programmatic dependent that claims its tiles dynamically.
    // SCB_FITNESS_PER_WARP=k overrides (experiments).
    int per_warp = 3;  // measured on H1M, population 4096: 22.8 / 19.6 / 18.0 / 18.8 / 20.1 us per step for 1 / 2 / 3 / 4 / 5
    if (const char* e = getenv("SCB_FITNESS_PER_WARP")) per_warp = atoi(e) > 0 ? atoi(e) : 1;
    int64_t blocks = (n_individuals + 32 * per_warp - 1) / (32 * per_warp);
    int64_t cap = (int64_t)sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    // staged truth tables (see the kernel): `chunk` consecutive individuals per CTA, rows in shared memory
    const int64_t chunk = (n_individuals + blocks - 1) / blocks;
    int lut_off = 0;
    if (n_groups <= 256 && !active && !out_diff_pg && chunk * n_groups <= 64 * 1024 && !getenv("SCB_FITNESS_NO_STAGE")) {
        lut_off = (int)((smem + 15) & ~(size_t)15);
        smem = (size_t)lut_off + (size_t)((chunk * n_groups + 15) & ~(int64_t)15);
    }
    static thread_local SmemMarks marks;
    if (int rc = ensure_dynamic_smem((const void*)population_fitness_kernel, smem, marks)) return rc;
    // programmatic dependent launch: the kernel may start while its producer (scb_rule_counts in the
    // same stream) drains; it waits for it with griddepcontrol.wait before it reads the counts
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3(1024);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[2];
    int n_attr = 0;
    if (!getenv("SCB_NO_PDL")) {
        attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
        ++n_attr;
    }
    // Sharded cells, experiment (SCB_FITNESS_CLUSTER=1): clusters of up to 8 CTAs share the summing of
    // the ranks' slots.  Measured at N=8 on H1M: 37.5 us per step against 28.7 us without clusters --
    // a cluster only starts when 8 SMs of one GPC are free of rule_counts CTAs, and the two cluster
    // barriers cost more than the 7/8 of the L2 reads they save.  Off by default.
    int cluster = 1;
    if (region && n_ranks > 1 && getenv("SCB_FITNESS_CLUSTER"))
        for (cluster = 8; cluster > 1 && blocks % cluster != 0; cluster >>= 1) {}
    if (cluster > 1) {
        attr[n_attr].id = cudaLaunchAttributeClusterDimension;
        attr[n_attr].val.clusterDim.x = (unsigned)cluster;
        attr[n_attr].val.clusterDim.y = 1;
        attr[n_attr].val.clusterDim.z = 1;
        ++n_attr;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n_attr;
    SCB_CUDA(cudaLaunchKernelEx(&cfg, population_fitness_kernel, counts, n_groups, n_individuals, lut, active,
                                out_diff, out_diff_pg, region, rank, n_ranks, sums, lut_off, (int)chunk));
    return SCB_OK;
}

static SumsInput no_sums() {
    SumsInput s;
    s.mode = 0;
    s.plan = nullptr;
    s.acc = nullptr;
    s.ticket = nullptr;
    s.n_rows = s.zero_row = 0;
    s.n_cells = 0;
    return s;
}

extern "C" int scb_population_fitness(const int64_t* counts, int n_groups, int64_t n_individuals,
                                      const uint8_t* lut, const uint8_t* active, int64_t* out_diff,
                                      int64_t* out_diff_pg, scb_stream_t stream) {
    SCB_REQUIRE(n_groups > 0 && n_individuals >= 0, "population_fitness: bad size");
    if (n_individuals == 0) return SCB_OK;
    SCB_REQUIRE(counts && lut && out_diff, "population_fitness: null pointer");
    return launch_population_fitness(counts, n_groups, n_individuals, lut, active, out_diff, out_diff_pg, nullptr,
                                     0, 1, no_sums(), stream);
}

extern "C" int scb_population_fitness_sums(const void* plan, void* scratch, int n_rows, int n_groups, int64_t n_cells,
                                           const void* local_region, int rank, int n_ranks, int64_t n_individuals,
                                           const uint8_t* lut, const uint8_t* active, int64_t* out_diff,
                                           int64_t* out_diff_pg, scb_stream_t stream) {
    SCB_REQUIRE(n_groups > 0 && n_rows > 0 && n_individuals >= 0 && n_cells >= 0, "population_fitness_sums: bad size");
    SCB_REQUIRE(plan && scratch && lut && out_diff, "population_fitness_sums: null pointer");
    SCB_REQUIRE(!local_region || (n_ranks >= 1 && n_ranks <= SCB_MAX_PEERS && rank >= 0 && rank < n_ranks), "population_fitness_sums: rank %d of %d", rank, n_ranks);
    // (an empty population still has to consume -- and clear -- the sums of its pass)
    SumsInput sums;
    sums.mode = local_region ? 2 : 1;
    if (const char* e = getenv("SCB_SUMS_DBG")) sums.mode |= atoi(e) << 4;
    sums.plan = static_cast<const unsigned char*>(plan);
    sums.acc = static_cast<unsigned long long*>(scratch);
    sums.ticket = reinterpret_cast<unsigned int*>(sums.acc + (size_t)n_rows + (size_t)kTermsPerGroup * n_groups);
    sums.n_rows = n_rows;
    sums.zero_row = plan_zero_row(n_rows);
    sums.n_cells = n_cells;
    return launch_population_fitness(nullptr, n_groups, n_individuals, lut, active, out_diff, out_diff_pg,
                                     static_cast<const unsigned char*>(local_region), rank, local_region ? n_ranks : 1,
                                     sums, stream);
}

extern "C" int scb_population_fitness_peers(const void* local_region, int rank, int n_ranks, int n_groups,
                                            int64_t n_individuals, const uint8_t* lut, const uint8_t* active,
                                            int64_t* out_diff, int64_t* out_diff_pg, scb_stream_t stream) {
    SCB_REQUIRE(n_groups > 0 && n_individuals >= 0, "population_fitness_peers: bad size");
    SCB_REQUIRE(n_ranks >= 1 && n_ranks <= SCB_MAX_PEERS && rank >= 0 && rank < n_ranks, "population_fitness_peers: rank %d of %d", rank, n_ranks);
    if (n_individuals == 0) return SCB_OK;
    SCB_REQUIRE(local_region && lut && out_diff, "population_fitness_peers: null pointer");
    return launch_population_fitness(nullptr, n_groups, n_individuals, lut, active, out_diff, out_diff_pg,
                                     static_cast<const unsigned char*>(local_region), rank, n_ranks, no_sums(), stream);
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
