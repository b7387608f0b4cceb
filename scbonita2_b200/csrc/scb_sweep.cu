// Knock-out / knock-in importance sweep and per-cell trajectories.
//
// Replaces CalculateImportanceScore (importance_scores.py:47-143,195-340) and
// attractor_analysis.vectorized_run_simulation (attractor_analysis.py:43-99).
//
// Representation: bit-sliced.  One 32-bit word holds one node's state for 32 cells, so one
// truth-table evaluation (7 LOP3) updates a node for 32 cells.  A CTA owns SLOTS words
// ("slots"); the network state of a slot is N words in shared memory.  Thread (slot, part)
// updates the part-th slice of the nodes for its slot; lanes of a warp are consecutive slots,
// so every shared-memory access is bank-conflict free and every per-node constant (parent
// rows, truth table) is warp-uniform.
//
// Exact first-repeat semantics with O(1) state copies (no history):
//   phase 1  forward simulation.  lambda (cycle length) per cell: 1 if state_t == state_{t-1},
//            otherwise Brent's scheme - compare with a checkpoint taken at t = 2^k - 1.  All
//            per-cell integers (lambda) are kept bit-sliced too (6 bit-planes per slot).
//   phase 2  (only if some cell in the CTA cycles with lambda > 1) restart twice from the data:
//            P runs lambda_cell steps ahead of Q (per-cell stall masks); the first j with
//            P == Q is mu, the first-repeat pair is (mu, mu+lambda), valid iff mu+lambda < steps
//            (find_attractors, importance_scores.py:134-143); Q freezes at state_mu.
//   scoring  cells frozen at their own state_mu are released together and stepped in lock-step
//            with the stored normal-run states state_{mu_n + t}; XOR popcounts are accumulated
//            for t <= min(lambda_x, lambda_n) (align_attractors, importance_scores.py:325-340).
// The normal run is analysed first (MODE_NORMAL) and leaves found/lambda planes and the
// aligned states in scratch; KO and KI of one node share a CTA (slots [0,SLOTS/2) = KO,
// [SLOTS/2,SLOTS) = KI of the same words) because a cell only counts if BOTH repeat
// (importance_scores.py:278).
#include <stdlib.h>

#include <type_traits>

#include "scb_common.cuh"

namespace scb {

constexpr int kLamPlanes = 6;  // lambda < steps <= 64
#ifndef SCB_SWEEP_THREADS
#define SCB_SWEEP_THREADS 512
#endif
constexpr int kSweepThreads = SCB_SWEEP_THREADS;

struct __align__(16) NodeEntry {
    int pa, pb, pc;
    uint32_t lut;
};

// bit-sliced "value > c" / "value >= c" for a per-lane integer held in kLamPlanes planes
__device__ __forceinline__ uint32_t planes_gt(const uint32_t (&v)[kLamPlanes], int c) {
    uint32_t gt = 0, eq = 0xFFFFFFFFu;
#pragma unroll
    for (int b = kLamPlanes - 1; b >= 0; --b) {
        uint32_t cb = ((c >> b) & 1) ? 0xFFFFFFFFu : 0u;
        gt |= eq & v[b] & ~cb;
        eq &= ~(v[b] ^ cb);
    }
    return gt;
}
__device__ __forceinline__ uint32_t planes_ge(const uint32_t (&v)[kLamPlanes], int c) {
    return c <= 0 ? 0xFFFFFFFFu : planes_gt(v, c - 1);
}
__device__ __forceinline__ void planes_set(uint32_t (&v)[kLamPlanes], uint32_t lanes, int value) {
#pragma unroll
    for (int b = 0; b < kLamPlanes; ++b)
        if ((value >> b) & 1) v[b] |= lanes;
}

// Canonical form of a node: the bound parent slots packed to the left (A, AB, ABC) in their
// original order, the truth table rewritten for the new slots with every unbound slot reading
// constant 0 (the reference's missing parents), unbound slots aliased to row 0.  The new table does
// not depend on the unbound slots, so the general three-input evaluation stays valid, and a node
// with `arity` bound parents can be evaluated from table entries 0/4 (one parent) or 0/2/4/6 (two).
__device__ __forceinline__ NodeEntry make_entry(const int32_t* __restrict__ net_parent,
                                                const uint8_t* __restrict__ net_lut, int n, int* arity = nullptr) {
    NodeEntry e;
    const uint32_t lut = net_lut[n];
    int real[3], k = 0, p[3];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        p[s] = net_parent[3 * n + s];
        if (p[s] >= 0) real[k++] = s;
    }
    uint32_t canon = 0;
#pragma unroll
    for (int idx = 0; idx < 8; ++idx) {
        int orig = 0;  // index into the original table: slot s is bit (2 - s)
        for (int j = 0; j < k; ++j) orig |= ((idx >> (2 - j)) & 1) << (2 - real[j]);
        canon |= ((lut >> orig) & 1u) << idx;
    }
    e.pa = k > 0 ? p[real[0]] : 0;
    e.pb = k > 1 ? p[real[1]] : 0;
    e.pc = k > 2 ? p[real[2]] : 0;
    e.lut = canon;
    if (arity) *arity = k;
    return e;
}

template <int SLOTS>
__device__ __forceinline__ uint32_t eval_node(const NodeEntry& e, const uint32_t* __restrict__ st,
                                              int slot) {
    LutMasks m = expand_lut(e.lut);
    return eval_lut(m, st[e.pa * SLOTS + slot], st[e.pb * SLOTS + slot], st[e.pc * SLOTS + slot]);
}

// Per-node constants as the sweep kernel keeps them in shared memory: element offsets of the
// three parent rows inside the state array and the truth table expanded to 8 full-word masks.
// Everything is warp-uniform, so each load is a broadcast.
struct __align__(16) NodeOffsets {
    int pa, pb, pc, arity;  // arity: bound parents (packed left); the others alias row 0
};


struct SweepArgs {
    const uint32_t* planes;
    int n_nodes;
    int64_t wpad, n_cells;
    const int32_t* net_parent;
    const uint8_t* net_lut;
    int steps;
    uint32_t* np_states;            // [steps][N][wpad]  normal-run states from each cell's own mu
    uint32_t* np_lam;               // [6][wpad]         lambda of the normal run, bit-sliced
    uint32_t* np_found;             // [wpad]
    // cell-major copies (bit n%32 of word n/32 of cell c's record = node n), read by the passes that
    // simulate compacted lists of cells: a cell's whole state is ONE 16/32-byte record
    uint32_t* planes_t;             // [wpad*32][nwp]         the data
    uint32_t* np_states_t;          // [steps][wpad*32][nwp]  normal-run states from each cell's own mu
    int nwp;                        // words per record: ceil(N/32) rounded up to a multiple of 4
    uint32_t* np_info_t;            // [wpad*32]  bits 0..5 lambda, bit 6 found (np_lam | np_found, cell-major)
    // per-node lists of deferred cells (KO and KI of a cell travel together): a pass reads the
    // `in` lists (gather modes) and appends the cells it leaves open to the `out` lists
    const int32_t* in_cells;        // [N][in_cap]
    int64_t in_cap;
    const unsigned int* in_count;   // [N] entries used
    int32_t* out_cells;             // [N][out_cap]
    int64_t out_cap;
    unsigned int* out_count;        // [N]
    unsigned int* counters;         // [1] flags (bit 0: normal run has lambda >= 3 or an unsettled cell)
    int fast_steps;                 // KO/KI: steps of the ring fast path before open cells are deferred;
                                    // <= 0: chosen on the device from the normal run (counters[3])
    unsigned int* settle_hist;      // [64] normal run: cells whose first repeat (period <= 2) the ring saw in step t
    uint32_t* np_hard;              // [wpad] normal run: cells the deep ring could not settle (long cycle / no repeat)
    int32_t* hard_cells;            // KO/KI pass: such cells go straight to this list (the cleanup pass's input)
    int64_t hard_cap;
    unsigned int* hard_count;
    int recheck_follows;            // the KO/KI pass is followed by the compacted recheck pass (deep ring)
    // cells sorted by the settle time of their unperturbed run (launch_sweep): a probe run of MODE_NORMAL on
    // the caller's planes writes a 6-bit key per cell (bit-sliced) and the histogram of the keys
    int probe;                      // MODE_NORMAL: only write key_planes / key_hist
    uint32_t* key_planes;           // [6][wpad]
    unsigned int* key_hist;         // [64]
    // KO/KI register ring: a CTA stops stepping when at most `stop_open` of its lanes are still open, or when
    // lanes have started to settle and none did for `stop_stall` steps; open lanes go to the next pass
    int stop_open, stop_stall;
    int ring_stall;                 // the same stall rule in the shared-memory ring of the KO/KI and recheck passes
    // Single runs (koki_plane_kernel, koki_compact_kernel<.., SINGLES>): a cell with ONE open run is deferred as that
    // run alone -- entry = cell | which << 30 (which: 0 = KO, 1 = KI) -- in the same per-node arrays as the pairs,
    // filled from the BACK (entry j of node n at cells[n * cap + cap - 1 - j]): a cell is in at most one of the two.
    // The run that did settle is scored at once as if its partner will repeat; if the partner turns out not to, the
    // settled run goes to the retract list and a last pass subtracts its contribution again.
    int singles;                    // plane pass: defer single runs (else every open cell travels as a pair)
    const unsigned int* in_count1;  // [N] single runs in the `in` arrays
    unsigned int* out_count1;       // [N] single runs appended to the `out` arrays
    int32_t* rt_cells;              // [N][rt_cap] retract list (filled from the back like the singles)
    int64_t rt_cap;
    unsigned int* rt_count;         // [N]
    int negate;                     // compacted singles pass: subtract the scores (the retract pass), nothing else
    int brent_mid;                  // cleanup pass: keep the permanent mid-window checkpoint (SCB_SWEEP_BRENT_MID=0: off)
    unsigned long long* out_raw;
    unsigned long long* out_missing;
    unsigned long long* out_keyerror;
};
constexpr int kWhichShift = 30;
constexpr int32_t kCellMask = (1 << kWhichShift) - 1;

// 32 x 32 bit transpose across a warp: lane i passes row i (bit j = element (i, j)) and receives
// column i (bit j = element (j, i)).  Five exchange steps: at distance d the lanes i and i^d swap the
// off-diagonal d x d blocks.
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, int lane) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        // bits whose index has bit d clear: 0x0000FFFF, 0x00FF00FF, 0x0F0F0F0F, 0x33333333, 0x55555555
        const uint32_t lo_mask = d == 16 ? 0x0000FFFFu : d == 8 ? 0x00FF00FFu : d == 4 ? 0x0F0F0F0Fu : d == 2 ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xFFFFFFFFu, x, d);
        if (lane & d) x = (x & ~lo_mask) | ((y >> d) & lo_mask);   // upper row: keep its high block, take the partner's high block as low
        else x = (x & lo_mask) | ((y << d) & ~lo_mask);            // lower row: keep its low block, take the partner's low block as high
    }
    return x;
}

// Bit-matrix transpose: node-major bitplanes src[slice][n_rows][wpad] (bit c%32 of word c/32 of row
// n) -> cell-major records dst[slice][wpad*32][nwp] (bit n%32 of word n/32 of cell c).  One CTA
// moves a tile of 32 rows x 32 words: coalesced row reads into shared memory, then one warp per
// word column -- lane b holds row b's word, bit l of every lane is collected with a ballot and
// kept by lane l, which ends up with the 32 row bits of cell l.  Slices past *last_slice (when
// given) are skipped: the normal run only fills the slices its longest cycle needs.
__global__ void __launch_bounds__(256)
cell_major_kernel(const uint32_t* __restrict__ src, int n_rows, int64_t wpad, uint32_t* __restrict__ dst,
                  int nwp, const unsigned int* __restrict__ last_slice, int n_slices) {
    __shared__ uint32_t tile[32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t w0 = (int64_t)blockIdx.x * 32;
    const int j = blockIdx.y;
    // (a CTA walks the slices itself: a grid with one z-layer per slice launches 50 x as many CTAs, nearly all of
    // which find their slice past *last_slice and leave -- 0.4 ms of launch overhead on the 1M-cell sweep)
    const int z_end = last_slice ? min(n_slices, (int)*last_slice + 1) : n_slices;
    for (int z = 0; z < z_end; ++z, src += (size_t)n_rows * wpad, dst += (size_t)wpad * 32 * nwp) {
    if (z) __syncthreads();
    for (int r = warp; r < 32; r += 8) {
        const int row = 32 * j + r;
        tile[r][lane] = (row < n_rows && w0 + lane < wpad) ? __ldg(src + (size_t)row * wpad + w0 + lane) : 0u;
    }
    __syncthreads();
    for (int wc = warp; wc < 32; wc += 8) {
        if (w0 + wc >= wpad) break;
        const uint32_t v = tile[lane][wc];
        uint32_t mine = 0;
#pragma unroll
        for (int l = 0; l < 32; ++l) {
            const uint32_t bits = __ballot_sync(0xFFFFFFFFu, (v >> l) & 1u);
            if (lane == l) mine = bits;
        }
        dst[((size_t)(w0 + wc) * 32 + lane) * nwp + j] = mine;
    }
    }
}

// Inverse of cell_major_kernel for one slice: cell-major records src[wpad*32][nwp] -> node-major
// bitplanes dst[n_rows][wpad].  One CTA makes 32 rows x 32 words: lane l of a warp reads word j of
// cell 32w + l's record, the warp transpose hands lane b the 32 cell bits of node 32j + b, and the
// tile goes out as coalesced row segments.  Cells at and past n_cells read as zero.
__global__ void __launch_bounds__(256)
node_major_kernel(const uint32_t* __restrict__ src, int nwp, int n_rows, int64_t wpad, int64_t n_cells,
                  uint32_t* __restrict__ dst) {
    __shared__ uint32_t tile[32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t w0 = (int64_t)blockIdx.x * 32;
    const int j = blockIdx.y;
    for (int wc = warp; wc < 32; wc += 8) {
        const int64_t cell = (w0 + wc) * 32 + lane;
        const uint32_t x = (w0 + wc < wpad && cell < n_cells) ? __ldg(src + (size_t)cell * nwp + j) : 0u;
        tile[lane][wc] = warp_transpose32(x, lane);
    }
    __syncthreads();
    for (int r = warp; r < 32; r += 8) {
        const int row = 32 * j + r;
        if (row < n_rows && w0 + lane < wpad) dst[(size_t)row * wpad + w0 + lane] = tile[r][lane];
    }
}

// Counting sort of the cells by their 6-bit key (settle step of the unperturbed run, MODE_NORMAL probe):
// cell c's record moves to position base[key] + (arrival order within the key).  The order inside a
// key is arbitrary -- every output of the sweep is a sum over cells.
__global__ void __launch_bounds__(256)
sort_scatter_kernel(const uint32_t* __restrict__ key_planes, int64_t wpad, int64_t n_cells,
                    const unsigned int* __restrict__ key_hist, unsigned int* __restrict__ bucket_fill,
                    const uint32_t* __restrict__ rec_in, uint32_t* __restrict__ rec_out, int nwp) {
    __shared__ unsigned int base[64];
    if (threadIdx.x == 0) {
        unsigned int acc = 0;
        for (int k = 0; k < 64; ++k) {
            base[k] = acc;
            acc += key_hist[k];
        }
    }
    __syncthreads();
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int64_t w = c >> 5;
    const bool active = c < n_cells;
    unsigned int key = 0;
    if (active) {
#pragma unroll
        for (int b = 0; b < 6; ++b) key |= ((__ldg(key_planes + (size_t)b * wpad + w) >> lane) & 1u) << b;
    }
    const unsigned int peers = __match_any_sync(0xFFFFFFFFu, active ? key : 0x100u + lane);
    const int leader = __ffs(peers) - 1;
    const unsigned int rank = __popc(peers & ((1u << lane) - 1u));
    unsigned int start = 0;
    if (active && lane == leader) start = atomicAdd(&bucket_fill[key], (unsigned int)__popc(peers));
    start = __shfl_sync(0xFFFFFFFFu, start, leader);
    if (active) {
        const size_t pos = (size_t)base[key] + start + rank;
        const uint4* in = reinterpret_cast<const uint4*>(rec_in + (size_t)c * nwp);
        uint4* out = reinterpret_cast<uint4*>(rec_out + pos * nwp);
        for (int q = 0; q < nwp / 4; ++q) out[q] = __ldg(in + q);
    }
}

#ifdef SCB_TRACE
// phase clock sums per MODE (tools/trace_sweep.py): [mode][phase] cycles of thread 0, [mode][15] CTAs/batches
__device__ unsigned long long g_sweep_phase[4][16];
#define SCB_SWEEP_PHASE(i)                                                         \
    do {                                                                           \
        if (tid == 0) {                                                            \
            const long long now_ = clock64();                                      \
            atomicAdd(&g_sweep_phase[MODE][i], (unsigned long long)(now_ - trace_t_)); \
            trace_t_ = now_;                                                       \
        }                                                                          \
    } while (0)
#else
#define SCB_SWEEP_PHASE(i) do { } while (0)
#endif

// How many ring steps the KO/KI pass runs on the plane words before it hands the cells that are
// still open to the compacted recheck pass.  A step on the plane words costs the same whether one
// lane of a word is open or all 32, a rechecked cell is simulated again from the start: with u(T) the
// share of cells still open after T steps the work is ~ T + kRecheckSteps * (u(T) - u(steps)) plane-word steps.
// u is taken from the NORMAL run (settle-time histogram), the KO/KI runs of a cell settle at about
// the same time as its unperturbed run.  Any choice gives the same results.
constexpr int kRecheckSteps = 115;  // cost of rechecking every cell, in plane-word steps (measured, H1M)
__global__ void choose_fast_steps_kernel(const unsigned int* __restrict__ hist, int steps, int64_t n_cells,
                                         unsigned int* __restrict__ counters, int sorted) {
    if (threadIdx.x != 0) return;
    if (sorted) {  // cells sorted by settle step: every CTA stops at its own cells' pace (ring2_registers)
        counters[3] = (unsigned int)steps;
        return;
    }
    // cells that never settle within the window are rechecked whatever T is: only the cells that
    // WOULD have settled between T and `steps` are the price of stopping early
    long long never = n_cells;
    for (int t = 0; t < steps; ++t) never -= hist[t];
    long long open = n_cells;
    long long best_cost = (long long)steps * n_cells;  // T = steps
    int best = steps;
    for (int t = 0; t < steps; ++t) {
        open -= hist[t];  // open after T = t + 1 steps
        const int T = t + 1;
        if (T < 4) continue;
        const long long cost = (long long)T * n_cells + (long long)kRecheckSteps * (open - never);
        if (cost < best_cost) {
            best_cost = cost;
            best = T;
        }
    }
    counters[3] = (unsigned int)best;
}

// MODE_NORMAL  the unperturbed run (ring fast path, slow path in place)
// MODE_KOKI    KO/KI of every node on the plane words: ring fast path for the first `fast_steps`
//              steps, cells still open are appended to the out lists
// MODE_RECHECK deferred cells, compacted 32 to a virtual word: the full-length ring fast path; what is
//              still open (long cycles, no repeat) goes to the out lists
// MODE_CLEANUP deferred cells, compacted: the general slow path (Brent + replays)
enum { MODE_NORMAL = 0, MODE_KOKI = 1, MODE_CLEANUP = 2, MODE_RECHECK = 3 };

#ifndef SCB_REG_RING
#define SCB_REG_RING 1
#endif
#ifndef SCB_REG_BRENT
#define SCB_REG_BRENT 1
#endif
constexpr bool kRegBrent = SCB_REG_BRENT != 0;  // experiments: 0 keeps the Brent phase's compare operands in shared memory
constexpr bool kRegRing = SCB_REG_RING != 0;  // experiments: 0 keeps the ring of the fast path in shared memory
#ifndef SCB_RING_ARITY
#define SCB_RING_ARITY 1
#endif
// experiments: 0 evaluates every node of the register layouts as a three-parent one (branch-free: the
// loads of all of a thread's nodes can fly together, at the price of unused rows and table words)
constexpr bool kRingArity = SCB_RING_ARITY != 0;
constexpr int kMaskPad = 15;  // all-zero truth tables after the last node
constexpr int kRing = 4;  // state copies per slot: ring of the fast path, P/Q pairs of the slow path

// Truth table of a canonical node as the sweep kernels evaluate it.  The table words of the mux tree
// are all-ones or all-zeros and warp-uniform, so the FIRST mux level -- "x ? m1 : m0" on such words --
// is one of 0, ~0, x, ~x, i.e. the affine map x * S + O with S = (m1 & 1) - (m0 & 1) and O = m0: an
// IMAD on the FMA pipe instead of a LOP3 on the ALU pipe, which the ring compares and freeze merges
// already keep busy.  Only the deeper levels mix two non-constant words and need LOP3.
//   arity <= 1: t[0], t[1]       = S, O of (m4, m0)                        f = a * S + O
//   arity == 2: t[0..3]          = S, O of (m2, m0) and (m6, m4)           f = a ? (b*S1+O1) : (b*S0+O0)
//   arity == 3: t[2k], t[2k + 1] = S, O of (m[2k+1], m[2k]), k = 0..3      g_k = c*S_k+O_k, then two mux levels
__device__ __forceinline__ uint32_t affine(uint32_t x, uint32_t s, uint32_t o) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(s), "r"(o));
    return r;
}
__device__ __forceinline__ LutMasks affine_table(uint32_t lut, int arity) {
    const LutMasks m = expand_lut(lut);
    LutMasks t;
#pragma unroll
    for (int k = 0; k < 8; ++k) t.m[k] = 0u;
    auto put = [&](int at, uint32_t m1, uint32_t m0) {
        t.m[at] = (m1 & 1u) - (m0 & 1u);
        t.m[at + 1] = m0;
    };
    if (arity <= 1) {
        put(0, m.m[4], m.m[0]);
    } else if (arity == 2) {
        put(0, m.m[2], m.m[0]);
        put(2, m.m[6], m.m[4]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) put(2 * k, m.m[2 * k + 1], m.m[2 * k]);
    }
    return t;
}

// Value of one canonical node (NodeOffsets / affine_table entry) for the two slots a thread of the
// register layouts owns: `srd` points at (node 0, copy, slot lane), the second slot is 32 words on.
// The node index is warp-uniform, so is the branch on its arity: one- and two-parent nodes read 2 / 4
// rows and 2 / 4 table words.
__device__ __forceinline__ void eval_pair(const NodeOffsets& o, const LutMasks& mk, const uint32_t* __restrict__ srd,
                                          uint32_t& f0, uint32_t& f1) {
    if (kRingArity && o.arity <= 1) {
        const uint32_t s = mk.m[0], c = mk.m[1];
        f0 = affine(srd[o.pa], s, c);
        f1 = affine(srd[o.pa + 32], s, c);
        return;
    }
    const uint4 lo = *reinterpret_cast<const uint4*>(&mk.m[0]);
    const uint32_t a0 = srd[o.pa], b0 = srd[o.pb];
    const uint32_t a1 = srd[o.pa + 32], b1 = srd[o.pb + 32];
    if (kRingArity && o.arity == 2) {
        f0 = lop3<0xCA>(a0, affine(b0, lo.z, lo.w), affine(b0, lo.x, lo.y));
        f1 = lop3<0xCA>(a1, affine(b1, lo.z, lo.w), affine(b1, lo.x, lo.y));
        return;
    }
    const uint4 hi = *reinterpret_cast<const uint4*>(&mk.m[4]);
    const uint32_t c0 = srd[o.pc], c1 = srd[o.pc + 32];
    f0 = lop3<0xCA>(a0, lop3<0xCA>(b0, affine(c0, hi.z, hi.w), affine(c0, hi.x, hi.y)),
                    lop3<0xCA>(b0, affine(c0, lo.z, lo.w), affine(c0, lo.x, lo.y)));
    f1 = lop3<0xCA>(a1, lop3<0xCA>(b1, affine(c1, hi.z, hi.w), affine(c1, hi.x, hi.y)),
                    lop3<0xCA>(b1, affine(c1, lo.z, lo.w), affine(c1, lo.x, lo.y)));
}

// Ring-of-2 fast path with the ring in REGISTERS (KO/KI and recheck passes, 64-slot CTAs of 16 warps,
// N <= 16 * CH, N <= 256).  Thread (lane, warp p) owns nodes [p*CH, p*CH + CH) of the KO slot `lane`
// AND the KI slot `lane + 32` of the same word: the truth-table masks are fetched once for both, the
// parent rows of the thread's nodes sit in registers (three bytes of one word per node), the two
// previous states of its own nodes never leave its registers, and shared memory only holds the
// current and the next state (copies 3 and 0, swapped every step) for the other warps to read --
// 10 shared-memory wavefronts per (node, 64 slots, step) instead of 18, and one CTA barrier per step
// (three sets of vote words: set t%3 is written in step t, read after its barrier, cleared in step
// t+1 -- by then every reader is past the barrier of step t+1).  The barrier of step t also carries
// "every lane was frozen after step t-1", so a CTA leaves one (idempotent) step late.
// Same semantics as the shared-memory ring: a lane whose new state equals state_{t-1} (t >= 1) or
// state_{t-2} (t >= 2) has its first repeat there and is frozen at it.
// Results per slot go to votes[0..63] (frozen), [64..127] (lambda bit 0), [128..191] (bit 1);
// returns the state copy that holds the final states.
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds_u128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u32_if(uint32_t addr, uint32_t v, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u32 [%0], %1;\n\t}" ::"r"(addr), "r"(v), "r"((uint32_t)pred) : "memory");
}
__device__ __forceinline__ uint32_t shared_addr(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

template <int CH>
__device__ __forceinline__ int ring2_registers(uint32_t* __restrict__ state, const NodeOffsets* __restrict__ offs,
                                               const LutMasks* __restrict__ masks, uint32_t* __restrict__ votes,
                                               int N, int forced_node, uint32_t cellmask, int t_fast, int tid,
                                               int stop_open, int stop_stall) {
    constexpr int SL = 64, STRIDE = 4 * SL, SET = 2 * SL;
    const int lane = tid & 31;
    const int n0 = (tid >> 5) * CH;
    const int n_mine = N - n0;              // node i of this thread exists iff i < n_mine
    const int forced_i = forced_node - n0;  // ... and is the forced one iff i == forced_i
    uint32_t xa0[CH], xb0[CH], xa1[CH], xb1[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        xa0[i] = xa1[i] = xb0[i] = xb1[i] = 0u;
        if (i < n_mine) {
            xa0[i] = state[((n0 + i) * 4 + 3) * SL + lane];
            xa1[i] = state[((n0 + i) * 4 + 3) * SL + lane + 32];
        }
    }
    for (int i = tid; i < 3 * SET; i += blockDim.x) votes[i] = 0;
    __syncthreads();
    uint32_t frozen0 = ~cellmask, frozen1 = ~cellmask;
    uint32_t lam0_0 = 0, lam1_0 = 0, lam0_1 = 0, lam1_1 = 0;
    int vset = 0;  // t % 3
    // every warp holds the frozen masks of all 64 slots, so "how many lanes are still open" is one warp
    // reduction and the same number everywhere: the CTA leaves the loop without another barrier
    unsigned int last_open = 0xFFFFFFFFu;
    int stall = 0;
    bool moved = false;

    // one step: reads copy RD, writes copy WR; cur* hold state_{t-1} of the own nodes, prv* state_{t-2}
    // and receive state_t.  Returns true when the CTA is done (copy WR holds the result).
    auto step = [&](auto rd_tag, uint32_t (&cur0)[CH], uint32_t (&prv0)[CH], uint32_t (&cur1)[CH],
                    uint32_t (&prv1)[CH], int t) -> bool {
        constexpr int RD = decltype(rd_tag)::value ? 3 : 0, WR = decltype(rd_tag)::value ? 0 : 3;
        uint32_t q10 = 0, q20 = 0, q11 = 0, q21 = 0;
        // copy RD is only read and copy WR only written in a step: telling the compiler lets it start
        // the loads of the next node before the stores of the current one
        const uint32_t* __restrict__ srd = state + RD * SL + lane;
        uint32_t* __restrict__ swr = state + (n0 * 4 + WR) * SL + lane;
        if (n_mine > 0) {  // warps past the last node only join the barrier
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                // a node past the end (only in the last warp with nodes: the offset and mask tables are
                // padded with zero entries) has an all-zero truth table and all-zero registers: it votes
                // "equal" and only its stores are suppressed
                uint32_t f0, f1;
                eval_pair(offs[n0 + i], masks[n0 + i], srd, f0, f1);
                if (i == forced_i) {  // never true past the end
                    f0 = 0u;
                    f1 = 0xFFFFFFFFu;
                }
                q10 |= f0 ^ cur0[i];
                q20 |= f0 ^ prv0[i];
                q11 |= f1 ^ cur1[i];
                q21 |= f1 ^ prv1[i];
                const uint32_t w0 = (f0 & ~frozen0) | (cur0[i] & frozen0);
                const uint32_t w1 = (f1 & ~frozen1) | (cur1[i] & frozen1);
                if (i < n_mine) {
                    swr[i * STRIDE] = w0;
                    swr[i * STRIDE + 32] = w1;
                }
                prv0[i] = w0;
                prv1[i] = w1;
            }
        }
        uint32_t* v = votes + vset * SET;
        if (q10) atomicOr(v + lane, q10);
        if (q11) atomicOr(v + 32 + lane, q11);
        if (q20) atomicOr(v + SL + lane, q20);
        if (q21) atomicOr(v + SL + 32 + lane, q21);
        __syncthreads();
        // state_{t-d} exists for t >= d only (the raw data state is never compared)
        uint32_t open = ~frozen0;
        uint32_t e1 = t >= 1 ? ~v[lane] & open : 0u;
        open &= ~e1;
        uint32_t e2 = t >= 2 ? ~v[SL + lane] & open : 0u;
        lam0_0 |= e1;
        lam1_0 |= e2;
        frozen0 |= e1 | e2;
        open = ~frozen1;
        e1 = t >= 1 ? ~v[32 + lane] & open : 0u;
        open &= ~e1;
        e2 = t >= 2 ? ~v[SL + 32 + lane] & open : 0u;
        lam0_1 |= e1;
        lam1_1 |= e2;
        frozen1 |= e1 | e2;
        // A step costs the same whether one lane of the CTA is open or all 2048: stop when few are left
        // (they are re-simulated, compacted, by the next pass), or when lanes had started to settle and
        // none did for a while (what is left then are cycles this ring cannot see).
        const unsigned int n_open = __reduce_add_sync(0xFFFFFFFFu, (unsigned int)(__popc(~frozen0) + __popc(~frozen1)));
        if (n_open <= (unsigned int)stop_open) return true;
        if (n_open < last_open) {
            moved = last_open != 0xFFFFFFFFu;
            last_open = n_open;
            stall = 0;
        } else if (moved && ++stall >= stop_stall) {
            return true;
        }
        const int vclear = vset == 0 ? 2 : vset - 1;  // (t + 2) % 3
        if (tid < SET) votes[vclear * SET + tid] = 0;
        vset = vset == 2 ? 0 : vset + 1;
        return false;
    };
    using Odd = std::integral_constant<bool, true>;    // reads copy 3
    using Even = std::integral_constant<bool, false>;  // reads copy 0
    int result_copy = 3;
    for (int t = 0; t < t_fast; t += 2) {
        result_copy = 0;
        if (step(Odd{}, xa0, xb0, xa1, xb1, t)) break;  // xa = state_{t-1}, xb <- state_t, written to copy 0
        if (t + 1 >= t_fast) break;
        result_copy = 3;
        if (step(Even{}, xb0, xa0, xb1, xa1, t + 1)) break;  // xb = state_t, xa <- state_{t+1}, written to copy 3
    }
    __syncthreads();  // every vote word has been read
    if (tid < 32) {
        votes[lane] = frozen0;
        votes[32 + lane] = frozen1;
        votes[64 + lane] = lam0_0;
        votes[96 + lane] = lam0_1;
        votes[128 + lane] = lam1_0;
        votes[160 + lane] = lam1_1;
    }
    __syncthreads();
    return result_copy;
}

// Brent phase of the general scheme with the compare operands in REGISTERS (same thread layout as
// ring2_registers: thread (lane, warp p) owns nodes [p*CH, p*CH + CH) of slots `lane` and
// `lane + 32`).  The previous state and the checkpoint of the thread's own nodes stay in its
// registers, shared memory only carries the current and the next state (copies 0 and 1), and there
// is one CTA barrier per step.  Per slot it delivers lambda (6 planes), `cyc` (lanes with a
// period > 1 found through a checkpoint) and `found_fp` (lanes that reached a fixed point inside
// the window) in res[plane * 64 + slot], planes 0..5 = lambda, 6 = cyc, 7 = found_fp; returns the copy
// that holds the last state.  No lane is frozen here: the replay that follows restarts from the data.
template <int CH>
__device__ __forceinline__ int brent_registers(uint32_t* __restrict__ state, const NodeOffsets* __restrict__ offs,
                                               const LutMasks* __restrict__ masks, uint32_t* __restrict__ votes,
                                               int N, int forced_node, uint32_t cellmask, int steps, int tid) {
    constexpr int SL = 64, STRIDE = 4 * SL, SET = 2 * SL;
    const int lane = tid & 31;
    const int n0 = (tid >> 5) * CH;
    const int n_mine = N - n0;
    const int forced_i = forced_node - n0;
    uint32_t cur0[CH], cur1[CH], chk0[CH], chk1[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        cur0[i] = cur1[i] = chk0[i] = chk1[i] = 0u;
        if (i < n_mine) {
            cur0[i] = state[((n0 + i) * 4 + 0) * SL + lane];
            cur1[i] = state[((n0 + i) * 4 + 0) * SL + lane + 32];
        }
    }
    for (int i = tid; i < 3 * SET; i += blockDim.x) votes[i] = 0;
    __syncthreads();
    uint32_t resolved0 = ~cellmask, resolved1 = ~cellmask;
    uint32_t lam0[kLamPlanes], lam1[kLamPlanes];
#pragma unroll
    for (int b = 0; b < kLamPlanes; ++b) lam0[b] = lam1[b] = 0u;
    uint32_t cyc0 = 0, cyc1 = 0, fp0 = 0, fp1 = 0;
    int chk_time = -1;
    bool all_prev = false;
    int vset = 0;
    const int t1_max = 2 * steps - 1;  // see the checkpoint schedule below

    auto step = [&](auto rd_tag, int t) -> bool {
        constexpr int RD = decltype(rd_tag)::value ? 0 : 1, WR = decltype(rd_tag)::value ? 1 : 0;
        uint32_t qp0 = 0, qp1 = 0, qc0 = 0, qc1 = 0;
        const uint32_t* __restrict__ srd = state + RD * SL + lane;
        uint32_t* __restrict__ swr = state + (n0 * 4 + WR) * SL + lane;
        if (n_mine > 0) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                uint32_t f0, f1;
                eval_pair(offs[n0 + i], masks[n0 + i], srd, f0, f1);  // tables padded with all-zero entries past the last node
                if (i == forced_i) {
                    f0 = 0u;
                    f1 = 0xFFFFFFFFu;
                }
                qp0 |= f0 ^ cur0[i];
                qp1 |= f1 ^ cur1[i];
                qc0 |= f0 ^ chk0[i];
                qc1 |= f1 ^ chk1[i];
                if (i < n_mine) {
                    swr[i * STRIDE] = f0;
                    swr[i * STRIDE + 32] = f1;
                }
                cur0[i] = f0;
                cur1[i] = f1;
            }
        }
        uint32_t* v = votes + vset * SET;
        if (qp0) atomicOr(v + lane, qp0);
        if (qp1) atomicOr(v + 32 + lane, qp1);
        if (chk_time >= 0) {
            if (qc0) atomicOr(v + SL + lane, qc0);
            if (qc1) atomicOr(v + SL + 32 + lane, qc1);
        }
        if (__syncthreads_and(all_prev)) return true;  // every lane was resolved after the step before: copy RD is the last state
        const int l = t - chk_time;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t& resolved = h ? resolved1 : resolved0;
            uint32_t(&lam)[kLamPlanes] = h ? lam1 : lam0;
            uint32_t& cyc = h ? cyc1 : cyc0;
            uint32_t& fp = h ? fp1 : fp0;
            const uint32_t neq_prev = v[32 * h + lane], neq_chk = v[SL + 32 * h + lane];
            if (t >= 1) {
                const uint32_t newly = ~neq_prev & ~resolved;
                lam[0] |= newly;  // lambda = 1
                resolved |= newly;
                if (t <= steps - 1) fp |= newly;
            }
            if (chk_time >= 0) {
                const uint32_t newly = ~neq_chk & ~resolved;
                if (l <= steps - 1) {
                    planes_set(lam, newly, l);
                    cyc |= newly;
                }
                resolved |= newly;  // l >= steps: such a cell can never repeat inside the window
            }
        }
        // checkpoints at t = 2^k - 1 < steps - 1 and a last one at t = steps - 1 (see the shared-memory version)
        if ((((t + 1) & t) == 0 && t < steps - 1) || t == steps - 1) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                chk0[i] = cur0[i];
                chk1[i] = cur1[i];
            }
            chk_time = t;
        }
        const int vclear = vset == 0 ? 2 : vset - 1;
        if (tid < SET) votes[vclear * SET + tid] = 0;
        vset = vset == 2 ? 0 : vset + 1;
        all_prev = (resolved0 & resolved1) == 0xFFFFFFFFu;
        return false;
    };
    using FromData = std::integral_constant<bool, true>;   // reads copy 0
    using FromOther = std::integral_constant<bool, false>;  // reads copy 1
    int last_copy = 0;
    for (int t = 0; t < t1_max; t += 2) {
        last_copy = 0;
        if (step(FromData{}, t)) break;
        last_copy = 1;
        if (t + 1 >= t1_max) break;
        if (step(FromOther{}, t + 1)) break;
        last_copy = 0;
    }
    __syncthreads();  // every vote word has been read
    if (tid < 32) {
#pragma unroll
        for (int b = 0; b < kLamPlanes; ++b) {
            votes[b * SL + lane] = lam0[b];
            votes[b * SL + 32 + lane] = lam1[b];
        }
        votes[6 * SL + lane] = cyc0;
        votes[6 * SL + 32 + lane] = cyc1;
        votes[7 * SL + lane] = fp0;
        votes[7 * SL + 32 + lane] = fp1;
    }
    __syncthreads();
    return last_copy;
}

// Replay phase of the general scheme in the same register layout (thread (lane, warp p): nodes
// [p*CH, p*CH + CH) of slots `lane` and `lane + 32`).  On entry copies 0 and 2 hold the data, the
// lambda planes of every slot sit in res[b * 64 + slot] (brent_registers).  P (copies 0/1) first
// runs lambda_cell steps ahead under per-cell stall masks, then P and Q (copies 2/3) step together
// -- one fetch of a node's table serves both and both slots -- until they meet: the first j with
// P == Q is mu; a repeat counts iff mu + lambda <= steps - 1; Q freezes at state_mu.  The own-node
// words that the stall / freeze merges need stay in registers.  Leaves `found` per slot in
// res[slot]; returns the copy that holds Q (its partner is that copy ^ 1).
template <int CH>
__device__ __forceinline__ int replay_registers(uint32_t* __restrict__ state, const NodeOffsets* __restrict__ offs,
                                                const LutMasks* __restrict__ masks, uint32_t* __restrict__ votes,
                                                int N, int forced_node, uint32_t cellmask, int steps, int tid) {
    constexpr int SL = 64, STRIDE = 4 * SL;
    const int lane = tid & 31;
    const int n0 = (tid >> 5) * CH;
    const int n_mine = N - n0;
    const int forced_i = forced_node - n0;
    uint32_t lam0[kLamPlanes], lam1[kLamPlanes];
#pragma unroll
    for (int b = 0; b < kLamPlanes; ++b) {
        lam0[b] = votes[b * SL + lane];
        lam1[b] = votes[b * SL + 32 + lane];
    }
    __syncthreads();
    for (int i = tid; i < 2 * SL; i += blockDim.x) votes[i] = 0;  // two sets of one vote word per slot
    uint32_t pq0[CH], pq1[CH];  // own nodes: P during the advance, Q afterwards
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        pq0[i] = pq1[i] = 0u;
        if (i < n_mine) {
            pq0[i] = state[((n0 + i) * 4 + 0) * SL + lane];
            pq1[i] = state[((n0 + i) * 4 + 0) * SL + lane + 32];
        }
    }
    __syncthreads();
    const uint32_t has0 = planes_ge(lam0, 1) & cellmask, has1 = planes_ge(lam1, 1) & cellmask;

    // value of node n0 + i for both slots, reading copy `rd`
    auto eval2 = [&](int i, const uint32_t* __restrict__ srd, uint32_t& f0, uint32_t& f1) {
        eval_pair(offs[n0 + i], masks[n0 + i], srd, f0, f1);
        if (i == forced_i) {
            f0 = 0u;
            f1 = 0xFFFFFFFFu;
        }
    };

    int pc_ = 0, pn_ = 1;
    for (int s = 0; s < steps; ++s) {
        const uint32_t adv0 = planes_gt(lam0, s) & has0, adv1 = planes_gt(lam1, s) & has1;  // lanes that still owe P a step
        if (!__syncthreads_or((adv0 | adv1) != 0)) break;
        if (n_mine > 0) {
            const uint32_t* __restrict__ srd = state + pc_ * SL + lane;
            uint32_t* __restrict__ swr = state + (n0 * 4 + pn_) * SL + lane;
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                uint32_t f0, f1;
                eval2(i, srd, f0, f1);
                pq0[i] = (f0 & adv0) | (pq0[i] & ~adv0);
                pq1[i] = (f1 & adv1) | (pq1[i] & ~adv1);
                if (i < n_mine) {
                    swr[i * STRIDE] = pq0[i];
                    swr[i * STRIDE + 32] = pq1[i];
                }
            }
        }
        __syncthreads();
        const int tmp = pc_;
        pc_ = pn_;
        pn_ = tmp;
    }
    // from here on the registers hold Q's own nodes (copy 2 = the data)
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        pq0[i] = pq1[i] = 0u;
        if (i < n_mine) {
            pq0[i] = state[((n0 + i) * 4 + 2) * SL + lane];
            pq1[i] = state[((n0 + i) * 4 + 2) * SL + lane + 32];
        }
    }
    int qc_ = 2, qn_ = 3;
    uint32_t done0 = ~has0, done1 = ~has1, found0 = 0, found1 = 0;
    for (int j = 0; j <= steps - 2; ++j) {
        uint32_t neq0 = 0, neq1 = 0;
        if (n_mine > 0) {
            const uint32_t* __restrict__ prd = state + pc_ * SL + lane;
            const uint32_t* __restrict__ qrd = state + qc_ * SL + lane;
            uint32_t* __restrict__ pwr = state + (n0 * 4 + pn_) * SL + lane;
            uint32_t* __restrict__ qwr = state + (n0 * 4 + qn_) * SL + lane;
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                uint32_t fp0, fp1, fq0, fq1;
                eval2(i, prd, fp0, fp1);
                eval2(i, qrd, fq0, fq1);
                neq0 |= fp0 ^ fq0;
                neq1 |= fp1 ^ fq1;
                pq0[i] = (pq0[i] & done0) | (fq0 & ~done0);
                pq1[i] = (pq1[i] & done1) | (fq1 & ~done1);
                if (i < n_mine) {
                    pwr[i * STRIDE] = fp0;
                    pwr[i * STRIDE + 32] = fp1;
                    qwr[i * STRIDE] = pq0[i];
                    qwr[i * STRIDE + 32] = pq1[i];
                }
            }
        }
        uint32_t* v = votes + (j & 1) * SL;
        if (neq0) atomicOr(v + lane, neq0);
        if (neq1) atomicOr(v + 32 + lane, neq1);
        __syncthreads();
        // first repeat is (j, j + lambda); it only exists if j + lambda <= steps - 1
        uint32_t newly = ~v[lane] & ~done0;
        found0 |= newly & ~planes_gt(lam0, steps - 1 - j);
        done0 |= newly;
        newly = ~v[32 + lane] & ~done1;
        found1 |= newly & ~planes_gt(lam1, steps - 1 - j);
        done1 |= newly;
        if (tid < SL) votes[((j + 1) & 1) * SL + tid] = 0;
        int tmp = pc_;
        pc_ = pn_;
        pn_ = tmp;
        tmp = qc_;
        qc_ = qn_;
        qn_ = tmp;
        if (__syncthreads_and((done0 & done1) == 0xFFFFFFFFu)) break;
    }
    __syncthreads();
    if (tid < 32) {
        votes[lane] = found0;
        votes[32 + lane] = found1;
    }
    __syncthreads();
    return qc_;
}

// State layout in shared memory: word (node n, copy k, slot s) lives at ((n*4 + k)*SLOTS + s), so
// the four copies of a node sit at compile-time offsets from each other.
template <int SLOTS, int MODE>
__global__ void __launch_bounds__(kSweepThreads, (SLOTS >= 64 || kSweepThreads > 512) ? 1 : 2)
sweep_kernel(const SweepArgs args) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int N = args.n_nodes;
    const int steps = args.steps;
    const int64_t wpad = args.wpad;
    const int S = blockDim.x / SLOTS;
    uint32_t* state = smem;                                                            // [N][4][SLOTS]
    NodeOffsets* offs = reinterpret_cast<NodeOffsets*>(smem + (size_t)4 * N * SLOTS);  // [N]
    LutMasks* masks = reinterpret_cast<LutMasks*>(offs + N + kMaskPad);                // [N + pad]
    uint32_t* votes = reinterpret_cast<uint32_t*>(masks + N + kMaskPad);  // [2][kRing][SLOTS] OR-accumulators
    uint32_t* xchg = votes + 2 * kRing * SLOTS;                // [2][SLOTS]
    // cleanup only: the 32 cell indices behind each virtual word, their normal-run planes, and
    // the per-node batch prefix
    constexpr int HALF = SLOTS / 2;
    int* cellidx = reinterpret_cast<int*>(xchg + 2 * SLOTS);   // [HALF][32]
    uint32_t* nrm = reinterpret_cast<uint32_t*>(cellidx + HALF * 32);  // [1 + kLamPlanes][HALF]
    int* pre = reinterpret_cast<int*>(nrm + (1 + kLamPlanes) * HALF);  // [N + 1]
    __shared__ int s_node;
    __shared__ unsigned int s_open[2];  // DEFERS modes, shared-memory ring: lanes still open after step t (parity t & 1)

    constexpr bool IS_GATHER = MODE == MODE_CLEANUP || MODE == MODE_RECHECK;
    constexpr bool DEFERS = MODE == MODE_KOKI || MODE == MODE_RECHECK;
    const int tid = threadIdx.x;
#ifdef SCB_TRACE
    long long trace_t_ = clock64();
#endif
    const int slot = tid % SLOTS;
    const int part = tid / SLOTS;
    const int chunk = (N + S - 1) / S;
    const int n_lo = min(N, part * chunk);
    const int n_hi = min(N, n_lo + chunk);
    constexpr int NODE_STRIDE = 4 * SLOTS;
    constexpr int BATCH_CELLS = HALF * 32;  // cells of one node a cleanup batch simulates

    if (tid < kMaskPad) {  // all-zero entries past the end (ring2_registers)
        masks[N + tid] = affine_table(0u, 0);
        NodeOffsets z;
        z.pa = z.pb = z.pc = z.arity = 0;
        offs[N + tid] = z;
    }
    for (int n = tid; n < N; n += blockDim.x) {
        int arity;
        NodeEntry e = make_entry(args.net_parent, args.net_lut, n, &arity);
        NodeOffsets o;
        if (!kRingArity) arity = 3;  // (the canonical table does not depend on the unbound slots, which alias row 0)
        o.pa = e.pa * NODE_STRIDE; o.pb = e.pb * NODE_STRIDE; o.pc = e.pc * NODE_STRIDE; o.arity = arity;
        offs[n] = o;
        masks[n] = affine_table(e.lut, arity);
    }
    int total_batches = 0;
    if (IS_GATHER) {
        if (tid == 0) {
            int acc = 0;
            for (int k = 0; k < N; ++k) {
                pre[k] = acc;
                int64_t cnt = args.in_count[k];
                if (cnt > args.in_cap) cnt = args.in_cap;
                acc += (int)((cnt + BATCH_CELLS - 1) / BATCH_CELLS);
            }
            pre[N] = acc;
        }
        __syncthreads();
        total_batches = pre[N];
    }
    // KO/KI on the plane words: the short ring when the unperturbed run showed no cycle longer than 2.
    // The compacted recheck pass always runs the deep ring: a perturbation often creates cycles of
    // length 3 or 4 (half of the cells the short ring leaves open on the H1M data), and the ring settles
    // them at a fifth of the cost of the general scheme.
    // When the recheck pass follows and the register ring applies, the KO/KI pass keeps the short
    // ring whatever the normal run showed: periods 3 and 4 are the recheck pass's business.
    const bool koki_short = SLOTS == 64 && kSweepThreads >= 512 && kRegRing && args.recheck_follows &&
                            N <= (kSweepThreads / 32) * ((208 + kSweepThreads / 32 - 1) / (kSweepThreads / 32));
    const bool deep_ring = MODE == MODE_KOKI ? (!koki_short && (args.counters[1] & 1u) != 0) : true;

    int batch = blockIdx.x;
    do {
        if (IS_GATHER && batch >= total_batches) break;
        // ------------------------------------------------------------ what this slot simulates
        int64_t word;
        int forced_node = -1;
        uint32_t forced_value = 0;
        uint32_t cellmask;
        if (MODE == MODE_NORMAL) {
            word = (int64_t)blockIdx.x * SLOTS + slot;
            cellmask = word < wpad ? tail_mask(word, args.n_cells) : 0u;
        } else if (MODE == MODE_KOKI) {
            // consecutive CTAs take the SAME 32 words under consecutive forced nodes: the rows of the
            // data and of the normal-run states they read stay in L2 (and L1) for all N of them
            word = (int64_t)(blockIdx.x / (unsigned)N) * HALF + (slot % HALF);
            forced_node = (int)(blockIdx.x % (unsigned)N);
            forced_value = slot >= HALF ? 0xFFFFFFFFu : 0u;
            cellmask = word < wpad ? tail_mask(word, args.n_cells) : 0u;
            // Cells whose UNPERTURBED run already needed the general scheme (period > 4 or no repeat)
            // almost always need it under a perturbation too: they skip both ring passes and go
            // straight to the cleanup list.  (Any split of the cells among the passes gives the same
            // result -- every pass is exact -- so this is only about where the time goes.)
            const uint32_t hard = word < wpad ? __ldg(args.np_hard + word) & cellmask : 0u;
            if (hard && part == 0 && slot < HALF) {
                const unsigned int k = __popc(hard);
                const unsigned int at0 = atomicAdd(&args.hard_count[forced_node], k);
                int32_t* dst = args.hard_cells + (size_t)forced_node * args.hard_cap + at0;  // cap = n_cells: always room
                for (uint32_t m = hard; m; m &= m - 1) *dst++ = (int32_t)(word * 32 + (__ffs(m) - 1));
            }
            cellmask &= ~hard;
        } else {
            // a virtual word: 32 deferred cells of one node, gathered from anywhere in the data
            if (tid == 0) {
                int lo = 0, hi = N;  // largest k with pre[k] <= batch
                while (hi - lo > 1) {
                    int mid = (lo + hi) >> 1;
                    if (pre[mid] <= batch) lo = mid; else hi = mid;
                }
                s_node = lo;
            }
            __syncthreads();
            forced_node = s_node;
            forced_value = slot >= HALF ? 0xFFFFFFFFu : 0u;
            int64_t cnt = args.in_count[forced_node];
            if (cnt > args.in_cap) cnt = args.in_cap;
            // the batch's cells are consecutive list entries: coalesced copy, virtual word v = entries
            // [32v, 32v + 32)
            const int64_t batch_first = (int64_t)(batch - pre[forced_node]) * BATCH_CELLS;
            for (int i = tid; i < BATCH_CELLS; i += blockDim.x)
                cellidx[i] = batch_first + i < cnt ? args.in_cells[(size_t)forced_node * args.in_cap + batch_first + i] : -1;
            __syncthreads();
            word = 0;
            const int64_t left = cnt - (batch_first + (int64_t)(slot % HALF) * 32);
            cellmask = left >= 32 ? 0xFFFFFFFFu : (left > 0 ? (1u << (int)left) - 1u : 0u);
        }
        const bool word_ok = word < wpad;
        for (int i = tid; i < 2 * kRing * SLOTS; i += blockDim.x) votes[i] = 0;

        auto at = [&](int k, int n) -> uint32_t& { return state[(n * 4 + k) * SLOTS + slot]; };
        // gather modes: copy k of every node, both halves, from the cell-major records of the batch's
        // cells; `kstage` names a state copy that is dead at the call (staging area).  One warp per
        // virtual word: lane = cell, 128-bit loads fetch its record, every record word goes through a
        // 32x32 warp transpose (lane b gets the word of node 32j + b) and is parked in copy kstage at a rotated slot
        // position -- lanes hold different nodes, whose rows all start on the same bank, and the
        // rotation by the node index spreads them over the banks.  After a barrier every thread
        // moves the words of its own (slot, nodes) into copy k, again conflict-free.
        auto load_cells = [&](const uint32_t* __restrict__ src_t, int k, int kstage) {
            const int lane = tid & 31;
            const int nwp = args.nwp;
            const int n_warps = blockDim.x >> 5;
            for (int v0 = tid >> 5; v0 < HALF; v0 += 2 * n_warps) {
                // two virtual words per round: their record loads fly together
                const int v1 = v0 + n_warps;
                const int c0 = cellidx[v0 * 32 + lane];
                const int c1 = v1 < HALF ? cellidx[v1 * 32 + lane] : -1;
                for (int j4 = 0; j4 < nwp; j4 += 4) {
                    uint4 x0 = make_uint4(0u, 0u, 0u, 0u), x1 = x0;
                    if (c0 >= 0) x0 = __ldg(reinterpret_cast<const uint4*>(src_t + (size_t)c0 * nwp + j4));
                    if (c1 >= 0) x1 = __ldg(reinterpret_cast<const uint4*>(src_t + (size_t)c1 * nwp + j4));
                    const uint32_t xw[2][4] = {{x0.x, x0.y, x0.z, x0.w}, {x1.x, x1.y, x1.z, x1.w}};
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int v = h ? v1 : v0;
                        if (v >= HALF) break;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int nb = (j4 + q) * 32;
                            if (nb < N) {
                                // lane = cell holds 32 node bits; after the transpose lane b holds the 32
                                // cell bits of node nb + b
                                const uint32_t mine = warp_transpose32(xw[h][q], lane);
                                const int n = nb + lane;
                                if (n < N) state[(n * 4 + kstage) * SLOTS + ((v + n) % HALF)] = mine;
                            }
                        }
                    }
                }
            }
            __syncthreads();
            for (int n = n_lo; n < n_hi; ++n)
                at(k, n) = state[(n * 4 + kstage) * SLOTS + ((slot + n) % HALF)];
        };
        auto load_data = [&](int k, int kstage) {
            if (IS_GATHER) {
                load_cells(args.planes_t, k, kstage);
                return;
            }
            for (int n = n_lo; n < n_hi; ++n)
                at(k, n) = word_ok ? __ldg(args.planes + (size_t)n * wpad + word) : 0u;
        };
        // next value of node n for my slot, reading state copy k
        auto step_node = [&](int k, int n) {
            const NodeOffsets o = offs[n];
            const uint32_t* st = state + k * SLOTS + slot;
            uint32_t f;
            if (o.arity <= 1) {
                f = affine(st[o.pa], masks[n].m[0], masks[n].m[1]);
            } else {
                const uint4 lo = *reinterpret_cast<const uint4*>(&masks[n].m[0]);
                const uint32_t a = st[o.pa], b = st[o.pb];
                if (o.arity == 2) {
                    f = lop3<0xCA>(a, affine(b, lo.z, lo.w), affine(b, lo.x, lo.y));
                } else {
                    const uint4 hi = *reinterpret_cast<const uint4*>(&masks[n].m[4]);
                    const uint32_t c = st[o.pc];
                    f = lop3<0xCA>(a, lop3<0xCA>(b, affine(c, hi.z, hi.w), affine(c, hi.x, hi.y)),
                                   lop3<0xCA>(b, affine(c, lo.z, lo.w), affine(c, lo.x, lo.y)));
                }
            }
            if (MODE != MODE_NORMAL && n == forced_node) f = forced_value;
            return f;
        };

        uint32_t found = 0;        // lanes with a first repeat inside the window
        uint32_t lam[kLamPlanes];  // its cycle length, bit-sliced
#pragma unroll
        for (int b = 0; b < kLamPlanes; ++b) lam[b] = 0;
        int xs = 0, xn = 1;        // state copy frozen at state_mu for found lanes / its scratch partner
        uint32_t unresolved = cellmask;

        // -------------------------------------------------------------------- fast path
        // Ring of the last states: the new state is compared with its predecessors while it is
        // computed (the oldest one lives in the very copy that is being overwritten).  A lane
        // whose state equals state_{t-d} (smallest d) has its first repeat at (t-d, t): it is at
        // state_mu right now and is frozen there.  Exact for cycle lengths <= ring depth (4, or
        // 2 when the unperturbed run showed no longer cycles; anything else is settled later).
        if (MODE != MODE_CLEANUP) {
            const bool deep = deep_ring;
            SCB_SWEEP_PHASE(0);
            load_data(kRing - 1, 0);
            __syncthreads();
            SCB_SWEEP_PHASE(1);
            uint32_t frozen = ~cellmask;
            uint32_t key[6] = {0u, 0u, 0u, 0u, 0u, 0u};  // probe: settle step of every lane, bit-sliced
            // shared-memory ring of a pass that can hand lanes on: stop once lanes had started to settle and
            // none did for a while (what is left are cycles the ring cannot see, or nothing at all)
            unsigned int last_open = 0xFFFFFFFFu;
            int stall = 0;
            bool moved = false;
            if (DEFERS && tid < 2) s_open[tid] = 0u;  // (ordered before the first add by the barrier after load_data)
            const int t_fast = MODE != MODE_KOKI      ? steps
                               : args.fast_steps > 0 ? min(steps, args.fast_steps)
                                                     : min(steps, (int)args.counters[3]);
            // nodes per warp of the register ring: 13 / 8 / 4 with 16 warps, 7 / 4 / 2 with 32
            constexpr int kWarps = kSweepThreads / 32;
            constexpr int CH_L = (208 + kWarps - 1) / kWarps, CH_M = (128 + kWarps - 1) / kWarps, CH_S = (64 + kWarps - 1) / kWarps;
            const bool reg_ring = DEFERS && SLOTS == 64 && kWarps >= 16 && !deep && N <= kWarps * CH_L && kRegRing;
            if (reg_ring) {
                // the recheck pass is the last ring pass: it only stops early when nothing is left
                const int so = MODE == MODE_KOKI ? args.stop_open : 0, ss = MODE == MODE_KOKI ? args.stop_stall : (1 << 30);
                xs = N <= kWarps * CH_S   ? ring2_registers<CH_S>(state, offs, masks, votes, N, forced_node, cellmask, t_fast, tid, so, ss)
                     : N <= kWarps * CH_M ? ring2_registers<CH_M>(state, offs, masks, votes, N, forced_node, cellmask, t_fast, tid, so, ss)
                                          : ring2_registers<CH_L>(state, offs, masks, votes, N, forced_node, cellmask, t_fast, tid, so, ss);
                frozen = votes[slot];
                lam[0] = votes[64 + slot];
                lam[1] = votes[128 + slot];
                __syncthreads();  // votes is reused below
            }
            for (int t = 0; t < (reg_ring ? 0 : t_fast); ++t) {
                const int ks = (t + 3) & 3, kd = t & 3, k2 = (t + 2) & 3, k3 = (t + 1) & 3;
                uint32_t neq1 = 0, neq2 = 0, neq3 = 0, neq4 = 0;
                if (deep) {
#pragma unroll 2
                    for (int n = n_lo; n < n_hi; ++n) {
                        const uint32_t f = step_node(ks, n);
                        uint32_t* base = state + n * NODE_STRIDE + slot;
                        const uint32_t old1 = base[ks * SLOTS];
                        neq1 |= f ^ old1;
                        neq2 |= f ^ base[k2 * SLOTS];
                        neq3 |= f ^ base[k3 * SLOTS];
                        neq4 |= f ^ base[kd * SLOTS];
                        base[kd * SLOTS] = (f & ~frozen) | (old1 & frozen);
                    }
                } else {
#pragma unroll 2
                    for (int n = n_lo; n < n_hi; ++n) {
                        const uint32_t f = step_node(ks, n);
                        uint32_t* base = state + n * NODE_STRIDE + slot;
                        const uint32_t old1 = base[ks * SLOTS];
                        neq1 |= f ^ old1;
                        neq2 |= f ^ base[k2 * SLOTS];
                        base[kd * SLOTS] = (f & ~frozen) | (old1 & frozen);
                    }
                    neq3 = neq4 = 0xFFFFFFFFu;
                }
                uint32_t* v = votes + (t & 1) * kRing * SLOTS;
                if (neq1) atomicOr(v + 0 * SLOTS + slot, neq1);
                if (neq2) atomicOr(v + 1 * SLOTS + slot, neq2);
                if (neq3) atomicOr(v + 2 * SLOTS + slot, neq3);
                if (neq4) atomicOr(v + 3 * SLOTS + slot, neq4);
                __syncthreads();
                uint32_t open = ~frozen;
                // state_{t-d} exists for t >= d only (the raw data state is never compared)
                uint32_t e1 = t >= 1 ? ~v[0 * SLOTS + slot] & open : 0u;
                open &= ~e1;
                uint32_t e2 = t >= 2 ? ~v[1 * SLOTS + slot] & open : 0u;
                open &= ~e2;
                uint32_t e3 = t >= 3 ? ~v[2 * SLOTS + slot] & open : 0u;
                open &= ~e3;
                uint32_t e4 = t >= 4 ? ~v[3 * SLOTS + slot] & open : 0u;
                lam[0] |= e1 | e3;
                lam[1] |= e2 | e3;
                lam[2] |= e4;
                frozen |= e1 | e2 | e3 | e4;
                if (MODE == MODE_NORMAL && part == 0) {  // settle-time histogram (choose_fast_steps_kernel)
                    // only periods <= 2 count as settled: that is what the KO/KI pass's short ring can see
                    const unsigned int c = __reduce_add_sync(0xFFFFFFFFu, (unsigned int)__popc(e1 | e2));
                    if (args.probe) {
                        // sort key: the settle step for periods 1 and 2 (at most 61), 62 for periods 3 and 4
                        const int kt = min(t, 61);
                        const unsigned int c34 = __reduce_add_sync(0xFFFFFFFFu, (unsigned int)__popc(e3 | e4));
#pragma unroll
                        for (int b = 0; b < 6; ++b) key[b] |= (((kt >> b) & 1) ? (e1 | e2) : 0u) | (((62 >> b) & 1) ? (e3 | e4) : 0u);
                        if ((tid & 31) == 0 && c) atomicAdd(&args.key_hist[kt], c);
                        if ((tid & 31) == 0 && c34) atomicAdd(&args.key_hist[62], c34);
                    } else if ((tid & 31) == 0 && c) {
                        atomicAdd(&args.settle_hist[t], c);
                    }
                }
                // the other accumulator set is idle between the two barriers: clear it for step t+1
                if (part < kRing) votes[((t + 1) & 1) * kRing * SLOTS + part * SLOTS + slot] = 0;
                xs = kd;
                if (DEFERS) {
                    if (part == 0) {
                        const unsigned int c = __reduce_add_sync(0xFFFFFFFFu, (unsigned int)__popc(~frozen));
                        if ((tid & 31) == 0 && c) atomicAdd(&s_open[t & 1], c);
                    }
                    if (tid == 0) s_open[(t + 1) & 1] = 0u;  // last read after the previous step's closing barrier
                }
                if (__syncthreads_and(frozen == 0xFFFFFFFFu)) break;
                if (DEFERS) {
                    const unsigned int n_open = s_open[t & 1];  // the same number in every thread
                    if (n_open < last_open) {
                        moved = last_open != 0xFFFFFFFFu;
                        last_open = n_open;
                        stall = 0;
                    } else if (moved && ++stall >= args.ring_stall) {
                        break;
                    }
                }
            }
            xn = (xs + 1) & 3;
            found = frozen & cellmask;
            unresolved = ~frozen;
            if (MODE == MODE_NORMAL && args.probe) {
                // key 63: the ring settled nothing (long cycle or no repeat); such cells end up last
                if (part == 0) {
                    const unsigned int c = __reduce_add_sync(0xFFFFFFFFu, (unsigned int)__popc(unresolved & cellmask));
                    if ((tid & 31) == 0 && c) atomicAdd(&args.key_hist[63], c);
                    if (word_ok) {
#pragma unroll
                        for (int b = 0; b < 6; ++b) args.key_planes[(size_t)b * wpad + word] = key[b] | (unresolved & cellmask);
                    }
                }
                return;
            }
            if (MODE == MODE_NORMAL && part == 0 && word_ok) args.np_hard[word] = unresolved & cellmask;
            SCB_SWEEP_PHASE(2);
        }

        // ------------------------------------- KO/KI, recheck: hand unsettled cells to the next pass
        if (DEFERS) {
            if (part == 0) xchg[slot] = unresolved;
            __syncthreads();
            const uint32_t both = unresolved | xchg[(slot + HALF) % SLOTS];  // KO and KI travel together
            if (part == 0 && slot < HALF) {
                uint32_t deferred = 0;
                if (both) {  // reserve room in the node's list; a full list keeps the cells here
                    const unsigned int k = __popc(both);
                    const unsigned int at0 = atomicAdd(&args.out_count[forced_node], k);
                    if ((int64_t)at0 + k <= args.out_cap) {
                        int32_t* dst = args.out_cells + (size_t)forced_node * args.out_cap + at0;
                        for (uint32_t m = both; m; m &= m - 1) {
                            const int ln = __ffs(m) - 1;
                            *dst++ = IS_GATHER ? cellidx[slot * 32 + ln] : (int32_t)(word * 32 + ln);
                        }
                        deferred = both;
                    } else {
                        atomicSub(&args.out_count[forced_node], k);
                    }
                }
                xchg[SLOTS + slot] = deferred;
            }
            __syncthreads();
            const uint32_t deferred = xchg[SLOTS + (slot % HALF)];
            cellmask &= ~deferred;
            found &= cellmask;
            unresolved = both & ~deferred;
            __syncthreads();  // xchg is reused below
            SCB_SWEEP_PHASE(3);
        }

        // -------------------------------------------------------------------- slow path
        // General scheme for the lanes in `cellmask` that are still open: Brent checkpoints at
        // t = 2^k - 1 give lambda, then two replays from the data (P runs lambda_cell steps ahead
        // of Q under per-cell stall masks) give mu and leave Q frozen at state_mu.
        if (MODE == MODE_CLEANUP || __syncthreads_or(unresolved != 0)) {
            const int cur0 = 0, nxt0 = 1, chk = 2;
            int cur = cur0, nxt = nxt0;
            for (int i = tid; i < 2 * kRing * SLOTS; i += blockDim.x) votes[i] = 0;
            load_data(cur, nxt);
            __syncthreads();
            uint32_t resolved = ~cellmask;
#pragma unroll
            for (int b = 0; b < kLamPlanes; ++b) lam[b] = 0;
            uint32_t cyc = 0, found_fp = 0;
            int chk_time = -1;
            constexpr int kWarpsB = kSweepThreads / 32;
            constexpr int CHB_L = (208 + kWarpsB - 1) / kWarpsB, CHB_M = (128 + kWarpsB - 1) / kWarpsB, CHB_S = (64 + kWarpsB - 1) / kWarpsB;
            const bool reg_brent = MODE == MODE_CLEANUP && SLOTS == 64 && kWarpsB >= 16 && N <= kWarpsB * CHB_L && kRegRing && kRegBrent;
            if (reg_brent) {
                cur = N <= kWarpsB * CHB_S   ? brent_registers<CHB_S>(state, offs, masks, votes, N, forced_node, cellmask, steps, tid)
                      : N <= kWarpsB * CHB_M ? brent_registers<CHB_M>(state, offs, masks, votes, N, forced_node, cellmask, steps, tid)
                                             : brent_registers<CHB_L>(state, offs, masks, votes, N, forced_node, cellmask, steps, tid);
                nxt = cur ^ 1;
#pragma unroll
                for (int b = 0; b < kLamPlanes; ++b) lam[b] = votes[b * SLOTS + slot];
                cyc = votes[6 * SLOTS + slot];
                found_fp = votes[7 * SLOTS + slot];
                __syncthreads();  // (the lambda planes stay in `votes` for replay_registers)
            }
            // Checkpoints at t = 2^k - 1 < steps - 1 and a LAST one at t = steps - 1: a cell with a
            // repeat inside the window (mu + lambda <= steps - 1) is on its cycle by then and is back
            // at that state lambda <= steps - 1 steps later, so 2 * steps - 1 steps settle every cell
            // (a power-of-two schedule alone needs up to 2^k - 1 + steps with 2^k - 1 >= steps - 2).
            const int t1_max = reg_brent ? 0 : 2 * steps - 1;
            for (int t = 0; t < t1_max; ++t) {
                uint32_t neq_prev = 0, neq_chk = 0;
#pragma unroll 2
                for (int n = n_lo; n < n_hi; ++n) {
                    const uint32_t f = step_node(cur, n);
                    neq_prev |= f ^ at(cur, n);
                    if (chk_time >= 0) neq_chk |= f ^ at(chk, n);
                    at(nxt, n) = f;
                }
                uint32_t* v = votes + (t & 1) * kRing * SLOTS;
                if (neq_prev) atomicOr(v + slot, neq_prev);
                if (neq_chk) atomicOr(v + SLOTS + slot, neq_chk);
                __syncthreads();
                neq_prev = v[slot];
                neq_chk = v[SLOTS + slot];
                if (t >= 1) {
                    uint32_t newly = ~neq_prev & ~resolved;
                    lam[0] |= newly;  // lambda = 1
                    resolved |= newly;
                    if (t <= steps - 1) found_fp |= newly;
                }
                if (chk_time >= 0) {
                    uint32_t newly = ~neq_chk & ~resolved;
                    int l = t - chk_time;
                    if (l <= steps - 1) {
                        planes_set(lam, newly, l);
                        cyc |= newly;
                    }
                    resolved |= newly;  // l >= steps: such a cell can never repeat inside the window
                }
                if ((((t + 1) & t) == 0 && t < steps - 1) || t == steps - 1) {  // new checkpoint = state_t (own nodes only)
                    for (int n = n_lo; n < n_hi; ++n) at(chk, n) = at(nxt, n);
                    chk_time = t;
                }
                if (part < 2) votes[((t + 1) & 1) * kRing * SLOTS + part * SLOTS + slot] = 0;
                int tmp = cur; cur = nxt; nxt = tmp;
                if (__syncthreads_and(resolved == 0xFFFFFFFFu)) break;
            }
            if (reg_brent && __syncthreads_or(cyc != 0)) {
                load_data(0, 1);
                load_data(2, 3);
                __syncthreads();
                xs = N <= kWarpsB * CHB_S   ? replay_registers<CHB_S>(state, offs, masks, votes, N, forced_node, cellmask, steps, tid)
                     : N <= kWarpsB * CHB_M ? replay_registers<CHB_M>(state, offs, masks, votes, N, forced_node, cellmask, steps, tid)
                                            : replay_registers<CHB_L>(state, offs, masks, votes, N, forced_node, cellmask, steps, tid);
                xn = xs ^ 1;
                found = votes[slot];
                __syncthreads();
            } else if (!reg_brent && __syncthreads_or(cyc != 0)) {
                int pc_ = 0, pn_ = 1;  // P: runs lambda ahead
                int qc_ = 2, qn_ = 3;  // Q: freezes at state_mu
                load_data(pc_, pn_);
                load_data(qc_, qn_);
                for (int i = tid; i < 2 * kRing * SLOTS; i += blockDim.x) votes[i] = 0;
                __syncthreads();
                const uint32_t has_lam = planes_ge(lam, 1) & cellmask;
                for (int s = 0; s < steps; ++s) {
                    uint32_t adv = planes_gt(lam, s) & has_lam;  // lanes that still owe P a step
                    if (!__syncthreads_or(adv != 0)) break;
#pragma unroll 2
                    for (int n = n_lo; n < n_hi; ++n) {
                        const uint32_t f = step_node(pc_, n);
                        at(pn_, n) = (f & adv) | (at(pc_, n) & ~adv);
                    }
                    __syncthreads();
                    int tmp = pc_; pc_ = pn_; pn_ = tmp;
                }
                uint32_t done = ~has_lam;
                found = 0;
                for (int j = 0; j <= steps - 2; ++j) {
                    uint32_t neq = 0;
#pragma unroll 2
                    for (int n = n_lo; n < n_hi; ++n) {
                        const uint32_t fp = step_node(pc_, n);
                        const uint32_t fq = step_node(qc_, n);
                        neq |= fp ^ fq;
                        at(pn_, n) = fp;
                        at(qn_, n) = (at(qc_, n) & done) | (fq & ~done);
                    }
                    uint32_t* v = votes + (j & 1) * kRing * SLOTS;
                    if (neq) atomicOr(v + slot, neq);
                    __syncthreads();
                    neq = v[slot];
                    uint32_t newly = ~neq & ~done;
                    // first repeat is (j, j + lambda); it only exists if j + lambda <= steps - 1
                    uint32_t eligible = ~planes_gt(lam, steps - 1 - j);
                    found |= newly & eligible;
                    done |= newly;
                    if (part == 0) votes[((j + 1) & 1) * kRing * SLOTS + slot] = 0;
                    int tmp = pc_; pc_ = pn_; pn_ = tmp;
                    tmp = qc_; qc_ = qn_; qn_ = tmp;
                    if (__syncthreads_and(done == 0xFFFFFFFFu)) break;
                }
                xs = qc_;
                xn = qn_;
            } else {
                found = found_fp;  // every settled lane sits on its fixed point already
                xs = cur;
                xn = nxt;
            }
            found &= cellmask;
        }

        SCB_SWEEP_PHASE(4);
#ifdef SCB_TRACE
        if (MODE == MODE_CLEANUP && part == 0) {  // histogram of what the general scheme had to settle
            const uint32_t f = found & cellmask;
            atomicAdd(&g_sweep_phase[MODE][8], (unsigned long long)__popc(~found & cellmask));           // no repeat
            atomicAdd(&g_sweep_phase[MODE][9], (unsigned long long)__popc(f & ~planes_ge(lam, 3)));        // 1, 2
            atomicAdd(&g_sweep_phase[MODE][10], (unsigned long long)__popc(f & planes_ge(lam, 3) & ~planes_ge(lam, 5)));  // 3, 4
            atomicAdd(&g_sweep_phase[MODE][11], (unsigned long long)__popc(f & planes_ge(lam, 5) & ~planes_ge(lam, 9)));  // 5..8
            atomicAdd(&g_sweep_phase[MODE][12], (unsigned long long)__popc(f & planes_ge(lam, 9)));        // > 8
        }
#endif
        // -------------------------------------------------------------------- outputs
        if (MODE == MODE_NORMAL) {
            if (part == 0 && word_ok) {
                args.np_found[word] = found;
#pragma unroll
                for (int b = 0; b < kLamPlanes; ++b) args.np_lam[(size_t)b * wpad + word] = lam[b];
                uint32_t miss = __popc(~found & cellmask);
                if (miss) atomicAdd(args.out_missing, (unsigned long long)miss);
                // tell the KO/KI pass whether the short ring is enough for this data set
                if ((~found & cellmask) || (planes_ge(lam, 3) & found)) atomicOr(&args.counters[1], 1u);
                // last slice of np_states any cell of this word needs (offsets 0..lambda): an upper bound
                // from the highest lambda plane in use is enough
                int hb = -1;
#pragma unroll
                for (int b = 0; b < kLamPlanes; ++b)
                    if (lam[b] & found) hb = b;
                if (hb >= 0) atomicMax(&args.counters[2], (unsigned int)min(steps - 1, (2 << hb) - 1));
            }
            for (int t = 0; t < steps; ++t) {
                uint32_t need = found & planes_ge(lam, t);
                if (!__syncthreads_or(need != 0)) break;
                if (word_ok)
                    for (int n = n_lo; n < n_hi; ++n)
                        args.np_states[((size_t)t * N + n) * wpad + word] = at(xs, n);
                uint32_t need_next = found & planes_ge(lam, t + 1);
                if (!__syncthreads_or(need_next != 0)) break;
#pragma unroll 2
                for (int n = n_lo; n < n_hi; ++n) at(xn, n) = step_node(xs, n);
                __syncthreads();
                int tmp = xs; xs = xn; xn = tmp;
            }
        } else {
            // a cell counts only if its KO run AND its KI run repeat (and the normal one)
            if (part == 0) xchg[slot] = found;
            __syncthreads();
            const uint32_t partner = xchg[(slot + HALF) % SLOTS];
            uint32_t found_n;
            uint32_t lamn[kLamPlanes];
            if (IS_GATHER) {
                // one warp per virtual word: lane = cell, its normal-run record (bit 0..5 = lambda, bit 6 =
                // found) is one 4-byte load; seven ballots make the planes
                for (int v = tid >> 5; v < HALF; v += blockDim.x >> 5) {
                    const int c = cellidx[v * 32 + (tid & 31)];
                    const uint32_t rec = c >= 0 ? __ldg(args.np_info_t + c) : 0u;
#pragma unroll
                    for (int b = 0; b <= kLamPlanes; ++b) {
                        const uint32_t bits = __ballot_sync(0xFFFFFFFFu, (rec >> b) & 1u);
                        if ((tid & 31) == 0) nrm[(b == kLamPlanes ? 0 : 1 + b) * HALF + v] = bits;
                    }
                }
                __syncthreads();
                found_n = nrm[slot % HALF];
#pragma unroll
                for (int b = 0; b < kLamPlanes; ++b) lamn[b] = nrm[(1 + b) * HALF + (slot % HALF)];
            } else {
                found_n = word_ok ? __ldg(args.np_found + word) : 0u;
#pragma unroll
                for (int b = 0; b < kLamPlanes; ++b)
                    lamn[b] = word_ok ? __ldg(args.np_lam + (size_t)b * wpad + word) : 0u;
            }
            const uint32_t valid = found & partner & found_n;
            if (part == 0) {
                uint32_t miss = __popc(~found & cellmask);
                if (miss)
                    atomicAdd(args.out_missing + 1 + 2 * forced_node + (slot >= HALF ? 1 : 0),
                              (unsigned long long)miss);
                if (slot < HALF) {
                    uint32_t bad = __popc(found & partner & ~found_n);
                    if (bad) atomicAdd(args.out_keyerror, (unsigned long long)bad);
                }
            }
#ifdef SCB_TRACE
            if (part == 0) {  // scoring statistics: lanes scored, lanes where either run cycles, lanes with lambda 2
                atomicAdd(&g_sweep_phase[MODE][8 + 5], (unsigned long long)__popc(valid));
                atomicAdd(&g_sweep_phase[MODE][8 + 6],
                          (unsigned long long)__popc(valid & (planes_ge(lam, 2) | planes_ge(lamn, 2))));
            }
#endif
            unsigned long long score = 0;
            // offset 0: the frozen states against the normal run's state_mu.  Gather modes assemble the
            // normal-run states of the batch's cells in the state copy that neither xs nor xn uses.
            if (__syncthreads_or(valid != 0)) {
                if (IS_GATHER) {
                    load_cells(args.np_states_t, xs ^ 2, xn ^ 2);
                    __syncthreads();
                }
                // lanes whose perturbed AND normal run have period 2: offset 2 repeats offset 0
                const uint32_t both2 = valid & planes_ge(lam, 2) & planes_ge(lamn, 2);
                uint32_t sc_both2 = 0;
                if (valid) {
                    uint32_t sc = 0;
#pragma unroll 4
                    for (int n = n_lo; n < n_hi; ++n) {
                        const uint32_t ref = IS_GATHER ? at(xs ^ 2, n) : __ldg(args.np_states + (size_t)n * wpad + word);
                        const uint32_t x = (at(xs, n) ^ ref) & valid;
                        sc += __popc(x);
                        sc_both2 += __popc(x & both2);
                    }
                    score = sc;
                }
                if (!__syncthreads_or((valid & (planes_ge(lam, 2) | planes_ge(lamn, 2))) != 0)) {
                    score *= 2;  // every counted cell: fixed point vs fixed point, offset 1 repeats offset 0
                } else {
                    // offsets 1..min(lambda_x, lambda_n): step, and compare the new state while it is made.
                    // With no period above 2 in the CTA, offset 2 is offset 0 again for the period-2 pairs
                    // (state_mu+2 = state_mu on both sides) and nothing else is active: one step suffices.
                    const bool short_cycles = !__syncthreads_or((valid & (planes_ge(lam, 3) | planes_ge(lamn, 3))) != 0);
                    for (int t = 1; t < steps; ++t) {
                        if (short_cycles && t == 2) {
                            score += sc_both2;
                            break;
                        }
                        const uint32_t act = valid & planes_ge(lam, t) & planes_ge(lamn, t);
                        if (!__syncthreads_or(act != 0)) break;
                        const int kref = xs ^ 2;
                        if (IS_GATHER) {
                            load_cells(args.np_states_t + (size_t)t * wpad * 32 * args.nwp, kref, xn ^ 2);
                            __syncthreads();
                        }
                        uint32_t sc = 0;
#pragma unroll 2
                        for (int n = n_lo; n < n_hi; ++n) {
                            const uint32_t ref =
                                IS_GATHER ? at(kref, n) : __ldg(args.np_states + ((size_t)t * N + n) * wpad + word);
                            const uint32_t f = step_node(xs, n);
                            at(xn, n) = f;
                            sc += __popc((f ^ ref) & act);
                        }
                        score += sc;
                        __syncthreads();
                        int tmp = xs; xs = xn; xn = tmp;
                    }
                }
            }
            // one node per CTA (and per cleanup batch): fold the warp first
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) score += __shfl_xor_sync(0xFFFFFFFFu, score, off);
            if ((tid & 31) == 0 && score) atomicAdd(args.out_raw + forced_node, score);
        }
        SCB_SWEEP_PHASE(5);
#ifdef SCB_TRACE
        if (tid == 0) atomicAdd(&g_sweep_phase[MODE][15], 1ull);
#endif
        batch += gridDim.x;
        if (IS_GATHER) __syncthreads();
    } while (IS_GATHER);
}

// ---------------------------------------------------------------------------------------
// KO/KI on the plane words, two CTAs per SM (the default KO/KI pass when the recheck pass follows).
//
// Same work and the same results as sweep_kernel<64, MODE_KOKI> with the register ring -- a CTA is
// (forced node, 32 words), slot 2l = KO and slot 2l + 1 = KI of word l, ring of 2: a lane whose new
// state equals state_{t-1} or state_{t-2} has its first repeat there and is frozen at it, lanes still
// open when the CTA stops go to the recheck list -- laid out for occupancy instead of for one fat CTA:
//  * shared memory holds TWO state copies, not four: the copy a step writes holds state_{t-2}, and the
//    owner of a (node, slot) word reads it just before it overwrites it -- nobody else touches that
//    copy during the step.  102 KB for 200 nodes x 64 slots: two CTAs (32 warps) per SM, so the
//    dependent LDS -> IMAD -> LOP3 chain of one warp is covered by seven others per scheduler
//    instead of three;
//  * the KO and the KI word of a (node, word) are neighbours: one 64-bit load / store serves both;
//  * parent offsets are byte offsets (one add per parent), the forced node is an ordinary one-parent
//    node whose parent is a constant row (KO slots 0, KI slots ~0): nothing about it in the loop;
//  * nodes are dealt to the 16 warps evenly (N/16 or N/16 + 1 each);
//  * no slow path: the deferral lists hold every cell (sweep_list_capacity), a lane the ring cannot
//    settle is always handed on.
// Scoring needs at most one more step: a lane frozen by this ring has period 1 or 2, so its states at
// offsets 0, 1, 2 are s, step(s), s (align_attractors, importance_scores.py:325-340).
// ---------------------------------------------------------------------------------------
#ifndef SCB_KK_REG_NODES
#define SCB_KK_REG_NODES 10  // measured on imp-1M: 58.6 / 58.0 / 58.1 / 57.2 / 58.8 ms for 0 / 4 / 6 / 10 / 13
#endif
constexpr int kKkThreads = 512, kKkWarps = 16, kKkSlots = 64;
constexpr int kKkRow = 2 * kKkSlots;  // words per node: [copy][slot]
constexpr int kKkRowBytes = kKkRow * 4, kKkCopyBytes = kKkSlots * 4;
constexpr int kKkTabPad = 16;

// Table entry of a node, 16 bytes, read with broadcast loads (one wavefront for 8 bytes, two for 16):
//   w0  parent rows A, B, C (bytes 0-2) and the number of bound parents (byte 3)
//   w1  affine table pairs 0 and 1 as bytes S0, O0, S1, O1 (affine_table: S in {0, 1, -1}, O in {0, ~0})
//   w2  pairs 2 and 3 (three-parent nodes only)
// A byte becomes a register operand with one PRMT (sign extension for S, sign replication for O), so a
// node costs one or two wavefronts of table traffic instead of four to six.
struct __align__(16) KkEntry {
    uint32_t w0, w1, w2, w3;
};
constexpr int kKkEntryBytes = (int)sizeof(KkEntry);  // 16
__device__ __forceinline__ uint32_t kk_pack_pairs(const LutMasks& t, int first) {
    return (t.m[first] & 0xFFu) | ((t.m[first + 1] & 0xFFu) << 8) | ((t.m[first + 2] & 0xFFu) << 16) | ((t.m[first + 3] & 0xFFu) << 24);
}
__device__ __forceinline__ KkEntry kk_make_entry(int pa, int pb, int pc, int arity, const LutMasks& t) {
    KkEntry e;
    e.w0 = (uint32_t)pa | ((uint32_t)pb << 8) | ((uint32_t)pc << 16) | ((uint32_t)arity << 24);
    e.w1 = kk_pack_pairs(t, 0);
    e.w2 = kk_pack_pairs(t, 4);
    e.w3 = 0u;
    return e;
}
// byte k of w as a multiplier (sign-extended) / as a full-word mask (its sign bit replicated): prmt.b32 with
// bit 3 of a selector nibble set copies the sign of the selected byte (__byte_perm drops that bit)
template <uint32_t SEL>
__device__ __forceinline__ uint32_t kk_prmt(uint32_t w) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0u), "n"(SEL));
    return r;
}
template <int K>
__device__ __forceinline__ uint32_t kk_s(uint32_t w) { return kk_prmt<0x8880u | (uint32_t)K | ((uint32_t)K << 4) | ((uint32_t)K << 8) | ((uint32_t)K << 12)>(w); }
template <int K>
__device__ __forceinline__ uint32_t kk_o(uint32_t w) { return kk_prmt<0x8888u | (uint32_t)K | ((uint32_t)K << 4) | ((uint32_t)K << 8) | ((uint32_t)K << 12)>(w); }
// shared-space address of parent row K of entry word w0 for the slot pair behind `srd`
template <int K, int ROWB>
__device__ __forceinline__ uint32_t kk_parent(uint32_t w0, uint32_t srd) { return srd + kk_prmt<0x4440u | (uint32_t)K>(w0) * (uint32_t)ROWB; }

// nodes per warp: the node rows are padded with do-nothing nodes to a multiple of the warp count, so every warp
// owns exactly the same number and the unrolled loop has no "does this node exist" test
static inline int koki_plane_cnt(int n_nodes) { return (n_nodes + kKkWarps - 1) / kKkWarps; }
static size_t koki_plane_smem_bytes(int n_nodes) {
    const int n_pad = koki_plane_cnt(n_nodes) * kKkWarps;
    return (size_t)(n_pad + 1) * kKkRowBytes + (size_t)(n_pad + kKkTabPad) * sizeof(KkEntry) + (size_t)3 * 2 * kKkSlots * 4;
}

// shared-memory accesses by 32-bit shared-space address.  Loads and stores are volatile asm: they stay in the
// order they are written, which is how a node's loads get issued together ahead of the previous node's store.
__device__ __forceinline__ uint2 kk_lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void kk_sts64(uint32_t addr, uint2 v) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}

// Value of a node for one slot pair: words w0, w1 of its table entry were fetched one node ahead, parents are
// read at `srd` (copy, slot pair).  The branch on the number of parents is warp-uniform and does not wait
// for a load; each arm starts by issuing all of its loads.
__device__ __forceinline__ uint32_t kk_lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <int ROWB = kKkRowBytes>  // bytes between the rows of two nodes
__device__ __forceinline__ uint2 kk_eval(const uint2 e, uint32_t entry, uint32_t srd) {
    uint2 f;
    if (e.x < (2u << 24)) {
        const uint2 a = kk_lds64(kk_parent<0, ROWB>(e.x, srd));
        const uint32_t s = kk_s<0>(e.y), o = kk_o<1>(e.y);
        f.x = affine(a.x, s, o);
        f.y = affine(a.y, s, o);
    } else if (e.x < (3u << 24)) {
        const uint2 a = kk_lds64(kk_parent<0, ROWB>(e.x, srd));
        const uint2 b = kk_lds64(kk_parent<1, ROWB>(e.x, srd));
        const uint32_t s0 = kk_s<0>(e.y), o0 = kk_o<1>(e.y), s1 = kk_s<2>(e.y), o1 = kk_o<3>(e.y);
        f.x = lop3<0xCA>(a.x, affine(b.x, s1, o1), affine(b.x, s0, o0));
        f.y = lop3<0xCA>(a.y, affine(b.y, s1, o1), affine(b.y, s0, o0));
    } else {
        const uint2 a = kk_lds64(kk_parent<0, ROWB>(e.x, srd));
        const uint2 b = kk_lds64(kk_parent<1, ROWB>(e.x, srd));
        const uint2 c = kk_lds64(kk_parent<2, ROWB>(e.x, srd));
        const uint32_t hi = kk_lds32(entry + 8);
        const uint32_t s0 = kk_s<0>(e.y), o0 = kk_o<1>(e.y), s1 = kk_s<2>(e.y), o1 = kk_o<3>(e.y);
        const uint32_t s2 = kk_s<0>(hi), o2 = kk_o<1>(hi), s3 = kk_s<2>(hi), o3 = kk_o<3>(hi);
        f.x = lop3<0xCA>(a.x, lop3<0xCA>(b.x, affine(c.x, s3, o3), affine(c.x, s2, o2)),
                         lop3<0xCA>(b.x, affine(c.x, s1, o1), affine(c.x, s0, o0)));
        f.y = lop3<0xCA>(a.y, lop3<0xCA>(b.y, affine(c.y, s3, o3), affine(c.y, s2, o2)),
                         lop3<0xCA>(b.y, affine(c.y, s1, o1), affine(c.y, s0, o0)));
    }
    return f;
}

template <int CNT>  // nodes per warp
__global__ void __launch_bounds__(kKkThreads, 2)
koki_plane_kernel(const SweepArgs args) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int N = args.n_nodes;
    constexpr int NP = CNT * kKkWarps;  // node rows incl. the padding; row NP = forced values
    const int64_t wpad = args.wpad;
    uint32_t* state = smem;                                                       // [NP + 1][2][64]
    KkEntry* table = reinterpret_cast<KkEntry*>(smem + (size_t)(NP + 1) * kKkRow);  // [NP + pad]
    uint32_t* votes = reinterpret_cast<uint32_t*>(table + NP + kKkTabPad);        // [3][2][64]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = warp * CNT;
    const int forced_node = (int)(blockIdx.x % (unsigned)N);
    const int64_t word = (int64_t)(blockIdx.x / (unsigned)N) * 32 + lane;
    const bool word_ok = word < wpad;
    uint32_t cellmask = word_ok ? tail_mask(word, args.n_cells) : 0u;

    // ---- tables (the forced node reads the constant row; a padding node is constant 0), constant row, vote words
    for (int n = tid; n < NP + kKkTabPad; n += kKkThreads) {
        KkEntry e = kk_make_entry(n < NP ? n : 0, 0, 0, 0, affine_table(0u, 0));
        if (n < N) {
            int arity;
            const NodeEntry ne = make_entry(args.net_parent, args.net_lut, n, &arity);
            e = kk_make_entry(ne.pa, ne.pb, ne.pc, arity, affine_table(ne.lut, arity));
            if (n == forced_node) e = kk_make_entry(NP, NP, NP, 1, affine_table(0xF0u, 1));  // f = A, A = the constant row
        }
        table[n] = e;
    }
    if (tid < 2 * kKkSlots) state[(size_t)NP * kKkRow + tid] = (tid & 1) ? 0xFFFFFFFFu : 0u;
    for (int i = tid; i < 3 * 2 * kKkSlots; i += kKkThreads) votes[i] = 0u;
    // cells whose unperturbed run needed the general scheme go straight to the cleanup list (see sweep_kernel)
    {
        const uint32_t hard = word_ok ? __ldg(args.np_hard + word) & cellmask : 0u;
        if (hard && warp == 0) {
            const unsigned int k = __popc(hard);
            const unsigned int at0 = atomicAdd(&args.hard_count[forced_node], k);
            int32_t* dst = args.hard_cells + (size_t)forced_node * args.hard_cap + at0;  // cap = n_cells: always room
            for (uint32_t m = hard; m; m &= m - 1) *dst++ = (int32_t)(word * 32 + (__ffs(m) - 1));
        }
        cellmask &= ~hard;
    }
    // shared-space addresses: (node n, copy k, slot pair of this lane) = s_lane + n * kKkRowBytes + k * kKkCopyBytes
    const uint32_t s_lane = shared_addr(state) + lane * 8;
    const uint32_t s_own = s_lane + n0 * kKkRowBytes;
    const uint32_t s_tab = shared_addr(table) + n0 * kKkEntryBytes;
    // ---- data -> copy 1 (both slots of a word start from the same state); padding rows: zero in both copies
    constexpr int KR = CNT < SCB_KK_REG_NODES ? CNT : SCB_KK_REG_NODES;
    uint2 cur[KR > 0 ? KR : 1];
#pragma unroll
    for (int i = 0; i < CNT; ++i) {
        const bool real = n0 + i < N;
        const uint32_t v = (real && word_ok) ? __ldg(args.planes + (size_t)(n0 + i) * wpad + word) : 0u;
        kk_sts64(s_own + i * kKkRowBytes + kKkCopyBytes, make_uint2(v, v));
        if (!real) kk_sts64(s_own + i * kKkRowBytes, make_uint2(0u, 0u));
        if (i < KR) cur[i < KR ? i : 0] = make_uint2(v, v);
    }
    __syncthreads();

    // ---- ring of 2
    uint32_t frozen0 = ~cellmask, frozen1 = ~cellmask;
    uint32_t lam0_0 = 0, lam1_0 = 0, lam0_1 = 0, lam1_1 = 0;  // lamB_h: bit B of lambda, slot h
    int vset = 0;
    unsigned int last_open = 0xFFFFFFFFu;
    int stall = 0;
    bool moved = false;
    const int t_fast = args.fast_steps > 0 ? min(args.steps, args.fast_steps) : min(args.steps, (int)args.counters[3]);

    // one step: reads copy RD (state_{t-1}), writes copy RD ^ 1 (which holds state_{t-2} of the own nodes).
    // Per node: the table entry was fetched one node ahead; parents, state_{t-2} and state_{t-1} of the node are
    // requested together, then the value, the two compares, the freeze merge and the store.
    auto step = [&](auto rd_tag, int t) -> bool {
        constexpr int RD = decltype(rd_tag)::value, WR = RD ^ 1;
        const uint32_t srd = s_lane + RD * kKkCopyBytes;
        uint32_t q1x = 0, q1y = 0, q2x = 0, q2y = 0;
        uint2 e_next = kk_lds64(s_tab);
#pragma unroll
        for (int i = 0; i < CNT; ++i) {
            const uint2 e = e_next;
            if (i + 1 < CNT) e_next = kk_lds64(s_tab + (i + 1) * kKkEntryBytes);
            const uint2 old2 = kk_lds64(s_own + i * kKkRowBytes + WR * kKkCopyBytes);
            // state_{t-1} of the first KR own nodes stays in registers (the ones the 64-register budget has room for)
            const uint2 c1 = i < KR ? cur[i < KR ? i : 0] : kk_lds64(s_own + i * kKkRowBytes + RD * kKkCopyBytes);
            const uint2 f = kk_eval(e, s_tab + i * kKkEntryBytes, srd);
            q1x |= f.x ^ c1.x;
            q1y |= f.y ^ c1.y;
            q2x |= f.x ^ old2.x;
            q2y |= f.y ^ old2.y;
            uint2 w;
            w.x = (f.x & ~frozen0) | (c1.x & frozen0);
            w.y = (f.y & ~frozen1) | (c1.y & frozen1);
            kk_sts64(s_own + i * kKkRowBytes + WR * kKkCopyBytes, w);
            if (i < KR) cur[i < KR ? i : 0] = w;
        }
        uint32_t* v = votes + vset * 2 * kKkSlots;
        q1x &= ~frozen0; q2x &= ~frozen0;  // only open lanes matter: a word whose lanes have all settled votes nothing
        q1y &= ~frozen1; q2y &= ~frozen1;
        if (q1x) atomicOr(v + 2 * lane, q1x);
        if (q1y) atomicOr(v + 2 * lane + 1, q1y);
        if (q2x) atomicOr(v + kKkSlots + 2 * lane, q2x);
        if (q2y) atomicOr(v + kKkSlots + 2 * lane + 1, q2y);
        __syncthreads();
        const uint2 n1 = *reinterpret_cast<const uint2*>(v + 2 * lane);
        const uint2 n2 = *reinterpret_cast<const uint2*>(v + kKkSlots + 2 * lane);
        // state_{t-d} exists for t >= d only (the raw data state is never compared)
        uint32_t open = ~frozen0;
        uint32_t e1 = t >= 1 ? ~n1.x & open : 0u;
        open &= ~e1;
        uint32_t e2 = t >= 2 ? ~n2.x & open : 0u;
        lam0_0 |= e1;
        lam1_0 |= e2;
        frozen0 |= e1 | e2;
        open = ~frozen1;
        e1 = t >= 1 ? ~n1.y & open : 0u;
        open &= ~e1;
        e2 = t >= 2 ? ~n2.y & open : 0u;
        lam0_1 |= e1;
        lam1_1 |= e2;
        frozen1 |= e1 | e2;
        // every warp holds the frozen masks of all 64 slots: the stop rule needs no barrier (ring2_registers)
        const unsigned int n_open = __reduce_add_sync(0xFFFFFFFFu, (unsigned int)(__popc(~frozen0) + __popc(~frozen1)));
        if (n_open <= (unsigned int)args.stop_open) return true;
        if (n_open < last_open) {
            moved = last_open != 0xFFFFFFFFu;
            last_open = n_open;
            stall = 0;
        } else if (moved && ++stall >= args.stop_stall) {
            return true;
        }
        const int vclear = vset == 0 ? 2 : vset - 1;  // (t + 2) % 3: every reader is past the barrier of step t + 1 by then
        if (tid < 2 * kKkSlots) votes[vclear * 2 * kKkSlots + tid] = 0u;
        vset = vset == 2 ? 0 : vset + 1;
        return false;
    };
    using FromOne = std::integral_constant<int, 1>;
    using FromZero = std::integral_constant<int, 0>;
    int xs = 1;  // copy that holds the final states
    for (int t = 0; t < t_fast; t += 2) {
        xs = 0;
        if (step(FromOne{}, t)) break;
        if (t + 1 >= t_fast) break;
        xs = 1;
        if (step(FromZero{}, t + 1)) break;
    }
    // (the closing barrier of the last step ordered every store of copy xs; nothing below writes the state)

    // ---- lanes still open go to the recheck list: a cell with both runs open as a pair, a cell with one run open
    // as that run alone (args.singles), its settled partner being scored here as if the open run will repeat
    const uint32_t open0 = ~frozen0, open1 = ~frozen1;
    const uint32_t pairs = args.singles ? (open0 & open1) : (open0 | open1);
    const uint32_t single0 = args.singles ? (open0 & ~open1) : 0u, single1 = args.singles ? (open1 & ~open0) : 0u;
    if (warp == 0) {
        if (pairs) {
            const unsigned int k = __popc(pairs);
            const unsigned int at0 = atomicAdd(&args.out_count[forced_node], k);
            int32_t* dst = args.out_cells + (size_t)forced_node * args.out_cap + at0;  // cap = n_cells: always room
            for (uint32_t m = pairs; m; m &= m - 1) *dst++ = (int32_t)(word * 32 + (__ffs(m) - 1));
        }
        if (single0 | single1) {
            const unsigned int k = __popc(single0) + __popc(single1);
            const unsigned int at0 = atomicAdd(&args.out_count1[forced_node], k);
            int32_t* dst = args.out_cells + (size_t)forced_node * args.out_cap + (args.out_cap - 1 - at0);  // from the back
            for (uint32_t m = single0; m; m &= m - 1) *dst-- = (int32_t)(word * 32 + (__ffs(m) - 1));
            for (uint32_t m = single1; m; m &= m - 1) *dst-- = (int32_t)(word * 32 + (__ffs(m) - 1)) | (1 << kWhichShift);
        }
    }
    const uint32_t cm0 = cellmask & ~pairs & ~single0, cm1 = cellmask & ~pairs & ~single1;  // lanes scored here, per slot

    // ---- scoring: a cell counts only if its KO run AND its KI run repeat (and the normal one)
    const uint32_t found_n = word_ok ? __ldg(args.np_found + word) : 0u;
    uint32_t lamn_ge2 = 0u;
    if (word_ok) {
#pragma unroll
        for (int b = 1; b < kLamPlanes; ++b) lamn_ge2 |= __ldg(args.np_lam + (size_t)b * wpad + word);
    }
    const uint32_t valid0 = cm0 & found_n, valid1 = cm1 & found_n;  // every lane left in cm0 / cm1 is frozen in that slot
    const uint32_t valid = valid0 | valid1;
    if (warp == 0) {
        const uint32_t bad = __popc(cm0 & cm1 & ~found_n);  // (a single run's cell is judged where that run settles)
        if (bad) atomicAdd(args.out_keyerror, (unsigned long long)bad);
    }
    unsigned long long score = 0;
    if (__any_sync(0xFFFFFFFFu, valid != 0)) {
        const uint32_t srd = s_lane + xs * kKkCopyBytes;
        const bool cyc = __any_sync(0xFFFFFFFFu, ((valid0 & (lam1_0 | lamn_ge2)) | (valid1 & (lam1_1 | lamn_ge2))) != 0);
        const uint32_t act2_0 = valid0 & lam1_0 & lamn_ge2, act2_1 = valid1 & lam1_1 & lamn_ge2;
        const bool any2 = __any_sync(0xFFFFFFFFu, (act2_0 | act2_1) != 0);
        uint32_t sc = 0, sc1 = 0;
#pragma unroll
        for (int i = 0; i < CNT; ++i) {
            if (n0 + i < N) {
                const size_t at = (size_t)(n0 + i) * wpad + word;
                const uint2 s = kk_lds64(s_own + i * kKkRowBytes + xs * kKkCopyBytes);
                const uint32_t ref0 = word_ok ? __ldg(args.np_states + at) : 0u;
                sc += __popc((s.x ^ ref0) & valid0) + __popc((s.y ^ ref0) & valid1);
                if (cyc) {
                    const uint2 f = kk_eval(kk_lds64(s_tab + i * kKkEntryBytes), s_tab + i * kKkEntryBytes, srd);
                    const uint32_t ref1 = word_ok ? __ldg(args.np_states + (size_t)N * wpad + at) : 0u;
                    sc1 += __popc((f.x ^ ref1) & valid0) + __popc((f.y ^ ref1) & valid1);
                    if (any2) {
                        const uint32_t ref2 = word_ok ? __ldg(args.np_states + (size_t)2 * N * wpad + at) : 0u;
                        sc1 += __popc((s.x ^ ref2) & act2_0) + __popc((s.y ^ ref2) & act2_1);
                    }
                }
            }
        }
        score = cyc ? (unsigned long long)sc + sc1 : 2ull * sc;  // fixed point vs fixed point: offset 1 repeats offset 0
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) score += __shfl_xor_sync(0xFFFFFFFFu, score, off);
    if (lane == 0 && score) atomicAdd(args.out_raw + forced_node, score);
}

// ---------------------------------------------------------------------------------------
// Cleanup pass (compacted cells, the general scheme), written like koki_plane_kernel.
//
// Same work and the same results as sweep_kernel<64, MODE_CLEANUP> with brent_registers / replay_registers:
// a batch is 1024 deferred cells of one forced node gathered into 32 virtual words, slot 2v = KO and 2v + 1 = KI
// of virtual word v; Brent checkpoints give lambda, two replays from the data (P runs lambda_cell steps ahead
// of Q under per-cell stall masks) give mu and leave Q frozen at state_mu; scoring steps the frozen states
// along the stored normal-run states.  What changed is the machinery: 768 threads (24 warps, six per
// scheduler instead of four; 80 registers), nodes dealt evenly (N / 24 or N / 24 + 1 per warp, the kernel is
// instantiated per N / 24 so only the last position carries a test), 64-bit KO/KI pairs, 16-byte packed table
// entries fetched one node ahead, the forced node as a constant row, parents by shared-space address -- and
// every warp holds the masks of all 64 slots (a lane is a virtual word, all warps read the same votes), so
// "is any lane left" is a warp vote instead of a barrier reduction: ONE barrier per step in every phase.
// The cells are gathered once per batch (a pristine copy of the data stays in state copy 2 while the Brent
// phase runs in copies 0 / 1), lambda / cyc / fixed-point planes are kept by warp 0 in shared memory during
// the Brent phase (registers there hold the previous state and the checkpoint of the own nodes).
// ---------------------------------------------------------------------------------------
constexpr int kKcThreads = 768, kKcWarps = kKcThreads / 32;
constexpr int kKcRowBytes = 4 * kKkSlots * 4, kKcCopyBytes = kKkSlots * 4;  // [node][4 copies][64 slots]
constexpr int kKcBatchCells = 32 * 32;
constexpr int kKcVoteWords = 3 * 4 * kKkSlots;  // three sets of up to four vote words per slot

static size_t koki_cleanup_smem_bytes(int n_nodes) {
    return (size_t)(n_nodes + 1) * kKcRowBytes + (size_t)(n_nodes + kKkTabPad) * sizeof(KkEntry) + (size_t)kKcVoteWords * 4 +
           (size_t)8 * kKkSlots * 4 + (size_t)2 * kKcBatchCells * 4 + (size_t)(1 + kLamPlanes) * 64 * 4 + 64 * 4 +
           (size_t)(n_nodes + 1) * 4;
}

template <int BASE, typename F>
__device__ __forceinline__ void kc_for_own(bool extra, F&& f) {
#pragma unroll
    for (int i = 0; i < BASE; ++i) f(i);
    if (extra) f(BASE);
}

// BASE = N / 24: every warp owns BASE nodes, the first N % 24 warps one more.
// RECHECK: the compacted recheck pass instead -- ring of 4 for all `steps` steps on the cells the KO/KI pass left
// open (a lane whose new state equals state_{t-d}, smallest d <= 4, has its first repeat at (t-d, t) and is frozen
// there; lanes still open -- longer cycles, no repeat -- go to the cleanup list): state_{t-1} and state_{t-2} of the
// own nodes in registers, state_{t-3} and state_{t-4} read from the copies they live in (the oldest one is the copy
// being overwritten), four vote words per slot.
// SINGLES: the batch is 2 048 single runs (SweepArgs::singles) instead of 1 024 KO/KI pairs: 64 virtual words, the
// two slots of a lane are two unrelated words, the constant row holds per slot which of its runs are knock-ins, a
// run is scored as if its partner (scored where it settled) repeats, and a run without a repeat sends that partner
// to the retract list.  args.negate (with RECHECK): the retract pass -- scores are subtracted, nothing else is written.
template <int BASE, bool RECHECK, bool SINGLES>
__global__ void __launch_bounds__(kKcThreads, 1)
koki_compact_kernel(const SweepArgs args) {
    constexpr int NV = SINGLES ? 64 : 32;     // virtual words per batch
    constexpr int kBatch = NV * 32;           // list entries per batch
    constexpr int MODE = RECHECK ? MODE_RECHECK : MODE_CLEANUP;  // (trace macros)
    (void)MODE;
    extern __shared__ __align__(16) uint32_t smem[];
    const int N = args.n_nodes;
    const int steps = args.steps;
    uint32_t* state = smem;                                                         // [N + 1][4][64], row N = forced values
    KkEntry* table = reinterpret_cast<KkEntry*>(smem + (size_t)(N + 1) * 4 * kKkSlots);  // [N + pad]
    uint32_t* votes = reinterpret_cast<uint32_t*>(table + N + kKkTabPad);           // [3][2][64], recheck [3][4][64]
    uint32_t* res = votes + kKcVoteWords;                                       // [8][64]: lambda planes, cyc, fixed point
    int* cellidx = reinterpret_cast<int*>(res + 8 * kKkSlots);                      // [NV][32] list entries (-1: none)
    uint32_t* nrm = reinterpret_cast<uint32_t*>(cellidx + 2 * kKcBatchCells);       // [1 + kLamPlanes][NV]
    uint32_t* kim = nrm + (1 + kLamPlanes) * 64;                                    // [64] singles: knock-in lanes of every slot
    int* pre = reinterpret_cast<int*>(kim + 64);                                    // [N + 1]
    __shared__ int s_node;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef SCB_TRACE
    long long trace_t_ = clock64();
#endif
    const int rem = N - BASE * kKcWarps;  // 0 <= rem < 24 by the choice of BASE
    const bool extra = warp < rem;
    const int n0 = warp * BASE + min(warp, rem);

    for (int n = tid; n < N + kKkTabPad; n += kKcThreads) {
        KkEntry e = kk_make_entry(0, 0, 0, 0, affine_table(0u, 0));
        if (n < N) {
            int arity;
            const NodeEntry ne = make_entry(args.net_parent, args.net_lut, n, &arity);
            e = kk_make_entry(ne.pa, ne.pb, ne.pc, arity, affine_table(ne.lut, arity));
        }
        table[n] = e;
    }
    if (tid < 4 * kKkSlots) state[(size_t)N * 4 * kKkSlots + tid] = (tid & 1) ? 0xFFFFFFFFu : 0u;
    if (tid == 0) {
        int acc = 0;
        for (int k = 0; k < N; ++k) {
            pre[k] = acc;
            int64_t c = SINGLES ? args.in_count1[k] : args.in_count[k];
            if (c > args.in_cap) c = args.in_cap;
            acc += (int)((c + kBatch - 1) / kBatch);
        }
        pre[N] = acc;
    }
    __syncthreads();
    const int total_batches = pre[N];
    const uint32_t s_lane = shared_addr(state) + lane * 8;
    const uint32_t s_own = s_lane + n0 * kKcRowBytes;
    const uint32_t s_tab = shared_addr(table) + n0 * kKkEntryBytes;
    int prev_forced = -1;

    // copy k of every node, both slots, from the cell-major records of the batch's cells (`src_t`), through the
    // staging copy kstage: one warp per virtual word, lane = cell, 128-bit loads fetch its record, every record
    // word goes through a 32 x 32 warp transpose (lane b gets node 32j + b) and is parked at a pair position
    // rotated by the node index (the lanes of a warp hold different nodes, whose rows start on the same bank);
    // after a barrier every thread moves the pairs of its own nodes into place -- in copy k and, if k2 >= 0, in k2.
    auto gather = [&](const uint32_t* __restrict__ src_t, int k, int k2, int kstage) {
        const int nwp = args.nwp;
        for (int v0 = warp; v0 < NV; v0 += 2 * kKcWarps) {
            const int v1 = v0 + kKcWarps;
            const int e0 = cellidx[v0 * 32 + lane];
            const int e1 = v1 < NV ? cellidx[v1 * 32 + lane] : -1;
            const int c0 = e0 >= 0 ? (e0 & kCellMask) : -1, c1 = e1 >= 0 ? (e1 & kCellMask) : -1;
            for (int j4 = 0; j4 < nwp; j4 += 4) {
                uint4 x0 = make_uint4(0u, 0u, 0u, 0u), x1 = x0;
                if (c0 >= 0) x0 = __ldg(reinterpret_cast<const uint4*>(src_t + (size_t)c0 * nwp + j4));
                if (c1 >= 0) x1 = __ldg(reinterpret_cast<const uint4*>(src_t + (size_t)c1 * nwp + j4));
                const uint32_t xw[2][4] = {{x0.x, x0.y, x0.z, x0.w}, {x1.x, x1.y, x1.z, x1.w}};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int v = h ? v1 : v0;
                    if (v >= NV) break;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int nb = (j4 + q) * 32;
                        if (nb < N) {
                            const uint32_t mine = warp_transpose32(xw[h][q], lane);
                            const int n = nb + lane;
                            if (n < N) {
                                if (SINGLES) state[((size_t)n * 4 + kstage) * kKkSlots + ((v + n) & 63)] = mine;
                                else *reinterpret_cast<uint2*>(state + ((size_t)n * 4 + kstage) * kKkSlots + 2 * ((v + n) & 31)) = make_uint2(mine, mine);
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        kc_for_own<BASE>(extra, [&](int i) {
            uint2 x;
            const uint32_t row = shared_addr(state) + (uint32_t)(n0 + i) * kKcRowBytes + kstage * kKcCopyBytes;
            if (SINGLES) {
                x.x = kk_lds32(row + 4 * ((2 * lane + n0 + i) & 63));
                x.y = kk_lds32(row + 4 * ((2 * lane + 1 + n0 + i) & 63));
            } else {
                x = kk_lds64(row + 8 * ((lane + n0 + i) & 31));
            }
            kk_sts64(s_own + i * kKcRowBytes + k * kKcCopyBytes, x);
            if (k2 >= 0) kk_sts64(s_own + i * kKcRowBytes + k2 * kKcCopyBytes, x);
        });
    };

    for (int batch = blockIdx.x; batch < total_batches; batch += gridDim.x) {
        // ------------------------------------------------------------ the batch: node, cells, table patch
        if (tid == 0) {
            int lo = 0, hi = N;  // largest k with pre[k] <= batch
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (pre[mid] <= batch) lo = mid; else hi = mid;
            }
            s_node = lo;
        }
        __syncthreads();
        const int forced_node = s_node;
        int64_t n_in = SINGLES ? args.in_count1[forced_node] : args.in_count[forced_node];
        if (n_in > args.in_cap) n_in = args.in_cap;
        const int64_t batch_first = (int64_t)(batch - pre[forced_node]) * kBatch;
        for (int i = tid; i < kBatch; i += kKcThreads) {
            const int64_t j = batch_first + i;  // (single runs are listed from the back of the node's array)
            cellidx[i] = j < n_in ? args.in_cells[(size_t)forced_node * args.in_cap + (SINGLES ? args.in_cap - 1 - j : j)] : -1;
        }
        if (forced_node != prev_forced && tid == 0) {
            if (prev_forced >= 0) {
                int arity;
                const NodeEntry ne = make_entry(args.net_parent, args.net_lut, prev_forced, &arity);
                table[prev_forced] = kk_make_entry(ne.pa, ne.pb, ne.pc, arity, affine_table(ne.lut, arity));
            }
            table[forced_node] = kk_make_entry(N, N, N, 1, affine_table(0xF0u, 1));  // f = A, A = the constant row
        }
        prev_forced = forced_node;
        for (int i = tid; i < kKcVoteWords + 8 * kKkSlots; i += kKcThreads) votes[i] = 0u;  // votes and res
        // lanes in use, per slot: a pair batch has the same cells in both slots of a lane
        auto lanes_of = [&](int v) -> uint32_t {
            const int64_t left = n_in - (batch_first + (int64_t)v * 32);
            return left >= 32 ? 0xFFFFFFFFu : (left > 0 ? (1u << (int)left) - 1u : 0u);
        };
        uint32_t cm0 = lanes_of(SINGLES ? 2 * lane : lane), cm1 = SINGLES ? lanes_of(2 * lane + 1) : cm0;
        __syncthreads();
        uint32_t kim0 = 0u, kim1 = 0u;  // singles: the knock-in runs of my two slots
        if constexpr (SINGLES) {
            for (int v = warp; v < NV; v += kKcWarps) {
                const int e = cellidx[v * 32 + lane];
                const uint32_t bits = __ballot_sync(0xFFFFFFFFu, e >= 0 && ((e >> kWhichShift) & 1));
                if (lane == 0) kim[v] = bits;
            }
            __syncthreads();
            if (tid < 4 * kKkSlots) state[(size_t)N * 4 * kKkSlots + tid] = kim[tid & 63];  // the constant row, all four copies
            kim0 = kim[2 * lane];
            kim1 = kim[2 * lane + 1];
            __syncthreads();
        }
        SCB_SWEEP_PHASE(0);

        uint32_t lam0[kLamPlanes], lam1[kLamPlanes];  // cycle length of every lane, bit-sliced
        uint32_t found0, found1;
        int xs, xn;  // copy frozen at state_mu for found lanes / its scratch partner
        if constexpr (RECHECK) {
            // ------------------------------------------------------------ ring of 4 (data into copy 3)
            gather(args.planes_t, 3, -1, 0);
            uint2 ra[BASE + 1], rb[BASE + 1];  // state_{t-1} / state_{t-2} of the own nodes, roles alternate
            kc_for_own<BASE>(extra, [&](int i) {
                ra[i] = kk_lds64(s_own + i * kKcRowBytes + 3 * kKcCopyBytes);
                rb[i] = make_uint2(0u, 0u);
            });
            __syncthreads();
            uint32_t frozen0 = ~cm0, frozen1 = ~cm1;
#pragma unroll
            for (int b = 0; b < kLamPlanes; ++b) lam0[b] = lam1[b] = 0u;
            unsigned int last_open = 0xFFFFFFFFu;
            int stall = 0, vset = 0;
            bool moved = false;
            // step t writes copy KD = t & 3 (it holds state_{t-4}) and reads copy KD + 3 (state_{t-1}); cur / prv are
            // state_{t-1} / state_{t-2} of the own nodes, prv receives state_t.  Returns true when the ring is done.
            auto ring_step = [&](auto kd_tag, uint2 (&cur)[BASE + 1], uint2 (&prv)[BASE + 1], int t) -> bool {
                constexpr int KD = decltype(kd_tag)::value, KS = (KD + 3) & 3, K3 = (KD + 1) & 3;
                const uint32_t srd = s_lane + KS * kKcCopyBytes;
                uint32_t q1x = 0, q1y = 0, q2x = 0, q2y = 0, q3x = 0, q3y = 0, q4x = 0, q4y = 0;
                uint2 e_next = kk_lds64(s_tab);
                kc_for_own<BASE>(extra, [&](int i) {
                    const uint2 e = e_next;
                    e_next = kk_lds64(s_tab + (i + 1) * kKkEntryBytes);  // (the table is padded)
                    const uint2 o3 = kk_lds64(s_own + i * kKcRowBytes + K3 * kKcCopyBytes);
                    const uint2 o4 = kk_lds64(s_own + i * kKcRowBytes + KD * kKcCopyBytes);
                    const uint2 f = kk_eval<kKcRowBytes>(e, s_tab + i * kKkEntryBytes, srd);
                    q1x |= f.x ^ cur[i].x;
                    q1y |= f.y ^ cur[i].y;
                    q2x |= f.x ^ prv[i].x;
                    q2y |= f.y ^ prv[i].y;
                    q3x |= f.x ^ o3.x;
                    q3y |= f.y ^ o3.y;
                    q4x |= f.x ^ o4.x;
                    q4y |= f.y ^ o4.y;
                    uint2 w;
                    w.x = (f.x & ~frozen0) | (cur[i].x & frozen0);
                    w.y = (f.y & ~frozen1) | (cur[i].y & frozen1);
                    kk_sts64(s_own + i * kKcRowBytes + KD * kKcCopyBytes, w);
                    prv[i] = w;
                });
                uint32_t* v = votes + vset * 4 * kKkSlots;  // [d - 1][slot]; only open lanes matter
                q1x &= ~frozen0; q2x &= ~frozen0; q3x &= ~frozen0; q4x &= ~frozen0;
                q1y &= ~frozen1; q2y &= ~frozen1; q3y &= ~frozen1; q4y &= ~frozen1;
                if (q1x) atomicOr(v + 2 * lane, q1x);
                if (q1y) atomicOr(v + 2 * lane + 1, q1y);
                if (q2x) atomicOr(v + kKkSlots + 2 * lane, q2x);
                if (q2y) atomicOr(v + kKkSlots + 2 * lane + 1, q2y);
                if (q3x) atomicOr(v + 2 * kKkSlots + 2 * lane, q3x);
                if (q3y) atomicOr(v + 2 * kKkSlots + 2 * lane + 1, q3y);
                if (q4x) atomicOr(v + 3 * kKkSlots + 2 * lane, q4x);
                if (q4y) atomicOr(v + 3 * kKkSlots + 2 * lane + 1, q4y);
                __syncthreads();
                const uint2 n1 = *reinterpret_cast<const uint2*>(v + 2 * lane);
                const uint2 n2 = *reinterpret_cast<const uint2*>(v + kKkSlots + 2 * lane);
                const uint2 n3 = *reinterpret_cast<const uint2*>(v + 2 * kKkSlots + 2 * lane);
                const uint2 n4 = *reinterpret_cast<const uint2*>(v + 3 * kKkSlots + 2 * lane);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t& frozen = h ? frozen1 : frozen0;
                    uint32_t(&lam)[kLamPlanes] = h ? lam1 : lam0;
                    // state_{t-d} exists for t >= d only (the raw data state is never compared)
                    uint32_t open = ~frozen;
                    const uint32_t e1 = t >= 1 ? ~(h ? n1.y : n1.x) & open : 0u;
                    open &= ~e1;
                    const uint32_t e2 = t >= 2 ? ~(h ? n2.y : n2.x) & open : 0u;
                    open &= ~e2;
                    const uint32_t e3 = t >= 3 ? ~(h ? n3.y : n3.x) & open : 0u;
                    open &= ~e3;
                    const uint32_t e4 = t >= 4 ? ~(h ? n4.y : n4.x) & open : 0u;
                    lam[0] |= e1 | e3;
                    lam[1] |= e2 | e3;
                    lam[2] |= e4;
                    frozen |= e1 | e2 | e3 | e4;
                }
                // set t % 3 is written in step t and read after its barrier; set (t + 2) % 3 is cleared here, after
                // the barrier of step t (its readers finished before it) and written in step t + 2
                const int vclear = vset == 0 ? 2 : vset - 1;
                if (tid < 4 * kKkSlots) votes[vclear * 4 * kKkSlots + tid] = 0u;
                vset = vset == 2 ? 0 : vset + 1;
                // stop when nothing is open, or once lanes had started to settle and none did for a while (what is left
                // are cycles the ring cannot see)
                const unsigned int n_open = __reduce_add_sync(0xFFFFFFFFu, (unsigned int)(__popc(~frozen0) + __popc(~frozen1)));
                if (n_open == 0) return true;
                if (n_open < last_open) {
                    moved = last_open != 0xFFFFFFFFu;
                    last_open = n_open;
                    stall = 0;
                } else if (moved && ++stall >= args.ring_stall && !args.negate) {
                    return true;  // (the retract pass runs on: every one of its runs settles, it did before)
                }
                return false;
            };
            using K0 = std::integral_constant<int, 0>;
            using K1 = std::integral_constant<int, 1>;
            using K2 = std::integral_constant<int, 2>;
            using K3 = std::integral_constant<int, 3>;
            xs = 3;
            for (int t = 0; t < steps; t += 4) {
                xs = 0;
                if (ring_step(K0{}, ra, rb, t)) break;
                if (t + 1 >= steps) break;
                xs = 1;
                if (ring_step(K1{}, rb, ra, t + 1)) break;
                if (t + 2 >= steps) break;
                xs = 2;
                if (ring_step(K2{}, ra, rb, t + 2)) break;
                if (t + 3 >= steps) break;
                xs = 3;
                if (ring_step(K3{}, rb, ra, t + 3)) break;
            }
            xn = (xs + 1) & 3;
            SCB_SWEEP_PHASE(2);
            if constexpr (SINGLES) {
                // runs still open go on alone (to the back of the cleanup arrays)
                const uint32_t open0 = ~frozen0, open1 = ~frozen1;
                if ((open0 | open1) && warp == 0 && !args.negate) {
                    const unsigned int k = __popc(open0) + __popc(open1);
                    const unsigned int at0 = atomicAdd(&args.out_count1[forced_node], k);
                    int32_t* dst = args.out_cells + (size_t)forced_node * args.out_cap + (args.out_cap - 1 - at0);
                    for (uint32_t m = open0; m; m &= m - 1) *dst-- = cellidx[(2 * lane) * 32 + (__ffs(m) - 1)];
                    for (uint32_t m = open1; m; m &= m - 1) *dst-- = cellidx[(2 * lane + 1) * 32 + (__ffs(m) - 1)];
                }
                cm0 &= ~open0;
                cm1 &= ~open1;
            } else {
                // lanes still open: KO and KI of a cell travel together to the cleanup list
                const uint32_t both = ~frozen0 | ~frozen1;
                if (both && warp == 0) {
                    const unsigned int k = __popc(both);
                    const unsigned int at0 = atomicAdd(&args.out_count[forced_node], k);
                    int32_t* dst = args.out_cells + (size_t)forced_node * args.out_cap + at0;  // cap = n_cells: always room
                    for (uint32_t m = both; m; m &= m - 1) *dst++ = cellidx[lane * 32 + (__ffs(m) - 1)];
                }
                cm0 &= ~both;
                cm1 &= ~both;
            }
            found0 = frozen0;
            found1 = frozen1;
        }
        if constexpr (!RECHECK) {
            // ------------------------------------------------------------ Brent phase: copies 0 / 1, data kept in copy 2
            gather(args.planes_t, 0, 2, 3);
            uint2 cur[BASE + 1], chk[BASE + 1];
            kc_for_own<BASE>(extra, [&](int i) {
                cur[i] = kk_lds64(s_own + i * kKcRowBytes);
                chk[i] = make_uint2(0u, 0u);
            });
            __syncthreads();
            uint32_t resolved0 = ~cm0, resolved1 = ~cm1;
            int chk_time = -1;
            int vset = 0;
            // Checkpoints at t = 2^k - 1 < steps - 1, a last one at steps - 1 (see sweep_kernel), and a PERMANENT one at
            // p = (steps - 1) / 2 kept in state copy 3 (the staging copy, free after the gather): a cell with a repeat
            // inside the window (mu + lambda <= steps - 1) is on its cycle at p if mu <= p, and back at that state by
            // p + lambda <= p + steps - 1; otherwise lambda < steps - 1 - p and the checkpoint at steps - 1 sees it by
            // 2 (steps - 1) - p.  75 steps settle every cell for steps = 50, where the last checkpoint alone needs 99.
            // (Any checkpoint that matches gives lambda exactly: it is compared every step from the one after it.)
            const int p1 = args.brent_mid ? (steps - 1) / 2 : (1 << 20);
            const int t1_max = args.brent_mid ? (steps - 1) + max(p1, steps - 1 - p1) + 1 : 2 * steps - 1;
            auto brent_step = [&](auto rd_tag, int t) -> bool {
                constexpr int RD = decltype(rd_tag)::value, WR = RD ^ 1;
                const uint32_t srd = s_lane + RD * kKcCopyBytes;
                uint32_t qpx = 0, qpy = 0, qcx = 0, qcy = 0, qfx = 0, qfy = 0;
                const bool p1_on = t > p1;
                uint2 e_next = kk_lds64(s_tab);
                kc_for_own<BASE>(extra, [&](int i) {
                    const uint2 e = e_next;
                    e_next = kk_lds64(s_tab + (i + 1) * kKkEntryBytes);  // (the table is padded)
                    const uint2 f = kk_eval<kKcRowBytes>(e, s_tab + i * kKkEntryBytes, srd);
                    qpx |= f.x ^ cur[i].x;
                    qpy |= f.y ^ cur[i].y;
                    qcx |= f.x ^ chk[i].x;
                    qcy |= f.y ^ chk[i].y;
                    if (p1_on) {
                        const uint2 fix = kk_lds64(s_own + i * kKcRowBytes + 3 * kKcCopyBytes);
                        qfx |= f.x ^ fix.x;
                        qfy |= f.y ^ fix.y;
                    }
                    kk_sts64(s_own + i * kKcRowBytes + WR * kKcCopyBytes, f);
                    if (t == p1) kk_sts64(s_own + i * kKcRowBytes + 3 * kKcCopyBytes, f);
                    cur[i] = f;
                });
                uint32_t* v = votes + vset * 4 * kKkSlots;
                if (qpx) atomicOr(v + 2 * lane, qpx);
                if (qpy) atomicOr(v + 2 * lane + 1, qpy);
                if (chk_time >= 0) {
                    if (qcx) atomicOr(v + kKkSlots + 2 * lane, qcx);
                    if (qcy) atomicOr(v + kKkSlots + 2 * lane + 1, qcy);
                }
                if (p1_on) {
                    if (qfx) atomicOr(v + 2 * kKkSlots + 2 * lane, qfx);
                    if (qfy) atomicOr(v + 2 * kKkSlots + 2 * lane + 1, qfy);
                }
                __syncthreads();
                const uint2 np = *reinterpret_cast<const uint2*>(v + 2 * lane);
                const uint2 nc = *reinterpret_cast<const uint2*>(v + kKkSlots + 2 * lane);
                const uint2 nf = *reinterpret_cast<const uint2*>(v + 2 * kKkSlots + 2 * lane);
                const int l = t - chk_time;
    #pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t& resolved = h ? resolved1 : resolved0;
                    const uint32_t neq_prev = h ? np.y : np.x, neq_chk = h ? nc.y : nc.x;
                    uint32_t* r = res + 2 * lane + h;  // plane b of this slot: r[b * 64]
                    if (t >= 1) {
                        const uint32_t newly = ~neq_prev & ~resolved;
                        resolved |= newly;
                        if (warp == 0 && newly) {
                            r[0] |= newly;  // lambda = 1
                            if (t <= steps - 1) r[7 * kKkSlots] |= newly;  // reached a fixed point inside the window
                        }
                    }
                    if (chk_time >= 0) {
                        const uint32_t newly = ~neq_chk & ~resolved;
                        resolved |= newly;  // l >= steps: such a cell can never repeat inside the window
                        if (warp == 0 && newly && l <= steps - 1) {
    #pragma unroll
                            for (int b = 0; b < kLamPlanes; ++b)
                                if ((l >> b) & 1) r[b * kKkSlots] |= newly;
                            r[6 * kKkSlots] |= newly;  // a period found through a checkpoint
                        }
                    }
                    if (p1_on) {
                        const uint32_t newly = ~(h ? nf.y : nf.x) & ~resolved;
                        const int l1 = t - p1;
                        resolved |= newly;
                        if (warp == 0 && newly && l1 <= steps - 1) {
#pragma unroll
                            for (int b = 0; b < kLamPlanes; ++b)
                                if ((l1 >> b) & 1) r[b * kKkSlots] |= newly;
                            r[6 * kKkSlots] |= newly;
                        }
                    }
                }
                if ((((t + 1) & t) == 0 && t < steps - 1) || t == steps - 1) {
                    kc_for_own<BASE>(extra, [&](int i) { chk[i] = cur[i]; });
                    chk_time = t;
                }
                const int vclear = vset == 0 ? 2 : vset - 1;
                if (tid < 4 * kKkSlots) votes[vclear * 4 * kKkSlots + tid] = 0u;
                vset = vset == 2 ? 0 : vset + 1;
                return __all_sync(0xFFFFFFFFu, (resolved0 & resolved1) == 0xFFFFFFFFu);
            };
            using C0 = std::integral_constant<int, 0>;
            using C1 = std::integral_constant<int, 1>;
            int last = 0;  // copy that holds the last state
            for (int t = 0; t < t1_max; t += 2) {
                last = 1;
                if (brent_step(C0{}, t)) break;
                if (t + 1 >= t1_max) break;
                last = 0;
                if (brent_step(C1{}, t + 1)) break;
            }
            __syncthreads();  // res is complete (warp 0), every vote word has been read
    #pragma unroll
            for (int b = 0; b < kLamPlanes; ++b) {
                const uint2 x = *reinterpret_cast<const uint2*>(res + b * kKkSlots + 2 * lane);
                lam0[b] = x.x;
                lam1[b] = x.y;
            }
            const uint2 cyc = *reinterpret_cast<const uint2*>(res + 6 * kKkSlots + 2 * lane);
            const uint2 fpt = *reinterpret_cast<const uint2*>(res + 7 * kKkSlots + 2 * lane);
            if (!__any_sync(0xFFFFFFFFu, (cyc.x | cyc.y) != 0)) {
                found0 = fpt.x;  // every settled lane sits on its fixed point already
                found1 = fpt.y;
                xs = last;
                xn = last ^ 1;
            } else {
                // ------------------------------------------------------------ replay: P in copies 0 / 1, Q in copies 2 / 3
                for (int i = tid; i < 3 * 2 * kKkSlots; i += kKcThreads) votes[i] = 0u;
                uint2 pq[BASE + 1];  // own nodes: P during the advance, Q afterwards
                kc_for_own<BASE>(extra, [&](int i) {
                    pq[i] = kk_lds64(s_own + i * kKcRowBytes + 2 * kKcCopyBytes);
                    kk_sts64(s_own + i * kKcRowBytes, pq[i]);
                });
                __syncthreads();
                const uint32_t has0 = planes_ge(lam0, 1) & cm0, has1 = planes_ge(lam1, 1) & cm1;
                int pc = 0;
                for (int s = 0; s < steps; ++s) {
                    const uint32_t adv0 = planes_gt(lam0, s) & has0, adv1 = planes_gt(lam1, s) & has1;  // lanes that still owe P a step
                    if (!__any_sync(0xFFFFFFFFu, (adv0 | adv1) != 0)) break;
                    const uint32_t srd = s_lane + pc * kKcCopyBytes;
                    uint2 e_next = kk_lds64(s_tab);
                    kc_for_own<BASE>(extra, [&](int i) {
                        const uint2 e = e_next;
                        e_next = kk_lds64(s_tab + (i + 1) * kKkEntryBytes);
                        const uint2 f = kk_eval<kKcRowBytes>(e, s_tab + i * kKkEntryBytes, srd);
                        pq[i].x = (f.x & adv0) | (pq[i].x & ~adv0);
                        pq[i].y = (f.y & adv1) | (pq[i].y & ~adv1);
                        kk_sts64(s_own + i * kKcRowBytes + (pc ^ 1) * kKcCopyBytes, pq[i]);
                    });
                    __syncthreads();
                    pc ^= 1;
                }
                kc_for_own<BASE>(extra, [&](int i) { pq[i] = kk_lds64(s_own + i * kKcRowBytes + 2 * kKcCopyBytes); });
                int qc = 2;
                uint32_t done0 = ~has0, done1 = ~has1;
                found0 = found1 = 0u;
                vset = 0;
                for (int j = 0; j <= steps - 2; ++j) {
                    uint32_t neq0 = 0, neq1 = 0;
                    const uint32_t prd = s_lane + pc * kKcCopyBytes, qrd = s_lane + qc * kKcCopyBytes;
                    uint2 e_next = kk_lds64(s_tab);
                    kc_for_own<BASE>(extra, [&](int i) {
                        const uint2 e = e_next;
                        e_next = kk_lds64(s_tab + (i + 1) * kKkEntryBytes);
                        const uint2 fp = kk_eval<kKcRowBytes>(e, s_tab + i * kKkEntryBytes, prd);
                        const uint2 fq = kk_eval<kKcRowBytes>(e, s_tab + i * kKkEntryBytes, qrd);
                        neq0 |= fp.x ^ fq.x;
                        neq1 |= fp.y ^ fq.y;
                        pq[i].x = (pq[i].x & done0) | (fq.x & ~done0);
                        pq[i].y = (pq[i].y & done1) | (fq.y & ~done1);
                        kk_sts64(s_own + i * kKcRowBytes + (pc ^ 1) * kKcCopyBytes, fp);
                        kk_sts64(s_own + i * kKcRowBytes + (qc ^ 1) * kKcCopyBytes, pq[i]);
                    });
                    uint32_t* v = votes + vset * 2 * kKkSlots;
                    if (neq0) atomicOr(v + 2 * lane, neq0);
                    if (neq1) atomicOr(v + 2 * lane + 1, neq1);
                    __syncthreads();
                    const uint2 nq = *reinterpret_cast<const uint2*>(v + 2 * lane);
                    // first repeat is (j, j + lambda); it only exists if j + lambda <= steps - 1
                    uint32_t newly = ~nq.x & ~done0;
                    found0 |= newly & ~planes_gt(lam0, steps - 1 - j);
                    done0 |= newly;
                    newly = ~nq.y & ~done1;
                    found1 |= newly & ~planes_gt(lam1, steps - 1 - j);
                    done1 |= newly;
                    const int vclear = vset == 0 ? 2 : vset - 1;
                    if (tid < kKkSlots) votes[vclear * 2 * kKkSlots + tid] = 0u;
                    vset = vset == 2 ? 0 : vset + 1;
                    pc ^= 1;
                    qc ^= 1;
                    if (__all_sync(0xFFFFFFFFu, (done0 & done1) == 0xFFFFFFFFu)) break;
                }
                xs = qc;
                xn = qc ^ 1;
            }
        }
        found0 &= cm0;
        found1 &= cm1;
        SCB_SWEEP_PHASE(4);

        // ------------------------------------------------------------ outputs
        // normal-run info of the batch's cells: one 4-byte record per cell (bits 0..5 lambda, bit 6 found)
        for (int v = warp; v < NV; v += kKcWarps) {
            const int e = cellidx[v * 32 + lane];
            const uint32_t rec = e >= 0 ? __ldg(args.np_info_t + (e & kCellMask)) : 0u;
#pragma unroll
            for (int b = 0; b <= kLamPlanes; ++b) {
                const uint32_t bits = __ballot_sync(0xFFFFFFFFu, (rec >> b) & 1u);
                if (lane == 0) nrm[(b == kLamPlanes ? 0 : 1 + b) * NV + v] = bits;
            }
        }
        __syncthreads();
        // (pairs: both slots of a lane hold the same cells; singles: two unrelated words)
        const int nv0 = SINGLES ? 2 * lane : lane, nv1 = SINGLES ? 2 * lane + 1 : lane;
        const uint32_t found_n0 = nrm[nv0], found_n1 = nrm[nv1];
        // "the normal run's cycle is at least t long", per slot, from the planes in shared memory
        auto lamn_ge = [&](int nv, int t) -> uint32_t {
            uint32_t l[kLamPlanes];
#pragma unroll
            for (int b = 0; b < kLamPlanes; ++b) l[b] = nrm[(1 + b) * NV + nv];
            return planes_ge(l, t);
        };
        // a cell counts only if its KO run AND its KI run repeat (and the normal one); a single run's partner has
        // repeated already, where it settled
        const uint32_t valid0 = SINGLES ? (found0 & found_n0) : (found0 & found1 & found_n0);
        const uint32_t valid1 = SINGLES ? (found1 & found_n1) : valid0;
        if (warp == 0 && !args.negate) {
            const uint32_t miss0 = ~found0 & cm0, miss1 = ~found1 & cm1;
            // (pairs: slot 0 is the knock-out, slot 1 the knock-in; singles: kim0 / kim1 say which run is which)
            const uint32_t m_ko = SINGLES ? __popc(miss0 & ~kim0) + __popc(miss1 & ~kim1) : __popc(miss0);
            const uint32_t m_ki = SINGLES ? __popc(miss0 & kim0) + __popc(miss1 & kim1) : __popc(miss1);
            if (m_ko) atomicAdd(args.out_missing + 1 + 2 * forced_node, (unsigned long long)m_ko);
            if (m_ki) atomicAdd(args.out_missing + 2 + 2 * forced_node, (unsigned long long)m_ki);
            const uint32_t bad = SINGLES ? __popc(found0 & ~found_n0) + __popc(found1 & ~found_n1) : __popc(found0 & found1 & ~found_n0);
            if (bad) atomicAdd(args.out_keyerror, (unsigned long long)bad);
            if constexpr (SINGLES) {
                // a run without a repeat: its partner was scored as if there were one -- the partner goes to the retract list
                const uint32_t r0 = miss0 & found_n0, r1 = miss1 & found_n1;
                if (r0 | r1) {
                    const unsigned int k = __popc(r0) + __popc(r1);
                    const unsigned int at0 = atomicAdd(&args.rt_count[forced_node], k);
                    int32_t* dst = args.rt_cells + (size_t)forced_node * args.rt_cap + (args.rt_cap - 1 - at0);
                    for (uint32_t m = r0; m; m &= m - 1) *dst-- = cellidx[(2 * lane) * 32 + (__ffs(m) - 1)] ^ (1 << kWhichShift);
                    for (uint32_t m = r1; m; m &= m - 1) *dst-- = cellidx[(2 * lane + 1) * 32 + (__ffs(m) - 1)] ^ (1 << kWhichShift);
                }
            }
        }
        unsigned long long score = 0;
        if (__any_sync(0xFFFFFFFFu, (valid0 | valid1) != 0)) {
            // offset 0: the frozen states against the normal run's state_mu, gathered into the pair of copies
            // that xs / xn do not use
            gather(args.np_states_t, xs ^ 2, -1, xn ^ 2);
            const uint32_t ge2_0 = planes_ge(lam0, 2), ge2_1 = planes_ge(lam1, 2);
            const uint32_t gen2_0 = lamn_ge(nv0, 2), gen2_1 = SINGLES ? lamn_ge(nv1, 2) : gen2_0;
            const uint32_t both2_0 = valid0 & ge2_0 & gen2_0, both2_1 = valid1 & ge2_1 & gen2_1;
            uint32_t sc = 0, sc_both2 = 0;
            kc_for_own<BASE>(extra, [&](int i) {
                const uint2 ref = kk_lds64(s_own + i * kKcRowBytes + (xs ^ 2) * kKcCopyBytes);
                const uint2 s = kk_lds64(s_own + i * kKcRowBytes + xs * kKcCopyBytes);
                const uint32_t x0 = (s.x ^ ref.x) & valid0, x1 = (s.y ^ ref.y) & valid1;
                sc += __popc(x0) + __popc(x1);
                sc_both2 += __popc(x0 & both2_0) + __popc(x1 & both2_1);
            });
            score = sc;
            if (!__any_sync(0xFFFFFFFFu, ((valid0 & (ge2_0 | gen2_0)) | (valid1 & (ge2_1 | gen2_1))) != 0)) {
                score *= 2;  // every counted cell: fixed point vs fixed point, offset 1 repeats offset 0
            } else {
                // offsets 1..min(lambda_x, lambda_n): step, and compare the new state while it is made.  With no
                // period above 2 in the batch, offset 2 is offset 0 again for the period-2 pairs: one step suffices.
                const uint32_t gen3_0 = lamn_ge(nv0, 3), gen3_1 = SINGLES ? lamn_ge(nv1, 3) : gen3_0;
                const bool short_cycles = !__any_sync(0xFFFFFFFFu, ((valid0 & (planes_ge(lam0, 3) | gen3_0)) | (valid1 & (planes_ge(lam1, 3) | gen3_1))) != 0);
                __syncthreads();  // (offset 0 read the reference copy; the next gather parks words in its partner)
                for (int t = 1; t < steps; ++t) {
                    if (short_cycles && t == 2) {
                        score += sc_both2;
                        break;
                    }
                    const uint32_t gn0 = lamn_ge(nv0, t), gn1 = SINGLES ? lamn_ge(nv1, t) : gn0;
                    const uint32_t act0 = valid0 & planes_ge(lam0, t) & gn0, act1 = valid1 & planes_ge(lam1, t) & gn1;
                    if (!__any_sync(0xFFFFFFFFu, (act0 | act1) != 0)) break;
                    gather(args.np_states_t + (size_t)t * args.wpad * 32 * args.nwp, xs ^ 2, -1, xn ^ 2);
                    const uint32_t srd = s_lane + xs * kKcCopyBytes;
                    uint32_t sct = 0;
                    uint2 e_next = kk_lds64(s_tab);
                    kc_for_own<BASE>(extra, [&](int i) {
                        const uint2 e = e_next;
                        e_next = kk_lds64(s_tab + (i + 1) * kKkEntryBytes);
                        const uint2 f = kk_eval<kKcRowBytes>(e, s_tab + i * kKkEntryBytes, srd);
                        const uint2 ref = kk_lds64(s_own + i * kKcRowBytes + (xs ^ 2) * kKcCopyBytes);
                        kk_sts64(s_own + i * kKcRowBytes + xn * kKcCopyBytes, f);
                        sct += __popc((f.x ^ ref.x) & act0) + __popc((f.y ^ ref.y) & act1);
                    });
                    score += sct;
                    __syncthreads();
                    const int tmp = xs;
                    xs = xn;
                    xn = tmp;
                }
            }
        }
        if (args.negate) score = 0ull - score;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) score += __shfl_xor_sync(0xFFFFFFFFu, score, off);
        if (lane == 0 && score) atomicAdd(args.out_raw + forced_node, score);
        SCB_SWEEP_PHASE(5);
#ifdef SCB_TRACE
        if (tid == 0) atomicAdd(&g_sweep_phase[MODE][15], 1ull);
#endif
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// trajectories: slot word = 32 selected cells; out[m][n][t] bytes
// ---------------------------------------------------------------------------------------
constexpr int kTrajSlots = 32;

__global__ void __launch_bounds__(256)
trajectories_kernel(const uint32_t* __restrict__ planes, int n_nodes, int64_t wpad,
                    const int64_t* __restrict__ cell_idx, int64_t n_selected,
                    const int32_t* __restrict__ net_parent, const uint8_t* __restrict__ net_lut,
                    int steps, int forced_node, uint32_t forced_value, uint8_t* __restrict__ out) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int N = n_nodes;
    const int S = blockDim.x / kTrajSlots;
    uint32_t* cur = smem;
    uint32_t* nxt = cur + (size_t)N * kTrajSlots;
    NodeEntry* tab = reinterpret_cast<NodeEntry*>(nxt + (size_t)N * kTrajSlots);
    const int tid = threadIdx.x;
    const int slot = tid % kTrajSlots;
    const int part = tid / kTrajSlots;
    const int chunk = (N + S - 1) / S;
    const int n_lo = min(N, part * chunk), n_hi = min(N, n_lo + chunk);
    const int64_t first = ((int64_t)blockIdx.x * kTrajSlots + slot) * 32;  // first selected cell of my slot

    for (int n = tid; n < N; n += blockDim.x) tab[n] = make_entry(net_parent, net_lut, n);
    for (int n = n_lo; n < n_hi; ++n) {
        uint32_t w = 0;
        const uint32_t* row = planes + (size_t)n * wpad;
        for (int b = 0; b < 32; ++b) {
            int64_t m = first + b;
            if (m < n_selected) {
                int64_t c = __ldg(cell_idx + m);
                w |= ((__ldg(row + (c >> 5)) >> (c & 31)) & 1u) << b;
            }
        }
        cur[n * kTrajSlots + slot] = w;
    }
    __syncthreads();
    for (int t = 0; t < steps; ++t) {
        for (int n = n_lo; n < n_hi; ++n) {
            uint32_t f = eval_node<kTrajSlots>(tab[n], cur, slot);
            if (n == forced_node) f = forced_value;  // knock-out / knock-in from the first update on
            nxt[n * kTrajSlots + slot] = f;
            for (int b = 0; b < 32; ++b) {
                int64_t m = first + b;
                if (m < n_selected) out[((size_t)m * N + n) * steps + t] = (f >> b) & 1u;
            }
        }
        __syncthreads();
        uint32_t* tmp = cur; cur = nxt; nxt = tmp;
    }
}

// words of a cell-major record: one bit per node, padded to whole 128-bit loads
static int record_words(int n_nodes) { return (((n_nodes + 31) / 32) + 3) & ~3; }

static size_t sweep_smem_bytes(int n_nodes, int slots) {
    return (size_t)4 * n_nodes * slots * 4 + (size_t)n_nodes * (sizeof(NodeOffsets) + sizeof(LutMasks)) +
           (size_t)kMaskPad * (sizeof(NodeOffsets) + sizeof(LutMasks)) +
           (size_t)2 * kRing * slots * 4 + (size_t)2 * slots * 4 +
           (size_t)(slots / 2) * 32 * 4 + (size_t)(1 + kLamPlanes) * (slots / 2) * 4 + (size_t)(n_nodes + 1) * 4;
}

// room per node for deferred cells: every cell (div = 1) -- single nodes send up to 89 % of the
// cells to the cleanup pass on the H1M data, and a list that overflowed would settle its cells
// inside the plane-word CTAs at many times the cost
static int64_t sweep_list_capacity(int64_t n_cells, int div) {
    int64_t cap = n_cells / div;
    if (cap < 65536) cap = 65536;
    return cap < n_cells ? cap : n_cells;
}
constexpr int kListDivA = 1, kListDivB = 1;

constexpr size_t kMaxDynSmem = 232448;  // 227 KB

// Lists of deferred cells: list A takes what the KO/KI pass leaves open after `fast_steps` ring steps
// (slow convergers, periods 3 and 4, hard cells alike), list B the cells the normal run could not
// settle with the ring (straight from the KO/KI pass) and what the recheck pass leaves open after
// the full number of steps (cycles longer than the ring, no repeat).
struct SweepLists {
    int32_t* cells_a;
    int64_t cap_a;
    unsigned int* count_a;
    int32_t* cells_b;
    int64_t cap_b;
    unsigned int* count_b;
    unsigned int* count_a1;  // single runs in list A / list B (filled from the back), retract list (in list A's arrays)
    unsigned int* count_b1;
    unsigned int* count_r;
};

template <int SLOTS>
static int launch_sweep(SweepArgs args, const SweepLists& lists, uint32_t* planes_sorted, unsigned int* bucket_fill,
                        bool sort_cells, cudaStream_t stream) {
    size_t smem = sweep_smem_bytes(args.n_nodes, SLOTS);
    int64_t n_words = (args.n_cells + 31) / 32;
    {
        const void* kernels[4] = {(const void*)sweep_kernel<SLOTS, MODE_NORMAL>, (const void*)sweep_kernel<SLOTS, MODE_KOKI>,
                                  (const void*)sweep_kernel<SLOTS, MODE_RECHECK>, (const void*)sweep_kernel<SLOTS, MODE_CLEANUP>};
        static thread_local SmemMarks marks[4];
        for (int k = 0; k < 4; ++k)
            if (int rc = ensure_dynamic_smem(kernels[k], smem, marks[k])) return rc;
    }
    // count_a1 | count_b1 | count_r | settle_hist | key_hist | bucket_fill | count_a | count_b | counters are contiguous
    SCB_CUDA(cudaMemsetAsync(lists.count_a1, 0, (size_t)args.n_nodes * 12 + 3 * 64 * 4 + (size_t)args.n_nodes * 8 + 16, stream));
    args.in_cells = nullptr;
    args.in_cap = 0;
    args.in_count = nullptr;
    args.out_cells = nullptr;
    args.out_cap = 0;
    args.out_count = nullptr;
    args.probe = 0;
    dim3 grid_n((unsigned)((n_words + SLOTS - 1) / SLOTS));
    const int n_blocks_j = (args.n_nodes + 31) / 32;
    dim3 grid_t((unsigned)((args.wpad + 31) / 32), (unsigned)n_blocks_j, 1);
    if (sort_cells) {
        // Cells sorted by the settle step of their unperturbed run.  A perturbed run of a cell settles at
        // about the same step as its unperturbed one, and a KO/KI CTA steps until (nearly) all of its 1024
        // cells have settled: with similar cells side by side every CTA stops at its own cells' pace instead
        // of the slowest cell's of the whole data set.  Every output is a sum over cells, so the order is
        // free.  probe (key per cell) -> counting sort of the cell-major records -> back to bitplanes.
        uint32_t* rec0 = args.np_states_t;  // records in the caller's order; dead before np_states_t is written
        cell_major_kernel<<<grid_t, 256, 0, stream>>>(args.planes, args.n_nodes, args.wpad, rec0, args.nwp, nullptr, 1);
        SCB_LAUNCH_CHECK("cell_major_kernel<data>");
        args.probe = 1;
        sweep_kernel<SLOTS, MODE_NORMAL><<<grid_n, kSweepThreads, smem, stream>>>(args);
        SCB_LAUNCH_CHECK("sweep_kernel<probe>");
        args.probe = 0;
        sort_scatter_kernel<<<(unsigned)((args.n_cells + 255) / 256), 256, 0, stream>>>(
            args.key_planes, args.wpad, args.n_cells, args.key_hist, bucket_fill, rec0, args.planes_t, args.nwp);
        SCB_LAUNCH_CHECK("sort_scatter_kernel");
        node_major_kernel<<<grid_t, 256, 0, stream>>>(args.planes_t, args.nwp, args.n_nodes, args.wpad, args.n_cells, planes_sorted);
        SCB_LAUNCH_CHECK("node_major_kernel");
        args.planes = planes_sorted;
    }
    sweep_kernel<SLOTS, MODE_NORMAL><<<grid_n, kSweepThreads, smem, stream>>>(args);
    SCB_LAUNCH_CHECK("sweep_kernel<normal>");
    if (args.fast_steps <= 0) {
        choose_fast_steps_kernel<<<1, 32, 0, stream>>>(args.settle_hist, args.steps, args.n_cells, args.counters, sort_cells ? 1 : 0);
        SCB_LAUNCH_CHECK("choose_fast_steps_kernel");
    }
    // cell-major copies of the data and of the normal-run states for the compacted passes
    {
        if (!sort_cells) {  // (sorted: planes_t is what the sort made)
            cell_major_kernel<<<grid_t, 256, 0, stream>>>(args.planes, args.n_nodes, args.wpad, args.planes_t, args.nwp, nullptr, 1);
            SCB_LAUNCH_CHECK("cell_major_kernel<data>");
        }
        grid_t.y = 1;   // np_lam[6][wpad] and np_found[wpad] are contiguous: a 7-row matrix, one word per cell
        cell_major_kernel<<<grid_t, 256, 0, stream>>>(args.np_lam, kLamPlanes + 1, args.wpad, args.np_info_t, 1, nullptr, 1);
        SCB_LAUNCH_CHECK("cell_major_kernel<normal info>");
        grid_t.y = (unsigned)n_blocks_j;
        cell_major_kernel<<<grid_t, 256, 0, stream>>>(args.np_states, args.n_nodes, args.wpad, args.np_states_t, args.nwp,
                                                      args.counters + 2, args.steps);
        SCB_LAUNCH_CHECK("cell_major_kernel<normal run>");
    }
    if (const char* ring = getenv("SCB_SWEEP_RING")) {  // experiments: force the KO/KI ring depth
        if (ring[0] == '4') SCB_CUDA(cudaMemsetAsync(args.counters + 1, 1, 1, stream));
        if (ring[0] == '2') SCB_CUDA(cudaMemsetAsync(args.counters + 1, 0, 1, stream));
    }
    constexpr int HALF = SLOTS / 2;
    const bool two_level = args.fast_steps < args.steps;  // also when the device chooses (fast_steps <= 0)
    args.hard_cells = lists.cells_b;
    args.hard_cap = lists.cap_b;
    args.hard_count = lists.count_b;
    args.recheck_follows = two_level ? 1 : 0;
    // KO/KI on the plane words; open cells go to list A (or straight to B without a recheck pass)
    args.out_cells = two_level ? lists.cells_a : lists.cells_b;
    args.out_cap = two_level ? lists.cap_a : lists.cap_b;
    args.out_count = two_level ? lists.count_a : lists.count_b;
    args.out_count1 = lists.count_a1;
    args.singles = 0;
    args.negate = 0;
    const int64_t n_koki = ((n_words + HALF - 1) / HALF) * args.n_nodes;
    if (n_koki >= ((int64_t)1 << 31)) {
        set_error("ko_ki_sweep: %lld (word block, node) pairs exceed one grid; shard the cells", (long long)n_koki);
        return SCB_ERR_UNSUPPORTED;
    }
    dim3 grid_k((unsigned)n_koki);
    // the two-CTAs-per-SM kernel whenever the recheck pass follows and the nodes fit its register arrays
    // (SCB_SWEEP_KOKI=old keeps sweep_kernel<SLOTS, MODE_KOKI>: experiments, tests of both paths)
    const char* koki_env = getenv("SCB_SWEEP_KOKI");
    const bool plane2 = SLOTS == 64 && two_level && args.n_nodes <= kKkWarps * 13 && !(koki_env && koki_env[0] == 'o') &&
                        koki_plane_smem_bytes(args.n_nodes) <= kMaxDynSmem;
    {
        // single runs need all three new kernels (the sweep_kernel modes know pairs only) and 30-bit cell indices
        const char* rc_e = getenv("SCB_SWEEP_RECHECK");
        const char* cl_e = getenv("SCB_SWEEP_CLEANUP");
        const char* sg_e = getenv("SCB_SWEEP_SINGLES");  // =0: every open cell travels as a pair
        args.singles = plane2 && !(rc_e && rc_e[0] == 'o') && !(cl_e && cl_e[0] == 'o') && !(sg_e && sg_e[0] == '0') &&
                       args.n_nodes / kKcWarps <= 8 && koki_cleanup_smem_bytes(args.n_nodes) <= kMaxDynSmem &&
                       args.wpad * 32 < ((int64_t)1 << kWhichShift);
    }
    if (plane2) {
        static const void* const kk[13] = {
            (const void*)koki_plane_kernel<1>, (const void*)koki_plane_kernel<2>, (const void*)koki_plane_kernel<3>,
            (const void*)koki_plane_kernel<4>, (const void*)koki_plane_kernel<5>, (const void*)koki_plane_kernel<6>,
            (const void*)koki_plane_kernel<7>, (const void*)koki_plane_kernel<8>, (const void*)koki_plane_kernel<9>,
            (const void*)koki_plane_kernel<10>, (const void*)koki_plane_kernel<11>, (const void*)koki_plane_kernel<12>,
            (const void*)koki_plane_kernel<13>};
        static thread_local SmemMarks kk_marks[13];
        static thread_local SmemMarks kk_carve[13];
        const int v = koki_plane_cnt(args.n_nodes) - 1;
        const size_t kk_smem = koki_plane_smem_bytes(args.n_nodes);
        if (int rc = ensure_dynamic_smem(kk[v], kk_smem, kk_marks[v])) return rc;
        if (int rc = ensure_max_carveout(kk[v], kk_carve[v])) return rc;
        void* kargs[1] = {(void*)&args};
        SCB_CUDA(cudaLaunchKernel(kk[v], grid_k, dim3(kKkThreads), kargs, kk_smem, stream));
        SCB_LAUNCH_CHECK("koki_plane_kernel");
    } else {
        sweep_kernel<SLOTS, MODE_KOKI><<<grid_k, kSweepThreads, smem, stream>>>(args);
        SCB_LAUNCH_CHECK("sweep_kernel<koki>");
    }
    int per_sm = smem <= kMaxDynSmem / 2 ? 2 : 1;
    // the compacted passes: koki_compact_kernel<BASE, RECHECK, SINGLES> whenever the nodes fit
    // (SCB_SWEEP_RECHECK=old / SCB_SWEEP_CLEANUP=old keep the sweep_kernel modes; they know pairs only)
    const int kc_base = args.n_nodes / kKcWarps;
    const bool compact_fits = SLOTS == 64 && kc_base <= 8 && koki_cleanup_smem_bytes(args.n_nodes) <= kMaxDynSmem;
    auto launch_compact = [&](bool recheck, bool single_runs, const char* what) -> int {
        static const void* const table[2][2][9] = {
            {{(const void*)koki_compact_kernel<0, false, false>, (const void*)koki_compact_kernel<1, false, false>, (const void*)koki_compact_kernel<2, false, false>,
              (const void*)koki_compact_kernel<3, false, false>, (const void*)koki_compact_kernel<4, false, false>, (const void*)koki_compact_kernel<5, false, false>,
              (const void*)koki_compact_kernel<6, false, false>, (const void*)koki_compact_kernel<7, false, false>, (const void*)koki_compact_kernel<8, false, false>},
             {(const void*)koki_compact_kernel<0, false, true>, (const void*)koki_compact_kernel<1, false, true>, (const void*)koki_compact_kernel<2, false, true>,
              (const void*)koki_compact_kernel<3, false, true>, (const void*)koki_compact_kernel<4, false, true>, (const void*)koki_compact_kernel<5, false, true>,
              (const void*)koki_compact_kernel<6, false, true>, (const void*)koki_compact_kernel<7, false, true>, (const void*)koki_compact_kernel<8, false, true>}},
            {{(const void*)koki_compact_kernel<0, true, false>, (const void*)koki_compact_kernel<1, true, false>, (const void*)koki_compact_kernel<2, true, false>,
              (const void*)koki_compact_kernel<3, true, false>, (const void*)koki_compact_kernel<4, true, false>, (const void*)koki_compact_kernel<5, true, false>,
              (const void*)koki_compact_kernel<6, true, false>, (const void*)koki_compact_kernel<7, true, false>, (const void*)koki_compact_kernel<8, true, false>},
             {(const void*)koki_compact_kernel<0, true, true>, (const void*)koki_compact_kernel<1, true, true>, (const void*)koki_compact_kernel<2, true, true>,
              (const void*)koki_compact_kernel<3, true, true>, (const void*)koki_compact_kernel<4, true, true>, (const void*)koki_compact_kernel<5, true, true>,
              (const void*)koki_compact_kernel<6, true, true>, (const void*)koki_compact_kernel<7, true, true>, (const void*)koki_compact_kernel<8, true, true>}}};
        static thread_local SmemMarks marks[2][2][9];
        const void* kernel = table[recheck ? 1 : 0][single_runs ? 1 : 0][kc_base];
        const size_t kc_smem = koki_cleanup_smem_bytes(args.n_nodes);
        if (int rc = ensure_dynamic_smem(kernel, kc_smem, marks[recheck ? 1 : 0][single_runs ? 1 : 0][kc_base])) return rc;
        void* kargs[1] = {(void*)&args};
        SCB_CUDA(cudaLaunchKernel(kernel, dim3((unsigned)sm_count()), dim3(kKcThreads), kargs, kc_smem, stream));
        SCB_LAUNCH_CHECK(what);
        return SCB_OK;
    };
    const char* rc_env = getenv("SCB_SWEEP_RECHECK");
    const char* cl_env = getenv("SCB_SWEEP_CLEANUP");
    const bool new_recheck = compact_fits && !(rc_env && rc_env[0] == 'o');
    const bool new_cleanup = compact_fits && !(cl_env && cl_env[0] == 'o');
    args.negate = 0;
    if (two_level) {
        // slow convergers, compacted: a persistent grid runs the full-length ring on list A (pairs, then single runs)
        args.in_cells = lists.cells_a;
        args.in_cap = lists.cap_a;
        args.in_count = lists.count_a;
        args.in_count1 = lists.count_a1;
        args.out_cells = lists.cells_b;
        args.out_cap = lists.cap_b;
        args.out_count = lists.count_b;
        args.out_count1 = lists.count_b1;
        if (new_recheck) {
            if (int rc = launch_compact(true, false, "koki_compact_kernel<recheck>")) return rc;
            if (args.singles)
                if (int rc = launch_compact(true, true, "koki_compact_kernel<recheck, singles>")) return rc;
        } else {
            sweep_kernel<SLOTS, MODE_RECHECK><<<(unsigned)(sm_count() * per_sm), kSweepThreads, smem, stream>>>(args);
            SCB_LAUNCH_CHECK("sweep_kernel<recheck>");
        }
    }
    // stragglers, compacted: the general slow path on list B (pairs, then single runs, then the retract list)
    args.in_cells = lists.cells_b;
    args.in_cap = lists.cap_b;
    args.in_count = lists.count_b;
    args.in_count1 = lists.count_b1;
    args.out_cells = nullptr;
    args.out_cap = 0;
    args.out_count = nullptr;
    args.out_count1 = nullptr;
    args.rt_cells = lists.cells_a;  // (list A is spent by now)
    args.rt_cap = lists.cap_a;
    args.rt_count = lists.count_r;
    if (new_cleanup) {
        if (int rc = launch_compact(false, false, "koki_compact_kernel<cleanup>")) return rc;
        if (args.singles) {
            if (int rc = launch_compact(false, true, "koki_compact_kernel<cleanup, singles>")) return rc;
            // runs whose partner did not repeat after all: simulate them once more and take their scores back
            args.in_cells = lists.cells_a;
            args.in_cap = lists.cap_a;
            args.in_count1 = lists.count_r;
            args.negate = 1;
            if (int rc = launch_compact(true, true, "koki_compact_kernel<retract>")) return rc;
            args.negate = 0;
        }
    } else {
        sweep_kernel<SLOTS, MODE_CLEANUP><<<(unsigned)(sm_count() * per_sm), kSweepThreads, smem, stream>>>(args);
        SCB_LAUNCH_CHECK("sweep_kernel<cleanup>");
    }
    return SCB_OK;
}

}  // namespace scb

using namespace scb;

#ifdef SCB_TRACE
extern "C" int scb_debug_sweep_phases(unsigned long long* out /*HOST [4][16]*/, int reset) {
    SCB_CUDA(cudaDeviceSynchronize());
    SCB_CUDA(cudaMemcpyFromSymbol(out, g_sweep_phase, sizeof(unsigned long long) * 64));
    if (reset) {
        unsigned long long zero[64] = {0};
        SCB_CUDA(cudaMemcpyToSymbol(g_sweep_phase, zero, sizeof(zero)));
    }
    return SCB_OK;
}
#endif

extern "C" int64_t scb_ko_ki_sweep_scratch_bytes(int n_nodes, int64_t wpad, int steps) {
    if (n_nodes <= 0 || wpad <= 0 || steps <= 0) return 0;
    // normal-run states + lambda planes + found plane + the two per-node lists of deferred cells +
    // their fill counts + counters
    // ... + the bitplanes in sorted cell order and the 6 key planes of the sort
    return ((int64_t)steps * n_nodes + kLamPlanes + 2) * wpad * 4 +
           (int64_t)(steps + 1) * wpad * 32 * record_words(n_nodes) * 4 + wpad * 32 * 4 +
           (int64_t)n_nodes * (sweep_list_capacity(wpad * 32, kListDivA) + sweep_list_capacity(wpad * 32, kListDivB)) * 4 +
           ((int64_t)n_nodes + 6) * wpad * 4 +
           (int64_t)n_nodes * 12 + 3 * 64 * 4 + (int64_t)n_nodes * 8 + 16;
}

extern "C" int scb_ko_ki_sweep(const uint32_t* planes, int n_nodes, int64_t wpad, int64_t n_cells,
                               const int32_t* net_parent, const uint8_t* net_lut, int steps,
                               void* scratch, int64_t scratch_bytes, int64_t* out_raw,
                               int64_t* out_missing, int64_t* out_keyerror, scb_stream_t stream) {
    SCB_REQUIRE(n_nodes > 0 && n_nodes <= 65535, "ko_ki_sweep: n_nodes=%d out of range", n_nodes);
    SCB_REQUIRE(steps >= 2 && steps <= 64, "ko_ki_sweep: steps=%d must be in [2, 64]", steps);
    SCB_REQUIRE(wpad > 0 && wpad % 4 == 0 && n_cells >= 0 && wpad * 32 >= n_cells, "ko_ki_sweep: wpad must be a positive multiple of 4 covering n_cells");
    SCB_REQUIRE(wpad * 32 < ((int64_t)1 << 31), "ko_ki_sweep: at most 2^31-1 cells per call (shard the cells)");
    SCB_REQUIRE(planes && net_parent && net_lut && scratch && out_raw && out_missing && out_keyerror, "ko_ki_sweep: null pointer");
    SCB_REQUIRE(reinterpret_cast<uintptr_t>(scratch) % 16 == 0, "ko_ki_sweep: scratch must be 16-byte aligned");
    if (n_cells == 0) return SCB_OK;
    int64_t need = scb_ko_ki_sweep_scratch_bytes(n_nodes, wpad, steps);
    if (scratch_bytes < need) {
        set_error("ko_ki_sweep: scratch too small (%lld < %lld bytes)", (long long)scratch_bytes, (long long)need);
        return SCB_ERR_SCRATCH;
    }
    SweepArgs args;
    args.planes = planes;
    args.n_nodes = n_nodes;
    args.wpad = wpad;
    args.n_cells = n_cells;
    args.net_parent = net_parent;
    args.net_lut = net_lut;
    args.steps = steps;
    args.np_states = static_cast<uint32_t*>(scratch);
    args.np_lam = args.np_states + (size_t)steps * n_nodes * wpad;
    args.np_found = args.np_lam + (size_t)kLamPlanes * wpad;
    SweepLists lists;
    lists.cap_a = sweep_list_capacity(wpad * 32, kListDivA);
    lists.cap_b = sweep_list_capacity(wpad * 32, kListDivB);
    args.nwp = record_words(n_nodes);
    args.np_hard = args.np_found + wpad;
    args.planes_t = args.np_hard + wpad;
    args.np_states_t = args.planes_t + (size_t)wpad * 32 * args.nwp;
    args.np_info_t = args.np_states_t + (size_t)steps * wpad * 32 * args.nwp;
    lists.cells_a = reinterpret_cast<int32_t*>(args.np_info_t + (size_t)wpad * 32);
    lists.cells_b = lists.cells_a + (size_t)n_nodes * lists.cap_a;
    uint32_t* planes_sorted = reinterpret_cast<uint32_t*>(lists.cells_b + (size_t)n_nodes * lists.cap_b);
    args.key_planes = planes_sorted + (size_t)n_nodes * wpad;
    lists.count_a1 = reinterpret_cast<unsigned int*>(args.key_planes + (size_t)6 * wpad);
    lists.count_b1 = lists.count_a1 + n_nodes;
    lists.count_r = lists.count_b1 + n_nodes;
    args.settle_hist = lists.count_r + n_nodes;
    args.key_hist = args.settle_hist + 64;
    unsigned int* bucket_fill = args.key_hist + 64;
    lists.count_a = bucket_fill + 64;
    lists.count_b = lists.count_a + n_nodes;
    args.counters = lists.count_b + n_nodes;  // followed by 16 spare bytes
    // Default: the KO/KI pass runs the ring on the plane words for a number of steps chosen on the
    // device from the normal run (choose_fast_steps_kernel), cells still open go through the compacted
    // recheck pass (deep ring, all `steps`), what that leaves open through the general scheme.
    // SCB_SWEEP_FAST_STEPS=k fixes the number (k >= steps: no recheck pass).  Measured on imp-1M:
    // 186 ms (k=50), 164 ms (k=24).
    args.fast_steps = 0;
    if (const char* e = getenv("SCB_SWEEP_FAST_STEPS")) args.fast_steps = atoi(e) > 0 ? atoi(e) : 0;
    if (args.fast_steps > steps) args.fast_steps = steps;
    // SCB_SWEEP_SORT=0 keeps the caller's cell order (experiments, tests of both paths)
    bool sort_cells = true;
    if (const char* e = getenv("SCB_SWEEP_SORT")) sort_cells = e[0] != '0';
    args.stop_open = 64;  // (24 before single runs made a deferred run half as expensive; flat between 64 and 160)
    args.stop_stall = 4;
    if (const char* e = getenv("SCB_SWEEP_STOP_OPEN")) args.stop_open = atoi(e) >= 0 ? atoi(e) : 0;
    if (const char* e = getenv("SCB_SWEEP_STOP_STALL")) args.stop_stall = atoi(e) > 0 ? atoi(e) : (1 << 30);
    args.brent_mid = 1;
    if (const char* e = getenv("SCB_SWEEP_BRENT_MID")) args.brent_mid = e[0] != '0';
    args.ring_stall = 6;
    if (const char* e = getenv("SCB_SWEEP_RING_STALL")) args.ring_stall = atoi(e) > 0 ? atoi(e) : (1 << 30);
    if (!sort_cells && !getenv("SCB_SWEEP_STOP_OPEN")) {  // unsorted: the global step cap does the stopping
        args.stop_open = 0;
        args.stop_stall = 1 << 30;
    }
    args.out_raw = reinterpret_cast<unsigned long long*>(out_raw);
    args.out_missing = reinterpret_cast<unsigned long long*>(out_missing);
    args.out_keyerror = reinterpret_cast<unsigned long long*>(out_keyerror);
    const char* force = getenv("SCB_SWEEP_SLOTS");  // experiments
    const bool prefer32 = force && force[0] == '3';
    if (!prefer32 && sweep_smem_bytes(n_nodes, 64) <= kMaxDynSmem)
        return launch_sweep<64>(args, lists, planes_sorted, bucket_fill, sort_cells, as_stream(stream));
    if (sweep_smem_bytes(n_nodes, 32) <= kMaxDynSmem)
        return launch_sweep<32>(args, lists, planes_sorted, bucket_fill, sort_cells, as_stream(stream));
    set_error("ko_ki_sweep: %d nodes exceed the shared-memory resident limit (%d)", n_nodes, SCB_MAX_SWEEP_NODES);
    return SCB_ERR_UNSUPPORTED;
}

extern "C" int scb_trajectories(const uint32_t* planes, int n_nodes, int64_t wpad, int64_t n_cells,
                                const int64_t* cell_idx, int64_t n_selected,
                                const int32_t* net_parent, const uint8_t* net_lut, int steps,
                                uint8_t* out, scb_stream_t stream) {
    return scb_trajectories_forced(planes, n_nodes, wpad, n_cells, cell_idx, n_selected, net_parent, net_lut,
                                   steps, -1, 0, out, stream);
}

extern "C" int scb_trajectories_forced(const uint32_t* planes, int n_nodes, int64_t wpad, int64_t n_cells,
                                       const int64_t* cell_idx, int64_t n_selected,
                                       const int32_t* net_parent, const uint8_t* net_lut, int steps,
                                       int forced_node, int forced_value, uint8_t* out,
                                       scb_stream_t stream) {
    SCB_REQUIRE(forced_node >= -1 && forced_node < n_nodes, "trajectories: forced node out of range");
    SCB_REQUIRE(n_nodes > 0 && steps >= 1 && n_selected >= 0, "trajectories: bad size");
    SCB_REQUIRE(wpad > 0 && wpad * 32 >= n_cells, "trajectories: wpad does not cover n_cells");
    if (n_selected == 0) return SCB_OK;
    SCB_REQUIRE(planes && cell_idx && net_parent && net_lut && out, "trajectories: null pointer");
    size_t smem = (size_t)2 * n_nodes * kTrajSlots * 4 + (size_t)n_nodes * sizeof(NodeEntry);
    if (smem > kMaxDynSmem) {
        set_error("trajectories: %d nodes exceed the shared-memory resident limit", n_nodes);
        return SCB_ERR_UNSUPPORTED;
    }
    SCB_CUDA(cudaFuncSetAttribute(trajectories_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t slots = (n_selected + 31) / 32;
    trajectories_kernel<<<(unsigned)((slots + kTrajSlots - 1) / kTrajSlots), 256, smem, as_stream(stream)>>>(
        planes, n_nodes, wpad, cell_idx, n_selected, net_parent, net_lut, steps, forced_node,
        forced_value ? 0xFFFFFFFFu : 0u, out);
    SCB_LAUNCH_CHECK("trajectories_kernel");
    return SCB_OK;
}
