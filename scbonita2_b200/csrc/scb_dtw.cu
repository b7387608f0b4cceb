// Pairwise trajectory distances between cells.
//
// Replaces compute_dtw_distance_pair / compute_dtw_distances (attractor_analysis.py:304-377): for
// every pair of simulated cells, the sum over the genes of
//     squared difference of the two down-sampled 0/1 series        if it is < EUCLIDEAN_THRESHOLD (5)
//     fastdtw(series1, series2, radius, squared pointwise distance)  otherwise.
// A gene's down-sampled series is a PATTERN of L bits (L = 10 for 20 steps, every 2nd), so both
// branches are a function of two L-bit patterns: a table of 2^L x 2^L bytes (1 MB) built once on
// the host, after which a pair of cells costs one table read per gene.  The reference spends a
// Python loop with NumPy slicing per gene and a process pool over the M(M-1)/2 pairs.
//
// FastDTW is the third-party package `fastdtw` 0.3.4 (environment.yml:103), absent from the
// reference tree: its published algorithm is restated here (halve by pair averages down to radius+2
// points, exact DTW, project the path, widen by the radius, windowed DTW one level up; ties in the
// order up / left / diagonal).  PARITY UNPINNED for the table entries of that branch; the squared-
// difference branch is exact by construction.  The test oracle holds an independent restatement.
#include <limits>

#include "scb_common.cuh"

namespace scb {

namespace {

constexpr int kMaxPoints = 12;

struct Series {
    double v[kMaxPoints];
    int n;
};
struct Path {
    signed char i[2 * kMaxPoints + 2], j[2 * kMaxPoints + 2];
    int n;
};
// cells (i, j) the DTW may visit, row by row in increasing j; full = every cell
struct Window {
    bool full;
    unsigned short row[kMaxPoints];  // bit j of row[i]
};

// DTW restricted to `window`; returns the distance and the path (first minimum in the order up,
// left, diagonal -- Python's min() over that tuple)
double windowed_dtw(const Series& x, const Series& y, const Window& w, Path* path) {
    const double inf = std::numeric_limits<double>::infinity();
    double cost[kMaxPoints + 1][kMaxPoints + 1];
    signed char pi[kMaxPoints + 1][kMaxPoints + 1], pj[kMaxPoints + 1][kMaxPoints + 1];
    for (int i = 0; i <= x.n; ++i)
        for (int j = 0; j <= y.n; ++j) cost[i][j] = inf;
    cost[0][0] = 0.0;
    pi[0][0] = pj[0][0] = 0;
    for (int i = 1; i <= x.n; ++i)
        for (int j = 1; j <= y.n; ++j) {
            if (!w.full && !((w.row[i - 1] >> (j - 1)) & 1)) continue;
            const double d = x.v[i - 1] - y.v[j - 1];
            const double dt = d * d;
            double best = cost[i - 1][j] + dt;
            int bi = i - 1, bj = j;
            if (cost[i][j - 1] + dt < best) { best = cost[i][j - 1] + dt; bi = i; bj = j - 1; }
            if (cost[i - 1][j - 1] + dt < best) { best = cost[i - 1][j - 1] + dt; bi = i - 1; bj = j - 1; }
            cost[i][j] = best;
            pi[i][j] = (signed char)bi;
            pj[i][j] = (signed char)bj;
        }
    if (path) {
        Path rev;
        rev.n = 0;
        int i = x.n, j = y.n;
        while (!(i == 0 && j == 0)) {
            rev.i[rev.n] = (signed char)(i - 1);
            rev.j[rev.n] = (signed char)(j - 1);
            ++rev.n;
            const int ni = pi[i][j], nj = pj[i][j];
            i = ni;
            j = nj;
        }
        path->n = rev.n;
        for (int k = 0; k < rev.n; ++k) {
            path->i[k] = rev.i[rev.n - 1 - k];
            path->j[k] = rev.j[rev.n - 1 - k];
        }
    }
    return cost[x.n][y.n];
}

Series reduce_by_half(const Series& x) {
    Series r;
    r.n = 0;
    for (int i = 0; i + 1 < x.n; i += 2) r.v[r.n++] = (x.v[i] + x.v[i + 1]) / 2;  // an odd tail point is dropped
    return r;
}

// the coarse path widened by `radius`, every coarse cell blown up to 2 x 2 fine cells; a row keeps
// the first contiguous run of member cells at or after the previous row's start (the package's scan)
Window expand_window(const Path& path, int lx, int ly, int radius) {
    bool fine[kMaxPoints + 8][kMaxPoints + 8] = {};
    for (int k = 0; k < path.n; ++k)
        for (int a = -radius; a <= radius; ++a)
            for (int b = -radius; b <= radius; ++b) {
                const int ci = path.i[k] + a, cj = path.j[k] + b;
                if (ci < 0 || cj < 0) continue;  // negative cells never reach the fine grid
                for (int u = 0; u < 2; ++u)
                    for (int v = 0; v < 2; ++v)
                        if (2 * ci + u < kMaxPoints + 8 && 2 * cj + v < kMaxPoints + 8) fine[2 * ci + u][2 * cj + v] = true;
            }
    Window w;
    w.full = false;
    int start_j = 0;
    for (int i = 0; i < lx; ++i) {
        w.row[i] = 0;
        int new_start = -1;
        for (int j = start_j; j < ly; ++j) {
            if (fine[i][j]) {
                w.row[i] |= (unsigned short)(1u << j);
                if (new_start < 0) new_start = j;
            } else if (new_start >= 0) {
                break;
            }
        }
        // the package carries None forward here (and then fails); a connected path never leaves a row empty
        start_j = new_start < 0 ? 0 : new_start;
    }
    return w;
}

double fast_dtw(const Series& x, const Series& y, int radius, Path* path) {
    const int min_size = radius + 2;
    Window w;
    w.full = true;
    if (x.n < min_size || y.n < min_size) return windowed_dtw(x, y, w, path);
    Path coarse;
    fast_dtw(reduce_by_half(x), reduce_by_half(y), radius, &coarse);
    w = expand_window(coarse, x.n, y.n, radius);
    return windowed_dtw(x, y, w, path);
}

}  // namespace

// pattern of one (cell, gene): bit k = traj[m][n][k * factor], k < ceil(steps / factor)
__global__ void trajectory_patterns_kernel(const uint8_t* __restrict__ traj, int64_t n_pairs, int steps, int factor,
                                           uint16_t* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const uint8_t* row = traj + (size_t)i * steps;
    uint32_t p = 0;
    for (int k = 0, t = 0; t < steps; ++k, t += factor) p |= (uint32_t)(row[t] != 0) << k;
    out[i] = (uint16_t)p;
}

// out[a][b] = sum over genes of table[pattern[a][g]][pattern[b][g]].  A CTA owns a 16 x 16 block of
// cell pairs; the patterns of its 32 cells are staged in shared memory gene-major, so the two
// pattern reads of a thread are broadcast / conflict free and the table reads go to L1 / L2 (the
// table is 1 MB at 10 points).  Only blocks on or above the diagonal run; they write both halves.
constexpr int kPairTile = 16;
__global__ void __launch_bounds__(kPairTile* kPairTile)
pair_distance_kernel(const uint16_t* __restrict__ patterns, int64_t n_cells, int n_genes,
                     const uint8_t* __restrict__ table, int bits, int32_t* __restrict__ out) {
    if (blockIdx.x < blockIdx.y) return;  // lower triangle of blocks
    extern __shared__ uint16_t sh[];      // [2][n_genes][kPairTile]
    uint16_t* pa = sh;
    uint16_t* pb = sh + (size_t)n_genes * kPairTile;
    const int64_t a0 = (int64_t)blockIdx.y * kPairTile, b0 = (int64_t)blockIdx.x * kPairTile;
    for (int e = threadIdx.x; e < n_genes * kPairTile; e += blockDim.x) {
        const int c = e / n_genes, g = e % n_genes;  // coalesced along the genes of a cell
        pa[g * kPairTile + c] = a0 + c < n_cells ? patterns[(size_t)(a0 + c) * n_genes + g] : 0;
        pb[g * kPairTile + c] = b0 + c < n_cells ? patterns[(size_t)(b0 + c) * n_genes + g] : 0;
    }
    __syncthreads();
    const int ia = threadIdx.x / kPairTile, ib = threadIdx.x % kPairTile;
    int32_t sum = 0;
#pragma unroll 4
    for (int g = 0; g < n_genes; ++g)
        sum += __ldg(table + (((size_t)pa[g * kPairTile + ia]) << bits) + pb[g * kPairTile + ib]);
    const int64_t a = a0 + ia, b = b0 + ib;
    // FastDTW is not symmetric in its arguments: like the reference (itertools.combinations, first cell
    // = first argument; create_distance_matrix mirrors the value) only a <= b is computed
    if (a <= b && b < n_cells) {
        out[(size_t)a * n_cells + b] = sum;
        out[(size_t)b * n_cells + a] = sum;
    }
}

}  // namespace scb

using namespace scb;

extern "C" int scb_dtw_table_build(int bits, int radius, int threshold, uint8_t* host_table) {
    SCB_REQUIRE(bits >= 1 && bits <= 12, "dtw_table_build: %d points per series (supported: 1..12)", bits);
    SCB_REQUIRE(radius >= 0 && threshold >= 0 && host_table, "dtw_table_build: bad argument");
    const int n = 1 << bits;
    Series x, y;
    x.n = y.n = bits;
    for (int px = 0; px < n; ++px) {
        for (int k = 0; k < bits; ++k) x.v[k] = (px >> k) & 1;
        for (int py = 0; py < n; ++py) {
            const int hamming = __builtin_popcount((unsigned)(px ^ py));
            double d = hamming;  // squared difference of 0/1 series
            if (hamming >= threshold) {
                for (int k = 0; k < bits; ++k) y.v[k] = (py >> k) & 1;
                d = fast_dtw(x, y, radius, nullptr);
            }
            if (!(d >= 0 && d <= 255 && d == (double)(int)d)) {
                set_error("dtw_table_build: distance %g of patterns %d, %d does not fit the byte table", d, px, py);
                return SCB_ERR_UNSUPPORTED;
            }
            host_table[(size_t)px * n + py] = (uint8_t)d;
        }
    }
    return SCB_OK;
}

extern "C" int scb_trajectory_patterns(const uint8_t* traj, int64_t n_cells, int n_genes, int steps, int factor,
                                       uint16_t* out, scb_stream_t stream) {
    SCB_REQUIRE(n_cells >= 0 && n_genes > 0 && steps >= 1 && factor >= 1, "trajectory_patterns: bad size");
    SCB_REQUIRE((steps + factor - 1) / factor <= 12, "trajectory_patterns: more than 12 points per series");
    if (n_cells == 0) return SCB_OK;
    SCB_REQUIRE(traj && out, "trajectory_patterns: null pointer");
    const int64_t n_pairs = n_cells * n_genes;
    trajectory_patterns_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, as_stream(stream)>>>(traj, n_pairs, steps,
                                                                                                factor, out);
    SCB_LAUNCH_CHECK("trajectory_patterns_kernel");
    return SCB_OK;
}

extern "C" int scb_pattern_pair_distances(const uint16_t* patterns, int64_t n_cells, int n_genes,
                                          const uint8_t* table, int bits, int32_t* out, scb_stream_t stream) {
    SCB_REQUIRE(n_cells >= 0 && n_genes > 0 && bits >= 1 && bits <= 12, "pattern_pair_distances: bad size");
    if (n_cells == 0) return SCB_OK;
    SCB_REQUIRE(patterns && table && out, "pattern_pair_distances: null pointer");
    const size_t smem = (size_t)2 * n_genes * kPairTile * sizeof(uint16_t);
    SCB_REQUIRE(smem <= 200 * 1024, "pattern_pair_distances: %d genes exceed the shared-memory tile", n_genes);
    const int64_t blocks = (n_cells + kPairTile - 1) / kPairTile;
    SCB_REQUIRE(blocks <= 65535, "pattern_pair_distances: at most %d cells per call", 65535 * kPairTile);
    static thread_local SmemMarks marks;
    if (int rc = ensure_dynamic_smem((const void*)pair_distance_kernel, smem, marks)) return rc;
    pair_distance_kernel<<<dim3((unsigned)blocks, (unsigned)blocks), kPairTile * kPairTile, smem, as_stream(stream)>>>(
        patterns, n_cells, n_genes, table, bits, out);
    SCB_LAUNCH_CHECK("pair_distance_kernel");
    return SCB_OK;
}
