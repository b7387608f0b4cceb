/*
 * scb_oracle.c — plain C restatement of the scBONITA2 hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Same algorithms as oracle/ref_port.py (which is pinned to the unmodified reference through
 * tests/golden/), written cell by cell with byte arrays so that mid-size cases (10^4..10^5
 * cells) finish in seconds.  Nothing here is bit-sliced or shares code with the CUDA path.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library.
 *
 *   oracle_rule_difference   RuleDetermination.calculate_error   (rule_determination.py:509-582)
 *   oracle_importance_raw    CalculateImportanceScore.*          (importance_scores.py:47-143,256-340)
 *   oracle_trajectory        attractor_analysis.vectorized_run_simulation (attractor_analysis.py:43-99)
 *
 * dense: uint8 [n_rows][n_cells] of 0/1.  A rule is an 8-bit truth table over the slots A,B,C
 * (bit A<<2|B<<1|C); parent row -1 = unbound slot, reads 0.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int lut_bit(uint8_t lut, int a, int b, int c) { return (lut >> ((a << 2) | (b << 1) | c)) & 1; }

/* number of cells whose predicted value differs from the target row (:577) */
int64_t oracle_rule_difference(const uint8_t* dense, int64_t n_cells, int target, const int32_t* parent,
                               uint8_t lut) {
    int64_t diff = 0;
    for (int64_t c = 0; c < n_cells; ++c) {
        int a = parent[0] >= 0 ? dense[(int64_t)parent[0] * n_cells + c] : 0;
        int b = parent[1] >= 0 ? dense[(int64_t)parent[1] * n_cells + c] : 0;
        int cc = parent[2] >= 0 ? dense[(int64_t)parent[2] * n_cells + c] : 0;
        diff += lut_bit(lut, a, b, cc) != dense[(int64_t)target * n_cells + c];
    }
    return diff;
}

/* one synchronous update of one cell (importance_scores.py:60-105); forced node: -1 none */
static void step_cell(const uint8_t* cur, uint8_t* nxt, int n, const int32_t* parent, const uint8_t* lut,
                      int forced, int forced_value) {
    for (int i = 0; i < n; ++i) {
        if (i == forced) { nxt[i] = (uint8_t)forced_value; continue; }
        const int32_t* p = parent + 3 * i;
        int a = p[0] >= 0 ? cur[p[0]] : 0, b = p[1] >= 0 ? cur[p[1]] : 0, c = p[2] >= 0 ? cur[p[2]] : 0;
        nxt[i] = (uint8_t)lut_bit(lut[i], a, b, c);
    }
}

/* history[steps][n] of one cell; returns 1 and (start, end) of the first repeat (:134-143) */
static int simulate_cell(const uint8_t* start, int n, const int32_t* parent, const uint8_t* lut, int steps,
                         int forced, int forced_value, uint8_t* hist, int* j_out, int* i_out) {
    for (int t = 0; t < steps; ++t)
        step_cell(t == 0 ? start : hist + (size_t)(t - 1) * n, hist + (size_t)t * n, n, parent, lut, forced,
                  forced_value);
    for (int i = 0; i < steps; ++i)
        for (int j = 0; j < i; ++j)
            if (memcmp(hist + (size_t)i * n, hist + (size_t)j * n, (size_t)n) == 0) {
                *j_out = j;
                *i_out = i;
                return 1;
            }
    return 0;
}

/* XOR count of two attractors aligned at their own starts, trimmed to the shorter (:325-340) */
static int64_t aligned_xor(const uint8_t* hx, int jx, int ix, const uint8_t* hn, int jn, int in_, int n) {
    int len_x = ix - jx + 1, len_n = in_ - jn + 1;
    int m = len_x < len_n ? len_x : len_n;
    int64_t d = 0;
    for (int t = 0; t < m; ++t)
        for (int k = 0; k < n; ++k) d += hx[(size_t)(jx + t) * n + k] ^ hn[(size_t)(jn + t) * n + k];
    return d;
}

/* raw[n], missing[2n+1] (cells without a repeat per condition), returns #(cell,node) pairs that
 * repeat under KO and KI but not normally (the reference raises KeyError there, :281) */
int64_t oracle_importance_raw(const uint8_t* dense, int n, int64_t n_cells, const int32_t* parent,
                              const uint8_t* lut, int steps, int64_t* raw, int64_t* missing) {
    uint8_t* start = (uint8_t*)malloc((size_t)n);
    uint8_t* hn = (uint8_t*)malloc((size_t)steps * n);
    uint8_t* ho = (uint8_t*)malloc((size_t)steps * n);
    uint8_t* hi = (uint8_t*)malloc((size_t)steps * n);
    int64_t keyerror = 0;
    memset(raw, 0, sizeof(int64_t) * (size_t)n);
    memset(missing, 0, sizeof(int64_t) * (size_t)(2 * n + 1));
    for (int64_t c = 0; c < n_cells; ++c) {
        for (int k = 0; k < n; ++k) start[k] = dense[(int64_t)k * n_cells + c];
        int jn = 0, in_ = 0, jo, io, ji, ii;
        int has_n = simulate_cell(start, n, parent, lut, steps, -1, 0, hn, &jn, &in_);
        if (!has_n) missing[0]++;
        for (int k = 0; k < n; ++k) {
            int has_o = simulate_cell(start, n, parent, lut, steps, k, 0, ho, &jo, &io);
            int has_i = simulate_cell(start, n, parent, lut, steps, k, 1, hi, &ji, &ii);
            if (!has_o) missing[1 + 2 * k]++;
            if (!has_i) missing[2 + 2 * k]++;
            if (has_o && has_i) {
                if (!has_n) { keyerror++; continue; }
                raw[k] += aligned_xor(hi, ji, ii, hn, jn, in_, n) + aligned_xor(ho, jo, io, hn, jn, in_, n);
            }
        }
    }
    free(start); free(hn); free(ho); free(hi);
    return keyerror;
}

/* out[steps][n]: states after 1..steps updates of one cell (attractor_analysis.py:43-99) */
void oracle_trajectory(const uint8_t* start, int n, const int32_t* parent, const uint8_t* lut, int steps,
                       uint8_t* out) {
    for (int t = 0; t < steps; ++t)
        step_cell(t == 0 ? start : out + (size_t)(t - 1) * n, out + (size_t)t * n, n, parent, lut, -1, 0);
}
