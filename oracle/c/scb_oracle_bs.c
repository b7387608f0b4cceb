/*
 * scb_oracle_bs.c — 64-cells-per-word C restatement of the importance sweep.  TEST INFRASTRUCTURE ONLY.
 *
 * Same routines as scb_oracle.c / oracle/ref_port.py (CalculateImportanceScore,
 * importance_scores.py:47-143 simulation, :134-143 find_attractors, :256-340 scoring and
 * align_attractors), but 64 cells share a machine word so that 10^5..10^6 cells finish in seconds
 * to minutes on the host cores.  Deliberately NOT the CUDA path's scheme: the FULL history
 * hist[steps][n] of every run is kept and the first repeat is found by the reference's own
 * all-pairs search (i ascending, j < i ascending) on it; nothing here detects cycles with rings,
 * checkpoints or replays.  Pinned to the reference goldens in tests/test_oracle_c.py.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library.
 *
 * planes: uint64 [n][w64], bit b of word w = cell 64*w + b, tail bits zero.  A rule is an 8-bit
 * truth table over the slots A,B,C (bit A<<2|B<<1|C); parent row -1 = unbound slot, reads 0.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    const uint64_t* planes;
    int n;
    int64_t w64, n_cells;
    const int32_t* parent;
    const uint8_t* lut;
    int steps;
    int64_t w_lo, w_hi;       /* this worker's word range */
    int64_t* raw;             /* [n] private */
    int64_t* missing;         /* [2n+1] private */
    int64_t keyerror;
    int64_t lam_sum, lam_cnt; /* sum / count of min(lambda_x, lambda_n) + 1 over the scored runs (roofline bookkeeping) */
} job_t;

/* one synchronous update of 64 cells (importance_scores.py:60-105) */
static void step_word(const uint64_t* cur, uint64_t* nxt, int n, const int32_t* parent, const uint8_t* lut,
                      int forced, uint64_t forced_word) {
    for (int i = 0; i < n; ++i) {
        if (i == forced) { nxt[i] = forced_word; continue; }
        const int32_t* p = parent + 3 * i;
        const uint64_t a = p[0] >= 0 ? cur[p[0]] : 0, b = p[1] >= 0 ? cur[p[1]] : 0, c = p[2] >= 0 ? cur[p[2]] : 0;
        const uint8_t t = lut[i];
        uint64_t f = 0;
        for (int k = 0; k < 8; ++k)
            if ((t >> k) & 1) f |= ((k & 4) ? a : ~a) & ((k & 2) ? b : ~b) & ((k & 1) ? c : ~c);
        nxt[i] = f;
    }
}

/* full history of one run; then per lane the first (i, j<i) with equal states (:134-143) */
static uint64_t simulate(const uint64_t* start, int n, const int32_t* parent, const uint8_t* lut, int steps,
                         int forced, uint64_t forced_word, uint64_t valid, uint64_t* hist, uint8_t* j_of,
                         uint8_t* i_of) {
    for (int t = 0; t < steps; ++t)
        step_word(t == 0 ? start : hist + (size_t)(t - 1) * n, hist + (size_t)t * n, n, parent, lut, forced, forced_word);
    uint64_t open = valid, found = 0;
    for (int i = 1; i < steps && open; ++i) {
        const uint64_t* hi = hist + (size_t)i * n;
        for (int j = 0; j < i && open; ++j) {
            const uint64_t* hj = hist + (size_t)j * n;
            uint64_t eq = open;
            for (int k = 0; k < n && eq; ++k) eq &= ~(hi[k] ^ hj[k]);
            if (!eq) continue;
            for (uint64_t m = eq; m; m &= m - 1) {
                const int l = __builtin_ctzll(m);
                j_of[l] = (uint8_t)j;
                i_of[l] = (uint8_t)i;
            }
            found |= eq;
            open &= ~eq;
        }
    }
    return found;
}

/* sum over the lanes in `lanes` of the XOR count of the two attractors, each taken from its own
 * start, trimmed to the shorter (:325-340) */
static int64_t aligned_xor(const uint64_t* hx, const uint8_t* jx, const uint8_t* ix, const uint64_t* hn,
                           const uint8_t* jn, const uint8_t* in_, int n, uint64_t lanes, int64_t* lam_sum) {
    int64_t d = 0;
    for (uint64_t m = lanes; m; m &= m - 1) {
        const int l = __builtin_ctzll(m);
        const int len_x = ix[l] - jx[l] + 1, len_n = in_[l] - jn[l] + 1;
        const int len = len_x < len_n ? len_x : len_n;
        *lam_sum += len;
        for (int t = 0; t < len; ++t) {
            const uint64_t* a = hx + (size_t)(jx[l] + t) * n;
            const uint64_t* b = hn + (size_t)(jn[l] + t) * n;
            for (int k = 0; k < n; ++k) d += ((a[k] ^ b[k]) >> l) & 1;
        }
    }
    return d;
}

/* the same total, word-parallel: lanes grouped by (jx, jn, len) share their XOR words */
static int64_t aligned_xor_grouped(const uint64_t* hx, const uint8_t* jx, const uint8_t* ix, const uint64_t* hn,
                                   const uint8_t* jn, const uint8_t* in_, int n, uint64_t lanes, int64_t* lam_sum) {
    int64_t d = 0;
    uint64_t left = lanes;
    while (left) {
        const int l0 = __builtin_ctzll(left);
        const int len_x0 = ix[l0] - jx[l0] + 1, len_n0 = in_[l0] - jn[l0] + 1;
        const int len0 = len_x0 < len_n0 ? len_x0 : len_n0;
        uint64_t grp = 0;
        for (uint64_t m = left; m; m &= m - 1) {
            const int l = __builtin_ctzll(m);
            const int len_x = ix[l] - jx[l] + 1, len_n = in_[l] - jn[l] + 1;
            const int len = len_x < len_n ? len_x : len_n;
            if (jx[l] == jx[l0] && jn[l] == jn[l0] && len == len0) grp |= 1ull << l;
        }
        *lam_sum += (int64_t)len0 * __builtin_popcountll(grp);
        for (int t = 0; t < len0; ++t) {
            const uint64_t* a = hx + (size_t)(jx[l0] + t) * n;
            const uint64_t* b = hn + (size_t)(jn[l0] + t) * n;
            for (int k = 0; k < n; ++k) d += __builtin_popcountll((a[k] ^ b[k]) & grp);
        }
        left &= ~grp;
    }
    return d;
}

static void* worker(void* arg) {
    job_t* J = (job_t*)arg;
    const int n = J->n, steps = J->steps;
    uint64_t* start = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)n);
    uint64_t* hn = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)steps * n);
    uint64_t* ho = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)steps * n);
    uint64_t* hi = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)steps * n);
    uint8_t jn[64], in_[64], jo[64], io[64], ji[64], ii[64];
    for (int64_t w = J->w_lo; w < J->w_hi; ++w) {
        const int64_t left = J->n_cells - 64 * w;
        if (left <= 0) break;
        const uint64_t valid = left >= 64 ? ~0ull : ((1ull << left) - 1);
        for (int k = 0; k < n; ++k) start[k] = J->planes[(size_t)k * J->w64 + w];
        const uint64_t fn = simulate(start, n, J->parent, J->lut, steps, -1, 0, valid, hn, jn, in_);
        J->missing[0] += __builtin_popcountll(valid & ~fn);
        for (int k = 0; k < n; ++k) {
            const uint64_t fo = simulate(start, n, J->parent, J->lut, steps, k, 0ull, valid, ho, jo, io);
            const uint64_t fi = simulate(start, n, J->parent, J->lut, steps, k, ~0ull, valid, hi, ji, ii);
            J->missing[1 + 2 * k] += __builtin_popcountll(valid & ~fo);
            J->missing[2 + 2 * k] += __builtin_popcountll(valid & ~fi);
            const uint64_t both = fo & fi;
            J->keyerror += __builtin_popcountll(both & ~fn);   /* the reference raises KeyError there (:281) */
            const uint64_t lanes = both & fn;
            if (!lanes) continue;
            J->lam_cnt += 2 * __builtin_popcountll(lanes);
            J->raw[k] += aligned_xor_grouped(hi, ji, ii, hn, jn, in_, n, lanes, &J->lam_sum) +
                         aligned_xor_grouped(ho, jo, io, hn, jn, in_, n, lanes, &J->lam_sum);
        }
    }
    free(start); free(hn); free(ho); free(hi);
    return NULL;
}

/* raw[n], missing[2n+1]; returns the number of (cell, node) pairs that repeat under KO and KI but
 * not normally.  stats (optional, [2]): sum and count of the aligned attractor lengths that were scored. */
int64_t oracle_bs_importance_raw(const uint64_t* planes, int n, int64_t w64, int64_t n_cells, const int32_t* parent,
                                 const uint8_t* lut, int steps, int n_threads, int64_t* raw, int64_t* missing,
                                 int64_t* stats) {
    const int64_t n_words = (n_cells + 63) / 64;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_words) n_threads = n_words > 0 ? (int)n_words : 1;
    job_t* jobs = (job_t*)calloc((size_t)n_threads, sizeof(job_t));
    pthread_t* tids = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; ++t) {
        job_t* J = jobs + t;
        J->planes = planes; J->n = n; J->w64 = w64; J->n_cells = n_cells; J->parent = parent; J->lut = lut; J->steps = steps;
        J->w_lo = n_words * t / n_threads;
        J->w_hi = n_words * (t + 1) / n_threads;
        J->raw = (int64_t*)calloc((size_t)n, sizeof(int64_t));
        J->missing = (int64_t*)calloc((size_t)(2 * n + 1), sizeof(int64_t));
        if (n_threads > 1) pthread_create(tids + t, NULL, worker, J);
    }
    if (n_threads == 1) worker(jobs);
    memset(raw, 0, sizeof(int64_t) * (size_t)n);
    memset(missing, 0, sizeof(int64_t) * (size_t)(2 * n + 1));
    int64_t keyerror = 0, lam_sum = 0, lam_cnt = 0;
    for (int t = 0; t < n_threads; ++t) {
        job_t* J = jobs + t;
        if (n_threads > 1) pthread_join(tids[t], NULL);
        for (int k = 0; k < n; ++k) raw[k] += J->raw[k];
        for (int k = 0; k < 2 * n + 1; ++k) missing[k] += J->missing[k];
        keyerror += J->keyerror; lam_sum += J->lam_sum; lam_cnt += J->lam_cnt;
        free(J->raw); free(J->missing);
    }
    if (stats) { stats[0] = lam_sum; stats[1] = lam_cnt; }
    free(jobs); free(tids);
    return keyerror;
}

/* self-check hook for the tests: the per-lane and the grouped scoring agree */
int64_t oracle_bs_selfcheck_scoring(const uint64_t* hx, const uint8_t* jx, const uint8_t* ix, const uint64_t* hn,
                                    const uint8_t* jn, const uint8_t* in_, int n, uint64_t lanes) {
    int64_t s1 = 0, s2 = 0;
    const int64_t a = aligned_xor(hx, jx, ix, hn, jn, in_, n, lanes, &s1);
    const int64_t b = aligned_xor_grouped(hx, jx, ix, hn, jn, in_, n, lanes, &s2);
    return (a == b && s1 == s2) ? a : -1;
}
