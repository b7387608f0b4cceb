/*
 * scbonita_b200.h — C-ABI of the B200-native scBONITA2 hot path.
 *
 * The reference (Luminarada80/scBONITA2) is pure Python and has no FFI layer; its boundary
 * for this path is a set of Python call signatures.  Every entry point below names the
 * reference routine whose arithmetic it replaces (paths relative to the reference repo);
 * INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 * -----------
 *  - plain C, no torch types.  Unless a parameter is marked HOST, every pointer is a DEVICE
 *    pointer owned by the caller.  The library never frees caller memory.
 *  - work is enqueued on the caller's stream and is asynchronous (scb_host_* calls are the
 *    exception: they take HOST buffers, copy, compute, copy back and synchronise).
 *  - every function returns 0 on success or a negative SCB_ERR_* code; scb_last_error()
 *    returns a thread-local description.  There is NO CPU fallback anywhere.
 *  - "planes": uint32 cell bitplanes [n_rows][wpad]; bit b of word w of row r is cell 32*w+b
 *    of gene row r; bits at and beyond n_cells are ZERO; wpad is a multiple of 4 so that every
 *    row is 16-byte aligned (rows are read with 128-bit loads / bulk copies).
 *  - "lut": 8-bit truth table of a rule over its parent slots A,B,C; bit (A<<2|B<<1|C) is the
 *    rule's value (the lop3.b32 immediate convention).  An unused slot has parent row -1 and
 *    reads as constant 0.
 */
#ifndef SCBONITA_B200_H
#define SCBONITA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* scb_stream_t; /* == cudaStream_t */

#define SCB_OK 0
#define SCB_ERR_INVALID_ARG (-1)
#define SCB_ERR_CUDA (-2)
#define SCB_ERR_UNSUPPORTED (-3)
#define SCB_ERR_SCRATCH (-4)

#define SCB_ABI_VERSION 1
#define SCB_COUNTS_PER_GROUP 16 /* [minterm 0..7][target 0..1] */
#define SCB_MAX_SWEEP_NODES 413

int scb_abi_version(void);
const char* scb_last_error(void);
/* Number of SMs of the current device (used by callers to size persistent grids). */
int scb_device_sm_count(void);

/* ---- (1) bitplane repack ------------------------------------------------------------------
 * Replaces the dense/csr 0-1 matrix the reference carries around
 * (rule_inference.py:61-73 `binarized_matrix`, importance_scores.py:36-41 `.todense()`).
 * dense01: uint8 [n_rows][row_stride] (row_stride >= n_cells, bytes); any non-zero byte is 1. */
int scb_pack_bitplanes(const uint8_t* dense01, int n_rows, int64_t n_cells, int64_t row_stride,
                       uint32_t* planes, int64_t wpad, scb_stream_t stream);
/* Threshold + repack in one pass over a dense FLOAT matrix [n_rows][row_stride] (elements): bit =
 * value > threshold, as sklearn.preprocessing.binarize does on the csr matrix
 * (rule_inference.py:61-73; NaN compares false).  Skips the uint8 / csr intermediates. */
int scb_binarize_pack_f32(const float* dense, int n_rows, int64_t n_cells, int64_t row_stride,
                          float threshold, uint32_t* planes, int64_t wpad, scb_stream_t stream);
int scb_binarize_pack_f64(const double* dense, int n_rows, int64_t n_cells, int64_t row_stride,
                          double threshold, uint32_t* planes, int64_t wpad, scb_stream_t stream);
/* inverse, for round-trip tests and for rebuilding dense intermediates */
int scb_unpack_bitplanes(const uint32_t* planes, int n_rows, int64_t n_cells, int64_t wpad,
                         uint8_t* dense01, int64_t row_stride, scb_stream_t stream);

/* ---- (1b) majority chunking ---------------------------------------------------------------
 * RuleDetermination.chunk_data (rule_determination.py:181-233): chunk k of row r is 1 iff at
 * least `min_ones` of the cells perm[k*chunk_size .. (k+1)*chunk_size) are 1.  `perm` is the
 * host-generated np.random.seed(42) permutation; `min_ones` comes from the float-mean recipe
 * (SURVEY Q7), not from 2*count>=size.  out_planes: [n_rows][wpad_out], tail bits zero. */
int scb_chunk_majority(const uint32_t* planes, int n_rows, int64_t wpad, const int32_t* perm,
                       int64_t n_chunks, int chunk_size, int min_ones, uint32_t* out_planes,
                       int64_t wpad_out, scb_stream_t stream);

/* ---- (2) rule scoring ---------------------------------------------------------------------
 * RuleDetermination.calculate_error (rule_determination.py:509-582) evaluates ONE rule text on
 * every column with Python eval.  Here one pass over the cells produces, for every "group"
 * (a target row with its <=3 parent rows, i.e. one node), the 16 joint counts
 *     counts[g][2*k + t] = #cells with (A,B,C) == k and target == t ,
 * from which the mismatch count of ANY truth table follows exactly:
 *     difference(lut) = sum_k counts[g][2*k + (1 - lut_k)].
 * scratch: caller-provided device buffer of scb_rule_counts_scratch_bytes(n_rows, n_groups)
 * bytes (16-byte aligned). */
int64_t scb_rule_counts_scratch_bytes(int n_rows, int n_groups);
int scb_rule_counts(const uint32_t* planes, int n_rows, int64_t wpad, int64_t n_cells,
                    int n_groups, const int32_t* target_row /*[G]*/,
                    const int32_t* parent_row /*[G][3]*/, void* scratch, int64_t scratch_bytes,
                    int64_t* out_counts /*[G][16]*/, scb_stream_t stream);
/* Same with flags.  The kernel always leaves `scratch` zeroed behind it; a caller that zeroed the
 * scratch once (or reuses it after a successful call, in stream order) passes
 * SCB_COUNTS_SCRATCH_CLEAN and the call enqueues no memset. */
#define SCB_COUNTS_SCRATCH_CLEAN 1
/* SCB_COUNTS_CHAINED: the caller vouches that the kernel BEFORE this call in the stream wrote neither
 * the bitplanes nor the plan (e.g. it is the consumer of the previous pass): the counting kernel is
 * then launched as a programmatic dependent and its main loop overlaps that kernel's tail.  Without
 * the flag the launch is fully serialised. */
#define SCB_COUNTS_CHAINED 2
int scb_rule_counts_ex(const uint32_t* planes, int n_rows, int64_t wpad, int64_t n_cells,
                       int n_groups, const int32_t* target_row /*[G]*/,
                       const int32_t* parent_row /*[G][3]*/, void* scratch, int64_t scratch_bytes,
                       int64_t* out_counts /*[G][16]*/, int flags, scb_stream_t stream);

/* Repeated passes over the same groups (GA generations, the cell shards of one pathway): the
 * cell-independent part -- groups sorted by arity, shared-memory row offsets, which group also
 * counts its target row -- is built once into a caller-owned `plan` buffer
 * (scb_rule_counts_plan_bytes bytes, 16-byte aligned); scb_rule_counts_run then enqueues ONE
 * kernel per 1024 groups and nothing else.  scb_rule_counts == plan_build + run. */
int64_t scb_rule_counts_plan_bytes(int n_rows, int n_groups);
int scb_rule_counts_plan_build(int n_rows, int n_groups, const int32_t* target_row /*[G]*/,
                               const int32_t* parent_row /*[G][3]*/, void* plan, int64_t plan_bytes,
                               scb_stream_t stream);
int scb_rule_counts_run(const uint32_t* planes, int n_rows, int64_t wpad, int64_t n_cells,
                        int n_groups, const void* plan, void* scratch, int64_t scratch_bytes,
                        int64_t* out_counts /*[G][16]*/, int flags, scb_stream_t stream);

/* The GA-step pair (one pass over the cells + the fitness of a whole population, nothing else on
 * the critical path).  scb_rule_sums_run leaves the SUBSET SUMS of the groups (popcounts of the AND
 * of every subset of {A, B, C, target}) in `scratch` instead of inverting them into counts;
 * scb_population_fitness_sums consumes them -- it inverts them while it builds its tables and leaves
 * `scratch` zeroed -- so every scb_rule_sums_run must be followed, in stream order, by exactly one
 * scb_population_fitness_sums on the same plan / scratch (same results as scb_rule_counts_run +
 * scb_population_fitness).  scb_rule_sums_supported: one launch of the counting kernel and a fitness
 * kernel whose tables plus a copy of the sums fit in shared memory (about 660 groups); larger sets take
 * scb_rule_counts_run + scb_population_fitness.  n_untargeted_rows = rows that are no group's
 * target (informational).  With peer_regions (sharded cells) the kernel's last CTA pushes the sums to every
 * rank's exchange region and scb_population_fitness_sums(region != NULL) sums them over the ranks. */
int scb_rule_sums_supported(int n_rows, int n_groups, int n_untargeted_rows);
int scb_rule_sums_run(const uint32_t* planes, int n_rows, int64_t wpad, int64_t n_cells, int n_groups,
                      int n_untargeted_rows, const void* plan, void* scratch, int64_t scratch_bytes, int flags,
                      void* const* peer_regions /*[n_ranks] or NULL*/, int rank, int n_ranks, scb_stream_t stream);
int scb_population_fitness_sums(const void* plan, void* scratch, int n_rows, int n_groups, int64_t n_cells,
                                const void* local_region /*NULL: one GPU*/, int rank, int n_ranks,
                                int64_t n_individuals, const uint8_t* lut /*[P][G]*/,
                                const uint8_t* active /*[P][G] or NULL*/, int64_t* out_diff /*[P]*/,
                                int64_t* out_diff_pg /*[P][G] or NULL*/, scb_stream_t stream);

/* difference for T (group, lut) tasks: the batched calculate_error.  count == n_cells. */
int scb_rule_error_table(const int64_t* counts /*[G][16]*/, int n_groups, int64_t n_tasks,
                         const int32_t* task_group /*[T]*/, const uint8_t* task_lut /*[T]*/,
                         int64_t* out_diff /*[T]*/, scb_stream_t stream);

/* Brute-force variant with arbitrary per-task rows (no grouping): every task streams its
 * rows and evaluates the truth table word by word.  Same results; kept as an independent
 * in-library cross-check and as the integer-pipe reference point for the roofline. */
int scb_rule_error_table_direct(const uint32_t* planes, int n_rows, int64_t wpad,
                                int64_t n_cells, int64_t n_tasks,
                                const int32_t* target_row /*[T]*/,
                                const int32_t* parent_row /*[T][3]*/,
                                const uint8_t* lut /*[T]*/, int64_t* out_diff /*[T]*/,
                                scb_stream_t stream);

/* ---- (2b) GA population fitness -----------------------------------------------------------
 * Historic CustomDeap.fitness_calculation (code/__pycache__/updated_deap_class.cpython-36.pyc):
 * per individual, the summed mismatches of its selected rule for every node.
 * lut[p][g] is individual p's truth table for group g; active[p][g] (optional, may be NULL)
 * switches a (p,g) pair off.  out_diff[p] = sum_g difference; out_diff_pg (optional) keeps
 * the per-node terms. */
int scb_population_fitness(const int64_t* counts /*[G][16]*/, int n_groups, int64_t n_individuals,
                           const uint8_t* lut /*[P][G]*/, const uint8_t* active /*[P][G]|NULL*/,
                           int64_t* out_diff /*[P]*/, int64_t* out_diff_pg /*[P][G]|NULL*/,
                           scb_stream_t stream);

/* ---- (2c) multi-GPU: the joint counts of the cell shards summed over PEER MEMORY --------------
 * Cells shard across the GPUs of one box; the only exchange on the GA-fitness path is the sum of
 * the per-shard joint counts.  Instead of a library all-reduce between the two kernels, every rank
 * owns a small "region" mapped into all its peers (CUDA IPC, NVLink): scb_peer_push_counts (or the
 * last CTA of scb_rule_counts_run_push) stores the rank's counts into its slot of every region as
 * 64-bit words {epoch tag : count} -- an aligned 8-byte store is atomic, so every word validates
 * itself: no fence, no flag; scb_population_fitness_peers sums the slots of all ranks while it
 * builds its tables, spinning on a word until its tag is the current epoch.  Integer sums:
 * bit-identical to the single-GPU result at any GPU count.  Counts must be < 2^31 (they are:
 * at most 2^31-1 cells per shard).
 *   region : scb_peer_region_bytes(n_ranks, n_values) bytes from scb_peer_region_alloc (zeroed);
 *            `ipc_handle` (64 bytes) goes to the peers, which map it with scb_peer_region_open
 *   peer_regions : DEVICE array of n_ranks region pointers as seen from this rank (own region at
 *            index `rank`)
 * Both calls only enqueue work on `stream`; the same sequence must run on every rank. */
#define SCB_MAX_PEERS 8
#define SCB_IPC_HANDLE_BYTES 64
int64_t scb_peer_region_bytes(int n_ranks, int64_t n_values);
int scb_peer_region_alloc(int64_t bytes, void** region, unsigned char* ipc_handle /*HOST [64]*/);
int scb_peer_region_open(const unsigned char* ipc_handle /*HOST [64]*/, void** region);
int scb_peer_region_close(void* region);
int scb_peer_region_free(void* region);
int scb_peer_push_counts(const int64_t* counts /*[n_values]*/, int64_t n_values, void* local_region,
                         void* const* peer_regions /*DEVICE [n_ranks]*/, int rank, int n_ranks,
                         scb_stream_t stream);
/* scb_rule_counts_run with the push folded into the kernel's last CTA: no kernel of its own */
int scb_rule_counts_run_push(const uint32_t* planes, int n_rows, int64_t wpad, int64_t n_cells,
                             int n_groups, const void* plan, void* scratch, int64_t scratch_bytes,
                             int64_t* out_counts /*[G][16]*/, int flags,
                             void* const* peer_regions /*DEVICE [n_ranks]*/, int rank, int n_ranks,
                             scb_stream_t stream);
int scb_population_fitness_peers(const void* local_region, int rank, int n_ranks, int n_groups,
                                 int64_t n_individuals, const uint8_t* lut /*[P][G]*/,
                                 const uint8_t* active /*[P][G]|NULL*/, int64_t* out_diff /*[P]*/,
                                 int64_t* out_diff_pg /*[P][G]|NULL*/, scb_stream_t stream);
/* 1 if a kernel gave up waiting for a peer since the last call (the flag is cleared by the call);
 * synchronises the device.  Such a kernel also writes INT64_MIN into every result of its launch, so a
 * timed-out step can never pass for a valid one.  scb_peer_set_timeout: how long a kernel waits for a
 * peer's word (milliseconds, 0 = the default of 30 s; SCB_PEER_TIMEOUT_MS sets the process default). */
int scb_peer_error(const void* local_region, int* out_error /*HOST*/);
int scb_peer_set_timeout(void* local_region, unsigned int timeout_ms);

/* Multi-rule individuals of the historic GA: an individual is a 0/1 string over the concatenated
 * candidate lists of all nodes; out_diff[p] = sum of task_diff[j] over its set bits, out_nsel[p] =
 * number of set bits (the GA's complexity penalty).  task_diff is scb_rule_error_table's output. */
int scb_bitstring_fitness(const int64_t* task_diff /*[T]*/, int64_t n_tasks, int64_t n_individuals,
                          const uint8_t* bits /*[P][T]*/, int64_t* out_diff /*[P]*/,
                          int32_t* out_nsel /*[P]*/, scb_stream_t stream);

/* ---- (3) knock-out / knock-in importance sweep --------------------------------------------
 * CalculateImportanceScore.calculate_importance_scores (importance_scores.py:47-143,195-340):
 * synchronous update of the whole network for `steps` steps under 2N+1 conditions (normal,
 * node i forced 0, node i forced 1), first-repeat attractor per cell, XOR distance of the
 * aligned attractors to the normal one.
 *   net_parent[n][3] : rows read by node n's slots A,B,C (-1 unused)
 *   net_lut[n]       : truth table of node n's rule
 *   out_raw[n]       : integer XOR total of node n (KO + KI), before max-normalisation
 *   out_missing[c]   : cells without a repeat within `steps`; c = 0 normal, 1+2n KO n, 2+2n KI n
 *   out_keyerror     : #(cell,node) pairs that repeat under KO and KI but NOT under normal
 *                      (the reference raises KeyError there, importance_scores.py:281)
 * All outputs are ACCUMULATED INTO (+=), so shards of cells can share one buffer; zero them
 * first.  scratch: scb_ko_ki_sweep_scratch_bytes(...) bytes, 16-byte aligned, private to the call
 * while it runs (normal-run states node-major and cell-major, two per-node lists of deferred cells:
 * about (2 * steps * N / 8 + 8 * N) bytes per cell, 4.5 GB for 200 nodes x 1M cells x 50 steps).
 * Environment knobs for experiments and tests only (every setting gives the same integers): SCB_SWEEP_FAST_STEPS,
 * SCB_SWEEP_RING, SCB_SWEEP_SLOTS, SCB_SWEEP_SORT, SCB_SWEEP_STOP_OPEN / _STOP_STALL / _RING_STALL,
 * SCB_SWEEP_KOKI=old, SCB_SWEEP_RECHECK=old, SCB_SWEEP_CLEANUP=old (the round-1 kernels for a pass),
 * SCB_SWEEP_SINGLES=0 (open cells always travel as KO/KI pairs), SCB_SWEEP_BRENT_MID=0. */
int64_t scb_ko_ki_sweep_scratch_bytes(int n_nodes, int64_t wpad, int steps);
int scb_ko_ki_sweep(const uint32_t* planes, int n_nodes, int64_t wpad, int64_t n_cells,
                    const int32_t* net_parent /*[N][3]*/, const uint8_t* net_lut /*[N]*/,
                    int steps, void* scratch, int64_t scratch_bytes, int64_t* out_raw /*[N]*/,
                    int64_t* out_missing /*[2N+1]*/, int64_t* out_keyerror /*[1]*/,
                    scb_stream_t stream);

/* ---- (3b) per-cell trajectories -----------------------------------------------------------
 * attractor_analysis.vectorized_run_simulation (attractor_analysis.py:43-99) for M cells:
 * out[m][n][t] = state of node n of cell cell_idx[m] after t+1 updates, t < steps. */
int scb_trajectories(const uint32_t* planes, int n_nodes, int64_t wpad, int64_t n_cells,
                     const int64_t* cell_idx /*[M]*/, int64_t n_selected,
                     const int32_t* net_parent /*[N][3]*/, const uint8_t* net_lut /*[N]*/,
                     int steps, uint8_t* out /*[M][N][steps]*/, scb_stream_t stream);

/* Same with one node forced to 0 (knock-out) or 1 (knock-in) from the first update on
 * (CalculateImportanceScore.vectorized_run_simulation, importance_scores.py:47-111);
 * forced_node = -1 forces nothing. */
int scb_trajectories_forced(const uint32_t* planes, int n_nodes, int64_t wpad, int64_t n_cells,
                            const int64_t* cell_idx /*[M]*/, int64_t n_selected,
                            const int32_t* net_parent /*[N][3]*/, const uint8_t* net_lut /*[N]*/,
                            int steps, int forced_node, int forced_value,
                            uint8_t* out /*[M][N][steps]*/, scb_stream_t stream);

/* ---- (3c) pairwise trajectory distances ---------------------------------
 * compute_dtw_distance_pair / compute_dtw_distances (attractor_analysis.py:304-377): per pair of
 * cells the sum over the genes of the squared difference of the two down-sampled 0/1 series if it
 * is below `threshold` (EUCLIDEAN_THRESHOLD = 5, :41), else fastdtw(radius, squared distance).
 * A down-sampled series is a pattern of `bits` points (10 for 20 steps, every 2nd), so the per-gene
 * distance is a table over two patterns.  fastdtw 0.3.4 is a third-party package absent from the
 * reference tree: its algorithm is restated, parity of that branch is unpinned. */
/* HOST: table[2^bits][2^bits] bytes; host code only, no GPU needed */
int scb_dtw_table_build(int bits, int radius, int threshold, uint8_t* host_table);
/* traj [M][N][steps] bytes (scb_trajectories) -> patterns [M][N]: bit k = traj[..][k * factor] */
int scb_trajectory_patterns(const uint8_t* traj, int64_t n_cells, int n_genes, int steps, int factor,
                            uint16_t* out /*[M][N]*/, scb_stream_t stream);
/* out[a][b] = sum_g table[patterns[a][g]][patterns[b][g]] (symmetric, diagonal included) */
int scb_pattern_pair_distances(const uint16_t* patterns /*[M][N]*/, int64_t n_cells, int n_genes,
                               const uint8_t* table /*DEVICE [2^bits][2^bits]*/, int bits,
                               int32_t* out /*[M][M]*/, scb_stream_t stream);

/* ---- host-buffer entry points (the end-to-end path; all pointers HOST) ---------------------
 * One workspace owns the device buffers of one (device, stream); not thread-safe. */
typedef struct scb_workspace scb_workspace;
int scb_workspace_create(scb_workspace** out);
void scb_workspace_destroy(scb_workspace* ws);

/* dense 0/1 matrix (HOST uint8 [n_rows][n_cells]) -> bitplanes resident in the workspace */
int scb_host_load_dense(scb_workspace* ws, const uint8_t* dense01, int n_rows, int64_t n_cells);
/* already-packed planes (HOST uint32 [n_rows][n_words], n_words = ceil(n_cells/32)) */
int scb_host_load_planes(scb_workspace* ws, const uint32_t* planes, int n_rows, int64_t n_cells);
/* counts for G groups on the resident planes, kept resident; optional HOST copy out */
int scb_host_rule_counts(scb_workspace* ws, int n_groups, const int32_t* target_row,
                         const int32_t* parent_row, int64_t* out_counts /*HOST [G][16]|NULL*/);
int scb_host_rule_error_table(scb_workspace* ws, int64_t n_tasks, const int32_t* task_group,
                              const uint8_t* task_lut, int64_t* out_diff /*HOST [T]*/);
int scb_host_population_fitness(scb_workspace* ws, int64_t n_individuals, const uint8_t* lut,
                                const uint8_t* active, int64_t* out_diff /*HOST [P]*/,
                                int64_t* out_diff_pg /*HOST [P][G]|NULL*/);
int scb_host_ko_ki_sweep(scb_workspace* ws, const int32_t* net_parent, const uint8_t* net_lut,
                         int steps, int64_t* out_raw /*HOST [N]*/,
                         int64_t* out_missing /*HOST [2N+1]*/, int64_t* out_keyerror /*HOST [1]*/);

#ifdef __cplusplus
}
#endif
#endif /* SCBONITA_B200_H */
