// C-ABI plumbing: error reporting, device queries and the host-buffer ("end-to-end") entry
// points that own their device buffers through an scb_workspace.
#include <string.h>

#include <new>

#include "scb_common.cuh"

namespace scb {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t err, const char* what) {
    set_error("CUDA error in %s: %s", what, cudaGetErrorString(err));
    return SCB_ERR_CUDA;
}

int sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace scb

using namespace scb;

extern "C" int scb_abi_version(void) { return SCB_ABI_VERSION; }
extern "C" const char* scb_last_error(void) { return g_error; }
extern "C" int scb_device_sm_count(void) { return sm_count(); }

// ------------------------------------------------------------------------------------------
// workspace: device buffers that grow on demand, one stream
// ------------------------------------------------------------------------------------------
struct DevBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return SCB_OK;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
        cap = want;
        return SCB_OK;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
};

struct scb_workspace {
    cudaStream_t stream = nullptr;
    DevBuf dense, planes, scratch, counts, meta_a, meta_b, meta_c, out_a, out_b, sweep;
    int n_rows = 0;
    int64_t n_cells = 0, wpad = 0;
    int n_groups = 0;
};

static int64_t padded_words(int64_t n_cells) {
    int64_t w = (n_cells + 31) / 32;
    w = (w + 3) / 4 * 4;
    return w < 4 ? 4 : w;
}

extern "C" int scb_workspace_create(scb_workspace** out) {
    SCB_REQUIRE(out, "workspace_create: null out pointer");
    scb_workspace* ws = new (std::nothrow) scb_workspace();
    SCB_REQUIRE(ws, "workspace_create: out of host memory");
    cudaError_t e = cudaStreamCreateWithFlags(&ws->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ws;
        return cuda_fail(e, "cudaStreamCreate");
    }
    *out = ws;
    return SCB_OK;
}

extern "C" void scb_workspace_destroy(scb_workspace* ws) {
    if (!ws) return;
    cudaStreamSynchronize(ws->stream);
    DevBuf* all[] = {&ws->dense, &ws->planes, &ws->scratch, &ws->counts, &ws->meta_a,
                     &ws->meta_b, &ws->meta_c, &ws->out_a, &ws->out_b, &ws->sweep};
    for (DevBuf* b : all) b->release();
    cudaStreamDestroy(ws->stream);
    delete ws;
}

#define WS_TRY(expr)                 \
    do {                             \
        int _rc = (expr);            \
        if (_rc != SCB_OK) return _rc; \
    } while (0)

extern "C" int scb_host_load_dense(scb_workspace* ws, const uint8_t* dense01, int n_rows,
                                   int64_t n_cells) {
    SCB_REQUIRE(ws && dense01 && n_rows > 0 && n_cells > 0, "host_load_dense: bad arguments");
    int64_t wpad = padded_words(n_cells);
    // device copy keeps a 16-byte aligned row stride so the packer takes its 128-bit path
    int64_t stride = (n_cells + 15) / 16 * 16;
    WS_TRY(ws->dense.reserve((size_t)n_rows * stride));
    WS_TRY(ws->planes.reserve((size_t)n_rows * wpad * 4));
    SCB_CUDA(cudaMemcpy2DAsync(ws->dense.ptr, stride, dense01, n_cells, n_cells, n_rows,
                               cudaMemcpyHostToDevice, ws->stream));
    WS_TRY(scb_pack_bitplanes(static_cast<const uint8_t*>(ws->dense.ptr), n_rows, n_cells, stride,
                              static_cast<uint32_t*>(ws->planes.ptr), wpad,
                              reinterpret_cast<scb_stream_t>(ws->stream)));
    ws->n_rows = n_rows;
    ws->n_cells = n_cells;
    ws->wpad = wpad;
    ws->n_groups = 0;
    SCB_CUDA(cudaStreamSynchronize(ws->stream));
    return SCB_OK;
}

extern "C" int scb_host_load_planes(scb_workspace* ws, const uint32_t* planes, int n_rows,
                                    int64_t n_cells) {
    SCB_REQUIRE(ws && planes && n_rows > 0 && n_cells > 0, "host_load_planes: bad arguments");
    int64_t wpad = padded_words(n_cells);
    int64_t n_words = (n_cells + 31) / 32;
    WS_TRY(ws->planes.reserve((size_t)n_rows * wpad * 4));
    if (wpad != n_words) SCB_CUDA(cudaMemsetAsync(ws->planes.ptr, 0, (size_t)n_rows * wpad * 4, ws->stream));
    SCB_CUDA(cudaMemcpy2DAsync(ws->planes.ptr, wpad * 4, planes, n_words * 4, n_words * 4, n_rows,
                               cudaMemcpyHostToDevice, ws->stream));
    ws->n_rows = n_rows;
    ws->n_cells = n_cells;
    ws->wpad = wpad;
    ws->n_groups = 0;
    SCB_CUDA(cudaStreamSynchronize(ws->stream));
    return SCB_OK;
}

extern "C" int scb_host_rule_counts(scb_workspace* ws, int n_groups, const int32_t* target_row,
                                    const int32_t* parent_row, int64_t* out_counts) {
    SCB_REQUIRE(ws && ws->n_rows > 0, "host_rule_counts: load a matrix first");
    SCB_REQUIRE(n_groups > 0 && target_row && parent_row, "host_rule_counts: bad arguments");
    for (int g = 0; g < n_groups; ++g) {
        SCB_REQUIRE(target_row[g] >= 0 && target_row[g] < ws->n_rows, "host_rule_counts: target row %d of group %d out of range", target_row[g], g);
        for (int k = 0; k < 3; ++k)
            SCB_REQUIRE(parent_row[3 * g + k] >= -1 && parent_row[3 * g + k] < ws->n_rows, "host_rule_counts: parent row %d of group %d out of range", parent_row[3 * g + k], g);
    }
    auto st = reinterpret_cast<scb_stream_t>(ws->stream);
    int64_t sbytes = scb_rule_counts_scratch_bytes(ws->n_rows, n_groups);
    WS_TRY(ws->scratch.reserve((size_t)sbytes));
    WS_TRY(ws->counts.reserve((size_t)n_groups * 16 * 8));
    WS_TRY(ws->meta_a.reserve((size_t)n_groups * 4));
    WS_TRY(ws->meta_b.reserve((size_t)n_groups * 12));
    SCB_CUDA(cudaMemcpyAsync(ws->meta_a.ptr, target_row, (size_t)n_groups * 4, cudaMemcpyHostToDevice, ws->stream));
    SCB_CUDA(cudaMemcpyAsync(ws->meta_b.ptr, parent_row, (size_t)n_groups * 12, cudaMemcpyHostToDevice, ws->stream));
    WS_TRY(scb_rule_counts(static_cast<const uint32_t*>(ws->planes.ptr), ws->n_rows, ws->wpad,
                           ws->n_cells, n_groups, static_cast<const int32_t*>(ws->meta_a.ptr),
                           static_cast<const int32_t*>(ws->meta_b.ptr), ws->scratch.ptr, sbytes,
                           static_cast<int64_t*>(ws->counts.ptr), st));
    ws->n_groups = n_groups;
    if (out_counts)
        SCB_CUDA(cudaMemcpyAsync(out_counts, ws->counts.ptr, (size_t)n_groups * 16 * 8, cudaMemcpyDeviceToHost, ws->stream));
    SCB_CUDA(cudaStreamSynchronize(ws->stream));
    return SCB_OK;
}

extern "C" int scb_host_rule_error_table(scb_workspace* ws, int64_t n_tasks,
                                         const int32_t* task_group, const uint8_t* task_lut,
                                         int64_t* out_diff) {
    SCB_REQUIRE(ws && ws->n_groups > 0, "host_rule_error_table: compute counts first");
    SCB_REQUIRE(n_tasks >= 0 && (n_tasks == 0 || (task_group && task_lut && out_diff)), "host_rule_error_table: bad arguments");
    if (n_tasks == 0) return SCB_OK;
    for (int64_t t = 0; t < n_tasks; ++t)
        SCB_REQUIRE(task_group[t] >= 0 && task_group[t] < ws->n_groups, "host_rule_error_table: task %lld names group %d of %d", (long long)t, task_group[t], ws->n_groups);
    WS_TRY(ws->meta_a.reserve((size_t)n_tasks * 4));
    WS_TRY(ws->meta_c.reserve((size_t)n_tasks));
    WS_TRY(ws->out_a.reserve((size_t)n_tasks * 8));
    SCB_CUDA(cudaMemcpyAsync(ws->meta_a.ptr, task_group, (size_t)n_tasks * 4, cudaMemcpyHostToDevice, ws->stream));
    SCB_CUDA(cudaMemcpyAsync(ws->meta_c.ptr, task_lut, (size_t)n_tasks, cudaMemcpyHostToDevice, ws->stream));
    WS_TRY(scb_rule_error_table(static_cast<const int64_t*>(ws->counts.ptr), ws->n_groups, n_tasks,
                                static_cast<const int32_t*>(ws->meta_a.ptr),
                                static_cast<const uint8_t*>(ws->meta_c.ptr),
                                static_cast<int64_t*>(ws->out_a.ptr),
                                reinterpret_cast<scb_stream_t>(ws->stream)));
    SCB_CUDA(cudaMemcpyAsync(out_diff, ws->out_a.ptr, (size_t)n_tasks * 8, cudaMemcpyDeviceToHost, ws->stream));
    SCB_CUDA(cudaStreamSynchronize(ws->stream));
    return SCB_OK;
}

extern "C" int scb_host_population_fitness(scb_workspace* ws, int64_t n_individuals,
                                           const uint8_t* lut, const uint8_t* active,
                                           int64_t* out_diff, int64_t* out_diff_pg) {
    SCB_REQUIRE(ws && ws->n_groups > 0, "host_population_fitness: compute counts first");
    SCB_REQUIRE(n_individuals > 0 && lut && out_diff, "host_population_fitness: bad arguments");
    size_t pg = (size_t)n_individuals * ws->n_groups;
    WS_TRY(ws->meta_c.reserve(pg));
    WS_TRY(ws->out_a.reserve((size_t)n_individuals * 8));
    if (active) WS_TRY(ws->meta_a.reserve(pg));
    if (out_diff_pg) WS_TRY(ws->out_b.reserve(pg * 8));
    SCB_CUDA(cudaMemcpyAsync(ws->meta_c.ptr, lut, pg, cudaMemcpyHostToDevice, ws->stream));
    if (active) SCB_CUDA(cudaMemcpyAsync(ws->meta_a.ptr, active, pg, cudaMemcpyHostToDevice, ws->stream));
    WS_TRY(scb_population_fitness(static_cast<const int64_t*>(ws->counts.ptr), ws->n_groups,
                                  n_individuals, static_cast<const uint8_t*>(ws->meta_c.ptr),
                                  active ? static_cast<const uint8_t*>(ws->meta_a.ptr) : nullptr,
                                  static_cast<int64_t*>(ws->out_a.ptr),
                                  out_diff_pg ? static_cast<int64_t*>(ws->out_b.ptr) : nullptr,
                                  reinterpret_cast<scb_stream_t>(ws->stream)));
    SCB_CUDA(cudaMemcpyAsync(out_diff, ws->out_a.ptr, (size_t)n_individuals * 8, cudaMemcpyDeviceToHost, ws->stream));
    if (out_diff_pg) SCB_CUDA(cudaMemcpyAsync(out_diff_pg, ws->out_b.ptr, pg * 8, cudaMemcpyDeviceToHost, ws->stream));
    SCB_CUDA(cudaStreamSynchronize(ws->stream));
    return SCB_OK;
}

extern "C" int scb_host_ko_ki_sweep(scb_workspace* ws, const int32_t* net_parent,
                                    const uint8_t* net_lut, int steps, int64_t* out_raw,
                                    int64_t* out_missing, int64_t* out_keyerror) {
    SCB_REQUIRE(ws && ws->n_rows > 0, "host_ko_ki_sweep: load a matrix first");
    SCB_REQUIRE(net_parent && net_lut && out_raw && out_missing && out_keyerror, "host_ko_ki_sweep: null pointer");
    const int N = ws->n_rows;
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < 3; ++k)
            SCB_REQUIRE(net_parent[3 * n + k] >= -1 && net_parent[3 * n + k] < N, "host_ko_ki_sweep: parent row %d of node %d out of range", net_parent[3 * n + k], n);
    int64_t sbytes = scb_ko_ki_sweep_scratch_bytes(N, ws->wpad, steps);
    SCB_REQUIRE(sbytes > 0, "host_ko_ki_sweep: bad steps");
    WS_TRY(ws->sweep.reserve((size_t)sbytes));
    WS_TRY(ws->meta_b.reserve((size_t)N * 12));
    WS_TRY(ws->meta_c.reserve((size_t)N));
    size_t n_out = (size_t)N + (2 * (size_t)N + 1) + 1;
    WS_TRY(ws->out_a.reserve(n_out * 8));
    SCB_CUDA(cudaMemcpyAsync(ws->meta_b.ptr, net_parent, (size_t)N * 12, cudaMemcpyHostToDevice, ws->stream));
    SCB_CUDA(cudaMemcpyAsync(ws->meta_c.ptr, net_lut, (size_t)N, cudaMemcpyHostToDevice, ws->stream));
    SCB_CUDA(cudaMemsetAsync(ws->out_a.ptr, 0, n_out * 8, ws->stream));
    int64_t* d_raw = static_cast<int64_t*>(ws->out_a.ptr);
    int64_t* d_missing = d_raw + N;
    int64_t* d_key = d_missing + (2 * N + 1);
    WS_TRY(scb_ko_ki_sweep(static_cast<const uint32_t*>(ws->planes.ptr), N, ws->wpad, ws->n_cells,
                           static_cast<const int32_t*>(ws->meta_b.ptr),
                           static_cast<const uint8_t*>(ws->meta_c.ptr), steps, ws->sweep.ptr, sbytes,
                           d_raw, d_missing, d_key, reinterpret_cast<scb_stream_t>(ws->stream)));
    SCB_CUDA(cudaMemcpyAsync(out_raw, d_raw, (size_t)N * 8, cudaMemcpyDeviceToHost, ws->stream));
    SCB_CUDA(cudaMemcpyAsync(out_missing, d_missing, (2 * (size_t)N + 1) * 8, cudaMemcpyDeviceToHost, ws->stream));
    SCB_CUDA(cudaMemcpyAsync(out_keyerror, d_key, 8, cudaMemcpyDeviceToHost, ws->stream));
    SCB_CUDA(cudaStreamSynchronize(ws->stream));
    return SCB_OK;
}
