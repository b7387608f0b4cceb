// Bitplane repack, unpack and majority chunking (HBM-bound byte/bit shuffling).
//
// Replaces the dense/csr 0-1 matrix of the reference (rule_inference.py:61-73) by uint32 cell
// bitplanes, and RuleDetermination.chunk_data (rule_determination.py:181-233).
#include "scb_common.cuh"

namespace scb {

// ---------------------------------------------------------------------------------------
// pack: one warp produces 32 consecutive words (1024 cells) of one row per iteration.
// Fast path (16-byte aligned rows): every lane reads 2 x 16 bytes = its own 32 cells.
// Generic path: 32 coalesced byte loads + ballot.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t nz_bits16(uint4 v) {
    // 16 bytes -> 16 bits (bit i = byte i != 0)
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t nz = __vcmpne4(w[i], 0u) & 0x01010101u;      // 0/1 per byte
        uint32_t nib = (nz * 0x01020408u) >> 24;              // flags at bits 0,8,16,24 -> bits 24..27
        out |= (nib & 0xFu) << (4 * i);
    }
    return out;
}

__global__ void __launch_bounds__(256) pack_aligned_kernel(const uint8_t* __restrict__ dense,
                                                           int n_rows, int64_t n_cells,
                                                           int64_t row_stride,
                                                           uint32_t* __restrict__ planes,
                                                           int64_t wpad) {
    // thread <-> (row, word); consecutive threads take consecutive words of a row
    int64_t words_per_row = wpad;
    int64_t total = (int64_t)n_rows * words_per_row;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t row = idx / words_per_row;
        int64_t w = idx - row * words_per_row;
        int64_t first = w * 32;
        uint32_t bits = 0;
        if (first < n_cells) {
            const uint8_t* src = dense + row * row_stride + first;
            if (first + 32 <= n_cells) {
                const uint4* p = reinterpret_cast<const uint4*>(src);
                uint4 lo = __ldg(p), hi = __ldg(p + 1);
                bits = nz_bits16(lo) | (nz_bits16(hi) << 16);
            } else {
                int left = (int)(n_cells - first);
                for (int b = 0; b < left; ++b) bits |= (uint32_t)(src[b] != 0) << b;
            }
        }
        planes[row * wpad + w] = bits;
    }
}

__global__ void __launch_bounds__(256) pack_generic_kernel(const uint8_t* __restrict__ dense,
                                                           int n_rows, int64_t n_cells,
                                                           int64_t row_stride,
                                                           uint32_t* __restrict__ planes,
                                                           int64_t wpad) {
    // warp <-> (row, block of 32 words); lane i of iteration j reads cell base + 32*j + i
    const int lane = threadIdx.x & 31;
    int64_t blocks_per_row = (wpad + 31) / 32;
    int64_t total = (int64_t)n_rows * blocks_per_row;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t item = warp; item < total; item += n_warps) {
        int64_t row = item / blocks_per_row;
        int64_t wb = (item - row * blocks_per_row) * 32;
        const uint8_t* src = dense + row * row_stride;
        uint32_t mine = 0;
        for (int j = 0; j < 32; ++j) {
            int64_t cell = (wb + j) * 32 + lane;
            bool one = cell < n_cells && src[cell] != 0;
            uint32_t word = __ballot_sync(0xFFFFFFFFu, one);
            if (lane == j) mine = word;
        }
        if (wb + lane < wpad) planes[row * wpad + wb + lane] = mine;
    }
}

__global__ void __launch_bounds__(256) unpack_kernel(const uint32_t* __restrict__ planes,
                                                     int n_rows, int64_t n_cells, int64_t wpad,
                                                     uint8_t* __restrict__ dense,
                                                     int64_t row_stride) {
    int64_t total = (int64_t)n_rows * n_cells;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t row = idx / n_cells;
        int64_t cell = idx - row * n_cells;
        uint32_t w = __ldg(planes + row * wpad + (cell >> 5));
        dense[row * row_stride + cell] = (w >> (cell & 31)) & 1u;
    }
}

// ---------------------------------------------------------------------------------------
// chunk majority: warp <-> (row, 32 consecutive chunks); each lane gathers its chunk's cells
// through the permutation and votes; ballot assembles the output word.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) chunk_majority_kernel(
    const uint32_t* __restrict__ planes, int n_rows, int64_t wpad, const int32_t* __restrict__ perm,
    int64_t n_chunks, int chunk_size, int min_ones, uint32_t* __restrict__ out, int64_t wpad_out) {
    const int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int64_t total = (int64_t)n_rows * wpad_out;
    for (int64_t item = warp; item < total; item += n_warps) {
        int64_t row = item / wpad_out;
        int64_t ow = item - row * wpad_out;
        int64_t chunk = ow * 32 + lane;
        bool vote = false;
        if (chunk < n_chunks) {
            const uint32_t* src = planes + row * wpad;
            const int32_t* cells = perm + chunk * chunk_size;
            int ones = 0;
            for (int i = 0; i < chunk_size; ++i) {
                int32_t c = __ldg(cells + i);
                ones += (__ldg(src + (c >> 5)) >> (c & 31)) & 1u;
            }
            vote = ones >= min_ones;
        }
        uint32_t word = __ballot_sync(0xFFFFFFFFu, vote);
        if (lane == 0) out[row * wpad_out + ow] = word;
    }
}

static int grid_for(int64_t threads_needed, int block) {
    int64_t blocks = (threads_needed + block - 1) / block;
    int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// ---------------------------------------------------------------------------------------
// threshold + pack in one pass over a dense FLOAT matrix (the reference builds a csr matrix and
// calls sklearn binarize, rule_inference.py:61-73: stored values > threshold become 1).
// warp <-> (row, block of 32 words); lane i of iteration j reads cell base + 32*j + i, so every
// load instruction of the warp is one contiguous 128/256-byte segment.  NaN > t is false.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) binarize_pack_kernel(const T* __restrict__ dense, int n_rows,
                                                            int64_t n_cells, int64_t row_stride, T threshold,
                                                            uint32_t* __restrict__ planes, int64_t wpad) {
    const int lane = threadIdx.x & 31;
    int64_t blocks_per_row = (wpad + 31) / 32;
    int64_t total = (int64_t)n_rows * blocks_per_row;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t item = warp; item < total; item += n_warps) {
        int64_t row = item / blocks_per_row;
        int64_t wb = (item - row * blocks_per_row) * 32;
        const T* src = dense + row * row_stride;
        uint32_t mine = 0;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
            int64_t cell = (wb + j) * 32 + lane;
            bool one = cell < n_cells && __ldg(src + cell) > threshold;
            uint32_t word = __ballot_sync(0xFFFFFFFFu, one);
            if (lane == j) mine = word;
        }
        if (wb + lane < wpad) planes[row * wpad + wb + lane] = mine;
    }
}

}  // namespace scb

using namespace scb;

extern "C" int scb_pack_bitplanes(const uint8_t* dense01, int n_rows, int64_t n_cells,
                                  int64_t row_stride, uint32_t* planes, int64_t wpad,
                                  scb_stream_t stream) {
    SCB_REQUIRE(n_rows >= 0 && n_cells >= 0, "pack_bitplanes: negative shape");
    SCB_REQUIRE(wpad % 4 == 0 && wpad * 32 >= n_cells, "pack_bitplanes: wpad=%lld must be a multiple of 4 covering %lld cells", (long long)wpad, (long long)n_cells);
    SCB_REQUIRE(row_stride >= n_cells, "pack_bitplanes: row_stride < n_cells");
    if (n_rows == 0 || wpad == 0) return SCB_OK;
    SCB_REQUIRE(dense01 && planes, "pack_bitplanes: null pointer");
    bool aligned = (reinterpret_cast<uintptr_t>(dense01) % 16 == 0) && (row_stride % 16 == 0);
    int64_t total = (int64_t)n_rows * wpad;
    if (aligned) {
        pack_aligned_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
            dense01, n_rows, n_cells, row_stride, planes, wpad);
    } else {
        int64_t warps = (int64_t)n_rows * ((wpad + 31) / 32);
        pack_generic_kernel<<<grid_for(warps * 32, 256), 256, 0, as_stream(stream)>>>(
            dense01, n_rows, n_cells, row_stride, planes, wpad);
    }
    SCB_LAUNCH_CHECK("pack_bitplanes");
    return SCB_OK;
}

extern "C" int scb_unpack_bitplanes(const uint32_t* planes, int n_rows, int64_t n_cells,
                                    int64_t wpad, uint8_t* dense01, int64_t row_stride,
                                    scb_stream_t stream) {
    SCB_REQUIRE(n_rows >= 0 && n_cells >= 0, "unpack_bitplanes: negative shape");
    SCB_REQUIRE(wpad * 32 >= n_cells && row_stride >= n_cells, "unpack_bitplanes: bad strides");
    if (n_rows == 0 || n_cells == 0) return SCB_OK;
    SCB_REQUIRE(dense01 && planes, "unpack_bitplanes: null pointer");
    unpack_kernel<<<grid_for((int64_t)n_rows * n_cells, 256), 256, 0, as_stream(stream)>>>(
        planes, n_rows, n_cells, wpad, dense01, row_stride);
    SCB_LAUNCH_CHECK("unpack_bitplanes");
    return SCB_OK;
}

extern "C" int scb_chunk_majority(const uint32_t* planes, int n_rows, int64_t wpad,
                                  const int32_t* perm, int64_t n_chunks, int chunk_size,
                                  int min_ones, uint32_t* out_planes, int64_t wpad_out,
                                  scb_stream_t stream) {
    SCB_REQUIRE(n_rows >= 0 && n_chunks >= 0 && chunk_size >= 1, "chunk_majority: bad shape");
    SCB_REQUIRE(wpad_out % 4 == 0 && wpad_out * 32 >= n_chunks, "chunk_majority: wpad_out must be a multiple of 4 covering n_chunks");
    SCB_REQUIRE(n_chunks * chunk_size <= wpad * 32, "chunk_majority: chunks exceed the cells");
    if (n_rows == 0 || wpad_out == 0) return SCB_OK;
    SCB_REQUIRE(planes && perm && out_planes, "chunk_majority: null pointer");
    int64_t warps = (int64_t)n_rows * wpad_out;
    chunk_majority_kernel<<<grid_for(warps * 32, 256), 256, 0, as_stream(stream)>>>(
        planes, n_rows, wpad, perm, n_chunks, chunk_size, min_ones, out_planes, wpad_out);
    SCB_LAUNCH_CHECK("chunk_majority");
    return SCB_OK;
}

template <typename T>
static int binarize_pack(const T* dense, int n_rows, int64_t n_cells, int64_t row_stride, T threshold,
                         uint32_t* planes, int64_t wpad, scb_stream_t stream) {
    SCB_REQUIRE(n_rows > 0 && n_cells >= 0 && row_stride >= n_cells, "binarize_pack: bad matrix shape");
    SCB_REQUIRE(wpad > 0 && wpad % 4 == 0 && wpad * 32 >= n_cells, "binarize_pack: wpad must be a positive multiple of 4 covering n_cells");
    SCB_REQUIRE(dense && planes, "binarize_pack: null pointer");
    int64_t items = (int64_t)n_rows * ((wpad + 31) / 32);
    int64_t blocks = (items + 7) / 8;
    int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    binarize_pack_kernel<T><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(dense, n_rows, n_cells, row_stride,
                                                                            threshold, planes, wpad);
    SCB_LAUNCH_CHECK("binarize_pack_kernel");
    return SCB_OK;
}

extern "C" int scb_binarize_pack_f32(const float* dense, int n_rows, int64_t n_cells, int64_t row_stride,
                                     float threshold, uint32_t* planes, int64_t wpad, scb_stream_t stream) {
    return binarize_pack<float>(dense, n_rows, n_cells, row_stride, threshold, planes, wpad, stream);
}

extern "C" int scb_binarize_pack_f64(const double* dense, int n_rows, int64_t n_cells, int64_t row_stride,
                                     double threshold, uint32_t* planes, int64_t wpad, scb_stream_t stream) {
    return binarize_pack<double>(dense, n_rows, n_cells, row_stride, threshold, planes, wpad, stream);
}
