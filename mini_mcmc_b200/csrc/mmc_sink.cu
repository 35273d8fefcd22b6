// Sample sinks, device half (src/io/arrow.rs:53-117, src/io/parquet.rs:49-122,154-221).
// The reference walks the [chain, observation, dim] array element by element into per-column Arrow builders
// (u32 chain, u32 observation, f64 dim_i).  Here a block of chains is turned into the f64 columns on the device
// (a widening transpose, HBM bound: 4-8 B read + 8 B written per element), copied into pinned staging and handed to
// the Arrow / Parquet writer as zero-copy column buffers while the next block is transposed and copied.
#include "mmc_common.cuh"

namespace mmc {
namespace {

// dim >= 32: 32 x 32 tiles through shared memory, coalesced on both sides
template <typename InT>
__global__ void __launch_bounds__(256) sink_transpose_kernel(const InT *__restrict__ in, double *__restrict__ out, int64_t rows, int dim) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int d0 = blockIdx.y * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = r0 + ty + 8 * i;
        const int d = d0 + tx;
        if (r < rows && d < dim) tile[ty + 8 * i][tx] = (double)in[r * dim + d];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int d = d0 + ty + 8 * i;
        const int64_t r = r0 + tx;
        if (r < rows && d < dim) out[(int64_t)d * rows + r] = tile[tx][ty + 8 * i];
    }
}

// dim < 32: flat coalesced read, column runs of 32 / dim rows on the write side
template <typename InT>
__global__ void __launch_bounds__(256) sink_flat_kernel(const InT *__restrict__ in, double *__restrict__ out, int64_t rows, int dim) {
    const int64_t total = rows * dim;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / dim;
        const int d = (int)(e - r * dim);
        out[(int64_t)d * rows + r] = (double)in[e];
    }
}

template <typename InT>
int launch_sink(const void *sample, int64_t n, int dim, int64_t c0, int64_t c_count, double *out, cudaStream_t s) {
    const int64_t rows = c_count * n;
    if (rows == 0) return MMC_OK;
    const InT *in = static_cast<const InT *>(sample) + c0 * n * dim;
    if (dim >= 32) {
        const dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((dim + 31) / 32));
        sink_transpose_kernel<InT><<<grid, 256, 0, s>>>(in, out, rows, dim);
    } else {
        const int64_t total = rows * dim;
        const int64_t blocks = std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
        sink_flat_kernel<InT><<<(unsigned)blocks, 256, 0, s>>>(in, out, rows, dim);
    }
    MMC_CUDA(cudaGetLastError());
    return MMC_OK;
}

}  // namespace
}  // namespace mmc

using namespace mmc;

extern "C" int mmc_sink_columns_dev(const void *sample_dev, int32_t dtype, int64_t chains, int64_t n, int32_t dim, int64_t c0,
                                    int64_t c_count, double *dim_cols_dev, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    MMC_REQUIRE(sample_dev && dim_cols_dev && chains >= 0 && n >= 0 && dim > 0 && c0 >= 0 && c_count >= 0 && c0 + c_count <= chains,
                "mmc_sink_columns_dev: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == MMC_F32) return launch_sink<float>(sample_dev, n, dim, c0, c_count, dim_cols_dev, s);
    if (dtype == MMC_F64) return launch_sink<double>(sample_dev, n, dim, c0, c_count, dim_cols_dev, s);
    if (dtype == MMC_U64) return launch_sink<unsigned long long>(sample_dev, n, dim, c0, c_count, dim_cols_dev, s);
    set_error("mmc_sink_columns_dev: unknown dtype %d", dtype);
    return MMC_ERR_INVALID;
}
