// rowops.cu — one-pass row kernels behind the legacy ABI and index maintenance.
//
//   row_norms       replaces n x cublasSnrm2                  (cuda_bridge.go:231-246)
//   normalize_rows  replaces n x cublasSnrm2 + D2H + n x cublasSscal (cuda_bridge.go:249-284)
//   row_scores      replaces cublasSgemv                      (cuda_bridge.go:290-318)
//   fill_uniform    synthetic corpora generated in HBM (bench / tests), bit-identical to the oracle
//   gather_rows     row gather for ScoreSubset (gpu.go:1578-1589 does this on the host + re-upload)
// All are warp-per-row, lanes striding the row (coalesced), 64-bit row offsets (the reference's orphan
// kernels overflow uint32 at n*dim >= 2^32, cuda_kernels.cu:195,274,298).
#include "kernels.cuh"

namespace nk {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void row_norms_kernel(const float *rows, float *norms, uint32_t n, uint32_t dim) {
    uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = warp; r < n; r += nwarps) {
        const float *p = rows + (size_t)r * dim;
        float a = 0.0f;
        for (uint32_t j = lane; j < dim; j += 32) {
            float x = __ldg(p + j);
            a = fmaf(x, x, a);
        }
        a = warp_sum(a);
        if (lane == 0) norms[r] = sqrtf(a);
    }
}

__global__ void normalize_rows_kernel(float *rows, uint32_t n, uint32_t dim) {
    uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = warp; r < n; r += nwarps) {
        float *p = rows + (size_t)r * dim;
        float a = 0.0f;
        for (uint32_t j = lane; j < dim; j += 32) a = fmaf(p[j], p[j], a);
        float nrm = sqrtf(warp_sum(a));
        if (nrm > 1e-10f) {  // cuda_bridge.go:267: rows with norm <= 1e-10 are left untouched
            float inv = 1.0f / nrm;
            for (uint32_t j = lane; j < dim; j += 32) p[j] *= inv;
        }
    }
}

__global__ void row_scores_kernel(const float *rows, const float *query, float *scores, uint32_t n, uint32_t dim,
                                  int normalized) {
    uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    float qq = 0.0f;
    if (!normalized) {
        for (uint32_t j = lane; j < dim; j += 32) qq = fmaf(query[j], query[j], qq);
        qq = warp_sum(qq);
    }
    for (uint32_t r = warp; r < n; r += nwarps) {
        const float *p = rows + (size_t)r * dim;
        float d = 0.0f, xx = 0.0f;
        for (uint32_t j = lane; j < dim; j += 32) {
            float x = __ldg(p + j);
            d = fmaf(x, __ldg(query + j), d);
            xx = fmaf(x, x, xx);
        }
        d = warp_sum(d);
        xx = warp_sum(xx);
        if (lane == 0) {
            if (normalized) {
                scores[r] = d;
            } else {
                float den = sqrtf(xx * qq);
                scores[r] = den > 0.0f ? d / den : 0.0f;
            }
        }
    }
}

template <typename T>
__global__ void fill_uniform_kernel(T *out, uint64_t total, uint32_t dim, uint64_t seed, uint64_t row_base) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        float v = uniform_at(seed, row_base * dim + i);
        if constexpr (sizeof(T) == 2) out[i] = __float2half_rn(v);
        else out[i] = v;
    }
}

template <typename T>
__global__ void gather_rows_kernel(const T *rows, uint32_t dim, const uint32_t *idx, uint32_t n_idx, T *out) {
    uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = warp; r < n_idx; r += nwarps) {
        const T *src = rows + (size_t)idx[r] * dim;
        T *dst = out + (size_t)r * dim;
        for (uint32_t j = lane; j < dim; j += 32) dst[j] = src[j];
    }
}

// ---- k-means update step (pkg/gpu/kmeans.go:585-618) -------------------------------------------------------------
// Per-cluster float64 sums of the member rows: one warp per row, lanes stride the dimensions, fp64 atomic adds (the
// K x dim accumulator is spread over L2, so contention is low; fp64 keeps the sum order-independent to ~1e-16, far
// below the float32 the mean is rounded to).  Rows whose assignment is outside [0, K) are skipped.
__global__ void cluster_sums_kernel(const float *rows, uint64_t n, uint32_t dim, const int32_t *assign, uint32_t K, double *sums,
                                    unsigned long long *counts) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t row = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n; row += warps) {
        const int32_t c = assign[row];
        if (c < 0 || (uint32_t)c >= K) continue;
        const float *x = rows + row * dim;
        double *dst = sums + (size_t)c * dim;
        for (uint32_t j = lane; j < dim; j += 32) atomicAdd(dst + j, (double)x[j]);
        if (lane == 0) atomicAdd(counts + c, 1ull);
    }
}
__global__ void convert_f32_to_f16_kernel(const float *src, __half *dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = __float2half_rn(src[i]);
}
// changed += (a[i] != b[i])
__global__ void count_changed_kernel(const int32_t *a, const uint32_t *b, uint64_t n, unsigned long long *changed) {
    unsigned long long mine = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        mine += a[i] != (int32_t)b[i];
    mine = __reduce_add_sync(0xffffffffu, (unsigned)mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(changed, mine);
}

static inline unsigned warp_grid(uint32_t n) {
    uint64_t blocks = ((uint64_t)n + 7) / 8;  // 8 warps per 256-thread block
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks == 0) blocks = 1;
    return (unsigned)blocks;
}

int row_norms(const float *rows, float *norms, uint32_t n, uint32_t dim, cudaStream_t s) {
    if (n == 0) return 0;
    row_norms_kernel<<<warp_grid(n), 256, 0, s>>>(rows, norms, n, dim);
    NK_CUDA_OK(cudaGetLastError());
    return 0;
}
int normalize_rows(float *rows, uint32_t n, uint32_t dim, cudaStream_t s) {
    if (n == 0) return 0;
    normalize_rows_kernel<<<warp_grid(n), 256, 0, s>>>(rows, n, dim);
    NK_CUDA_OK(cudaGetLastError());
    return 0;
}
int row_scores(const float *rows, const float *query, float *scores, uint32_t n, uint32_t dim, int normalized,
               cudaStream_t s) {
    if (n == 0) return 0;
    row_scores_kernel<<<warp_grid(n), 256, 0, s>>>(rows, query, scores, n, dim, normalized);
    NK_CUDA_OK(cudaGetLastError());
    return 0;
}
int fill_uniform(void *out, int dtype, uint64_t n_rows, uint32_t dim, uint64_t seed, uint64_t row_base, cudaStream_t s) {
    uint64_t total = n_rows * dim;
    if (total == 0) return 0;
    unsigned grid = 148 * 16;
    if (dtype == NK_DTYPE_F16)
        fill_uniform_kernel<__half><<<grid, 256, 0, s>>>((__half *)out, total, dim, seed, row_base);
    else
        fill_uniform_kernel<float><<<grid, 256, 0, s>>>((float *)out, total, dim, seed, row_base);
    NK_CUDA_OK(cudaGetLastError());
    return 0;
}
int cluster_sums(const float *rows, uint64_t n, uint32_t dim, const int32_t *assign, uint32_t K, double *sums,
                 unsigned long long *counts, cudaStream_t s) {
    if (n == 0) return 0;
    cluster_sums_kernel<<<148 * 8, 256, 0, s>>>(rows, n, dim, assign, K, sums, counts);
    NK_CUDA_OK(cudaGetLastError());
    return 0;
}
int convert_f32_to_f16(const float *src, void *dst, size_t n, cudaStream_t s) {
    if (n == 0) return 0;
    convert_f32_to_f16_kernel<<<148 * 8, 256, 0, s>>>(src, static_cast<__half *>(dst), n);
    NK_CUDA_OK(cudaGetLastError());
    return 0;
}
int count_changed(const int32_t *a, const uint32_t *b, uint64_t n, unsigned long long *changed, cudaStream_t s) {
    if (n == 0) return 0;
    count_changed_kernel<<<148 * 4, 256, 0, s>>>(a, b, n, changed);
    NK_CUDA_OK(cudaGetLastError());
    return 0;
}
int gather_rows(const void *rows, int dtype, uint32_t dim, const uint32_t *idx, uint32_t n_idx, void *out,
                cudaStream_t s) {
    if (n_idx == 0) return 0;
    if (dtype == NK_DTYPE_F16)
        gather_rows_kernel<__half><<<warp_grid(n_idx), 256, 0, s>>>((const __half *)rows, dim, idx, n_idx, (__half *)out);
    else
        gather_rows_kernel<float><<<warp_grid(n_idx), 256, 0, s>>>((const float *)rows, dim, idx, n_idx, (float *)out);
    NK_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace nk
