// mma_rate.cu — microbenchmark: cycles per tcgen05.mma kind::tf32 (M=128, K=8) for N in {64,128,256}, A operand from
// TMEM (TS) or shared memory (SS), and kind::f16 (bf16, M=128, K=16, A from TMEM) for N in {128,256}, issued back-to-back
// by one elected thread, one CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../nornicdb_b200/csrc -o mma_rate mma_rate.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "ptx_sm100.cuh"
using namespace nk::ptx;

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long *out_cycles) {
    extern __shared__ unsigned char dyn[];
    unsigned char *base = dyn + ((1024u - (smem_u32(dyn) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<float *>(base)[i] = 1.0f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_base, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (threadIdx.x < 32) {
        const uint32_t idesc = make_idesc_tf32(128, N);
        const uint64_t bdesc = make_smem_desc_sw128(smem_u32(base));             // B: N rows x 128 B
        const uint64_t adesc = make_smem_desc_sw128(smem_u32(base + 32768));     // A: 128 rows x 128 B (SS)
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (elect_one_sync()) {
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            if (TS) mma_tf32_ts(tmem + m * N, tmem + 256 + 64 * ((it & 1) * 2 + m) + (r == 1 ? 32 : 0) + kk * 8, bdesc + kk * 2, idesc, 1);
                            else mma_tf32_ss(tmem + m * N, adesc + kk * 2, bdesc + kk * 2, idesc, 1);
                        }
            }
            __syncwarp();
        }
        if (elect_one_sync()) tc_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) *out_cycles = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel_bf16(int iters, long long *out_cycles) {
    extern __shared__ unsigned char dyn[];
    unsigned char *base = dyn + ((1024u - (smem_u32(dyn) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(base)[i] = 0x3f803f80u;  // bf16 1.0 pairs
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_base, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (threadIdx.x < 32) {
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint64_t bdesc = make_smem_desc_sw128(smem_u32(base));  // B: N rows x 128 B (64 bf16)
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (elect_one_sync()) {
#pragma unroll
                for (int r = 0; r < 6; ++r)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) mma_bf16_ts(tmem, tmem + 256 + (r & 3) * 32 + kk * 8, bdesc + kk * 2, idesc, 1);
            }
            __syncwarp();
        }
        if (elect_one_sync()) tc_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) *out_cycles = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N>
void run_bf16(long long *d_out) {
    const int iters = 2000;
    cudaFuncSetAttribute(rate_kernel_bf16<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
    rate_kernel_bf16<N><<<148, 128, 66 * 1024>>>(iters, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
    printf("bf16 N=%3d A=TMEM: %s  %.1f cycles per MMA (M=128,K=16)\n", N, cudaGetErrorString(e), (double)c / (iters * 24.0));
}

template <int N, bool TS>
void run(long long *d_out) {
    const int iters = 2000;
    cudaFuncSetAttribute(rate_kernel<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
    rate_kernel<N, TS><<<148, 128, 66 * 1024>>>(iters, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
    printf("N=%3d %s: %s  %.1f cycles per MMA (M=128,K=8), %.1f per 24-MMA slab\n", N, TS ? "A=TMEM" : "A=SMEM", cudaGetErrorString(e),
           (double)c / (iters * 24.0), (double)c / iters);
}

int main() {
    long long *d_out;
    cudaMalloc(&d_out, 8);
    run<64, true>(d_out);
    run<128, true>(d_out);
    run<256, true>(d_out);
    run<64, false>(d_out);
    run<128, false>(d_out);
    run<256, false>(d_out);
    run_bf16<128>(d_out);
    run_bf16<256>(d_out);
    return 0;
}
