// tmem_contention.cu — does tcgen05.st traffic from other warps slow tcgen05.mma (A from TMEM), and vice versa?
// One CTA per SM: warp 0 issues 24-MMA "slabs" back to back (M=128,N=64,K=8 tf32, A in TMEM); warps 4..11 each
// store 2 x STTM.x32 per slab-equivalent (the split warps' TMEM traffic), free-running.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "ptx_sm100.cuh"
using namespace nk::ptx;

__global__ void __launch_bounds__(384, 1) k(int iters, int do_mma, int st_per_iter, int do_lds, long long *out) {
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
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long t0 = clock64();
    if (warp == 0) {
        if (do_mma) {
            const uint32_t idesc = make_idesc_tf32(128, 64);
            const uint64_t bdesc = make_smem_desc_sw128(smem_u32(base));
            for (int it = 0; it < iters; ++it) {
                if (elect_one_sync()) {
#pragma unroll
                    for (int m = 0; m < 2; ++m)
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                            for (int r = 0; r < 3; ++r)
                                mma_tf32_ts(tmem + m * 64, tmem + 128 + 64 * (m) + (r == 1 ? 32 : 0) + kk * 8, bdesc + kk * 2, idesc, 1);
                }
                __syncwarp();
            }
            if (elect_one_sync()) tc_commit(&bar);
            __syncwarp();
            mbar_wait(&bar, 0);
        }
        long long t1 = clock64();
        if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    } else if (warp >= 4) {
        const uint32_t quad = warp & 3, m = (warp - 4) >> 2;
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = lane + i;
        float acc = 0.f;
        for (int it = 0; it < iters; ++it) {
            if (do_lds) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float4 x = *reinterpret_cast<const float4 *>(base + ((it & 7) * 4096) + lane * 128 + ((c ^ (lane & 7)) << 4));
                    acc += x.x + x.y + x.z + x.w;
                }
            }
            for (int s = 0; s < st_per_iter; ++s) tmem_st_32x32b_x32(tmem + ((quad * 32u) << 16) + 256 + (((it + s) & 3) * 2 + m) * 32, v);
            if (st_per_iter) tmem_wait_st();
        }
        if (acc == 1234.5f) out[3] = 1;
        long long t1 = clock64();
        if (warp == 4 && lane == 0 && blockIdx.x == 0) out[1] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
    long long *d, h[4];
    cudaMalloc(&d, 32);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
    const int iters = 4000;
    int cfg[][3] = {{1, 0, 0}, {0, 2, 0}, {1, 2, 0}, {0, 2, 1}, {1, 2, 1}, {1, 1, 1}, {1, 4, 0}};
    for (auto &c : cfg) {
        cudaMemset(d, 0, 32);
        k<<<148, 384, 66 * 1024>>>(iters, c[0], c[1], c[2], d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("mma=%d sttm_per_iter=%d lds=%d : %s  mma warp %.0f cyc/slab, split warp %.0f cyc/iter\n", c[0], c[1], c[2], cudaGetErrorString(e),
               (double)h[0] / iters, (double)h[1] / iters);
    }
    return 0;
}
