// common.cuh — shared device/host utilities of libnornic_knn (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nornic_knn.h"

namespace nk {

// ---------------------------------------------------------------------------------------------
// Error plumbing: thread-local message (the reference's is a racy process-global,
// pkg/gpu/cuda/cuda_bridge.go:21-33).
// ---------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
const char *get_error();
void clear_error();

#define NK_CUDA_OK(call)                                                                      \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess) {                                                              \
            nk::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return -1;                                                                        \
        }                                                                                     \
    } while (0)

#define NK_CUDA_OK_PTR(call)                                                                  \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess) {                                                              \
            nk::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return nullptr;                                                                   \
        }                                                                                     \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Candidate keys.  A (score, row) pair is one u64: order-preserving score bits in the high word,
// ~row in the low word, so that a plain unsigned compare implements the boundary's ordering
// "(score desc, row index asc)" (strict '>' forward scan of cuda_bridge.go:356-371).  Euclidean
// search stores score = -dist^2, giving (distance asc, row asc).  Key 0 is below every real key
// (ord(-inf) = 0x007fffff) and marks an empty slot.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t ord_bits(float s) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(s);
#else
    union { float f; uint32_t u; } c; c.f = s; uint32_t b = c.u;
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord_to_float(uint32_t o) {
    uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    union { float f; uint32_t u; } c; c.u = b; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float s, uint32_t row) {
    return ((uint64_t)ord_bits(s) << 32) | (uint64_t)(0xffffffffu - row);
}
__host__ __device__ __forceinline__ float key_score(uint64_t k) { return ord_to_float((uint32_t)(k >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t k) { return 0xffffffffu - (uint32_t)k; }

// Counter-based U[-1,1) generator, bit-identical to oracle/knn_oracle.c orc_uniform_at().
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27; z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}
__host__ __device__ __forceinline__ float uniform_at(uint64_t seed, uint64_t elem) {
    uint64_t z = mix64((seed + 1) * 0x9E3779B97F4A7C15ULL + elem * 0xD1B54A32D192ED03ULL);
    uint32_t m = (uint32_t)(z >> 40);
    return (float)m * (1.0f / 8388608.0f) - 1.0f;
}

static inline uint32_t next_pow2(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// Block-wide bitonic sort (descending) of P u64 keys in shared memory.  P is a power of two.
// Every thread of the block must call it.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_bitonic_sort_desc(uint64_t *s, int P) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool desc = (lo & size) == 0;
                uint64_t a = s[lo], b = s[hi];
                if ((a < b) == desc) {
                    s[lo] = b;
                    s[hi] = a;
                }
            }
        }
    }
    __syncthreads();
}

// Prune one candidate buffer (global, capacity cap, *cnt live entries) down to its best k entries,
// sorted descending, and refresh the pass threshold.  Block-wide; sbuf has P >= cap slots.
// tau is the score a new candidate must reach (>=) to be worth buffering.
__device__ __forceinline__ void block_prune(uint64_t *cand, int cap, int *cnt, float *tau, uint32_t k,
                                            uint64_t *sbuf, int P) {
    __syncthreads();
    int n = *cnt;
    if (n > cap) n = cap;
    for (int i = threadIdx.x; i < P; i += blockDim.x) sbuf[i] = i < n ? cand[i] : 0ull;
    block_bitonic_sort_desc(sbuf, P);
    int keep = n < (int)k ? n : (int)k;
    for (int i = threadIdx.x; i < keep; i += blockDim.x) cand[i] = sbuf[i];
    if (threadIdx.x == 0) {
        *cnt = keep;
        *tau = (n >= (int)k) ? key_score(sbuf[k - 1]) : -INFINITY;
    }
    __syncthreads();
}
// Same, for a sub-group of `nthreads` threads (a multiple of 32, all of whole warps) synchronising on
// named barrier `bar_id` instead of the whole CTA: used by the epilogue warps of the tensor-core scan.
__device__ __forceinline__ void group_sync(int bar_id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void group_bitonic_sort_desc(uint64_t *s, int P, int gtid, int nthreads, int bar_id) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            group_sync(bar_id, nthreads);
            for (int t = gtid; t < (P >> 1); t += nthreads) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool desc = (lo & size) == 0;
                uint64_t a = s[lo], b = s[hi];
                if ((a < b) == desc) {
                    s[lo] = b;
                    s[hi] = a;
                }
            }
        }
    }
    group_sync(bar_id, nthreads);
}
__device__ __forceinline__ void group_prune(uint64_t *cand, int cap, int *cnt, float *tau, uint32_t k, uint64_t *sbuf,
                                            int P, int gtid, int nthreads, int bar_id) {
    group_sync(bar_id, nthreads);
    int n = *cnt;
    if (n > cap) n = cap;
    for (int i = gtid; i < P; i += nthreads) sbuf[i] = i < n ? cand[i] : 0ull;
    group_bitonic_sort_desc(sbuf, P, gtid, nthreads, bar_id);
    int keep = n < (int)k ? n : (int)k;
    for (int i = gtid; i < keep; i += nthreads) cand[i] = sbuf[i];
    if (gtid == 0) {
        *cnt = keep;
        *tau = (n >= (int)k) ? key_score(sbuf[k - 1]) : -INFINITY;
    }
    group_sync(bar_id, nthreads);
}
// ---------------------------------------------------------------------------------------------
// Warp-level prune: select the best k of n <= 32*PER_LANE buffered keys with the keys held in registers
// (PER_LANE per lane) by repeated warp-wide max extraction; writes them back sorted descending.  No block
// barrier, so several warps prune different queries concurrently.  Cost ~ k * 60 instructions.
// Caller guarantees every producer of cand[] has finished (barrier) and that *cnt <= 32*PER_LANE.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, o);
        uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), o);
        uint64_t w = ((uint64_t)hi << 32) | lo;
        v = w > v ? w : v;
    }
    return v;
}
// margin_mode (tensor-core filter scan): keys carry UPPER bounds; after the k-th best everything whose bound is
// still >= (k-th bound - margin2) is kept as well (it may beat the k-th once re-scored exactly), up to max_keep
// entries, and tau = k-th bound - margin2.  Exact mode keeps exactly the best k and tau = k-th score.
template <int PER_LANE>
__device__ __forceinline__ void warp_prune(uint64_t *cand, int *cnt, float *tau, uint32_t k, int lane, uint64_t *out,
                                           int out_len, bool margin_mode, float margin2, int max_keep) {
    int n = *cnt;
    if (n > 32 * PER_LANE) n = 32 * PER_LANE;
    uint64_t v[PER_LANE];
    uint64_t lmax = 0;
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) {
        int idx = lane + 32 * i;
        v[i] = idx < n ? cand[idx] : 0ull;
        lmax = v[i] > lmax ? v[i] : lmax;
    }
    __syncwarp();
    const int limit = margin_mode ? (n < max_keep ? n : max_keep) : (n < (int)k ? n : (int)k);
    float kth = -INFINITY;
    int keep = 0;
    for (int j = 0; j < limit; ++j) {
        const uint64_t m = warp_max_u64(lmax);
        if (m == 0ull) break;
        if (j >= (int)k && key_score(m) < kth - margin2) break;  // margin mode only (exact mode stops at limit == k)
        if (j == (int)k - 1) kth = key_score(m);
        keep = j + 1;
        if (lane == 0) {
            cand[j] = m;
            if (out && j < out_len) out[j] = m;
        }
        if (lmax == m) {  // keys are unique (row id in the low word): exactly one lane owns it
            lmax = 0;
#pragma unroll
            for (int i = 0; i < PER_LANE; ++i) {
                if (v[i] == m) v[i] = 0ull;
                lmax = v[i] > lmax ? v[i] : lmax;
            }
        }
    }
    if (out)
        for (int j = keep + lane; j < out_len; j += 32) out[j] = 0ull;
    if (lane == 0) {
        *cnt = keep;
        *tau = (n >= (int)k) ? kth - (margin_mode ? margin2 : 0.0f) : -INFINITY;
    }
    __syncwarp();
}
#endif  // __CUDACC__

}  // namespace nk
