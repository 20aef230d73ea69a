// scan_tensor.cu — fused distance + top-k scan over the FP32 rows on the 5th-gen tensor cores (tcgen05 + TMEM + TMA):
// the path of shards without a BF16 shadow, the retry stage of the shadow filter (scan_tensor_shadow.cu), the exact
// 3xTF32 scan, plus the finish step and the host logic both filters share.
//
// For Q > ~16 queries the CUDA-core scan stops being HBM-bound (fp32 FMA ridge ~10 flop/byte, the batch needs
// Q/2 flop/byte), so the Q x N^T contraction moves to tcgen05.mma.  fp32 inputs on tensor cores mean TF32
// (10-bit mantissa), which by itself cannot meet "identical index sets, 1e-4".  Two modes (template NT):
//
//   NT = 3  "exact":  3xTF32 split  x*q ~= xl*qh + xh*ql + xh*qh  (xh = rn_tf32(x), xl = x - xh exactly),
//           fp32 accumulation in TMEM, ~2^-21 relative per product.  Per-CTA exact top-k lists.
//   NT = 1  "filter": ONE TF32 product per element (a third of the tensor work and energy — the chip is
//           power-capped when HBM and the tensor pipes both run flat out) with a RIGOROUS error margin:
//           |s_hat - s| <= c*|x|*|q|, c = 2^-10 + d*2^-22  (rounding of both operands + fp32 accumulation), so
//           every row whose upper bound s_hat + B can still reach the running k-th best lower bound is kept.
//           The few survivors (k + a handful) are re-scored EXACTLY in fp32 by filter_finish_kernel, so final
//           scores/indices carry no TF32 error at all.  If a margin buffer overflows (adversarial near-ties) a flag
//           is raised ON DEVICE and the NT=3 kernels — enqueued behind, early-exiting when the flag is clear —
//           redo the search exactly.  No host round trip.
//
// One persistent CTA per SM, warp-specialised (16 warps):
//   warp 0      TMA producer: corpus K-slabs [256 rows x 32 floats] (128B-swizzled), ring released by the split
//               warps as soon as the slab is in registers (EVICT_FIRST);
//   warp 3      TMA producer: query slabs (pre-rounded hi [+ lo], [64 x 32 floats] each, L2-resident, EVICT_LAST);
//   warps 4-11  "split" warps, one thread per corpus row: conflict-free swizzled LDS.128, xh (and xl) with packed
//               FADD2/FFMA2, |x|^2 on the side (no separate norm pass), tcgen05.st into a TMEM ring: the corpus is
//               the MMA's A operand FROM TENSOR MEMORY (smem-operand MMAs of this shape are 1.5x slower, see
//               profiles/experiments/mma_rate.cu);
//   warps 1,2   MMA issuers, one per 128-row M-tile (tcgen05.mma.kind::tf32 M=128 N=64 K=8, accumulators in TMEM);
//               two issuers because one warp's per-slab poll/fence/commit overhead lets the shallow MMA queue drain;
//   warps 12-15 epilogue: tcgen05.ld the [128 rows x 64 queries] accumulator (thread = corpus row), release it,
//               scale / bound, compare against each query's running threshold with a compact mask pass, append the
//               few survivors; warp-level register top-k prune.  Distances never go to memory.
// Exact mode: per-CTA lists are folded by merge_keys().  Filter mode: CTAs share their thresholds through a per-query
// atomic max, append their survivors to one list per query, and filter_finish_kernel (below) selects, re-scores and
// sorts.  Large batches: 2 / 4 query groups per launch share every corpus tile through L2 (Params::qgroups).
//
// Algorithmic HBM traffic per launch: n*dim*4 (corpus, once); query re-reads are served from L2.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>

#include "kernels.cuh"
#include "ptx_sm100.cuh"
#include "scan_tensor_shared.cuh"

namespace nk {

namespace tc {
constexpr int ROWS = 256;        // corpus rows per tile (2 M-tiles of 128)
constexpr int QT_MAX = 128;      // queries per launch (MMA N): 64, or 128 for large batches (filter mode)
constexpr int ASTAGES = 4;       // corpus-slab smem ring
constexpr int A_BYTES = ROWS * BK * 4;   // 32 KB
constexpr int ACC_COL = 0;       // [mtile] x QT columns (single-buffered, drained per M-tile); the A ring follows
constexpr int SPLIT_WARP0 = 4, EPI_WARP0 = 12;
constexpr int MAX_BSTAGES = 8, MAX_TSTAGES = 6;
constexpr int XX_RING = 8;       // >= MAX_TSTAGES / 1 slab-per-tile + 2

template <int NT, int QT> struct Cfg {
    static constexpr int PARTS = NT == 3 ? 2 : 1;          // hi (+ lo)
    static constexpr int B_BYTES = QT * BK * 4;            // one part of one query slab (8 / 16 KB)
    static constexpr int BST_BYTES = PARTS * B_BYTES;
    static constexpr int BSTAGES = (80 * 1024) / BST_BYTES > MAX_BSTAGES ? MAX_BSTAGES : (80 * 1024) / BST_BYTES;
    static constexpr int A_COL = 2 * QT;                   // TMEM: accumulators [2][QT], then the A-operand ring
    static constexpr int TSTAGES = (TMEM_COLS - A_COL) / (2 * PARTS * BK);  // 3 / 6 (QT=64), 4 (QT=128 filter)
    static constexpr int RING_BYTES = ASTAGES * A_BYTES + BSTAGES * BST_BYTES;
    static_assert(BSTAGES >= 2 && TSTAGES >= 2 && TSTAGES <= MAX_TSTAGES, "ring too shallow");
};

struct __align__(8) Shared {
    uint64_t afull_s[ASTAGES], aempty_s[ASTAGES];    // corpus slab in smem: TMA -> split warps -> TMA
    uint64_t bfull[MAX_BSTAGES], bempty[MAX_BSTAGES];  // query slabs in smem: TMA -> MMA -> TMA
    uint64_t afull[MAX_TSTAGES][2], aempty[MAX_TSTAGES][2];  // A operand in TMEM, per (stage, M-tile)
    uint64_t accfull[2], accempty[2];                // accumulators, per M-tile
    uint32_t tmem_base;
    unsigned int maxxx;       // running max of |x|^2 (float bits) over the rows this CTA has scored
    float xx[XX_RING][ROWS];  // |x|^2 per row, ring over tiles: the split warps run up to TSTAGES slabs (several tiles when
                              // dim is small) ahead of the MMA, which runs one tile ahead of the epilogue
    float tau[QT_MAX];
    float qn[QT_MAX];
    int cnt[QT_MAX];
};
}  // namespace tc

// Wait-time instrumentation (NK_TC_DEBUG bit 64): cycles CTA 0 spends blocked at each hand-off.
__device__ long long g_tc_prof[32];
#define TC_PROF_BEGIN() long long _t0 = prof ? clock64() : 0
#define TC_PROF_END(slot) do { if (prof) { long long _t1 = clock64(); acc_##slot += _t1 - _t0; } } while (0)

// Exact fp32 score of one corpus row against the query staged in shared memory, by one warp: the arithmetic every
// filter path re-scores its survivors with (fp32 FMA, 128-bit loads; fp16 rows are widened exactly).  Returns the key
// score: dot | dot / sqrt(|x|^2 |q|^2) (0 for zero vectors, simd_amd64.go:31-35) | -|x - q|^2; NaN -> -inf.  *xx_out = |x|^2.
__device__ __forceinline__ float warp_exact_score(const void *rows, int dtype, size_t local, uint32_t dim, const float *qs, float qq,
                                                  int metric, int lane, float *xx_out = nullptr) {
    float d = 0.0f, xx = 0.0f;
    const float4 *q4 = reinterpret_cast<const float4 *>(qs);
    auto acc4 = [&](const float4 v, const float4 u) {
        if (metric == NK_METRIC_EUCLIDEAN) {
            float t;
            t = v.x - u.x; d = fmaf(t, t, d);
            t = v.y - u.y; d = fmaf(t, t, d);
            t = v.z - u.z; d = fmaf(t, t, d);
            t = v.w - u.w; d = fmaf(t, t, d);
            xx = fmaf(v.x, v.x, xx); xx = fmaf(v.y, v.y, xx); xx = fmaf(v.z, v.z, xx); xx = fmaf(v.w, v.w, xx);
        } else {
            d = fmaf(v.x, u.x, d); xx = fmaf(v.x, v.x, xx);
            d = fmaf(v.y, u.y, d); xx = fmaf(v.y, v.y, xx);
            d = fmaf(v.z, u.z, d); xx = fmaf(v.z, v.z, xx);
            d = fmaf(v.w, u.w, d); xx = fmaf(v.w, v.w, xx);
        }
    };
    if (dtype == NK_DTYPE_F32) {
        const float4 *x4 = reinterpret_cast<const float4 *>(static_cast<const float *>(rows) + local * dim);
#pragma unroll 8
        for (uint32_t j = lane; j < dim / 4; j += 32) acc4(__ldg(x4 + j), q4[j]);
    } else if (dtype == NK_DTYPE_F16) {  // fp16 rows, dim % 8 == 0: 8 halves per 128-bit load
        const uint4 *x8 = reinterpret_cast<const uint4 *>(static_cast<const __half *>(rows) + local * dim);
#pragma unroll 4
        for (uint32_t j = lane; j < dim / 8; j += 32) {
            const uint4 w = __ldg(x8 + j);
            const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&w.x)), f1 = __half22float2(*reinterpret_cast<const __half2 *>(&w.y));
            const float2 f2 = __half22float2(*reinterpret_cast<const __half2 *>(&w.z)), f3 = __half22float2(*reinterpret_cast<const __half2 *>(&w.w));
            acc4(make_float4(f0.x, f0.y, f1.x, f1.y), q4[2 * j]);
            acc4(make_float4(f2.x, f2.y, f3.x, f3.y), q4[2 * j + 1]);
        }
    } else {  // bf16 rows: widening is a 16-bit shift
        const uint4 *x8 = reinterpret_cast<const uint4 *>(static_cast<const uint16_t *>(rows) + local * dim);
#pragma unroll 4
        for (uint32_t j = lane; j < dim / 8; j += 32) {
            const uint4 w = __ldg(x8 + j);
            acc4(make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u), __uint_as_float(w.y << 16), __uint_as_float(w.y & 0xffff0000u)), q4[2 * j]);
            acc4(make_float4(__uint_as_float(w.z << 16), __uint_as_float(w.z & 0xffff0000u), __uint_as_float(w.w << 16), __uint_as_float(w.w & 0xffff0000u)), q4[2 * j + 1]);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        d += __shfl_xor_sync(0xffffffffu, d, o);
        xx += __shfl_xor_sync(0xffffffffu, xx, o);
    }
    float sc = d;
    if (metric == NK_METRIC_EUCLIDEAN) sc = -d;
    else if (metric == NK_METRIC_COSINE) {
        const float den = sqrtf(xx * qq);
        sc = den > 0.0f ? d / den : 0.0f;
    }
    if (sc != sc) sc = -INFINITY;
    if (xx_out) *xx_out = xx;
    return sc;
}

// ---------------------------------------------------------------------------------------------------
// filter_prep_kernel — everything a filter search needs before its scan, in ONE launch (one CTA per padded query row):
//   * |q|, cosine normalisation; TF32 hi / lo split (TF32 filter + exact 3xTF32 stages); BF16 copy with the measured
//     rounding residues qa = |bf16(q)|, qb = |q - bf16(q)| + acc_c |q| (shadow stage);
//   * clears the search state the scan kernels accumulate into (shared thresholds, list fills, status words 1..7) —
//     previously two cudaMemsetAsync;
//   * SAMPLED INITIAL THRESHOLD: the query's exact fp32 score against `sample` evenly spaced rows of the shard; the k-th
//     largest of their LOWER bounds (score minus the fp32 summation allowance) is a lower bound of the true k-th best
//     score, so the scan starts with a real threshold instead of buffering every (row, query) pair of its first two
//     tiles and pruning 64-128 full buffers per CTA (~25 us per launch, the largest fixed cost of a small-shard search).
// ---------------------------------------------------------------------------------------------------
constexpr int PREP_THREADS = 1024;  // 32 warps: the sample scoring is latency-bound (one 4 KB row per warp at a time)
struct PrepParams {
    const float *q;
    uint32_t Q, dim, dimpad;
    int normalise, metric;
    float acc_c;
    float *qhi, *qlo;      // [Qpad x dim] (nullable)
    uint16_t *qbf;         // [Qpad x dimpad] (nullable); fp16 bits when qbf_f16
    int qbf_f16;
    float *qnorm, *qa, *qb;
    uint32_t *state;       // gtau[QA] then gcount[QA]
    uint32_t Qpad, QA;
    int *flags;
    // sampled initial threshold (sample == 0: none)
    const void *rows;
    int dtype;
    uint32_t n, sample, k;
    const uint32_t *mask;
    float min_score;
    const int *only_if;
};

__global__ void __launch_bounds__(PREP_THREADS) filter_prep_kernel(PrepParams p) {
    extern __shared__ __align__(16) unsigned char prep_smem[];
    float *qs = reinterpret_cast<float *>(prep_smem);                                  // raw query [dim]
    uint64_t *skeys = reinterpret_cast<uint64_t *>(prep_smem + (size_t)((p.dim + 3) & ~3u) * 4);  // [sample]
    __shared__ float red[3][PREP_THREADS / 32];
    if (p.only_if && *p.only_if == 0) return;
    const uint32_t row = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    constexpr int NW = PREP_THREADS / 32;
    // ---- clear the search state: own entry, block 0 also the padding and the status words
    if (p.state) {
        if (tid == 0) { p.state[row] = 0u; p.state[p.QA + row] = 0u; }
        if (row == 0) {
            for (uint32_t i = p.Qpad + tid; i < p.QA; i += PREP_THREADS) { p.state[i] = 0u; p.state[p.QA + i] = 0u; }
            if (tid >= 1 && tid <= 7) p.flags[tid] = 0;
        }
    }
    float t = 0.0f;
    if (row < p.Q) {
        float a = 0.0f;
        for (uint32_t j = tid; j < p.dim; j += PREP_THREADS) {
            const float v = p.q[(size_t)row * p.dim + j];
            qs[j] = v;
            a = fmaf(v, v, a);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) red[0][w] = a;
    }
    __syncthreads();
    if (row < p.Q)
        for (int i = 0; i < NW; ++i) t += red[0][i];
    const float nrm = sqrtf(t);
    // zero query -> all cosine scores 0 (simd_amd64.go:31-35)
    const float inv = p.normalise ? (t > 0.0f ? 1.0f / nrm : 0.0f) : 1.0f;
    const float qn = p.normalise ? (t > 0.0f ? 1.0f : 0.0f) : nrm;
    if (p.qhi) {
        for (uint32_t j = tid; j < p.dim; j += PREP_THREADS) {
            const float v = row < p.Q ? qs[j] * inv : 0.0f;
            const float h = __uint_as_float(ptx::tf32_round_bits(__float_as_uint(v)));
            p.qhi[(size_t)row * p.dim + j] = h;
            if (p.qlo) p.qlo[(size_t)row * p.dim + j] = v - h;
        }
    }
    float hh = 0.0f, dd = 0.0f;
    if (p.qbf) {
        for (uint32_t j = tid; j < p.dimpad; j += PREP_THREADS) {
            const float v = (row < p.Q && j < p.dim) ? qs[j] * inv : 0.0f;
            uint16_t b;
            float vb;
            if (p.qbf_f16) {
                const __half h = __float2half_rn(v);
                b = __half_as_ushort(h);
                vb = __half2float(h);
            } else {
                b = ptx::f32_to_bf16_bits(v);
                vb = __uint_as_float((uint32_t)b << 16);
            }
            p.qbf[(size_t)row * p.dimpad + j] = b;
            hh = fmaf(vb, vb, hh);
            dd = fmaf(v - vb, v - vb, dd);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            hh += __shfl_xor_sync(0xffffffffu, hh, o);
            dd += __shfl_xor_sync(0xffffffffu, dd, o);
        }
        if (lane == 0) { red[1][w] = hh; red[2][w] = dd; }
    }
    __syncthreads();
    if (tid == 0) {
        if (p.qnorm) p.qnorm[row] = qn;
        if (p.qbf) {
            hh = dd = 0.0f;
            for (int i = 0; i < NW; ++i) { hh += red[1][i]; dd += red[2][i]; }
            p.qa[row] = sqrtf(hh) * 1.0001f;
            p.qb[row] = (sqrtf(dd) + p.acc_c * qn) * 1.0001f;
        }
    }
    // ---- sampled initial threshold
    if (p.sample == 0 || row >= p.Q) return;
    const uint32_t S = p.sample;  // power of two
    auto sample_row = [&](uint32_t i) -> uint64_t { return p.n >= S ? (uint64_t)i * p.n / S : i; };
    auto lower_key = [&](float sc, float xx) -> uint64_t {
        // lower bound of the real-valued score: the fp32 summation allowance the filters use (acc_c |x||q|)
        const float lo = sc - (p.metric == NK_METRIC_COSINE ? p.acc_c : p.metric == NK_METRIC_DOT ? p.acc_c * sqrtf(xx * t) : p.acc_c * -sc);
        return (lo == lo && lo > -INFINITY && lo < INFINITY) ? (uint64_t)ord_bits(lo) : 0ull;
    };
    auto admissible = [&](uint64_t r) -> bool { return r < p.n && (!p.mask || ((__ldg(p.mask + (r >> 5)) >> (r & 31)) & 1u)); };
    if (p.dtype == NK_DTYPE_F32 && p.metric != NK_METRIC_EUCLIDEAN) {
        // two rows per warp at a time: both rows' 128-bit loads are in flight together (the scoring is latency-bound:
        // one 4 KB row per warp per round trip)
        const float4 *q4 = reinterpret_cast<const float4 *>(qs);
        for (uint32_t i = 2 * w; i < S; i += 2 * NW) {
            const uint64_t r0 = sample_row(i), r1 = sample_row(i + 1);
            const bool ok0 = admissible(r0), ok1 = i + 1 < S && admissible(r1);
            const float4 *x0 = reinterpret_cast<const float4 *>(static_cast<const float *>(p.rows) + (ok0 ? r0 : 0) * p.dim);
            const float4 *x1 = reinterpret_cast<const float4 *>(static_cast<const float *>(p.rows) + (ok1 ? r1 : 0) * p.dim);
            float d0 = 0.f, d1 = 0.f, n0 = 0.f, n1 = 0.f;
#pragma unroll 4
            for (uint32_t j = lane; j < p.dim / 4; j += 32) {
                const float4 a = __ldg(x0 + j), b = __ldg(x1 + j), u = q4[j];
                d0 = fmaf(a.x, u.x, d0); n0 = fmaf(a.x, a.x, n0); d0 = fmaf(a.y, u.y, d0); n0 = fmaf(a.y, a.y, n0);
                d0 = fmaf(a.z, u.z, d0); n0 = fmaf(a.z, a.z, n0); d0 = fmaf(a.w, u.w, d0); n0 = fmaf(a.w, a.w, n0);
                d1 = fmaf(b.x, u.x, d1); n1 = fmaf(b.x, b.x, n1); d1 = fmaf(b.y, u.y, d1); n1 = fmaf(b.y, b.y, n1);
                d1 = fmaf(b.z, u.z, d1); n1 = fmaf(b.z, b.z, n1); d1 = fmaf(b.w, u.w, d1); n1 = fmaf(b.w, b.w, n1);
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                d0 += __shfl_xor_sync(0xffffffffu, d0, o); n0 += __shfl_xor_sync(0xffffffffu, n0, o);
                d1 += __shfl_xor_sync(0xffffffffu, d1, o); n1 += __shfl_xor_sync(0xffffffffu, n1, o);
            }
            if (lane == 0) {
                float s0 = d0, s1 = d1;
                if (p.metric == NK_METRIC_COSINE) {
                    const float e0 = sqrtf(n0 * t), e1 = sqrtf(n1 * t);
                    s0 = e0 > 0.0f ? d0 / e0 : 0.0f;
                    s1 = e1 > 0.0f ? d1 / e1 : 0.0f;
                }
                skeys[i] = ok0 ? lower_key(s0, n0) : 0ull;
                if (i + 1 < S) skeys[i + 1] = ok1 ? lower_key(s1, n1) : 0ull;
            }
        }
    } else {
        for (uint32_t i = w; i < S; i += NW) {
            uint64_t key = 0ull;
            const uint64_t r = sample_row(i);
            if (admissible(r)) {
                float xx;
                const float sc = warp_exact_score(p.rows, p.dtype, (size_t)r, p.dim, qs, t, p.metric, lane, &xx);
                key = lower_key(sc, xx);
            }
            if (lane == 0) skeys[i] = key;
        }
    }
    __syncthreads();
    // k-th largest of the S lower bounds by counting ranks (S <= 2048: one barrier instead of a bitonic network)
    if (p.k > S) return;
    for (uint32_t i = tid; i < S; i += PREP_THREADS) {
        const uint64_t mine = skeys[i];
        if (!mine) continue;
        uint32_t greater = 0, equal_before = 0;
        for (uint32_t j = 0; j < S; ++j) {
            const uint64_t o = skeys[j];
            greater += o > mine ? 1u : 0u;
            equal_before += (o == mine && j < i) ? 1u : 0u;
        }
        if (greater + equal_before == p.k - 1) {  // exactly one entry holds rank k-1 (ties ordered by position)
            const float tau0 = fmaxf(ord_to_float((uint32_t)mine), p.min_score);
            if (tau0 > -INFINITY) p.state[row] = ord_bits(tau0);
        }
    }
}

template <int NT, int QT, bool DUMP>
__global__ void __launch_bounds__(tc::THREADS, 1)
knn_scan_tc_kernel(const __grid_constant__ CUtensorMap map_rows, const __grid_constant__ CUtensorMap map_qhi,
                   const __grid_constant__ CUtensorMap map_qlo, tc::Params p) {
    using namespace tc;
    using C = Cfg<NT, QT>;
    constexpr bool FILTER = NT == 1;
    constexpr int A_COL = C::A_COL, B_BYTES = C::B_BYTES;
    pdl_trigger();
    pdl_wait();
    if (p.only_if && *p.only_if == 0) return;  // exact fallback not needed (uniform over the grid)

    extern __shared__ unsigned char smem_dyn[];
    // 128-byte-swizzled tiles need 1024-byte alignment: align by hand (the launch reserves the slack).
    unsigned char *smem_raw = smem_dyn + ((1024u - (ptx::smem_u32(smem_dyn) & 1023u)) & 1023u);
    // layout: [ASTAGES x A 32K] [BSTAGES x (Bhi 8K [| Blo 8K])] [Shared]
    unsigned char *a_base = smem_raw;
    unsigned char *b_base = smem_raw + (size_t)ASTAGES * A_BYTES;
    Shared &sh = *reinterpret_cast<Shared *>(smem_raw + (size_t)C::RING_BYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t num_tiles = (p.n + ROWS - 1) / ROWS;
    // Query groups: CTA b serves query block (b % G) over the tile subset (b / G).  The G sibling CTAs stream the SAME
    // corpus tiles at the same pace, so a tile crosses HBM once and its other G-1 readers hit the 126 MB L2: large
    // batches (Q > 128) stop being bound by re-reading the corpus once per 128 queries.
    const uint32_t grp = blockIdx.x % p.qgroups, sub = blockIdx.x / p.qgroups, sgrid = gridDim.x / p.qgroups;
    const uint32_t q0 = p.q0 + grp * QT, qpad_off = p.qpad_off + grp * QT;
    const uint32_t nq = p.nq - grp * QT < (uint32_t)QT ? p.nq - grp * QT : (uint32_t)QT;
    const uint64_t a_policy = p.qgroups > 1 ? ptx::CACHE_EVICT_NORMAL : ptx::CACHE_EVICT_FIRST;
    const bool prof = (p.debug & 64) && blockIdx.x == 0;
    long long acc_a = 0, acc_b = 0, acc_c = 0, acc_d = 0;
    const long long t_start = prof ? clock64() : 0;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_rows);
        ptx::prefetch_tensormap(&map_qhi);
        if (NT == 3) ptx::prefetch_tensormap(&map_qlo);
        for (int i = 0; i < ASTAGES; ++i) { ptx::mbar_init(&sh.afull_s[i], 1); ptx::mbar_init(&sh.aempty_s[i], 8); }
        for (int i = 0; i < C::BSTAGES; ++i) { ptx::mbar_init(&sh.bfull[i], 1); ptx::mbar_init(&sh.bempty[i], 2); }
        for (int i = 0; i < C::TSTAGES; ++i)
            for (int m = 0; m < 2; ++m) { ptx::mbar_init(&sh.afull[i][m], 4); ptx::mbar_init(&sh.aempty[i][m], 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&sh.accfull[i], 1); ptx::mbar_init(&sh.accempty[i], 4); }
        sh.maxxx = 0u;
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc(&sh.tmem_base, TMEM_COLS);
    if (tid < QT) {
        // start threshold: the caller's score floor and, when the prep kernel sampled the shard, a lower bound of the
        // query's k-th best score (no flood tiles then)
        float t0 = p.min_score;
        if (FILTER && p.presampled && (uint32_t)tid < nq) {
            const uint32_t g = __ldcg(p.gtau + q0 + tid);
            if (g) t0 = fmaxf(t0, ord_to_float(g));
        }
        sh.tau[tid] = t0;
        sh.cnt[tid] = 0;
        sh.qn[tid] = (FILTER && p.qnorm) ? p.qnorm[qpad_off + tid] : 1.0f;
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = sh.tmem_base;
    const bool flood = !(FILTER && p.presampled);  // first two tiles: every (row, query) pair is placed directly

    if (warp == 0) {
        // ===================================== TMA producer: corpus slabs ========================
        uint32_t g = 0;
        for (uint32_t tile = sub; tile < num_tiles; tile += sgrid) {
            for (uint32_t j = 0; j < p.nslab; ++j, ++g) {
                const uint32_t s = g % ASTAGES;
                { TC_PROF_BEGIN(); ptx::mbar_wait(&sh.aempty_s[s], ((g / ASTAGES) & 1) ^ 1); TC_PROF_END(a); }
                if (ptx::elect_one_sync()) {
                    ptx::mbar_arrive_expect_tx(&sh.afull_s[s], A_BYTES);
                    ptx::tma_load_2d(&map_rows, &sh.afull_s[s], a_base + (size_t)s * A_BYTES, (int32_t)(j * BK), (int32_t)(tile * ROWS), a_policy);
                }
                __syncwarp();
            }
        }
        if (prof && lane == 0) { g_tc_prof[0] = acc_a; g_tc_prof[1] = clock64() - t_start; }
    } else if (warp == 3) {
        // ===================================== TMA producer: query slabs (L2-resident) ===========
        uint32_t g = 0;
        for (uint32_t tile = sub; tile < num_tiles; tile += sgrid) {
            for (uint32_t j = 0; j < p.nslab; ++j, ++g) {
                const uint32_t s = g % C::BSTAGES;
                { TC_PROF_BEGIN(); ptx::mbar_wait(&sh.bempty[s], ((g / C::BSTAGES) & 1) ^ 1); TC_PROF_END(a); }
                if (ptx::elect_one_sync()) {
                    unsigned char *st = b_base + (size_t)s * C::BST_BYTES;
                    ptx::mbar_arrive_expect_tx(&sh.bfull[s], C::BST_BYTES);
                    ptx::tma_load_2d(&map_qhi, &sh.bfull[s], st, (int32_t)(j * BK), (int32_t)qpad_off, ptx::CACHE_EVICT_LAST);
                    if (NT == 3) ptx::tma_load_2d(&map_qlo, &sh.bfull[s], st + B_BYTES, (int32_t)(j * BK), (int32_t)qpad_off, ptx::CACHE_EVICT_LAST);
                }
                __syncwarp();
            }
        }
        if (prof && lane == 0) { g_tc_prof[2] = acc_a; }
    } else if (warp == 1 || warp == 2) {
        // ===================================== MMA issuers (one warp per M-tile) ==================
        const uint32_t m = warp - 1;
        const uint32_t idesc = ptx::make_idesc_tf32(128, QT);
        const uint32_t d = tmem + ACC_COL + m * QT;
        uint32_t g = 0, it = 0;
        for (uint32_t tile = sub; tile < num_tiles; tile += sgrid, ++it) {
            for (uint32_t j = 0; j < p.nslab; ++j, ++g) {
                const uint32_t s = g % C::BSTAGES, ts = g % C::TSTAGES;
                if (j == 0) { TC_PROF_BEGIN(); ptx::mbar_wait(&sh.accempty[m], (it & 1) ^ 1); TC_PROF_END(b); }  // epilogue drained
                { TC_PROF_BEGIN(); ptx::mbar_wait(&sh.bfull[s], (g / C::BSTAGES) & 1); TC_PROF_END(a); }  // query slabs landed
                { TC_PROF_BEGIN(); ptx::mbar_wait(&sh.afull[ts][m], (g / C::TSTAGES) & 1); TC_PROF_END(c); }  // A operand in TMEM
                ptx::tc_fence_after();
                if (ptx::elect_one_sync()) {
                    const uint32_t bhi = ptx::smem_u32(b_base + (size_t)s * C::BST_BYTES);
                    const uint64_t dhi = ptx::make_smem_desc_sw128(bhi);
                    const uint32_t ahi = tmem + A_COL + (ts * 2 + m) * C::PARTS * BK;
                    // K advance per MMA = 8 floats = 32 B = 2 descriptor units = 8 TMEM columns
                    if (NT == 3) {
                        const uint64_t dlo = ptx::make_smem_desc_sw128(bhi + B_BYTES);
                        const uint32_t alo = ahi + BK;
#pragma unroll
                        for (uint32_t kk = 0; kk < BK / 8; ++kk) {  // smallest terms first
                            ptx::mma_tf32_ts(d, alo + kk * 8, dhi + kk * 2, idesc, (j | kk) != 0);
                            ptx::mma_tf32_ts(d, ahi + kk * 8, dlo + kk * 2, idesc, 1);
                            ptx::mma_tf32_ts(d, ahi + kk * 8, dhi + kk * 2, idesc, 1);
                        }
                    } else {
#pragma unroll
                        for (uint32_t kk = 0; kk < BK / 8; ++kk) ptx::mma_tf32_ts(d, ahi + kk * 8, dhi + kk * 2, idesc, (j | kk) != 0);
                    }
                    ptx::tc_commit(&sh.aempty[ts][m]);                        // TMEM A slot of this M-tile reusable
                    ptx::tc_commit(&sh.bempty[s]);                            // query slabs: both issuers must be done
                    if (j + 1 == p.nslab) ptx::tc_commit(&sh.accfull[m]);     // accumulator of this M-tile complete
                }
                __syncwarp();
            }
        }
        if (prof && lane == 0 && warp == 1) { g_tc_prof[4] = acc_a; g_tc_prof[5] = acc_b; g_tc_prof[6] = acc_c; g_tc_prof[7] = clock64() - t_start; }
    } else if (warp >= SPLIT_WARP0 && warp < SPLIT_WARP0 + 8) {
        // ===================================== split warps ======================================
        const uint32_t m = (warp - SPLIT_WARP0) >> 2, quad = warp & 3;
        const uint32_t r = m * 128 + quad * 32 + lane;  // row within the tile
        const uint32_t lane_base = (quad * 32u) << 16;
        uint32_t g = 0, it = 0;
        for (uint32_t tile = sub; tile < num_tiles; tile += sgrid, ++it) {
            uint64_t xx2 = 0;  // two partial sums of |x|^2 (packed f32x2)
            for (uint32_t j = 0; j < p.nslab; ++j, ++g) {
                const uint32_t s = g % ASTAGES, ts = g % C::TSTAGES;
                { TC_PROF_BEGIN(); ptx::mbar_wait(&sh.afull_s[s], (g / ASTAGES) & 1); TC_PROF_END(a); }
                long long _tw = prof ? clock64() : 0;
                const unsigned char *rowp = a_base + (size_t)s * A_BYTES + (size_t)r * 128;
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    // 128B swizzle: logical 16-byte chunk c of row r sits at chunk c ^ (r & 7)
                    const uint4 v = *reinterpret_cast<const uint4 *>(rowp + ((c ^ (r & 7)) << 4));
                    const uint32_t e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int u = 0; u < 4; u += 2) {
                        const uint32_t h0 = ptx::tf32_round_bits(e[u]), h1 = ptx::tf32_round_bits(e[u + 1]);
                        const uint64_t x2 = ptx::pack2(e[u], e[u + 1]);
                        xx2 = ptx::fma_f32x2(x2, x2, xx2);  // |x|^2 on the side
                        hi[c * 4 + u] = h0; hi[c * 4 + u + 1] = h1;
                        if (NT == 3) {
                            const uint64_t l2 = ptx::sub_f32x2(x2, ptx::pack2(h0, h1));  // exact residual x - xh
                            lo[c * 4 + u] = (uint32_t)l2; lo[c * 4 + u + 1] = (uint32_t)(l2 >> 32);
                        }
                    }
                }
                // the slab now lives in registers: hand the smem slot straight back to the TMA producer
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&sh.aempty_s[s]);
                if (prof) acc_c += clock64() - _tw;
                { TC_PROF_BEGIN(); ptx::mbar_wait(&sh.aempty[ts][m], ((g / C::TSTAGES) & 1) ^ 1); TC_PROF_END(b); }
                _tw = prof ? clock64() : 0;
                ptx::tc_fence_after();
                const uint32_t acol = tmem + lane_base + A_COL + (ts * 2 + m) * C::PARTS * BK;
                ptx::tmem_st_32x32b_x32(acol, hi);
                if (NT == 3) ptx::tmem_st_32x32b_x32(acol + BK, lo);
                ptx::tmem_wait_st();
                if (j + 1 == p.nslab)  // published by the afull arrive below
                    sh.xx[it % XX_RING][r] = __uint_as_float((uint32_t)xx2) + __uint_as_float((uint32_t)(xx2 >> 32));
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&sh.afull[ts][m]);
                if (prof) acc_d += clock64() - _tw;
            }
        }
        if (prof && lane == 0 && warp == SPLIT_WARP0) { g_tc_prof[8] = acc_a; g_tc_prof[9] = acc_b; g_tc_prof[10] = acc_c; g_tc_prof[11] = acc_d; g_tc_prof[12] = clock64() - t_start; }
        if (prof && lane == 0 && warp == SPLIT_WARP0 + 4) { g_tc_prof[13] = acc_a; g_tc_prof[14] = acc_b; g_tc_prof[15] = acc_c; g_tc_prof[16] = acc_d; }
    } else if (warp >= EPI_WARP0) {
        // ===================================== epilogue =========================================
        const uint32_t quad = warp & 3;
        const uint32_t lane_base = (quad * 32u) << 16;
        uint64_t *my_cand = p.cand + (size_t)blockIdx.x * QT * P;
        const int prune_at = P - ROWS;
        const bool cosine = p.metric == NK_METRIC_COSINE, euclid = p.metric == NK_METRIC_EUCLIDEAN;
        const float bfac = p.margin_c * (euclid ? 2.0f : 1.0f);  // bound on -dist^2 = -(|x|^2+|q|^2-2x.q) is 2c|x||q|
        uint32_t it = 0;
        for (uint32_t tile = sub; tile < num_tiles; tile += sgrid, ++it) {
#pragma unroll 1
            for (uint32_t m = 0; m < 2; ++m) {
                const uint32_t rt = m * 128 + quad * 32 + lane;
                const uint32_t row = tile * ROWS + rt;
                const bool alive = row < p.n && (!p.mask || ((__ldg(p.mask + (row >> 5)) >> (row & 31)) & 1u));
                { TC_PROF_BEGIN(); ptx::mbar_wait(&sh.accfull[m], it & 1); TC_PROF_END(a); }
                ptx::tc_fence_after();
                // score(row, query c):
                //   cosine     acc / |x|  (queries pre-normalised)       filter bound c
                //   dot        acc                                        filter bound c |x| |q|
                //   euclidean  -(|x|^2 + |q|^2 - 2 acc)   (filter only)   filter bound 2c |x| |q|
                // filter mode buffers the UPPER bound score + bound.
                const float x2 = sh.xx[it % XX_RING][rt];
                const float xn = sqrtf(x2);
                float mul = 1.0f, bnd = 0.0f;
                if (cosine) mul = x2 > 0.0f ? 1.0f / xn : 0.0f;
                if (FILTER) {
                    bnd = cosine ? p.margin_c : bfac * xn;
                    if (!cosine && alive && x2 < INFINITY) atomicMax(&sh.maxxx, __float_as_uint(x2));  // NaN / Inf rows: kept, judged exactly
                    if (euclid) mul = 2.0f;
                }
#pragma unroll 1
                for (uint32_t half = 0; half < QT / 64; ++half) {
                    // drain 64 accumulator columns into registers; after the last half hand the accumulator back,
                    // then score from registers
                    const uint32_t cb = half * 64;
                    uint32_t v0[32], v1[32];
                    ptx::tmem_ld_32x32b_x32(tmem + lane_base + ACC_COL + m * QT + cb, v0);
                    ptx::tmem_ld_32x32b_x32(tmem + lane_base + ACC_COL + m * QT + cb + 32, v1);
                    ptx::tmem_wait_ld();
                    if (half + 1 == QT / 64) {
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&sh.accempty[m]);
                    }
                    if (DUMP && FILTER && row < p.n) {  // tests only: score estimate and error bound per (row, query)
                        for (uint32_t c = 0; c < 64 && cb + c < nq; ++c) {
                            uint32_t bits = 0;
#pragma unroll
                            for (uint32_t i = 0; i < 32; ++i) {
                                if (c == i) bits = v0[i];
                                if (c == 32 + i) bits = v1[i];
                            }
                            const float qn = sh.qn[cb + c];
                            float est = __uint_as_float(bits) * mul, b = bnd * qn;
                            if (euclid) { est -= fmaf(qn, qn, x2); b += EUC_EPS * fmaf(qn, qn, x2); }
                            p.dump_est[(size_t)row * p.dump_ld + q0 + cb + c] = est;
                            p.dump_bnd[(size_t)row * p.dump_ld + q0 + cb + c] = b;
                        }
                    }
                    if (flood && it < 2 && cb < nq) {
                        // Flood tiles: until the first prune (after this CTA's second tile) every threshold is -inf and
                        // EVERY (row, query) pair is buffered.  Place them directly - slot = tile-local row, no atomics,
                        // no register select - instead of 256 x QT trips through the rare-push loop (~60 us per launch).
                        const uint32_t slot = it * ROWS + rt;
                        const uint32_t grow = (uint32_t)(p.row_base + row);
#pragma unroll
                        for (uint32_t c = 0; c < 64; ++c) {
                            const uint32_t qi = cb + c;
                            if (qi < nq) {
                                float sc = __uint_as_float(c < 32 ? v0[c & 31] : v1[c & 31]) * mul;
                                if (FILTER) {
                                    const float qn = sh.qn[qi];
                                    if (euclid) sc -= EUC_KEEP * fmaf(qn, qn, x2);
                                    sc = fmaf(bnd, qn, sc);
                                    if (sc != sc) sc = INFINITY;
                                } else if (sc != sc) {
                                    sc = -INFINITY;
                                }
                                my_cand[(size_t)qi * P + slot] = (alive && sc >= p.min_score) ? make_key(sc, grow) : 0ull;  // 0 = empty slot
                            }
                        }
                        if (rt == 0 && half == 0)
                            for (uint32_t qi = m == 0 ? 0 : nq; qi < nq; ++qi) sh.cnt[qi] = (int)((it + 1) * ROWS);
                    } else if (alive && cb < nq) {
                        // Compact compare pass -> 64-bit mask of columns worth buffering (NaN passes); the rare pushes
                        // run in a small out-of-line loop so the hot code stays a few hundred instructions (a fully
                        // unrolled push per column was ~40 KB of SASS: I-cache thrash).
                        uint32_t pass0 = 0, pass1 = 0;
#pragma unroll
                        for (uint32_t c = 0; c < 32; ++c) {
                            float s0 = __uint_as_float(v0[c]) * mul, s1 = __uint_as_float(v1[c]) * mul;
                            if (FILTER) {
                                const float q0n = sh.qn[cb + c], q1n = sh.qn[cb + 32 + c];
                                if (euclid) { s0 -= EUC_KEEP * fmaf(q0n, q0n, x2); s1 -= EUC_KEEP * fmaf(q1n, q1n, x2); }
                                s0 = fmaf(bnd, q0n, s0);
                                s1 = fmaf(bnd, q1n, s1);
                            }
                            pass0 |= !(s0 < sh.tau[cb + c]) ? (1u << c) : 0u;
                            pass1 |= !(s1 < sh.tau[cb + 32 + c]) ? (1u << c) : 0u;
                        }
                        uint64_t pass = (uint64_t)pass0 | ((uint64_t)pass1 << 32);
                        if (nq - cb < 64) pass &= (1ull << (nq - cb)) - 1ull;
#pragma unroll 1
                        while (pass) {
                            const uint32_t c = (uint32_t)__ffsll((long long)pass) - 1u;
                            pass &= pass - 1ull;
                            uint32_t bits = 0;
#pragma unroll
                            for (uint32_t i = 0; i < 32; ++i) {  // register file is not indexable: select by compare
                                if (c == i) bits = v0[i];
                                if (c == 32 + i) bits = v1[i];
                            }
                            const uint32_t qi = cb + c;
                            float sc = __uint_as_float(bits) * mul;
                            if (FILTER) {
                                const float qn = sh.qn[qi];
                                if (euclid) sc -= EUC_KEEP * fmaf(qn, qn, x2);
                                sc = fmaf(bnd, qn, sc);
                                if (sc != sc) sc = INFINITY;  // undecidable here: keep it, the exact rescoring judges
                            } else if (sc != sc) {
                                sc = -INFINITY;
                            }
                            if (sc >= sh.tau[qi]) {
                                int pos = atomicAdd(&sh.cnt[qi], 1);
                                if (pos < P) my_cand[(size_t)qi * P + pos] = make_key(sc, (uint32_t)(p.row_base + row));
                                else atomicExch(p.flags, 1);
                            }
                        }
                    }
                }
            }
            // prune any buffer that could overflow during the next tile; the 4 epilogue warps prune different
            // queries concurrently with a register-resident warp selection
            group_sync(EPI_BAR, EPI_THREADS);  // every push of this tile is visible
            for (uint32_t qi = quad; qi < nq; qi += 4)
                if (sh.cnt[qi] > prune_at) {
                    const float margin2 = !FILTER ? 0.0f : filter_margin2(p.metric, p.margin_c, __uint_as_float(sh.maxxx), sh.qn[qi]);
                    float floor_tau = p.min_score;
                    if (FILTER) {
                        const uint32_t g = __ldcg(p.gtau + q0 + qi);
                        if (g) floor_tau = fmaxf(floor_tau, ord_to_float(g));
                    }
                    warp_prune<16>(my_cand + (size_t)qi * P, &sh.cnt[qi], &sh.tau[qi], p.k, lane, nullptr, 0, FILTER, margin2, prune_at, floor_tau,
                                   nullptr, p.flags + FLAG_OVERFLOW);
                    // everything inside the margin must fit below prune_at, or the next tile could overflow the buffer
                    if (FILTER && lane == 0 && sh.cnt[qi] >= prune_at) atomicOr(p.flags + 1, 1);
                    // publish: this CTA's threshold is a lower bound on the true global k-th best score, so every CTA
                    // may filter with the largest one any CTA has found
                    if (FILTER && lane == 0 && sh.tau[qi] > -INFINITY) atomicMax(p.gtau + q0 + qi, ord_bits(sh.tau[qi]));
                }
            group_sync(EPI_BAR, EPI_THREADS);
            if (FILTER) {  // adopt the shared thresholds (one L2 read per query per tile)
                for (uint32_t qi = tid - EPI_WARP0 * 32; qi < nq; qi += EPI_THREADS) {
                    const uint32_t g = __ldcg(p.gtau + q0 + qi);
                    if (g) sh.tau[qi] = fmaxf(sh.tau[qi], ord_to_float(g));
                }
            }
        }
    }

    // ---- emit this CTA's list per query: best k (exact) or everything inside the margin (filter).  Every role has
    // finished its tile loop here, so all 16 warps share the final prunes (4x shorter tail than the epilogue warps alone).
    __syncthreads();
    {
        const bool cosine = p.metric == NK_METRIC_COSINE;
        uint64_t *my_cand = p.cand + (size_t)blockIdx.x * QT * P;
        for (uint32_t qi = warp; qi < nq; qi += THREADS / 32) {
            const float margin2 = !FILTER ? 0.0f : filter_margin2(p.metric, p.margin_c, __uint_as_float(sh.maxxx), sh.qn[qi]);
            if (FILTER) {
                // survivors (inside this CTA's margin AND above the shared threshold) go to the query's shared list
                float floor_tau = p.min_score;
                const uint32_t g = __ldcg(p.gtau + q0 + qi);
                if (g) floor_tau = fmaxf(floor_tau, ord_to_float(g));
                if (sh.cnt[qi] <= (int)p.k_emit) {
                    // few entries: no local selection, the finish kernel selects globally — append what still reaches the
                    // current threshold
                    const float t = fmaxf(sh.tau[qi], floor_tau);
                    uint64_t thr = t > -INFINITY ? (uint64_t)ord_bits(t) << 32 : 1ull;
                    if (thr == 0ull) thr = 1ull;
                    warp_emit_above(my_cand + (size_t)qi * P, sh.cnt[qi], thr, lane, p.partial + (size_t)(q0 + qi) * p.list_cap,
                                    (int)p.list_cap, p.gcount + q0 + qi);
                    continue;
                }
                warp_prune<16>(my_cand + (size_t)qi * P, &sh.cnt[qi], &sh.tau[qi], p.k, lane,
                               p.partial + (size_t)(q0 + qi) * p.list_cap, (int)p.list_cap, true, margin2,
                               (int)p.k_emit, floor_tau, p.gcount + q0 + qi, p.flags + FLAG_OVERFLOW);
                // the list was cut at k_emit while rows inside the margin remained -> exact fallback
                if (lane == 0 && sh.cnt[qi] >= (int)p.k_emit && (int)p.k_emit > (int)p.k) atomicOr(p.flags + FLAG_OVERFLOW, 2);
                if (lane == 0 && sh.tau[qi] > -INFINITY) atomicMax(p.gtau + q0 + qi, ord_bits(sh.tau[qi]));
            } else {
                warp_prune<16>(my_cand + (size_t)qi * P, &sh.cnt[qi], &sh.tau[qi], p.k, lane,
                               p.partial + ((size_t)(q0 + qi) * gridDim.x + blockIdx.x) * p.k_emit, (int)p.k_emit, false, 0.0f, (int)p.k_emit,
                               p.min_score);
            }
        }
        if (FILTER && !cosine && tid == 0) atomicMax(reinterpret_cast<unsigned int *>(p.flags + 2), sh.maxxx);
    }

    if (prof && tid == EPI_WARP0 * 32) { g_tc_prof[17] = acc_a; g_tc_prof[18] = clock64() - t_start; g_tc_prof[19] = (long long)num_tiles; }
    // ---- teardown ----------------------------------------------------------------------------------
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------
// Filter-mode finish (one CTA per query): the query's shared list holds every row whose upper bound reached the
// cross-CTA threshold (k plus a few dozen).  Radix-select the k-th largest bound; everything with bound >= (k-th bound -
// 2*Bmax) may belong to the true top-k: those rows are re-scored EXACTLY in fp32 with the same arithmetic as the
// CUDA-core scan (dot / sqrt(|x|^2 |q|^2) etc.), sorted by (score desc, row asc), and the best k written out — as keys
// and, when the caller wants them, already decoded (index, score), so no separate decode launch follows.  Survivor
// sets of any size are handled in rounds of FINISH_WIN list entries (near-tie data: thousands of rows inside the margin)
// with the running best k carried from round to round.
// The LAST CTA to finish does the stage bookkeeping that used to be two more launches: it turns the stage's overflow
// flag into the retry marker and, if a retry is due, wipes the shared thresholds and list fills.
// ---------------------------------------------------------------------------------------------------
constexpr int FINISH_THREADS = 1024;  // one CTA per query has an SM to itself: 32 warps for the list passes and the re-scoring
constexpr int FINISH_CAP = 4096;
constexpr int FINISH_WIN = FINISH_CAP - 256;  // list entries per round; the carried best k (<= 192) fits in the rest
constexpr int FINISH_HI = 16384;              // score words of the list cached in shared memory (longer lists: read from L2)
constexpr uint32_t ORD_POS_INF = 0xFF800000u;  // ord_bits(+inf): the bound of a row whose score is undecidable (NaN)
struct FinishParams {
    const void *rows;        // corpus shard (fp32, or fp16 for the fp16 tensor path)
    int dtype;
    uint32_t dim;
    uint64_t row_base;
    const float *queries;    // raw fp32 queries [Q x dim]
    const uint64_t *lists;   // [Q][list_cap] shared append lists (keys carry upper bounds)
    const int *gcount;       // [Q]
    uint32_t list_cap, k;
    int metric;
    float margin_c;          // TF32 passes: c in |s_hat - s| <= c |x||q|
    uint32_t q_big;          // queries [0, q_big) went through the BF16 / FP16 kernel: bound factors qa / qb / qn below
    const float *qa, *qb, *qn;
    int *flags;
    const int *only_if;      // retry stage: run only if *only_if != 0
    uint64_t *out;           // [Q][k] keys
    uint32_t *out_idx;       // optional fused decode (nullable): [Q][k]
    float *out_score;
    float min_score;         // caller's score floor: exact scores below it are dropped
    int mark_retry;          // first stage of a two-stage filter: overflow -> retry marker + state wipe
    uint32_t *state;         // gtau[QA] + gcount[QA]
    uint32_t state_words;
    int debug;               // NK_TC_DEBUG & 64: per-CTA phase stamps (g_fin_prof)
};

// Debug timeline of the finish kernel (NK_TC_DEBUG & 64), printed after the first-stage finish launch:
//   [0] entry  [1] list cached  [2] k-th bound selected  [3] survivors gathered  [4] re-scored  [5] ranked  [6] exit  [7] list length
__device__ unsigned long long g_fin_prof[256][8];
__device__ __forceinline__ unsigned long long fin_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    return t;
}

__global__ void __launch_bounds__(FINISH_THREADS) filter_finish_kernel(FinishParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *shi = reinterpret_cast<uint32_t *>(smem_raw);                                   // FINISH_HI score words of the list
    uint64_t *sb = reinterpret_cast<uint64_t *>(smem_raw + FINISH_HI * 4);                    // FINISH_CAP bound keys
    uint64_t *se = reinterpret_cast<uint64_t *>(smem_raw + FINISH_HI * 4 + FINISH_CAP * 8);   // FINISH_CAP exact keys
    float *qs = reinterpret_cast<float *>(smem_raw + FINISH_HI * 4 + FINISH_CAP * 16);        // query
    __shared__ float s_qq;
    __shared__ int s_count, s_last;
    pdl_trigger();
    pdl_wait();  // launched as a programmatic dependent of the scan (common.cuh)
    if (p.only_if && *p.only_if == 0) return;
    const uint32_t q = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool prof = p.debug && q < 256 && tid == 0;
    if (prof) g_fin_prof[q][0] = fin_now();
    for (uint32_t j = tid; j < p.dim; j += FINISH_THREADS) qs[j] = p.queries[(size_t)q * p.dim + j];
    int n = p.gcount[q];
    if (tid == 0) atomicMax(p.flags + FLAG_LONGEST, n);  // diagnostics: longest shared list of this search
    if (n > (int)p.list_cap) n = (int)p.list_cap;
    const uint64_t *list = p.lists + (size_t)q * p.list_cap;
    // ONE pass over the list (L2): the score words go to shared memory — every selection pass below reads them from there
    // (the list was read five times from L2, each time a chain of dependent ~1 us loads: that, not the arithmetic, was the
    // finish step's cost).  Four independent loads per thread per iteration.
    const bool cached = n <= FINISH_HI;
    uint32_t umax = 0u, umin = 0xffffffffu;
    int ninf = 0;
    for (int i0 = tid; i0 < n; i0 += 4 * FINISH_THREADS) {
        uint32_t h[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * FINISH_THREADS;
            h[u] = i < n ? (uint32_t)(__ldcg(list + i) >> 32) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * FINISH_THREADS;
            if (i < n) {
                if (cached) shi[i] = h[u];
                umax = max(umax, h[u]);
                umin = min(umin, h[u]);
                ninf += h[u] >= ORD_POS_INF ? 1 : 0;
            }
        }
    }
    auto hi_at = [&](int i) -> uint32_t { return cached ? shi[i] : (uint32_t)(__ldcg(list + i) >> 32); };
    __syncthreads();
    if (prof) { g_fin_prof[q][1] = fin_now(); g_fin_prof[q][7] = (unsigned long long)n; }
    if (warp == 0) {
        float a = 0.0f;
        for (uint32_t j = lane; j < p.dim; j += 32) a = fmaf(qs[j], qs[j], a);
#pragma unroll
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) s_qq = a;
    }
    // k-th largest bound by radix select over the score bits that actually vary (8 bits per pass, smem histogram), then
    // everything inside the margin below it is gathered for exact re-scoring.
    __shared__ int hist[256];
    __shared__ uint32_t s_prefix, s_red[3][FINISH_THREADS / 32];
    __shared__ int s_krem;
    // Rows with an undecidable score (NaN -> bound +inf) are always kept but say nothing about the k-th best score: the
    // threshold comes from the k-th largest FINITE bound = the (k + n_inf)-th largest overall.
    umax = __reduce_max_sync(0xffffffffu, umax);
    umin = __reduce_min_sync(0xffffffffu, umin);
    ninf = __reduce_add_sync(0xffffffffu, ninf);
    if (lane == 0) { s_red[0][warp] = umax; s_red[1][warp] = umin; s_red[2][warp] = (uint32_t)ninf; }
    if (tid == 0) s_count = 0;
    __syncthreads();
    const float maxxx = __uint_as_float((unsigned int)p.flags[FLAG_MAXXX]);
    const float margin2 = q < p.q_big ? bf16_margin2(p.metric, __uint_as_float((unsigned int)p.flags[FLAG_MAX_RA]),
                                                     __uint_as_float((unsigned int)p.flags[FLAG_MAX_RB]), maxxx, p.qa[q], p.qb[q], p.qn[q])
                                      : filter_margin2(p.metric, p.margin_c, maxxx, sqrtf(s_qq));
    ninf = 0;
#pragma unroll
    for (int w = 0; w < FINISH_THREADS / 32; ++w) { umax = max(umax, s_red[0][w]); umin = min(umin, s_red[1][w]); ninf += (int)s_red[2][w]; }
    const int k_eff = (int)p.k + ninf;
    if (tid == 0) s_krem = k_eff;
    uint64_t thr_key = 1ull;  // fewer than k decidable entries: keep them all
    if (n >= k_eff) {
        int rem = 32 - __clz(umax ^ umin);  // low bits in which the bounds differ (0: all equal)
        if (tid == 0) s_prefix = rem >= 32 ? 0u : (umax >> rem) << rem;
        __syncthreads();
        // The select stops with <= 10 low bits unresolved: the prefix (low bits zero) is then a LOWER bound of the k-th largest
        // bound, at most 2^10 ulps (~1e-4 relative) below it — far inside the filter margin subtracted next, so at most a
        // stray extra survivor is re-scored, and two of the four histogram passes (3 barriers of 1024 threads each) are gone.
        while (rem > 10) {
            const int w = rem < 8 ? rem : 8, shift = rem - w;
            if (tid < 256) hist[tid] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            for (int i = tid; i < n; i += FINISH_THREADS) {
                const uint32_t hi = hi_at(i);
                if (rem >= 32 || (hi >> rem) == (prefix >> rem)) atomicAdd(&hist[(hi >> shift) & ((1u << w) - 1u)], 1);
            }
            __syncthreads();
            if (warp == 0) {  // digit holding the krem-th largest: lane l owns the 8 digits below nb - 8l, descending
                const int krem = s_krem, nb = 1 << w;
                int loc[8], sum = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int dgt = nb - 1 - (lane * 8 + j);
                    loc[j] = dgt >= 0 ? hist[dgt] : 0;
                    sum += loc[j];
                }
                __syncwarp();  // every lane has read s_krem / hist before the owner lane rewrites s_krem below
                int incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                int c = incl - sum;
                if (c < krem && incl >= krem) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (c + loc[j] >= krem) {
                            s_prefix = prefix | ((uint32_t)(nb - 1 - (lane * 8 + j)) << shift);
                            s_krem = krem - c;
                            break;
                        }
                        c += loc[j];
                    }
                }
            }
            __syncthreads();
            rem = shift;
        }
        __syncthreads();
        const float thr = ord_to_float(s_prefix) - margin2;
        if (thr < INFINITY) {
            thr_key = (uint64_t)ord_bits(thr) << 32;
            if (thr_key == 0ull) thr_key = 1ull;
        } else if (tid == 0) {
            // k-th bound +inf / NaN (k or more NaN rows) or a non-finite margin: no threshold can be trusted -> keep every
            // listed row here and let the exact stage redo the search
            atomicOr(p.flags + FLAG_OVERFLOW, 8);
        }
    }
    // ---- how many rows are inside the margin?  Almost always a few dozen: ONE gather + re-score round.  Only adversarial
    // near-tie data (thousands of survivors) goes through the list in windows, the running best k carried over.
    const uint32_t thr_hi = (uint32_t)(thr_key >> 32);  // keys >= thr_key  <=>  score word >= thr_hi (thr_key's low word is 0 or 1)
    if (prof) g_fin_prof[q][2] = fin_now();
    int mine = 0;
    for (int i = tid; i < n; i += FINISH_THREADS) mine += hi_at(i) >= thr_hi ? 1 : 0;
    mine = __reduce_add_sync(0xffffffffu, mine);
    if (lane == 0 && mine) atomicAdd(&s_count, mine);
    __syncthreads();
    const int survivors = s_count;
    const int win = survivors <= FINISH_WIN ? (n > 0 ? n : 1) : FINISH_WIN;  // whole list in one round when the survivors fit
    int carry = 0;
    for (int w0 = 0; w0 < n || w0 == 0; w0 += win) {
        __syncthreads();
        if (tid == 0) s_count = 0;
        __syncthreads();
        const int w1 = min(n, w0 + win);
        for (int i = w0 + tid; i < w1; i += FINISH_THREADS) {
            if (hi_at(i) >= thr_hi) {  // the full key is fetched only for survivors
                const uint64_t key = __ldcg(list + i);
                if (key >= thr_key) sb[atomicAdd(&s_count, 1)] = key;
            }
        }
        __syncthreads();
        if (prof && w0 == 0) g_fin_prof[q][3] = fin_now();
        const int count = s_count;  // <= FINISH_WIN
        for (int i = warp; i < count; i += FINISH_THREADS / 32) {
            const uint32_t grow = key_row(sb[i]);
            const size_t local = (size_t)(grow - (uint32_t)p.row_base);
            const float sc = warp_exact_score(p.rows, p.dtype, local, p.dim, qs, s_qq, p.metric, lane);
            if (lane == 0) se[carry + i] = sc >= p.min_score ? make_key(sc, grow) : 0ull;
        }
        const int total = carry + count;
        __syncthreads();
        if (prof && w0 == 0) g_fin_prof[q][4] = fin_now();
        if (total <= FINISH_THREADS) {
            // small sets (the common case): rank by counting — one barrier instead of a 15-45 stage bitonic network
            const uint64_t mykey = tid < total ? se[tid] : 0ull;
            int rank = 0;
            for (int j = 0; j < total; ++j) rank += se[j] > mykey ? 1 : 0;  // keys are unique (row in the low word) or 0
            __syncthreads();
            if (tid < total && mykey) se[rank] = mykey;
            // dropped (0) keys: fill the tail behind the live ones
            int live = __syncthreads_count(tid < total && mykey != 0ull);
            for (int i = live + tid; i < total; i += FINISH_THREADS) se[i] = 0ull;
            __syncthreads();
        } else {
            int P2 = 32;
            while (P2 < total) P2 <<= 1;
            for (int i = total + tid; i < P2; i += FINISH_THREADS) se[i] = 0ull;
            block_bitonic_sort_desc(se, P2);
        }
        carry = total < (int)p.k ? total : (int)p.k;
    }
    if (prof) g_fin_prof[q][5] = fin_now();
    for (uint32_t i = tid; i < p.k; i += FINISH_THREADS) {
        const uint64_t key = (int)i < carry ? se[i] : 0ull;
        p.out[(size_t)q * p.k + i] = key;
        if (p.out_idx) {
            float sc = key_score(key);
            if (p.metric == NK_METRIC_EUCLIDEAN) sc = sqrtf(fmaxf(-sc, 0.0f));
            p.out_idx[(size_t)q * p.k + i] = key ? key_row(key) : 0xffffffffu;
            p.out_score[(size_t)q * p.k + i] = key ? sc : 0.0f;
        }
    }
    // ---- stage bookkeeping by the last CTA
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(p.flags + FLAG_FINISH_CTAS, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (prof) g_fin_prof[q][6] = fin_now();
    if (!s_last) return;
    __threadfence();
    const int ovf = atomicOr(p.flags + FLAG_OVERFLOW, 0);
    if (tid == 0 && ovf) atomicOr(p.flags + FLAG_OVF_BITS, ovf);
    if (p.mark_retry) {
        if (ovf)
            for (uint32_t i = tid; i < p.state_words; i += FINISH_THREADS) p.state[i] = 0u;
        if (tid == 0) {
            p.flags[FLAG_RETRY] = ovf != 0;
            p.flags[FLAG_OVERFLOW] = 0;
            if (ovf) atomicAdd(p.flags + FLAG_N_RETRY, 1);
        }
    } else if (tid == 0 && ovf) {
        atomicAdd(p.flags + FLAG_N_EXACT, 1);
    }
    if (tid == 0) p.flags[FLAG_FINISH_CTAS] = 0;
}

// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) {
        cudaGetLastError();
        return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

int tc_make_map(CUtensorMap *m, const void *base, uint64_t rows, uint32_t cols, uint32_t elem_bytes, uint32_t box_cols,
                uint32_t box_rows, uint64_t row_stride_bytes, int map_dtype) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return -1;
    }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {(cuuint64_t)row_stride_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                   : map_dtype == NK_DTYPE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUresult r = enc(m, dt, 2,
                     const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%u", (int)r, (unsigned long long)rows, cols);
        return -1;
    }
    return 0;
}
// Tensor maps are cached per shard: cuTensorMapEncodeTiled (1-2 us of host time each, three per launch) runs only when
// the base pointer or the shape changed since the previous search.
const CUtensorMap *tc_cached_map(Workspace &ws, int slot, const void *base, uint64_t rows, uint32_t cols, uint32_t elem_bytes,
                                 uint32_t box_cols, uint32_t box_rows, uint64_t row_stride_bytes, int map_dtype) {
    static_assert(sizeof(CUtensorMap) == sizeof(ws.maps[0].bytes), "CUtensorMap is 128 bytes");
    Workspace::MapSlot &m = ws.maps[slot];
    const uint64_t key[6] = {reinterpret_cast<uint64_t>(base), rows, ((uint64_t)cols << 32) | elem_bytes,
                             ((uint64_t)box_cols << 32) | box_rows, row_stride_bytes, (uint64_t)map_dtype};
    if (!m.valid || memcmp(m.key, key, sizeof(key)) != 0) {
        if (tc_make_map(reinterpret_cast<CUtensorMap *>(m.bytes), base, rows, cols, elem_bytes, box_cols, box_rows, row_stride_bytes, map_dtype))
            return nullptr;
        memcpy(m.key, key, sizeof(key));
        m.valid = true;
    }
    return reinterpret_cast<const CUtensorMap *>(m.bytes);
}

// getenv once per variable (the search path used to call getenv per search)
int tc_env_int(const char *name, int dflt) {
    static std::mutex mu;
    static std::map<std::string, int> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(name);
    if (it != cache.end()) return it->second;
    const char *e = getenv(name);
    const int v = e ? atoi(e) : dflt;
    cache[name] = v;
    return v;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device) instead of once per launch
int tc_ensure_smem(const void *func, int device, size_t smem) {
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, size_t> done;
    std::lock_guard<std::mutex> lk(mu);
    size_t &cur = done[{func, device}];
    if (cur >= smem && cur != 0) return 0;
    NK_CUDA_OK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cur = smem;
    return 0;
}

static bool tc_common_ok(const DeviceInfo &di, const ScanArgs &a) {
    if (di.cc < 100) return false;
    if (a.dtype != NK_DTYPE_F32) return false;                      // fp16 / bf16 corpus: 16-bit tensor pass or CUDA cores
    if (a.dim % 4 != 0 || a.dim < 32) return false;                 // TMA: 16-byte global stride
    if ((reinterpret_cast<uintptr_t>(a.rows) & 15) != 0) return false;
    if (a.n == 0 || a.k == 0) return false;
    return true;
}
// exact (3xTF32) mode: cosine / dot, k <= 255 (register-resident prune over a 512-slot buffer)
bool scan_tensor_supported(const DeviceInfo &di, const ScanArgs &a) {
    return tc_common_ok(di, a) && a.metric != NK_METRIC_EUCLIDEAN && a.k + tc::ROWS + 1 <= (uint32_t)tc::P;
}
// filter mode (1xTF32 over fp32 rows, or the 16-bit pass over a shadow / a 16-bit corpus) + exact rescoring: all three
// metrics, k <= 192
bool scan_tensor_filter_supported(const DeviceInfo &di, const ScanArgs &a) {
    if (a.k > 192 || a.dim > 32768) return false;  // finish kernel keeps the query in shared memory
    return tc_common_ok(di, a) || shadow_pass_supported(di, a);
}

int tc_debug_flags() { return tc_env_int("NK_TC_DEBUG", 0); }

// One launch: queries [q0, q0+nq) against the whole shard, QT = 64 or 128 query columns per MMA.
struct TcPassArgs {
    uint32_t grid, k_emit, Qpad, q0, nq, qgroups;
    float margin_c;
    const float *qhi, *qlo, *qnorm;
    const int *only_if;
    bool count_main;
    int presampled;
    float *dump_est = nullptr, *dump_bnd = nullptr;
    uint32_t dump_ld = 0;
};
template <int NT, int QT, bool DUMP = false>
static int launch_pass(const DeviceInfo &di, const ScanArgs &a, Workspace &ws, const TcPassArgs &t, uint64_t *launches) {
    using namespace tc;
    const CUtensorMap *map_rows = tc_cached_map(ws, 0, a.rows, a.n, a.dim, 4, (uint32_t)BK, ROWS, (uint64_t)a.dim * 4);
    // query rows past Qpad are zero-filled by TMA
    const CUtensorMap *map_qhi = tc_cached_map(ws, QT == 64 ? 1 : 2, t.qhi, t.Qpad, a.dim, 4, (uint32_t)BK, QT, (uint64_t)a.dim * 4);
    const CUtensorMap *map_qlo = NT == 3 ? tc_cached_map(ws, 3, t.qlo, t.Qpad, a.dim, 4, (uint32_t)BK, QT, (uint64_t)a.dim * 4) : map_qhi;
    if (!map_rows || !map_qhi || !map_qlo) return -1;
    const size_t smem = (size_t)Cfg<NT, QT>::RING_BYTES + sizeof(Shared) + 1024;
    if (smem > di.max_smem_optin) {
        set_error("tensor path needs %zu B shared memory (> %zu)", smem, di.max_smem_optin);
        return -1;
    }
    if (tc_ensure_smem(reinterpret_cast<const void *>(knn_scan_tc_kernel<NT, QT, DUMP>), di.device_id, smem)) return -1;
    Params p{};
    p.n = a.n; p.dim = a.dim; p.nslab = (a.dim + BK - 1) / BK; p.row_base = a.row_base;
    p.q0 = t.q0; p.nq = t.nq; p.k = a.k; p.qpad_off = t.q0; p.qgroups = t.qgroups; p.list_cap = t.grid * t.k_emit;
    p.metric = a.metric; p.k_emit = t.k_emit; p.margin_c = t.margin_c; p.qnorm = t.qnorm;
    p.cand = ws.cand; p.partial = ws.partial; p.flags = ws.flags; p.only_if = t.only_if; p.debug = tc_debug_flags(); p.mask = a.row_mask;
    p.gtau = reinterpret_cast<uint32_t *>(ws.keys2); p.gcount = reinterpret_cast<int *>(ws.keys2) + (t.Qpad + QT_BIG);
    p.presampled = t.presampled; p.min_score = a.min_score;
    p.dump_est = t.dump_est; p.dump_bnd = t.dump_bnd; p.dump_ld = t.dump_ld;
    // (asynchronous API: the early-exit retry / exact stages are programmatic dependents — their launch overlaps the kernel
    // before them.  Not for the host-synchronous API: measured 5-15 us slower per nk_search, the copy-out behind it waits longer)
    NK_CUDA_OK(launch_pdl(knn_scan_tc_kernel<NT, QT, DUMP>, dim3(t.grid), dim3(THREADS), smem, a.stream, t.only_if != nullptr && !a.defer_tail, *map_rows, *map_qhi,
                          *map_qlo, p));
    if (launches) ++*launches;
    if (t.count_main && a.main_launches) ++*a.main_launches;
    return 0;
}

static void fin_print_prof(cudaStream_t stream, uint32_t Q) {
    if (!(tc_debug_flags() & 64)) return;
    static unsigned long long h[256][8];
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(h, g_fin_prof, sizeof(h));
    const uint32_t g = Q < 256 ? Q : 256;
    unsigned long long t0 = ~0ull, tmax = 0;
    for (uint32_t b = 0; b < g; ++b) { t0 = h[b][0] < t0 ? h[b][0] : t0; tmax = h[b][6] > tmax ? h[b][6] : tmax; }
    double a[8] = {0};
    for (uint32_t b = 0; b < g; ++b) {
        for (int i = 1; i < 7; ++i) a[i] += (double)(h[b][i] - h[b][i - 1]);
        a[0] += (double)(h[b][0] - t0);
        a[7] += (double)h[b][7];
    }
    fprintf(stderr, "[finish prof, %u CTAs] span %.1f us | entry +%.1f | list cached %.1f | k-th bound %.1f | gather %.1f | re-score %.1f | rank %.1f | write + count %.1f | "
            "list length %.0f\n", g, (tmax - t0) / 1e3, a[0] / g / 1e3, a[1] / g / 1e3, a[2] / g / 1e3, a[3] / g / 1e3, a[4] / g / 1e3, a[5] / g / 1e3, a[6] / g / 1e3, a[7] / g);
}

static void tc_print_prof(cudaStream_t stream, uint32_t num_tiles, uint32_t grid, uint32_t nslab) {
    if (!(tc_debug_flags() & 64)) return;
    long long h[32];
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(h, g_tc_prof, sizeof(h));
    uint32_t slabs = ((num_tiles + grid - 1) / grid) * nslab;
    fprintf(stderr, "[tc prof CTA0, ~%u slabs] total %lld cyc (%.0f/slab)\n  tmaA wait aempty_s %lld | tmaB wait bempty %lld\n"
            "  mma: wait bfull %lld, wait accempty %lld, wait afull %lld, total %lld\n"
            "  split m0: wait afull_s %lld, wait aempty %lld, read+math %lld, st+arrive %lld, total %lld\n"
            "  split m1: wait afull_s %lld, wait aempty %lld, read+math %lld, st+arrive %lld\n  epi: wait accfull %lld total %lld\n",
            slabs, h[7], (double)h[7] / slabs, h[0], h[2], h[4], h[5], h[6], h[7], h[8], h[9], h[10], h[11], h[12], h[13], h[14], h[15], h[16], h[17], h[18]);
}

// ---- the plan of one filter search: everything both the primary stage and the (possibly deferred) tail derive from
// (device, args).  Deterministic, so the host-driven retry of nk_search can rebuild it after the fact.
struct FilterPlan {
    uint32_t Qpad, QA, num_tiles, grid, k_emit, dimpad, sample;
    bool pair;          // batches of >= 256 queries run on CTA pairs (scan_tensor_pair.cu)
    bool big;           // first stage = 16-bit pass (BF16 shadow of an fp32 shard, or an fp16 / bf16 corpus itself)
    bool stage2;        // the TF32 filter over the fp32 rows exists as the retry stage (fp32 shards with a shadow)
    bool can_exact_tc;  // exact stage = 3xTF32 kernel (else the CUDA-core scan)
    float acc_c, margin_tf32;
    float *qhi, *qlo, *qnorm, *qa, *qb;
    void *qbf16;
    uint32_t max_groups, max_groups_shadow;
    size_t fsmem, psmem;
};

static int make_filter_plan(const DeviceInfo &di, const ScanArgs &a, Workspace &ws, FilterPlan *fp) {
    using namespace tc;
    FilterPlan &f = *fp;
    f.Qpad = (a.Q + 63) / 64 * 64;
    f.QA = f.Qpad + QT_BIG;  // padded per-query arrays
    f.num_tiles = (a.n + ROWS - 1) / ROWS;
    f.grid = (uint32_t)di.num_sms < f.num_tiles ? (uint32_t)di.num_sms : f.num_tiles;
    // per-CTA contribution: at most k + room for the rows inside the margin
    f.k_emit = next_pow2(a.k + a.k / 2 + 32);
    if (f.k_emit < 64) f.k_emit = 64;
    if (f.k_emit > (uint32_t)(P - ROWS)) f.k_emit = P - ROWS;
    // TF32: 2^-10 (rounding of both operands, unit roundoff 2^-11 each) + d * 2^-22 (fp32 accumulation, truncating
    // adders) + fp32 rounding of the norms.  The 16-bit kernel measures its rounding residues instead.
    f.acc_c = (float)a.dim * 2.384185791015625e-7f + 4e-6f;
    f.margin_tf32 = 9.765625e-4f + f.acc_c;
    f.big = shadow_pass_supported(di, a);
    f.stage2 = f.big && a.dtype == NK_DTYPE_F32 && tc_common_ok(di, a);
    f.can_exact_tc = scan_tensor_supported(di, a);  // euclidean / 16-bit rows have no 3xTF32 twin: overflow -> CUDA-core scan
    // 16-bit margins are ~6x wider than TF32's and clustered corpora put hundreds of rows per CTA inside them: the 16-bit
    // kernel has 1024-slot buffers and emits up to 512 entries per (CTA, query) (the finish step re-scores in rounds)
    if (f.big) f.k_emit = a.k <= 160 ? 512u : (uint32_t)(P_SHADOW - ROWS);
    f.dimpad = f.big ? a.shadow_dimpad : (a.dim + 63) / 64 * 64;
    const bool need_f32q = a.dtype == NK_DTYPE_F32;  // hi / lo arrays only serve passes over fp32 rows
    const size_t qaux_floats = (need_f32q ? (size_t)2 * f.Qpad * a.dim : 0) + (size_t)3 * f.QA;
    if (ws_reserve((void **)&ws.qaux, &ws.qaux_bytes, qaux_floats * 4 + (f.big ? (size_t)f.Qpad * f.dimpad * 2 : 0))) return -1;
    f.pair = f.big && a.Q >= 256 && pair_pass_supported(di, a, f.grid);
    if (ws_reserve((void **)&ws.cand, &ws.cand_bytes, (size_t)f.grid * (f.pair ? 256 : QT_MAX) * (f.big ? P_SHADOW : P) * 8)) return -1;
    if (ws_reserve((void **)&ws.partial, &ws.partial_bytes, (size_t)a.Q * f.grid * f.k_emit * 8)) return -1;
    if (ws_reserve((void **)&ws.keys2, &ws.keys2_bytes, (size_t)f.QA * 8)) return -1;  // gtau[] + gcount[]
    f.qhi = need_f32q ? ws.qaux : nullptr;
    f.qlo = need_f32q ? ws.qaux + (size_t)f.Qpad * a.dim : nullptr;
    f.qnorm = ws.qaux + (need_f32q ? (size_t)2 * f.Qpad * a.dim : 0);
    f.qa = f.qnorm + f.QA;
    f.qb = f.qa + f.QA;
    f.qbf16 = ws.qaux + qaux_floats;
    f.max_groups = (uint32_t)tc_env_int("NK_TC_QGROUPS", 4);
    // G query groups leave each CTA 1/G of the grid for its queries, i.e. G times the rows — and G times the rows inside
    // the BF16 margin — per (CTA, query) buffer: large k keeps fewer groups so that k + margin rows stay inside k_emit
    f.max_groups_shadow = a.k <= 128 ? 4u : a.k <= 160 ? 2u : 1u;
    if (f.max_groups_shadow > f.max_groups) f.max_groups_shadow = f.max_groups;
    // Sampled initial threshold: single-pass batches on shards with enough tiles per CTA for the flood tiles to matter.
    // (Large batches re-score the sample once per query — more L2 traffic than the flood tiles cost.)
    f.sample = 0;
    if (tc_env_int("NK_TAU_SAMPLE", 1) && a.Q <= 128 && a.n >= 4096) {
        f.sample = next_pow2(8 * a.k);  // k-th best of the sample = top k/S of the shard: ~8-12% of the rows pass at first
        if (f.sample < 128) f.sample = 128;
        if (f.sample > 2048) f.sample = 2048;
        const int s_env = tc_env_int("NK_TAU_SAMPLE_S", 0);  // experiments: sample size override (power of two)
        if (s_env >= 32 && s_env <= 2048) f.sample = next_pow2((uint32_t)s_env);
    }
    f.fsmem = (size_t)FINISH_HI * 4 + (size_t)FINISH_CAP * 16 + (size_t)a.dim * 4;
    f.psmem = (size_t)((a.dim + 3) & ~3u) * 4 + (size_t)f.sample * 8;
    if (tc_ensure_smem(reinterpret_cast<const void *>(filter_finish_kernel), di.device_id, f.fsmem)) return -1;
    if (f.psmem > 48 * 1024 && tc_ensure_smem(reinterpret_cast<const void *>(filter_prep_kernel), di.device_id, f.psmem)) return -1;
    return 0;
}

// Exact 3xTF32 scan.  only_if != nullptr: every kernel early-exits unless *only_if != 0 (device-side fallback).
// prep: convert the queries here (stand-alone use); the filter path has already done it in its fused prep.
static int scan_tensor_exact(const DeviceInfo &di, const ScanArgs &a, Workspace &ws, uint64_t *out_keys, uint64_t *launches,
                             const int *only_if, bool prep, bool count_main) {
    using namespace tc;
    const uint32_t Qpad = (a.Q + 63) / 64 * 64;
    const uint32_t num_tiles = (a.n + ROWS - 1) / ROWS;
    uint32_t grid = (uint32_t)di.num_sms;
    if (grid > num_tiles) grid = num_tiles;
    if (ws_reserve((void **)&ws.qaux, &ws.qaux_bytes, ((size_t)2 * Qpad * a.dim + Qpad + QT_MAX) * 4)) return -1;
    if (ws_reserve((void **)&ws.cand, &ws.cand_bytes, (size_t)grid * QT_MAX * P * 8)) return -1;
    if (ws_reserve((void **)&ws.partial, &ws.partial_bytes, (size_t)a.Q * grid * a.k * 8)) return -1;
    float *qhi = ws.qaux, *qlo = ws.qaux + (size_t)Qpad * a.dim;
    if (prep) {
        PrepParams pp{};
        pp.q = a.queries; pp.Q = a.Q; pp.dim = a.dim; pp.dimpad = a.dim; pp.normalise = a.metric == NK_METRIC_COSINE; pp.metric = a.metric;
        pp.qhi = qhi; pp.qlo = qlo; pp.Qpad = Qpad; pp.only_if = only_if;
        const size_t psmem = (size_t)((a.dim + 3) & ~3u) * 4;
        if (psmem > 48 * 1024 && tc_ensure_smem(reinterpret_cast<const void *>(filter_prep_kernel), di.device_id, psmem)) return -1;
        filter_prep_kernel<<<Qpad, PREP_THREADS, psmem, a.stream>>>(pp);
        NK_CUDA_OK(cudaGetLastError());
        if (launches) ++*launches;
    }
    if (count_main && a.ev_begin) NK_CUDA_OK(cudaEventRecord(a.ev_begin, a.stream));
    for (uint32_t q0 = 0; q0 < a.Q; q0 += 64) {
        TcPassArgs t{};
        t.grid = grid; t.k_emit = a.k; t.Qpad = Qpad; t.q0 = q0; t.nq = a.Q - q0 < 64u ? a.Q - q0 : 64u; t.qgroups = 1;
        t.qhi = qhi; t.qlo = qlo; t.only_if = only_if; t.count_main = count_main;
        if (launch_pass<3, 64>(di, a, ws, t, launches)) return -1;
    }
    if (count_main && a.ev_end) NK_CUDA_OK(cudaEventRecord(a.ev_end, a.stream));
    // fold the per-CTA lists and, when the caller wants them, write the decoded (index, score) arrays in the same launch
    if (merge_keys(ws.partial, grid, a.k, (size_t)grid * a.k, a.Q, a.k, out_keys, a.stream, only_if, 0, a.out_idx, a.out_score, a.metric)) return -1;
    if (launches) ++*launches;
    if (count_main) tc_print_prof(a.stream, num_tiles, grid, (a.dim + BK - 1) / BK);
    return 0;
}

int scan_tensor(const DeviceInfo &di, const ScanArgs &a, Workspace &ws, uint64_t *out_keys, uint64_t *launches) {
    if (a.n == 0 || a.Q == 0 || a.k == 0) return 0;
    if (!scan_tensor_supported(di, a)) {
        set_error("tensor path: unsupported shape");
        return -1;
    }
    return scan_tensor_exact(di, a, ws, out_keys, launches, nullptr, true, true);
}

// TF32 passes over queries [qfirst, Q): 128 query columns per MMA while more than 64 queries remain (twice the queries per
// corpus byte streamed), a 64-column launch for the tail; 2 or 4 query blocks per launch share every corpus tile through
// L2 (sibling CTAs).
static int tf32_passes(const DeviceInfo &di, const ScanArgs &a, Workspace &ws, const FilterPlan &f, uint32_t qfirst, const int *only_if,
                       bool count_main, int presampled, uint64_t *launches) {
    for (uint32_t q0 = qfirst; q0 < a.Q;) {
        const uint32_t left = a.Q - q0;
        TcPassArgs t{};
        t.grid = f.grid; t.k_emit = f.k_emit; t.Qpad = f.Qpad; t.q0 = q0; t.margin_c = f.margin_tf32;
        t.qhi = f.qhi; t.qnorm = f.qnorm; t.only_if = only_if; t.count_main = count_main; t.presampled = presampled;
        if (left > 64) {
            uint32_t groups = 1;
            if (left > 3 * 128 && f.max_groups >= 4 && f.grid % 4 == 0 && f.grid >= 8) groups = 4;
            else if (left > 128 && f.max_groups >= 2 && f.grid % 2 == 0 && f.grid >= 4) groups = 2;
            t.nq = left < 128u * groups ? left : 128u * groups;
            t.qgroups = groups;
            if (launch_pass<1, 128>(di, a, ws, t, launches)) return -1;
        } else {
            t.nq = left; t.qgroups = 1;
            if (launch_pass<1, 64>(di, a, ws, t, launches)) return -1;
        }
        q0 += t.nq;
    }
    return 0;
}

static void fill_finish(FinishParams &fp, const ScanArgs &a, Workspace &ws, const FilterPlan &f, uint64_t *out_keys) {
    fp.rows = a.rows; fp.dtype = a.dtype; fp.dim = a.dim; fp.row_base = a.row_base; fp.queries = a.queries;
    fp.lists = ws.partial; fp.gcount = reinterpret_cast<const int *>(ws.keys2) + f.QA; fp.list_cap = f.grid * f.k_emit; fp.k = a.k;
    fp.metric = a.metric; fp.margin_c = f.margin_tf32; fp.flags = ws.flags; fp.out = out_keys;
    fp.out_idx = a.out_idx; fp.out_score = a.out_score; fp.min_score = a.min_score;
    fp.qa = f.qa; fp.qb = f.qb; fp.qn = f.qnorm;
    fp.state = reinterpret_cast<uint32_t *>(ws.keys2); fp.state_words = 2 * f.QA;
}

// The retry / exact stages of a filter search.  Every kernel returns at once unless the stage before raised its flag on
// the device, so the tail can be queued blindly behind the first stage (asynchronous API: no host round trip) — or, for
// the host-synchronous API, only after the host has seen a flag (a_defer_tail): the common case then launches nothing.
int scan_tensor_filter_tail(const DeviceInfo &di, const ScanArgs &a, Workspace &ws, uint64_t *out_keys, uint64_t *launches) {
    using namespace tc;
    FilterPlan f;
    if (make_filter_plan(di, a, ws, &f)) return -1;
    if (f.stage2 && a.defer_tail) {
        // the same search over the fp32 rows with the (much tighter) TF32 margins, if a 16-bit margin buffer overflowed.
        // Host-driven tail only: queued blindly (asynchronous API) the stage would cost two early-exit launches on EVERY
        // search to make the rare overflow cheaper — there an overflow goes straight to the exact stage.
        if (tf32_passes(di, a, ws, f, 0, ws.flags + FLAG_RETRY, false, 0, launches)) return -1;
        FinishParams fp{};
        fill_finish(fp, a, ws, f, out_keys);
        fp.q_big = 0; fp.only_if = ws.flags + FLAG_RETRY; fp.mark_retry = 0;
        NK_CUDA_OK(launch_pdl(filter_finish_kernel, dim3(a.Q), dim3(FINISH_THREADS), f.fsmem, a.stream, !a.defer_tail, fp));
        if (launches) ++*launches;
    }
    ScanArgs b = a;
    b.ev_begin = b.ev_end = nullptr;
    b.main_launches = nullptr;
    if (f.can_exact_tc) {
        if (scan_tensor_exact(di, b, ws, out_keys, launches, ws.flags + FLAG_OVERFLOW, false, false)) return -1;
    } else {  // euclidean / large k / 16-bit rows: the CUDA-core scan is the exact twin
        b.only_if = ws.flags + FLAG_OVERFLOW;
        if (scan_simt(di, b, ws, out_keys, launches)) return -1;
    }
    return 0;  // (both exact twins write the decoded arrays from their merge launch)
}

// Filter mode: prep -> 16-bit (or 1xTF32) scan with rigorous margins -> finish (select by upper bound, exact fp32
// rescoring, decode).  Three launches in the common case; the retry / exact stages follow as scan_tensor_filter_tail.
int scan_tensor_filter(const DeviceInfo &di, const ScanArgs &a, Workspace &ws, uint64_t *out_keys, uint64_t *launches) {
    using namespace tc;
    if (a.n == 0 || a.Q == 0 || a.k == 0) return 0;
    if (!scan_tensor_filter_supported(di, a)) {
        set_error("tensor filter path: unsupported shape");
        return -1;
    }
    FilterPlan f;
    if (make_filter_plan(di, a, ws, &f)) return -1;
    const uint32_t q_big = f.big ? a.Q : 0;  // queries served by the 16-bit kernel in the first stage (all or none)

    PrepParams pp{};
    pp.q = a.queries; pp.Q = a.Q; pp.dim = a.dim; pp.dimpad = f.dimpad; pp.normalise = a.metric == NK_METRIC_COSINE; pp.metric = a.metric;
    pp.acc_c = f.acc_c;
    pp.qhi = f.qhi; pp.qlo = f.can_exact_tc ? f.qlo : nullptr;
    pp.qbf = f.big ? static_cast<uint16_t *>(f.qbf16) : nullptr; pp.qbf_f16 = a.dtype == NK_DTYPE_F16;
    pp.qnorm = f.qnorm; pp.qa = f.qa; pp.qb = f.qb;
    pp.state = reinterpret_cast<uint32_t *>(ws.keys2); pp.Qpad = f.Qpad; pp.QA = f.QA; pp.flags = ws.flags;
    pp.rows = a.rows; pp.dtype = a.dtype; pp.n = a.n; pp.sample = f.sample; pp.k = a.k; pp.mask = a.row_mask; pp.min_score = a.min_score;
    filter_prep_kernel<<<f.Qpad, PREP_THREADS, f.psmem, a.stream>>>(pp);
    NK_CUDA_OK(cudaGetLastError());
    if (launches) ++*launches;

    if (a.ev_begin) NK_CUDA_OK(cudaEventRecord(a.ev_begin, a.stream));
    for (uint32_t q0 = 0; q0 < q_big;) {  // 16-bit passes: 128 query columns (up to 4 query groups per launch), 64 for the tail
        const uint32_t left = q_big - q0;
        ShadowPassArgs sp{};
        sp.grid = f.grid; sp.k_emit = f.k_emit; sp.dimpad = f.dimpad; sp.Qpad = f.Qpad; sp.q0 = q0;
        sp.qbf16 = f.qbf16; sp.qnorm = f.qnorm; sp.qa = f.qa; sp.qb = f.qb; sp.presampled = f.sample != 0;
        if (left >= 256 && f.pair) {
            // large batches: CTA pairs, 256 query columns per MMA, two query blocks per launch sharing tiles through L2
            const uint32_t gmax = (uint32_t)tc_env_int("NK_PAIR_GROUPS", 4);
            sp.qgroups = (left >= 1024 && f.grid >= 16 && gmax >= 4) ? 4u : (left >= 512 && f.grid >= 8 && gmax >= 2) ? 2u : 1u;
            sp.nq = left < 256u * sp.qgroups ? left / 256u * 256u : 256u * sp.qgroups;
            if (launch_pair_pass(di, a, ws, sp, launches)) return -1;
        } else if (left > 64) {
            uint32_t groups = 1;
            if (left > 3 * 128 && f.max_groups_shadow >= 4 && f.grid % 4 == 0 && f.grid >= 8) groups = 4;
            else if (left > 128 && f.max_groups_shadow >= 2 && f.grid % 2 == 0 && f.grid >= 4) groups = 2;
            sp.nq = left < 128u * groups ? left : 128u * groups;
            sp.qgroups = groups;
            if (launch_shadow_pass(128, di, a, ws, sp, launches)) return -1;
        } else {
            sp.nq = left; sp.qgroups = 1;
            if (launch_shadow_pass(64, di, a, ws, sp, launches)) return -1;
        }
        q0 += sp.nq;
    }
    if (!f.big && tf32_passes(di, a, ws, f, 0, nullptr, true, f.sample != 0, launches)) return -1;
    if (a.ev_end) NK_CUDA_OK(cudaEventRecord(a.ev_end, a.stream));

    FinishParams fp{};
    fill_finish(fp, a, ws, f, out_keys);
    fp.q_big = q_big; fp.only_if = nullptr; fp.mark_retry = (f.stage2 && a.defer_tail) ? 1 : 0;
    fp.debug = tc_debug_flags() & 64;
    NK_CUDA_OK(launch_pdl(filter_finish_kernel, dim3(a.Q), dim3(FINISH_THREADS), f.fsmem, a.stream, !a.defer_tail, fp));
    if (launches) ++*launches;
    tc_print_prof(a.stream, f.num_tiles, f.grid, (a.dim + BK - 1) / BK);
    fin_print_prof(a.stream, a.Q);
    if (a.defer_tail) return 0;
    return scan_tensor_filter_tail(di, a, ws, out_keys, launches);
}

// ---- tests only: the filters' raw score estimates and error bounds for every (row, query) pair --------------------
// which: NK_PATH_TENSOR_FILTER (1xTF32 over the fp32 rows) or NK_PATH_TENSOR_SHADOW (16-bit pass).  est / bnd: device
// [n x ld] floats.  The estimate is what the kernel compares (before adding the bound); |est - exact| <= bnd is the
// invariant the filter's soundness rests on (tests/test_gpu_error_model.py measures it against fp64).
int scan_filter_dump(const DeviceInfo &di, const ScanArgs &a, Workspace &ws, int which, float *est, float *bnd, uint32_t ld,
                     uint64_t *launches) {
    using namespace tc;
    if (a.Q == 0 || a.Q > 64 || a.n == 0) { set_error("dump: 1 <= Q <= 64 and n > 0 required"); return -1; }
    ScanArgs b = a;
    b.k = 1; b.min_score = -INFINITY; b.defer_tail = true; b.out_idx = nullptr; b.out_score = nullptr;
    if (!scan_tensor_filter_supported(di, b)) { set_error("dump: unsupported shape"); return -1; }
    FilterPlan f;
    if (make_filter_plan(di, b, ws, &f)) return -1;
    if (which == NK_PATH_TENSOR_SHADOW && !f.big) { set_error("dump: no 16-bit pass for this shard"); return -1; }
    if (which == NK_PATH_TENSOR_FILTER && a.dtype != NK_DTYPE_F32) { set_error("dump: TF32 pass needs fp32 rows"); return -1; }
    PrepParams pp{};
    pp.q = b.queries; pp.Q = b.Q; pp.dim = b.dim; pp.dimpad = f.dimpad; pp.normalise = b.metric == NK_METRIC_COSINE; pp.metric = b.metric;
    pp.acc_c = f.acc_c; pp.qhi = f.qhi; pp.qlo = nullptr;
    pp.qbf = f.big ? static_cast<uint16_t *>(f.qbf16) : nullptr; pp.qbf_f16 = b.dtype == NK_DTYPE_F16;
    pp.qnorm = f.qnorm; pp.qa = f.qa; pp.qb = f.qb;
    pp.state = reinterpret_cast<uint32_t *>(ws.keys2); pp.Qpad = f.Qpad; pp.QA = f.QA; pp.flags = ws.flags;
    pp.rows = b.rows; pp.dtype = b.dtype; pp.n = b.n; pp.sample = 0; pp.k = 1; pp.min_score = -INFINITY;
    filter_prep_kernel<<<f.Qpad, PREP_THREADS, f.psmem, b.stream>>>(pp);
    NK_CUDA_OK(cudaGetLastError());
    if (launches) ++*launches;
    if (which == NK_PATH_TENSOR_SHADOW) {
        ShadowPassArgs sp{};
        sp.grid = f.grid; sp.k_emit = f.k_emit; sp.dimpad = f.dimpad; sp.Qpad = f.Qpad; sp.q0 = 0; sp.nq = b.Q; sp.qgroups = 1;
        sp.qbf16 = f.qbf16; sp.qnorm = f.qnorm; sp.qa = f.qa; sp.qb = f.qb; sp.presampled = 0;
        sp.dump_est = est; sp.dump_bnd = bnd; sp.dump_ld = ld;
        if (launch_shadow_pass(64, di, b, ws, sp, launches)) return -1;
    } else {
        TcPassArgs t{};
        t.grid = f.grid; t.k_emit = f.k_emit; t.Qpad = f.Qpad; t.q0 = 0; t.nq = b.Q; t.qgroups = 1; t.margin_c = f.margin_tf32;
        t.qhi = f.qhi; t.qnorm = f.qnorm; t.dump_est = est; t.dump_bnd = bnd; t.dump_ld = ld;
        if (launch_pass<1, 64, true>(di, b, ws, t, launches)) return -1;
    }
    return 0;
}

}  // namespace nk
