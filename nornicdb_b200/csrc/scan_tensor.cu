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

// Queries -> tf32 hi (/ lo) arrays, zero padded to a multiple of 64 rows; cosine normalises first; also |q|.
__global__ void tc_prep_queries_kernel(const float *q, uint32_t Q, uint32_t dim, int normalise, float *qhi, float *qlo,
                                       float *qnorm) {
    const uint32_t row = blockIdx.x;
    __shared__ float red[32];
    float t = 0.0f;
    if (row < Q) {
        float a = 0.0f;
        for (uint32_t j = threadIdx.x; j < dim; j += blockDim.x) a = fmaf(q[(size_t)row * dim + j], q[(size_t)row * dim + j], a);
#pragma unroll
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
        __syncthreads();
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    }
    const float nrm = sqrtf(t);
    // zero query -> all cosine scores 0 (simd_amd64.go:31-35)
    const float inv = normalise ? (t > 0.0f ? 1.0f / nrm : 0.0f) : 1.0f;
    if (threadIdx.x == 0 && qnorm) qnorm[row] = normalise ? (t > 0.0f ? 1.0f : 0.0f) : nrm;
    for (uint32_t j = threadIdx.x; j < dim; j += blockDim.x) {
        float v = row < Q ? q[(size_t)row * dim + j] * inv : 0.0f;
        float h = __uint_as_float(ptx::tf32_round_bits(__float_as_uint(v)));
        qhi[(size_t)row * dim + j] = h;
        if (qlo) qlo[(size_t)row * dim + j] = v - h;
    }
}

template <int NT, int QT>
__global__ void __launch_bounds__(tc::THREADS, 1)
knn_scan_tc_kernel(const __grid_constant__ CUtensorMap map_rows, const __grid_constant__ CUtensorMap map_qhi,
                   const __grid_constant__ CUtensorMap map_qlo, tc::Params p) {
    using namespace tc;
    using C = Cfg<NT, QT>;
    constexpr bool FILTER = NT == 1;
    constexpr int A_COL = C::A_COL, B_BYTES = C::B_BYTES;
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
        sh.tau[tid] = -INFINITY;
        sh.cnt[tid] = 0;
        sh.qn[tid] = (FILTER && p.qnorm) ? p.qnorm[qpad_off + tid] : 1.0f;
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = sh.tmem_base;

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
                    if (!cosine && alive) atomicMax(&sh.maxxx, __float_as_uint(x2));
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
                    if (it < 2 && cb < nq) {
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
                                my_cand[(size_t)qi * P + slot] = alive ? make_key(sc, grow) : 0ull;  // 0 = empty slot
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
                    float floor_tau = -INFINITY;
                    if (FILTER) {
                        const uint32_t g = __ldcg(p.gtau + q0 + qi);
                        if (g) floor_tau = ord_to_float(g);
                    }
                    warp_prune<16>(my_cand + (size_t)qi * P, &sh.cnt[qi], &sh.tau[qi], p.k, lane, nullptr, 0, FILTER, margin2, prune_at, floor_tau);
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
                float floor_tau = -INFINITY;
                const uint32_t g = __ldcg(p.gtau + q0 + qi);
                if (g) floor_tau = ord_to_float(g);
                warp_prune<16>(my_cand + (size_t)qi * P, &sh.cnt[qi], &sh.tau[qi], p.k, lane,
                               p.partial + (size_t)(q0 + qi) * p.list_cap, (int)p.list_cap, true, margin2,
                               (int)p.k_emit, floor_tau, p.gcount + q0 + qi);
                // the list was cut at k_emit while rows inside the margin remained -> exact fallback
                if (lane == 0 && sh.cnt[qi] >= (int)p.k_emit && (int)p.k_emit > (int)p.k) atomicOr(p.flags + 1, 2);
                if (lane == 0 && sh.tau[qi] > -INFINITY) atomicMax(p.gtau + q0 + qi, ord_bits(sh.tau[qi]));
            } else {
                warp_prune<16>(my_cand + (size_t)qi * P, &sh.cnt[qi], &sh.tau[qi], p.k, lane,
                               p.partial + ((size_t)(q0 + qi) * gridDim.x + blockIdx.x) * p.k_emit, (int)p.k_emit, false, 0.0f, (int)p.k_emit);
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
// cross-CTA threshold (k plus a few dozen).  Sort it by bound; everything with bound >= (k-th bound - 2*Bmax) may belong
// to the true top-k: those rows are re-scored EXACTLY in fp32 with the same arithmetic as the CUDA-core scan
// (dot / sqrt(|x|^2 |q|^2) etc.), sorted by (score desc, row asc), and the best k written out.  A list longer than
// FINISH_CAP (tiny shards with large k, or adversarial near-ties) raises the overflow flag; the exact kernels queued
// behind redo the search.
// ---------------------------------------------------------------------------------------------------
constexpr int FINISH_THREADS = 256;
constexpr int FINISH_CAP = 4096;
struct FinishParams {
    const void *rows;        // fp32 corpus shard
    uint32_t dim;
    uint64_t row_base;
    const float *queries;    // raw fp32 queries [Q x dim]
    const uint64_t *lists;   // [Q][list_cap] shared append lists (keys carry upper bounds)
    const int *gcount;       // [Q]
    uint32_t list_cap, k;
    int metric;
    float margin_c;          // TF32 passes: c in |s_hat - s| <= c |x||q|
    uint32_t q_big;          // queries [0, q_big) went through the BF16 kernel: bound factors qa / qb / qn below
    const float *qa, *qb, *qn;
    int *flags;
    const int *only_if;      // retry stage: run only if *only_if != 0
    uint64_t *out;           // [Q][k]
};

__global__ void __launch_bounds__(FINISH_THREADS) filter_finish_kernel(FinishParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *sb = reinterpret_cast<uint64_t *>(smem_raw);                      // FINISH_CAP bound keys
    uint64_t *se = reinterpret_cast<uint64_t *>(smem_raw + FINISH_CAP * 8);     // FINISH_CAP exact keys
    float *qs = reinterpret_cast<float *>(smem_raw + FINISH_CAP * 16);          // query
    __shared__ float s_qq;
    __shared__ int s_count;
    if (p.only_if && *p.only_if == 0) return;
    const uint32_t q = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t j = tid; j < p.dim; j += FINISH_THREADS) qs[j] = p.queries[(size_t)q * p.dim + j];
    int n = p.gcount[q];
    if (tid == 0) atomicMax(p.flags + 7, n);  // diagnostics: longest shared list of this search
    if (n > (int)p.list_cap) n = (int)p.list_cap;
    __syncthreads();
    if (warp == 0) {
        float a = 0.0f;
        for (uint32_t j = lane; j < p.dim; j += 32) a = fmaf(qs[j], qs[j], a);
#pragma unroll
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) s_qq = a;
    }
    __syncthreads();
    const float maxxx = __uint_as_float((unsigned int)p.flags[2]);
    const float margin2 = q < p.q_big ? bf16_margin2(p.metric, __uint_as_float((unsigned int)p.flags[4]), __uint_as_float((unsigned int)p.flags[6]),
                                                     maxxx, p.qa[q], p.qb[q], p.qn[q])
                                      : filter_margin2(p.metric, p.margin_c, maxxx, sqrtf(s_qq));
    const uint64_t *list = p.lists + (size_t)q * p.list_cap;
    // k-th largest bound by radix select over the score bits that actually vary (8 bits per pass, smem histogram; the
    // list is read from L2), then everything inside the margin below it is gathered for exact re-scoring.
    __shared__ int hist[256];
    __shared__ uint32_t s_prefix, s_red[2][FINISH_THREADS / 32];
    __shared__ int s_krem;
    uint32_t umax = 0u, umin = 0xffffffffu;
    for (int i = tid; i < n; i += FINISH_THREADS) {
        const uint32_t hi = (uint32_t)(__ldcg(list + i) >> 32);
        umax = max(umax, hi);
        umin = min(umin, hi);
    }
    umax = __reduce_max_sync(0xffffffffu, umax);
    umin = __reduce_min_sync(0xffffffffu, umin);
    if (lane == 0) { s_red[0][warp] = umax; s_red[1][warp] = umin; }
    if (tid == 0) { s_count = 0; s_krem = (int)p.k; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < FINISH_THREADS / 32; ++w) { umax = max(umax, s_red[0][w]); umin = min(umin, s_red[1][w]); }
    uint64_t thr_key = 1ull;  // fewer than k entries: keep them all
    if (n >= (int)p.k) {
        int rem = 32 - __clz(umax ^ umin);  // low bits in which the bounds differ (0: all equal)
        if (tid == 0) s_prefix = rem >= 32 ? 0u : (umax >> rem) << rem;
        while (rem > 0) {
            const int w = rem < 8 ? rem : 8, shift = rem - w;
            hist[tid] = 0;  // FINISH_THREADS == 256
            __syncthreads();
            const uint32_t prefix = s_prefix;
            for (int i = tid; i < n; i += FINISH_THREADS) {
                const uint32_t hi = (uint32_t)(__ldcg(list + i) >> 32);
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
        thr_key = (uint64_t)ord_bits(ord_to_float(s_prefix) - margin2) << 32;
        if (thr_key == 0ull) thr_key = 1ull;
    }
    for (int i = tid; i < n; i += FINISH_THREADS) {
        const uint64_t key = __ldcg(list + i);
        if (key >= thr_key) {
            const int pos = atomicAdd(&s_count, 1);
            if (pos < FINISH_CAP) sb[pos] = key;
        }
    }
    __syncthreads();
    if (s_count > FINISH_CAP) {  // thousands of rows inside the margin: adversarial near-ties -> exact fallback
        if (tid == 0) atomicOr(p.flags + 1, 4);
    }
    const int count = s_count < FINISH_CAP ? s_count : FINISH_CAP;
    // exact scores
    for (int i = warp; i < count; i += FINISH_THREADS / 32) {
        const uint32_t grow = key_row(sb[i]);
        const size_t local = (size_t)(grow - (uint32_t)p.row_base);
        float d = 0.0f, xx = 0.0f;
        // fp32 rows, dim % 4 == 0, 16-byte aligned (tc_common_ok): 128-bit loads, all of a row's requests in flight
        const float4 *x4 = reinterpret_cast<const float4 *>(static_cast<const float *>(p.rows) + local * p.dim);
        const float4 *q4 = reinterpret_cast<const float4 *>(qs);
#pragma unroll 8
        for (uint32_t j = lane; j < p.dim / 4; j += 32) {
            const float4 v = __ldg(x4 + j), u = q4[j];
            if (p.metric == NK_METRIC_EUCLIDEAN) {
                float t;
                t = v.x - u.x; d = fmaf(t, t, d);
                t = v.y - u.y; d = fmaf(t, t, d);
                t = v.z - u.z; d = fmaf(t, t, d);
                t = v.w - u.w; d = fmaf(t, t, d);
            } else {
                d = fmaf(v.x, u.x, d); xx = fmaf(v.x, v.x, xx);
                d = fmaf(v.y, u.y, d); xx = fmaf(v.y, v.y, xx);
                d = fmaf(v.z, u.z, d); xx = fmaf(v.z, v.z, xx);
                d = fmaf(v.w, u.w, d); xx = fmaf(v.w, v.w, xx);
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            d += __shfl_xor_sync(0xffffffffu, d, o);
            xx += __shfl_xor_sync(0xffffffffu, xx, o);
        }
        if (lane == 0) {
            float sc = d;
            if (p.metric == NK_METRIC_EUCLIDEAN) sc = -d;
            else if (p.metric == NK_METRIC_COSINE) {
                float den = sqrtf(xx * s_qq);
                sc = den > 0.0f ? d / den : 0.0f;
            }
            if (sc != sc) sc = -INFINITY;
            se[i] = make_key(sc, grow);
        }
    }
    int P2 = 32;
    while (P2 < count) P2 <<= 1;
    for (int i = count + tid; i < P2; i += FINISH_THREADS) se[i] = 0ull;
    block_bitonic_sort_desc(se, P2);
    for (uint32_t i = tid; i < p.k; i += FINISH_THREADS) p.out[(size_t)q * p.k + i] = (int)i < count ? se[i] : 0ull;
}

// Staged fallback (all on the device, no host round trip):  BF16 filter -> TF32 filter -> exact.  After a stage,
// retry_mark moves its overflow flag [1] to the retry marker [5] and clears [1]; retry_zero wipes the shared thresholds
// and list fills if a retry is due.  The next stage's kernels run only_if flags[5] != 0.
__global__ void retry_mark_kernel(int *flags) {
    flags[5] = flags[1] != 0;
    flags[1] = 0;
}
__global__ void retry_zero_kernel(const int *only_if, unsigned long long *words, uint32_t n) {
    if (*only_if == 0) return;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) words[i] = 0ull;
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
                uint32_t box_rows, uint64_t row_stride_bytes) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return -1;
    }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {(cuuint64_t)row_stride_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%u", (int)r, (unsigned long long)rows, cols);
        return -1;
    }
    return 0;
}
static int make_map(CUtensorMap *m, const void *base, uint64_t rows, uint32_t dim, uint32_t box_rows) {
    return tc_make_map(m, base, rows, dim, 4, (uint32_t)tc::BK, box_rows, (uint64_t)dim * 4);
}

static bool tc_common_ok(const DeviceInfo &di, const ScanArgs &a) {
    if (di.cc < 100) return false;
    if (a.dtype != NK_DTYPE_F32) return false;                      // fp16 corpus: CUDA-core scan (HBM-bound at small Q)
    if (a.dim % 4 != 0 || a.dim < 32) return false;                 // TMA: 16-byte global stride
    if ((reinterpret_cast<uintptr_t>(a.rows) & 15) != 0) return false;
    if (a.n == 0 || a.k == 0) return false;
    return true;
}
// exact (3xTF32) mode: cosine / dot, k <= 255 (register-resident prune over a 512-slot buffer)
bool scan_tensor_supported(const DeviceInfo &di, const ScanArgs &a) {
    return tc_common_ok(di, a) && a.metric != NK_METRIC_EUCLIDEAN && a.k + tc::ROWS + 1 <= (uint32_t)tc::P;
}
// filter (1xTF32 + exact rescoring) mode: all three metrics, k <= 192
bool scan_tensor_filter_supported(const DeviceInfo &di, const ScanArgs &a) {
    return tc_common_ok(di, a) && a.k <= 192 && a.dim <= 32768;  // finish kernel keeps the query in shared memory
}

int tc_debug_flags() {
    const char *dbg = getenv("NK_TC_DEBUG");
    return dbg ? atoi(dbg) : 0;
}

// One launch: queries [q0, q0+nq) against the whole shard, QT = 64 or 128 query columns per MMA.
template <int NT, int QT>
static int launch_pass(const DeviceInfo &di, const ScanArgs &a, Workspace &ws, uint32_t grid, uint32_t k_emit, float margin_c,
                       const float *qhi, const float *qlo, const float *qnorm, uint32_t Qpad, uint32_t q0, uint32_t nq,
                       const int *only_if, uint64_t *launches, bool count_main, uint32_t qgroups = 1) {
    using namespace tc;
    CUtensorMap map_rows, map_qhi, map_qlo;
    if (make_map(&map_rows, a.rows, a.n, a.dim, ROWS)) return -1;
    if (make_map(&map_qhi, qhi, Qpad, a.dim, QT)) return -1;          // rows past Qpad are zero-filled by TMA
    if (make_map(&map_qlo, qlo ? qlo : qhi, Qpad, a.dim, QT)) return -1;
    const size_t smem = (size_t)Cfg<NT, QT>::RING_BYTES + sizeof(Shared) + 1024;
    if (smem > di.max_smem_optin) {
        set_error("tensor path needs %zu B shared memory (> %zu)", smem, di.max_smem_optin);
        return -1;
    }
    NK_CUDA_OK(cudaFuncSetAttribute(knn_scan_tc_kernel<NT, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    Params p;
    p.n = a.n; p.dim = a.dim; p.nslab = (a.dim + BK - 1) / BK; p.row_base = a.row_base;
    p.q0 = q0; p.nq = nq; p.k = a.k; p.qpad_off = q0; p.qgroups = qgroups; p.list_cap = grid * k_emit;
    p.metric = a.metric; p.k_emit = k_emit; p.margin_c = margin_c; p.qnorm = qnorm;
    p.cand = ws.cand; p.partial = ws.partial; p.flags = ws.flags; p.only_if = only_if; p.debug = tc_debug_flags(); p.mask = a.row_mask;
    p.gtau = reinterpret_cast<uint32_t *>(ws.keys2); p.gcount = reinterpret_cast<int *>(ws.keys2) + (Qpad + QT_BIG);
    knn_scan_tc_kernel<NT, QT><<<grid, THREADS, smem, a.stream>>>(map_rows, map_qhi, map_qlo, p);
    NK_CUDA_OK(cudaGetLastError());
    if (launches) ++*launches;
    if (count_main && a.main_launches) ++*a.main_launches;
    return 0;
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

// Exact 3xTF32 scan.  only_if != nullptr: every kernel early-exits unless *only_if != 0 (device-side fallback).
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
    float *qhi = ws.qaux, *qlo = ws.qaux + (size_t)Qpad * a.dim, *qnorm = ws.qaux + (size_t)2 * Qpad * a.dim;
    if (prep) {
        tc_prep_queries_kernel<<<Qpad, 256, 0, a.stream>>>(a.queries, a.Q, a.dim, a.metric == NK_METRIC_COSINE, qhi, qlo, qnorm);
        NK_CUDA_OK(cudaGetLastError());
        if (launches) ++*launches;
    }
    if (count_main && a.ev_begin) NK_CUDA_OK(cudaEventRecord(a.ev_begin, a.stream));
    for (uint32_t q0 = 0; q0 < a.Q; q0 += 64) {
        const uint32_t nq = a.Q - q0 < 64u ? a.Q - q0 : 64u;
        if (launch_pass<3, 64>(di, a, ws, grid, a.k, 0.0f, qhi, qlo, nullptr, Qpad, q0, nq, only_if, launches, count_main)) return -1;
    }
    if (count_main && a.ev_end) NK_CUDA_OK(cudaEventRecord(a.ev_end, a.stream));
    if (merge_keys(ws.partial, grid, a.k, (size_t)grid * a.k, a.Q, a.k, out_keys, a.stream, only_if)) return -1;
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

// Filter mode: 1xTF32 scan with rigorous margins -> merge by upper bound -> exact fp32 rescoring; the exact 3xTF32
// search is enqueued behind it and runs only if the device-side overflow flag was raised.
int scan_tensor_filter(const DeviceInfo &di, const ScanArgs &a, Workspace &ws, uint64_t *out_keys, uint64_t *launches) {
    using namespace tc;
    if (a.n == 0 || a.Q == 0 || a.k == 0) return 0;
    if (!scan_tensor_filter_supported(di, a)) {
        set_error("tensor filter path: unsupported shape");
        return -1;
    }
    const uint32_t Qpad = (a.Q + 63) / 64 * 64;
    const uint32_t num_tiles = (a.n + ROWS - 1) / ROWS;
    uint32_t grid = (uint32_t)di.num_sms;
    if (grid > num_tiles) grid = num_tiles;
    // per-CTA contribution: at most k + room for the rows inside the margin
    uint32_t k_emit = next_pow2(a.k + a.k / 2 + 32);
    if (k_emit < 64) k_emit = 64;
    if (k_emit > (uint32_t)(P - ROWS)) k_emit = P - ROWS;
    // TF32: 2^-10 (rounding of both operands, unit roundoff 2^-11 each) + d * 2^-22 (fp32 accumulation, truncating
    // adders) + fp32 rounding of the norms.  The BF16 kernel (big batches) measures its rounding residues instead.
    const float acc_c = (float)a.dim * 2.384185791015625e-7f + 4e-6f;
    const float margin_tf32 = 9.765625e-4f + acc_c;
    // a BF16 shadow of the shard (scan_tensor_shadow.cu) halves the bytes the filter streams; the TF32 scan over the fp32
    // rows is then the retry stage
    const bool big = shadow_pass_supported(di, a);
    if (big && k_emit < 128) k_emit = 128;  // BF16 margins are ~3x wider: more rows per CTA sit inside them
    const uint32_t dimpad = big ? a.shadow_dimpad : (a.dim + 63) / 64 * 64;
    const uint32_t QA = Qpad + QT_BIG;  // padded per-query arrays

    const size_t qaux_floats = (size_t)2 * Qpad * a.dim + (size_t)3 * QA;
    if (ws_reserve((void **)&ws.qaux, &ws.qaux_bytes, qaux_floats * 4 + (big ? (size_t)Qpad * dimpad * 2 : 0))) return -1;
    if (ws_reserve((void **)&ws.cand, &ws.cand_bytes, (size_t)grid * QT_MAX * P * 8)) return -1;
    if (ws_reserve((void **)&ws.partial, &ws.partial_bytes, (size_t)a.Q * grid * k_emit * 8)) return -1;
    if (ws_reserve((void **)&ws.keys2, &ws.keys2_bytes, (size_t)QA * 8)) return -1;  // gtau[] + gcount[]
    float *qhi = ws.qaux, *qlo = ws.qaux + (size_t)Qpad * a.dim, *qnorm = ws.qaux + (size_t)2 * Qpad * a.dim;
    float *qa = qnorm + QA, *qb = qa + QA;
    void *qbf16 = ws.qaux + qaux_floats;

    NK_CUDA_OK(cudaMemsetAsync(ws.flags + 1, 0, 7 * sizeof(int), a.stream));  // overflow, running maxima, retry marker
    NK_CUDA_OK(cudaMemsetAsync(ws.keys2, 0, (size_t)QA * 8, a.stream));        // shared thresholds + list fills
    const bool can_fallback = scan_tensor_supported(di, a);  // euclidean has no 3xTF32 twin: overflow -> CUDA-core scan
    uint32_t max_groups = 4;
    if (const char *e = getenv("NK_TC_QGROUPS")) max_groups = (uint32_t)atoi(e);
    // G query groups leave each CTA 1/G of the grid for its queries, i.e. G times the rows — and G times the rows inside
    // the BF16 margin — per (CTA, query) buffer: large k keeps fewer groups so that k + margin rows stay inside k_emit
    uint32_t max_groups_shadow = a.k <= 32 ? 4u : a.k <= 64 ? 2u : 1u;
    if (max_groups_shadow > max_groups) max_groups_shadow = max_groups;
    const uint32_t q_big = big ? a.Q : 0;  // queries served by the shadow kernel in the first stage (all or none)
    tc_prep_queries_kernel<<<Qpad, 256, 0, a.stream>>>(a.queries, a.Q, a.dim, a.metric == NK_METRIC_COSINE, qhi,
                                                       can_fallback ? qlo : nullptr, qnorm);
    NK_CUDA_OK(cudaGetLastError());
    if (launches) ++*launches;
    if (q_big && bf16_prep_queries(a, Qpad, dimpad, acc_c, qbf16, qnorm, qa, qb, launches)) return -1;

    FinishParams fp;
    fp.rows = a.rows; fp.dim = a.dim; fp.row_base = a.row_base; fp.queries = a.queries;
    fp.lists = ws.partial; fp.gcount = reinterpret_cast<const int *>(ws.keys2) + QA; fp.list_cap = grid * k_emit; fp.k = a.k;
    fp.metric = a.metric; fp.margin_c = margin_tf32; fp.flags = ws.flags; fp.out = out_keys;
    fp.qa = qa; fp.qb = qb; fp.qn = qnorm;
    const size_t fsmem = (size_t)FINISH_CAP * 16 + (size_t)a.dim * 4;
    NK_CUDA_OK(cudaFuncSetAttribute(filter_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));

    // TF32 passes over queries [qfirst, Q): 128 query columns per MMA while more than 64 queries remain (twice the queries
    // per corpus byte streamed), a 64-column launch for the tail; 2 or 4 query blocks per launch share every corpus tile
    // through L2 (sibling CTAs).
    auto tf32_passes = [&](uint32_t qfirst, const int *only_if, bool count_main) -> int {
        for (uint32_t q0 = qfirst; q0 < a.Q;) {
            const uint32_t left = a.Q - q0;
            if (left > 64) {
                uint32_t groups = 1;
                if (left > 3 * 128 && max_groups >= 4 && grid % 4 == 0 && grid >= 8) groups = 4;
                else if (left > 128 && max_groups >= 2 && grid % 2 == 0 && grid >= 4) groups = 2;
                const uint32_t nq = left < 128u * groups ? left : 128u * groups;
                if (launch_pass<1, 128>(di, a, ws, grid, k_emit, margin_tf32, qhi, nullptr, qnorm, Qpad, q0, nq, only_if, launches, count_main, groups)) return -1;
                q0 += nq;
            } else {
                if (launch_pass<1, 64>(di, a, ws, grid, k_emit, margin_tf32, qhi, nullptr, qnorm, Qpad, q0, left, only_if, launches, count_main)) return -1;
                q0 += left;
            }
        }
        return 0;
    };

    if (a.ev_begin) NK_CUDA_OK(cudaEventRecord(a.ev_begin, a.stream));
    for (uint32_t q0 = 0; q0 < q_big;) {  // shadow passes: 128 query columns (up to 4 query groups per launch), 64 for the tail
        const uint32_t left = q_big - q0;
        if (left > 64) {
            uint32_t groups = 1;
            if (left > 3 * 128 && max_groups_shadow >= 4 && grid % 4 == 0 && grid >= 8) groups = 4;
            else if (left > 128 && max_groups_shadow >= 2 && grid % 2 == 0 && grid >= 4) groups = 2;
            const uint32_t nq = left < 128u * groups ? left : 128u * groups;
            if (launch_shadow_pass(128, di, a, ws, grid, k_emit, qbf16, dimpad, qnorm, qa, qb, Qpad, q0, nq, groups, launches)) return -1;
            q0 += nq;
        } else {
            if (launch_shadow_pass(64, di, a, ws, grid, k_emit, qbf16, dimpad, qnorm, qa, qb, Qpad, q0, left, 1, launches)) return -1;
            q0 += left;
        }
    }
    if (tf32_passes(q_big, nullptr, true)) return -1;
    if (a.ev_end) NK_CUDA_OK(cudaEventRecord(a.ev_end, a.stream));
    fp.q_big = q_big; fp.only_if = nullptr;
    filter_finish_kernel<<<a.Q, FINISH_THREADS, fsmem, a.stream>>>(fp);
    NK_CUDA_OK(cudaGetLastError());
    if (launches) ++*launches;
    if (q_big) {
        // retry stage, queued behind and skipped on the device unless a BF16 margin buffer overflowed: the same search over
        // the fp32 rows with the (much tighter) TF32 margins
        retry_mark_kernel<<<1, 1, 0, a.stream>>>(ws.flags);
        retry_zero_kernel<<<(QA + 255) / 256, 256, 0, a.stream>>>(ws.flags + 5, reinterpret_cast<unsigned long long *>(ws.keys2), QA);
        NK_CUDA_OK(cudaGetLastError());
        if (launches) *launches += 2;
        if (tf32_passes(0, ws.flags + 5, false)) return -1;
        fp.q_big = 0; fp.only_if = ws.flags + 5;
        filter_finish_kernel<<<a.Q, FINISH_THREADS, fsmem, a.stream>>>(fp);
        NK_CUDA_OK(cudaGetLastError());
        if (launches) ++*launches;
    }
    tc_print_prof(a.stream, num_tiles, grid, (a.dim + BK - 1) / BK);
    // queued behind; every kernel returns at once unless flags[1] was raised on the device
    if (can_fallback) {
        if (scan_tensor_exact(di, a, ws, out_keys, launches, ws.flags + 1, false, false)) return -1;
    } else {
        ScanArgs b = a;  // euclidean / large k: the CUDA-core scan is the exact twin
        b.only_if = ws.flags + 1;
        b.ev_begin = b.ev_end = nullptr;
        b.main_launches = nullptr;
        if (scan_simt(di, b, ws, out_keys, launches)) return -1;
    }
    return 0;
}

}  // namespace nk
