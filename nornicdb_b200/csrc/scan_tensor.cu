// scan_tensor.cu — fused distance + top-k scan on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
// For Q > ~16 queries the CUDA-core scan stops being HBM-bound (fp32 FMA ridge is ~10 flop/byte, the
// batch needs Q/2 flop/byte), so the Q x N^T contraction moves to tcgen05.mma.  fp32 inputs on the
// tensor cores mean TF32 (10-bit mantissa), which cannot meet the 1e-4 / identical-index-set parity bar
// by itself, so each product is the 3xTF32 split  x*q ~= xh*qh + xl*qh + xh*ql  (xh = rna_tf32(x),
// xl = x - xh), which restores ~2^-21 relative accuracy per product with fp32 accumulation in TMEM.
//
// One persistent CTA per SM, warp-specialised (16 warps):
//   warp 0      TMA producer: corpus K-slabs [256 rows x 32 floats] (128B-swizzled) + the query slabs
//               (pre-split hi/lo, [64 x 32 floats] each) into a 4-stage shared-memory ring;
//   warps 4-11  "split" warps, one thread per corpus row: read the row's 128 B of the slab from shared
//               memory (conflict-free thanks to the 128B swizzle), form xh / xl, accumulate |x|^2 on the
//               side (cosine needs it - no separate norm pass over the corpus), and store xh / xl with
//               tcgen05.st into a 2-stage TMEM ring: the corpus is the MMA's A operand FROM TENSOR MEMORY,
//               so shared-memory bandwidth is spent once per corpus byte, not three times;
//   warp 1      MMA issuer (one thread): per slab 2 M-tiles x 4 k-steps x 3 MMAs (M=128 rows, N=64 queries,
//               K=8), D accumulates in TMEM (double-buffered per row tile);
//   warps 12-15 epilogue: tcgen05.ld the [128 rows x 64 queries] accumulators (thread = corpus row),
//               scale by 1/|x| for cosine, compare against each query's running threshold and append the
//               few survivors to the per-(CTA, query) candidate buffers; prune with a bitonic sort when a
//               buffer could overflow.  Distances never go to memory.
// Per-CTA best-k lists are folded by merge_keys(), exactly like the CUDA-core scan.
//
// Algorithmic HBM traffic per launch: n*dim*4 (corpus, once) + small query re-reads served from L2.
#include <cuda.h>
#include <stdlib.h>

#include "kernels.cuh"
#include "ptx_sm100.cuh"

namespace nk {

namespace tc {
constexpr int THREADS = 512;
constexpr int ROWS = 256;        // corpus rows per tile (2 M-tiles of 128)
constexpr int QT = 64;           // queries per launch (MMA N)
constexpr int BK = 32;           // floats per K-slab = one 128-byte swizzle row
constexpr int ASTAGES = 4;       // corpus-slab ring (freed by the split warps as soon as they have read it)
constexpr int BSTAGES = 5;       // query-slab ring (freed when the MMAs that read it complete)
constexpr int TSTAGES = 3;       // TMEM A-operand ring depth
constexpr int A_BYTES = ROWS * BK * 4;   // 32 KB
constexpr int B_BYTES = QT * BK * 4;     // 8 KB (hi) + 8 KB (lo)
constexpr int BST_BYTES = 2 * B_BYTES;              // hi + lo
constexpr int RING_BYTES = ASTAGES * A_BYTES + BSTAGES * BST_BYTES;  // 160 KB + 48 KB
constexpr int TMEM_COLS = 512;
constexpr int ACC_COL = 0;       // [mtile] x 64 columns -> 128 columns (single-buffered, drained per M-tile)
constexpr int A_COL = 128;       // [tstage][mtile][hi|lo] x 32 columns -> 384 columns
constexpr int EPI_THREADS = 128;
constexpr int EPI_BAR = 1;
constexpr int SPLIT_WARP0 = 4, EPI_WARP0 = 12;

struct Params {
    uint32_t n, dim, nslab;
    uint64_t row_base;
    uint32_t q0, nq, k;
    int metric;
    int P;
    uint64_t *cand;     // [grid][QT][P]
    uint64_t *partial;  // [Q][grid][k]
    int *flags;
    int debug;  // timing experiments only (NK_TC_DEBUG): 1 = no split math, 2 = no MMA, 4 = no epilogue scoring, 8 = no TMEM stores
};

struct __align__(8) Shared {
    uint64_t afull_s[ASTAGES], aempty_s[ASTAGES];  // corpus slab in smem: TMA -> split warps -> TMA
    uint64_t bfull[BSTAGES], bempty[BSTAGES];      // query slabs in smem: TMA -> MMA -> TMA
    uint64_t afull[TSTAGES][2], aempty[TSTAGES][2];  // per (TMEM stage, M-tile)
    uint64_t accfull[2], accempty[2];                // per M-tile
    uint32_t tmem_base;
    float xx[2][ROWS];
    float tau[QT];
    int cnt[QT];
};
}  // namespace tc

// Wait-time instrumentation (NK_TC_DEBUG bit 64): cycles CTA 0 spends blocked at each hand-off.
__device__ long long g_tc_prof[32];
__device__ unsigned long long g_tc_cta_ns[2][160];  // [0] = start, [1] = end (globaltimer) per CTA
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TC_PROF_BEGIN() long long _t0 = prof ? clock64() : 0
#define TC_PROF_END(slot) do { if (prof) { long long _t1 = clock64(); acc_##slot += _t1 - _t0; } } while (0)

// Queries -> (optionally normalised) tf32 hi / lo arrays, zero padded to a multiple of 64 rows.
__global__ void tc_prep_queries_kernel(const float *q, uint32_t Q, uint32_t Qpad, uint32_t dim, int normalise,
                                       float *qhi, float *qlo) {
    uint32_t row = blockIdx.x;
    float inv = 1.0f;
    if (row < Q && normalise) {
        float a = 0.0f;
        for (uint32_t j = threadIdx.x; j < dim; j += blockDim.x) a = fmaf(q[(size_t)row * dim + j], q[(size_t)row * dim + j], a);
        __shared__ float red[32];
#pragma unroll
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
        __syncthreads();
        float t = 0.0f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        inv = t > 0.0f ? 1.0f / sqrtf(t) : 0.0f;  // zero query -> all scores 0 (simd_amd64.go:31-35)
    }
    for (uint32_t j = threadIdx.x; j < dim; j += blockDim.x) {
        float v = row < Q ? q[(size_t)row * dim + j] * inv : 0.0f;
        float h = ptx::cvt_rna_tf32(v);
        qhi[(size_t)row * dim + j] = h;
        qlo[(size_t)row * dim + j] = v - h;
    }
    (void)Qpad;
}

__global__ void __launch_bounds__(tc::THREADS, 1)
knn_scan_tc_kernel(const __grid_constant__ CUtensorMap map_rows, const __grid_constant__ CUtensorMap map_qhi,
                   const __grid_constant__ CUtensorMap map_qlo, tc::Params p) {
    using namespace tc;
    extern __shared__ unsigned char smem_dyn[];
    // 128-byte-swizzled tiles need 1024-byte alignment: align by hand (the launch reserves the slack).
    unsigned char *smem_raw = smem_dyn + ((1024u - (ptx::smem_u32(smem_dyn) & 1023u)) & 1023u);
    // layout: [ASTAGES x A 32K] [BSTAGES x (Bhi 8K | Blo 8K)] [sort buffer P x 8] [Shared]
    unsigned char *a_base = smem_raw;
    unsigned char *b_base = smem_raw + (size_t)ASTAGES * A_BYTES;
    uint64_t *sbuf = reinterpret_cast<uint64_t *>(smem_raw + (size_t)RING_BYTES);
    Shared &sh = *reinterpret_cast<Shared *>(smem_raw + (size_t)RING_BYTES + (size_t)p.P * 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t num_tiles = (p.n + ROWS - 1) / ROWS;
    const bool prof = (p.debug & 64) && blockIdx.x == 0;
    long long acc_a = 0, acc_b = 0, acc_c = 0, acc_d = 0;
    const long long t_start = prof ? clock64() : 0;
    if ((p.debug & 64) && tid == 0 && blockIdx.x < 160) g_tc_cta_ns[0][blockIdx.x] = globaltimer_ns();

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_rows);
        ptx::prefetch_tensormap(&map_qhi);
        ptx::prefetch_tensormap(&map_qlo);
        for (int i = 0; i < ASTAGES; ++i) { ptx::mbar_init(&sh.afull_s[i], 1); ptx::mbar_init(&sh.aempty_s[i], 8); }
        for (int i = 0; i < BSTAGES; ++i) { ptx::mbar_init(&sh.bfull[i], 1); ptx::mbar_init(&sh.bempty[i], 2); }
        for (int i = 0; i < TSTAGES; ++i)
            for (int m = 0; m < 2; ++m) { ptx::mbar_init(&sh.afull[i][m], 4); ptx::mbar_init(&sh.aempty[i][m], 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&sh.accfull[i], 1); ptx::mbar_init(&sh.accempty[i], 4); }
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc(&sh.tmem_base, TMEM_COLS);
    if (tid < QT) {
        sh.tau[tid] = -INFINITY;
        sh.cnt[tid] = 0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = sh.tmem_base;

    if (warp == 0) {
        // ===================================== TMA producer: corpus slabs ========================
        uint32_t g = 0;
        for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (uint32_t j = 0; j < p.nslab; ++j, ++g) {
                const uint32_t s = g % ASTAGES;
                { TC_PROF_BEGIN(); ptx::mbar_wait(&sh.aempty_s[s], ((g / ASTAGES) & 1) ^ 1); TC_PROF_END(a); }
                if (ptx::elect_one_sync()) {
                    ptx::mbar_arrive_expect_tx(&sh.afull_s[s], A_BYTES);
                    ptx::tma_load_2d(&map_rows, &sh.afull_s[s], a_base + (size_t)s * A_BYTES, (int32_t)(j * BK), (int32_t)(tile * ROWS), ptx::CACHE_EVICT_FIRST);
                }
                __syncwarp();
            }
        }
        if (prof && lane == 0) { g_tc_prof[0] = acc_a; g_tc_prof[1] = clock64() - t_start; }
    } else if (warp == 3) {
        // ===================================== TMA producer: query slabs (L2-resident) ===========
        uint32_t g = 0;
        for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (uint32_t j = 0; j < p.nslab; ++j, ++g) {
                const uint32_t s = g % BSTAGES;
                { TC_PROF_BEGIN(); ptx::mbar_wait(&sh.bempty[s], ((g / BSTAGES) & 1) ^ 1); TC_PROF_END(a); }
                if (ptx::elect_one_sync()) {
                    unsigned char *st = b_base + (size_t)s * BST_BYTES;
                    ptx::mbar_arrive_expect_tx(&sh.bfull[s], BST_BYTES);
                    ptx::tma_load_2d(&map_qhi, &sh.bfull[s], st, (int32_t)(j * BK), (int32_t)p.q0, ptx::CACHE_EVICT_LAST);
                    ptx::tma_load_2d(&map_qlo, &sh.bfull[s], st + B_BYTES, (int32_t)(j * BK), (int32_t)p.q0, ptx::CACHE_EVICT_LAST);
                }
                __syncwarp();
            }
        }
        if (prof && lane == 0) { g_tc_prof[2] = acc_a; }
    } else if (warp == 1 || warp == 2) {
        // ===================================== MMA issuers (one warp per M-tile) ==================
        // Two issuing warps: the per-slab serial overhead of one (barrier polls, fences, commits) overlaps with
        // the other's MMAs, so the shallow tcgen05 queue never drains.  tcgen05.commit is per issuing thread.
        {
            const uint32_t m = warp - 1;
            const uint32_t idesc = ptx::make_idesc_tf32(128, QT);
            const uint32_t d = tmem + ACC_COL + m * QT;
            uint32_t g = 0, it = 0;
            for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                for (uint32_t j = 0; j < p.nslab; ++j, ++g) {
                    const uint32_t s = g % BSTAGES, ts = g % TSTAGES;
                    if (j == 0) { TC_PROF_BEGIN(); ptx::mbar_wait(&sh.accempty[m], (it & 1) ^ 1); TC_PROF_END(b); }  // epilogue drained
                    { TC_PROF_BEGIN(); ptx::mbar_wait(&sh.bfull[s], (g / BSTAGES) & 1); TC_PROF_END(a); }  // query slabs landed
                    { TC_PROF_BEGIN(); ptx::mbar_wait(&sh.afull[ts][m], (g / TSTAGES) & 1); TC_PROF_END(c); }  // xh / xl are in TMEM
                    ptx::tc_fence_after();
                    if (ptx::elect_one_sync()) {
                        const uint32_t bhi = ptx::smem_u32(b_base + (size_t)s * BST_BYTES);
                        const uint64_t dhi = ptx::make_smem_desc_sw128(bhi), dlo = ptx::make_smem_desc_sw128(bhi + B_BYTES);
                        const uint32_t ahi = tmem + A_COL + ((ts * 2 + m) * 2) * BK, alo = ahi + BK;
                        if (!(p.debug & 2)) {
#pragma unroll
                            for (uint32_t kk = 0; kk < BK / 8; ++kk) {
                                // smallest terms first; K advance = 8 floats = 32 B = 2 descriptor units / 8 TMEM columns
                                ptx::mma_tf32_ts(d, alo + kk * 8, dhi + kk * 2, idesc, (j | kk) != 0);
                                ptx::mma_tf32_ts(d, ahi + kk * 8, dlo + kk * 2, idesc, 1);
                                ptx::mma_tf32_ts(d, ahi + kk * 8, dhi + kk * 2, idesc, 1);
                            }
                        }
                        ptx::tc_commit(&sh.aempty[ts][m]);                        // TMEM A slot of this M-tile reusable
                        ptx::tc_commit(&sh.bempty[s]);                            // query slabs: both issuers must be done
                        if (j + 1 == p.nslab) ptx::tc_commit(&sh.accfull[m]);     // accumulator of this M-tile complete
                    }
                    __syncwarp();
                }
            }
        }
        if (prof && lane == 0 && warp == 1) { g_tc_prof[4] = acc_a; g_tc_prof[5] = acc_b; g_tc_prof[6] = acc_c; g_tc_prof[7] = clock64() - t_start; }
    } else if (warp >= SPLIT_WARP0 && warp < SPLIT_WARP0 + 8) {
        // ===================================== split warps ======================================
        const uint32_t m = (warp - SPLIT_WARP0) >> 2, quad = warp & 3;
        const uint32_t r = m * 128 + quad * 32 + lane;  // row within the tile
        const uint32_t lane_base = (quad * 32u) << 16;
        uint32_t g = 0, it = 0;
        for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            uint64_t xx2 = 0;  // two partial sums of |x|^2 (packed f32x2)
            for (uint32_t j = 0; j < p.nslab; ++j, ++g) {
                const uint32_t s = g % ASTAGES, ts = g % TSTAGES;
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
                        const uint64_t l2 = ptx::sub_f32x2(x2, ptx::pack2(h0, h1));  // exact residual x - xh
                        xx2 = ptx::fma_f32x2(x2, x2, xx2);                           // |x|^2 on the side (cosine)
                        hi[c * 4 + u] = h0; hi[c * 4 + u + 1] = h1;
                        lo[c * 4 + u] = (uint32_t)l2; lo[c * 4 + u + 1] = (uint32_t)(l2 >> 32);
                    }
                }
                // the slab now lives in registers: hand the smem slot straight back to the TMA producer
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&sh.aempty_s[s]);
                if (prof) acc_c += clock64() - _tw;
                { TC_PROF_BEGIN(); ptx::mbar_wait(&sh.aempty[ts][m], ((g / TSTAGES) & 1) ^ 1); TC_PROF_END(b); }
                _tw = prof ? clock64() : 0;
                ptx::tc_fence_after();
                const uint32_t acol = tmem + lane_base + A_COL + ((ts * 2 + m) * 2) * BK;
                if (!(p.debug & 8)) {
                    ptx::tmem_st_32x32b_x32(acol, hi);
                    ptx::tmem_st_32x32b_x32(acol + BK, lo);
                    ptx::tmem_wait_st();
                } else if (hi[0] == 0x12345678u && lo[3] == 77u) {
                    sh.xx[0][r] = 1.0f;
                }
                if (j + 1 == p.nslab) sh.xx[it & 1][r] = __uint_as_float((uint32_t)xx2) + __uint_as_float((uint32_t)(xx2 >> 32));  // published by the afull arrive below
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
        const uint32_t quad = warp & 3, gtid = tid - EPI_WARP0 * 32;
        const uint32_t lane_base = (quad * 32u) << 16;
        uint64_t *my_cand = p.cand + (size_t)blockIdx.x * QT * p.P;
        const int prune_at = p.P - ROWS;
        uint32_t it = 0;
        for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
#pragma unroll 1
            for (uint32_t m = 0; m < 2; ++m) {
                const uint32_t rt = m * 128 + quad * 32 + lane;
                const uint32_t row = tile * ROWS + rt;
                { TC_PROF_BEGIN(); ptx::mbar_wait(&sh.accfull[m], it & 1); TC_PROF_END(a); }
                ptx::tc_fence_after();
                // drain the accumulator first and hand it back, then score from registers
                uint32_t v0[32], v1[32];
                ptx::tmem_ld_32x32b_x32(tmem + lane_base + ACC_COL + m * QT, v0);
                ptx::tmem_ld_32x32b_x32(tmem + lane_base + ACC_COL + m * QT + 32, v1);
                ptx::tmem_wait_ld();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&sh.accempty[m]);
                float scale = 1.0f;
                if (p.metric == NK_METRIC_COSINE) {
                    float x2 = sh.xx[it & 1][rt];
                    scale = x2 > 0.0f ? 1.0f / sqrtf(x2) : 0.0f;
                }
                if (row < p.n && !(p.debug & 4)) {
                    // Compact compare pass -> 64-bit mask of columns worth buffering (NaN passes and is mapped to
                    // -inf below); the rare pushes run in a small out-of-line loop so the hot code stays a few
                    // hundred instructions (a fully unrolled push per column was ~40 KB of SASS: I-cache thrash).
                    uint32_t pass0 = 0, pass1 = 0;
#pragma unroll
                    for (uint32_t c = 0; c < 32; ++c) {
                        pass0 |= !(__uint_as_float(v0[c]) * scale < sh.tau[c]) ? (1u << c) : 0u;
                        pass1 |= !(__uint_as_float(v1[c]) * scale < sh.tau[32 + c]) ? (1u << c) : 0u;
                    }
                    uint64_t pass = (uint64_t)pass0 | ((uint64_t)pass1 << 32);
                    if (p.nq < 64) pass &= (1ull << p.nq) - 1ull;
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
                        float sc = __uint_as_float(bits) * scale;
                        if (sc != sc) sc = -INFINITY;
                        if (sc >= sh.tau[c]) {
                            int pos = atomicAdd(&sh.cnt[c], 1);
                            if (pos < p.P) my_cand[(size_t)c * p.P + pos] = make_key(sc, (uint32_t)(p.row_base + row));
                            else atomicExch(p.flags, 1);
                        }
                    }
                }
            }
            // prune any buffer that could overflow during the next tile
            group_sync(EPI_BAR, EPI_THREADS);  // every push of this tile is visible
            if (p.P == 512) {
                // register-resident warp selection; the 4 epilogue warps prune different queries concurrently
                for (uint32_t qi = quad; qi < p.nq; qi += 4)
                    if (sh.cnt[qi] > prune_at)
                        warp_prune<16>(my_cand + (size_t)qi * p.P, &sh.cnt[qi], &sh.tau[qi], p.k, lane, nullptr);
                group_sync(EPI_BAR, EPI_THREADS);
            } else {
                uint64_t need = 0;
                for (uint32_t qi = 0; qi < p.nq; ++qi) need |= (uint64_t)(sh.cnt[qi] > prune_at ? 1 : 0) << qi;
                group_sync(EPI_BAR, EPI_THREADS);
                if (need) {
                    for (uint32_t qi = 0; qi < p.nq; ++qi)
                        if (need & (1ull << qi))
                            group_prune(my_cand + (size_t)qi * p.P, p.P, &sh.cnt[qi], &sh.tau[qi], p.k, sbuf, p.P, gtid, EPI_THREADS, EPI_BAR);
                }
            }
        }
        // emit this CTA's best k per query
        if (p.P == 512) {
            for (uint32_t qi = quad; qi < p.nq; qi += 4)
                warp_prune<16>(my_cand + (size_t)qi * p.P, &sh.cnt[qi], &sh.tau[qi], p.k, lane,
                               p.partial + ((size_t)(p.q0 + qi) * gridDim.x + blockIdx.x) * p.k);
        } else {
            for (uint32_t qi = 0; qi < p.nq; ++qi) {
                group_prune(my_cand + (size_t)qi * p.P, p.P, &sh.cnt[qi], &sh.tau[qi], p.k, sbuf, p.P, gtid, EPI_THREADS, EPI_BAR);
                uint64_t *dst = p.partial + ((size_t)(p.q0 + qi) * gridDim.x + blockIdx.x) * p.k;
                for (uint32_t i = gtid; i < p.k; i += EPI_THREADS) dst[i] = sbuf[i];
                group_sync(EPI_BAR, EPI_THREADS);
            }
        }
    }

    if (prof && tid == EPI_WARP0 * 32) { g_tc_prof[17] = acc_a; g_tc_prof[18] = clock64() - t_start; g_tc_prof[19] = (long long)num_tiles; }
    // ---- teardown ----------------------------------------------------------------------------------
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem, TMEM_COLS);
    if ((p.debug & 64) && tid == 0 && blockIdx.x < 160) g_tc_cta_ns[1][blockIdx.x] = globaltimer_ns();
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

// 2-D fp32 row-major [rows x dim] tensor, box = [box_rows x 32 floats], 128-byte swizzle, OOB -> 0.
static int make_map(CUtensorMap *m, const void *base, uint64_t rows, uint32_t dim, uint32_t box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return -1;
    }
    cuuint64_t gdim[2] = {dim, rows};
    cuuint64_t gstride[1] = {(cuuint64_t)dim * 4};
    cuuint32_t box[2] = {(cuuint32_t)tc::BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu dim=%u", (int)r, (unsigned long long)rows, dim);
        return -1;
    }
    return 0;
}

bool scan_tensor_supported(const DeviceInfo &di, const ScanArgs &a) {
    if (di.cc < 100) return false;
    if (a.dtype != NK_DTYPE_F32) return false;                      // fp16 corpus: CUDA-core scan (HBM-bound at Q=1)
    if (a.metric == NK_METRIC_EUCLIDEAN) return false;              // needs exact rescoring of |x|^2+|q|^2-2xq; SIMT path
    if (a.dim % 4 != 0 || a.dim < 32) return false;                 // TMA: 16-byte global stride
    if ((reinterpret_cast<uintptr_t>(a.rows) & 15) != 0) return false;
    if (a.k > NK_MAX_K || a.n == 0) return false;
    if (next_pow2(a.k + tc::ROWS + 1) > 1024) return false;          // sort buffer must fit beside the rings (k <= 767)
    return true;
}

int scan_tensor(const DeviceInfo &di, const ScanArgs &a, Workspace &ws, uint64_t *out_keys, uint64_t *launches) {
    using namespace tc;
    if (a.n == 0 || a.Q == 0 || a.k == 0) return 0;
    if (!scan_tensor_supported(di, a)) {
        set_error("tensor path: unsupported shape");
        return -1;
    }
    const uint32_t Qpad = (a.Q + QT - 1) / QT * QT;
    const uint32_t nslab = (a.dim + BK - 1) / BK;
    const uint32_t P = next_pow2(a.k + ROWS + 1) < 512 ? 512 : next_pow2(a.k + ROWS + 1);
    const uint32_t num_tiles = (a.n + ROWS - 1) / ROWS;
    uint32_t grid = (uint32_t)di.num_sms;
    if (grid > num_tiles) grid = num_tiles;

    if (ws_reserve((void **)&ws.qaux, &ws.qaux_bytes, (size_t)2 * Qpad * a.dim * 4)) return -1;
    if (ws_reserve((void **)&ws.cand, &ws.cand_bytes, (size_t)grid * QT * P * 8)) return -1;
    if (ws_reserve((void **)&ws.partial, &ws.partial_bytes, (size_t)a.Q * grid * a.k * 8)) return -1;
    float *qhi = ws.qaux, *qlo = ws.qaux + (size_t)Qpad * a.dim;

    tc_prep_queries_kernel<<<Qpad, 256, 0, a.stream>>>(a.queries, a.Q, Qpad, a.dim, a.metric == NK_METRIC_COSINE, qhi, qlo);
    NK_CUDA_OK(cudaGetLastError());
    if (launches) ++*launches;

    CUtensorMap map_rows, map_qhi, map_qlo;
    if (make_map(&map_rows, a.rows, a.n, a.dim, ROWS)) return -1;
    if (make_map(&map_qhi, qhi, Qpad, a.dim, QT)) return -1;
    if (make_map(&map_qlo, qlo, Qpad, a.dim, QT)) return -1;

    const size_t smem = (size_t)RING_BYTES + (size_t)P * 8 + sizeof(Shared) + 1024;
    if (smem > di.max_smem_optin) {
        set_error("tensor path needs %zu B shared memory (> %zu)", smem, di.max_smem_optin);
        return -1;
    }
    NK_CUDA_OK(cudaFuncSetAttribute(knn_scan_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    if (a.ev_begin) NK_CUDA_OK(cudaEventRecord(a.ev_begin, a.stream));
    for (uint32_t q0 = 0; q0 < a.Q; q0 += QT) {
        Params p;
        p.n = a.n; p.dim = a.dim; p.nslab = nslab; p.row_base = a.row_base;
        p.q0 = q0; p.nq = a.Q - q0 < (uint32_t)QT ? a.Q - q0 : (uint32_t)QT; p.k = a.k;
        p.metric = a.metric; p.P = (int)P; p.cand = ws.cand; p.partial = ws.partial; p.flags = ws.flags;
        {
            const char *dbg = getenv("NK_TC_DEBUG");
            p.debug = dbg ? atoi(dbg) : 0;
        }
        knn_scan_tc_kernel<<<grid, THREADS, smem, a.stream>>>(map_rows, map_qhi, map_qlo, p);
        NK_CUDA_OK(cudaGetLastError());
        if (launches) ++*launches;
        if (a.main_launches) ++*a.main_launches;
    }
    if (a.ev_end) NK_CUDA_OK(cudaEventRecord(a.ev_end, a.stream));
    if (merge_keys(ws.partial, grid, a.k, (size_t)grid * a.k, a.Q, a.k, out_keys, a.stream)) return -1;
    if (launches) ++*launches;
    {
        const char *dbg = getenv("NK_TC_DEBUG");
        if (dbg && (atoi(dbg) & 64)) {
            long long h[32];
            cudaStreamSynchronize(a.stream);
            cudaMemcpyFromSymbol(h, g_tc_prof, sizeof(h));
            uint32_t slabs = ((num_tiles + grid - 1) / grid) * nslab;
            {
                unsigned long long c[2][160];
                cudaMemcpyFromSymbol(c, g_tc_cta_ns, sizeof(c));
                unsigned long long t0 = ~0ull;
                for (uint32_t i = 0; i < grid && i < 160; ++i) t0 = c[0][i] < t0 ? c[0][i] : t0;
                fprintf(stderr, "[tc per-CTA end times, us since first start]");
                for (uint32_t i = 0; i < grid && i < 160; ++i) fprintf(stderr, "%s%.0f", i % 16 ? " " : "\n  ", (double)(c[1][i] - t0) / 1e3);
                fprintf(stderr, "\n");
            }
            fprintf(stderr, "[tc prof CTA0, ~%u slabs] total %lld cyc (%.0f/slab)\n  tmaA wait aempty_s %lld | tmaB wait bempty %lld\n"
                    "  mma: wait bfull %lld, wait accempty %lld, wait afull %lld, total %lld\n"
                    "  split m0: wait afull_s %lld, wait aempty %lld, read+math %lld, st+arrive %lld, total %lld\n"
                    "  split m1: wait afull_s %lld, wait aempty %lld, read+math %lld, st+arrive %lld\n  epi: wait accfull %lld total %lld\n",
                    slabs, h[7], (double)h[7] / slabs, h[0], h[2], h[4], h[5], h[6], h[7], h[8], h[9], h[10], h[11], h[12], h[13], h[14], h[15], h[16], h[17], h[18]);
        }
    }
    return 0;
}

}  // namespace nk
