// scan_tensor_shadow.cu — filter scan over the BF16 SHADOW of the corpus (tcgen05 / TMEM / TMA).
//
// An fp32 shard optionally carries a bf16 copy of itself ("shadow", +50% HBM, built once at upload and kept in step by
// append / update / remove) plus two floats per row: |x|^2 and |x - bf16(x)|^2.  The filter scan streams the SHADOW —
// half the bytes of the fp32 corpus, so the HBM-bound regime (Q <= 128) runs at twice the queries per second — and
// exact fp32 re-scoring of the few survivors (filter_finish_kernel, reading the fp32 rows) makes the results identical
// to a full-precision scan.  This is the default filter where a shadow exists; the TF32 scan over the fp32 rows
// (scan_tensor.cu) is its retry stage and serves shards without a shadow (caller-owned device rows, fp16 corpora).
//
// Rigorous bound.  With x = xb + dx, q = qb + dq (xb, qb the BF16 roundings):
//     x.q - xb.qb = dx.qb + x.dq   =>   |s_hat - s| <= |dx| |qb| + |x| |dq| + acc_c |x| |q|   (fp32 accumulation)
// |dx| per row and |qb|, |dq| per query are the MEASURED rounding residues (~0.4 * 2^-8 relative for typical data, not
// the worst case 2^-8), so the margin is as tight as Cauchy-Schwarz allows.  Rows whose upper bound reaches the running
// k-th bound are kept; margin overflow raises the device flag and the TF32 filter, then the exact kernels, retry.
//
// Both operands are plain bf16 tiles in shared memory, so there is no conversion stage and TMEM holds only
// accumulators — double-buffered: the epilogue of tile t overlaps the MMAs of tile t+1.
//   warp 0      TMA: shadow slabs [256 rows x 64 bf16] (32 KB, 128B-swizzled), 4-stage ring;
//   warp 3      TMA: query slabs [QT x 64 bf16] (8 / 16 KB), L2-resident;
//   warp 1      MMA issuer: per slab 2 M-tiles x 4 tcgen05.mma.kind::f16 (M=128, N=QT, K=16), A and B from smem;
//   warps 4-11  epilogue, one warpgroup per 128-row M-tile: row norms from global, tcgen05.ld 64 columns at a time,
//               bound / compare / push, warp-level prunes.
// Shared-memory traffic per [256 rows x 128 queries x 64 dims]: 112 KB (A 32 w + 32 r, B 16 w + 32 r) for twice the K
// extent of a TF32 slab: half the bytes per product of the TF32 kernel, which is shared-memory-bandwidth bound at 128
// query columns.  Algorithmic HBM bytes per launch: n * dimpad * 2 (shadow) + n * 8 (norms).
#include <cuda.h>
#include <stdlib.h>

#include "kernels.cuh"
#include "ptx_sm100.cuh"
#include "scan_tensor_shared.cuh"

namespace nk {

namespace sb {
using namespace tc;
constexpr int NTHREADS = 384;
constexpr int ROWS = 256;                 // corpus rows per tile (2 M-tiles of 128)
constexpr int QT_MAX = 128;
constexpr int BKB = 64;                   // bf16 per row per slab = one 128-byte swizzle row
constexpr int ASTAGES = 4;
constexpr int A_BYTES = ROWS * 128;       // 32 KB
constexpr int MAX_BSTAGES = 8;
constexpr int EPI_WARP0 = 4, EPI_WARPS = 8, EPI_NT = EPI_WARPS * 32;
constexpr int PB = tc::P_SHADOW;          // candidate buffer slots per (CTA, query)
constexpr int PRUNE_LANE = PB / 32;       // keys per lane of a warp prune

template <int QT> struct Cfg {
    static constexpr int B_BYTES = QT * 128;
    static constexpr int BSTAGES = (80 * 1024) / B_BYTES > MAX_BSTAGES ? MAX_BSTAGES : (80 * 1024) / B_BYTES;
    static constexpr int RING_BYTES = ASTAGES * A_BYTES + BSTAGES * B_BYTES;
    static_assert(4 * QT <= TMEM_COLS, "two accumulator buffers of two M-tiles");
};

struct __align__(8) Shared {
    uint64_t afull[ASTAGES], aempty[ASTAGES];
    uint64_t bfull[MAX_BSTAGES], bempty[MAX_BSTAGES];
    uint64_t accfull[2], accempty[2][2];          // per accumulator buffer (, M-tile)
    uint32_t tmem_base;
    unsigned int maxxx, max_ra, max_rb;            // running maxima (float bits) of |x|^2 and the per-row bound factors
    float tau[QT_MAX];
    float qn[QT_MAX];                              // |q| (1 for cosine)
    float qa[QT_MAX], qb[QT_MAX];                  // bound(row, q) = ra(row) qa[q] + rb(row) qb[q]
    int cnt[QT_MAX];
    uint32_t ethr[QT_MAX];                         // emission: score-word threshold per query
    int eoff[QT_MAX + 1];                          // emission: prefix of 32-entry chunk counts
};
}  // namespace sb

// Queries -> bf16 [Qpad x dimpad] (zero padded), cosine normalises first.  Per query: qnorm = |q| (1 for cosine),
// qa = |bf16(q)|, qb = |q - bf16(q)| + acc_c |q|  (both inflated by 1e-4 for the fp32 rounding of the sums).
__global__ void bf16_prep_queries_kernel(const float *q, uint32_t Q, uint32_t dim, uint32_t dimpad, int normalise, float acc_c,
                                         uint16_t *out, float *qnorm, float *qa, float *qb) {
    const uint32_t row = blockIdx.x;
    __shared__ float red[3][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float t = 0.0f;
    if (row < Q) {
        float a = 0.0f;
        for (uint32_t j = threadIdx.x; j < dim; j += blockDim.x) a = fmaf(q[(size_t)row * dim + j], q[(size_t)row * dim + j], a);
#pragma unroll
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) red[0][w] = a;
        __syncthreads();
        for (int i = 0; i < nw; ++i) t += red[0][i];
        __syncthreads();
    }
    const float nrm = sqrtf(t);
    const float inv = normalise ? (t > 0.0f ? 1.0f / nrm : 0.0f) : 1.0f;  // zero query -> all cosine scores 0
    const float qn = normalise ? (t > 0.0f ? 1.0f : 0.0f) : nrm;
    float hh = 0.0f, dd = 0.0f;
    for (uint32_t j = threadIdx.x; j < dimpad; j += blockDim.x) {
        const float v = (row < Q && j < dim) ? q[(size_t)row * dim + j] * inv : 0.0f;
        const uint16_t b = ptx::f32_to_bf16_bits(v);
        out[(size_t)row * dimpad + j] = b;
        const float vb = __uint_as_float((uint32_t)b << 16);
        hh = fmaf(vb, vb, hh);
        dd = fmaf(v - vb, v - vb, dd);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        hh += __shfl_xor_sync(0xffffffffu, hh, o);
        dd += __shfl_xor_sync(0xffffffffu, dd, o);
    }
    if (lane == 0) { red[1][w] = hh; red[2][w] = dd; }
    __syncthreads();
    if (threadIdx.x == 0) {
        hh = dd = 0.0f;
        for (int i = 0; i < nw; ++i) { hh += red[1][i]; dd += red[2][i]; }
        qnorm[row] = qn;
        qa[row] = sqrtf(hh) * 1.0001f;
        qb[row] = (sqrtf(dd) + acc_c * qn) * 1.0001f;
    }
}

// fp32 rows -> bf16 shadow rows (row stride dimpad, zero padded) + |x|^2 and |x - bf16(x)|^2 per row.  One warp per row.
__global__ void build_shadow_kernel(const float *rows, uint64_t n, uint32_t dim, uint32_t dimpad, uint16_t *shadow, float *xnorm2,
                                    float *dnorm2) {
    const uint64_t row = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const float *x = rows + row * dim;
    uint16_t *o = shadow + row * dimpad;
    float xx = 0.0f, dd = 0.0f;
    if ((dim & 3u) == 0 && ((reinterpret_cast<uintptr_t>(rows) & 15) == 0)) {
        for (uint32_t j = lane * 4; j < dimpad; j += 128) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < dim) v = __ldg(reinterpret_cast<const float4 *>(x + j));
            const uint32_t b01 = ptx::pack_bf16x2(v.x, v.y), b23 = ptx::pack_bf16x2(v.z, v.w);
            *reinterpret_cast<uint2 *>(o + j) = make_uint2(b01, b23);
            const float e0 = v.x - __uint_as_float(b01 << 16), e1 = v.y - __uint_as_float(b01 & 0xffff0000u);
            const float e2 = v.z - __uint_as_float(b23 << 16), e3 = v.w - __uint_as_float(b23 & 0xffff0000u);
            xx = fmaf(v.x, v.x, xx); xx = fmaf(v.y, v.y, xx); xx = fmaf(v.z, v.z, xx); xx = fmaf(v.w, v.w, xx);
            dd = fmaf(e0, e0, dd); dd = fmaf(e1, e1, dd); dd = fmaf(e2, e2, dd); dd = fmaf(e3, e3, dd);
        }
    } else {
        for (uint32_t j = lane; j < dimpad; j += 32) {
            const float v = j < dim ? x[j] : 0.0f;
            const uint16_t b = ptx::f32_to_bf16_bits(v);
            o[j] = b;
            const float e = v - __uint_as_float((uint32_t)b << 16);
            xx = fmaf(v, v, xx);
            dd = fmaf(e, e, dd);
        }
    }
#pragma unroll
    for (int s = 16; s; s >>= 1) {
        xx += __shfl_xor_sync(0xffffffffu, xx, s);
        dd += __shfl_xor_sync(0xffffffffu, dd, s);
    }
    if (lane == 0) { xnorm2[row] = xx; dnorm2[row] = dd; }
}

int build_shadow(const float *rows, uint64_t first, uint64_t count, uint32_t dim, uint32_t dimpad, void *shadow, float *xnorm2,
                 float *dnorm2, cudaStream_t stream) {
    if (count == 0) return 0;
    const unsigned int blocks = (unsigned int)((count + 7) / 8);
    build_shadow_kernel<<<blocks, 256, 0, stream>>>(rows + first * dim, count, dim, dimpad,
                                                    static_cast<uint16_t *>(shadow) + first * dimpad, xnorm2 + first, dnorm2 + first);
    NK_CUDA_OK(cudaGetLastError());
    return 0;
}

// Debug timeline (NK_TC_DEBUG & 64): per-CTA globaltimer stamps, printed by launch_shadow_pass_t after a synchronize.
//   [0] entry  [1] set-up done  [2] first accumulator ready  [3] last tile's epilogue done  [4] exit
//   [5] ns in the prune section (epilogue warp 0)  [6] tile index of the first prune  [7] ns waiting for accumulators
__device__ unsigned long long g_sb_prof[256][16];  // [8] all warps past the tile loop  [9] warp 0 done emitting  [10] live chunks  [11] entries emitted by warp 0
__device__ __forceinline__ unsigned long long sb_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    return t;
}

template <int QT, bool DUMP>
__global__ void __launch_bounds__(sb::NTHREADS, 1)
knn_scan_shadow_kernel(const __grid_constant__ CUtensorMap map_rows, const __grid_constant__ CUtensorMap map_q, tc::Params p) {
    using namespace sb;
    using C = Cfg<QT>;
    pdl_trigger();  // the finish kernel's launch may begin now (it waits for this grid's completion before reading anything)
    if (p.only_if && *p.only_if == 0) return;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem_raw = smem_dyn + ((1024u - (ptx::smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char *a_base = smem_raw;
    unsigned char *b_base = smem_raw + (size_t)ASTAGES * A_BYTES;
    Shared &sh = *reinterpret_cast<Shared *>(smem_raw + (size_t)C::RING_BYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool prof = (p.debug & 64) && blockIdx.x < 256;
    if (prof && tid == 0) { g_sb_prof[blockIdx.x][0] = sb_now(); g_sb_prof[blockIdx.x][6] = ~0ull; }
    const uint32_t num_tiles = (p.n + ROWS - 1) / ROWS;
    // query groups: CTA b serves query block (b % G) over the tile subset (b / G); siblings share tiles through L2
    const uint32_t grp = blockIdx.x % p.qgroups, sub = blockIdx.x / p.qgroups, sgrid = gridDim.x / p.qgroups;
    const uint32_t q0 = p.q0 + grp * QT, qpad_off = p.qpad_off + grp * QT;
    const uint32_t nq = p.nq - grp * QT < (uint32_t)QT ? p.nq - grp * QT : (uint32_t)QT;
    const uint64_t a_policy = p.qgroups > 1 ? ptx::CACHE_EVICT_NORMAL : ptx::CACHE_EVICT_FIRST;
    const uint32_t nslab = p.nslab;  // dimpad / 64

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_rows);
        ptx::prefetch_tensormap(&map_q);
        for (int i = 0; i < ASTAGES; ++i) { ptx::mbar_init(&sh.afull[i], 1); ptx::mbar_init(&sh.aempty[i], 1); }
        for (int i = 0; i < C::BSTAGES; ++i) { ptx::mbar_init(&sh.bfull[i], 1); ptx::mbar_init(&sh.bempty[i], 1); }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(&sh.accfull[b], 1);
            ptx::mbar_init(&sh.accempty[b][0], 4);
            ptx::mbar_init(&sh.accempty[b][1], 4);
        }
        sh.maxxx = 0u; sh.max_ra = 0u; sh.max_rb = 0u;
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc(&sh.tmem_base, TMEM_COLS);
    if (tid < QT) {
        // start threshold: the caller's score floor and, when the prep kernel sampled the shard, a lower bound of the
        // query's k-th best score (no flood tiles then)
        float t0 = p.min_score;
        if (p.presampled && (uint32_t)tid < nq) {
            const uint32_t g = __ldcg(p.gtau + q0 + tid);
            if (g) t0 = fmaxf(t0, ord_to_float(g));
        }
        sh.tau[tid] = t0;
        sh.cnt[tid] = 0;
        sh.qn[tid] = p.qnorm[qpad_off + tid];
        sh.qa[tid] = p.qa[qpad_off + tid];
        sh.qb[tid] = p.qb[qpad_off + tid];
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = sh.tmem_base;
    const bool flood = !p.presampled;  // first two tiles: every (row, query) pair is placed directly
    if (prof && tid == 0) g_sb_prof[blockIdx.x][1] = sb_now();

    if (warp == 0) {
        // ===================================== TMA producer: shadow slabs =========================
        uint32_t g = 0;
        for (uint32_t tile = sub; tile < num_tiles; tile += sgrid) {
            for (uint32_t j = 0; j < nslab; ++j, ++g) {
                const uint32_t s = g % ASTAGES;
                ptx::mbar_wait(&sh.aempty[s], ((g / ASTAGES) & 1) ^ 1);
                if (ptx::elect_one_sync()) {
                    ptx::mbar_arrive_expect_tx(&sh.afull[s], A_BYTES);
                    ptx::tma_load_2d(&map_rows, &sh.afull[s], a_base + (size_t)s * A_BYTES, (int32_t)(j * BKB), (int32_t)(tile * ROWS), a_policy);
                }
                __syncwarp();
            }
        }
    } else if (warp == 3) {
        // ===================================== TMA producer: query slabs (L2-resident) ===========
        uint32_t g = 0;
        for (uint32_t tile = sub; tile < num_tiles; tile += sgrid) {
            for (uint32_t j = 0; j < nslab; ++j, ++g) {
                const uint32_t s = g % C::BSTAGES;
                ptx::mbar_wait(&sh.bempty[s], ((g / C::BSTAGES) & 1) ^ 1);
                if (ptx::elect_one_sync()) {
                    ptx::mbar_arrive_expect_tx(&sh.bfull[s], C::B_BYTES);
                    ptx::tma_load_2d(&map_q, &sh.bfull[s], b_base + (size_t)s * C::B_BYTES, (int32_t)(j * BKB), (int32_t)qpad_off, ptx::CACHE_EVICT_LAST);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =========================================
        const uint32_t idesc = p.op_f16 ? ptx::make_idesc_f16(128, QT) : ptx::make_idesc_bf16(128, QT);
        uint32_t g = 0, it = 0;
        for (uint32_t tile = sub; tile < num_tiles; tile += sgrid, ++it) {
            const uint32_t buf = it & 1;
            ptx::mbar_wait(&sh.accempty[buf][0], ((it >> 1) & 1) ^ 1);  // the epilogue has drained this buffer (two tiles ago)
            ptx::mbar_wait(&sh.accempty[buf][1], ((it >> 1) & 1) ^ 1);
            for (uint32_t j = 0; j < nslab; ++j, ++g) {
                const uint32_t sa = g % ASTAGES, sbq = g % C::BSTAGES;
                ptx::mbar_wait(&sh.bfull[sbq], (g / C::BSTAGES) & 1);
                ptx::mbar_wait(&sh.afull[sa], (g / ASTAGES) & 1);
                ptx::tc_fence_after();
                if (ptx::elect_one_sync()) {
                    const uint64_t bdesc = ptx::make_smem_desc_sw128(ptx::smem_u32(b_base + (size_t)sbq * C::B_BYTES));
                    const uint32_t abase = ptx::smem_u32(a_base + (size_t)sa * A_BYTES);
                    // K advance per MMA = 16 bf16 = 32 B = 2 descriptor units
#pragma unroll
                    for (uint32_t m = 0; m < 2; ++m) {
                        const uint64_t adesc = ptx::make_smem_desc_sw128(abase + m * (A_BYTES / 2));
                        const uint32_t d = tmem + (buf * 2 + m) * QT;
#pragma unroll
                        for (uint32_t kk = 0; kk < 4; ++kk) ptx::mma_bf16_ss(d, adesc + kk * 2, bdesc + kk * 2, idesc, (j | kk) != 0);
                    }
                    ptx::tc_commit(&sh.aempty[sa]);
                    ptx::tc_commit(&sh.bempty[sbq]);
                    if (j + 1 == nslab) ptx::tc_commit(&sh.accfull[buf]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ===================================== epilogue =========================================
        const uint32_t m = (uint32_t)(warp - EPI_WARP0) >> 2, quad = warp & 3, ewarp = warp - EPI_WARP0;
        const uint32_t lane_base = (quad * 32u) << 16;
        const uint32_t rt = m * 128 + quad * 32 + lane;  // row within the tile
        uint64_t *my_cand = p.cand + (size_t)blockIdx.x * QT * PB;
        // capacity rule: a buffer must be prunable below PB - ROWS, or the next tile could overflow it.  Prunes are TRIGGERED
        // much earlier (hidden behind the MMAs): an early, tight threshold keeps the buffers short, so the emission at the
        // end of the launch needs no selection and the finish step reads short lists (small shards: 26 tiles per CTA at
        // N = 1M would otherwise never prune at all).
        const int prune_at = PB - ROWS;
        const int prune_trigger = p.prune_trigger > 0 ? min(prune_at, p.prune_trigger) : min(prune_at, max(192, 4 * (int)p.k));  // (192 / 256 / 384 measured at k = 10: 0.445 / 0.448 / 0.476 ms per 1.25M-row search)
        const bool cosine = p.metric == NK_METRIC_COSINE, euclid = p.metric == NK_METRIC_EUCLIDEAN;
        const bool eprof = prof && ewarp == 0 && lane == 0;
        unsigned long long t_prune = 0, t_wait = 0, t_a = 0;
        uint32_t it = 0;
        for (uint32_t tile = sub; tile < num_tiles; tile += sgrid, ++it) {
            const uint32_t buf = it & 1;
            const uint32_t row = tile * ROWS + rt;
            const bool alive = row < p.n && (!p.mask || ((__ldg(p.mask + (row >> 5)) >> (row & 31)) & 1u));
            // the buffered key is the UPPER bound of the score, bound(row, q) = ra qa[q] + rb qb[q]:
            //   cosine     acc/|x|                        ra = |dx|/|x|   rb = 1
            //   dot        acc                            ra = |dx|       rb = |x|
            //   euclidean  -(|x|^2 + |q|^2 - 2 acc)        ra = 2|dx|      rb = 2|x|   (+ eps (|x|^2 + |q|^2))
            const float x2 = row < p.n ? __ldg(p.xnorm2 + row) : 0.0f;
            const float xn = sqrtf(x2);
            const float dxn = sqrtf((row < p.n && p.dnorm2) ? __ldg(p.dnorm2 + row) : 0.0f) * 1.0001f;  // 16-bit corpus: no residue
            float mul = 1.0f, ra = dxn, rb = xn;
            if (cosine) {
                mul = x2 > 0.0f ? 1.0f / xn : 0.0f;
                ra = dxn * mul * 1.000001f;
                rb = 1.0f;
            } else if (euclid) {
                mul = 2.0f; ra = 2.0f * dxn; rb = 2.0f * xn;
            }
            if (alive && ra < INFINITY && rb < INFINITY) {  // NaN / Inf rows are kept anyway (score NaN -> +inf below)
                atomicMax(&sh.max_ra, __float_as_uint(ra));
                if (!cosine) { atomicMax(&sh.max_rb, __float_as_uint(rb)); atomicMax(&sh.maxxx, __float_as_uint(x2)); }
            }
            if (eprof) t_a = sb_now();
            ptx::mbar_wait(&sh.accfull[buf], (it >> 1) & 1);
            ptx::tc_fence_after();
            if (eprof) {
                const unsigned long long t = sb_now();
                t_wait += t - t_a;
                if (it == 0) g_sb_prof[blockIdx.x][2] = t;
            }
#pragma unroll 1
            for (uint32_t chunk = 0; chunk < QT / 64; ++chunk) {
                const uint32_t cb = chunk * 64;
                uint32_t v0[32], v1[32];
                ptx::tmem_ld_32x32b_x32(tmem + lane_base + (buf * 2 + m) * QT + cb, v0);
                ptx::tmem_ld_32x32b_x32(tmem + lane_base + (buf * 2 + m) * QT + cb + 32, v1);
                ptx::tmem_wait_ld();
                if (chunk + 1 == QT / 64) {  // this M-tile's accumulator is in registers / scored: hand it back
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&sh.accempty[buf][m]);
                }
                if (DUMP && row < p.n) {  // tests only: score estimate and error bound per (row, query)
                    for (uint32_t c = 0; c < 64 && cb + c < nq; ++c) {
                        uint32_t bits = 0;
#pragma unroll
                        for (uint32_t i = 0; i < 32; ++i) {
                            if (c == i) bits = v0[i];
                            if (c == 32 + i) bits = v1[i];
                        }
                        const uint32_t qi = cb + c;
                        const float qn = sh.qn[qi];
                        float est = __uint_as_float(bits) * mul, b = fmaf(ra, sh.qa[qi], rb * sh.qb[qi]);
                        if (euclid) { est -= fmaf(qn, qn, x2); b += EUC_EPS * fmaf(qn, qn, x2); }
                        p.dump_est[(size_t)row * p.dump_ld + q0 + qi] = est;
                        p.dump_bnd[(size_t)row * p.dump_ld + q0 + qi] = b;
                    }
                }
                if (flood && it < 2 && cb < nq) {
                    // Flood tiles: until the first prune every threshold is -inf and EVERY (row, query) pair is buffered:
                    // place them directly (slot = tile-local row), no atomics, no register select.
                    const uint32_t slot = it * ROWS + rt;
                    const uint32_t grow = (uint32_t)(p.row_base + row);
#pragma unroll
                    for (uint32_t c = 0; c < 64; ++c) {
                        const uint32_t qi = cb + c;
                        if (qi < nq) {
                            float sc = fmaf(__uint_as_float(c < 32 ? v0[c & 31] : v1[c & 31]), mul, fmaf(ra, sh.qa[qi], rb * sh.qb[qi]));
                            if (euclid) { const float qn = sh.qn[qi]; sc -= EUC_KEEP * fmaf(qn, qn, x2); }
                            if (sc != sc) sc = INFINITY;
                            my_cand[(size_t)qi * PB + slot] = (alive && sc >= p.min_score) ? make_key(sc, grow) : 0ull;  // 0 = empty slot
                        }
                    }
                    if (rt == 0 && chunk == 0)
                        for (uint32_t qi = 0; qi < nq; ++qi) sh.cnt[qi] = (int)((it + 1) * ROWS);
                } else if (alive && cb < nq) {
                    // compact compare pass -> 64-bit mask of columns worth buffering (NaN passes); rare pushes out of line
                    uint32_t pass0 = 0, pass1 = 0;
#pragma unroll
                    for (uint32_t c = 0; c < 32; ++c) {
                        float s0 = fmaf(__uint_as_float(v0[c]), mul, fmaf(ra, sh.qa[cb + c], rb * sh.qb[cb + c]));
                        float s1 = fmaf(__uint_as_float(v1[c]), mul, fmaf(ra, sh.qa[cb + 32 + c], rb * sh.qb[cb + 32 + c]));
                        if (euclid) {
                            const float q0n = sh.qn[cb + c], q1n = sh.qn[cb + 32 + c];
                            s0 -= EUC_KEEP * fmaf(q0n, q0n, x2);
                            s1 -= EUC_KEEP * fmaf(q1n, q1n, x2);
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
                        // the register file is not indexable: a 6-level select tree (63 SEL) picks column c
                        uint32_t t[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) t[i] = (c & 32u) ? v1[i] : v0[i];
#pragma unroll
                        for (int w = 16; w >= 1; w >>= 1) {
#pragma unroll
                            for (int i = 0; i < w; ++i) t[i] = (c & (uint32_t)w) ? t[i + w] : t[i];
                        }
                        const uint32_t bits = t[0];
                        const uint32_t qi = cb + c;
                        float sc = fmaf(__uint_as_float(bits), mul, fmaf(ra, sh.qa[qi], rb * sh.qb[qi]));
                        if (euclid) { const float qn = sh.qn[qi]; sc -= EUC_KEEP * fmaf(qn, qn, x2); }
                        if (sc != sc) sc = INFINITY;  // undecidable here: keep it, the exact rescoring judges
                        if (sc >= sh.tau[qi]) {
                            int pos = atomicAdd(&sh.cnt[qi], 1);
                            if (pos < PB) my_cand[(size_t)qi * PB + pos] = make_key(sc, (uint32_t)(p.row_base + row));
                            else atomicExch(p.flags, 1);
                        }
                    }
                }
            }
            // prune any buffer that could overflow during the next tile (8 warps, different queries concurrently)
            group_sync(EPI_BAR, EPI_NT);  // every push of this tile is visible
            if (eprof) t_a = sb_now();
            // (never between the two flood tiles: the second one is placed by slot, behind the first one's 256 entries)
            for (uint32_t qi = ewarp; qi < nq; qi += EPI_WARPS)
                if (sh.cnt[qi] > prune_trigger && !(flood && it == 0)) {
                    if (eprof && g_sb_prof[blockIdx.x][6] == ~0ull) g_sb_prof[blockIdx.x][6] = it;
                    const float margin2 = bf16_margin2(p.metric, __uint_as_float(sh.max_ra), cosine ? 1.0f : __uint_as_float(sh.max_rb),
                                                       __uint_as_float(sh.maxxx), sh.qa[qi], sh.qb[qi], sh.qn[qi]);
                    float floor_tau = p.min_score;
                    const uint32_t gt = __ldcg(p.gtau + q0 + qi);
                    if (gt) floor_tau = fmaxf(floor_tau, ord_to_float(gt));
                    // the select's cost is the register-resident key count: an early-trigger prune holds ~260-300 keys, not PB
                    const int have = sh.cnt[qi];
                    if (have <= 320)
                        warp_prune<10>(my_cand + (size_t)qi * PB, &sh.cnt[qi], &sh.tau[qi], p.k, lane, nullptr, 0, true, margin2, prune_at, floor_tau,
                                       nullptr, p.flags + FLAG_OVERFLOW);
                    else if (have <= 512)
                        warp_prune<16>(my_cand + (size_t)qi * PB, &sh.cnt[qi], &sh.tau[qi], p.k, lane, nullptr, 0, true, margin2, prune_at, floor_tau,
                                       nullptr, p.flags + FLAG_OVERFLOW);
                    else
                        warp_prune<PRUNE_LANE>(my_cand + (size_t)qi * PB, &sh.cnt[qi], &sh.tau[qi], p.k, lane, nullptr, 0, true, margin2, prune_at, floor_tau,
                                               nullptr, p.flags + FLAG_OVERFLOW);
                    // everything inside the margin must fit below prune_at, or the next tile could overflow the buffer
                    if (lane == 0 && sh.cnt[qi] >= prune_at) atomicOr(p.flags + FLAG_OVERFLOW, 1);
                    if (lane == 0 && sh.tau[qi] > -INFINITY) atomicMax(p.gtau + q0 + qi, ord_bits(sh.tau[qi]));
                }
            group_sync(EPI_BAR, EPI_NT);
            if (eprof) t_prune += sb_now() - t_a;
            for (uint32_t qi = tid - EPI_WARP0 * 32; qi < nq; qi += EPI_NT) {  // adopt the shared thresholds
                const uint32_t gt = __ldcg(p.gtau + q0 + qi);
                if (gt) sh.tau[qi] = fmaxf(sh.tau[qi], ord_to_float(gt));
            }
        }
        if (eprof) { g_sb_prof[blockIdx.x][3] = sb_now(); g_sb_prof[blockIdx.x][5] = t_prune; g_sb_prof[blockIdx.x][7] = t_wait; }
    }

    // ---- emit: everything inside this CTA's margin AND above the shared threshold goes to the query's shared list.
    __syncthreads();
    {
        const bool cosine = p.metric == NK_METRIC_COSINE;
        uint64_t *my_cand = p.cand + (size_t)blockIdx.x * QT * PB;
        // Fast path (buffers with <= k_emit entries, i.e. almost always all of them): no selection, the finish kernel selects
        // globally — append what still reaches the current threshold.  The buffers live in global memory (L2); walking them
        // query by query was a chain of dependent round trips (~19 us at the end of every launch, measured).  Instead: the
        // thresholds of all queries in one round, then the live 32-entry chunks of ALL buffers as one flat work list, 16
        // independent loads in flight per warp, then the list-fill atomics of the whole batch in one more round.
        constexpr int NW = NTHREADS / 32, JQ = (QT + NW - 1) / NW;
        constexpr int STAGE_CHUNKS = Cfg<QT>::RING_BYTES / 256;  // 32-key chunks the (contiguous) operand rings hold
        if (prof && tid == 0) g_sb_prof[blockIdx.x][8] = sb_now();
        int n_emitted = 0;
        if (tid < QT) {
            // (the shared threshold is fetched asynchronously: its round trip — several microseconds while the other CTAs
            // still stream — overlaps the staging of the buffers below instead of preceding it)
            int chunks = 0;
            sh.ethr[tid] = 0u;
            if ((uint32_t)tid < nq && sh.cnt[tid] <= (int)p.k_emit) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(ptx::smem_u32(&sh.ethr[tid])), "l"(p.gtau + q0 + tid) : "memory");
                chunks = (sh.cnt[tid] + 31) >> 5;
            }
            sh.eoff[tid + 1] = chunks;
        }
        __syncthreads();
        if (warp == 0) {  // inclusive prefix of the chunk counts -> eoff[q + 1]; eoff[0] = 0
            constexpr int PERL = QT / 32;
            int loc[PERL], sum = 0;
#pragma unroll
            for (int i = 0; i < PERL; ++i) { loc[i] = sh.eoff[lane * PERL + i + 1]; sum += loc[i]; }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            int run = incl - sum;
            if (lane == 0) sh.eoff[0] = 0;
#pragma unroll
            for (int i = 0; i < PERL; ++i) { run += loc[i]; sh.eoff[lane * PERL + i + 1] = run; }
        }
        __syncthreads();
        {
            // The operand rings are idle now (every MMA has completed: the epilogue saw the last accumulator): stage the live
            // chunks there with asynchronous copies — every load of the warp in flight at once, no registers held.  Chunk
            // c of query q lands in stage slot eoff[q] + c; queries whose chunks do not fit the rings take the per-query path.
            if (prof && tid == 0) g_sb_prof[blockIdx.x][12] = sb_now();  // prefix done
            uint64_t *stage = reinterpret_cast<uint64_t *>(a_base);
            for (uint32_t qi = warp; qi < nq; qi += NW) {
                const int c0 = sh.eoff[qi], c1 = sh.eoff[qi + 1], n = sh.cnt[qi];
                if (c1 > STAGE_CHUNKS) continue;
                for (int ch = 0; ch < c1 - c0; ++ch) {
                    const int slot = (ch << 5) + lane;
                    uint64_t *dst = stage + ((size_t)(c0 + ch) << 5) + lane;
                    if (slot < n)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(ptx::smem_u32(dst)), "l"(my_cand + (size_t)qi * PB + slot) : "memory");
                    else
                        *dst = 0ull;
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            if (tid < QT) {  // the shared thresholds have landed too -> final score-word threshold per query
                float t = fmaxf(sh.tau[tid], p.min_score);
                const uint32_t gt = sh.ethr[tid];
                if (gt) t = fmaxf(t, ord_to_float(gt));
                sh.ethr[tid] = t > -INFINITY ? ord_bits(t) : 0u;
            }
            __syncthreads();
            if (prof && tid == 0) g_sb_prof[blockIdx.x][13] = sb_now();  // staged
            // survivors per query (pass A), ONE list-fill atomic per query — lane j owns the warp's j-th query, all of them
            // in one round trip — then the stores (pass B)
            int my = 0;
#pragma unroll
            for (int j = 0; j < JQ; ++j) {
                const uint32_t qi = warp + j * NW;
                if (qi >= nq) continue;
                const int c0 = sh.eoff[qi], c1 = sh.eoff[qi + 1];
                if (c1 > STAGE_CHUNKS) continue;
                const uint32_t th = sh.ethr[qi];
                int tot = 0;
                for (int c = c0; c < c1; ++c) {
                    const uint64_t v = stage[((size_t)c << 5) + lane];
                    tot += __popc(__ballot_sync(0xffffffffu, v != 0ull && (uint32_t)(v >> 32) >= th));
                }
                if (lane == j) my = tot;
            }
            int off = 0;
            if (my > 0) off = atomicAdd(p.gcount + q0 + warp + lane * NW, my);
            if (prof && tid == 0) g_sb_prof[blockIdx.x][14] = sb_now() + (off & 0);  // list-fill atomics answered
#pragma unroll
            for (int j = 0; j < JQ; ++j) {
                const uint32_t qi = warp + j * NW;
                int goff = __shfl_sync(0xffffffffu, off, j);
                if (qi >= nq) continue;
                const int c0 = sh.eoff[qi], c1 = sh.eoff[qi + 1];
                if (c1 > STAGE_CHUNKS) continue;
                const uint32_t th = sh.ethr[qi];
                uint64_t *out = p.partial + (size_t)(q0 + qi) * p.list_cap;
                for (int c = c0; c < c1; ++c) {
                    const uint64_t v = stage[((size_t)c << 5) + lane];
                    const bool keep = v != 0ull && (uint32_t)(v >> 32) >= th;
                    const uint32_t m = __ballot_sync(0xffffffffu, keep);
                    const int pos = goff + __popc(m & ((1u << lane) - 1u));
                    if (keep && pos < (int)p.list_cap) out[pos] = v;
                    goff += __popc(m);
                    n_emitted += __popc(m);
                }
            }
            if (prof && tid == 0) { g_sb_prof[blockIdx.x][9] = sb_now(); g_sb_prof[blockIdx.x][10] = (unsigned long long)sh.eoff[QT]; g_sb_prof[blockIdx.x][11] = (unsigned long long)n_emitted; }
        }
        for (uint32_t qi = warp; qi < nq; qi += NW) {  // buffers beyond k_emit (near-tie data): select, then emit
            if (sh.cnt[qi] <= (int)p.k_emit && sh.eoff[qi + 1] <= STAGE_CHUNKS) continue;
            const float margin2 = bf16_margin2(p.metric, __uint_as_float(sh.max_ra), cosine ? 1.0f : __uint_as_float(sh.max_rb),
                                               __uint_as_float(sh.maxxx), sh.qa[qi], sh.qb[qi], sh.qn[qi]);
            float floor_tau = p.min_score;
            const uint32_t gt = __ldcg(p.gtau + q0 + qi);
            if (gt) floor_tau = fmaxf(floor_tau, ord_to_float(gt));
            if (sh.cnt[qi] <= (int)p.k_emit) {  // (did not fit the staging area)
                const float t = fmaxf(sh.tau[qi], floor_tau);
                uint64_t thr = t > -INFINITY ? (uint64_t)ord_bits(t) << 32 : 1ull;
                if (thr == 0ull) thr = 1ull;
                warp_emit_above(my_cand + (size_t)qi * PB, sh.cnt[qi], thr, lane, p.partial + (size_t)(q0 + qi) * p.list_cap,
                                (int)p.list_cap, p.gcount + q0 + qi);
                continue;
            }
            warp_prune<PRUNE_LANE>(my_cand + (size_t)qi * PB, &sh.cnt[qi], &sh.tau[qi], p.k, lane,
                           p.partial + (size_t)(q0 + qi) * p.list_cap, (int)p.list_cap, true, margin2,
                           (int)p.k_emit, floor_tau, p.gcount + q0 + qi, p.flags + FLAG_OVERFLOW);
            if (lane == 0 && sh.cnt[qi] >= (int)p.k_emit && (int)p.k_emit > (int)p.k) atomicOr(p.flags + FLAG_OVERFLOW, 2);
            if (lane == 0 && sh.tau[qi] > -INFINITY) atomicMax(p.gtau + q0 + qi, ord_bits(sh.tau[qi]));
        }
        if (tid == 0) {
            atomicMax(reinterpret_cast<unsigned int *>(p.flags + 2), sh.maxxx);
            atomicMax(reinterpret_cast<unsigned int *>(p.flags + 4), sh.max_ra);
            atomicMax(reinterpret_cast<unsigned int *>(p.flags + 6), cosine ? __float_as_uint(1.0f) : sh.max_rb);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem, TMEM_COLS);
    if (prof && tid == 0) g_sb_prof[blockIdx.x][4] = sb_now();
}

static void sb_print_prof(cudaStream_t stream, uint32_t grid) {
    static unsigned long long h[256][16];
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(h, g_sb_prof, sizeof(h));
    const uint32_t g = grid < 256 ? grid : 256;
    unsigned long long t0 = ~0ull, tend_min = ~0ull, tend_max = 0, e3_max = 0;
    double setup = 0, first = 0, prune = 0, wait = 0, ends = 0, e3 = 0, ftile = 0;
    for (uint32_t b = 0; b < g; ++b) t0 = h[b][0] < t0 ? h[b][0] : t0;
    unsigned long long start_max = 0;
    for (uint32_t b = 0; b < g; ++b) {
        start_max = h[b][0] - t0 > start_max ? h[b][0] - t0 : start_max;
        setup += (double)(h[b][1] - h[b][0]); first += (double)(h[b][2] - h[b][0]); prune += (double)h[b][5]; wait += (double)h[b][7];
        ends += (double)(h[b][4] - t0); e3 += (double)(h[b][3] - t0); ftile += h[b][6] == ~0ull ? -1.0 : (double)h[b][6];
        tend_min = h[b][4] - t0 < tend_min ? h[b][4] - t0 : tend_min;
        tend_max = h[b][4] - t0 > tend_max ? h[b][4] - t0 : tend_max;
        e3_max = h[b][3] - t0 > e3_max ? h[b][3] - t0 : e3_max;
    }
    double sync_at = 0, emit_at = 0, chunks = 0, emitted = 0;
    for (uint32_t b = 0; b < g; ++b) { sync_at += (double)(h[b][8] - t0); emit_at += (double)(h[b][9] - t0); chunks += (double)h[b][10]; emitted += (double)h[b][11]; }
    double t12 = 0, t13 = 0, t14 = 0;
    for (uint32_t b = 0; b < g; ++b) { t12 += (double)(h[b][12] - t0); t13 += (double)(h[b][13] - t0); t14 += (double)(h[b][14] - t0); }
    fprintf(stderr, "[shadow prof] emission of warp 0: thresholds + prefix +%.1f | staged +%.1f | atomics answered +%.1f\n", t12 / g / 1e3, t13 / g / 1e3, t14 / g / 1e3);
    fprintf(stderr, "[shadow prof] all warps past the loop avg +%.1f | warp 0 emitted avg +%.1f | live chunks per CTA %.1f | entries emitted by warp 0 %.1f\n",
            sync_at / g / 1e3, emit_at / g / 1e3, chunks / g, emitted / g);
    fprintf(stderr, "[shadow prof, %u CTAs] span %.1f us | last CTA start +%.1f | set-up %.1f | first accumulator +%.1f | epilogue done avg +%.1f max +%.1f | "
            "exit min +%.1f avg +%.1f max +%.1f | prune section %.1f us (first at tile %.1f) | accumulator wait %.1f us\n",
            g, tend_max / 1e3, start_max / 1e3, setup / g / 1e3, first / g / 1e3, e3 / g / 1e3, e3_max / 1e3, tend_min / 1e3, ends / g / 1e3,
            tend_max / 1e3, prune / g / 1e3, ftile / g, wait / g / 1e3);
}

// The 16-bit pass needs a 16-bit image of the rows: the BF16 shadow of an fp32 shard, or — fp16 / bf16 corpora — the rows
// themselves (a.shadow == a.rows, shadow_native: no rounding residue on the row side, only the per-row |x|^2 array).
bool shadow_pass_supported(const DeviceInfo &di, const ScanArgs &a) {
    if (di.cc < 100 || a.shadow == nullptr || a.xnorm2 == nullptr || a.k > 192 || a.dim > 32768 || a.dim < 32 || a.n == 0) return false;
    if ((reinterpret_cast<uintptr_t>(a.rows) & 15) != 0 || (reinterpret_cast<uintptr_t>(a.shadow) & 15) != 0) return false;
    if (a.dtype == NK_DTYPE_F32) return a.dim % 4 == 0;
    return a.dim % 8 == 0;  // 16-bit rows: 16-byte TMA row stride and 128-bit exact re-scoring loads
}

// Host: stand-alone bf16 conversion of a query block (assign_tensor.cu: the centroids).
int bf16_prep_queries(const ScanArgs &a, uint32_t Qpad, uint32_t dimpad, float acc_c, void *qbf16, float *qnorm, float *qa,
                      float *qb, uint64_t *launches) {
    bf16_prep_queries_kernel<<<Qpad, 256, 0, a.stream>>>(a.queries, a.Q, a.dim, dimpad, a.metric == NK_METRIC_COSINE, acc_c,
                                                         static_cast<uint16_t *>(qbf16), qnorm, qa, qb);
    NK_CUDA_OK(cudaGetLastError());
    if (launches) ++*launches;
    return 0;
}

template <int QT, bool DUMP>
static int launch_shadow_pass_t(const DeviceInfo &di, const ScanArgs &a, Workspace &ws, const ShadowPassArgs &sp, uint64_t *launches) {
    using namespace sb;
    const bool native = a.shadow_native;  // 16-bit corpus scanned in place: row stride dim * 2, columns past dim read as zero (TMA OOB fill)
    const CUtensorMap *map_rows = tc_cached_map(ws, 4, a.shadow, a.n, native ? a.dim : sp.dimpad, 2, BKB, ROWS,
                                                (uint64_t)(native ? a.dim : sp.dimpad) * 2, a.dtype);
    const CUtensorMap *map_q = tc_cached_map(ws, QT == 64 ? 5 : 6, sp.qbf16, sp.Qpad, sp.dimpad, 2, BKB, QT, (uint64_t)sp.dimpad * 2,
                                             a.dtype);  // rows past Qpad: zero
    if (!map_rows || !map_q) return -1;
    const size_t smem = (size_t)Cfg<QT>::RING_BYTES + sizeof(Shared) + 1024;
    if (smem > di.max_smem_optin) {
        set_error("shadow tensor path needs %zu B shared memory (> %zu)", smem, di.max_smem_optin);
        return -1;
    }
    if (tc_ensure_smem(reinterpret_cast<const void *>(knn_scan_shadow_kernel<QT, DUMP>), di.device_id, smem)) return -1;
    tc::Params p{};
    p.n = a.n; p.dim = a.dim; p.nslab = sp.dimpad / BKB; p.row_base = a.row_base;
    p.q0 = sp.q0; p.nq = sp.nq; p.k = a.k; p.qpad_off = sp.q0; p.qgroups = sp.qgroups; p.list_cap = sp.grid * sp.k_emit;
    p.metric = a.metric; p.k_emit = sp.k_emit; p.margin_c = 0.0f; p.qnorm = sp.qnorm; p.qa = sp.qa; p.qb = sp.qb;
    p.xnorm2 = a.xnorm2; p.dnorm2 = a.dnorm2;
    p.cand = ws.cand; p.partial = ws.partial; p.flags = ws.flags; p.only_if = nullptr; p.debug = tc_debug_flags(); p.mask = a.row_mask;
    p.gtau = reinterpret_cast<uint32_t *>(ws.keys2); p.gcount = reinterpret_cast<int *>(ws.keys2) + (sp.Qpad + QT_BIG);
    p.presampled = sp.presampled; p.min_score = a.min_score; p.op_f16 = a.dtype == NK_DTYPE_F16;
    p.prune_trigger = tc_env_int("NK_PRUNE_TRIGGER", 0);
    p.dump_est = sp.dump_est; p.dump_bnd = sp.dump_bnd; p.dump_ld = sp.dump_ld;
    knn_scan_shadow_kernel<QT, DUMP><<<sp.grid, NTHREADS, smem, a.stream>>>(*map_rows, *map_q, p);
    NK_CUDA_OK(cudaGetLastError());
    if (launches) ++*launches;
    if (a.main_launches) ++*a.main_launches;
    if (p.debug & 64) sb_print_prof(a.stream, sp.grid);
    return 0;
}

int launch_shadow_pass(int qt, const DeviceInfo &di, const ScanArgs &a, Workspace &ws, const ShadowPassArgs &sp, uint64_t *launches) {
    if (sp.dump_est) return launch_shadow_pass_t<64, true>(di, a, ws, sp, launches);
    return qt == 128 ? launch_shadow_pass_t<128, false>(di, a, ws, sp, launches) : launch_shadow_pass_t<64, false>(di, a, ws, sp, launches);
}

}  // namespace nk
