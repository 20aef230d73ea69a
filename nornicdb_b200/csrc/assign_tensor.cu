// assign_tensor.cu — k-means assignment step on the tensor cores (pkg/gpu/kmeans.go:458-546).
//
// nearest centroid of every corpus row = the Q x N^T contraction of the scan with Q = #centroids and an ARGMAX over the
// centroid columns instead of a per-query top-k.  The kernel reuses the shadow scan's pipeline (scan_tensor_shadow.cu:
// TMA slabs of the BF16 shadow and of the bf16 centroids, tcgen05.mma.kind::f16 from shared memory, double-buffered TMEM
// accumulators) with one more loop level — every corpus tile is multiplied with each block of 128 centroids in turn (the
// tile is re-read from L2, it crossed HBM once) — and a register-only epilogue: thread = corpus row keeps the four best
// UPPER bounds and the best LOWER bound of its row's scores across all centroid blocks.
//   cosine    (assignToCentroidsGPU, kmeans.go:491-546)  score = x.c/|c|            (|x| > 0 does not change the argmax)
//   euclidean (assignToCentroids,    kmeans.go:458-489)  score = 2 x.c - |c|^2      (argmin |x-c|^2)
// with the rigorous BF16 bound of scan_tensor_shadow.cu.  A row whose runner-up cannot reach the winner's lower bound is
// decided; the others (near-ties) are listed with their <= 4 candidates and assign_fixup_kernel re-scores those candidates
// exactly in fp32 (strict comparison in ascending centroid order = lowest index wins ties, kmeans.go:470-476,529-534).
#include <cuda.h>
#include <stdlib.h>

#include "kernels.cuh"
#include "ptx_sm100.cuh"
#include "scan_tensor_shared.cuh"

namespace nk {

namespace asg {
using namespace tc;
constexpr int NTHREADS = 384;
constexpr int ROWS = 256;
constexpr int QT = 128;
constexpr int BKB = 64;
constexpr int ASTAGES = 4;
constexpr int A_BYTES = ROWS * 128;
constexpr int B_BYTES = QT * 128;
constexpr int BSTAGES = 5;
constexpr int RING_BYTES = ASTAGES * A_BYTES + BSTAGES * B_BYTES;
constexpr int EPI_WARP0 = 4;
constexpr uint32_t NONE = 0xffffffffu, ALL = 0xfffffffeu;

struct Params {
    uint32_t n, nslab, K, nqb;   // rows, 64-element slabs per row, centroids, blocks of 128 centroids
    int metric;
    const float *qn, *qa, *qb;   // per centroid (padded): |c| (1 for cosine), bound factors
    const float *xnorm2, *dnorm2;
    uint32_t *assign;            // [n] out
    uint32_t *amb_count, *amb_rows, *amb_cand;  // near-ties: count, rows, 4 candidates each (ALL = every centroid)
    uint32_t amb_cap;
};

struct __align__(8) Shared {
    uint64_t afull[ASTAGES], aempty[ASTAGES];
    uint64_t bfull[BSTAGES], bempty[BSTAGES];
    uint64_t accfull[2], accempty[2][2];
    uint32_t tmem_base;
};
}  // namespace asg

__global__ void __launch_bounds__(asg::NTHREADS, 1)
assign_scan_kernel(const __grid_constant__ CUtensorMap map_rows, const __grid_constant__ CUtensorMap map_q, asg::Params p) {
    using namespace asg;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem_raw = smem_dyn + ((1024u - (ptx::smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char *a_base = smem_raw;
    unsigned char *b_base = smem_raw + (size_t)ASTAGES * A_BYTES;
    Shared &sh = *reinterpret_cast<Shared *>(smem_raw + (size_t)RING_BYTES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t num_tiles = (p.n + ROWS - 1) / ROWS, nslab = p.nslab, nqb = p.nqb;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_rows);
        ptx::prefetch_tensormap(&map_q);
        for (int i = 0; i < ASTAGES; ++i) { ptx::mbar_init(&sh.afull[i], 1); ptx::mbar_init(&sh.aempty[i], 1); }
        for (int i = 0; i < BSTAGES; ++i) { ptx::mbar_init(&sh.bfull[i], 1); ptx::mbar_init(&sh.bempty[i], 1); }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(&sh.accfull[b], 1);
            ptx::mbar_init(&sh.accempty[b][0], 4);
            ptx::mbar_init(&sh.accempty[b][1], 4);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) ptx::tmem_alloc(&sh.tmem_base, TMEM_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = sh.tmem_base;

    if (warp == 0) {
        // TMA producer: shadow slabs; a tile is streamed once per centroid block (first from HBM, then from L2)
        uint32_t g = 0;
        for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
            for (uint32_t qbk = 0; qbk < nqb; ++qbk)
                for (uint32_t j = 0; j < nslab; ++j, ++g) {
                    const uint32_t s = g % ASTAGES;
                    ptx::mbar_wait(&sh.aempty[s], ((g / ASTAGES) & 1) ^ 1);
                    if (ptx::elect_one_sync()) {
                        ptx::mbar_arrive_expect_tx(&sh.afull[s], A_BYTES);
                        ptx::tma_load_2d(&map_rows, &sh.afull[s], a_base + (size_t)s * A_BYTES, (int32_t)(j * BKB), (int32_t)(tile * ROWS),
                                         qbk + 1 == nqb ? ptx::CACHE_EVICT_FIRST : ptx::CACHE_EVICT_NORMAL);
                    }
                    __syncwarp();
                }
    } else if (warp == 3) {
        // TMA producer: centroid slabs (L2-resident)
        uint32_t g = 0;
        for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
            for (uint32_t qbk = 0; qbk < nqb; ++qbk)
                for (uint32_t j = 0; j < nslab; ++j, ++g) {
                    const uint32_t s = g % BSTAGES;
                    ptx::mbar_wait(&sh.bempty[s], ((g / BSTAGES) & 1) ^ 1);
                    if (ptx::elect_one_sync()) {
                        ptx::mbar_arrive_expect_tx(&sh.bfull[s], B_BYTES);
                        ptx::tma_load_2d(&map_q, &sh.bfull[s], b_base + (size_t)s * B_BYTES, (int32_t)(j * BKB), (int32_t)(qbk * QT), ptx::CACHE_EVICT_LAST);
                    }
                    __syncwarp();
                }
    } else if (warp == 1) {
        // MMA issuer
        const uint32_t idesc = ptx::make_idesc_bf16(128, QT);
        uint32_t g = 0, it = 0;
        for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
            for (uint32_t qbk = 0; qbk < nqb; ++qbk, ++it) {
                const uint32_t buf = it & 1;
                ptx::mbar_wait(&sh.accempty[buf][0], ((it >> 1) & 1) ^ 1);
                ptx::mbar_wait(&sh.accempty[buf][1], ((it >> 1) & 1) ^ 1);
                for (uint32_t j = 0; j < nslab; ++j, ++g) {
                    const uint32_t sa = g % ASTAGES, sbq = g % BSTAGES;
                    ptx::mbar_wait(&sh.bfull[sbq], (g / BSTAGES) & 1);
                    ptx::mbar_wait(&sh.afull[sa], (g / ASTAGES) & 1);
                    ptx::tc_fence_after();
                    if (ptx::elect_one_sync()) {
                        const uint64_t bdesc = ptx::make_smem_desc_sw128(ptx::smem_u32(b_base + (size_t)sbq * B_BYTES));
                        const uint32_t abase = ptx::smem_u32(a_base + (size_t)sa * A_BYTES);
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
        // epilogue: one warpgroup per 128-row M-tile, thread = corpus row; everything stays in registers
        const uint32_t m = (uint32_t)(warp - EPI_WARP0) >> 2, quad = warp & 3;
        const uint32_t lane_base = (quad * 32u) << 16;
        const uint32_t rt = m * 128 + quad * 32 + lane;
        const bool euclid = p.metric == NK_METRIC_EUCLIDEAN;
        uint32_t it = 0;
        for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const uint32_t row = tile * ROWS + rt;
            const float x2 = row < p.n ? __ldg(p.xnorm2 + row) : 0.0f;
            const float xn = sqrtf(x2);
            const float dxn = sqrtf(row < p.n ? __ldg(p.dnorm2 + row) : 0.0f) * 1.0001f;
            // score = mul * acc - (euclid ? |c|^2 : 0);  bound = ra qa[c] + rb qb[c] (+ 1e-6 |c|^2)
            const float mul = euclid ? 2.0f : 1.0f, ra = mul * dxn, rb = mul * xn;
            float u0 = -INFINITY, u1 = -INFINITY, u2 = -INFINITY, u3 = -INFINITY, maxlo = -INFINITY;
            uint32_t i0 = NONE, i1 = NONE, i2 = NONE, i3 = NONE;
            for (uint32_t qbk = 0; qbk < nqb; ++qbk, ++it) {
                const uint32_t buf = it & 1;
                ptx::mbar_wait(&sh.accfull[buf], (it >> 1) & 1);
                ptx::tc_fence_after();
#pragma unroll 1
                for (uint32_t chunk = 0; chunk < QT / 64; ++chunk) {
                    const uint32_t cb = qbk * QT + chunk * 64;
                    uint32_t v0[32], v1[32];
                    ptx::tmem_ld_32x32b_x32(tmem + lane_base + (buf * 2 + m) * QT + chunk * 64, v0);
                    ptx::tmem_ld_32x32b_x32(tmem + lane_base + (buf * 2 + m) * QT + chunk * 64 + 32, v1);
                    ptx::tmem_wait_ld();
                    if (chunk + 1 == QT / 64) {
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&sh.accempty[buf][m]);
                    }
                    if (cb < p.K) {
#pragma unroll
                        for (uint32_t c = 0; c < 64; ++c) {
                            const uint32_t ci = cb + c;  // centroid index (uniform over the warp)
                            if (ci < p.K) {
                                const float acc = __uint_as_float(c < 32 ? v0[c & 31] : v1[c & 31]);
                                const float qn = __ldg(p.qn + ci);
                                const float cc = euclid ? qn * qn : 0.0f;
                                const float s = fmaf(acc, mul, -cc);
                                const float bnd = fmaf(ra, __ldg(p.qa + ci), rb * __ldg(p.qb + ci)) + 1e-6f * cc;
                                const float up = s + bnd, lo = s - bnd;
                                maxlo = fmaxf(maxlo, lo);  // NaN scores never enter: such rows end up in the exact fix-up
                                if (up > u3) {  // strict: among equal bounds the lower centroid index stays ahead
                                    u3 = up; i3 = ci;
                                    if (u3 > u2) { float t = u2; u2 = u3; u3 = t; uint32_t ti = i2; i2 = i3; i3 = ti; }
                                    if (u2 > u1) { float t = u1; u1 = u2; u2 = t; uint32_t ti = i1; i1 = i2; i2 = ti; }
                                    if (u1 > u0) { float t = u0; u0 = u1; u1 = t; uint32_t ti = i0; i0 = i1; i1 = ti; }
                                }
                            }
                        }
                    }
                }
            }
            if (row < p.n) {
                const bool decided = i0 != NONE && !(u1 >= maxlo) && x2 > 0.0f && x2 < INFINITY;
                if (decided) {
                    p.assign[row] = i0;
                } else {
                    // near-tie (or a zero / non-finite row): exact fp32 re-scoring of the candidates decides
                    const uint32_t pos = atomicAdd(p.amb_count, 1u);
                    if (pos < p.amb_cap) {
                        p.amb_rows[pos] = row;
                        const bool all = i0 == NONE || u3 >= maxlo || !(x2 > 0.0f) || !(x2 < INFINITY);
                        p.amb_cand[4 * pos + 0] = all ? ALL : i0;
                        p.amb_cand[4 * pos + 1] = all ? NONE : i1;
                        p.amb_cand[4 * pos + 2] = (all || !(u2 >= maxlo)) ? NONE : i2;
                        p.amb_cand[4 * pos + 3] = NONE;
                    }
                    p.assign[row] = i0 == NONE ? 0u : i0;  // overwritten by the fix-up
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem, TMEM_COLS);
}

// Exact fp32 decision for the listed rows: one warp per row, candidates in ascending centroid order, strict comparison
// (kmeans.go:470-476: dist < minDist; :529-534: sim > maxSim), so the lowest index wins ties.
__global__ void assign_fixup_kernel(const float *rows, uint32_t dim, const float *centroids, uint32_t K, int metric,
                                    const uint32_t *amb_count, uint32_t amb_cap, const uint32_t *amb_rows, const uint32_t *amb_cand,
                                    uint32_t *assign) {
    const int lane = threadIdx.x & 31;
    uint32_t total = *amb_count;
    if (total > amb_cap) total = amb_cap;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < total; e += warps) {
        const uint32_t row = amb_rows[e];
        const float *x = rows + (size_t)row * dim;
        uint32_t cand[3] = {amb_cand[4 * e], amb_cand[4 * e + 1], amb_cand[4 * e + 2]};
        const bool all = cand[0] == asg::ALL;
        if (!all) {  // ascending order (3-element sorting network; NONE = 0xffffffff sorts last)
            if (cand[0] > cand[1]) { uint32_t t = cand[0]; cand[0] = cand[1]; cand[1] = t; }
            if (cand[1] > cand[2]) { uint32_t t = cand[1]; cand[1] = cand[2]; cand[2] = t; }
            if (cand[0] > cand[1]) { uint32_t t = cand[0]; cand[0] = cand[1]; cand[1] = t; }
        }
        float xx = 0.0f;
        if (metric != NK_METRIC_EUCLIDEAN) {
            for (uint32_t j = lane; j < dim; j += 32) xx = fmaf(x[j], x[j], xx);
#pragma unroll
            for (int o = 16; o; o >>= 1) xx += __shfl_xor_sync(0xffffffffu, xx, o);
        }
        float best = metric == NK_METRIC_EUCLIDEAN ? INFINITY : -INFINITY;
        uint32_t nearest = 0;
        const uint32_t count = all ? K : 3u;
        for (uint32_t t = 0; t < count; ++t) {
            const uint32_t ci = all ? t : cand[t];
            if (ci >= K) continue;
            const float *c = centroids + (size_t)ci * dim;
            float a = 0.0f, cc = 0.0f;
            if (metric == NK_METRIC_EUCLIDEAN) {
                for (uint32_t j = lane; j < dim; j += 32) { const float d = x[j] - c[j]; a = fmaf(d, d, a); }
            } else {
                for (uint32_t j = lane; j < dim; j += 32) { a = fmaf(x[j], c[j], a); cc = fmaf(c[j], c[j], cc); }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                cc += __shfl_xor_sync(0xffffffffu, cc, o);
            }
            if (metric == NK_METRIC_EUCLIDEAN) {
                if (a < best) { best = a; nearest = ci; }
            } else {
                const float den = sqrtf(xx * cc);
                const float sim = den > 0.0f ? a / den : 0.0f;  // cosineSimilarityFlat: 0 for zero vectors (gpu.go:2467-2484)
                if (sim > best) { best = sim; nearest = ci; }
            }
        }
        if (lane == 0) assign[row] = nearest;
    }
}

bool assign_tensor_supported(const DeviceInfo &di, uint32_t dim, uint32_t K, int metric) {
    if (const char *e = getenv("NK_ASSIGN_TENSOR"))
        if (atoi(e) == 0) return false;
    return di.cc >= 100 && dim % 4 == 0 && dim >= 32 && dim <= 32768 && K >= 1 && (metric == NK_METRIC_COSINE || metric == NK_METRIC_EUCLIDEAN);
}

// rows / shadow / norms: one shard (device pointers); centroids_dev: [K x dim] fp32 on the same device; assign_dev: [n].
int assign_tensor(const DeviceInfo &di, const float *rows, const void *shadow, uint32_t dimpad, const float *xnorm2, const float *dnorm2,
                  uint32_t n, uint32_t dim, const float *centroids_dev, uint32_t K, int metric, uint32_t *assign_dev, cudaStream_t stream,
                  uint64_t *launches, void **scratch, size_t *scratch_bytes) {
    using namespace asg;
    if (n == 0) return 0;
    const uint32_t nqb = (K + QT - 1) / QT, Kpad = nqb * QT;
    const float acc_c = (float)dim * 2.384185791015625e-7f + 4e-6f;
    // scratch: bf16 centroids [Kpad x dimpad], qn / qa / qb [Kpad], near-tie list (every row could be one)
    const size_t off_q = 0, off_f = ((size_t)Kpad * dimpad * 2 + 255) & ~(size_t)255, off_amb = off_f + (((size_t)3 * Kpad * 4 + 255) & ~(size_t)255);
    const size_t total = off_amb + 4 + (size_t)n * 4 + (size_t)n * 16 + 64;
    if (ws_reserve(scratch, scratch_bytes, total)) return -1;  // grow-only, reused across calls
    unsigned char *buf = static_cast<unsigned char *>(*scratch);
    float *qn = reinterpret_cast<float *>(buf + off_f), *qa = qn + Kpad, *qb = qa + Kpad;
    uint32_t *amb_count = reinterpret_cast<uint32_t *>(buf + off_amb);
    uint32_t *amb_rows = amb_count + 1;
    uint32_t *amb_cand = amb_rows + n;
    int rc = 0;
    do {
        if (cudaMemsetAsync(amb_count, 0, 4, stream) != cudaSuccess) { rc = -1; break; }
        ScanArgs a;
        a.queries = centroids_dev; a.Q = K; a.dim = dim; a.metric = metric; a.stream = stream;
        if (bf16_prep_queries(a, Kpad, dimpad, acc_c, buf + off_q, qn, qa, qb, launches)) { rc = -1; break; }
        CUtensorMap map_rows, map_q;
        if (tc_make_map(&map_rows, shadow, n, dimpad, 2, BKB, ROWS, (uint64_t)dimpad * 2)) { rc = -1; break; }
        if (tc_make_map(&map_q, buf + off_q, Kpad, dimpad, 2, BKB, QT, (uint64_t)dimpad * 2)) { rc = -1; break; }
        const size_t smem = (size_t)RING_BYTES + sizeof(Shared) + 1024;
        if (cudaFuncSetAttribute(assign_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { rc = -1; break; }
        asg::Params p;
        p.n = n; p.nslab = dimpad / BKB; p.K = K; p.nqb = nqb; p.metric = metric;
        p.qn = qn; p.qa = qa; p.qb = qb; p.xnorm2 = xnorm2; p.dnorm2 = dnorm2;
        p.assign = assign_dev; p.amb_count = amb_count; p.amb_rows = amb_rows; p.amb_cand = amb_cand; p.amb_cap = n;
        const uint32_t num_tiles = (n + ROWS - 1) / ROWS;
        const uint32_t grid = num_tiles < (uint32_t)di.num_sms ? num_tiles : (uint32_t)di.num_sms;
        assign_scan_kernel<<<grid, NTHREADS, smem, stream>>>(map_rows, map_q, p);
        assign_fixup_kernel<<<di.num_sms * 4, 256, 0, stream>>>(rows, dim, centroids_dev, K, metric, amb_count, n, amb_rows, amb_cand, assign_dev);
        if (cudaGetLastError() != cudaSuccess) { rc = -1; break; }
        if (launches) *launches += 2;
    } while (0);
    if (rc != 0) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) set_error("assign_tensor: %s", cudaGetErrorString(e));
    }
    cudaStreamSynchronize(stream);
    return rc;
}

}  // namespace nk
