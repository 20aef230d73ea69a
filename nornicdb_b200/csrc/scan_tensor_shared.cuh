// scan_tensor_shared.cuh — declarations shared by the tensor-core scan kernels (scan_tensor.cu: TF32 exact / filter over
// the fp32 rows, scan_tensor_shadow.cu: BF16 filter over the shadow corpus) and their common finish step.
#pragma once
#include <cuda.h>

#include "kernels.cuh"

namespace nk {

namespace tc {
constexpr int THREADS = 512;
constexpr int BK = 32;           // floats per corpus K-slab = one 128-byte swizzle row
constexpr int TMEM_COLS = 512;
constexpr int EPI_THREADS = 128;
constexpr int EPI_BAR = 1;
constexpr int P = 512;           // candidate buffer capacity per (CTA, query): warp_prune<16>
constexpr int QT_BIG = 256;      // padding of the per-query arrays (a CTA reads up to 128 entries past the last query)
constexpr float EUC_EPS = 4e-6f;  // fp32 rounding of |x|^2 + |q|^2 in the euclidean upper bound
constexpr float EUC_KEEP = 1.0f - EUC_EPS;

struct Params {
    uint32_t n, dim, nslab;
    uint64_t row_base;
    uint32_t q0, nq, k;       // queries [q0, q0+nq) in this launch; nq <= qgroups * QT
    uint32_t qgroups;         // 1, 2 or 4 query blocks (sibling CTAs b, b+1, ..) share each corpus tile (grid % qgroups == 0)
    uint32_t list_cap;        // filter: entries per query in `partial`
    int metric;
    uint32_t k_emit;          // slots per (CTA, query) in `partial` (k for exact; k + margin room for filter)
    float margin_c;           // filter mode: c in |s_hat - s| <= c |x| |q|
    const float *qnorm;       // filter mode: |q| per query (1 for cosine), padded
    const float *qa, *qb;     // shadow kernel: per-query bound factors |bf16(q)| and |q - bf16(q)| + acc_c |q|
    const float *xnorm2, *dnorm2;  // shadow kernel: per-row |x|^2 and |x - bf16(x)|^2
    uint64_t *cand;           // [grid][QT][P]
    uint32_t qpad_off;        // row of this launch's first query inside the padded query arrays
    uint64_t *partial;        // exact: [Q][grid][k_emit] fixed slots; filter: [Q][grid*k_emit] shared append lists
    uint32_t *gtau;           // filter: [Q] cross-CTA shared threshold (order-preserving bits, atomicMax; 0 = none yet)
    int *gcount;              // filter: [Q] fill of the shared append lists
    int *flags;               // [0] fatal buffer overflow, [1] filter-margin overflow (-> next stage), [2] max |x|^2 bits,
                              // [4] / [6] BF16 max ra / rb bits, [5] retry-stage marker, [7] longest list (diagnostics)
    const int *only_if;       // exact fallback: run only if *only_if != 0
    const uint32_t *mask;     // row bitmask (bit set = row takes part) or nullptr
    int debug;                // NK_TC_DEBUG bit 64: clock64 wait-time instrumentation of CTA 0
};
}  // namespace tc

#ifdef __CUDACC__
// Filter mode: 2 x (largest possible gap between an upper bound and the true score) for rows with |x|^2 <= maxxx.
//   cosine 2c | dot 2c|x||q| | euclidean (on -dist^2) 2(2c|x||q| + eps(|x|^2+|q|^2))
__device__ __forceinline__ float filter_margin2(int metric, float c, float maxxx, float qn) {
    if (metric == NK_METRIC_COSINE) return 2.0f * c;
    const float xq = sqrtf(maxxx) * qn;
    if (metric == NK_METRIC_DOT) return 2.0f * c * xq;
    return 2.0f * (2.0f * c * xq + tc::EUC_EPS * (maxxx + qn * qn));
}
// BF16 filter: bound(row, q) = ra(row) qa[q] + rb(row) qb[q] (+ eps (|x|^2 + |q|^2) for euclidean); margin2 is twice its
// largest value over the rows seen.
__device__ __forceinline__ float bf16_margin2(int metric, float max_ra, float max_rb, float maxxx, float qa, float qb, float qn) {
    float m = 2.0f * fmaf(max_ra, qa, max_rb * qb);
    if (metric == NK_METRIC_EUCLIDEAN) m += 2.0f * tc::EUC_EPS * (maxxx + qn * qn);
    return m;
}
#endif

// 2-D row-major [rows x cols] tensor of 4-byte (fp32) or 2-byte (bf16) elements, box = [box_rows x box_cols] with
// box_cols * elem_bytes == 128 (one swizzle row), 128-byte swizzle, OOB -> 0.
int tc_make_map(CUtensorMap *m, const void *base, uint64_t rows, uint32_t cols, uint32_t elem_bytes, uint32_t box_cols,
                uint32_t box_rows, uint64_t row_stride_bytes);
int tc_debug_flags();

// BF16 shadow filter pass (scan_tensor_shadow.cu): queries [q0, q0+nq), nq <= qgroups * qt, qt = 64 or 128.
int launch_shadow_pass(int qt, const DeviceInfo &di, const ScanArgs &a, Workspace &ws, uint32_t grid, uint32_t k_emit,
                       const void *qbf16, uint32_t dimpad, const float *qnorm, const float *qa, const float *qb, uint32_t Qpad,
                       uint32_t q0, uint32_t nq, uint32_t qgroups, uint64_t *launches);
bool shadow_pass_supported(const DeviceInfo &di, const ScanArgs &a);
int bf16_prep_queries(const ScanArgs &a, uint32_t Qpad, uint32_t dimpad, float acc_c, void *qbf16, float *qnorm, float *qa,
                      float *qb, uint64_t *launches);
// fp32 rows [first, first+count) -> bf16 shadow rows (stride dimpad) + |x|^2, |x - bf16(x)|^2
int build_shadow(const float *rows, uint64_t first, uint64_t count, uint32_t dim, uint32_t dimpad, void *shadow, float *xnorm2,
                 float *dnorm2, cudaStream_t stream);

}  // namespace nk
