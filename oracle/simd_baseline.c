/*
 * simd_baseline.c — the reference's amd64 CPU brute force, restated for TIMING (bench.py
 * cpu_baseline / --impl reference) and as a second, lane-parallel fp32 opinion in tests.
 *
 * TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE (same rule as knn_oracle.c).
 *
 * What the reference runs on amd64: for one query, loop over N contiguous rows calling an
 * AVX2+FMA kernel from github.com/viterin/vek v0.4.3 (vek32.Dot / CosineSimilarity / Distance;
 * pkg/simd/simd_amd64.go:20-44) in the Batch* loop shape of pkg/simd/simd.go:165-169,199-203,
 * 227-231, followed by the bounded insertion top-k of pkg/gpu/cuda/cuda_bridge.go:350-371.
 * vek is not vendored and Go is not installed, so the kernel is restated: 4 independent 8-lane
 * FMA accumulators over 32 floats per step (lane-parallel accumulation, which is what
 * "-ffast-math AVX2 assembly" gives — pkg/simd/doc.go:71-73), horizontal add at the end.
 * Compiled with -O3 -mavx2 -mfma -ffast-math (oracle/Makefile).
 *
 * Threading: the reference scores one query on ONE goroutine (pkg/search/vector_index.go:330-342).
 * threads=1 reproduces that; threads>1 partitions rows with OpenMP — an upper bound the reference
 * does not implement, reported with the core count.
 */
#include <immintrin.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define SB_COSINE 0
#define SB_DOT 1
#define SB_EUCLIDEAN 2

static inline float hsum8(__m256 v) {
    __m128 lo = _mm256_castps256_ps128(v), hi = _mm256_extractf128_ps(v, 1);
    lo = _mm_add_ps(lo, hi);
    lo = _mm_add_ps(lo, _mm_movehl_ps(lo, lo));
    lo = _mm_add_ss(lo, _mm_shuffle_ps(lo, lo, 1));
    return _mm_cvtss_f32(lo);
}

float sb_dot(const float *a, const float *b, size_t n) {
    __m256 s0 = _mm256_setzero_ps(), s1 = s0, s2 = s0, s3 = s0;
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        s0 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i), s0);
        s1 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i + 8), _mm256_loadu_ps(b + i + 8), s1);
        s2 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i + 16), _mm256_loadu_ps(b + i + 16), s2);
        s3 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i + 24), _mm256_loadu_ps(b + i + 24), s3);
    }
    for (; i + 8 <= n; i += 8) s0 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i), s0);
    float sum = hsum8(_mm256_add_ps(_mm256_add_ps(s0, s1), _mm256_add_ps(s2, s3)));
    for (; i < n; i++) sum += a[i] * b[i];
    return sum;
}

float sb_cosine(const float *a, const float *b, size_t n) {
    __m256 d0 = _mm256_setzero_ps(), d1 = d0, x0 = d0, x1 = d0, y0 = d0, y1 = d0;
    size_t i = 0;
    for (; i + 16 <= n; i += 16) {
        __m256 a0 = _mm256_loadu_ps(a + i), a1 = _mm256_loadu_ps(a + i + 8);
        __m256 b0 = _mm256_loadu_ps(b + i), b1 = _mm256_loadu_ps(b + i + 8);
        d0 = _mm256_fmadd_ps(a0, b0, d0);
        d1 = _mm256_fmadd_ps(a1, b1, d1);
        x0 = _mm256_fmadd_ps(a0, a0, x0);
        x1 = _mm256_fmadd_ps(a1, a1, x1);
        y0 = _mm256_fmadd_ps(b0, b0, y0);
        y1 = _mm256_fmadd_ps(b1, b1, y1);
    }
    float dot = hsum8(_mm256_add_ps(d0, d1)), xa = hsum8(_mm256_add_ps(x0, x1)), xb = hsum8(_mm256_add_ps(y0, y1));
    for (; i < n; i++) {
        dot += a[i] * b[i];
        xa += a[i] * a[i];
        xb += b[i] * b[i];
    }
    float den = sqrtf(xa * xb);
    /* simd_amd64.go:31-35 maps NaN (zero vector) to 0; written as a compare because -ffast-math
     * may fold isnan(). */
    if (!(den > 0.0f)) return 0.0f;
    return dot / den;
}

float sb_euclid(const float *a, const float *b, size_t n) {
    __m256 s0 = _mm256_setzero_ps(), s1 = s0, s2 = s0, s3 = s0;
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        __m256 t0 = _mm256_sub_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i));
        __m256 t1 = _mm256_sub_ps(_mm256_loadu_ps(a + i + 8), _mm256_loadu_ps(b + i + 8));
        __m256 t2 = _mm256_sub_ps(_mm256_loadu_ps(a + i + 16), _mm256_loadu_ps(b + i + 16));
        __m256 t3 = _mm256_sub_ps(_mm256_loadu_ps(a + i + 24), _mm256_loadu_ps(b + i + 24));
        s0 = _mm256_fmadd_ps(t0, t0, s0);
        s1 = _mm256_fmadd_ps(t1, t1, s1);
        s2 = _mm256_fmadd_ps(t2, t2, s2);
        s3 = _mm256_fmadd_ps(t3, t3, s3);
    }
    for (; i + 8 <= n; i += 8) {
        __m256 t0 = _mm256_sub_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i));
        s0 = _mm256_fmadd_ps(t0, t0, s0);
    }
    float sum = hsum8(_mm256_add_ps(_mm256_add_ps(s0, s1), _mm256_add_ps(s2, s3)));
    for (; i < n; i++) {
        float t = a[i] - b[i];
        sum += t * t;
    }
    return sqrtf(sum);
}

static inline float sb_score(const float *row, const float *q, size_t d, int metric) {
    if (metric == SB_COSINE) return sb_cosine(row, q, d);
    if (metric == SB_DOT) return sb_dot(row, q, d);
    return -sb_euclid(row, q, d); /* rank distance ascending */
}

typedef struct {
    float s;
    uint32_t i;
} sb_cand;

static inline int sb_better(float s, uint32_t i, const sb_cand *c) { return s > c->s || (s == c->s && i < c->i); }

static void sb_insert(sb_cand *top, uint32_t *cnt, uint32_t k, float s, uint32_t gi) {
    uint32_t pos;
    if (*cnt < k)
        pos = (*cnt)++;
    else if (sb_better(s, gi, &top[k - 1]))
        pos = k - 1;
    else
        return;
    while (pos > 0 && sb_better(s, gi, &top[pos - 1])) {
        top[pos] = top[pos - 1];
        pos--;
    }
    top[pos].s = s;
    top[pos].i = gi;
}

/* Brute-force kNN the way the reference's CPU path does it: per query, a serial scan over rows with
 * the SIMD kernel, bounded insertion top-k with (score desc, index asc).  threads<=1: one thread per
 * the whole job (queries processed one after another, as SearchBatch does — points_service.go:697-725).
 * threads>1: rows partitioned across OpenMP threads, per-thread top-k lists merged.
 * Euclidean: out_score is the distance.  Returns k clamped to n. */
unsigned sb_knn(const float *rows, uint64_t n, uint32_t dim, const float *queries, uint32_t Q, uint32_t k,
                int metric, int threads, uint32_t *out_idx, float *out_score) {
    if (k == 0 || n == 0 || Q == 0) return 0;
    uint32_t ke = k > n ? (uint32_t)n : k;
#ifdef _OPENMP
    int nt = threads > 1 ? threads : 1;
#else
    int nt = 1;
#endif
    sb_cand *lists = (sb_cand *)malloc(sizeof(sb_cand) * (size_t)nt * ke);
    uint32_t *cnts = (uint32_t *)malloc(sizeof(uint32_t) * nt);
    for (uint32_t qi = 0; qi < Q; qi++) {
        const float *q = queries + (uint64_t)qi * dim;
        memset(cnts, 0, sizeof(uint32_t) * nt);
#pragma omp parallel num_threads(nt)
        {
#ifdef _OPENMP
            int t = omp_get_thread_num();
#else
            int t = 0;
#endif
            uint64_t lo = n * (uint64_t)t / nt, hi = n * (uint64_t)(t + 1) / nt;
            sb_cand *top = lists + (size_t)t * ke;
            uint32_t cnt = 0;
            for (uint64_t r = lo; r < hi; r++) sb_insert(top, &cnt, ke, sb_score(rows + r * dim, q, dim, metric), (uint32_t)r);
            cnts[t] = cnt;
        }
        sb_cand *fin = lists; /* merge thread lists into thread 0's list */
        uint32_t fc = cnts[0];
        for (int t = 1; t < nt; t++)
            for (uint32_t j = 0; j < cnts[t]; j++) sb_insert(fin, &fc, ke, lists[(size_t)t * ke + j].s, lists[(size_t)t * ke + j].i);
        for (uint32_t j = 0; j < ke; j++) {
            out_idx[(uint64_t)qi * k + j] = fin[j].i;
            out_score[(uint64_t)qi * k + j] = metric == SB_EUCLIDEAN ? -fin[j].s : fin[j].s;
        }
    }
    free(lists);
    free(cnts);
    return ke;
}

int sb_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
