// l2_stream.cu — microbenchmark: L2 -> SM rate of TMA loads when the working set is L2-resident.
//   mode 0: every CTA streams its own slice of a [n x 1024] fp32 buffer `reps` times (n sized to fit the 126 MB L2)
//   mode 1: every CTA streams the SAME 512 KB block (what the query operand of the scan does)
//   mode 2: groups of 4 adjacent CTAs stream the same tiles of a 8 GB buffer (HBM once, L2 three times: query groups)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_stream l2_stream.cu -lcuda && ./l2_stream
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma3(const CUtensorMap *m, uint64_t *bar, void *dst, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

constexpr int STAGES = 4;
// box = {32 floats, KA atoms, ROWS rows} = 32 KB per stage when KA*ROWS = 256.
__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap map, uint32_t n, int KA, int ROWS, float *sink, int reps, int mode) {
    extern __shared__ unsigned char dyn[];
    unsigned char *base = dyn + ((1024u - (smem_u32(dyn) & 1023u)) & 1023u);
    __shared__ uint64_t full[STAGES], empty[STAGES];
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t tiles = (n + ROWS - 1) / ROWS;
    const uint32_t first = mode == 1 ? 0 : (mode == 2 ? blockIdx.x / 4 : blockIdx.x), stride = mode == 1 ? 1 : (mode == 2 ? gridDim.x / 4 : gridDim.x);
    const int katoms = 32 / KA;  // K steps per tile (1024 floats = 32 atoms)
    if (threadIdx.x == 0) {
        uint32_t g = 0;
        for (int r = 0; r < reps; ++r)
        for (uint32_t t = first; t < tiles; t += stride)
            for (int j = 0; j < katoms; ++j, ++g) {
                uint32_t s = g % STAGES;
                mbar_wait(&empty[s], ((g / STAGES) & 1) ^ 1);
                mbar_expect(&full[s], 32768);
                tma3(&map, &full[s], base + s * 32768, 0, j * KA, t * ROWS);
            }
    } else if (threadIdx.x == 32) {
        uint32_t g = 0;
        float acc = 0.f;
        for (int r = 0; r < reps; ++r)
        for (uint32_t t = first; t < tiles; t += stride)
            for (int j = 0; j < katoms; ++j, ++g) {
                uint32_t s = g % STAGES;
                mbar_wait(&full[s], (g / STAGES) & 1);
                acc += *reinterpret_cast<float *>(base + s * 32768 + (g & 1023) * 4);
                mbar_arrive(&empty[s]);
            }
        if (acc == 12345.678f) *sink = acc;
    }
}

typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                        const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const uint64_t d = 1024;
    const uint64_t nmax = 2000000;
    float *corpus, *sink;
    CK(cudaMalloc(&corpus, nmax * d * 4));
    CK(cudaMemset(corpus, 0, nmax * d * 4));
    CK(cudaMalloc(&sink, 4));
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    Enc enc = (Enc)fp;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * 32768 + 1024));
    struct Case { int mode; uint64_t n; int reps; const char *what; } cases[] = {
        {0, 148 * 256 / 4, 400, "own slice, 9.7 MB working set"},
        {0, 148 * 256, 100, "own slice, 38.8 MB working set"},
        {0, 148 * 256 * 2, 50, "own slice, 77.6 MB working set"},
        {0, 148 * 256 * 3, 34, "own slice, 116 MB working set"},
        {1, 128, 3000, "all CTAs read the same 512 KB"},
        {1, 1024, 400, "all CTAs read the same 4 MB"},
        {2, nmax, 1, "groups of 4 CTAs share tiles, 8.2 GB from HBM"},
        {0, nmax, 1, "own slice, 8.2 GB from HBM (reference)"},
    };
    for (auto &c : cases) {
        const int KA = 1, ROWS = 256;
        CUtensorMap map;
        cuuint64_t gdim[3] = {32, d / 32, c.n};
        cuuint64_t gstr[2] = {128, d * 4};
        cuuint32_t box[3] = {32, (cuuint32_t)KA, (cuuint32_t)ROWS};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, corpus, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e9f;
        for (int it = 0; it < 4; ++it) {
            cudaEventRecord(e0);
            stream_kernel<<<sms, 128, STAGES * 32768 + 1024>>>(map, (uint32_t)c.n, KA, ROWS, sink, c.reps, c.mode);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (it > 0 && ms < best) best = ms;
        }
        // bytes delivered to the SMs
        double tiles = (double)((c.n + ROWS - 1) / ROWS);
        double per_cta_tiles = c.mode == 1 ? tiles : (c.mode == 2 ? tiles / (sms / 4) : tiles / sms);
        double bytes = per_cta_tiles * sms * ROWS * d * 4 * c.reps;
        printf("mode %d %-48s : %.3f ms  %.0f GB/s into the SMs\n", c.mode, c.what, best, bytes / best / 1e6);
    }
    return 0;
}
