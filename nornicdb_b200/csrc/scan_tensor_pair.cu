// scan_tensor_pair.cu — the 16-bit filter scan for LARGE batches on CTA PAIRS (tcgen05.mma.cta_group::2).
//
// At 128 query columns the single-CTA kernel (scan_tensor_shadow.cu) is bound by shared-memory bandwidth, not by the
// tensor pipe: every [128 rows x 128 queries x 16] MMA reads 4 KB of A and 4 KB of B from shared memory in 64 cycles
// (= the SM's whole 128 B/cycle) while TMA writes another 96 B/cycle of operands — ncu: tensor pipe 50 % active
// (profiles/ncu_r2_scan_shadow_c3_k10.md).  A CTA pair (two SMs of one TPC, a 2-CTA cluster) shares the QUERY operand:
//   pair tile  = 256 corpus rows x 256 queries: CTA r stages rows [128r, 128r+128) of the tile (A) and queries
//   [128r, 128r+128) of the block (its half of B); one tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16, 128 cycles)
//   reads per CTA 4 KB of A + 4 KB of B-half = 64 B/cycle, TMA writes 32 KB per CTA per 512 cycles = 64 B/cycle:
//   exactly the shared-memory budget instead of 1.75x over it.
// Each CTA's TMEM receives the accumulators of ITS 128 rows against all 256 queries (256 columns, double-buffered = all 512
// columns), so the epilogue is the same thread-per-row threshold / push / warp-prune machinery as the single-CTA kernel —
// 8 warps: 4 lane quadrants x 2 column halves.
//
// Roles per CTA (384 threads): warp 0 TMA producer (own A half + own B half, completion counted on the LEADER's full
// barrier: cp.async.bulk.tensor ... .cta_group::2), warp 1 MMA issuer (leader CTA only), warp 2 TMEM allocation, warps 4-11
// epilogue.  Barriers: full[s] (leader; 64 KB of transactions from both CTAs), empty[s] and accfull[b] (each CTA; arrived
// by tcgen05.commit ... multicast::cluster to both CTAs), accempty[b] (leader; 8 epilogue warps of EACH CTA arrive, the
// peer's through the cluster shared window).
//
// Optional (NK_PAIR_CLUSTER = 2 | 4): a cluster of C pairs on the SAME query block and C consecutive row tiles.  The two
// query half-slabs are then loaded ONCE per cluster and TMA-multicast to the C even / C odd CTAs (-25 % / -37 % of the L2->SM
// operand traffic that bounds this kernel); a ring stage is released when all C pairs have consumed it.
//
// Algorithmic HBM bytes per launch: n * dimpad * 2 + 8 n (one launch serves up to 2 groups x 256 queries; sibling pairs of
// the two groups stream the same tiles at the same pace and share them through L2).
#include <cuda.h>

#include "kernels.cuh"
#include "ptx_sm100.cuh"
#include "scan_tensor_shared.cuh"

namespace nk {

namespace pr {
using namespace tc;
constexpr int NTHREADS = 384;
constexpr int ROWS_CTA = 128;             // corpus rows per CTA per tile (pair tile = 256 rows)
constexpr int QT = 256;                   // query columns per pair MMA
constexpr int QH = QT / 2;                // staged per CTA
constexpr int BKB = 64;                   // bf16 per row per slab = one 128-byte swizzle row
constexpr int STAGES = 6;
constexpr int A_BYTES = ROWS_CTA * 128;   // 16 KB
constexpr int B_BYTES = QH * 128;         // 16 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int RING_BYTES = STAGES * STAGE_BYTES;  // 192 KB
constexpr int EPI_WARP0 = 4, EPI_WARPS = 8, EPI_NT = EPI_WARPS * 32;
constexpr int PB = tc::P_SHADOW;
constexpr int PRUNE_LANE = PB / 32;
constexpr int FLOOD_TILES = 4;            // first tiles of a CTA: every (row, query) pair is placed directly (4 x 128 slots)
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> the even (leader) CTA

struct __align__(8) Shared {
    uint64_t full[STAGES], empty[STAGES];
    uint64_t accfull[2], accempty[2];
    uint32_t tmem_base;
    unsigned int maxxx, max_ra, max_rb;
    float tau[QT], qn[QT], qa[QT], qb[QT];
    int cnt[QT];
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are counted on the LEADER CTA's mbarrier (same offset, rank bit cleared)
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap *m, uint64_t *bar_local, void *smem_dst, int32_t c0, int32_t c1,
                                                 uint64_t cache_hint) {
    const uint32_t bar = ptx::smem_u32(bar_local) & PEER_MASK;
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "l"(cache_hint)
        : "memory");
}
// the same, multicast to every CTA of `cta_mask` (same smem offset in each; completion counted on each destination pair's leader)
__device__ __forceinline__ void tma_load_2d_pair_mc(const CUtensorMap *m, uint64_t *bar_local, void *smem_dst, int32_t c0, int32_t c1,
                                                    uint16_t cta_mask, uint64_t cache_hint) {
    const uint32_t bar = ptx::smem_u32(bar_local) & PEER_MASK;
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5, %6;"
        ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask), "l"(cache_hint)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void mma_bf16_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the barrier at the same offset in BOTH CTAs of the pair once all MMAs issued so far have completed
__device__ __forceinline__ void tc_commit_pair(uint64_t *bar_local, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(ptx::smem_u32(bar_local)), "h"(cta_mask) : "memory");
}
// arrive on the LEADER's copy of a barrier from either CTA
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar_local, uint32_t leader_rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(ptx::smem_u32(bar_local)), "r"(leader_rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *smem_result, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(smem_result)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
}  // namespace pr

__global__ void __launch_bounds__(pr::NTHREADS, 1)
knn_scan_pair_kernel(const __grid_constant__ CUtensorMap map_rows, const __grid_constant__ CUtensorMap map_q, tc::Params p) {
    using namespace pr;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem_raw = smem_dyn + ((1024u - (ptx::smem_u32(smem_dyn) & 1023u)) & 1023u);
    Shared &sh = *reinterpret_cast<Shared *>(smem_raw + (size_t)RING_BYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();  // cluster = C pairs (C = 1, 2 or 4)
    const uint32_t rank = crank & 1u, C = csize >> 1, cpair = crank >> 1;
    const bool leader = rank == 0;
    const uint16_t pair_mask = (uint16_t)(3u << (crank & ~1u)), all_mask = (uint16_t)((1u << csize) - 1u);
    // query groups: cluster c serves query block (c % G) over the cluster-tile subset (c / G); siblings share tiles through L2.
    // A cluster-tile is C consecutive 256-row tiles, one per pair.
    const uint32_t num_tiles = (p.n + 2 * ROWS_CTA * C - 1) / (2 * ROWS_CTA * C);
    const uint32_t cl = blockIdx.x / csize, ncl = gridDim.x / csize;
    const uint32_t grp = cl % p.qgroups, sub = cl / p.qgroups, sgrid = ncl / p.qgroups;
    const uint32_t q0 = p.q0 + grp * QT, qpad_off = p.qpad_off + grp * QT;
    const uint32_t nq = p.nq - grp * QT < (uint32_t)QT ? p.nq - grp * QT : (uint32_t)QT;
    const uint64_t a_policy = p.qgroups > 1 ? ptx::CACHE_EVICT_NORMAL : ptx::CACHE_EVICT_FIRST;
    const uint32_t nslab = p.nslab;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_rows);
        ptx::prefetch_tensormap(&map_q);
        for (int i = 0; i < STAGES; ++i) { ptx::mbar_init(&sh.full[i], 1); ptx::mbar_init(&sh.empty[i], C); }  // every pair releases a stage
        for (int b = 0; b < 2; ++b) { ptx::mbar_init(&sh.accfull[b], 1); ptx::mbar_init(&sh.accempty[b], 2 * EPI_WARPS); }
        sh.maxxx = 0u; sh.max_ra = 0u; sh.max_rb = 0u;
        ptx::fence_barrier_init();
    }
    if (tid < QT) {
        sh.tau[tid] = p.min_score;
        sh.cnt[tid] = 0;
        sh.qn[tid] = p.qnorm[qpad_off + tid];
        sh.qa[tid] = p.qa[qpad_off + tid];
        sh.qb[tid] = p.qb[qpad_off + tid];
    }
    __syncthreads();
    cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit can land
    if (warp == 2) tmem_alloc_pair(&sh.tmem_base, TMEM_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = sh.tmem_base;

    if (warp == 0) {
        // ===================================== TMA producer: own A half + own B half ===================
        uint32_t g = 0, epoch = 0;
        const bool lockstep = p.sync_every != 0 && p.qgroups > 1;
        for (uint32_t tile = sub; tile < num_tiles; tile += sgrid) {
            for (uint32_t j = 0; j < nslab; ++j, ++g) {
                const uint32_t s = g % STAGES;
                ptx::mbar_wait(&sh.empty[s], ((g / STAGES) & 1) ^ 1);
                if (lockstep && (j % p.sync_every) == 0) {
                    // sibling lockstep (experiment, NK_PAIR_SYNC): the producers of the G clusters that stream this tile
                    // subset issue the slab together, so that their requests for the same corpus lines meet in L2
                    if (lane == 0) {
                        const uint32_t target = (epoch + 1) * csize * p.qgroups;
                        atomicAdd(p.sync + sub, 1u);
                        uint32_t v;
                        do {
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.sync + sub) : "memory");
                        } while ((int32_t)(v - target) < 0);
                    }
                    ++epoch;
                    __syncwarp();
                }
                if (ptx::elect_one_sync()) {
                    unsigned char *st = smem_raw + (size_t)s * STAGE_BYTES;
                    if (leader) ptx::mbar_arrive_expect_tx(&sh.full[s], 2 * STAGE_BYTES);  // both CTAs' bytes land on this barrier
                    tma_load_2d_pair(&map_rows, &sh.full[s], st, (int32_t)(j * BKB),
                                     (int32_t)((tile * C + cpair) * 2 * ROWS_CTA + rank * ROWS_CTA), a_policy);
                    if (C == 1) {
                        tma_load_2d_pair(&map_q, &sh.full[s], st + A_BYTES, (int32_t)(j * BKB), (int32_t)(qpad_off + rank * QH), ptx::CACHE_EVICT_LAST);
                    } else if (cpair == 0) {
                        // one L2 read per query half-slab per cluster: multicast to the C even (rank 0) / odd (rank 1) CTAs
                        uint16_t mask = 0;
                        for (uint32_t c = 0; c < C; ++c) mask |= (uint16_t)(1u << (2 * c + rank));
                        tma_load_2d_pair_mc(&map_q, &sh.full[s], st + A_BYTES, (int32_t)(j * BKB), (int32_t)(qpad_off + rank * QH), mask,
                                            ptx::CACHE_EVICT_LAST);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1 && leader) {
        // ===================================== MMA issuer (leader CTA) =================================
        const uint32_t idesc = p.op_f16 ? ptx::make_idesc_f16(2 * ROWS_CTA, QT) : ptx::make_idesc_bf16(2 * ROWS_CTA, QT);
        uint32_t g = 0, it = 0;
        for (uint32_t tile = sub; tile < num_tiles; tile += sgrid, ++it) {
            const uint32_t buf = it & 1;
            ptx::mbar_wait(&sh.accempty[buf], ((it >> 1) & 1) ^ 1);  // both CTAs' epilogues have drained this buffer
            for (uint32_t j = 0; j < nslab; ++j, ++g) {
                const uint32_t s = g % STAGES;
                ptx::mbar_wait(&sh.full[s], (g / STAGES) & 1);
                ptx::tc_fence_after();
                if (ptx::elect_one_sync()) {
                    const uint32_t base = ptx::smem_u32(smem_raw + (size_t)s * STAGE_BYTES);
                    const uint64_t adesc = ptx::make_smem_desc_sw128(base), bdesc = ptx::make_smem_desc_sw128(base + A_BYTES);
                    const uint32_t d = tmem + buf * QT;
#pragma unroll
                    for (uint32_t kk = 0; kk < 4; ++kk) mma_bf16_ss_pair(d, adesc + kk * 2, bdesc + kk * 2, idesc, (j | kk) != 0);
                    tc_commit_pair(&sh.empty[s], all_mask);            // this pair is done with stage s (every CTA of the cluster counts C)
                    if (j + 1 == nslab) tc_commit_pair(&sh.accfull[buf], pair_mask);
                }
                __syncwarp();
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ===================================== epilogue: 4 lane quadrants x 2 column halves ===========
        const uint32_t ewarp = warp - EPI_WARP0, quad = warp & 3, chalf = ewarp >> 2;
        const uint32_t lane_base = (quad * 32u) << 16;
        const uint32_t rt = quad * 32 + lane;  // row within this CTA's half of the tile
        uint64_t *my_cand = p.cand + (size_t)blockIdx.x * QT * PB;
        const int prune_at = PB - ROWS_CTA;
        const int prune_trigger = min(prune_at, max(FLOOD_TILES * ROWS_CTA - 1, 4 * (int)p.k));
        const bool cosine = p.metric == NK_METRIC_COSINE, euclid = p.metric == NK_METRIC_EUCLIDEAN;
        uint32_t it = 0;
        for (uint32_t tile = sub; tile < num_tiles; tile += sgrid, ++it) {
            const uint32_t buf = it & 1;
            const uint32_t row = (tile * C + cpair) * 2 * ROWS_CTA + rank * ROWS_CTA + rt;
            const bool alive = row < p.n && (!p.mask || ((__ldg(p.mask + (row >> 5)) >> (row & 31)) & 1u));
            const float x2 = row < p.n ? __ldg(p.xnorm2 + row) : 0.0f;
            const float xn = sqrtf(x2);
            const float dxn = sqrtf((row < p.n && p.dnorm2) ? __ldg(p.dnorm2 + row) : 0.0f) * 1.0001f;
            float mul = 1.0f, ra = dxn, rb = xn;
            if (cosine) {
                mul = x2 > 0.0f ? 1.0f / xn : 0.0f;
                ra = dxn * mul * 1.000001f;
                rb = 1.0f;
            } else if (euclid) {
                mul = 2.0f; ra = 2.0f * dxn; rb = 2.0f * xn;
            }
            if (alive && ra < INFINITY && rb < INFINITY) {
                atomicMax(&sh.max_ra, __float_as_uint(ra));
                if (!cosine) { atomicMax(&sh.max_rb, __float_as_uint(rb)); atomicMax(&sh.maxxx, __float_as_uint(x2)); }
            }
            ptx::mbar_wait(&sh.accfull[buf], (it >> 1) & 1);
            ptx::tc_fence_after();
#pragma unroll 1
            for (uint32_t chunk = 0; chunk < 2; ++chunk) {
                const uint32_t cb = chalf * QH + chunk * 64;  // first query column of this chunk
                uint32_t v0[32], v1[32];
                ptx::tmem_ld_32x32b_x32(tmem + lane_base + buf * QT + cb, v0);
                ptx::tmem_ld_32x32b_x32(tmem + lane_base + buf * QT + cb + 32, v1);
                ptx::tmem_wait_ld();
                if (chunk == 1) {  // this warp's share of the accumulator is in registers: hand it back to the leader's issuer
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(&sh.accempty[buf], crank & ~1u);
                }
                if (it < (uint32_t)FLOOD_TILES && cb < nq) {
                    const uint32_t slot = it * ROWS_CTA + rt;
                    const uint32_t grow = (uint32_t)(p.row_base + row);
#pragma unroll
                    for (uint32_t c = 0; c < 64; ++c) {
                        const uint32_t qi = cb + c;
                        if (qi < nq) {
                            float sc = fmaf(__uint_as_float(c < 32 ? v0[c & 31] : v1[c & 31]), mul, fmaf(ra, sh.qa[qi], rb * sh.qb[qi]));
                            if (euclid) { const float qn = sh.qn[qi]; sc -= EUC_KEEP * fmaf(qn, qn, x2); }
                            if (sc != sc) sc = INFINITY;
                            my_cand[(size_t)qi * PB + slot] = (alive && sc >= p.min_score) ? make_key(sc, grow) : 0ull;
                        }
                    }
                } else if (alive && cb < nq) {
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
                        uint32_t t[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) t[i] = (c & 32u) ? v1[i] : v0[i];
#pragma unroll
                        for (int w = 16; w >= 1; w >>= 1) {
#pragma unroll
                            for (int i = 0; i < w; ++i) t[i] = (c & (uint32_t)w) ? t[i + w] : t[i];
                        }
                        const uint32_t qi = cb + c;
                        float sc = fmaf(__uint_as_float(t[0]), mul, fmaf(ra, sh.qa[qi], rb * sh.qb[qi]));
                        if (euclid) { const float qn = sh.qn[qi]; sc -= EUC_KEEP * fmaf(qn, qn, x2); }
                        if (sc != sc) sc = INFINITY;
                        if (sc >= sh.tau[qi]) {
                            int pos = atomicAdd(&sh.cnt[qi], 1);
                            if (pos < PB) my_cand[(size_t)qi * PB + pos] = make_key(sc, (uint32_t)(p.row_base + row));
                            else atomicExch(p.flags, 1);
                        }
                    }
                }
            }
            group_sync(EPI_BAR, EPI_NT);  // every push of this tile is visible
            if (it < (uint32_t)FLOOD_TILES)
                for (uint32_t qi = tid - EPI_WARP0 * 32; qi < nq; qi += EPI_NT) sh.cnt[qi] = (int)((it + 1) * ROWS_CTA);
            group_sync(EPI_BAR, EPI_NT);
            for (uint32_t qi = ewarp; qi < nq; qi += EPI_WARPS)
                if (sh.cnt[qi] > prune_trigger) {
                    const float margin2 = bf16_margin2(p.metric, __uint_as_float(sh.max_ra), cosine ? 1.0f : __uint_as_float(sh.max_rb),
                                                       __uint_as_float(sh.maxxx), sh.qa[qi], sh.qb[qi], sh.qn[qi]);
                    float floor_tau = p.min_score;
                    const uint32_t gt = __ldcg(p.gtau + q0 + qi);
                    if (gt) floor_tau = fmaxf(floor_tau, ord_to_float(gt));
                    // the select's cost is the register-resident key count: size it to what the buffer holds (512 right after
                    // the flood tiles, 512-700 at a later trigger), not to its 1024 slots
                    const int have = sh.cnt[qi];
                    if (have <= 512 && p.prune_trigger != -1)  // (NK_PRUNE_TRIGGER=-1: A/B switch, always the full-width select)
                        warp_prune<16>(my_cand + (size_t)qi * PB, &sh.cnt[qi], &sh.tau[qi], p.k, lane, nullptr, 0, true, margin2, prune_at,
                                       floor_tau, nullptr, p.flags + FLAG_OVERFLOW);
                    else if (have <= 704 && p.prune_trigger != -1)
                        warp_prune<22>(my_cand + (size_t)qi * PB, &sh.cnt[qi], &sh.tau[qi], p.k, lane, nullptr, 0, true, margin2, prune_at,
                                       floor_tau, nullptr, p.flags + FLAG_OVERFLOW);
                    else
                        warp_prune<PRUNE_LANE>(my_cand + (size_t)qi * PB, &sh.cnt[qi], &sh.tau[qi], p.k, lane, nullptr, 0, true, margin2, prune_at,
                                               floor_tau, nullptr, p.flags + FLAG_OVERFLOW);
                    if (lane == 0 && sh.cnt[qi] >= prune_at) atomicOr(p.flags + FLAG_OVERFLOW, 1);
                    if (lane == 0 && sh.tau[qi] > -INFINITY) atomicMax(p.gtau + q0 + qi, ord_bits(sh.tau[qi]));
                }
            group_sync(EPI_BAR, EPI_NT);
            for (uint32_t qi = tid - EPI_WARP0 * 32; qi < nq; qi += EPI_NT) {  // adopt the shared thresholds
                const uint32_t gt = __ldcg(p.gtau + q0 + qi);
                if (gt) sh.tau[qi] = fmaxf(sh.tau[qi], ord_to_float(gt));
            }
        }
    }

    // ---- emit: everything inside this CTA's margin AND above the shared threshold goes to the query's shared list
    __syncthreads();
    {
        const bool cosine = p.metric == NK_METRIC_COSINE;
        uint64_t *my_cand = p.cand + (size_t)blockIdx.x * QT * PB;
        for (uint32_t qi = warp; qi < nq; qi += NTHREADS / 32) {
            const float margin2 = bf16_margin2(p.metric, __uint_as_float(sh.max_ra), cosine ? 1.0f : __uint_as_float(sh.max_rb),
                                               __uint_as_float(sh.maxxx), sh.qa[qi], sh.qb[qi], sh.qn[qi]);
            float floor_tau = p.min_score;
            const uint32_t gt = __ldcg(p.gtau + q0 + qi);
            if (gt) floor_tau = fmaxf(floor_tau, ord_to_float(gt));
            if (sh.cnt[qi] <= (int)p.k_emit) {
                const float t = fmaxf(sh.tau[qi], floor_tau);
                uint64_t thr = t > -INFINITY ? (uint64_t)ord_bits(t) << 32 : 1ull;
                if (thr == 0ull) thr = 1ull;
                warp_emit_above(my_cand + (size_t)qi * PB, sh.cnt[qi], thr, lane, p.partial + (size_t)(q0 + qi) * p.list_cap,
                                (int)p.list_cap, p.gcount + q0 + qi);
                continue;
            }
            warp_prune<PRUNE_LANE>(my_cand + (size_t)qi * PB, &sh.cnt[qi], &sh.tau[qi], p.k, lane,
                                   p.partial + (size_t)(q0 + qi) * p.list_cap, (int)p.list_cap, true, margin2, (int)p.k_emit, floor_tau,
                                   p.gcount + q0 + qi, p.flags + FLAG_OVERFLOW);
            if (lane == 0 && sh.cnt[qi] >= (int)p.k_emit && (int)p.k_emit > (int)p.k) atomicOr(p.flags + FLAG_OVERFLOW, 2);
            if (lane == 0 && sh.tau[qi] > -INFINITY) atomicMax(p.gtau + q0 + qi, ord_bits(sh.tau[qi]));
        }
        if (tid == 0) {
            atomicMax(reinterpret_cast<unsigned int *>(p.flags + FLAG_MAXXX), sh.maxxx);
            atomicMax(reinterpret_cast<unsigned int *>(p.flags + FLAG_MAX_RA), sh.max_ra);
            atomicMax(reinterpret_cast<unsigned int *>(p.flags + FLAG_MAX_RB), cosine ? __float_as_uint(1.0f) : sh.max_rb);
        }
    }
    // ---- teardown: both CTAs must be done with TMEM and with each other's barriers
    ptx::tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_pair(tmem, TMEM_COLS);
}

bool pair_pass_supported(const DeviceInfo &di, const ScanArgs &a, uint32_t grid) {
    if (!tc_env_int("NK_PAIR", 1)) return false;
    return shadow_pass_supported(di, a) && grid >= 4 && di.max_smem_optin >= (size_t)pr::RING_BYTES + sizeof(pr::Shared) + 1024;
}

// One launch: queries [q0, q0 + nq), nq <= qgroups * 256, on CTA pairs (grid rounded down to whole pairs x groups).
int launch_pair_pass(const DeviceInfo &di, const ScanArgs &a, Workspace &ws, const ShadowPassArgs &sp, uint64_t *launches) {
    using namespace pr;
    const bool native = a.shadow_native;
    const CUtensorMap *map_rows = tc_cached_map(ws, 7, a.shadow, a.n, native ? a.dim : sp.dimpad, 2, BKB, ROWS_CTA,
                                                (uint64_t)(native ? a.dim : sp.dimpad) * 2, a.dtype);
    const CUtensorMap *map_q = tc_cached_map(ws, 6, sp.qbf16, sp.Qpad, sp.dimpad, 2, BKB, QH, (uint64_t)sp.dimpad * 2, a.dtype);
    if (!map_rows || !map_q) return -1;
    const size_t smem = (size_t)RING_BYTES + sizeof(Shared) + 1024;
    if (tc_ensure_smem(reinterpret_cast<const void *>(knn_scan_pair_kernel), di.device_id, smem)) return -1;
    uint32_t C = (uint32_t)tc_env_int("NK_PAIR_CLUSTER", 1);  // pairs per cluster (query half-slabs multicast across them)
    if (C != 2 && C != 4) C = 1;
    const uint32_t unit = 2 * C * sp.qgroups;  // whole clusters, the same number per query group
    uint32_t grid = sp.grid / unit * unit;
    if (grid == 0) { C = 1; grid = sp.grid / (2 * sp.qgroups) * (2 * sp.qgroups); }
    tc::Params p{};
    p.n = a.n; p.dim = a.dim; p.nslab = sp.dimpad / BKB; p.row_base = a.row_base;
    p.q0 = sp.q0; p.nq = sp.nq; p.k = a.k; p.qpad_off = sp.q0; p.qgroups = sp.qgroups; p.list_cap = sp.grid * sp.k_emit;
    p.metric = a.metric; p.k_emit = sp.k_emit; p.qnorm = sp.qnorm; p.qa = sp.qa; p.qb = sp.qb;
    p.xnorm2 = a.xnorm2; p.dnorm2 = a.dnorm2;
    p.cand = ws.cand; p.partial = ws.partial; p.flags = ws.flags; p.mask = a.row_mask;
    p.gtau = reinterpret_cast<uint32_t *>(ws.keys2); p.gcount = reinterpret_cast<int *>(ws.keys2) + (sp.Qpad + QT_BIG);
    p.presampled = 0; p.min_score = a.min_score; p.op_f16 = a.dtype == NK_DTYPE_F16;
    p.prune_trigger = tc_env_int("NK_PRUNE_TRIGGER", 0);
    p.sync_every = (uint32_t)tc_env_int("NK_PAIR_SYNC", 0);
    if (p.sync_every && sp.qgroups > 1) {
        if (ws_reserve((void **)&ws.below, &ws.below_bytes, 1024)) return -1;  // (the big-k bound array doubles as the counter block)
        NK_CUDA_OK(cudaMemsetAsync(ws.below, 0, 1024, a.stream));
        p.sync = reinterpret_cast<uint32_t *>(ws.below);
    } else {
        p.sync_every = 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = smem; cfg.stream = a.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2 * C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (C > 1) {
        // a GPC may not fit as many clusters of 2C SMs as it has SMs: never launch more clusters than can be co-resident
        // (the kernel is persistent: every CTA must be running for the cluster barriers to complete)
        int max_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&max_clusters, knn_scan_pair_kernel, &cfg) == cudaSuccess && max_clusters > 0) {
            uint32_t ncl = grid / (2 * C);
            if ((uint32_t)max_clusters < ncl) ncl = (uint32_t)max_clusters / sp.qgroups * sp.qgroups;
            if (ncl == 0) { set_error("pair kernel: no room for a cluster of %u CTAs", 2 * C); return -1; }
            grid = ncl * 2 * C;
            cfg.gridDim = dim3(grid);
        } else {
            cudaGetLastError();
        }
    }
    NK_CUDA_OK(cudaLaunchKernelEx(&cfg, knn_scan_pair_kernel, *map_rows, *map_q, p));
    NK_CUDA_OK(cudaGetLastError());
    if (launches) ++*launches;
    if (a.main_launches) ++*a.main_launches;
    return 0;
}

}  // namespace nk
