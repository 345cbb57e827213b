// Coarse scoring as a bf16 tcgen05 GEMM with the top-K selection fused into the epilogue, so the
// [Q, N] score matrix of `scores = np.dot(vecs.T, qvecs)` (src/main_retrieve.py:175) never
// reaches HBM; what leaves the kernel is, per query and per database split, a short list of
// candidate rows that stage 2 (finalise.cu) rescores exactly.
//
// Tile: D[128 queries x 256 rows] += Q16[128 x 64] * DB16[256 x 64]^T, K swept in 64-element
// (128-byte, one swizzle atom) steps.  Queries are the UMMA M dimension = TMEM lanes, database
// rows the N dimension = TMEM columns, so every epilogue thread owns ONE query and streams that
// query's scores: the running threshold is a register compare, survivors are appended to the
// thread's own list (no atomics), and a list that fills up is trimmed by its warp (radix select).
//
// Warp roles (192 threads, 1 CTA/SM, persistent over jobs):
//   warp 0      TMA producer   4-stage smem ring, 48 KB per stage, mbarrier full/empty
//   warp 1      MMA issuer     one lane issues tcgen05.mma (M128 N256 K16, 4 per stage);
//                              owns the 512-column TMEM allocation = 2 accumulators
//   warps 2..5  epilogue       tcgen05.ld 32x32b.x32 -> filter -> lists; overlap the next tile's MMA
// A job is (query tile, database split); jobs are dealt round-robin to the CTAs.
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "select.cuh"
#include "internal.h"

namespace xs {

// Two shapes of the same kernel:
//   PAIR = false  one CTA per 128 x 256 tile (cta_group::1), 4 stages x 48 KB.  Small query batches: HBM-bound.
//   PAIR = true   a cluster of two CTAs per 256 x 256 tile (cta_group::2): each CTA stages ITS 128 query rows and
//                 ITS 128 of the 256 database rows (16 + 16 KB per stage, 6 stages), the leader issues M=256 MMAs
//                 that read both CTAs' shared memory and write both CTAs' TMEM.  Halves the L2->SM operand
//                 traffic per flop and deepens the ring: large query batches, tensor-pipe-bound.
constexpr int MAX_STAGES = 6;
constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;          // 16 KB
//   HALF          one CTA per 128 x 128 tile (cta_group::1), 6 stages x 32 KB: the per-CTA staging of PAIR without
//                 the pairing.  Used by the threshold-bootstrap pass, which is one tile per CTA and purely
//                 latency-bound: a deeper ring of smaller stages finishes the 32 K-steps sooner.
enum { MODE_FULL = 0, MODE_PAIR = 1, MODE_HALF = 2, MODE_FULL3 = 3 };
//   FULL3         FULL with a 3-stage ring (144 KB): leaves 70 KB of the SM's shared memory free, so that the selection /
//                 rescoring kernel of the PREVIOUS batch can be resident next to it and run under this batch's scan
template <int MODE> struct Shape {
    static constexpr bool PAIR = MODE == MODE_PAIR;
    static constexpr int STAGES = MODE == MODE_FULL ? 4 : (MODE == MODE_FULL3 ? 3 : 6);
    static constexpr int B_ROWS = (MODE == MODE_FULL || MODE == MODE_FULL3) ? GEMM_BN : GEMM_BN / 2;     // database rows this CTA stages per step
    static constexpr int TILE_N = MODE == MODE_HALF ? GEMM_BN / 2 : GEMM_BN;     // database rows per tile (UMMA N)
    static constexpr int B_BYTES = B_ROWS * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int UMMA_M = PAIR ? 256 : 128;
    // kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N = TILE_N, M = 128 | 256
    static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TILE_N >> 3) << 17) | ((uint32_t)(UMMA_M >> 4) << 24);
};
constexpr int GEMM_THREADS = 192;
constexpr int EPI_WARP0 = 2;
constexpr int TMEM_COLS = 512;
constexpr uint64_t HINT_EVICT_FIRST = 0x12F0000000000000ull;   // database tiles: streamed once
constexpr uint64_t HINT_EVICT_LAST  = 0x14F0000000000000ull;   // query tiles: re-read by every CTA

struct __align__(8) GemmBarriers {
    uint64_t full[MAX_STAGES], empty[MAX_STAGES], tfull[2], tempty[2];
    uint32_t tmem_base;
    uint32_t pad;
};
// both shapes stage 192 KB of operands
constexpr size_t GEMM_SMEM_EXTRA = 1024 /*alignment slack*/ + sizeof(GemmBarriers) + 4 * 256 * sizeof(uint32_t) + 16 + (size_t)BOOT_MAX_GRID * 8 * sizeof(float);
constexpr size_t GEMM_SMEM = (size_t)4 * (A_BYTES + GEMM_BN * GEMM_BK * 2) + GEMM_SMEM_EXTRA;
constexpr size_t GEMM_SMEM3 = (size_t)3 * (A_BYTES + GEMM_BN * GEMM_BK * 2) + GEMM_SMEM_EXTRA;
static_assert(Shape<MODE_PAIR>::STAGES * Shape<MODE_PAIR>::STAGE_BYTES == Shape<MODE_FULL>::STAGES * Shape<MODE_FULL>::STAGE_BYTES, "ring sizes differ");
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;        // shared::cluster address -> same offset in the pair's leader CTA

// ---- PTX wrappers ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must trap, not hang the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000ll) { printf("xs gemm_topk: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint64_t hint) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
                 " [%0], [%1, {%3, %4}], [%2], %5;"
                 :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "l"(hint) : "memory");
}
// cta_group::2 load: lands in THIS CTA's shared memory, completes bytes on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint64_t hint) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
                 " [%0], [%1, {%3, %4}], [%2], %5;"
                 :: "r"(dst), "l"(map), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1), "l"(hint) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
                 :: "r"(bar), "r"(cta) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {      // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// device-scope handshake words of the in-kernel threshold bootstrap
__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
    uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ uint64_t ld_acquire_gpu_u64(const uint64_t* p) {
    uint64_t v; asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_relaxed_gpu_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
// counter += v with release semantics: everything this thread wrote (and observed through a CTA barrier) before it is
// visible to whoever acquires the new count -- one operation instead of fence + atomic
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define BOOT_STAMP(i) do { if (boot.trace && sub == 0 && lane == 0) boot.trace[(size_t)blockIdx.x * 8 + (i)] = globaltimer_ns(); } while (0)

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu)       // start address, 16-byte units
         | ((uint64_t)(1024u >> 4) << 32)               // stride byte offset
         | (1ull << 46)                                 // descriptor version (sm_100)
         | (2ull << 61);                                // SWIZZLE_128B
}

// ---- the kernel ------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_db, const __grid_constant__ CUtensorMap tmap_db_h,
                 int m_tiles, int n_tiles, int tile_stride, int splits, int half_units, int k_blocks, int a_row0, int64_t nq, int64_t n_valid,
                 int k, int k_keep, int cap, int sample_mode, int db_tiled, uint64_t hint_db, const float* __restrict__ eps, const float* __restrict__ thr0,
                 uint64_t* __restrict__ pool_items, int* __restrict__ pool_count, uint32_t* __restrict__ pool_thr, const InlineBoot boot) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    using S = Shape<MODE>;
    constexpr bool PAIR = S::PAIR;
    constexpr int STAGES = S::STAGES, B_BYTES = S::B_BYTES, STAGE_BYTES = S::STAGE_BYTES, TILE_N = S::TILE_N;
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    GemmBarriers* bars = reinterpret_cast<GemmBarriers*>(smem + STAGES * STAGE_BYTES);
    uint32_t* whist = reinterpret_cast<uint32_t*>(bars + 1);          // [4 warps][256]
    // [BOOT_MAX_GRID * 8] one query's threshold sample (in-kernel bootstrap), 16-byte aligned for cp.async
    float* bsamp = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(whist + 4 * 256) + 15) & ~(uintptr_t)15);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // scheduling unit = CTA (or CTA pair); a unit works on jobs (query-tile unit mu, database split sp), dealt
    // split-major so that units running at the same time sweep the SAME database tiles (L2 reuse) for different queries
    const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
    const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int n_units = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int m_units = PAIR ? (m_tiles + 1) / 2 : m_tiles;
    const int n_jobs = m_units * splits;
    // The database tiles of a job.  Default: whole tiles t0 .. t1.  half_units (single-CTA shapes, main pass): the database is
    // dealt in HALF tiles of 128 rows so that the jobs differ by at most half a tile (3 or 4 tiles per CTA on an 8-GPU shard is a
    // 33 % longer critical path, 3 or 3.5 is 17 %); a job is then [lone half] whole tiles ... [lone half], a lone half being a
    // 128-row TMA box (second tensor map) and an N = 128 MMA into the same accumulator.
    constexpr uint32_t IDESC_HALF = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((GEMM_BN / 2) >> 3) << 17) | ((uint32_t)(S::UMMA_M >> 4) << 24);
    auto job_range = [&](int sp, int& u, int& u1) {
        const int total = half_units ? 2 * n_tiles : n_tiles;
        u = (int)((int64_t)sp * total / splits); u1 = (int)((int64_t)(sp + 1) * total / splits);
    };
    auto next_unit = [&](int& u, int u1, int& row0, int& ncols) {
        if (half_units) {
            const bool half = (u & 1) || (u + 2 > u1);
            row0 = u * (GEMM_BN / 2); ncols = half ? GEMM_BN / 2 : GEMM_BN; u += half ? 1 : 2;
        } else { row0 = u * tile_stride * TILE_N; ncols = TILE_N; u += 1; }
    };

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_db) : "memory");
        if (half_units) asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_db_h) : "memory");
        // pair: the leader's `full` collects its own arrive(+expect_tx) and the peer's remote arrive; the
        // leader's `tempty` collects the 128 epilogue threads of BOTH CTAs
        for (int i = 0; i < STAGES; ++i) { mbar_init(smem_u32(&bars->full[i]), PAIR ? 2 : 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->tfull[i]), 1); mbar_init(smem_u32(&bars->tempty[i]), PAIR ? 256 : 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"(smem_u32(&bars->tmem_base)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"(smem_u32(&bars->tmem_base)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();              // the peer's barriers exist before anything signals them
    tc_fence_after();
    pdl_wait();                                          // barriers, TMEM and descriptors are set up; now the predecessor's data
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int job = unit; job < n_jobs; job += n_units) {
                const int sp = job / m_units, mu = job - sp * m_units;
                const int mt = PAIR ? 2 * mu + (int)cta_rank : mu;
                int u, u1;
                job_range(sp, u, u1);
                while (u < u1) {
                    int row0, ncols;
                    next_unit(u, u1, row0, ncols);
                    for (int kb = 0; kb < k_blocks; ++kb) {
                        mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
                        const uint32_t full = smem_u32(&bars->full[stage]);
                        // database operand coordinates: row-major [n_pad][d_pad] -> (k, row); tiled
                        // [n_pad/256][d_pad/64][256][64] (every box one contiguous 16/32 KB run) -> (0, ((row/256)*KB + kb)*256 + row%256)
                        const int brow = row0 + (PAIR ? (int)cta_rank * S::B_ROWS : 0);
                        const int bc0 = db_tiled ? 0 : kb * GEMM_BK;
                        const int bc1 = db_tiled ? (((brow >> 8) * k_blocks + kb) << 8) + (brow & 255) : brow;
                        if constexpr (PAIR) {
                            if (cta_rank == 0) mbar_expect_tx(full, 2 * STAGE_BYTES);      // both CTAs' bytes land on the leader's barrier
                            else mbar_arrive_remote(full, 0);
                            tma_load_2d_pair(smem_u32(sA + stage * A_BYTES), &tmap_q, full, kb * GEMM_BK, a_row0 + mt * GEMM_BM, HINT_EVICT_LAST);
                            tma_load_2d_pair(smem_u32(sB + stage * B_BYTES), &tmap_db, full, bc0, bc1, HINT_EVICT_LAST);
                        } else {
                            const bool lone_half = ncols != TILE_N;
                            mbar_expect_tx(full, A_BYTES + ncols * GEMM_BK * 2);
                            tma_load_2d(smem_u32(sA + stage * A_BYTES), &tmap_q, full, kb * GEMM_BK, a_row0 + mt * GEMM_BM, HINT_EVICT_LAST);
                            tma_load_2d(smem_u32(sB + stage * B_BYTES), lone_half ? &tmap_db_h : &tmap_db, full, bc0, bc1, hint_db);
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ====================================== MMA issuer ======================================
        if (lane == 0 && cta_rank == 0) {                      // pair: only the leader issues
            int stage = 0; uint32_t phase = 0; uint32_t it = 0;
            for (int job = unit; job < n_jobs; job += n_units) {
                const int sp = job / m_units;
                int u, u1;
                job_range(sp, u, u1);
                for (; u < u1; ++it) {
                    int row0, ncols;
                    next_unit(u, u1, row0, ncols);
                    const uint32_t idesc = (ncols == TILE_N) ? S::IDESC : IDESC_HALF;
                    const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
                    mbar_wait(smem_u32(&bars->tempty[acc]), acc_phase ^ 1);      // epilogue(s) drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * GEMM_BN;
                    for (int kb = 0; kb < k_blocks; ++kb) {
                        mbar_wait(smem_u32(&bars->full[stage]), phase);          // TMA bytes landed (in both CTAs)
                        tc_fence_after();
                        const uint64_t da = umma_desc_sw128(smem_u32(sA + stage * A_BYTES));
                        const uint64_t db = umma_desc_sw128(smem_u32(sB + stage * B_BYTES));
#pragma unroll
                        for (int kk = 0; kk < GEMM_BK / 16; ++kk) {               // +32 B per K=16 step inside the atom
                            if constexpr (PAIR) tc_mma_bf16_pair(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (kb | kk) != 0);
                            else                tc_mma_bf16(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (kb | kk) != 0);
                        }
                        // frees the smem slot (in both CTAs) when the MMAs retire
                        if constexpr (PAIR) tc_commit_pair(smem_u32(&bars->empty[stage])); else tc_commit(smem_u32(&bars->empty[stage]));
                        if (kb == k_blocks - 1) {
                            if constexpr (PAIR) tc_commit_pair(smem_u32(&bars->tfull[acc])); else tc_commit(smem_u32(&bars->tfull[acc]));
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ======================================= epilogue =======================================
        const int ew = warp - EPI_WARP0;                       // 0..3, private histogram
        const int sub = warp & 3;                              // TMEM sub-partition this warp may read
        const int m = sub * 32 + lane;                         // query row inside the tile = TMEM lane
        uint32_t* my_hist = whist + ew * 256;
        uint32_t it = 0;
        // accumulator hand-back: the leader's MMA thread waits for both CTAs' epilogues
        auto release_acc = [&](uint32_t acc) {
            tc_fence_before();
            if (PAIR && cta_rank != 0) mbar_arrive_remote(smem_u32(&bars->tempty[acc]), 0);
            else mbar_arrive(smem_u32(&bars->tempty[acc]));
        };
        for (int job = unit; job < n_jobs; job += n_units) {
            const int sp = job / m_units, mu = job - sp * m_units;
            const int mt = PAIR ? 2 * mu + (int)cta_rank : mu;
            int u, u1;
            job_range(sp, u, u1);
            const int64_t q = (int64_t)mt * GEMM_BM + m;
            const bool active = q < nq;
            const int64_t slot = ((int64_t)mt * splits + sp) * GEMM_BM + m;
            uint64_t* list = pool_items + slot * cap;
            int cnt = 0;
            float thr = active ? (thr0 ? thr0[q] : -INFINITY) : INFINITY;
            uint32_t thr_rec = 0;
            if (__ballot_sync(0xffffffffu, active) == 0u) {
                // all 32 query rows of this warp are tile padding (e.g. rows 96..127 of a 70-query batch): nothing to
                // read back -- only hand the accumulators over so that the MMA thread can go on
                for (; u < u1; ++it) {
                    int row0, ncols;
                    next_unit(u, u1, row0, ncols);
                    const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
                    mbar_wait(smem_u32(&bars->tfull[acc]), acc_phase);
                    release_acc(acc);
                }
                pool_count[slot] = 0;
                pool_thr[slot] = 0u;
                continue;
            }
            if (sample_mode) {
                // Threshold bootstrap: per tile only 8 scores are kept -- the maxima of its 8 column groups (branch-free:
                // one max per score).  They are real scores, so the k-th best of the union over all sampled tiles is a
                // valid lower bound of the database's k-th best.  No lists, no trims.  (One tile per job.)
                for (; u < u1; ++it) {
                    int row0, ncols;
                    next_unit(u, u1, row0, ncols);
                    const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
                    mbar_wait(smem_u32(&bars->tfull[acc]), acc_phase);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + acc * GEMM_BN;
#pragma unroll 1
                    for (int c = 0; c < TILE_N / 32; ++c) {
                        uint32_t v[32];
                        tc_ld32(taddr + c * 32, v);
                        tc_ld_wait();
                        const int64_t lim = n_valid - ((int64_t)row0 + c * 32);
                        constexpr int G = TILE_N / 8;                          // columns per group: 32 (256-row tiles) or 16
                        float gmax[32 / G];
#pragma unroll
                        for (int g = 0; g < 32 / G; ++g) gmax[g] = -INFINITY;
#pragma unroll
                        for (int i = 0; i < 32; ++i) gmax[i / G] = fmaxf(gmax[i / G], (i < lim) ? __uint_as_float(v[i]) : -INFINITY);
#pragma unroll
                        for (int g = 0; g < 32 / G; ++g) if (active && gmax[g] > -INFINITY && cnt < 8) list[cnt++] = make_item(gmax[g], 0u);
                    }
                    release_acc(acc);
                }
                pool_count[slot] = cnt;
                pool_thr[slot] = 0u;
                continue;
            }
            // one accumulator -> this thread's list: keep what beats the running threshold, trim a list that could overflow
            auto filter_tile = [&](uint32_t acc, int row0, int ncols) {
                const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + acc * GEMM_BN;
#pragma unroll 1
                for (int c = 0; c < ncols / 32; ++c) {
                    uint32_t v[32];
                    tc_ld32(taddr + c * 32, v);
                    tc_ld_wait();
                    const int64_t row_base = (int64_t)row0 + c * 32;
                    const int64_t lim = n_valid - row_base;                 // rows >= n_valid are zero padding
                    if (lim >= 32) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float s = __uint_as_float(v[i]);
                            if (s > thr) list[cnt++] = make_item(s, (uint32_t)(row_base + i));
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float s = __uint_as_float(v[i]);
                            if (s > thr && i < lim) list[cnt++] = make_item(s, (uint32_t)(row_base + i));
                        }
                    }
                    // a list that could overflow on the next chunk is trimmed to its k_keep best by the warp
                    uint32_t need = __ballot_sync(0xffffffffu, cnt > cap - 32);
                    while (need) {
                        const int src = __ffs(need) - 1;
                        need &= need - 1;
                        uint64_t* l = reinterpret_cast<uint64_t*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(list), src));
                        const int c_src = __shfl_sync(0xffffffffu, cnt, src);
                        uint32_t T;
                        const int c_new = warp_prune_edge(l, c_src, k_keep, -1.f, cap, 64, my_hist, T);
                        if (lane == src) { cnt = c_new; thr = fmaxf(thr, key_score(T)); thr_rec = max(thr_rec, T); }
                    }
                }
            };
            if (boot.on) {
                int row_first, ncols_first;
                next_unit(u, u1, row_first, ncols_first);          // the job's first unit (a whole tile or a lone half) is the sample
                // ---- in-kernel threshold bootstrap (one job per CTA, every CTA resident) ------------------------------
                // The FIRST tile of every CTA doubles as the sample: (A) its 8 best scores per query go to global memory,
                // (B) once all CTAs have arrived, CTA c takes the k-th best of query c's sample (a valid lower bound of the
                // database's k-th best) minus the band and publishes it, (C) every thread picks up its query's threshold
                // and only then filters the tile, which has been waiting in its TMEM accumulator.  Meanwhile TMA and MMA
                // run ahead into the ring and the second accumulator, so the exchange hides behind the second tile's
                // HBM stream.  Replaces the separate bootstrap GEMM + threshold-select launches.
                BOOT_STAMP(0);                                   // job start
                mbar_wait(smem_u32(&bars->tfull[0]), 0);
                tc_fence_after();
                BOOT_STAMP(1);                                   // first tile accumulated
                {
                    const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16);
                    float* dst = boot.samp + ((size_t)q * gridDim.x + blockIdx.x) * 8;
#pragma unroll 1
                    for (int c = 0; c < 8; ++c) {                            // the maximum of every 32-column group: one max per score
                        float gmax = -INFINITY;
                        if (c < ncols_first / 32) {                          // (uniform)
                            uint32_t v[32];
                            tc_ld32(taddr + c * 32, v);
                            tc_ld_wait();
                            const int64_t lim = n_valid - ((int64_t)row_first + c * 32);
#pragma unroll
                            for (int i = 0; i < 32; ++i) gmax = fmaxf(gmax, (i < lim) ? __uint_as_float(v[i]) : -INFINITY);
                        }
                        if (active) dst[c] = gmax;
                    }
                }
                // (no fence per thread: the barrier orders these stores before the leader's fence + atomic below)
                // warps that hold queries are those with sub < n_warps_on (a query's tile row is sub * 32 + lane); only they get here
                const int n_warps_on = (int)((min(nq, (int64_t)GEMM_BM) + 31) >> 5);
                asm volatile("bar.sync 1, %0;" :: "r"(n_warps_on * 32) : "memory");
                if (sub == 0) {                                  // the warp that holds queries 0..31 (TMEM sub-partition 0): always populated
                    if (lane == 0) red_release_gpu_add(boot.arrive, 1u);
                    BOOT_STAMP(2);                               // sample written, arrival posted
                    if ((int64_t)blockIdx.x < nq) {
                        // (B) this CTA owns queries blockIdx.x, blockIdx.x + grid, ...
                        if (lane == 0) {
                            const long long w0 = clock64();
                            while ((int32_t)(ld_acquire_gpu_u32(boot.arrive) - boot.arrive_target) < 0) {
                                __nanosleep(40);
                                if (clock64() - w0 > 8000000000ll) { printf("xs gemm_topk: bootstrap arrival timeout (block %d)\n", blockIdx.x); __trap(); }
                            }
                        }
                        __syncwarp();
                        BOOT_STAMP(3);                           // (owners) every CTA has arrived
                        const int n_samp = (int)gridDim.x * 8;
                        for (int64_t qq = blockIdx.x; qq < nq; qq += gridDim.x) {
                            // the query's sample (grid x 8 scores, ~5 KB) comes in with ONE round trip: cp.async straight
                            // from L2 into shared memory (forty dependent register loads per lane took 20 us here)
                            const float* src = boot.samp + (size_t)qq * gridDim.x * 8;
                            const uint32_t sdst = smem_u32(bsamp);
                            for (int i = lane * 4; i < n_samp; i += 128)
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sdst + (uint32_t)i * 4u), "l"(src + i) : "memory");
                            asm volatile("cp.async.commit_group;" ::: "memory");
                            asm volatile("cp.async.wait_group 0;" ::: "memory");
                            __syncwarp();
                            int valid_cnt = 0;
                            uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
                            for (int i = lane; i < n_samp; i += 32) {
                                const float s = bsamp[i];
                                if (s > -INFINITY) { const uint32_t key = score_key(s); kmin = min(kmin, key); kmax = max(kmax, key); ++valid_cnt; }
                            }
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) {
                                valid_cnt += __shfl_xor_sync(0xffffffffu, valid_cnt, o);
                                kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
                                kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
                            }
                            float t_pub = -INFINITY;
                            if (valid_cnt >= k) {
                                // conservative k-th best: ONE pass into 256 bins spread over [min, max] of the keys, lower edge of
                                // the bin that holds the k-th largest -- never above it, at most one bin below
                                const uint32_t range = kmax - kmin;
                                const int shift = (range >> 8) ? (32 - __clz(range) - 8) : 0;
#pragma unroll
                                for (int j = 0; j < 8; ++j) my_hist[lane * 8 + j] = 0;
                                __syncwarp();
                                for (int i = lane; i < n_samp; i += 32) {
                                    const float s = bsamp[i];
                                    if (s > -INFINITY) atomicAdd(&my_hist[(score_key(s) - kmin) >> shift], 1u);
                                }
                                __syncwarp();
                                uint32_t dg, kr;
                                warp_pick_digit(my_hist, (uint32_t)k, dg, kr);
                                __syncwarp();
                                // the main pass keeps scores strictly above the threshold: that edge, minus the band
                                t_pub = nextafterf(key_score(kmin + (dg << shift)) - 2.f * eps[qq], -INFINITY);
                            }
                            BOOT_STAMP(7);                       // (owners) own threshold selected
                            if (lane == 0) {
                                st_relaxed_gpu_u64(boot.thr_pub + qq, ((uint64_t)boot.epoch << 32) | (uint64_t)__float_as_uint(t_pub));
                                red_release_gpu_add(boot.published, 1u);
                            }
                        }
                    }
                    // (C) ONE thread per CTA waits until every query's threshold is out (ten thousand threads polling their
                    // own words would swamp the L2 that the operand stream needs), the named barrier releases the rest
                    if (lane == 0) {
                        const long long w0 = clock64();
                        while ((int32_t)(ld_acquire_gpu_u32(boot.published) - boot.published_target) < 0) {
                            __nanosleep(100);
                            if (clock64() - w0 > 8000000000ll) { printf("xs gemm_topk: bootstrap threshold timeout (block %d)\n", blockIdx.x); __trap(); }
                        }
                    }
                    __syncwarp();
                    BOOT_STAMP(4);                               // every threshold is out
                }
                asm volatile("bar.sync 1, %0;" :: "r"(n_warps_on * 32) : "memory");
                if (active) {
                    const uint64_t w = __ldcg(reinterpret_cast<const unsigned long long*>(boot.thr_pub + q));
                    if ((uint32_t)(w >> 32) != boot.epoch) { printf("xs gemm_topk: stale bootstrap threshold (block %d)\n", blockIdx.x); __trap(); }
                    thr = __uint_as_float((uint32_t)w);
                }
                __syncwarp();
                filter_tile(0u, row_first, ncols_first);
                release_acc(0u);
                BOOT_STAMP(5);                                   // first accumulator handed back
                it = 1;
            }
            for (; u < u1; ++it) {
                int row0, ncols;
                next_unit(u, u1, row0, ncols);
                const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
                mbar_wait(smem_u32(&bars->tfull[acc]), acc_phase);
                tc_fence_after();
                filter_tile(acc, row0, ncols);
                release_acc(acc);
            }
            // end of job: keep only what can still matter -- everything within 2*eps below this
            // split's k-th best (the global k-th best is at least as large)
            const float band = active ? 2.f * eps[q] : 0.f;
            // (short lists are left alone: the finaliser reads them anyway and a trim costs more than it saves)
            uint32_t need = __ballot_sync(0xffffffffu, cnt > k_keep);
            while (need) {
                const int src = __ffs(need) - 1;
                need &= need - 1;
                uint64_t* l = reinterpret_cast<uint64_t*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(list), src));
                const int c_src = __shfl_sync(0xffffffffu, cnt, src);
                const float b_src = __shfl_sync(0xffffffffu, band, src);
                uint32_t T;
                const int c_new = warp_prune_edge(l, c_src, k, b_src, cap, 0, my_hist, T);
                if (lane == src) cnt = c_new;
            }
            pool_count[slot] = cnt;
            pool_thr[slot] = thr_rec;
            if (boot.on) BOOT_STAMP(6);                          // job done
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();              // the leader's MMAs wrote the peer's TMEM: both must be done
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else                asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- host side -------------------------------------------------------------------------------------
GemmPlan plan_gemm(int64_t nq, int64_t n_pad, int k, int num_sms, int forced_splits, bool allow_pair) {
    GemmPlan p{};
    p.m_tiles = (int)((nq + GEMM_BM - 1) / GEMM_BM);
    p.n_tiles = (int)(n_pad / GEMM_BN);
    p.pair = (allow_pair && p.m_tiles >= 2) ? 1 : 0;
    const int units = p.pair ? num_sms / 2 : num_sms;             // CTAs or CTA pairs
    const int m_units = p.pair ? (p.m_tiles + 1) / 2 : p.m_tiles;
    int best = 1;
    if (forced_splits > 0) best = forced_splits;
    else {
        // pick the split count that fills whole waves of `units` jobs while keeping jobs long
        double best_score = -1.0;
        const int s_max = (int)((p.n_tiles < 4096) ? p.n_tiles : 4096);
        for (int s = 1; s <= s_max; ++s) {
            const int64_t jobs = (int64_t)m_units * s;
            const int64_t rounds = (jobs + units - 1) / units;
            const double tiles_per_job = (double)((p.n_tiles + s - 1) / s);     // tiles per unit on the critical path
            const double crit = (double)rounds * tiles_per_job;
            const double ideal = (double)m_units * p.n_tiles / units;
            const double score = ideal / crit - 0.0005 * s;      // prefer fewer, longer jobs on ties
            if (score > best_score) { best_score = score; best = s; }
            if (jobs > (int64_t)units * 64) break;
        }
    }
    if (best > p.n_tiles) best = p.n_tiles > 0 ? p.n_tiles : 1;
    p.splits = best;
    int keep = k + (k / 2 > 156 ? k / 2 : 156);
    p.k_keep = (keep + 31) & ~31;
    p.cap = 2 * p.k_keep;
    const int64_t jobs = (int64_t)m_units * p.splits;
    const int used = (int)(jobs < units ? jobs : units);
    p.grid = p.pair ? 2 * used : used;
    p.tile_stride = 1;
    p.sample_mode = 0;
    p.half = 0;
    // in-kernel threshold bootstrap: one query tile, exactly one job per CTA (so every CTA is resident and its first tile
    // can serve as the sample), and a sample that holds well over k scores
    p.inline_boot = 0;
    p.half_units = p.pair ? 0 : 1;                    // single-CTA shapes deal the database in half tiles (balance)
    if (!p.pair && p.m_tiles == 1 && forced_splits <= 0 && p.n_tiles >= 16) {
        const int s = p.n_tiles < num_sms ? p.n_tiles : num_sms;
        if (s <= BOOT_MAX_GRID && 8 * s >= (5 * k + 3) / 4 + 8) { p.splits = s; p.grid = s; p.inline_boot = 1; }
    }
    return p;
}

// Threshold bootstrap: S database tiles spread evenly over the database (stride), one tile per job.
// The k-th best score of that sample is a valid lower bound of the k-th best of the whole database.
GemmPlan plan_gemm_sample(const GemmPlan& main_plan, int num_sms, int k) {
    GemmPlan p = main_plan;
    p.pair = 0;
    p.half = 1;                                      // 128-row tiles, 6-stage ring: one tile per CTA, latency-bound
    const int all_tiles = main_plan.n_tiles * 2;     // in 128-row tiles
    // tiles per query tile: at least 32 and 1.25 k (8 scores per tile: the union must hold well over k), then
    // rounded UP to fill whole waves of one-tile jobs
    const int mt = main_plan.m_tiles > 0 ? main_plan.m_tiles : 1;
    int s0 = (5 * k + 3) / 4;
    if (s0 < 32) s0 = 32;
    const int64_t rounds = ((int64_t)mt * s0 + num_sms - 1) / num_sms;
    int s = (int)(rounds * num_sms / mt);
    if (s < s0) s = s0;
    if (s > all_tiles) s = all_tiles;
    p.n_tiles = s;
    p.tile_stride = all_tiles / s;
    p.splits = s;                                   // one tile per job
    p.sample_mode = 1;                              // 8 group maxima per (query, tile)
    p.inline_boot = 0;
    p.stages3 = 0;
    p.half_units = 0;
    const int64_t jobs = (int64_t)p.m_tiles * p.splits;
    p.grid = (int)(jobs < num_sms ? jobs : num_sms);
    return p;
}

cudaError_t launch_gemm_topk(const CUtensorMap& tmap_q, const CUtensorMap& tmap_db, const CUtensorMap* tmap_db_half, const GemmPlan& plan,
                             int64_t nq, int64_t n_valid, int d_pad, int k, const float* eps,
                             uint64_t* pool_items, int* pool_count, uint32_t* pool_thr, int a_row0,
                             const float* thr0, const InlineBoot* boot, cudaStream_t st) {
    const bool three = plan.stages3 && !plan.pair && !plan.half;
    auto kern = plan.pair ? gemm_topk_kernel<MODE_PAIR> : (plan.half ? gemm_topk_kernel<MODE_HALF> : (three ? gemm_topk_kernel<MODE_FULL3> : gemm_topk_kernel<MODE_FULL>));
    const size_t smem_bytes = three ? GEMM_SMEM3 : GEMM_SMEM;
    // function attributes are set once per kernel shape and device (each call is a driver round trip on the per-step host path)
    const int shape = plan.pair ? 1 : (plan.half ? 2 : (three ? 3 : 0));
    static unsigned char attr_done[4][64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    cudaError_t e = cudaSuccess;
    if (dev < 0 || dev >= 64 || !attr_done[shape][dev]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
        // the SM's whole 228 KB as shared memory whatever this kernel needs: with the 3-stage ring the 70 KB it leaves are for a
        // finalise CTA of the other lane (the driver would otherwise carve out just enough for this kernel and nothing else fits)
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
        if (dev >= 0 && dev < 64) attr_done[shape][dev] = 1;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)plan.grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = plan.pair ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    const int k_blocks = d_pad / GEMM_BK;
    // streamed-once database tiles: evict-first measured 0.603 ms per pass against 0.642 ms with evict-normal / evict-last
    const uint64_t hint_db = HINT_EVICT_FIRST;
    InlineBoot ib{};
    if (plan.inline_boot) {
        if (!boot || !boot->samp || !boot->arrive || !boot->thr_pub || plan.grid != plan.splits || plan.pair || plan.half) return cudaErrorInvalidValue;
        ib = *boot; ib.on = 1;
    }
    const int half_units = (plan.half_units && tmap_db_half && !plan.pair && !plan.half && !plan.sample_mode && plan.tile_stride == 1) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, tmap_q, tmap_db, half_units ? *tmap_db_half : tmap_db, plan.m_tiles, plan.n_tiles, plan.tile_stride, plan.splits, half_units, k_blocks,
                              a_row0, nq, n_valid, k, plan.k_keep, plan.cap, plan.sample_mode, plan.db_tiled, hint_db, eps, thr0,
                              pool_items, pool_count, pool_thr, ib);
}

}  // namespace xs
