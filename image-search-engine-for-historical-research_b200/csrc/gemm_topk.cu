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
#include "common.cuh"
#include "select.cuh"
#include "internal.h"

namespace xs {

constexpr int STAGES = 4;
constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;          // 16 KB
constexpr int B_BYTES = GEMM_BN * GEMM_BK * 2;          // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int GEMM_THREADS = 192;
constexpr int EPI_WARP0 = 2;
constexpr int TMEM_COLS = 512;
constexpr uint64_t HINT_EVICT_FIRST = 0x12F0000000000000ull;   // database tiles: streamed once
constexpr uint64_t HINT_EVICT_LAST  = 0x14F0000000000000ull;   // query tiles: re-read by every CTA

struct __align__(8) GemmBarriers {
    uint64_t full[STAGES], empty[STAGES], tfull[2], tempty[2];
    uint32_t tmem_base;
    uint32_t pad;
};
constexpr size_t GEMM_SMEM = 1024 /*alignment slack*/ + (size_t)STAGES * STAGE_BYTES + sizeof(GemmBarriers) + 4 * 256 * sizeof(uint32_t);
size_t gemm_smem_bytes() { return GEMM_SMEM; }

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

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu)       // start address, 16-byte units
         | ((uint64_t)(1024u >> 4) << 32)               // stride byte offset
         | (1ull << 46)                                 // descriptor version (sm_100)
         | (2ull << 61);                                // SWIZZLE_128B
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N=256, M=128.
constexpr uint32_t UMMA_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(GEMM_BN >> 3) << 17) | ((uint32_t)(GEMM_BM >> 4) << 24);

// ---- the kernel ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_db,
                 int m_tiles, int n_tiles, int tile_stride, int splits, int k_blocks, int a_row0, int64_t nq, int64_t n_valid,
                 int k, int k_keep, int cap, int sample_mode, const float* __restrict__ eps, const float* __restrict__ thr0,
                 uint64_t* __restrict__ pool_items, int* __restrict__ pool_count, uint32_t* __restrict__ pool_thr) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    GemmBarriers* bars = reinterpret_cast<GemmBarriers*>(smem + STAGES * STAGE_BYTES);
    uint32_t* whist = reinterpret_cast<uint32_t*>(bars + 1);          // [4 warps][256]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_jobs = m_tiles * splits;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_db) : "memory");
        for (int i = 0; i < STAGES; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->tfull[i]), 1); mbar_init(smem_u32(&bars->tempty[i]), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&bars->tmem_base)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int job = blockIdx.x; job < n_jobs; job += gridDim.x) {
                const int mt = job / splits, sp = job - mt * splits;
                const int t0 = (int)((int64_t)sp * n_tiles / splits), t1 = (int)((int64_t)(sp + 1) * n_tiles / splits);
                for (int t = t0; t < t1; ++t) {
                    for (int kb = 0; kb < k_blocks; ++kb) {
                        mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
                        const uint32_t full = smem_u32(&bars->full[stage]);
                        mbar_expect_tx(full, STAGE_BYTES);
                        tma_load_2d(smem_u32(sA + stage * A_BYTES), &tmap_q, full, kb * GEMM_BK, a_row0 + mt * GEMM_BM, HINT_EVICT_LAST);
                        tma_load_2d(smem_u32(sB + stage * B_BYTES), &tmap_db, full, kb * GEMM_BK, t * tile_stride * GEMM_BN, HINT_EVICT_FIRST);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ====================================== MMA issuer ======================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0; uint32_t it = 0;
            for (int job = blockIdx.x; job < n_jobs; job += gridDim.x) {
                const int mt = job / splits, sp = job - mt * splits;
                const int t0 = (int)((int64_t)sp * n_tiles / splits), t1 = (int)((int64_t)(sp + 1) * n_tiles / splits);
                for (int t = t0; t < t1; ++t, ++it) {
                    const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
                    mbar_wait(smem_u32(&bars->tempty[acc]), acc_phase ^ 1);      // epilogue drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * GEMM_BN;
                    for (int kb = 0; kb < k_blocks; ++kb) {
                        mbar_wait(smem_u32(&bars->full[stage]), phase);          // TMA bytes landed
                        tc_fence_after();
                        const uint64_t da = umma_desc_sw128(smem_u32(sA + stage * A_BYTES));
                        const uint64_t db = umma_desc_sw128(smem_u32(sB + stage * B_BYTES));
#pragma unroll
                        for (int kk = 0; kk < GEMM_BK / 16; ++kk)                 // +32 B per K=16 step inside the atom
                            tc_mma_bf16(d_tmem, da + 2 * kk, db + 2 * kk, UMMA_IDESC, (kb | kk) != 0);
                        tc_commit(smem_u32(&bars->empty[stage]));                 // frees the smem slot when the MMAs retire
                        if (kb == k_blocks - 1) tc_commit(smem_u32(&bars->tfull[acc]));
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
        for (int job = blockIdx.x; job < n_jobs; job += gridDim.x) {
            const int mt = job / splits, sp = job - mt * splits;
            const int t0 = (int)((int64_t)sp * n_tiles / splits), t1 = (int)((int64_t)(sp + 1) * n_tiles / splits);
            const int64_t q = (int64_t)mt * GEMM_BM + m;
            const bool active = q < nq;
            const int64_t slot = ((int64_t)mt * splits + sp) * GEMM_BM + m;
            uint64_t* list = pool_items + slot * cap;
            int cnt = 0;
            float thr = active ? (thr0 ? thr0[q] : -INFINITY) : INFINITY;
            uint32_t thr_rec = 0;
            if (sample_mode) {
                // Threshold bootstrap: only the 8 best scores of this lane's tile(s) are needed (the
                // union of per-tile top-8 lists holds >= k items, and its k-th best is a valid lower
                // bound of the database's k-th best).  Sorted in registers, no lists, no trims.
                float top[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) top[j] = -INFINITY;
                for (int t = t0; t < t1; ++t, ++it) {
                    const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
                    mbar_wait(smem_u32(&bars->tfull[acc]), acc_phase);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + acc * GEMM_BN;
#pragma unroll 1
                    for (int c = 0; c < GEMM_BN / 32; ++c) {
                        uint32_t v[32];
                        tc_ld32(taddr + c * 32, v);
                        tc_ld_wait();
                        const int64_t lim = n_valid - ((int64_t)t * tile_stride * GEMM_BN + c * 32);
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            float s = __uint_as_float(v[i]);
                            if (s > top[7] && i < lim) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) { const float hi = fmaxf(top[j], s); s = fminf(top[j], s); top[j] = hi; }
                            }
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(smem_u32(&bars->tempty[acc]));
                }
                if (active) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) if (top[j] > -INFINITY) list[cnt++] = make_item(top[j], 0u);
                }
                pool_count[slot] = cnt;
                pool_thr[slot] = 0u;
                continue;
            }
            for (int t = t0; t < t1; ++t, ++it) {
                const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
                mbar_wait(smem_u32(&bars->tfull[acc]), acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + acc * GEMM_BN;
#pragma unroll 1
                for (int c = 0; c < GEMM_BN / 32; ++c) {
                    uint32_t v[32];
                    tc_ld32(taddr + c * 32, v);
                    tc_ld_wait();
                    const int64_t row_base = (int64_t)t * tile_stride * GEMM_BN + c * 32;
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
                        const int c_new = warp_prune(l, c_src, k_keep, true, 0.f, my_hist, T);
                        if (lane == src) { cnt = c_new; thr = fmaxf(thr, key_score(T)); thr_rec = max(thr_rec, T); }
                    }
                }
                tc_fence_before();
                mbar_arrive(smem_u32(&bars->tempty[acc]));
            }
            // end of job: keep only what can still matter -- everything within 2*eps below this
            // split's k-th best (the global k-th best is at least as large)
            const float band = active ? 2.f * eps[q] : 0.f;
            uint32_t need = __ballot_sync(0xffffffffu, cnt > k);
            while (need) {
                const int src = __ffs(need) - 1;
                need &= need - 1;
                uint64_t* l = reinterpret_cast<uint64_t*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(list), src));
                const int c_src = __shfl_sync(0xffffffffu, cnt, src);
                const float b_src = __shfl_sync(0xffffffffu, band, src);
                uint32_t T;
                const int c_new = warp_prune(l, c_src, k, false, b_src, my_hist, T);
                if (lane == src) cnt = c_new;
            }
            pool_count[slot] = cnt;
            pool_thr[slot] = thr_rec;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- host side -------------------------------------------------------------------------------------
GemmPlan plan_gemm(int64_t nq, int64_t n_pad, int k, int num_sms, int forced_splits) {
    GemmPlan p{};
    p.m_tiles = (int)((nq + GEMM_BM - 1) / GEMM_BM);
    p.n_tiles = (int)(n_pad / GEMM_BN);
    int best = 1;
    if (forced_splits > 0) best = forced_splits;
    else {
        // pick the split count that fills whole waves of `num_sms` jobs while keeping jobs long
        double best_score = -1.0;
        const int s_max = (int)((p.n_tiles < 4096) ? p.n_tiles : 4096);
        for (int s = 1; s <= s_max; ++s) {
            const int64_t jobs = (int64_t)p.m_tiles * s;
            const int64_t rounds = (jobs + num_sms - 1) / num_sms;
            // tiles per CTA on the critical path
            const double tiles_per_job = (double)((p.n_tiles + s - 1) / s);
            const double crit = (double)rounds * tiles_per_job;
            const double ideal = (double)p.m_tiles * p.n_tiles / num_sms;
            const double score = ideal / crit - 0.0005 * s;      // prefer fewer, longer jobs on ties
            if (score > best_score) { best_score = score; best = s; }
            if (jobs > (int64_t)num_sms * 64) break;
        }
    }
    if (best > p.n_tiles) best = p.n_tiles > 0 ? p.n_tiles : 1;
    p.splits = best;
    int keep = k + (k / 2 > 156 ? k / 2 : 156);
    p.k_keep = (keep + 31) & ~31;
    p.cap = 2 * p.k_keep;
    const int64_t jobs = (int64_t)p.m_tiles * p.splits;
    p.grid = (int)(jobs < num_sms ? jobs : num_sms);
    p.tile_stride = 1;
    p.sample_mode = 0;
    return p;
}

// Threshold bootstrap: S database tiles spread evenly over the database (stride), one tile per job.
// The k-th best score of that sample is a valid lower bound of the k-th best of the whole database.
GemmPlan plan_gemm_sample(const GemmPlan& main_plan, int num_sms) {
    GemmPlan p = main_plan;
    const int all_tiles = main_plan.n_tiles;
    int s = (num_sms / 2) / (main_plan.m_tiles > 0 ? main_plan.m_tiles : 1);
    if (s < 16) s = 16;
    if (s > all_tiles) s = all_tiles;
    p.n_tiles = s;
    p.tile_stride = all_tiles / s;
    p.splits = s;                                   // one tile per job
    p.sample_mode = 1;                              // register top-8 per (query, tile)
    const int64_t jobs = (int64_t)p.m_tiles * p.splits;
    p.grid = (int)(jobs < num_sms ? jobs : num_sms);
    return p;
}

cudaError_t launch_gemm_topk(const CUtensorMap& tmap_q, const CUtensorMap& tmap_db, const GemmPlan& plan,
                             int64_t nq, int64_t n_valid, int d_pad, int k, const float* eps,
                             uint64_t* pool_items, int* pool_count, uint32_t* pool_thr, int a_row0,
                             const float* thr0, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(gemm_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    if (e != cudaSuccess) return e;
    gemm_topk_kernel<<<plan.grid, GEMM_THREADS, GEMM_SMEM, st>>>(tmap_q, tmap_db, plan.m_tiles, plan.n_tiles, plan.tile_stride, plan.splits,
                                                                  d_pad / GEMM_BK, a_row0, nq, n_valid, k, plan.k_keep, plan.cap, plan.sample_mode, eps, thr0,
                                                                  pool_items, pool_count, pool_thr);
    return cudaGetLastError();
}

}  // namespace xs
