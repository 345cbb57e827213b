// Stage 2: candidate pools -> exact top-k.
//
// One CTA per query: (1) radix-select the k-th best coarse score over all partial lists,
// (2) gather every candidate inside the exactness band below it, (3) rescore those rows exactly
// (fp32 operands, fp64 accumulation in a fixed order, one rounding to fp32 -- the deterministic
// stand-in for OpenBLAS' sgemm at src/main_retrieve.py:175), (4) sort by (score desc, id asc),
// (5) emit the first k, plus a certificate bit when anything outside the band may have been lost.
#include <cstdio>
#include <cooperative_groups.h>
#include "common.cuh"
#include "select.cuh"
#include "internal.h"

namespace xs {

constexpr int FIN_THREADS = 512;
constexpr int FIN_MAX_LISTS = 4096;          // partial lists per query that the shared-memory gather supports
constexpr int FIN_SMEM_BUDGET = 200 * 1024;  // dynamic shared memory the finalise kernel may ask for
constexpr int FIN_FUSED_MAX_Q = 128;         // coarse-mode batches up to this size: ONE kernel (select + rescore + emit, several CTAs per query)
constexpr float MODEL_CHECK = 0.75f;         // a rescored candidate whose |coarse - exact| exceeds this fraction of eps voids the certificate

// ---- peer-exchange handshake words (system scope: written by kernels running on OTHER GPUs over NVLink) ----
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
// Spin until *p >= want.  A peer that never arrives is a protocol error: trap after ~20 s instead of hanging the GPU.
__device__ __forceinline__ void wait_word_sys(const uint32_t* p, uint32_t want) {
    if (ld_acquire_sys(p) >= want) return;
    const long long t0 = clock64();
    while (ld_acquire_sys(p) < want) {
        __nanosleep(64);
        if (clock64() - t0 > 40000000000ll) { printf("xs: peer exchange timed out (want %u, have %u)\n", want, ld_acquire_sys(p)); __trap(); }
    }
}

// Exact inner products of TWO database rows with the query row held in shared memory: fp32
// operands, fp64 accumulation, fixed order (lane-strided float4s, chunk by chunk), warp-reduced.
// All loads of a chunk are issued before any arithmetic so that 8 x 512 B are in flight per warp.
__device__ __forceinline__ void exact_dot2(const float* __restrict__ v0, const float* __restrict__ v1, bool has1,
                                           const float* qs, int d_pad, double& r0, double& r1) {
    const int lane = lane_id();
    double a0 = 0.0, a1 = 0.0;
    for (int base = 0; base < d_pad; base += 512) {
        float4 x[4], y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = base + j * 128 + lane * 4;
            const bool in = i < d_pad;
            x[j] = in ? ld_stream_f4(v0 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            y[j] = (in && has1) ? ld_stream_f4(v1 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("" ::: "memory");                  // keep every load of the chunk ahead of the arithmetic
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = base + j * 128 + lane * 4;
            if (i < d_pad) {
                const float4 w = *reinterpret_cast<const float4*>(qs + i);
                a0 = fma((double)x[j].x, (double)w.x, a0); a0 = fma((double)x[j].y, (double)w.y, a0);
                a0 = fma((double)x[j].z, (double)w.z, a0); a0 = fma((double)x[j].w, (double)w.w, a0);
                a1 = fma((double)y[j].x, (double)w.x, a1); a1 = fma((double)y[j].y, (double)w.y, a1);
                a1 = fma((double)y[j].z, (double)w.z, a1); a1 = fma((double)y[j].w, (double)w.w, a1);
            }
        }
    }
    r0 = warp_sum(a0);
    r1 = warp_sum(a1);
}

// One row, the query row read through the read-only path (L1-resident after the first warp): up to
// 16 x 512 B in flight per warp -- used where many CTAs share the rescoring of one query.
__device__ __forceinline__ double exact_dot1(const float* __restrict__ v, const float* __restrict__ qrow, int d_pad) {
    const int lane = lane_id();
    double a = 0.0;
    for (int base = 0; base < d_pad; base += 2048) {
        float4 x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int i = base + j * 128 + lane * 4;
            x[j] = (i < d_pad) ? ld_stream_f4(v + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("" ::: "memory");                  // keep every load of the chunk ahead of the arithmetic
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int i = base + j * 128 + lane * 4;
            if (i < d_pad) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(qrow + i));
                a = fma((double)x[j].x, (double)w.x, a); a = fma((double)x[j].y, (double)w.y, a);
                a = fma((double)x[j].z, (double)w.z, a); a = fma((double)x[j].w, (double)w.w, a);
            }
        }
    }
    return warp_sum(a);
}

// Descending bitonic sort of m (power of two) 64-bit items in shared memory.
__device__ void block_sort_desc(uint64_t* a, int m) {
    for (int size = 2; size <= m; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (m >> 1); t += blockDim.x) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                uint64_t x = a[lo], y = a[hi];
                if ((x < y) == desc) { a[lo] = y; a[hi] = x; }
            }
        }
    }
    __syncthreads();
}

// Conservative k-th-largest for COARSE scores (whole CTA of FIN_THREADS threads): one min/max pass
// and one pass into 4096 bins spread over [min, max]; returns the lower edge of the bin that holds
// the kk-th largest key -- never above it, at most one bin (typically 1-3 items) below.  The exactness
// band is cut from that edge, so a lower value only adds a few candidates.  Replaces four 8-bit radix
// passes with warp-match aggregation (28k cycles at 6.6k items) by ~4k cycles.
template <int NT, typename Each>
__device__ uint32_t block_kth_edge(Each each, uint32_t kk, uint32_t* bins, uint32_t* sh, uint32_t* wsum) {
    constexpr int NB = 4096, PER = NB / NT;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { sh[0] = 0xFFFFFFFFu; sh[1] = 0u; sh[2] = 0u; }
    for (int i = threadIdx.x; i < NB; i += NT) bins[i] = 0;
    __syncthreads();
    uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
    each([&](uint64_t it, bool valid) { if (valid) { const uint32_t key = item_key(it); kmin = min(kmin, key); kmax = max(kmax, key); } });
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if (lane == 0) { atomicMin(&sh[0], kmin); atomicMax(&sh[1], kmax); }
    __syncthreads();
    kmin = sh[0]; kmax = sh[1];
    const uint32_t range = kmax - kmin;
    const int shift = (range >> 12) ? (32 - __clz(range) - 12) : 0;
    each([&](uint64_t it, bool valid) { if (valid) atomicAdd(&bins[(item_key(it) - kmin) >> shift], 1u); });
    __syncthreads();
    uint32_t h[PER], mine = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) { h[j] = bins[threadIdx.x * PER + j]; mine += h[j]; }
    uint32_t suf = mine;                                // suffix over lanes >= lane
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_down_sync(0xffffffffu, suf, o); if (lane + o < 32) suf += t; }
    if (lane == 0) wsum[warp] = suf;
    __syncthreads();
    uint32_t above = 0;
    for (int w = warp + 1; w < NT / 32; ++w) above += wsum[w];
    const uint32_t incl = above + suf, excl = incl - mine;
    if (incl >= kk && excl < kk) {                      // exactly one thread
        uint32_t run = excl;
#pragma unroll
        for (int j = PER - 1; j >= 0; --j) {
            run += h[j];
            if (run >= kk) { sh[2] = (uint32_t)(threadIdx.x * PER + j); break; }
        }
    }
    __syncthreads();
    return kmin + (sh[2] << shift);
}

// Exclusive prefix sum of offs[0..P) in place (offs[P] = total), whole CTA.  scratch: 33 words.
__device__ void block_exclusive_scan(int* offs, int P, int* scratch) {
    const int nth = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = (P + nth - 1) / nth;
    const int lo = min(tid * chunk, P), hi = min(lo + chunk, P);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += offs[i];
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = (lane < (nth >> 5)) ? scratch[lane] : 0, winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
        scratch[lane] = winc - w;
        if (lane == 31) scratch[32] = winc;
    }
    __syncthreads();
    int run = scratch[warp] + inc - sum;
    for (int i = lo; i < hi; ++i) { int c = offs[i]; offs[i] = run; run += c; }
    if (tid == 0) offs[P] = scratch[32];
    __syncthreads();
}

// Copies every pooled item of query q into shared memory (flat).  offs[] holds the exclusive scan of
// the list counts.  Lists are short (tens of items), so the copy is laid out as a flat loop over
// (list, position < 32) pairs -- every load independent of every other -- plus a remainder loop for
// the rare longer list.
__device__ void gather_pool(const uint64_t* __restrict__ pool_items, int64_t q, int P, int cap,
                            const int* offs, int total, uint64_t* items) {
    // eight independent loads per thread in flight before the first store: a load-store pair per iteration would
    // serialise on the L2 round trip (it was 12 % of the kernel's samples)
    const int flat = P * 32, nt = blockDim.x;
    for (int base = threadIdx.x; base < flat; base += nt * 8) {
        uint64_t v[8];
        int dst[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = base + u * nt;
            dst[u] = -1;
            if (i < flat) {
                const int p = i >> 5, j = i & 31;
                const int o = offs[p], c = offs[p + 1] - o;
                if (j < c) { dst[u] = o + j; v[u] = pool_items[pool_slot(q, p, P) * cap + j]; }
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) if (dst[u] >= 0) items[dst[u]] = v[u];
    }
    for (int p = threadIdx.x >> 5; p < P; p += blockDim.x >> 5) {
        const int o = offs[p], c = offs[p + 1] - o;
        for (int j = 32 + (threadIdx.x & 31); j < c; j += 32) items[o + j] = pool_items[pool_slot(q, p, P) * cap + j];
    }
    (void)total;
    __syncthreads();
}

// Last step of every finalise form, whole CTA: `items` (shared memory) holds the query's candidates sorted by
// (exact score desc, id asc).  Writes the first k to the caller's arrays and / or straight into every rank's mailbox
// (peer exchange: ids, scores and the certificate word, then the query's arrival flag per rank).
__device__ void emit_topk(const FinaliseArgs& a, int64_t q, const uint64_t* items, int ncand, uint32_t selfkey, int status) {
    const int kout = min(a.k, ncand);
    if (kout < a.k) status |= ST_UNCERTIFIED;           // a short list (NaN scores, a bad threshold) is never handed out as certified
    const PushTarget& p = a.push;
    if (p.world > 0) {
        // the receivers must have merged this slot's previous epoch before it is overwritten
        if ((int)threadIdx.x < p.world && p.epoch > 1) wait_word_sys(p.my_acks + threadIdx.x, p.epoch - 1);
        __syncthreads();
    }
    for (int r = threadIdx.x; r < a.k; r += blockDim.x) {
        int64_t id = -1;
        float sc = -INFINITY;
        if (r < kout) {
            const uint64_t it = items[r];
            uint32_t key = item_key(it);
            if (a.self_base >= 0 && key == 0xFFFFFFFFu) key = selfkey;
            id = (int64_t)item_row(it) + a.id_offset;
            sc = key_score(key);
        }
        if (a.out_idx) a.out_idx[q * a.out_pitch + r] = id;
        if (a.out_score) a.out_score[q * a.out_pitch + r] = sc;
        for (int g = 0; g < p.world; ++g) {
            p.ids[g][q * a.k + r] = id;
            p.scores[g][q * a.k + r] = sc;
        }
    }
    if (threadIdx.x == 0) {
        if (a.status) a.status[q] = status;
        for (int g = 0; g < p.world; ++g) p.status[g][q] = status;
    }
    if (p.world > 0) {
        // the barrier orders every thread's stores before the releasing threads' system-scope fence (cumulative), so one
        // fence per destination is enough -- a __threadfence_system() in all 256 threads first cost another NVLink round trip
        __syncthreads();
        if ((int)threadIdx.x < p.world) st_release_sys(p.flags[threadIdx.x] + q, p.epoch);
    }
}

// SPLIT = false: the whole of stage 2 in one CTA per query (large batches, exact mode).
// SPLIT = true : stops after the candidate gather and hands the list to rescore / emit kernels that use
//                every SM (small batches, where one SM per query would be latency-bound on ~120 row reads).
template <bool SPLIT>
__global__ void __launch_bounds__(FIN_THREADS, 1)
finalise_kernel(FinaliseArgs a, int cand_max, int item_cap) {
    extern __shared__ uint64_t fin_smem[];              // [cand_max] candidates | [item_cap] gathered items | [P+1] offsets
    uint64_t* cand = fin_smem;
    uint64_t* items = fin_smem + cand_max;
    int* offs = reinterpret_cast<int*>(items + item_cap);
    float* qs = reinterpret_cast<float*>(offs + ((a.P <= FIN_MAX_LISTS) ? ((a.P + 4) & ~3) : 4));   // [d_pad] query row
    uint32_t* bins = reinterpret_cast<uint32_t*>(qs + a.d_pad);                                     // [4096] (coarse mode)
    __shared__ uint32_t hist[256];
    __shared__ uint32_t edge_sh[4];
    __shared__ uint32_t edge_wsum[FIN_THREADS / 32];
    __shared__ uint32_t misc[2];
    __shared__ int scan_scratch[33];
    __shared__ uint32_t sh_ncand, sh_flag, sh_selfkey;
    pdl_wait();
    const int64_t q = blockIdx.x;
    const int lane = lane_id(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const bool staged = a.P <= FIN_MAX_LISTS;           // offsets fit -> try the shared-memory gather

    if (threadIdx.x == 0) { sh_ncand = 0; sh_flag = 0; sh_selfkey = 0; }
    if (!SPLIT)
        for (int i = threadIdx.x * 4; i < a.d_pad; i += blockDim.x * 4)
            *reinterpret_cast<float4*>(qs + i) = *reinterpret_cast<const float4*>(a.q32 + q * a.d_pad + i);
    __syncthreads();
    uint32_t total = 0;
    {   // list sizes, and the largest "something above this key was dropped upstream" mark
        uint32_t thr = 0, t = 0;
        for (int p = threadIdx.x; p < a.P; p += blockDim.x) {
            const int64_t slot = pool_slot(q, p, a.P);
            const int c = a.pool_count[slot];
            if (staged) offs[p] = c; else t += (uint32_t)c;
            thr = max(thr, a.pool_thr[slot]);
        }
        if (thr) atomicMax(&sh_flag, thr);
        if (!staged && t) atomicAdd(&sh_ncand, t);
        __syncthreads();
        if (staged) { block_exclusive_scan(offs, a.P, scan_scratch); total = (uint32_t)offs[a.P]; }
        else { total = sh_ncand; __syncthreads(); if (threadIdx.x == 0) sh_ncand = 0; __syncthreads(); }
    }
    const uint32_t max_thr = sh_flag;
    const bool in_smem = staged && total <= (uint32_t)item_cap;
    if (in_smem) gather_pool(a.pool_items, q, a.P, a.cap, offs, (int)total, items);

    auto each = [&](auto fn) {
        if (in_smem) {
            const int n_up = ((int)total + (int)blockDim.x - 1) / (int)blockDim.x * (int)blockDim.x;
            for (int i = threadIdx.x; i < n_up; i += blockDim.x) {
                bool valid = i < (int)total;
                fn(valid ? items[i] : 0ull, valid);
            }
        } else {
            for (int p = warp; p < a.P; p += nwarps) {
                const int64_t slot = pool_slot(q, p, a.P);
                const int cnt = a.pool_count[slot];
                const uint64_t* lst = a.pool_items + slot * a.cap;
                for (int b = 0; b < cnt; b += 32) {
                    int i = b + lane;
                    bool valid = i < cnt;
                    fn(valid ? lst[i] : 0ull, valid);
                }
            }
        }
    };

    const uint32_t kk = min((uint32_t)a.k, total);
    uint64_t cut = 0;
    uint32_t cut_key = 0;
    if (total > kk) {
        if (a.exact) {
            const uint64_t T = block_kth_largest(each, kk, 8, hist, misc);
            cut = T; cut_key = (uint32_t)(T >> 32);
        } else {
            const uint32_t edge = block_kth_edge<FIN_THREADS>(each, kk, bins, edge_sh, edge_wsum);
            cut_key = score_key(key_score(edge) - 2.f * a.eps[q]);
            cut = (uint64_t)cut_key << 32;
        }
    }
    __syncthreads();
    each([&](uint64_t it, bool valid) {
        bool take = valid && it >= cut;
        uint32_t m = __ballot_sync(0xffffffffu, take);
        uint32_t base = 0;
        if (lane == 0 && m) base = atomicAdd(&sh_ncand, (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        uint32_t pos = base + __popc(m & lanemask_lt());
        if (take && pos < (uint32_t)cand_max) cand[pos] = it;
    });
    __syncthreads();
    const uint32_t found = sh_ncand;
    const int ncand = (int)min(found, (uint32_t)cand_max);
    // certificate: nothing that could belong to the exact top-k was dropped upstream
    bool uncertified = false;
    if (!a.exact) uncertified = (found > (uint32_t)cand_max) || (max_thr != 0 && max_thr >= cut_key);
    else          uncertified = (max_thr != 0);

    if (SPLIT) {      // hand over to finalise_rescore_kernel / finalise_emit_kernel
        for (int c = threadIdx.x; c < ncand; c += blockDim.x) a.w_cand[q * cand_max + c] = cand[c];
        if (threadIdx.x == 0) { a.w_ncand[q] = ncand; a.w_flag[q] = uncertified ? ST_UNCERTIFIED : 0; }
        return;
    }
    if (!a.exact) {   // exact rescoring, one warp per PAIR of candidates
        for (int c = warp * 2; c < ncand; c += nwarps * 2) {
            const bool has1 = c + 1 < ncand;
            const uint32_t row0 = item_row(cand[c]), row1 = has1 ? item_row(cand[c + 1]) : row0;
            double s0, s1;
            exact_dot2(a.db32 + (int64_t)row0 * a.d_pad, a.db32 + (int64_t)row1 * a.d_pad, has1, qs, a.d_pad, s0, s1);
            if (lane == 0) {
                cand[c] = make_item((float)s0, row0);
                if (has1) cand[c + 1] = make_item((float)s1, row1);
            }
        }
    }
    __syncthreads();
    if (a.self_base >= 0) {   // self-kNN: the query's own row ranks first whatever the ties
        const uint32_t self_row = (uint32_t)(a.self_base + q);
        for (int c = threadIdx.x; c < ncand; c += blockDim.x)
            if (item_row(cand[c]) == self_row)
                cand[c] = (0xFFFFFFFFull << 32) | (uint64_t)(0xFFFFFFFFu - self_row);
        if (warp == 0) {      // its exact score <v, v>, recomputed here (the pooled one may be a +inf boost)
            const float* vrow = a.db32 + (int64_t)self_row * a.d_pad;
            double s0, s1;
            exact_dot2(vrow, vrow, false, qs, a.d_pad, s0, s1);
            if (lane == 0) sh_selfkey = score_key((float)s0);
        }
    }
    int m = 1;
    while (m < ncand) m <<= 1;
    if (m < 2) m = 2;
    for (int c = ncand + threadIdx.x; c < m; c += blockDim.x) cand[c] = 0ull;
    block_sort_desc(cand, m);

    emit_topk(a, q, cand, ncand, sh_selfkey, uncertified ? ST_UNCERTIFIED : 0);
    if (threadIdx.x == 0 && a.n_cand) atomicAdd(a.n_cand, ncand);
}

// ---- split pipeline: rescore on every SM, then sort + emit ---------------------------------------------
constexpr int RS_WARPS = 8;

// grid (nq, rs_split): warp w of CTA (q, j) rescores candidates j*RS_WARPS + w, + rs_split*RS_WARPS, ...
// The row is staged with cp.async (16 B per lane and step, all steps issued back to back) so the whole
// 8 KB row is in flight at once -- with register loads ptxas interleaves each load with its use and a
// warp never has more than ~3 outstanding.  24 resident warps per SM -> ~190 KB in flight per SM.
// The LAST of a query's rs_split CTAs to finish (atomic ticket) sorts the rescored candidates, applies
// the self-first rule and emits the first k -- no separate launch.
__global__ void __launch_bounds__(RS_WARPS * 32)
finalise_rescore_emit_kernel(FinaliseArgs a, int cand_max) {
    extern __shared__ float rs_rows[];                  // [RS_WARPS][d_pad] row staging; reused as [cand_max] items by the emitter
    __shared__ uint32_t sh_selfkey, sh_last;
    pdl_wait();
    const int64_t q = blockIdx.x;
    const int ncand = a.w_ncand[q];
    uint64_t* cand = a.w_cand + q * cand_max;
    const float* qrow = a.q32 + q * a.d_pad;
    const int warp = threadIdx.x >> 5, lane = lane_id();
    float* buf = rs_rows + (size_t)warp * a.d_pad;
    const uint32_t sbuf = (uint32_t)__cvta_generic_to_shared(buf);
    const int rs_split = gridDim.y;                     // CTAs per query: 16 for small batches (latency), 4 for large ones (launch count)
    for (int c = blockIdx.y * RS_WARPS + warp; c < ncand; c += rs_split * RS_WARPS) {
        const uint32_t row = item_row(cand[c]);
        const float* v = a.db32 + (int64_t)row * a.d_pad;
        for (int i = lane * 4; i < a.d_pad; i += 128)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sbuf + (uint32_t)i * 4u), "l"(v + i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        double acc = 0.0;
        for (int i = lane * 4; i < a.d_pad; i += 128) {
            const float4 x = *reinterpret_cast<const float4*>(buf + i);
            const float4 w = __ldg(reinterpret_cast<const float4*>(qrow + i));
            acc = fma((double)x.x, (double)w.x, acc); acc = fma((double)x.y, (double)w.y, acc);
            acc = fma((double)x.z, (double)w.z, acc); acc = fma((double)x.w, (double)w.w, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            // model check: the band assumes |coarse - exact| stays well inside eps; a candidate that says otherwise voids it
            if (fabsf(key_score(item_key(cand[c])) - (float)acc) > MODEL_CHECK * a.eps[q]) atomicOr(&a.w_flag[q], ST_UNCERTIFIED);
            cand[c] = make_item((float)acc, row);
        }
        __syncwarp();
    }
    // ---- completion ticket ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(&a.w_ticket[q], 1);
        sh_last = (t == rs_split - 1) ? 1u : 0u;
        if (sh_last) a.w_ticket[q] = 0;                 // ready for the next call
        sh_selfkey = 0;
    }
    __syncthreads();
    if (!sh_last) return;
    __threadfence();
    // ---- sort + emit (this CTA only) ----
    uint64_t* items = reinterpret_cast<uint64_t*>(rs_rows);
    int m = 1;
    while (m < ncand) m <<= 1;
    if (m < 2) m = 2;
    for (int c = threadIdx.x; c < m; c += blockDim.x) items[c] = (c < ncand) ? __ldcg(cand + c) : 0ull;
    __syncthreads();
    if (a.self_base >= 0) {
        const uint32_t self_row = (uint32_t)(a.self_base + q);
        for (int c = threadIdx.x; c < ncand; c += blockDim.x)
            if (item_row(items[c]) == self_row)
                items[c] = (0xFFFFFFFFull << 32) | (uint64_t)(0xFFFFFFFFu - self_row);
        if (threadIdx.x < 32) {
            const float* vrow = a.db32 + (int64_t)self_row * a.d_pad;
            const double s = exact_dot1(vrow, qrow, a.d_pad);
            if (threadIdx.x == 0) sh_selfkey = score_key((float)s);
        }
    }
    block_sort_desc(items, m);
    const int flags = __ldcg(&a.w_flag[q]);
    emit_topk(a, q, items, ncand, sh_selfkey, flags);
    if (threadIdx.x == 0) {
        a.w_ncand[q] = 0; a.w_flag[q] = 0;              // the counters are shared with the fused form: leave them zero
        if (a.n_cand) atomicAdd(a.n_cand, ncand);
    }
}

// ---- fused form for small batches: select + rescore + emit in ONE launch, one thread-block CLUSTER per query ------------
// grid (S, nq), cluster (S, 1, 1).  The S CTAs of a query split its partial lists between them and meet through
// distributed shared memory:
//   1. every CTA gathers ITS lists (p = rank, rank + S, ...) into shared memory, finds their min / max key;
//   2. cluster barrier; the global min / max come from the S exchange blocks; every CTA fills a 1024-bin histogram of its
//      items over that range; cluster barrier; every thread sums its 4 bins over the S CTAs (DSMEM reads) and the usual
//      suffix scan gives the conservative k-th best and the cut of the band -- the same value in all CTAs;
//   3. every CTA collects its items inside the band; cluster barrier; the band members are dealt evenly: warp w of CTA r
//      takes global positions r * found / S + w, + 8, ..., fetches the item from whichever CTA holds it (DSMEM), rescores it
//      exactly (the 8 KB fp32 row staged with cp.async) and stores the result into CTA 0's list (DSMEM);
//   4. cluster barrier; CTA 0 ranks the <= cand_max results by counting, emits, and -- with the peer exchange on -- stores
//      the k results straight into every rank's mailbox.
// No redundant work, no global atomics or tickets between the phases, S times the shared-memory capacity for the lists.
// It is what bounds the per-step fixed cost and with it strong scaling.
constexpr int FF_THREADS = 256, FF_WARPS = FF_THREADS / 32;
constexpr int FF_ITEMS = 2048;               // pooled items one CTA of the cluster can hold
constexpr int FF_MAX_CLUSTER = 8;
// Histogram of the cluster form: 1024 bins over [min, max] of the query's pooled keys.  The pooled scores span a few
// hundredths, so a bin is ~5e-5 wide -- far inside the band -- and a CTA's histogram is 4 KB: distributed shared memory
// moves ~20 bytes per cycle and SM, and every CTA reads every other CTA's histogram (4096 bins cost 30 us here).
constexpr int FF_BIN_BITS = 10, FF_BINS = 1 << FF_BIN_BITS;
__device__ __forceinline__ unsigned long long ff_timer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define FF_STAMP(i) do { if (a.trace && threadIdx.x == 0) a.trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 10 + (i)] = ff_timer(); } while (0)

struct FfExchange {              // one per CTA, read by the whole cluster
    uint32_t kmin, kmax, total, thr, ncand, overflow, bad, pad;
};

__global__ void __launch_bounds__(FF_THREADS)
finalise_cluster_kernel(FinaliseArgs a, int cand_max, int item_cap, int rs_warps) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ uint64_t ff_smem[];
    uint64_t* cands = ff_smem;                          // [cand_max] this CTA's band members
    uint64_t* result = ff_smem + cand_max;              // [cand_max] (CTA 0's copy collects the cluster's rescored candidates)
    uint64_t* items = ff_smem + 2 * cand_max;           // phase 1: [item_cap] gathered pool items | FF_BINS bins
    uint32_t* bins = reinterpret_cast<uint32_t*>(items + item_cap);
    float* rows = reinterpret_cast<float*>(ff_smem + 2 * cand_max);      // phase 3 (aliases phase 1): [rs_warps][d_pad] row staging
    __shared__ FfExchange xch;
    __shared__ int offs[FF_THREADS + 1];
    __shared__ int scan_scratch[33];
    __shared__ uint32_t wsum[FF_WARPS];
    __shared__ uint32_t sh_min, sh_max, sh_pick, sh_n, sh_selfkey;
    pdl_wait();
    FF_STAMP(0);
    const int64_t q = blockIdx.y;
    const int S = (int)gridDim.x, r = (int)blockIdx.x;
    const int lane = lane_id(), warp = threadIdx.x >> 5, tid = threadIdx.x;

    // ---- 1. my lists: p = r + t * S ------------------------------------------------------------------------------------
    if (tid == 0) { sh_min = 0xFFFFFFFFu; sh_max = 0u; sh_pick = 0u; sh_n = 0u; sh_selfkey = 0u; }
    const int n_lists = (a.P - r + S - 1) / S;          // <= FF_THREADS guaranteed by the launcher
    uint32_t thr = 0;
    {
        int c = 0;
        if (tid < n_lists) {
            const int64_t slot = pool_slot(q, r + tid * S, a.P);
            c = a.pool_count[slot];
            thr = a.pool_thr[slot];
        }
        offs[tid] = c;
    }
    __syncthreads();
    block_exclusive_scan(offs, FF_THREADS, scan_scratch);            // offs[FF_THREADS] = total
    const int total_all = offs[FF_THREADS];
    const int total_r = min(total_all, item_cap);                    // what does not fit voids the certificate below
    {   // gather: warp per list round-robin would serialise on short lists; flat (list, position < 32) slots, 8 loads in flight
        const int flat = n_lists * 32;
        for (int base = tid; base < flat; base += FF_THREADS * 8) {
            uint64_t v[8];
            int dst[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = base + u * FF_THREADS;
                dst[u] = -1;
                if (i < flat) {
                    const int l = i >> 5, jj = i & 31;
                    const int o = offs[l], c = offs[l + 1] - o;
                    if (jj < c && o + jj < item_cap) { dst[u] = o + jj; v[u] = a.pool_items[pool_slot(q, r + l * S, a.P) * a.cap + jj]; }
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) if (dst[u] >= 0) items[dst[u]] = v[u];
        }
        for (int l = warp; l < n_lists; l += FF_WARPS) {             // the rare list longer than 32
            const int o = offs[l], c = offs[l + 1] - o;
            for (int jj = 32 + lane; jj < c; jj += 32) if (o + jj < item_cap) items[o + jj] = a.pool_items[pool_slot(q, r + l * S, a.P) * a.cap + jj];
        }
    }
    for (int i = tid; i < FF_BINS; i += FF_THREADS) bins[i] = 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) thr = max(thr, __shfl_xor_sync(0xffffffffu, thr, o));
    if (lane == 0 && thr) atomicMax(&sh_pick, thr);                 // (sh_pick doubles as the max of pool_thr until step 2)
    __syncthreads();
    FF_STAMP(1);
    {
        uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
        for (int i = tid; i < total_r; i += FF_THREADS) { const uint32_t key = item_key(items[i]); kmin = min(kmin, key); kmax = max(kmax, key); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
            kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
        }
        if (lane == 0) { atomicMin(&sh_min, kmin); atomicMax(&sh_max, kmax); }
    }
    __syncthreads();
    if (tid == 0) {
        xch.kmin = sh_min; xch.kmax = sh_max; xch.total = (uint32_t)total_r; xch.thr = sh_pick;
        xch.overflow = total_all > item_cap ? 1u : 0u; xch.ncand = 0; xch.bad = 0;
        sh_pick = 0u;
    }
    cluster.sync();                                                  // ---- every CTA's exchange block is readable ----
    FF_STAMP(2);
    uint32_t kmin = 0xFFFFFFFFu, kmax = 0u, total = 0, max_thr = 0, overflow = 0;
    for (int g = 0; g < S; ++g) {
        const FfExchange* x = cluster.map_shared_rank(&xch, g);
        kmin = min(kmin, x->kmin); kmax = max(kmax, x->kmax); total += x->total; max_thr = max(max_thr, x->thr); overflow |= x->overflow;
    }
    // ---- 2. conservative k-th best of the query's pooled coarse scores --------------------------------------------------
    const uint32_t kk = min((uint32_t)a.k, total);
    uint64_t cut = 0;
    uint32_t cut_key = 0;
    if (total > kk) {                                                // (uniform across the cluster)
        const uint32_t range = kmax - kmin;
        const int shift = (range >> FF_BIN_BITS) ? (32 - __clz(range) - FF_BIN_BITS) : 0;
        for (int i = tid; i < total_r; i += FF_THREADS) atomicAdd(&bins[(item_key(items[i]) - kmin) >> shift], 1u);
        cluster.sync();                                              // ---- every CTA's histogram is complete ----
        constexpr int PER = FF_BINS / FF_THREADS;                    // 4 bins per thread (one 16-byte DSMEM read per CTA), summed over the cluster
        uint32_t h[PER];
#pragma unroll
        for (int jj = 0; jj < PER; ++jj) h[jj] = 0;
        for (int g = 0; g < S; ++g) {
            const uint4* rb = reinterpret_cast<const uint4*>(cluster.map_shared_rank(bins, g) + tid * PER);
#pragma unroll
            for (int jj = 0; jj < PER / 4; ++jj) {
                const uint4 w = rb[jj];
                h[4 * jj] += w.x; h[4 * jj + 1] += w.y; h[4 * jj + 2] += w.z; h[4 * jj + 3] += w.w;
            }
        }
        uint32_t minesum = 0;
#pragma unroll
        for (int jj = 0; jj < PER; ++jj) minesum += h[jj];
        uint32_t suf = minesum;                                      // suffix over lanes >= lane
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_down_sync(0xffffffffu, suf, o); if (lane + o < 32) suf += t; }
        if (lane == 0) wsum[warp] = suf;
        __syncthreads();
        uint32_t above = 0;
        for (int w = warp + 1; w < FF_WARPS; ++w) above += wsum[w];
        const uint32_t incl = above + suf, excl = incl - minesum;
        if (incl >= kk && excl < kk) {                               // exactly one thread
            uint32_t run = excl;
#pragma unroll
            for (int jj = PER - 1; jj >= 0; --jj) {
                run += h[jj];
                if (run >= kk) { sh_pick = (uint32_t)(tid * PER + jj); break; }
            }
        }
        __syncthreads();
        const uint32_t edge = kmin + (sh_pick << shift);
        cut_key = score_key(key_score(edge) - 2.f * a.eps[q]);
        cut = (uint64_t)cut_key << 32;
    }
    FF_STAMP(3);
    // ---- 3. my band members ------------------------------------------------------------------------------------------------
    {
        const int n_up = (total_r + FF_THREADS - 1) / FF_THREADS * FF_THREADS;
        for (int i = tid; i < n_up; i += FF_THREADS) {
            const uint64_t it = (i < total_r) ? items[i] : 0ull;
            const bool take = (i < total_r) && it >= cut;
            const uint32_t m = __ballot_sync(0xffffffffu, take);
            uint32_t base = 0;
            if (lane == 0 && m) base = atomicAdd(&sh_n, (uint32_t)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            const uint32_t pos = base + __popc(m & lanemask_lt());
            if (take && pos < (uint32_t)cand_max) cands[pos] = it;
        }
    }
    __syncthreads();
    if (tid == 0) xch.ncand = sh_n;                                  // may exceed cand_max: counted, not stored
    cluster.sync();                                                  // ---- band members + counts readable; histograms no longer needed ----
    FF_STAMP(4);
    uint32_t off[FF_MAX_CLUSTER + 1];
    off[0] = 0;
    bool lost = false;
    for (int g = 0; g < S; ++g) {
        const uint32_t c = cluster.map_shared_rank(&xch, g)->ncand;
        lost |= c > (uint32_t)cand_max;
        off[g + 1] = off[g] + min(c, (uint32_t)cand_max);
    }
    const uint32_t found = off[S];
    const int ncand = (int)min(found, (uint32_t)cand_max);
    // certificate: nothing that could belong to the exact top-k was dropped upstream or does not fit
    const bool uncertified = lost || found > (uint32_t)cand_max || overflow != 0 || (max_thr != 0 && max_thr >= cut_key);
    // exact rescoring of an even share: global positions lo + warp, + 8, ... of the concatenated band lists
    const int lo = (int)((int64_t)r * ncand / S), hi = (int)((int64_t)(r + 1) * ncand / S);
    const float* qrow = a.q32 + q * a.d_pad;
    const float band_check = MODEL_CHECK * a.eps[q];
    float* buf = rows + (size_t)warp * a.d_pad;
    const uint32_t sbuf = (uint32_t)__cvta_generic_to_shared(buf);
    uint64_t* result0 = cluster.map_shared_rank(result, 0);
    bool bad = false;
    for (int c = lo + warp; c < hi && warp < rs_warps; c += rs_warps) {
        int g = 0;
        while (g + 1 < S && (uint32_t)c >= off[g + 1]) ++g;
        const uint64_t it = cluster.map_shared_rank(cands, g)[c - off[g]];
        const uint32_t row = item_row(it);
        const float* v = a.db32 + (int64_t)row * a.d_pad;
        for (int i = lane * 4; i < a.d_pad; i += 128)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sbuf + (uint32_t)i * 4u), "l"(v + i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        double acc = 0.0;
#pragma unroll 4
        for (int i = lane * 4; i < a.d_pad; i += 128) {
            const float4 x = *reinterpret_cast<const float4*>(buf + i);
            const float4 w = __ldg(reinterpret_cast<const float4*>(qrow + i));
            acc = fma((double)x.x, (double)w.x, acc); acc = fma((double)x.y, (double)w.y, acc);
            acc = fma((double)x.z, (double)w.z, acc); acc = fma((double)x.w, (double)w.w, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            const float exact = (float)acc;
            bad |= fabsf(key_score(item_key(it)) - exact) > band_check;      // model check
            result0[c] = make_item(exact, row);
        }
        __syncwarp();
    }
    if (lane == 0 && bad) cluster.map_shared_rank(&xch, 0)->bad = 1u;      // (CTA 0's block: the others have exited by the time it is read)
    FF_STAMP(5);
    cluster.sync();                                                  // ---- CTA 0 holds every rescored candidate ----
    FF_STAMP(6);
    if (r != 0) return;
    // ---- 4. rank by counting (items are distinct 64-bit words), emit ---------------------------------------------------------
    const int flags = (uncertified || xch.bad) ? ST_UNCERTIFIED : 0;
    uint64_t* sorted = cands;                                        // CTA 0's own band list is no longer needed ... by CTA 0
    if (a.self_base >= 0) {
        const uint32_t self_row = (uint32_t)(a.self_base + q);
        for (int c = tid; c < ncand; c += FF_THREADS)
            if (item_row(result[c]) == self_row)
                result[c] = (0xFFFFFFFFull << 32) | (uint64_t)(0xFFFFFFFFu - self_row);
        if (tid < 32) {
            const double sd = exact_dot1(a.db32 + (int64_t)self_row * a.d_pad, qrow, a.d_pad);
            if (tid == 0) sh_selfkey = score_key((float)sd);
        }
        __syncthreads();
    }
    for (int c = tid; c < ncand; c += FF_THREADS) {
        const uint64_t me = result[c];
        int rank = 0;
        for (int o = 0; o < ncand; ++o) rank += result[o] > me ? 1 : 0;
        sorted[rank] = me;
    }
    __syncthreads();
    FF_STAMP(7);
    emit_topk(a, q, sorted, ncand, sh_selfkey, flags);
    FF_STAMP(8);
    if (tid == 0 && a.n_cand) atomicAdd(a.n_cand, ncand);
    if (!a.merge.on) return;
    // ---- 5. receiving end of the exchange, same kernel: this query's lists from every rank -> the final k ----------------
    // (the other clusters keep working; the only thing waited for here is another GPU)
    const MergeTarget& mg = a.merge;
    const int world = a.push.world, k = a.k;
    const int64_t qg = mg.q0 + q;                                    // query index within the whole call
    if (tid < world) wait_word_sys(mg.flags + (int64_t)tid * mg.flag_stride + qg, a.push.epoch);
    __syncthreads();
    uint64_t* mitems = reinterpret_cast<uint64_t*>(rows);            // [world * k] (the staging region is free now)
    uint64_t* mrank = mitems + world * k;
    const int mtotal = world * k;
    for (int i = tid; i < mtotal; i += FF_THREADS) {
        const int g = i / k, rr = i - g * k;
        const char* part = mg.base + (int64_t)g * mg.part_bytes;
        const int64_t id = reinterpret_cast<const int64_t*>(part)[qg * k + rr];
        const float sc = reinterpret_cast<const float*>(part + mg.nq_total * k * 8)[qg * k + rr];
        // position i orders equal scores by ascending id: parts own increasing id ranges and are sorted inside
        mitems[i] = id >= 0 ? make_item(sc, (uint32_t)i) : 0ull;
    }
    __syncthreads();
    for (int i = tid; i < mtotal; i += FF_THREADS) {
        const uint64_t me = mitems[i];
        if (me == 0ull) continue;
        int rank = 0;
        for (int o = 0; o < mtotal; ++o) rank += mitems[o] > me ? 1 : 0;
        if (rank < k) mrank[rank] = me;
    }
    __syncthreads();
    int valid = 0;
    for (int i = tid; i < mtotal; i += FF_THREADS) valid += mitems[i] != 0ull ? 1 : 0;
    __shared__ int sh_valid;
    if (tid == 0) sh_valid = 0;
    __syncthreads();
    if (valid) atomicAdd(&sh_valid, valid);
    __syncthreads();
    const int nvalid = min(sh_valid, k);
    for (int rr = tid; rr < k; rr += FF_THREADS) {
        int64_t id = -1;
        float sc = -INFINITY;
        if (rr < nvalid) {
            const int i = (int)item_row(mrank[rr]);
            const int g = i / k, r2 = i - g * k;
            const char* part = mg.base + (int64_t)g * mg.part_bytes;
            id = reinterpret_cast<const int64_t*>(part)[qg * k + r2];
            sc = reinterpret_cast<const float*>(part + mg.nq_total * k * 8)[qg * k + r2];
        }
        mg.out_idx[qg * k + rr] = id;
        if (mg.out_score) mg.out_score[qg * k + rr] = sc;
    }
    if (tid == 0 && mg.out_status) {                                 // certified only if every shard certified its list
        int32_t stw = 0;
        for (int g = 0; g < world; ++g) stw |= reinterpret_cast<const int32_t*>(mg.base + (int64_t)g * mg.part_bytes + mg.nq_total * k * 12)[qg];
        mg.out_status[qg] = stw;
    }
    __syncthreads();                                                 // every read of the mailbox by this CTA has completed
    __shared__ int sh_lastq;
    if (tid == 0) {
        __threadfence();
        sh_lastq = (atomicAdd(mg.ticket, 1u) == (uint32_t)(mg.nq_total - 1)) ? 1 : 0;
    }
    __syncthreads();
    if (sh_lastq) {                                                  // the whole call has been merged here: the slot may be overwritten
        if (tid == 0) *mg.ticket = 0;
        if (tid < world) st_release_sys(mg.ack[tid], a.push.epoch);
    }
}

// Candidates per query the rescoring stage can hold: the band of the statistical certificate is a few dozen rows wide
// on top of k; the worst-case band (mode 1) is ~8x wider.
int finalise_cand_max(int k, int mode) {
    int m = 512;
    while (m < 2 * k) m <<= 1;
    if (mode == 1) m *= 2;
    return m;
}
size_t finalise_work_bytes(int64_t nq, int k, int cand_max) {
    (void)k;
    return (size_t)nq * (size_t)cand_max * sizeof(uint64_t);
}

static int finalise_form(const FinaliseArgs& a, int64_t nq) {     // 0: one CTA per query, 1: select -> rescore+emit, 2: fused
    if (a.exact || !a.work || !a.ticket) return 0;
    return nq <= FIN_FUSED_MAX_Q ? 2 : 1;
}
int finalise_launches(const FinaliseArgs& a, int64_t nq) { return finalise_form(a, nq) == 1 ? 2 : 1; }
// The receiving end of the exchange can ride in the same launch when the cluster form is used (one emitting CTA per query)
// and the world's lists of a query fit its staging region twice over (items + ranks).
bool finalise_can_merge(int64_t nq, int k, int world, int P, int d_pad) {
    if (nq > FIN_FUSED_MAX_Q || (P + FF_MAX_CLUSTER - 1) / FF_MAX_CLUSTER > FF_THREADS) return false;
    size_t region = (size_t)FF_WARPS * d_pad * sizeof(float);
    const size_t phase1 = (size_t)FF_ITEMS * sizeof(uint64_t) + FF_BINS * sizeof(uint32_t);
    if (region < phase1) region = phase1;
    return (size_t)world * k * 16 + 64 <= region && (int64_t)world * k <= 2048;
}

bool launch_finalise(const FinaliseArgs& a_in, int64_t nq, cudaStream_t st) {
    if (nq <= 0) return false;
    FinaliseArgs a = a_in;
    const int cand_max = a.cand_max > 0 ? a.cand_max : finalise_cand_max(a.k, 0);
    int form = finalise_form(a, nq);
    int form_fallback = 0;
    if (form) {
        a.w_cand = static_cast<uint64_t*>(a.work);
        a.w_ticket = a.ticket;                       // zero-initialised once by the owner, self-resetting
        a.w_ncand = a.ticket + nq;
        a.w_flag = a.ticket + 2 * nq;
    }
    const size_t offs_bytes = ((a.P <= FIN_MAX_LISTS) ? (size_t)((a.P + 4) & ~3) : 4) * sizeof(int);
    if (form == 2) {
        // phase 1 (items + bins) and phase 3 (row staging) share one region.  Latency mode (default): 8 staging warps, 72 KB
        // per CTA, three CTAs per SM, as many CTAs per query as fit ONE wave.  Throughput mode (a.per_sm == 1, set by the
        // two-lane pipeline together with the 3-stage GEMM ring): 7 staging warps, 64 KB -- one such CTA fits on an SM NEXT
        // to a scanning GEMM CTA of the other lane, so the previous batch is finalised underneath the next batch's scan.
        const bool slim = a.per_sm == 1;
        const int rs_warps = slim ? FF_WARPS - 1 : FF_WARPS;
        size_t region = (size_t)rs_warps * a.d_pad * sizeof(float);
        const size_t phase1 = (size_t)FF_ITEMS * sizeof(uint64_t) + FF_BINS * sizeof(uint32_t);
        if (region < phase1) region = phase1;
        const int item_cap = (int)((region - FF_BINS * sizeof(uint32_t)) / sizeof(uint64_t)) & ~1;
        const size_t smem = (size_t)2 * cand_max * sizeof(uint64_t) + region;
        static int num_sms = 0;
        if (!num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); if (num_sms <= 0) num_sms = 148; }
        // CTAs per query (= cluster size): as many as fit in ONE wave, and enough that every CTA's lists fit
        static const int sizes[] = {1, 2, 3, 4, 6, 8};
        const int want = (int)((int64_t)(slim ? 1 : 3) * num_sms / nq);
        int S = 0;
        for (int cand : sizes) {
            if ((a.P + cand - 1) / cand > FF_THREADS) continue;    // every CTA must be able to take its share of the lists
            if (S == 0 || cand <= want) S = cand;
        }                                                         // S == 0: more lists than a cluster can take -> the two-kernel form below
        if (S) {
            static size_t attr_smem[64] = {};                     // largest size the attribute has been raised to, per device
            int dev = 0;
            cudaGetDevice(&dev);
            const size_t want_attr = smem > 96 * 1024 ? smem : 96 * 1024;
            if (dev < 0 || dev >= 64 || attr_smem[dev] < want_attr) {
                cudaFuncSetAttribute(finalise_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want_attr);
                cudaFuncSetAttribute(finalise_cluster_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
                if (dev >= 0 && dev < 64) attr_smem[dev] = want_attr;
            }
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)S, (unsigned)nq); cfg.blockDim = dim3(FF_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
            cudaLaunchAttribute attr[2];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[1].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = 2;
            if (a.merge.on && !(a.push.world > 0 && finalise_can_merge(nq, a.k, a.push.world, a.P, a.d_pad))) a.merge.on = 0;
            cudaLaunchKernelEx(&cfg, finalise_cluster_kernel, a, cand_max, item_cap, rs_warps);
            return a.merge.on != 0;
        }
    }
    a.merge.on = 0;                                   // the other forms emit (and push) only: the caller launches the merge kernel
    if (form == 2) form_fallback = 1;
    const size_t fixed = (size_t)cand_max * sizeof(uint64_t) + offs_bytes + (size_t)a.d_pad * sizeof(float) + 4096 * sizeof(uint32_t);
    // room for the gathered pool items: what the pools can hold, capped by the shared-memory budget
    size_t want_items = (size_t)a.P * (size_t)a.cap;
    size_t max_items = (FIN_SMEM_BUDGET > fixed) ? (FIN_SMEM_BUDGET - fixed) / sizeof(uint64_t) : 0;
    if (max_items > 12288) max_items = 12288;      // 96 KB: two CTAs per SM; larger pools take the global-memory path
    int item_cap = (int)(want_items < max_items ? want_items : max_items);
    item_cap = (item_cap + 1) & ~1;                  // keeps the fp32 query row behind it 16-byte aligned
    if (a.P > FIN_MAX_LISTS) item_cap = 0;
    const size_t smem = fixed + (size_t)item_cap * sizeof(uint64_t);
    if (form_fallback) form = 1;
    if (form == 1) {
        cudaFuncSetAttribute(finalise_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(FIN_SMEM_BUDGET + 1024));
        launch_pdl(finalise_kernel<true>, dim3((unsigned)nq), dim3(FIN_THREADS), smem, st, a, cand_max, item_cap);
        size_t rsm = (size_t)RS_WARPS * a.d_pad * sizeof(float);
        if (rsm < (size_t)cand_max * sizeof(uint64_t)) rsm = (size_t)cand_max * sizeof(uint64_t);
        if (rsm > 48 * 1024) cudaFuncSetAttribute(finalise_rescore_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsm);
        const unsigned rs_split = nq <= 512 ? 8u : 4u;
        launch_pdl(finalise_rescore_emit_kernel, dim3((unsigned)nq, rs_split), dim3(RS_WARPS * 32), rsm, st, a, cand_max);
    } else {
        cudaFuncSetAttribute(finalise_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(FIN_SMEM_BUDGET + 1024));
        launch_pdl(finalise_kernel<false>, dim3((unsigned)nq), dim3(FIN_THREADS), smem, st, a, cand_max, item_cap);
    }
    return false;
}

// ---- threshold bootstrap ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
sample_threshold_kernel(const uint64_t* __restrict__ pool_items, const int* __restrict__ pool_count, int P, int cap,
                        int k, const float* __restrict__ eps, float* __restrict__ thr0) {
    extern __shared__ uint64_t smp_items[];            // [8 * P]: the sample pass leaves at most 8 items per list
    __shared__ uint32_t hist[256];
    __shared__ uint32_t misc[2];
    __shared__ uint32_t sh_total;
    pdl_wait();
    const int64_t q = blockIdx.x;
    if (threadIdx.x == 0) sh_total = 0;
    __syncthreads();
    const int n = 8 * P;
    uint32_t mine = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int p = i >> 3, j = i & 7;
        const int64_t slot = pool_slot(q, p, P);
        const bool valid = j < pool_count[slot];
        smp_items[i] = valid ? pool_items[slot * cap + j] : 0ull;
        mine += valid ? 1u : 0u;
    }
    if (mine) atomicAdd(&sh_total, mine);
    __syncthreads();
    const uint32_t total = sh_total;
    if (total < (uint32_t)k) { if (threadIdx.x == 0) thr0[q] = -INFINITY; return; }
    const int n_up = (n + (int)blockDim.x - 1) / (int)blockDim.x * (int)blockDim.x;
    auto each = [&](auto fn) {
        for (int i = threadIdx.x; i < n_up; i += blockDim.x) {
            const uint64_t it = (i < n) ? smp_items[i] : 0ull;
            fn(it, it != 0ull);
        }
    };
    const uint64_t T = block_kth_largest(each, (uint32_t)k, 4, hist, misc);
    if (threadIdx.x == 0) {
        const float cut = key_score((uint32_t)(T >> 32)) - 2.f * eps[q];
        thr0[q] = nextafterf(cut, -INFINITY);        // the main pass keeps scores strictly above thr0
    }
}

void launch_sample_threshold(const uint64_t* pool_items, const int* pool_count, int P, int cap, int k,
                             const float* eps, float* thr0, int64_t nq, cudaStream_t st) {
    if (nq <= 0) return;
    const size_t smem = (size_t)8 * P * sizeof(uint64_t);
    if (smem > 48 * 1024) cudaFuncSetAttribute(sample_threshold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_pdl(sample_threshold_kernel, dim3((unsigned)nq), dim3(1024), smem, st, pool_items, pool_count, P, cap, k, eps, thr0);
}

// ---- multi-GPU merge -------------------------------------------------------------------------------
// in: [parts][nq][k] (score desc, id asc inside every part; parts own increasing id ranges, so the
// flat position p*k + r orders equal scores by ascending id).  One CTA per query.
// With sync.flags set this is the receiving end of the peer exchange: query q of part p was stored into this GPU's
// mailbox by rank p (its emit step or its push kernel), which then released flags[p][q] = epoch, so a query is merged
// as soon as ITS lists are in; the last CTA to finish reading acknowledges the epoch to every peer (their next store
// into this mailbox slot waits for it).
__global__ void __launch_bounds__(512)
merge_parts_kernel(const char* in_idx, const char* in_score, const char* in_status, int64_t idx_stride, int64_t score_stride, int64_t status_stride,
                   int parts, int64_t nq, int k, int64_t* __restrict__ out_idx, float* __restrict__ out_score, int32_t* __restrict__ out_status,
                   int m, const MergeSync sync) {
    extern __shared__ uint64_t cand[];                  // [m] power of two >= parts*k
    __shared__ int s_last;
    const int64_t q = blockIdx.x;
    const int total = parts * k;
    pdl_wait();                                         // launched early (programmatic dependent launch): the local search ahead of it has finished
    if (sync.flags) {
        if ((int)threadIdx.x < parts) wait_word_sys(sync.flags + (int64_t)threadIdx.x * sync.flag_stride + q, sync.epoch);
        __syncthreads();
    }
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        uint64_t it = 0ull;
        if (i < total) {
            const int p = i / k, r = i - p * k;
            const int64_t* ids = reinterpret_cast<const int64_t*>(in_idx + (int64_t)p * idx_stride);
            const float* sc = reinterpret_cast<const float*>(in_score + (int64_t)p * score_stride);
            if (ids[q * k + r] >= 0) it = make_item(sc[q * k + r], (uint32_t)i);
        }
        cand[i] = it;
    }
    block_sort_desc(cand, m);
    for (int r = threadIdx.x; r < k; r += blockDim.x) {
        const uint64_t it = cand[r];
        int64_t id = -1;
        float s = -INFINITY;
        if (it != 0ull) {
            const int i = (int)item_row(it);
            const int p = i / k, rr = i - p * k;
            id = reinterpret_cast<const int64_t*>(in_idx + (int64_t)p * idx_stride)[q * k + rr];
            s = reinterpret_cast<const float*>(in_score + (int64_t)p * score_stride)[q * k + rr];
        }
        out_idx[q * k + r] = id;
        if (out_score) out_score[q * k + r] = s;
    }
    if (out_status && threadIdx.x == 0) {               // a query is certified only if every shard certified its list
        int32_t st = 0;
        if (in_status)
            for (int p = 0; p < parts; ++p) st |= reinterpret_cast<const int32_t*>(in_status + (int64_t)p * status_stride)[q];
        out_status[q] = st;
    }
    if (sync.flags) {
        __syncthreads();                                // every read of the mailbox by this CTA has completed
        if (threadIdx.x == 0) {
            __threadfence();
            s_last = (atomicAdd(sync.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
        }
        __syncthreads();
        if (s_last) {
            if (threadIdx.x == 0) *sync.ticket = 0;
            if ((int)threadIdx.x < parts) st_release_sys(sync.ack[threadIdx.x], sync.epoch);
        }
    }
}

void launch_merge_parts(const void* in_idx, const void* in_score, const void* in_status, int64_t idx_stride, int64_t score_stride,
                        int64_t status_stride, int parts, int64_t nq, int k,
                        int64_t* out_idx, float* out_score, int32_t* out_status, cudaStream_t st, const MergeSync* sync) {
    if (nq <= 0) return;
    int m = 2;
    while (m < parts * k) m <<= 1;
    const size_t smem = (size_t)m * sizeof(uint64_t);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(merge_parts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    MergeSync none{};
    launch_pdl(merge_parts_kernel, dim3((unsigned)nq), dim3(512), smem, st, static_cast<const char*>(in_idx), static_cast<const char*>(in_score),
               static_cast<const char*>(in_status), idx_stride, score_stride, status_stride, parts, nq, k, out_idx, out_score, out_status, m,
               sync ? *sync : none);
}

// Sending end of the peer exchange as a kernel of its own (for payloads that were not produced by an emit step with
// the push fused in): the CTAs of column g copy this rank's packed result into rank g's mailbox (plain 16-byte stores
// through the NVLink peer mapping, every load issued before the first store; g == rank is a local copy) and the last
// of them releases rank g's per-query arrival flags.  Before overwriting the slot they wait for rank g's
// acknowledgement of the slot's previous epoch.
constexpr int PUSH_THREADS = 1024, PUSH_UNROLL = 8;
__global__ void __launch_bounds__(PUSH_THREADS)
exchange_push_kernel(const uint4* __restrict__ src, int64_t n16, const PushArgs a) {
    __shared__ int s_last;
    const int g = blockIdx.x;
    if (threadIdx.x == 0 && a.epoch > 1) wait_word_sys(a.my_acks + g, a.epoch - 1);
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(a.dst[g]);
    for (int64_t base = (int64_t)blockIdx.y * PUSH_THREADS * PUSH_UNROLL; base < n16; base += (int64_t)gridDim.y * PUSH_THREADS * PUSH_UNROLL) {
        uint4 v[PUSH_UNROLL];
#pragma unroll
        for (int u = 0; u < PUSH_UNROLL; ++u) {
            const int64_t i = base + u * PUSH_THREADS + threadIdx.x;
            if (i < n16) v[u] = src[i];
        }
#pragma unroll
        for (int u = 0; u < PUSH_UNROLL; ++u) {
            const int64_t i = base + u * PUSH_THREADS + threadIdx.x;
            if (i < n16) dst[i] = v[u];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        s_last = 1;
        if (gridDim.y > 1) {
            s_last = (atomicAdd(a.tickets + g, 1u) == gridDim.y - 1) ? 1 : 0;
            if (s_last) { a.tickets[g] = 0; __threadfence_system(); }
        }
    }
    __syncthreads();
    if (s_last)             // fence.sys above + relaxed stores = release of the whole payload, one flag per query
        for (int64_t i = threadIdx.x; i < a.nq; i += PUSH_THREADS) st_relaxed_sys(a.flag[g] + i, a.epoch);
}

void launch_exchange_push(const void* src, int64_t bytes, const PushArgs& a, int world, cudaStream_t st) {
    const int64_t n16 = bytes / 16;
    int64_t nblk = (n16 + (int64_t)PUSH_THREADS * PUSH_UNROLL - 1) / ((int64_t)PUSH_THREADS * PUSH_UNROLL);
    nblk = nblk < 1 ? 1 : (nblk > 16 ? 16 : nblk);
    exchange_push_kernel<<<dim3(world, (unsigned)nblk), PUSH_THREADS, 0, st>>>(static_cast<const uint4*>(src), n16, a);
}

}  // namespace xs
