// Stage 2: candidate pools -> exact top-k.
//
// One CTA per query: (1) radix-select the k-th best coarse score over all partial lists,
// (2) gather every candidate inside the exactness band below it, (3) rescore those rows exactly
// (fp32 operands, fp64 accumulation in a fixed order, one rounding to fp32 -- the deterministic
// stand-in for OpenBLAS' sgemm at src/main_retrieve.py:175), (4) sort by (score desc, id asc),
// (5) emit the first k, plus a certificate bit when anything outside the band may have been lost.
#include "common.cuh"
#include "select.cuh"
#include "internal.h"

namespace xs {

constexpr int FIN_THREADS = 512;
constexpr int FIN_MAX_LISTS = 4096;          // partial lists per query that the shared-memory gather supports
constexpr int FIN_SMEM_BUDGET = 200 * 1024;  // dynamic shared memory the finalise kernel may ask for

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
    const int flat = P * 32;
    for (int i = threadIdx.x; i < flat; i += blockDim.x) {
        const int p = i >> 5, j = i & 31;
        const int o = offs[p], c = offs[p + 1] - o;
        if (j < c) items[o + j] = pool_items[pool_slot(q, p, P) * cap + j];
    }
    for (int p = threadIdx.x >> 5; p < P; p += blockDim.x >> 5) {
        const int o = offs[p], c = offs[p + 1] - o;
        for (int j = 32 + (threadIdx.x & 31); j < c; j += 32) items[o + j] = pool_items[pool_slot(q, p, P) * cap + j];
    }
    (void)total;
    __syncthreads();
}

__global__ void __launch_bounds__(FIN_THREADS, 1)
finalise_kernel(FinaliseArgs a, int cand_max, int item_cap) {
    extern __shared__ uint64_t fin_smem[];              // [cand_max] candidates | [item_cap] gathered items | [P+1] offsets
    uint64_t* cand = fin_smem;
    uint64_t* items = fin_smem + cand_max;
    int* offs = reinterpret_cast<int*>(items + item_cap);
    float* qs = reinterpret_cast<float*>(offs + ((a.P <= FIN_MAX_LISTS) ? ((a.P + 4) & ~3) : 4));   // [d_pad] query row
    __shared__ uint32_t hist[256];
    __shared__ uint32_t misc[2];
    __shared__ int scan_scratch[33];
    __shared__ uint32_t sh_ncand, sh_flag, sh_selfkey;
    const int64_t q = blockIdx.x;
    const int lane = lane_id(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const bool staged = a.P <= FIN_MAX_LISTS;           // offsets fit -> try the shared-memory gather

    if (threadIdx.x == 0) { sh_ncand = 0; sh_flag = 0; sh_selfkey = 0; }
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
        uint64_t T = block_kth_largest(each, kk, a.exact ? 8 : 4, hist, misc);
        if (a.exact) { cut = T; cut_key = (uint32_t)(T >> 32); }
        else {
            cut_key = score_key(key_score((uint32_t)(T >> 32)) - 2.f * a.eps[q]);
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

    const int kout = min(a.k, ncand);
    for (int r = threadIdx.x; r < a.k; r += blockDim.x) {
        int64_t id = -1;
        float sc = -INFINITY;
        if (r < kout) {
            uint64_t it = cand[r];
            uint32_t key = item_key(it);
            if (a.self_base >= 0 && key == 0xFFFFFFFFu) key = sh_selfkey;
            id = (int64_t)item_row(it) + a.id_offset;
            sc = key_score(key);
        }
        a.out_idx[q * a.out_pitch + r] = id;
        if (a.out_score) a.out_score[q * a.out_pitch + r] = sc;
    }
    if (threadIdx.x == 0) {
        if (a.status) a.status[q] = uncertified ? ST_UNCERTIFIED : 0;
        if (a.n_cand) atomicAdd(a.n_cand, ncand);
    }
}

int finalise_cand_max(int k) {
    int m = 256;
    while (m < 2 * k) m <<= 1;
    return m;
}

void launch_finalise(const FinaliseArgs& a, int64_t nq, cudaStream_t st) {
    if (nq <= 0) return;
    const int cand_max = finalise_cand_max(a.k);
    const size_t offs_bytes = ((a.P <= FIN_MAX_LISTS) ? (size_t)((a.P + 4) & ~3) : 4) * sizeof(int);
    const size_t fixed = (size_t)cand_max * sizeof(uint64_t) + offs_bytes + (size_t)a.d_pad * sizeof(float);
    // room for the gathered pool items: what the pools can hold, capped by the shared-memory budget
    size_t want_items = (size_t)a.P * (size_t)a.cap;
    size_t max_items = (FIN_SMEM_BUDGET > fixed) ? (FIN_SMEM_BUDGET - fixed) / sizeof(uint64_t) : 0;
    if (max_items > 12288) max_items = 12288;      // 96 KB: two CTAs per SM; larger pools take the global-memory path
    int item_cap = (int)(want_items < max_items ? want_items : max_items);
    item_cap = (item_cap + 1) & ~1;                  // keeps the fp32 query row behind it 16-byte aligned
    if (a.P > FIN_MAX_LISTS) item_cap = 0;
    const size_t smem = fixed + (size_t)item_cap * sizeof(uint64_t);
    cudaFuncSetAttribute(finalise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(FIN_SMEM_BUDGET + 1024));
    finalise_kernel<<<(unsigned)nq, FIN_THREADS, smem, st>>>(a, cand_max, item_cap);
}

// ---- threshold bootstrap ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
sample_threshold_kernel(const uint64_t* __restrict__ pool_items, const int* __restrict__ pool_count, int P, int cap,
                        int k, const float* __restrict__ eps, float* __restrict__ thr0) {
    extern __shared__ uint64_t smp_items[];            // [8 * P]: the sample pass leaves at most 8 items per list
    __shared__ uint32_t hist[256];
    __shared__ uint32_t misc[2];
    __shared__ uint32_t sh_total;
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
    sample_threshold_kernel<<<(unsigned)nq, 1024, (size_t)8 * P * sizeof(uint64_t), st>>>(pool_items, pool_count, P, cap, k, eps, thr0);
}

// ---- multi-GPU merge -------------------------------------------------------------------------------
// in: [parts][nq][k] (score desc, id asc inside every part; parts own increasing id ranges, so the
// flat position p*k + r orders equal scores by ascending id).  One CTA per query.
__global__ void __launch_bounds__(512)
merge_parts_kernel(const int64_t* __restrict__ in_idx, const float* __restrict__ in_score, int parts,
                   int64_t nq, int k, int64_t* __restrict__ out_idx, float* __restrict__ out_score, int m) {
    extern __shared__ uint64_t cand[];                  // [m] power of two >= parts*k
    const int64_t q = blockIdx.x;
    const int total = parts * k;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        uint64_t it = 0ull;
        if (i < total) {
            int p = i / k, r = i - p * k;
            int64_t src = ((int64_t)p * nq + q) * k + r;
            if (in_idx[src] >= 0) it = make_item(in_score[src], (uint32_t)i);
        }
        cand[i] = it;
    }
    block_sort_desc(cand, m);
    for (int r = threadIdx.x; r < k; r += blockDim.x) {
        uint64_t it = cand[r];
        int64_t id = -1;
        float sc = -INFINITY;
        if (it != 0ull) {
            int i = (int)item_row(it);
            int p = i / k, rr = i - p * k;
            int64_t src = ((int64_t)p * nq + q) * k + rr;
            id = in_idx[src];
            sc = in_score[src];
        }
        out_idx[q * k + r] = id;
        if (out_score) out_score[q * k + r] = sc;
    }
}

void launch_merge_parts(const int64_t* in_idx, const float* in_score, int parts, int64_t nq, int k,
                        int64_t* out_idx, float* out_score, cudaStream_t st) {
    if (nq <= 0) return;
    int m = 2;
    while (m < parts * k) m <<= 1;
    const size_t smem = (size_t)m * sizeof(uint64_t);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(merge_parts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    merge_parts_kernel<<<(unsigned)nq, 512, smem, st>>>(in_idx, in_score, parts, nq, k, out_idx, out_score, m);
}

}  // namespace xs
