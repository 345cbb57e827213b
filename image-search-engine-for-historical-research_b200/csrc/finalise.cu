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

__global__ void __launch_bounds__(FIN_THREADS)
finalise_kernel(FinaliseArgs a, int cand_max) {
    extern __shared__ uint64_t cand[];                  // [cand_max], cand_max is a power of two
    __shared__ uint32_t hist[256];
    __shared__ uint32_t misc[2];
    __shared__ uint32_t sh_total, sh_ncand, sh_flag, sh_selfkey;
    const int64_t q = blockIdx.x;
    const int lane = lane_id(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

    if (threadIdx.x == 0) { sh_total = 0; sh_ncand = 0; sh_flag = 0; sh_selfkey = 0; }
    __syncthreads();
    {   // total number of pooled candidates + largest "dropped above this" mark
        uint32_t t = 0, thr = 0;
        for (int p = threadIdx.x; p < a.P; p += blockDim.x) {
            int64_t slot = pool_slot(q, p, a.P);
            t += (uint32_t)a.pool_count[slot];
            thr = max(thr, a.pool_thr[slot]);
        }
        if (t) atomicAdd(&sh_total, t);
        if (thr) atomicMax(&sh_flag, thr);
    }
    __syncthreads();
    const uint32_t total = sh_total;
    const uint32_t max_thr = sh_flag;
    __syncthreads();
    if (threadIdx.x == 0) sh_flag = 0;

    auto each = [&](auto fn) {
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

    if (!a.exact) {   // exact rescoring, one warp per candidate
        const float* qrow = a.q32 + q * a.d_pad;
        for (int c = warp; c < ncand; c += nwarps) {
            const uint32_t row = item_row(cand[c]);
            const float* vrow = a.db32 + (int64_t)row * a.d_pad;
            double acc = 0.0;
            for (int i = lane * 4; i < a.d_pad; i += 128) {
                float4 v = ld_stream_f4(vrow + i);
                float4 w = *reinterpret_cast<const float4*>(qrow + i);
                acc = fma((double)v.x, (double)w.x, acc);
                acc = fma((double)v.y, (double)w.y, acc);
                acc = fma((double)v.z, (double)w.z, acc);
                acc = fma((double)v.w, (double)w.w, acc);
            }
            acc = warp_sum(acc);
            if (lane == 0) cand[c] = make_item((float)acc, row);
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
            const float* qrow = a.q32 + q * a.d_pad;
            double acc = 0.0;
            for (int i = lane * 4; i < a.d_pad; i += 128) {
                float4 v = *reinterpret_cast<const float4*>(vrow + i);
                float4 w = *reinterpret_cast<const float4*>(qrow + i);
                acc = fma((double)v.x, (double)w.x, acc);
                acc = fma((double)v.y, (double)w.y, acc);
                acc = fma((double)v.z, (double)w.z, acc);
                acc = fma((double)v.w, (double)w.w, acc);
            }
            acc = warp_sum(acc);
            if (lane == 0) sh_selfkey = score_key((float)acc);
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
    const size_t smem = (size_t)cand_max * sizeof(uint64_t);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(finalise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    finalise_kernel<<<(unsigned)nq, FIN_THREADS, smem, st>>>(a, cand_max);
}

// ---- threshold bootstrap ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sample_threshold_kernel(const uint64_t* __restrict__ pool_items, const int* __restrict__ pool_count, int P, int cap,
                        int k, const float* __restrict__ eps, float* __restrict__ thr0) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t misc[2];
    __shared__ uint32_t sh_total;
    const int64_t q = blockIdx.x;
    const int lane = lane_id(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (threadIdx.x == 0) sh_total = 0;
    __syncthreads();
    uint32_t t = 0;
    for (int p = threadIdx.x; p < P; p += blockDim.x) t += (uint32_t)pool_count[pool_slot(q, p, P)];
    if (t) atomicAdd(&sh_total, t);
    __syncthreads();
    const uint32_t total = sh_total;
    if (total < (uint32_t)k) { if (threadIdx.x == 0) thr0[q] = -INFINITY; return; }
    auto each = [&](auto fn) {
        for (int p = warp; p < P; p += nwarps) {
            const int64_t slot = pool_slot(q, p, P);
            const int cnt = pool_count[slot];
            const uint64_t* lst = pool_items + slot * cap;
            for (int b = 0; b < cnt; b += 32) {
                int i = b + lane;
                bool valid = i < cnt;
                fn(valid ? lst[i] : 0ull, valid);
            }
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
    sample_threshold_kernel<<<(unsigned)nq, 256, 0, st>>>(pool_items, pool_count, P, cap, k, eps, thr0);
}

// ---- multi-GPU merge -------------------------------------------------------------------------------
// in: [parts][nq][k] (score desc, id asc inside every part; parts own increasing id ranges, so the
// flat position p*k + r orders equal scores by ascending id).  One CTA per query.
__global__ void __launch_bounds__(FIN_THREADS)
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
    merge_parts_kernel<<<(unsigned)nq, FIN_THREADS, smem, st>>>(in_idx, in_score, parts, nq, k, out_idx, out_score, m);
}

}  // namespace xs
