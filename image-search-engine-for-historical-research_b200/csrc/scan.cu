// HBM-bound scoring kernels and the score-matrix -> candidate-pool reduction.
//
//   scan_scores   batch-1 online path (src/online.py:133 -> nnsearch.py:699-703): one pass over the
//                 bf16 database, fp32 query, fp32 accumulate.  Algorithmic bytes: N * D * 2.
//   exact_scores  exact path: fp32 database rows, fp64 accumulate, result rounded once to fp32.
//                 Used when a query cannot be certified by the bf16 pass and for full rankings.
//   scores_to_pools  per 4096-row slice, radix-select the k-th best score and emit the survivors.
#include "common.cuh"
#include "select.cuh"
#include "internal.h"

namespace xs {

// ---- scan_scores ---------------------------------------------------------------------------------
// One warp owns groups of R consecutive rows (R * d_pad * 2 contiguous bytes).  Per step every
// lane issues R 16-byte loads per 8-element chunk, two chunks in flight, so a warp keeps
// 2*R*512 B outstanding; with 16 resident warps per SM that is ~64 KB in flight per SM.
template <int QB, int R>
__global__ void __launch_bounds__(256, 2)
scan_scores_kernel(const uint4* __restrict__ db16, const float* __restrict__ q32, int64_t n, int d_pad,
                   float* __restrict__ scores, int64_t pitch) {
    extern __shared__ float qs[];                       // [QB][d_pad]
    for (int i = threadIdx.x; i < QB * d_pad; i += blockDim.x) qs[i] = q32[i];
    __syncthreads();
    const int chunks = d_pad >> 3;                      // 16-byte chunks per row
    const int lane = lane_id();
    const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t groups = (n + R - 1) / R;
    for (int64_t g = gw; g < groups; g += nw) {
        const int64_t row0 = g * R;
        const uint4* base = db16 + row0 * chunks;       // rows >= n are zero padding (n_pad), safe to read
        float acc[R][QB];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int b = 0; b < QB; ++b) acc[r][b] = 0.f;
        for (int c = lane; c < chunks; c += 64) {
            uint4 w0[R], w1[R];
            const bool two = (c + 32) < chunks;
#pragma unroll
            for (int r = 0; r < R; ++r) w0[r] = ld_stream_u4(base + (int64_t)r * chunks + c);
#pragma unroll
            for (int r = 0; r < R; ++r) w1[r] = two ? ld_stream_u4(base + (int64_t)r * chunks + c + 32) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int b = 0; b < QB; ++b) {
                const float4 qa = *reinterpret_cast<const float4*>(qs + b * d_pad + c * 8);
                const float4 qb = *reinterpret_cast<const float4*>(qs + b * d_pad + c * 8 + 4);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float a = acc[r][b];
                    a = fmaf(bf16lo(w0[r].x), qa.x, a); a = fmaf(bf16hi(w0[r].x), qa.y, a);
                    a = fmaf(bf16lo(w0[r].y), qa.z, a); a = fmaf(bf16hi(w0[r].y), qa.w, a);
                    a = fmaf(bf16lo(w0[r].z), qb.x, a); a = fmaf(bf16hi(w0[r].z), qb.y, a);
                    a = fmaf(bf16lo(w0[r].w), qb.z, a); a = fmaf(bf16hi(w0[r].w), qb.w, a);
                    acc[r][b] = a;
                }
                if (two) {
                    const float4 qc = *reinterpret_cast<const float4*>(qs + b * d_pad + (c + 32) * 8);
                    const float4 qd = *reinterpret_cast<const float4*>(qs + b * d_pad + (c + 32) * 8 + 4);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float a = acc[r][b];
                        a = fmaf(bf16lo(w1[r].x), qc.x, a); a = fmaf(bf16hi(w1[r].x), qc.y, a);
                        a = fmaf(bf16lo(w1[r].y), qc.z, a); a = fmaf(bf16hi(w1[r].y), qc.w, a);
                        a = fmaf(bf16lo(w1[r].z), qd.x, a); a = fmaf(bf16hi(w1[r].z), qd.y, a);
                        a = fmaf(bf16lo(w1[r].w), qd.z, a); a = fmaf(bf16hi(w1[r].w), qd.w, a);
                        acc[r][b] = a;
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int b = 0; b < QB; ++b) {
                float s = warp_sum(acc[r][b]);
                if (lane == 0 && row0 + r < n) scores[(int64_t)b * pitch + row0 + r] = s;
            }
    }
}

void launch_scan_scores(const __nv_bfloat16* db16, const float* q32, int nq, int64_t n, int d_pad,
                        float* scores, int64_t score_pitch, int num_sms, cudaStream_t st) {
    const int grid = num_sms * 2;
    const uint4* db = reinterpret_cast<const uint4*>(db16);
    for (int q0 = 0; q0 < nq;) {
        const int left = nq - q0;
        const float* q = q32 + (int64_t)q0 * d_pad;
        float* s = scores + (int64_t)q0 * score_pitch;
        if (left >= 2) {
            scan_scores_kernel<2, 4><<<grid, 256, 2 * d_pad * sizeof(float), st>>>(db, q, n, d_pad, s, score_pitch);
            q0 += 2;
        } else {
            scan_scores_kernel<1, 4><<<grid, 256, 1 * d_pad * sizeof(float), st>>>(db, q, n, d_pad, s, score_pitch);
            q0 += 1;
        }
    }
}

// ---- exact_scores --------------------------------------------------------------------------------
template <int QB, int R>
__global__ void __launch_bounds__(256, 2)
exact_scores_kernel(const float4* __restrict__ db32, const float* __restrict__ q32, int64_t n, int d_pad,
                    float* __restrict__ scores, int64_t pitch) {
    extern __shared__ float qs[];
    for (int i = threadIdx.x; i < QB * d_pad; i += blockDim.x) qs[i] = q32[i];
    __syncthreads();
    const int chunks = d_pad >> 2;                      // float4 per row
    const int lane = lane_id();
    const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t groups = (n + R - 1) / R;
    for (int64_t g = gw; g < groups; g += nw) {
        const int64_t row0 = g * R;
        double acc[R][QB];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int b = 0; b < QB; ++b) acc[r][b] = 0.0;
        for (int c = lane; c < chunks; c += 32) {
            float4 w[R];
#pragma unroll
            for (int r = 0; r < R; ++r)
                w[r] = (row0 + r < n) ? ld_stream_f4(db32 + (row0 + r) * chunks + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int b = 0; b < QB; ++b) {
                const float4 qv = *reinterpret_cast<const float4*>(qs + b * d_pad + c * 4);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    double a = acc[r][b];
                    a = fma((double)w[r].x, (double)qv.x, a);
                    a = fma((double)w[r].y, (double)qv.y, a);
                    a = fma((double)w[r].z, (double)qv.z, a);
                    a = fma((double)w[r].w, (double)qv.w, a);
                    acc[r][b] = a;
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int b = 0; b < QB; ++b) {
                double s = warp_sum(acc[r][b]);
                if (lane == 0 && row0 + r < n) scores[(int64_t)b * pitch + row0 + r] = (float)s;
            }
    }
}

void launch_exact_scores(const float* db32, const float* q32, int nq, int64_t n, int d_pad,
                         float* scores, int64_t score_pitch, int num_sms, cudaStream_t st) {
    const int grid = num_sms * 2;
    const float4* db = reinterpret_cast<const float4*>(db32);
    for (int q0 = 0; q0 < nq;) {
        const int left = nq - q0;
        const float* q = q32 + (int64_t)q0 * d_pad;
        float* s = scores + (int64_t)q0 * score_pitch;
        if (left >= 4) {
            exact_scores_kernel<4, 2><<<grid, 256, 4 * d_pad * sizeof(float), st>>>(db, q, n, d_pad, s, score_pitch);
            q0 += 4;
        } else if (left >= 2) {
            exact_scores_kernel<2, 2><<<grid, 256, 2 * d_pad * sizeof(float), st>>>(db, q, n, d_pad, s, score_pitch);
            q0 += 2;
        } else {
            exact_scores_kernel<1, 2><<<grid, 256, 1 * d_pad * sizeof(float), st>>>(db, q, n, d_pad, s, score_pitch);
            q0 += 1;
        }
    }
}

// ---- scores_to_pools -----------------------------------------------------------------------------
// grid = (P, nq); CTA (p, q) reduces rows [p*SLICE_ROWS, ...) of query q to one partial list.
__global__ void __launch_bounds__(256)
scores_to_pools_kernel(const float* __restrict__ scores, int64_t pitch, int64_t n, int k,
                       const float* __restrict__ eps, const float* __restrict__ thr0, int exact,
                       uint64_t* __restrict__ pool_items, int* __restrict__ pool_count, uint32_t* __restrict__ pool_thr,
                       int P, int cap) {
    __shared__ uint32_t keys[SLICE_ROWS];
    __shared__ uint32_t hist[256];
    __shared__ uint32_t misc[2];
    __shared__ uint32_t n_out;
    const int p = blockIdx.x;
    const int64_t q = blockIdx.y;
    const int64_t row0 = (int64_t)p * SLICE_ROWS;
    const int cnt = (int)min((int64_t)SLICE_ROWS, n - row0);
    const float* s = scores + q * pitch + row0;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) keys[i] = score_key(s[i]);
    if (threadIdx.x == 0) n_out = 0;
    __syncthreads();
    const int cnt_up = (cnt + 255) & ~255;
    auto each = [&](auto fn) {
        for (int b = 0; b < cnt_up; b += 256) {
            int i = b + threadIdx.x;
            bool valid = i < cnt;
            uint64_t it = valid ? ((uint64_t)keys[i] << 32) | (uint64_t)(0xFFFFFFFFu - (uint32_t)(row0 + i)) : 0ull;
            fn(it, valid);
        }
    };
    uint64_t cut = 0;                                   // keep items >= cut
    if (cnt > k) {
        uint64_t T = block_kth_largest(each, (uint32_t)k, exact ? 8 : 4, hist, misc);
        if (exact) cut = T;
        else cut = (uint64_t)score_key(key_score((uint32_t)(T >> 32)) - 2.f * eps[q]) << 32;
    }
    if (thr0) {                                         // database-wide lower bound from the sampled scores
        const uint64_t c0 = (uint64_t)score_key(thr0[q]) << 32;
        cut = cut > c0 ? cut : c0;
    }
    const int64_t slot = pool_slot(q, p, P);
    uint64_t* out = pool_items + slot * cap;
    each([&](uint64_t it, bool valid) {
        bool take = valid && it >= cut;
        uint32_t m = __ballot_sync(0xffffffffu, take);
        uint32_t base = 0;
        if (lane_id() == 0 && m) base = atomicAdd(&n_out, (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        uint32_t pos = base + __popc(m & lanemask_lt());
        if (take && pos < (uint32_t)cap) out[pos] = it;
    });
    __syncthreads();
    if (threadIdx.x == 0) {
        pool_count[slot] = (int)min(n_out, (uint32_t)cap);
        pool_thr[slot] = (n_out > (uint32_t)cap) ? 0xFFFFFFFFu : 0u;   // overflow: something was dropped
    }
}

void launch_scores_to_pools(const float* scores, int64_t score_pitch, int nq, int64_t n, int k,
                            const float* eps, const float* thr0, bool exact, uint64_t* pool_items, int* pool_count,
                            uint32_t* pool_thr, int P, int cap, cudaStream_t st) {
    dim3 grid((unsigned)P, (unsigned)nq);
    scores_to_pools_kernel<<<grid, 256, 0, st>>>(scores, score_pitch, n, k, eps, thr0, exact ? 1 : 0,
                                                 pool_items, pool_count, pool_thr, P, cap);
}

// ---- scores_sample_threshold -----------------------------------------------------------------------
// One CTA per query: the k-th best of SCORE_SAMPLE scores taken at a fixed stride over the row range
// is a lower bound of the k-th best of all rows; slices then emit only what can still matter.
constexpr int SCORE_SAMPLE = 16384;
__global__ void __launch_bounds__(1024)
scores_sample_threshold_kernel(const float* __restrict__ scores, int64_t pitch, int64_t n, int k,
                               const float* __restrict__ eps, int exact, float* __restrict__ thr0) {
    extern __shared__ uint32_t skeys[];                 // [SCORE_SAMPLE]
    __shared__ uint32_t hist[256];
    __shared__ uint32_t misc[2];
    const int64_t q = blockIdx.x;
    const int64_t stride = n / SCORE_SAMPLE > 0 ? n / SCORE_SAMPLE : 1;
    const int cnt = (int)min((int64_t)SCORE_SAMPLE, n / stride);
    const float* s = scores + q * pitch;
    for (int i0 = threadIdx.x; i0 < cnt; i0 += blockDim.x * 8) {      // 8 independent strided loads in flight per thread
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int i = i0 + j * blockDim.x; v[j] = (i < cnt) ? s[(int64_t)i * stride] : 0.f; }
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int i = i0 + j * blockDim.x; if (i < cnt) skeys[i] = score_key(v[j]); }
    }
    __syncthreads();
    if (cnt < 4 * k) { if (threadIdx.x == 0) thr0[q] = -INFINITY; return; }
    const int cnt_up = (cnt + (int)blockDim.x - 1) / (int)blockDim.x * (int)blockDim.x;
    auto each = [&](auto fn) {
        for (int i = threadIdx.x; i < cnt_up; i += blockDim.x) {
            bool valid = i < cnt;
            fn(valid ? ((uint64_t)skeys[i] << 32) : 0ull, valid);
        }
    };
    const uint64_t T = block_kth_largest(each, (uint32_t)k, 4, hist, misc);
    if (threadIdx.x == 0) {
        const float kth = key_score((uint32_t)(T >> 32));
        thr0[q] = exact ? kth : nextafterf(kth - 2.f * eps[q], -INFINITY);
    }
}

void launch_scores_sample_threshold(const float* scores, int64_t score_pitch, int nq, int64_t n, int k,
                                    const float* eps, bool exact, float* thr0, cudaStream_t st) {
    cudaFuncSetAttribute(scores_sample_threshold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SCORE_SAMPLE * 4);
    scores_sample_threshold_kernel<<<(unsigned)nq, 1024, SCORE_SAMPLE * 4, st>>>(scores, score_pitch, n, k, eps, exact ? 1 : 0, thr0);
}

}  // namespace xs
