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

// ---- database-wide score histogram ---------------------------------------------------------------------------------
// Both scoring kernels count every score into HIST_BINS bins that are LINEAR over [-B, +B], B = ||q|| * max ||v|| (no score
// can leave that interval): 4096 bins of width B / 2048 whatever the score level.  (Bins cut from the float's exponent
// and leading mantissa bits are 0.06 wide around 0.6 -- un-whitened descriptors score 0.6..0.7 against everything, the
// k-th best's bin then holds most of the database and the batch-1 path falls back to the exact scan on every query.)
__device__ __forceinline__ int score_bin(float s, float inv_bound) {
    const float t = fmaf(s, inv_bound, 1.0f) * (float)(HIST_BINS / 2);
    const int b = (int)t;                               // NaN -> 0
    return min(max(b, 0), HIST_BINS - 1);
}
__device__ __forceinline__ float bin_lower_edge(int b, float bound) {      // a score counted in bin b is >= this (minus rounding, see caller)
    return ((float)b * (1.0f / (float)(HIST_BINS / 2)) - 1.0f) * bound;
}
// ||q|| * max||v|| (slightly inflated) for the QB query rows in shared memory; every CTA computes the same values, CTA 0
// publishes them for scores_to_pools.  `red`: QB * 8 doubles.  Ends with a barrier.
template <int QB>
__device__ __forceinline__ void query_bounds(const float* qs, int d_pad, const DevStats* stats, double* red, float* bound_out, float (&inv)[QB]) {
    double a[QB];
#pragma unroll
    for (int b = 0; b < QB; ++b) {
        a[b] = 0.0;
        for (int i = threadIdx.x; i < d_pad; i += blockDim.x) a[b] += (double)qs[b * d_pad + i] * qs[b * d_pad + i];
        a[b] = warp_sum(a[b]);
    }
    if (lane_id() == 0)
#pragma unroll
        for (int b = 0; b < QB; ++b) red[b * 8 + (threadIdx.x >> 5)] = a[b];
    __syncthreads();
    const float vn = __uint_as_float(stats->vnmax_bits);
#pragma unroll
    for (int b = 0; b < QB; ++b) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[b * 8 + w];
        const float bound = (float)sqrt(t) * vn * 1.0001f;
        inv[b] = bound > 0.f ? 1.0f / bound : 0.f;
        if (blockIdx.x == 0 && threadIdx.x == 0) bound_out[b] = bound;
    }
}

// ---- scan_scores ---------------------------------------------------------------------------------
// One warp owns groups of R consecutive rows (R * d_pad * 2 contiguous bytes).  Per step every
// lane issues R 16-byte loads per 8-element chunk, two chunks in flight, so a warp keeps
// 2*R*512 B outstanding; with 16 resident warps per SM that is ~64 KB in flight per SM.
template <int QB, int R>
__global__ void __launch_bounds__(256, 2)
scan_scores_kernel(const uint4* __restrict__ db16, const float* __restrict__ q32, int64_t n, int d_pad,
                   float* __restrict__ scores, int64_t pitch, uint32_t* __restrict__ ghist, const DevStats* __restrict__ stats,
                   float* __restrict__ bounds) {
    extern __shared__ float qs[];                       // [QB][d_pad] query rows | [QB][HIST_BINS] score histogram
    __shared__ double red[QB * 8];
    pdl_wait();
    uint32_t* sh = reinterpret_cast<uint32_t*>(qs + QB * d_pad);
    for (int i = threadIdx.x; i < QB * d_pad; i += blockDim.x) qs[i] = q32[i];
    for (int i = threadIdx.x; i < QB * HIST_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    float inv[QB];
    query_bounds<QB>(qs, d_pad, stats, red, bounds, inv);
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
                if (lane == 0 && row0 + r < n) {
                    scores[(int64_t)b * pitch + row0 + r] = s;
                    atomicAdd(&sh[b * HIST_BINS + score_bin(s, inv[b])], 1u);
                }
            }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < QB * HIST_BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&ghist[i], sh[i]);
}

// The same pass over the TILED twin (tile_db16_kernel: [row >> 8][k-block][row & 255][64 bf16], what the GEMM's TMA boxes
// read), so an index that keeps only the twin still serves the batch-1 path.  Per k-block, 16 consecutive rows are 2 KB
// contiguous: a warp owns groups of 16 rows and issues four 512-byte loads per k-block (load j: rows 4j .. 4j+3, lane l
// holds row 4j + (l >> 3), 16-byte chunk l & 7), two k-blocks in flight -- the same 8 x 16 B outstanding per lane as above.
// A row's sum is spread over the 8 lanes that share l >> 3.
template <int QB>
__global__ void __launch_bounds__(256, 2)
scan_scores_tiled_kernel(const uint4* __restrict__ db16t, const float* __restrict__ q32, int64_t n, int d_pad,
                         float* __restrict__ scores, int64_t pitch, uint32_t* __restrict__ ghist, const DevStats* __restrict__ stats,
                         float* __restrict__ bounds) {
    extern __shared__ float qs[];                       // [QB][d_pad] query rows | [QB][HIST_BINS] score histogram
    __shared__ double red[QB * 8];
    pdl_wait();
    uint32_t* sh = reinterpret_cast<uint32_t*>(qs + QB * d_pad);
    for (int i = threadIdx.x; i < QB * d_pad; i += blockDim.x) qs[i] = q32[i];
    for (int i = threadIdx.x; i < QB * HIST_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    float inv[QB];
    query_bounds<QB>(qs, d_pad, stats, red, bounds, inv);
    constexpr int G = 16;                               // rows per warp group
    const int kbs = d_pad >> 6;                         // 64-column k-blocks
    const int lane = lane_id();
    const int sub = lane >> 3;                          // row within a 4-row load
    const float* qlane = qs + (lane & 7) * 8;           // this lane's 8 columns of every k-block
    const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t groups = (n + G - 1) / G;
    for (int64_t g = gw; g < groups; g += nw) {
        const int64_t row0 = g * G;                     // rows >= n are zero padding (n_pad is a multiple of 256)
        const uint4* base = db16t + (((row0 >> 8) * kbs) << 11) + ((row0 & 255) << 3) + lane;    // k-block 0; next k-block: + 2048 chunks
        float acc[4][QB];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int b = 0; b < QB; ++b) acc[j][b] = 0.f;
        for (int kb = 0; kb < kbs; kb += 2) {
            uint4 w0[4], w1[4];
            const bool two = (kb + 1) < kbs;
            const uint4* p = base + ((int64_t)kb << 11);
#pragma unroll
            for (int j = 0; j < 4; ++j) w0[j] = ld_stream_u4(p + j * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) w1[j] = two ? ld_stream_u4(p + 2048 + j * 32) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int b = 0; b < QB; ++b) {
                const float4 qa = *reinterpret_cast<const float4*>(qlane + b * d_pad + kb * 64);
                const float4 qb = *reinterpret_cast<const float4*>(qlane + b * d_pad + kb * 64 + 4);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float a = acc[j][b];
                    a = fmaf(bf16lo(w0[j].x), qa.x, a); a = fmaf(bf16hi(w0[j].x), qa.y, a);
                    a = fmaf(bf16lo(w0[j].y), qa.z, a); a = fmaf(bf16hi(w0[j].y), qa.w, a);
                    a = fmaf(bf16lo(w0[j].z), qb.x, a); a = fmaf(bf16hi(w0[j].z), qb.y, a);
                    a = fmaf(bf16lo(w0[j].w), qb.z, a); a = fmaf(bf16hi(w0[j].w), qb.w, a);
                    acc[j][b] = a;
                }
                if (two) {
                    const float4 qc = *reinterpret_cast<const float4*>(qlane + b * d_pad + kb * 64 + 64);
                    const float4 qd = *reinterpret_cast<const float4*>(qlane + b * d_pad + kb * 64 + 68);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float a = acc[j][b];
                        a = fmaf(bf16lo(w1[j].x), qc.x, a); a = fmaf(bf16hi(w1[j].x), qc.y, a);
                        a = fmaf(bf16lo(w1[j].y), qc.z, a); a = fmaf(bf16hi(w1[j].y), qc.w, a);
                        a = fmaf(bf16lo(w1[j].z), qd.x, a); a = fmaf(bf16hi(w1[j].z), qd.y, a);
                        a = fmaf(bf16lo(w1[j].w), qd.z, a); a = fmaf(bf16hi(w1[j].w), qd.w, a);
                        acc[j][b] = a;
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int b = 0; b < QB; ++b) {
                float s = acc[j][b];
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                const int64_t row = row0 + 4 * j + sub;
                if ((lane & 7) == 0 && row < n) {
                    scores[(int64_t)b * pitch + row] = s;
                    atomicAdd(&sh[b * HIST_BINS + score_bin(s, inv[b])], 1u);
                }
            }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < QB * HIST_BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&ghist[i], sh[i]);
}

void launch_scan_scores(const __nv_bfloat16* db16, bool tiled, const float* q32, int nq, int64_t n, int d_pad,
                        float* scores, int64_t score_pitch, uint32_t* ghist, const DevStats* stats, float* bounds, int num_sms, cudaStream_t st) {
    const int grid = num_sms * 2;
    const uint4* db = reinterpret_cast<const uint4*>(db16);
    const size_t per_q = (size_t)d_pad * sizeof(float) + HIST_BINS * sizeof(uint32_t);
    {   // dynamic shared memory limit: raised once per device (and again only if a wider index shows up), not on every query
        static int attr_bytes[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        const int need = (int)(2 * per_q);
        if (dev < 0 || dev >= 64 || attr_bytes[dev] < need) {
            cudaFuncSetAttribute(scan_scores_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
            cudaFuncSetAttribute(scan_scores_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
            cudaFuncSetAttribute(scan_scores_tiled_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
            cudaFuncSetAttribute(scan_scores_tiled_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
            if (dev >= 0 && dev < 64) attr_bytes[dev] = need;
        }
    }
    for (int q0 = 0; q0 < nq;) {
        const int left = nq - q0;
        const float* q = q32 + (int64_t)q0 * d_pad;
        float* s = scores + (int64_t)q0 * score_pitch;
        uint32_t* h = ghist + (size_t)q0 * HIST_BINS;
        if (left >= 2) {
            if (tiled) launch_pdl(scan_scores_tiled_kernel<2>, dim3(grid), dim3(256), 2 * per_q, st, db, q, n, d_pad, s, score_pitch, h, stats, bounds + q0);
            else launch_pdl(scan_scores_kernel<2, 4>, dim3(grid), dim3(256), 2 * per_q, st, db, q, n, d_pad, s, score_pitch, h, stats, bounds + q0);
            q0 += 2;
        } else {
            if (tiled) launch_pdl(scan_scores_tiled_kernel<1>, dim3(grid), dim3(256), per_q, st, db, q, n, d_pad, s, score_pitch, h, stats, bounds + q0);
            else launch_pdl(scan_scores_kernel<1, 4>, dim3(grid), dim3(256), per_q, st, db, q, n, d_pad, s, score_pitch, h, stats, bounds + q0);
            q0 += 1;
        }
    }
}

// ---- exact_scores --------------------------------------------------------------------------------
template <int QB, int R>
__global__ void __launch_bounds__(256, 2)
exact_scores_kernel(const float4* __restrict__ db32, const float* __restrict__ q32, int64_t n, int d_pad,
                    float* __restrict__ scores, int64_t pitch, uint32_t* __restrict__ ghist, const DevStats* __restrict__ stats,
                    float* __restrict__ bounds) {
    extern __shared__ float qs[];
    __shared__ double red[QB * 8];
    uint32_t* sh = reinterpret_cast<uint32_t*>(qs + QB * d_pad);
    for (int i = threadIdx.x; i < QB * d_pad; i += blockDim.x) qs[i] = q32[i];
    for (int i = threadIdx.x; i < QB * HIST_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    float inv[QB];
    query_bounds<QB>(qs, d_pad, stats, red, bounds, inv);
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
                if (lane == 0 && row0 + r < n) {
                    scores[(int64_t)b * pitch + row0 + r] = (float)s;
                    atomicAdd(&sh[b * HIST_BINS + score_bin((float)s, inv[b])], 1u);
                }
            }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < QB * HIST_BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&ghist[i], sh[i]);
}

void launch_exact_scores(const float* db32, const float* q32, int nq, int64_t n, int d_pad,
                         float* scores, int64_t score_pitch, uint32_t* ghist, const DevStats* stats, float* bounds, int num_sms, cudaStream_t st) {
    const int grid = num_sms * 2;
    const float4* db = reinterpret_cast<const float4*>(db32);
    const size_t per_q = (size_t)d_pad * sizeof(float) + HIST_BINS * sizeof(uint32_t);
    cudaFuncSetAttribute(exact_scores_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * per_q));
    cudaFuncSetAttribute(exact_scores_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * per_q));
    cudaFuncSetAttribute(exact_scores_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)per_q);
    for (int q0 = 0; q0 < nq;) {
        const int left = nq - q0;
        const float* q = q32 + (int64_t)q0 * d_pad;
        float* s = scores + (int64_t)q0 * score_pitch;
        uint32_t* h = ghist + (size_t)q0 * HIST_BINS;
        if (left >= 4) {
            exact_scores_kernel<4, 2><<<grid, 256, 4 * per_q, st>>>(db, q, n, d_pad, s, score_pitch, h, stats, bounds + q0);
            q0 += 4;
        } else if (left >= 2) {
            exact_scores_kernel<2, 2><<<grid, 256, 2 * per_q, st>>>(db, q, n, d_pad, s, score_pitch, h, stats, bounds + q0);
            q0 += 2;
        } else {
            exact_scores_kernel<1, 2><<<grid, 256, per_q, st>>>(db, q, n, d_pad, s, score_pitch, h, stats, bounds + q0);
            q0 += 1;
        }
    }
}

// ---- scores_to_pools -----------------------------------------------------------------------------
// grid = (P, nq); CTA (p, q) reduces rows [p*SLICE_ROWS, ...) of query q to one partial list.
// The database-wide histogram of score keys (filled by the scoring kernel) gives every CTA the bin
// that holds the k-th best score; its lower edge is a lower bound of that score, so
//   coarse scores: emit what lies within 2*eps below the edge (a pure filter, no per-slice select)
//   exact scores : emit the slice's own k best that are not below the edge (bounded by k per slice)
__global__ void __launch_bounds__(256)
scores_to_pools_kernel(const float* __restrict__ scores, int64_t pitch, int64_t n, int k,
                       const float* __restrict__ eps, const uint32_t* __restrict__ ghist, const float* __restrict__ bounds, int exact,
                       uint64_t* __restrict__ pool_items, int* __restrict__ pool_count, uint32_t* __restrict__ pool_thr,
                       int P, int cap) {
    __shared__ uint32_t keys[SLICE_ROWS];
    __shared__ uint32_t hist[256];
    __shared__ uint32_t misc[2];
    __shared__ uint32_t wsum[8];
    __shared__ uint32_t n_out, edge_key;
    pdl_wait();
    const int p = blockIdx.x;
    const int64_t q = blockIdx.y;
    const int64_t row0 = (int64_t)p * SLICE_ROWS;
    const int cnt = (int)min((int64_t)SLICE_ROWS, n - row0);
    const float* s = scores + q * pitch + row0;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) keys[i] = score_key(s[i]);
    if (threadIdx.x == 0) { n_out = 0; edge_key = 0; }
    {   // bin of the database's k-th best score: thread t owns HIST_BINS/256 consecutive bins, suffix sums from the top
        constexpr int PER = HIST_BINS / 256;
        const uint32_t* gh = ghist + q * HIST_BINS + threadIdx.x * PER;
        uint32_t h[PER], mine = 0;
#pragma unroll
        for (int j = 0; j < PER; ++j) { h[j] = gh[j]; mine += h[j]; }
        const int lane = lane_id(), warp = threadIdx.x >> 5;
        uint32_t suf = mine;                              // suffix over lanes >= lane within the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_down_sync(0xffffffffu, suf, o); if (lane + o < 32) suf += t; }
        if (lane == 0) wsum[warp] = suf;
        __syncthreads();
        uint32_t above = 0;                               // everything in higher warps
        for (int w = warp + 1; w < 8; ++w) above += wsum[w];
        const uint32_t incl = above + suf, excl = incl - mine;   // counts of bins >= my first bin / > my last bin
        if (incl >= (uint32_t)k && excl < (uint32_t)k) {          // exactly one thread (if the database holds >= k rows)
            uint32_t run = excl;
#pragma unroll
            for (int j = PER - 1; j >= 0; --j) {
                run += h[j];
                if (run >= (uint32_t)k) {
                    // lower edge of that bin, lowered by the rounding of the bin arithmetic: no score counted in it or above lies below
                    const float b = bounds[q];
                    edge_key = max(1u, score_key(bin_lower_edge(threadIdx.x * PER + j, b) - 1e-6f * b));
                    break;
                }
            }
        }
    }
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
    uint64_t cut;                                        // keep items >= cut
    if (!exact) {
        const uint32_t ck = edge_key ? score_key(key_score(edge_key) - 2.f * eps[q]) : 0u;
        cut = (uint64_t)ck << 32;
    } else {
        cut = (uint64_t)edge_key << 32;
        if (cnt > k) {
            const uint64_t T = block_kth_largest(each, (uint32_t)k, 8, hist, misc);
            cut = cut > T ? cut : T;
        }
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
                            const float* eps, const uint32_t* ghist, const float* bounds, bool exact, uint64_t* pool_items, int* pool_count,
                            uint32_t* pool_thr, int P, int cap, cudaStream_t st) {
    dim3 grid((unsigned)P, (unsigned)nq);
    launch_pdl(scores_to_pools_kernel, grid, dim3(256), 0, st, scores, score_pitch, n, k, eps, ghist, bounds, exact ? 1 : 0,
               pool_items, pool_count, pool_thr, P, cap);
}

}  // namespace xs
