// Index build and query preparation kernels.
//
// The reference hands the matcher `vecs.T` -- an F-order (N,D) view of a (D,N) C-contiguous
// array (src/networks/imageretrievalnet.py:370, src/online.py:133) in fp32 or, by accident of
// np.concatenate on an fp64 seed, fp64 (src/online.py:96-100).  The device wants row-major rows:
//   db32 [N][d_pad] fp32  -- exact rescoring operand
//   db16 [n_pad][d_pad] bf16 -- coarse scoring operand (TMA / UMMA tiles, batch-1 scan)
// so the build is: stage a block of rows -> layout_rows (transpose + convert) -> finish_rows
// (optional L2 normalisation as nnsearch.py:693-697, bf16 rounding, norm statistics).
#include "common.cuh"
#include "internal.h"

namespace xs {

// ---- layout_rows ---------------------------------------------------------------------------------
template <typename T>
__global__ void layout_colmajor_kernel(const T* __restrict__ src, int64_t pitch, int64_t rows, int d,
                                       int d_pad, float* __restrict__ dst) {
    // src[c * pitch + r]  ->  dst[r * d_pad + c]; 32x32 tile through shared memory
    __shared__ float tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int c = c0 + j;
        int64_t r = r0 + threadIdx.x;
        tile[j][threadIdx.x] = (c < d && r < rows) ? (float)src[(int64_t)c * pitch + r] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int64_t r = r0 + j;
        int c = c0 + threadIdx.x;
        if (r < rows && c < d_pad) dst[r * d_pad + c] = tile[threadIdx.x][j];
    }
}

template <typename T>
__global__ void layout_rowmajor_kernel(const T* __restrict__ src, int64_t pitch, int64_t rows, int d,
                                       int d_pad, float* __restrict__ dst) {
    const int64_t r = blockIdx.x;
    if (r >= rows) return;
    for (int c = threadIdx.x; c < d_pad; c += blockDim.x)
        dst[r * d_pad + c] = (c < d) ? (float)src[r * pitch + c] : 0.f;
}

void launch_layout_rows(const void* src_tile, int dtype, bool colmajor, int64_t pitch,
                        int64_t rows, int d, int d_pad, float* dst32, cudaStream_t st) {
    if (rows <= 0) return;
    if (colmajor) {
        dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((d_pad + 31) / 32)), block(32, 8);
        if (dtype == 0) layout_colmajor_kernel<float><<<grid, block, 0, st>>>((const float*)src_tile, pitch, rows, d, d_pad, dst32);
        else            layout_colmajor_kernel<double><<<grid, block, 0, st>>>((const double*)src_tile, pitch, rows, d, d_pad, dst32);
    } else {
        if (dtype == 0) layout_rowmajor_kernel<float><<<(unsigned)rows, 256, 0, st>>>((const float*)src_tile, pitch, rows, d, d_pad, dst32);
        else            layout_rowmajor_kernel<double><<<(unsigned)rows, 256, 0, st>>>((const double*)src_tile, pitch, rows, d, d_pad, dst32);
    }
}

// ---- finish_rows / prep_queries ------------------------------------------------------------------
// One warp per row.  Sum of squares and of fourth powers in fp64 (fixed lane-strided order).
__device__ __forceinline__ void row_moments(const float* row, int d_pad, double& s2, double& s4) {
    double a2 = 0.0, a4 = 0.0;
    for (int c = lane_id() * 4; c < d_pad; c += 128) {
        float4 v = *reinterpret_cast<const float4*>(row + c);
        double x0 = v.x, x1 = v.y, x2 = v.z, x3 = v.w;
        double q0 = x0 * x0, q1 = x1 * x1, q2 = x2 * x2, q3 = x3 * x3;
        a2 += (q0 + q1) + (q2 + q3);
        a4 += (q0 * q0 + q1 * q1) + (q2 * q2 + q3 * q3);
    }
    s2 = warp_sum(a2);
    s4 = warp_sum(a4);
}

__device__ __forceinline__ void scale_and_round(float* row32, __nv_bfloat16* row16, int d_pad, float scale, bool do_scale) {
    for (int c = lane_id() * 4; c < d_pad; c += 128) {
        float4 v = *reinterpret_cast<const float4*>(row32 + c);
        if (do_scale) {
            v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
            *reinterpret_cast<float4*>(row32 + c) = v;
        }
        if (row16) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&lo);
            pk.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(row16 + c) = pk;
        }
    }
}

__global__ void finish_rows_kernel(float* rows32, __nv_bfloat16* rows16, int64_t rows, int d_pad,
                                   int renorm, DevStats* stats) {
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    float* row = rows32 + r * d_pad;
    double s2, s4;
    row_moments(row, d_pad, s2, s4);
    float scale = 1.f;
    if (renorm) {
        // a zero row stays zero (the reference would produce NaNs, nnsearch.py:697)
        scale = (s2 > 0.0) ? (float)(1.0 / sqrt(s2)) : 0.f;
        double sc = scale;
        s4 *= sc * sc * sc * sc;
        s2 *= sc * sc;
    }
    scale_and_round(row, rows16 ? rows16 + r * d_pad : nullptr, d_pad, scale, renorm != 0);
    if (lane_id() == 0) {
        float n4 = (float)sqrt(sqrt(s4)) * 1.000001f, n2 = (float)sqrt(s2) * 1.000001f;
        atomicMax(&stats->v4max_bits, __float_as_uint(n4));
        atomicMax(&stats->vnmax_bits, __float_as_uint(n2));
    }
}

void launch_finish_rows(float* rows32, __nv_bfloat16* rows16, int64_t rows, int d_pad, bool renorm,
                        DevStats* stats, cudaStream_t st) {
    if (rows <= 0) return;
    const int wpb = 8;
    finish_rows_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(rows32, rows16, rows, d_pad, renorm ? 1 : 0, stats);
}

// eps[q] bounds |bf16 coarse score - exact score| for query q against ANY database row:
// both operands are rounded to bf16 (unit roundoff u = 2^-9), so the error of one product is
// ~ v_i q_i (d1 + d2) with independent roundings of variance <= u^2/3; over the row the standard
// deviation is <= u sqrt(2/3) sqrt(sum (v_i q_i)^2) <= u sqrt(2/3) ||v||_4 ||q||_4 (Cauchy-Schwarz).
// eps = sigmas * that bound + a small absolute term for the fp32 accumulation inside the tensor core.
__global__ void prep_queries_kernel(float* q32, __nv_bfloat16* q16, int64_t nq, int d_pad, int renorm,
                                    const DevStats* stats, float eps_sigmas, float* eps) {
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= nq) return;
    float* row = q32 + r * d_pad;
    double s2, s4;
    row_moments(row, d_pad, s2, s4);
    float scale = 1.f;
    if (renorm) {
        scale = (s2 > 0.0) ? (float)(1.0 / sqrt(s2)) : 0.f;
        double sc = scale;
        s4 *= sc * sc * sc * sc;
        s2 *= sc * sc;
    }
    scale_and_round(row, q16 ? q16 + r * d_pad : nullptr, d_pad, scale, renorm != 0);
    if (lane_id() == 0) {
        const float v4 = __uint_as_float(stats->v4max_bits), vn = __uint_as_float(stats->vnmax_bits);
        const float q4 = (float)sqrt(sqrt(s4)), qn = (float)sqrt(s2);
        const float u = 1.0f / 512.0f;
        eps[r] = eps_sigmas * u * 0.8165f * q4 * v4 + 2e-5f * qn * vn;
    }
}

// Fast path for d == d_pad <= 2048: the raw row goes straight into registers (16 float4 per lane,
// all loads issued up front), moments, scaling, fp32 + bf16 stores and eps in one kernel; the rows
// [nq, nq_pad) of the bf16 copy (GEMM tile padding) are zero-filled by the same launch.
__global__ void prep_queries_fused_kernel(const float* __restrict__ raw, float* __restrict__ q32, __nv_bfloat16* __restrict__ q16,
                                          int64_t nq, int64_t nq_pad, int d_pad, int renorm, const DevStats* stats,
                                          float eps_sigmas, float* __restrict__ eps) {
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = lane_id();
    if (r >= nq_pad) return;
    if (r >= nq) {
        if (q16) for (int c = lane * 8; c < d_pad; c += 256) *reinterpret_cast<uint4*>(q16 + r * d_pad + c) = make_uint4(0, 0, 0, 0);
        return;
    }
    float4 v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int c = j * 128 + lane * 4;
        v[j] = (c < d_pad) ? *reinterpret_cast<const float4*>(raw + r * d_pad + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    double a2 = 0.0, a4 = 0.0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        double x0 = v[j].x, x1 = v[j].y, x2 = v[j].z, x3 = v[j].w;
        double q0 = x0 * x0, q1 = x1 * x1, q2 = x2 * x2, q3 = x3 * x3;
        a2 += (q0 + q1) + (q2 + q3);
        a4 += (q0 * q0 + q1 * q1) + (q2 * q2 + q3 * q3);
    }
    double s2 = warp_sum(a2), s4 = warp_sum(a4);
    float scale = 1.f;
    if (renorm) {
        scale = (s2 > 0.0) ? (float)(1.0 / sqrt(s2)) : 0.f;
        double sc = scale;
        s4 *= sc * sc * sc * sc;
        s2 *= sc * sc;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int c = j * 128 + lane * 4;
        if (c < d_pad) {
            float4 w = v[j];
            if (renorm) { w.x *= scale; w.y *= scale; w.z *= scale; w.w *= scale; }
            *reinterpret_cast<float4*>(q32 + r * d_pad + c) = w;
            if (q16) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(w.x, w.y), hi = __floats2bfloat162_rn(w.z, w.w);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&lo);
                pk.y = *reinterpret_cast<uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(q16 + r * d_pad + c) = pk;
            }
        }
    }
    if (lane == 0) {
        const float v4 = __uint_as_float(stats->v4max_bits), vn = __uint_as_float(stats->vnmax_bits);
        const float q4 = (float)sqrt(sqrt(s4)), qn = (float)sqrt(s2);
        eps[r] = eps_sigmas * (1.0f / 512.0f) * 0.8165f * q4 * v4 + 2e-5f * qn * vn;
    }
}

bool launch_prep_queries_fused(const float* raw, float* q32, __nv_bfloat16* q16, int64_t nq, int64_t nq_pad, int d, int d_pad,
                               bool renorm, const DevStats* stats, float eps_sigmas, float* eps, cudaStream_t st) {
    if (d != d_pad || d_pad > 2048 || (reinterpret_cast<uintptr_t>(raw) & 15)) return false;
    const int wpb = 4;
    const int64_t rows = q16 ? nq_pad : nq;
    prep_queries_fused_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(raw, q32, q16, nq, rows, d_pad, renorm ? 1 : 0,
                                                                                       stats, eps_sigmas, eps);
    return true;
}

// ---- tiled bf16 copy ---------------------------------------------------------------------------------
__global__ void tile_db16_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t n_pad, int d_pad) {
    // one thread per 16-byte chunk: row r, k-block kb, chunk c (8 chunks of 8 bf16 per 128-byte k-block row)
    const int chunks_per_row = d_pad >> 3;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad * chunks_per_row) return;
    const int64_t r = i / chunks_per_row;
    const int cc = (int)(i - r * chunks_per_row);
    const int kb = cc >> 3, c = cc & 7;
    const int KB = d_pad >> 6;
    const int64_t out = ((((r >> 8) * KB + kb) << 8) + (r & 255)) * 8 + c;
    dst[out] = src[i];
}

void launch_tile_db16(const __nv_bfloat16* db16, __nv_bfloat16* db16t, int64_t n_pad, int d_pad, cudaStream_t st) {
    const int64_t total = n_pad * (d_pad >> 3);
    tile_db16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(db16), reinterpret_cast<uint4*>(db16t), n_pad, d_pad);
}

// ---- AQE query construction ------------------------------------------------------------------------
// feature_enhancement of src/utils/Reranking.py:195-208 / :288-301 up to the re-score: for every query
// take its kq best database rows, weight them ((kq-j)/kq)^w (j = 0 best), sum, divide by (norm + 1e-6).
// One CTA per query, float64 accumulation (the reference's weights are float64, so numpy promotes).
__global__ void __launch_bounds__(256)
aqe_queries_kernel(const float* __restrict__ db32, const int64_t* __restrict__ top_ids, int kq, double w, int64_t n,
                   int d_pad, float* __restrict__ q_out) {
    __shared__ double red[8];
    __shared__ double sh_scale;
    const int64_t q = blockIdx.x;
    const int64_t* ids = top_ids + q * kq;
    double acc[8];                                      // d_pad <= 8 * 256 handled per pass; larger rows loop
    double ss = 0.0;
    for (int base = 0; base < d_pad; base += 8 * 256) {
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] = 0.0;
        for (int j = 0; j < kq; ++j) {
            const int64_t id = ids[j];
            if (id < 0 || id >= n) continue;
            const double wj = pow((double)(kq - j) / (double)kq, w);
            const float* row = db32 + id * d_pad;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int c = base + t * 256 + threadIdx.x;
                if (c < d_pad) acc[t] += wj * (double)row[c];
            }
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int c = base + t * 256 + threadIdx.x;
            if (c < d_pad) { ss += acc[t] * acc[t]; q_out[q * d_pad + c] = (float)acc[t]; }   // unscaled for now
        }
    }
    ss = warp_sum(ss);
    if (lane_id() == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += red[i];
        sh_scale = 1.0 / (sqrt(t) + 1e-6);
    }
    __syncthreads();
    // second pass: recompute in float64 and scale before the single rounding to fp32
    for (int base = 0; base < d_pad; base += 8 * 256) {
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] = 0.0;
        for (int j = 0; j < kq; ++j) {
            const int64_t id = ids[j];
            if (id < 0 || id >= n) continue;
            const double wj = pow((double)(kq - j) / (double)kq, w);
            const float* row = db32 + id * d_pad;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int c = base + t * 256 + threadIdx.x;
                if (c < d_pad) acc[t] += wj * (double)row[c];
            }
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int c = base + t * 256 + threadIdx.x;
            if (c < d_pad) q_out[q * d_pad + c] = (float)(acc[t] * sh_scale);
        }
    }
}

void launch_aqe_queries(const float* db32, const int64_t* top_ids, int64_t nq, int kq, double w, int64_t n, int d_pad,
                        float* q_out, cudaStream_t st) {
    if (nq <= 0) return;
    aqe_queries_kernel<<<(unsigned)nq, 256, 0, st>>>(db32, top_ids, kq, w, n, d_pad, q_out);
}

void launch_prep_queries(float* q32, __nv_bfloat16* q16, int64_t nq, int d_pad, bool renorm,
                         const DevStats* stats, float eps_sigmas, float* eps, cudaStream_t st) {
    if (nq <= 0) return;
    const int wpb = 4;
    prep_queries_kernel<<<(unsigned)((nq + wpb - 1) / wpb), wpb * 32, 0, st>>>(q32, q16, nq, d_pad, renorm ? 1 : 0, stats, eps_sigmas, eps);
}

}  // namespace xs
