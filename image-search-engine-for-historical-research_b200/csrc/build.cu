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

// ---- random rotation ------------------------------------------------------------------------------------
// Before anything is rounded to bf16, database rows and queries go through the same orthogonal map
//     R = H D2 H D1        (D1, D2: random sign flips from the index's private seed; H: normalised Walsh-Hadamard)
// so that <Rv, Rq> = <v, q> while the bf16 rounding errors of Rv / Rq no longer depend on any structure of the data
// (constant components, low-entropy mantissas, rows built against a known query ...): for ANY fixed pair (v, q) the
// coordinates of Rv, Rq are generic, and the statistical error model behind eps (see prep_queries_kernel) holds with
// the stated probability over the seed instead of "for data that happens to look random".  The exact stage never sees
// R: rescoring uses the original fp32 rows and queries.
// d_pad is a multiple of 64, not necessarily a power of two: H is block diagonal over the binary decomposition of
// d_pad (largest block first in round 1, smallest first in round 2, so the two rounds mix across block borders).
__device__ __forceinline__ float rot_sign(uint32_t seed, uint32_t round, uint32_t i) {
    uint32_t h = (i + round * 0x9E3779B9u) * 0x85EBCA6Bu ^ seed;
    h ^= h >> 15; h *= 0xC2B2AE35u; h ^= h >> 13; h *= 0x27D4EB2Fu; h ^= h >> 16;
    return (h & 0x80000000u) ? -1.f : 1.f;
}
// Whole CTA, row x[0..d_pad) in shared memory.
__device__ void block_rotate(float* x, int d_pad, uint32_t seed) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int round = 0; round < 2; ++round) {
        for (int i = tid; i < d_pad; i += nt) x[i] *= rot_sign(seed, (uint32_t)round, (uint32_t)i);
        __syncthreads();
        int o = 0, rem = d_pad;
        while (rem > 0) {
            const int B = (round == 0) ? (1 << (31 - __clz(rem))) : (rem & -rem);      // largest / smallest power of two first
            for (int lh = 0; (1 << lh) < B; ++lh) {
                const int h = 1 << lh;
                for (int i = tid; i < (B >> 1); i += nt) {
                    const int j = o + ((i >> lh) << (lh + 1)) + (i & (h - 1));
                    const float a = x[j], b = x[j + h];
                    x[j] = a + b; x[j + h] = a - b;
                }
                __syncthreads();
            }
            const float scale = rsqrtf((float)B);
            for (int i = tid; i < B; i += nt) x[o + i] *= scale;
            o += B; rem -= B;
        }
        __syncthreads();
    }
}

// ---- finish_rows / prep_queries ------------------------------------------------------------------
// One CTA of ROW_THREADS threads per row, the row held in shared memory; sums in fp64, fixed order.
constexpr int ROW_THREADS = 128;
__device__ __forceinline__ void block_sum2(double& a, double& b, double* red) {      // red: 2 * ROW_THREADS/32 doubles
    a = warp_sum(a); b = warp_sum(b);
    const int warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane_id() == 0) { red[warp] = a; red[ROW_THREADS / 32 + warp] = b; }
    __syncthreads();
    double ta = 0.0, tb = 0.0;
#pragma unroll
    for (int w = 0; w < ROW_THREADS / 32; ++w) { ta += red[w]; tb += red[ROW_THREADS / 32 + w]; }
    a = ta; b = tb;
}
__device__ __forceinline__ void smem_moments(const float* x, int d_pad, double& s2, double& s4, double* red) {
    double a2 = 0.0, a4 = 0.0;
    for (int c = threadIdx.x; c < d_pad; c += ROW_THREADS) {
        const double v = x[c], q = v * v;
        a2 += q; a4 += q * q;
    }
    block_sum2(a2, a4, red);
    s2 = a2; s4 = a4;
}
// bf16 rounding of x[0..d_pad) into dst (row-major); returns the squared norms of the rounding residual and of the rounded row
__device__ __forceinline__ void smem_round_bf16(const float* x, int d_pad, __nv_bfloat16* dst, double& res2, double& r2, double* red) {
    double a = 0.0, b = 0.0;
    for (int c = threadIdx.x * 2; c < d_pad; c += 2 * ROW_THREADS) {
        const float v0 = x[c], v1 = x[c + 1];
        const __nv_bfloat162 pk = __floats2bfloat162_rn(v0, v1);
        if (dst) *reinterpret_cast<__nv_bfloat162*>(dst + c) = pk;
        const float w0 = __low2float(pk), w1 = __high2float(pk);
        const double e0 = (double)v0 - (double)w0, e1 = (double)v1 - (double)w1;
        a += e0 * e0 + e1 * e1;
        b += (double)w0 * w0 + (double)w1 * w1;
    }
    block_sum2(a, b, red);
    res2 = a; r2 = b;
}

__global__ void __launch_bounds__(ROW_THREADS)
finish_rows_kernel(float* rows32, __nv_bfloat16* rows16, int64_t rows, int d_pad, int renorm, int rotate, uint32_t seed, DevStats* stats) {
    extern __shared__ float fr_smem[];
    __shared__ double red[2 * ROW_THREADS / 32];
    float* x = fr_smem;
    const int64_t r = blockIdx.x;
    float* row = rows32 + r * d_pad;
    for (int c = threadIdx.x * 4; c < d_pad; c += 4 * ROW_THREADS) *reinterpret_cast<float4*>(x + c) = *reinterpret_cast<const float4*>(row + c);
    __syncthreads();
    double s2, s4;
    smem_moments(x, d_pad, s2, s4, red);
    if (renorm) {
        // a zero row stays zero (the reference would produce NaNs, nnsearch.py:697)
        const float scale = (s2 > 0.0) ? (float)(1.0 / sqrt(s2)) : 0.f;
        for (int c = threadIdx.x; c < d_pad; c += ROW_THREADS) { const float v = x[c] * scale; x[c] = v; row[c] = v; }
        __syncthreads();
        smem_moments(x, d_pad, s2, s4, red);
    }
    if (rotate) { block_rotate(x, d_pad, seed); double t2; smem_moments(x, d_pad, t2, s4, red); }
    double res2, r2;
    smem_round_bf16(x, d_pad, rows16 ? rows16 + r * d_pad : nullptr, res2, r2, red);
    if (threadIdx.x == 0) {
        const float n4 = (float)sqrt(sqrt(s4)) * 1.000001f, n2 = (float)sqrt(s2) * 1.000001f, rho = (float)sqrt(res2) * 1.000001f;
        atomicMax(&stats->v4max_bits, __float_as_uint(n4));
        atomicMax(&stats->vnmax_bits, __float_as_uint(n2));
        atomicMax(&stats->rhomax_bits, __float_as_uint(rho));
    }
}


// Query preparation, one launch: optional copy from the caller's raw fp32 rows (pitch d), optional normalisation
// (nnsearch.py:693-697), the rotated copies the coarse kernels read (bf16 for the GEMM, fp32 for the batch-1 scan),
// zero rows [nq, nq_pad) of the bf16 copy (GEMM tile padding) and the per-query band eps.
//
// eps[q] bounds |coarse score - exact score| of query q against ANY database row.
//   mode 0 (statistical, default): both operands are rounded to bf16.  The relative rounding error of one element is
//     uniform in +-2^-8/m (m = its mantissa in [1,2)): standard deviation 2^-8 sqrt(E[1/m^2]/12) = 0.85 * 2^-9 for
//     log-uniform mantissas.  A product v'_i q'_i carries the sum of two such errors -- independent for unrelated
//     vectors, IDENTICAL when the row is a copy of the query -- so its standard deviation is at most 1.7 * 2^-9 |v'_i q'_i|,
//     and over a row (independent across coordinates: that is what the random rotation buys)
//         sigma <= 1.7 * 2^-9 * sqrt(sum (v'_i q'_i)^2) <= 1.7 * 2^-9 ||v'||_4 ||q'||_4      (Cauchy-Schwarz; primes = rotated).
//     eps = sigmas * that + a small absolute term for fp32 accumulation and the rotation's own rounding.  For unrelated
//     unit vectors the true sigma is ~3x smaller than the bound (the default 8 "sigmas" are ~24 of theirs); the bound is
//     attained by near-duplicates of the query.
//   mode 1 (worst case): |<v16,q16> - <v',q'>| <= rho_v ||q16|| + ||v'|| rho_q by Cauchy-Schwarz with the stored
//     residual norms rho = ||x' - bf16(x')||, plus d_pad 2^-22 ||v|| ||q|| for ANY order of fp32 accumulation and
//     1e-5 ||v|| ||q|| for the rotation arithmetic.  No assumption at all, ~3x the candidates.
__global__ void __launch_bounds__(ROW_THREADS)
prep_queries_kernel(const float* __restrict__ raw, int raw_pitch, int d, float* q32, __nv_bfloat16* q16, float* q32r,
                    int64_t nq, int d_pad, int renorm, int rotate, uint32_t seed, const DevStats* stats,
                    float eps_sigmas, int eps_mode, float* __restrict__ eps) {
    extern __shared__ float pq_smem[];
    __shared__ double red[2 * ROW_THREADS / 32];
    pdl_wait();
    const int64_t r = blockIdx.x;
    if (r >= nq) {                                      // GEMM tile padding
        if (q16) for (int c = threadIdx.x * 8; c < d_pad; c += 8 * ROW_THREADS) *reinterpret_cast<uint4*>(q16 + r * d_pad + c) = make_uint4(0, 0, 0, 0);
        return;
    }
    float* x = pq_smem;
    if (raw) { for (int c = threadIdx.x; c < d_pad; c += ROW_THREADS) x[c] = (c < d) ? raw[r * raw_pitch + c] : 0.f; }
    else     { for (int c = threadIdx.x * 4; c < d_pad; c += 4 * ROW_THREADS) *reinterpret_cast<float4*>(x + c) = *reinterpret_cast<const float4*>(q32 + r * d_pad + c); }
    __syncthreads();
    double s2, s4;
    smem_moments(x, d_pad, s2, s4, red);
    if (renorm) {
        const float scale = (s2 > 0.0) ? (float)(1.0 / sqrt(s2)) : 0.f;
        for (int c = threadIdx.x; c < d_pad; c += ROW_THREADS) x[c] *= scale;
        __syncthreads();
        smem_moments(x, d_pad, s2, s4, red);
    }
    if (raw || renorm) for (int c = threadIdx.x * 4; c < d_pad; c += 4 * ROW_THREADS) *reinterpret_cast<float4*>(q32 + r * d_pad + c) = *reinterpret_cast<const float4*>(x + c);
    if (rotate) { __syncthreads(); block_rotate(x, d_pad, seed); double t2; smem_moments(x, d_pad, t2, s4, red); }
    if (q32r) for (int c = threadIdx.x * 4; c < d_pad; c += 4 * ROW_THREADS) *reinterpret_cast<float4*>(q32r + r * d_pad + c) = *reinterpret_cast<const float4*>(x + c);
    double res2, r2;
    smem_round_bf16(x, d_pad, q16 ? q16 + r * d_pad : nullptr, res2, r2, red);
    if (threadIdx.x == 0) {
        const float v4 = __uint_as_float(stats->v4max_bits), vn = __uint_as_float(stats->vnmax_bits), vrho = __uint_as_float(stats->rhomax_bits);
        const float q4 = (float)sqrt(sqrt(s4)), qn = (float)sqrt(s2);
        float e;
        if (eps_mode == 1) {
            const float q16n = (float)sqrt(r2) * 1.000001f, qrho = (float)sqrt(res2) * 1.000001f;
            e = vrho * q16n + vn * qrho + ((float)d_pad * 2.3841858e-7f + 1e-5f) * qn * vn;
        } else {
            e = eps_sigmas * (1.0f / 512.0f) * 1.7f * q4 * v4 + 2e-5f * qn * vn;
        }
        eps[r] = e;
    }
}


// ---- register / shuffle form of the same kernels for d_pad = 128 * J, J a power of two (128 .. 4096 columns) ---------
// Thread t of the 128-thread CTA holds elements j * 128 + t (coalesced rows): the Walsh-Hadamard butterflies on the j bits
// are register-to-register, those on the five lane bits are shuffles, and only the two warp bits go through shared memory
// -- four barriers per round instead of one per butterfly stage, which is what the per-request latency of prep feels.
template <int J>
__device__ __forceinline__ void cta_rotate_regs(float (&v)[J], float* xch, uint32_t seed) {
    const int tid = threadIdx.x, lane = tid & 31;
    const float scale = rsqrtf((float)(J * ROW_THREADS));
#pragma unroll
    for (int round = 0; round < 2; ++round) {
#pragma unroll
        for (int j = 0; j < J; ++j) v[j] *= rot_sign(seed, (uint32_t)round, (uint32_t)(j * ROW_THREADS + tid));
#pragma unroll
        for (int m = 1; m < J; m <<= 1)
#pragma unroll
            for (int j = 0; j < J; ++j)
                if (!(j & m)) { const float a = v[j], b = v[j | m]; v[j] = a + b; v[j | m] = a - b; }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const float other = __shfl_xor_sync(0xffffffffu, v[j], o);
                v[j] = (lane & o) ? other - v[j] : v[j] + other;
            }
#pragma unroll
        for (int wb = 32; wb < ROW_THREADS; wb <<= 1) {
            __syncthreads();
#pragma unroll
            for (int j = 0; j < J; ++j) xch[j * ROW_THREADS + tid] = v[j];
            __syncthreads();
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const float other = xch[j * ROW_THREADS + (tid ^ wb)];
                v[j] = (tid & wb) ? other - v[j] : v[j] + other;
            }
        }
#pragma unroll
        for (int j = 0; j < J; ++j) v[j] *= scale;
    }
}

__device__ __forceinline__ void block_sum4(double (&a)[4], double* red) {      // red: 4 * ROW_THREADS/32 doubles
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = warp_sum(a[i]);
    const int warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane_id() == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) red[i * (ROW_THREADS / 32) + warp] = a[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < ROW_THREADS / 32; ++w) t += red[i * (ROW_THREADS / 32) + w];
        a[i] = t;
    }
}

// Shared tail of the two fast kernels: v = the (normalised) row.  Rotates (optional), rounds to bf16 and returns
// sums[0] = ||v'||^2, sums[1] = sum v'^4, sums[2] = ||v' - bf16(v')||^2, sums[3] = ||bf16(v')||^2.
template <int J>
__device__ __forceinline__ void rotate_round_reduce(float (&v)[J], float* xch, double* red, int rotate, uint32_t seed,
                                                    __nv_bfloat16* dst16, float* dst32r, double (&sums)[4]) {
    const int tid = threadIdx.x;
    if (rotate) cta_rotate_regs<J>(v, xch, seed);
    sums[0] = sums[1] = sums[2] = sums[3] = 0.0;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const float x = v[j];
        const __nv_bfloat16 h = __float2bfloat16_rn(x);
        const float w = __bfloat162float(h);
        if (dst16) dst16[j * ROW_THREADS + tid] = h;
        if (dst32r) dst32r[j * ROW_THREADS + tid] = x;
        const double q = (double)x * x, e = (double)x - (double)w;
        sums[0] += q; sums[1] += q * q; sums[2] += e * e; sums[3] += (double)w * w;
    }
    block_sum4(sums, red);
}

template <int J>
__global__ void __launch_bounds__(ROW_THREADS)
finish_rows_fast_kernel(float* rows32, __nv_bfloat16* rows16, int64_t rows, int renorm, int rotate, uint32_t seed, DevStats* stats) {
    __shared__ float xch[J * ROW_THREADS];
    __shared__ double red[4 * ROW_THREADS / 32];
    constexpr int d_pad = J * ROW_THREADS;
    const int tid = threadIdx.x;
    const int64_t r = blockIdx.x;
    float* row = rows32 + r * d_pad;
    float v[J];
#pragma unroll
    for (int j = 0; j < J; ++j) v[j] = row[j * ROW_THREADS + tid];
    double s2 = 0.0;
    if (renorm) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int j = 0; j < J; ++j) a += (double)v[j] * v[j];
        block_sum2(a, b, red);
        const float scale = (a > 0.0) ? (float)(1.0 / sqrt(a)) : 0.f;       // a zero row stays zero (the reference would produce NaNs, nnsearch.py:697)
#pragma unroll
        for (int j = 0; j < J; ++j) { v[j] *= scale; row[j * ROW_THREADS + tid] = v[j]; }
    }
    if (!rotate) { double a = 0.0, b = 0.0; for (int j = 0; j < J; ++j) a += (double)v[j] * v[j]; block_sum2(a, b, red); s2 = a; }
    double sums[4];
    rotate_round_reduce<J>(v, xch, red, rotate, seed, rows16 ? rows16 + r * d_pad : nullptr, nullptr, sums);
    if (rotate) s2 = sums[0];
    if (tid == 0) {
        const float n4 = (float)sqrt(sqrt(sums[1])) * 1.000001f, n2 = (float)sqrt(s2) * 1.00001f, rho = (float)sqrt(sums[2]) * 1.000001f;
        atomicMax(&stats->v4max_bits, __float_as_uint(n4));
        atomicMax(&stats->vnmax_bits, __float_as_uint(n2));
        atomicMax(&stats->rhomax_bits, __float_as_uint(rho));
    }
}

template <int J>
__global__ void __launch_bounds__(ROW_THREADS)
prep_queries_fast_kernel(const float* __restrict__ raw, int raw_pitch, int d, float* q32, __nv_bfloat16* q16, float* q32r,
                         int64_t nq, int renorm, int rotate, uint32_t seed, const DevStats* stats,
                         float eps_sigmas, int eps_mode, float* __restrict__ eps) {
    __shared__ float xch[J * ROW_THREADS];
    __shared__ double red[4 * ROW_THREADS / 32];
    constexpr int d_pad = J * ROW_THREADS;
    pdl_wait();
    const int tid = threadIdx.x;
    const int64_t r = blockIdx.x;
    if (r >= nq) {                                      // GEMM tile padding
        if (q16) for (int c = tid * 8; c < d_pad; c += 8 * ROW_THREADS) *reinterpret_cast<uint4*>(q16 + r * d_pad + c) = make_uint4(0, 0, 0, 0);
        return;
    }
    float v[J];
    if (raw) {
#pragma unroll
        for (int j = 0; j < J; ++j) { const int c = j * ROW_THREADS + tid; v[j] = (c < d) ? raw[r * raw_pitch + c] : 0.f; }
    } else {
#pragma unroll
        for (int j = 0; j < J; ++j) v[j] = q32[r * d_pad + j * ROW_THREADS + tid];
    }
    if (renorm) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int j = 0; j < J; ++j) a += (double)v[j] * v[j];
        block_sum2(a, b, red);
        const float scale = (a > 0.0) ? (float)(1.0 / sqrt(a)) : 0.f;
#pragma unroll
        for (int j = 0; j < J; ++j) v[j] *= scale;
    }
    if (raw || renorm) {
#pragma unroll
        for (int j = 0; j < J; ++j) q32[r * d_pad + j * ROW_THREADS + tid] = v[j];
    }
    double sums[4];
    rotate_round_reduce<J>(v, xch, red, rotate, seed, q16 ? q16 + r * d_pad : nullptr, q32r ? q32r + r * d_pad : nullptr, sums);
    if (tid == 0) {
        const float v4 = __uint_as_float(stats->v4max_bits), vn = __uint_as_float(stats->vnmax_bits), vrho = __uint_as_float(stats->rhomax_bits);
        const float q4 = (float)sqrt(sqrt(sums[1])), qn = (float)sqrt(sums[0]) * 1.00001f;    // the rotation preserves the norm (to rounding)
        float e;
        if (eps_mode == 1) {
            const float q16n = (float)sqrt(sums[3]) * 1.000001f, qrho = (float)sqrt(sums[2]) * 1.000001f;
            e = vrho * q16n + vn * qrho + ((float)d_pad * 2.3841858e-7f + 1e-5f) * qn * vn;
        } else {
            e = eps_sigmas * (1.0f / 512.0f) * 1.7f * q4 * v4 + 2e-5f * qn * vn;
        }
        eps[r] = e;
    }
}

template <int J>
static void launch_prep_fast(const float* raw, int d, float* q32, __nv_bfloat16* q16, float* q32r, int64_t nq, int64_t rows,
                             bool renorm, bool rotate, uint32_t seed, const DevStats* stats, float eps_sigmas, int eps_mode, float* eps, cudaStream_t st) {
    launch_pdl(prep_queries_fast_kernel<J>, dim3((unsigned)rows), dim3(ROW_THREADS), 0, st,
               raw, d, d, q32, q16, q32r, nq, renorm ? 1 : 0, rotate ? 1 : 0, seed, stats, eps_sigmas, eps_mode, eps);
}

void launch_prep_queries(const float* raw, int d, float* q32, __nv_bfloat16* q16, float* q32r, int64_t nq, int64_t nq_pad, int d_pad,
                         bool renorm, bool rotate, uint32_t seed, const DevStats* stats, float eps_sigmas, int eps_mode, float* eps,
                         cudaStream_t st) {
    if (nq <= 0) return;
    const int64_t rows = q16 ? nq_pad : nq;
    switch (d_pad) {
        case 128:  return launch_prep_fast<1>(raw, d, q32, q16, q32r, nq, rows, renorm, rotate, seed, stats, eps_sigmas, eps_mode, eps, st);
        case 256:  return launch_prep_fast<2>(raw, d, q32, q16, q32r, nq, rows, renorm, rotate, seed, stats, eps_sigmas, eps_mode, eps, st);
        case 512:  return launch_prep_fast<4>(raw, d, q32, q16, q32r, nq, rows, renorm, rotate, seed, stats, eps_sigmas, eps_mode, eps, st);
        case 1024: return launch_prep_fast<8>(raw, d, q32, q16, q32r, nq, rows, renorm, rotate, seed, stats, eps_sigmas, eps_mode, eps, st);
        case 2048: return launch_prep_fast<16>(raw, d, q32, q16, q32r, nq, rows, renorm, rotate, seed, stats, eps_sigmas, eps_mode, eps, st);
        default: break;
    }
    const size_t smem = (size_t)d_pad * sizeof(float);
    if (smem > 48 * 1024) cudaFuncSetAttribute(prep_queries_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_pdl(prep_queries_kernel, dim3((unsigned)rows), dim3(ROW_THREADS), smem, st,
               raw, d, d, q32, q16, q32r, nq, d_pad, renorm ? 1 : 0, rotate ? 1 : 0, seed, stats, eps_sigmas, eps_mode, eps);
}

void launch_finish_rows(float* rows32, __nv_bfloat16* rows16, int64_t rows, int d_pad, bool renorm, bool rotate, uint32_t seed,
                        DevStats* stats, cudaStream_t st) {
    if (rows <= 0) return;
    const unsigned g = (unsigned)rows;
    const int rn = renorm ? 1 : 0, ro = rotate ? 1 : 0;
    switch (d_pad) {
        case 128:  finish_rows_fast_kernel<1><<<g, ROW_THREADS, 0, st>>>(rows32, rows16, rows, rn, ro, seed, stats); return;
        case 256:  finish_rows_fast_kernel<2><<<g, ROW_THREADS, 0, st>>>(rows32, rows16, rows, rn, ro, seed, stats); return;
        case 512:  finish_rows_fast_kernel<4><<<g, ROW_THREADS, 0, st>>>(rows32, rows16, rows, rn, ro, seed, stats); return;
        case 1024: finish_rows_fast_kernel<8><<<g, ROW_THREADS, 0, st>>>(rows32, rows16, rows, rn, ro, seed, stats); return;
        case 2048: finish_rows_fast_kernel<16><<<g, ROW_THREADS, 0, st>>>(rows32, rows16, rows, rn, ro, seed, stats); return;
        default: break;
    }
    const size_t smem = (size_t)d_pad * sizeof(float);
    if (smem > 48 * 1024) cudaFuncSetAttribute(finish_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    finish_rows_kernel<<<g, ROW_THREADS, smem, st>>>(rows32, rows16, rows, d_pad, rn, ro, seed, stats);
}

// ---- tiled bf16 copy ---------------------------------------------------------------------------------
__global__ void tile_db16_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t r0, int64_t rows, int d_pad) {
    // one thread per 16-byte chunk: row r0 + rl of the database = row rl of `src`, k-block kb, chunk c (8 chunks of 8 bf16
    // per 128-byte k-block row)
    const int chunks_per_row = d_pad >> 3;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * chunks_per_row) return;
    const int64_t rl = i / chunks_per_row;
    const int cc = (int)(i - rl * chunks_per_row);
    const int64_t r = r0 + rl;
    const int kb = cc >> 3, c = cc & 7;
    const int KB = d_pad >> 6;
    const int64_t out = ((((r >> 8) * KB + kb) << 8) + (r & 255)) * 8 + c;
    dst[out] = src[i];
}

void launch_tile_db16(const __nv_bfloat16* rows16, __nv_bfloat16* db16t, int64_t r0, int64_t rows, int d_pad, cudaStream_t st) {
    const int64_t total = rows * (d_pad >> 3);
    if (total <= 0) return;
    tile_db16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(rows16), reinterpret_cast<uint4*>(db16t), r0, rows, d_pad);
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

}  // namespace xs
