// Gallery-side diffusion: one truncated conjugate-gradient solve per database row
// (src/utils/diffusion.py:15-19 `get_offline_result`, driven from :74-76):
//     ids      = trunc_ids[i]                       the row's n_trunc nearest neighbours, itself first
//     trunc_lap = lap_alpha[ids][:, ids]            n_trunc x n_trunc principal sub-matrix of the Laplacian
//     scores, _ = cg(trunc_lap, e_0, tol=1e-6, maxiter=20)
// The reference slices a scipy matrix and calls scipy's CG once per row from joblib threads.  Here one CTA owns
// one row at a time:
//   1. the row's ids go into a shared-memory hash (global id -> local position);
//   2. the sub-matrix is assembled ONCE into a per-CTA scratch (local column u16 + value f32, one fixed-stride
//      segment per local row) by probing the hash with the CSR columns of every selected row;
//   3. up to `maxiter` CG steps run with x, r, p, q in shared memory in fp64 (scipy upcasts the fp32 matrix
//      against the fp64 right-hand side, so the arithmetic type is the same); one warp per local row for A*p.
// The iteration is the textbook one scipy implements (no preconditioner): stop BEFORE a step when
// ||r|| < tol * ||b|| (= tol, b = e_0), otherwise return the iterate after `maxiter` steps.
#include "common.cuh"
#include "internal.h"

namespace xs {

constexpr int CG_THREADS = 512;
constexpr int CG_WARPS = CG_THREADS / 32;

struct CgArgs {
    const int64_t* indptr;      // Laplacian CSR (n rows), device
    const int32_t* indices;
    const float* values;
    const int32_t* trunc_ids;   // [rows, T] device, ids in [0, n)
    int64_t n;
    int64_t rows;
    int T, H, hash_shift, stride, maxiter;
    double atol;
    uint16_t* s_cols;           // scratch: gridDim.x * T * stride
    float* s_vals;
    float* out;                 // [rows, T]
    int* bad;                   // set to 1 when an id is out of range
};

__device__ __forceinline__ uint32_t cg_hash(int32_t g, int shift) { return ((uint32_t)g * 2654435761u) >> shift; }

// Deterministic CTA-wide sum; `buf` holds CG_WARPS doubles.  The caller alternates between two buffers so that
// one barrier per call is enough.
__device__ __forceinline__ double cg_block_sum(double v, double* buf) {
    v = warp_sum(v);
    if (lane_id() == 0) buf[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < CG_WARPS; ++w) s += buf[w];
    return s;
}

__global__ void __launch_bounds__(CG_THREADS)
diffusion_cg_kernel(const CgArgs a) {
    extern __shared__ __align__(16) unsigned char cg_smem[];
    const int T = a.T, H = a.H;
    double* x = reinterpret_cast<double*>(cg_smem);
    double* r = x + T;
    double* p = r + T;
    double* q = p + T;
    int32_t* hkeys = reinterpret_cast<int32_t*>(q + T);
    uint16_t* hvals = reinterpret_cast<uint16_t*>(hkeys + H);
    uint16_t* cnt = hvals + H;
    __shared__ double red[2][CG_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint16_t* cols = a.s_cols + (size_t)blockIdx.x * T * a.stride;
    float* vals = a.s_vals + (size_t)blockIdx.x * T * a.stride;

    for (int64_t i = blockIdx.x; i < a.rows; i += gridDim.x) {
        const int32_t* ids = a.trunc_ids + i * T;

        // 1. global id -> local position
        for (int h = tid; h < H; h += CG_THREADS) hkeys[h] = -1;
        __syncthreads();
        for (int t = tid; t < T; t += CG_THREADS) {
            const int32_t g = ids[t];
            if (g < 0 || g >= a.n) { *a.bad = 1; continue; }
            uint32_t h = cg_hash(g, a.hash_shift);
            while (true) {
                const int32_t prev = atomicCAS(&hkeys[h], -1, g);
                if (prev == -1) { hvals[h] = (uint16_t)t; break; }
                if (prev == g) break;                              // repeated id: the first position owns the column
                h = (h + 1) & (H - 1);
            }
        }
        __syncthreads();

        // 2. the principal sub-matrix, local row by local row
        for (int rr = warp; rr < T; rr += CG_WARPS) {
            const int32_t g = ids[rr];
            int c = 0;
            if (g >= 0 && g < a.n) {
                const int64_t b = a.indptr[g], e = a.indptr[g + 1];
                uint16_t* crow = cols + (size_t)rr * a.stride;
                float* vrow = vals + (size_t)rr * a.stride;
                for (int64_t base = b; base < e; base += 32) {
                    const int64_t k = base + lane;
                    bool ok = false;
                    uint16_t lc = 0;
                    float v = 0.f;
                    if (k < e) {
                        const int32_t col = a.indices[k];
                        v = a.values[k];
                        uint32_t h = cg_hash(col, a.hash_shift);
                        while (true) {
                            const int32_t key = hkeys[h];
                            if (key == col) { ok = true; lc = hvals[h]; break; }
                            if (key == -1) break;
                            h = (h + 1) & (H - 1);
                        }
                    }
                    const uint32_t m = __ballot_sync(0xffffffffu, ok);
                    const int pos = c + __popc(m & lanemask_lt());
                    if (ok && pos < a.stride) { crow[pos] = lc; vrow[pos] = v; }
                    c += __popc(m);
                }
                c = min(c, a.stride);
            }
            if (lane == 0) cnt[rr] = (uint16_t)c;
        }
        for (int t = tid; t < T; t += CG_THREADS) { x[t] = 0.0; r[t] = (t == 0) ? 1.0 : 0.0; }
        __syncthreads();

        // 3. conjugate gradients on  A x = e_0,  x0 = 0
        double rho = 1.0, rho_prev = 1.0;
        for (int it = 0; it < a.maxiter; ++it) {
            if (!(sqrt(rho) >= a.atol)) break;
            if (it == 0) {
                for (int t = tid; t < T; t += CG_THREADS) p[t] = r[t];
            } else {
                const double beta = rho / rho_prev;
                for (int t = tid; t < T; t += CG_THREADS) p[t] = beta * p[t] + r[t];
            }
            __syncthreads();
            double part = 0.0;
            for (int rr = warp; rr < T; rr += CG_WARPS) {
                const int c = cnt[rr];
                const uint16_t* crow = cols + (size_t)rr * a.stride;
                const float* vrow = vals + (size_t)rr * a.stride;
                double acc = 0.0;
                for (int e = lane; e < c; e += 32) acc = fma((double)vrow[e], p[crow[e]], acc);
                acc = warp_sum(acc);
                if (lane == 0) { q[rr] = acc; part = fma(p[rr], acc, part); }
            }
            const double pq = cg_block_sum(part, red[0]);
            const double alpha = rho / pq;
            part = 0.0;
            for (int t = tid; t < T; t += CG_THREADS) {
                x[t] = fma(alpha, p[t], x[t]);
                const double rn = fma(-alpha, q[t], r[t]);
                r[t] = rn;
                part = fma(rn, rn, part);
            }
            rho_prev = rho;
            rho = cg_block_sum(part, red[1]);
        }
        __syncthreads();
        for (int t = tid; t < T; t += CG_THREADS) a.out[i * T + t] = (float)x[t];
        __syncthreads();
    }
}

size_t diffusion_cg_smem(int T, int H) { return (size_t)T * 4 * sizeof(double) + (size_t)H * 6 + (size_t)T * 2 + 16; }

int diffusion_cg_hash_size(int T) {
    int H = 64;
    while (H < 2 * T) H <<= 1;
    return H;
}

// Resident CTAs of the solver for this problem size (the scratch is sized per resident CTA).
cudaError_t diffusion_cg_grid(int T, int* grid_out) {
    const int H = diffusion_cg_hash_size(T);
    const size_t smem = diffusion_cg_smem(T, H);
    cudaError_t e = cudaFuncSetAttribute(diffusion_cg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0, dev = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, diffusion_cg_kernel, CG_THREADS, smem);
    if (e != cudaSuccess) return e;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    *grid_out = sms * (per_sm > 0 ? per_sm : 1);
    return cudaSuccess;
}

cudaError_t launch_diffusion_cg(const int64_t* indptr, const int32_t* indices, const float* values, int64_t n,
                                const int32_t* trunc_ids, int64_t rows, int T, int stride, int maxiter, double atol,
                                uint16_t* s_cols, float* s_vals, int grid, float* out, int* bad, cudaStream_t st) {
    CgArgs a{};
    a.indptr = indptr; a.indices = indices; a.values = values; a.trunc_ids = trunc_ids;
    a.n = n; a.rows = rows; a.T = T; a.H = diffusion_cg_hash_size(T);
    int lg = 0;
    while ((1 << lg) < a.H) ++lg;
    a.hash_shift = 32 - lg;
    a.stride = stride; a.maxiter = maxiter; a.atol = atol;
    a.s_cols = s_cols; a.s_vals = s_vals; a.out = out; a.bad = bad;
    const size_t smem = diffusion_cg_smem(T, a.H);
    const int g = (int)(rows < grid ? rows : grid);
    diffusion_cg_kernel<<<g, CG_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace xs
