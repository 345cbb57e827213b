// Mutual-kNN test of the diffusion affinity graph (src/utils/diffusion.py:101-116):
//   ismutual[i][j] = (i is among the kd neighbours of ids[i][j]),  j >= 1   (slot 0 is the row itself, :108)
// The reference evaluates it with a Python loop over rows and np.isin over kd x kd ids -- O(N kd^2) in
// the interpreter, the reason the method is capped at N < 120 000 (Reranking.py:212).  Here: one warp per
// row, lanes over the neighbour slots, each lane scanning its neighbour's id list (int32, L2-resident for
// the reference's sizes).
#include "common.cuh"
#include "internal.h"

namespace xs {

__global__ void ids_to_i32_kernel(const int64_t* __restrict__ in, int32_t* __restrict__ out, int64_t count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = (int32_t)in[i];
}

__global__ void __launch_bounds__(256)
mutual_knn_kernel(const int32_t* __restrict__ ids, int64_t n, int kd, uint8_t* __restrict__ mutual) {
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int32_t* mine = ids + i * kd;
    for (int j = lane_id(); j < kd; j += 32) {
        bool found = false;
        const int32_t nb = mine[j];
        if (j > 0 && nb >= 0 && nb < n) {
            const int32_t* theirs = ids + (int64_t)nb * kd;
            for (int t = 0; t < kd; ++t) found |= (theirs[t] == (int32_t)i);
        }
        mutual[i * kd + j] = found ? 1 : 0;
    }
}

// same test on kNN lists that are the leading kd columns of wider rows (the device-resident pipeline keeps n_trunc columns)
__global__ void __launch_bounds__(256)
mutual_knn_strided_kernel(const int32_t* __restrict__ ids, int64_t pitch, int64_t n, int kd, uint8_t* __restrict__ mutual) {
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int32_t* mine = ids + i * pitch;
    for (int j = lane_id(); j < kd; j += 32) {
        bool found = false;
        const int32_t nb = mine[j];
        if (j > 0 && nb >= 0 && nb < n) {
            const int32_t* theirs = ids + (int64_t)nb * pitch;
            for (int t = 0; t < kd; ++t) found |= (theirs[t] == (int32_t)i);
        }
        mutual[i * kd + j] = found ? 1 : 0;
    }
}

void launch_ids_to_i32(const int64_t* ids64, int32_t* ids32, int64_t count, cudaStream_t st) {
    ids_to_i32_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(ids64, ids32, count);
}

void launch_mutual_knn(const int64_t* ids64, int32_t* ids32, int64_t n, int kd, uint8_t* mutual, cudaStream_t st) {
    const int64_t count = n * kd;
    ids_to_i32_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(ids64, ids32, count);
    mutual_knn_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(ids32, n, kd, mutual);
}

}  // namespace xs

// ---- affinity, degrees and the normalised Laplacian on the device (src/utils/diffusion.py:87-116) ------------------
// The reference assembles these with scipy on the host:
//     sims[sims < 0] = 0; A[i, ids[i][j]] = sims[i][j] ** gamma  for mutual slots j >= 1        (get_affinity, :101-116)
//     degrees = A @ ones + 1e-12;  L = I - alpha * D^-1/2 A D^-1/2                              (get_laplacian, :87-98)
// with float32 matrices throughout (dia_matrix(..., dtype=np.float32)), i.e. every entry of L is
//     -( f32(alpha) * ( (f32(deg_i^-1/2) * a_ij) * f32(deg_c^-1/2) ) )       each product rounded to float32,
// degrees summed in float64.  Kept in that exact order here so that the device graph equals the reference's bit for bit.
namespace xs {

__global__ void __launch_bounds__(256)
affinity_kernel(const float* __restrict__ sims, int64_t sim_pitch, const uint8_t* __restrict__ mutual, int64_t n, int kd, double gamma,
                float* __restrict__ aff, float* __restrict__ dinv) {
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int ig = (int)gamma;
    const bool integral = (double)ig == gamma && ig >= 1 && ig <= 8;
    double deg = 0.0;
    for (int j = lane_id(); j < kd; j += 32) {
        float a = 0.f;
        if (mutual[i * kd + j]) {
            const float s = fmaxf(sims[i * sim_pitch + j], 0.f);
            double p;
            if (integral) { p = s; for (int e = 1; e < ig; ++e) p *= (double)s; }     // exact product, one rounding: what a correctly rounded powf gives
            else p = pow((double)s, gamma);
            a = (float)p;
        }
        aff[i * kd + j] = a;
        deg += (double)a;
    }
    deg = warp_sum(deg);
    if (lane_id() == 0) dinv[i] = (float)(1.0 / sqrt(deg + 1e-12));
}

// ELL form of the Laplacian: row i holds cnt[i] <= kd entries (diagonal first), the rest is padding (column INT32_MIN).
__global__ void __launch_bounds__(256)
laplacian_kernel(const float* __restrict__ aff, const float* __restrict__ dinv, const int32_t* __restrict__ ids, int64_t id_pitch,
                 int64_t n, int kd, float alpha, int32_t* __restrict__ cols, float* __restrict__ vals, int32_t* __restrict__ cnt) {
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int lane = lane_id();
    const float di = dinv[i];
    int c = 1;
    if (lane == 0) { cols[i * kd] = (int32_t)i; vals[i * kd] = 1.0f; }
    for (int base = 1; base < kd; base += 32) {
        const int j = base + lane;
        float a = 0.f;
        int32_t col = 0;
        if (j < kd) { a = aff[i * kd + j]; col = ids[i * id_pitch + j]; }
        const bool on = (j < kd) && a != 0.f && col >= 0 && col < n;
        const uint32_t m = __ballot_sync(0xffffffffu, on);
        if (on) {
            const int pos = c + __popc(m & lanemask_lt());
            const float v = __fmul_rn(alpha, __fmul_rn(__fmul_rn(di, a), dinv[col]));
            cols[i * kd + pos] = col;
            vals[i * kd + pos] = -v;
        }
        c += __popc(m);
    }
    for (int j = c + lane; j < kd; j += 32) { cols[i * kd + j] = INT32_MIN; vals[i * kd + j] = 0.f; }
    if (lane == 0) cnt[i] = c;
}

__global__ void ell_indptr_kernel(int64_t* indptr, int64_t n, int kd) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) indptr[i] = i * kd;
}

__global__ void ids_strided_to_i32_kernel(const int64_t* __restrict__ in, int64_t in_pitch, int32_t* __restrict__ out, int64_t out_pitch,
                                          int64_t rows, int cols, int64_t sub) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * cols) return;
    const int64_t r = t / cols; const int c = (int)(t - r * cols);
    out[r * out_pitch + c] = (int32_t)(in[r * in_pitch + c] - sub);
}

void launch_ids_strided_to_i32(const int64_t* in, int64_t in_pitch, int32_t* out, int64_t out_pitch, int64_t rows, int cols, int64_t sub, cudaStream_t st) {
    const int64_t total = rows * cols;
    if (total <= 0) return;
    ids_strided_to_i32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, in_pitch, out, out_pitch, rows, cols, sub);
}

// ids32: [n][id_pitch] (first kd columns used), sims: [n][sim_pitch].  Scratch: mutual [n*kd] bytes, aff [n*kd] f32, dinv [n] f32.
void launch_diffusion_graph(const int32_t* ids32, int64_t id_pitch, const float* sims, int64_t sim_pitch, int64_t n, int kd, double alpha, double gamma,
                            uint8_t* mutual, float* aff, float* dinv, int32_t* cols, float* vals, int32_t* cnt, int64_t* indptr, cudaStream_t st) {
    const unsigned rows_grid = (unsigned)((n + 7) / 8);
    mutual_knn_strided_kernel<<<rows_grid, 256, 0, st>>>(ids32, id_pitch, n, kd, mutual);
    affinity_kernel<<<rows_grid, 256, 0, st>>>(sims, sim_pitch, mutual, n, kd, gamma, aff, dinv);
    laplacian_kernel<<<rows_grid, 256, 0, st>>>(aff, dinv, ids32, id_pitch, n, kd, (float)alpha, cols, vals, cnt);
    if (indptr) ell_indptr_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(indptr, n, kd);
}

}  // namespace xs
