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

void launch_ids_to_i32(const int64_t* ids64, int32_t* ids32, int64_t count, cudaStream_t st) {
    ids_to_i32_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(ids64, ids32, count);
}

void launch_mutual_knn(const int64_t* ids64, int32_t* ids32, int64_t n, int kd, uint8_t* mutual, cudaStream_t st) {
    const int64_t count = n * kd;
    ids_to_i32_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(ids64, ids32, count);
    mutual_knn_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(ids32, n, kd, mutual);
}

}  // namespace xs
