// Full ranking (K == N): the `ranks = np.argsort(-scores, axis=0)` of src/main_retrieve.py:176 in
// mAP mode (src/test_rOP1m.py:147-148), as a segmented, stable LSD radix sort of the exact scores.
//
// One segment per query.  Keys are ~score_key(score) (ascending key = descending score), values the
// row ids in their natural order, so stability yields the documented tie rule (ascending id).
// Four 8-bit passes; per pass: per-tile digit histograms -> per-segment exclusive scan over
// (digit, tile) -> stable scatter (warps own consecutive sub-tiles and walk them in order).
#include "common.cuh"
#include "select.cuh"
#include "internal.h"

namespace xs {

constexpr int SORT_TILE = 4096;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_SUB = SORT_TILE / SORT_WARPS;      // elements per warp, walked 32 at a time

template <bool FIRST>
__device__ __forceinline__ uint32_t sort_key_at(const float* scores, const uint32_t* keys_in, int64_t i) {
    return FIRST ? ~score_key(scores[i]) : keys_in[i];
}

template <bool FIRST>
__global__ void __launch_bounds__(SORT_THREADS)
sort_hist_kernel(const float* __restrict__ scores, int64_t pitch, const uint32_t* __restrict__ keys_in, int64_t n,
                 int shift, uint32_t* __restrict__ block_hist, int nblk) {
    __shared__ uint32_t h[256];
    const int q = blockIdx.y, blk = blockIdx.x;
    h[threadIdx.x] = 0;
    __syncthreads();
    const float* s = scores + (int64_t)q * pitch;
    const uint32_t* kin = keys_in + (int64_t)q * n;
    const int64_t base = (int64_t)blk * SORT_TILE;
    for (int j = threadIdx.x; j < SORT_TILE; j += SORT_THREADS) {
        const int64_t i = base + j;
        const bool valid = i < n;
        const uint32_t key = valid ? sort_key_at<FIRST>(s, kin, i) : 0u;
        hist_add(h, (key >> shift) & 255u, valid);
    }
    __syncthreads();
    block_hist[((int64_t)q * nblk + blk) * 256 + threadIdx.x] = h[threadIdx.x];
}

// grid = nq, 256 threads: thread d turns the per-tile counts of digit d into global output offsets.
__global__ void __launch_bounds__(256)
sort_scan_kernel(uint32_t* __restrict__ block_hist, int nblk) {
    __shared__ uint32_t tot[256];
    const int q = blockIdx.x, d = threadIdx.x;
    uint32_t* h = block_hist + (int64_t)q * nblk * 256;
    uint32_t t = 0;
    for (int b = 0; b < nblk; ++b) t += h[b * 256 + d];
    tot[d] = t;
    __syncthreads();
    if (d == 0) {                                     // 256-entry exclusive scan, serial (negligible)
        uint32_t run = 0;
        for (int i = 0; i < 256; ++i) { uint32_t c = tot[i]; tot[i] = run; run += c; }
    }
    __syncthreads();
    uint32_t run = tot[d];
    for (int b = 0; b < nblk; ++b) { uint32_t c = h[b * 256 + d]; h[b * 256 + d] = run; run += c; }
}

template <bool FIRST>
__global__ void __launch_bounds__(SORT_THREADS)
sort_scatter_kernel(const float* __restrict__ scores, int64_t pitch, const uint32_t* __restrict__ keys_in,
                    const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                    int64_t n, int shift, const uint32_t* __restrict__ block_offs, int nblk) {
    __shared__ uint32_t wcount[SORT_WARPS][256];
    const int q = blockIdx.y, blk = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* s = scores + (int64_t)q * pitch;
    const uint32_t* kin = keys_in + (int64_t)q * n;
    const uint32_t* vin = vals_in + (int64_t)q * n;
    uint32_t* kout = keys_out + (int64_t)q * n;
    uint32_t* vout = vals_out + (int64_t)q * n;
    for (int i = threadIdx.x; i < SORT_WARPS * 256; i += SORT_THREADS) (&wcount[0][0])[i] = 0;
    __syncthreads();
    const int64_t wbase = (int64_t)blk * SORT_TILE + warp * SORT_SUB;
    // 1. digit counts of this warp's sub-tile
    for (int j = lane; j < SORT_SUB; j += 32) {
        const int64_t i = wbase + j;
        const bool valid = i < n;
        const uint32_t key = valid ? sort_key_at<FIRST>(s, kin, i) : 0u;
        hist_add(wcount[warp], (key >> shift) & 255u, valid);
    }
    __syncthreads();
    // 2. per (warp, digit) output base = tile offset of the digit + counts of the warps before
    {
        const int d = threadIdx.x;
        uint32_t run = block_offs[((int64_t)q * nblk + blk) * 256 + d];
        for (int w = 0; w < SORT_WARPS; ++w) { uint32_t c = wcount[w][d]; wcount[w][d] = run; run += c; }
    }
    __syncthreads();
    // 3. stable scatter: 32 elements at a time, in order
    for (int j0 = 0; j0 < SORT_SUB; j0 += 32) {
        const int64_t i = wbase + j0 + lane;
        const bool valid = i < n;
        const uint32_t key = valid ? sort_key_at<FIRST>(s, kin, i) : 0u;
        const uint32_t val = valid ? (FIRST ? (uint32_t)i : vin[i]) : 0u;
        const uint32_t dg = (key >> shift) & 255u;
        const uint32_t peers = __match_any_sync(0xffffffffu, valid ? dg : 0xffffffffu);
        const uint32_t rank = __popc(peers & lanemask_lt());
        if (valid) {
            const uint32_t pos = wcount[warp][dg] + rank;
            kout[pos] = key;
            vout[pos] = val;
        }
        __syncwarp();
        if (valid && rank == 0) wcount[warp][dg] += __popc(peers);
        __syncwarp();
    }
}

// ranks32 [c][n] (sorted row ids), keys [c][n] (sorted ~score keys) -> out_ranks[n][nq_total] (+ id_offset), column q0 + j
__global__ void rank_transpose_kernel(const uint32_t* __restrict__ vals, const uint32_t* __restrict__ keys, int64_t n, int c,
                                      int q0, int nq_total, int64_t id_offset, int64_t* __restrict__ out_ranks,
                                      float* __restrict__ out_scores) {
    __shared__ uint32_t tv[32][33], tk[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int j0 = blockIdx.y * 32;
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
        const int j = j0 + jj;
        const int64_t r = r0 + threadIdx.x;
        const bool ok = j < c && r < n;
        tv[jj][threadIdx.x] = ok ? vals[(int64_t)j * n + r] : 0u;
        tk[jj][threadIdx.x] = ok ? keys[(int64_t)j * n + r] : 0u;
    }
    __syncthreads();
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
        const int64_t r = r0 + rr;
        const int j = j0 + threadIdx.x;
        if (r < n && j < c) {
            out_ranks[r * nq_total + q0 + j] = (int64_t)tv[threadIdx.x][rr] + id_offset;
            if (out_scores) out_scores[r * nq_total + q0 + j] = key_score(~tk[threadIdx.x][rr]);
        }
    }
}

size_t rank_all_work_bytes(int c, int64_t n) {
    const int nblk = (int)((n + SORT_TILE - 1) / SORT_TILE);
    return (size_t)c * n * 4 * 4 + (size_t)c * nblk * 256 * 4;
}

// scores: [c][pitch] exact fp32 scores.  work: rank_all_work_bytes(c, n).  Output columns q0..q0+c of
// out_ranks [n][nq_total] int64 / out_scores [n][nq_total] fp32 (device).
void launch_rank_all(const float* scores, int64_t pitch, int c, int64_t n, int q0, int nq_total, int64_t id_offset,
                     void* work, int64_t* out_ranks, float* out_scores, cudaStream_t st) {
    const int nblk = (int)((n + SORT_TILE - 1) / SORT_TILE);
    uint32_t* ka = static_cast<uint32_t*>(work);
    uint32_t* va = ka + (size_t)c * n;
    uint32_t* kb = va + (size_t)c * n;
    uint32_t* vb = kb + (size_t)c * n;
    uint32_t* bh = vb + (size_t)c * n;
    dim3 grid((unsigned)nblk, (unsigned)c);
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 8 * pass;
        const uint32_t* kin = (pass & 1) ? kb : ka;   // pass 0 reads the scores; 1 reads b; 2 reads a; 3 reads b
        const uint32_t* vin = (pass & 1) ? vb : va;
        uint32_t* kout = (pass & 1) ? ka : kb;
        uint32_t* vout = (pass & 1) ? va : vb;
        if (pass == 0) {
            sort_hist_kernel<true><<<grid, SORT_THREADS, 0, st>>>(scores, pitch, nullptr, n, shift, bh, nblk);
            sort_scan_kernel<<<(unsigned)c, 256, 0, st>>>(bh, nblk);
            sort_scatter_kernel<true><<<grid, SORT_THREADS, 0, st>>>(scores, pitch, nullptr, nullptr, kout, vout, n, shift, bh, nblk);
        } else {
            sort_hist_kernel<false><<<grid, SORT_THREADS, 0, st>>>(nullptr, 0, kin, n, shift, bh, nblk);
            sort_scan_kernel<<<(unsigned)c, 256, 0, st>>>(bh, nblk);
            sort_scatter_kernel<false><<<grid, SORT_THREADS, 0, st>>>(nullptr, 0, kin, vin, kout, vout, n, shift, bh, nblk);
        }
    }
    // after 4 passes the result sits in (ka, va)
    dim3 tg((unsigned)((n + 31) / 32), (unsigned)((c + 31) / 32)), tb(32, 8);
    rank_transpose_kernel<<<tg, tb, 0, st>>>(va, ka, n, c, q0, nq_total, id_offset, out_ranks, out_scores);
}

}  // namespace xs
