// Radix-select primitives over 64-bit candidate items (see common.cuh for the item format).
//   block_kth_largest : whole CTA, items enumerated by a caller-supplied functor
//   warp_prune        : one warp trims one in-memory candidate list (used by the GEMM epilogue)
#pragma once
#include "common.cuh"

namespace xs {

// Histogram increment with intra-warp aggregation: lanes that hit the same bin elect one leader,
// so a degenerate digit (all scores share their top byte) costs one shared atomic, not 32.
__device__ __forceinline__ void hist_add(uint32_t* hist, uint32_t bin, bool valid) {
    uint32_t peers = __match_any_sync(0xffffffffu, valid ? bin : 0xffffffffu);
    if (valid && (lane_id() == (uint32_t)(__ffs(peers) - 1))) atomicAdd(&hist[bin], (uint32_t)__popc(peers));
}

// One warp finds, from a 256-bin histogram, the bin holding the k-th largest element counted
// from the top.  Returns (digit, k_within_digit) to all lanes.  Requires sum(hist) >= k >= 1.
__device__ __forceinline__ void warp_pick_digit(const uint32_t* hist, uint32_t k, uint32_t& digit, uint32_t& k_rem) {
    const uint32_t lane = lane_id();
    uint32_t h[8], s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { h[j] = hist[lane * 8 + j]; s += h[j]; }
    uint32_t suf = s;                                   // suffix sum over lanes >= lane
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_down_sync(0xffffffffu, suf, o);
        if (lane + o < 32) suf += t;
    }
    uint32_t ball = __ballot_sync(0xffffffffu, suf >= k);
    uint32_t L = ball ? (31 - __clz(ball)) : 0;         // highest lane whose suffix still reaches k
    uint32_t dg = 0, kr = k;
    if (lane == L) {
        uint32_t above = suf - s;
#pragma unroll
        for (int j = 7; j >= 0; --j) {
            if (above + h[j] >= k) { dg = lane * 8 + j; kr = k - above; break; }
            above += h[j];
        }
    }
    digit = __shfl_sync(0xffffffffu, dg, L);
    k_rem = __shfl_sync(0xffffffffu, kr, L);
}

// Block-wide k-th largest.  `each(fn)` must call fn(item, valid) CONVERGENTLY per warp (all 32
// lanes together; lanes without an item pass valid=false) and visit every item exactly once per
// call.  npass leading bytes are resolved (4 = score key only, 8 = whole item); the returned
// value has the unresolved low bytes cleared.  Caller guarantees total >= k >= 1.
// hist: 256 words of shared memory, misc: 2 words.
template <typename Each>
__device__ uint64_t block_kth_largest(Each each, uint32_t k, int npass, uint32_t* hist, uint32_t* misc) {
    uint64_t prefix = 0, mask = 0;
    for (int pass = 0; pass < npass; ++pass) {
        const int shift = 56 - 8 * pass;
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        each([&](uint64_t it, bool valid) {
            bool in = valid && ((it & mask) == prefix);
            hist_add(hist, (uint32_t)(it >> shift) & 255u, in);
        });
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t dg, kr;
            warp_pick_digit(hist, k, dg, kr);
            if (threadIdx.x == 0) { misc[0] = dg; misc[1] = kr; }
        }
        __syncthreads();
        k = misc[1];
        prefix |= (uint64_t)misc[0] << shift;
        mask |= 0xFFull << shift;
    }
    return prefix;
}

// One warp trims the list `list[0..cnt)` in place.
//   by_count:  keep the `keep` largest items (score-key order; among equal keys the earliest kept)
//   !by_count: find the `keep`-th largest key a, then keep every item whose score is
//              >= key_score(a) - band   (the exactness band of the bf16 pass)
// Returns the new count; thr_key receives the `keep`-th largest key.  whist = 256 words of
// shared memory private to this warp.  Must be called by all 32 lanes.  cnt >= keep >= 1.
__device__ __forceinline__ int warp_prune(uint64_t* list, int cnt, int keep, bool by_count, float band,
                                          uint32_t* whist, uint32_t& thr_key) {
    const uint32_t lane = lane_id();
    uint32_t prefix = 0, mask = 0, k = keep;
    __syncwarp();
#pragma unroll 1
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
#pragma unroll
        for (int j = 0; j < 8; ++j) whist[lane * 8 + j] = 0;
        __syncwarp();
        for (int b = 0; b < cnt; b += 32) {
            int i = b + lane;
            uint32_t key = (i < cnt) ? item_key(list[i]) : 0u;
            bool in = (i < cnt) && ((key & mask) == prefix);
            hist_add(whist, (key >> shift) & 255u, in);
        }
        __syncwarp();
        uint32_t dg, kr;
        warp_pick_digit(whist, k, dg, kr);
        k = kr;
        prefix |= dg << shift;
        mask |= 255u << shift;
        __syncwarp();
    }
    thr_key = prefix;
    uint32_t cut_key = prefix;
    if (!by_count) cut_key = score_key(key_score(prefix) - band);
    int out = 0;
    uint32_t quota = by_count ? k : 0xffffffffu;        // how many items equal to the threshold survive
    for (int b = 0; b < cnt; b += 32) {
        int i = b + lane;
        uint64_t it = (i < cnt) ? list[i] : 0ull;
        uint32_t key = item_key(it);
        bool gt = (i < cnt) && (by_count ? (key > prefix) : (key >= cut_key));
        bool eq = by_count && (i < cnt) && (key == prefix);
        uint32_t eqm = __ballot_sync(0xffffffffu, eq);
        bool take = gt || (eq && (uint32_t)__popc(eqm & lanemask_lt()) < quota);
        uint32_t tm = __ballot_sync(0xffffffffu, take);
        int pos = out + __popc(tm & lanemask_lt());
        __syncwarp();
        if (take) list[pos] = it;
        out += __popc(tm);
        uint32_t ne = __popc(eqm);
        quota = (by_count) ? (quota > ne ? quota - ne : 0u) : quota;
    }
    __syncwarp();
    return out;
}

// Cheaper trim for the common case: min/max of the keys, ONE pass into 256 bins spread over [min, max],
// keep everything from the bin that holds the `keep`-th largest key upwards.  The kept count is >= keep
// and exceeds it by at most that bin's population; when that would leave fewer than `slack` free slots
// (heavy ties) the exact radix trim above takes over.  `band` < 0: trim by count at the bin edge;
// `band` >= 0: keep every item whose score is >= key_score(edge) - band.  thr_key = the edge (<= the
// true keep-th largest key), i.e. nothing at or above it was dropped.
__device__ __forceinline__ int warp_prune_edge(uint64_t* list, int cnt, int keep, float band, int cap, int slack,
                                               uint32_t* whist, uint32_t& thr_key) {
    const uint32_t lane = lane_id();
    __syncwarp();
    uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
    for (int i = lane; i < cnt; i += 32) { const uint32_t key = item_key(list[i]); kmin = min(kmin, key); kmax = max(kmax, key); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    const uint32_t range = kmax - kmin;
    const int shift = (range >> 8) ? (32 - __clz(range) - 8) : 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) whist[lane * 8 + j] = 0;
    __syncwarp();
    for (int i = lane; i < cnt; i += 32) atomicAdd(&whist[(item_key(list[i]) - kmin) >> shift], 1u);
    __syncwarp();
    uint32_t dg, kr;
    warp_pick_digit(whist, (uint32_t)keep, dg, kr);
    // population at or above the edge = keep - kr + whist[dg]
    const uint32_t kept = (uint32_t)keep - kr + whist[dg];
    __syncwarp();
    if (band < 0.f && kept + (uint32_t)slack > (uint32_t)cap)
        return warp_prune(list, cnt, keep, true, 0.f, whist, thr_key);      // ties: exact trim by count
    const uint32_t edge = kmin + (dg << shift);
    thr_key = edge;
    const uint32_t cut_key = (band >= 0.f) ? score_key(key_score(edge) - band) : edge;
    int out = 0;
    for (int b = 0; b < cnt; b += 32) {
        const int i = b + lane;
        const uint64_t it = (i < cnt) ? list[i] : 0ull;
        const bool take = (i < cnt) && item_key(it) >= cut_key;
        const uint32_t tm = __ballot_sync(0xffffffffu, take);
        const int pos = out + __popc(tm & lanemask_lt());
        __syncwarp();
        if (take) list[pos] = it;
        out += __popc(tm);
    }
    __syncwarp();
    return out;
}

}  // namespace xs
