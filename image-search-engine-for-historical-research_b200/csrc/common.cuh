// Shared device helpers for the exhaustive-matching kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace xs {

// ----------------------------------------------------------------------------------------------
// Candidate items.  A candidate is one 64-bit word whose unsigned order is the ranking order:
//   bits 63..32  order-preserving image of the fp32 score (bigger = better)
//   bits 31..0   0xFFFFFFFF - row   (so that on equal scores the LOWER row id is the bigger item)
// ----------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t score_key(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float key_score(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_item(float score, uint32_t row) {
    return (static_cast<uint64_t>(score_key(score)) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - row);
}
__host__ __device__ __forceinline__ uint32_t item_row(uint64_t it) { return 0xFFFFFFFFu - static_cast<uint32_t>(it); }
__host__ __device__ __forceinline__ uint32_t item_key(uint64_t it) { return static_cast<uint32_t>(it >> 32); }

// Pool addressing.  Stage 1 (scan / GEMM) leaves, for every query q and partial list p, up to
// `cap` items.  The layout is the one the GEMM epilogue wants (a query is a TMEM lane, 128 lanes
// per query tile):   slot(q, p) = ((q / 128) * P + p) * 128 + (q % 128)
__host__ __device__ __forceinline__ int64_t pool_slot(int64_t q, int p, int P) {
    return ((q >> 7) * P + p) * 128 + (q & 127);
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m;
}

// 16-byte streaming load that does not allocate in L1 (database rows are read exactly once).
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_stream_f4(const void* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
// bf16 pair packed in a 32-bit word -> two fp32 (exact: bf16 is the top half of fp32).
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Programmatic dependent launch: a kernel launched with launch_pdl() may start (and run its prologue)
// while its predecessor in the stream is still finishing; pdl_wait() blocks until the predecessor has
// completed and its writes are visible.  EVERY kernel of a chain calls it before its first global access,
// so completion order along the chain is preserved.  Without the launch attribute it is a no-op.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

}  // namespace xs
