// C ABI (include/xs_b200.h) and host-side orchestration of the exhaustive matcher.
//
// Data in HBM per index:  db32 [n][d_pad] fp32 (exact operand), db16t = the rotated rows in bf16, tiled
// [n_pad/256][d_pad/64][256][64] (coarse operand of the GEMM and of the batch-1 scan; n_pad = multiple of 256 so TMA
// boxes never straddle the end; the row-major form db16 only with "compact" 0), a grow-only workspace
// (query copies, score rows, candidate pools, results).  One search = prep_queries -> coarse
// kernel (batch-1 scan | tcgen05 GEMM with fused top-K) -> finalise (exact rescoring + sort),
// all on one stream; uncertified queries are re-run on the exact fp32 path.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <fcntl.h>
#include <unistd.h>

#include "../../include/xs_b200.h"
#include "common.cuh"
#include "internal.h"

using namespace xs;

// ---- errors ------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}
#define CU_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { cudaGetLastError(); \
    return fail(e_ == cudaErrorMemoryAllocation ? XS_ERR_NOMEM : XS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); } } while (0)
#define XS_TRY(expr) do { int r_ = (expr); if (r_ != XS_OK) return r_; } while (0)

// ---- workspace buffers -------------------------------------------------------------------------------
struct Buf {
    void* p = nullptr; size_t cap = 0;
    int ensure(size_t need) {
        if (need <= cap) return XS_OK;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = need + need / 8;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { cudaGetLastError(); e = cudaMalloc(&p, need); want = need; }
        if (e != cudaSuccess) { cudaGetLastError(); p = nullptr; return fail(XS_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e)); }
        cap = want;
        return XS_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinnedBuf {
    void* p = nullptr; size_t cap = 0;
    int ensure(size_t need) {
        if (need <= cap) return XS_OK;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMallocHost(&p, need + need / 4);
        if (e != cudaSuccess) { cudaGetLastError(); p = nullptr; return fail(XS_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", need, cudaGetErrorString(e)); }
        cap = need + need / 4;
        return XS_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// The database arrays of an index, shared by the index and its clones (xs_index_clone): freed with the last holder.
struct DbShare {
    std::mutex mu; int refs = 1;
    __nv_bfloat16* db16 = nullptr; __nv_bfloat16* db16t = nullptr; float* db32 = nullptr; void* dstats = nullptr;
};

struct xs_index {
    DbShare* share = nullptr;
    int device = 0; int num_sms = 0;
    int64_t n = 0, n_pad = 0, id_offset = 0;
    int d = 0, d_pad = 0;
    __nv_bfloat16* db16 = nullptr; float* db32 = nullptr; DevStats* dstats = nullptr;
    __nv_bfloat16* db16t = nullptr;              // tiled twin of db16 for the GEMM's database operand (optional)
    CUtensorMap tmap_db_b, tmap_db_a;            // db16 as GEMM operand B (box 256 rows) / A (box 128 rows, self-kNN)
    CUtensorMap tmap_dbt_b, tmap_dbt_h;          // tiled twin: boxes of 256 / 128 rows, each one contiguous run
    CUtensorMap tmap_q_cached; const void* tmap_q_base = nullptr; int64_t tmap_q_rows = 0;   // query operand map of the last GEMM call
    // tunables
    float eps_sigmas = 8.f; int scan_max_q = 1; int force_path = 0; int gemm_splits = 0; int sample_pass = 1; int pair_mode = 1; int timing = 0;
    int eps_mode = 0;                             // certificate: 0 = statistical band (8 sigma, random rotation, model check), 1 = worst-case band
    int inline_boot = 1;                          // small batches: threshold bootstrap inside the GEMM launch (0: separate sample pass)
    int half_units = 1;                           // single-CTA GEMM shapes: deal the database in half tiles (0: whole tiles)
    int gemm_stages = 4;                          // operand ring of the single-CTA GEMM shape: 4 stages, or 3 to leave room for a co-resident finalise CTA
    int scan_tiled = 1;                           // batch-1 scan reads the tiled twin when there is one (0: the row-major copy, while it is kept)
    int fin_per_sm = 0;                           // cluster finalise: 0 = latency mode (three CTAs per SM), 1 = one slim CTA per SM, resident next to a GEMM CTA
    bool rotate = true; uint32_t rot_seed = 0;    // random rotation applied before bf16 rounding (fixed at build time)
    uint32_t boot_arrived = 0, boot_published = 0, boot_epoch = 0;    // host mirrors of the in-kernel bootstrap's counters / epoch
    int self_lanes = 1;                           // xs_self_knn: 2 = batches alternate between this index and an internal clone
                                                  // (measured: 1.03 s either way at 500k x 500k -- the loop is tensor/power bound)
    xs_index* self_lane = nullptr;                // that clone (created on first use, freed with the index)
    // workspace
    Buf boot_samp, boot_sync, boot_trace, q32r, sbound;
    int boot_trace_on = 0, boot_trace_grid = 0;
    Buf fin_trace; int fin_trace_rows = 0;
    Buf fin_work, fin_ticket, aqe_ids, q_raw, q32, q16, eps, thr0, ghist, sort_work, rank_out, rank_scores, scores, pool_items, pool_count, pool_thr, status, ncand, out_idx, out_score, stage, stage16;
    PinnedBuf h_idx, h_score, h_status;           // pinned landing zone of the host API (one sync per call)
    PinnedBuf h_aqe;                              // pinned staging of xs_aqe_search's id lists
    cudaStream_t stream = nullptr;                // the index's own stream (host API, build)
    cudaStream_t cur = nullptr;                   // stream of the call in progress (the caller's for *_dev entry points)
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool ev_valid = false;
    std::mutex mu;
    xs_stats stats{};
    int64_t bytes = 0;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int get_encoder() {
    if (g_encode) return XS_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(XS_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return XS_OK;
}

// 2-D bf16 tensor [rows][d_pad], box {64 columns, box_rows}, 128-byte swizzle (what umma_desc_sw128 expects).
static int make_tmap(CUtensorMap* map, const void* base, int64_t rows, int d_pad, int box_rows) {
    XS_TRY(get_encoder());
    cuuint64_t dims[2] = {(cuuint64_t)d_pad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)d_pad * 2};
    cuuint32_t box[2] = {(cuuint32_t)GEMM_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    // 256-byte L2 promotion: measured 0.603 ms per database pass against 0.66 ms with 128 B / 64 B / none
    const CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(XS_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return XS_OK;
}

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---- process-wide defaults picked up by xs_index_create* --------------------------------------------------
static std::mutex g_cfg_mu;
static int g_rotate = -1;                         // -1: not set -> on unless XS_NO_ROTATE=1
static uint32_t g_rot_seed = 0x5EEDB200u;
static int g_compact = -1;                        // -1: not set -> on unless XS_KEEP_ROWMAJOR=1
static bool default_compact() {
    std::lock_guard<std::mutex> lk(g_cfg_mu);
    if (g_compact < 0) { const char* e = getenv("XS_KEEP_ROWMAJOR"); g_compact = (e && atoi(e)) ? 0 : 1; }
    return g_compact != 0;
}
static void default_rotation(bool* on, uint32_t* seed) {
    std::lock_guard<std::mutex> lk(g_cfg_mu);
    if (g_rotate < 0) { const char* e = getenv("XS_NO_ROTATE"); g_rotate = (e && atoi(e)) ? 0 : 1; }
    *on = g_rotate != 0; *seed = g_rot_seed;
}
extern "C" int xs_config_set(const char* name, double value) {
    if (!name) return fail(XS_ERR_ARG, "null pointer");
    std::lock_guard<std::mutex> lk(g_cfg_mu);
    if (!strcmp(name, "rotation")) g_rotate = value != 0.0 ? 1 : 0;
    else if (!strcmp(name, "rotation_seed")) g_rot_seed = (uint32_t)(uint64_t)value;
    else if (!strcmp(name, "compact")) g_compact = value != 0.0 ? 1 : 0;
    else return fail(XS_ERR_ARG, "unknown configuration key '%s'", name);
    return XS_OK;
}

// ---- ABI: misc -----------------------------------------------------------------------------------------
extern "C" const char* xs_last_error(void) { return g_err.c_str(); }
extern "C" int xs_abi_version(void) { return 2; }
extern "C" int xs_device_count(int* count) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { cudaGetLastError(); if (count) *count = 0; return fail(XS_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    if (count) *count = c;
    return XS_OK;
}

// ---- index lifecycle -------------------------------------------------------------------------------------
static int index_alloc(xs_index* ix, int64_t n, int d, int device, int64_t id_offset, bool rowmajor16 = true) {
    if (n <= 0 || d <= 0) return fail(XS_ERR_ARG, "empty database (n=%lld, d=%d)", (long long)n, d);
    if (n >= (int64_t)0xFFFFFF00u) return fail(XS_ERR_UNSUPPORTED, "more than 2^32-256 rows per index; shard the database");
    // the rescoring kernels stage eight fp32 rows per CTA in shared memory and the row kernels one: 4096 columns is what fits
    if (d > 4096) return fail(XS_ERR_UNSUPPORTED, "d=%d > 4096 columns is not supported by the rescoring kernels", d);
    CU_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(XS_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    ix->device = device; ix->num_sms = prop.multiProcessorCount;
    ix->n = n; ix->d = d; ix->id_offset = id_offset;
    ix->d_pad = (int)round_up(d, COL_ALIGN);
    ix->n_pad = round_up(n, ROW_ALIGN);
    default_rotation(&ix->rotate, &ix->rot_seed);
    const size_t b32 = (size_t)n * ix->d_pad * sizeof(float), b16 = (size_t)ix->n_pad * ix->d_pad * 2;
    CU_TRY(cudaMalloc(&ix->db32, b32));
    if (rowmajor16) CU_TRY(cudaMalloc(&ix->db16, b16));
    CU_TRY(cudaMalloc(&ix->dstats, sizeof(DevStats)));
    ix->share = new DbShare();
    ix->share->db16 = ix->db16; ix->share->db32 = ix->db32; ix->share->dstats = ix->dstats;
    ix->bytes = (int64_t)(b32 + (rowmajor16 ? b16 : 0));
    CU_TRY(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
    for (auto& e : ix->ev) CU_TRY(cudaEventCreate(&e));
    CU_TRY(cudaMemsetAsync(ix->dstats, 0, sizeof(DevStats), ix->stream));
    // zero the padding rows of the bf16 copy (they are read by TMA boxes and by the scan's row groups)
    if (!rowmajor16) return XS_OK;
    if (ix->n_pad > n) CU_TRY(cudaMemsetAsync(ix->db16 + (size_t)n * ix->d_pad, 0, (size_t)(ix->n_pad - n) * ix->d_pad * 2, ix->stream));
    XS_TRY(make_tmap(&ix->tmap_db_b, ix->db16, ix->n_pad, ix->d_pad, GEMM_BN));
    XS_TRY(make_tmap(&ix->tmap_db_a, ix->db16, ix->n_pad, ix->d_pad, GEMM_BM));
    return XS_OK;
}

static void index_free(xs_index* ix) {
    if (!ix) return;
    if (ix->self_lane) { index_free(ix->self_lane); ix->self_lane = nullptr; }
    cudaSetDevice(ix->device);
    if (ix->stream) cudaStreamSynchronize(ix->stream);
    for (Buf* b : {&ix->boot_samp, &ix->boot_sync, &ix->boot_trace, &ix->fin_trace, &ix->q32r, &ix->sbound, &ix->fin_work, &ix->fin_ticket, &ix->aqe_ids, &ix->q_raw, &ix->q32, &ix->q16, &ix->eps, &ix->thr0, &ix->ghist, &ix->sort_work, &ix->rank_out, &ix->rank_scores, &ix->scores, &ix->pool_items, &ix->pool_count, &ix->pool_thr,
                   &ix->status, &ix->ncand, &ix->out_idx, &ix->out_score, &ix->stage, &ix->stage16}) b->release();
    ix->h_idx.release(); ix->h_score.release(); ix->h_status.release(); ix->h_aqe.release();
    if (ix->share) {
        bool last;
        { std::lock_guard<std::mutex> lk(ix->share->mu); last = (--ix->share->refs == 0); }
        if (last) {
            if (ix->share->db16) cudaFree(ix->share->db16);
            if (ix->share->db16t) cudaFree(ix->share->db16t);
            if (ix->share->db32) cudaFree(ix->share->db32);
            if (ix->share->dstats) cudaFree(ix->share->dstats);
            delete ix->share;
        }
    } else {                                      // allocation failed before the share existed
        if (ix->db16) cudaFree(ix->db16);
        if (ix->db32) cudaFree(ix->db32);
        if (ix->dstats) cudaFree(ix->dstats);
    }
    for (auto& e : ix->ev) if (e) cudaEventDestroy(e);
    if (ix->stream) cudaStreamDestroy(ix->stream);
    cudaGetLastError();
    delete ix;
}

// 2-D map over the tiled twin viewed as [n_pad * KB rows][64 columns] (stride 128 B): box {64, box_rows}.
static int make_tmap_tiled(CUtensorMap* map, const void* base, int64_t n_pad, int d_pad, int box_rows) {
    XS_TRY(get_encoder());
    cuuint64_t dims[2] = {(cuuint64_t)GEMM_BK, (cuuint64_t)n_pad * (cuuint64_t)(d_pad / GEMM_BK)};
    cuuint64_t strides[1] = {(cuuint64_t)GEMM_BK * 2};
    cuuint32_t box[2] = {(cuuint32_t)GEMM_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(XS_ERR_CUDA, "cuTensorMapEncodeTiled (tiled twin) failed with CUresult %d", (int)r);
    return XS_OK;
}

// After all rows are in place: the tiled twin of the bf16 copy (skipped with XS_NO_TILED=1 or when memory is short).
static int build_tiled_twin(xs_index* ix) {
    if (const char* e = getenv("XS_NO_TILED")) if (atoi(e)) return XS_OK;
    const size_t b16 = (size_t)ix->n_pad * ix->d_pad * 2;
    if (cudaMalloc(&ix->db16t, b16) != cudaSuccess) { cudaGetLastError(); ix->db16t = nullptr; return XS_OK; }   // optional: fall back to the row-major maps
    ix->share->db16t = ix->db16t;
    launch_tile_db16(ix->db16, ix->db16t, 0, ix->n_pad, ix->d_pad, ix->stream);
    CU_TRY(cudaStreamSynchronize(ix->stream));
    XS_TRY(make_tmap_tiled(&ix->tmap_dbt_b, ix->db16t, ix->n_pad, ix->d_pad, GEMM_BN));
    XS_TRY(make_tmap_tiled(&ix->tmap_dbt_h, ix->db16t, ix->n_pad, ix->d_pad, GEMM_BM));
    ix->bytes += (int64_t)b16;
    // Compact index (the default): every reader of the bf16 copy -- the GEMM's TMA boxes and the batch-1 scan -- now has the
    // tiled form, so the row-major one goes back to the allocator: 2 * d_pad bytes per row, 4.1 GB per million 2048-d rows.
    if (default_compact()) {
        CU_TRY(cudaFree(ix->db16));
        ix->db16 = nullptr; ix->share->db16 = nullptr;
        ix->bytes -= (int64_t)b16;
    }
    return XS_OK;
}

// Compact build: no row-major bf16 copy is ever allocated.  The tiled array is allocated up front (padding rows of the last
// 256-row block zeroed); each staged chunk's bf16 rows land in a chunk-sized scratch and are tiled from there.
static bool want_compact_build() {
    const char* e = getenv("XS_NO_TILED");
    return !(e && atoi(e)) && default_compact();
}
static int alloc_tiled(xs_index* ix) {
    const size_t b16 = (size_t)ix->n_pad * ix->d_pad * 2;
    if (cudaMalloc(&ix->db16t, b16) != cudaSuccess) { cudaGetLastError(); ix->db16t = nullptr; return fail(XS_ERR_NOMEM, "no memory for the bf16 database (%zu bytes)", b16); }
    ix->share->db16t = ix->db16t;
    ix->bytes += (int64_t)b16;
    if (ix->n_pad > ix->n) {
        const size_t block = (size_t)ROW_ALIGN * ix->d_pad * 2;          // one 256-row block, all its k-blocks: contiguous
        CU_TRY(cudaMemsetAsync(reinterpret_cast<char*>(ix->db16t) + b16 - block, 0, block, ix->stream));
    }
    XS_TRY(make_tmap_tiled(&ix->tmap_dbt_b, ix->db16t, ix->n_pad, ix->d_pad, GEMM_BN));
    XS_TRY(make_tmap_tiled(&ix->tmap_dbt_h, ix->db16t, ix->n_pad, ix->d_pad, GEMM_BM));
    return XS_OK;
}
// fp32 rows r0 .. r0 + rows are in place in db32: finish them (normalise / rotate / statistics / bf16) into the layout(s) kept.
static int finish_chunk(xs_index* ix, int64_t r0, int64_t rows, bool renormalise) {
    if (ix->db16) {
        launch_finish_rows(ix->db32 + (size_t)r0 * ix->d_pad, ix->db16 + (size_t)r0 * ix->d_pad, rows, ix->d_pad, renormalise, ix->rotate, ix->rot_seed, ix->dstats, ix->stream);
        return XS_OK;
    }
    XS_TRY(ix->stage16.ensure((size_t)rows * ix->d_pad * 2));
    launch_finish_rows(ix->db32 + (size_t)r0 * ix->d_pad, ix->stage16.as<__nv_bfloat16>(), rows, ix->d_pad, renormalise, ix->rotate, ix->rot_seed, ix->dstats, ix->stream);
    launch_tile_db16(ix->stage16.as<__nv_bfloat16>(), ix->db16t, r0, rows, ix->d_pad, ix->stream);
    return XS_OK;
}

// Copies `rows` rows starting at r0 of a strided host matrix into a dense device tile.
//   column-major source (stride_row == 1): tile is [d][rows]   (pitch = rows)
//   row-major source    (stride_col == 1): tile is [rows][d]   (pitch = d)
static int stage_host_rows(const void* src, int dtype, bool colmajor, int64_t stride, int64_t r0, int64_t rows,
                           int d, void* tile, cudaStream_t st) {
    const size_t es = dtype == XS_F64 ? 8 : 4;
    const char* s = static_cast<const char*>(src);
    if (colmajor)   // stride = elements between consecutive columns
        CU_TRY(cudaMemcpy2DAsync(tile, (size_t)rows * es, s + (size_t)r0 * es, (size_t)stride * es, (size_t)rows * es, (size_t)d,
                                 cudaMemcpyHostToDevice, st));
    else            // stride = elements between consecutive rows
        CU_TRY(cudaMemcpy2DAsync(tile, (size_t)d * es, s + (size_t)r0 * stride * es, (size_t)stride * es, (size_t)d * es, (size_t)rows,
                                 cudaMemcpyHostToDevice, st));
    return XS_OK;
}

static int check_layout(int dtype, int64_t rows, int d, int64_t stride_row, int64_t stride_col, bool* colmajor) {
    if (dtype != XS_F32 && dtype != XS_F64) return fail(XS_ERR_ARG, "dtype must be XS_F32 or XS_F64");
    if (stride_col == 1 && stride_row >= d) { *colmajor = false; return XS_OK; }
    if (stride_row == 1 && stride_col >= rows) { *colmajor = true; return XS_OK; }
    if (rows == 1 && stride_col == 1) { *colmajor = false; return XS_OK; }
    return fail(XS_ERR_ARG, "unsupported strides (row %lld, col %lld): pass a row-major matrix or the F-order view vecs.T",
                (long long)stride_row, (long long)stride_col);
}

// ---- on-disk image of an index: what lives in HBM, byte for byte -------------------------------------------------------
// File = 4 KB header | db32 [n][d_pad] fp32 | db16 [n_pad][d_pad] bf16 | db16t (same bytes, tiled) -- sections 4 KB
// aligned.  Loading is then a straight upload: reader threads pread() slices of the file into a ring of pinned buffers and
// every slice goes out with its own cudaMemcpyAsync as soon as it is staged; no transpose, no conversion, no tiling kernel
// (the reference unpickles, concatenates and transposes on every start, src/online.py:93-102).
namespace {
struct IndexFileHeader {
    char magic[8];               // "XSB200\0\1"
    uint32_t version, flags;     // flags: 1 rotate, 2 has tiled twin
    int64_t n, n_pad;
    int32_t d, d_pad;
    uint32_t rot_seed, pad0;
    DevStats stats; uint32_t pad1;
    uint64_t off32, off16, off16t, bytes32, bytes16;
};
constexpr size_t FILE_ALIGN = 4096;
constexpr size_t STAGE_BYTES_IO = (size_t)32 << 20;
constexpr int STAGE_RING = 4, IO_THREADS = 8;

struct StageRing {               // pinned staging shared by save / load (one transfer at a time)
    std::mutex mu;
    void* buf[STAGE_RING] = {};
    cudaEvent_t ev[STAGE_RING] = {};
    bool ready = false;
    int ensure() {
        if (ready) return XS_OK;
        for (int i = 0; i < STAGE_RING; ++i) {
            CU_TRY(cudaMallocHost(&buf[i], STAGE_BYTES_IO));
            CU_TRY(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        }
        ready = true;
        return XS_OK;
    }
};
StageRing g_ring;

// fn(part_offset, part_bytes, dst/src pointer inside the pinned buffer) on IO_THREADS threads
template <typename F>
bool parallel_parts(size_t bytes, char* pinned, F fn) {
    std::vector<std::thread> th;
    std::vector<int> ok(IO_THREADS, 1);
    const size_t part = ((bytes + IO_THREADS - 1) / IO_THREADS + 4095) & ~(size_t)4095;
    for (int t = 0; t < IO_THREADS; ++t) {
        const size_t lo = (size_t)t * part;
        if (lo >= bytes) break;
        const size_t nb = std::min(part, bytes - lo);
        th.emplace_back([&, t, lo, nb] { ok[t] = fn(lo, nb, pinned + lo) ? 1 : 0; });
    }
    for (auto& x : th) x.join();
    for (int v : ok) if (!v) return false;
    return true;
}

bool pread_all(int fd, void* dst, size_t nb, off_t off) {
    char* p = static_cast<char*>(dst);
    while (nb) { const ssize_t r = pread(fd, p, nb, off); if (r <= 0) return false; p += r; off += r; nb -= (size_t)r; }
    return true;
}
bool pwrite_all(int fd, const void* src, size_t nb, off_t off) {
    const char* p = static_cast<const char*>(src);
    while (nb) { const ssize_t r = pwrite(fd, p, nb, off); if (r <= 0) return false; p += r; off += r; nb -= (size_t)r; }
    return true;
}

// file section -> device array
int upload_section(int fd, uint64_t off, void* dev, size_t bytes, cudaStream_t st) {
    for (size_t done = 0, i = 0; done < bytes; done += STAGE_BYTES_IO, ++i) {
        const int slot = (int)(i % STAGE_RING);
        const size_t nb = std::min(STAGE_BYTES_IO, bytes - done);
        CU_TRY(cudaEventSynchronize(g_ring.ev[slot]));               // the copy that last used this buffer has left it
        if (!parallel_parts(nb, static_cast<char*>(g_ring.buf[slot]), [&](size_t lo, size_t n2, char* dst) { return pread_all(fd, dst, n2, (off_t)(off + done + lo)); }))
            return fail(XS_ERR_ARG, "short read from the index file");
        CU_TRY(cudaMemcpyAsync(static_cast<char*>(dev) + done, g_ring.buf[slot], nb, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaEventRecord(g_ring.ev[slot], st));
    }
    return XS_OK;
}
// device array -> file section
int download_section(int fd, uint64_t off, const void* dev, size_t bytes, cudaStream_t st) {
    for (size_t done = 0; done < bytes; done += STAGE_BYTES_IO) {
        const size_t nb = std::min(STAGE_BYTES_IO, bytes - done);
        CU_TRY(cudaMemcpyAsync(g_ring.buf[0], static_cast<const char*>(dev) + done, nb, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        if (!parallel_parts(nb, static_cast<char*>(g_ring.buf[0]), [&](size_t lo, size_t n2, char* src) { return pwrite_all(fd, src, n2, (off_t)(off + done + lo)); }))
            return fail(XS_ERR_ARG, "short write to the index file");
    }
    return XS_OK;
}
}  // namespace

extern "C" int xs_index_create(const void* db, int dtype, int64_t n, int d, int64_t stride_row, int64_t stride_col,
                               int device, int renormalise, int64_t id_offset, xs_index** out) {
    if (!db || !out) return fail(XS_ERR_ARG, "null pointer");
    bool colmajor = false;
    XS_TRY(check_layout(dtype, n, d, stride_row, stride_col, &colmajor));
    xs_index* ix = new xs_index();
    const bool compact = want_compact_build();
    int rc = index_alloc(ix, n, d, device, id_offset, !compact);
    if (rc == XS_OK && compact) rc = alloc_tiled(ix);
    if (rc != XS_OK) { index_free(ix); return rc; }
    const size_t es = dtype == XS_F64 ? 8 : 4;
    // The caller's matrix is pageable.  Worker threads copy a chunk of rows into a ring of pinned buffers (contiguous runs
    // either way: whole rows of a row-major source, one run per column of the F-order view vecs.T), the chunk goes out with
    // one cudaMemcpyAsync and the layout / finish kernels follow it on the stream -- the next chunk is being staged by the
    // CPU meanwhile.  (A cudaMemcpy2DAsync from pageable memory with a stream sync per chunk ran at 2.6-4.6 GB/s.)
    std::lock_guard<std::mutex> ring_lock(g_ring.mu);
    rc = g_ring.ensure();
    int64_t chunk = (int64_t)(STAGE_BYTES_IO / ((size_t)d * es));
    chunk = chunk < 1 ? 1 : (chunk > n ? n : chunk);
    if (rc == XS_OK) rc = ix->stage.ensure((size_t)chunk * d * es);
    const char* src = static_cast<const char*>(db);
    const int64_t stride = colmajor ? stride_col : stride_row;
    for (int64_t r0 = 0, i = 0; rc == XS_OK && r0 < n; r0 += chunk, ++i) {
        const int64_t rows = (n - r0 < chunk) ? n - r0 : chunk;
        const int slot = (int)(i % STAGE_RING);
        cudaError_t e = cudaEventSynchronize(g_ring.ev[slot]);          // the copy that last used this pinned buffer has left it
        if (e != cudaSuccess) { rc = fail(XS_ERR_CUDA, "index build failed: %s", cudaGetErrorString(e)); break; }
        char* pin = static_cast<char*>(g_ring.buf[slot]);
        const int64_t runs = colmajor ? d : rows;                       // contiguous runs of the source in this chunk
        const size_t run_bytes = (size_t)(colmajor ? rows : d) * es;
        std::vector<std::thread> th;
        const int nth = (int)std::min<int64_t>(IO_THREADS, runs);
        for (int t = 0; t < nth; ++t)
            th.emplace_back([=] {
                for (int64_t u = runs * t / nth; u < runs * (t + 1) / nth; ++u)
                    memcpy(pin + (size_t)u * run_bytes, src + ((size_t)u * stride + (colmajor ? (size_t)r0 : (size_t)r0 * stride)) * es, run_bytes);
            });
        for (auto& x : th) x.join();
        e = cudaMemcpyAsync(ix->stage.p, pin, (size_t)rows * d * es, cudaMemcpyHostToDevice, ix->stream);
        if (e == cudaSuccess) e = cudaEventRecord(g_ring.ev[slot], ix->stream);
        if (e != cudaSuccess) { rc = fail(XS_ERR_CUDA, "index build failed: %s", cudaGetErrorString(e)); break; }
        launch_layout_rows(ix->stage.p, dtype, colmajor, colmajor ? rows : d, rows, d, ix->d_pad, ix->db32 + (size_t)r0 * ix->d_pad, ix->stream);
        rc = finish_chunk(ix, r0, rows, renormalise != 0);
    }
    if (rc == XS_OK) { cudaError_t e = cudaStreamSynchronize(ix->stream); if (e != cudaSuccess) rc = fail(XS_ERR_CUDA, "index build failed: %s", cudaGetErrorString(e)); }
    ix->stage.release(); ix->stage16.release();
    if (rc == XS_OK && !compact) rc = build_tiled_twin(ix);
    if (rc != XS_OK) { index_free(ix); return rc; }
    *out = ix;
    return XS_OK;
}

extern "C" int xs_index_create_dev(const float* db_dev, int64_t n, int d, int device, int renormalise, int64_t id_offset, xs_index** out) {
    if (!db_dev || !out) return fail(XS_ERR_ARG, "null pointer");
    xs_index* ix = new xs_index();
    const bool compact = want_compact_build();
    int rc = index_alloc(ix, n, d, device, id_offset, !compact);
    if (rc == XS_OK && compact) rc = alloc_tiled(ix);
    if (rc != XS_OK) { index_free(ix); return rc; }
    const int64_t chunk = compact ? (1 << 16) : (1 << 20);                // compact: the chunk's bf16 rows pass through a scratch of this many rows
    for (int64_t r0 = 0; r0 < n && rc == XS_OK; r0 += chunk) {
        const int64_t rows = (n - r0 < chunk) ? n - r0 : chunk;
        launch_layout_rows(db_dev + (size_t)r0 * d, XS_F32, false, d, rows, d, ix->d_pad, ix->db32 + (size_t)r0 * ix->d_pad, ix->stream);
        rc = finish_chunk(ix, r0, rows, renormalise != 0);
    }
    if (rc != XS_OK) { index_free(ix); return rc; }
    cudaError_t e = cudaStreamSynchronize(ix->stream);
    if (e != cudaSuccess) { rc = fail(XS_ERR_CUDA, "index build failed: %s", cudaGetErrorString(e)); index_free(ix); return rc; }
    ix->stage16.release();
    if (!compact) rc = build_tiled_twin(ix);
    if (rc != XS_OK) { index_free(ix); return rc; }
    *out = ix;
    return XS_OK;
}

extern "C" int xs_index_save(xs_index* ix, const char* path) {
    if (!ix || !path) return fail(XS_ERR_ARG, "null pointer");
    std::lock_guard<std::mutex> lk(ix->mu);
    std::lock_guard<std::mutex> lk2(g_ring.mu);
    CU_TRY(cudaSetDevice(ix->device));
    XS_TRY(g_ring.ensure());
    IndexFileHeader h{};
    memcpy(h.magic, "XSB200\0\1", 8);
    h.version = 1; h.flags = (ix->rotate ? 1u : 0u) | (ix->db16t ? 2u : 0u) | (ix->db16 ? 0u : 4u);      // 4: no row-major bf16 section (compact index)
    h.n = ix->n; h.n_pad = ix->n_pad; h.d = ix->d; h.d_pad = ix->d_pad; h.rot_seed = ix->rot_seed;
    CU_TRY(cudaMemcpy(&h.stats, ix->dstats, sizeof(DevStats), cudaMemcpyDeviceToHost));
    h.bytes32 = (uint64_t)ix->n * ix->d_pad * 4; h.bytes16 = (uint64_t)ix->n_pad * ix->d_pad * 2;
    h.off32 = FILE_ALIGN;
    h.off16 = (h.off32 + h.bytes32 + FILE_ALIGN - 1) / FILE_ALIGN * FILE_ALIGN;
    h.off16t = ix->db16 ? (h.off16 + h.bytes16 + FILE_ALIGN - 1) / FILE_ALIGN * FILE_ALIGN : h.off16;
    const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return fail(XS_ERR_ARG, "cannot create %s", path);
    char page[FILE_ALIGN] = {};
    memcpy(page, &h, sizeof h);
    int rc = pwrite_all(fd, page, FILE_ALIGN, 0) ? XS_OK : fail(XS_ERR_ARG, "cannot write %s", path);
    if (rc == XS_OK) rc = download_section(fd, h.off32, ix->db32, h.bytes32, ix->stream);
    if (rc == XS_OK && ix->db16) rc = download_section(fd, h.off16, ix->db16, h.bytes16, ix->stream);
    if (rc == XS_OK && ix->db16t) rc = download_section(fd, h.off16t, ix->db16t, h.bytes16, ix->stream);
    close(fd);
    return rc;
}

extern "C" int xs_index_load(const char* path, int device, int64_t id_offset, xs_index** out) {
    if (!path || !out) return fail(XS_ERR_ARG, "null pointer");
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(XS_ERR_ARG, "cannot open %s", path);
    IndexFileHeader h{};
    if (!pread_all(fd, &h, sizeof h, 0) || memcmp(h.magic, "XSB200\0\1", 8) != 0 || h.version != 1) { close(fd); return fail(XS_ERR_ARG, "%s is not an index file of this library", path); }
    std::lock_guard<std::mutex> lk2(g_ring.mu);
    const char* no_tiled_env = getenv("XS_NO_TILED");
    const bool no_tiled = no_tiled_env && atoi(no_tiled_env);
    if ((h.flags & 4u) && !(h.flags & 2u)) { close(fd); return fail(XS_ERR_ARG, "%s holds no bf16 section at all", path); }
    // compact: only the tiled twin is uploaded (a file written by a compact index holds nothing else)
    const bool compact = (h.flags & 4u) || ((h.flags & 2u) && !no_tiled && default_compact());
    xs_index* ix = new xs_index();
    int rc = index_alloc(ix, h.n, h.d, device, id_offset, !compact);
    if (rc == XS_OK && (ix->d_pad != h.d_pad || ix->n_pad != h.n_pad)) rc = fail(XS_ERR_ARG, "%s was written with another padding", path);
    if (rc == XS_OK) rc = g_ring.ensure();
    if (rc == XS_OK) {
        ix->rotate = (h.flags & 1u) != 0; ix->rot_seed = h.rot_seed;
        cudaError_t e = cudaMemcpyAsync(ix->dstats, &h.stats, sizeof(DevStats), cudaMemcpyHostToDevice, ix->stream);
        if (e != cudaSuccess) rc = fail(XS_ERR_CUDA, "xs_index_load: %s", cudaGetErrorString(e));
    }
    if (rc == XS_OK) rc = upload_section(fd, h.off32, ix->db32, h.bytes32, ix->stream);
    if (rc == XS_OK && !compact) rc = upload_section(fd, h.off16, ix->db16, h.bytes16, ix->stream);
    if (rc == XS_OK && (h.flags & 2u)) {
        if ((compact || !no_tiled) && cudaMalloc(&ix->db16t, h.bytes16) == cudaSuccess) {
            ix->share->db16t = ix->db16t;
            rc = upload_section(fd, h.off16t, ix->db16t, h.bytes16, ix->stream);
            if (rc == XS_OK) rc = make_tmap_tiled(&ix->tmap_dbt_b, ix->db16t, ix->n_pad, ix->d_pad, GEMM_BN);
            if (rc == XS_OK) rc = make_tmap_tiled(&ix->tmap_dbt_h, ix->db16t, ix->n_pad, ix->d_pad, GEMM_BM);
            if (rc == XS_OK) ix->bytes += (int64_t)h.bytes16;
        } else {
            cudaGetLastError(); ix->db16t = nullptr;
            if (compact && rc == XS_OK) rc = fail(XS_ERR_NOMEM, "xs_index_load: no memory for the bf16 database");
        }
    }
    close(fd);
    if (rc == XS_OK) { cudaError_t e = cudaStreamSynchronize(ix->stream); if (e != cudaSuccess) rc = fail(XS_ERR_CUDA, "xs_index_load: %s", cudaGetErrorString(e)); }
    if (rc == XS_OK && !(h.flags & 2u)) rc = build_tiled_twin(ix);
    if (rc != XS_OK) { index_free(ix); return rc; }
    *out = ix;
    return XS_OK;
}

extern "C" int xs_index_destroy(xs_index* ix) {
    if (ix) { std::lock_guard<std::mutex> lk(ix->mu); }      // a call still inside the library on another thread finishes first
    index_free(ix);
    return XS_OK;
}

// A second search lane over the same database: shares the (read-only) database arrays and tensor maps, owns its
// workspaces, stream, tunables and statistics -- so two searches can be in flight on two streams at once.
static void copy_tunables(xs_index* dst, const xs_index* src) {
    dst->eps_sigmas = src->eps_sigmas; dst->scan_max_q = src->scan_max_q; dst->force_path = src->force_path;
    dst->gemm_splits = src->gemm_splits; dst->sample_pass = src->sample_pass; dst->pair_mode = src->pair_mode;
    dst->eps_mode = src->eps_mode; dst->inline_boot = src->inline_boot; dst->gemm_stages = src->gemm_stages; dst->half_units = src->half_units; dst->fin_per_sm = src->fin_per_sm; dst->scan_tiled = src->scan_tiled;
}

// caller holds src->mu
static int clone_locked(xs_index* src, xs_index** out) {
    *out = nullptr;
    CU_TRY(cudaSetDevice(src->device));
    xs_index* ix = new xs_index();
    ix->device = src->device; ix->num_sms = src->num_sms;
    ix->n = src->n; ix->n_pad = src->n_pad; ix->id_offset = src->id_offset; ix->d = src->d; ix->d_pad = src->d_pad;
    ix->db16 = src->db16; ix->db32 = src->db32; ix->dstats = src->dstats; ix->db16t = src->db16t;
    ix->tmap_db_b = src->tmap_db_b; ix->tmap_db_a = src->tmap_db_a; ix->tmap_dbt_b = src->tmap_dbt_b; ix->tmap_dbt_h = src->tmap_dbt_h;
    copy_tunables(ix, src);
    ix->rotate = src->rotate; ix->rot_seed = src->rot_seed;
    ix->share = src->share;
    { std::lock_guard<std::mutex> lk2(ix->share->mu); ++ix->share->refs; }
    cudaError_t e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
    for (auto& ev : ix->ev) if (e == cudaSuccess) e = cudaEventCreate(&ev);
    if (e != cudaSuccess) { cudaGetLastError(); index_free(ix); return fail(XS_ERR_CUDA, "xs_index_clone: %s", cudaGetErrorString(e)); }
    *out = ix;
    return XS_OK;
}

extern "C" int xs_index_clone(xs_index* src, xs_index** out) {
    if (!src || !out) return fail(XS_ERR_ARG, "null pointer");
    std::lock_guard<std::mutex> lk(src->mu);
    return clone_locked(src, out);
}

extern "C" int xs_index_info(const xs_index* ix, int64_t* n, int* d, int* device, int64_t* device_bytes) {
    if (!ix) return fail(XS_ERR_ARG, "null index");
    if (n) *n = ix->n;
    if (d) *d = ix->d;
    if (device) *device = ix->device;
    if (device_bytes) *device_bytes = ix->bytes;
    return XS_OK;
}

extern "C" int xs_set_param(xs_index* ix, const char* name, double value) {
    if (!ix || !name) return fail(XS_ERR_ARG, "null pointer");
    std::lock_guard<std::mutex> lk(ix->mu);
    if (!strcmp(name, "eps_sigmas")) ix->eps_sigmas = (float)value;
    else if (!strcmp(name, "scan_max_q")) ix->scan_max_q = (int)value;
    else if (!strcmp(name, "force_path")) ix->force_path = (int)value;
    else if (!strcmp(name, "gemm_splits")) ix->gemm_splits = (int)value;
    else if (!strcmp(name, "sample_pass")) ix->sample_pass = (int)value;
    else if (!strcmp(name, "pair_mode")) ix->pair_mode = (int)value;
    else if (!strcmp(name, "timing")) ix->timing = (int)value;
    else if (!strcmp(name, "certificate")) ix->eps_mode = ((int)value == 1) ? 1 : 0;
    else if (!strcmp(name, "inline_boot")) ix->inline_boot = (int)value != 0;
    else if (!strcmp(name, "boot_trace")) ix->boot_trace_on = (int)value != 0;
    else if (!strcmp(name, "gemm_stages")) ix->gemm_stages = ((int)value == 3) ? 3 : 4;
    else if (!strcmp(name, "half_units")) ix->half_units = (int)value != 0;
    else if (!strcmp(name, "fin_per_sm")) ix->fin_per_sm = (int)value == 1 ? 1 : 0;
    else if (!strcmp(name, "scan_tiled")) ix->scan_tiled = (int)value != 0;
    else if (!strcmp(name, "self_lanes")) ix->self_lanes = ((int)value >= 2) ? 2 : 1;
    else return fail(XS_ERR_ARG, "unknown parameter '%s'", name);
    return XS_OK;
}

// Debugging aid (xs_set_param "boot_trace" 1): globaltimer stamps (ns) of the last small-batch search.
//   which = 0: the GEMM launch with the in-kernel bootstrap, [grid][8] (0 job start, 1 first tile accumulated, 2 arrival posted,
//              3 all arrived (owners), 4 thresholds out, 5 first accumulator released, 6 job done, 7 own threshold selected (owners))
//   which = 1: the fused finalise launch, rows of 10 per (query, CTA) slot in launch order, unused slots zero (0 start, 1 sizes known,
//              2 pools gathered, 3 cut known, 4 share collected, 5 share rescored, 6 ticket taken, 7 sorted (last CTA), 8 emitted)
extern "C" int xs_debug_trace(xs_index* ix, int which, unsigned long long* out, int max_rows, int* rows) {
    if (!ix || !out || !rows) return fail(XS_ERR_ARG, "null pointer");
    std::lock_guard<std::mutex> lk(ix->mu);
    *rows = 0;
    Buf& b = which == 0 ? ix->boot_trace : ix->fin_trace;
    const int have = which == 0 ? ix->boot_trace_grid : ix->fin_trace_rows;
    const int width = which == 0 ? 8 : 10;
    if (!b.p || !have) return XS_OK;
    CU_TRY(cudaSetDevice(ix->device));
    CU_TRY(cudaDeviceSynchronize());
    const int g = have < max_rows ? have : max_rows;
    CU_TRY(cudaMemcpy(out, b.p, (size_t)g * width * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    *rows = g;
    return XS_OK;
}

extern "C" int xs_index_stats(const xs_index* cix, xs_stats* out) {
    if (!cix || !out) return fail(XS_ERR_ARG, "null pointer");
    xs_index* ix = const_cast<xs_index*>(cix);
    std::lock_guard<std::mutex> lk(ix->mu);
    if (ix->ev_valid) {
        cudaSetDevice(ix->device);
        if (cudaEventSynchronize(ix->ev[3]) == cudaSuccess) {
            cudaEventElapsedTime(&ix->stats.ms_coarse, ix->ev[1], ix->ev[2]);
            cudaEventElapsedTime(&ix->stats.ms_total, ix->ev[0], ix->ev[3]);
        }
        cudaGetLastError();
    }
    *out = ix->stats;
    return XS_OK;
}

// ---- search core -----------------------------------------------------------------------------------------
enum { PATH_SCAN = 1, PATH_GEMM = 2, PATH_EXACT = 3 };

struct CoreArgs {
    const float* raw;           // optional: caller's fp32 row-major [nq][d] device matrix, not yet copied into q32
    float* q32;                 // [nq][d_pad] device, fp32 (normalised in place when prep_renorm)
    int64_t nq;
    int k;
    bool prep;                  // run prep_queries (false: q32 is final, eps must be valid for nq queries)
    bool prep_renorm;
    const CUtensorMap* tmap_a;  // non-null: bf16 query operand comes from this map (self-kNN) at row a_row0
    int64_t a_row0;
    int64_t self_base;          // -1 or first database row of the queries
    int64_t* out_idx; float* out_score; int* status;   // device; pitch = k
    int* ncand;                 // device counter of rescored candidates (nullptr: the index's own)
    int path;                   // PATH_*
    const PushTarget* push;     // non-null: results also go straight into the peer-exchange mailboxes (out_idx etc. may be null)
    const MergeTarget* merge;   // non-null (with push): try to merge the world's lists in the same launch
    bool* merged;               // out: the merge did ride along (else the caller launches the merge kernel)
};

__global__ void boost_self_kernel(float* scores, int64_t pitch, int nq, int64_t self_base) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) scores[(int64_t)q * pitch + self_base + q] = INFINITY;
}
__global__ void zero_rows_bf16_kernel(__nv_bfloat16* p, int64_t count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) p[i] = __float2bfloat16(0.f);
}

// Exact path for queries q32[0..nq): fp32 rows, fp64 accumulation, per-slice exact top-k, merge.
static int run_exact(xs_index* ix, const float* q32, int64_t nq, int k, int64_t self_base,
                     int64_t* out_idx, float* out_score, int* status, int* launches) {
    const int P = (int)((ix->n + SLICE_ROWS - 1) / SLICE_ROWS);
    const int cap = k;
    const int64_t chunk_max = 16;
    XS_TRY(ix->scores.ensure((size_t)chunk_max * ix->n * sizeof(float)));
    XS_TRY(ix->ghist.ensure((size_t)chunk_max * HIST_BINS * sizeof(uint32_t)));
    XS_TRY(ix->sbound.ensure(64 * sizeof(float)));
    const int64_t slots = round_up(chunk_max, 128) * P;
    XS_TRY(ix->pool_items.ensure((size_t)slots * cap * 8));
    XS_TRY(ix->pool_count.ensure((size_t)slots * 4));
    XS_TRY(ix->pool_thr.ensure((size_t)slots * 4));
    for (int64_t q0 = 0; q0 < nq; q0 += chunk_max) {
        const int c = (int)((nq - q0 < chunk_max) ? nq - q0 : chunk_max);
        CU_TRY(cudaMemsetAsync(ix->ghist.p, 0, (size_t)c * HIST_BINS * sizeof(uint32_t), ix->cur));
        launch_exact_scores(ix->db32, q32 + q0 * ix->d_pad, c, ix->n, ix->d_pad, ix->scores.as<float>(), ix->n, ix->ghist.as<uint32_t>(), ix->dstats, ix->sbound.as<float>(), ix->num_sms, ix->cur);
        *launches += (c + 3) / 4;
        if (self_base >= 0) { boost_self_kernel<<<(c + 127) / 128, 128, 0, ix->cur>>>(ix->scores.as<float>(), ix->n, c, self_base + q0); ++*launches; }
        launch_scores_to_pools(ix->scores.as<float>(), ix->n, c, ix->n, k, nullptr, ix->ghist.as<uint32_t>(), ix->sbound.as<float>(), true, ix->pool_items.as<uint64_t>(),
                               ix->pool_count.as<int>(), ix->pool_thr.as<uint32_t>(), P, cap, ix->cur);
        FinaliseArgs fa{};
        fa.pool_items = ix->pool_items.as<uint64_t>(); fa.pool_count = ix->pool_count.as<int>(); fa.pool_thr = ix->pool_thr.as<uint32_t>();
        fa.P = P; fa.cap = cap; fa.db32 = ix->db32; fa.q32 = q32 + q0 * ix->d_pad; fa.d_pad = ix->d_pad; fa.eps = nullptr;
        fa.k = k; fa.exact = true; fa.id_offset = ix->id_offset; fa.self_base = self_base >= 0 ? self_base + q0 : -1;
        fa.out_idx = out_idx + q0 * k; fa.out_score = out_score ? out_score + q0 * k : nullptr;
        fa.status = status ? status + q0 : nullptr; fa.n_cand = nullptr; fa.out_pitch = k;
        launch_finalise(fa, c, ix->cur);
        *launches += 2;
    }
    CU_TRY(cudaGetLastError());
    return XS_OK;
}

// Completion counters of the split finalise: zeroed when (re)allocated, the kernel resets them after use.
static int ensure_tickets(xs_index* ix, int64_t nq) {
    const size_t need = (size_t)nq * 3 * sizeof(int);      // completion tickets, candidate counters, certificate flags
    if (need <= ix->fin_ticket.cap) return XS_OK;
    XS_TRY(ix->fin_ticket.ensure(need));
    CU_TRY(cudaMemsetAsync(ix->fin_ticket.p, 0, ix->fin_ticket.cap, ix->cur));
    return XS_OK;
}

// Query preparation (one launch): q32 (normalised when asked), rotated bf16 / fp32 copies for the coarse kernels, eps.
static int prepare_queries(xs_index* ix, const CoreArgs& a, __nv_bfloat16* q16, float* q32r, int* launches) {
    if (!a.prep) return XS_OK;
    launch_prep_queries(a.raw, ix->d, a.q32, q16, q32r, a.nq, round_up(a.nq, GEMM_BM), ix->d_pad, a.prep_renorm, ix->rotate, ix->rot_seed,
                        ix->dstats, ix->eps_sigmas, ix->eps_mode, ix->eps.as<float>(), ix->cur);
    ++*launches;
    return XS_OK;
}

// Everything the finalise stage needs that does not depend on the coarse path.
static int fill_finalise(xs_index* ix, const CoreArgs& a, int64_t q0, int64_t c, FinaliseArgs* fa, int* ncand) {
    const int k = a.k;
    fa->db32 = ix->db32; fa->q32 = a.q32 + q0 * ix->d_pad; fa->d_pad = ix->d_pad; fa->eps = ix->eps.as<float>() + q0;
    fa->k = k; fa->exact = false; fa->id_offset = ix->id_offset;
    fa->self_base = a.self_base >= 0 ? a.self_base + q0 : -1;
    fa->out_idx = a.out_idx ? a.out_idx + q0 * k : nullptr; fa->out_score = a.out_score ? a.out_score + q0 * k : nullptr;
    fa->status = a.status ? a.status + q0 : nullptr; fa->n_cand = ncand; fa->out_pitch = k;
    fa->cand_max = finalise_cand_max(k, ix->eps_mode);
    fa->per_sm = ix->fin_per_sm;
    XS_TRY(ix->fin_work.ensure(finalise_work_bytes(c, k, fa->cand_max)));
    XS_TRY(ensure_tickets(ix, c));
    fa->work = ix->fin_work.p; fa->ticket = ix->fin_ticket.as<int>();
    if (ix->boot_trace_on && c <= 128) {
        XS_TRY(ix->fin_trace.ensure((size_t)128 * 16 * 10 * sizeof(unsigned long long)));
        CU_TRY(cudaMemsetAsync(ix->fin_trace.p, 0, (size_t)128 * 16 * 10 * sizeof(unsigned long long), ix->cur));
        fa->trace = ix->fin_trace.as<unsigned long long>();
        ix->fin_trace_rows = (int)c * 16;
    }
    if (a.push) {
        // the mailbox part holds ALL nq queries of the call: advance the per-query pointers to this batch
        fa->push = *a.push;
        for (int g = 0; g < fa->push.world; ++g) {
            fa->push.ids[g] += q0 * k; fa->push.scores[g] += q0 * k; fa->push.status[g] += q0; fa->push.flags[g] += q0;
        }
        if (a.merge && c == a.nq) { fa->merge = *a.merge; fa->merge.on = 1; fa->merge.q0 = 0; }     // single-batch calls only
    }
    return XS_OK;
}

static int search_core(xs_index* ix, const CoreArgs& a) {
    const int64_t nq = a.nq; const int k = a.k;
    int launches = 0;
    ix->stats = xs_stats{};
    ix->stats.n_queries = nq; ix->stats.path = a.path;
    XS_TRY(ix->eps.ensure((size_t)nq * sizeof(float)));
    XS_TRY(ix->ncand.ensure(sizeof(int)));
    int* const ncand = a.ncand ? a.ncand : ix->ncand.as<int>();
    CU_TRY(cudaMemsetAsync(ncand, 0, sizeof(int), ix->cur));
    const bool timing = ix->timing != 0;            // CUDA events around the coarse kernel / the call (off: nothing between the launches)
    if (timing) CU_TRY(cudaEventRecord(ix->ev[0], ix->cur));
    ix->ev_valid = timing;

    if (a.path == PATH_EXACT) {
        if (a.push) return fail(XS_ERR_UNSUPPORTED, "the exact path does not feed the peer exchange directly (k=%d too large for the coarse filter?)", k);
        XS_TRY(prepare_queries(ix, a, nullptr, nullptr, &launches));
        if (timing) CU_TRY(cudaEventRecord(ix->ev[1], ix->cur));
        XS_TRY(run_exact(ix, a.q32, nq, k, a.self_base, a.out_idx, a.out_score, a.status, &launches));
        if (timing) CU_TRY(cudaEventRecord(ix->ev[2], ix->cur));
    } else if (a.path == PATH_SCAN) {
        XS_TRY(ix->q32r.ensure((size_t)nq * ix->d_pad * sizeof(float)));
        XS_TRY(prepare_queries(ix, a, nullptr, ix->q32r.as<float>(), &launches));
        const int P = (int)((ix->n + SLICE_ROWS - 1) / SLICE_ROWS);
        const int cap = (ix->eps_mode == 1 ? 4 : 2) * k + 256;
        const int64_t chunk_max = 8;
        XS_TRY(ix->scores.ensure((size_t)chunk_max * ix->n * sizeof(float)));
        XS_TRY(ix->ghist.ensure((size_t)chunk_max * HIST_BINS * sizeof(uint32_t)));
        XS_TRY(ix->sbound.ensure(64 * sizeof(float)));
        const int64_t slots = round_up(chunk_max, 128) * P;
        XS_TRY(ix->pool_items.ensure((size_t)slots * cap * 8));
        XS_TRY(ix->pool_count.ensure((size_t)slots * 4));
        XS_TRY(ix->pool_thr.ensure((size_t)slots * 4));
        const bool scan_tiled = ix->db16t && (ix->scan_tiled || !ix->db16);
        for (int64_t q0 = 0; q0 < nq; q0 += chunk_max) {
            const int c = (int)((nq - q0 < chunk_max) ? nq - q0 : chunk_max);
            CU_TRY(cudaMemsetAsync(ix->ghist.p, 0, (size_t)c * HIST_BINS * sizeof(uint32_t), ix->cur));
            if (timing && q0 == 0) CU_TRY(cudaEventRecord(ix->ev[1], ix->cur));
            launch_scan_scores(scan_tiled ? ix->db16t : ix->db16, scan_tiled, ix->q32r.as<float>() + q0 * ix->d_pad, c, ix->n, ix->d_pad, ix->scores.as<float>(), ix->n, ix->ghist.as<uint32_t>(), ix->dstats, ix->sbound.as<float>(), ix->num_sms, ix->cur);
            launches += (c + 1) / 2;
            if (timing && q0 == 0) CU_TRY(cudaEventRecord(ix->ev[2], ix->cur));
            launch_scores_to_pools(ix->scores.as<float>(), ix->n, c, ix->n, k, ix->eps.as<float>() + q0, ix->ghist.as<uint32_t>(), ix->sbound.as<float>(), false, ix->pool_items.as<uint64_t>(),
                                   ix->pool_count.as<int>(), ix->pool_thr.as<uint32_t>(), P, cap, ix->cur);
            FinaliseArgs fa{};
            fa.pool_items = ix->pool_items.as<uint64_t>(); fa.pool_count = ix->pool_count.as<int>(); fa.pool_thr = ix->pool_thr.as<uint32_t>();
            fa.P = P; fa.cap = cap;
            XS_TRY(fill_finalise(ix, a, q0, c, &fa, ncand));
            const bool mg = launch_finalise(fa, c, ix->cur);
            if (a.merged) *a.merged = mg;
            launches += 1 + finalise_launches(fa, c);
        }
    } else {
        // tcgen05 GEMM with fused top-K, queries in batches that bound the pool workspace
        // at most 8192 queries per batch (bounds the pool workspace); several batches are made equal (10,000 queries = 2 x 5,120,
        // not 8,192 + 1,808: the short tail batch would run the tensor pipe at a fraction of its rate)
        int64_t batch_max = 8192;
        if (nq > batch_max) { const int64_t nb = (nq + batch_max - 1) / batch_max; batch_max = round_up((nq + nb - 1) / nb, 2 * GEMM_BM); }
        const int64_t nq_pad = round_up(nq < batch_max ? nq : batch_max, GEMM_BM);
        CUtensorMap tmap_q;
        if (!a.tmap_a) XS_TRY(ix->q16.ensure((size_t)round_up(nq, GEMM_BM) * ix->d_pad * 2));
        XS_TRY(prepare_queries(ix, a, a.tmap_a ? nullptr : ix->q16.as<__nv_bfloat16>(), nullptr, &launches));
        (void)nq_pad;
        for (int64_t q0 = 0; q0 < nq; q0 += batch_max) {
            const int64_t c = (nq - q0 < batch_max) ? nq - q0 : batch_max;
            GemmPlan plan = plan_gemm(c, ix->n_pad, k, ix->num_sms, ix->gemm_splits, ix->pair_mode != 0);
            if (!ix->inline_boot) plan.inline_boot = 0;
            plan.stages3 = ix->gemm_stages == 3 ? 1 : 0;
            InlineBoot boot{};
            if (plan.inline_boot) {
                // the CTAs of this launch meet once inside the kernel: sample lists, arrival counter, published thresholds
                XS_TRY(ix->boot_samp.ensure((size_t)GEMM_BM * BOOT_MAX_GRID * 8 * sizeof(float)));
                if (!ix->boot_sync.p) {
                    XS_TRY(ix->boot_sync.ensure(16 + GEMM_BM * sizeof(uint64_t)));
                    CU_TRY(cudaMemsetAsync(ix->boot_sync.p, 0, ix->boot_sync.cap, ix->cur));
                    ix->boot_arrived = 0; ix->boot_published = 0; ix->boot_epoch = 0;
                }
                boot.samp = ix->boot_samp.as<float>();
                boot.arrive = ix->boot_sync.as<uint32_t>();
                boot.thr_pub = reinterpret_cast<uint64_t*>(static_cast<char*>(ix->boot_sync.p) + 16);
                ix->boot_arrived += (uint32_t)plan.grid;
                boot.arrive_target = ix->boot_arrived;
                boot.published = ix->boot_sync.as<uint32_t>() + 1;
                ix->boot_published += (uint32_t)c;
                boot.published_target = ix->boot_published;
                boot.epoch = ++ix->boot_epoch;
                if (ix->boot_trace_on) {
                    XS_TRY(ix->boot_trace.ensure((size_t)BOOT_MAX_GRID * 8 * sizeof(unsigned long long)));
                    CU_TRY(cudaMemsetAsync(ix->boot_trace.p, 0, (size_t)BOOT_MAX_GRID * 8 * sizeof(unsigned long long), ix->cur));
                    boot.trace = ix->boot_trace.as<unsigned long long>();
                    ix->boot_trace_grid = plan.grid;
                }
            }
            const int64_t slots = (int64_t)((plan.m_tiles + 1) & ~1) * plan.splits * GEMM_BM;
            XS_TRY(ix->pool_items.ensure((size_t)slots * plan.cap * 8));
            XS_TRY(ix->pool_count.ensure((size_t)slots * 4));
            XS_TRY(ix->pool_thr.ensure((size_t)slots * 4));
            const CUtensorMap* ta = a.tmap_a;
            int64_t row0 = a.a_row0 + q0;
            if (!ta) {
                // the query operand's tensor map only depends on (address, rows): encoded once per shape, not per step
                const void* qbase = ix->q16.as<__nv_bfloat16>() + q0 * ix->d_pad;
                const int64_t qrows = round_up(c, GEMM_BM);
                if (ix->tmap_q_base != qbase || ix->tmap_q_rows != qrows) {
                    XS_TRY(make_tmap(&ix->tmap_q_cached, qbase, qrows, ix->d_pad, GEMM_BM));
                    ix->tmap_q_base = qbase; ix->tmap_q_rows = qrows;
                }
                tmap_q = ix->tmap_q_cached;
                ta = &tmap_q; row0 = 0;
            }
            // threshold bootstrap on a strided sample of database tiles (skipped when the sample would be the whole database)
            const float* thr0 = nullptr;
            XS_TRY(ix->thr0.ensure((size_t)c * sizeof(float)));
            GemmPlan sp = plan_gemm_sample(plan, ix->num_sms, k);
            if (!plan.inline_boot && ix->sample_pass && plan.n_tiles >= 2 * ix->num_sms && 8 * sp.splits >= 3 * k && (size_t)sp.splits * 64 <= 200 * 1024) {
                const int64_t sslots = (int64_t)sp.m_tiles * sp.splits * GEMM_BM;
                XS_TRY(ix->pool_items.ensure((size_t)(sslots > slots ? sslots : slots) * plan.cap * 8));
                XS_TRY(ix->pool_count.ensure((size_t)(sslots > slots ? sslots : slots) * 4));
                XS_TRY(ix->pool_thr.ensure((size_t)(sslots > slots ? sslots : slots) * 4));
                sp.db_tiled = ix->db16t ? 1 : 0;
                cudaError_t es = launch_gemm_topk(*ta, ix->db16t ? ix->tmap_dbt_h : ix->tmap_db_a, nullptr, sp, c, ix->n, ix->d_pad, k, ix->eps.as<float>() + q0,
                                                  ix->pool_items.as<uint64_t>(), ix->pool_count.as<int>(), ix->pool_thr.as<uint32_t>(),
                                                  (int)row0, nullptr, nullptr, ix->cur);
                if (es != cudaSuccess) return fail(XS_ERR_CUDA, "gemm_topk (sample) launch failed: %s", cudaGetErrorString(es));
                launch_sample_threshold(ix->pool_items.as<uint64_t>(), ix->pool_count.as<int>(), sp.splits, sp.cap, k,
                                        ix->eps.as<float>() + q0, ix->thr0.as<float>(), c, ix->cur);
                thr0 = ix->thr0.as<float>();
                launches += 2;
            }
            if (timing && q0 == 0) CU_TRY(cudaEventRecord(ix->ev[1], ix->cur));
            plan.db_tiled = ix->db16t ? 1 : 0;
            if (!ix->half_units) plan.half_units = 0;
            cudaError_t e = launch_gemm_topk(*ta, ix->db16t ? (plan.pair ? ix->tmap_dbt_h : ix->tmap_dbt_b) : (plan.pair ? ix->tmap_db_a : ix->tmap_db_b),
                                             ix->db16t ? &ix->tmap_dbt_h : &ix->tmap_db_a, plan, c, ix->n, ix->d_pad, k, ix->eps.as<float>() + q0,
                                             ix->pool_items.as<uint64_t>(), ix->pool_count.as<int>(), ix->pool_thr.as<uint32_t>(),
                                             (int)row0, thr0, plan.inline_boot ? &boot : nullptr, ix->cur);
            if (e != cudaSuccess) return fail(XS_ERR_CUDA, "gemm_topk launch failed: %s", cudaGetErrorString(e));
            if (timing && q0 == 0) CU_TRY(cudaEventRecord(ix->ev[2], ix->cur));
            FinaliseArgs fa{};
            fa.pool_items = ix->pool_items.as<uint64_t>(); fa.pool_count = ix->pool_count.as<int>(); fa.pool_thr = ix->pool_thr.as<uint32_t>();
            fa.P = plan.splits; fa.cap = plan.cap;
            XS_TRY(fill_finalise(ix, a, q0, c, &fa, ncand));
            const bool mg = launch_finalise(fa, c, ix->cur);
            if (a.merged) *a.merged = mg;
            launches += 1 + finalise_launches(fa, c);
        }
    }
    if (timing) CU_TRY(cudaEventRecord(ix->ev[3], ix->cur));
    CU_TRY(cudaGetLastError());
    ix->stats.gpu_launches = launches;
    return XS_OK;
}

static int choose_path(const xs_index* ix, int64_t nq, int k) {
    if (ix->force_path >= PATH_SCAN && ix->force_path <= PATH_EXACT) return ix->force_path;
    if (k > 2048 || (int64_t)k * 4 > ix->n) return PATH_EXACT;     // huge k: the coarse filter has nothing to discard
    if (nq <= ix->scan_max_q) return PATH_SCAN;
    return PATH_GEMM;
}

// Device-API epilogue when the caller did not ask for the certificate bits: read them (one
// synchronisation), re-run uncertified queries exactly into the caller's device buffers.
static int rerun_uncertified(xs_index* ix, float* q32, int64_t nq, int k, int64_t self_base, int64_t* out_idx, float* out_score, int* status_dev) {
    XS_TRY(ix->h_status.ensure((size_t)(nq + 1) * sizeof(int)));
    CU_TRY(cudaMemcpyAsync(ix->h_status.as<int>() + 1, status_dev, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost, ix->cur));
    CU_TRY(cudaMemcpyAsync(ix->h_status.p, ix->ncand.p, sizeof(int), cudaMemcpyDeviceToHost, ix->cur));
    CU_TRY(cudaStreamSynchronize(ix->cur));
    const int* st = ix->h_status.as<int>() + 1;
    ix->stats.n_candidates = ix->h_status.as<int>()[0];
    int launches = 0;
    int64_t reruns = 0;
    for (int64_t q = 0; q < nq;) {
        if (!(st[q] & ST_UNCERTIFIED)) { ++q; continue; }
        int64_t e = q + 1;
        while (e < nq && e - q < 16 && (st[e] & ST_UNCERTIFIED)) ++e;
        XS_TRY(run_exact(ix, q32 + q * ix->d_pad, e - q, k, self_base >= 0 ? self_base + q : -1, out_idx + q * k,
                         out_score ? out_score + q * k : nullptr, nullptr, &launches));
        reruns += e - q;
        q = e;
    }
    ix->stats.n_exact_rerun = reruns;
    ix->stats.gpu_launches += launches;
    return XS_OK;
}

// Host-API epilogue: results + certificate bits land in pinned memory with ONE synchronisation;
// uncertified queries (rare) are re-run on the exact path and re-copied; then a plain memcpy to the
// caller's (possibly pageable) buffers.
// One device buffer for everything a host call returns, so that ONE copy brings it back:
//   [ids int64 nq*k | scores f32 nq*k | certificate bits int32 nq | candidate counter int32 (+ pad)]
struct OutPack {
    size_t off_score, off_status, off_ncand, bytes;
    OutPack(int64_t nq, int k) {
        off_score = (size_t)nq * k * sizeof(int64_t);
        off_status = off_score + (size_t)nq * k * sizeof(float);
        off_ncand = off_status + (size_t)nq * sizeof(int);
        bytes = (off_ncand + sizeof(int) + 15) & ~(size_t)15;
    }
};

static int finish_to_host(xs_index* ix, float* q32, int64_t nq, int k, int64_t self_base, int64_t* dev_idx, float* dev_score,
                          int* dev_status, bool coarse, int64_t* out_idx, float* out_score) {
    const size_t nb_i = (size_t)nq * k * sizeof(int64_t), nb_s = (size_t)nq * k * sizeof(float);
    const OutPack pk(nq, k);
    // the caller laid its device results out as an OutPack starting at dev_idx (xs_search / xs_aqe_search do)
    const bool packed = reinterpret_cast<char*>(dev_score) == reinterpret_cast<char*>(dev_idx) + pk.off_score &&
                        reinterpret_cast<char*>(dev_status) == reinterpret_cast<char*>(dev_idx) + pk.off_status;
    XS_TRY(ix->h_idx.ensure(packed ? pk.bytes : nb_i));
    XS_TRY(ix->h_score.ensure(nb_s));
    XS_TRY(ix->h_status.ensure((size_t)(nq + 1) * sizeof(int)));
    int64_t* h_idx = ix->h_idx.as<int64_t>();
    float* h_score = ix->h_score.as<float>();
    int* h_status = ix->h_status.as<int>() + 1;
    int* h_ncand = ix->h_status.as<int>();
    if (packed) {
        CU_TRY(cudaMemcpyAsync(ix->h_idx.p, dev_idx, pk.bytes, cudaMemcpyDeviceToHost, ix->stream));
        char* base = static_cast<char*>(ix->h_idx.p);
        h_score = reinterpret_cast<float*>(base + pk.off_score);
        h_status = reinterpret_cast<int*>(base + pk.off_status);
        h_ncand = reinterpret_cast<int*>(base + pk.off_ncand);
    } else {
        CU_TRY(cudaMemcpyAsync(h_idx, dev_idx, nb_i, cudaMemcpyDeviceToHost, ix->stream));
        if (out_score) CU_TRY(cudaMemcpyAsync(h_score, dev_score, nb_s, cudaMemcpyDeviceToHost, ix->stream));
        if (coarse) {
            CU_TRY(cudaMemcpyAsync(h_status, dev_status, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost, ix->stream));
            CU_TRY(cudaMemcpyAsync(h_ncand, ix->ncand.p, sizeof(int), cudaMemcpyDeviceToHost, ix->stream));
        }
    }
    CU_TRY(cudaStreamSynchronize(ix->stream));
    if (coarse) {
        const int* st = h_status;
        ix->stats.n_candidates = *h_ncand;
        int launches = 0;
        int64_t reruns = 0;
        for (int64_t q = 0; q < nq;) {
            if (!(st[q] & ST_UNCERTIFIED)) { ++q; continue; }
            int64_t e = q + 1;
            while (e < nq && e - q < 16 && (st[e] & ST_UNCERTIFIED)) ++e;
            XS_TRY(run_exact(ix, q32 + q * ix->d_pad, e - q, k, self_base >= 0 ? self_base + q : -1, dev_idx + q * k,
                             dev_score + q * k, nullptr, &launches));
            CU_TRY(cudaMemcpyAsync(h_idx + q * k, dev_idx + q * k, (size_t)(e - q) * k * sizeof(int64_t), cudaMemcpyDeviceToHost, ix->stream));
            if (out_score) CU_TRY(cudaMemcpyAsync(h_score + q * k, dev_score + q * k, (size_t)(e - q) * k * sizeof(float), cudaMemcpyDeviceToHost, ix->stream));
            reruns += e - q;
            q = e;
        }
        if (reruns) CU_TRY(cudaStreamSynchronize(ix->stream));
        ix->stats.n_exact_rerun = reruns;
        ix->stats.gpu_launches += launches;
    }
    memcpy(out_idx, h_idx, nb_i);
    if (out_score) memcpy(out_score, h_score, nb_s);
    return XS_OK;
}

static int check_search_args(const xs_index* ix, int64_t nq, int k) {
    if (!ix) return fail(XS_ERR_ARG, "null index");
    if (nq <= 0) return fail(XS_ERR_ARG, "no queries (nq=%lld)", (long long)nq);
    if (k <= 0 || k > ix->n) return fail(XS_ERR_ARG, "k=%d out of range for a database of %lld rows", k, (long long)ix->n);
    if (k > 4096) return fail(XS_ERR_UNSUPPORTED, "k=%d > 4096: use xs_rank_all for full rankings", k);
    return XS_OK;
}

extern "C" int xs_search_dev(xs_index* ix, const float* q_dev, int64_t nq, int renormalise_q, int k,
                             int64_t* out_idx_dev, float* out_score_dev, int32_t* out_status_dev, void* stream) {
    XS_TRY(check_search_args(ix, nq, k));
    if (!q_dev || !out_idx_dev) return fail(XS_ERR_ARG, "null pointer");
    std::lock_guard<std::mutex> lk(ix->mu);
    CU_TRY(cudaSetDevice(ix->device));
    // everything is enqueued on the caller's stream: no cross-stream hops on the hot path.  The workspace is
    // reused from call to call, so one index serves one stream at a time.
    ix->cur = static_cast<cudaStream_t>(stream);
    XS_TRY(ix->q32.ensure((size_t)nq * ix->d_pad * sizeof(float)));
    XS_TRY(ix->status.ensure((size_t)nq * sizeof(int)));
    CoreArgs a{};
    a.raw = q_dev;
    a.q32 = ix->q32.as<float>(); a.nq = nq; a.k = k; a.prep = true; a.prep_renorm = renormalise_q != 0; a.tmap_a = nullptr; a.a_row0 = 0;
    a.self_base = -1; a.out_idx = out_idx_dev; a.out_score = out_score_dev;
    a.status = out_status_dev ? out_status_dev : ix->status.as<int>();
    a.path = choose_path(ix, nq, k);
    int rc = search_core(ix, a);
    if (rc == XS_OK && !out_status_dev && a.path != PATH_EXACT)
        rc = rerun_uncertified(ix, a.q32, nq, k, -1, out_idx_dev, out_score_dev, a.status);
    return rc;
}

extern "C" int xs_search(xs_index* ix, const void* q, int dtype, int64_t nq, int64_t stride_row, int64_t stride_col,
                         int renormalise_q, int k, int64_t* out_idx, float* out_score) {
    XS_TRY(check_search_args(ix, nq, k));
    if (!q || !out_idx) return fail(XS_ERR_ARG, "null pointer");
    bool colmajor = false;
    XS_TRY(check_layout(dtype, nq, ix->d, stride_row, stride_col, &colmajor));
    std::lock_guard<std::mutex> lk(ix->mu);
    CU_TRY(cudaSetDevice(ix->device));
    ix->cur = ix->stream;
    const size_t es = dtype == XS_F64 ? 8 : 4;
    XS_TRY(ix->q_raw.ensure((size_t)nq * ix->d * es));
    XS_TRY(ix->q32.ensure((size_t)nq * ix->d_pad * sizeof(float)));
    const OutPack pk(nq, k);
    XS_TRY(ix->out_idx.ensure(pk.bytes));                 // ids | scores | certificate bits | counter, copied back in one go
    XS_TRY(stage_host_rows(q, dtype, colmajor, colmajor ? stride_col : stride_row, 0, nq, ix->d, ix->q_raw.p, ix->stream));
    CoreArgs a{};
    int extra_launches = 0;
    if (dtype == XS_F32 && !colmajor) a.raw = ix->q_raw.as<float>();          // dense fp32 rows: prepared in one fused kernel
    else { launch_layout_rows(ix->q_raw.p, dtype, colmajor, colmajor ? nq : ix->d, nq, ix->d, ix->d_pad, ix->q32.as<float>(), ix->stream); extra_launches = 1; }
    a.q32 = ix->q32.as<float>(); a.nq = nq; a.k = k; a.prep = true; a.prep_renorm = renormalise_q != 0; a.tmap_a = nullptr; a.a_row0 = 0;
    char* const pack = static_cast<char*>(ix->out_idx.p);
    a.self_base = -1; a.out_idx = reinterpret_cast<int64_t*>(pack); a.out_score = reinterpret_cast<float*>(pack + pk.off_score);
    a.status = reinterpret_cast<int*>(pack + pk.off_status); a.ncand = reinterpret_cast<int*>(pack + pk.off_ncand);
    a.path = choose_path(ix, nq, k);
    XS_TRY(search_core(ix, a));
    ix->stats.gpu_launches += extra_launches;
    return finish_to_host(ix, a.q32, nq, k, -1, a.out_idx, a.out_score, a.status, a.path != PATH_EXACT, out_idx, out_score);
}

extern "C" int xs_aqe_search(xs_index* ix, const int64_t* top_ids, int64_t nq, int kq, double w, int k,
                             int64_t* out_idx, float* out_score, float* out_queries) {
    XS_TRY(check_search_args(ix, nq, k));
    if (!top_ids || !out_idx) return fail(XS_ERR_ARG, "null pointer");
    if (kq < 1 || kq > 64) return fail(XS_ERR_ARG, "kq=%d out of range [1, 64]", kq);
    std::lock_guard<std::mutex> lk(ix->mu);
    CU_TRY(cudaSetDevice(ix->device));
    ix->cur = ix->stream;
    XS_TRY(ix->aqe_ids.ensure((size_t)nq * kq * sizeof(int64_t)));
    XS_TRY(ix->q32.ensure((size_t)nq * ix->d_pad * sizeof(float)));
    XS_TRY(ix->status.ensure((size_t)nq * sizeof(int)));
    XS_TRY(ix->out_idx.ensure((size_t)nq * k * sizeof(int64_t)));
    XS_TRY(ix->out_score.ensure((size_t)nq * k * sizeof(float)));
    // ids arrive as the caller saw them (row + id_offset); the kernel wants local rows
    // (staged in pinned memory owned by the index: the copy is asynchronous and nothing waits for it on the host)
    const size_t n_ids = (size_t)nq * kq;
    XS_TRY(ix->h_aqe.ensure(n_ids * sizeof(int64_t)));
    int64_t* local = ix->h_aqe.as<int64_t>();
    for (size_t i = 0; i < n_ids; ++i) local[i] = top_ids[i] - ix->id_offset;
    CU_TRY(cudaMemcpyAsync(ix->aqe_ids.p, local, n_ids * sizeof(int64_t), cudaMemcpyHostToDevice, ix->stream));
    launch_aqe_queries(ix->db32, ix->aqe_ids.as<int64_t>(), nq, kq, w, ix->n, ix->d_pad, ix->q32.as<float>(), ix->stream);
    CoreArgs a{};
    a.q32 = ix->q32.as<float>(); a.nq = nq; a.k = k; a.prep = true; a.prep_renorm = false; a.tmap_a = nullptr; a.a_row0 = 0;
    a.self_base = -1; a.out_idx = ix->out_idx.as<int64_t>(); a.out_score = ix->out_score.as<float>(); a.status = ix->status.as<int>();
    a.path = choose_path(ix, nq, k);
    XS_TRY(search_core(ix, a));
    ix->stats.gpu_launches += 1;
    if (out_queries)
        CU_TRY(cudaMemcpy2DAsync(out_queries, (size_t)ix->d * sizeof(float), ix->q32.p, (size_t)ix->d_pad * sizeof(float),
                                 (size_t)ix->d * sizeof(float), (size_t)nq, cudaMemcpyDeviceToHost, ix->stream));
    return finish_to_host(ix, a.q32, nq, k, -1, a.out_idx, a.out_score, a.status, a.path != PATH_EXACT, out_idx, out_score);
}

// Self-kNN runs as a two-deep software pipeline: while the GPU works on batch b, the host drains batch
// b-1 (certificate bits, the rare exact re-runs, copy into the caller's arrays) from the other half of
// a double-buffered pinned landing zone.
extern "C" int xs_self_knn(xs_index* ix, int64_t q_begin, int64_t q_end, int k, int64_t* out_idx, float* out_score) {
    if (!ix) return fail(XS_ERR_ARG, "null index");
    if (q_begin < 0 || q_end > ix->n || q_begin >= q_end) return fail(XS_ERR_ARG, "bad row range [%lld, %lld)", (long long)q_begin, (long long)q_end);
    XS_TRY(check_search_args(ix, q_end - q_begin, k));
    if (!out_idx) return fail(XS_ERR_ARG, "null pointer");
    std::lock_guard<std::mutex> lk(ix->mu);
    CU_TRY(cudaSetDevice(ix->device));
    const int64_t batch = 8192;
    // Two lanes: even batches run on this index, odd ones on an internal clone (same database arrays, own workspaces
    // and stream), so the selection / rescoring / copy-out of one batch overlaps the GEMM of the next.
    const bool two = ix->self_lanes >= 2 && q_end - q_begin > batch;
    if (two && !ix->self_lane) XS_TRY(clone_locked(ix, &ix->self_lane));
    xs_index* lane[2] = {ix, two ? ix->self_lane : ix};
    if (two) copy_tunables(lane[1], ix);
    const size_t nb_i = (size_t)batch * k * sizeof(int64_t), nb_s = (size_t)batch * k * sizeof(float), nb_st = (size_t)(batch + 1) * sizeof(int);
    for (int l = 0; l < (two ? 2 : 1); ++l) {
        lane[l]->cur = lane[l]->stream;
        XS_TRY(lane[l]->status.ensure((size_t)batch * sizeof(int)));
        XS_TRY(lane[l]->out_idx.ensure(nb_i));
        XS_TRY(lane[l]->out_score.ensure(nb_s));
    }
    XS_TRY(ix->h_idx.ensure(2 * nb_i));
    XS_TRY(ix->h_score.ensure(2 * nb_s));
    XS_TRY(ix->h_status.ensure(2 * nb_st));
    cudaEvent_t done[2];
    CU_TRY(cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming));
    xs_stats total{};
    int rc = XS_OK;

    // drain batch `pb` (rows [pr0, pr0+pc)) that was copied into landing-zone half `ph`
    auto drain = [&](int ph, int64_t pr0, int64_t pc, bool coarse) -> int {
        CU_TRY(cudaEventSynchronize(done[ph]));
        int64_t* hi = reinterpret_cast<int64_t*>(static_cast<char*>(ix->h_idx.p) + ph * nb_i);
        float* hs = reinterpret_cast<float*>(static_cast<char*>(ix->h_score.p) + ph * nb_s);
        int* hst = reinterpret_cast<int*>(static_cast<char*>(ix->h_status.p) + ph * nb_st);
        if (coarse) {
            total.n_candidates += hst[0];
            for (int64_t q = 0; q < pc;) {
                if (!(hst[1 + q] & ST_UNCERTIFIED)) { ++q; continue; }
                int64_t e = q + 1;
                while (e < pc && e - q < 16 && (hst[1 + e] & ST_UNCERTIFIED)) ++e;
                // rare: exact re-run on the lane that produced the batch, after its work in flight has finished (the
                // lane's device result buffers are reused by its next batch)
                xs_index* L = lane[ph];
                CU_TRY(cudaStreamSynchronize(L->stream));
                int launches = 0;
                XS_TRY(run_exact(L, L->db32 + (size_t)(pr0 + q) * L->d_pad, e - q, k, pr0 + q, L->out_idx.as<int64_t>(),
                                 L->out_score.as<float>(), nullptr, &launches));
                CU_TRY(cudaMemcpyAsync(hi + q * k, L->out_idx.p, (size_t)(e - q) * k * sizeof(int64_t), cudaMemcpyDeviceToHost, L->stream));
                CU_TRY(cudaMemcpyAsync(hs + q * k, L->out_score.p, (size_t)(e - q) * k * sizeof(float), cudaMemcpyDeviceToHost, L->stream));
                CU_TRY(cudaStreamSynchronize(L->stream));
                total.n_exact_rerun += e - q; total.gpu_launches += launches;
                q = e;
            }
        }
        memcpy(out_idx + (pr0 - q_begin) * k, hi, (size_t)pc * k * sizeof(int64_t));
        if (out_score) memcpy(out_score + (pr0 - q_begin) * k, hs, (size_t)pc * k * sizeof(float));
        return XS_OK;
    };

    int64_t prev_r0 = -1, prev_c = 0; int prev_h = 0; bool prev_coarse = false; bool have_prev = false;
    int b = 0;
    for (int64_t r0 = q_begin; r0 < q_end && rc == XS_OK; r0 += batch, ++b) {
        const int64_t c = (q_end - r0 < batch) ? q_end - r0 : batch;
        const int h = b & 1;
        xs_index* L = lane[h];
        CoreArgs a{};
        a.q32 = L->db32 + (size_t)r0 * L->d_pad; a.nq = c; a.k = k; a.prep = true; a.prep_renorm = false;   // rows are used as stored
        a.path = choose_path(L, c, k);
        a.tmap_a = (a.path == PATH_GEMM && L->db16) ? &L->tmap_db_a : nullptr;     // compact index: the rows go through prep like any query
        a.a_row0 = r0; a.self_base = r0;
        a.out_idx = L->out_idx.as<int64_t>(); a.out_score = L->out_score.as<float>(); a.status = L->status.as<int>();
        rc = search_core(L, a);
        if (rc != XS_OK) break;
        total.n_queries += c; total.gpu_launches += L->stats.gpu_launches; total.path = L->stats.path;
        char* hi = static_cast<char*>(ix->h_idx.p) + h * nb_i;
        char* hs = static_cast<char*>(ix->h_score.p) + h * nb_s;
        int* hst = reinterpret_cast<int*>(static_cast<char*>(ix->h_status.p) + h * nb_st);
        cudaError_t e = cudaMemcpyAsync(hi, a.out_idx, (size_t)c * k * sizeof(int64_t), cudaMemcpyDeviceToHost, L->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hs, a.out_score, (size_t)c * k * sizeof(float), cudaMemcpyDeviceToHost, L->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hst + 1, a.status, (size_t)c * sizeof(int), cudaMemcpyDeviceToHost, L->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hst, L->ncand.p, sizeof(int), cudaMemcpyDeviceToHost, L->stream);
        if (e == cudaSuccess) e = cudaEventRecord(done[h], L->stream);
        if (e != cudaSuccess) { rc = fail(XS_ERR_CUDA, "xs_self_knn: %s", cudaGetErrorString(e)); break; }
        if (have_prev) rc = drain(prev_h, prev_r0, prev_c, prev_coarse);
        prev_r0 = r0; prev_c = c; prev_h = h; prev_coarse = a.path != PATH_EXACT; have_prev = true;
    }
    if (rc == XS_OK && have_prev) rc = drain(prev_h, prev_r0, prev_c, prev_coarse);
    for (int l = 0; l < (two ? 2 : 1); ++l) { cudaStreamSynchronize(lane[l]->stream); lane[l]->ev_valid = false; }
    cudaEventDestroy(done[0]);
    cudaEventDestroy(done[1]);
    ix->stats = total;
    return rc;
}

extern "C" int xs_rank_all(xs_index* ix, const void* q, int dtype, int64_t nq, int64_t stride_row, int64_t stride_col,
                           int renormalise_q, int64_t* out_ranks, float* out_scores_sorted) {
    if (!ix) return fail(XS_ERR_ARG, "null index");
    if (!q || !out_ranks) return fail(XS_ERR_ARG, "null pointer");
    if (nq <= 0) return fail(XS_ERR_ARG, "no queries (nq=%lld)", (long long)nq);
    bool colmajor = false;
    XS_TRY(check_layout(dtype, nq, ix->d, stride_row, stride_col, &colmajor));
    std::lock_guard<std::mutex> lk(ix->mu);
    CU_TRY(cudaSetDevice(ix->device));
    ix->cur = ix->stream;
    const size_t es = dtype == XS_F64 ? 8 : 4;
    const int64_t n = ix->n;
    const int64_t col_block = 128;                     // queries per device-resident output block
    const int chunk = 8;                               // queries per scoring + sort pass
    XS_TRY(ix->q_raw.ensure((size_t)nq * ix->d * es));
    XS_TRY(ix->q32.ensure((size_t)nq * ix->d_pad * sizeof(float)));
    XS_TRY(ix->eps.ensure((size_t)nq * sizeof(float)));
    XS_TRY(ix->scores.ensure((size_t)chunk * n * sizeof(float)));
    XS_TRY(ix->ghist.ensure((size_t)chunk * HIST_BINS * sizeof(uint32_t)));
    XS_TRY(ix->sort_work.ensure(rank_all_work_bytes(chunk, n)));
    XS_TRY(ix->sbound.ensure(64 * sizeof(float)));
    const int64_t cb_max = nq < col_block ? nq : col_block;
    XS_TRY(ix->rank_out.ensure((size_t)n * cb_max * sizeof(int64_t)));
    if (out_scores_sorted) XS_TRY(ix->rank_scores.ensure((size_t)n * cb_max * sizeof(float)));
    XS_TRY(stage_host_rows(q, dtype, colmajor, colmajor ? stride_col : stride_row, 0, nq, ix->d, ix->q_raw.p, ix->stream));
    int launches = 0;
    ix->stats = xs_stats{};
    ix->stats.n_queries = nq; ix->stats.path = PATH_EXACT;
    ix->ev_valid = false;
    CoreArgs a{};
    if (dtype == XS_F32 && !colmajor) a.raw = ix->q_raw.as<float>();
    else { launch_layout_rows(ix->q_raw.p, dtype, colmajor, colmajor ? nq : ix->d, nq, ix->d, ix->d_pad, ix->q32.as<float>(), ix->stream); ++launches; }
    a.q32 = ix->q32.as<float>(); a.nq = nq; a.prep = true; a.prep_renorm = renormalise_q != 0;
    XS_TRY(prepare_queries(ix, a, nullptr, nullptr, &launches));
    for (int64_t b0 = 0; b0 < nq; b0 += col_block) {
        const int64_t cb = (nq - b0 < col_block) ? nq - b0 : col_block;
        for (int64_t q0 = 0; q0 < cb; q0 += chunk) {
            const int c = (int)((cb - q0 < chunk) ? cb - q0 : chunk);
            CU_TRY(cudaMemsetAsync(ix->ghist.p, 0, (size_t)c * HIST_BINS * sizeof(uint32_t), ix->stream));
            launch_exact_scores(ix->db32, ix->q32.as<float>() + (b0 + q0) * ix->d_pad, c, n, ix->d_pad, ix->scores.as<float>(), n,
                                ix->ghist.as<uint32_t>(), ix->dstats, ix->sbound.as<float>(), ix->num_sms, ix->stream);
            launch_rank_all(ix->scores.as<float>(), n, c, n, (int)q0, (int)cb, ix->id_offset, ix->sort_work.p,
                            ix->rank_out.as<int64_t>(), out_scores_sorted ? ix->rank_scores.as<float>() : nullptr, ix->stream);
            launches += (c + 3) / 4 + 13;
        }
        CU_TRY(cudaGetLastError());
        // block of columns [b0, b0+cb) of the caller's [n][nq] arrays
        CU_TRY(cudaMemcpy2DAsync(out_ranks + b0, (size_t)nq * sizeof(int64_t), ix->rank_out.p, (size_t)cb * sizeof(int64_t),
                                 (size_t)cb * sizeof(int64_t), (size_t)n, cudaMemcpyDeviceToHost, ix->stream));
        if (out_scores_sorted)
            CU_TRY(cudaMemcpy2DAsync(out_scores_sorted + b0, (size_t)nq * sizeof(float), ix->rank_scores.p, (size_t)cb * sizeof(float),
                                     (size_t)cb * sizeof(float), (size_t)n, cudaMemcpyDeviceToHost, ix->stream));
        CU_TRY(cudaStreamSynchronize(ix->stream));
    }
    ix->stats.gpu_launches = launches;
    return XS_OK;
}

extern "C" int xs_merge_candidates(int device, const int64_t* in_idx, const float* in_score, int n_parts, int64_t nq, int k,
                                   int64_t* out_idx, float* out_score, void* stream) {
    if (!in_idx || !in_score || !out_idx) return fail(XS_ERR_ARG, "null pointer");
    if (n_parts <= 0 || nq <= 0 || k <= 0) return fail(XS_ERR_ARG, "bad sizes");
    // two dense arrays: ids advance nq*k*8 bytes per part, scores nq*k*4
    return xs_merge_candidates_strided(device, in_idx, in_score, nullptr, nq * k * 8, nq * k * 4, 0, n_parts, nq, k, out_idx, out_score, nullptr, stream);
}

extern "C" int xs_merge_candidates_strided(int device, const void* in_idx, const void* in_score, const void* in_status,
                                           int64_t idx_part_stride, int64_t score_part_stride, int64_t status_part_stride,
                                           int n_parts, int64_t nq, int k,
                                           int64_t* out_idx, float* out_score, int32_t* out_status, void* stream) {
    if (!in_idx || !in_score || !out_idx) return fail(XS_ERR_ARG, "null pointer");
    if (n_parts <= 0 || nq <= 0 || k <= 0) return fail(XS_ERR_ARG, "bad sizes");
    if ((int64_t)n_parts * k > 16384) return fail(XS_ERR_UNSUPPORTED, "n_parts*k = %lld > 16384", (long long)n_parts * k);
    CU_TRY(cudaSetDevice(device));
    launch_merge_parts(in_idx, in_score, in_status, idx_part_stride, score_part_stride, status_part_stride, n_parts, nq, k,
                       out_idx, out_score, out_status, static_cast<cudaStream_t>(stream));
    CU_TRY(cudaGetLastError());
    return XS_OK;
}

// ---- peer exchange: per-shard result lists stored into every rank's mailbox over NVLink peer mappings --------
// Mailbox of one rank (its own HBM, exported through CUDA IPC):
//   header   acks [2 slots][W] | merge tickets [2] | push tickets [2][W]
//   flags    [2 slots][W senders][max_q]   per-query arrival epochs
//   data     [2 slots][W senders][part_bytes]; a part = ids int64 [nq][k] | scores f32 [nq][k] | certificate bits int32 [nq]
struct xs_exchange {
    int device = 0, world = 1, rank = 0;
    int64_t part_bytes = 0;                       // mailbox bytes per (slot, sender), multiple of 16
    int64_t max_q = 0;                            // queries per search the flag area is sized for
    char* local = nullptr;                        // this rank's mailbox
    char* peer[XCHG_MAX_WORLD] = {};              // every rank's mailbox as mapped here (peer[rank] == local)
    bool connected = false;
    uint32_t push_epoch[2] = {0, 0}, merge_epoch[2] = {0, 0};
    std::mutex mu;                                // epochs and launches of one exchange are serialised
    static constexpr size_t ACKS_OFF = 0, TICKET_OFF = 2 * XCHG_MAX_WORLD * 4, PUSH_TICKET_OFF = TICKET_OFF + 64, FLAGS_OFF = 1024;
    size_t flags_bytes() const { return (((size_t)2 * world * max_q * 4) + 255) & ~(size_t)255; }
    size_t data_off() const { return FLAGS_OFF + flags_bytes(); }
    size_t total() const { return data_off() + (size_t)2 * world * part_bytes; }
    char* part(int g, int slot, int sender) const { return peer[g] + data_off() + ((size_t)slot * world + sender) * part_bytes; }
    uint32_t* flags(int g, int slot, int sender) const { return reinterpret_cast<uint32_t*>(peer[g] + FLAGS_OFF) + ((size_t)slot * world + sender) * max_q; }
};

extern "C" int64_t xs_exchange_part_bytes(int64_t nq, int k) {
    return (nq * k * 12 + nq * 4 + 15) / 16 * 16;
}

extern "C" int xs_exchange_create(int device, int world, int rank, int64_t max_queries, int k_max, xs_exchange** out, unsigned char* handle_out) {
    if (!out || !handle_out) return fail(XS_ERR_ARG, "null pointer");
    if (world < 1 || world > XCHG_MAX_WORLD || rank < 0 || rank >= world) return fail(XS_ERR_ARG, "bad world/rank (%d/%d, at most %d ranks)", rank, world, XCHG_MAX_WORLD);
    if (max_queries <= 0 || k_max <= 0) return fail(XS_ERR_ARG, "max_queries and k_max must be positive");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
    CU_TRY(cudaSetDevice(device));
    xs_exchange* ex = new xs_exchange();
    ex->device = device; ex->world = world; ex->rank = rank;
    ex->max_q = max_queries; ex->part_bytes = xs_exchange_part_bytes(max_queries, k_max);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ex->local), ex->total());
    if (e == cudaSuccess) e = cudaMemset(ex->local, 0, ex->total());
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, ex->local);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (ex->local) cudaFree(ex->local);
        delete ex;
        return fail(e == cudaErrorMemoryAllocation ? XS_ERR_NOMEM : XS_ERR_CUDA, "xs_exchange_create: %s", cudaGetErrorString(e));
    }
    memcpy(handle_out, &h, 64);
    ex->peer[rank] = ex->local;
    *out = ex;
    return XS_OK;
}

extern "C" int xs_exchange_connect(xs_exchange* ex, const unsigned char* handles) {
    if (!ex || !handles) return fail(XS_ERR_ARG, "null pointer");
    if (ex->connected) return fail(XS_ERR_ARG, "already connected");
    CU_TRY(cudaSetDevice(ex->device));
    for (int g = 0; g < ex->world; ++g) {
        if (g == ex->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)g * 64, 64);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int j = 0; j < g; ++j) if (j != ex->rank && ex->peer[j]) { cudaIpcCloseMemHandle(ex->peer[j]); ex->peer[j] = nullptr; }
            return fail(XS_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", g, cudaGetErrorString(e));
        }
        ex->peer[g] = static_cast<char*>(p);
    }
    ex->connected = true;
    return XS_OK;
}

static int exchange_check_use(xs_exchange* ex, int slot, int64_t nq, int k) {
    if (!ex->connected && ex->world > 1) return fail(XS_ERR_ARG, "exchange not connected");
    if (slot < 0 || slot > 1 || nq <= 0 || k <= 0 || nq > ex->max_q || xs_exchange_part_bytes(nq, k) > ex->part_bytes)
        return fail(XS_ERR_ARG, "bad slot/sizes (slot %d, nq %lld of %lld, %lld of %lld bytes)", slot, (long long)nq, (long long)ex->max_q,
                    (long long)xs_exchange_part_bytes(nq, k), (long long)ex->part_bytes);
    return XS_OK;
}

// The search with the exchange fused into its last kernel.  Sending end: the emit step of every query stores the k results
// into all mailboxes and releases that query's flag -- no packed local result, no push kernel, no collective.  Receiving end
// (merge_out != null): the same CTA then waits for the other ranks' lists of its query and merges; when that is not possible
// (large batches, huge k) the merge kernel is launched behind the search instead.
static int search_exchange(xs_index* ix, const float* q_dev, int64_t nq, int renormalise_q, int k, xs_exchange* ex, int slot,
                           int64_t* out_idx, float* out_score, int32_t* out_status, bool want_merge, void* stream) {
    XS_TRY(check_search_args(ix, nq, k));
    if (!q_dev || !ex) return fail(XS_ERR_ARG, "null pointer");
    if (want_merge && !out_idx) return fail(XS_ERR_ARG, "null pointer");
    std::lock_guard<std::mutex> lk(ix->mu);
    std::unique_lock<std::mutex> lk2(ex->mu);
    XS_TRY(exchange_check_use(ex, slot, nq, k));
    if (ex->device != ix->device) return fail(XS_ERR_ARG, "index and exchange live on different devices");
    if (ex->push_epoch[slot] != ex->merge_epoch[slot]) return fail(XS_ERR_ARG, "search into slot %d before its previous result was merged", slot);
    if (want_merge && (int64_t)ex->world * k > 16384) return fail(XS_ERR_UNSUPPORTED, "world*k = %lld > 16384", (long long)ex->world * k);
    CU_TRY(cudaSetDevice(ix->device));
    ix->cur = static_cast<cudaStream_t>(stream);
    XS_TRY(ix->q32.ensure((size_t)nq * ix->d_pad * sizeof(float)));
    XS_TRY(ix->status.ensure((size_t)nq * sizeof(int)));
    PushTarget pt{};
    pt.world = ex->world;
    for (int g = 0; g < ex->world; ++g) {
        char* part = ex->part(g, slot, ex->rank);
        pt.ids[g] = reinterpret_cast<int64_t*>(part);
        pt.scores[g] = reinterpret_cast<float*>(part + (size_t)nq * k * 8);
        pt.status[g] = reinterpret_cast<int32_t*>(part + (size_t)nq * k * 12);
        pt.flags[g] = ex->flags(g, slot, ex->rank);
    }
    pt.my_acks = reinterpret_cast<const uint32_t*>(ex->local + xs_exchange::ACKS_OFF) + slot * XCHG_MAX_WORLD;
    pt.epoch = ex->push_epoch[slot] + 1;
    MergeTarget mt{};
    if (want_merge) {
        mt.base = ex->part(ex->rank, slot, 0); mt.part_bytes = ex->part_bytes;
        mt.flags = ex->flags(ex->rank, slot, 0); mt.flag_stride = ex->max_q;
        mt.nq_total = nq;
        mt.ticket = reinterpret_cast<uint32_t*>(ex->local + xs_exchange::TICKET_OFF) + slot;
        for (int g = 0; g < ex->world; ++g)
            mt.ack[g] = reinterpret_cast<uint32_t*>(ex->peer[g] + xs_exchange::ACKS_OFF) + slot * XCHG_MAX_WORLD + ex->rank;
        mt.out_idx = out_idx; mt.out_score = out_score; mt.out_status = out_status;
    }
    bool merged = false;
    CoreArgs a{};
    a.raw = q_dev;
    a.q32 = ix->q32.as<float>(); a.nq = nq; a.k = k; a.prep = true; a.prep_renorm = renormalise_q != 0; a.tmap_a = nullptr; a.a_row0 = 0;
    a.self_base = -1; a.out_idx = nullptr; a.out_score = nullptr; a.status = ix->status.as<int>();
    a.path = choose_path(ix, nq, k);
    a.push = &pt;
    a.merge = want_merge ? &mt : nullptr;
    a.merged = &merged;
    XS_TRY(search_core(ix, a));
    ++ex->push_epoch[slot];
    if (want_merge) {
        if (merged) ++ex->merge_epoch[slot];
        else {
            lk2.unlock();
            XS_TRY(xs_exchange_merge(ex, slot, nq, k, out_idx, out_score, out_status, stream));
            ix->stats.gpu_launches += 1;
        }
    }
    return XS_OK;
}

extern "C" int xs_search_dev_push(xs_index* ix, const float* q_dev, int64_t nq, int renormalise_q, int k,
                                  xs_exchange* ex, int slot, void* stream) {
    return search_exchange(ix, q_dev, nq, renormalise_q, k, ex, slot, nullptr, nullptr, nullptr, false, stream);
}

extern "C" int xs_search_dev_exchange(xs_index* ix, const float* q_dev, int64_t nq, int renormalise_q, int k, xs_exchange* ex, int slot,
                                      int64_t* out_idx_dev, float* out_score_dev, int32_t* out_status_dev, void* stream) {
    return search_exchange(ix, q_dev, nq, renormalise_q, k, ex, slot, out_idx_dev, out_score_dev, out_status_dev, true, stream);
}

extern "C" int xs_exchange_push(xs_exchange* ex, const void* packed_dev, int64_t nq, int k, int slot, void* stream) {
    if (!ex || !packed_dev) return fail(XS_ERR_ARG, "null pointer");
    std::lock_guard<std::mutex> lock(ex->mu);
    XS_TRY(exchange_check_use(ex, slot, nq, k));
    if (ex->push_epoch[slot] != ex->merge_epoch[slot]) return fail(XS_ERR_ARG, "push into slot %d before its previous result was merged", slot);
    CU_TRY(cudaSetDevice(ex->device));
    PushArgs a{};
    for (int g = 0; g < ex->world; ++g) {
        a.dst[g] = ex->part(g, slot, ex->rank);
        a.flag[g] = ex->flags(g, slot, ex->rank);
    }
    a.my_acks = reinterpret_cast<const uint32_t*>(ex->local + xs_exchange::ACKS_OFF) + slot * XCHG_MAX_WORLD;
    a.tickets = reinterpret_cast<uint32_t*>(ex->local + xs_exchange::PUSH_TICKET_OFF) + slot * XCHG_MAX_WORLD;
    a.epoch = ++ex->push_epoch[slot];
    a.nq = nq;
    launch_exchange_push(packed_dev, xs_exchange_part_bytes(nq, k), a, ex->world, static_cast<cudaStream_t>(stream));
    CU_TRY(cudaGetLastError());
    return XS_OK;
}

extern "C" int xs_exchange_merge(xs_exchange* ex, int slot, int64_t nq, int k, int64_t* out_idx, float* out_score, int32_t* out_status, void* stream) {
    if (!ex || !out_idx) return fail(XS_ERR_ARG, "null pointer");
    std::lock_guard<std::mutex> lock(ex->mu);
    XS_TRY(exchange_check_use(ex, slot, nq, k));
    if ((int64_t)ex->world * k > 16384) return fail(XS_ERR_UNSUPPORTED, "world*k = %lld > 16384", (long long)ex->world * k);
    if (ex->merge_epoch[slot] >= ex->push_epoch[slot]) return fail(XS_ERR_ARG, "merge of slot %d without a matching push", slot);
    CU_TRY(cudaSetDevice(ex->device));
    MergeSync ms{};
    ms.flags = ex->flags(ex->rank, slot, 0);
    ms.flag_stride = ex->max_q;
    ms.ticket = reinterpret_cast<uint32_t*>(ex->local + xs_exchange::TICKET_OFF) + slot;
    for (int g = 0; g < ex->world; ++g)
        ms.ack[g] = reinterpret_cast<uint32_t*>(ex->peer[g] + xs_exchange::ACKS_OFF) + slot * XCHG_MAX_WORLD + ex->rank;
    ms.epoch = ++ex->merge_epoch[slot];
    const char* base = ex->part(ex->rank, slot, 0);
    launch_merge_parts(base, base + nq * k * 8, base + nq * k * 12, ex->part_bytes, ex->part_bytes, ex->part_bytes, ex->world, nq, k,
                       out_idx, out_score, out_status, static_cast<cudaStream_t>(stream), &ms);
    CU_TRY(cudaGetLastError());
    return XS_OK;
}

extern "C" int xs_exchange_destroy(xs_exchange* ex) {
    if (!ex) return XS_OK;
    cudaSetDevice(ex->device);
    cudaDeviceSynchronize();
    for (int g = 0; g < ex->world; ++g)
        if (g != ex->rank && ex->peer[g]) cudaIpcCloseMemHandle(ex->peer[g]);
    if (ex->local) cudaFree(ex->local);
    cudaGetLastError();
    delete ex;
    return XS_OK;
}

// ---- native two-slot pipeline: one host call per sharded step ---------------------------------------------------------------
// At eight GPUs a 70-query step is ~0.12 ms of device time; driving it from Python (stream contexts, four ctypes calls, a
// tensor .cpu() for the certificate words) costs more than that on the host.  xs_pipeline_submit enqueues the whole step --
// local search with the exchange's sending end fused in, the merge, the copy of the merged certificate words into pinned
// memory -- on one of two lane streams and returns; xs_pipeline_collect waits for that slot's event and hands back device
// pointers to the merged result plus the number of uncertified queries (the caller re-runs those collectively).
struct xs_pipeline {
    xs_index* lane[2] = {nullptr, nullptr};       // one lane: the caller's index; two lanes: two internal clones (own workspaces, same database arrays)
    bool owns_lanes = false;
    xs_exchange* ex = nullptr;                    // null: single shard, no exchange step
    int n_lanes = 1, device = 0;
    int64_t nq_max = 0; int k_max = 0;
    cudaStream_t stream[2] = {nullptr, nullptr};
    cudaEvent_t in_ev = nullptr, done[2] = {nullptr, nullptr};
    Buf out_idx[2], out_score[2], out_status[2];
    PinnedBuf h_status[2];
    int64_t nq[2] = {0, 0}; int k[2] = {0, 0}; bool busy[2] = {false, false};
    int next = 0;
    std::mutex mu;
};

extern "C" int xs_pipeline_create(xs_index* ix, xs_exchange* ex, int64_t nq_max, int k_max, int lanes, xs_pipeline** out) {
    if (!ix || !out) return fail(XS_ERR_ARG, "null pointer");
    if (nq_max <= 0 || k_max <= 0) return fail(XS_ERR_ARG, "bad sizes");
    if (ex && ex->device != ix->device) return fail(XS_ERR_ARG, "index and exchange live on different devices");
    CU_TRY(cudaSetDevice(ix->device));
    xs_pipeline* p = new xs_pipeline();
    p->ex = ex; p->device = ix->device; p->nq_max = nq_max; p->k_max = k_max; p->n_lanes = lanes >= 2 ? 2 : 1;
    p->lane[0] = ix;
    int rc = XS_OK;
    if (p->n_lanes == 2) {
        // Two lanes overlap for real only if the previous batch's finalise CTAs can sit on the SMs that the next batch's scan
        // occupies: 3-stage operand ring for the scan (158 KB), one slim finalise CTA per SM (64 KB).  Both lanes are clones
        // (own workspaces and tunables, the caller's index keeps its latency-mode shapes).  XS_PIPE_OVERLAP=0 keeps the
        // default shapes on the lanes as well (they then merely alternate).
        const char* e = getenv("XS_PIPE_OVERLAP");
        const bool overlap = !(e && !atoi(e));
        std::lock_guard<std::mutex> lk(ix->mu);
        for (int l = 0; l < 2 && rc == XS_OK; ++l) {
            rc = clone_locked(ix, &p->lane[l]);
            if (rc == XS_OK && overlap) { p->lane[l]->gemm_stages = 3; p->lane[l]->fin_per_sm = 1; }
        }
        p->owns_lanes = true;
    }
    cudaError_t e = cudaSuccess;
    for (int s = 0; s < 2 && rc == XS_OK && e == cudaSuccess; ++s) {
        if (s < p->n_lanes) e = cudaStreamCreateWithFlags(&p->stream[s], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->done[s], cudaEventDisableTiming);
        if (e == cudaSuccess) rc = p->out_idx[s].ensure((size_t)nq_max * k_max * 8);
        if (rc == XS_OK) rc = p->out_score[s].ensure((size_t)nq_max * k_max * 4);
        if (rc == XS_OK) rc = p->out_status[s].ensure((size_t)nq_max * 4);
        if (rc == XS_OK) rc = p->h_status[s].ensure((size_t)nq_max * 4);
    }
    if (e == cudaSuccess && rc == XS_OK) e = cudaEventCreateWithFlags(&p->in_ev, cudaEventDisableTiming);
    if (e != cudaSuccess && rc == XS_OK) { cudaGetLastError(); rc = fail(XS_ERR_CUDA, "xs_pipeline_create: %s", cudaGetErrorString(e)); }
    if (rc != XS_OK) { xs_pipeline_destroy(p); return rc; }
    *out = p;
    return XS_OK;
}

// Debugging aid: the index behind lane `lane` (lane 1 is the internal clone), e.g. for xs_set_param "boot_trace" / xs_debug_trace.
extern "C" xs_index* xs_pipeline_lane(xs_pipeline* p, int lane) { return (p && lane >= 0 && lane < p->n_lanes) ? p->lane[lane] : nullptr; }

extern "C" int xs_pipeline_destroy(xs_pipeline* p) {
    if (!p) return XS_OK;
    cudaSetDevice(p->device);
    for (int s = 0; s < 2; ++s) {
        if (p->stream[s]) { cudaStreamSynchronize(p->stream[s]); cudaStreamDestroy(p->stream[s]); }
        if (p->done[s]) cudaEventDestroy(p->done[s]);
        p->out_idx[s].release(); p->out_score[s].release(); p->out_status[s].release(); p->h_status[s].release();
    }
    if (p->in_ev) cudaEventDestroy(p->in_ev);
    if (p->owns_lanes) for (int l = 0; l < 2; ++l) if (p->lane[l]) index_free(p->lane[l]);
    cudaGetLastError();
    delete p;
    return XS_OK;
}

// q_dev: fp32 row-major [nq][d] on the pipeline's device, ready on `caller_stream`.  *slot_out identifies the step.
extern "C" int xs_pipeline_submit(xs_pipeline* p, const float* q_dev, int64_t nq, int k, void* caller_stream, int* slot_out) {
    if (!p || !q_dev || !slot_out) return fail(XS_ERR_ARG, "null pointer");
    std::lock_guard<std::mutex> lk(p->mu);
    if (!p->lane[0]) return fail(XS_ERR_ARG, "pipeline without lanes");
    XS_TRY(check_search_args(p->lane[0], nq, k));
    if (nq > p->nq_max || k > p->k_max) return fail(XS_ERR_ARG, "step of %lld queries x %d exceeds the pipeline's %lld x %d", (long long)nq, k, (long long)p->nq_max, p->k_max);
    const int slot = p->next;
    if (p->busy[slot]) return fail(XS_ERR_ARG, "slot %d has not been collected (at most two steps in flight)", slot);
    CU_TRY(cudaSetDevice(p->device));
    const int l = p->n_lanes == 2 ? slot : 0;
    xs_index* ix = p->lane[l];
    cudaStream_t st = p->stream[l];
    CU_TRY(cudaEventRecord(p->in_ev, static_cast<cudaStream_t>(caller_stream)));     // the queries come first
    CU_TRY(cudaStreamWaitEvent(st, p->in_ev, 0));
    int64_t* oi = p->out_idx[slot].as<int64_t>(); float* os = p->out_score[slot].as<float>(); int32_t* ost = p->out_status[slot].as<int32_t>();
    if (p->ex) {
        // fused merge (the finalise CTA of a query waits for the peers' lists): fewest launches, best latency.  With two lanes the
        // merge is its own small kernel instead, so that the other lane's scan -- which needs every SM to itself -- is not held
        // up behind finalise CTAs that are only waiting for another GPU.
        // Measured on 8 B200s (70 queries, 300 steps): separate merge 182 / 186 us per step (2 / 1 lanes), fused 254 / 224 us --
        // with seven peers to wait for, the waiting CTAs delay the launch's completion by the slowest rank's skew.
        static const int force = [] { const char* e = getenv("XS_PIPE_MERGE"); return e ? atoi(e) : 0; }();     // 1 fused, 2 separate (default)
        const bool separate = force != 1;
        if (separate) {
            XS_TRY(xs_search_dev_push(ix, q_dev, nq, 0, k, p->ex, slot, st));
            XS_TRY(xs_exchange_merge(p->ex, slot, nq, k, oi, os, ost, st));
        } else {
            XS_TRY(xs_search_dev_exchange(ix, q_dev, nq, 0, k, p->ex, slot, oi, os, ost, st));
        }
    } else {
        XS_TRY(xs_search_dev(ix, q_dev, nq, 0, k, oi, os, ost, st));
    }
    CU_TRY(cudaMemcpyAsync(p->h_status[slot].p, ost, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaEventRecord(p->done[slot], st));
    p->nq[slot] = nq; p->k[slot] = k; p->busy[slot] = true;
    p->next = slot ^ 1;
    *slot_out = slot;
    return XS_OK;
}

// Waits for the step in `slot`; the caller's stream is made to wait for it as well, so the device pointers can be
// consumed there.  flagged (optional, nq entries) receives the indices of the uncertified queries.
extern "C" int xs_pipeline_collect(xs_pipeline* p, int slot, void* caller_stream, int64_t** out_idx_dev, float** out_score_dev,
                                   int64_t* n_flagged, int32_t* flagged) {
    if (!p || !out_idx_dev || !out_score_dev || !n_flagged) return fail(XS_ERR_ARG, "null pointer");
    std::lock_guard<std::mutex> lk(p->mu);
    if (slot < 0 || slot > 1 || !p->busy[slot]) return fail(XS_ERR_ARG, "nothing in flight in slot %d", slot);
    CU_TRY(cudaSetDevice(p->device));
    CU_TRY(cudaStreamWaitEvent(static_cast<cudaStream_t>(caller_stream), p->done[slot], 0));
    CU_TRY(cudaEventSynchronize(p->done[slot]));
    const int32_t* st = p->h_status[slot].as<int32_t>();
    int64_t nf = 0;
    for (int64_t q = 0; q < p->nq[slot]; ++q)
        if (st[q]) { if (flagged) flagged[nf] = (int32_t)q; ++nf; }
    *n_flagged = nf;
    *out_idx_dev = p->out_idx[slot].as<int64_t>();
    *out_score_dev = p->out_score[slot].as<float>();
    p->busy[slot] = false;
    return XS_OK;
}

namespace {
struct DevMem {                                   // scope-bound cudaMalloc for the one-shot entry points
    void* p = nullptr;
    ~DevMem() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};
}  // namespace

extern "C" int xs_diffusion_cg(int device, const int64_t* indptr, const int32_t* indices, const float* values, int64_t n,
                               const int64_t* trunc_ids, int64_t rows, int n_trunc, int maxiter, double tol,
                               float* out_scores) {
    if (!indptr || !indices || !values || !trunc_ids || !out_scores) return fail(XS_ERR_ARG, "null pointer");
    if (n <= 0 || n > 0x7FFFFFFF || rows <= 0) return fail(XS_ERR_ARG, "bad sizes (n=%lld, rows=%lld)", (long long)n, (long long)rows);
    if (n_trunc < 1 || n_trunc > 4096 || n_trunc > n) return fail(XS_ERR_ARG, "n_trunc must be in [1, min(n, 4096)] (got %d)", n_trunc);
    if (maxiter < 0 || !(tol >= 0.0)) return fail(XS_ERR_ARG, "bad maxiter/tol");
    if (indptr[0] != 0) return fail(XS_ERR_ARG, "indptr[0] must be 0");
    int64_t maxdeg = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t d = indptr[i + 1] - indptr[i];
        if (d < 0) return fail(XS_ERR_ARG, "indptr must be non-decreasing");
        if (d > maxdeg) maxdeg = d;
    }
    const int64_t nnz = indptr[n];
    const int T = n_trunc;
    const int stride = (int)std::max<int64_t>(1, std::min<int64_t>(maxdeg, T));
    CU_TRY(cudaSetDevice(device));
    int grid = 0;
    CU_TRY(diffusion_cg_grid(T, &grid));
    const int64_t chunk = std::min<int64_t>(rows, 32768);
    if ((int64_t)grid > chunk) grid = (int)chunk;
    DevMem d_ptr, d_ind, d_val, d_ids64, d_ids32, d_out, d_cols, d_vals, d_bad;
    cudaError_t e = d_ptr.alloc((size_t)(n + 1) * 8);
    if (e == cudaSuccess) e = d_ind.alloc((size_t)nnz * 4);
    if (e == cudaSuccess) e = d_val.alloc((size_t)nnz * 4);
    if (e == cudaSuccess) e = d_ids64.alloc((size_t)chunk * T * 8);
    if (e == cudaSuccess) e = d_ids32.alloc((size_t)chunk * T * 4);
    if (e == cudaSuccess) e = d_out.alloc((size_t)chunk * T * 4);
    if (e == cudaSuccess) e = d_cols.alloc((size_t)grid * T * stride * 2);
    if (e == cudaSuccess) e = d_vals.alloc((size_t)grid * T * stride * 4);
    if (e == cudaSuccess) e = d_bad.alloc(4);
    if (e == cudaSuccess) e = cudaMemset(d_bad.p, 0, 4);
    if (e == cudaSuccess) e = cudaMemcpy(d_ptr.p, indptr, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && nnz) e = cudaMemcpy(d_ind.p, indices, (size_t)nnz * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && nnz) e = cudaMemcpy(d_val.p, values, (size_t)nnz * 4, cudaMemcpyHostToDevice);
    for (int64_t r0 = 0; e == cudaSuccess && r0 < rows; r0 += chunk) {
        const int64_t rc = std::min<int64_t>(chunk, rows - r0);
        e = cudaMemcpy(d_ids64.p, trunc_ids + r0 * T, (size_t)rc * T * 8, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) break;
        launch_ids_to_i32(d_ids64.as<int64_t>(), d_ids32.as<int32_t>(), rc * T, nullptr);
        e = launch_diffusion_cg(d_ptr.as<int64_t>(), d_ind.as<int32_t>(), d_val.as<float>(), n, d_ids32.as<int32_t>(), rc, T,
                                stride, maxiter, tol, d_cols.as<uint16_t>(), d_vals.as<float>(), grid, d_out.as<float>(),
                                d_bad.as<int>(), nullptr);
        if (e == cudaSuccess) e = cudaMemcpy(out_scores + r0 * T, d_out.p, (size_t)rc * T * 4, cudaMemcpyDeviceToHost);
    }
    int bad = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&bad, d_bad.p, 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(e == cudaErrorMemoryAllocation ? XS_ERR_NOMEM : XS_ERR_CUDA, "xs_diffusion_cg: %s", cudaGetErrorString(e)); }
    if (bad) return fail(XS_ERR_ARG, "trunc_ids holds an id outside [0, n)");
    return XS_OK;
}

// ---- diffusion graph + gallery-side solves, device resident (src/utils/diffusion.py:52-116) ---------------------------
extern "C" int xs_diffusion_laplacian(int device, const float* sims, const int64_t* ids, int64_t n, int kd, double alpha, double gamma,
                                      int32_t* out_cols, float* out_vals, int32_t* out_cnt, float* out_affinity) {
    if (!sims || !ids || !out_cols || !out_vals || !out_cnt) return fail(XS_ERR_ARG, "null pointer");
    if (n <= 0 || kd <= 0 || n > 0x7FFFFFFF) return fail(XS_ERR_ARG, "bad sizes (n=%lld, kd=%d)", (long long)n, kd);
    CU_TRY(cudaSetDevice(device));
    const size_t count = (size_t)n * kd;
    DevMem d_ids64, d_ids32, d_sims, d_mut, d_aff, d_dinv, d_cols, d_vals, d_cnt;
    cudaError_t e = d_ids64.alloc(count * 8);
    if (e == cudaSuccess) e = d_ids32.alloc(count * 4);
    if (e == cudaSuccess) e = d_sims.alloc(count * 4);
    if (e == cudaSuccess) e = d_mut.alloc(count);
    if (e == cudaSuccess) e = d_aff.alloc(count * 4);
    if (e == cudaSuccess) e = d_dinv.alloc((size_t)n * 4);
    if (e == cudaSuccess) e = d_cols.alloc(count * 4);
    if (e == cudaSuccess) e = d_vals.alloc(count * 4);
    if (e == cudaSuccess) e = d_cnt.alloc((size_t)n * 4);
    if (e == cudaSuccess) e = cudaMemcpy(d_ids64.p, ids, count * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_sims.p, sims, count * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        launch_ids_to_i32(d_ids64.as<int64_t>(), d_ids32.as<int32_t>(), (int64_t)count, nullptr);
        launch_diffusion_graph(d_ids32.as<int32_t>(), kd, d_sims.as<float>(), kd, n, kd, alpha, gamma, d_mut.as<uint8_t>(), d_aff.as<float>(),
                               d_dinv.as<float>(), d_cols.as<int32_t>(), d_vals.as<float>(), d_cnt.as<int32_t>(), nullptr, nullptr);
        e = cudaMemcpy(out_cols, d_cols.p, count * 4, cudaMemcpyDeviceToHost);
    }
    if (e == cudaSuccess) e = cudaMemcpy(out_vals, d_vals.p, count * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(out_cnt, d_cnt.p, (size_t)n * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && out_affinity) e = cudaMemcpy(out_affinity, d_aff.p, count * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(e == cudaErrorMemoryAllocation ? XS_ERR_NOMEM : XS_ERR_CUDA, "xs_diffusion_laplacian: %s", cudaGetErrorString(e)); }
    return XS_OK;
}

// The whole gallery side without leaving the device: N x N self-kNN truncated at n_trunc -> mutual-kNN graph of the first
// kd neighbours -> Laplacian -> one truncated CG per row.  The host receives only the final (ids, scores) arrays.
extern "C" int xs_diffusion_offline(xs_index* ix, int n_trunc, int kd, double alpha, double gamma, int maxiter, double tol,
                                    int64_t* out_ids, float* out_sims, float* out_scores) {
    if (!ix) return fail(XS_ERR_ARG, "null index");
    if (!out_ids || !out_scores) return fail(XS_ERR_ARG, "null pointer");
    XS_TRY(check_search_args(ix, ix->n, n_trunc));
    if (kd < 1 || kd > n_trunc) return fail(XS_ERR_ARG, "kd must be in [1, n_trunc] (got %d)", kd);
    if (n_trunc > 4096 || ix->n > 0x7FFFFFFF) return fail(XS_ERR_ARG, "n_trunc <= 4096 and n < 2^31 required");
    if (maxiter < 0 || !(tol >= 0.0)) return fail(XS_ERR_ARG, "bad maxiter/tol");
    std::lock_guard<std::mutex> lk(ix->mu);
    CU_TRY(cudaSetDevice(ix->device));
    ix->cur = ix->stream;
    const int64_t n = ix->n;
    const int T = n_trunc;
    const int64_t batch = 8192;
    DevMem d_ids32, d_sims, d_mut, d_aff, d_dinv, d_cols, d_vals, d_cnt, d_ptr, d_out, d_scols, d_svals, d_bad;
    int grid = 0;
    CU_TRY(diffusion_cg_grid(T, &grid));
    const int stride = std::max(1, std::min(kd, T));
    cudaError_t e = d_ids32.alloc((size_t)n * T * 4);
    if (e == cudaSuccess) e = d_sims.alloc((size_t)n * T * 4);
    if (e == cudaSuccess) e = d_mut.alloc((size_t)n * kd);
    if (e == cudaSuccess) e = d_aff.alloc((size_t)n * kd * 4);
    if (e == cudaSuccess) e = d_dinv.alloc((size_t)n * 4);
    if (e == cudaSuccess) e = d_cols.alloc((size_t)n * kd * 4);
    if (e == cudaSuccess) e = d_vals.alloc((size_t)n * kd * 4);
    if (e == cudaSuccess) e = d_cnt.alloc((size_t)n * 4);
    if (e == cudaSuccess) e = d_ptr.alloc((size_t)(n + 1) * 8);
    if (e == cudaSuccess) e = d_out.alloc((size_t)n * T * 4);
    if (e == cudaSuccess) e = d_scols.alloc((size_t)grid * T * stride * 2);
    if (e == cudaSuccess) e = d_svals.alloc((size_t)grid * T * stride * 4);
    if (e == cudaSuccess) e = d_bad.alloc(4);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_bad.p, 0, 4, ix->stream);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(e == cudaErrorMemoryAllocation ? XS_ERR_NOMEM : XS_ERR_CUDA, "xs_diffusion_offline: %s", cudaGetErrorString(e)); }
    // 1. self-kNN, batch by batch, results stay on the device (ids as local int32 for the graph kernels)
    XS_TRY(ix->status.ensure((size_t)batch * sizeof(int)));
    XS_TRY(ix->out_idx.ensure((size_t)batch * T * sizeof(int64_t)));
    XS_TRY(ix->out_score.ensure((size_t)batch * T * sizeof(float)));
    XS_TRY(ix->h_status.ensure((size_t)(batch + 1) * sizeof(int)));
    xs_stats total{};
    for (int64_t r0 = 0; r0 < n; r0 += batch) {
        const int64_t c = (n - r0 < batch) ? n - r0 : batch;
        CoreArgs a{};
        a.q32 = ix->db32 + (size_t)r0 * ix->d_pad; a.nq = c; a.k = T; a.prep = true; a.prep_renorm = false;
        a.path = choose_path(ix, c, T);
        a.tmap_a = (a.path == PATH_GEMM && ix->db16) ? &ix->tmap_db_a : nullptr; a.a_row0 = r0; a.self_base = r0;
        a.out_idx = ix->out_idx.as<int64_t>(); a.out_score = ix->out_score.as<float>(); a.status = ix->status.as<int>();
        XS_TRY(search_core(ix, a));
        total.n_queries += c; total.gpu_launches += ix->stats.gpu_launches; total.path = ix->stats.path;
        if (a.path != PATH_EXACT) {
            int* hst = ix->h_status.as<int>();
            CU_TRY(cudaMemcpyAsync(hst, a.status, (size_t)c * sizeof(int), cudaMemcpyDeviceToHost, ix->stream));
            CU_TRY(cudaStreamSynchronize(ix->stream));
            int launches = 0;
            for (int64_t q = 0; q < c;) {
                if (!(hst[q] & ST_UNCERTIFIED)) { ++q; continue; }
                int64_t qe = q + 1;
                while (qe < c && qe - q < 16 && (hst[qe] & ST_UNCERTIFIED)) ++qe;
                XS_TRY(run_exact(ix, ix->db32 + (size_t)(r0 + q) * ix->d_pad, qe - q, T, r0 + q, a.out_idx + q * T, a.out_score + q * T, nullptr, &launches));
                total.n_exact_rerun += qe - q;
                q = qe;
            }
            total.gpu_launches += launches;
        }
        launch_ids_strided_to_i32(a.out_idx, T, d_ids32.as<int32_t>() + (size_t)r0 * T, T, c, T, ix->id_offset, ix->stream);
        CU_TRY(cudaMemcpyAsync(d_sims.as<float>() + (size_t)r0 * T, a.out_score, (size_t)c * T * sizeof(float), cudaMemcpyDeviceToDevice, ix->stream));
        CU_TRY(cudaMemcpyAsync(out_ids + (size_t)r0 * T, a.out_idx, (size_t)c * T * sizeof(int64_t), cudaMemcpyDeviceToHost, ix->stream));
        if (out_sims) CU_TRY(cudaMemcpyAsync(out_sims + (size_t)r0 * T, a.out_score, (size_t)c * T * sizeof(float), cudaMemcpyDeviceToHost, ix->stream));
        CU_TRY(cudaStreamSynchronize(ix->stream));          // the batch buffers are reused
    }
    // 2. graph of the first kd neighbours, 3. CG per row -- consuming the lists where they are
    launch_diffusion_graph(d_ids32.as<int32_t>(), T, d_sims.as<float>(), T, n, kd, alpha, gamma, d_mut.as<uint8_t>(), d_aff.as<float>(), d_dinv.as<float>(),
                           d_cols.as<int32_t>(), d_vals.as<float>(), d_cnt.as<int32_t>(), d_ptr.as<int64_t>(), ix->stream);
    e = launch_diffusion_cg(d_ptr.as<int64_t>(), d_cols.as<int32_t>(), d_vals.as<float>(), n, d_ids32.as<int32_t>(), n, T, stride, maxiter, tol,
                            d_scols.as<uint16_t>(), d_svals.as<float>(), grid, d_out.as<float>(), d_bad.as<int>(), ix->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_scores, d_out.p, (size_t)n * T * 4, cudaMemcpyDeviceToHost, ix->stream);
    int bad = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, d_bad.p, 4, cudaMemcpyDeviceToHost, ix->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(XS_ERR_CUDA, "xs_diffusion_offline: %s", cudaGetErrorString(e)); }
    if (bad) return fail(XS_ERR_CUDA, "xs_diffusion_offline: a neighbour id fell outside the database");
    ix->stats = total;
    ix->ev_valid = false;
    return XS_OK;
}

extern "C" int xs_mutual_knn(int device, const int64_t* ids, int64_t n, int kd, uint8_t* out_mutual) {
    if (!ids || !out_mutual) return fail(XS_ERR_ARG, "null pointer");
    if (n <= 0 || kd <= 0 || n > 0x7FFFFFFF) return fail(XS_ERR_ARG, "bad sizes (n=%lld, kd=%d)", (long long)n, kd);
    CU_TRY(cudaSetDevice(device));
    const size_t count = (size_t)n * kd;
    void *d64 = nullptr, *d32 = nullptr, *dm = nullptr;
    cudaError_t e = cudaMalloc(&d64, count * 8);
    if (e == cudaSuccess) e = cudaMalloc(&d32, count * 4);
    if (e == cudaSuccess) e = cudaMalloc(&dm, count);
    if (e == cudaSuccess) e = cudaMemcpy(d64, ids, count * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        launch_mutual_knn(static_cast<const int64_t*>(d64), static_cast<int32_t*>(d32), n, kd, static_cast<uint8_t*>(dm), nullptr);
        e = cudaMemcpy(out_mutual, dm, count, cudaMemcpyDeviceToHost);
    }
    if (d64) cudaFree(d64);
    if (d32) cudaFree(d32);
    if (dm) cudaFree(dm);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(e == cudaErrorMemoryAllocation ? XS_ERR_NOMEM : XS_ERR_CUDA, "xs_mutual_knn: %s", cudaGetErrorString(e)); }
    return XS_OK;
}
