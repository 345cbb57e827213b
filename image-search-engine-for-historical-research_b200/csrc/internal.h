// Host-side declarations shared by the translation units of libxs_b200.so (not part of the ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace xs {

// Geometry of the coarse (bf16) GEMM kernel -- see gemm_topk.cu.
constexpr int GEMM_BM = 128;      // queries per tile  (UMMA M, TMEM lanes)
constexpr int GEMM_BN = 256;      // database rows per tile (UMMA N, TMEM columns)
constexpr int GEMM_BK = 64;       // K elements per pipeline stage (= one 128-byte swizzle atom of bf16)
constexpr int ROW_ALIGN = 256;    // database rows are padded to a multiple of this
constexpr int COL_ALIGN = 64;     // descriptor length is padded to a multiple of this

constexpr int SLICE_ROWS = 4096;
constexpr int HIST_BINS = 4096;   // database-wide histogram of score keys (top 12 bits), filled by the scoring kernels
constexpr int HIST_SHIFT = 20;  // rows per partial list in the score-matrix -> pools kernel

// status bits written by the finalise kernel (one int32 per query)
constexpr int ST_UNCERTIFIED = 1;

struct DevStats {                 // device-side database statistics (uint-ordered positive floats)
    unsigned int v4max_bits;      // max over rows of ||v||_4
    unsigned int vnmax_bits;      // max over rows of ||v||_2
};

// ---- build.cu -----------------------------------------------------------------------------------
// src_tile is a device buffer holding `rows` database rows either row-major ([rows][pitch]) or
// column-major ([d][pitch], the reference's (D,N) layout); writes fp32 rows [rows][d_pad].
void launch_layout_rows(const void* src_tile, int dtype, bool colmajor, int64_t pitch,
                        int64_t rows, int d, int d_pad, float* dst32, cudaStream_t st);
// Optional L2 normalisation, bf16 copy, ||.||_4 / ||.||_2 maxima.  dst16 rows have pitch d_pad.
void launch_finish_rows(float* rows32, __nv_bfloat16* rows16, int64_t rows, int d_pad, bool renorm,
                        DevStats* stats, cudaStream_t st);
// Query preparation: normalise (optional), fp32 + bf16 copies, per-query error band eps.
void launch_prep_queries(float* q32, __nv_bfloat16* q16, int64_t nq, int d_pad, bool renorm,
                         const DevStats* stats, float eps_sigmas, float* eps, cudaStream_t st);

// One-kernel variant reading the caller's raw fp32 row-major queries (d == d_pad <= 2048, 16-byte aligned);
// also zero-fills bf16 rows [nq, nq_pad).  Returns false (nothing launched) when the fast path does not apply.
bool launch_prep_queries_fused(const float* raw, float* q32, __nv_bfloat16* q16, int64_t nq, int64_t nq_pad, int d, int d_pad,
                               bool renorm, const DevStats* stats, float eps_sigmas, float* eps, cudaStream_t st);

// AQE (Reranking.py:195-208): q_out[q] = normalise( sum_j ((kq-j)/kq)^w * db32[top_ids[q][j]] ), float64 inside.
void launch_aqe_queries(const float* db32, const int64_t* top_ids, int64_t nq, int kq, double w, int64_t n, int d_pad,
                        float* q_out, cudaStream_t st);

// Tiled copy of the bf16 database for the GEMM's B operand: [n_pad/256][d_pad/64][256][64], so that every TMA box
// (128 or 256 rows x 64 columns) is one contiguous 16/32 KB run in HBM instead of 128-byte pieces 4 KB apart.
void launch_tile_db16(const __nv_bfloat16* db16, __nv_bfloat16* db16t, int64_t n_pad, int d_pad, cudaStream_t st);

// ---- scan.cu ------------------------------------------------------------------------------------
// Batch-1 HBM scan: scores[q][row] = <db16[row], q32[q]>, fp32 accumulate.  Both scoring kernels also
// add every score to ghist[q][score_key >> HIST_SHIFT] (zeroed by the caller).
void launch_scan_scores(const __nv_bfloat16* db16, const float* q32, int nq, int64_t n, int d_pad,
                        float* scores, int64_t score_pitch, uint32_t* ghist, int num_sms, cudaStream_t st);
// Exact scoring: scores[q][row] = fp32( sum_fp64 db32[row][i] * q32[q][i] ).  Any nq (looped in 4s).
void launch_exact_scores(const float* db32, const float* q32, int nq, int64_t n, int d_pad,
                         float* scores, int64_t score_pitch, uint32_t* ghist, int num_sms, cudaStream_t st);
// Score matrix -> candidate pools (one partial list per SLICE_ROWS rows).
//   exact = false: keep items within the eps band below the slice's k-th best
//   exact = true : keep exactly the slice's k best (full 64-bit item order)
void launch_scores_to_pools(const float* scores, int64_t score_pitch, int nq, int64_t n, int k,
                            const float* eps, const uint32_t* ghist, bool exact, uint64_t* pool_items, int* pool_count,
                            uint32_t* pool_thr, int P, int cap, cudaStream_t st);

// ---- gemm_topk.cu -------------------------------------------------------------------------------
struct GemmPlan {
    int m_tiles;      // ceil(nq / 128)
    int n_tiles;      // n_pad / 256
    int splits;       // partial lists per query (P)
    int k_keep;       // items kept by a mid-job trim
    int cap;          // capacity of one partial list (>= 2 * k_keep)
    int grid;         // CTAs launched
    int db_tiled;     // 1: the database tensor map is over the tiled copy [n_pad/256][d_pad/64][256][64]
    int half;         // 1: 128-row database tiles, one CTA each, 6-stage ring (bootstrap pass; needs the box-128 database map)
    int pair;         // 1: cta_group::2 kernel, clusters of two CTAs, 256 x 256 tiles (needs the box-128 database map)
    int sample_mode;  // 1: threshold bootstrap pass (8 best scores per query and tile, no ids)
    int tile_stride;  // database tile t of the plan is tile t * tile_stride of the matrix (sample pass > 1)
};
GemmPlan plan_gemm(int64_t nq, int64_t n_pad, int k, int num_sms, int forced_splits, bool allow_pair);
GemmPlan plan_gemm_sample(const GemmPlan& main_plan, int num_sms, int k);
// tmap_q: [nq_pad128][d_pad] bf16, box {64,128};  tmap_db: [n_pad][d_pad] bf16, box {64,256} (box {64,128} when plan.pair)
cudaError_t launch_gemm_topk(const CUtensorMap& tmap_q, const CUtensorMap& tmap_db, const GemmPlan& plan,
                             int64_t nq, int64_t n_valid, int d_pad, int k, const float* eps,
                             uint64_t* pool_items, int* pool_count, uint32_t* pool_thr, int a_row0,
                             const float* thr0, cudaStream_t st);

// ---- finalise.cu --------------------------------------------------------------------------------
struct FinaliseArgs {
    const uint64_t* pool_items; const int* pool_count; const uint32_t* pool_thr;
    int P, cap;
    const float* db32; const float* q32; int d_pad;
    const float* eps;           // per-query band (ignored when exact)
    int k; bool exact;          // exact: pool scores are already fp32-exact -> no rescoring
    int64_t id_offset;
    int64_t self_base;          // >= 0: query q is database row self_base + q and must rank first
    int64_t* out_idx; float* out_score; int* status; int* n_cand;
    int64_t out_pitch;          // elements between consecutive queries in out_idx / out_score
    void* work;                 // optional: finalise_work_bytes(nq, k) of scratch -> enables the 3-kernel split for small batches
    int* ticket;                // [nq] zero-initialised completion counters (self-resetting), required for the split
    uint64_t* w_cand; int* w_ncand; int* w_flag; int* w_ticket;   // carved out by launch_finalise
};
int  finalise_cand_max(int k);
size_t finalise_work_bytes(int64_t nq, int k);
int  finalise_launches(const FinaliseArgs& a, int64_t nq);
void launch_finalise(const FinaliseArgs& a, int64_t nq, cudaStream_t st);
// thr0[q] = (k-th best pooled coarse score) - 2 eps[q], one ulp lower; -inf when fewer than k items.
void launch_sample_threshold(const uint64_t* pool_items, const int* pool_count, int P, int cap, int k,
                             const float* eps, float* thr0, int64_t nq, cudaStream_t st);
// Multi-GPU merge of [parts][nq][k] lists.
constexpr int XCHG_MAX_WORLD = 16;
struct MergeSync {                       // receiving end of the peer exchange (all zero = plain merge)
    const uint32_t* flags;               // local arrival flags of this slot, one per sending rank
    uint32_t epoch;
    uint32_t* ticket;                    // local CTA counter of this slot
    uint32_t* ack[XCHG_MAX_WORLD];       // rank g's acknowledgement word for (slot, this rank)
};
struct PushArgs {                        // sending end
    void* dst[XCHG_MAX_WORLD];           // rank g's mailbox part for (slot, this rank)
    uint32_t* flag[XCHG_MAX_WORLD];      // rank g's arrival flag for (slot, this rank)
    const uint32_t* my_acks;             // local acknowledgement words of this slot, one per receiving rank
    uint32_t* tickets;                   // local CTA counters of this slot, one per receiving rank (large payloads)
    uint32_t epoch;
};
void launch_merge_parts(const void* in_idx, const void* in_score, int64_t idx_stride_bytes, int64_t score_stride_bytes, int parts, int64_t nq, int k,
                        int64_t* out_idx, float* out_score, cudaStream_t st, const MergeSync* sync = nullptr);
void launch_exchange_push(const void* src, int64_t bytes, const PushArgs& a, int world, cudaStream_t st);
// ---- sort.cu ------------------------------------------------------------------------------------
// Full ranking (K == N): stable segmented radix sort of c exact score rows; writes columns q0..q0+c of
// out_ranks [n][nq_total] int64 (+ id_offset) and, optionally, out_scores [n][nq_total] fp32.
size_t rank_all_work_bytes(int c, int64_t n);
void launch_rank_all(const float* scores, int64_t pitch, int c, int64_t n, int q0, int nq_total, int64_t id_offset,
                     void* work, int64_t* out_ranks, float* out_scores, cudaStream_t st);

// ---- graph.cu -----------------------------------------------------------------------------------
// mutual[i][j] = 1 iff j >= 1 and i is among the kd neighbours of ids[i][j]  (diffusion.py:107-108)
void launch_mutual_knn(const int64_t* ids64, int32_t* ids32_scratch, int64_t n, int kd, uint8_t* mutual, cudaStream_t st);
void launch_ids_to_i32(const int64_t* ids64, int32_t* ids32, int64_t count, cudaStream_t st);

// diffusion.cu -- truncated CG solves (one CTA per database row)
cudaError_t diffusion_cg_grid(int T, int* grid_out);
cudaError_t launch_diffusion_cg(const int64_t* indptr, const int32_t* indices, const float* values, int64_t n,
                                const int32_t* trunc_ids, int64_t rows, int T, int stride, int maxiter, double atol,
                                uint16_t* s_cols, float* s_vals, int grid, float* out, int* bad, cudaStream_t st);

}  // namespace xs
