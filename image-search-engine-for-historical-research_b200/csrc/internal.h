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
constexpr int HIST_BINS = 4096;   // database-wide score histogram filled by the scoring kernels (linear bins, scan.cu)

// status bits written by the finalise kernel (one int32 per query)
constexpr int ST_UNCERTIFIED = 1;

struct DevStats {                 // device-side database statistics (uint-ordered positive floats)
    unsigned int v4max_bits;      // max over rows of ||v||_4
    unsigned int vnmax_bits;      // max over rows of ||v||_2
    unsigned int rhomax_bits;     // max over rows of ||v' - bf16(v')||_2  (v' = rotated row): worst-case band
};

// ---- build.cu -----------------------------------------------------------------------------------
// src_tile is a device buffer holding `rows` database rows either row-major ([rows][pitch]) or
// column-major ([d][pitch], the reference's (D,N) layout); writes fp32 rows [rows][d_pad].
void launch_layout_rows(const void* src_tile, int dtype, bool colmajor, int64_t pitch,
                        int64_t rows, int d, int d_pad, float* dst32, cudaStream_t st);
// Optional L2 normalisation (in place), then the bf16 copy of the ROTATED row (build.cu: random rotation) and the
// database statistics behind eps.  dst16 rows have pitch d_pad.
void launch_finish_rows(float* rows32, __nv_bfloat16* rows16, int64_t rows, int d_pad, bool renorm, bool rotate, uint32_t seed,
                        DevStats* stats, cudaStream_t st);
// Query preparation in one launch: raw (optional, fp32 row-major, pitch d) -> q32 [nq][d_pad] (normalised when asked),
// q16 [nq_pad][d_pad] bf16 of the rotated queries (rows >= nq zero), q32r fp32 rotated queries (batch-1 scan), eps[q]
// (eps_mode 0: statistical band, 1: worst-case band).  q16 / q32r may be null.
void launch_prep_queries(const float* raw, int d, float* q32, __nv_bfloat16* q16, float* q32r, int64_t nq, int64_t nq_pad, int d_pad,
                         bool renorm, bool rotate, uint32_t seed, const DevStats* stats, float eps_sigmas, int eps_mode, float* eps,
                         cudaStream_t st);

// AQE (Reranking.py:195-208): q_out[q] = normalise( sum_j ((kq-j)/kq)^w * db32[top_ids[q][j]] ), float64 inside.
void launch_aqe_queries(const float* db32, const int64_t* top_ids, int64_t nq, int kq, double w, int64_t n, int d_pad,
                        float* q_out, cudaStream_t st);

// Tiled copy of the bf16 database for the GEMM's B operand: [n_pad/256][d_pad/64][256][64], so that every TMA box
// (128 or 256 rows x 64 columns) is one contiguous 16/32 KB run in HBM instead of 128-byte pieces 4 KB apart.
// `rows16`: `rows` row-major bf16 rows that are rows r0 .. r0 + rows of the database.
void launch_tile_db16(const __nv_bfloat16* rows16, __nv_bfloat16* db16t, int64_t r0, int64_t rows, int d_pad, cudaStream_t st);

// ---- scan.cu ------------------------------------------------------------------------------------
// Batch-1 HBM scan: scores[q][row] = <db16[row], q32[q]>, fp32 accumulate.  Both scoring kernels also
// add every score to the linear histogram ghist[q][.] (zeroed by the caller).
//   The histogram is linear over [-B, B], B = ||q|| max||v||; the kernels publish B per query in bounds[q] for scores_to_pools.
void launch_scan_scores(const __nv_bfloat16* db16, bool tiled, const float* q32, int nq, int64_t n, int d_pad,
                        float* scores, int64_t score_pitch, uint32_t* ghist, const DevStats* stats, float* bounds, int num_sms, cudaStream_t st);
// Exact scoring: scores[q][row] = fp32( sum_fp64 db32[row][i] * q32[q][i] ).  Any nq (looped in 4s).
void launch_exact_scores(const float* db32, const float* q32, int nq, int64_t n, int d_pad,
                         float* scores, int64_t score_pitch, uint32_t* ghist, const DevStats* stats, float* bounds, int num_sms, cudaStream_t st);
// Score matrix -> candidate pools (one partial list per SLICE_ROWS rows).
//   exact = false: keep items within the eps band below the slice's k-th best
//   exact = true : keep exactly the slice's k best (full 64-bit item order)
void launch_scores_to_pools(const float* scores, int64_t score_pitch, int nq, int64_t n, int k,
                            const float* eps, const uint32_t* ghist, const float* bounds, bool exact, uint64_t* pool_items, int* pool_count,
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
    int inline_boot;  // 1: one job per CTA; the first tile of every CTA is the threshold sample (no separate bootstrap launch)
    int half_units;   // 1: single-CTA shapes deal the database in half tiles of 128 rows (jobs differ by at most half a tile)
    int stages3;      // 1: single-CTA shape with a 3-stage operand ring (leaves shared memory for a co-resident finalise CTA)
};
// In-kernel threshold bootstrap of gemm_topk_kernel (plan.inline_boot): the CTAs meet once through these words.
constexpr int BOOT_MAX_GRID = 160;        // CTAs (= sample lists per query) the selecting warp holds in registers
struct InlineBoot {
    int on;
    float* samp;                  // [128 queries][grid][8] best scores of every CTA's first tile
    uint32_t* arrive;             // CTA arrival counter (monotonic across launches)
    uint32_t arrive_target;       // counter value once every CTA of THIS launch has arrived
    uint32_t* published;          // thresholds published so far (monotonic across launches)
    uint32_t published_target;    // its value once every query of THIS launch has its threshold
    uint64_t* thr_pub;            // [128] (epoch << 32 | threshold bits), written by the CTA that owns the query
    uint32_t epoch;               // distinguishes this launch's thresholds from the previous launch's
    unsigned long long* trace;    // optional [grid][8] globaltimer stamps (debugging: xs_set_param "boot_trace")
};
GemmPlan plan_gemm(int64_t nq, int64_t n_pad, int k, int num_sms, int forced_splits, bool allow_pair);
GemmPlan plan_gemm_sample(const GemmPlan& main_plan, int num_sms, int k);
// tmap_q: [nq_pad128][d_pad] bf16, box {64,128};  tmap_db: [n_pad][d_pad] bf16, box {64,256} (box {64,128} when plan.pair)
//   tmap_db_half: the 128-row-box map of the same database (for plan.half_units; may be null)
cudaError_t launch_gemm_topk(const CUtensorMap& tmap_q, const CUtensorMap& tmap_db, const CUtensorMap* tmap_db_half, const GemmPlan& plan,
                             int64_t nq, int64_t n_valid, int d_pad, int k, const float* eps,
                             uint64_t* pool_items, int* pool_count, uint32_t* pool_thr, int a_row0,
                             const float* thr0, const InlineBoot* boot, cudaStream_t st);

// ---- finalise.cu --------------------------------------------------------------------------------
constexpr int XCHG_MAX_WORLD = 16;
// Peer exchange, sending end fused into the emit step: the K results of a query go straight into every rank's mailbox
// (NVLink peer mappings; g == this rank is the local mailbox) and the query's arrival flag is released per rank.
struct PushTarget {
    int world;                           // 0: off
    int64_t* ids[XCHG_MAX_WORLD];        // rank g's mailbox part for (slot, this rank): ids [nq][k]
    float* scores[XCHG_MAX_WORLD];       //   scores [nq][k]
    int32_t* status[XCHG_MAX_WORLD];     //   certificate bits [nq]
    uint32_t* flags[XCHG_MAX_WORLD];     // rank g's per-query arrival flags for (slot, this rank)
    const uint32_t* my_acks;             // local acknowledgement words of this slot, one per receiving rank
    uint32_t epoch;
};
// Receiving end of the peer exchange fused into the same kernel (cluster form only): after a query's own results have gone
// out, its emitting CTA waits for the other ranks' lists of THAT query, merges world x k -> k and writes the final answer.
struct MergeTarget {
    int on;
    const char* base;                    // this rank's mailbox, slot base: part g at base + g * part_bytes (ids | scores | status)
    int64_t part_bytes;
    const uint32_t* flags;               // local per-query arrival flags of this slot: flags[g * flag_stride + q]
    int64_t flag_stride;
    int64_t nq_total;                    // queries of the whole call (layout of a part)
    int64_t q0;                          // first query of this launch within the call
    uint32_t* ticket;                    // local counter of merged queries (the last one acknowledges the slot)
    uint32_t* ack[XCHG_MAX_WORLD];       // rank g's acknowledgement word for (slot, this rank)
    int64_t* out_idx; float* out_score; int32_t* out_status;    // [nq_total][k] / [nq_total]
};
struct FinaliseArgs {
    const uint64_t* pool_items; const int* pool_count; const uint32_t* pool_thr;
    int P, cap;
    const float* db32; const float* q32; int d_pad;
    const float* eps;           // per-query band (ignored when exact)
    int k; bool exact;          // exact: pool scores are already fp32-exact -> no rescoring
    int64_t id_offset;
    int64_t self_base;          // >= 0: query q is database row self_base + q and must rank first
    int64_t* out_idx; float* out_score; int* status; int* n_cand;   // out_idx / out_score / status may be null when push is on
    int64_t out_pitch;          // elements between consecutive queries in out_idx / out_score
    void* work;                 // finalise_work_bytes(nq, k, cand_max) of scratch for the multi-CTA-per-query forms
    int* ticket;                // [3 * nq] zero-initialised words (completion tickets, candidate counters, flags; self-resetting)
    int cand_max;               // candidates per query that can be rescored (power of two); 0 = default for k
    int per_sm;                 // cluster form: 0 / 3 = latency mode (three CTAs per SM), 1 = throughput mode (one slim CTA per SM, next to a GEMM CTA)
    uint64_t* w_cand; int* w_ncand; int* w_flag; int* w_ticket;   // carved out by launch_finalise
    PushTarget push;
    MergeTarget merge;          // merge.on requires push.world > 0
    unsigned long long* trace;  // optional [nq * ctas_per_query][10] globaltimer stamps of the fused form (debugging)
};
int  finalise_cand_max(int k, int mode);                  // mode 0: statistical band, 1: worst-case band (more candidates)
size_t finalise_work_bytes(int64_t nq, int k, int cand_max);
int  finalise_launches(const FinaliseArgs& a, int64_t nq);
bool finalise_can_merge(int64_t nq, int k, int world, int P, int d_pad);
bool launch_finalise(const FinaliseArgs& a, int64_t nq, cudaStream_t st);     // true: the exchange's merge rode in the same launch (a.merge)
// thr0[q] = (k-th best pooled coarse score) - 2 eps[q], one ulp lower; -inf when fewer than k items.
void launch_sample_threshold(const uint64_t* pool_items, const int* pool_count, int P, int cap, int k,
                             const float* eps, float* thr0, int64_t nq, cudaStream_t st);
// Multi-GPU merge of [parts][nq][k] lists.
struct MergeSync {                       // receiving end of the peer exchange (all zero = plain merge)
    const uint32_t* flags;               // local per-query arrival flags of this slot: flags[g * flag_stride + q]
    int64_t flag_stride;
    uint32_t epoch;
    uint32_t* ticket;                    // local CTA counter of this slot
    uint32_t* ack[XCHG_MAX_WORLD];       // rank g's acknowledgement word for (slot, this rank)
};
struct PushArgs {                        // sending end as a kernel of its own (whole packed payload)
    void* dst[XCHG_MAX_WORLD];           // rank g's mailbox part for (slot, this rank)
    uint32_t* flag[XCHG_MAX_WORLD];      // rank g's per-query arrival flags for (slot, this rank)
    const uint32_t* my_acks;             // local acknowledgement words of this slot, one per receiving rank
    uint32_t* tickets;                   // local CTA counters of this slot, one per receiving rank (large payloads)
    uint32_t epoch;
    int64_t nq;                          // flags [0, nq) are released
};
// part p: ids at in_idx + p*idx_stride BYTES, scores at in_score + p*score_stride BYTES, optional certificate bits at
// in_status + p*status_stride BYTES; out_status[q] = OR over the parts.
void launch_merge_parts(const void* in_idx, const void* in_score, const void* in_status, int64_t idx_stride_bytes, int64_t score_stride_bytes,
                        int64_t status_stride_bytes, int parts, int64_t nq, int k,
                        int64_t* out_idx, float* out_score, int32_t* out_status, cudaStream_t st, const MergeSync* sync = nullptr);
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

// Diffusion graph on the device (diffusion.py:87-116): mutual test on the first kd columns of ids32 [n][id_pitch], affinity
// max(sim,0)^gamma, degrees, and the Laplacian I - alpha D^-1/2 A D^-1/2 in ELL form (kd slots per row, diagonal first,
// padding column INT32_MIN); indptr (optional) = i * kd so that the ELL arrays read as CSR for the CG kernel.
void launch_diffusion_graph(const int32_t* ids32, int64_t id_pitch, const float* sims, int64_t sim_pitch, int64_t n, int kd, double alpha, double gamma,
                            uint8_t* mutual, float* aff, float* dinv, int32_t* cols, float* vals, int32_t* cnt, int64_t* indptr, cudaStream_t st);
void launch_ids_strided_to_i32(const int64_t* in, int64_t in_pitch, int32_t* out, int64_t out_pitch, int64_t rows, int cols, int64_t sub, cudaStream_t st);

// diffusion.cu -- truncated CG solves (one CTA per database row)
cudaError_t diffusion_cg_grid(int T, int* grid_out);
cudaError_t launch_diffusion_cg(const int64_t* indptr, const int32_t* indices, const float* values, int64_t n,
                                const int32_t* trunc_ids, int64_t rows, int T, int stride, int maxiter, double atol,
                                uint16_t* s_cols, float* s_vals, int grid, float* out, int* bad, cudaStream_t st);

}  // namespace xs
