/*
 * xs_b200.h -- C ABI of the B200-native exhaustive (exact top-K) matcher.
 *
 * This is the drop-in boundary for ONE path of YYao-42/Image-Search-Engine-for-Historical-Research:
 * L2-normalised global descriptors scored as scores = vecs.T @ qvecs and ranked to the top K.
 * The reference has no FFI of its own (it is 100 % Python); the entry points below are what a
 * ctypes binding for its three call shapes needs, and each one names the reference interface it
 * replaces (paths relative to the reference checkout):
 *
 *   B1  matching_L2(K, train[N,D], test[Q,D]) -> (idx int64[Q,K], time_per_query)
 *                                                   src/utils/nnsearch.py:687-706
 *   B2  scores = np.dot(vecs.T, qvecs); ranks = np.argsort(-scores, axis=0)
 *                                                   src/main_retrieve.py:175-176
 *   B3  KNN(database, 'cosine').search(queries, k) -> (sims f32[nq,k], ids int64[nq,k])
 *                                                   src/utils/knn.py:8-40
 *       and its N x N use for the diffusion graph   src/utils/diffusion.py:67
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a
 * non-zero code otherwise, with a thread-local message behind xs_last_error(); no exception
 * crosses the boundary.  Host entry points block until results are in the caller's buffers;
 * *_dev entry points enqueue on the given CUDA stream and return.  There is no CPU fallback:
 * without a CUDA device every compute entry point fails with XS_ERR_CUDA.
 */
#ifndef XS_B200_H
#define XS_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define XS_API __attribute__((visibility("default")))
#else
#define XS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xs_index xs_index;

enum { XS_OK = 0, XS_ERR_ARG = 1, XS_ERR_CUDA = 2, XS_ERR_NOMEM = 3, XS_ERR_UNSUPPORTED = 4 };
enum { XS_F32 = 0, XS_F64 = 1 };

/* Query statistics of the last search on an index (see xs_index_stats). */
typedef struct xs_stats {
    int64_t n_queries;       /* queries in the last call                                          */
    int64_t n_exact_rerun;   /* queries whose bf16 candidate set could not be certified and were   */
                             /* re-run on the exact fp32 path                                     */
    int64_t n_candidates;    /* total candidates rescored in fp32 (sum over queries)              */
    int32_t path;            /* 1 = batch-1 HBM scan, 2 = tcgen05 GEMM + fused top-K, 3 = exact   */
    int32_t gpu_launches;    /* kernels launched by the last call                                 */
    float   ms_coarse;       /* device time of the coarse kernel (scan or GEMM), CUDA events      */
    float   ms_total;        /* device time of the whole call                                     */
} xs_stats;

/* Thread-local description of the last failure on this thread ("" if none). */
XS_API const char* xs_last_error(void);
/* ABI version of this library (bumped on any signature change). */
XS_API int xs_abi_version(void);
/* Number of visible CUDA devices (0 and XS_ERR_CUDA if the driver is absent). */
XS_API int xs_device_count(int* count);
/*
 * Process-wide defaults picked up by xs_index_create* (no reference counterpart):
 *   "rotation"       1 (default) / 0: apply the seeded random rotation (sign flips + Walsh-Hadamard, twice) to rows
 *                    and queries before they are rounded to bf16, so the rounding errors cannot line up with
 *                    structure in the data (the exact stage always uses the unrotated fp32 values)
 *   "rotation_seed"  the 32-bit seed of that rotation; results never depend on it, only which inputs could defeat
 *                    the statistical certificate does
 *   "compact"        1 (default; XS_KEEP_ROWMAJOR=1 in the environment = 0): keep ONE bf16 copy of the database, in
 *                    the tiled layout the GEMM's TMA boxes read; the batch-1 scan reads it too and xs_self_knn sends
 *                    its rows through the ordinary query preparation.  0: also keep the row-major copy (2 * d_pad
 *                    bytes per row more).  Applies to xs_index_create* and to xs_index_load of a two-copy image; an
 *                    image written by a compact index always loads compact.  Results do not depend on it.
 */
XS_API int xs_config_set(const char* name, double value);

/*
 * Build a device-resident index from a HOST matrix of n rows x d columns.
 *   replaces: the implicit "index" of matching_L2 (the raw train matrix, nnsearch.py:687-698)
 *             and BaseKNN.__init__/add (fp32 copy + IndexFlatIP.add, knn.py:8-23,34-40).
 * db           element (r, c) lives at db[r*stride_row + c*stride_col] (strides in ELEMENTS).
 *              Accepted: row-major (stride_col == 1) and the reference's F-order view vecs.T of a
 *              (D,N) C array (stride_row == 1, stride_col >= n).  The caller keeps ownership;
 *              the matrix is not modified.
 * dtype        XS_F32 or XS_F64 (online.py:96-100 accidentally builds float64).
 * renormalise  1: divide every row by its L2 norm first (matching_L2 semantics, :693-697);
 *              0: use rows as given (KNN / np.dot semantics).
 * id_offset    added to every returned id (row-sharding across GPUs: shard g passes its first row).
 * Device layout: bf16 copies of the (rotated) rows for coarse scoring + fp32 row-major copy (exact rescoring).
 * d <= 4096 (XS_ERR_UNSUPPORTED beyond: the rescoring kernels stage whole rows in shared memory).
 */
XS_API int xs_index_create(const void* db, int dtype, int64_t n, int d,
                    int64_t stride_row, int64_t stride_col,
                    int device, int renormalise, int64_t id_offset, xs_index** out);

/* Same, from a DEVICE fp32 row-major [n, d] matrix on `device` (rows contiguous). */
XS_API int xs_index_create_dev(const float* db_dev, int64_t n, int d,
                        int device, int renormalise, int64_t id_offset, xs_index** out);

XS_API int xs_index_destroy(xs_index* index);

/*
 * On-disk image of an index: the arrays exactly as they live in HBM (fp32 rows, the bf16 rotated vectors in the tiled
 * layout -- plus their row-major form when the index keeps it, see "compact" --, the statistics behind the error band,
 * the rotation seed): 6 * d_pad bytes per row for a compact index.
 *   replaces: save_path_feature / load_path_features pickles and the .pt distractor file      src/utils/general.py:67-92
 *             + the unpickle / concatenate / transpose of every start-up                       src/online.py:93-102
 * xs_index_load is a straight upload -- reader threads fill a ring of pinned buffers, one cudaMemcpyAsync per 32 MB slice,
 * no layout, conversion or tiling kernel -- and yields an index indistinguishable from the one that was saved.
 */
XS_API int xs_index_save(xs_index* index, const char* path);
XS_API int xs_index_load(const char* path, int device, int64_t id_offset, xs_index** out);

/*
 * A second search lane over the same database (no reference counterpart).  The clone shares the read-only
 * database arrays of `src` (they are freed when the last of the index and its clones is destroyed, in any
 * order) and owns its workspaces, stream, tunables and statistics, so searches on the index and on the clone
 * may run concurrently on two streams: the latency-bound selection / rescoring of one batch then overlaps the
 * database scan of the next.  Costs workspace memory only.
 */
XS_API int xs_index_clone(xs_index* src, xs_index** out);

/* n rows, d columns, device ordinal and device bytes held; any out pointer may be NULL. */
XS_API int xs_index_info(const xs_index* index, int64_t* n, int* d, int* device, int64_t* device_bytes);
XS_API int xs_index_stats(const xs_index* index, xs_stats* out);

/*
 * Exact top-k by inner product, HOST buffers (the end-to-end call).
 *   replaces: the per-query loop of matching_L2 (nnsearch.py:699-703) with renormalise_q = 1,
 *             np.dot + argsort[:k] (main_retrieve.py:175-176) and IndexFlatIP.search
 *             (knn.py:25-31) with renormalise_q = 0.
 * q            nq x d host matrix, element (r,c) at q[r*stride_row + c*stride_col]; same two
 *              layouts as xs_index_create.
 * out_idx      [nq, k] int64, best first; ids are row numbers + id_offset.
 * out_score    [nq, k] fp32 exact scores (may be NULL).
 * Ordering: descending fp32 score; exact ties by ascending id.  k must be <= n.
 */
XS_API int xs_search(xs_index* index, const void* q, int dtype, int64_t nq,
              int64_t stride_row, int64_t stride_col, int renormalise_q, int k,
              int64_t* out_idx, float* out_score);

/*
 * Same with DEVICE buffers on the index's device, enqueued on `stream` (a cudaStream_t).
 * q_dev is fp32 row-major [nq, d].  out_status_dev (may be NULL) receives one int32 per query:
 * 0 = certified by the bf16 coarse pass, 1 = needs the exact path -- the CALLER must then re-run those
 * queries (xs_set_param "force_path" 3) before using them.  With out_status_dev == NULL the call
 * synchronises the stream once and re-runs uncertified queries itself.
 */
XS_API int xs_search_dev(xs_index* index, const float* q_dev, int64_t nq, int renormalise_q, int k,
                  int64_t* out_idx_dev, float* out_score_dev, int32_t* out_status_dev,
                  void* stream);

/*
 * Self-kNN over database rows [q_begin, q_end): top-k neighbours of each row among ALL rows.
 *   replaces: self.knn.search(self.features, n_trunc)              src/utils/diffusion.py:67
 * Row i's own id is always returned first (get_affinity relies on ids[i][0] == i, :108).
 * Outputs are HOST buffers [q_end-q_begin, k].
 */
XS_API int xs_self_knn(xs_index* index, int64_t q_begin, int64_t q_end, int k,
                int64_t* out_idx, float* out_score);

/*
 * Average-query-expansion re-score, on the device end to end.
 *   replaces: feature_enhancement + re-ranking inside qge1 / QGE   src/utils/Reranking.py:195-208, 287-306
 *             (the second database scan of every web request, src/online.py:148)
 * top_ids      HOST [nq, kq] int64: the kq best ids of each query from a previous search (ranks[:kq].T).
 * The new query is normalise( sum_j ((kq-j)/kq)^w * row(top_ids[q][j]) ), divided by (norm + 1e-6) as at
 * Reranking.py:203, built in float64 from the index's fp32 rows; it is then searched like xs_search
 * (no further normalisation).  out_queries (may be NULL): HOST [nq, d] fp32, the expanded queries.
 */
XS_API int xs_aqe_search(xs_index* index, const int64_t* top_ids, int64_t nq, int kq, double w, int k,
                         int64_t* out_idx, float* out_score, float* out_queries);

/*
 * Full ranking (k == n) for nq host queries: out_ranks is [n, nq] int64, one COLUMN per query,
 * best first -- the exact shape of `ranks` at main_retrieve.py:176 / evaluate.py:52-55.
 */
XS_API int xs_rank_all(xs_index* index, const void* q, int dtype, int64_t nq,
                int64_t stride_row, int64_t stride_col, int renormalise_q,
                int64_t* out_ranks, float* out_scores_sorted);

/*
 * Merge per-shard results after an all-gather (row-sharded multi-GPU search).
 * in_idx / in_score: DEVICE [n_parts, nq, k] (part-major, as all_gather lays them out);
 * out: DEVICE [nq, k].  Descending score, ties by ascending id.  Enqueued on `stream`.
 */
XS_API int xs_merge_candidates(int device, const int64_t* in_idx, const float* in_score,
                        int n_parts, int64_t nq, int k,
                        int64_t* out_idx, float* out_score, void* stream);

/*
 * Same merge for lists that were gathered as ONE packed buffer per rank (ids | scores | certificate bits back to
 * back): part p's ids start idx_part_stride BYTES after part p-1's, its scores score_part_stride bytes, its
 * certificate words (int32 per query, may be NULL) status_part_stride bytes.  out_status_dev (may be NULL)
 * receives the OR over the parts: a merged query is certified only if every shard certified its list.
 */
XS_API int xs_merge_candidates_strided(int device, const void* in_idx, const void* in_score, const void* in_status,
                                       int64_t idx_part_stride, int64_t score_part_stride, int64_t status_part_stride,
                                       int n_parts, int64_t nq, int k,
                                       int64_t* out_idx, float* out_score, int32_t* out_status_dev, void* stream);

/*
 * Peer exchange: the all-gather + merge of the row-sharded search without a collective library on the data
 * path.  Every rank owns a mailbox in its own HBM, exported through CUDA IPC.  The SENDING end is fused into the
 * search: xs_search_dev_push runs the local search and its last kernel stores every query's k results (ids,
 * scores, certificate word) straight into all ranks' mailboxes with plain stores over the NVLink peer mappings,
 * then releases that query's arrival flag on every rank.  The MERGE kernel of each rank waits per query for the
 * world's flags, merges, and acknowledges the slot so that it can be overwritten.  xs_exchange_push is the
 * sending end as a kernel of its own, for a packed result produced elsewhere (xs_search_dev into a buffer of
 * xs_exchange_part_bytes(nq, k) bytes laid out  ids int64 [nq*k] | scores f32 [nq*k] | status int32 [nq]).
 * One process per GPU; the 64-byte handles travel through whatever channel the host processes share.
 * (No reference counterpart: the reference is single-process.  Replaces ncclAllGather + xs_merge_candidates
 * of SURVEY.md section 8e.)
 *
 * Protocol: every rank calls search_dev_push|push(slot) / merge(slot) in the same order; slots 0 and 1 alternate,
 * and the merge of a slot must be enqueued before the next push into the same slot (at most two searches in
 * flight).  Enqueue the merge on the stream of the push (or behind it): then the only thing a merge kernel ever
 * waits for is another GPU.  A peer that never arrives traps the waiting kernel after ~20 s instead of hanging
 * the GPU.  xs_exchange_destroy synchronises the device; the caller must barrier across ranks before calling it.
 */
typedef struct xs_exchange xs_exchange;
XS_API int64_t xs_exchange_part_bytes(int64_t nq, int k);
XS_API int xs_exchange_create(int device, int world, int rank, int64_t max_queries, int k_max, xs_exchange** out,
                              unsigned char* handle_out /* 64 bytes */);
XS_API int xs_exchange_connect(xs_exchange* ex, const unsigned char* handles /* world x 64 bytes, by rank */);
XS_API int xs_search_dev_push(xs_index* index, const float* q_dev, int64_t nq, int renormalise_q, int k,
                              xs_exchange* ex, int slot, void* stream);
/* search + push + merge in one call: equivalent to xs_search_dev_push followed by xs_exchange_merge on the same stream.  For
 * batches of at most 128 queries the merge rides in the search's last kernel too: the CTA that emitted a query waits for
 * the other ranks' lists of that query and merges them, so a sharded step is three launches and the only wait is on
 * another GPU. */
XS_API int xs_search_dev_exchange(xs_index* index, const float* q_dev, int64_t nq, int renormalise_q, int k, xs_exchange* ex, int slot,
                                  int64_t* out_idx_dev, float* out_score_dev, int32_t* out_status_dev, void* stream);
XS_API int xs_exchange_push(xs_exchange* ex, const void* packed_dev, int64_t nq, int k, int slot, void* stream);
XS_API int xs_exchange_merge(xs_exchange* ex, int slot, int64_t nq, int k, int64_t* out_idx_dev, float* out_score_dev,
                             int32_t* out_status_dev, void* stream);
XS_API int xs_exchange_destroy(xs_exchange* ex);

/*
 * Two-slot pipeline for the sharded (or single-shard) search: ONE host call enqueues a whole step -- local search with the
 * exchange's sending end fused in, merge, copy of the merged certificate words to pinned memory -- on one of two lane
 * streams (lanes = 2: the second lane is an internal clone of the index); a second call waits for the step and returns
 * device pointers to the merged [nq, k] result (valid until the slot's next submit) and the uncertified queries, which
 * the caller re-runs (collectively when sharded: every rank sees the same merged certificate words).  At most two steps
 * in flight; with an exchange every rank submits and collects in the same order.  ex may be NULL (no exchange step).
 * (No reference counterpart: it exists because at eight GPUs a step is shorter than the host work of driving it.)
 */
typedef struct xs_pipeline xs_pipeline;
XS_API int xs_pipeline_create(xs_index* index, xs_exchange* ex, int64_t nq_max, int k_max, int lanes, xs_pipeline** out);
XS_API int xs_pipeline_submit(xs_pipeline* p, const float* q_dev, int64_t nq, int k, void* caller_stream, int* slot_out);
XS_API int xs_pipeline_collect(xs_pipeline* p, int slot, void* caller_stream, int64_t** out_idx_dev, float** out_score_dev,
                               int64_t* n_flagged, int32_t* flagged);
XS_API int xs_pipeline_destroy(xs_pipeline* p);
/* Debugging aid: the index behind a lane of the pipeline (lane 1 = the internal clone; owned by the pipeline). */
XS_API xs_index* xs_pipeline_lane(xs_pipeline* p, int lane);

/*
 * Mutual-kNN test of the diffusion affinity graph.
 *   replaces: the per-row loop `np.isin(ids[ids[i]], i).any(axis=1)` of get_affinity   src/utils/diffusion.py:106-108
 * ids: HOST [n, kd] int64, the kNN lists of every row (slot 0 = the row itself, as xs_self_knn returns them).
 * out_mutual: HOST [n, kd] bytes, 1 where slot j >= 1 is a mutual neighbour, 0 elsewhere (slot 0 always 0).
 */
XS_API int xs_mutual_knn(int device, const int64_t* ids, int64_t n, int kd, uint8_t* out_mutual);

/*
 * Diffusion graph on the device: mutual-kNN affinity and the normalised Laplacian from kNN lists.
 *   replaces: Diffusion.get_affinity + get_laplacian (scipy on the host)              src/utils/diffusion.py:87-116
 * sims / ids: HOST [n, kd] (slot 0 = the row itself, as xs_self_knn returns them; sims are not modified).
 * A[i, ids[i][j]] = max(sims[i][j], 0)^gamma for mutual slots j >= 1; L = I - alpha D^-1/2 A D^-1/2 with float32
 * entries computed in the reference's order.  Outputs (HOST, ELL form): out_cols int32 [n, kd], out_vals f32 [n, kd],
 * out_cnt int32 [n] -- row i has out_cnt[i] entries, the diagonal first; out_affinity (may be NULL) f32 [n, kd], 0 at
 * non-mutual slots.
 */
XS_API int xs_diffusion_laplacian(int device, const float* sims, const int64_t* ids, int64_t n, int kd, double alpha, double gamma,
                                  int32_t* out_cols, float* out_vals, int32_t* out_cnt, float* out_affinity);

/*
 * The whole gallery side of the diffusion re-ranking without leaving the device.
 *   replaces: Diffusion.get_offline_results up to the final csr_matrix merge           src/utils/diffusion.py:52-85
 * N x N self-kNN truncated at n_trunc (row i's own id first) -> mutual-kNN graph of the first kd neighbours ->
 * Laplacian -> one truncated CG per row on L[ids][:, ids] x = e_0 (see xs_diffusion_cg).  The lists never visit the
 * host between the stages.  out_ids int64 [n, n_trunc], out_sims f32 [n, n_trunc] (may be NULL), out_scores f32
 * [n, n_trunc]: HOST.  n_trunc <= 4096.
 */
XS_API int xs_diffusion_offline(xs_index* index, int n_trunc, int kd, double alpha, double gamma, int maxiter, double tol,
                                int64_t* out_ids, float* out_sims, float* out_scores);

/*
 * Gallery-side diffusion: one truncated conjugate-gradient solve per database row.
 *   replaces: get_offline_result (`lap_alpha[ids][:, ids]`, `linalg.cg(trunc_lap, trunc_init, tol=1e-6, maxiter=20)`)
 *             src/utils/diffusion.py:15-19 and the joblib loop over all rows :74-76
 * indptr/indices/values: HOST CSR of the n x n Laplacian (int64 / int32 / float32; canonical: no repeated
 *   columns inside a row).  trunc_ids: HOST [rows, n_trunc] int64, each row's truncation set (distinct ids in
 *   [0, n); slot 0 carries the right-hand side e_0 -- the row itself in the reference).  n_trunc <= 4096.
 * Solves  L[ids][:, ids] x = e_0  from x0 = 0 with plain CG in fp64: stops before a step once
 * ||r|| < tol (||b|| = 1), else returns the iterate after `maxiter` steps.  out_scores: HOST [rows, n_trunc]
 * float32 (the dtype the reference stores them in, diffusion.py:81-83).
 */
XS_API int xs_diffusion_cg(int device, const int64_t* indptr, const int32_t* indices, const float* values, int64_t n,
                           const int64_t* trunc_ids, int64_t rows, int n_trunc, int maxiter, double tol,
                           float* out_scores);

/*
 * Tunables (set before searching; all have safe defaults):
 *   "eps_sigmas"   float  width of the bf16 error band in standard deviations      (8.0)
 *   "certificate"  int    0 = statistical band: eps_sigmas standard deviations of the bf16 rounding error, made
 *                             data-independent by the random rotation, plus a model check on every rescored
 *                             candidate; 1 = worst-case band from stored residual norms (Cauchy-Schwarz, no
 *                             assumption; about 3x the rescored candidates)                                 (0)
 *   "inline_boot"  int    1 = batches <= 128 queries bootstrap their score threshold inside the GEMM launch  (1)
 *   "scan_max_q"   int    largest batch served by the batch-1 HBM scan kernel      (1)
 *   "force_path"   int    0 = auto, 1 = scan, 2 = tcgen05 GEMM, 3 = exact fp32      (0)
 *   "gemm_splits"  int    database splits per query tile, 0 = auto                 (0)
 *   "timing"       int    1 = record CUDA events so xs_index_stats reports ms_coarse/ms_total (0)
 *   "pair_mode"    int    1 = CTA-pair (cta_group::2) GEMM shape for batches > 128 queries   (1)
 *   "sample_pass"  int    1 = threshold bootstrap pass before the GEMM                    (1)
 *   "scan_tiled"   int    1 = the batch-1 scan reads the tiled bf16 array (always, on a compact index)   (1)
 *   "self_lanes"   int    2 = xs_self_knn alternates its batches between the index and an internal clone (1)
 */
XS_API int xs_set_param(xs_index* index, const char* name, double value);

/* Debugging aid (after xs_set_param "boot_trace" 1): per-CTA globaltimer stamps of the last small-batch search --
 * which = 0: the GEMM launch that bootstrapped its threshold in-kernel, rows of 8; which = 1: the fused finalise
 * launch, rows of 10.  *rows receives the number of rows copied (0 if nothing was recorded). */
XS_API int xs_debug_trace(xs_index* index, int which, unsigned long long* out, int max_rows, int* rows);

#ifdef __cplusplus
}
#endif
#endif /* XS_B200_H */
