"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT.

A numpy restatement of the reference's exhaustive-matching path (the reference itself is pure
numpy here, so numpy is the faithful language for the restatement).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module -- and only as the checker or as the timed CPU baseline.  The product
package never imports it and has no CPU fallback.

Parity status: **pinned against outputs of the reference's own code run in the build
container** -- ``tests/golden/make_golden.py`` extracts ``matching_L2`` (nnsearch.py:687-706),
the two inline ranking lines (main_retrieve.py:175-176) and ``qge1`` (Reranking.py:287-306)
from ``/root/reference`` by AST / line number, imports ``src/utils/evaluate.py`` directly, runs
them on seeded inputs and commits inputs-by-seed + outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function below against those files.  The faiss
boundary (``IndexFlatIP.search``, knn.py:30,36) is **unpinned**: faiss is an un-vendored,
un-versioned dependency (requirements.txt:4 is a commented-out ``faiss==1.5.3``; README.md:19
installs the latest faiss-gpu) that is absent from the image, so ``knn_search`` restates its
published contract (fp32 inner product, k best in descending order, int64 labels).

All ``file:line`` citations are relative to ``/root/reference``.
"""
from __future__ import annotations

import time

import numpy as np


# --------------------------------------------------------------------------------------
# a1 -- matching_L2                                          src/utils/nnsearch.py:687-706
# --------------------------------------------------------------------------------------
def matching_L2(K, embedded_features_train, embedded_features_test):
    """Restates nnsearch.py:687-706 step for step.

    Row-normalise both inputs (:693-698); for each query take the Euclidean distance to every
    train row (:701) and keep the first K of a full ascending argsort (:703).  Returns
    ``(idx int64 [Q,K], time_per_query)`` with the timer placed exactly as in :688,704-705
    (it includes the normalisation).
    """
    t_start = time.time()
    n_train, _ = embedded_features_train.shape
    n_test, _ = embedded_features_test.shape
    idx = np.zeros((n_test, K), dtype=np.int64)
    train_norm = np.expand_dims(np.linalg.norm(embedded_features_train, axis=1), axis=1)
    test_norm = np.expand_dims(np.linalg.norm(embedded_features_test, axis=1), axis=1)
    train = embedded_features_train / train_norm
    test = embedded_features_test / test_norm
    for r in range(n_test):
        d = np.linalg.norm(test[r, :] - train, axis=1)
        idx[r, :] = np.argsort(d)[:K]
    t_end = time.time()
    return idx, (t_end - t_start) / n_test


# --------------------------------------------------------------------------------------
# a2 + a3 -- inline inner-product scoring and full ranking     src/main_retrieve.py:175-176
#            (same two lines at main_train.py:703-704,715-716, Reranking.py:206-207,299-300)
# --------------------------------------------------------------------------------------
def rank_ip(vecs, qvecs):
    """``scores = np.dot(vecs.T, qvecs)`` (:175); ``ranks = np.argsort(-scores, axis=0)`` (:176).

    ``vecs`` is ``(D,N)``, ``qvecs`` is ``(D,Q)``; returns ``(scores (N,Q), ranks int64 (N,Q))``,
    one column per query, best first.  This is the canonical parity oracle of the north star.
    """
    scores = np.dot(vecs.T, qvecs)
    ranks = np.argsort(-scores, axis=0)
    return scores, ranks


def scores_f64(vecs, qvecs):
    """float64 adjudicator for near-ties: the same contraction with every product exact."""
    return np.dot(vecs.T.astype(np.float64), qvecs.astype(np.float64))


def topk_ip(vecs, qvecs, k):
    """Top-k slice of :func:`rank_ip` with an explicit, total tie rule.

    Descending fp32 score, ties broken by ascending id (``np.lexsort``); this is the rule the
    CUDA path implements, written down here so that both sides of a parity test agree on
    exact ties where the reference's introsort leaves the order unspecified.
    Returns ``(ids int64 [Q,k], sims f32 [Q,k])``.
    """
    scores = np.dot(vecs.T, qvecs).astype(np.float32, copy=False)
    n, q = scores.shape
    k = min(k, n)
    ids = np.empty((q, k), dtype=np.int64)
    sims = np.empty((q, k), dtype=np.float32)
    for j in range(q):
        s = scores[:, j]
        if k < n:
            # candidates: everything >= the k-th largest value (keeps all ties at the boundary)
            kth = np.partition(s, n - k)[n - k]
            cand = np.nonzero(s >= kth)[0]
        else:
            cand = np.arange(n)
        order = np.lexsort((cand, -s[cand].astype(np.float64)))[:k]
        ids[j] = cand[order]
        sims[j] = s[cand[order]]
    return ids, sims


# --------------------------------------------------------------------------------------
# a4/a5/a6 -- faiss flat kNN wrapper                                src/utils/knn.py:8-40
# --------------------------------------------------------------------------------------
def knn_search(database, queries, k, method="cosine"):
    """``KNN(database, method).search(queries, k)`` (knn.py:25-40), faiss replaced by numpy.

    ``BaseKNN.__init__`` casts the DB to C-contiguous fp32 (:10-15); ``search`` does the same to
    the queries (:26-29) and returns faiss' ``(sims f32 [nq,k], ids int64 [nq,k])`` (:30-31).
    ``'cosine'`` = ``IndexFlatIP`` (descending inner product), ``'euclidean'`` = ``IndexFlatL2``
    (ascending *squared* L2) (:36-37).  Ties: ascending id.
    """
    db = np.ascontiguousarray(database, dtype=np.float32)
    qs = np.ascontiguousarray(queries, dtype=np.float32)
    if method == "cosine":
        ids, sims = topk_ip(db.T, qs.T, k)
        return sims, ids
    if method == "euclidean":
        d2 = ((qs.astype(np.float64) ** 2).sum(1)[:, None]
              + (db.astype(np.float64) ** 2).sum(1)[None, :]
              - 2.0 * qs.astype(np.float64) @ db.astype(np.float64).T)
        ids = np.empty((qs.shape[0], k), dtype=np.int64)
        dist = np.empty((qs.shape[0], k), dtype=np.float32)
        for j in range(qs.shape[0]):
            order = np.lexsort((np.arange(db.shape[0]), d2[j]))[:k]
            ids[j] = order
            dist[j] = d2[j, order]
        return dist, ids
    raise KeyError(method)


# --------------------------------------------------------------------------------------
# a7 -- AQE feature enhancement + re-score                 src/utils/Reranking.py:287-306
# --------------------------------------------------------------------------------------
def feature_enhancement(it_times, k, ranks, qvecs, vecs, w):
    """Restates the nested helper at Reranking.py:195-208 / :288-301.

    Weighted mean of the top-k DB columns with weights ``((k..1)/k) ** w`` (:197,200), divided by
    ``(norm + 1e-6)`` (:203), then the a2+a3 ranking again (:206-207).  Note the reference never
    feeds ``ranks_aqe`` back into the loop, so every iteration recomputes the same thing; the
    restatement keeps that behaviour.
    """
    for _ in range(it_times):
        qe_weight = (np.arange(k, 0, -1) / k).reshape(1, k, 1)
        top = vecs[:, ranks[:k, 0:int(ranks.shape[1])]]
        qtop = (top * (qe_weight ** w)).sum(axis=1)
        qtop = qtop / (np.linalg.norm(qtop, ord=2, axis=0, keepdims=True) + 1e-6)
        qvecs_qe = qtop
        ranks_aqe = np.argsort(-np.dot(vecs.T, qvecs_qe), axis=0)
    return qvecs_qe, ranks_aqe


def qge1(ranks, qvec, vecs, K):
    """Reranking.py:287-306 -- k=3, w=4, one iteration; returns the full re-ranking."""
    _, ranks_aqe = feature_enhancement(1, 3, ranks, qvec, vecs, 8. / 2)
    return ranks_aqe


# --------------------------------------------------------------------------------------
# a9 -- consumers of `ranks`: junk-adjusted AP / mAP        src/utils/evaluate.py:4-112
# --------------------------------------------------------------------------------------
def compute_ap(pos, nres):
    """Trapezoidal AP over zero-based positive positions (evaluate.py:4-38)."""
    # sequential accumulation in the reference's order (:25-36) so the float64 result is
    # bit-identical, not merely close
    step = 1. / nres
    ap = 0
    for j, r in enumerate(pos):
        before = 1. if r == 0 else float(j) / r
        after = float(j + 1) / (r + 1)
        ap += (before + after) * step / 2.
    return ap


def compute_map(ranks, gnd, kappas=()):
    """mAP, per-query AP, mean precision@kappas (evaluate.py:40-112).

    ``ranks`` is ``(K_or_N, nq)`` zero-based ids, best first.  Positives are the ``ok`` ids, the
    ``junk`` ids are removed from the ranking before positions are counted (:84-94); queries
    with no positives are skipped (:68-72).
    """
    nq = len(gnd)
    aps = np.zeros(nq)
    prs = np.zeros((nq, len(kappas)))
    pr = np.zeros(len(kappas))
    total, nempty = 0.0, 0
    for i in range(nq):
        ok = np.asarray(gnd[i]["ok"])
        if ok.shape[0] == 0:
            aps[i] = np.nan
            prs[i, :] = np.nan
            nempty += 1
            continue
        junk_ids = np.asarray(gnd[i].get("junk", np.empty(0)))
        col = ranks[:, i]
        is_pos = np.isin(col, ok)
        is_junk = np.isin(col, junk_ids)
        # position of each positive once the junk entries ranked before it are dropped
        pos = np.nonzero(is_pos)[0] - np.cumsum(is_junk)[is_pos]
        ap = compute_ap(pos, len(ok))
        total += ap
        aps[i] = ap
        pos1 = pos + 1
        for j, kap in enumerate(kappas):
            kq = min(pos1.max(), kap)
            prs[i, j] = (pos1 <= kq).sum() / kq
        pr += prs[i, :]
    return total / (nq - nempty), aps, pr / (nq - nempty), prs


def protocol_maps(ranks, gnd, kappas=(1, 5, 10)):
    """The Easy / Medium / Hard split of evaluate.py:123-147 -> ``(mapE, mapM, mapH)``."""
    def regroup(ok_keys, junk_keys):
        return [{"ok": np.concatenate([g[k] for k in ok_keys]),
                 "junk": np.concatenate([g[k] for k in junk_keys])} for g in gnd]
    e = compute_map(ranks, regroup(["easy"], ["junk", "hard"]), kappas)[0]
    m = compute_map(ranks, regroup(["easy", "hard"], ["junk"]), kappas)[0]
    h = compute_map(ranks, regroup(["hard"], ["junk", "easy"]), kappas)[0]
    return e, m, h


# --------------------------------------------------------------------------------------
# 8f-2 -- mutual-kNN affinity of the diffusion graph          src/utils/diffusion.py:101-116
# --------------------------------------------------------------------------------------
def mutual_mask(ids):
    """``ismutual`` of diffusion.py:107-108 for every row: slot j of row i is mutual when i occurs in
    the neighbour list of ``ids[i, j]``; slot 0 (the row itself) is cleared."""
    ids = np.asarray(ids)
    out = np.zeros(ids.shape, dtype=bool)
    for i in range(ids.shape[0]):
        m = np.isin(ids[ids[i]], i).any(axis=1)
        m[0] = False
        out[i] = m
    return out


def affinity_dense(sims, ids, gamma=3):
    """Dense restatement of ``get_affinity`` (diffusion.py:101-116) for small N."""
    sims = np.where(sims < 0, 0, sims).astype(np.float32) ** gamma
    n = sims.shape[0]
    a = np.zeros((n, n), dtype=np.float32)
    m = mutual_mask(ids)
    for i in range(n):
        a[i, np.asarray(ids)[i, m[i]]] = sims[i, m[i]]
    return a


# --------------------------------------------------------------------------------------
# 8f-4 -- average query expansion / database augmentation     src/utils/Reranking.py:314-365, 375-440
# --------------------------------------------------------------------------------------
def _db_postprocess(query_vecs, reference_vecs):
    """``postprocess`` (Reranking.py:326-332 / :387-398): move the origin to the mean of all rows of both
    sets, then L2-normalise each set (left untouched when it holds a zero row, :322-324)."""
    center = np.mean(np.concatenate([query_vecs, reference_vecs], axis=0), axis=0)
    out = []
    for v in (query_vecs - center, reference_vecs - center):
        norm = np.expand_dims(np.linalg.norm(v, axis=1), axis=1)
        out.append(v if np.any(norm == 0) else v / norm)
    return out[0], out[1]


def _db_sim_order(query_vecs, reference_vecs):
    """``calculate_sim_matrix`` + ``np.argsort(sim_mat, axis=1)`` (Reranking.py:334-345)."""
    q, r = _db_postprocess(query_vecs, reference_vecs)
    return np.argsort(2 - 2 * np.dot(q, r.T), axis=1)


def average_query_expansion(qvecs, vecs, K, top_k=3):
    """Reranking.py:339-358 without the printing: returns ``(ranks (K, Q), vecs_aug (N, 2D), qvecs_aug (Q, 2D))``."""
    indices = _db_sim_order(qvecs.T, vecs.T)
    q_aug = np.concatenate([qvecs.T, np.mean(vecs.T[indices[:, :top_k], :], axis=1)], axis=1)
    indices = _db_sim_order(vecs.T, vecs.T)
    v_aug = np.concatenate([vecs.T, np.mean(vecs.T[indices[:, 1:top_k + 1], :], axis=1)], axis=1)
    idx, _ = matching_L2(K, v_aug, q_aug)
    return idx.T, v_aug, q_aug


def database_augmentation(qvecs, vecs, K, top_k=3):
    """Reranking.py:404-431 without the printing: returns ``(ranks (K, Q), vecs_aug (N, D), qvecs_aug (Q, D))``."""
    weights = np.logspace(0, -2., top_k + 1)
    indices = _db_sim_order(qvecs.T, vecs.T)
    top_k_ref = vecs.T[indices[:, :top_k], :]
    q_aug = np.tensordot(weights, np.concatenate([np.expand_dims(qvecs.T, 1), top_k_ref], axis=1), axes=(0, 1))
    indices = _db_sim_order(vecs.T, vecs.T)
    v_aug = np.tensordot(weights, vecs.T[indices[:, :top_k + 1], :], axes=(0, 1))
    idx, _ = matching_L2(K, v_aug, q_aug)
    return idx.T, v_aug, q_aug


def cg_plain(a, b, tol=1e-6, maxiter=20):
    """The solver behind ``linalg.cg(trunc_lap, trunc_init, tol=1e-6, maxiter=20)`` (diffusion.py:18; scipy is a
    third-party dependency, ``scipy==1.9.0`` in requirements.txt:19): un-preconditioned conjugate gradients
    from x0 = 0, fp64, stop before a step when ``||r|| < tol * ||b||``, else return the iterate after
    ``maxiter`` steps.  tests/test_oracle_golden.py pins it against the scipy installed here."""
    b = np.asarray(b, dtype=np.float64)
    x = np.zeros_like(b)
    r = b.copy()
    atol = tol * float(np.linalg.norm(b))
    p = None
    rho_prev = 1.0
    for it in range(maxiter):
        if np.linalg.norm(r) < atol:
            break
        rho = float(r @ r)
        p = r.copy() if it == 0 else r + (rho / rho_prev) * p
        q = a @ p
        alpha = rho / float(p @ q)
        x += alpha * p
        r -= alpha * q
        rho_prev = rho
    return x


def offline_scores(lap, trunc_ids, tol=1e-6, maxiter=20):
    """``get_offline_result`` for every row (diffusion.py:15-19, 74-76): ``lap[ids][:, ids]`` then CG on e_0."""
    lap = lap.tocsr()
    trunc_ids = np.asarray(trunc_ids)
    out = np.empty(trunc_ids.shape, dtype=np.float64)
    e0 = np.zeros(trunc_ids.shape[1])
    e0[0] = 1
    for i in range(trunc_ids.shape[0]):
        ids = trunc_ids[i]
        out[i] = cg_plain(lap[ids][:, ids], e0, tol, maxiter)
    return out


# --------------------------------------------------------------------------------------
# 8e -- merge of per-shard top-k lists (no reference counterpart: the reference is single-process)
# --------------------------------------------------------------------------------------
def merge_parts(ids_parts, sims_parts, k):
    """Semantics the CUDA merge kernel must have: ``[G, nq, k] -> [nq, k]``, descending score, ties by
    ascending id, entries with id < 0 (a shard with fewer than k rows) ignored."""
    g, nq, kk = ids_parts.shape
    ids = np.empty((nq, k), dtype=np.int64)
    sims = np.empty((nq, k), dtype=np.float32)
    for j in range(nq):
        i = ids_parts[:, j, :].reshape(-1)
        s = sims_parts[:, j, :].reshape(-1)
        keep = i >= 0
        i, s = i[keep], s[keep]
        order = np.lexsort((i, -s.astype(np.float64)))[:k]
        ids[j, :len(order)] = i[order]
        sims[j, :len(order)] = s[order]
        ids[j, len(order):] = -1
        sims[j, len(order):] = -np.inf
    return ids, sims


# --------------------------------------------------------------------------------------
# comparator used by every parity test
# --------------------------------------------------------------------------------------
def compare_topk(ids, ref_ids, score_of, rtol=1e-6, atol=1e-7):
    """Tie-tolerant comparison of two top-K id lists for ONE query.

    ``score_of(id_array) -> float64 scores`` adjudicates.  The lists agree when, position by
    position, the ids are equal or their adjudicated scores differ by no more than
    ``atol + rtol*|score|`` (fp32 near-tie: OpenBLAS' summation order is not reproducible on a
    GPU, SURVEY.md section 7 "hard parts"), and the same holds at the K / K+1 boundary for set
    differences.  Returns ``(ok, message)``.
    """
    ids = np.asarray(ids)
    ref_ids = np.asarray(ref_ids)
    if ids.shape != ref_ids.shape:
        return False, f"shape {ids.shape} vs {ref_ids.shape}"
    if len(np.unique(ids)) != len(ids):
        return False, "duplicate ids in result"
    s_got = score_of(ids)
    s_ref = score_of(ref_ids)
    tol = atol + rtol * np.abs(s_ref)
    bad = (ids != ref_ids) & (np.abs(s_got - s_ref) > tol)
    if bad.any():
        p = int(np.nonzero(bad)[0][0])
        return False, (f"pos {p}: id {ids[p]} (score {s_got[p]:.9g}) vs ref id {ref_ids[p]} "
                       f"(score {s_ref[p]:.9g})")
    return True, ""
