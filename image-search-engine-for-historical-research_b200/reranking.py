"""AQE feature enhancement + re-score: the exhaustive-search half of the reference's QGE.

``qge1`` (src/utils/Reranking.py:287-306, called per web request at src/online.py:148) and the
``feature_enhancement`` helper of ``QGE`` (:195-208) build an expanded query from the top-k database
vectors and then repeat the full ``np.dot`` + ``argsort`` -- a second exhaustive scan.  Here the
expanded query is built on the device from the index's own fp32 rows and searched in place, so the
second scan costs one more pass of the same kernels and nothing crosses PCIe but ``k`` ids.

The diffusion / random-walk part of QGE (:212-265) lives in ``diffusion.py``.  The two older re-ranking
schemes that the reference builds on an N x N similarity matrix plus a full argsort --
``average_query_expansion`` (:314-365) and ``database_augmentation`` (:375-440) -- only ever read the first
3-4 columns of that argsort, so here they are a top-k search / self-kNN on the same kernels; ``initial_rank``
is the k-reciprocal method's ``batch_torch_topk`` (:487-511).
"""
from __future__ import annotations

import numpy as np

from .index import ExactIndex
from .nnsearch import leased_index, matching_L2_once


def feature_enhancement(it_times, k, ranks, qvecs, vecs, w, K=None, index=None):
    """Same arguments as the nested helper at Reranking.py:195 (``ranks`` is ``(>=k, Q)``, ``vecs``
    ``(D, N)``); returns ``(qvecs_qe (D, Q), ranks_aqe)``.

    ``ranks_aqe`` is the full ``(N, Q)`` ranking when ``K`` is None (as the reference), else only its
    first K rows.  As in the reference, ``ranks`` is not fed back between iterations (:196-207 never
    reassign it), so every iteration yields the same result and one pass is computed.
    """
    if index is None:
        with leased_index(np.asarray(vecs).T, renormalise=False) as ix:
            return feature_enhancement(it_times, k, ranks, qvecs, vecs, w, K=K, index=ix)
    ix = index
    top = np.ascontiguousarray(np.asarray(ranks)[:k, :].T, dtype=np.int64)          # (Q, k), best first
    if K is None or K >= ix.N or K > 4096:
        _, _, qe = ix.aqe_search(top, 1, w=w, return_queries=True)
        r = ix.rank_all(qe)
        return np.ascontiguousarray(qe.T), (r if K is None else r[:K])
    ids, _, qe = ix.aqe_search(top, K, w=w, return_queries=True)
    return np.ascontiguousarray(qe.T), np.ascontiguousarray(ids.T)


def qge1(ranks, qvec, vecs, K, full=False):
    """Drop-in for ``qge1(ranks, qvec, vecs, K)`` (Reranking.py:287-306): k=3, w=4, one iteration.

    Returns ``ranks_aqe`` with one column per query.  The reference returns all N rows and its caller
    keeps ``[:K]`` (online.py:149-152); by default only those K rows are produced -- pass
    ``full=True`` for the complete ``(N, Q)`` ranking.
    """
    _, ranks_aqe = feature_enhancement(1, 3, ranks, qvec, vecs, 8. / 2, K=None if full else K)
    return ranks_aqe


def _postprocess(query_vecs, reference_vecs, same=False):
    """``postprocess`` of Reranking.py:326-332 (= :387-398): centre on the mean of all rows of both sets, then
    L2-normalise each set (a set with a zero row is left unnormalised, :320-324)."""
    if same:
        center = np.mean(reference_vecs, axis=0)           # mean of [v; v] is the mean of v
    else:
        center = np.mean(np.concatenate([query_vecs, reference_vecs], axis=0), axis=0)
    out = []
    for v in ((reference_vecs,) if same else (query_vecs, reference_vecs)):
        v = v - center
        norm = np.expand_dims(np.linalg.norm(v, axis=1), axis=1)
        out.append(v if np.any(norm == 0) else v / norm)
    return (out[0], out[0]) if same else (out[0], out[1])


def _nearest_rows(query_vecs, reference_vecs, k, device=0):
    """First ``k`` columns of ``np.argsort(calculate_sim_matrix(query_vecs, reference_vecs), axis=1)``
    (Reranking.py:334-345): the k most similar reference rows per query row after ``postprocess``.
    ``query_vecs is reference_vecs`` selects the N x N case (a row's own id comes first)."""
    same = query_vecs is reference_vecs
    qn, rn = _postprocess(query_vecs, reference_vecs, same)
    with ExactIndex(rn, device=device, renormalise=False) as ix:
        if same:
            return ix.self_knn(k)[1]
        return ix.search(qn, k)[0]


def average_query_expansion(qvecs, vecs, K, dataset=None, gnd=None, top_k=3, device=0):
    """``average_query_expansion(qvecs, vecs, K, dataset, gnd)`` (Reranking.py:314-365): queries and database
    rows are extended by the mean of their ``top_k`` nearest database rows (2D-dimensional vectors), then
    ``matching_L2``.  The reference prints the mAP of ``ranks_qe2`` and returns nothing; this returns
    ``ranks_qe2`` ``(K, Q)`` (``dataset`` / ``gnd`` are accepted and ignored -- evaluation is the caller's)."""
    q, v = np.asarray(qvecs).T, np.asarray(vecs).T
    ids = _nearest_rows(q, v, top_k, device)
    q_aug = np.concatenate([q, np.mean(v[ids, :], axis=1)], axis=1)
    ids = _nearest_rows(v, v, top_k + 1, device)
    v_aug = np.concatenate([v, np.mean(v[ids[:, 1:top_k + 1], :], axis=1)], axis=1)
    match_idx, _ = matching_L2_once(K, v_aug, q_aug, device)      # a one-off database: not worth a cache slot
    return match_idx.T


def database_augmentation(qvecs, vecs, K, dataset=None, gnd=None, top_k=3, device=0):
    """``database_augmentation(qvecs, vecs, K, dataset, gnd)`` (Reranking.py:375-440): every query / database
    row becomes the ``logspace(0, -2, top_k+1)``-weighted sum of itself and its nearest database rows, then
    ``matching_L2``.  Returns ``ranks_dba`` ``(K, Q)``."""
    q, v = np.asarray(qvecs).T, np.asarray(vecs).T
    weights = np.logspace(0, -2., top_k + 1)
    ids = _nearest_rows(q, v, top_k, device)
    q_aug = np.tensordot(weights, np.concatenate([np.expand_dims(q, 1), v[ids, :]], axis=1), axes=(0, 1))
    ids = _nearest_rows(v, v, top_k + 1, device)
    v_aug = np.tensordot(weights, v[ids, :], axes=(0, 1))
    match_idx, _ = matching_L2_once(K, v_aug, q_aug, device)
    return match_idx.T


def initial_rank(feat, k1, device=0):
    """``batch_torch_topk(feat, feat, k1)`` of the k-reciprocal re-ranking (Reranking.py:487-511, called on the
    concatenated query+gallery features): per row the ``k1`` nearest rows by ``2 - 2 * feat @ feat.T``.  The
    reference divides every row of the distance matrix by that row's maximum before ``topk`` -- a positive
    per-row scale that cannot change the row's order -- so this is the self-kNN id list.  ``feat`` is
    ``(M, D)`` L2-normalised (numpy or a torch tensor); returns int64 ``(M, k1)``."""
    if hasattr(feat, "detach"):
        feat = feat.detach().cpu().numpy()
    with ExactIndex(np.asarray(feat), device=device, renormalise=False) as ix:
        return ix.self_knn(k1)[1]
