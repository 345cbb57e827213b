"""AQE feature enhancement + re-score: the exhaustive-search half of the reference's QGE.

``qge1`` (src/utils/Reranking.py:287-306, called per web request at src/online.py:148) and the
``feature_enhancement`` helper of ``QGE`` (:195-208) build an expanded query from the top-k database
vectors and then repeat the full ``np.dot`` + ``argsort`` -- a second exhaustive scan.  Here the
expanded query is built on the device from the index's own fp32 rows and searched in place, so the
second scan costs one more pass of the same kernels and nothing crosses PCIe but ``k`` ids.

The diffusion / random-walk part of QGE (:212-265) is outside the exact path.
"""
from __future__ import annotations

import numpy as np

from .nnsearch import cached_index


def feature_enhancement(it_times, k, ranks, qvecs, vecs, w, K=None, index=None):
    """Same arguments as the nested helper at Reranking.py:195 (``ranks`` is ``(>=k, Q)``, ``vecs``
    ``(D, N)``); returns ``(qvecs_qe (D, Q), ranks_aqe)``.

    ``ranks_aqe`` is the full ``(N, Q)`` ranking when ``K`` is None (as the reference), else only its
    first K rows.  As in the reference, ``ranks`` is not fed back between iterations (:196-207 never
    reassign it), so every iteration yields the same result and one pass is computed.
    """
    ix = index if index is not None else cached_index(np.asarray(vecs).T, renormalise=False)
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
