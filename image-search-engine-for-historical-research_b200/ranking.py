"""The inline scoring + ranking sites of the reference, as one call.

``scores = np.dot(vecs.T, qvecs); ranks = np.argsort(-scores, axis=0)`` appears at
src/main_retrieve.py:175-176, src/main_train.py:703-704,715-716 and
src/utils/Reranking.py:206-207,299-300, always on ``vecs (D,N)`` / ``qvecs (D,Q)`` with unit-norm
columns and no re-normalisation.
"""
from __future__ import annotations

import numpy as np

from .nnsearch import cached_index


def rank_ip(vecs, qvecs, K=None, return_scores=False, index=None):
    """``ranks int64 (K or N, Q)``, one column per query, best first (evaluate.py:52-55).

    ``K=None`` reproduces the full ``argsort``; an integer K returns only the first K rows,
    which is all that ``compute_map`` on truncated ranks or the web UI (online.py:152) reads.
    With ``return_scores`` also returns the matching fp32 scores, same shape.
    """
    ix = index if index is not None else cached_index(np.asarray(vecs).T, renormalise=False)
    q = np.asarray(qvecs).T
    if K is None or K >= ix.N or K > 4096:
        out = ix.rank_all(q, return_scores=return_scores)
        if return_scores:
            r, s = out
            return (r, s) if K is None else (r[:K], s[:K])
        return out if K is None else out[:K]
    ids, sims = ix.search(q, K)
    ranks = np.ascontiguousarray(ids.T)
    return (ranks, np.ascontiguousarray(sims.T)) if return_scores else ranks
