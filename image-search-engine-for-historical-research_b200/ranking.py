"""The inline scoring + ranking sites of the reference, as one call.

``scores = np.dot(vecs.T, qvecs); ranks = np.argsort(-scores, axis=0)`` appears at
src/main_retrieve.py:175-176, src/main_train.py:703-704,715-716 and
src/utils/Reranking.py:206-207,299-300, always on ``vecs (D,N)`` / ``qvecs (D,Q)`` with unit-norm
columns and no re-normalisation.
"""
from __future__ import annotations

import numpy as np

from .nnsearch import leased_index


def rank_ip(vecs, qvecs, K=None, return_scores=False, index=None):
    """``ranks int64 (K or N, Q)``, one column per query, best first (evaluate.py:52-55).

    ``K=None`` reproduces the full ``argsort``; an integer K returns only the first K rows,
    which is all that ``compute_map`` on truncated ranks or the web UI (online.py:152) reads.
    With ``return_scores`` also returns the matching fp32 scores, same shape.
    """
    if index is None:
        with leased_index(np.asarray(vecs).T, renormalise=False) as ix:
            return rank_ip(vecs, qvecs, K=K, return_scores=return_scores, index=ix)
    ix = index
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


def rank_ip_torch(vecs, qvecs, K, index=None):
    """The same ranking for torch CUDA tensors without leaving the device -- the shape of the
    reference's other dense score+sort sites, e.g. hard-negative mining
    ``scores = torch.mm(poolvecs.t(), qvecs); scores, ranks = torch.sort(scores, dim=0, descending=True)``
    (src/datasets/traindataset.py:221-222, 468-469), which only reads the leading ranks.

    ``vecs`` is ``(D, N)``, ``qvecs`` ``(D, Q)`` (float tensors on one CUDA device); returns
    ``(scores f32 (K, Q), ranks int64 (K, Q))`` tensors on that device, best first.  Pass ``index`` (an
    ``ExactIndex``) to reuse a database already resident in the matcher's layout.
    """
    import torch
    from .index import ExactIndex
    if not (vecs.is_cuda and qvecs.is_cuda):
        raise ValueError("rank_ip_torch takes CUDA tensors; use rank_ip for numpy arrays")
    dev = vecs.device.index if vecs.device.index is not None else torch.cuda.current_device()
    own = index is None
    if own:
        rows = vecs.t().contiguous().float()
        torch.cuda.current_stream(dev).synchronize()
        index = ExactIndex.from_device(rows.data_ptr(), rows.shape[0], rows.shape[1], dev)
    q = qvecs.t().contiguous().float()
    nq = int(q.shape[0])
    ids = torch.empty((nq, int(K)), dtype=torch.int64, device=q.device)
    sims = torch.empty((nq, int(K)), dtype=torch.float32, device=q.device)
    index.search_device(q.data_ptr(), nq, int(K), ids.data_ptr(), sims.data_ptr(),
                        stream=torch.cuda.current_stream(dev).cuda_stream)
    if own:
        torch.cuda.current_stream(dev).synchronize()
        index.close()
    return sims.t().contiguous(), ids.t().contiguous()
