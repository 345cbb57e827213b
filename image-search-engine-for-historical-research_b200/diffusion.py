"""Graph construction for diffusion re-ranking on top of the GPU self-kNN.

The reference builds the N x N kNN graph with faiss (``self.knn.search(self.features, n_trunc)``,
src/utils/diffusion.py:67) and then decides mutual neighbourhood row by row in Python
(``get_affinity``, :101-116).  Here the kNN lists come from the tcgen05 GEMM + fused top-K self search (a row's
own id first); the mutual test, the affinity values, the degrees and the normalised Laplacian are CUDA kernels
(xs_diffusion_laplacian) that round exactly where the reference's float32 scipy matrices do, and the per-row
truncated conjugate-gradient solves (:15-19, 74-76) run one CTA per database row (xs_diffusion_cg).
``Diffusion.get_offline_results`` strings the three together ON the device (xs_diffusion_offline): the kNN lists go
from the search kernels to the graph kernels to the solver without visiting the host.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat


def mutual_mask(ids, device: int = 0) -> np.ndarray:
    """``mask[i, j]`` is True iff ``j >= 1`` and ``i`` appears among the neighbours of ``ids[i, j]``
    -- exactly ``np.isin(ids[ids[i]], i).any(axis=1)`` with slot 0 cleared (diffusion.py:107-108)."""
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    n, kd = ids.shape
    out = np.empty((n, kd), dtype=np.uint8)
    nat.check(nat.load().xs_mutual_knn(int(device), ids.ctypes.data, n, kd, out.ctypes.data), "xs_mutual_knn")
    return out.view(np.bool_)


def _device_graph(sims, ids, alpha: float, gamma: float, device: int):
    """xs_diffusion_laplacian: mutual test, affinity, degrees and Laplacian on the device; returns the ELL arrays
    ``(cols int32 [n,kd], vals f32 [n,kd], cnt int32 [n], affinity f32 [n,kd])``."""
    sims = np.ascontiguousarray(sims, dtype=np.float32)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    n, kd = ids.shape
    cols = np.empty((n, kd), dtype=np.int32)
    vals = np.empty((n, kd), dtype=np.float32)
    cnt = np.empty((n,), dtype=np.int32)
    aff = np.empty((n, kd), dtype=np.float32)
    nat.check(nat.load().xs_diffusion_laplacian(int(device), sims.ctypes.data, ids.ctypes.data, n, kd, float(alpha), float(gamma),
                                                cols.ctypes.data, vals.ctypes.data, cnt.ctypes.data, aff.ctypes.data),
              "xs_diffusion_laplacian")
    return cols, vals, cnt, aff


def get_affinity(sims, ids, gamma=3, device: int = 0):
    """Drop-in for ``Diffusion.get_affinity(sims, ids, gamma=3)`` (diffusion.py:101-116): mutual-kNN affinity
    ``csc_matrix`` with values ``max(sim, 0) ** gamma``, computed on the device.  Like the reference it clamps negative
    similarities in the caller's ``sims`` array in place (:103)."""
    import scipy.sparse as sparse
    num = sims.shape[0]
    sims[sims < 0] = 0
    _, _, _, aff = _device_graph(sims, ids, 0.99, gamma, device)
    mask = aff != 0
    rows = np.repeat(np.arange(num), mask.sum(axis=1))
    return sparse.csc_matrix((aff[mask], (rows, np.asarray(ids)[mask])), shape=(num, num), dtype=np.float32)


def get_laplacian(sims, ids, alpha=0.99, device: int = 0):
    """``Diffusion.get_laplacian`` (diffusion.py:87-98): ``I - alpha * D^-1/2 A D^-1/2`` as a float32 ``csr_matrix``.
    Every entry is computed on the device in the reference's order of float32 roundings (csrc/graph.cu); the host only
    wraps the device's row lists into a scipy matrix."""
    import scipy.sparse as sparse
    num = sims.shape[0]
    sims[sims < 0] = 0                                       # the reference's get_affinity clamps in place (:103)
    cols, vals, cnt, _ = _device_graph(sims, ids, alpha, 3, device)
    keep = np.arange(cols.shape[1])[None, :] < cnt[:, None]
    indptr = np.zeros(num + 1, dtype=np.int64)
    np.cumsum(cnt, out=indptr[1:])
    lap = sparse.csr_matrix((vals[keep], cols[keep], indptr), shape=(num, num), dtype=np.float32)
    lap.sort_indices()
    return lap


def knn_graph(features, n_trunc: int, kd: int = 50, device: int = 0):
    """The offline front half of ``Diffusion.get_offline_results`` (diffusion.py:52-71, exact branch):
    ``(sims, ids)`` of the N x N self search truncated at ``n_trunc`` and the Laplacian of the first
    ``kd`` neighbours."""
    from .knn import KNN
    knn = KNN(np.asarray(features), "cosine", device=device)
    sims, ids = knn.self_search(n_trunc)
    lap = get_laplacian(sims[:, :kd].copy(), ids[:, :kd], device=device)
    return sims, ids, lap


def offline_cg(lap, trunc_ids, tol: float = 1e-6, maxiter: int = 20, device: int = 0) -> np.ndarray:
    """All rows of ``get_offline_result`` (diffusion.py:15-19) at once: for every row ``i`` solve
    ``lap[ids][:, ids] x = e_0`` with ``ids = trunc_ids[i]`` by at most ``maxiter`` CG steps
    (``linalg.cg(trunc_lap, trunc_init, tol=1e-6, maxiter=20)``).  Returns float32 ``(rows, n_trunc)``."""
    import scipy.sparse as sparse
    lap = sparse.csr_matrix(lap)
    lap.sum_duplicates()
    if lap.shape[0] != lap.shape[1]:
        raise ValueError("the Laplacian must be square")
    trunc_ids = np.ascontiguousarray(trunc_ids, dtype=np.int64)
    if trunc_ids.ndim != 2:
        raise ValueError("trunc_ids must be (rows, n_trunc)")
    rows, n_trunc = trunc_ids.shape
    indptr = np.ascontiguousarray(lap.indptr, dtype=np.int64)
    indices = np.ascontiguousarray(lap.indices, dtype=np.int32)
    values = np.ascontiguousarray(lap.data, dtype=np.float32)
    out = np.empty((rows, n_trunc), dtype=np.float32)
    nat.check(nat.load().xs_diffusion_cg(int(device), indptr.ctypes.data, indices.ctypes.data, values.ctypes.data,
                                         lap.shape[0], trunc_ids.ctypes.data, rows, n_trunc, int(maxiter), float(tol),
                                         out.ctypes.data), "xs_diffusion_cg")
    return out


def offline_device(index, n_trunc: int, kd: int = 50, alpha: float = 0.99, gamma: float = 3, tol: float = 1e-6, maxiter: int = 20,
                   return_sims: bool = False):
    """Steps 1-2 of ``get_offline_results`` (diffusion.py:52-76) in one device-resident pass over an ``ExactIndex``:
    N x N self-kNN truncated at ``n_trunc``, Laplacian of the first ``kd`` neighbours, one truncated CG per row.
    Returns ``(ids int64 (N, n_trunc), sims f32 or None, scores f32 (N, n_trunc))``."""
    n = index.N
    ids = np.empty((n, int(n_trunc)), dtype=np.int64)
    sims = np.empty((n, int(n_trunc)), dtype=np.float32) if return_sims else None
    scores = np.empty((n, int(n_trunc)), dtype=np.float32)
    nat.check(index._lib.xs_diffusion_offline(index._h, int(n_trunc), int(kd), float(alpha), float(gamma), int(maxiter), float(tol),
                                              ids.ctypes.data, sims.ctypes.data if return_sims else None, scores.ctypes.data),
              "xs_diffusion_offline")
    return ids, sims, scores


class Diffusion(object):
    """Mirror of ``src/utils/diffusion.py:42-116`` (exact-kNN branch; the ANN branch for N >= 110 000 exists in
    the reference only because the exhaustive search was too slow -- here the exhaustive one is used at
    every size).  ``features`` is ``(N, D)``; ``cache_dir`` may be None (no joblib cache)."""

    def __init__(self, features, cache_dir=None, device: int = 0):
        from .knn import KNN
        self.features = np.asarray(features)
        self.N = len(self.features)
        self.cache_dir = cache_dir
        self.device = device
        self.knn = KNN(self.features, method="cosine", device=device)

    def get_offline_results(self, n_trunc, kd=50):
        """diffusion.py:52-85: self-kNN truncated at ``n_trunc``, Laplacian of the first ``kd`` neighbours,
        one truncated CG per row, merged into an ``(N, N)`` float32 ``csr_matrix``."""
        import os
        import scipy.sparse as sparse
        path = os.path.join(self.cache_dir, "offline.jbl") if self.cache_dir else None
        if path and os.path.exists(path):
            import joblib
            return joblib.load(path)
        # self-kNN -> graph -> CG without leaving the device (xs_diffusion_offline); the host sees the final arrays only
        ids, _, all_scores = offline_device(self.knn.index, n_trunc, kd)
        rows = np.repeat(np.arange(self.N), n_trunc)
        offline = sparse.csr_matrix((all_scores.reshape(-1), (rows, ids.reshape(-1))), shape=(self.N, self.N),
                                    dtype=np.float32)
        if path:
            import joblib
            joblib.dump(offline, path)
        return offline

    def get_laplacian(self, sims, ids, alpha=0.99):
        return get_laplacian(sims, ids, alpha=alpha, device=self.device)

    def get_affinity(self, sims, ids, gamma=3):
        return get_affinity(sims, ids, gamma=gamma, device=self.device)


def search_offline(offline, sims, idx, n_trunc: int):
    """The query side of the diffusion re-ranking (Reranking.py:243-256): per query, the ``sims ** 3``-weighted
    sum of the offline rows of its ``k_query`` nearest database items, then the ``n_trunc`` best columns.
    ``sims``/``idx`` are what ``diffusion.knn.search(qvecs.T, k_query)`` returned.  Returns
    ``(truncation_scores f32 (Q, n_trunc), truncation_ranks int64 (Q, n_trunc))``; the caller transposes the
    ranks like the reference (:258).  Host-side: 3 sparse rows per query."""
    sims = np.asarray(sims) ** 3
    nq = idx.shape[0]
    out_s = np.empty((nq, n_trunc), dtype=np.float32)
    out_r = np.empty((nq, n_trunc), dtype=np.int64)
    for i in range(nq):
        scores = np.asarray(sims[i] @ offline[idx[i]]).reshape(-1)
        order = np.lexsort((np.arange(scores.size), -scores))[:n_trunc]
        out_s[i] = scores[order]
        out_r[i] = order
    return out_s, out_r
