"""Graph construction for diffusion re-ranking on top of the GPU self-kNN.

The reference builds the N x N kNN graph with faiss (``self.knn.search(self.features, n_trunc)``,
src/utils/diffusion.py:67) and then decides mutual neighbourhood row by row in Python
(``get_affinity``, :101-116).  Here the kNN lists come from ``KNN.self_search`` (tcgen05 GEMM + fused
top-K, a row's own id first) and the mutual test runs as one CUDA kernel (xs_mutual_knn); the sparse
assembly and the normalised Laplacian keep the reference's scipy formulation (:87-98).  The per-row
conjugate-gradient solves (:15-19, 74-76) are outside the exact-matching path and are not rebuilt.
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


def get_affinity(sims, ids, gamma=3, device: int = 0):
    """Drop-in for ``Diffusion.get_affinity(sims, ids, gamma=3)`` (diffusion.py:101-116): mutual-kNN
    affinity ``csc_matrix`` with values ``max(sim, 0) ** gamma``.  Like the reference it clamps
    negative similarities in the caller's ``sims`` array in place (:103)."""
    import scipy.sparse as sparse
    num = sims.shape[0]
    sims[sims < 0] = 0
    powed = sims ** gamma
    mask = mutual_mask(ids, device)
    rows = np.repeat(np.arange(num), mask.sum(axis=1))
    return sparse.csc_matrix((powed[mask], (rows, np.asarray(ids)[mask])), shape=(num, num), dtype=np.float32)


def get_laplacian(sims, ids, alpha=0.99, device: int = 0):
    """``Diffusion.get_laplacian`` (diffusion.py:87-98): ``I - alpha * D^-1/2 A D^-1/2``."""
    import scipy.sparse as sparse
    affinity = get_affinity(sims, ids, device=device)
    num = affinity.shape[0]
    degrees = affinity @ np.ones(num) + 1e-12
    mat = sparse.dia_matrix((degrees ** (-0.5), [0]), shape=(num, num), dtype=np.float32)
    stochastic = mat @ affinity @ mat
    sparse_eye = sparse.dia_matrix((np.ones(num), [0]), shape=(num, num), dtype=np.float32)
    return sparse_eye - alpha * stochastic


def knn_graph(features, n_trunc: int, kd: int = 50, device: int = 0):
    """The offline front half of ``Diffusion.get_offline_results`` (diffusion.py:52-71, exact branch):
    ``(sims, ids)`` of the N x N self search truncated at ``n_trunc`` and the Laplacian of the first
    ``kd`` neighbours."""
    from .knn import KNN
    knn = KNN(np.asarray(features), "cosine", device=device)
    sims, ids = knn.self_search(n_trunc)
    lap = get_laplacian(sims[:, :kd].copy(), ids[:, :kd], device=device)
    return sims, ids, lap
