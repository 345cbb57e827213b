"""Drop-in for the reference's faiss flat wrapper, ``src/utils/knn.py:8-40``.

``KNN(database, 'cosine').search(queries, k) -> (sims f32 [nq,k], ids int64 [nq,k])`` with the
database resident on the GPU.  ``'euclidean'`` (faiss ``IndexFlatL2``: ascending squared L2) is
served by the same inner-product kernels on augmented vectors ``[v, -|v|^2/2]`` / ``[q, 1]``.
"""
from __future__ import annotations

import numpy as np

from .index import ExactIndex


class BaseKNN(object):
    def __init__(self, database, method):
        # same contract as knn.py:10-15: the object owns an fp32, C-contiguous host copy and exposes N, D, database
        self.database = np.ascontiguousarray(database, dtype=np.float32)
        self.N, self.D = int(self.database.shape[0]), int(self.database.shape[-1])
        self.method = method

    def add(self, batch_size=10000):
        """knn.py:17-23 adds to the faiss index in 10k-row batches; here the whole matrix goes
        to the device in one build (staged in 8k-row tiles inside xs_index_create)."""
        if self.method == 'cosine':
            self.index = ExactIndex(self.database, renormalise=False, device=self.device)
        else:
            aug = np.empty((self.N, self.D + 1), dtype=np.float32)
            aug[:, :self.D] = self.database
            aug[:, self.D] = -0.5 * np.einsum('ij,ij->i', self.database, self.database, dtype=np.float64)
            self.index = ExactIndex(aug, renormalise=False, device=self.device)

    def search(self, queries, k):
        queries = np.ascontiguousarray(queries, dtype=np.float32)       # the cast knn.py:25-29 applies before faiss
        if self.method == 'cosine':
            ids, sims = self.index.search(queries, k)
            return sims, ids
        aug = np.ones((len(queries), self.D + 1), dtype=np.float32)
        aug[:, :self.D] = queries
        ids, s = self.index.search(aug, k)
        qn = np.einsum('ij,ij->i', queries, queries, dtype=np.float64)[:, None]
        return (qn - 2.0 * s.astype(np.float64)).astype(np.float32), ids

    def self_search(self, k):
        """``self.search(self.database, k)`` without shipping the database back through the
        host: the N x N kNN graph for diffusion (src/utils/diffusion.py:67)."""
        if self.method != 'cosine':
            return self.search(self.database, k)
        return self.index.self_knn(k)


class KNN(BaseKNN):
    def __init__(self, database, method, device=0):
        super().__init__(database, method)
        if method not in ('cosine', 'euclidean'):
            raise KeyError(method)          # the reference's dict lookup fails the same way, knn.py:36-37
        self.device = device
        self.add()
