"""Device-resident exact index: the object behind matching_L2 / KNN / rank_ip."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat


class ExactIndex:
    """bf16 + fp32 row-major copy of an (N, D) descriptor matrix in HBM, searched exactly.

    ``database`` may be any float array with the strides the reference produces (notably the
    F-order view ``vecs.T`` of a ``(D, N)`` array, src/online.py:133).  ``renormalise=True`` gives
    ``matching_L2`` semantics (rows divided by their norm, src/utils/nnsearch.py:693-697);
    ``False`` gives ``np.dot`` / ``IndexFlatIP`` semantics (rows used as they are).
    """

    def __init__(self, database, renormalise: bool = False, device: int = 0, id_offset: int = 0):
        self._lib = nat.load()
        self._h = C.c_void_p()
        a, code, sr, sc = nat.as_matrix(database, "database")
        self.N, self.D = int(a.shape[0]), int(a.shape[1])
        self.device = int(device)
        self.renormalised = bool(renormalise)
        nat.check(self._lib.xs_index_create(a.ctypes.data, code, self.N, self.D, sr, sc, self.device,
                                            int(bool(renormalise)), int(id_offset), C.byref(self._h)),
                  "xs_index_create")

    @classmethod
    def from_device(cls, data_ptr: int, n: int, d: int, device: int, renormalise: bool = False, id_offset: int = 0):
        """Build from an fp32 row-major ``[n, d]`` matrix that already lives on ``device`` (e.g. a
        torch tensor's ``data_ptr()``) -- used by bench.py to synthesise 1M+ rows on the GPU."""
        self = cls.__new__(cls)
        self._lib = nat.load()
        self._h = C.c_void_p()
        self.N, self.D, self.device, self.renormalised = int(n), int(d), int(device), bool(renormalise)
        nat.check(self._lib.xs_index_create_dev(C.c_void_p(data_ptr), self.N, self.D, self.device,
                                                int(bool(renormalise)), int(id_offset), C.byref(self._h)),
                  "xs_index_create_dev")
        return self

    def save(self, path: str):
        """Write the device arrays as they are (xs_index_save); ``ExactIndex.load`` brings them back with plain copies."""
        nat.check(self._lib.xs_index_save(self._h, str(path).encode()), "xs_index_save")

    @classmethod
    def load(cls, path: str, device: int = 0, id_offset: int = 0):
        """An index from its on-disk image (xs_index_load): pinned, multi-threaded, double-buffered upload; no kernels."""
        self = cls.__new__(cls)
        self._lib = nat.load()
        self._h = C.c_void_p()
        nat.check(self._lib.xs_index_load(str(path).encode(), int(device), int(id_offset), C.byref(self._h)), "xs_index_load")
        n, d = C.c_int64(), C.c_int()
        nat.check(self._lib.xs_index_info(self._h, C.byref(n), C.byref(d), None, None), "xs_index_info")
        self.N, self.D, self.device, self.renormalised = int(n.value), int(d.value), int(device), False
        return self

    def clone(self) -> "ExactIndex":
        """A second search lane over the same device-resident database (xs_index_clone): shares the database
        arrays, owns its workspaces -- searches on ``self`` and on the clone may overlap on two streams."""
        other = type(self).__new__(type(self))
        other._lib = self._lib
        other._h = C.c_void_p()
        other.N, other.D, other.device, other.renormalised = self.N, self.D, self.device, self.renormalised
        nat.check(self._lib.xs_index_clone(self._h, C.byref(other._h)), "xs_index_clone")
        return other

    # -- lifetime -----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.xs_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- tunables / stats ---------------------------------------------------------------------
    def set_param(self, name: str, value: float):
        nat.check(self._lib.xs_set_param(self._h, name.encode(), float(value)), "xs_set_param")

    @property
    def device_bytes(self) -> int:
        b = C.c_int64()
        nat.check(self._lib.xs_index_info(self._h, None, None, None, C.byref(b)), "xs_index_info")
        return int(b.value)

    def stats(self) -> dict:
        s = nat.XsStats()
        nat.check(self._lib.xs_index_stats(self._h, C.byref(s)), "xs_index_stats")
        return {f: getattr(s, f) for f, _ in s._fields_}

    # -- search -------------------------------------------------------------------------------
    def search(self, queries, k: int, renormalise: bool = False):
        """Exact top-k by inner product: ``(ids int64 [nq,k], sims f32 [nq,k])``, best first,
        exact ties by ascending id."""
        q, code, sr, sc = nat.as_matrix(queries, "queries")
        if q.shape[1] != self.D:
            raise ValueError(f"queries have {q.shape[1]} columns, the index has {self.D}")
        nq = int(q.shape[0])
        ids = np.empty((nq, int(k)), dtype=np.int64)
        sims = np.empty((nq, int(k)), dtype=np.float32)
        if nq == 0:
            return ids, sims
        nat.check(self._lib.xs_search(self._h, q.ctypes.data, code, nq, sr, sc, int(bool(renormalise)), int(k),
                                      ids.ctypes.data, sims.ctypes.data), "xs_search")
        return ids, sims

    def search_device(self, q_ptr: int, nq: int, k: int, out_idx_ptr: int, out_score_ptr: int,
                      status_ptr: int = 0, stream: int = 0, renormalise: bool = False):
        """Device-pointer variant (fp32 row-major queries on the index's device), enqueued on
        ``stream``; see xs_search_dev in include/xs_b200.h."""
        nat.check(self._lib.xs_search_dev(self._h, C.c_void_p(q_ptr), int(nq), int(bool(renormalise)), int(k),
                                          C.c_void_p(out_idx_ptr), C.c_void_p(out_score_ptr),
                                          C.c_void_p(status_ptr) if status_ptr else None,
                                          C.c_void_p(stream) if stream else None), "xs_search_dev")

    def search_device_push(self, q_ptr: int, nq: int, k: int, exchange_handle, slot: int, stream: int = 0, renormalise: bool = False):
        """Device search whose last kernel stores the results into every rank's peer-exchange mailbox
        (xs_search_dev_push); collect them with ``PeerExchange.merge``."""
        nat.check(self._lib.xs_search_dev_push(self._h, C.c_void_p(q_ptr), int(nq), int(bool(renormalise)), int(k),
                                               exchange_handle, int(slot), C.c_void_p(stream) if stream else None), "xs_search_dev_push")

    def search_device_exchange(self, q_ptr: int, nq: int, k: int, exchange_handle, slot: int, out_idx_ptr: int, out_score_ptr: int,
                               out_status_ptr: int, stream: int = 0, renormalise: bool = False):
        """Search + exchange in one call (xs_search_dev_exchange): local search, results into every rank's mailbox, and the
        merged ``[nq, k]`` answer (plus certificate words) into the given device buffers."""
        nat.check(self._lib.xs_search_dev_exchange(self._h, C.c_void_p(q_ptr), int(nq), int(bool(renormalise)), int(k), exchange_handle, int(slot),
                                                   C.c_void_p(out_idx_ptr), C.c_void_p(out_score_ptr), C.c_void_p(out_status_ptr),
                                                   C.c_void_p(stream) if stream else None), "xs_search_dev_exchange")

    def self_knn(self, k: int, begin: int = 0, end: int | None = None):
        """Top-k neighbours of database rows ``[begin, end)`` among all rows; a row's own id is
        first (src/utils/diffusion.py:67,108).  Returns ``(sims, ids)`` like ``KNN.search``."""
        end = self.N if end is None else int(end)
        nq = end - int(begin)
        ids = np.empty((nq, int(k)), dtype=np.int64)
        sims = np.empty((nq, int(k)), dtype=np.float32)
        nat.check(self._lib.xs_self_knn(self._h, int(begin), end, int(k), ids.ctypes.data, sims.ctypes.data), "xs_self_knn")
        return sims, ids

    def aqe_search(self, top_ids, k: int, w: float = 4.0, return_queries: bool = False):
        """Average-query-expansion re-score (src/utils/Reranking.py:195-208): ``top_ids`` is
        ``(nq, kq)`` -- each query's kq best ids from a previous search; returns ``(ids, sims)`` of the
        expanded queries (and the expanded queries ``(nq, D)`` fp32 with ``return_queries``)."""
        t = np.ascontiguousarray(top_ids, dtype=np.int64)
        nq, kq = t.shape
        ids = np.empty((nq, int(k)), dtype=np.int64)
        sims = np.empty((nq, int(k)), dtype=np.float32)
        qe = np.empty((nq, self.D), dtype=np.float32) if return_queries else None
        nat.check(self._lib.xs_aqe_search(self._h, t.ctypes.data, nq, kq, float(w), int(k), ids.ctypes.data, sims.ctypes.data,
                                          qe.ctypes.data if return_queries else None), "xs_aqe_search")
        return (ids, sims, qe) if return_queries else (ids, sims)

    def rank_all(self, queries, renormalise: bool = False, return_scores: bool = False):
        """Full ranking: ``ranks int64 (N, nq)``, one column per query, best first -- the array
        ``np.argsort(-scores, axis=0)`` yields at src/main_retrieve.py:176."""
        q, code, sr, sc = nat.as_matrix(queries, "queries")
        if q.shape[1] != self.D:
            raise ValueError(f"queries have {q.shape[1]} columns, the index has {self.D}")
        nq = int(q.shape[0])
        ranks = np.empty((self.N, nq), dtype=np.int64)
        sc_sorted = np.empty((self.N, nq), dtype=np.float32) if return_scores else None
        nat.check(self._lib.xs_rank_all(self._h, q.ctypes.data, code, nq, sr, sc, int(bool(renormalise)),
                                        ranks.ctypes.data, sc_sorted.ctypes.data if return_scores else None), "xs_rank_all")
        return (ranks, sc_sorted) if return_scores else ranks
