"""Drop-in for the exhaustive branch of the reference's ``src/utils/nnsearch.py``.

``matching_L2`` keeps the reference signature and return contract (nnsearch.py:687-706):
``(idx int64 [Q, K], time_per_query seconds)``, inputs any float dtype / strides, not mutated,
rows normalised inside.  The reference function is stateless and ``src/online.py:133`` calls it
per request with the same ``vecs.T``; uploading 8-16 GB per call would dominate, so the device
index is cached per database array (identity of the owning buffer + a content fingerprint).
"""
from __future__ import annotations

import threading
import time
import weakref
from collections import OrderedDict

import numpy as np

from .index import ExactIndex

_CACHE_SLOTS = 2
_cache: "OrderedDict[tuple, _Entry]" = OrderedDict()
_cache_lock = threading.Lock()
default_device = 0


class _Entry:
    """One cached device index.  ``users`` counts the calls currently searching it: an entry that is evicted (or
    found stale) while in use is closed by its last user, never under a running search."""
    __slots__ = ("ref", "fp", "index", "users", "retired")

    def __init__(self, ref, fp, index):
        self.ref, self.fp, self.index, self.users, self.retired = ref, fp, index, 0, False


def _root(a: np.ndarray):
    while isinstance(getattr(a, "base", None), np.ndarray):
        a = a.base
    return a


def _fingerprint(a: np.ndarray) -> bytes:
    """256 probes spread over the matrix: catches most in-place edits without reading 8 GB per request (the reference
    function is stateless and would simply see the new values).  An edit that misses every probe is served from the
    old device copy -- call ``clear_index_cache()`` after mutating a database array in place."""
    n, d = a.shape
    rows = np.linspace(0, n - 1, num=min(n, 256)).astype(np.int64)
    cols = (rows * 131 + 7) % d
    return np.asarray(a[rows, cols], dtype=np.float64).tobytes()


def _retire(entry: _Entry):
    """Caller holds the cache lock.  Close now if nobody is searching the index, else leave it to the last user."""
    entry.retired = True
    if entry.users == 0:
        entry.index.close()


def clear_index_cache():
    """Drop every cached device index (call after mutating a database array in place)."""
    with _cache_lock:
        for entry in list(_cache.values()):
            _retire(entry)
        _cache.clear()


def _lease(database, renormalise: bool, device: int | None = None) -> _Entry:
    a = np.asarray(database)
    dev = default_device if device is None else device
    key = (a.__array_interface__["data"][0], a.shape, a.strides, a.dtype.str, bool(renormalise), dev)
    fp = _fingerprint(a)
    with _cache_lock:
        # entries whose array has been garbage-collected only pin HBM: drop them first
        for k in [k for k, e in _cache.items() if e.ref() is None]:
            _retire(_cache.pop(k))
        hit = _cache.get(key)
        if hit is not None:
            if hit.fp == fp:
                _cache.move_to_end(key)
                hit.users += 1
                return hit
            _retire(_cache.pop(key))
        # make room BEFORE building: a 1M-row index is 16.5 GB, three of them at once is how a 10M-row job runs out of memory
        while len(_cache) >= _CACHE_SLOTS:
            _, old = _cache.popitem(last=False)
            _retire(old)
        entry = _Entry(weakref.ref(_root(a)), fp, ExactIndex(a, renormalise=renormalise, device=dev))
        entry.users = 1
        _cache[key] = entry
        return entry


def _release(entry: _Entry):
    with _cache_lock:
        entry.users -= 1
        if entry.retired and entry.users == 0:
            entry.index.close()


class leased_index:
    """``with leased_index(db, renormalise) as ix:`` -- the cached device index of ``db``, protected against eviction
    by another thread for the duration of the block."""

    def __init__(self, database, renormalise: bool, device: int | None = None):
        self._args = (database, renormalise, device)
        self._entry = None

    def __enter__(self) -> ExactIndex:
        self._entry = _lease(*self._args)
        return self._entry.index

    def __exit__(self, *exc):
        _release(self._entry)
        self._entry = None


def cached_index(database, renormalise: bool, device: int | None = None) -> ExactIndex:
    """The cached index without a lease (single-threaded callers).  Prefer ``leased_index`` where requests run on
    threads (the reference's Flask server, src/online.py:163)."""
    entry = _lease(database, renormalise, device)
    _release(entry)
    return entry.index


def matching_L2(K, embedded_features_train, embedded_features_test):
    """Exact K nearest neighbours by normalised L2 distance == descending cosine.

    Same call as ``src/utils/nnsearch.py:687``: ``train`` is ``(N, D)`` (typically ``vecs.T``),
    ``test`` is ``(Q, D)``; returns ``(idx, time_per_query)`` with ``idx`` int64 ``(Q, K)`` and
    the wall-clock timer around the whole call as at :688,704-705.  Raises ``ValueError`` for
    ``K > N`` (the reference fails there too, with a numpy broadcast error at :703).

    The device copy of ``train`` is cached between calls (see ``_fingerprint`` for what that means for arrays that are
    edited in place).
    """
    t1 = time.time()
    num_train, _ = np.shape(embedded_features_train)
    num_test, _ = np.shape(embedded_features_test)
    if K > num_train:
        raise ValueError(f"K={K} exceeds the {num_train} database rows")
    with leased_index(embedded_features_train, renormalise=True) as index:
        if K == num_train or K > 4096:
            idx = np.ascontiguousarray(index.rank_all(embedded_features_test, renormalise=True)[:K].T)
        else:
            idx, _ = index.search(embedded_features_test, K, renormalise=True)
    t2 = time.time()
    return idx, (t2 - t1) / num_test


def matching_L2_once(K, embedded_features_train, embedded_features_test, device: int | None = None):
    """``matching_L2`` on a database that is used once (the augmented matrices of the re-ranking helpers): the device
    index lives for this call only and never enters the cache."""
    t1 = time.time()
    num_train, _ = np.shape(embedded_features_train)
    num_test, _ = np.shape(embedded_features_test)
    if K > num_train:
        raise ValueError(f"K={K} exceeds the {num_train} database rows")
    with ExactIndex(embedded_features_train, renormalise=True, device=default_device if device is None else device) as index:
        if K == num_train or K > 4096:
            idx = np.ascontiguousarray(index.rank_all(embedded_features_test, renormalise=True)[:K].T)
        else:
            idx, _ = index.search(embedded_features_test, K, renormalise=True)
    return idx, (time.time() - t1) / num_test


EXHAUSTIVE_NAMES = ("L2", "exhaustive")


def matching(method, K, embedded_features_train, embedded_features_test, **kwargs):
    """The ``--matching_method`` string dispatch of src/online.py:132-143 / src/offline.py:107-118
    for the one branch this framework implements.  ``'L2'`` is the reference's flag value;
    ``'exhaustive'`` is accepted as an alias.  The approximate methods (PQ, ANNOY, HNSW, PQ_HNSW)
    are outside the exact path."""
    if method in EXHAUSTIVE_NAMES:
        return matching_L2(K, embedded_features_train, embedded_features_test)
    raise NotImplementedError(f"matching_method {method!r}: only the exhaustive ('L2') path is built here")
