"""B200-native exhaustive (exact top-K) matching path of the historical-image search engine.

Public surface = the reference's three call shapes for this path:

* ``matching_L2(K, train, test) -> (idx, time_per_query)``      src/utils/nnsearch.py:687
* ``rank_ip(vecs, qvecs, K=None) -> ranks``                     src/main_retrieve.py:175-176
* ``KNN(database, 'cosine').search(queries, k) -> (sims, ids)`` src/utils/knn.py:8-40
* ``qge1(ranks, qvec, vecs, K) -> ranks_aqe`` (AQE second pass)   src/utils/Reranking.py:287-306

backed by hand-written sm_100a CUDA behind the C ABI in ``include/xs_b200.h``.  No CPU fallback.
"""
from .index import ExactIndex
from .knn import KNN, BaseKNN
from .nnsearch import matching, matching_L2, cached_index, leased_index, clear_index_cache
from .ranking import rank_ip, rank_ip_torch
from .reranking import (feature_enhancement, qge1, average_query_expansion, database_augmentation,
                        initial_rank)
from . import diffusion, store

__all__ = ["ExactIndex", "KNN", "BaseKNN", "matching", "matching_L2", "rank_ip", "rank_ip_torch",
           "cached_index", "leased_index", "clear_index_cache", "feature_enhancement", "qge1", "average_query_expansion",
           "database_augmentation", "initial_rank", "diffusion", "store"]
