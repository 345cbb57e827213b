"""Seeded synthetic descriptor sets for the exhaustive matching path.

The reference ships no data (its descriptors come from a ResNet101-SOLAR/GeM
extractor, /root/reference/src/networks/imageretrievalnet.py:356-386, that emits one
unit-norm 2048-d fp32 column per image).  These generators produce arrays of exactly
that contract -- ``vecs`` is ``(D, N)`` C-contiguous fp32 with unit-norm columns,
``qvecs`` is ``(D, Q)`` -- in the four families SURVEY.md section 8(d) names:

* ``G`` isotropic Gaussian          (scores ~ N(0, 1/D))
* ``P`` non-negative |Gaussian|     (un-whitened GeM look-alike, all scores crowded)
* ``C`` clustered + ground truth    (centroids + noise; labels give ok/junk/easy/hard)
* ``T`` ties / duplicates           (repeated rows -> exactly equal scores)

Everything is generated in 50k-row chunks from ``numpy.random.default_rng(seed)`` so
that a 1M-row set never needs a float64 temporary of full size.  DB seed 0 and query
seed 1 are the conventions used by bench.py and the tests.
"""
from __future__ import annotations

import numpy as np

CHUNK = 50_000


def _unit_rows(x: np.ndarray) -> np.ndarray:
    n = np.sqrt(np.einsum("ij,ij->i", x, x, dtype=np.float64)).astype(np.float32)
    n[n == 0] = 1.0
    x /= n[:, None]
    return x


def rows(n: int, d: int, seed: int, family: str = "G") -> np.ndarray:
    """``(n, d)`` C-contiguous fp32, unit-norm rows (row-major twin of the reference layout)."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, d), dtype=np.float32)
    for lo in range(0, n, CHUNK):
        hi = min(n, lo + CHUNK)
        blk = rng.standard_normal((hi - lo, d), dtype=np.float32)
        if family == "P":
            np.abs(blk, out=blk)
        elif family != "G":
            raise ValueError(f"unknown family {family!r}")
        out[lo:hi] = _unit_rows(blk)
    return out


def reference_layout(rowmajor: np.ndarray) -> np.ndarray:
    """Row-major ``(N, D)`` -> the reference's ``vecs`` array: ``(D, N)`` C-contiguous.

    ``vecs.T`` (what the scripts hand to ``matching_L2``, online.py:133) is then an F-order
    ``(N, D)`` view, exactly as in the reference.
    """
    return np.ascontiguousarray(rowmajor.T)


def gaussian(n: int, q: int, d: int = 2048, db_seed: int = 0, q_seed: int = 1, family: str = "G"):
    """Returns ``(vecs (D,N), qvecs (D,Q))`` in the reference layout."""
    return reference_layout(rows(n, d, db_seed, family)), reference_layout(rows(q, d, q_seed, family))


def clustered(n: int, q: int, d: int = 2048, n_clusters: int = 200, noise: float = 0.7,
              db_seed: int = 0, q_seed: int = 1, spread: float | None = None):
    """Clustered set with synthetic ground truth in the rOxford/rParis ``gnd`` shape.

    Returns ``(vecs (D,N), qvecs (D,Q), gnd)`` where ``gnd[i]`` has ``easy``/``hard``/``junk``
    (new protocol, evaluate.py:123-147) *and* ``ok`` (old protocol, evaluate.py:118-120) id
    lists.  Members of the query's cluster are split by id modulo 4: 0,1 -> easy, 2 -> hard,
    3 -> junk; ``ok`` = easy + hard.  ``spread`` (optional) pulls all centroids towards one
    common direction (``unit(common + spread * g_i)``) so that clusters overlap and the
    ranking -- hence the mAP -- is not trivially perfect.
    """
    rng = np.random.default_rng(db_seed)
    cent = _unit_rows(rng.standard_normal((n_clusters, d), dtype=np.float32))
    if spread is not None:
        common = _unit_rows(rng.standard_normal((1, d), dtype=np.float32))
        cent = _unit_rows(common + np.float32(spread) * cent)
    labels = rng.integers(0, n_clusters, size=n)
    db = np.empty((n, d), dtype=np.float32)
    for lo in range(0, n, CHUNK):
        hi = min(n, lo + CHUNK)
        blk = rng.standard_normal((hi - lo, d), dtype=np.float32)
        blk *= np.float32(noise / np.sqrt(d))
        blk += cent[labels[lo:hi]]
        db[lo:hi] = _unit_rows(blk)
    rq = np.random.default_rng(q_seed)
    qlab = rq.integers(0, n_clusters, size=q)
    qn = rq.standard_normal((q, d), dtype=np.float32)
    qn *= np.float32(noise / np.sqrt(d))
    qn += cent[qlab]
    qn = _unit_rows(qn)
    gnd = []
    for i in range(q):
        members = np.nonzero(labels == qlab[i])[0]
        m4 = members % 4
        easy, hard, junk = members[m4 <= 1], members[m4 == 2], members[m4 == 3]
        gnd.append({"easy": easy, "hard": hard, "junk": junk,
                    "ok": np.concatenate([easy, hard])})
    return reference_layout(db), reference_layout(qn), gnd


def ties(n: int, q: int, d: int = 2048, n_distinct: int = 64, db_seed: int = 0, q_seed: int = 1):
    """Heavily duplicated DB: only ``n_distinct`` different rows, repeated round-robin.

    Every score therefore occurs ``n / n_distinct`` times -- the exact-tie stress case.  The
    framework's tie rule (documented in DESIGN.md) is *lower id first*; numpy's introsort leaves
    ties unspecified (main_retrieve.py:176), so parity on this family is by score, not by id.
    """
    base = rows(n_distinct, d, db_seed)
    db = base[np.arange(n) % n_distinct].copy()
    return reference_layout(db), reference_layout(rows(q, d, q_seed))
