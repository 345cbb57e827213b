"""Seeded random-shape sweep of every CUDA path against the oracle (run on a GPU box).

    python tests/fuzz_gpu.py [n_cases] [seed]
"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("image-search-engine-for-historical-research_b200")
from oracle import oracle  # noqa: E402  (checker)

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
bad = 0
t0 = time.time()
for case in range(n_cases):
    n = int(rng.choice([3, 17, 255, 256, 257, 1000, 4097, 9000, 30000]))
    d = int(rng.choice([1, 5, 63, 64, 65, 128, 200, 512, 2048]))
    nq = int(rng.choice([1, 2, 3, 7, 64, 127, 128, 129, 300]))
    k = int(min(n, rng.choice([1, 2, 10, 100, 257, 1000])))
    fam = rng.choice(["gauss", "pos", "dup", "scaled", "f64"])
    if n * d > 40_000_000:
        d = 64
    v = rng.standard_normal((d, n)).astype(np.float32)
    q = rng.standard_normal((d, nq)).astype(np.float32)
    if fam == "pos":
        v, q = np.abs(v), np.abs(q)
    elif fam == "dup":
        v = v[:, rng.integers(0, max(1, n // 7), size=n)]
    elif fam == "scaled":
        v *= rng.uniform(0.01, 100.0, size=(1, n)).astype(np.float32)
        q *= rng.uniform(0.01, 100.0, size=(1, nq)).astype(np.float32)
    renorm = bool(rng.integers(0, 2))
    vin = v.astype(np.float64) if fam == "f64" else v
    if renorm:
        vn = (v.astype(np.float64) / np.maximum(np.linalg.norm(v.astype(np.float64), axis=0, keepdims=True), 1e-300)).astype(np.float32)
        qn = (q.astype(np.float64) / np.maximum(np.linalg.norm(q.astype(np.float64), axis=0, keepdims=True), 1e-300)).astype(np.float32)
    else:
        vn, qn = v, q
    s64 = oracle.scores_f64(vn, qn)
    order = np.lexsort((np.broadcast_to(np.arange(n)[:, None], s64.shape), -s64), axis=0)[:k].T      # fp64 order, ties by id
    with pkg.ExactIndex(vin.T, renormalise=renorm) as ix:
        for path in (0, 1, 2, 3):
            if path == 1 and nq > 16:
                continue
            ix.set_param("force_path", path)
            ids, sims = ix.search(q.T, k, renormalise=renorm)
            scale = np.abs(s64).max() + 1e-30
            for j in range(nq):
                ok, msg = oracle.compare_topk(ids[j], order[j], lambda i, j=j: s64[i, j], rtol=2e-6, atol=2e-6 * scale)
                if not ok:
                    bad += 1
                    print(f"MISMATCH case {case} n={n} d={d} nq={nq} k={k} fam={fam} renorm={renorm} path={path} q{j}: {msg}", flush=True)
                    break
            ref_s = np.take_along_axis(s64, order.T, axis=0).T
            if not np.allclose(sims, ref_s, rtol=2e-5, atol=2e-6 * scale):
                bad += 1
                print(f"SCORES case {case} n={n} d={d} nq={nq} k={k} fam={fam} renorm={renorm} path={path}: max err {np.abs(sims-ref_s).max():.3e}", flush=True)
print(f"{n_cases} cases, {bad} failures, {time.time()-t0:.1f} s")
sys.exit(1 if bad else 0)
