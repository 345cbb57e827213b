"""Diagnostic run on a GPU box (lives under tests/ because it uses the oracle as its checker): every path
against the oracle on cfg1-sized inputs, verbose.

    python tests/gpu_check.py [scan|exact|gemm|all] [N] [Q] [D] [K]
"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("image-search-engine-for-historical-research_b200")
synth = importlib.import_module("image-search-engine-for-historical-research_b200.synth")
from oracle import oracle  # noqa: E402  (checker only)

which = sys.argv[1] if len(sys.argv) > 1 else "all"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4993
Q = int(sys.argv[3]) if len(sys.argv) > 3 else 70
D = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
K = int(sys.argv[5]) if len(sys.argv) > 5 else 100

vecs, qvecs = synth.gaussian(N, Q, d=D)
t0 = time.time()
ref_ids, ref_sims = oracle.topk_ip(vecs, qvecs, K)
s64 = oracle.scores_f64(vecs, qvecs)
print(f"oracle: {time.time()-t0:.2f}s  N={N} Q={Q} D={D} K={K}", flush=True)

ix = pkg.ExactIndex(vecs.T, renormalise=False)
print("index built, device bytes", ix.device_bytes, flush=True)
paths = {"scan": 1, "gemm": 2, "exact": 3}
rc = 0
for name, code in paths.items():
    if which not in ("all", name):
        continue
    ix.set_param("force_path", code)
    t0 = time.time()
    try:
        ids, sims = ix.search(qvecs.T, K)
    except Exception as e:  # noqa: BLE001
        print(f"[{name}] FAILED: {type(e).__name__}: {e}", flush=True)
        rc = 1
        continue
    dt = time.time() - t0
    bad = 0
    first = ""
    for j in range(Q):
        ok, msg = oracle.compare_topk(ids[j], ref_ids[j], lambda i, j=j: s64[i, j])
        if not ok:
            bad += 1
            first = first or f"q{j}: {msg}"
    exact_eq = int((ids == ref_ids).all(axis=1).sum())
    rel = np.max(np.abs(sims - ref_sims) / np.maximum(np.abs(ref_sims), 1e-30))
    st = ix.stats()
    print(f"[{name}] {dt*1e3:.1f} ms  queries ok {Q-bad}/{Q} (identical lists {exact_eq}/{Q})  max rel score err {rel:.2e}  stats {st}  {first}", flush=True)
    if bad:
        rc = 1
sys.exit(rc)
