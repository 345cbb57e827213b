"""Row a3 at scale: the full ranking ranks = argsort(-scores, axis=0) for 70 queries over the 1,007,000-row database
(src/main_retrieve.py:175-176; 7.8 s with numpy on the survey box), through rank_ip(K=None) / xs_rank_all.

    python tools/rank_probe.py [N] [Q]
"""
import importlib, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("image-search-engine-for-historical-research_b200")
bench = importlib.import_module("bench")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_007_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 70
dev = torch.device("cuda", 0)
rows = bench.synth_rows_device(torch, N, 2048, dev, 0)
q = bench.synth_rows_device(torch, Q, 2048, dev, 1).cpu().numpy()
ix = pkg.ExactIndex.from_device(rows.data_ptr(), N, 2048, 0)
ix.rank_all(q[:2])
for rep in range(2):
    t0 = time.perf_counter()
    ranks = ix.rank_all(q)
    dt = time.perf_counter() - t0
    print(f"xs_rank_all: {Q} queries x {N} rows, full ranking into a host int64 [{N}, {Q}] array: {dt:.3f} s ({dt/Q*1e3:.2f} ms per query)", flush=True)
top, _ = ix.search(q, 100)
print("first 100 ranks equal the top-100 search:", bool((ranks[:100].T == top).all()))
