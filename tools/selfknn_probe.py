"""Self-kNN (diffusion.py:67 shape) on one GPU: N x N exhaustive top-k, timing + spot parity.

    python tools/selfknn_probe.py [N] [k]
"""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("image-search-engine-for-historical-research_b200")
bench = importlib.import_module("bench")

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda", 0)
rows = bench.synth_rows_device(torch, N, 2048, dev, 0)
ix = pkg.ExactIndex.from_device(rows.data_ptr(), N, 2048, 0)
LANES = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ix.set_param("self_lanes", LANES)
ix.self_knn(K, 0, min(N, 20000))                     # warm-up: workspaces (and the second lane) exist before the timed run
torch.cuda.synchronize()
t0 = time.perf_counter()
sims, ids = ix.self_knn(K)
dt = time.perf_counter() - t0
print(f"lanes={LANES}", end=" ")
st = ix.stats()
print(f"self-kNN N={N} k={K}: {dt:.3f} s wall ({N/dt:.0f} rows/s, {2.0*N*N*2048/dt/1e12:.0f} TFLOP/s incl. D2H of {ids.nbytes/1e6:.0f}+{sims.nbytes/1e6:.0f} MB), stats {st}")
assert (ids[:, 0] == np.arange(N)).all(), "own id must come first"
# spot parity: 64 rows through the exact fp32 path
pick = np.linspace(0, N - 1, 64).astype(np.int64)
ix.set_param("force_path", 3)
q = rows[torch.from_numpy(pick).to(dev)].cpu().numpy()
xi, xs = ix.search(q, K)
bad = 0
for j, r in enumerate(pick):
    a, b = ids[r], xi[j]
    # the self row is forced first in self_knn; in a plain search it is first anyway (score ~1)
    if not (a == b).all():
        diff = np.nonzero(a != b)[0]
        if np.abs(sims[r][diff] - xs[j][diff]).max() > 1e-6 * np.abs(xs[j][diff]).max():
            bad += 1
print(f"spot parity vs exact path: {64-bad}/64 rows ok, max |score diff| {np.abs(sims[pick]-xs).max():.2e}")
sys.exit(1 if bad else 0)
