"""How much headroom does the bf16 error band have?  Shrink eps_sigmas and count queries whose GEMM-path
result differs from the exact fp32 path (1,007,000 x 2048, 70 queries, top-100; families G and P).

    python tools/eps_sweep.py
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("image-search-engine-for-historical-research_b200")

N, D, Q, K = 1_007_000, 2048, 70, 100


def rows(n, seed, fam):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    out = torch.empty((n, D), dtype=torch.float32, device="cuda")
    for lo in range(0, n, 65536):
        hi = min(n, lo + 65536)
        b = torch.randn((hi - lo, D), generator=g, dtype=torch.float32, device="cuda")
        if fam == "P":
            b = b.abs()
        out[lo:hi] = b / b.norm(dim=1, keepdim=True)
    return out


for fam in ("G", "P"):
    db = rows(N, 0, fam)
    q = rows(Q, 1, fam).cpu().numpy()
    ix = pkg.ExactIndex.from_device(db.data_ptr(), N, D, 0)
    ix.set_param("force_path", 3)
    xi, xs = ix.search(q, K)
    ix.set_param("force_path", 2)
    for sig in (8.0, 4.0, 2.0, 1.0, 0.5, 0.25):
        ix.set_param("eps_sigmas", sig)
        gi, gs = ix.search(q, K)
        st = ix.stats()
        wrong = int((gi != xi).any(axis=1).sum())
        print(f"family {fam} eps_sigmas {sig:5.2f}: queries differing from exact {wrong:2d}/70, uncertified->rerun {st['n_exact_rerun']:2d}, "
              f"candidates rescored per query {st['n_candidates']/Q:6.1f}", flush=True)
    ix.close()
    del db
    torch.cuda.empty_cache()
