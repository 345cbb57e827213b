"""Short fixed command for ncu: build the cfg2 index on the device, then a few 70-query and batch-1 steps.

    python tools/profile_step.py [N] [steps]
"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("image-search-engine-for-historical-research_b200")
bench = importlib.import_module("bench")

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_007_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
rows = bench.synth_rows_device(torch, N, 2048, dev, 0)
queries = bench.synth_rows_device(torch, 70, 2048, dev, 1)
ix = pkg.ExactIndex.from_device(rows.data_ptr(), N, 2048, 0)
ids = torch.empty((70, 100), dtype=torch.int64, device=dev)
sims = torch.empty((70, 100), dtype=torch.float32, device=dev)
status = torch.zeros((70,), dtype=torch.int32, device=dev)
big = bench.synth_rows_device(torch, 1024, 2048, dev, 2)
ids_b = torch.empty((1024, 100), dtype=torch.int64, device=dev)
sims_b = torch.empty((1024, 100), dtype=torch.float32, device=dev)
status_b = torch.zeros((1024,), dtype=torch.int32, device=dev)
ix.set_param("timing", 1)
for nq in (70, 1, 1024):                      # GEMM path (HBM-bound), scan path, CTA-pair GEMM (tensor-bound)
    q, i, s, st = (big, ids_b, sims_b, status_b) if nq == 1024 else (queries, ids, sims, status)
    for _ in range(steps):
        ix.search_device(q.data_ptr(), nq, 100, i.data_ptr(), s.data_ptr(), status_ptr=st.data_ptr())
    torch.cuda.synchronize()
    print(nq, ix.stats(), int(st[:nq].sum()))
