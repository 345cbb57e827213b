"""Diagnostic: self-kNN on compact / two-copy indexes (re-run counts, own-id-first), small multi-batch case and 1M rows."""
import importlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

pkg = importlib.import_module("image-search-engine-for-historical-research_b200")
nat = importlib.import_module(pkg.__name__ + "._native")
synth = importlib.import_module(pkg.__name__ + ".synth")

v, _ = synth.gaussian(20000, 1, d=64)
for compact in (1, 0, 1):
    nat.config_set("compact", compact)
    ix = pkg.ExactIndex(v.T)
    for rep in range(2):
        sims, ids = ix.self_knn(10)
        bad = np.nonzero(ids[:, 0] != np.arange(20000))[0]
        st = ix.stats()
        print(f"small: compact {compact} rep {rep}: rows whose own id is not first: {bad.size} {bad[:12].tolist()} first cols {ids[bad[:3], :3].tolist()} sims {sims[bad[:3], :3].tolist()} reruns {st['n_exact_rerun']} launches {st['gpu_launches']} path {st['path']}", flush=True)
    ix.close()

dev = torch.device("cuda:0")
n = 1007000
rows = bench.synth_rows_device(torch, n, bench.DIM, dev, seed=0)
q = bench.synth_rows_device(torch, 2, bench.DIM, dev, seed=1)
ids = torch.empty((2, 100), dtype=torch.int64, device=dev)
sims = torch.empty((2, 100), dtype=torch.float32, device=dev)
stt = torch.zeros((2,), dtype=torch.int32, device=dev)
for compact in (0, 1):
    nat.config_set("compact", compact)
    ix = pkg.ExactIndex.from_device(rows.data_ptr(), n, bench.DIM, 0)
    for stage in ("fresh", "after scan calls", "after scan_max_q=2 calls"):
        if stage == "after scan calls":
            for _ in range(3):
                ix.search_device(q.data_ptr(), 1, 100, ids.data_ptr(), sims.data_ptr(), status_ptr=stt.data_ptr())
        if stage == "after scan_max_q=2 calls":
            ix.set_param("scan_max_q", 2)
            for _ in range(3):
                ix.search_device(q.data_ptr(), 2, 100, ids.data_ptr(), sims.data_ptr(), status_ptr=stt.data_ptr())
        torch.cuda.synchronize()
        for k in (50, 100):
            t0 = time.perf_counter()
            s_, i_ = ix.self_knn(k, 0, 8192)
            dt = time.perf_counter() - t0
            st = ix.stats()
            print(f"1M: compact {compact} {stage} k {k}: 8192 rows in {dt:.3f} s, reruns {st['n_exact_rerun']} launches {st['gpu_launches']} own-first {bool((i_[:, 0] == np.arange(8192)).all())}", flush=True)
    ix.close()
