"""Batch-1 scan over the row-major bf16 copy against the tiled twin, same index, same queries (two-copy index built on
purpose: `compact` 0), then the self-kNN loop on a compact and on a two-copy index.  python tools/scan_probe.py [rows]"""
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

rows_n = int(sys.argv[1]) if len(sys.argv) > 1 else 1007000
dev = torch.device("cuda:0")
rows = bench.synth_rows_device(torch, rows_n, bench.DIM, dev, seed=0)
q = bench.synth_rows_device(torch, 2, bench.DIM, dev, seed=1)
ids = torch.empty((2, 100), dtype=torch.int64, device=dev)
sims = torch.empty((2, 100), dtype=torch.float32, device=dev)
st = torch.zeros((2,), dtype=torch.int32, device=dev)


def time_scan(ix, nq, tiled):
    ix.set_param("force_path", 1)
    ix.set_param("scan_max_q", 2)
    ix.set_param("scan_tiled", tiled)
    for _ in range(10):
        ix.search_device(q.data_ptr(), nq, 100, ids.data_ptr(), sims.data_ptr(), status_ptr=st.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        ix.search_device(q.data_ptr(), nq, 100, ids.data_ptr(), sims.data_ptr(), status_ptr=st.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    whole = e0.elapsed_time(e1) / 100
    ix.set_param("timing", 1)
    ks = []
    for _ in range(20):
        ix.search_device(q.data_ptr(), nq, 100, ids.data_ptr(), sims.data_ptr(), status_ptr=st.data_ptr())
        ks.append(ix.stats()["ms_coarse"])
    ix.set_param("timing", 0)
    return whole, sum(ks) / len(ks), ids.clone(), int(st.sum().item())


nat.config_set("compact", 0)
fat = pkg.ExactIndex.from_device(rows.data_ptr(), rows_n, bench.DIM, 0)
algo = rows_n * bench.DIM * 2
for nq in (1, 2):
    res = {}
    for tiled in (0, 1, 0, 1):
        whole, kern, got, bad = time_scan(fat, nq, tiled)
        res[tiled] = got
        print(f"rows {rows_n} nq {nq} scan over {'tiled twin' if tiled else 'row-major'}: query {whole * 1e3:7.1f} us, scan kernel {kern * 1e3:7.1f} us "
              f"= {algo / kern / 1e6:7.1f} GB/s, uncertified {bad}", flush=True)
    print("  same ids:", bool((res[0][:nq] == res[1][:nq]).all()))
print("two-copy index bytes", fat.device_bytes)
n_knn = min(rows_n, 60000)
for label in ("two-copy", "compact"):
    if label == "compact":
        fat.close()
        nat.config_set("compact", 1)
        ix = pkg.ExactIndex.from_device(rows.data_ptr(), rows_n, bench.DIM, 0)
        print("compact index bytes", ix.device_bytes)
    else:
        ix = fat
    ix.set_param("force_path", 0)
    ix.self_knn(50, 0, 8192)
    t0 = time.perf_counter()
    s_, i_ = ix.self_knn(50, 0, n_knn)
    dt = time.perf_counter() - t0
    print(f"self-kNN rows [0, {n_knn}) of {rows_n}, k=50 on the {label} index: {dt:.3f} s = {2.0 * n_knn * rows_n * bench.DIM / dt / 1e12:.0f} TFLOP/s, first ids {i_[1, :4].tolist()}", flush=True)
