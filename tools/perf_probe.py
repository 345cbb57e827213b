"""Timing probe on a GPU box: cfg2/cfg3-sized database built on the device, several knob settings.

    python tools/perf_probe.py [N] [Q]
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
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 70
K = 100
dev = torch.device("cuda", 0)
rows = bench.synth_rows_device(torch, N, 2048, dev, 0)
queries = bench.synth_rows_device(torch, max(Q, 70), 2048, dev, 1)
ix = pkg.ExactIndex.from_device(rows.data_ptr(), N, 2048, 0)
ids = torch.empty((max(Q, 70), K), dtype=torch.int64, device=dev)
sims = torch.empty((max(Q, 70), K), dtype=torch.float32, device=dev)
status = torch.zeros((max(Q, 70),), dtype=torch.int32, device=dev)


def run(label, nq, reps=10, **params):
    for k_, v_ in params.items():
        ix.set_param(k_, v_)
    ix.set_param("timing", 1)
    coarse, total = [], []
    for i in range(reps + 2):
        ix.search_device(queries.data_ptr(), nq, K, ids.data_ptr(), sims.data_ptr(), status_ptr=status.data_ptr())
        st = ix.stats()
        if i >= 2:
            coarse.append(st["ms_coarse"]); total.append(st["ms_total"])
    torch.cuda.synchronize()
    ix.set_param("timing", 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ix.search_device(queries.data_ptr(), nq, K, ids.data_ptr(), sims.data_ptr(), status_ptr=status.data_ptr())
    e1.record(); torch.cuda.synchronize()
    c = sum(coarse) / len(coarse); t = sum(total) / len(total)
    gbs = N * 2048 * 2 / (c * 1e-3) / 1e9
    print(f"{label:46s} nq={nq:5d} coarse {c:7.3f} ms ({gbs:6.0f} GB/s)  call {t:7.3f} ms  pipelined {e0.elapsed_time(e1)/reps:7.3f} ms/step  "
          f"uncert {int(status[:nq].sum())} cand {st['n_candidates']} launches {st['gpu_launches']}", flush=True)
    for k_ in params:
        ix.set_param(k_, {"sample_pass": 1}.get(k_, 0))


run("gemm default (sample pass on)", Q, reps=40, force_path=2)
if len(sys.argv) > 3 and sys.argv[3] == "only_first":
    sys.exit(0)
run("gemm, no sample pass", Q, force_path=2, sample_pass=0)
run("gemm default nq=1", 1, force_path=2)
run("gemm default nq=16", 16, force_path=2)
run("scan nq=1", 1, force_path=1)
run("scan nq=2", 2, force_path=1)
run("exact nq=4", 4, reps=3, force_path=3)

# compute-bound regime: large query batches through the tcgen05 path
if len(sys.argv) > 3:
    for nq_big in (1024, 4096, 8192):
        qb = bench.synth_rows_device(torch, nq_big, 2048, dev, 2)
        ib = torch.empty((nq_big, K), dtype=torch.int64, device=dev)
        sb = torch.empty((nq_big, K), dtype=torch.float32, device=dev)
        stb = torch.zeros((nq_big,), dtype=torch.int32, device=dev)
        ix.set_param("timing", 1)
        for _ in range(2):
            ix.search_device(qb.data_ptr(), nq_big, K, ib.data_ptr(), sb.data_ptr(), status_ptr=stb.data_ptr())
        st = ix.stats()
        ix.set_param("timing", 0)
        tf = 2.0 * N * 2048 * nq_big / (st["ms_coarse"] * 1e-3) / 1e12
        print(f"gemm nq={nq_big}: coarse {st['ms_coarse']:.3f} ms = {tf:.0f} TFLOP/s (bf16 dense), call {st['ms_total']:.3f} ms, "
              f"uncert {int(stb.sum())}, launches {st['gpu_launches']}", flush=True)

# host-buffer API (e2e): pinned queries in, ids + scores back to the host
import time
qh = queries.cpu().pin_memory().numpy()
for nq in (70, 1):
    for _ in range(3):
        ix.search(qh[:nq], K)
    t0 = time.perf_counter()
    for _ in range(20):
        ix.search(qh[:nq], K)
    dt = (time.perf_counter() - t0) / 20
    print(f"host API nq={nq}: {dt*1e3:.3f} ms/call  ({nq/dt:.0f} QPS)  stats {ix.stats()}", flush=True)
