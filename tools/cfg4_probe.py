"""BASELINE config 4 under torchrun: a large database row-sharded across the GPUs of the box, a large query batch,
exact top-100, per-shard lists exchanged through the peer mailboxes (or NCCL) and merged.  Prints one JSON line.

    python -m torch.distributed.run --nproc-per-node G tools/cfg4_probe.py [N_total] [Q] [reps] [peer|nccl]
"""
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("image-search-engine-for-historical-research_b200")
sharded = importlib.import_module("image-search-engine-for-historical-research_b200.sharded")
bench = importlib.import_module("bench")
oracle = importlib.import_module("oracle.oracle")           # checker only (tools/ is not product)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
how = sys.argv[4] if len(sys.argv) > 4 else "peer"
K, D = 100, 2048
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
b = sharded.shard_bounds(N, world)
lo, hi = b[rank], b[rank + 1]
rows = bench.synth_rows_device(torch, hi - lo, D, dev, seed=100 + rank)       # each rank draws its own shard
index = pkg.ExactIndex.from_device(rows.data_ptr(), hi - lo, D, local, id_offset=lo)
queries = bench.synth_rows_device(torch, Q, D, dev, seed=1)
exchange = sharded.PeerExchange(local, Q, K) if (world > 1 and how == "peer") else None
shard, searcher = sharded.make_searcher(index, local, exchange=exchange)
for _ in range(2):
    ids, sims = searcher.search(queries, K)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ids, sims = searcher.search(queries, K)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
index.set_param("timing", 1)
shard.local_search(queries, K)
st = index.stats()
index.set_param("timing", 0)
# parity 1: the merged answer against a host merge of every shard's EXACT fp32 path, 8 queries
merged_ok = bench.merged_parity_check(torch, dist, index, queries, ids, world, rank) if world > 1 else None
# parity 2: rank 0's shard against the oracle on the host (np.dot + top-k over its rows), 8 queries
oracle_ok = None
if rank == 0:
    rows_h = rows.cpu().numpy()
    q8 = queries[:8].cpu().numpy()
    ref_i, ref_s = oracle.topk_ip(rows_h.T, q8.T, K)
    li, ls = index.search(q8, K)
    oracle_ok = True
    for j in range(8):
        ok, msg = oracle.compare_topk(li[j] - lo, ref_i[j], lambda i, j=j: rows_h[i].astype(np.float64) @ q8[j].astype(np.float64))
        oracle_ok = oracle_ok and ok
    oracle_ok = bool(oracle_ok and np.allclose(ls, ref_s, rtol=1e-5, atol=1e-7))
if rank == 0:
    t = float(ms.item()) * 1e-3
    fl = 2.0 * N * D * Q
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    sust, burst = peaks.get("bf16_tflops_sustained"), peaks.get("bf16_tflops")
    per_gpu = fl / t / 1e12 / world
    print(json.dumps({
        "probe": "cfg4", "n_rows": N, "n_gpus": world, "rows_per_gpu": hi - lo, "queries": Q, "k": K, "exchange": how if world > 1 else "none",
        "ms_per_batch": t * 1e3, "queries_per_s": Q / t, "tflops_aggregate": fl / t / 1e12, "tflops_per_gpu": per_gpu,
        "frac_of_sustained_peak": per_gpu / sust if sust else None, "frac_of_burst_peak": per_gpu / burst if burst else None,
        "local_gemm_ms": st["ms_coarse"], "local_call_ms": st["ms_total"], "local_gemm_tflops": 2.0 * (hi - lo) * D * Q / (st["ms_coarse"] * 1e-3) / 1e12,
        "exact_reruns": int(searcher.n_rerun), "merged_parity_vs_exact_shards": merged_ok, "shard0_parity_vs_oracle": oracle_ok,
        "ids_range": [int(ids.min()), int(ids.max())]}), flush=True)
if exchange is not None:
    exchange.close()
if world > 1:
    dist.destroy_process_group()
