"""BASELINE config 5 shape under torchrun: the N x N self-kNN graph (src/utils/diffusion.py:67) with the
database REPLICATED on every GPU and the query rows sharded -- no data-path collective.

    python -m torch.distributed.run --nproc-per-node G tools/cfg5_probe.py [N] [k]
"""
import importlib
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("image-search-engine-for-historical-research_b200")
sharded = importlib.import_module("image-search-engine-for-historical-research_b200.sharded")
bench = importlib.import_module("bench")

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 100
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rows = bench.synth_rows_device(torch, N, 2048, dev, seed=0)        # identical on every rank
ix = pkg.ExactIndex.from_device(rows.data_ptr(), N, 2048, local)
del rows
torch.cuda.empty_cache()
b = sharded.shard_bounds(N, world)
lo, hi = b[rank], b[rank + 1]
ix.self_knn(K, lo, min(hi, lo + 8192))                              # warm-up
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
sims, ids = ix.self_knn(K, lo, hi)
dt = time.perf_counter() - t0
st = ix.stats()
ok = bool((ids[:, 0] == np.arange(lo, hi)).all())
t = torch.tensor([dt], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    tm = float(t.item())
    print(f"cfg5-shape: {N} x {N} self-kNN k={K} on {world} GPU(s), {hi-lo} query rows per GPU: {tm:.3f} s wall (max over ranks, results in host memory), "
          f"{N/tm:.0f} rows/s, {2.0*N*N*2048/tm/1e12:.0f} TFLOP/s aggregate, exact reruns on rank 0: {st['n_exact_rerun']}, own id first: {ok}", flush=True)
if world > 1:
    dist.destroy_process_group()
