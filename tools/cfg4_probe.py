"""BASELINE config 4 shape under torchrun: a large database row-sharded across the GPUs of the box, a
large query batch, exact top-100 with the NCCL candidate merge.

    python -m torch.distributed.run --nproc-per-node G tools/cfg4_probe.py [N_total] [Q] [reps]
"""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("image-search-engine-for-historical-research_b200")
sharded = importlib.import_module("image-search-engine-for-historical-research_b200.sharded")
bench = importlib.import_module("bench")

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
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
del rows
torch.cuda.empty_cache()
queries = bench.synth_rows_device(torch, Q, D, dev, seed=1)
shard = sharded.CudaShard(index, local)
searcher = sharded.ShardedSearcher(shard.local_search, shard.merge)
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
unc = shard.uncertified(Q, K)
index.set_param("timing", 1)
shard.local_search(queries, K)
st = index.stats()
if rank == 0:
    t = float(ms.item()) * 1e-3
    fl = 2.0 * N * D * Q
    print(f"cfg4-shape: N={N} rows over {world} GPUs ({hi-lo} rows/GPU), Q={Q}, top-{K}: {t*1e3:.1f} ms/batch -> {Q/t:.0f} QPS, "
          f"{fl/t/1e12:.0f} TFLOP/s aggregate ({fl/t/1e12/world:.0f} per GPU), local stats {st}, uncertified {unc}, "
          f"ids range [{int(ids.min())}, {int(ids.max())}]", flush=True)
if world > 1:
    dist.destroy_process_group()
