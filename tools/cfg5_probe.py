"""BASELINE config 5 under torchrun: the N x N self-kNN graph (src/utils/diffusion.py:67), two ways:

  replicated   the database REPLICATED on every GPU, the query rows sharded -- no data-path collective
  rowsharded   the database ROW-SHARDED (each GPU holds N/G rows): blocks of query rows are broadcast by their owner,
               searched by every shard, the per-shard lists exchanged through the peer mailboxes and merged

    python -m torch.distributed.run --nproc-per-node G tools/cfg5_probe.py [N] [k] [replicated|rowsharded|both]

Prints one JSON line per variant.  Parity: 128 sample rows of rank 0's range against the oracle on the host.
"""
import importlib
import json
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
oracle = importlib.import_module("oracle.oracle")           # checker only (tools/ is not product)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 100
which = sys.argv[3] if len(sys.argv) > 3 else "both"
D = 2048
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rows = bench.synth_rows_device(torch, N, D, dev, seed=0)        # identical on every rank
b = sharded.shard_bounds(N, world)
lo, hi = b[rank], b[rank + 1]
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
rows_h = rows.cpu().numpy() if rank == 0 else None
sample = np.linspace(lo, hi - 1, num=min(128, hi - lo)).astype(np.int64)


def parity(ids, sims):
    """rank 0: its sample rows against np.dot + top-k on the host (own id first; the rest as the oracle ranks them)."""
    q = rows_h[sample]
    ref_s, ref_i = oracle.knn_search(rows_h, q, K)
    ok = True
    for j, r in enumerate(sample):
        got = ids[r - lo]
        good, _ = oracle.compare_topk(got, ref_i[j], lambda i, j=j: rows_h[i].astype(np.float64) @ q[j].astype(np.float64))
        ok = ok and good and int(got[0]) == int(r)
    return bool(ok and np.allclose(sims[sample - lo], ref_s, rtol=1e-5, atol=1e-7))


def report(variant, seconds, reruns, ok_first, par):
    t = torch.tensor([seconds], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        tm = float(t.item())
        fl = 2.0 * N * N * D
        sust = peaks.get("bf16_tflops_sustained")
        print(json.dumps({"probe": "cfg5", "variant": variant, "n_rows": N, "k": K, "n_gpus": world, "query_rows_per_gpu": hi - lo,
                          "seconds": tm, "rows_per_s": N / tm, "tflops_aggregate": fl / tm / 1e12, "tflops_per_gpu": fl / tm / 1e12 / world,
                          "frac_of_sustained_peak": fl / tm / 1e12 / world / sust if sust else None,
                          "exact_reruns_rank0": int(reruns), "own_id_first": ok_first, "parity_128_rows_vs_oracle": par}), flush=True)


if which in ("replicated", "both"):
    ix = pkg.ExactIndex.from_device(rows.data_ptr(), N, D, local)
    ix.self_knn(K, lo, min(hi, lo + 8192))                              # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    sims, ids = ix.self_knn(K, lo, hi)                                  # results land in host memory
    dt = time.perf_counter() - t0
    st = ix.stats()
    report("replicated database, query rows sharded, no collective", dt, st["n_exact_rerun"],
           bool((ids[:, 0] == np.arange(lo, hi)).all()), parity(ids, sims) if rank == 0 else None)
    ix.close()
    del ix
    torch.cuda.empty_cache()

if which in ("rowsharded", "both"):
    local_rows = rows[lo:hi].contiguous()
    ix = pkg.ExactIndex.from_device(local_rows.data_ptr(), hi - lo, D, local, id_offset=lo)
    block = 8192
    exchange = sharded.PeerExchange(local, block, K) if world > 1 else None
    shard, searcher = sharded.make_searcher(ix, local, lanes=2, exchange=exchange)
    for _ in range(2):                                                  # warm-up: both lanes / mailbox slots
        searcher.search(rows[:block].contiguous(), K)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    sims_t, ids_t = sharded.self_knn_rowsharded(searcher, local_rows, b, K, rank, block=block)
    ids, sims = ids_t.cpu().numpy(), sims_t.cpu().numpy()              # results to host memory, as the other variant
    dt = time.perf_counter() - t0
    report("row-sharded database, query blocks broadcast by their owner, peer-mailbox exchange + merge", dt, searcher.n_rerun,
           bool((ids[:, 0] == np.arange(lo, hi)).all()), parity(ids, sims) if rank == 0 else None)
    if exchange is not None:
        exchange.close()
if world > 1:
    dist.destroy_process_group()
