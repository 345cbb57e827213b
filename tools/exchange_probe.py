"""Where a sharded 70-query step goes (torchrun, one process per GPU): device time of the local search alone, of the
search with the exchange's sending end fused in, of the merge, and of whole steps -- blocking and pipelined.

    python -m torch.distributed.run --nproc-per-node G tools/exchange_probe.py [steps]
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

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N, D, Q, K = 1_007_000, 2048, 70, 100
b = sharded.shard_bounds(N, world)
lo, hi = b[rank], b[rank + 1]
rows = bench.synth_rows_device(torch, hi - lo, D, dev, seed=100 + rank)
index = pkg.ExactIndex.from_device(rows.data_ptr(), hi - lo, D, local, id_offset=lo)
queries = bench.synth_rows_device(torch, Q, D, dev, seed=1)
ex = sharded.PeerExchange(local, Q, K)
shard, searcher = sharded.make_searcher(index, local, exchange=ex)
ex2 = sharded.PeerExchange(local, Q, K)
_, pipe = sharded.make_searcher(index, local, lanes=int(os.environ.get("XS_LANES", "1")), exchange=ex2, pipeline=(Q, K))


def timed(fn, n=steps, sync_each=False):
    for _ in range(10):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
        if sync_each:
            torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


slot = [0]
def local_only():
    shard.local_search(queries, K, 0)
def push_merge():
    shard.local_push(queries, K, ex, slot[0])
    ex.merge(Q, K, slot[0])
    slot[0] ^= 1
pend = [None]
def piped():
    nxt = pipe.search_async(queries, K)
    if pend[0] is not None:
        pend[0].result()
    pend[0] = nxt

t_local = timed(local_only)
t_pm = timed(push_merge)
t_pm_sync = timed(push_merge, sync_each=True)
t_pipe = timed(piped)
pend[0].result()
index.set_param("timing", 1)
shard.local_search(queries, K, 0)
st = index.stats()
# timeline of the finalise launch with the push fused in (rank 0)
import ctypes as C
import numpy as np
nat = importlib.import_module("image-search-engine-for-historical-research_b200._native")
index.set_param("timing", 0)
index.set_param("boot_trace", 1)
for _ in range(6):
    push_merge()
torch.cuda.synchronize(); dist.barrier()
if rank == 0:
    buf = np.zeros((2048, 10), dtype=np.uint64)
    g = C.c_int(0)
    nat.check(nat.load().xs_debug_trace(index._h, 1, buf.ctypes.data, 2048, C.byref(g)), "trace")
    t = buf[: g.value].astype(np.int64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3
    for j, nm in enumerate(["start", "sizes", "gathered", "cut", "collected", "rescored", "ticket", "sorted(last)", "emitted+pushed(last)"]):
        col = rel[:, j][t[:, j] > 0]
        if col.size:
            print(f"  finalise+push {nm:24s} min {col.min():7.1f}  median {np.median(col):7.1f}  max {col.max():7.1f}   ({col.size} CTAs)")
index.set_param("boot_trace", 0)
# two consecutive pipelined steps on the device clock: when does each kernel of each lane start and end?
lib = nat.load()
lanes_h = [C.c_void_p(lib.xs_pipeline_lane(pipe._h, l)) for l in range(int(os.environ.get("XS_LANES", "1")))]
for hnd in lanes_h:
    nat.check(lib.xs_set_param(hnd, b"boot_trace", 1.0), "set")
for _ in range(12):
    piped()
pend[0].result(); pend[0] = None
torch.cuda.synchronize(); dist.barrier()
if rank == 0:
    ev = []
    for l, hnd in enumerate(lanes_h):
        for which, width, name, c_end in ((0, 8, "scan+select GEMM", 6), (1, 10, "finalise+push", 8)):
            buf = np.zeros((2048, width), dtype=np.uint64)
            g = C.c_int(0)
            nat.check(lib.xs_debug_trace(hnd, which, buf.ctypes.data, 2048, C.byref(g)), "trace")
            t = buf[: g.value].astype(np.int64)
            t = t[t[:, 0] > 0]
            if len(t):
                ends = t[:, c_end][t[:, c_end] > 0]
                ev.append((int(t[:, 0].min()), int(ends.max() if len(ends) else t.max()), f"lane {l} {name}"))
    if ev:
        t0 = min(e[0] for e in ev)
        for a_, b_, nm in sorted(ev):
            print(f"  pipelined timeline: {nm:28s} {(a_ - t0) / 1e3:8.1f} .. {(b_ - t0) / 1e3:8.1f} us")
for hnd in lanes_h:
    nat.check(lib.xs_set_param(hnd, b"boot_trace", 0.0), "set")
if rank == 0:
    print(f"{world} GPUs, {hi - lo} rows/GPU: local search {t_local:.1f} us | search+push+merge back to back {t_pm:.1f} us | "
          f"the same, host-synchronised per step {t_pm_sync:.1f} us | native pipeline (2 in flight) {t_pipe:.1f} us | "
          f"coarse kernel {st['ms_coarse']*1e3:.1f} us, local call {st['ms_total']*1e3:.1f} us", flush=True)
pipe.close(); ex2.close(); ex.close()
dist.destroy_process_group()
