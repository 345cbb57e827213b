"""Run under torchrun (one process per GPU): the peer-memory exchange (xs_exchange_*) against the NCCL
all-gather path and the oracle, row-sharded, blocking and pipelined.  Launched by test_gpu_exchange.py:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tests/exchange_check.py
"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("image-search-engine-for-historical-research_b200")
sharded = importlib.import_module("image-search-engine-for-historical-research_b200.sharded")
synth = importlib.import_module("image-search-engine-for-historical-research_b200.synth")
oracle = importlib.import_module("oracle.oracle")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, d = 20011, 256
    vecs, qvecs = synth.gaussian(n, 70, d=d)                       # (d, n), (d, 70): same on every rank (seeded)
    b = sharded.shard_bounds(n, world)
    lo, hi = b[rank], b[rank + 1]
    index = pkg.ExactIndex(np.ascontiguousarray(vecs.T[lo:hi]), device=local, id_offset=lo)
    shard = sharded.CudaShard(index, local)
    q_all = torch.from_numpy(np.ascontiguousarray(qvecs.T)).to(dev)
    exchange = sharded.PeerExchange(local, sharded.packed_bytes(70, 100))
    peer = sharded.ShardedSearcher(shard.local_search, shard.merge, exchange=exchange, exchange_pipelined=True)
    nccl = sharded.ShardedSearcher(shard.local_search, shard.merge)
    s64 = oracle.scores_f64(vecs, qvecs)
    for nq, k in ((70, 100), (1, 100), (33, 7), (70, 100)):
        q = q_all[:nq].contiguous()
        want_i, want_s = [t.cpu().numpy().copy() for t in nccl.search(q, k)]
        got_i, got_s = [t.cpu().numpy().copy() for t in peer.search(q, k)]
        assert np.array_equal(got_i, want_i) and np.array_equal(got_s, want_s), f"rank {rank}: peer != nccl at nq={nq} k={k}"
        ref_i, _ = oracle.topk_ip(vecs, qvecs[:, :nq], k)
        for j in range(nq):
            ok, msg = oracle.compare_topk(got_i[j], ref_i[j], lambda i, j=j: s64[i, j])
            assert ok, f"rank {rank} nq={nq} k={k} query {j}: {msg}"
    # pipelined: two searches in flight, ranks deliberately out of step, 40 epochs over both slots
    q = q_all.contiguous()
    want_i = nccl.search(q, 100)[0].cpu().numpy().copy()
    pending = None
    for it in range(40):
        if it % 7 == rank:
            torch.cuda._sleep(20_000_000)                            # ~10 ms of device time: this rank lags
        nxt = peer.search_async(q, 100)
        if pending is not None:
            assert np.array_equal(pending.result()[0].cpu().numpy(), want_i), f"rank {rank}: pipelined step {it}"
        pending = nxt
    assert np.array_equal(pending.result()[0].cpu().numpy(), want_i)
    # a third search without collecting the second: the searcher enqueues the overdue merge itself
    h1, h2, h3 = peer.search_async(q, 100), peer.search_async(q, 100), peer.search_async(q, 100)
    for h in (h1, h2, h3):
        assert np.array_equal(h.result()[0].cpu().numpy(), want_i)
    exchange.close()
    dist.barrier()
    if rank == 0:
        print("exchange_check ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
