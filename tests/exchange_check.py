"""Run under torchrun (one process per rank): the row-sharded search with the peer-memory exchange (xs_search_dev_push /
xs_exchange_*) against the NCCL all-gather path and the oracle, blocking and pipelined, including the families whose
queries need the collective exact re-run.  Launched by test_gpu_exchange.py:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tests/exchange_check.py

XS_CHECK_ONE_DEVICE=1: every rank uses device 0 and the process group is gloo (NCCL refuses two ranks on one GPU);
the mailboxes are mapped through CUDA IPC all the same, so the exchange kernels run exactly as they do across GPUs.
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


def check_against_oracle(tag, rank, got_i, got_s, vecs, qvecs, k):
    ref_i, ref_s = oracle.topk_ip(vecs, qvecs, k)
    s64 = oracle.scores_f64(vecs, qvecs)
    for j in range(qvecs.shape[1]):
        ok, msg = oracle.compare_topk(got_i[j], ref_i[j], lambda i, j=j: s64[i, j])
        assert ok, f"rank {rank} {tag} query {j}: {msg}"
    np.testing.assert_allclose(got_s, ref_s, rtol=1e-5, atol=1e-7)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    one_device = os.environ.get("XS_CHECK_ONE_DEVICE") == "1"
    if one_device:
        local = 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if one_device:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)
    n, d = 20011, 256
    vecs, qvecs = synth.gaussian(n, 70, d=d)                       # (d, n), (d, 70): same on every rank (seeded)
    b = sharded.shard_bounds(n, world)
    lo, hi = b[rank], b[rank + 1]
    index = pkg.ExactIndex(np.ascontiguousarray(vecs.T[lo:hi]), device=local, id_offset=lo)
    q_all = torch.from_numpy(np.ascontiguousarray(qvecs.T)).to(dev)
    exchange = sharded.PeerExchange(local, 70, 100)
    shard, peer = sharded.make_searcher(index, local, lanes=2, exchange=exchange)
    nccl = None if one_device else sharded.ShardedSearcher(shard.local_search, shard.merge)
    for nq, k in ((70, 100), (1, 100), (33, 7), (70, 100)):
        q = q_all[:nq].contiguous()
        got_i, got_s = [t.cpu().numpy().copy() for t in peer.search(q, k)]
        if nccl is not None:
            want_i, want_s = [t.cpu().numpy().copy() for t in nccl.search(q, k)]
            assert np.array_equal(got_i, want_i) and np.array_equal(got_s, want_s), f"rank {rank}: peer != nccl at nq={nq} k={k}"
        check_against_oracle(f"nq={nq} k={k}", rank, got_i, got_s, vecs, qvecs[:, :nq], k)
    # pipelined on two lanes: two searches in flight, ranks deliberately out of step, 40 epochs over both slots
    q = q_all.contiguous()
    want_i = peer.search(q, 100)[0].cpu().numpy().copy()
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
    assert peer.n_rerun == 0
    # the native two-slot pipeline (xs_pipeline_*: one C call per step) over its own mailboxes: same answers
    ex_p = sharded.PeerExchange(local, 70, 100)
    _, pipe = sharded.make_searcher(index, local, lanes=2, exchange=ex_p, pipeline=(70, 100))
    for nq, k in ((70, 100), (1, 100), (33, 7)):
        gi, gs = [t.cpu().numpy().copy() for t in pipe.search(q_all[:nq].contiguous(), k)]
        wi, ws = [t.cpu().numpy().copy() for t in peer.search(q_all[:nq].contiguous(), k)]
        assert np.array_equal(gi, wi) and np.array_equal(gs, ws), f"rank {rank}: pipeline != searcher at nq={nq} k={k}"
    pending = None
    for it in range(30):
        if it % 5 == rank:
            torch.cuda._sleep(10_000_000)
        nxt = pipe.search_async(q, 100)
        if pending is not None:
            assert np.array_equal(pending.result()[0].cpu().numpy(), want_i), f"rank {rank}: native pipeline step {it}"
        pending = nxt
    assert np.array_equal(pending.result()[0].cpu().numpy(), want_i)
    h1, h2, h3 = pipe.search_async(q, 100), pipe.search_async(q, 100), pipe.search_async(q, 100)     # the third collects the first itself
    for h in (h1, h2, h3):
        assert np.array_equal(h.result()[0].cpu().numpy(), want_i)
    pipe.close()
    ex_p.close()
    # the one-call form with the merge riding in the search's last kernel (xs_search_dev_exchange), both slots, twice round
    ex_f = sharded.PeerExchange(local, 70, 100)
    for step in range(4):
        fi, fs, fst = shard.local_exchange(q, 100, ex_f, step & 1)
        torch.cuda.synchronize()
        assert np.array_equal(fi.cpu().numpy(), want_i) and int(fst.sum().item()) == 0, f"rank {rank}: fused exchange step {step}"
    ex_f.close()
    index.close()

    # families where shards cannot certify every query: the certificate words travel with the lists and the flagged
    # queries are re-run on the exact path by every rank together
    for name, (v2, q2), k in (("crowded", synth.clustered(12000, 12, d=256, n_clusters=6, noise=0.02)[:2], 100),
                              ("duplicates", synth.ties(3000, 8, d=64, n_distinct=6), 40)):
        n2 = v2.shape[1]
        b2 = sharded.shard_bounds(n2, world)
        ix2 = pkg.ExactIndex(np.ascontiguousarray(v2.T[b2[rank]:b2[rank + 1]]), device=local, id_offset=b2[rank])
        ex2 = sharded.PeerExchange(local, q2.shape[1], k)
        _, s2 = sharded.make_searcher(ix2, local, exchange=ex2, pipeline=(q2.shape[1], k) if name == "crowded" else None)
        qd = torch.from_numpy(np.ascontiguousarray(q2.T)).to(dev)
        for _ in range(2):
            gi, gs = [t.cpu().numpy().copy() for t in s2.search(qd, k)]
            check_against_oracle(name, rank, gi, gs, v2, q2, k)
        if rank == 0:
            print(f"{name}: {s2.n_rerun} queries re-run on the exact path", flush=True)
        ex2.close()
        ix2.close()
    exchange.close()
    dist.barrier()
    if rank == 0:
        print("exchange_check ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
