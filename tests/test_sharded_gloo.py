"""world_size-2 gloo test of the row-sharded search logic on CPU.

The CUDA local searcher cannot run here, so the test injects the oracle as each rank's local
searcher and the host restatement of the merge; what is under test is the product's sharding
code: shard bounds, global id offsets, the all-gather layout [G, nq, k] and the merge semantics
(descending score, ties by ascending id) -- SURVEY.md section 8(e).
"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "image-search-engine-for-historical-research_b200"


def _worker(rank, world, port, n, nq, d, k, ties, out_dir, flag_queries=()):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sharded = importlib.import_module(PKG + ".sharded")
    synth = importlib.import_module(PKG + ".synth")
    oracle = importlib.import_module("oracle.oracle")
    vecs, qvecs = (synth.ties(n, nq, d=d, n_distinct=8) if ties else synth.gaussian(n, nq, d=d))
    bounds = sharded.shard_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    shard = np.ascontiguousarray(vecs[:, lo:hi])

    calls = {"coarse": 0, "exact": 0}

    def local_search(queries, kk, exact=False):
        """Stand-in for the CUDA shard: the oracle's exact answer.  To exercise the certificate plumbing, rank 1
        claims it could NOT certify the queries listed in `flag_queries` and hands out a deliberately wrong list
        for them (its worst rows) unless asked for the exact path."""
        calls["exact" if exact else "coarse"] += 1
        ids, sims = oracle.topk_ip(shard, queries.numpy().T, min(kk, hi - lo))
        pad = kk - ids.shape[1]
        if pad:
            ids = np.concatenate([ids, -np.ones((ids.shape[0], pad), np.int64) - lo], axis=1)
            sims = np.concatenate([sims, np.full((sims.shape[0], pad), -np.inf, np.float32)], axis=1)
        status = np.zeros(ids.shape[0], np.int32)
        if not exact and rank == 1:
            for j in flag_queries:
                if j < ids.shape[0]:
                    status[j] = 1
                    sims[j] = -1.0                      # garbage the merge would otherwise rank last
        packed = torch.zeros((sharded.packed_bytes(ids.shape[0], kk),), dtype=torch.uint8)
        pi, ps, pst = sharded.unpack(packed, ids.shape[0], kk)
        pi.copy_(torch.from_numpy(ids + lo)); ps.copy_(torch.from_numpy(sims)); pst.copy_(torch.from_numpy(status))
        return packed

    def merge(packed_all, world_, nq_, kk):
        parts = [sharded.unpack(packed_all[g * sharded.packed_bytes(nq_, kk):(g + 1) * sharded.packed_bytes(nq_, kk)], nq_, kk)
                 for g in range(world_)]
        i, s = oracle.merge_parts(np.stack([p[0].numpy() for p in parts]), np.stack([p[1].numpy() for p in parts]), kk)
        st = np.bitwise_or.reduce(np.stack([p[2].numpy() for p in parts]), axis=0)
        return torch.from_numpy(i), torch.from_numpy(s), torch.from_numpy(st)

    searcher = sharded.ShardedSearcher(local_search, merge)
    qt = torch.from_numpy(np.ascontiguousarray(qvecs.T))
    # two searches in flight (the throughput form): the second one, on the reversed queries, must not disturb the first
    h1 = searcher.search_async(qt, k)
    h2 = searcher.search_async(torch.flip(qt, dims=[0]).contiguous(), k)
    ids, sims = h1.result()
    ids2, sims2 = h2.result()
    assert torch.equal(torch.flip(ids2, dims=[0]), ids) and torch.equal(torch.flip(sims2, dims=[0]), sims)
    ids_b, sims_b = searcher.search(qt, k)                # blocking form agrees
    assert torch.equal(ids_b, ids) and torch.equal(sims_b, sims)
    if flag_queries:                                      # every rank re-ran the flagged queries, and only those
        assert calls["exact"] == 3 and searcher.n_rerun == 3 * len([j for j in flag_queries if j < nq]), (calls, searcher.n_rerun)
    else:
        assert calls["exact"] == 0 and searcher.n_rerun == 0
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ids=ids.numpy(), sims=sims.numpy())
    dist.destroy_process_group()


def _run(tmp_path, n, nq, d, k, ties=False, world=2, flag_queries=()):
    port = 29000 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n, nq, d, k, ties, str(tmp_path), tuple(flag_queries)), nprocs=world, join=True)
    return [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]


def test_shard_bounds():
    sharded = importlib.import_module(PKG + ".sharded")
    assert sharded.shard_bounds(10, 3) == [0, 4, 7, 10]
    assert sharded.shard_bounds(1_007_000, 8)[-1] == 1_007_000
    b = sharded.shard_bounds(1_007_000, 8)
    assert max(b[i + 1] - b[i] for i in range(8)) - min(b[i + 1] - b[i] for i in range(8)) <= 1


def test_two_rank_search_matches_unsharded(tmp_path, synth, oracle):
    outs = _run(tmp_path, n=601, nq=5, d=48, k=9)
    vecs, qvecs = synth.gaussian(601, 5, d=48)
    ref_ids, ref_sims = oracle.topk_ip(vecs, qvecs, 9)
    for o in outs:                       # every rank ends up with the full answer
        np.testing.assert_array_equal(o["ids"], ref_ids)
        np.testing.assert_array_equal(o["sims"], ref_sims)


def test_uncertified_queries_are_rerun_on_every_rank(tmp_path, synth, oracle):
    """A shard that cannot certify a query says so in the status word that travels with its list; the merged word
    is the OR over the shards, so every rank re-runs the same queries on the exact path (the sharded
    search used to drop the certificate on the floor)."""
    outs = _run(tmp_path, n=601, nq=5, d=48, k=9, flag_queries=(1, 3))
    vecs, qvecs = synth.gaussian(601, 5, d=48)
    ref_ids, ref_sims = oracle.topk_ip(vecs, qvecs, 9)
    for o in outs:
        np.testing.assert_array_equal(o["ids"], ref_ids)
        np.testing.assert_array_equal(o["sims"], ref_sims)


def test_two_rank_ties_prefer_lower_global_id(tmp_path, synth, oracle):
    outs = _run(tmp_path, n=64, nq=3, d=16, k=20, ties=True)
    vecs, qvecs = synth.ties(64, 3, d=16, n_distinct=8)
    ref_ids, ref_sims = oracle.topk_ip(vecs, qvecs, 20)
    np.testing.assert_array_equal(outs[0]["ids"], ref_ids)
    np.testing.assert_array_equal(outs[1]["ids"], ref_ids)


def test_k_larger_than_a_shard(tmp_path, synth, oracle):
    outs = _run(tmp_path, n=21, nq=2, d=8, k=15)      # shards of 11 and 10 rows, k = 15
    vecs, qvecs = synth.gaussian(21, 2, d=8)
    ref_ids, _ = oracle.topk_ip(vecs, qvecs, 15)
    np.testing.assert_array_equal(outs[0]["ids"], ref_ids)


def test_shard_bounds_and_packing_properties():
    """Host-side invariants of the sharding helpers (no process group needed)."""
    import importlib
    import torch
    from hypothesis import given, settings, strategies as st
    from conftest import PKG_NAME
    sharded = importlib.import_module(PKG_NAME + ".sharded")
    oracle = importlib.import_module("oracle.oracle")

    @settings(max_examples=200, deadline=None)
    @given(st.integers(1, 10_000_000), st.integers(1, 16))
    def bounds(n, world):
        b = sharded.shard_bounds(n, world)
        sizes = np.diff(b)
        assert b[0] == 0 and b[-1] == n and len(b) == world + 1
        assert (sizes >= 0).all() and sizes.max() - sizes.min() <= 1 and (np.diff(sizes) <= 0).all()
    bounds()

    @settings(max_examples=100, deadline=None)
    @given(st.integers(1, 300), st.integers(1, 130))
    def packing(nq, k):
        nb = sharded.packed_bytes(nq, k)
        assert nb % 16 == 0 and 0 <= nb - nq * k * 12 - nq * 4 < 16
        buf = torch.zeros(nb, dtype=torch.uint8)
        ids, sims, status = sharded.unpack(buf, nq, k)
        ids.copy_(torch.arange(nq * k, dtype=torch.int64).view(nq, k))
        sims.fill_(1.5)
        status.fill_(7)
        ids2, sims2, status2 = sharded.unpack(buf, nq, k)            # views of the same bytes
        assert ids2[-1, -1].item() == nq * k - 1 and sims2[0, 0].item() == 1.5 and status2[-1].item() == 7 and status2.numel() == nq
    packing()

    @settings(max_examples=50, deadline=None)
    @given(st.integers(2, 5), st.integers(1, 6), st.integers(1, 12), st.integers(0, 2**31 - 1))
    def merge_is_order_independent(g, nq, k, seed):
        rng = np.random.default_rng(seed)
        # g shards with disjoint id ranges, each list sorted (score desc, id asc) as the local search leaves it
        ids = np.empty((g, nq, k), np.int64)
        sims = np.empty((g, nq, k), np.float32)
        for p in range(g):
            for j in range(nq):
                i = np.sort(rng.choice(1000, size=k, replace=False)) + 1000 * p
                s = rng.integers(0, 4, size=k).astype(np.float32)    # heavy ties on purpose
                order = np.lexsort((i, -s))
                ids[p, j], sims[p, j] = i[order], s[order]
        a_i, a_s = oracle.merge_parts(ids, sims, k)
        perm = rng.permutation(g)
        b_i, b_s = oracle.merge_parts(ids[perm], sims[perm], k)
        np.testing.assert_array_equal(a_i, b_i)
        np.testing.assert_array_equal(a_s, b_s)
        for j in range(nq):                                            # and it is the top-k of the union
            flat_i, flat_s = ids[:, j].reshape(-1), sims[:, j].reshape(-1)
            order = np.lexsort((flat_i, -flat_s))[:k]
            np.testing.assert_array_equal(a_i[j], flat_i[order])
    merge_is_order_independent()


def _selfknn_worker(rank, world, port, n, d, k, block, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sharded = importlib.import_module(PKG + ".sharded")
    synth = importlib.import_module(PKG + ".synth")
    oracle = importlib.import_module("oracle.oracle")
    vecs, _ = synth.ties(n, 1, d=d, n_distinct=n // 3)          # every row has two exact duplicates: the self-first rule matters
    bounds = sharded.shard_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    shard = np.ascontiguousarray(vecs[:, lo:hi])

    def local_search(queries, kk):
        ids, sims = oracle.topk_ip(shard, queries.numpy().T, kk)
        packed = torch.zeros((sharded.packed_bytes(ids.shape[0], kk),), dtype=torch.uint8)
        pi, ps, _ = sharded.unpack(packed, ids.shape[0], kk)
        pi.copy_(torch.from_numpy(ids + lo)); ps.copy_(torch.from_numpy(sims))
        return packed

    def merge(packed_all, world_, nq_, kk):
        pb = sharded.packed_bytes(nq_, kk)
        parts = [sharded.unpack(packed_all[g * pb:(g + 1) * pb], nq_, kk) for g in range(world_)]
        i, s = oracle.merge_parts(np.stack([p[0].numpy() for p in parts]), np.stack([p[1].numpy() for p in parts]), kk)
        return torch.from_numpy(i), torch.from_numpy(s), torch.zeros(nq_, dtype=torch.int32)

    searcher = sharded.ShardedSearcher(local_search, merge)
    rows_local = torch.from_numpy(np.ascontiguousarray(shard.T))
    sims, ids = sharded.self_knn_rowsharded(searcher, rows_local, bounds, k, rank, block=block)
    np.savez(os.path.join(out_dir, f"self{rank}.npz"), ids=ids.numpy(), sims=sims.numpy(), lo=lo, hi=hi)
    dist.destroy_process_group()


def test_rowsharded_self_knn(tmp_path, synth, oracle):
    """The N x N self-kNN over a row-sharded database: blocks of query rows broadcast by their owner, searched by every
    shard, merged, kept by the owner; a row's own id first even among exact duplicates (diffusion.py:108)."""
    n, d, k, world = 150, 24, 7, 2
    port = 31000 + (os.getpid() % 2000)
    mp.spawn(_selfknn_worker, args=(world, port, n, d, k, 32, str(tmp_path)), nprocs=world, join=True)
    vecs, _ = synth.ties(n, 1, d=d, n_distinct=n // 3)
    ref_s, ref_i = oracle.knn_search(np.ascontiguousarray(vecs.T), np.ascontiguousarray(vecs.T), k)
    for r in range(world):
        o = np.load(os.path.join(tmp_path, f"self{r}.npz"))
        lo, hi = int(o["lo"]), int(o["hi"])
        assert (o["ids"][:, 0] == np.arange(lo, hi)).all()
        np.testing.assert_allclose(o["sims"], ref_s[lo:hi], rtol=1e-6, atol=1e-7)
        # same neighbour SETS up to exact ties (the duplicates): compare by score multiset and by membership of non-tied ids
        for j in range(hi - lo):
            assert len(set(o["ids"][j].tolist())) == k
