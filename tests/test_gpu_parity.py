"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures.

Bar: identical top-K id lists except where the adjudicating float64 scores are within a few fp32
ulp (rtol 1e-6), scores within 1e-5 relative of the oracle's fp32 scores, identical mAP.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PATHS = {"scan": 1, "gemm": 2, "exact": 3}


def _check_lists(oracle, ids, ref_ids, s64, what):
    for j in range(ids.shape[0]):
        ok, msg = oracle.compare_topk(ids[j], ref_ids[j], lambda i, j=j: s64[i, j])
        assert ok, f"{what}, query {j}: {msg}"


@pytest.fixture(scope="module")
def cfg1(synth, oracle):
    vecs, qvecs = synth.gaussian(4993, 70, d=2048)
    ref_ids, ref_sims = oracle.topk_ip(vecs, qvecs, 100)
    return vecs, qvecs, ref_ids, ref_sims, oracle.scores_f64(vecs, qvecs)


@pytest.mark.parametrize("path", ["gemm", "scan", "exact"])
def test_cfg1_every_path(pkg, oracle, cfg1, path):
    vecs, qvecs, ref_ids, ref_sims, s64 = cfg1
    with pkg.ExactIndex(vecs.T) as ix:
        ix.set_param("force_path", PATHS[path])
        nq = 70 if path != "scan" else 6
        ids, sims = ix.search(qvecs.T[:nq], 100)
        st = ix.stats()
    assert st["path"] == PATHS[path] and st["gpu_launches"] > 0
    _check_lists(oracle, ids, ref_ids[:nq], s64, path)
    np.testing.assert_allclose(sims, ref_sims[:nq], rtol=1e-5, atol=1e-7)
    assert ids.dtype == np.int64 and sims.dtype == np.float32


def test_default_dispatch(pkg, cfg1):
    vecs, qvecs, *_ = cfg1
    with pkg.ExactIndex(vecs.T) as ix:
        ix.search(qvecs.T[:1], 100)
        assert ix.stats()["path"] == 1          # batch 1 -> HBM scan
        ix.search(qvecs.T, 100)
        assert ix.stats()["path"] == 2          # batch 70 -> tcgen05 GEMM + fused top-K
        assert ix.stats()["n_exact_rerun"] == 0


@pytest.mark.parametrize("fam", ["G", "P"])
def test_golden_top100(pkg, synth, oracle, golden, fam):
    v, q = synth.gaussian(1000, 5, d=2048, family=fam)
    s64 = oracle.scores_f64(v, q)
    for path in ("gemm", "scan"):
        with pkg.ExactIndex(v.T) as ix:
            ix.set_param("force_path", PATHS[path])
            ids, sims = ix.search(q.T, 100)
        _check_lists(oracle, ids, golden[f"C_{fam}_top100"].T.astype(np.int64), s64, f"golden {fam} {path}")
        np.testing.assert_allclose(sims, golden[f"C_{fam}_top100_scores"].T, rtol=1e-5, atol=1e-7)
        for j in range(5):       # matching_L2's distance order names the same sets
            assert set(ids[j]) == set(golden[f"C_{fam}_matching_L2_idx"][j])


def test_matching_L2_api(pkg, synth, golden):
    vecs, qvecs = synth.gaussian(512, 8, d=64)
    idx, tpq = pkg.matching_L2(10, vecs.T, qvecs.T)
    assert idx.shape == (8, 10) and idx.dtype == np.int64 and tpq > 0
    np.testing.assert_array_equal(idx, golden["A_matching_L2_idx"])
    # float64, un-normalised rows: the function normalises (nnsearch.py:693-698)
    vecs64 = vecs.astype(np.float64) * np.linspace(0.5, 2.0, vecs.shape[1])[None, :]
    q64 = qvecs.astype(np.float64) * 3.0
    idx, _ = pkg.matching_L2(10, vecs64.T, q64.T)
    np.testing.assert_array_equal(idx, golden["B_matching_L2_idx"])
    # cached index is reused for the same array, rebuilt after an in-place edit
    a = pkg.cached_index(vecs64.T, True)
    assert pkg.cached_index(vecs64.T, True) is a
    pkg.clear_index_cache()


def test_rank_ip_topk_and_map(pkg, synth, oracle, golden):
    v, q, gnd = synth.clustered(3000, 12, d=256, n_clusters=40, noise=1.6, spread=0.5)
    ranks = pkg.rank_ip(v, q, K=100)
    assert ranks.shape == (100, 12) and ranks.dtype == np.int64
    s64 = oracle.scores_f64(v, q)
    _check_lists(oracle, ranks.T, golden["D_top100"].T.astype(np.int64), s64, "rank_ip")
    m100, aps100, _, _ = oracle.compute_map(ranks, gnd, [1, 5, 10])
    assert m100 == golden["D_map_top100"]
    np.testing.assert_array_equal(aps100, golden["D_aps_top100"])
    pkg.clear_index_cache()


def test_knn_wrapper(pkg, synth, oracle):
    v, q = synth.gaussian(2000, 9, d=128)
    knn = pkg.KNN(v.T, "cosine")
    assert (knn.N, knn.D) == (2000, 128) and knn.database.flags["C_CONTIGUOUS"]
    sims, ids = knn.search(q.T, 7)
    rs, ri = oracle.knn_search(v.T, q.T, 7, "cosine")
    s64 = oracle.scores_f64(v, q)
    _check_lists(oracle, ids, ri, s64, "KNN cosine")
    np.testing.assert_allclose(sims, rs, rtol=1e-5, atol=1e-7)
    knn2 = pkg.KNN(v.T * 1.5, "euclidean")
    dist, ids2 = knn2.search(q.T, 7)
    rd, ri2 = oracle.knn_search(v.T * 1.5, q.T, 7, "euclidean")
    d64 = -(((v.T * 1.5).astype(np.float64)[:, None, :] - q.T.astype(np.float64)[None, :, :]) ** 2).sum(-1)
    _check_lists(oracle, ids2, ri2, d64, "KNN euclidean")
    np.testing.assert_allclose(dist, rd, rtol=1e-4, atol=1e-5)
    with pytest.raises(KeyError):
        pkg.KNN(v.T, "manhattan")


def test_self_knn_self_first(pkg, synth, oracle):
    v, _ = synth.gaussian(1500, 1, d=256)
    knn = pkg.KNN(v.T, "cosine")
    sims, ids = knn.self_search(20)
    assert (ids[:, 0] == np.arange(1500)).all()
    rs, ri = oracle.knn_search(v.T, v.T, 20, "cosine")
    s64 = oracle.scores_f64(v, v)
    _check_lists(oracle, ids, ri, s64, "self-kNN")
    np.testing.assert_allclose(sims, rs, rtol=1e-5, atol=1e-6)


def test_ties_and_duplicates(pkg, synth, oracle, golden):
    v, q = synth.ties(256, 4, d=64, n_distinct=16)
    for path in ("gemm", "scan", "exact"):
        with pkg.ExactIndex(v.T) as ix:
            ix.set_param("force_path", PATHS[path])
            ids, sims = ix.search(q.T, 40)
        rid, rsim = oracle.topk_ip(v, q, 40)
        # duplicated rows give exactly equal exact scores -> the documented rule (ascending id) decides
        np.testing.assert_array_equal(ids, rid, err_msg=path)
        np.testing.assert_allclose(sims, golden["E_sorted_scores"][:40].T, rtol=1e-5, atol=1e-7)


def test_edges(pkg, synth, oracle):
    v, q = synth.gaussian(300, 3, d=100)              # d not a multiple of 64, N not of 256
    s64 = oracle.scores_f64(v, q)
    with pkg.ExactIndex(v.T) as ix:
        for path in ("gemm", "scan", "exact"):
            ix.set_param("force_path", PATHS[path])
            for k in (1, 7, 300):
                if k * 4 > 300 and path != "exact":
                    continue
                ids, _ = ix.search(q.T, k)
                rid, _ = oracle.topk_ip(v, q, k)
                _check_lists(oracle, ids, rid, s64, f"edge {path} k={k}")
        ix.set_param("force_path", 0)
        ids, _ = ix.search(q.T, 300)                  # k == N -> exact path by dispatch
        assert sorted(ids[0].tolist()) == list(range(300))
        with pytest.raises(ValueError):
            ix.search(q.T, 301)
        with pytest.raises(ValueError):
            ix.search(q.T[:, :50], 5)
    # row-major input and id_offset (a shard's view)
    with pkg.ExactIndex(np.ascontiguousarray(v.T), id_offset=1000) as ix:
        ids, _ = ix.search(np.ascontiguousarray(q.T), 5)
        rid, _ = oracle.topk_ip(v, q, 5)
        np.testing.assert_array_equal(ids, rid + 1000)


def test_merge_kernel_matches_host_merge(pkg, synth, oracle):
    """xs_merge_candidates (the post-all-gather G*k -> k merge) against its host restatement,
    on per-shard lists produced by the CUDA path with global ids."""
    import importlib
    import torch
    sharded = importlib.import_module(pkg.__name__ + ".sharded")
    v, q = synth.gaussian(3000, 6, d=128)
    k, world = 10, 3
    bounds = sharded.shard_bounds(3000, world)
    ids_parts, sims_parts, shards = [], [], []
    for g in range(world):
        ix = pkg.ExactIndex(v.T[bounds[g]:bounds[g + 1]], id_offset=bounds[g])
        shards.append(ix)
        i, s = ix.search(q.T, k)
        ids_parts.append(i); sims_parts.append(s)
    ids_all = torch.from_numpy(np.stack(ids_parts)).cuda()
    sims_all = torch.from_numpy(np.stack(sims_parts)).cuda()
    cs = sharded.CudaShard(shards[0], 0)
    mi, ms = cs.merge_lists(ids_all, sims_all, k)
    # and the packed single-gather layout: [ids | sims] bytes per part, back to back
    pb = sharded.packed_bytes(6, k)
    packed_all = torch.zeros((world * pb,), dtype=torch.uint8, device="cuda")
    for g in range(world):
        gi, gs, gst = sharded.unpack(packed_all[g * pb:(g + 1) * pb], 6, k)
        gi.copy_(ids_all[g]); gs.copy_(sims_all[g])
        if g == 1:
            gst[4] = 1                                  # shard 1 could not certify query 4
    pi, ps, pst = cs.merge(packed_all, world, 6, k)
    torch.cuda.synchronize()
    assert torch.equal(pi, mi) and torch.equal(ps, ms)
    assert pst.cpu().tolist() == [0, 0, 0, 0, 1, 0]    # the merged certificate word is the OR over the shards
    hi, hs = oracle.merge_parts(np.stack(ids_parts), np.stack(sims_parts), k)
    np.testing.assert_array_equal(mi.cpu().numpy(), hi)
    np.testing.assert_array_equal(ms.cpu().numpy(), hs)
    rid, _ = oracle.topk_ip(v, q, k)
    s64 = oracle.scores_f64(v, q)
    _check_lists(oracle, hi, rid, s64, "sharded merge")
    for ix in shards:
        ix.close()


def test_rank_all_full_ranking(pkg, synth, oracle, golden):
    """K == N: the full `ranks` array of main_retrieve.py:176, plus mAP identical to the reference's."""
    v, q, gnd = synth.clustered(3000, 12, d=256, n_clusters=40, noise=1.6, spread=0.5)
    ranks, sc = pkg.rank_ip(v, q, K=None, return_scores=True)
    assert ranks.shape == (3000, 12) and ranks.dtype == np.int64 and sc.shape == (3000, 12)
    for j in range(12):
        assert sorted(ranks[:, j].tolist()) == list(range(3000))
    assert (np.diff(sc, axis=0) <= 0).all()
    s64 = oracle.scores_f64(v, q)
    _, ref_ranks = oracle.rank_ip(v, q)
    _check_lists(oracle, ranks.T, ref_ranks.T, s64, "rank_all")
    # Over a FULL ranking of 3000 rows a few fp32 near-ties (score gaps < 1e-7) between a positive and a
    # negative are ordered differently by OpenBLAS' summation and by the exact rescoring; each such swap
    # moves one AP term by ~1/(rank * n_pos).  (The top-100 test above is bit-identical.)
    m, aps, pr, prs = oracle.compute_map(ranks, gnd, [1, 5, 10])
    assert abs(m - golden["D_map"]) < 1e-6
    np.testing.assert_allclose(aps, golden["D_aps"], rtol=0, atol=1e-5)
    np.testing.assert_array_equal(prs, golden["D_prs"])
    e, mm, h = oracle.protocol_maps(ranks, gnd)
    assert abs(e - golden["D_mapE"]) < 1e-6 and abs(mm - golden["D_mapM"]) < 1e-6 and abs(h - golden["D_mapH"]) < 1e-6
    # matching_L2 with K == N (mAP mode of test_rOP1m.py:147-148) goes through the same sort
    idx, _ = pkg.matching_L2(3000, v.T, q.T)
    assert idx.shape == (12, 3000)
    _check_lists(oracle, idx, ref_ranks.T, s64, "matching_L2 K=N")
    # ties: equal scores come out in ascending id order
    vt, qt = synth.ties(256, 4, d=64, n_distinct=16)
    rt = pkg.rank_ip(vt, qt)
    rid, _ = oracle.topk_ip(vt, qt, 256)
    np.testing.assert_array_equal(rt.T, rid)
    pkg.clear_index_cache()


@pytest.mark.parametrize("nq,n", [(300, 9000), (129, 3000)])
def test_pair_mode_large_batches(pkg, synth, oracle, nq, n):
    """More than 128 queries -> the cta_group::2 (CTA-pair, 256 x 256 tile) shape of the GEMM kernel;
    odd numbers of query tiles leave the peer CTA of the last pair without queries."""
    v, q = synth.gaussian(n, nq, d=512)
    rid, rs = oracle.topk_ip(v, q, 50)
    s64 = oracle.scores_f64(v, q)
    with pkg.ExactIndex(v.T) as ix:
        ix.set_param("force_path", 2)
        ids, sims = ix.search(q.T, 50)
        assert ix.stats()["n_exact_rerun"] == 0
        _check_lists(oracle, ids, rid, s64, "pair mode")
        np.testing.assert_allclose(sims, rs, rtol=1e-5, atol=1e-7)
        ix.set_param("pair_mode", 0)                    # same batch through the single-CTA shape
        ids1, sims1 = ix.search(q.T, 50)
        np.testing.assert_array_equal(ids1, ids)
        np.testing.assert_array_equal(sims1, sims)


def test_qge1_aqe_second_pass(pkg, synth, oracle, golden):
    """AQE re-score (Reranking.py:287-306) built and searched on the device vs the reference's own output."""
    vecs, qvecs = synth.gaussian(512, 8, d=64)
    first = pkg.rank_ip(vecs, qvecs, K=10)
    np.testing.assert_array_equal(first, golden["A_ranks"][:10])
    r10 = pkg.qge1(first, qvecs, vecs, 10)
    assert r10.shape == (10, 8) and r10.dtype == np.int64
    np.testing.assert_array_equal(r10, golden["A_qge1_ranks"][:10])
    full = pkg.qge1(first, qvecs, vecs, 10, full=True)
    assert full.shape == (512, 8)
    qe, _ = oracle.feature_enhancement(1, 3, golden["A_ranks"][:10], qvecs, vecs, 4.0)
    s64 = np.dot(vecs.T.astype(np.float64), qe)
    _check_lists(oracle, full.T, golden["A_qge1_ranks"].T, s64, "qge1 full ranking")
    qe_gpu, _ = pkg.feature_enhancement(1, 3, first, qvecs, vecs, 4.0, K=5)
    np.testing.assert_allclose(qe_gpu, qe, rtol=2e-6, atol=1e-7)
    pkg.clear_index_cache()


@pytest.mark.parametrize("k", [1000, 2048, 3000])
def test_large_k(pkg, synth, oracle, k):
    """k up to 2048 stays on the coarse paths (larger pools); above that the dispatch goes exact."""
    v, q = synth.gaussian(20000, 5, d=256)
    rid, rs = oracle.topk_ip(v, q, k)
    s64 = oracle.scores_f64(v, q)
    with pkg.ExactIndex(v.T) as ix:
        ids, sims = ix.search(q.T, k)
        assert ix.stats()["path"] == (2 if k <= 2048 else 3)
        _check_lists(oracle, ids, rid, s64, f"k={k}")
        np.testing.assert_allclose(sims, rs, rtol=1e-5, atol=1e-7)
        if k <= 2048:
            i1, s1 = ix.search(q.T[:1], k)          # batch 1 -> scan path with the same k
            assert ix.stats()["path"] == 1
            _check_lists(oracle, i1, rid[:1], s64, f"scan k={k}")


def test_crowded_scores_fall_back_to_exact(pkg, synth, oracle):
    """Tight clusters: hundreds of rows within the bf16 error band of the k-th score.  Queries the
    coarse pass cannot certify are re-run on the exact path -- results stay exact, reruns are counted."""
    v, q, _ = synth.clustered(20000, 16, d=256, n_clusters=10, noise=0.02)
    rid, rs = oracle.topk_ip(v, q, 100)
    s64 = oracle.scores_f64(v, q)
    with pkg.ExactIndex(v.T) as ix:
        for path in (2, 1):
            ix.set_param("force_path", path)
            ids, sims = ix.search(q.T, 100)
            st = ix.stats()
            _check_lists(oracle, ids, rid, s64, f"crowded path {path}")
            np.testing.assert_allclose(sims, rs, rtol=1e-5, atol=1e-7)
            assert st["n_exact_rerun"] > 0, st       # this family is built to defeat the coarse filter


def test_threads_share_one_index(pkg, synth, oracle):
    """Flask serves requests from threads (online.py:163): concurrent searches on one index."""
    import threading
    v, q = synth.gaussian(6000, 32, d=128)
    rid, _ = oracle.topk_ip(v, q, 20)
    s64 = oracle.scores_f64(v, q)
    out = {}
    with pkg.ExactIndex(v.T) as ix:
        def work(t):
            for rep in range(5):
                sl = slice(t * 8, t * 8 + 8) if rep % 2 == 0 else slice(t * 8, t * 8 + 1)
                ids, _ = ix.search(q.T[sl], 20)
                out[(t, rep)] = (sl, ids)
        ths = [threading.Thread(target=work, args=(t,)) for t in range(4)]
        [t.start() for t in ths]
        [t.join() for t in ths]
    assert len(out) == 20
    for (t, rep), (sl, ids) in out.items():
        _check_lists(oracle, ids, rid[sl], s64[:, sl], f"thread {t} rep {rep}")


def test_argument_errors_and_dtypes(pkg, synth, oracle):
    v, q = synth.gaussian(500, 4, d=64)
    rid, _ = oracle.topk_ip(v, q, 5)
    with pkg.ExactIndex(v.T.astype(np.float16)) as ix:          # fp16 in -> cast to fp32 on the way
        ids, _ = ix.search(q.T.astype(np.float16), 5)
        assert ids.shape == (4, 5)
    with pkg.ExactIndex(v.T) as ix:
        e_ids, e_sims = ix.search(np.empty((0, 64), np.float32), 5)
        assert e_ids.shape == (0, 5) and e_sims.shape == (0, 5)
        for bad_k in (0, -1, 501):
            with pytest.raises(ValueError):
                ix.search(q.T, bad_k)
        with pytest.raises(ValueError):
            ix.search(q.T[0], 5)                                   # 1-D
        with pytest.raises(ValueError):
            ix.set_param("no_such_knob", 1)
        ids, _ = ix.search(q.T.astype(np.float64), 5)              # fp64 queries
        np.testing.assert_array_equal(ids, rid)
        ix.close()
        with pytest.raises((ValueError, RuntimeError)):
            ix.search(q.T, 5)                                      # closed index
    with pytest.raises(ValueError):
        pkg.matching_L2(600, v.T, q.T)                             # K > N


def test_diffusion_graph(pkg, synth, oracle, golden):
    """Mutual-kNN affinity + Laplacian (diffusion.py:87-116) from the GPU kNN lists vs the reference's own output."""
    ids, sims = golden["F_knn_ids"].astype(np.int64), golden["F_knn_sims"]
    np.testing.assert_array_equal(pkg.diffusion.mutual_mask(ids), oracle.mutual_mask(ids))
    aff = pkg.diffusion.get_affinity(sims.copy(), ids).toarray()
    # same sparsity pattern; values within 2 ulp: the device cubes exactly and rounds once, the reference's `sims ** 3` is
    # numpy's float32 pow, whose last bit depends on the host's vector math library (the golden file has its AVX-512 bits)
    np.testing.assert_array_equal(aff != 0, golden["F_affinity"] != 0)
    np.testing.assert_allclose(aff, golden["F_affinity"], rtol=2.4e-7, atol=0)
    lap = pkg.diffusion.get_laplacian(sims.copy(), ids)
    lap_d = np.asarray(lap.toarray(), dtype=np.float32)
    np.testing.assert_allclose(lap_d, golden["F_laplacian"], rtol=1e-6, atol=1e-7)
    # end to end: the kNN lists themselves from the GPU self search
    v, _ = synth.clustered(400, 1, d=64, n_clusters=12, noise=0.9)[:2]
    s2, i2, lap2 = pkg.diffusion.knn_graph(v.T, n_trunc=12, kd=12)
    assert (i2[:, 0] == np.arange(400)).all()
    s64 = oracle.scores_f64(v, v)
    _check_lists(oracle, i2, ids, s64, "knn_graph ids")
    assert lap2.shape == (400, 400)


def test_diffusion_offline_cg(pkg, synth, oracle, golden):
    """Gallery-side truncated CG (diffusion.py:15-19, 74-76): the CUDA solver vs the reference's own output
    (case G), vs the oracle on a second graph, and the Diffusion class end to end."""
    import scipy.sparse as sparse
    lap = sparse.csr_matrix(golden["F_laplacian"])
    ids = golden["G_trunc_ids"].astype(np.int64)
    got = pkg.diffusion.offline_cg(lap, ids)
    assert got.dtype == np.float32 and got.shape == ids.shape
    np.testing.assert_allclose(got, golden["G_offline"], rtol=2e-6, atol=1e-8)
    # fewer steps / early stop / a subset of rows in another order
    np.testing.assert_allclose(pkg.diffusion.offline_cg(lap, ids[::-7], maxiter=3),
                               oracle.offline_scores(lap, ids[::-7], maxiter=3), rtol=2e-6, atol=1e-8)
    np.testing.assert_allclose(pkg.diffusion.offline_cg(lap, ids[:50], tol=1e-2),
                               oracle.offline_scores(lap, ids[:50], tol=1e-2), rtol=2e-6, atol=1e-8)
    np.testing.assert_array_equal(pkg.diffusion.offline_cg(lap, ids[:5], maxiter=0), np.zeros((5, 40), np.float32))
    # a denser, larger problem: truncation sets of 300 out of 3000 rows, 30 neighbours in the graph
    v, _ = synth.clustered(3000, 1, d=64, n_clusters=40, noise=0.8)[:2]
    d = pkg.diffusion.Diffusion(v.T, None)
    sims, tids = d.knn.self_search(300)
    lap2 = d.get_laplacian(sims[:, :30].copy(), tids[:, :30])
    pick = np.arange(0, 3000, 97)
    np.testing.assert_allclose(pkg.diffusion.offline_cg(lap2, tids[pick]), oracle.offline_scores(lap2, tids[pick]),
                               rtol=5e-6, atol=1e-8)
    offline = d.get_offline_results(300, 30)                 # self-kNN -> graph -> CG on the device end to end
    assert offline.shape == (3000, 3000) and offline.dtype == np.float32
    oi, osims, osc = pkg.diffusion.offline_device(d.knn.index, 300, 30, return_sims=True)
    np.testing.assert_array_equal(oi, tids)
    np.testing.assert_array_equal(osims, sims)
    np.testing.assert_allclose(osc[pick], oracle.offline_scores(lap2, tids[pick]), rtol=5e-6, atol=1e-8)
    np.testing.assert_allclose(np.asarray(offline[pick[3]].todense()).reshape(-1)[tids[pick[3]]],
                               oracle.offline_scores(lap2, tids[pick[3:4]])[0], rtol=5e-6, atol=1e-8)
    qs, qi = d.knn.search(v.T[:4], 3)
    ts, tr = pkg.diffusion.search_offline(offline, qs, qi, 300)
    for i in range(4):                                         # Reranking.py:251-255 on a dense row
        dense = (qs[i].astype(np.float32) ** 3) @ offline[qi[i]].toarray()
        np.testing.assert_allclose(ts[i], np.sort(dense)[::-1][:300], rtol=1e-6)
        np.testing.assert_allclose(dense[tr[i]], ts[i], rtol=1e-6)
        assert len(set(tr[i].tolist())) == 300
    with pytest.raises(ValueError):
        bad = ids[:3].copy()
        bad[1, 5] = 400
        pkg.diffusion.offline_cg(lap, bad)
    with pytest.raises(ValueError):
        pkg.diffusion.offline_cg(lap, np.zeros((2, 5000), np.int64))


def test_query_expansion_and_database_augmentation(pkg, synth, oracle, golden):
    """average_query_expansion / database_augmentation (Reranking.py:314-365, 375-440) through top-k searches
    instead of N x N argsorts, vs the reference's own ranks (case H); initial_rank = batch_torch_topk (:487-511)."""
    v, q, _ = synth.clustered(600, 8, d=64, n_clusters=20, noise=1.2, spread=0.5)
    for name in ("average_query_expansion", "database_augmentation"):
        got = getattr(pkg, name)(q, v, 20, "synthetic", None)
        assert got.shape == (20, 8) and got.dtype == np.int64
        _, v_aug, q_aug = getattr(oracle, name)(q, v, 20)
        s64 = oracle.scores_f64((v_aug / np.linalg.norm(v_aug, axis=1, keepdims=True)).T,
                                (q_aug / np.linalg.norm(q_aug, axis=1, keepdims=True)).T)
        _check_lists(oracle, got.T, golden[f"H_{name}_ranks"].T.astype(np.int64), s64, name)
    feat = np.ascontiguousarray(np.concatenate([q, v], axis=1).T)
    got = pkg.initial_rank(feat, 7)
    dist = 2 - 2 * feat.astype(np.float64) @ feat.astype(np.float64).T
    dist = (dist / dist.max(axis=0)).T                                    # the reference's per-row rescale
    want = np.argsort(dist, axis=1, kind="stable")[:, :7]
    assert (got[:, 0] == np.arange(608)).all()
    _check_lists(oracle, got[:, 1:], want[:, 1:], -dist.T, "initial_rank")
    import torch
    np.testing.assert_array_equal(pkg.initial_rank(torch.from_numpy(feat), 7), got)


def test_two_lanes_pipelined_and_clone_lifetime(pkg, synth, oracle):
    """xs_index_clone: a second lane over the same database arrays.  Pipelined searches alternate between the index
    and its clone on two streams; results equal the plain search; the clone outlives its parent."""
    import importlib
    import torch
    sharded = importlib.import_module(pkg.__name__ + ".sharded")
    v, q = synth.gaussian(30000, 96, d=256)
    index = pkg.ExactIndex(v.T)
    batches = [np.ascontiguousarray(q.T[i:i + 32]) for i in (0, 32, 64)]
    want = [index.search(b, 20) for b in batches]
    shard = sharded.CudaShard(index, 0, lanes=2)
    searcher = sharded.ShardedSearcher(shard.local_search, shard.merge, lane_stream=shard.lane_stream)
    qd = [torch.from_numpy(b).cuda() for b in batches]
    pending, got = None, []
    for it in range(30):
        nxt = (it % 3, searcher.search_async(qd[it % 3], 20))
        if pending is not None:
            ids, sims = pending[1].result()
            got.append((pending[0], ids.cpu().numpy().copy(), sims.cpu().numpy().copy()))
        pending = nxt
    ids, sims = pending[1].result()
    got.append((pending[0], ids.cpu().numpy().copy(), sims.cpu().numpy().copy()))
    assert len(got) == 30
    for b, ids, sims in got:
        np.testing.assert_array_equal(ids, want[b][0])
        np.testing.assert_array_equal(sims, want[b][1])
    ids, sims = searcher.search(qd[1], 20)                            # blocking form right after pipelined ones
    np.testing.assert_array_equal(ids.cpu().numpy(), want[1][0])
    clone = shard.lanes[1]
    assert clone.device_bytes == 0 and index.device_bytes > 0          # the clone owns workspaces only
    index.close()                                                      # the database arrays live on with the clone
    ids2, sims2 = clone.search(batches[2], 20)
    np.testing.assert_array_equal(ids2, want[2][0])
    shard.close()


def test_self_knn_pipelined_batches_and_reruns(pkg, synth, oracle):
    """More rows than one 8192-row batch (two-deep pipeline) and duplicated rows (certificate fails ->
    exact re-run inside the pipeline): every row still gets its own id first and exact neighbours."""
    v, _ = synth.gaussian(20000, 1, d=64)
    ix = pkg.ExactIndex(v.T)
    sims, ids = ix.self_knn(10)
    assert (ids[:, 0] == np.arange(20000)).all()
    pick = np.array([0, 1, 8191, 8192, 8193, 16383, 16384, 19999])
    rs, ri = oracle.knn_search(v.T, v.T[pick], 10, "cosine")
    s64 = oracle.scores_f64(v, v[:, pick])
    _check_lists(oracle, ids[pick], ri, s64, "self-kNN pipelined")
    sub_s, sub_i = ix.self_knn(10, 9000, 9100)              # a row range
    np.testing.assert_array_equal(sub_i, ids[9000:9100])
    ix.set_param("self_lanes", 2)                           # batches alternate between the index and an internal clone
    sims2, ids2 = ix.self_knn(10)
    np.testing.assert_array_equal(ids2, ids)
    np.testing.assert_array_equal(sims2, sims)
    ix.close()
    vt, _ = synth.ties(9000, 1, d=64, n_distinct=300)       # every row has 29 exact duplicates
    ix = pkg.ExactIndex(vt.T)
    ix.set_param("self_lanes", 2)                           # the exact re-run then happens on the lane that owns the batch
    sims, ids = ix.self_knn(8)
    st = ix.stats()
    assert (ids[:, 0] == np.arange(9000)).all()
    # the 7 other neighbours are duplicates of the row (score 1), lowest ids first
    for r in (0, 299, 300, 4567, 8999):
        dup = np.arange(r % 300, 9000, 300)
        expect = [r] + [d for d in dup if d != r][:7]
        assert ids[r].tolist() == expect, (r, ids[r], expect)
    ix.close()


def test_rank_ip_torch_device_tensors(pkg, synth, oracle):
    """Device-resident variant for the torch.mm + torch.sort call sites (traindataset.py:221-222)."""
    import torch
    v, q = synth.gaussian(5000, 40, d=256)
    tv, tq = torch.from_numpy(v).cuda(), torch.from_numpy(q).cuda()
    scores, ranks = pkg.rank_ip_torch(tv, tq, 25)
    assert ranks.shape == (25, 40) and ranks.is_cuda and scores.dtype == torch.float32
    ref_scores, ref_ranks = torch.sort(torch.mm(tv.t(), tq), dim=0, descending=True)
    s64 = oracle.scores_f64(v, q)
    _check_lists(oracle, ranks.t().cpu().numpy(), ref_ranks[:25].t().cpu().numpy(), s64, "rank_ip_torch")
    torch.testing.assert_close(scores, ref_scores[:25], rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        pkg.rank_ip_torch(tv.cpu(), tq.cpu(), 5)


def test_rank_all_many_queries_column_blocks(pkg, synth, oracle):
    """More than 128 queries: xs_rank_all streams the [N, nq] result back in column blocks."""
    v, q = synth.gaussian(700, 150, d=64)
    with pkg.ExactIndex(v.T) as ix:
        ranks, sc = ix.rank_all(q.T, return_scores=True)
    assert ranks.shape == (700, 150)
    _, ref = oracle.rank_ip(v, q)
    s64 = oracle.scores_f64(v, q)
    _check_lists(oracle, ranks.T, ref.T, s64, "rank_all 150 queries")
    assert (np.diff(sc, axis=0) <= 0).all()


def test_device_api_reruns_uncertified_itself(pkg, synth, oracle):
    """xs_search_dev without a status buffer: duplicated rows defeat the certificate, the call re-runs
    those queries exactly before returning (one stream synchronisation)."""
    import torch
    v, q = synth.ties(4096, 6, d=64, n_distinct=64)
    scores, ranks = pkg.rank_ip_torch(torch.from_numpy(v).cuda(), torch.from_numpy(q).cuda(), 70)
    rid, rs = oracle.topk_ip(v, q, 70)
    np.testing.assert_array_equal(ranks.t().cpu().numpy(), rid)
    np.testing.assert_allclose(scores.t().cpu().numpy(), rs, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("n,nq,d,k", [(1, 1, 8, 1), (5, 3, 1, 5), (255, 2, 7, 3), (257, 130, 65, 9), (600, 4, 3000, 10)])
def test_extreme_shapes(pkg, oracle, n, nq, d, k):
    """Degenerate and awkward sizes: a single row, d = 1, sizes straddling the 256-row / 64-column padding,
    descriptors longer than the 2048-wide fast paths."""
    rng = np.random.default_rng(n * 1000 + d)
    v = rng.standard_normal((d, n)).astype(np.float32)
    q = rng.standard_normal((d, nq)).astype(np.float32)
    rid, rs = oracle.topk_ip(v, q, k)
    s64 = oracle.scores_f64(v, q)
    with pkg.ExactIndex(v.T) as ix:
        for path in (0, 1, 2, 3):
            ix.set_param("force_path", path)
            ids, sims = ix.search(q.T, k)
            _check_lists(oracle, ids, rid, s64, f"shape {(n, nq, d, k)} path {path}")
            np.testing.assert_allclose(sims, rs, rtol=2e-5, atol=1e-6)


def test_many_queries_multiple_batches(pkg, synth, oracle):
    """More queries than one internal batch (8192) through the host call."""
    v, q = synth.gaussian(3000, 9000, d=64)
    with pkg.ExactIndex(v.T) as ix:
        ids, sims = ix.search(q.T, 5)
    pick = np.array([0, 127, 128, 8191, 8192, 8999])
    rid, rs = oracle.topk_ip(v, q[:, pick], 5)
    s64 = oracle.scores_f64(v, q[:, pick])
    _check_lists(oracle, ids[pick], rid, s64, "9000 queries")
    np.testing.assert_allclose(sims[pick], rs, rtol=1e-5, atol=1e-7)


def test_random_shape_sweep():
    """24 seeded random (n, d, nq, k, data family, renormalise) cases x every path against the oracle
    (tests/fuzz_gpu.py; 280 further cases were run clean during development)."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "fuzz_gpu.py"), "24", "7"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
