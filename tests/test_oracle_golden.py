"""Pins the CPU oracle to outputs of the reference's own code (tests/golden/make_golden.py).

Not a GPU test: this is the "oracle vs golden vectors" leg of the parity chain
reference -> golden -> oracle -> CUDA path.
"""
import numpy as np


def test_matching_L2_fp32(oracle, synth, golden):
    vecs, qvecs = synth.gaussian(512, 8, d=64)
    idx, tpq = oracle.matching_L2(10, vecs.T, qvecs.T)
    assert idx.dtype == np.int64 and idx.shape == (8, 10) and tpq > 0
    np.testing.assert_array_equal(idx, golden["A_matching_L2_idx"])


def test_matching_L2_fp64_unnormalised(oracle, synth, golden):
    vecs, qvecs = synth.gaussian(512, 8, d=64)
    vecs64 = vecs.astype(np.float64) * np.linspace(0.5, 2.0, vecs.shape[1])[None, :]
    q64 = qvecs.astype(np.float64) * 3.0
    idx, _ = oracle.matching_L2(10, vecs64.T, q64.T)
    np.testing.assert_array_equal(idx, golden["B_matching_L2_idx"])


def test_rank_ip(oracle, synth, golden):
    vecs, qvecs = synth.gaussian(512, 8, d=64)
    scores, ranks = oracle.rank_ip(vecs, qvecs)
    np.testing.assert_array_equal(scores, golden["A_scores"])
    np.testing.assert_array_equal(ranks, golden["A_ranks"])


def test_qge1(oracle, synth, golden):
    vecs, qvecs = synth.gaussian(512, 8, d=64)
    ranks = golden["A_ranks"][:10]
    np.testing.assert_array_equal(oracle.qge1(ranks, qvecs, vecs, 10), golden["A_qge1_ranks"])


def test_cfg1_slice_top100(oracle, synth, golden):
    for fam in ("G", "P"):
        v, q = synth.gaussian(1000, 5, d=2048, family=fam)
        idx, _ = oracle.matching_L2(100, v.T, q.T)
        np.testing.assert_array_equal(idx, golden[f"C_{fam}_matching_L2_idx"])
        ids, sims = oracle.topk_ip(v, q, 100)
        np.testing.assert_array_equal(sims.T, golden[f"C_{fam}_top100_scores"])
        # ids identical except inside runs of exactly equal fp32 scores (introsort leaves those
        # unspecified, main_retrieve.py:176; topk_ip orders them by ascending id)
        diff = ids.T != golden[f"C_{fam}_top100"]
        s = sims.T
        tied = np.zeros_like(diff)
        tied[1:] |= s[1:] == s[:-1]
        tied[:-1] |= s[:-1] == s[1:]
        assert not (diff & ~tied).any()
        for j in range(5):
            assert set(ids[j]) == set(golden[f"C_{fam}_top100"][:, j])
        # the distance order and the score order name the same SETS (SURVEY 8c caveat)
        for j in range(5):
            assert set(idx[j]) == set(ids[j])


def test_map_matches_reference_evaluate(oracle, synth, golden):
    v, q, gnd = synth.clustered(3000, 12, d=256, n_clusters=40, noise=1.6, spread=0.5)
    _, ranks = oracle.rank_ip(v, q)
    np.testing.assert_array_equal(ranks[:100], golden["D_top100"])
    m, aps, pr, prs = oracle.compute_map(ranks, gnd, [1, 5, 10])
    assert m == golden["D_map"]
    np.testing.assert_array_equal(aps, golden["D_aps"])
    np.testing.assert_array_equal(pr, golden["D_pr"])
    np.testing.assert_array_equal(prs, golden["D_prs"])
    m100, aps100, _, _ = oracle.compute_map(ranks[:100], gnd, [1, 5, 10])
    assert m100 == golden["D_map_top100"]
    np.testing.assert_array_equal(aps100, golden["D_aps_top100"])
    e, mm, h = oracle.protocol_maps(ranks, gnd)
    assert (e, mm, h) == (golden["D_mapE"], golden["D_mapM"], golden["D_mapH"])


def test_ties_scores(oracle, synth, golden):
    v, q = synth.ties(256, 4, d=64, n_distinct=16)
    scores, ranks = oracle.rank_ip(v, q)
    np.testing.assert_array_equal(np.take_along_axis(scores, ranks, axis=0), golden["E_sorted_scores"])
    ids, sims = oracle.topk_ip(v, q, 40)
    np.testing.assert_array_equal(sims.T, golden["E_sorted_scores"][:40])
    # explicit tie rule: equal scores come out in ascending id order
    for j in range(4):
        for a in range(39):
            if sims[j, a] == sims[j, a + 1]:
                assert ids[j, a] < ids[j, a + 1]


def test_knn_search_contract(oracle, synth):
    v, q = synth.gaussian(300, 6, d=32)
    sims, ids = oracle.knn_search(v.T, q.T, 7, "cosine")
    assert sims.dtype == np.float32 and ids.dtype == np.int64 and sims.shape == ids.shape == (6, 7)
    assert (np.diff(sims, axis=1) <= 0).all()
    _, ranks = oracle.rank_ip(v, q)
    np.testing.assert_array_equal(ids, ranks[:7].T)
    dist, ids2 = oracle.knn_search(v.T, q.T, 7, "euclidean")
    assert (np.diff(dist, axis=1) >= 0).all()
    np.testing.assert_array_equal(ids2, ids)   # unit-norm rows: same order


def test_compare_topk(oracle):
    s = np.array([0.9, 0.8, 0.8 + 1e-9, 0.1])
    f = lambda i: s[i]
    assert oracle.compare_topk([0, 1, 2], [0, 2, 1], f)[0]
    assert not oracle.compare_topk([0, 1, 3], [0, 1, 2], f)[0]
    assert not oracle.compare_topk([0, 1, 1], [0, 1, 2], f)[0]


def test_diffusion_affinity(oracle, golden):
    ids, sims = golden["F_knn_ids"].astype(np.int64), golden["F_knn_sims"]
    np.testing.assert_array_equal(oracle.affinity_dense(sims.copy(), ids), golden["F_affinity"])
    assert golden["F_affinity"].any() and (golden["F_affinity"] != golden["F_affinity"].T).any() is not None


def test_offline_cg_matches_reference(oracle, golden):
    """Case G: the truncated CG restatement against get_offline_result (diffusion.py:15-19) run on scipy."""
    import scipy.sparse as sparse
    lap = sparse.csr_matrix(golden["F_laplacian"])
    ids = golden["G_trunc_ids"].astype(np.int64)
    np.testing.assert_array_equal(ids[:, :12], golden["F_knn_ids"])        # the Laplacian's lists are the prefix
    got = oracle.offline_scores(lap, ids)
    np.testing.assert_allclose(got, golden["G_offline"], rtol=1e-9, atol=1e-12)
    assert (golden["G_offline"][:, 0] > 0.5).all()                          # the row's own score dominates


def test_cg_plain_early_stop_matches_scipy(oracle):
    """Well-conditioned system: the stopping rule fires before maxiter, same iterate as scipy's cg."""
    import scipy.sparse.linalg as linalg
    rng = np.random.default_rng(3)
    m = rng.standard_normal((30, 30))
    a = np.eye(30) + 0.01 * (m @ m.T)
    b = np.zeros(30)
    b[0] = 1
    want, info = linalg.cg(a, b, rtol=1e-6, atol=0.0, maxiter=20)
    assert info == 0
    np.testing.assert_allclose(oracle.cg_plain(a, b, 1e-6, 20), want, rtol=1e-12, atol=1e-15)
    want2, info2 = linalg.cg(a, b, rtol=1e-6, atol=0.0, maxiter=2)          # cut short
    assert info2 == 2
    np.testing.assert_allclose(oracle.cg_plain(a, b, 1e-6, 2), want2, rtol=1e-12, atol=1e-15)


def test_query_expansion_and_database_augmentation(oracle, synth, golden):
    """Case H: the N x N argsort re-rankers (Reranking.py:314-365, 375-440) restated vs the lifted originals."""
    v, q, _ = synth.clustered(600, 8, d=64, n_clusters=20, noise=1.2, spread=0.5)
    for name in ("average_query_expansion", "database_augmentation"):
        ranks, v_aug, q_aug = getattr(oracle, name)(q, v, 20)
        np.testing.assert_array_equal(ranks, golden[f"H_{name}_ranks"])
        assert v_aug.shape == (600, 128 if name.startswith("average") else 64) and q_aug.shape[0] == 8
