"""Feature store round trip (CPU) and its use as an index source (GPU)."""
import os
import pickle

import numpy as np
import pytest


def test_store_roundtrip_and_reference_pickle(tmp_path, pkg, synth):
    vecs, _ = synth.gaussian(300, 1, d=40)
    paths = [f"img_{i}.jpg" for i in range(300)]
    # the reference's own format (general.py:67-81)
    pkl = os.path.join(tmp_path, "db_path_feature.pkl")
    with open(pkl, "wb") as f:
        pickle.dump({"path": paths, "feature": vecs}, f)
    v2, p2 = pkg.store.load_path_features(pkl)
    np.testing.assert_array_equal(v2, vecs)
    assert p2 == paths
    d = pkg.store.convert_pickle(pkl, os.path.join(tmp_path, "store"))
    rows, p3 = pkg.store.open_store(d)
    assert isinstance(rows, np.memmap) and rows.shape == (300, 40) and rows.dtype == np.float32
    np.testing.assert_array_equal(np.asarray(rows), vecs.T)
    assert p3 == paths and rows.flags["C_CONTIGUOUS"]
    # float64 input (online.py:96-100) is stored as fp32
    pkg.store.save_store(os.path.join(tmp_path, "s64"), vecs.astype(np.float64))
    r64, _ = pkg.store.open_store(os.path.join(tmp_path, "s64"))
    np.testing.assert_array_equal(np.asarray(r64), vecs.T)


@pytest.mark.gpu
def test_index_from_store(tmp_path, pkg, synth, oracle):
    vecs, q = synth.gaussian(2000, 4, d=128)
    d = pkg.store.save_store(os.path.join(tmp_path, "store"), vecs, [str(i) for i in range(2000)])
    ix, paths = pkg.store.index_from_store(d)
    ids, _ = ix.search(q.T, 10)
    rid, _ = oracle.topk_ip(vecs, q, 10)
    np.testing.assert_array_equal(ids, rid)
    assert len(paths) == 2000
    ix.close()


def test_search_offline_matches_reference_statements(pkg):
    """Query side of the diffusion re-ranking (Reranking.py:243-256) -- host-only code, so it runs on CPU:
    ``scores = sims[i] @ offline[idx[i]]`` then argpartition + argsort of the ``n_trunc`` best."""
    import scipy.sparse as sparse
    rng = np.random.default_rng(5)
    n, n_trunc = 300, 40
    offline = sparse.random(n, n, density=0.3, random_state=7, dtype=np.float32, format="csr")
    sims = rng.random((6, 3)).astype(np.float32)
    idx = rng.integers(0, n, size=(6, 3))
    got_s, got_r = pkg.diffusion.search_offline(offline, sims, idx, n_trunc)
    cubed = sims ** 3
    for i in range(6):
        scores = cubed[i] @ offline[idx[i]]                                   # the reference's statements
        parts = np.argpartition(-scores, n_trunc)[:n_trunc]
        ranks = np.argsort(-scores[parts])
        np.testing.assert_allclose(got_s[i], scores[parts][ranks], rtol=1e-6)
        np.testing.assert_allclose(scores[got_r[i]], got_s[i], rtol=1e-6)
    assert got_r.dtype == np.int64 and got_s.dtype == np.float32


def test_distractor_file_and_append(tmp_path, pkg, synth):
    """The R1M flow of test_rOP1m.py:137-139: dataset vectors + a torch-saved distractor matrix, concatenated."""
    import torch
    vecs, _ = synth.gaussian(300, 1, d=32)
    extra, _ = synth.gaussian(500, 1, d=32)
    pt = os.path.join(tmp_path, "net_vecs_revisitop1m.pt")
    torch.save(torch.from_numpy(extra), pt)
    np.testing.assert_array_equal(pkg.store.load_distractors(pt), extra)
    d = pkg.store.save_store(os.path.join(tmp_path, "db"), vecs, [f"a{i}" for i in range(300)])
    pkg.store.append_store(d, pkg.store.load_distractors(pt), [f"b{i}" for i in range(500)])
    rows, paths = pkg.store.open_store(d)
    assert rows.shape == (800, 32) and rows.dtype == np.float32 and len(paths) == 800 and paths[300] == "b0"
    np.testing.assert_array_equal(np.asarray(rows).T, np.concatenate([vecs, extra], axis=1))
    d2 = pkg.store.convert_pt(pt, os.path.join(tmp_path, "only1m"))
    assert pkg.store.open_store(d2)[0].shape == (500, 32)
    with pytest.raises(ValueError):
        pkg.store.append_store(d, np.zeros((31, 4), np.float32))
