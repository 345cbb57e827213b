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
    ix, paths = pkg.store.index_from_store(d)                 # built from the mapped rows; writes the device image
    ids, sims = ix.search(q.T, 10)
    rid, _ = oracle.topk_ip(vecs, q, 10)
    np.testing.assert_array_equal(ids, rid)
    assert len(paths) == 2000
    img = os.path.join(d, pkg.store.INDEX_FILE)
    assert os.path.exists(img) and os.path.getsize(img) >= ix.device_bytes
    ix2, _ = pkg.store.index_from_store(d)                    # straight upload of the image, no kernels
    assert ix2.N == 2000 and ix2.D == 128 and ix2.device_bytes == ix.device_bytes
    for nq in (4, 1):                                         # GEMM path and scan path read the uploaded arrays
        i2, s2 = ix2.search(q.T[:nq], 10)
        np.testing.assert_array_equal(i2, ids[:nq])
        np.testing.assert_array_equal(s2, sims[:nq])
    k1, k2 = ix.self_knn(5)[1], ix2.self_knn(5)[1]
    np.testing.assert_array_equal(k1, k2)
    ix.close(); ix2.close()
    # a renormalised index keeps its own image; a stale image (rows rewritten) is rebuilt
    ixn, _ = pkg.store.index_from_store(d, renormalise=True)
    ixn.close()
    assert os.path.exists(img + ".n")
    with pytest.raises(ValueError):
        pkg.ExactIndex.load(os.path.join(d, pkg.store.ROWS_FILE))     # not an index image
