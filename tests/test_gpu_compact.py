"""The compact index: only the TILED bf16 array is kept in HBM (second half).

Round 1 kept two bf16 copies of the database -- row-major for the batch-1 scan and the self-kNN's query operand, tiled for
the GEMM's TMA boxes.  The scan now reads the tiled array too (scan_scores_tiled_kernel), the self-kNN's rows go through
the ordinary query preparation, the build tiles chunk by chunk and the image file holds one bf16 section:
2 * d_pad bytes per row (4.1 GB per million 2048-d rows) come back, at build time as well.

Every other GPU test already runs on compact indexes (it is the default); these pin what is specific to it: the byte
count, equality with the two-copy index on every coarse path, ragged shapes of the tiled scan (rows not a multiple of
16 / 256, an odd number of 64-column blocks), self-kNN, and the image round trip in both directions.
"""
import importlib
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("XS_NO_TILED", "0") not in ("", "0"), reason="XS_NO_TILED=1: no tiled array, nothing to compact")]


@pytest.fixture()
def nat(pkg):
    n = importlib.import_module(pkg.__name__ + "._native")
    yield n
    n.config_set("compact", 1)


def _rows(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


def _bytes16(n, d):
    n_pad, d_pad = (n + 255) // 256 * 256, (d + 63) // 64 * 64
    return n_pad * d_pad * 2


@pytest.mark.parametrize("n,d", [(5000, 2048), (4097, 192), (777, 64), (12345, 520)])
def test_compact_equals_two_copy_index_on_every_path(pkg, oracle, nat, n, d):
    rows, q = _rows(n, d, 1), _rows(9, d, 2)
    vecs, qvecs = np.ascontiguousarray(rows.T), np.ascontiguousarray(q.T)
    k = 50
    ref_i, _ = oracle.topk_ip(vecs, qvecs, k)
    s64 = oracle.scores_f64(vecs, qvecs)
    nat.config_set("compact", 0)
    with pkg.ExactIndex(rows) as fat:
        nat.config_set("compact", 1)
        with pkg.ExactIndex(rows) as slim:
            assert fat.device_bytes - slim.device_bytes == _bytes16(n, d)
            for path in (0, 1, 2):                       # auto, scan, GEMM
                for ix in (fat, slim):
                    ix.set_param("force_path", path)
                a_i, a_s = fat.search(q, k)
                b_i, b_s = slim.search(q, k)
                assert slim.stats()["n_exact_rerun"] == 0
                np.testing.assert_array_equal(a_i, b_i)
                np.testing.assert_array_equal(a_s, b_s)
                for j in range(q.shape[0]):
                    ok, msg = oracle.compare_topk(b_i[j], ref_i[j], lambda i, j=j: s64[i, j])
                    assert ok, f"n={n} d={d} path={path} query {j}: {msg}"
            # one query at a time: the batch-1 online shape (scan path by default)
            slim.set_param("force_path", 0)
            for j in range(3):
                i1, _ = slim.search(q[j:j + 1], k)
                ok, msg = oracle.compare_topk(i1[0], ref_i[j], lambda i, j=j: s64[i, j])
                assert ok, msg
            # the two scan kernels on the two-copy index agree with each other as well
            fat.set_param("force_path", 1)
            fat.set_param("scan_tiled", 0)
            c_i, c_s = fat.search(q, k)
            np.testing.assert_array_equal(c_i, b_i)


def test_compact_self_knn(pkg, oracle, nat):
    """Three 8192-row batches through the two-deep pipeline: on the compact index the rows reach the GEMM through the
    ordinary query preparation, on the two-copy index straight from the row-major copy -- same lists, own id first,
    and no exact re-runs on plain data (a re-run is correct but costs a full fp32 scan: 35 s instead of 0.2 s at 1M rows)."""
    n = 20000
    rows = _rows(n, 128, 3)
    out = {}
    for compact in (1, 0):
        nat.config_set("compact", compact)
        with pkg.ExactIndex(rows) as ix:
            sims, ids = ix.self_knn(11)
            assert ix.stats()["n_exact_rerun"] <= n // 1000, f"compact {compact}: {ix.stats()['n_exact_rerun']} exact re-runs"
        np.testing.assert_array_equal(ids[:, 0], np.arange(n))          # a row's own id comes first
        out[compact] = (sims, ids)
    np.testing.assert_array_equal(out[0][1], out[1][1])
    np.testing.assert_array_equal(out[0][0], out[1][0])
    ids = out[1][1]
    pick = np.array([0, 1, 4097, 8191, 8192, 8193, 12345, 16383, 16384, 19999])
    s = rows[pick].astype(np.float64) @ rows.astype(np.float64).T
    s[np.arange(pick.size), pick] = -np.inf
    ref = np.argsort(-s, axis=1, kind="stable")[:, :10]
    for j, r in enumerate(pick):
        ok, msg = oracle.compare_topk(ids[r, 1:], ref[j], lambda i, j=j: s[j, i])
        assert ok, f"row {r}: {msg}"


def test_image_round_trip_both_ways(pkg, nat, tmp_path):
    n, d = 6001, 320
    rows, q = _rows(n, d, 4), _rows(5, d, 5)
    with pkg.ExactIndex(rows) as ix:
        want_i, want_s = ix.search(q, 20)
        ix.save(str(tmp_path / "slim.xsb"))
    nat.config_set("compact", 0)
    with pkg.ExactIndex(rows) as fat:
        fat.save(str(tmp_path / "fat.xsb"))
        fat_bytes = fat.device_bytes
    assert os.path.getsize(tmp_path / "fat.xsb") - os.path.getsize(tmp_path / "slim.xsb") >= _bytes16(n, d)
    # a compact file always loads compact; a two-copy file loads as the process default says
    for compact, name, expect_fat in ((0, "slim.xsb", False), (1, "slim.xsb", False), (1, "fat.xsb", False), (0, "fat.xsb", True)):
        nat.config_set("compact", compact)
        with pkg.ExactIndex.load(str(tmp_path / name)) as ix:
            assert (ix.device_bytes == fat_bytes) == expect_fat, (name, compact)
            for path in (1, 2):
                ix.set_param("force_path", path)
                got_i, got_s = ix.search(q, 20)
                np.testing.assert_array_equal(got_i, want_i)
                np.testing.assert_array_equal(got_s, want_s)
