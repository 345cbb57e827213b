"""BASELINE.json's full size (1,007,000 x 2048, 70 queries, top-100) on the GPU: the oracle's own top-100 for all
70 queries (family G), size-independent properties, agreement between the three independent CUDA paths (tcgen05
GEMM / bf16 scan / exact fp32) and a float64 recomputation of the returned scores with torch."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N, D, Q, K = 1_007_000, 2048, 70, 100


def _rows(torch, n, seed, family):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    out = torch.empty((n, D), dtype=torch.float32, device="cuda")
    for lo in range(0, n, 65536):
        hi = min(n, lo + 65536)
        blk = torch.randn((hi - lo, D), generator=g, dtype=torch.float32, device="cuda")
        if family == "P":
            blk = blk.abs()
        out[lo:hi] = blk / blk.norm(dim=1, keepdim=True)
    return out


@pytest.mark.parametrize("family", ["G", "P"])
def test_full_size_properties(pkg, oracle, family):
    import torch
    rows = _rows(torch, N, 0, family)
    queries = _rows(torch, Q, 1, family)
    q_np = queries.cpu().numpy()
    ix = pkg.ExactIndex.from_device(rows.data_ptr(), N, D, 0)
    try:
        ids, sims = ix.search(q_np, K)                       # default dispatch: GEMM + fused top-K
        st = ix.stats()
        assert st["path"] == 2
        # every query certified by the bf16 pass, or re-run exactly -- never silently wrong
        assert st["n_exact_rerun"] <= Q // 10, st
        # properties: shape, range, uniqueness, sortedness, tie rule
        assert ids.shape == (Q, K) and ids.min() >= 0 and ids.max() < N
        for j in range(Q):
            assert len(np.unique(ids[j])) == K
        d = np.diff(sims, axis=1)
        assert (d <= 0).all()
        tie = d == 0
        assert (np.diff(ids, axis=1)[tie] > 0).all()
        # returned scores are the exact inner products (float64 recomputation on the device)
        got = torch.from_numpy(ids).cuda()
        ref = torch.einsum("qkd,qd->qk", rows[got].double(), queries.double()).float().cpu().numpy()
        np.testing.assert_allclose(sims, ref, rtol=1e-6, atol=1e-7)
        # the K-th score really is a top-K boundary: no row outside the list beats it (checked for 4 queries, fp64)
        for j in (0, 23, 46, 69):
            s = (rows.double() @ queries[j].double())
            kth = torch.topk(s, K).values[-1].item()
            assert abs(kth - float(sims[j, -1])) <= 1e-6 * abs(kth) + 1e-7
        if family == "G":
            # the oracle itself at the full size, all 70 lists (np.dot + top-k on the host: ~10 s, 8 GB of host memory)
            rows_h = rows.cpu().numpy()
            ref_i, ref_s = oracle.topk_ip(rows_h.T, q_np.T, K)
            q64 = q_np.astype(np.float64)
            for j in range(Q):
                ok, msg = oracle.compare_topk(ids[j], ref_i[j], lambda i, j=j: rows_h[i].astype(np.float64) @ q64[j])
                assert ok, f"full size, query {j}: {msg}"
            np.testing.assert_allclose(sims, ref_s, rtol=1e-5, atol=1e-7)
            del rows_h
        # path agreement: exact fp32 path for 6 queries, bf16 scan path for 2
        ix.set_param("force_path", 3)
        xi, xs = ix.search(q_np[:6], K)
        np.testing.assert_array_equal(xi, ids[:6])
        np.testing.assert_array_equal(xs, sims[:6])
        ix.set_param("force_path", 1)
        si, ss = ix.search(q_np[:2], K)
        np.testing.assert_array_equal(si, ids[:2])
        np.testing.assert_array_equal(ss, sims[:2])
        # idempotence
        ix.set_param("force_path", 0)
        ids2, sims2 = ix.search(q_np, K)
        np.testing.assert_array_equal(ids2, ids)
        np.testing.assert_array_equal(sims2, sims)
        # large k at full size: the bootstrap pass sizes its sample to k; coarse paths vs exact path
        if family == "G":
            for kk in (1000, 2048):
                ix.set_param("force_path", 2)
                gi, gs = ix.search(q_np[:3], kk)
                assert ix.stats()["path"] == 2
                ix.set_param("force_path", 3)
                ei, es = ix.search(q_np[:3], kk)
                np.testing.assert_array_equal(gi, ei)
                np.testing.assert_array_equal(gs, es)
            ix.set_param("force_path", 0)
        # batch-1 dispatch takes the scan path and agrees
        i1, s1 = ix.search(q_np[5:6], K)
        assert ix.stats()["path"] == 1
        np.testing.assert_array_equal(i1[0], ids[5])
    finally:
        ix.close()
        del rows
        torch.cuda.empty_cache()
