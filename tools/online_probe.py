"""The online request path of the reference (src/online.py:121-152) through the drop-in Python API:
matching_L2(K, vecs.T, qvec.T) followed by qge1(ranks, qvec, vecs, K), one query at a time, host arrays.

    python tools/online_probe.py [N]
"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("image-search-engine-for-historical-research_b200")
synth = importlib.import_module("image-search-engine-for-historical-research_b200.synth")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_007_000
K = 100
t0 = time.time()
vecs = np.ascontiguousarray(synth.rows(N, 2048, 0).T)            # (D, N) like online.py's `vecs`
qs = synth.rows(32, 2048, 1)
print(f"host data ready in {time.time()-t0:.1f} s", flush=True)
t0 = time.perf_counter()
idx, _ = pkg.matching_L2(K, vecs.T, qs[:1])                        # first call uploads the database (renormalised index)
print(f"first matching_L2 call (index build from the F-order host view): {time.perf_counter()-t0:.2f} s", flush=True)
t0 = time.perf_counter()
r2 = pkg.qge1(idx.T, qs[:1].T, vecs, K)                            # first call uploads the un-normalised twin index
print(f"first qge1 call (second index): {time.perf_counter()-t0:.2f} s", flush=True)
lat_m, lat_q = [], []
for i in range(1, 32):
    qvec = qs[i:i + 1].T                                           # (2048, 1) like online.py:123
    t0 = time.perf_counter()
    match_idx, tpq = pkg.matching_L2(K, vecs.T, qvec.T)
    t1 = time.perf_counter()
    ranks2 = pkg.qge1(match_idx.T, qvec, vecs, K)
    t2 = time.perf_counter()
    lat_m.append(t1 - t0); lat_q.append(t2 - t1)
print(f"per request over {len(lat_m)} requests: matching_L2 {np.median(lat_m)*1e3:.3f} ms (reported time_per_query {tpq*1e3:.3f} ms), "
      f"qge1 {np.median(lat_q)*1e3:.3f} ms, total {np.median(np.add(lat_m, lat_q))*1e3:.3f} ms", flush=True)
