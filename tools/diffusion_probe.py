"""Gallery-side diffusion at the reference's parameters (Reranking.py:231-235: n_trunc 2000, kd 200 on N < 120 000):
GPU truncated CG per row vs scipy's slicing + cg on a sample of rows (the statement of diffusion.py:15-19)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("image-search-engine-for-historical-research_b200")
synth = importlib.import_module("image-search-engine-for-historical-research_b200.synth")
import scipy.sparse.linalg as linalg

n = int(os.environ.get("XS_PROBE_N", 30000)); T = int(os.environ.get("XS_PROBE_T", 2000)); kd = int(os.environ.get("XS_PROBE_KD", 200))
v, _ = synth.clustered(n, 1, d=256, n_clusters=max(8, n // 150), noise=0.7)[:2]
d = pkg.diffusion.Diffusion(v.T, None)
t0 = time.time(); sims, ids = d.knn.self_search(T); t1 = time.time()
lap = d.get_laplacian(sims[:, :kd].copy(), ids[:, :kd]).tocsr(); t2 = time.time()
print(f"N={n} n_trunc={T} kd={kd}: self-kNN {t1-t0:.2f}s, Laplacian {t2-t1:.2f}s, nnz/row {lap.nnz/n:.1f}", flush=True)
for rep in range(2):
    t0 = time.time(); out = pkg.diffusion.offline_cg(lap, ids); t3 = time.time()
    print(f"GPU truncated CG, all {n} rows: {t3-t0:.2f}s ({(t3-t0)/n*1e6:.1f} us/row)", flush=True)
b = np.zeros(T); b[0] = 1
sample = np.arange(0, n, max(1, n // 40))[:40]
t0 = time.time()
ref = np.stack([linalg.cg(lap[ids[i]][:, ids[i]], b, rtol=1e-6, atol=0.0, maxiter=20)[0] for i in sample])
t1 = time.time()
print(f"scipy slicing + cg, {len(sample)} rows on one thread: {(t1-t0)/len(sample)*1e3:.2f} ms/row -> {(t1-t0)/len(sample)*n:.0f}s for all rows")
print("max |gpu - scipy| =", np.abs(out[sample] - ref).max(), " max rel =", (np.abs(out[sample] - ref) / (np.abs(ref) + 1e-9)).max())

# the whole gallery side on the device: self-kNN -> mutual graph -> Laplacian -> CG with nothing visiting the host in between
for rep in range(2):
    t0 = time.time()
    oi, _, osc = pkg.diffusion.offline_device(d.knn.index, T, kd)
    t1 = time.time()
    print(f"device-resident pipeline (xs_diffusion_offline), all {n} rows: {t1-t0:.2f}s", flush=True)
print("device pipeline vs stagewise: ids equal", bool((oi == ids).all()), " max |score diff| =", float(np.abs(osc - out).max()))
