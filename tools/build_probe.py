"""Index build time from HOST memory in the reference's layouts (F-order view of a (D,N) array, fp32 / fp64).

    python tools/build_probe.py [N]
"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("image-search-engine-for-historical-research_b200")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000
rng = np.random.default_rng(0)
vecs = rng.standard_normal((2048, N), dtype=np.float32)          # the reference's (D, N) C-contiguous array
pkg.ExactIndex(vecs.T[:1000].copy()).close()                      # context + module load out of the timing
for name, arr in (("fp32 F-order view vecs.T", vecs.T), ("fp32 row-major copy", np.ascontiguousarray(vecs.T)),
                  ("fp64 F-order view (online.py:96-100)", vecs.astype(np.float64).T)):
    t0 = time.perf_counter()
    ix = pkg.ExactIndex(arr, renormalise=True)
    dt = time.perf_counter() - t0
    print(f"{name:40s}: {dt:.3f} s for {N} rows ({arr.nbytes/1e9:.2f} GB host -> {ix.device_bytes/1e9:.2f} GB device) = {arr.nbytes/dt/1e9:.2f} GB/s", flush=True)
    ix.close()

# device image: save once, then start-up = a straight pinned, multi-threaded upload (xs_index_load)
import tempfile
with tempfile.TemporaryDirectory() as tmp:
    ix = pkg.ExactIndex(np.ascontiguousarray(vecs.T), renormalise=True)
    img = os.path.join(tmp, "index.xsb")
    t0 = time.perf_counter(); ix.save(img); dt = time.perf_counter() - t0
    size = os.path.getsize(img)
    print(f"xs_index_save: {dt:.3f} s for {size/1e9:.2f} GB = {size/dt/1e9:.2f} GB/s", flush=True)
    want = ix.search(vecs.T[:4].copy(), 10)
    ix.close()
    for rep in range(3):
        t0 = time.perf_counter(); ix2 = pkg.ExactIndex.load(img); dt = time.perf_counter() - t0
        print(f"xs_index_load (file in the page cache): {dt:.3f} s for {size/1e9:.2f} GB = {size/dt/1e9:.2f} GB/s host -> device", flush=True)
        got = ix2.search(vecs.T[:4].copy(), 10)
        assert (got[0] == want[0]).all() and (got[1] == want[1]).all()
        ix2.close()
