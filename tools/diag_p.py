"""Why does a family need exact re-runs?  Per (k, path): re-runs and rescored candidates per query."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("image-search-engine-for-historical-research_b200")
synth = importlib.import_module("image-search-engine-for-historical-research_b200.synth")
for fam, n, d in (("P", 20000, 512), ("G", 20000, 512), ("P", 200000, 2048)):
    v, q = synth.gaussian(n, 12, d=d, family=fam)
    rows, queries = np.ascontiguousarray(v.T), np.ascontiguousarray(q.T)
    s = rows @ queries.T
    print(f"family {fam} n={n} d={d}: score mean {s.mean():.4f} std {s.std():.4f}; top-100 boundary {np.sort(s[:,0])[-100]:.4f}, max {s[:,0].max():.4f}")
    with pkg.ExactIndex(rows) as ix:
        for cert in (0, 1):
            ix.set_param("certificate", cert)
            for k in (1, 100):
                for path in (2, 1):
                    ix.set_param("force_path", path)
                    for nq in (12, 1):
                        ix.search(queries[:nq], k)
                        st = ix.stats()
                        print(f"  cert {cert} k {k:3d} path {path} nq {nq:2d}: reruns {st['n_exact_rerun']:2d}, candidates/query {st['n_candidates']/nq:7.1f}")
