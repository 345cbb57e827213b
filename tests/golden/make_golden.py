"""Generate the golden fixtures in this directory FROM THE REFERENCE'S OWN CODE.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

``src/utils/nnsearch.py`` and ``src/utils/Reranking.py`` cannot be imported (faiss / annoy /
nanopq / kornia are absent), so the functions on the path are lifted out of the reference
*source text* with ``ast`` and executed unchanged in a namespace that holds only numpy and
time; the two inline ranking statements of ``src/main_retrieve.py:175-176`` are taken by line
number.  ``src/utils/evaluate.py`` imports cleanly and is used directly.  Inputs are
regenerated from seeds by the tests (synth.py), so only the OUTPUTS are stored.
"""
from __future__ import annotations

import ast
import importlib
import os
import sys
import textwrap
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
synth = importlib.import_module("image-search-engine-for-historical-research_b200.synth")


def lift_function(path: str, name: str):
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            code = ast.get_source_segment(src, node)
            ns = {"np": np, "time": time}
            exec(compile(code, f"{path}:{name}", "exec"), ns)
            return ns[name]
    raise KeyError(name)


def lift_lines(path: str, first: int, last: int):
    lines = open(path).read().splitlines()[first - 1:last]
    return textwrap.dedent("\n".join(lines))


def main():
    warnings.simplefilter("ignore")
    ref_matching_L2 = lift_function(f"{REF}/src/utils/nnsearch.py", "matching_L2")
    ref_qge1 = lift_function(f"{REF}/src/utils/Reranking.py", "qge1")
    rank_src = lift_lines(f"{REF}/src/main_retrieve.py", 175, 176)
    assert "np.dot(vecs.T, qvecs)" in rank_src and "np.argsort(-scores, axis=0)" in rank_src, rank_src
    sys.path.insert(0, REF)
    from src.utils import evaluate as ref_eval  # noqa: E402

    out = {}

    # --- case A: small Gaussian, fp32, reference layout (D,N) -----------------------------
    vecs, qvecs = synth.gaussian(512, 8, d=64)
    idx, _ = ref_matching_L2(10, vecs.T, qvecs.T)
    out["A_matching_L2_idx"] = idx
    ns = {"np": np, "vecs": vecs, "qvecs": qvecs}
    exec(rank_src, ns)
    out["A_scores"] = ns["scores"]
    out["A_ranks"] = ns["ranks"]
    out["A_qge1_ranks"] = ref_qge1(ns["ranks"][:10], qvecs, vecs, 10)

    # --- case B: float64 inputs as online.py:96-100 builds them, un-normalised rows ---------
    vecs64 = (vecs.astype(np.float64) * np.linspace(0.5, 2.0, vecs.shape[1])[None, :])
    q64 = qvecs.astype(np.float64) * 3.0
    idx, _ = ref_matching_L2(10, vecs64.T, q64.T)
    out["B_matching_L2_idx"] = idx

    # --- case C: rOxford-shaped cfg1 slice, D=2048, Gaussian + non-negative ---------------
    for fam in ("G", "P"):
        v, q = synth.gaussian(1000, 5, d=2048, family=fam)
        idx, _ = ref_matching_L2(100, v.T, q.T)
        out[f"C_{fam}_matching_L2_idx"] = idx
        ns = {"np": np, "vecs": v, "qvecs": q}
        exec(rank_src, ns)
        out[f"C_{fam}_top100"] = ns["ranks"][:100].astype(np.int32)
        out[f"C_{fam}_top100_scores"] = np.take_along_axis(ns["scores"], ns["ranks"][:100], axis=0)

    # --- case D: clustered set with ground truth -> reference mAP code ---------------------
    v, q, gnd = synth.clustered(3000, 12, d=256, n_clusters=40, noise=1.6, spread=0.5)
    ns = {"np": np, "vecs": v, "qvecs": q}
    exec(rank_src, ns)
    ranks = ns["ranks"]
    m, aps, pr, prs = ref_eval.compute_map(ranks, gnd, [1, 5, 10])
    out["D_map"], out["D_aps"], out["D_pr"], out["D_prs"] = np.float64(m), aps, pr, prs
    m100, aps100, _, _ = ref_eval.compute_map(ranks[:100], gnd, [1, 5, 10])
    out["D_map_top100"], out["D_aps_top100"] = np.float64(m100), aps100
    # new protocol (E/M/H) -- evaluate.py:123-147 only prints, so redo its regrouping here
    for tag, okk, jk in (("E", ["easy"], ["junk", "hard"]), ("M", ["easy", "hard"], ["junk"]),
                         ("H", ["hard"], ["junk", "easy"])):
        g2 = [{"ok": np.concatenate([g[k] for k in okk]),
               "junk": np.concatenate([g[k] for k in jk])} for g in gnd]
        out[f"D_map{tag}"] = np.float64(ref_eval.compute_map(ranks, g2, [1, 5, 10])[0])
    out["D_top100"] = ranks[:100].astype(np.int32)

    # --- case E: duplicates / exact ties -> only scores are well defined -------------------
    v, q = synth.ties(256, 4, d=64, n_distinct=16)
    ns = {"np": np, "vecs": v, "qvecs": q}
    exec(rank_src, ns)
    out["E_sorted_scores"] = np.take_along_axis(ns["scores"], ns["ranks"], axis=0)
    out["E_ranks"] = ns["ranks"]

    # --- case F: diffusion graph -- Diffusion.get_affinity / get_laplacian lifted from the class body ---
    import scipy.sparse as sparse
    dsrc = open(f"{REF}/src/utils/diffusion.py").read()
    cls = [n for n in ast.parse(dsrc).body if isinstance(n, ast.ClassDef) and n.name == "Diffusion"][0]
    ns = {"np": np, "sparse": sparse}
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in ("get_affinity", "get_laplacian"):
            fn.decorator_list = []
            code = ast.get_source_segment(dsrc, fn)
            exec(compile(textwrap.dedent(code), f"diffusion.py:{fn.name}", "exec"), ns)

    class _Self:
        get_affinity = lambda self, sims, ids, gamma=3: ns["get_affinity"](self, sims, ids, gamma)
    v, _ = synth.clustered(400, 1, d=64, n_clusters=12, noise=0.9)[:2]
    oracle = importlib.import_module("oracle.oracle")
    sims, ids = oracle.knn_search(v.T, v.T, 12, "cosine")
    assert (ids[:, 0] == np.arange(400)).all()
    aff = ns["get_affinity"](_Self(), sims.copy(), ids)
    lap = ns["get_laplacian"](_Self(), sims.copy(), ids)
    out["F_knn_ids"] = ids.astype(np.int32)
    out["F_knn_sims"] = sims
    out["F_affinity"] = aff.toarray()
    out["F_laplacian"] = np.asarray(lap.toarray(), dtype=np.float32)

    # --- case G: gallery-side diffusion -- get_offline_result (diffusion.py:15-19) lifted as is.  The one
    # adaptation: the reference pins scipy 1.9 whose cg takes `tol=`; the scipy installed here (>= 1.14)
    # renamed it `rtol`, so the `linalg` name the lifted body sees forwards tol -> rtol (same criterion,
    # ||r|| <= tol * ||b||). ---
    import scipy.sparse.linalg as sp_linalg

    class _Linalg:
        @staticmethod
        def cg(a, b, tol=1e-5, maxiter=None):
            return sp_linalg.cg(a, b, rtol=tol, atol=0.0, maxiter=maxiter)
    n_trunc, kd = 40, 12
    sims40, ids40 = oracle.knn_search(v.T, v.T, n_trunc, "cosine")
    lap_g = ns["get_laplacian"](_Self(), sims40[:, :kd].copy(), ids40[:, :kd])
    trunc_init = np.zeros(n_trunc)
    trunc_init[0] = 1
    gns = {"np": np, "linalg": _Linalg, "trunc_ids": ids40, "trunc_init": trunc_init, "lap_alpha": lap_g}
    gfn = [n for n in ast.parse(dsrc).body if isinstance(n, ast.FunctionDef) and n.name == "get_offline_result"][0]
    exec(compile(ast.get_source_segment(dsrc, gfn), "diffusion.py:get_offline_result", "exec"), gns)
    out["G_trunc_ids"] = ids40.astype(np.int32)
    out["G_trunc_sims"] = sims40
    out["G_offline"] = np.stack([gns["get_offline_result"](i) for i in range(400)])

    # --- case H: average query expansion / database augmentation (Reranking.py:314-365, 375-440), lifted whole.
    # They print their mAP instead of returning the ranks, so the `compute_map_and_print2` name they call is a
    # recorder; `matching_L2` is the reference's own function lifted above. ---
    v, q, _ = synth.clustered(600, 8, d=64, n_clusters=20, noise=1.2, spread=0.5)
    for name in ("average_query_expansion", "database_augmentation"):
        fn = lift_function(f"{REF}/src/utils/Reranking.py", name)
        seen = {}
        fn.__globals__.update(matching_L2=ref_matching_L2, print=lambda *a, **k: None,
                              compute_map_and_print2=lambda dataset, ranks, gnd: seen.setdefault("ranks", ranks))
        fn(q.copy(), v.copy(), 20, "synthetic", None)
        out[f"H_{name}_ranks"] = seen["ranks"].astype(np.int32)

    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    for k, a in out.items():
        print(f"{k:28s} {str(a.dtype):8s} {a.shape}")


if __name__ == "__main__":
    main()
