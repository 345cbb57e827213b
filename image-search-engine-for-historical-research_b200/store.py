"""Feature store: a memory-mappable on-disk form of the descriptor matrix.

The reference persists descriptors as a pickle ``{'path': [...], 'feature': ndarray (D, N)}``
(``save_path_feature`` / ``load_path_features``, src/utils/general.py:67-92) and as a torch file
for the R1M distractors (src/extract_1m.py:98, src/test_rOP1m.py:137); at start-up ``online.py``
unpickles, concatenates into a float64 array and transposes views of it on every request
(src/online.py:93-102, 133).  The store keeps ROW-major fp32 ``(N, D)`` in a plain ``.npy`` (so
``np.load(mmap_mode='r')`` maps it without reading) next to the path list; building the device
index from it is then a straight staged copy -- no unpickle, no concatenate, no transpose, no
accidental float64.
"""
from __future__ import annotations

import json
import os
import pickle

import numpy as np

ROWS_FILE = "rows.npy"
PATHS_FILE = "paths.json"
INDEX_FILE = "index.xsb"          # device image: fp32 rows + tiled bf16 rows, as they live in HBM (xs_index_save)


def save_store(directory: str, vecs, paths=None, chunk: int = 65536) -> str:
    """Write ``vecs`` -- the reference's ``(D, N)`` array -- as a row-major fp32 store."""
    vecs = np.asarray(vecs)
    d, n = vecs.shape
    os.makedirs(directory, exist_ok=True)
    out = np.lib.format.open_memmap(os.path.join(directory, ROWS_FILE), mode="w+", dtype=np.float32, shape=(n, d))
    for lo in range(0, n, chunk):                      # transposed in blocks: no second full-size copy
        hi = min(n, lo + chunk)
        out[lo:hi] = vecs[:, lo:hi].T
    out.flush()
    del out
    with open(os.path.join(directory, PATHS_FILE), "w") as f:
        json.dump(list(paths) if paths is not None else [], f)
    return directory


def open_store(directory: str):
    """``(rows, paths)``: ``rows`` is a read-only memory map ``(N, D)`` fp32; ``rows.T`` is the
    reference's ``vecs`` view."""
    rows = np.load(os.path.join(directory, ROWS_FILE), mmap_mode="r")
    with open(os.path.join(directory, PATHS_FILE)) as f:
        paths = json.load(f)
    return rows, paths


def load_path_features(pickle_path: str):
    """Reader for the reference's own pickle (general.py:84-92): ``(vecs (D, N), img_r_path)``."""
    with open(pickle_path, "rb") as f:
        pf = pickle.load(f)
    return pf["feature"], pf["path"]


def convert_pickle(pickle_path: str, directory: str) -> str:
    """One-off conversion of a reference feature pickle into a store."""
    vecs, paths = load_path_features(pickle_path)
    return save_store(directory, vecs, paths)


def load_distractors(pt_path: str):
    """Reader for the reference's R1M distractor file (``torch.save(vecs, '<net>_vecs_revisitop1m.pt')``,
    src/extract_1m.py:98; read back at src/test_rOP1m.py:137-138): ``vecs (D, N)`` as numpy."""
    import torch
    t = torch.load(pt_path, map_location="cpu")
    return t.numpy() if hasattr(t, "numpy") else np.asarray(t)


def convert_pt(pt_path: str, directory: str, paths=None) -> str:
    """One-off conversion of the distractor ``.pt`` file into a store."""
    return save_store(directory, load_distractors(pt_path), paths)


def append_store(directory: str, vecs, paths=None, chunk: int = 65536) -> str:
    """Append the columns of ``vecs (D, M)`` to an existing store -- the store form of
    ``vecs = np.concatenate([vecs, vecs_1m], axis=1)`` (src/test_rOP1m.py:139).  The rows file is rewritten
    block by block into a new file and swapped in, so a reader holding the old map keeps a consistent view."""
    vecs = np.asarray(vecs)
    old, old_paths = open_store(directory)
    n, d = old.shape
    if vecs.shape[0] != d:
        raise ValueError(f"dimension mismatch: store has D={d}, got {vecs.shape[0]}")
    m = vecs.shape[1]
    tmp = os.path.join(directory, ROWS_FILE + ".tmp")
    out = np.lib.format.open_memmap(tmp, mode="w+", dtype=np.float32, shape=(n + m, d))
    for lo in range(0, n, chunk):
        out[lo:min(n, lo + chunk)] = old[lo:min(n, lo + chunk)]
    for lo in range(0, m, chunk):
        hi = min(m, lo + chunk)
        out[n + lo:n + hi] = vecs[:, lo:hi].T
    out.flush()
    del out, old
    os.replace(tmp, os.path.join(directory, ROWS_FILE))
    new_paths = list(old_paths) + (list(paths) if paths is not None else [])
    with open(os.path.join(directory, PATHS_FILE), "w") as f:
        json.dump(new_paths, f)
    return directory


def index_from_store(directory: str, renormalise: bool = False, device: int = 0, image: bool = True):
    """Device index of a store.  The first call builds it from the mapped rows and (``image=True``) writes the device
    image next to them; later calls upload that image as it is -- no layout / convert / tile kernels, pinned
    double-buffered copies at PCIe speed (xs_index_load).  The image is tied to ``renormalise`` and to the rows file's
    modification time."""
    from .index import ExactIndex
    rows, paths = open_store(directory)
    img = os.path.join(directory, INDEX_FILE + (".n" if renormalise else ""))
    rows_path = os.path.join(directory, ROWS_FILE)
    if image and os.path.exists(img) and os.path.getmtime(img) >= os.path.getmtime(rows_path):
        ix = ExactIndex.load(img, device=device)
        ix.renormalised = bool(renormalise)
        return ix, paths
    ix = ExactIndex(rows, renormalise=renormalise, device=device)
    if image:
        ix.save(img)
    return ix, paths
