"""ctypes binding of libxs_b200.so -- the only door between the Python host code and the CUDA path.

There is deliberately no fallback: if the shared library is missing or the machine has no
CUDA device, every compute entry point raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libxs_b200.so")

XS_F32, XS_F64 = 0, 1
PATH_AUTO, PATH_SCAN, PATH_GEMM, PATH_EXACT = 0, 1, 2, 3

# every symbol include/xs_b200.h declares (tests check the library exports exactly these)
ABI_SYMBOLS = (
    "xs_last_error", "xs_abi_version", "xs_device_count", "xs_index_create", "xs_index_create_dev",
    "xs_index_destroy", "xs_index_clone", "xs_index_info", "xs_index_stats", "xs_search", "xs_search_dev", "xs_self_knn",
    "xs_rank_all", "xs_merge_candidates", "xs_set_param", "xs_aqe_search", "xs_merge_candidates_strided", "xs_mutual_knn", "xs_diffusion_cg",
    "xs_exchange_create", "xs_exchange_connect", "xs_exchange_push", "xs_exchange_merge", "xs_exchange_destroy",
    "xs_exchange_part_bytes", "xs_search_dev_push", "xs_config_set", "xs_diffusion_laplacian", "xs_diffusion_offline", "xs_debug_trace", "xs_index_save", "xs_index_load",
    "xs_search_dev_exchange", "xs_pipeline_create", "xs_pipeline_submit", "xs_pipeline_collect", "xs_pipeline_destroy", "xs_pipeline_lane",
)


class XsStats(C.Structure):
    _fields_ = [("n_queries", C.c_int64), ("n_exact_rerun", C.c_int64), ("n_candidates", C.c_int64),
                ("path", C.c_int32), ("gpu_launches", C.c_int32), ("ms_coarse", C.c_float), ("ms_total", C.c_float)]


_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load the shared library (once).  Raises RuntimeError with the build hint if it is absent."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run "
                f"`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc with sm_100a). "
                f"There is no CPU fallback for the matching path.")
        lib = C.CDLL(LIB_PATH)
        p, i64, i32, f32p, i64p = C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int64)
        lib.xs_last_error.restype = C.c_char_p
        lib.xs_last_error.argtypes = []
        lib.xs_abi_version.restype = i32
        lib.xs_abi_version.argtypes = []
        lib.xs_device_count.argtypes = [C.POINTER(i32)]
        lib.xs_index_create.argtypes = [p, i32, i64, i32, i64, i64, i32, i32, i64, C.POINTER(p)]
        lib.xs_index_create_dev.argtypes = [p, i64, i32, i32, i32, i64, C.POINTER(p)]
        lib.xs_index_destroy.argtypes = [p]
        lib.xs_index_save.argtypes = [p, C.c_char_p]
        lib.xs_index_load.argtypes = [C.c_char_p, i32, i64, C.POINTER(p)]
        lib.xs_index_clone.argtypes = [p, C.POINTER(p)]
        lib.xs_index_info.argtypes = [p, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]
        lib.xs_index_stats.argtypes = [p, C.POINTER(XsStats)]
        lib.xs_search.argtypes = [p, p, i32, i64, i64, i64, i32, i32, p, p]
        lib.xs_search_dev.argtypes = [p, p, i64, i32, i32, p, p, p, p]
        lib.xs_self_knn.argtypes = [p, i64, i64, i32, p, p]
        lib.xs_aqe_search.argtypes = [p, p, i64, i32, C.c_double, i32, p, p, p]
        lib.xs_rank_all.argtypes = [p, p, i32, i64, i64, i64, i32, p, p]
        lib.xs_merge_candidates.argtypes = [i32, p, p, i32, i64, i32, p, p, p]
        lib.xs_merge_candidates_strided.argtypes = [i32, p, p, p, i64, i64, i64, i32, i64, i32, p, p, p, p]
        lib.xs_mutual_knn.argtypes = [i32, p, i64, i32, p]
        lib.xs_diffusion_cg.argtypes = [i32, p, p, p, i64, p, i64, i32, i32, C.c_double, p]
        lib.xs_diffusion_laplacian.argtypes = [i32, p, p, i64, i32, C.c_double, C.c_double, p, p, p, p]
        lib.xs_diffusion_offline.argtypes = [p, i32, i32, C.c_double, C.c_double, i32, C.c_double, p, p, p]
        lib.xs_set_param.argtypes = [p, C.c_char_p, C.c_double]
        lib.xs_debug_trace.argtypes = [p, i32, p, i32, C.POINTER(i32)]
        lib.xs_config_set.argtypes = [C.c_char_p, C.c_double]
        lib.xs_exchange_part_bytes.argtypes = [i64, i32]
        lib.xs_exchange_create.argtypes = [i32, i32, i32, i64, i32, C.POINTER(p), p]
        lib.xs_exchange_connect.argtypes = [p, p]
        lib.xs_search_dev_push.argtypes = [p, p, i64, i32, i32, p, i32, p]
        lib.xs_search_dev_exchange.argtypes = [p, p, i64, i32, i32, p, i32, p, p, p, p]
        lib.xs_exchange_push.argtypes = [p, p, i64, i32, i32, p]
        lib.xs_exchange_merge.argtypes = [p, i32, i64, i32, p, p, p, p]
        lib.xs_exchange_destroy.argtypes = [p]
        lib.xs_pipeline_create.argtypes = [p, p, i64, i32, i32, C.POINTER(p)]
        lib.xs_pipeline_submit.argtypes = [p, p, i64, i32, p, C.POINTER(i32)]
        lib.xs_pipeline_collect.argtypes = [p, i32, p, C.POINTER(p), C.POINTER(p), C.POINTER(i64), p]
        lib.xs_pipeline_destroy.argtypes = [p]
        lib.xs_pipeline_lane.argtypes = [p, i32]
        lib.xs_pipeline_lane.restype = p
        for name in ABI_SYMBOLS:
            if name not in ("xs_last_error", "xs_exchange_part_bytes", "xs_pipeline_lane"):
                getattr(lib, name).restype = i32
        lib.xs_exchange_part_bytes.restype = i64
        _lib = lib
        return lib


def config_set(name: str, value: float) -> None:
    """Process-wide defaults picked up by later index builds (xs_config_set): ``rotation`` 0/1, ``rotation_seed``,
    ``compact`` 1/0 (one tiled bf16 copy of the database, or the row-major copy next to it as well)."""
    check(load().xs_config_set(name.encode(), float(value)), "xs_config_set")


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().xs_last_error().decode("utf-8", "replace")
        kind = {1: ValueError, 3: MemoryError, 4: NotImplementedError}.get(rc, RuntimeError)
        raise kind(f"{what}: {msg}")


def as_matrix(a, name: str):
    """numpy (rows, cols) of any float dtype / strides -> (array kept alive, dtype code, stride_row, stride_col).

    The two layouts the C ABI takes directly are passed through untouched (row-major, and the
    reference's F-order view ``vecs.T``); anything else is made C-contiguous first.
    """
    a = np.asarray(a)
    if a.ndim != 2:
        raise ValueError(f"{name} must be 2-D, got shape {a.shape}")
    if a.dtype == np.float64:
        code = XS_F64
    else:
        if a.dtype != np.float32:
            a = a.astype(np.float32)
        code = XS_F32
    es = a.itemsize
    sr, sc = a.strides[0] // es, a.strides[1] // es
    rows, cols = a.shape
    ok_row = (a.strides[1] == es or cols == 1) and sr >= cols and a.strides[0] % es == 0
    ok_col = (a.strides[0] == es or rows == 1) and sc >= rows and a.strides[1] % es == 0
    if ok_row:
        return a, code, (sr if rows > 1 else cols), 1
    if ok_col:
        return a, code, 1, (sc if cols > 1 else rows)
    a = np.ascontiguousarray(a)
    return a, code, cols, 1
