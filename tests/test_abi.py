"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol that
include/xs_b200.h declares, and -- with no CUDA device -- fails loudly instead of falling back."""
import importlib
import os
import re

import numpy as np
import pytest

from conftest import PKG_NAME, ROOT


@pytest.fixture(scope="module")
def nat():
    entry = importlib.import_module("__graft_entry__")
    entry.build()
    return importlib.import_module(PKG_NAME + "._native")


def test_header_and_library_agree(nat):
    hdr = open(os.path.join(ROOT, "include", "xs_b200.h")).read()
    declared = set(re.findall(r"XS_API\s+[\w\s\*]+?\b(xs_\w+)\s*\(", hdr))
    assert declared == set(nat.ABI_SYMBOLS), declared ^ set(nat.ABI_SYMBOLS)
    lib = nat.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.xs_abi_version() == 2


def test_ctypes_signatures_match_the_header(nat):
    """Every prototype in the header against the argtypes the ctypes mirror declares: same arity, and pointer /
    64-bit / 32-bit / double parameters in the same positions (a mismatch here corrupts the stack silently)."""
    import ctypes as C
    hdr = open(os.path.join(ROOT, "include", "xs_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    lib = nat.load()
    protos = re.findall(r"XS_API\s+[\w\s\*]+?\b(xs_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S)
    assert len(protos) == len(nat.ABI_SYMBOLS)
    for name, args in protos:
        params = [a.strip() for a in args.split(",")] if args.strip() not in ("", "void") else []
        argtypes = getattr(lib, name).argtypes
        assert argtypes is not None and len(argtypes) == len(params), (name, params, argtypes)
        for ptext, ctype in zip(params, argtypes):
            if "*" in ptext:
                kind = "ptr"
            elif re.search(r"\bint64_t\b", ptext):
                kind = "i64"
            elif re.search(r"\bdouble\b", ptext):
                kind = "f64"
            else:
                kind = "i32"
            if kind == "ptr":
                ok = ctype in (C.c_void_p, C.c_char_p) or hasattr(ctype, "contents") or issubclass(ctype, C._Pointer)
            else:
                ok = ctype is {"i64": C.c_int64, "f64": C.c_double, "i32": C.c_int}[kind]
            assert ok, (name, ptext, ctype)


def test_no_torch_types_in_abi():
    hdr = open(os.path.join(ROOT, "include", "xs_b200.h")).read()
    assert "torch" not in hdr.lower().replace("torch-extension", "") and "at::" not in hdr


def test_layout_classification(nat):
    a = np.zeros((5, 7), dtype=np.float32)
    assert nat.as_matrix(a, "a")[1:] == (nat.XS_F32, 7, 1)
    v = np.zeros((7, 5), dtype=np.float64)              # the reference's (D,N) array; v.T is (N,D) F-order
    arr, code, sr, sc = nat.as_matrix(v.T, "v")
    assert (code, sr, sc) == (nat.XS_F64, 1, 5) and arr.ctypes.data == v.ctypes.data   # no copy
    arr, code, sr, sc = nat.as_matrix(a[::2, ::2], "s")  # odd strides -> contiguous copy
    assert (sr, sc) == (4, 1) and arr.flags["C_CONTIGUOUS"]
    h = nat.as_matrix(np.zeros((3, 4), dtype=np.float16), "h")
    assert h[0].dtype == np.float32
    with pytest.raises(ValueError):
        nat.as_matrix(np.zeros(4), "x")


def test_fails_loudly_without_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        pkg.ExactIndex(np.eye(8, dtype=np.float32))
    with pytest.raises(RuntimeError):
        pkg.matching_L2(2, np.eye(8, dtype=np.float32), np.eye(8, dtype=np.float32)[:2])
    with pytest.raises(NotImplementedError):
        pkg.matching("ANNOY", 2, np.eye(8), np.eye(8))


def test_product_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, PKG_NAME)
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
