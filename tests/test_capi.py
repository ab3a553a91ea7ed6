"""The C-ABI library loads and exports every symbol include/probing_rag.h declares (no GPU:
nothing is computed here)."""
import ctypes
import os
import re

from probing_rag_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "probing_rag.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pr_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_all_declared_symbols():
    path = build.build()
    L = ctypes.CDLL(path)
    syms = declared_symbols()
    assert "pr_bm25_topk" in syms and "pr_index_create" in syms and "pr_topk_merge" in syms
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert set(_lib.SIGNATURES) == set(syms)


def test_version_and_argument_errors_without_gpu():
    L = _lib.lib()
    assert L.pr_version() == 100
    # argument validation happens before any CUDA call
    assert L.pr_topk_merge(4, 0, 2, None, None, None, None, None) == _lib.PR_EINVAL
    assert b"bad argument" in L.pr_last_error() or b"k must be" in L.pr_last_error()
    assert L.pr_bm25_workspace_bytes(None, 4, 10) == 0
    h = ctypes.c_void_p()
    assert L.pr_index_create(ctypes.byref(h), 0, 10, 0, 10, 4, 0, None, None, None) == _lib.PR_EINVAL
