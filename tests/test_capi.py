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
    assert L.pr_version() == 201
    # argument validation happens before any CUDA call
    assert L.pr_topk_merge(4, 0, 2, None, None, None, None, None) == _lib.PR_EINVAL
    assert b"bad argument" in L.pr_last_error() or b"k must be" in L.pr_last_error()
    assert L.pr_bm25_workspace_bytes(None, 4, 10) == 0
    assert L.pr_bm25_num_launches(None, 4, 10, -1) == -1 and L.pr_bm25_theta_offset(None, 4, 10) == 0
    assert L.pr_bm25_running_scores_offset(None, 4, 10) == 0
    assert L.pr_bm25_raise_union_bound(None, 4, 10, None, 2, None, 0, None) == _lib.PR_EINVAL
    h = ctypes.c_void_p()
    assert L.pr_index_create(ctypes.byref(h), 0, 10, 0, 10, 4, 0, None, None, None) == _lib.PR_EINVAL


def test_default_scoring_kernel_resource_budget():
    """Guard against a silently worse build of the default BM25 kernels (the three variants of <8 warps, E = 1>):
    3 CTAs x 8 warps per SM need <= 80 registers per thread; a stack frame beyond a few words means ptxas spilled
    inside the step loop (seen once with nvcc -split-compile: same source, 104 bytes of stack, -21% queries/s);
    and a YIELD in front of the epoch test of the step loop (ptxas adds one for some harmless-looking code
    shapes) costs 2% queries/s."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        import pytest
        pytest.skip("cuobjdump not on PATH")
    lib = build.build()
    out = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    lines = out.splitlines()
    names = sorted({m.group(0) for l in lines for m in [re.search(r"_ZN3prl16bm25_lean_kernelILi8ELi1ELi[012]E\w*", l)] if m})
    assert len(names) == 3, names
    for name in names:
        hit = next(lines[i + 1] for i, l in enumerate(lines) if name in l and i + 1 < len(lines))
        m = re.search(r"REG:(\d+) STACK:(\d+)", hit)
        assert m, hit
        assert int(m.group(1)) <= 80 and int(m.group(2)) <= 48, (name, hit)
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", name, lib], capture_output=True, text=True).stdout.splitlines()
        ops = [l for l in sass if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]
        assert len(ops) > 1000, "kernel SASS not found"
        for i, l in enumerate(ops):
            if "YIELD" in l:
                assert not any("0x40000000" in x or "0x20000000" in x for x in ops[max(0, i - 3):i]), f"{name}: YIELD in the step loop"


def test_pool_accumulate_argument_errors_without_gpu():
    """pr_pool_accumulate validates its arguments before any CUDA call."""
    L = _lib.lib()
    acc = ctypes.c_void_p(0x1000)
    act = ctypes.c_void_p(0x2000)
    assert L.pr_pool_accumulate(None, 4, 6, 0, 2048, act, 0, 4, 1, 2048, 2048, None, None) == _lib.PR_EINVAL
    assert L.pr_pool_accumulate(acc, 4, 6, 6, 2048, act, 0, 4, 1, 2048, 2048, None, None) == _lib.PR_EINVAL      # slot out of range
    assert L.pr_pool_accumulate(acc, 4, 6, 0, 2046, act, 0, 4, 1, 2046, 2046, None, None) == _lib.PR_EINVAL      # d_model % 4
    assert L.pr_pool_accumulate(acc, 4, 6, 0, 2048, act, 3, 4, 1, 2048, 2048, None, None) == _lib.PR_EINVAL      # dtype
    assert L.pr_pool_accumulate(acc, 4, 6, 0, 2048, act, 0, 5, 1, 2048, 2048, None, None) == _lib.PR_EINVAL      # more rows than the accumulator
    assert L.pr_pool_accumulate(acc, 4, 6, 0, 2048, act, 1, 4, 1, 2050, 2048, None, None) == _lib.PR_EINVAL      # stride not vectorisable
    assert L.pr_pool_accumulate(acc, 4, 6, 0, 2048, act, 0, 0, 1, 2048, 2048, None, None) == _lib.PR_OK          # nothing to do
