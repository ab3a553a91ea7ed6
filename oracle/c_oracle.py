"""ctypes loader for oracle/libbm25_oracle.so (plain-C BM25 oracle) -- TEST INFRASTRUCTURE ONLY.
Only tests/, __graft_entry__.smoke() and bench.py may import this module."""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    subprocess.run(["make", "-s", "-C", _HERE, "libbm25_oracle.so"], check=True)
    return os.path.join(_HERE, "libbm25_oracle.so")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libbm25_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.oracle_bm25_retrieve.restype = ctypes.c_int
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def retrieve_batch(index: dict, q_indptr, q_terms, k: int, n_threads: int = 1):
    """Same contract as oracle.bm25_oracle.retrieve_batch(canonical=True)."""
    L = lib()
    indptr = np.ascontiguousarray(index["indptr"], dtype=np.int64)
    indices = np.ascontiguousarray(index["indices"], dtype=np.int32)
    data = np.ascontiguousarray(index["data"], dtype=np.float32)
    q_indptr = np.ascontiguousarray(q_indptr, dtype=np.int64)
    q_terms = np.ascontiguousarray(q_terms, dtype=np.int32)
    nq = len(q_indptr) - 1
    out_s = np.zeros((nq, k), dtype=np.float32)
    out_d = np.zeros((nq, k), dtype=np.int32)
    if k > index["num_docs"]:
        raise ValueError(f"k of {k} is larger than the number of documents {index['num_docs']}")

    def run(lo, hi):
        rc = L.oracle_bm25_retrieve(
            _p(indptr, ctypes.c_int64), _p(indices, ctypes.c_int32), _p(data, ctypes.c_float),
            ctypes.c_int32(index["num_docs"]), ctypes.c_int32(len(indptr) - 1),
            ctypes.c_int32(index.get("doc_id_base", 0)),
            _p(q_indptr, ctypes.c_int64), _p(q_terms, ctypes.c_int32),
            ctypes.c_int32(lo), ctypes.c_int32(hi), ctypes.c_int32(k),
            _p(out_s, ctypes.c_float), _p(out_d, ctypes.c_int32))
        if rc == -1:
            raise ValueError("query token id out of range")
        if rc != 0:
            raise RuntimeError(f"oracle_bm25_retrieve rc={rc}")

    n_threads = max(1, min(n_threads, nq))
    if n_threads == 1:
        run(0, nq)
    else:
        bounds = np.linspace(0, nq, n_threads + 1).astype(int)
        with ThreadPoolExecutor(n_threads) as ex:
            list(ex.map(lambda i: run(int(bounds[i]), int(bounds[i + 1])), range(n_threads)))
    return out_s, out_d
