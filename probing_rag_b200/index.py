"""BM25 inverted index in HBM + the handle libprobingrag.so scores it through.

The arrays are exactly what `bm25s.BM25.index()` builds inside
`BM25Retriever.from_defaults` (/root/reference/exp_rag.py:242):
term-major CSC with precomputed per-(term, doc) weights

    data    f32[nnz]   w = idf(t) * tf / (k1*((1-b) + b*|d|/avgdl) + tf),  k1=1.5, b=0.75,
                       idf(t) = ln(1 + (N - df + 0.5)/(df + 0.5))   ("lucene")
    indices i32[nnz]   doc ids, ascending inside each term
    indptr  i64[V+1]

but resident on the GPU, built on the GPU, persisted as a flat binary (the reference
re-tokenises and re-indexes the corpus at every start), and shardable by doc range with
GLOBAL N / avgdl / df so a shard's weights equal the single index's bit for bit.
"""
from __future__ import annotations

import ctypes
import json
import math
import os

import numpy as np
import torch

from . import _lib

K1 = 1.5    # bm25s.BM25() defaults, SURVEY App. A.3
B = 0.75


def idf_lucene_table(df: np.ndarray, n_docs: int) -> np.ndarray:
    """idf[t] = math.log(1 + (N - df + 0.5)/(df + 0.5)) in float64, stored f32 (App. A.4).
    Evaluated once per distinct df value with the C library's log, like the CPU library."""
    df = np.asarray(df, dtype=np.int64)
    uniq, inv = np.unique(df, return_inverse=True)
    table = np.array([math.log(1 + (n_docs - int(v) + 0.5) / (int(v) + 0.5)) if v > 0 else 0.0
                      for v in uniq], dtype=np.float64)
    return table[inv].reshape(df.shape).astype(np.float32)


def count_postings(tokens: torch.Tensor, doc_lens: torch.Tensor, vocab: int):
    """(term i32[nnz], doc i32[nnz], tf i32[nnz], df_local i64[V]) sorted by (term, doc):
    one device radix sort of (term, doc) keys, then run-length counting."""
    dev = tokens.device
    n_docs = doc_lens.numel()
    shift = max(1, int(n_docs - 1).bit_length())
    if shift + max(1, int(vocab - 1).bit_length()) > 62:
        raise ValueError("vocab x n_docs too large for a 64-bit sort key")
    doc_of = torch.repeat_interleave(torch.arange(n_docs, device=dev, dtype=torch.int32),
                                     doc_lens.to(torch.int64))
    key = (tokens.to(torch.int64) << shift) | doc_of.to(torch.int64)
    del doc_of
    key = torch.sort(key).values
    uniq, tf = torch.unique_consecutive(key, return_counts=True)
    del key
    term = (uniq >> shift).to(torch.int32)
    doc = (uniq & ((1 << shift) - 1)).to(torch.int32)
    del uniq
    df_local = torch.bincount(term, minlength=vocab)
    return term, doc, tf.to(torch.int32), df_local


def bm25_weights(term: torch.Tensor, doc: torch.Tensor, tf: torch.Tensor, doc_lens: torch.Tensor,
                 idf: torch.Tensor, avgdl: float, k1: float = K1, b: float = B,
                 chunk: int = 1 << 26) -> torch.Tensor:
    """w = f32( f64(idf_f32[t]) * ( tf / (k1*((1-b) + b*|d|/avgdl) + tf) ) ), every operation
    a single correctly-rounded float64 op in the order bm25s's `_score_tfc_robertson` applies
    them under NumPy >= 2; the product is rounded to f32 once."""
    out = torch.empty(term.numel(), dtype=torch.float32, device=term.device)
    one_minus_b = 1 - b
    for s in range(0, term.numel(), chunk):
        e = min(s + chunk, term.numel())
        l_d = doc_lens[doc[s:e].long()].to(torch.float64)
        tff = tf[s:e].to(torch.float32).to(torch.float64)
        x = l_d.mul_(b).div_(avgdl).add_(one_minus_b).mul_(k1).add_(tff)   # k1*((1-b)+b*l_d/avgdl)+tf
        tfc = tff.div_(x)
        out[s:e] = idf[term[s:e].long()].to(torch.float64).mul_(tfc).to(torch.float32)
    return out


class BM25Index:
    """Device-resident index + pr_index_t handle.  All query entry points take token ids."""

    def __init__(self, indptr: torch.Tensor, doc_ids: torch.Tensor, weights: torch.Tensor,
                 n_docs: int, n_docs_global: int | None = None, doc_id_base: int = 0,
                 meta: dict | None = None, aux_budget_bytes: int | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("BM25Index needs a CUDA device: the retrieval hot path has no CPU fallback")
        if indptr.device.type != "cuda":
            raise ValueError("index arrays must live on a CUDA device")
        self.indptr = indptr.to(torch.int64).contiguous()
        # 128-bit loads: keep the posting arrays readable up to nnz rounded up to 4 elements
        nnz = doc_ids.numel()

        def padded(t, dtype):
            buf = torch.zeros((nnz + 3) // 4 * 4 + 4, dtype=dtype, device=t.device)
            buf[:nnz] = t
            return buf[:nnz]
        self.doc_ids = padded(doc_ids, torch.int32)
        self.weights = padded(weights, torch.float32)
        self.n_docs = int(n_docs)
        self.n_docs_global = int(n_docs if n_docs_global is None else n_docs_global)
        self.doc_id_base = int(doc_id_base)
        self.n_terms = self.indptr.numel() - 1
        self.nnz = self.doc_ids.numel()
        self.device = self.indptr.device
        self.meta = dict(meta or {})
        self._ws: dict = {}
        self._pinned: dict = {}
        self._handle = ctypes.c_void_p()
        L = _lib.lib()
        with torch.cuda.device(self.device):
            _lib.check(L.pr_index_create(
                ctypes.byref(self._handle), self.device.index or 0, self.n_docs_global, self.doc_id_base,
                self.n_docs, self.n_terms, self.nnz, self.indptr.data_ptr(),
                self.doc_ids.data_ptr() if self.nnz else None, self.weights.data_ptr() if self.nnz else None))
            # per-index structures of the scoring kernel: posting offsets at every 2048-doc boundary
            # for the frequent terms (a fifth of the budget) and the hot posting stream (the rest);
            # 12 bytes per posting by default (min 64 MB, max 48 GB)
            budget = aux_budget_bytes if aux_budget_bytes is not None else \
                int(min(max(self.nnz * 12, 64 << 20), 48 << 30))
            nbytes = int(L.pr_index_aux_bytes(self._handle, budget))
            self._aux = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            _lib.check(L.pr_index_build_aux(self._handle, self._aux.data_ptr(), nbytes,
                                            torch.cuda.current_stream(self.device).cuda_stream))

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            try:
                _lib.lib().pr_index_destroy(h)
            except Exception:
                pass
            self._handle = None

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_arrays(cls, data, indices, indptr, n_docs, device="cuda", **kw) -> "BM25Index":
        """From bm25s-style host arrays {data, indices, indptr, num_docs} (App. A.4)."""
        dev = torch.device(device)
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)
        return cls(t(indptr, torch.int64), t(indices, torch.int32), t(data, torch.float32), n_docs, **kw)

    @classmethod
    def from_tokens(cls, tokens: torch.Tensor, doc_lens: torch.Tensor, vocab: int,
                    n_docs_global: int | None = None, avgdl_global: float | None = None,
                    df_global: torch.Tensor | None = None, doc_id_base: int = 0,
                    k1: float = K1, b: float = B) -> "BM25Index":
        """Build on the device of `tokens` (flat i32 token ids, docs back to back)."""
        n_docs = doc_lens.numel()
        term, doc, tf, df_local = count_postings(tokens, doc_lens, vocab)
        n_glob = n_docs if n_docs_global is None else int(n_docs_global)
        avgdl = float(doc_lens.double().mean().item()) if avgdl_global is None else float(avgdl_global)
        df = df_local if df_global is None else df_global
        idf = torch.from_numpy(idf_lucene_table(df.cpu().numpy(), n_glob)).to(tokens.device)
        w = bm25_weights(term, doc, tf, doc_lens, idf, avgdl, k1, b)
        indptr = torch.zeros(vocab + 1, dtype=torch.int64, device=tokens.device)
        torch.cumsum(df_local, 0, out=indptr[1:])
        return cls(indptr, doc, w, n_docs, n_glob, doc_id_base,
                   meta={"avgdl": avgdl, "k1": k1, "b": b, "df": df})

    # ------------------------------------------------------------------ persistence
    def save(self, path: str) -> None:
        """Flat binary + JSON header; replaces re-indexing at every start (exp_rag.py:241-242)."""
        os.makedirs(path, exist_ok=True)
        hdr = {"format": "probing-rag-b200-csr-v1", "n_docs": self.n_docs, "n_docs_global": self.n_docs_global,
               "doc_id_base": self.doc_id_base, "n_terms": self.n_terms, "nnz": self.nnz,
               "meta": {k: v for k, v in self.meta.items() if isinstance(v, (int, float, str))}}
        with open(os.path.join(path, "index.json"), "w") as f:
            json.dump(hdr, f)
        self.indptr.cpu().numpy().tofile(os.path.join(path, "indptr.i64"))
        self.doc_ids.cpu().numpy().tofile(os.path.join(path, "doc_ids.i32"))
        self.weights.cpu().numpy().tofile(os.path.join(path, "weights.f32"))

    @classmethod
    def load(cls, path: str, device="cuda", doc_range: tuple[int, int] | None = None) -> "BM25Index":
        """Load a saved index.  `doc_range=(lo, hi)` cuts the doc-range shard [lo, hi) out of a saved WHOLE-corpus
        index on the device: a term's postings are ascending in doc id, so the shard's list is a
        contiguous piece of it, and the weights already carry the global N / avgdl / df -- they are the shard's
        weights bit for bit."""
        with open(os.path.join(path, "index.json")) as f:
            hdr = json.load(f)
        if hdr.get("format") != "probing-rag-b200-csr-v1":
            raise ValueError(f"{path}: not a probing-rag-b200 index")
        dev = torch.device(device)
        ld = lambda name, dt: torch.from_numpy(np.fromfile(os.path.join(path, name), dtype=dt)).to(dev)
        indptr, doc_ids, weights = ld("indptr.i64", np.int64), ld("doc_ids.i32", np.int32), ld("weights.f32", np.float32)
        if doc_range is None:
            return cls(indptr, doc_ids, weights, hdr["n_docs"], hdr["n_docs_global"], hdr["doc_id_base"], meta=hdr.get("meta"))
        lo, hi = int(doc_range[0]), int(doc_range[1])
        if hdr["doc_id_base"] != 0 or hdr["n_docs"] != hdr["n_docs_global"]:
            raise ValueError(f"{path} holds a shard already; doc_range cuts shards out of a whole-corpus index")
        if not 0 <= lo <= hi <= hdr["n_docs"]:
            raise ValueError(f"doc_range [{lo}, {hi}) outside [0, {hdr['n_docs']})")
        keep = (doc_ids >= lo) & (doc_ids < hi)
        # postings kept per term = difference of the running count of kept postings at the term boundaries
        run = torch.zeros(doc_ids.numel() + 1, dtype=torch.int64, device=dev)
        torch.cumsum(keep, 0, out=run[1:])
        s_indptr = run[indptr]
        del run
        s_docs = (doc_ids[keep] - lo).to(torch.int32)
        s_w = weights[keep]
        del keep, doc_ids, weights
        return cls(s_indptr, s_docs, s_w, hi - lo, hdr["n_docs_global"], lo, meta=hdr.get("meta"))

    # ------------------------------------------------------------------ tuning
    def set_tuning(self, **kw) -> None:
        t = _lib.Tuning(**{k: int(v) for k, v in kw.items()})
        _lib.check(_lib.lib().pr_index_set_tuning(self._handle, ctypes.byref(t)))
        self._ws.clear()

    def get_tuning(self) -> dict:
        t = _lib.Tuning()
        _lib.check(_lib.lib().pr_index_get_tuning(self._handle, ctypes.byref(t)))
        return {k: getattr(t, k) for k, _ in _lib.Tuning._fields_}

    def set_profiling(self, enable: bool) -> None:
        _lib.check(_lib.lib().pr_index_set_profiling(self._handle, int(enable)))

    def profile(self):
        """(device ms summed over the scoring-kernel launches of the last topk, launch count)."""
        ms, n = ctypes.c_float(), ctypes.c_int32()
        _lib.check(_lib.lib().pr_bm25_profile(self._handle, ctypes.byref(ms), ctypes.byref(n)))
        return float(ms.value), int(n.value)

    def aux_info(self) -> dict:
        info = _lib.AuxInfo()
        _lib.check(_lib.lib().pr_index_aux_info(self._handle, ctypes.byref(info)))
        out = {k: int(getattr(info, k)) for k, _ in _lib.AuxInfo._fields_}
        out["aux_bytes"] = int(self._aux.numel())
        return out

    @property
    def last_launches(self) -> int:
        return int(_lib.lib().pr_bm25_last_launches(self._handle))

    # ------------------------------------------------------------------ queries
    def _workspace(self, n_queries: int, k: int) -> torch.Tensor:
        """One scratch buffer, grown on demand: gate-compacted batches change size on every call, and the library
        re-initialises whatever it is handed (any buffer of at least pr_bm25_workspace_bytes works)."""
        key = (n_queries, k)
        nbytes = self._ws.get(key)
        if nbytes is None:
            nbytes = int(_lib.lib().pr_bm25_workspace_bytes(self._handle, n_queries, k))
            if nbytes == 0:
                raise ValueError(f"k must be in [1, {_lib.PR_MAX_K}] (got {k})")
            if len(self._ws) > 4096:
                self._ws.clear()
            self._ws[key] = nbytes
        buf = getattr(self, "_ws_buf", None)
        if buf is None or buf.numel() < nbytes:
            self._ws_buf = buf = torch.empty(nbytes + nbytes // 4, dtype=torch.uint8, device=self.device)
        return buf

    def _check_query_batch(self, q_indptr: torch.Tensor, q_terms: torch.Tensor, k: int, out):
        """Everything the C ABI cannot see from raw pointers: dtype, device, layout, sizes.  (The CSR's
        CONTENT -- monotone offsets inside q_terms, term ids in range -- is validated on the device.)"""
        if k > self.n_docs_global:
            raise ValueError(f"k of {k} is larger than the number of documents {self.n_docs_global}")
        if not (1 <= k <= _lib.PR_MAX_K):
            raise ValueError(f"k must be in [1, {_lib.PR_MAX_K}] (got {k})")
        if q_indptr.dtype != torch.int64 or q_terms.dtype != torch.int32:
            raise ValueError("q_indptr must be int64 and q_terms int32")
        if q_indptr.dim() != 1 or q_terms.dim() != 1 or q_indptr.numel() < 1:
            raise ValueError("q_indptr must be a 1-D array of n_queries+1 offsets and q_terms 1-D")
        for name, t in (("q_indptr", q_indptr), ("q_terms", q_terms)):
            if not t.is_cuda or t.device != self.device:
                raise ValueError(f"{name} must live on the index's device {self.device} (got {t.device})")
        nq = q_indptr.numel() - 1
        if out is None:
            out = (torch.empty((nq, k), dtype=torch.float32, device=self.device),
                   torch.empty((nq, k), dtype=torch.int32, device=self.device))
        else:
            for t, dt in zip(out, (torch.float32, torch.int32)):
                if (t.dtype != dt or tuple(t.shape) != (nq, k) or not t.is_cuda or t.device != self.device
                        or not t.is_contiguous()):
                    raise ValueError(f"out must be contiguous (float32[{nq},{k}], int32[{nq},{k}]) tensors on {self.device}")
        return q_indptr.contiguous(), q_terms.contiguous(), nq, out

    def topk(self, q_indptr: torch.Tensor, q_terms: torch.Tensor, k: int, out=None, check_status=True,
             exchange=None, exchange_rounds: int | None = None, list_exchange=None, list_rounds: int = 0):
        """Device CSR query batch -> (scores f32[B,k], doc_ids i32[B,k]) on the device.
        Enqueues on the current stream; with check_status the stream is synchronised and a bad
        term id / inconsistent CSR raises ValueError like bm25s does (App. A.5).

        exchange: optional callable(theta f32[B]) run between the launches of the call (doc-range
        shards: an all-reduce(MAX) of the per-query k-th-score bounds over the ranks).
        Every rank must join the same number of exchanges: `exchange_rounds` (>= this shard's launches - 1)
        is what the longest shard needs, see `ShardedBM25`.

        list_exchange: optional callable(run_scores f32[B,k]) -> gathered f32[G,B,k], run after each of the first
        `list_rounds` launches (all of them on every rank, also on a shard with fewer launches): the all-gather of
        the shards' running score lists; every query's bound is then raised to the k-th largest score of the union
        (pr_bm25_raise_union_bound).  Works with private and with peer-shared thresholds."""
        q_indptr, q_terms, nq, out = self._check_query_batch(q_indptr, q_terms, k, out)
        if nq == 0:
            return out
        ws = self._workspace(nq, k)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        L = _lib.lib()
        args = (self._handle, nq, q_indptr.data_ptr(), q_terms.data_ptr() if q_terms.numel() else None, q_terms.numel(), k,
                out[0].data_ptr(), out[1].data_ptr(), ws.data_ptr(), ws.numel())
        with torch.cuda.device(self.device):
            if exchange is None and (list_exchange is None or list_rounds < 1):
                _lib.check(L.pr_bm25_topk(*args, stream))
            elif exchange is None:
                n_launch = int(L.pr_bm25_num_launches(self._handle, nq, k, -1))
                off = int(L.pr_bm25_running_scores_offset(self._handle, nq, k))
                run_s = ws[off:off + 4 * nq * k].view(torch.float32).view(nq, k)
                li = done = 0
                while li < n_launch:
                    if done < list_rounds:
                        _lib.check(L.pr_bm25_topk_range(*args, li, li + 1, stream))
                        li += 1
                        gathered = list_exchange(run_s)
                        done += 1
                        if li < n_launch:             # (behind the call's last launch there is nothing left to filter)
                            _lib.check(L.pr_bm25_raise_union_bound(self._handle, nq, k, gathered.data_ptr(), gathered.shape[0],
                                                                   ws.data_ptr(), ws.numel(), stream))
                    else:                             # the launches behind the exchanged ones go out as one range
                        _lib.check(L.pr_bm25_topk_range(*args, li, n_launch, stream))
                        li = n_launch
                for _ in range(list_rounds - done):   # a shard with fewer launches still joins the rounds the others run
                    list_exchange(run_s)
            else:
                n_launch = int(L.pr_bm25_num_launches(self._handle, nq, k, -1))
                off = int(L.pr_bm25_theta_offset(self._handle, nq, k))
                theta = ws[off:off + 4 * nq].view(torch.float32)
                rounds = n_launch - 1 if exchange_rounds is None else max(int(exchange_rounds), n_launch - 1)
                done = 0
                for li in range(n_launch):
                    _lib.check(L.pr_bm25_topk_range(*args, li, li + 1, stream))
                    if done < rounds:
                        exchange(theta)
                        done += 1
                for _ in range(rounds - done):              # a short shard still joins the rounds the longest one needs
                    exchange(theta)
            if check_status and nq:
                _lib.check(L.pr_bm25_status(ws.data_ptr(), stream))
        return out

    def num_launches(self, n_queries: int, k: int, n_docs: int | None = None) -> int:
        """Scoring launches of a call of this shape (on a shard of n_docs documents with this index's tuning)."""
        n = int(_lib.lib().pr_bm25_num_launches(self._handle, n_queries, k, -1 if n_docs is None else int(n_docs)))
        if n < 0:
            raise ValueError(f"k must be in [1, {_lib.PR_MAX_K}] (got {k})")
        return n

    def check_status(self, n_queries: int, k: int) -> None:
        """Synchronise the current stream and raise what the last topk of this shape flagged on the device."""
        _lib.check(_lib.lib().pr_bm25_status(self._workspace(n_queries, k).data_ptr(),
                                              torch.cuda.current_stream(self.device).cuda_stream))

    def topk_host(self, q_indptr: np.ndarray, q_terms: np.ndarray, k: int, exchange=None):
        """Host CSR query batch -> host (scores, doc_ids): pinned H2D copy of the queries,
        scoring, pinned D2H copy of the ranked lists.  Returns (scores, ids, h2d_bytes, d2h_bytes);
        the arrays are the caller's own (copied out of the pinned staging buffers)."""
        d_qi, d_qt, n_terms = self._stage_queries(q_indptr, q_terms)
        s, d = self.topk(d_qi, d_qt, k, check_status=False, exchange=exchange)
        h_s, h_d = self._unstage_lists(s, d)
        _lib.check(_lib.lib().pr_bm25_status(self._workspace(len(q_indptr) - 1, k).data_ptr(),
                                              torch.cuda.current_stream(self.device).cuda_stream))
        return h_s.numpy().copy(), h_d.numpy().copy(), (len(q_indptr)) * 8 + n_terms * 4, (len(q_indptr) - 1) * k * 8

    def _stage_queries(self, q_indptr: np.ndarray, q_terms: np.ndarray):
        """Host CSR -> device tensors through pinned staging buffers (grown on demand, reused)."""
        q_indptr = np.asarray(q_indptr)
        q_terms = np.asarray(q_terms)
        if q_indptr.ndim != 1 or len(q_indptr) < 1 or q_terms.ndim != 1:
            raise ValueError("q_indptr must be a 1-D array of n_queries+1 offsets and q_terms 1-D")
        n_i, n_t = len(q_indptr), len(q_terms)
        st = self._pinned
        if st.get("qi") is None or st["qi"].numel() < n_i:
            st["qi"] = torch.empty(max(n_i, 1024), dtype=torch.int64).pin_memory()
        if st.get("qt") is None or st["qt"].numel() < max(n_t, 1):
            st["qt"] = torch.empty(max(n_t, 4096), dtype=torch.int32).pin_memory()
        st["qi"].numpy()[:n_i] = q_indptr
        st["qt"].numpy()[:n_t] = q_terms
        d_qi = st["qi"][:n_i].to(self.device, non_blocking=True)
        d_qt = st["qt"][:n_t].to(self.device, non_blocking=True)
        return d_qi, d_qt, n_t

    def _unstage_lists(self, s: torch.Tensor, d: torch.Tensor):
        """Device [B,k] lists -> pinned host tensors (views of the staging buffers; valid until the next call)."""
        n = s.numel()
        st = self._pinned
        if st.get("s") is None or st["s"].numel() < max(n, 1):
            st["s"] = torch.empty(max(n, 4096), dtype=torch.float32).pin_memory()
            st["d"] = torch.empty(max(n, 4096), dtype=torch.int32).pin_memory()
        h_s = st["s"][:n].view(s.shape)
        h_d = st["d"][:n].view(d.shape)
        h_s.copy_(s, non_blocking=True)
        h_d.copy_(d, non_blocking=True)
        return h_s, h_d

    def algorithmic_bytes(self, q_indptr, q_terms, k: int) -> int:
        """sum_q (8 * sum_{t in q} df_shard(t) + 8*k)."""
        qt = torch.as_tensor(q_terms).to(self.device).long()
        df = self.indptr[1:] - self.indptr[:-1]
        return int(8 * df[qt].sum().item() + 8 * k * (len(q_indptr) - 1))


def merge_topk(scores: torch.Tensor, ids: torch.Tensor):
    """[G, B, k] per-shard ranked lists -> [B, k] (pr_topk_merge)."""
    g, nq, k = scores.shape
    scores = scores.contiguous().float()
    ids = ids.contiguous().to(torch.int32)
    out_s = torch.empty((nq, k), dtype=torch.float32, device=scores.device)
    out_d = torch.empty((nq, k), dtype=torch.int32, device=scores.device)
    with torch.cuda.device(scores.device):
        _lib.check(_lib.lib().pr_topk_merge(nq, k, g, scores.data_ptr(), ids.data_ptr(), out_s.data_ptr(),
                                            out_d.data_ptr(), torch.cuda.current_stream(scores.device).cuda_stream))
    return out_s, out_d
