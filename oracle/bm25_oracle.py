"""CPU oracle for the BM25 retrieval hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this module.  The product path (probing_rag_b200) never does.

PARITY UNPINNED: the arithmetic this restates lives in the third-party,
un-vendored, un-pinned packages the reference merely calls
(`llama-index-retrievers-bm25` 0.2-0.5 -> `bm25s` 0.2.x; call sites
/root/reference/exp_rag.py:236-242, 426, 428, 492 and utils.py:640).  None of
them is installed here and the reference ships no tests or golden vectors, so
this oracle is a restatement of their published algorithm
and cannot be checked against the real library in this container.

Functions and what they follow:
  build_index_loop / build_index   bm25s.BM25.index           (App. A.3-A.4)
  score_query                      bm25s.BM25.get_scores      (App. A.5)
  topk_bm25s                       bm25s.selection.topk       (App. A.6)
  topk_canonical                   canonical order: score desc, doc id asc,
                                   zero-score tail = lowest doc ids
  retrieve / retrieve_batch        BM25.retrieve + llama-index `_retrieve`
                                   (exp_rag.py:426 -> list ranked by score)
"""
from __future__ import annotations

import math
from concurrent.futures import ThreadPoolExecutor

import numpy as np

K1 = 1.5      # bm25s.BM25() default (App. A.3)
B = 0.75      # bm25s.BM25() default (App. A.3)


def idf_lucene(df: np.ndarray, n_docs: int) -> np.ndarray:
    """bm25s `_score_idf_lucene`: math.log(1 + (N - df + 0.5)/(df + 0.5)) in
    Python float64, stored to an f32 array (App. A.4).  Terms with df == 0 keep 0."""
    df = np.asarray(df, dtype=np.int64)
    out = np.zeros(df.shape, dtype=np.float32)
    # math.log per distinct df value (there are few), exactly as the library does per term
    uniq, inv = np.unique(df, return_inverse=True)
    table = np.array([math.log(1 + (n_docs - int(v) + 0.5) / (int(v) + 0.5)) if v > 0 else 0.0
                      for v in uniq], dtype=np.float64)
    out[...] = table[inv].reshape(df.shape).astype(np.float32)
    return out


def idf_lucene_scalar(df: int, n_docs: int) -> np.float32:
    return np.float32(math.log(1 + (n_docs - df + 0.5) / (df + 0.5)))


def build_index_loop(corpus_token_ids, vocab_size: int, k1: float = K1, b: float = B) -> dict:
    """Literal per-document restatement of bm25s `_build_scores_and_indices_for_matrix`
    (App. A.4); slow, used on small corpora to pin `build_index`.

    corpus_token_ids: sequence of 1-D int arrays (token ids per doc, duplicates kept).
    """
    n_docs = len(corpus_token_ids)
    doc_lens = np.array([len(d) for d in corpus_token_ids])
    avgdl = doc_lens.mean() if n_docs else np.float64(0.0)          # np.float64 scalar
    df = np.zeros(vocab_size, dtype=np.int64)
    for d in corpus_token_ids:
        df[np.unique(np.asarray(d, dtype=np.int64))] += 1
    idf = np.zeros(vocab_size, dtype=np.float32)
    for t in np.flatnonzero(df):
        idf[t] = math.log(1 + (n_docs - int(df[t]) + 0.5) / (int(df[t]) + 0.5))
    rows, cols, vals = [], [], []
    for doc_idx, d in enumerate(corpus_token_ids):
        l_d = len(d)
        voc, tf = np.unique(np.asarray(d, dtype=np.int32), return_counts=True)
        tf = tf.astype(np.float32)
        # NumPy >= 2 (NEP 50): np.float64 scalar + f32 array -> float64 array
        tfc = tf / (k1 * ((1 - b) + b * l_d / avgdl) + tf)
        w = (idf[voc] * tfc).astype(np.float32)
        rows.append(np.full(len(voc), doc_idx, dtype=np.int32))
        cols.append(voc.astype(np.int32))
        vals.append(w)
    rows = np.concatenate(rows) if rows else np.zeros(0, np.int32)
    cols = np.concatenate(cols) if cols else np.zeros(0, np.int32)
    vals = np.concatenate(vals) if vals else np.zeros(0, np.float32)
    # scipy.sparse.csc_matrix((w,(doc,voc))) -> column (term) major, rows ascending
    order = np.lexsort((rows, cols))
    indptr = np.zeros(vocab_size + 1, dtype=np.int64)
    np.cumsum(np.bincount(cols, minlength=vocab_size), out=indptr[1:])
    return {"data": vals[order], "indices": rows[order], "indptr": indptr,
            "num_docs": n_docs, "avgdl": float(avgdl), "df": df, "doc_lens": doc_lens.astype(np.int32)}


def build_index(tokens: np.ndarray, doc_lens: np.ndarray, vocab_size: int,
                k1: float = K1, b: float = B, n_docs_global: int | None = None,
                avgdl_global: float | None = None, df_global: np.ndarray | None = None,
                doc_id_base: int = 0) -> dict:
    """Vectorised equivalent of `build_index_loop` (same arithmetic, same rounding).

    tokens: flat int array, the docs' token ids back to back; doc_lens[i] tokens each.
    The *_global arguments build a doc-range shard with global statistics
    (SURVEY 8e): weights then equal the single-index weights bit for bit.
    Stored doc ids are local (0-based); `doc_id_base` is recorded for the caller.
    """
    tokens = np.asarray(tokens, dtype=np.int64)
    doc_lens = np.asarray(doc_lens, dtype=np.int64)
    n_docs = len(doc_lens)
    doc_of = np.repeat(np.arange(n_docs, dtype=np.int64), doc_lens)
    key = tokens * np.int64(max(n_docs, 1)) + doc_of
    key.sort(kind="stable")
    uniq, tf = np.unique(key, return_counts=True)
    term = (uniq // max(n_docs, 1)).astype(np.int32)
    doc = (uniq % max(n_docs, 1)).astype(np.int32)
    df_local = np.bincount(term, minlength=vocab_size).astype(np.int64)
    n_glob = n_docs if n_docs_global is None else int(n_docs_global)
    avgdl = np.float64(doc_lens.mean()) if avgdl_global is None else np.float64(avgdl_global)
    df = df_local if df_global is None else np.asarray(df_global, dtype=np.int64)
    idf = idf_lucene(df, n_glob)
    l_d = doc_lens[doc].astype(np.float64)
    tf32 = tf.astype(np.float32)
    # identical operation order to `_score_tfc_robertson`: k1*((1-b) + b*l_d/avgdl) + tf
    denom = k1 * ((1 - b) + b * l_d / avgdl) + tf32
    tfc = tf32 / denom                                            # float64
    w = (idf[term] * tfc).astype(np.float32)                       # f32*f64 -> f64 -> round once
    indptr = np.zeros(vocab_size + 1, dtype=np.int64)
    np.cumsum(df_local, out=indptr[1:])
    return {"data": w, "indices": doc, "indptr": indptr, "num_docs": n_docs,
            "num_docs_global": n_glob, "avgdl": float(avgdl), "df": df,
            "doc_lens": doc_lens.astype(np.int32), "doc_id_base": int(doc_id_base)}


def score_query(index: dict, q_terms: np.ndarray) -> np.ndarray:
    """bm25s `_compute_relevance_from_scores` (App. A.5): dense f32 accumulator,
    one `np.add.at` per query token in query order, duplicates counted each time."""
    data, indices, indptr = index["data"], index["indices"], index["indptr"]
    n_terms = len(indptr) - 1
    scores = np.zeros(index["num_docs"], dtype=np.float32)
    for t in np.asarray(q_terms, dtype=np.int64):
        if t < 0 or t >= n_terms:
            raise ValueError(f"query token id {t} out of range [0, {n_terms})")
        s, e = indptr[t], indptr[t + 1]
        np.add.at(scores, indices[s:e], data[s:e])
    return scores


def topk_bm25s(scores: np.ndarray, k: int):
    """bm25s `selection.topk` NumPy backend (App. A.6): argpartition + argsort + flip.
    Tie order and the zero-score tail are implementation-defined here."""
    ind = np.argpartition(scores, -k)[-k:]
    vals = scores[ind]
    order = np.flip(np.argsort(vals))
    return vals[order], ind[order].astype(np.int32)


def topk_canonical(scores: np.ndarray, k: int, doc_id_base: int = 0):
    """Canonical total order (SURVEY 8c i-ii): score descending, doc id ascending;
    fewer than k positive scores -> the lowest doc ids with score 0 fill the tail."""
    n = len(scores)
    if k > n:
        raise ValueError(f"k of {k} is larger than the number of documents {n}")
    if k == 0:
        return np.zeros(0, np.float32), np.zeros(0, np.int32)
    kth = np.partition(scores, n - k)[n - k]
    above = np.flatnonzero(scores > kth)
    ties = np.flatnonzero(scores == kth)[: k - len(above)]
    cand = np.concatenate([above, ties])
    order = np.lexsort((cand, -scores[cand].astype(np.float64)))
    cand = cand[order]
    return scores[cand].astype(np.float32), (cand + doc_id_base).astype(np.int32)


def retrieve(index: dict, q_terms: np.ndarray, k: int):
    """BM25.retrieve for one query (App. A.5) with the canonical top-k order."""
    if k > index["num_docs"]:
        raise ValueError(f"k of {k} is larger than the number of documents {index['num_docs']}")
    return topk_canonical(score_query(index, q_terms), k, index.get("doc_id_base", 0))


def retrieve_batch(index: dict, q_indptr: np.ndarray, q_terms: np.ndarray, k: int,
                   n_threads: int = 0, canonical: bool = True):
    """Batch of queries in CSR form.  n_threads=0 is bm25s's default serial `map`;
    n_threads>0 mirrors its ThreadPoolExecutor over queries (App. A.5)."""
    nq = len(q_indptr) - 1
    out_s = np.zeros((nq, k), dtype=np.float32)
    out_d = np.zeros((nq, k), dtype=np.int32)
    sel = topk_canonical if canonical else topk_bm25s

    def one(i):
        sc = score_query(index, q_terms[q_indptr[i]:q_indptr[i + 1]])
        if canonical:
            out_s[i], out_d[i] = sel(sc, k, index.get("doc_id_base", 0))
        else:
            out_s[i], out_d[i] = sel(sc, k)

    if k > index["num_docs"]:
        raise ValueError(f"k of {k} is larger than the number of documents {index['num_docs']}")
    if n_threads and n_threads > 1:
        with ThreadPoolExecutor(max_workers=n_threads) as ex:
            list(ex.map(one, range(nq)))
    else:
        for i in range(nq):
            one(i)
    return out_s, out_d


def merge_topk(scores_lists: np.ndarray, ids_lists: np.ndarray, k: int):
    """Merge G per-shard lists [G, B, k] into [B, k] with the canonical order
    (SURVEY 8e).  Sentinel entries (doc id < 0) are ignored."""
    g, nq, kk = scores_lists.shape
    s = np.transpose(scores_lists, (1, 0, 2)).reshape(nq, g * kk)
    d = np.transpose(ids_lists, (1, 0, 2)).reshape(nq, g * kk)
    out_s = np.zeros((nq, k), np.float32)
    out_d = np.zeros((nq, k), np.int32)
    for i in range(nq):
        valid = d[i] >= 0
        si, di = s[i][valid], d[i][valid]
        order = np.lexsort((di, -si.astype(np.float64)))[:k]
        out_s[i, :len(order)], out_d[i, :len(order)] = si[order], di[order]
    return out_s, out_d


def same_modulo_ties(scores_a, ids_a, scores_b, ids_b) -> bool:
    """True when two ranked lists agree up to permutation inside equal-score groups
    whose members all have positive score (how real bm25s output must be compared)."""
    if not np.array_equal(scores_a, scores_b):
        return False
    for v in np.unique(scores_a):
        m = scores_a == v
        if v > 0 and set(ids_a[m].tolist()) != set(ids_b[m].tolist()):
            # the last group may be cut differently at the k boundary
            if not m[-1]:
                return False
    return True
