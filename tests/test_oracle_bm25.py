"""The CPU oracle against its fixtures and against itself (no GPU)."""
import math
import os

import numpy as np
import pytest

from oracle import bm25_oracle as bo
from oracle import c_oracle as co
from probing_rag_b200 import synth


def test_known_answer_three_docs():
    # hand-evaluated App. A.3-A.4 formulas, float64 Python -> f32
    docs = [np.array([0, 1, 1]), np.array([1, 2]), np.array([2, 2, 2, 0])]
    idx = bo.build_index_loop(docs, 3)
    n, avgdl = 3, 3.0
    def w(df, tf, ld):
        idf = np.float32(math.log(1 + (n - df + 0.5) / (df + 0.5)))
        return np.float32(float(idf) * (tf / (1.5 * ((1 - 0.75) + 0.75 * ld / avgdl) + tf)))
    # term 0: docs 0 (tf1,len3), 2 (tf1,len4); term 1: docs 0 (tf2), 1 (tf1,len2); term 2: docs 1, 2 (tf3)
    expect = [w(2, 1, 3), w(2, 1, 4), w(2, 2, 3), w(2, 1, 2), w(2, 1, 2), w(2, 3, 4)]
    assert idx["indptr"].tolist() == [0, 2, 4, 6]
    assert idx["indices"].tolist() == [0, 2, 0, 1, 1, 2]
    assert idx["data"].tolist() == [float(x) for x in expect]
    s, d = bo.retrieve(idx, np.array([1, 2]), 3)
    sc = np.zeros(3, np.float32)
    sc[0] += expect[2]; sc[1] += expect[3]; sc[1] += expect[4]; sc[2] += expect[5]
    order = sorted(range(3), key=lambda i: (-sc[i], i))
    assert d.tolist() == order and s.tolist() == [float(sc[i]) for i in order]


def test_golden_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "bm25_golden.npz"))
    idx = bo.build_index(g["tokens"], g["doc_lens"], int(g["vocab"]))
    assert np.array_equal(idx["data"], g["data"])
    assert np.array_equal(idx["indices"], g["indices"])
    assert np.array_equal(idx["indptr"], g["indptr"])
    s, d = bo.retrieve_batch(idx, g["q_indptr"], g["q_terms"], int(g["k"]))
    assert np.array_equal(s, g["scores"]) and np.array_equal(d, g["ids"])
    s2, d2 = co.retrieve_batch(idx, g["q_indptr"], g["q_terms"], int(g["k"]))
    assert np.array_equal(s2, g["scores"]) and np.array_equal(d2, g["ids"])


def test_vectorised_builder_equals_literal_loop():
    toks, lens = synth.corpus_np(3000, 1 << 12)
    off = np.concatenate([[0], np.cumsum(lens)])
    docs = [toks[off[i]:off[i + 1]] for i in range(len(lens))]
    a, b = bo.build_index_loop(docs, 1 << 12), bo.build_index(toks, lens, 1 << 12)
    for key in ("data", "indices", "indptr"):
        assert np.array_equal(a[key], b[key]), key


def test_canonical_topk_vs_bm25s_selection(small_corpus):
    idx = small_corpus["index"]
    qi, qt = small_corpus["q_indptr"], small_corpus["q_terms"]
    for q in range(0, 200, 7):
        sc = bo.score_query(idx, qt[qi[q]:qi[q + 1]])
        s1, d1 = bo.topk_canonical(sc, 10)
        s2, d2 = bo.topk_bm25s(sc, 10)
        assert np.array_equal(s1, s2)                 # same score multiset, descending
        assert bo.same_modulo_ties(s1, d1, s2, d2)
        assert np.all(np.diff(s1) <= 0)
        for i in range(9):
            if s1[i] == s1[i + 1]:
                assert d1[i] < d1[i + 1]


def test_zero_score_tail_and_empty_query():
    docs = [np.array([0]), np.array([1]), np.array([1, 0]), np.array([2]), np.array([2])]
    idx = bo.build_index_loop(docs, 4)
    s, d = bo.retrieve(idx, np.array([0], dtype=np.int32), 4)
    assert d[:2].tolist() == [0, 2] and d[2:].tolist() == [1, 3] and s[2:].tolist() == [0.0, 0.0]
    s, d = bo.retrieve(idx, np.array([], dtype=np.int32), 3)
    assert d.tolist() == [0, 1, 2] and s.tolist() == [0.0, 0.0, 0.0]
    s, d = bo.retrieve(idx, np.array([3], dtype=np.int32), 2)      # term with df == 0
    assert d.tolist() == [0, 1]
    with pytest.raises(ValueError):
        bo.retrieve(idx, np.array([0]), 6)
    with pytest.raises(ValueError):
        bo.retrieve(idx, np.array([4]), 2)


def test_c_oracle_matches_numpy(small_corpus):
    idx, qi, qt = small_corpus["index"], small_corpus["q_indptr"], small_corpus["q_terms"]
    s1, d1 = bo.retrieve_batch(idx, qi[:201], qt, 10)
    s2, d2 = co.retrieve_batch(idx, qi[:201], qt, 10, n_threads=4)
    assert np.array_equal(s1, s2) and np.array_equal(d1, d2)


def test_doc_range_shards_merge_to_single_index():
    n_docs, vocab = 6000, 1 << 12
    toks, lens = synth.corpus_np(n_docs, vocab)
    full = bo.build_index(toks, lens, vocab)
    qi, qt = synth.queries_np(64, vocab, full["df"])
    ref_s, ref_d = bo.retrieve_batch(full, qi, qt, 10)
    off = np.concatenate([[0], np.cumsum(lens, dtype=np.int64)])
    for g in (2, 3, 8):
        per = -(-n_docs // g)
        ss, dd = [], []
        for r in range(g):
            lo, hi = r * per, min((r + 1) * per, n_docs)
            sh = bo.build_index(toks[off[lo]:off[hi]], lens[lo:hi], vocab, n_docs_global=n_docs,
                                avgdl_global=full["avgdl"], df_global=full["df"], doc_id_base=lo)
            s, d = bo.retrieve_batch(sh, qi, qt, 10)
            ss.append(s); dd.append(d)
        ms, md = bo.merge_topk(np.stack(ss), np.stack(dd), 10)
        assert np.array_equal(ms, ref_s) and np.array_equal(md, ref_d), g


def test_bm25s_readme_quickstart_scores():
    """Sanity anchor on the variant constants (NOT a pin: recalled from the published bm25s README quickstart, which
    cannot be fetched offline).  Corpus of four sentences, query "does the fish purr like a cat?", tokenised with
    bm25s' English stop-word list and no stemming; the README prints the two hits as `score: 1.06` and `score: 0.48`.
    Lucene idf with k1 = 1.5, b = 0.75 reproduces both to the printed digits; Robertson idf would give 0.75 / 0.34,
    BM25+ / BM25L other values again."""
    from probing_rag_b200 import text
    corpus = ["a cat is a feline and likes to purr", "a dog is the human's best friend and loves to play",
              "a bird is a beautiful animal that can fly", "a fish is a creature that lives in water and swims"]
    docs = [text.split_tokens(d) for d in corpus]
    assert [len(d) for d in docs] == [4, 6, 5, 5]                 # after the 33-word stop list
    vocab = {}
    for d in docs:
        for w in d:
            vocab.setdefault(w, len(vocab))
    idx = bo.build_index_loop([np.array([vocab[w] for w in d]) for d in docs], len(vocab))
    q = np.array([vocab[w] for w in text.split_tokens("does the fish purr like a cat?") if w in vocab], dtype=np.int32)
    s = bo.score_query(idx, q)
    assert f"{s[0]:.2f}" == "1.06" and f"{s[3]:.2f}" == "0.48" and s[1] == 0 and s[2] == 0
    sc, ids = bo.topk_canonical(s, 2)
    assert ids.tolist() == [0, 3]
