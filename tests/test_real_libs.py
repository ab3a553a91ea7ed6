"""Opportunistic parity against the REAL third-party packages the reference calls
(/root/reference/exp_rag.py:236-242, 426: llama-index BM25Retriever -> bm25s -> PyStemmer).  None of them is
installed in the build container or the GPU image (no network), so every test here skips there; on any box
that has them this is the pin the oracle lacks (oracle/bm25_oracle.py header: PARITY UNPINNED).  bm25s's
top-k order inside equal-score groups is implementation-defined, so ranked lists are
compared modulo permutation inside tie groups."""
import numpy as np
import pytest

from oracle import bm25_oracle as bo
from probing_rag_b200.text import STOPWORDS_EN, BuiltinStemmer, Vocabulary, porter2_stem, split_tokens

CORPUS = [
    "a cat is a feline and likes to purr",
    "a dog is the human's best friend and loves to play",
    "a bird is a beautiful animal that can fly",
    "a fish is a creature that lives in water and swims",
    "The cats were running through the gardens of the national library, chasing birds",
    "Retrieval augmented generation probes the hidden states of a language model",
    "Paris is the capital city of France; the Eiffel tower stands by the river Seine",
    "London bridge crosses the river Thames in England",
]
QUERIES = ["does the fish purr like a cat?", "capital city of France", "running cats and birds", "river bridge tower",
           "unknownword"]

WORDS = ("consign consigned consigning consignment consist consisted consistency consistent consistently consisting "
         "consists consolation consolations consolatory console consoled consoles consolidate consolidated knack "
         "knackeries knacks knag knave knaves knavish kneaded kneading knee kneel kneeled kneeling kneels knees knell "
         "knelt knew knick knif knife knight knightly knights knit knits knitted knitting knives knob knobs knock "
         "generously generate generation communication communities caresses ponies ties cries gas gaps kiwis agreed "
         "feed hopping hoping luxuriated relational conditional rational happy cry by say sky dying news succeed "
         "running national electricity wikipedia retrieval probing augmented skies ugly early only singly idly").split()


def test_porter2_restatement_equals_pystemmer():
    Stemmer = pytest.importorskip("Stemmer")
    st = Stemmer.Stemmer("english")
    text_words = sorted({w for t in CORPUS + QUERIES for w in split_tokens(t)} | set(WORDS))
    bad = [(w, porter2_stem(w), s) for w, s in zip(text_words, st.stemWords(text_words)) if porter2_stem(w) != s]
    assert not bad, bad


def test_tokenizer_and_stop_list_equal_bm25s():
    bm25s = pytest.importorskip("bm25s")
    from bm25s.tokenization import STOPWORDS_EN as real_stop
    assert set(real_stop) == set(STOPWORDS_EN)
    tok = bm25s.tokenize(CORPUS + QUERIES, stopwords="en", stemmer=None, return_ids=False, show_progress=False)
    assert [list(t) for t in tok] == [split_tokens(t) for t in CORPUS + QUERIES]


def _our_index():
    v = Vocabulary(BuiltinStemmer())
    toks, lens = v.encode_corpus(CORPUS)
    return v, bo.build_index(toks, lens, len(v))


def test_index_arrays_equal_bm25s():
    """bm25s.BM25().index(): the CSC {data, indices, indptr} per term, compared through the vocabulary
    (term ids differ: bm25s numbers stems in set order, App. A.2)."""
    bm25s = pytest.importorskip("bm25s")
    pytest.importorskip("Stemmer")
    import Stemmer
    st = Stemmer.Stemmer("english")
    real = bm25s.BM25()
    ct = bm25s.tokenize(CORPUS, stopwords="en", stemmer=st, show_progress=False)
    real.index(ct, show_progress=False)
    v, ours = _our_index()
    real_vocab = {k: i for k, i in ct.vocab.items() if k != ""}
    assert set(real_vocab) == set(v.stem_to_id)
    rs = real.scores
    for stem, rid in real_vocab.items():
        oid = v.stem_to_id[stem]
        r_lo, r_hi = rs["indptr"][rid], rs["indptr"][rid + 1]
        o_lo, o_hi = ours["indptr"][oid], ours["indptr"][oid + 1]
        order = np.argsort(rs["indices"][r_lo:r_hi])
        assert np.array_equal(rs["indices"][r_lo:r_hi][order], ours["indices"][o_lo:o_hi]), stem
        assert np.allclose(rs["data"][r_lo:r_hi][order], ours["data"][o_lo:o_hi], rtol=1e-6, atol=0), stem


def test_retrieve_equals_bm25s_modulo_tie_groups():
    bm25s = pytest.importorskip("bm25s")
    import Stemmer
    st = Stemmer.Stemmer("english")
    real = bm25s.BM25()
    real.index(bm25s.tokenize(CORPUS, stopwords="en", stemmer=st, show_progress=False), show_progress=False)
    v, ours = _our_index()
    k = 4
    for q in QUERIES:
        ids, scores = real.retrieve(bm25s.tokenize(q, stemmer=st, show_progress=False), k=k, show_progress=False)
        os_, od = bo.retrieve(ours, np.array(v.encode_query(q), np.int32), k)
        assert np.allclose(scores[0], os_, rtol=1e-5, atol=0), q
        assert bo.same_modulo_ties(os_, np.asarray(ids[0], np.int32), os_, od), q


def test_llama_index_retriever_returns_the_same_passages():
    pytest.importorskip("llama_index.retrievers.bm25")
    from llama_index.core import Document
    from llama_index.core.storage.docstore import SimpleDocumentStore
    from llama_index.retrievers.bm25 import BM25Retriever as RealRetriever
    store = SimpleDocumentStore()
    store.add_documents([Document(text=t, doc_id=str(i)) for i, t in enumerate(CORPUS)])   # make_indexer.py:438-442
    real = RealRetriever.from_defaults(docstore=store, similarity_top_k=3)                 # exp_rag.py:242
    v, ours = _our_index()
    for q in QUERIES[:4]:
        res = real.retrieve(q)                                                             # exp_rag.py:426
        os_, od = bo.retrieve(ours, np.array(v.encode_query(q), np.int32), 3)
        assert np.allclose([r.score for r in res], os_, rtol=1e-5, atol=0), q
        got = np.array([int(r.node.node_id) for r in res], np.int32)
        assert bo.same_modulo_ties(os_, got, os_, od), q
        pos = os_ > 0
        assert [r.text for r, p in zip(res, pos) if p and (os_ == r.score).sum() == 1] == \
            [CORPUS[d] for d, s in zip(od, os_) if s > 0 and (os_ == s).sum() == 1]
