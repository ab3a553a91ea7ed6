"""The persisted retriever: a fresh process serves `retrieve(str)` -- ranked
nodes WITH their texts -- from a saved directory, without re-tokenising or re-indexing the corpus
(what /root/reference/exp_rag.py:241-242 does at every start).  Needs a B200."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import bm25_oracle as bo
from probing_rag_b200 import BM25Index, BM25Retriever, Document, SimpleDocumentStore
from probing_rag_b200.corpus import PassageStore, write_index_csv, read_index_csv

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORDS = ("retrieval augmented generation probing language model hidden state wikipedia passage "
         "question answer paris france capital city river tower london england bridge Zürich café "
         "running runs cats gardens national the of and is").split()
QUERIES = ["What is the capital city of France? The tower of Paris", "london bridge river", "",
           "cats running in Zürich gardens", "unknownword anotherunknown", "the of and"]


def make_texts(n=700, seed=3):
    rng = np.random.default_rng(seed)
    return [" ".join(rng.choice(WORDS, size=int(rng.integers(0, 30)))) for _ in range(n)]


def as_rows(res):
    return [[(r.node.id_, r.score, r.text) for r in one] for one in res]


def check_against_oracle(bm25, texts, queries, k):
    toks, lens = bm25.vocab.encode_corpus(texts)          # the vocabulary is complete: no new stems appear
    ora = bo.build_index(toks, lens, len(bm25.vocab))
    got = bm25.retrieve_batch(queries, k=k)
    for q, res in zip(queries, got):
        os_, od = bo.retrieve(ora, np.array(bm25.vocab.encode_query(q), np.int32), k)
        assert [int(r.node.id_) for r in res] == od.tolist(), q
        assert [r.score for r in res] == [float(x) for x in os_], q
        assert [r.text for r in res] == [texts[d] for d in od], q


def test_persist_then_load_serves_identical_nodes_with_text(tmp_path):
    texts = make_texts()
    store = SimpleDocumentStore()
    store.add_documents([Document(text=t, doc_id=str(i), metadata={"row": i} if i % 7 == 0 else None)
                         for i, t in enumerate(texts)])
    bm25 = BM25Retriever.from_defaults(docstore=store, similarity_top_k=5)
    before = as_rows(bm25.retrieve_batch(QUERIES))
    check_against_oracle(bm25, texts, QUERIES, 5)
    d = str(tmp_path / "wiki_bm25")
    bm25.persist(d)
    assert {"index.json", "indptr.i64", "doc_ids.i32", "weights.f32", "vocab.txt", "passages.bin", "passages.off",
            "retriever.json", "metadata.json"} <= set(os.listdir(d))
    loaded = BM25Retriever.from_persist_dir(d)
    assert loaded.similarity_top_k == 5 and loaded.corpus is None and len(loaded.passages) == len(texts)
    assert as_rows(loaded.retrieve_batch(QUERIES)) == before
    one = loaded.retrieve(QUERIES[0])                       # exp_rag.py:426
    assert [(r.node.id_, r.score, r.text) for r in one] == before[0]
    assert one[0].text == texts[int(one[0].node.id_)] and one[0].get_content() == one[0].text
    assert loaded.retrieve(QUERIES[3])[0].metadata == store.docs[loaded.retrieve(QUERIES[3])[0].node.id_].metadata
    check_against_oracle(loaded, texts, QUERIES, 3)
    assert as_rows(BM25Retriever.from_persist_dir(d, similarity_top_k=2).retrieve_batch(QUERIES[:2])) == \
        [r[:2] for r in before[:2]]


def test_fresh_process_serves_the_saved_directory(tmp_path):
    texts = make_texts(400, seed=9)
    d = str(tmp_path / "ix")
    bm25 = BM25Retriever.from_texts(iter(texts), similarity_top_k=4, persist_dir=d)
    want = as_rows(bm25.retrieve_batch(QUERIES))
    code = ("import json, sys; sys.path.insert(0, %r)\n"
            "from probing_rag_b200 import BM25Retriever\n"
            "r = BM25Retriever.from_persist_dir(%r)\n"
            "print(json.dumps([[(n.node.id_, n.score, n.text) for n in one] for one in r.retrieve_batch(%r)]))\n"
            % (ROOT, d, QUERIES))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    got = json.loads(out.stdout.strip().splitlines()[-1])
    assert [[tuple(x) for x in one] for one in got] == want


def test_streamed_build_equals_the_docstore_build(tmp_path):
    """from_texts over the (doc, doc_id) CSV stream (make_indexer.py:459-464) == from_defaults(docstore=...)."""
    texts = make_texts(300, seed=11)
    csv_path = str(tmp_path / "wiki_index_2.csv")
    write_index_csv(csv_path, texts)
    streamed = BM25Retriever.from_texts(read_index_csv(csv_path), similarity_top_k=5, persist_dir=str(tmp_path / "p"))
    in_memory = BM25Retriever.from_texts(iter(texts), similarity_top_k=5)
    store = SimpleDocumentStore()
    store.add_documents([Document(text=t, doc_id=str(i)) for i, t in enumerate(texts)])
    classic = BM25Retriever.from_defaults(docstore=store, similarity_top_k=5)
    a, b, c = (as_rows(r.retrieve_batch(QUERIES)) for r in (streamed, in_memory, classic))
    assert a == b == c
    for r in (streamed, in_memory):
        assert np.array_equal(r.index.weights.cpu().numpy(), classic.index.weights.cpu().numpy())
        assert np.array_equal(r.index.doc_ids.cpu().numpy(), classic.index.doc_ids.cpu().numpy())
    assert list(PassageStore.open(str(tmp_path / "p"))) == texts


def test_load_rejects_mismatched_pieces(tmp_path):
    texts = make_texts(100, seed=1)
    d = str(tmp_path / "ix")
    BM25Retriever.from_texts(iter(texts), persist_dir=d)
    with open(os.path.join(d, "vocab.txt"), "a") as f:
        f.write("extra\n")
    with pytest.raises(ValueError):
        BM25Retriever.from_persist_dir(d)


def test_bare_shard_does_not_return_short_lists_silently():
    """A doc-range shard with fewer than k documents leaves (-inf, -1) entries in its local list;
    llama-index always returns k nodes, so the text-level API refuses instead of dropping them."""
    texts = make_texts(64, seed=2)
    full = BM25Retriever.from_texts(iter(texts), similarity_top_k=5)
    toks, lens = full.vocab.encode_corpus(texts)
    whole = bo.build_index(toks, lens, len(full.vocab))
    ora = bo.build_index(toks[: int(lens[:3].sum())], lens[:3], len(full.vocab), n_docs_global=64,
                         avgdl_global=whole["avgdl"], df_global=whole["df"])
    shard = BM25Index.from_arrays(ora["data"], ora["indices"], ora["indptr"], 3, n_docs_global=64)
    r = BM25Retriever(None, 5, index=shard, vocab=full.vocab)
    with pytest.raises(ValueError, match="shard"):
        r.retrieve("paris tower")
    s, d = r.retrieve_ids(torch.tensor([0, 1], dtype=torch.int64, device="cuda"),
                          torch.tensor([0], dtype=torch.int32, device="cuda"), 5)
    assert (d.cpu().numpy()[0, 3:] == -1).all()
