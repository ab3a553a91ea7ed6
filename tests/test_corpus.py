"""Corpus ingestion, docstore formats and the persisted passage store / vocabulary.
CPU only: the on-disk formats either side of the hot path (make_indexer.py:252-293, 436-444, 459-464;
exp_rag.py:241)."""
import csv
import json

import numpy as np
import pytest

from probing_rag_b200.corpus import (PassageStore, dedup_stable, iter_docstore_json, read_index_csv,
                                     read_wiki_tsv, write_index_csv)
from probing_rag_b200.retriever import Document, SimpleDocumentStore
from probing_rag_b200.text import STOPWORDS_EN, BuiltinStemmer, Vocabulary, split_tokens

TEXTS = [
    "Aaron Aaron ( or ; \"Ahärôn\") is a prophet, high priest, and the brother of Moses",
    'He said, "quoted, with a comma"\tand a tab',
    "",
    "Zürich – naïve café, 東京 is the capital of Japan",
    "line one\nline two of the same passage",
    "The the of and",        # stop words only -> empty document
]


def test_wiki_tsv_reader_matches_the_reference_loop(tmp_path):
    """make_indexer.py:258-265: csv.reader(delimiter='\\t'), header skipped, text = column 1."""
    path = tmp_path / "psgs_w100.tsv"
    with open(path, "w", newline="", encoding="utf-8") as f:
        wr = csv.writer(f, delimiter="\t", lineterminator="\n")
        wr.writerow(["id", "text", "title"])
        for i, t in enumerate(TEXTS):
            wr.writerow([i + 1, t, f"title {i}"])
    got = list(read_wiki_tsv(str(path)))
    with open(path) as f:                       # the reference's own loop
        tr = csv.reader(f, delimiter="\t")
        next(tr)
        want = [line[1] for line in tr]
    assert got == want == TEXTS


def test_dedup_keeps_first_occurrence():
    assert list(dedup_stable(["b", "a", "b", "c", "a"])) == ["b", "a", "c"]


def test_index_csv_round_trip_and_pandas_shape(tmp_path):
    """make_indexer.py:459-464: DataFrame([texts, doc_ids]).T, columns ['doc','doc_id'], to_csv(index=False)."""
    pd = pytest.importorskip("pandas")
    ref = tmp_path / "ref.csv"
    df = pd.DataFrame([TEXTS, list(range(len(TEXTS)))]).T
    df.columns = ["doc", "doc_id"]
    df.to_csv(ref, index=False)
    # pandas writes the empty passage as an empty field and reads it back as NaN; the text itself is ""
    assert list(read_index_csv(str(ref))) == TEXTS
    ours = tmp_path / "ours.csv"
    assert write_index_csv(str(ours), TEXTS) == len(TEXTS)
    assert list(read_index_csv(str(ours))) == TEXTS
    back = pd.read_csv(ours, keep_default_na=False)
    assert list(back.columns) == ["doc", "doc_id"] and back["doc"].tolist() == TEXTS
    assert back["doc_id"].tolist() == list(range(len(TEXTS)))


def test_index_csv_rejects_shuffled_ids(tmp_path):
    p = tmp_path / "bad.csv"
    p.write_text("doc,doc_id\nfoo,1\nbar,0\n")
    with pytest.raises(ValueError):
        list(read_index_csv(str(p)))


def _docstore_blob(shape: str) -> dict:
    data = {}
    for i, t in enumerate(TEXTS):
        if shape == "text":                     # llama-index-core < 0.12
            d = {"id_": str(i), "embedding": None, "metadata": {}, "text": t, "class_name": "Document"}
        elif shape == "text_resource":          # core >= 0.12
            d = {"id_": str(i), "metadata": {"src": i}, "text_resource": {"text": t, "mimetype": None},
                 "class_name": "Document"}
        else:                                   # __data__ serialised as a string
            d = json.dumps({"id_": str(i), "text": t, "metadata": {}})
        data[str(i)] = {"__data__": d, "__type__": "4"}
    return {"docstore/data": data, "docstore/metadata": {k: {"doc_hash": "x"} for k in data},
            "docstore/ref_doc_info": {}}


@pytest.mark.parametrize("shape", ["text", "text_resource", "string"])
def test_docstore_json_shapes(tmp_path, shape):
    path = tmp_path / f"llama_index_bm25_model_{shape}.json"
    path.write_text(json.dumps(_docstore_blob(shape)))
    rows = list(iter_docstore_json(str(path)))
    assert [r[0] for r in rows] == [str(i) for i in range(len(TEXTS))]
    assert [r[1] for r in rows] == TEXTS
    store = SimpleDocumentStore.from_persist_path(str(path))          # exp_rag.py:241
    assert list(store.docs) == [str(i) for i in range(len(TEXTS))]    # insertion order = doc index
    assert [d.text for d in store.docs.values()] == TEXTS
    assert [d.get_content() for d in store.docs.values()] == TEXTS
    if shape == "text_resource":
        assert store.docs["3"].metadata == {"src": 3}


def test_docstore_persist_round_trip(tmp_path):
    """make_indexer.py:436-444 then exp_rag.py:241."""
    store = SimpleDocumentStore()
    store.add_documents([Document(text=t, doc_id=f"{n}") for n, t in enumerate(TEXTS)])
    path = tmp_path / "store.json"
    store.persist(str(path))
    blob = json.loads(path.read_text())
    assert set(blob) >= {"docstore/data", "docstore/metadata"}
    assert blob["docstore/data"]["1"]["__data__"]["text"] == TEXTS[1]
    back = SimpleDocumentStore.from_persist_path(str(path))
    assert [(d.id_, d.text) for d in back.docs.values()] == [(str(i), t) for i, t in enumerate(TEXTS)]


def test_passage_store_round_trip(tmp_path):
    n = PassageStore.write(str(tmp_path), iter(TEXTS))
    assert n == len(TEXTS) and PassageStore.exists(str(tmp_path))
    ps = PassageStore.open(str(tmp_path))
    assert len(ps) == len(TEXTS)
    assert [ps[i] for i in range(len(ps))] == TEXTS == list(ps)
    assert ps.doc_id(4) == "4"
    with pytest.raises(IndexError):
        ps.text(len(TEXTS))
    # custom ids are stored, default ids are not
    PassageStore.write(str(tmp_path), TEXTS, [f"doc-{i}" for i in range(len(TEXTS))])
    assert PassageStore.open(str(tmp_path)).doc_id(2) == "doc-2"
    PassageStore.write(str(tmp_path), TEXTS, [str(i) for i in range(len(TEXTS))])
    assert not (tmp_path / PassageStore.IDS).exists()
    # empty corpus
    empty = tmp_path / "empty"
    assert PassageStore.write(str(empty), []) == 0 and len(PassageStore.open(str(empty))) == 0


def test_passage_store_detects_truncation(tmp_path):
    PassageStore.write(str(tmp_path), TEXTS)
    with open(tmp_path / PassageStore.BIN, "ab") as f:
        f.write(b"x")
    with pytest.raises(ValueError):
        PassageStore.open(str(tmp_path))


def _encode_sequential(texts, stemmer):
    """The literal per-document loop of bm25s.tokenize (App. A.2): stop words dropped before
    stemming, stem ids in first-seen order."""
    s2i, toks, lens = {}, [], []
    for t in texts:
        n = 0
        for w in split_tokens(t):
            st = stemmer.stemWords([w])[0]
            toks.append(s2i.setdefault(st, len(s2i)))
            n += 1
        lens.append(n)
    return np.array(toks, np.int32), np.array(lens, np.int32), s2i


def test_batched_tokenisation_equals_the_sequential_loop():
    rng = np.random.default_rng(5)
    words = ("running runs ran cats cat the of and gardens gardening garden nationality national nation "
             "Zürich café 東京 capital x yy 42 isn't U.S. retrieval retrieved probing probes").split()
    texts = [" ".join(rng.choice(words, size=rng.integers(0, 40))) for _ in range(300)] + TEXTS
    want_t, want_l, want_vocab = _encode_sequential(texts, BuiltinStemmer())
    for batch_docs in (1, 7, 4096):
        v = Vocabulary(BuiltinStemmer())
        got_t, got_l = v.encode_corpus(iter(texts), batch_docs=batch_docs)
        assert got_t.dtype == np.int32 and got_l.dtype == np.int32
        assert np.array_equal(got_t, want_t) and np.array_equal(got_l, want_l)
        assert v.stem_to_id == want_vocab
    assert want_l[-1] == 0 and want_l[len(texts) - len(TEXTS) + 2] == 0      # stop-word-only and empty passages
    assert not (set(v.stem_to_id) & STOPWORDS_EN)


def test_vocabulary_save_load_serves_the_same_queries(tmp_path):
    v = Vocabulary(BuiltinStemmer())
    v.encode_corpus(TEXTS + ["cats are running in the gardens", "national retrieval of probes"])
    path = tmp_path / "vocab.txt"
    v.save(str(path))
    w = Vocabulary.load(str(path), BuiltinStemmer())
    assert w.stem_to_id == v.stem_to_id and len(w) == len(v)
    for q in ("Running CATS and dogs, cats!", "the of", "", "Zürich 東京 nation probing", "gardening retrieved"):
        assert w.encode_query(q) == v.encode_query(q)
    qi, qt = w.encode_queries(["running cats", "", "dogs", "garden garden"])
    assert qi.tolist() == [0, 2, 2, 2, 4] and qt.dtype == np.int32
    assert qt.tolist() == [v.stem_to_id["run"], v.stem_to_id["cat"], v.stem_to_id["garden"], v.stem_to_id["garden"]]
