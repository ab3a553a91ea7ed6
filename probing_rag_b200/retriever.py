"""Drop-in for the retriever boundary of /root/reference/exp_rag.py:

    bm25 = BM25Retriever.from_defaults(docstore=docstore2, similarity_top_k=5)    # :242
    retrieved_passages = bm25.retrieve(value['text'][0])                          # :426, :428, :492
    evidence.text                                                                  # :372

same names, same argument meaning, same result shape (llama-index `NodeWithScore` look-alikes,
score-descending) and the same exceptions, with the scoring done by libprobingrag.so on the
GPU.  The batched entry points (`retrieve_batch`, `retrieve_ids`) are what the reference's
one-query-at-a-time loop cannot express.
"""
from __future__ import annotations

import json
import os
from typing import Iterable, Sequence

import numpy as np
import torch

from .corpus import PassageStore, iter_docstore_json
from .index import BM25Index
from .text import BuiltinStemmer, Vocabulary


class Document:
    """llama_index.core.Document look-alike (make_indexer.py:439: Document(text=.., doc_id=..))."""

    def __init__(self, text: str = "", doc_id: str | None = None, id_: str | None = None,
                 metadata: dict | None = None):
        self.text = text
        self.id_ = doc_id if doc_id is not None else (id_ if id_ is not None else "")
        self.metadata = metadata or {}

    @property
    def node_id(self) -> str:
        return self.id_

    @property
    def doc_id(self) -> str:
        return self.id_

    def get_content(self, metadata_mode=None) -> str:
        return self.text

    def __repr__(self):
        return f"Document(id_={self.id_!r}, text={self.text[:40]!r})"


TextNode = Document


class NodeWithScore:
    """llama_index.core.schema.NodeWithScore look-alike: .node, .score, .text, .get_content()."""

    def __init__(self, node: Document, score: float):
        self.node = node
        self.score = score

    @property
    def text(self) -> str:
        return self.node.text

    @property
    def node_id(self) -> str:
        return self.node.node_id

    @property
    def id_(self) -> str:
        return self.node.id_

    @property
    def metadata(self) -> dict:
        return self.node.metadata

    def get_content(self, metadata_mode=None) -> str:
        return self.node.get_content()

    def get_text(self) -> str:
        return self.node.text

    def get_score(self, raise_error: bool = False) -> float:
        if self.score is None:
            if raise_error:
                raise ValueError("Score not set.")
            return 0.0
        return self.score

    def __repr__(self):
        return f"NodeWithScore(score={self.score:.6f}, node={self.node!r})"


class SimpleDocumentStore:
    """Reader/writer of the docstore JSON the reference persists (make_indexer.py:441-444) and
    loads (exp_rag.py:241); SURVEY App. A.8.  `.docs` keeps insertion order = doc index."""

    def __init__(self):
        self.docs: dict[str, Document] = {}

    def add_documents(self, documents: Iterable[Document]) -> None:
        for d in documents:
            self.docs[d.id_] = d

    def persist(self, persist_path: str) -> None:
        data = {d.id_: {"__data__": {"id_": d.id_, "text": d.text, "metadata": d.metadata}, "__type__": "4"}
                for d in self.docs.values()}
        with open(persist_path, "w") as f:
            json.dump({"docstore/data": data, "docstore/metadata": {k: {"doc_hash": ""} for k in data}}, f)

    @classmethod
    def from_persist_path(cls, persist_path: str) -> "SimpleDocumentStore":
        store = cls()
        for doc_id, text, metadata in iter_docstore_json(persist_path):
            store.docs[doc_id] = Document(text=text, doc_id=doc_id, metadata=metadata)
        return store


RETRIEVER_JSON = "retriever.json"
VOCAB_FILE = "vocab.txt"
METADATA_FILE = "metadata.json"


class BM25Retriever:
    """llama_index.retrievers.bm25.BM25Retriever look-alike over a GPU-resident index.

    `index` is a `BM25Index` (whole corpus on one GPU) or a `sharding.ShardedBM25` (this rank's
    doc-range shard + the exchange step; every rank then returns the same global lists).  The
    passages behind the doc ids are either the nodes the retriever was built from or a
    memory-mapped `PassageStore` (persisted retrievers)."""

    def __init__(self, nodes: Sequence[Document] | None, similarity_top_k: int = 2,
                 index=None, vocab: Vocabulary | None = None, stemmer=None, device="cuda",
                 passages: PassageStore | None = None, metadata: list | None = None):
        self.similarity_top_k = int(similarity_top_k)
        self.corpus = list(nodes) if nodes is not None else None
        self.passages = passages
        self._metadata = metadata
        if index is None:
            if not nodes:
                raise ValueError("Please pass exactly one of index, nodes, or docstore.")
            vocab = Vocabulary(stemmer)
            index = self._build_index(vocab, (n.get_content() for n in self.corpus), device)
        self.index = index
        self.vocab = vocab
        # global doc id of corpus[0]: nodes given next to a prebuilt shard are that shard's documents
        self._node_base = int(getattr(index, "doc_id_base", 0)) if self.corpus is not None else 0

    @staticmethod
    def _build_index(vocab: Vocabulary, texts: Iterable[str], device, progress=None) -> BM25Index:
        """Tokenise in batches into flat numpy arrays (no per-token Python objects survive a
        batch) and build the index on the GPU (bm25s `tokenize` + `index`, App. A.2-A.4)."""
        toks, lens = vocab.encode_corpus(texts, progress=progress)
        dev = torch.device(device)
        return BM25Index.from_tokens(torch.from_numpy(toks).to(dev), torch.from_numpy(lens).to(dev),
                                     max(len(vocab), 1))

    @classmethod
    def from_defaults(cls, index=None, nodes=None, docstore=None, stemmer=None, language: str = "en",
                      similarity_top_k: int = 2, verbose: bool = False, tokenizer=None, vocab=None,
                      device="cuda") -> "BM25Retriever":
        """exp_rag.py:242.  Exactly one of index / nodes / docstore (llama-index semantics,
        App. A.1); `index` here is a prebuilt `BM25Index` so the corpus is not re-indexed."""
        if sum(x is not None for x in (index, nodes, docstore)) != 1:
            raise ValueError("Please pass exactly one of index, nodes, or docstore.")
        if language not in ("en", "english"):
            raise ValueError("only the English pipeline the reference uses is implemented")
        if docstore is not None:
            nodes = list(docstore.docs.values())
        return cls(nodes, similarity_top_k=similarity_top_k, index=index, vocab=vocab, stemmer=stemmer,
                   device=device)

    @classmethod
    def from_texts(cls, texts: Iterable[str], similarity_top_k: int = 2, stemmer=None, device="cuda",
                   persist_dir: str | None = None, progress=None) -> "BM25Retriever":
        """Build from a STREAM of passage texts (doc id = position, as make_indexer.py:438-439
        assigns them) -- `corpus.read_wiki_tsv`, `corpus.read_index_csv`, ... -- without holding
        the corpus as Python objects: with `persist_dir` every text goes straight into the
        `PassageStore` there while it is tokenised, and the finished retriever (CSR, vocabulary,
        passages) is persisted in the same directory.  Without it the texts are kept in memory."""
        vocab = Vocabulary(stemmer)
        if persist_dir is None:
            kept: list[str] = []

            def tee():
                for t in texts:
                    kept.append(t)
                    yield t
            index = cls._build_index(vocab, tee(), device, progress)
            return cls([Document(text=t, doc_id=str(i)) for i, t in enumerate(kept)], similarity_top_k,
                       index=index, vocab=vocab)
        # one pass over the stream, two consumers: PassageStore.write pulls the texts, and every
        # 4096 of them are tokenised as they go by
        toks, lens, chunk, n_seen = [], [], [], [0]

        def flush():
            t, l = vocab.encode_corpus_batch(chunk)
            toks.append(t)
            lens.append(l)
            n_seen[0] += len(chunk)
            chunk.clear()
            if progress is not None:
                progress(n_seen[0])

        def passing():
            for t in texts:
                chunk.append(t)
                if len(chunk) >= 4096:
                    flush()
                yield t
            if chunk:
                flush()
        n = PassageStore.write(persist_dir, passing())
        if n == 0:
            raise ValueError("Please pass exactly one of index, nodes, or docstore.")
        dev = torch.device(device)
        index = BM25Index.from_tokens(torch.from_numpy(np.concatenate(toks)).to(dev),
                                      torch.from_numpy(np.concatenate(lens)).to(dev), max(len(vocab), 1))
        r = cls(None, similarity_top_k, index=index, vocab=vocab, passages=PassageStore.open(persist_dir))
        r.persist(persist_dir)
        return r

    # ---- persistence (llama-index BM25Retriever.persist / from_persist_dir, App. A.1; the reference
    # does not use them and re-indexes at every start, exp_rag.py:241-242 -- this removes that cost)
    def _persist_meta(self, path: str) -> None:
        stem = "builtin-porter2" if isinstance(self.vocab.stemmer, BuiltinStemmer) else type(self.vocab.stemmer).__module__
        with open(os.path.join(path, RETRIEVER_JSON), "w") as f:
            json.dump({"format": "probing-rag-b200-retriever-v1", "similarity_top_k": self.similarity_top_k,
                       "stemmer": stem, "n_terms": len(self.vocab)}, f)

    def persist(self, path: str) -> None:
        """CSR + vocabulary + passages + retriever settings into `path/`: everything a fresh
        process needs to serve `retrieve(str)` with texts, without re-tokenising the corpus."""
        if self.vocab is None:
            raise ValueError("this retriever has no vocabulary (built from a token-id index): nothing to serve text queries with")
        if not isinstance(self.index, BM25Index):
            raise ValueError("persist the whole-corpus BM25Index; shards are cut from it at load time")
        os.makedirs(path, exist_ok=True)
        self.index.save(path)
        self.vocab.save(os.path.join(path, VOCAB_FILE))
        self._persist_meta(path)
        if self.corpus is not None:
            PassageStore.write(path, (n.get_content() for n in self.corpus), [n.id_ for n in self.corpus])
            meta = [n.metadata for n in self.corpus]
            mpath = os.path.join(path, METADATA_FILE)
            if any(meta):
                with open(mpath, "w") as f:
                    json.dump(meta, f)
            elif os.path.exists(mpath):
                os.remove(mpath)
        elif self.passages is not None:
            if not PassageStore.exists(path) or len(PassageStore.open(path)) != len(self.passages):
                PassageStore.write(path, iter(self.passages), [self.passages.doc_id(i) for i in range(len(self.passages))])

    @classmethod
    def from_persist_dir(cls, path: str, device="cuda", similarity_top_k: int | None = None,
                         stemmer=None, sharded: bool = False, group=None, exchange="p2p",
                         max_queries: int = 65536) -> "BM25Retriever":
        """Load a persisted retriever.  `sharded=True` (one process per GPU, `torch.distributed` initialised): every
        rank cuts ITS doc-range shard out of the saved whole-corpus index (`BM25Index.load(doc_range=…)`) and the
        retriever scores through `sharding.ShardedBM25` -- every rank returns the same merged lists, with texts from
        the memory-mapped passage store all ranks share."""
        with open(os.path.join(path, RETRIEVER_JSON)) as f:
            cfg = json.load(f)
        if cfg.get("format") != "probing-rag-b200-retriever-v1":
            raise ValueError(f"{path}: not a persisted probing-rag-b200 retriever")
        vocab = Vocabulary.load(os.path.join(path, VOCAB_FILE), stemmer)
        have = "builtin-porter2" if isinstance(vocab.stemmer, BuiltinStemmer) else type(vocab.stemmer).__module__
        if have != cfg.get("stemmer"):
            raise ValueError(f"{path} was indexed with stemmer {cfg.get('stemmer')!r} but this process has {have!r}: "
                             "query stems would not match the vocabulary")
        if sharded:
            import torch.distributed as dist

            from .sharding import ShardedBM25, shard_range
            with open(os.path.join(path, "index.json")) as f:
                n_docs = json.load(f)["n_docs"]
            rng = shard_range(n_docs, dist.get_rank(group), dist.get_world_size(group))
            shard = BM25Index.load(path, device=device, doc_range=rng)
            index = ShardedBM25(shard, group=group, exchange=exchange, max_queries=max_queries)
            index.n_terms = shard.n_terms
        else:
            index = BM25Index.load(path, device=device)
        if index.n_terms != max(len(vocab), 1):
            raise ValueError(f"{path}: vocabulary of {len(vocab)} stems for an index of {index.n_terms} terms")
        meta = None
        mpath = os.path.join(path, METADATA_FILE)
        if os.path.exists(mpath):
            with open(mpath) as f:
                meta = json.load(f)
        return cls(None, cfg["similarity_top_k"] if similarity_top_k is None else similarity_top_k,
                   index=index, vocab=vocab, passages=PassageStore.open(path), metadata=meta)

    # ---- token-id level (the hot path)
    def retrieve_ids(self, q_indptr: torch.Tensor, q_terms: torch.Tensor, k: int | None = None):
        """CSR batch of term ids on the device -> (scores f32[B,k], doc_ids i32[B,k]) on the device."""
        return self.index.topk(q_indptr, q_terms, self.similarity_top_k if k is None else k)

    # ---- text level
    def _encode(self, queries: Sequence[str]):
        if self.vocab is None:
            raise ValueError("this retriever was built from a token-id index: use retrieve_ids")
        return self.vocab.encode_queries(queries)

    def _node(self, d: int) -> Document:
        if self.corpus is not None:
            return self.corpus[d - self._node_base]
        if self.passages is not None:
            return Document(text=self.passages.text(d), doc_id=self.passages.doc_id(d),
                            metadata=self._metadata[d] if self._metadata is not None else None)
        raise ValueError("this retriever has no passages behind its doc ids (built from a bare index): "
                         "use retrieve_ids, or load it with from_persist_dir")

    def _nodes(self, scores: np.ndarray, ids: np.ndarray) -> list[NodeWithScore]:
        if ids.size and int(ids.min()) < 0:
            # only a bare doc-range shard with fewer than k documents can leave a list short; llama-index
            # always returns k nodes, so this is a mis-use, not a result
            raise ValueError("ranked list shorter than k: this retriever wraps ONE shard of a sharded corpus; "
                             "wrap the shard in sharding.ShardedBM25 so the lists are merged over all shards")
        return [NodeWithScore(node=self._node(d), score=float(s)) for s, d in zip(scores.tolist(), ids.tolist())]

    def retrieve_batch(self, queries: Sequence[str], k: int | None = None) -> list[list[NodeWithScore]]:
        k = self.similarity_top_k if k is None else k
        q_indptr, q_terms = self._encode(queries)
        scores, ids, _, _ = self.index.topk_host(q_indptr, q_terms, k)
        return [self._nodes(scores[i], ids[i]) for i in range(len(queries))]

    def retrieve(self, str_or_query_bundle) -> list[NodeWithScore]:
        """exp_rag.py:426 / :428 / :492, utils.py:640: one query string -> k ranked nodes."""
        q = getattr(str_or_query_bundle, "query_str", str_or_query_bundle)
        return self.retrieve_batch([q])[0]
