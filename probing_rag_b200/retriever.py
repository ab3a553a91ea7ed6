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
from typing import Iterable, Sequence

import numpy as np
import torch

from .index import BM25Index
from .text import Vocabulary


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
        with open(persist_path) as f:
            blob = json.load(f)
        store = cls()
        for doc_id, entry in blob["docstore/data"].items():
            d = entry.get("__data__", entry)
            if isinstance(d, str):
                d = json.loads(d)
            text = d.get("text")
            if text is None:
                text = (d.get("text_resource") or {}).get("text", "")
            store.docs[doc_id] = Document(text=text, doc_id=d.get("id_", doc_id), metadata=d.get("metadata") or {})
        return store


class BM25Retriever:
    """llama_index.retrievers.bm25.BM25Retriever look-alike over a GPU-resident index."""

    def __init__(self, nodes: Sequence[Document] | None, similarity_top_k: int = 2,
                 index: BM25Index | None = None, vocab: Vocabulary | None = None,
                 stemmer=None, device="cuda"):
        self.similarity_top_k = int(similarity_top_k)
        self.corpus = list(nodes) if nodes is not None else None
        if index is None:
            if not nodes:
                raise ValueError("Please pass exactly one of index, nodes, or docstore.")
            vocab = Vocabulary(stemmer)
            toks, lens = [], []
            for n in self.corpus:
                ids = vocab.encode_corpus_doc(n.get_content())
                toks.extend(ids)
                lens.append(len(ids))
            dev = torch.device(device)
            index = BM25Index.from_tokens(
                torch.tensor(toks, dtype=torch.int32, device=dev),
                torch.tensor(lens, dtype=torch.int32, device=dev), max(len(vocab), 1))
        self.index = index
        self.vocab = vocab

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

    # ---- token-id level (the hot path)
    def retrieve_ids(self, q_indptr: torch.Tensor, q_terms: torch.Tensor, k: int | None = None):
        """CSR batch of term ids on the device -> (scores f32[B,k], doc_ids i32[B,k]) on the device."""
        return self.index.topk(q_indptr, q_terms, self.similarity_top_k if k is None else k)

    # ---- text level
    def _encode(self, queries: Sequence[str]):
        if self.vocab is None:
            raise ValueError("this retriever was built from a token-id index: use retrieve_ids")
        ids = [self.vocab.encode_query(q) for q in queries]
        q_indptr = np.zeros(len(ids) + 1, dtype=np.int64)
        np.cumsum([len(x) for x in ids], out=q_indptr[1:])
        q_terms = np.fromiter((t for x in ids for t in x), dtype=np.int32, count=int(q_indptr[-1]))
        return q_indptr, q_terms

    def _nodes(self, scores: np.ndarray, ids: np.ndarray) -> list[NodeWithScore]:
        out = []
        base = self.index.doc_id_base
        for s, d in zip(scores.tolist(), ids.tolist()):
            if d < 0:
                continue
            node = self.corpus[d - base] if self.corpus is not None else Document(text="", doc_id=str(d))
            out.append(NodeWithScore(node=node, score=float(s)))
        return out

    def retrieve_batch(self, queries: Sequence[str], k: int | None = None) -> list[list[NodeWithScore]]:
        k = self.similarity_top_k if k is None else k
        q_indptr, q_terms = self._encode(queries)
        scores, ids, _, _ = self.index.topk_host(q_indptr, q_terms, k)
        return [self._nodes(scores[i], ids[i]) for i in range(len(queries))]

    def retrieve(self, str_or_query_bundle) -> list[NodeWithScore]:
        """exp_rag.py:426 / :428 / :492, utils.py:640: one query string -> k ranked nodes."""
        q = getattr(str_or_query_bundle, "query_str", str_or_query_bundle)
        return self.retrieve_batch([q])[0]
