"""Corpus ingestion and the passage store next to the persisted index.

What the reference does around the retriever (/root/reference/make_indexer.py):

    :252-293  make_wiki_documents: csv.reader(psgs_w100.tsv, delimiter='\\t'), skip the header,
              passage text = column 1, then `list(set(texts))` (hash-order dedup) and doc_id = row
    :436-444  Document(text=text, doc_id=f'{num}') for every passage -> SimpleDocumentStore
              -> one JSON file, re-read and RE-INDEXED at every start (exp_rag.py:241-242)
    :459-464  pandas DataFrame([texts, doc_ids]).T, columns ['doc', 'doc_id'] -> CSV

Here the same inputs are read as streams of passage texts (never a list of 21M Python
objects per token), and what a retriever needs besides the CSR -- the passage texts, addressed
by doc index -- is kept as two flat files that are memory-mapped at load time:

    passages.bin   utf-8 bytes of all passages back to back
    passages.off   int64[n+1] byte offsets
    doc_ids.txt    only when the ids are not "0", "1", ... (the reference's are, :439)
"""
from __future__ import annotations

import csv
import json
import os
import sys
from typing import Iterable, Iterator

import numpy as np


def _raise_field_limit() -> None:
    try:
        csv.field_size_limit(sys.maxsize)
    except OverflowError:
        csv.field_size_limit(2 ** 31 - 1)


def read_wiki_tsv(path: str, text_column: int = 1, skip_header: bool = True) -> Iterator[str]:
    """Passage texts of a DPR `psgs_w100.tsv` (id, text, title), the way make_indexer.py:258-265
    reads it: csv.reader with a tab delimiter, first row skipped, column 1."""
    _raise_field_limit()
    with open(path, newline="", encoding="utf-8") as f:
        tr = csv.reader(f, delimiter="\t")
        if skip_header:
            next(tr, None)
        for row in tr:
            yield row[text_column]


def dedup_stable(texts: Iterable[str]) -> Iterator[str]:
    """Duplicate removal that keeps the first occurrence.  The reference uses `list(set(texts))`
    (make_indexer.py:289), whose order depends on PYTHONHASHSEED; only the persisted
    docstore / CSV fixes the doc-id assignment, so any stable order is equally valid."""
    seen = set()
    for t in texts:
        if t not in seen:
            seen.add(t)
            yield t


def read_index_csv(path: str) -> Iterator[str]:
    """Passage texts, in doc_id order, of the `(doc, doc_id)` CSV make_indexer.py:459-464 writes
    with pandas (`doc_id` = row number there, which is checked)."""
    _raise_field_limit()
    with open(path, newline="", encoding="utf-8") as f:
        rd = csv.reader(f)
        header = next(rd, None)
        if header is None:
            return
        try:
            c_doc, c_id = header.index("doc"), header.index("doc_id")
        except ValueError:
            raise ValueError(f"{path}: expected columns 'doc' and 'doc_id', got {header}") from None
        for i, row in enumerate(rd):
            if int(row[c_id]) != i:
                raise ValueError(f"{path}: doc_id {row[c_id]} at row {i}: the CSV is not in doc order")
            yield row[c_doc]


def write_index_csv(path: str, texts: Iterable[str]) -> int:
    """The same file `df.to_csv(path, index=False)` produces at make_indexer.py:464."""
    n = 0
    with open(path, "w", newline="", encoding="utf-8") as f:
        wr = csv.writer(f, lineterminator="\n")
        wr.writerow(["doc", "doc_id"])
        for n, t in enumerate(texts, 1):
            wr.writerow([t, n - 1])
    return n


def iter_docstore_json(path: str) -> Iterator[tuple[str, str, dict]]:
    """(doc_id, text, metadata) of a SimpleDocumentStore JSON (make_indexer.py:444, read at
    exp_rag.py:241) in key order = insertion order = doc index.  Handles both shapes of
    SURVEY App. A.8: `__data__.text` and `__data__.text_resource.text` (core >= 0.12); `__data__`
    may itself be a JSON string."""
    with open(path, encoding="utf-8") as f:
        blob = json.load(f)
    for doc_id, entry in blob["docstore/data"].items():
        d = entry.get("__data__", entry)
        if isinstance(d, str):
            d = json.loads(d)
        text = d.get("text")
        if text is None:
            text = (d.get("text_resource") or {}).get("text", "")
        yield d.get("id_", doc_id), text, d.get("metadata") or {}


class PassageStore:
    """Read-only passages addressed by doc index, memory-mapped (21M passages ~ 13 GB of text
    stay on disk / in the page cache instead of 21M Python strings)."""

    BIN, OFF, IDS = "passages.bin", "passages.off", "doc_ids.txt"

    def __init__(self, blob, offsets: np.ndarray, doc_ids: list[str] | None = None):
        self._blob = blob
        self._off = offsets
        self._ids = doc_ids

    def __len__(self) -> int:
        return len(self._off) - 1

    def text(self, i: int) -> str:
        if not 0 <= i < len(self):
            raise IndexError(i)
        return bytes(self._blob[int(self._off[i]):int(self._off[i + 1])]).decode("utf-8")

    __getitem__ = text

    def doc_id(self, i: int) -> str:
        return self._ids[i] if self._ids is not None else str(i)

    def __iter__(self) -> Iterator[str]:
        for i in range(len(self)):
            yield self.text(i)

    @classmethod
    def write(cls, path: str, texts: Iterable[str], doc_ids: Iterable[str] | None = None) -> int:
        """Stream passages into `path/`; returns how many were written."""
        os.makedirs(path, exist_ok=True)
        offs = [0]
        pos = 0
        with open(os.path.join(path, cls.BIN), "wb") as f:
            for t in texts:
                b = t.encode("utf-8")
                f.write(b)
                pos += len(b)
                offs.append(pos)
        np.asarray(offs, dtype=np.int64).tofile(os.path.join(path, cls.OFF))
        ids_path = os.path.join(path, cls.IDS)
        if os.path.exists(ids_path):
            os.remove(ids_path)
        if doc_ids is not None:
            ids = list(doc_ids)
            if len(ids) != len(offs) - 1:
                raise ValueError("doc_ids and texts differ in length")
            if any(s != str(i) for i, s in enumerate(ids)):     # the reference's ids are str(i): nothing to store
                if any("\n" in s for s in ids):
                    raise ValueError("a doc id contains a newline")
                with open(ids_path, "w", encoding="utf-8", newline="\n") as f:
                    f.write("\n".join(ids) + "\n")
        return len(offs) - 1

    @classmethod
    def open(cls, path: str) -> "PassageStore":
        off = np.fromfile(os.path.join(path, cls.OFF), dtype=np.int64)
        bin_path = os.path.join(path, cls.BIN)
        blob = np.memmap(bin_path, dtype=np.uint8, mode="r") if os.path.getsize(bin_path) else np.zeros(0, np.uint8)
        if len(off) < 1 or off[0] != 0 or int(off[-1]) != blob.size or np.any(np.diff(off) < 0):
            raise ValueError(f"{path}: passage offsets do not match the passage file")
        ids = None
        ids_path = os.path.join(path, cls.IDS)
        if os.path.exists(ids_path):
            with open(ids_path, encoding="utf-8", newline="\n") as f:
                ids = [line.rstrip("\n") for line in f]
            if len(ids) != len(off) - 1:
                raise ValueError(f"{path}: {len(ids)} doc ids for {len(off) - 1} passages")
        return cls(blob, off, ids)

    @classmethod
    def exists(cls, path: str) -> bool:
        return os.path.exists(os.path.join(path, cls.OFF)) and os.path.exists(os.path.join(path, cls.BIN))
