"""Text front-end of the retriever: `bm25s.tokenize` semantics as used by
llama-index's BM25Retriever (App. A.1) around /root/reference/exp_rag.py:242, 426.

    lower-case -> re.findall(r"(?u)\\b\\w\\w+\\b") -> drop the 33 English stop words ->
    Snowball-English stemming -> vocabulary ids; on the query side unknown tokens are
    dropped, duplicates and order are kept.

The reference stems with PyStemmer (`Stemmer.Stemmer("english")`), which is not installed
here (no network).  `get_stemmer()` uses it when importable and otherwise the built-in
restatement of the Snowball English (Porter2) algorithm below -- which therefore could not
be compared with PyStemmer in this container.
"""
from __future__ import annotations

import ctypes
import os
import re
import unicodedata

import numpy as np

TOKEN_PATTERN = re.compile(r"(?u)\b\w\w+\b")

# bm25s STOPWORDS_EN (the Lucene/Elasticsearch default list, App. A.2)
STOPWORDS_EN = frozenset(
    "a an and are as at be but by for if in into is it no not of on or such that the their "
    "then there these they this to was will with".split())

_VOWELS = frozenset("aeiouy")
_DOUBLES = ("bb", "dd", "ff", "gg", "mm", "nn", "pp", "rr", "tt")
_LI_ENDINGS = frozenset("cdeghkmnrt")

_EXCEPTIONS = {
    "skis": "ski", "skies": "sky", "dying": "die", "lying": "lie", "tying": "tie",
    "idly": "idl", "gently": "gentl", "ugly": "ugli", "early": "earli", "only": "onli",
    "singly": "singl",
    "sky": "sky", "news": "news", "howe": "howe", "atlas": "atlas", "cosmos": "cosmos",
    "bias": "bias", "andes": "andes",
}
_EXCEPTIONS_1A = frozenset(("inning", "outing", "canning", "herring", "earring", "proceed",
                            "exceed", "succeed"))

_STEP2 = (("ization", "ize"), ("ational", "ate"), ("fulness", "ful"), ("ousness", "ous"),
          ("iveness", "ive"), ("tional", "tion"), ("biliti", "ble"), ("lessli", "less"),
          ("entli", "ent"), ("ation", "ate"), ("alism", "al"), ("aliti", "al"), ("ousli", "ous"),
          ("iviti", "ive"), ("fulli", "ful"), ("enci", "ence"), ("anci", "ance"), ("abli", "able"),
          ("izer", "ize"), ("ator", "ate"), ("alli", "al"), ("bli", "ble"), ("ogi", "og"), ("li", ""))
_STEP3 = (("ational", "ate"), ("tional", "tion"), ("alize", "al"), ("icate", "ic"), ("iciti", "ic"),
          ("ative", ""), ("ical", "ic"), ("ness", ""), ("ful", ""))
_STEP4 = ("ement", "ance", "ence", "able", "ible", "ment", "ant", "ent", "ism", "ate", "iti", "ous",
          "ive", "ize", "ion", "al", "er", "ic")


def _is_vowel(w: str, i: int) -> bool:
    return w[i] in _VOWELS


def _r1_r2(w: str):
    def region(start: int) -> int:
        for i in range(start + 1, len(w)):
            if not _is_vowel(w, i) and _is_vowel(w, i - 1):
                return i + 1
        return len(w)
    if w.startswith(("gener", "arsen")):
        r1 = 5
    elif w.startswith("commun"):
        r1 = 6
    else:
        r1 = region(0)
    return r1, region(r1)


def _has_vowel(s: str) -> bool:
    return any(c in _VOWELS for c in s)


def _ends_short_syllable(w: str) -> bool:
    n = len(w)
    if n == 2:
        return _is_vowel(w, 0) and not _is_vowel(w, 1)
    if n >= 3:
        return (not _is_vowel(w, n - 3) and _is_vowel(w, n - 2) and not _is_vowel(w, n - 1)
                and w[n - 1] not in "wxY")
    return False


def porter2_stem(word: str) -> str:
    """Snowball English ("Porter2") stemmer for one lower-case word."""
    if len(word) <= 2:
        return word
    if word in _EXCEPTIONS:
        return _EXCEPTIONS[word]
    w = word[1:] if word.startswith("'") else word
    if len(w) <= 2:
        return w
    # mark consonant-y as Y
    chars = list(w)
    if chars[0] == "y":
        chars[0] = "Y"
    for i in range(1, len(chars)):
        if chars[i] == "y" and chars[i - 1] in _VOWELS:
            chars[i] = "Y"
    w = "".join(chars)
    r1, r2 = _r1_r2(w)

    # step 0
    for suf in ("'s'", "'s", "'"):
        if w.endswith(suf):
            w = w[:-len(suf)]
            break
    # step 1a
    if w.endswith("sses"):
        w = w[:-2]
    elif w.endswith(("ied", "ies")):
        w = w[:-2] if len(w) > 4 else w[:-1]
    elif w.endswith(("us", "ss")):
        pass
    elif w.endswith("s"):
        if _has_vowel(w[:-2]):
            w = w[:-1]
    if w in _EXCEPTIONS_1A:
        return w.replace("Y", "y")
    # step 1b
    if w.endswith("eedly"):
        if len(w) - 5 >= r1:
            w = w[:-3]
    elif w.endswith("eed"):
        if len(w) - 3 >= r1:
            w = w[:-1]
    else:
        for suf in ("ingly", "edly", "ing", "ed"):
            if w.endswith(suf):
                stem = w[:-len(suf)]
                if _has_vowel(stem):
                    w = stem
                    if w.endswith(("at", "bl", "iz")):
                        w += "e"
                    elif w.endswith(_DOUBLES):
                        w = w[:-1]
                    elif _ends_short_syllable(w) and r1 >= len(w):
                        w += "e"
                break
    # step 1c
    if len(w) > 2 and w[-1] in "yY" and w[-2] not in _VOWELS:
        w = w[:-1] + "i"
    # step 2
    for suf, rep in _STEP2:
        if w.endswith(suf):
            if len(w) - len(suf) >= r1:
                if suf == "ogi":
                    if w[:-3].endswith("l"):
                        w = w[:-3] + rep
                elif suf == "li":
                    if len(w) > 2 and w[-3] in _LI_ENDINGS:
                        w = w[:-2]
                else:
                    w = w[:-len(suf)] + rep
            break
    # step 3
    for suf, rep in _STEP3:
        if w.endswith(suf):
            if len(w) - len(suf) >= r1:
                if suf == "ative":
                    if len(w) - 5 >= r2:
                        w = w[:-5]
                else:
                    w = w[:-len(suf)] + rep
            break
    # step 4
    for suf in _STEP4:
        if w.endswith(suf):
            if len(w) - len(suf) >= r2:
                if suf == "ion":
                    if len(w) > 3 and w[-4] in "st":
                        w = w[:-3]
                else:
                    w = w[:-len(suf)]
            break
    # step 5
    if w.endswith("e"):
        if len(w) - 1 >= r2 or (len(w) - 1 >= r1 and not _ends_short_syllable(w[:-1])):
            w = w[:-1]
    elif w.endswith("l"):
        if len(w) - 1 >= r2 and len(w) > 1 and w[-2] == "l":
            w = w[:-1]
    return w.replace("Y", "y")


_TEXT_LIB = None


def _text_lib():
    """libprtext.so (csrc/textproc.c): the corpus tokenizer in C.  None when it has not been built or was generated
    from another Unicode database than this interpreter's -- the Python path below is then used for every document
    (it is also what the C path hands the documents it cannot express exactly to)."""
    global _TEXT_LIB
    if _TEXT_LIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libprtext.so")
        lib = False
        if os.path.exists(path) and os.environ.get("PROBING_RAG_PY_TOKENIZER", "0") != "1":
            lib = ctypes.CDLL(path)
            c, vp, i64 = ctypes, ctypes.c_void_p, ctypes.c_int64
            lib.pt_unidata_version.restype = c.c_char_p
            lib.pt_create.restype = vp
            lib.pt_destroy.argtypes = [vp]
            lib.pt_size.restype, lib.pt_size.argtypes = i64, [vp]
            lib.pt_encode.restype = i64
            lib.pt_encode.argtypes = [vp, vp, vp, i64, i64, vp, i64, c.POINTER(i64), vp]
            lib.pt_intern_many.restype, lib.pt_intern_many.argtypes = c.c_int, [vp, vp, vp, i64, vp]
            lib.pt_tokens_since.restype, lib.pt_tokens_since.argtypes = i64, [vp, i64, vp, i64, vp]
            lib.pt_stem_many.restype, lib.pt_stem_many.argtypes = None, [vp, vp, i64, vp, vp]
            if lib.pt_unidata_version().decode() != unicodedata.unidata_version:
                lib = False
        _TEXT_LIB = lib
    return _TEXT_LIB or None


class BuiltinStemmer:
    """PyStemmer-shaped object: `.stemWords(list[str]) -> list[str]`.  Batches go through the C restatement of the
    same algorithm (csrc/textproc.c:pt_stem_many) when libprtext.so is built; words with non-ASCII letters and
    single words stay on `porter2_stem`."""

    def __init__(self):
        self._cache: dict[str, str] = {}
        self._lib = _text_lib()

    def stemWord(self, word: str) -> str:
        s = self._cache.get(word)
        if s is None:
            s = self._cache[word] = porter2_stem(word)
        return s

    def stemWords(self, words):
        if self._lib is None or len(words) < 64:
            return [self.stemWord(w) for w in words]
        words = list(words)
        idx = [i for i, w in enumerate(words) if w.isascii()]
        enc = [words[i].encode("ascii") for i in idx]
        n = len(enc)
        offs = np.zeros(n + 1, np.int64)
        np.cumsum(np.fromiter(map(len, enc), np.int64, n), out=offs[1:])
        out = ctypes.create_string_buffer(int(offs[-1]) + 2 * n + 1)
        oo = np.empty(n + 1, np.int64)
        self._lib.pt_stem_many(b"".join(enc), offs.ctypes.data, n, out, oo.ctypes.data)
        text = out.raw[:int(oo[-1])].decode("ascii")
        ol = oo.tolist()
        res = [None] * len(words)
        for j, i in enumerate(idx):
            res[i] = text[ol[j]:ol[j + 1]]
        for i, w in enumerate(words):
            if res[i] is None:
                res[i] = self.stemWord(w)
        return res


def get_stemmer():
    try:
        import Stemmer  # PyStemmer, what llama-index uses (App. A.1)
        return Stemmer.Stemmer("english")
    except Exception:
        return BuiltinStemmer()


def split_tokens(text: str, stopwords=STOPWORDS_EN) -> list[str]:
    """lower -> regex -> stop-word removal (before stemming, App. A.2)."""
    return [t for t in TOKEN_PATTERN.findall(text.lower()) if t not in stopwords]


class _AutoId(dict):
    """surface token -> dense id in first-seen order; `map(d.__getitem__, tokens)` runs in C."""

    def __missing__(self, key):
        v = self[key] = len(self)
        return v


class Vocabulary:
    """stem -> term id.  bm25s assigns ids from a `set` of stems (hash order, App. A.2); here
    ids are first-seen order, which changes no score.

    The corpus side works on batches of documents and numpy arrays (21M passages are ~1.5 G
    tokens: a Python list of ints would be tens of GB): surface tokens get dense ids through a
    dict, and one int32 table maps surface id -> stem id (-1 = stop word), so stop-word
    removal, stemming and the id remap of a whole batch are three numpy gathers."""

    def __init__(self, stemmer=None):
        self.stemmer = stemmer if stemmer is not None else get_stemmer()
        self.stem_to_id: dict[str, int] = {}
        self._surf = _AutoId()                       # surface token -> surface id
        self._surf_stem = np.zeros(0, np.int32)      # surface id -> stem id, -1 = stop word
        self._n_mapped = 0
        self._lib = _text_lib()
        self._tab = self._lib.pt_create() if self._lib is not None else None   # C surface-token table

    def __del__(self):
        if getattr(self, "_tab", None):
            self._lib.pt_destroy(self._tab)
            self._tab = None

    def __len__(self) -> int:
        return len(self.stem_to_id)

    # ------------------------------------------------------------------ corpus side
    def _map_new_surfaces(self) -> None:
        """Stem the surface tokens seen since the last call (each unique token once, in
        first-seen order, stop words dropped BEFORE stemming -- App. A.2) and extend the table."""
        n = len(self._surf)
        if n == self._n_mapped:
            return
        if n > self._surf_stem.size:
            grown = np.full(max(n, 2 * self._surf_stem.size, 1024), -1, np.int32)
            grown[:self._n_mapped] = self._surf_stem[:self._n_mapped]
            self._surf_stem = grown
        new = list(self._surf)[self._n_mapped:] if self._n_mapped else list(self._surf)
        keep = [i for i, t in enumerate(new) if t not in STOPWORDS_EN]
        stems = self.stemmer.stemWords([new[i] for i in keep])
        s2i = self.stem_to_id
        out = np.full(len(new), -1, np.int32)
        for i, st in zip(keep, stems):
            j = s2i.get(st)
            if j is None:
                j = s2i[st] = len(s2i)
            out[i] = j
        self._surf_stem[self._n_mapped:n] = out
        self._n_mapped = n

    def _surface_ids_py(self, texts):
        find = TOKEN_PATTERN.findall
        toks, counts = [], []
        for t in texts:
            w = find(t.lower())
            toks.extend(w)
            counts.append(len(w))
        sid = np.fromiter(map(self._surf.__getitem__, toks), dtype=np.int64, count=len(toks))
        return sid, np.asarray(counts, np.int64)

    def _surface_ids_c(self, texts):
        """The same (surface ids, tokens per document) through csrc/textproc.c.  Documents the C tables cannot
        express exactly (see its header) go through `re` here, one at a time and in place, so surface ids keep
        their first-seen order; the new surface strings are copied back into `self._surf` afterwards (the query
        side and the stemmer work on Python strings)."""
        lib, tab = self._lib, self._tab
        enc = [t.encode("utf-8") for t in texts]
        n = len(enc)
        offs = np.zeros(n + 1, np.int64)
        np.cumsum(np.fromiter(map(len, enc), np.int64, n), out=offs[1:])
        buf = b"".join(enc)
        cap = len(buf) // 2 + 1                       # a token is at least two bytes
        ids = np.empty(cap, np.int32)
        counts = np.zeros(n, np.int32)
        n_tok = ctypes.c_int64(0)
        find = TOKEN_PATTERN.findall
        i = 0
        while i < n:
            i = lib.pt_encode(tab, buf, offs.ctypes.data, i, n, ids.ctypes.data, cap, ctypes.byref(n_tok),
                              counts.ctypes.data)
            if i < 0:
                raise MemoryError("libprtext: out of memory")
            if i < n:                                 # this document needs str.lower() / re
                w = [t.encode("utf-8") for t in find(texts[i].lower())]
                if w:
                    wo = np.zeros(len(w) + 1, np.int64)
                    np.cumsum(np.fromiter(map(len, w), np.int64, len(w)), out=wo[1:])
                    out = np.empty(len(w), np.int32)
                    if lib.pt_intern_many(tab, b"".join(w), wo.ctypes.data, len(w), out.ctypes.data) != 0:
                        raise MemoryError("libprtext: out of memory")
                    if n_tok.value + len(w) > cap:    # cannot happen (>= 2 bytes per token), but never overrun
                        raise RuntimeError("token buffer too small")
                    ids[n_tok.value:n_tok.value + len(w)] = out
                    n_tok.value += len(w)
                counts[i] = len(w)
                i += 1
        have = len(self._surf)
        total = lib.pt_size(tab)
        if total > have:
            nbytes = lib.pt_tokens_since(tab, have, None, 0, None)
            raw = ctypes.create_string_buffer(max(int(nbytes), 1))
            to = np.empty(total - have + 1, np.int64)
            if lib.pt_tokens_since(tab, have, raw, nbytes, to.ctypes.data) != nbytes:
                raise RuntimeError("libprtext: token table changed underneath")
            text = raw.raw[:nbytes]
            surf = self._surf
            tl = to.tolist()
            for j in range(total - have):
                surf[text[tl[j]:tl[j + 1]].decode("utf-8")] = have + j
        return ids[:n_tok.value].astype(np.int64), counts.astype(np.int64)

    def encode_corpus_batch(self, texts) -> tuple[np.ndarray, np.ndarray]:
        """Documents -> (term ids i32[total], doc lengths i32[n]), docs back to back; new
        stems are added to the vocabulary."""
        if not isinstance(texts, (list, tuple)):
            texts = list(texts)
        if self._tab is not None and len(self._surf) == self._lib.pt_size(self._tab):
            sid, counts = self._surface_ids_c(texts)
        else:                                         # no C library
            sid, counts = self._surface_ids_py(texts)
        self._map_new_surfaces()
        stem = self._surf_stem[sid]
        keep = stem >= 0
        doc_of = np.repeat(np.arange(len(counts)), counts)
        lens = np.bincount(doc_of[keep], minlength=len(counts)).astype(np.int32)
        return stem[keep], lens

    def encode_corpus(self, texts, batch_docs: int = 4096, progress=None) -> tuple[np.ndarray, np.ndarray]:
        """Stream an iterable of documents through `encode_corpus_batch`; the result is two
        flat numpy arrays (what `BM25Index.from_tokens` takes), never a Python token list."""
        tok_parts, len_parts, batch = [], [], []
        n_done = 0

        def flush():
            nonlocal n_done
            t, l = self.encode_corpus_batch(batch)
            tok_parts.append(t)
            len_parts.append(l)
            n_done += len(batch)
            batch.clear()
            if progress is not None:
                progress(n_done)

        for t in texts:
            batch.append(t)
            if len(batch) >= batch_docs:
                flush()
        if batch:
            flush()
        if not tok_parts:
            return np.zeros(0, np.int32), np.zeros(0, np.int32)
        return np.concatenate(tok_parts), np.concatenate(len_parts)

    def encode_corpus_doc(self, text: str) -> list[int]:
        return self.encode_corpus_batch([text])[0].tolist()

    # ------------------------------------------------------------------ query side
    def encode_query(self, text: str) -> list[int]:
        """Unknown stems are dropped; order and duplicates kept (App. A.5)."""
        out = []
        s2i = self.stem_to_id
        for tok in split_tokens(text):
            sid = self._surf.get(tok)
            if sid is not None and sid < self._n_mapped:
                i = int(self._surf_stem[sid])
            else:   # a surface form the corpus never had may still stem to a known term
                i = s2i.get(self.stemmer.stemWords([tok])[0], -1)
            if i >= 0:
                out.append(i)
        return out

    def encode_queries(self, queries) -> tuple[np.ndarray, np.ndarray]:
        """Query strings -> CSR batch (q_indptr i64[B+1], q_terms i32[nnz])."""
        ids = [self.encode_query(q) for q in queries]
        q_indptr = np.zeros(len(ids) + 1, dtype=np.int64)
        np.cumsum([len(x) for x in ids], out=q_indptr[1:])
        q_terms = np.fromiter((t for x in ids for t in x), dtype=np.int32, count=int(q_indptr[-1]))
        return q_indptr, q_terms

    # ------------------------------------------------------------------ persistence
    def save(self, path: str) -> None:
        """One stem per line in term-id order (stems are \\w+ tokens: no newline inside)."""
        stems = [None] * len(self.stem_to_id)
        for s, i in self.stem_to_id.items():
            stems[i] = s
        with open(path, "w", encoding="utf-8", newline="\n") as f:
            f.write("\n".join(stems))
            if stems:
                f.write("\n")

    @classmethod
    def load(cls, path: str, stemmer=None) -> "Vocabulary":
        v = cls(stemmer)
        with open(path, encoding="utf-8", newline="\n") as f:
            v.stem_to_id = {line.rstrip("\n"): i for i, line in enumerate(f)}
        return v
