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

import re

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


class BuiltinStemmer:
    """PyStemmer-shaped object: `.stemWords(list[str]) -> list[str]`."""

    def __init__(self):
        self._cache: dict[str, str] = {}

    def stemWord(self, word: str) -> str:
        s = self._cache.get(word)
        if s is None:
            s = self._cache[word] = porter2_stem(word)
        return s

    def stemWords(self, words):
        return [self.stemWord(w) for w in words]


def get_stemmer():
    try:
        import Stemmer  # PyStemmer, what llama-index uses (App. A.1)
        return Stemmer.Stemmer("english")
    except Exception:
        return BuiltinStemmer()


def split_tokens(text: str, stopwords=STOPWORDS_EN) -> list[str]:
    """lower -> regex -> stop-word removal (before stemming, App. A.2)."""
    return [t for t in TOKEN_PATTERN.findall(text.lower()) if t not in stopwords]


class Vocabulary:
    """stem -> term id.  bm25s assigns ids from a `set` of stems (hash order, App. A.2); here
    ids are first-seen order, which changes no score."""

    def __init__(self, stemmer=None):
        self.stemmer = stemmer if stemmer is not None else get_stemmer()
        self.stem_to_id: dict[str, int] = {}
        self._surface: dict[str, str] = {}

    def __len__(self) -> int:
        return len(self.stem_to_id)

    def _stem(self, tok: str) -> str:
        s = self._surface.get(tok)
        if s is None:
            s = self._surface[tok] = self.stemmer.stemWords([tok])[0]
        return s

    def encode_corpus_doc(self, text: str) -> list[int]:
        ids = []
        for tok in split_tokens(text):
            s = self._stem(tok)
            i = self.stem_to_id.get(s)
            if i is None:
                i = self.stem_to_id[s] = len(self.stem_to_id)
            ids.append(i)
        return ids

    def encode_query(self, text: str) -> list[int]:
        """Unknown stems are dropped; order and duplicates kept (App. A.5)."""
        out = []
        for tok in split_tokens(text):
            i = self.stem_to_id.get(self._stem(tok))
            if i is not None:
                out.append(i)
        return out
