"""Text front-end: bm25s.tokenize semantics and the Porter2 restatement."""
import os

import numpy as np

from probing_rag_b200.text import (STOPWORDS_EN, BuiltinStemmer, Vocabulary, porter2_stem,
                                   split_tokens)

# (word, stem) pairs from the published Snowball English vocabulary/output lists
PORTER2_PAIRS = [
    ("consign", "consign"), ("consigned", "consign"), ("consigning", "consign"),
    ("consist", "consist"), ("consisted", "consist"), ("consistency", "consist"),
    ("consolation", "consol"), ("consolations", "consol"), ("consolatory", "consolatori"),
    ("knack", "knack"), ("knackeries", "knackeri"), ("knaves", "knave"), ("knavish", "knavish"),
    ("kneel", "kneel"), ("kneeled", "kneel"), ("knew", "knew"), ("knife", "knife"), ("knightly", "knight"),
    ("knots", "knot"), ("knowing", "know"), ("knowledge", "knowledg"),
    ("generously", "generous"), ("generate", "generat"), ("communication", "communic"),
    ("caresses", "caress"), ("ponies", "poni"), ("ties", "tie"), ("cries", "cri"), ("gas", "gas"),
    ("gaps", "gap"), ("kiwis", "kiwi"), ("agreed", "agre"), ("feed", "feed"), ("hopping", "hop"),
    ("hoping", "hope"), ("luxuriated", "luxuri"), ("relational", "relat"), ("conditional", "condit"),
    ("rational", "ration"), ("happy", "happi"), ("cry", "cri"), ("by", "by"), ("say", "say"),
    ("sky", "sky"), ("dying", "die"), ("news", "news"), ("succeed", "succeed"), ("running", "run"),
    ("national", "nation"), ("electricity", "electr"), ("running", "run"), ("wikipedia", "wikipedia"),
    ("was", "was"), ("retrieval", "retriev"), ("probing", "probe"), ("augmented", "augment"),
]


def test_porter2_known_pairs():
    bad = [(w, porter2_stem(w), s) for w, s in PORTER2_PAIRS if porter2_stem(w) != s]
    assert not bad, bad


def test_split_tokens_semantics():
    assert len(STOPWORDS_EN) == 33
    toks = split_tokens("The U.S. is a BIG-country, isn't it? x yy 42 été")
    # >=2 word chars, lower-cased, stop words dropped, unicode kept
    assert toks == ["big", "country", "isn", "yy", "42", "été"]


def test_vocabulary_query_drops_unknown_and_keeps_duplicates():
    v = Vocabulary(BuiltinStemmer())
    d0 = v.encode_corpus_doc("Cats are running in the gardens, cats run")
    assert len(v) == 3 and d0 == [0, 1, 2, 0, 1]       # cat, run, garden
    assert v.encode_query("running cats and dogs cats") == [1, 0, 0]
    assert v.encode_query("the of and") == []


# ---------------------------------------------------------------------------------------------------------------
# csrc/textproc.c (libprtext.so): the corpus tokenizer and the stemmer in C must equal the Python restatement
# ---------------------------------------------------------------------------------------------------------------
def _c_text_lib():
    from probing_rag_b200 import build as b
    import probing_rag_b200.text as T
    b.build_text()
    T._TEXT_LIB = None
    os.environ.pop("PROBING_RAG_PY_TOKENIZER", None)
    lib = T._text_lib()
    assert lib is not None, "libprtext.so did not build / load"
    return lib


def _py_vocabulary():
    import probing_rag_b200.text as T
    os.environ["PROBING_RAG_PY_TOKENIZER"] = "1"
    T._TEXT_LIB = None
    try:
        v = T.Vocabulary()
        assert v._tab is None and v.stemmer._lib is None
        return v
    finally:
        os.environ.pop("PROBING_RAG_PY_TOKENIZER", None)
        T._TEXT_LIB = None


def test_c_tokenizer_equals_the_python_path_on_unicode_text():
    import random
    from probing_rag_b200.text import Vocabulary
    _c_text_lib()
    rnd = random.Random(11)
    special = [0x130, 0x3a3, 0x3c3, 0x3c2, 0xdf, 0x1e9e, 0xc9, 0xe9, 0x416, 0x4e2d, 0x660, 0x1F600, 0x10400, 0x2160, 0xb2,
               0x5f, 0x301, 0x200b, 0xa0, 0x1c5, 0x2c6, 0xaa, 0x345, 0x3a9, 0x212a, 0xfb01, 0x7ff, 0x800, 0xffff, 0xd7ff]
    pool = [chr(c) for c in list(range(32, 127)) * 5 + special]
    docs = ["".join(rnd.choice(pool) for _ in range(rnd.randint(0, 120))) for _ in range(4000)]
    docs += ["", "a", "ab", "Σ", "AΣ", "İstanbul ΑΣ ΣΑ", "x" * 3000, "y" * 1022 + " " + "z" * 1021, "日本語 テスト",
             "naïve café", "a_b __ _1", "The Eiffel Tower's height (330 m) — measured in 2022."]
    vc = Vocabulary()
    assert vc._tab is not None
    tc, lc = vc.encode_corpus(docs, batch_docs=257)
    vp = _py_vocabulary()
    tp, lp = vp.encode_corpus(docs, batch_docs=257)
    assert np.array_equal(tc, tp) and np.array_equal(lc, lp)
    assert vc.stem_to_id == vp.stem_to_id
    assert list(vc._surf) == list(vp._surf)                  # same surface ids, same first-seen order
    for q in ["İstanbul café ΑΣ", "naïve x", "AB ab tower's", "テスト"]:
        assert vc.encode_query(q) == vp.encode_query(q)


def test_c_stemmer_equals_porter2_stem_word_for_word():
    import random
    from probing_rag_b200.text import BuiltinStemmer, porter2_stem
    _c_text_lib()
    bases = ("gener arsen commun sky ski dy ly ty agree proceed exceed succeed inning herring canning hop hope happ cry say "
             "boy try relat condition ration val digit conform radic differ adopt adjust depend activ allow replac nation "
             "sensibl feudal decis formal callous operat electr rational vile fail fill controll roll beauti y yy yes by ay "
             "toy ied ies ss us a's o'clock 'tis").split()
    sufs = ("s es ed ing ingly edly eed eedly ly li y ies ied ness ful fulness ization ational tional alism aliti iviti "
            "biliti ousli entli lessli fulli enci anci abli izer ator alli bli ogi logi ative icate iciti ical alize ement "
            "ment ance ence able ible ant ent ism ate iti ous ive ize ion sion tion al er ic e l ll 's 's' '").split()
    words = [b + s for b in bases for s in sufs] + [b + s + t for b in bases for s in sufs for t in ("s", "ed", "ing")]
    words += bases + sufs + [w for w, _ in PORTER2_PAIRS]
    rnd = random.Random(5)
    words += ["".join(rnd.choice("aeiouylstrngdbe'") for _ in range(rnd.randint(1, 12))) for _ in range(60000)]
    words += ["naïve", "café", "日本語", "x9", "2022", "_a_"]
    words = list(dict.fromkeys(words))
    st = BuiltinStemmer()
    assert st._lib is not None
    got = st.stemWords(words)
    want = [porter2_stem(w) for w in words]
    bad = [(w, g, x) for w, g, x in zip(words, got, want) if g != x]
    assert not bad, bad[:10]
