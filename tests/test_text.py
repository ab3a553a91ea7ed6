"""Text front-end: bm25s.tokenize semantics and the Porter2 restatement."""
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
