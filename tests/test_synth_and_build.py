"""Synthetic workload generator and the torch index builder vs the oracle (CPU tensors)."""
import numpy as np
import torch

from oracle import bm25_oracle as bo
from probing_rag_b200 import synth
from probing_rag_b200.index import bm25_weights, count_postings, idf_lucene_table


def test_corpus_shape_and_determinism():
    t1, l1 = synth.corpus_np(20_000, 1 << 16)
    t2, l2 = synth.corpus_np(20_000, 1 << 16)
    assert np.array_equal(t1, t2) and np.array_equal(l1, l2)
    assert l1.min() >= 8 and l1.max() <= 128 and abs(l1.mean() - 70) < 1
    assert t1.min() >= 0 and t1.max() < (1 << 16)


def test_doc_range_slices_equal_full_corpus():
    n = synth.DOC_BLOCK + 5000          # spans two generation blocks
    t, l = synth.corpus_np(n, 1 << 14)
    off = np.concatenate([[0], np.cumsum(l, dtype=np.int64)])
    for lo, hi in ((0, 1000), (synth.DOC_BLOCK - 10, synth.DOC_BLOCK + 10), (synth.DOC_BLOCK, n)):
        ts, ls = synth.corpus_np(n, 1 << 14, doc_lo=lo, doc_hi=hi)
        assert np.array_equal(ls, l[lo:hi]) and np.array_equal(ts, t[off[lo]:off[hi]])


def test_queries_shape():
    qi, qt = synth.queries_np(5000, 1 << 20)
    ln = np.diff(qi)
    assert ln.min() >= 1 and ln.max() <= 32 and abs(ln.mean() - 6) < 0.2
    qi2, qt2 = synth.queries_np(200, 1 << 20, kind="later")
    ln2 = np.diff(qi2)
    assert ln2.min() >= 64 and ln2.max() <= 1024 and abs(ln2.mean() - 350) < 30


def test_torch_builder_bit_exact_vs_oracle(small_corpus):
    toks, lens, vocab, ora = (small_corpus[k] for k in ("tokens", "doc_lens", "vocab", "index"))
    term, doc, tf, df = count_postings(torch.from_numpy(toks), torch.from_numpy(lens), vocab)
    assert np.array_equal(doc.numpy(), ora["indices"])
    assert np.array_equal(df.numpy(), ora["df"])
    idf = idf_lucene_table(df.numpy(), len(lens))
    assert np.array_equal(idf, bo.idf_lucene(ora["df"], len(lens)))
    w = bm25_weights(term, doc, tf, torch.from_numpy(lens), torch.from_numpy(idf),
                     float(torch.from_numpy(lens).double().mean().item()), chunk=1 << 20)
    assert np.array_equal(w.numpy(), ora["data"])
    indptr = np.concatenate([[0], np.cumsum(df.numpy())])
    assert np.array_equal(indptr, ora["indptr"])


def test_torch_block_generator_shapes():
    cdf = torch.from_numpy(synth.zipf_mandelbrot_cdf(1 << 14))
    t, l = synth.corpus_block_torch(0, 1000, cdf)
    t2, l2 = synth.corpus_block_torch(0, 1000, cdf)
    assert torch.equal(t, t2) and torch.equal(l, l2)
    assert l.numel() == 1000 and int(l.sum()) == t.numel() and int(t.max()) < (1 << 14)
