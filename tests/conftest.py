import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the gpu-marked tests are skipped, not failed (they run on the B200 box)."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def small_corpus():
    """100k-doc config-1-shaped corpus + oracle index + 1,000 queries."""
    from oracle import bm25_oracle as bo
    from probing_rag_b200 import synth
    n_docs, vocab = 100_000, 1 << 20
    toks, lens = synth.corpus_np(n_docs, vocab)
    idx = bo.build_index(toks, lens, vocab)
    q_indptr, q_terms = synth.queries_np(1000, vocab, idx["df"])
    return {"tokens": toks, "doc_lens": lens, "vocab": vocab, "index": idx,
            "q_indptr": q_indptr, "q_terms": q_terms}
