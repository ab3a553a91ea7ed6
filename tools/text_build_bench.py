"""Corpus-ingestion timing at a real scale: N synthetic ~100-word passages in the DPR
`psgs_w100.tsv` shape -> read_wiki_tsv -> BM25Retriever.from_texts(persist_dir=...) (tokenise + stem + GPU index
build + persist) -> a FRESH load with from_persist_dir -> retrieve(str) with passage texts.

    python tools/text_build_bench.py [--n 1000000] [--out gpurun_out/text_build.json]
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from probing_rag_b200 import BM25Retriever  # noqa: E402
from probing_rag_b200.corpus import read_wiki_tsv  # noqa: E402

STEMS = ("retriev augment generat prob languag model hidden state wikipedia passag question answer capital citi river tower "
         "bridg nation histori scienc music film footbal univers govern presid centuri war famili compani station").split()
SUFFIXES = ["", "s", "ed", "ing", "ation", "al", "ly", "er", "ers", "ment"]
STOPS = "the of and in to a is was for as on with by that it at from an be this".split()


def make_words(n_types: int, rng) -> np.ndarray:
    letters = np.array(list("abcdefghijklmnopqrstuvwxyz"))
    words = [s + suf for s in STEMS for suf in SUFFIXES]
    while len(words) < n_types:
        k = int(rng.integers(3, 11))
        words.append("".join(rng.choice(letters, size=k)))
    return np.array(words[:n_types], dtype=object)


def write_tsv(path: str, n: int, seed: int = 7) -> int:
    rng = np.random.default_rng(seed)
    words = make_words(200_000, rng).tolist() + STOPS
    n_w = len(words) - len(STOPS)
    p = (np.arange(n_w) + 30.0) ** -1.1
    p /= p.sum()
    cdf = np.cumsum(p)
    get = words.__getitem__
    n_bytes = 0
    with open(path, "w", encoding="utf-8") as f:
        f.write("id\ttext\ttitle\n")
        for lo in range(0, n, 50000):
            m = min(50000, n - lo)
            ids = np.minimum(np.searchsorted(cdf, rng.random((m, 100))), n_w - 1)
            stop_pos = rng.random((m, 100)) < 0.3                       # ~30% stop words, like running text
            ids[stop_pos] = n_w + rng.integers(0, len(STOPS), size=int(stop_pos.sum()))
            rows = ids.tolist()
            lines = [f"{lo + i + 1}\t{' '.join(map(get, rows[i]))}, {lo + i}.\tTitle {lo + i}\n" for i in range(m)]
            blob = "".join(lines)
            f.write(blob)
            n_bytes += len(blob)
    return n_bytes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "text_build.json"))
    args = ap.parse_args()
    work = tempfile.mkdtemp(prefix="prtext_")
    tsv = os.path.join(work, "psgs_w100.tsv")
    t0 = time.perf_counter()
    n_bytes = write_tsv(tsv, args.n)
    t_gen = time.perf_counter() - t0
    persist = os.path.join(work, "index")
    marks = {}

    def progress(n_done):
        if n_done >= args.n and "tokenised" not in marks:
            marks["tokenised"] = time.perf_counter()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = BM25Retriever.from_texts(read_wiki_tsv(tsv), similarity_top_k=5, persist_dir=persist, progress=progress)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    t_tok = marks.get("tokenised", t0) - t0
    n_tokens = int(r.index.meta.get("avgdl", 0) * args.n)
    queries = ["what is the capital city of the nation", "history of football universities", "retrieval augmented generation models",
               "presidents and governments of the century", "music films and family companies"]
    want = [[(x.node.id_, x.score, x.text) for x in one] for one in r.retrieve_batch(queries)]
    nnz, n_terms = r.index.nnz, r.index.n_terms
    del r
    torch.cuda.empty_cache()
    t0 = time.perf_counter()
    r2 = BM25Retriever.from_persist_dir(persist)
    torch.cuda.synchronize()
    t_load = time.perf_counter() - t0
    got = [[(x.node.id_, x.score, x.text) for x in one] for one in r2.retrieve_batch(queries)]
    assert got == want, "the loaded retriever returns different nodes"
    r2.retrieve(queries[0])
    t0 = time.perf_counter()
    for q in queries * 4:
        res = r2.retrieve(q)
    t_q = (time.perf_counter() - t0) / (4 * len(queries))
    assert len(res) == 5 and all(x.text for x in res)
    sizes = {f: os.path.getsize(os.path.join(persist, f)) for f in sorted(os.listdir(persist))}
    out = {"n_passages": args.n, "tsv_bytes": n_bytes, "synthetic_tsv_written_s": t_gen,
           "build_total_s": t_build, "read_tokenise_stem_s": t_tok, "gpu_index_build_and_persist_s": t_build - t_tok,
           "passages_per_s_tokenise": args.n / max(t_tok, 1e-9), "mb_per_s_tokenise": n_bytes / 1e6 / max(t_tok, 1e-9),
           "tokens_after_stopwords": n_tokens, "n_terms": n_terms, "nnz": nnz,
           "load_from_persist_dir_s": t_load, "retrieve_str_ms": t_q * 1e3, "loaded_results_identical_with_text": True,
           "persisted_files_bytes": sizes, "host_cores": os.cpu_count(),
           "stemmer": type(r2.vocab.stemmer).__name__,
           "projection_21M_passages_tokenise_s": t_tok * 21_015_324 / args.n}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
