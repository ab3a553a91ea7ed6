"""Parity of the CUDA BM25 path (through the C ABI) with the CPU oracle.  Needs a B200.

Bar (BASELINE.json north_star): returned doc ids and ranks bit-exact; scores within 1e-5
relative.  The kernel adds postings in the oracle's order in fp32, so scores are asserted
bit-equal as well (SCORE_RTOL documents the contractual tolerance).
"""
import os

import numpy as np
import pytest
import torch

from oracle import bm25_oracle as bo
from oracle import c_oracle as co
from probing_rag_b200 import synth

pytestmark = pytest.mark.gpu

SCORE_RTOL = 1e-5

DEFAULTS = dict(subs_per_item=24, warps_per_cta=8, docs_per_launch=393216, min_items=32768, items_per_warp=1, tile_epochs=4, batch_variant=3)

# launch plans that exercise every path of the live-threshold protocol (bm25_lean.cuh): one launch with all
# thresholds travelling between warps, one chunk per launch (thresholds only through the merge), everything between
TUNINGS = [
    dict(),                                                                               # the default plan
    dict(subs_per_item=12, warps_per_cta=8, docs_per_launch=98304, min_items=2048),
    dict(subs_per_item=1, warps_per_cta=4, docs_per_launch=2048, min_items=1, items_per_warp=1),   # one sub-tile per item and launch: 49 launches
    dict(subs_per_item=5, warps_per_cta=12, docs_per_launch=1000000, min_items=100000),  # one launch
    dict(subs_per_item=3, warps_per_cta=12, docs_per_launch=20000, min_items=2048),
    dict(subs_per_item=24, warps_per_cta=8, items_per_warp=64),                           # items cut down to one sub-tile, one launch
    dict(subs_per_item=2, warps_per_cta=8, docs_per_launch=8192, min_items=1, items_per_warp=1),
    dict(subs_per_item=24, warps_per_cta=8, docs_per_launch=49152, min_items=1, items_per_warp=1),  # long items, 3 launches
    # the large-batch kernel variants (what a 65,536-query batch runs) on these small batches: four and two tile epochs
    dict(batch_variant=2),
    dict(batch_variant=2, tile_epochs=2),
    dict(batch_variant=2, subs_per_item=7, warps_per_cta=12, docs_per_launch=30000, min_items=1),
    dict(batch_variant=2, subs_per_item=1, warps_per_cta=4, docs_per_launch=2048, min_items=1, items_per_warp=1),
    dict(batch_variant=2, subs_per_item=24, warps_per_cta=8, docs_per_launch=49152, min_items=1, tile_epochs=2),
    dict(batch_variant=1, subs_per_item=24, docs_per_launch=49152, min_items=1),          # small-batch variant over several launches
]


def tune(gi, **kw):
    gi.set_tuning(**dict(DEFAULTS, **kw))


def gpu_index(idx, **kw):
    from probing_rag_b200 import BM25Index
    return BM25Index.from_arrays(idx["data"], idx["indices"], idx["indptr"], idx["num_docs"], **kw)


def to_dev(gi, q_indptr, q_terms):
    dev = gi.device
    return (torch.from_numpy(np.ascontiguousarray(q_indptr, dtype=np.int64)).to(dev),
            torch.from_numpy(np.ascontiguousarray(q_terms, dtype=np.int32)).to(dev))


def run_gpu(gi, q_indptr, q_terms, k, **kw):
    s, d = gi.topk(*to_dev(gi, q_indptr, q_terms), k, **kw)
    torch.cuda.synchronize()
    return s.cpu().numpy(), d.cpu().numpy()


def assert_parity(gs, gd, os_, od):
    assert np.array_equal(gd, od), f"doc ids differ in {np.flatnonzero((gd != od).any(axis=1))[:8]}"
    assert np.allclose(gs, os_, rtol=SCORE_RTOL, atol=0)
    assert np.array_equal(gs, os_), "scores not bit-identical"


@pytest.fixture(scope="module")
def corpus_gpu(small_corpus):
    return gpu_index(small_corpus["index"])


def test_golden_fixture_on_gpu(golden_dir):
    g = np.load(os.path.join(golden_dir, "bm25_golden.npz"))
    idx = {"data": g["data"], "indices": g["indices"], "indptr": g["indptr"], "num_docs": len(g["doc_lens"])}
    gi = gpu_index(idx)
    for tun in (dict(), dict(subs_per_item=1, warps_per_cta=4, docs_per_launch=2048, min_items=1, items_per_warp=1)):
        tune(gi, **tun)
        s, d = run_gpu(gi, g["q_indptr"], g["q_terms"], int(g["k"]))
        assert_parity(s, d, g["scores"], g["ids"])


@pytest.mark.parametrize("tun", TUNINGS, ids=lambda t: "-".join(f"{k[:4]}{v}" for k, v in t.items()) or "default")
def test_config1_100k_docs_1k_queries(small_corpus, corpus_gpu, tun):
    """BASELINE config 1: 100k passages, 1,000 queries, top-10, every launch plan."""
    tune(corpus_gpu, **tun)
    qi, qt = small_corpus["q_indptr"], small_corpus["q_terms"]
    os_, od = co.retrieve_batch(small_corpus["index"], qi, qt, 10, n_threads=8)
    gs, gd = run_gpu(corpus_gpu, qi, qt, 10)
    assert_parity(gs, gd, os_, od)


@pytest.mark.parametrize("k", [1, 5, 32, 33, 64, 100, 128])
def test_depth_sweep(small_corpus, corpus_gpu, k):
    qi, qt = small_corpus["q_indptr"][:65], small_corpus["q_terms"]
    os_, od = co.retrieve_batch(small_corpus["index"], qi, qt, k, n_threads=8)
    for tun in (TUNINGS[0], TUNINGS[1], TUNINGS[2], TUNINGS[4], TUNINGS[5]):
        tune(corpus_gpu, **tun)
        gs, gd = run_gpu(corpus_gpu, qi, qt, k)
        assert_parity(gs, gd, os_, od)


@pytest.mark.parametrize("nq", [1, 2, 37])
def test_small_batches(small_corpus, corpus_gpu, nq):
    """The reference's own shape: one query at a time (exp_rag.py:426).  A single query leaves one list per
    (sub-tile, query) item: the 32-warp merge takes over from 128 lists per query."""
    qi, qt = small_corpus["q_indptr"][:nq + 1], small_corpus["q_terms"]
    os_, od = co.retrieve_batch(small_corpus["index"], qi, qt, 10)
    for tun in TUNINGS:
        tune(corpus_gpu, **tun)
        gs, gd = run_gpu(corpus_gpu, qi, qt, 10)
        assert_parity(gs, gd, os_, od)


def test_single_query_over_many_sub_tiles_uses_the_wide_merge():
    """1 query x 400k documents = 196 one-sub-tile items in one launch (> 128 lists: bm25_merge_wide_kernel),
    for k up to 128 and queries of 1..24 terms."""
    n_docs, vocab = 400_000, 1 << 16
    toks, lens = synth.corpus_np(n_docs, vocab)
    idx = bo.build_index(toks, lens, vocab)
    gi = gpu_index(idx)
    assert gi.aux_info()["n_sub_tiles"] == 196
    qi, qt = synth.queries_np(24, vocab, idx["df"])
    for k in (1, 10, 100, 128):
        os_, od = co.retrieve_batch(idx, qi, qt, k, n_threads=8)
        for b in (1, 3, 24):
            sel = slice(0, b + 1)
            gs, gd = run_gpu(gi, qi[sel], qt[:qi[b]], k)
            assert gi.num_launches(b, k) == 1
            assert_parity(gs, gd, os_[:b], od[:b])


def test_large_batch_takes_the_large_batch_variant(small_corpus, corpus_gpu):
    """8,192 queries (more than twice the resident warps): the library picks the large-batch kernel variant by
    itself -- four tile epochs, bounds published once per item -- over one and over many launches."""
    idx = small_corpus["index"]
    qi, qt = synth.queries_np(8192, small_corpus["vocab"], idx["df"], seed=99)
    os_, od = co.retrieve_batch(idx, qi, qt, 10, n_threads=8)
    for tun in (dict(), dict(docs_per_launch=8192, min_items=1), dict(tile_epochs=2, docs_per_launch=16384, min_items=1),
                dict(subs_per_item=5, warps_per_cta=12, docs_per_launch=40000, min_items=1)):
        tune(corpus_gpu, **tun)
        gs, gd = run_gpu(corpus_gpu, qi, qt, 10)
        assert_parity(gs, gd, os_, od)
    os_, od = co.retrieve_batch(idx, qi, qt, 100, n_threads=8)
    tune(corpus_gpu, docs_per_launch=16384, min_items=1)
    gs, gd = run_gpu(corpus_gpu, qi, qt, 100)
    assert_parity(gs, gd, os_, od)


def test_long_transcript_queries(small_corpus, corpus_gpu):
    """Later-round queries are whole LM transcripts (exp_rag.py:428, 457): 64-1024 terms,
    many 32-term passes per sub-tile, posting cursors in the warp's scratch."""
    idx = small_corpus["index"]
    qi, qt = synth.queries_np(48, small_corpus["vocab"], idx["df"], kind="later")
    assert np.diff(qi).max() > 256
    os_, od = co.retrieve_batch(idx, qi, qt, 10, n_threads=8)
    for tun in (dict(), dict(subs_per_item=4, docs_per_launch=98304, min_items=2048),
                dict(subs_per_item=4, docs_per_launch=16384, min_items=1, items_per_warp=1),
                dict(batch_variant=2), dict(batch_variant=2, tile_epochs=2, subs_per_item=4, docs_per_launch=16384, min_items=1)):
        tune(corpus_gpu, **tun)
        gs, gd = run_gpu(corpus_gpu, qi, qt, 10)
        assert_parity(gs, gd, os_, od)


def test_edge_queries(small_corpus, corpus_gpu):
    """Empty query, duplicated terms, df==0 term, single rare term (zero-score tail)."""
    idx = small_corpus["index"]
    df = idx["df"]
    rare = int(np.flatnonzero(df == 1)[0])
    absent = int(np.flatnonzero(df == 0)[0])
    head = int(np.argmax(df))
    queries = [[], [head], [head, head, head], [rare], [absent], [rare, absent, rare], [head, rare, head],
               list(np.flatnonzero(df == 2)[:3])]
    qi = np.zeros(len(queries) + 1, np.int64)
    qi[1:] = np.cumsum([len(q) for q in queries])
    qt = np.array([t for q in queries for t in q], dtype=np.int32)
    os_, od = bo.retrieve_batch(idx, qi, qt, 10)
    for tun in TUNINGS:
        tune(corpus_gpu, **tun)
        gs, gd = run_gpu(corpus_gpu, qi, qt, 10)
        assert_parity(gs, gd, os_, od)
    assert od[0].tolist() == list(range(10)) and gs[0].tolist() == [0.0] * 10
    # a batch of nothing but empty queries, and an empty batch
    e_s, e_d = run_gpu(corpus_gpu, np.zeros(4, np.int64), np.zeros(0, np.int32), 10)
    assert e_d.tolist() == [list(range(10))] * 3 and not e_s.any()
    z_s, z_d = run_gpu(corpus_gpu, np.zeros(1, np.int64), np.zeros(0, np.int32), 10)
    assert z_s.shape == (0, 10) and z_d.shape == (0, 10)


def test_errors_match_bm25s(small_corpus, corpus_gpu):
    tune(corpus_gpu)
    qi = np.array([0, 1], np.int64)
    with pytest.raises(ValueError):                       # bm25s: k > num_docs -> ValueError
        run_gpu(corpus_gpu, qi, np.array([3], np.int32), 129)
    with pytest.raises(ValueError):                       # bm25s: token id out of range -> ValueError
        run_gpu(corpus_gpu, qi, np.array([small_corpus["vocab"]], np.int32), 10)
    with pytest.raises(ValueError):
        run_gpu(corpus_gpu, qi, np.array([-1], np.int32), 10)
    tiny = bo.build_index_loop([np.array([0, 1]), np.array([1])], 2)
    gt = gpu_index(tiny)
    with pytest.raises(ValueError):
        run_gpu(gt, qi, np.array([0], np.int32), 3)
    s, d = run_gpu(gt, qi, np.array([1], np.int32), 2)
    os_, od = bo.retrieve_batch(tiny, qi, np.array([1], np.int32), 2)
    assert_parity(s, d, os_, od)


def test_malformed_query_batches_raise_instead_of_reading_out_of_bounds(small_corpus, corpus_gpu):
    """Raw pointers cross the C ABI: the wrapper checks device / dtype / layout, the library checks the CSR's
    content on the device and then scores nothing (no illegal address, the context survives)."""
    tune(corpus_gpu)
    gi = corpus_gpu
    qi, qt = small_corpus["q_indptr"][:9], small_corpus["q_terms"][:int(small_corpus["q_indptr"][8])]
    d_qi, d_qt = to_dev(gi, qi, qt)
    good = gi.topk(d_qi, d_qt, 10)
    with pytest.raises(ValueError, match="device"):
        gi.topk(d_qi.cpu(), d_qt, 10)
    with pytest.raises(ValueError, match="device"):
        gi.topk(d_qi, d_qt.cpu(), 10)
    with pytest.raises(ValueError, match="int64"):
        gi.topk(d_qi.int(), d_qt, 10)
    with pytest.raises(ValueError, match="out must"):
        gi.topk(d_qi, d_qt, 10, out=(torch.empty((8, 10), device=gi.device), torch.empty((8, 9), dtype=torch.int32, device=gi.device)))
    with pytest.raises(ValueError, match="out must"):
        gi.topk(d_qi, d_qt, 10, out=(torch.empty((8, 10)), torch.empty((8, 10), dtype=torch.int32)))
    # a strided view is made contiguous, not mis-read
    wide = torch.stack([d_qt, d_qt], dim=1)
    s, d = gi.topk(d_qi, wide[:, 0], 10)
    assert torch.equal(s, good[0]) and torch.equal(d, good[1])
    # CSR content: q_terms shorter than q_indptr says, offsets not starting at 0, decreasing offsets
    with pytest.raises(ValueError, match="CSR"):
        gi.topk(d_qi, d_qt[:-3], 10)
    with pytest.raises(ValueError, match="CSR"):
        gi.topk(d_qi + 1, d_qt, 10)
    bad = d_qi.clone()
    bad[3], bad[4] = d_qi[4], d_qi[3] - 1
    with pytest.raises(ValueError, match="CSR"):
        gi.topk(bad, d_qt, 10)
    torch.cuda.synchronize()
    s, d = gi.topk(d_qi, d_qt, 10)                         # the context and the handle are intact
    assert torch.equal(s, good[0]) and torch.equal(d, good[1])


def test_topk_host_returns_arrays_the_caller_owns(small_corpus, corpus_gpu):
    """Two calls of the same shape (two retrieval rounds) must not alias each other's results."""
    tune(corpus_gpu)
    qi, qt = small_corpus["q_indptr"], small_corpus["q_terms"]
    a_s, a_d, h2d, d2h = corpus_gpu.topk_host(qi[:9], qt[:qi[8]], 10)
    keep_s, keep_d = a_s.copy(), a_d.copy()
    b_qi = qi[8:17] - qi[8]
    b_s, b_d, _, _ = corpus_gpu.topk_host(b_qi, qt[qi[8]:qi[16]], 10)
    assert np.array_equal(a_s, keep_s) and np.array_equal(a_d, keep_d)
    assert not np.array_equal(a_d, b_d)
    assert h2d == 9 * 8 + int(qi[8]) * 4 and d2h == 8 * 10 * 8
    os_, od = co.retrieve_batch(small_corpus["index"], qi[:17], qt, 10)
    assert_parity(np.concatenate([a_s, b_s]), np.concatenate([a_d, b_d]), os_, od)


def test_invalid_index_rejected():
    from probing_rag_b200 import BM25Index
    with pytest.raises(ValueError):
        BM25Index.from_arrays(np.array([1.0, 1.0], np.float32), np.array([1, 0], np.int32),
                              np.array([0, 2], np.int64), 2)          # doc ids not ascending
    with pytest.raises(ValueError):
        BM25Index.from_arrays(np.array([-1.0], np.float32), np.array([0], np.int32),
                              np.array([0, 1], np.int64), 2)          # negative weight
    with pytest.raises(ValueError):
        BM25Index.from_arrays(np.array([1.0], np.float32), np.array([5], np.int32),
                              np.array([0, 1], np.int64), 2)          # doc id out of range


def test_tie_heavy_corpus():
    """Many identical documents -> large exact-score tie groups cut at the k boundary.  The live thresholds are
    NON-strict bounds (a document tying with a bound found by another warp can still win on doc id): this is
    the corpus where a strict test would lose documents."""
    rng = np.random.default_rng(5)
    base = [rng.integers(0, 40, size=int(rng.integers(4, 12))).astype(np.int32) for _ in range(25)]
    docs = [base[int(rng.integers(0, 25))].copy() for _ in range(30000)]
    idx = bo.build_index(np.concatenate(docs), np.array([len(d) for d in docs]), 40)
    queries = [[int(t)] for t in range(0, 40, 5)] + [[1, 2, 3], [7, 7, 9, 30, 2], [39, 0]] + \
        [list(range(16)), list(range(20, 28)), [1] * 12, list(range(5, 22)), [3, 3, 4, 4, 5, 5, 6, 6, 7]]  # long step lists: chunked sub-tiles
    qi = np.zeros(len(queries) + 1, np.int64)
    qi[1:] = np.cumsum([len(q) for q in queries])
    qt = np.array([t for q in queries for t in q], dtype=np.int32)
    gi = gpu_index(idx)
    for k in (1, 10, 100):
        os_, od = bo.retrieve_batch(idx, qi, qt, k)
        for tun in (dict(), dict(subs_per_item=2, warps_per_cta=4, docs_per_launch=4096, min_items=1, items_per_warp=1),
                    dict(subs_per_item=12, warps_per_cta=8, docs_per_launch=98304, min_items=2048),
                    dict(subs_per_item=1, warps_per_cta=12, items_per_warp=64),
                    dict(subs_per_item=3, warps_per_cta=8, docs_per_launch=8192, min_items=2048),
                    dict(batch_variant=2), dict(batch_variant=2, tile_epochs=2, docs_per_launch=8192, min_items=1),
                    dict(batch_variant=2, subs_per_item=1, docs_per_launch=4096, min_items=1)):
            tune(gi, **tun)
            for _ in range(3):                      # the order in which warps publish bounds varies from run to run
                gs, gd = run_gpu(gi, qi, qt, k)
                assert_parity(gs, gd, os_, od)


def test_without_boundary_tables_every_term_takes_the_cursor_path(small_corpus):
    """aux budget 0 -> no term is tabulated: the head terms (thousands of postings per sub-tile,
    clustered far beyond the lane-local scan limit) go through the rare-term cursor + warp search."""
    gi = gpu_index(small_corpus["index"], aux_budget_bytes=0)
    info = gi.aux_info()
    assert info["table_rows"] == 0 and info["hot_rows"] == 0 and info["cold_stream_bytes"] == 8 * gi.nnz
    qi, qt = small_corpus["q_indptr"][:201], small_corpus["q_terms"]
    os_, od = co.retrieve_batch(small_corpus["index"], qi, qt, 10, n_threads=8)
    for tun in (dict(subs_per_item=12, warps_per_cta=8, docs_per_launch=98304),
                dict(subs_per_item=3, warps_per_cta=4, docs_per_launch=8192, min_items=1, items_per_warp=1)):
        tune(gi, **tun)
        gs, gd = run_gpu(gi, qi, qt, 10)
        assert_parity(gs, gd, os_, od)
    lq, lt = synth.queries_np(8, small_corpus["vocab"], small_corpus["index"]["df"], kind="later")
    os_, od = co.retrieve_batch(small_corpus["index"], lq, lt, 10, n_threads=8)
    gs, gd = run_gpu(gi, lq, lt, 10)
    assert_parity(gs, gd, os_, od)


def build_shards(small_corpus, g):
    idx, toks, lens = small_corpus["index"], small_corpus["tokens"], small_corpus["doc_lens"]
    off = np.concatenate([[0], np.cumsum(lens, dtype=np.int64)])
    n_docs = len(lens)
    per = -(-n_docs // g)
    shards = []
    for r in range(g):
        lo, hi = r * per, min((r + 1) * per, n_docs)
        sh = bo.build_index(toks[off[lo]:off[hi]], lens[lo:hi], small_corpus["vocab"], n_docs_global=n_docs,
                            avgdl_global=idx["avgdl"], df_global=idx["df"], doc_id_base=lo)
        shards.append(gpu_index(sh, n_docs_global=n_docs, doc_id_base=lo))
    return shards


def test_doc_range_shards_and_merge_equal_single_index(small_corpus):
    """SURVEY 8e on one GPU: G doc-range shards built with global statistics, local top-k each,
    pr_topk_merge -> bit-identical to the single index."""
    from probing_rag_b200 import merge_topk
    qi, qt = small_corpus["q_indptr"][:301], small_corpus["q_terms"]
    ref_s, ref_d = co.retrieve_batch(small_corpus["index"], qi, qt, 10, n_threads=8)
    for g in (2, 3, 8):
        ss, dd = [], []
        for gi in build_shards(small_corpus, g):
            tune(gi, **(TUNINGS[1] if g == 2 else TUNINGS[0] if g == 3 else dict(docs_per_launch=4096, min_items=1)))
            s, d = gi.topk(*to_dev(gi, qi, qt), 10)
            ss.append(s); dd.append(d)
        ms, md = merge_topk(torch.stack(ss), torch.stack(dd))
        torch.cuda.synchronize()
        assert_parity(ms.cpu().numpy(), md.cpu().numpy(), ref_s, ref_d)


@pytest.mark.parametrize("g", [2, 4])
def test_threshold_exchange_between_shards_keeps_the_merged_lists_bit_identical(small_corpus, g):
    """The multi-GPU protocol on one GPU: the shards run launch by launch in lock step and, between launches,
    every shard's per-query bound is raised to the MAXIMUM over the shards (what the all-reduce(MAX) of
    ShardedBM25 does over NCCL).  A shard may then hold fewer than k candidates of its own -- the merged lists
    must still equal the single index bit for bit, ties included."""
    from probing_rag_b200 import _lib, merge_topk
    qi, qt = small_corpus["q_indptr"][:301], small_corpus["q_terms"]
    ref_s, ref_d = co.retrieve_batch(small_corpus["index"], qi, qt, 10, n_threads=8)
    shards = build_shards(small_corpus, g)
    nq, k = len(qi) - 1, 10
    L = _lib.lib()
    for tun in (dict(subs_per_item=2, docs_per_launch=4096, min_items=1, items_per_warp=1), dict(docs_per_launch=16384, min_items=1, subs_per_item=4),
                dict(batch_variant=2, docs_per_launch=16384, min_items=1, subs_per_item=4)):
        outs, calls, thetas = [], [], []
        for gi in shards:
            tune(gi, **tun)
            d_qi, d_qt = to_dev(gi, qi, qt)
            out = (torch.empty((nq, k), dtype=torch.float32, device=gi.device), torch.empty((nq, k), dtype=torch.int32, device=gi.device))
            ws = gi._workspace(nq, k)
            off = int(L.pr_bm25_theta_offset(gi._handle, nq, k))
            thetas.append(ws[off:off + 4 * nq].view(torch.float32))
            calls.append((gi, (gi._handle, nq, d_qi.data_ptr(), d_qt.data_ptr(), d_qt.numel(), k, out[0].data_ptr(), out[1].data_ptr(),
                               ws.data_ptr(), ws.numel()), (d_qi, d_qt)))
            outs.append(out)
        n_launch = {gi.num_launches(nq, k) for gi in shards}
        assert len(n_launch) == 1 and min(n_launch) > 1      # equal shards -> the same plan on every "rank"
        stream = torch.cuda.current_stream().cuda_stream
        raised = 0
        for li in range(min(n_launch)):
            for gi, args, _ in calls:
                _lib.check(L.pr_bm25_topk_range(*args, li, li + 1, stream))
            best = torch.stack(thetas).max(dim=0).values
            raised += int(sum((best > t).sum() for t in thetas))
            for t in thetas:
                t.copy_(best)
        assert raised > 0                                    # the exchange really changed some shard's bound
        ms, md = merge_topk(torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs]))
        torch.cuda.synchronize()
        assert_parity(ms.cpu().numpy(), md.cpu().numpy(), ref_s, ref_d)


@pytest.mark.parametrize("g", [2, 4])
def test_union_bound_exchange_between_shards_keeps_the_merged_lists_bit_identical(small_corpus, g):
    """The stronger exchange of the first launches (pr_bm25_raise_union_bound), in lock step on one GPU: behind every
    launch the shards' running score lists are stacked (the all-gather) and every shard's bound is raised to the
    k-th largest score of the union.  The bound must be exactly that value where it beats the shard's own, and the
    merged lists must still equal the single index bit for bit."""
    from probing_rag_b200 import _lib, merge_topk
    qi, qt = small_corpus["q_indptr"][:301], small_corpus["q_terms"]
    nq = len(qi) - 1
    L = _lib.lib()
    for k in (10, 40):
        ref_s, ref_d = co.retrieve_batch(small_corpus["index"], qi, qt, k, n_threads=8)
        shards = build_shards(small_corpus, g)
        outs, calls, thetas, runs = [], [], [], []
        for gi in shards:
            tune(gi, subs_per_item=2, docs_per_launch=4096, min_items=1, items_per_warp=1)
            d_qi, d_qt = to_dev(gi, qi, qt)
            out = (torch.empty((nq, k), dtype=torch.float32, device=gi.device), torch.empty((nq, k), dtype=torch.int32, device=gi.device))
            ws = gi._workspace(nq, k)
            off = int(L.pr_bm25_theta_offset(gi._handle, nq, k))
            thetas.append(ws[off:off + 4 * nq].view(torch.float32))
            off = int(L.pr_bm25_running_scores_offset(gi._handle, nq, k))
            runs.append(ws[off:off + 4 * nq * k].view(torch.float32).view(nq, k))
            calls.append((gi, (gi._handle, nq, d_qi.data_ptr(), d_qt.data_ptr(), d_qt.numel(), k, out[0].data_ptr(), out[1].data_ptr(),
                               ws.data_ptr(), ws.numel()), (d_qi, d_qt), ws))
            outs.append(out)
        n_launch = {gi.num_launches(nq, k) for gi in shards}
        assert len(n_launch) == 1 and min(n_launch) > 2
        stream = torch.cuda.current_stream().cuda_stream
        raised = 0
        for li in range(min(n_launch)):
            for gi, args, _, _ in calls:
                _lib.check(L.pr_bm25_topk_range(*args, li, li + 1, stream))
            if li + 1 == min(n_launch):
                break
            gathered = torch.stack(runs).contiguous()                       # [G, B, k]
            union_kth = gathered.permute(1, 0, 2).reshape(nq, -1).topk(k, dim=1).values[:, k - 1]
            before = [t.clone() for t in thetas]
            for gi, args, _, ws in calls:
                _lib.check(L.pr_bm25_raise_union_bound(gi._handle, nq, k, gathered.data_ptr(), g, ws.data_ptr(), ws.numel(), stream))
            for t, b in zip(thetas, before):
                want = torch.where(union_kth > 0, torch.maximum(b, union_kth), b)
                assert torch.equal(t, want)
                raised += int((t > b).sum())
        assert raised > 0
        ms, md = merge_topk(torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs]))
        torch.cuda.synchronize()
        assert_parity(ms.cpu().numpy(), md.cpu().numpy(), ref_s, ref_d)


def test_gpu_index_builder_bit_exact(small_corpus):
    """BM25Index.from_tokens on the device == the oracle's builder (weights bit for bit)."""
    from probing_rag_b200 import BM25Index
    gi = BM25Index.from_tokens(torch.from_numpy(small_corpus["tokens"]).cuda(),
                               torch.from_numpy(small_corpus["doc_lens"]).cuda(), small_corpus["vocab"])
    ora = small_corpus["index"]
    assert np.array_equal(gi.indptr.cpu().numpy(), ora["indptr"])
    assert np.array_equal(gi.doc_ids.cpu().numpy(), ora["indices"])
    assert np.array_equal(gi.weights.cpu().numpy(), ora["data"])


def test_save_load_roundtrip(tmp_path, small_corpus, corpus_gpu):
    from probing_rag_b200 import BM25Index
    tune(corpus_gpu)
    corpus_gpu.save(str(tmp_path / "ix"))
    gi = BM25Index.load(str(tmp_path / "ix"))
    qi, qt = small_corpus["q_indptr"][:33], small_corpus["q_terms"]
    a = run_gpu(corpus_gpu, qi, qt, 10)
    b = run_gpu(gi, qi, qt, 10)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_shards_cut_from_a_saved_whole_index_equal_shards_built_from_the_corpus(tmp_path, small_corpus, corpus_gpu):
    """BM25Index.load(doc_range=...): the doc-range shard of a persisted whole-corpus index has exactly the arrays of a
    shard built from its documents with the global statistics, including empty and one-document ranges."""
    from probing_rag_b200 import BM25Index, merge_topk
    corpus_gpu.save(str(tmp_path / "ix"))
    idx, toks, lens = small_corpus["index"], small_corpus["tokens"], small_corpus["doc_lens"]
    off = np.concatenate([[0], np.cumsum(lens, dtype=np.int64)])
    n_docs = len(lens)
    qi, qt = small_corpus["q_indptr"][:129], small_corpus["q_terms"]
    ref_s, ref_d = co.retrieve_batch(idx, qi, qt, 10, n_threads=8)
    ss, dd = [], []
    for lo, hi in ((0, 33334), (33334, 33335), (33335, 33335), (33335, 90000), (90000, n_docs)):
        sh = BM25Index.load(str(tmp_path / "ix"), doc_range=(lo, hi))
        want = bo.build_index(toks[off[lo]:off[hi]], lens[lo:hi], small_corpus["vocab"], n_docs_global=n_docs,
                              avgdl_global=idx["avgdl"], df_global=idx["df"], doc_id_base=lo)
        assert (sh.n_docs, sh.n_docs_global, sh.doc_id_base) == (hi - lo, n_docs, lo)
        assert np.array_equal(sh.indptr.cpu().numpy(), want["indptr"])
        assert np.array_equal(sh.doc_ids.cpu().numpy(), want["indices"])
        assert np.array_equal(sh.weights.cpu().numpy(), want["data"])
        s, d = sh.topk(*to_dev(sh, qi, qt), 10)
        ss.append(s); dd.append(d)
    ms, md = merge_topk(torch.stack(ss), torch.stack(dd))
    torch.cuda.synchronize()
    assert_parity(ms.cpu().numpy(), md.cpu().numpy(), ref_s, ref_d)
    with pytest.raises(ValueError):
        BM25Index.load(str(tmp_path / "ix"), doc_range=(5, n_docs + 1))


def test_retriever_text_level_drop_in():
    """exp_rag.py:242 / :426 / :372 shape: from_defaults(docstore=...), retrieve(str) -> nodes with
    .text/.score, score-descending; checked against the oracle on the same token ids."""
    from probing_rag_b200 import BM25Retriever, Document, SimpleDocumentStore
    rng = np.random.default_rng(3)
    words = ("retrieval augmented generation probing language model hidden state wikipedia passage "
             "question answer paris france capital city river tower london england bridge").split()
    texts = [" ".join(rng.choice(words, size=int(rng.integers(5, 30)))) for _ in range(500)]
    store = SimpleDocumentStore()
    store.add_documents([Document(text=t, doc_id=str(i)) for i, t in enumerate(texts)])   # make_indexer.py:438-442
    bm25 = BM25Retriever.from_defaults(docstore=store, similarity_top_k=5)
    res = bm25.retrieve("What is the capital city of France? The tower of Paris")
    assert len(res) == 5 and all(res[i].score >= res[i + 1].score for i in range(4))
    toks, lens = bm25.vocab.encode_corpus(texts)
    ora = bo.build_index(toks, lens, len(bm25.vocab))
    q = np.array(bm25.vocab.encode_query("What is the capital city of France? The tower of Paris"), np.int32)
    os_, od = bo.retrieve(ora, q, 5)
    assert [int(r.node.id_) for r in res] == od.tolist()
    assert [r.score for r in res] == [float(x) for x in os_]
    assert res[0].text == texts[od[0]] and res[0].get_content() == res[0].node.text
    batch = bm25.retrieve_batch(["paris tower", "", "london bridge river"], k=3)
    assert len(batch) == 3 and [len(b) for b in batch] == [3, 3, 3]
    with pytest.raises(ValueError):
        BM25Retriever.from_defaults(similarity_top_k=5)


@pytest.mark.skipif(os.environ.get("PR_SKIP_FULL") == "1", reason="PR_SKIP_FULL=1")
def test_full_size_21m_properties():
    """BASELINE config 2 shape (21,015,324 passages): the oracle cannot score 64k queries here,
    so check a 24-query sample against the C oracle on the same arrays, plus size-independent
    properties on a larger batch: rank order, launch-plan invariance, batch-split invariance."""
    import bench
    gi, qi, qt = bench.build_workload(synth.N_DOCS_WIKI, 1 << 22, 2048, torch.device("cuda"))
    d_qi, d_qt = to_dev(gi, qi, qt)
    tune(gi)
    s2, d2 = gi.topk(d_qi, d_qt, 10)
    for tun in (dict(docs_per_launch=49152, min_items=1), dict(subs_per_item=6, warps_per_cta=12),
                dict(docs_per_launch=4000000, warps_per_cta=4), dict(batch_variant=2), dict(batch_variant=2, tile_epochs=2)):
        tune(gi, **tun)
        s1, d1 = gi.topk(d_qi, d_qt, 10)
        assert torch.equal(s1, s2) and torch.equal(d1, d2), tun
    tune(gi)
    assert bool((s2[:, :-1] >= s2[:, 1:]).all())
    tie = s2[:, :-1] == s2[:, 1:]
    assert bool((d2[:, :-1][tie] < d2[:, 1:][tie]).all())
    for b in (1, 8, 64):                                       # small batches: one launch, thousands of lists per query
        sa, da = gi.topk(d_qi[:b + 1], d_qt[:int(qi[b])], 10)
        assert gi.num_launches(b, 10) == 1
        assert torch.equal(sa, s2[:b]) and torch.equal(da, d2[:b])
    host = {"data": gi.weights.cpu().numpy(), "indices": gi.doc_ids.cpu().numpy(),
            "indptr": gi.indptr.cpu().numpy(), "num_docs": gi.n_docs}
    os_, od = co.retrieve_batch(host, qi[:25], qt, 10, n_threads=min(24, os.cpu_count() or 1))
    assert_parity(s2[:24].cpu().numpy(), d2[:24].cpu().numpy(), os_, od)


def test_extreme_weights(small_corpus):
    """Weights far outside the usual BM25 range (2^-40 .. 2^12 times a normal weight): the sign-epoch
    accumulators must still match the oracle bit for bit."""
    idx = dict(small_corpus["index"])
    data = idx["data"].copy()
    data[::7] *= np.float32(2.0 ** -40)          # tiny weights
    data[3::11] *= np.float32(4096.0)            # large weights
    idx["data"] = data
    gi = gpu_index(idx)
    qi, qt = small_corpus["q_indptr"][:129], small_corpus["q_terms"]
    os_, od = co.retrieve_batch(idx, qi, qt, 10, n_threads=8)
    assert gi.get_tuning()["tile_epochs"] == 4          # asked for, but the weight range forbids it: the library runs two
    for tun in (dict(), dict(docs_per_launch=16384, min_items=1), dict(batch_variant=2), dict(batch_variant=2, docs_per_launch=16384, min_items=1)):
        tune(gi, **tun)
        gs, gd = run_gpu(gi, qi, qt, 10)
        assert_parity(gs, gd, os_, od)


@pytest.mark.parametrize("seed", range(12))
def test_random_shapes(seed):
    """Ragged shapes the Zipf workload never produces: 1..9000 documents (one sub-tile, a partial last sub-tile,
    exactly 2048 / 4096), tiny vocabularies (every term is heavy: thousands of postings per sub-tile, many
    wide steps, lists longer than one descriptor list), empty and 1..48-term queries with duplicates."""
    rng = np.random.default_rng(900 + seed)
    n_docs = int([1, 31, 2047, 2048, 2049, 4096, 5000, 9000, 700, 4097, 6143, 8192][seed])
    vocab = int(rng.choice([4, 16, 200, 3000]))
    lens = rng.integers(1, 40, size=n_docs).astype(np.int32)
    toks = rng.integers(0, vocab, size=int(lens.sum())).astype(np.int32)
    idx = bo.build_index(toks, lens, vocab)
    nq = 97
    qlens = rng.integers(0, 49, size=nq)
    qlens[:3] = [0, 1, 48]
    qi = np.concatenate([[0], np.cumsum(qlens)]).astype(np.int64)
    qt = rng.integers(0, vocab, size=int(qlens.sum())).astype(np.int32)
    keep = idx["df"][qt] > 0                     # bm25s drops unknown terms on the query side
    qid = np.repeat(np.arange(nq), qlens)[keep]
    qt = qt[keep]
    qi = np.concatenate([[0], np.cumsum(np.bincount(qid, minlength=nq))]).astype(np.int64)
    gi = gpu_index(idx)
    for k in (1, min(10, n_docs), min(100, n_docs)):
        os_, od = co.retrieve_batch(idx, qi, qt, k, n_threads=8)
        for tun in (dict(), dict(subs_per_item=1, warps_per_cta=4, docs_per_launch=2048, min_items=1, items_per_warp=1),
                    dict(subs_per_item=2, warps_per_cta=12, docs_per_launch=4096, min_items=1, items_per_warp=1),
                    dict(batch_variant=2), dict(batch_variant=2, subs_per_item=1, docs_per_launch=2048, min_items=1),
                    dict(batch_variant=2, tile_epochs=2, subs_per_item=3, docs_per_launch=6144, min_items=1)):
            tune(gi, **tun)
            gs, gd = run_gpu(gi, qi, qt, k)
            assert_parity(gs, gd, os_, od)
