"""Fused prober (tcgen05 bf16x3 GEMMs + gate + compaction) against the reference's ImprovedProbe.
Tolerance (BASELINE.json north_star): prober probabilities within 1e-3 of the fp32 reference."""
import os

import numpy as np
import pytest
import torch

from oracle import bm25_oracle as bo
from oracle import c_oracle as co
from oracle import prober_oracle as po
from probing_rag_b200 import synth

pytestmark = pytest.mark.gpu

PROB_ATOL = 1e-3


def oracle_probers(layers=po.PROBE_LAYERS, d_model=po.D_MODEL):
    out = []
    for layer in layers:
        p = po.OracleImprovedProbe(d_model, po.N_CLASSES)
        p.load_state_dict(po.make_prober_state(layer, d_model=d_model))
        out.append(p.eval())
    return out


@pytest.fixture(scope="module")
def gate6():
    from probing_rag_b200.prober import ProberGate
    probers = oracle_probers()
    return probers, ProberGate([p.state_dict() for p in probers], device="cuda")


def check_gate(out, ref_logits, theta, ablation):
    psum_ref, ret_ref = po.gate(ref_logits, theta, ablation)
    assert torch.allclose(out.probsum.cpu(), psum_ref, atol=PROB_ATOL * (ref_logits.shape[1] - ablation) + 1e-6)
    margin = (psum_ref[:, 0] + theta - psum_ref[:, 1]).abs()
    agree = (out.retrieve.cpu() == ret_ref) | (margin < 6 * PROB_ATOL)
    assert bool(agree.all())
    assert torch.equal(out.retrieve_idx.cpu().long(), torch.nonzero(out.retrieve.cpu()).flatten())


def test_golden_reference_outputs(golden_dir, gate6):
    """tests/golden/prober_golden.npz holds the outputs of the REFERENCE class (utils.py:29-57)."""
    _, gate = gate6
    g = np.load(os.path.join(golden_dir, "prober_golden.npz"))
    x = po.make_hidden_states(int(g["n"]), seed=0)
    ref = torch.from_numpy(g["logits"])
    for j, th in enumerate(g["thetas"]):
        out = gate(x.cuda(), theta=float(th), want_logits=True)
        dp = (torch.softmax(out.logits.cpu(), -1) - torch.softmax(ref, -1)).abs().max().item()
        assert dp < PROB_ATOL, dp
        assert dp < 2e-4, f"bf16x3 should be far inside the tolerance, got {dp}"
        margin = np.abs(g["probsum"][:, 0] + float(th) - g["probsum"][:, 1])
        assert np.all((out.retrieve.cpu().numpy() == g["retrieve"][j]) | (margin < 6 * PROB_ATOL))


@pytest.mark.parametrize("n", [1, 7, 128, 129, 1000])
def test_batch_sizes_and_ablation(gate6, n):
    probers, gate = gate6
    x = po.make_hidden_states(n, seed=n)
    ref = po.prober_logits(probers, x)
    for theta, ablation in ((0.0, 0), (0.5, 2), (-0.25, 5)):
        out = gate(x.cuda(), theta=theta, ablation=ablation, want_logits=True)
        assert (torch.softmax(out.logits.cpu(), -1) - torch.softmax(ref, -1)).abs().max().item() < PROB_ATOL
        check_gate(out, ref, theta, ablation)


def test_large_mean_and_scale_inputs(gate6):
    """Sums of up to 149 residual vectors: large magnitude, non-zero mean (hard parts)."""
    probers, gate = gate6
    x = po.make_hidden_states(256, seed=11)
    x = x * 3.0 + 40.0 * x.std(dim=-1, keepdim=True)
    ref = po.prober_logits(probers, x)
    out = gate(x.cuda(), want_logits=True)
    assert (torch.softmax(out.logits.cpu(), -1) - torch.softmax(ref, -1)).abs().max().item() < PROB_ATOL


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_half_precision_hidden_states_are_read_as_they_are(gate6, dtype):
    """exp_rag.py:385-387 feeds the prober whatever dtype the LM produced: bf16 / f16 sums go through the C ABI
    unconverted (x_dtype) and must match the reference evaluated on the same (widened) values."""
    probers, gate = gate6
    x = po.make_hidden_states(300, seed=21)
    if dtype == torch.float16:
        x = x * (1.0 / 16.0)                       # keep the largest sums inside f16 range
    xh = x.to(dtype)
    ref = po.prober_logits(probers, xh.float())
    out = gate(xh.cuda(), want_logits=True)
    assert (torch.softmax(out.logits.cpu(), -1) - torch.softmax(ref, -1)).abs().max().item() < PROB_ATOL
    check_gate(out, ref, 0.0, 0)
    wide = gate(xh.float().cuda(), want_logits=True)           # the same values through the f32 path: same kernel maths
    assert torch.equal(wide.logits, out.logits) and torch.equal(wide.retrieve, out.retrieve)


def test_gate_compares_in_double_like_the_reference(gate6):
    """exp_rag.py:414 `P0.item() + threshold < P1.item()` is Python-float arithmetic.  With the gate's own f32
    sums as inputs the decision must match that expression EXACTLY for thresholds that are not f32 numbers
    (0.1) and for thresholds placed on a row's boundary, where an f32 comparison flips."""
    _, gate = gate6
    x = po.make_hidden_states(2000, seed=33).cuda()
    base = gate(x)
    p = base.probsum.cpu().double()
    for theta in (0.1, -0.1, 0.3, 1e-9, -1e-9):
        out = gate(x, theta=theta)
        assert torch.equal(out.probsum, base.probsum)
        want = ~(p[:, 0] + theta < p[:, 1])
        assert torch.equal(out.retrieve.cpu(), want), theta
    # thresholds exactly on / one double-ulp around the boundary of individual rows
    for r in (3, 500, 1999):
        edge = float(p[r, 1] - p[r, 0])                      # P0 + edge == P1 in double
        for theta in (edge, np.nextafter(edge, -np.inf), np.nextafter(edge, np.inf)):
            out = gate(x, theta=float(theta))
            want = ~(p[:, 0] + float(theta) < p[:, 1])
            assert torch.equal(out.retrieve.cpu(), want)


def test_async_gate_output_is_usable_without_a_host_sync(gate6):
    _, gate = gate6
    x = po.make_hidden_states(777, seed=5).cuda()
    a = gate(x, sync=True)
    b = gate(x, sync=False)
    n = int(b.n_retrieve.item())
    assert n == a.retrieve_idx.numel() == int(a.retrieve.sum())
    assert b.retrieve_idx.shape == (777,)
    assert torch.equal(b.retrieve_idx[:n], a.retrieve_idx) and bool((b.retrieve_idx[n:] == -1).all())


def test_other_shapes():
    from probing_rag_b200.prober import ProberGate
    for d_model, layers in ((256, (6,)), (1024, (6, 8, 10))):
        probers = oracle_probers(layers, d_model)
        gate = ProberGate([p.state_dict() for p in probers], device="cuda")
        x = po.make_hidden_states(300, seed=d_model, n_probers=len(layers), d_model=d_model)
        ref = po.prober_logits(probers, x)
        out = gate(x.cuda(), want_logits=True)
        assert (torch.softmax(out.logits.cpu(), -1) - torch.softmax(ref, -1)).abs().max().item() < PROB_ATOL
        check_gate(out, ref, 0.0, 0)


def test_improved_probe_is_a_drop_in_module(tmp_path):
    """utils.py:302-330: ImprovedProbe(input_size=d_model, output_size=2).to(device);
    load_state_dict(torch.load(path)); eval(); prober(x[B,2048]) -> logits[B,2]; .to('cpu')."""
    from probing_rag_b200.prober import STATE_KEYS, ImprovedProbe
    sd = po.make_prober_state(12)
    path = tmp_path / "in3_1.0_gemma-2b_tokens_mean_2_l12_resid_post_ep1.pt"     # utils.py:316 naming
    torch.save(sd, path)
    prober = ImprovedProbe(input_size=2048, output_size=2).to("cuda")
    assert set(prober.state_dict().keys()) == set(STATE_KEYS)
    assert sum(p.numel() for p in prober.parameters()) == 1318914
    prober.load_state_dict(torch.load(path))
    prober.eval()
    ref = po.OracleImprovedProbe(2048, 2)
    ref.load_state_dict(sd)
    ref.eval()
    x = po.make_hidden_states(5, seed=2)[:, 0]
    with torch.no_grad():
        logit = prober(x.cuda())
        assert logit.shape == (5, 2) and logit.is_cuda
        want = ref(x)
        assert (torch.softmax(logit.to("cpu"), -1) - torch.softmax(want, -1)).abs().max().item() < PROB_ATOL
        one = prober(x[:1].cuda()).to("cpu")                       # the reference's B=1 call (exp_rag.py:387)
        assert (torch.softmax(one, -1) - torch.softmax(want[:1], -1)).abs().max().item() < PROB_ATOL
    prober.train()
    assert prober(x.cuda()).shape == (5, 2)                        # training mode: plain module


def test_config4_prober_gated_bm25(small_corpus):
    """BASELINE config 4 shape: hidden states -> prober -> compacted BM25 top-10."""
    from probing_rag_b200 import BM25Index, BM25Retriever
    from probing_rag_b200.prober import ProberGate, gate_and_retrieve
    idx = small_corpus["index"]
    gi = BM25Index.from_arrays(idx["data"], idx["indices"], idx["indptr"], idx["num_docs"])
    retr = BM25Retriever.from_defaults(index=gi, similarity_top_k=10)
    nq = 512
    qi, qt = small_corpus["q_indptr"][:nq + 1], small_corpus["q_terms"][:small_corpus["q_indptr"][nq]]
    probers = oracle_probers()
    gate = ProberGate([p.state_dict() for p in probers], device="cuda")
    x = po.make_hidden_states(nq, seed=5)
    out, scores, ids = gate_and_retrieve(gate, retr, x.cuda(), torch.from_numpy(qi).cuda(), torch.from_numpy(qt).cuda())
    torch.cuda.synchronize()
    ref_logits = po.prober_logits(probers, x)
    check_gate(out, ref_logits, 0.0, 0)
    sel = out.retrieve_idx.cpu().numpy()
    assert 0 < len(sel) < nq
    os_, od = co.retrieve_batch(idx, qi, qt, 10, n_threads=8)
    assert np.array_equal(ids.cpu().numpy(), od[sel]) and np.array_equal(scores.cpu().numpy(), os_[sel])


# ---- on-device hidden-state pooling vs the reference's hook + concat + sum ------------
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_pooler_matches_reference_cache_concat_sum(dtype):
    """exp_rag.py:317-321 + 385-386: per layer, cache every forward call's activations on the host, drop
    the prefill entry, concat over tokens, sum.  The pooler adds on the device instead."""
    from functools import partial
    from probing_rag_b200.pooling import HiddenStatePooler
    g = torch.Generator().manual_seed(11)
    layers = [f"blocks.{l}.hook_resid_post" for l in po.PROBE_LAYERS]
    B, d = 5, po.D_MODEL
    pooler = HiddenStatePooler(B, layers, d)
    for gen in range(2):                         # two generations: reset() in between like `cache = {}`
        pooler.reset()
        cache = {}
        steps = [7] + [1] * (13 + gen)           # prefill of 7 tokens, then one token per decode step
        for T in steps:
            for name in layers:
                act = (torch.randn(B, T, d, generator=g) * 3).to(dtype).cuda()
                out = partial(pooler.hook_fn, layer=name)(act, None)
                assert out is act
                cache.setdefault(name, []).append(act.detach().cpu())
        ref = torch.stack([po.pool_hidden_states([e.float() for e in cache[name]]) for name in layers], dim=1)
        torch.cuda.synchronize()
        got = pooler.X.cpu()
        assert got.shape == (B, len(layers), d)
        # same addends, fp32 accumulation; only the summation order over <= 14 tokens may differ
        assert torch.allclose(got, ref, rtol=1e-5, atol=1e-4), (got - ref).abs().max()
        assert pooler.calls(layers[0]) == len(steps)


@pytest.mark.gpu
def test_pooler_row_map_strides_and_errors():
    from probing_rag_b200.pooling import HiddenStatePooler
    g = torch.Generator().manual_seed(5)
    pooler = HiddenStatePooler(6, ["a", "b"], 256)
    big = torch.randn(4, 9, 512, generator=g).cuda()
    view = big[:, 2:7, 128:384]                  # strided rows / tokens, offset features
    rm = torch.tensor([5, -1, 0, 2], dtype=torch.int32)
    pooler.add("b", view, row_map=rm)
    pooler.add("b", view, row_map=rm)
    torch.cuda.synchronize()
    want = torch.zeros(6, 2, 256)
    s = view.cpu().sum(dim=1)
    for r, dst in enumerate(rm.tolist()):
        if dst >= 0:
            want[dst, 1] = s[r] + s[r]
    assert torch.allclose(pooler.X.cpu(), want, rtol=1e-5, atol=1e-5)
    with pytest.raises(ValueError):
        pooler.add("a", torch.zeros(7, 1, 256, device="cuda"))      # more rows than the accumulator
    with pytest.raises(ValueError):
        pooler.add("a", torch.zeros(2, 1, 128, device="cuda"))      # wrong d_model
    with pytest.raises(KeyError):
        pooler.add("c", torch.zeros(2, 1, 256, device="cuda"))


@pytest.mark.gpu
def test_pooled_states_feed_the_gate_like_the_reference_loop():
    """hooks -> pooled X -> six probers -> gate, against the oracle chain on the host (exp_rag.py:381-415)."""
    from probing_rag_b200.pooling import HiddenStatePooler
    from probing_rag_b200.prober import ProberGate
    g = torch.Generator().manual_seed(2)
    layers = list(po.PROBE_LAYERS)
    probers = []
    for layer in layers:
        p = po.OracleImprovedProbe(po.D_MODEL, po.N_CLASSES)
        p.load_state_dict(po.make_prober_state(layer))
        probers.append(p.eval())
    B = 64
    pooler = HiddenStatePooler(B, layers, po.D_MODEL)
    pooler.reset()
    cache = {l: [] for l in layers}
    for T in [5] + [1] * 20:
        for l in layers:
            act = torch.randn(B, T, po.D_MODEL, generator=g) * 4
            pooler.hook_fn(act.cuda(), None, layer=l)
            cache[l].append(act)
    x_ref = torch.stack([po.pool_hidden_states(cache[l]) for l in layers], dim=1)
    ref_logits = po.prober_logits(probers, x_ref)
    psum_ref, ret_ref = po.gate(ref_logits, 0.0, 0)
    out = ProberGate([p.state_dict() for p in probers], device="cuda")(pooler.X, want_logits=True)
    err = (torch.softmax(out.logits.cpu(), -1) - torch.softmax(ref_logits, -1)).abs().max().item()
    assert err < 1e-3, err                        # north-star tolerance for prober probabilities
    margin = (psum_ref[:, 0] - psum_ref[:, 1]).abs()
    assert bool(((out.retrieve.cpu() == ret_ref) | (margin < 2e-3)).all())


# ---- BASELINE config 5: multi-step adaptive retrieval (batch, depth k, up to 4 retrieve calls) -------------
@pytest.mark.gpu
@pytest.mark.parametrize("nq,k", [(1, 1), (8, 5), (300, 10), (64, 100)])
def test_config5_adaptive_rounds_match_the_per_question_loop(small_corpus, nq, k):
    """exp_rag.py:396-468 for a batch on the GPU (fused gate -> compaction -> batched BM25 top-k, per round)
    against a literal per-question walk on the host: oracle probers + gate, C-oracle BM25, the same seeded
    'LM' (hidden states and LM-transcript-shaped next queries drawn per (question, call))."""
    from probing_rag_b200 import BM25Index, BM25Retriever, rounds
    from probing_rag_b200.prober import ProberGate
    idx = small_corpus["index"]
    vocab, df = small_corpus["vocab"], idx["df"]
    gi = BM25Index.from_arrays(idx["data"], idx["indices"], idx["indptr"], idx["num_docs"])
    retr = BM25Retriever.from_defaults(index=gi, similarity_top_k=k)
    probers = oracle_probers()
    gate = ProberGate([p.state_dict() for p in probers], device="cuda")
    qi, qt = small_corpus["q_indptr"][:nq + 1], small_corpus["q_terms"][:small_corpus["q_indptr"][nq]]

    def lm(q, call):
        """what the LM side hands back for question q after retrieve call `call` (call 0 = first generation)"""
        x = po.make_hidden_states(1, seed=1000 * call + q)[0]
        li, lt = synth.queries_np(1, vocab, df, seed=7000 * call + q, kind="later")
        return x, lt[li[0]:li[1]]

    # ---- host walk, one question at a time
    x0 = torch.stack([lm(q, 0)[0] for q in range(nq)])
    dec0 = po.gate(po.prober_logits(probers, x0), 0.0, 0)
    want_calls, want_last = [], {}
    # a question whose oracle gate margin |P0 - P1| ever falls inside the prober tolerance may legitimately take
    # another path on the GPU: it still runs, but is not compared
    fragile = ((dec0[0][:, 0] - dec0[0][:, 1]).abs() < 5e-3).tolist()
    for q in range(nq):
        need, calls, retr_count = bool(dec0[1][q]), 0, 0
        terms = qt[qi[q]:qi[q + 1]]
        while need:
            calls += 1
            os_, od = co.retrieve_batch(idx, np.array([0, len(terms)], dtype=np.int64), terms.astype(np.int32), k, n_threads=1)
            want_last[q] = (os_[0], od[0])
            x, terms = lm(q, calls)
            ps, again = po.gate(po.prober_logits(probers, x.unsqueeze(0)), 0.0, 0)
            fragile[q] = fragile[q] or float((ps[0, 0] - ps[0, 1]).abs()) < 5e-3
            if retr_count > 2:
                break
            retr_count += 1
            need = bool(again[0])
        want_calls.append(calls)
    assert sum(fragile) <= max(1, nq // 50)

    # ---- the batched path
    def step_fn(active, s, d, call):
        xs, terms = zip(*[lm(int(q), call) for q in active.tolist()]) if active.numel() else ((), ())
        X = torch.stack(xs).cuda() if xs else torch.zeros(0, 6, po.D_MODEL, device="cuda")
        lens = np.array([len(t) for t in terms], dtype=np.int64)
        n_indptr = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).cuda()
        n_terms = torch.from_numpy(np.concatenate(terms).astype(np.int32) if terms else np.zeros(0, np.int32)).cuda()
        return X, (n_indptr, n_terms)

    res = rounds.adaptive_retrieval(gate, retr.retrieve_ids, x0.cuda(), torch.from_numpy(qi).cuda(),
                                    torch.from_numpy(qt).cuda(), step_fn, k=k)
    got_calls, got_rc = res.calls.cpu().tolist(), res.retr_count.cpu().tolist()
    ls, ld = res.last_scores.cpu().numpy(), res.last_doc_ids.cpu().numpy()
    for q in range(nq):
        if fragile[q]:
            continue
        assert got_calls[q] == want_calls[q] and got_rc[q] == min(want_calls[q], 3), q
        if want_calls[q] == 0:
            assert ld[q, 0] == -1
        else:
            assert np.array_equal(ld[q], want_last[q][1]) and np.array_equal(ls[q], want_last[q][0]), q
