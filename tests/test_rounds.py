"""Round control (exp_rag.py:422-468) batched: host logic against the oracle's per-question
state machine, with fakes standing in for the CUDA gate / retriever (no GPU needed)."""
import numpy as np
import pytest
import torch

from oracle import prober_oracle as po
from probing_rag_b200 import rounds


class FakeGateOut:
    def __init__(self, retrieve):
        self.retrieve = retrieve


def make_fakes(decisions):
    """decisions[q] = list of gate outcomes after generation 0, 1, 2, ... of question q.
    X carries (question id, generation index) so the fake gate can look the outcome up."""
    def gate(X, theta=0.0, ablation=0):
        q = X[:, 0, 0].long().tolist()
        g = X[:, 0, 1].long().tolist()
        return FakeGateOut(torch.tensor([bool(decisions[a][b]) for a, b in zip(q, g)], dtype=torch.bool))

    log = []

    def retrieve_ids(q_indptr, q_terms, k):
        n = q_indptr.numel() - 1
        log.append((q_indptr.clone(), q_terms.clone()))
        first = torch.tensor([int(q_terms[q_indptr[i]]) if q_indptr[i + 1] > q_indptr[i] else -1 for i in range(n)])
        s = first.float().unsqueeze(1).repeat(1, k)
        d = first.to(torch.int32).unsqueeze(1).repeat(1, k)
        return s, d

    def step_fn(active, s, d, call):
        X = torch.zeros(active.numel(), 1, 2)
        X[:, 0, 0] = active.float()
        X[:, 0, 1] = call
        # next search input: two tokens [1000*call + q, 7]
        n = active.numel()
        qi = torch.arange(0, 2 * n + 1, 2, dtype=torch.int64)
        qt = torch.stack([1000 * call + active.to(torch.int32), torch.full((n,), 7, dtype=torch.int32)], 1).flatten()
        return X, (qi, qt)
    return gate, retrieve_ids, step_fn, log


def test_round_control_matches_reference_state_machine():
    rng = np.random.default_rng(0)
    nq = 200
    decisions = [list(rng.random(6) < 0.7) for _ in range(nq)]
    decisions[0] = [False] * 6                      # never retrieves
    decisions[1] = [True] * 6                       # retrieves until the cap
    decisions[2] = [True, False, True, True, True, True]
    gate, retrieve_ids, step_fn, log = make_fakes(decisions)
    X0 = torch.zeros(nq, 1, 2)
    X0[:, 0, 0] = torch.arange(nq).float()
    q_indptr = torch.arange(nq + 1, dtype=torch.int64)
    q_terms = torch.arange(nq, dtype=torch.int32)    # question q = single token q
    res = rounds.adaptive_retrieval(gate, retrieve_ids, X0, q_indptr, q_terms, step_fn, k=3)
    want_calls = [po.retrieval_rounds(d) for d in decisions]
    assert res.calls.tolist() == want_calls
    assert res.retr_count.tolist() == [min(c, 3) for c in want_calls]
    assert max(want_calls) == 4 and res.calls[1].item() == 4 and res.retr_count[1].item() == 3
    assert res.calls[0].item() == 0 and res.last_doc_ids[0, 0].item() == -1
    # first call searches the question, later calls the transcript of the previous generation
    for q in range(nq):
        c = want_calls[q]
        if c == 1:
            assert res.last_doc_ids[q, 0].item() == q
        elif c > 1:
            assert res.last_doc_ids[q, 0].item() == 1000 * (c - 1) + q
    assert len(log) == max(want_calls)
    assert [a.numel() for a in res.per_call_active] == [sum(1 for c in want_calls if c >= i) for i in range(1, 5)]


def test_select_queries_ragged_and_empty():
    qi = torch.tensor([0, 2, 2, 5, 6], dtype=torch.int64)
    qt = torch.tensor([10, 11, 20, 21, 22, 30], dtype=torch.int32)
    ci, ct = rounds.select_queries(qi, qt, torch.tensor([3, 1, 0]))
    assert ci.tolist() == [0, 1, 1, 3] and ct.tolist() == [30, 10, 11]
    ci, ct = rounds.select_queries(qi, qt, torch.tensor([], dtype=torch.int64))
    assert ci.tolist() == [0] and ct.numel() == 0
