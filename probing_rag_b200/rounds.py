"""Batched retrieval-round control: /root/reference/exp_rag.py:396-468 for a whole batch.

The reference walks one question at a time:

    generate -> gate (:406-415) -> while gate says retrieve:                      (:422)
        bm25.retrieve(question if first call else decoded transcript)             (:426 / :428)
        build prompt, generate (:439-443), gate again (:446-455)
        if retr_count > 2: break  else: retr_count += 1                           (:462-465)

so a question issues at most 4 `retrieve` calls and its recorded `retr_count` saturates at 3.
Here the same state machine runs over a batch: every round gates all still-active questions
with one fused prober call, compacts the ones that retrieve, scores them with one batched
BM25 top-k, and hands the passages to `step_fn` -- the LM side (prompt + generate + pooled
hidden states), which stays stock PyTorch and is not part of this package.

    step_fn(active: LongTensor[n], scores f32[n,k], doc_ids i32[n,k], call: int)
        -> (X_next f32[n, P, d_model], (q_indptr i64[n+1], q_terms i32[...]))

returns, for the n questions that just retrieved (ascending question index), the pooled hidden
states of the new generation and the next search input (the decoded transcript, :457) as term ids.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable

import torch

MAX_RETR_COUNT = 3      # exp_rag.py:462-465: `if retr_count > 2: break`
MAX_CALLS = 4           # ... which allows a 4th retrieve call before the break


@dataclass
class RoundsResult:
    retr_count: torch.Tensor                 # i32[B]  what the reference appends to retr_count_list (:468)
    calls: torch.Tensor                      # i32[B]  retrieve calls actually issued (0..4)
    last_scores: torch.Tensor | None         # f32[B,k] lists of each question's last retrieve call (rows of
    last_doc_ids: torch.Tensor | None        # i32[B,k] questions that never retrieved are -inf / -1)
    per_call_active: list = field(default_factory=list)   # questions scored by each call


def select_queries(q_indptr: torch.Tensor, q_terms: torch.Tensor, idx: torch.Tensor):
    """CSR sub-batch of the queries `idx` (ascending or not), on the arrays' device."""
    idx = idx.long()
    lens = (q_indptr[1:] - q_indptr[:-1])[idx]
    c_indptr = torch.zeros(idx.numel() + 1, dtype=torch.int64, device=q_indptr.device)
    torch.cumsum(lens, 0, out=c_indptr[1:])
    total = int(c_indptr[-1].item()) if idx.numel() else 0
    pos = (torch.arange(total, device=q_indptr.device) - torch.repeat_interleave(c_indptr[:-1], lens)
           + torch.repeat_interleave(q_indptr[:-1][idx], lens))
    return c_indptr, q_terms[pos].to(torch.int32)


def adaptive_retrieval(gate: Callable, retrieve_ids: Callable, X0: torch.Tensor, q_indptr: torch.Tensor,
                       q_terms: torch.Tensor, step_fn: Callable, k: int, theta: float = 0.0,
                       ablation: int = 0) -> RoundsResult:
    """`gate(X, theta=, ablation=)` -> object with `.retrieve` bool[n] (ProberGate);
    `retrieve_ids(q_indptr, q_terms, k)` -> (scores [n,k], doc_ids [n,k]) (BM25Retriever.retrieve_ids)."""
    dev = X0.device
    nq = X0.shape[0]
    retr_count = torch.zeros(nq, dtype=torch.int32, device=dev)
    calls = torch.zeros(nq, dtype=torch.int32, device=dev)
    last_s = torch.full((nq, k), float("-inf"), dtype=torch.float32, device=dev)
    last_d = torch.full((nq, k), -1, dtype=torch.int32, device=dev)
    per_call = []

    need = gate(X0, theta=theta, ablation=ablation).retrieve.to(dev)         # :406-415
    active = torch.nonzero(need, as_tuple=False).flatten()                   # ascending question ids
    cur_indptr, cur_terms = select_queries(q_indptr, q_terms, active)        # first call: the question (:426)
    call = 0
    while active.numel() > 0:
        call += 1
        s, d = retrieve_ids(cur_indptr, cur_terms, k)                        # :426 / :428
        last_s[active] = s
        last_d[active] = d
        calls[active] += 1
        per_call.append(active)
        X_next, (n_indptr, n_terms) = step_fn(active, s, d, call)            # :439-443, :457
        again = gate(X_next, theta=theta, ablation=ablation).retrieve.to(dev)   # :446-455
        capped = retr_count[active] > MAX_RETR_COUNT - 1                     # :462 `retr_count > 2`
        retr_count[active] += (~capped).to(torch.int32)                      # :465
        keep = again & ~capped
        sel = torch.nonzero(keep, as_tuple=False).flatten()
        active = active[sel]
        cur_indptr, cur_terms = select_queries(n_indptr, n_terms, sel)       # later calls: the transcript (:428)
    assert call <= MAX_CALLS
    return RoundsResult(retr_count=retr_count, calls=calls, last_scores=last_s, last_doc_ids=last_d,
                        per_call_active=per_call)
