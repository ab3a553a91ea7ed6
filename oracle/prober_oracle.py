"""CPU oracle for the prober + gate -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py may import this module.

Restates, in plain torch fp32 / numpy:
  OracleImprovedProbe.forward   /root/reference/utils.py:29-57 (ImprovedProbe)
  pool_hidden_states            /root/reference/exp_rag.py:381-389 (sum over generated tokens,
                                prefill entry dropped)
  gate                          /root/reference/exp_rag.py:393, 407-415
  retrieval_rounds              /root/reference/exp_rag.py:422-468 (round control)

PINNED: tests/golden/make_golden.py imports the reference's own `ImprovedProbe`
(AST-extracted from /root/reference/utils.py) in the build container, checks this
restatement against it bit for bit and writes tests/golden/prober_golden.npz from the
REFERENCE class's outputs.
"""
from __future__ import annotations

import hashlib

import numpy as np
import torch
from torch import nn

D_MODEL = 2048                       # gemma-2b d_model (Config_Maker, utils.py:288)
HIDDEN = 512                         # ImprovedProbe default hidden_size (utils.py:30)
N_CLASSES = 2                        # Config_Maker.num_classes (utils.py:290)
PROBE_LAYERS = tuple(range(6, 17, 2))  # exp_rag.py:311

STATE_KEYS = (
    "layer_norm_input.weight", "layer_norm_input.bias",
    "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc3.weight", "fc3.bias",
    "layer_norm1.weight", "layer_norm1.bias", "layer_norm2.weight", "layer_norm2.bias",
)


class OracleImprovedProbe(nn.Module):
    """Same modules, names and forward order as utils.py:29-57."""

    def __init__(self, input_size, output_size, hidden_size=HIDDEN):
        super().__init__()
        self.layer_norm_input = nn.LayerNorm(normalized_shape=input_size)
        self.fc1 = nn.Linear(input_size, hidden_size)
        self.fc2 = nn.Linear(hidden_size, hidden_size)
        self.fc3 = nn.Linear(hidden_size, output_size)
        self.silu = nn.SiLU()
        self.dropout = nn.Dropout(p=0.1)
        self.layer_norm1 = nn.LayerNorm(normalized_shape=hidden_size)
        self.layer_norm2 = nn.LayerNorm(normalized_shape=hidden_size)

    def forward(self, x):
        x = self.layer_norm_input(x)
        x = self.dropout(self.layer_norm1(self.silu(self.fc1(x))))
        x = self.dropout(self.layer_norm2(self.silu(self.fc2(x))))
        return self.fc3(x)


# synthetic checkpoints / hidden states are workload generators and live with the other ones
from probing_rag_b200.synth import make_hidden_states, make_prober_state  # noqa: E402,F401


def state_digest(sd: dict) -> str:
    h = hashlib.sha256()
    for k in STATE_KEYS:
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def pool_hidden_states(cache_entries):
    """exp_rag.py:385-386: cat(cache[name][1:], dim=1) then SUM over tokens -> [1, d]."""
    x = torch.concat(list(cache_entries)[1:], dim=1)
    return torch.sum(x, dim=1)


@torch.no_grad()
def prober_logits(probers, x: torch.Tensor) -> torch.Tensor:
    """x[B, P, d] -> logits[B, P, 2] with prober p applied to x[:, p] (exp_rag.py:406)."""
    return torch.stack([p(x[:, i]) for i, p in enumerate(probers)], dim=1)


def gate(logits: torch.Tensor, theta: float = 0.0, ablation: int = 0):
    """exp_rag.py:407-415: P = sum_{l>=ablation} softmax(logits_l); no-retrieve iff
    P[0] + theta < P[1].  Returns (probsum[B,2], retrieve[B] bool)."""
    probs = torch.softmax(logits.float(), dim=-1)
    psum = probs[:, ablation:].sum(dim=1)
    retrieve = ~(psum[:, 0] + theta < psum[:, 1])
    return psum, retrieve


def retrieval_rounds(gate_decisions) -> int:
    """exp_rag.py:414-468 round control for one question.  gate_decisions[i] is the gate
    outcome (True = retrieve) after generation i.  Returns the number of `retrieve` calls
    issued (0..4); the reference records min(calls, 3) as retr_count."""
    it = iter(gate_decisions)
    if not next(it):
        return 0
    calls, retr_count = 0, 0
    while True:
        calls += 1                      # bm25.retrieve (:426 / :428)
        again = next(it)                # gate after the new generation (:446-455)
        if retr_count > 2:              # :462-463
            break
        retr_count += 1                 # :465
        if not again:
            break
    return calls


# ---- training step, literal restatement of /root/reference/train.py --------------
def tokens_mean_reference(cache_activation, labels, pred_lens):
    """train.py:155-165 + 199-205: slice each sample's last pred_len tokens, concat, split, mean."""
    result = []
    for i in range(cache_activation.size(0)):
        result.append(cache_activation[i, -int(pred_lens[i]):, :])
    input_tensor = torch.cat(result, dim=0)
    parts = torch.split(input_tensor, [int(p) for p in pred_lens.tolist()])
    return torch.cat([torch.mean(t, dim=0, keepdim=True) for t in parts], dim=0)


def train_step_reference(model, optim, scheduler, activations, labels, pred_lens):
    """train.py:141-151, 210-220 (`method_2_train`) with criterion_ce on softmax(output)."""
    x = tokens_mean_reference(activations, labels, pred_lens)
    output = model(x)
    logit = torch.nn.Softmax(dim=-1)(output)
    loss = torch.nn.CrossEntropyLoss()(logit, labels)
    loss.backward()
    optim.step()
    scheduler.step()
    optim.zero_grad()
    return round(loss.item(), 4), optim.param_groups[0]["lr"]
