"""Prober side of the hot path: `ImprovedProbe` (same class name, constructor, state_dict keys
and call shape as /root/reference/utils.py:29-57, so `load_prober` / `load_state_dict(torch.load(..))`
at utils.py:302-326 keep working) and `ProberGate`, the batched replacement of the gating code at
/root/reference/exp_rag.py:381-415 backed by libprobingrag.so's tcgen05 kernels.

    probers = [ImprovedProbe(2048, 2) ...]            # utils.py:302, one per layer 6,8,..,16 (exp_rag.py:311)
    gate = ProberGate(probers)                         # packs + splits the weights once
    out = gate(X, theta=0.0, ablation=0)               # X[B, 6, 2048] pooled hidden states (exp_rag.py:385-386)
    out.retrieve_idx                                   # rows that must retrieve -> compacted BM25 batch
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import torch
from torch import nn

from . import _lib

_X_DTYPES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}   # pr_prober_forward x_dtype codes

STATE_KEYS = (
    "layer_norm_input.weight", "layer_norm_input.bias",
    "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc3.weight", "fc3.bias",
    "layer_norm1.weight", "layer_norm1.bias", "layer_norm2.weight", "layer_norm2.bias",
)


def split_bf16(w: torch.Tensor):
    """fp32 -> (hi, lo) bf16 with hi + lo == w to ~2^-17 relative: operands of the bf16x3 GEMM."""
    hi = w.to(torch.bfloat16)
    lo = (w - hi.to(torch.float32)).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


class ImprovedProbe(nn.Module):
    """utils.py:29-57.  LayerNorm -> fc1 -> SiLU -> LayerNorm -> (dropout) -> fc2 -> SiLU ->
    LayerNorm -> (dropout) -> fc3.  In eval mode on a CUDA tensor the forward runs through the
    fused kernels (as a one-prober ProberGate); in training mode it is the plain module."""

    def __init__(self, input_size, output_size, hidden_size=512):
        super().__init__()
        self.layer_norm_input = nn.LayerNorm(normalized_shape=input_size)
        self.fc1 = nn.Linear(input_size, hidden_size)
        self.fc2 = nn.Linear(hidden_size, hidden_size)
        self.fc3 = nn.Linear(hidden_size, output_size)
        self.silu = nn.SiLU()
        self.dropout = nn.Dropout(p=0.1)
        self.layer_norm1 = nn.LayerNorm(normalized_shape=hidden_size)
        self.layer_norm2 = nn.LayerNorm(normalized_shape=hidden_size)
        self._gate = None
        self._gate_version = None

    def forward_eager(self, x):
        x = self.layer_norm_input(x)
        x = self.dropout(self.layer_norm1(self.silu(self.fc1(x))))
        x = self.dropout(self.layer_norm2(self.silu(self.fc2(x))))
        return self.fc3(x)

    def _fusable(self, x) -> bool:
        return (not self.training and x.is_cuda and x.dim() == 2 and self.fc3.out_features == 2
                and self.fc1.out_features == 512 and self.fc1.in_features % 128 == 0
                and self.fc1.in_features <= 2048 and not torch.is_grad_enabled())

    def forward(self, x):
        if not self._fusable(x):
            return self.forward_eager(x)
        version = tuple(p._version for p in self.parameters()) + (str(x.device),)
        if self._gate is None or self._gate_version != version:
            self._gate = ProberGate([self])
            self._gate_version = version
        return self._gate(x.unsqueeze(1), want_logits=True).logits[:, 0]


@dataclass
class GateOutput:
    probsum: torch.Tensor          # f32[B, 2]   sum over probers >= ablation of softmax(logits)   (exp_rag.py:407-410)
    retrieve: torch.Tensor         # bool[B]     NOT (probsum[0] + theta < probsum[1])             (exp_rag.py:414-415)
    retrieve_idx: torch.Tensor     # i32[n]      rows that retrieve, ascending (device).  sync=True: exactly the n valid
                                   #             entries; sync=False: all B slots, the first n_retrieve valid, the rest -1
    n_retrieve: torch.Tensor       # i32[1]      on the device (no host sync needed to use the async result)
    logits: torch.Tensor | None    # f32[B, P, 2]


class ProberGate:
    """Six probers + softmax-sum gate + compaction in four kernel launches."""

    def __init__(self, probers, device=None):
        probers = list(probers)
        if not 1 <= len(probers) <= _lib.PR_PROBER_MAX:
            raise ValueError(f"between 1 and {_lib.PR_PROBER_MAX} probers")
        sds = [p.state_dict() if isinstance(p, nn.Module) else p for p in probers]
        dev = torch.device(device) if device is not None else sds[0]["fc1.weight"].device
        if dev.type != "cuda":
            raise RuntimeError("ProberGate needs a CUDA device: the fused prober has no CPU fallback")
        self.device = dev
        self.n_probers = len(sds)
        self.hidden, self.d_model = sds[0]["fc1.weight"].shape
        for sd in sds:
            if tuple(sd["fc1.weight"].shape) != (self.hidden, self.d_model) or sd["fc3.weight"].shape[0] != 2:
                raise ValueError("all probers must share d_model/hidden and have 2 classes")

        def stack(key):
            return torch.stack([sd[key].detach().to(dev, torch.float32) for sd in sds]).contiguous()

        self.t = {k: stack(k) for k in STATE_KEYS if k not in ("fc1.weight", "fc2.weight")}
        # The input LayerNorm is folded around fc1 (include/probing_rag.h): fc1(LN(x)) = rstd * (W' x - mean * rowsum(W'))
        # + (W beta + b1) with W' = W diag(gamma).  Folded once per checkpoint, in float64, rounded to f32 once.
        w1 = stack("fc1.weight").double()
        gamma, beta = self.t["layer_norm_input.weight"].double(), self.t["layer_norm_input.bias"].double()
        w1g = (w1 * gamma[:, None, :]).float()
        self.w1_hi, self.w1_lo = split_bf16(w1g)
        # row sums of exactly what the tensor cores multiply by (hi + lo), so the mean term cancels what they add
        self.w1_rowsum = (self.w1_hi.double() + self.w1_lo.double()).sum(-1).float().contiguous()
        self.b1_folded = (torch.einsum("pod,pd->po", w1, beta) + self.t["fc1.bias"].double()).float().contiguous()
        self.w2_hi, self.w2_lo = split_bf16(stack("fc2.weight"))
        t = self.t
        self._set = _lib.ProberSet(
            n_probers=self.n_probers, d_model=self.d_model, hidden=self.hidden,
            w1_rowsum=self.w1_rowsum.data_ptr(), b1=self.b1_folded.data_ptr(), ln1_w=t["layer_norm1.weight"].data_ptr(), ln1_b=t["layer_norm1.bias"].data_ptr(),
            b2=t["fc2.bias"].data_ptr(), ln2_w=t["layer_norm2.weight"].data_ptr(), ln2_b=t["layer_norm2.bias"].data_ptr(),
            w3=t["fc3.weight"].data_ptr(), b3=t["fc3.bias"].data_ptr(),
            w1_hi=self.w1_hi.data_ptr(), w1_lo=self.w1_lo.data_ptr(),
            w2_hi=self.w2_hi.data_ptr(), w2_lo=self.w2_lo.data_ptr())
        self._ws = {}

    def _workspace(self, n_rows: int) -> torch.Tensor:
        ws = self._ws.get(n_rows)
        if ws is None:
            nbytes = int(_lib.lib().pr_prober_workspace_bytes(self.n_probers, n_rows, self.d_model, self.hidden))
            if len(self._ws) > 4:
                self._ws.clear()
            ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
            self._ws[n_rows] = ws
        return ws

    def flops(self, n_rows: int) -> int:
        """Algorithmic flops: 2*(d*h + h*h + h*2) per prober and row."""
        return n_rows * self.n_probers * 2 * (self.d_model * self.hidden + self.hidden * self.hidden + self.hidden * 2)

    @torch.no_grad()
    def __call__(self, X: torch.Tensor, theta: float = 0.0, ablation: int = 0, want_logits: bool = False,
                 sync: bool = True) -> GateOutput:
        """X [B, P, d_model] pooled hidden states: f32, bf16 or f16 (read as they are -- a bf16 LM's sums
        need no widening pass), anything else is converted to f32."""
        if X.dim() != 3 or X.shape[1] != self.n_probers or X.shape[2] != self.d_model:
            raise ValueError(f"X must be [B, {self.n_probers}, {self.d_model}], got {tuple(X.shape)}")
        x_dtype = _X_DTYPES.get(X.dtype)
        if x_dtype is None:
            X, x_dtype = X.to(torch.float32), 0
        X = X.to(self.device).contiguous()
        n = X.shape[0]
        dev = self.device
        logits = torch.empty((n, self.n_probers, 2), dtype=torch.float32, device=dev) if want_logits else None
        probsum = torch.empty((n, 2), dtype=torch.float32, device=dev)
        mask = torch.empty(n, dtype=torch.uint8, device=dev)
        compact = torch.empty(n, dtype=torch.int32, device=dev)
        n_ret = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = self._workspace(n)
        ws_ptr = (ws.data_ptr() + 1023) // 1024 * 1024
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().pr_prober_forward(
                ctypes.byref(self._set), n, X.data_ptr(), x_dtype, float(theta), int(ablation),
                logits.data_ptr() if want_logits else None, probsum.data_ptr(), mask.data_ptr(), compact.data_ptr(),
                n_ret.data_ptr(), ws_ptr, ws.numel() - (ws_ptr - ws.data_ptr()),
                torch.cuda.current_stream(dev).cuda_stream))
        idx = compact[: int(n_ret.item())] if sync else compact
        return GateOutput(probsum=probsum, retrieve=mask.bool(), retrieve_idx=idx, n_retrieve=n_ret, logits=logits)


def gate_and_retrieve(gate: ProberGate, retriever, X: torch.Tensor, q_indptr: torch.Tensor, q_terms: torch.Tensor,
                      theta: float = 0.0, ablation: int = 0, k: int | None = None):
    """BASELINE config 4: prober forward -> retrieve/no-retrieve mask -> compacted BM25 top-k.
    Returns (GateOutput, scores f32[n_retrieve, k], doc_ids i32[n_retrieve, k]); row i of the
    results belongs to query out.retrieve_idx[i]."""
    from .rounds import select_queries
    out = gate(X, theta=theta, ablation=ablation)
    c_indptr, c_terms = select_queries(q_indptr, q_terms, out.retrieve_idx)
    scores, ids = retriever.retrieve_ids(c_indptr, c_terms, k)
    return out, scores, ids
