"""Prober training step: /root/reference/train.py:141-151, 155-165, 199-220, 344-345.

The reference trains one `ImprovedProbe` per (layer, hook position) on frozen-LM activations with
method `tokens_mean`: the input of a sample is the MEAN of the activations of its last `pred_len`
tokens (sequences are left-padded, train.py:97-100), the loss is `CrossEntropyLoss` applied to the
SOFTMAX of the logits (train.py:148-149 -- a second log-softmax inside the criterion; kept as is,
the checkpoints the hot path loads were trained that way), the optimiser is AdamW with
ExponentialLR(gamma=0.995) stepped after every batch (train.py:131-136, 212-214), and the result is
`torch.save(prober.to('cpu').state_dict(), ckpt/_3/in3_..._l{layer}_resid_post_ep{epoch}.pt)`
(train.py:344-345) -- the file `utils.load_prober` reads back (utils.py:316).

Here the ragged per-sample slicing + concat + split + mean loop (train.py:155-165, 199-208) is one
masked reduction on the device; the MLP in training mode is the plain torch module (autograd), the
fused tcgen05 forward is the eval path.
"""
from __future__ import annotations

import os

import torch
from torch import nn

from .prober import ImprovedProbe


def tokens_mean_inputs(activations: torch.Tensor, pred_lens: torch.Tensor) -> torch.Tensor:
    """[B, T, d] activations of left-padded sequences, pred_lens[B] -> [B, d]: mean over each sample's
    last pred_len tokens (train.py:159-160 slices `[i, -pred_len:, :]`, :203-205 takes the mean)."""
    B, T, _ = activations.shape
    pl = pred_lens.to(activations.device).long()
    if bool((pl < 1).any()) or bool((pl > T).any()):
        raise ValueError("pred_lens must be in [1, T]")
    pos = torch.arange(T, device=activations.device)
    mask = (pos.unsqueeze(0) >= (T - pl).unsqueeze(1)).to(activations.dtype)          # [B, T]
    summed = torch.einsum("bt,btd->bd", mask, activations)
    return summed / pl.to(activations.dtype).unsqueeze(1)


def make_loss(prober: nn.Module, x: torch.Tensor, labels: torch.Tensor):
    """train.py:141-151 for the 2-class probers: loss = CE(softmax(logits), labels); returns (loss, probs)."""
    probs = torch.softmax(prober(x), dim=-1)
    return nn.functional.cross_entropy(probs, labels.long()), probs


def accuracy(probs: torch.Tensor, labels: torch.Tensor) -> float:
    """train.py:172-184."""
    return float((probs.argmax(dim=-1) == labels.to(probs.device)).sum().item()) / labels.numel()


def checkpoint_name(train_ratio, model_id: str, method: str, num_classes: int, layer: int, position: str,
                    epoch: int, root: str = "ckpt/_3") -> str:
    """train.py:345 / utils.py:316 path pattern."""
    return os.path.join(root, f"in3_{train_ratio}_{model_id.split('/')[1]}_{method}_{num_classes}_l{layer}_{position}_ep{epoch}.pt")


class ProberTrainer:
    """One prober, AdamW + ExponentialLR(0.995) stepped per batch, method `tokens_mean`."""

    def __init__(self, d_model: int = 2048, num_classes: int = 2, lr: float = 1e-4, device="cuda",
                 prober: ImprovedProbe | None = None, gamma: float = 0.995):
        self.device = torch.device(device)
        self.prober = (prober if prober is not None else ImprovedProbe(d_model, num_classes)).to(self.device)
        self.optim = torch.optim.AdamW(self.prober.parameters(), lr=lr)
        self.sched = torch.optim.lr_scheduler.ExponentialLR(self.optim, gamma=gamma)

    def train_step(self, activations: torch.Tensor, labels: torch.Tensor, pred_lens: torch.Tensor):
        """train.py:210-220 (`method_2_train`): returns (loss, learning rate after the step)."""
        self.prober.train()
        x = tokens_mean_inputs(activations.to(self.device), pred_lens)
        loss, _ = make_loss(self.prober, x, labels.to(self.device))
        loss.backward()
        self.optim.step()
        self.sched.step()
        self.optim.zero_grad()
        return float(loss.item()), self.optim.param_groups[0]["lr"]

    @torch.no_grad()
    def eval_step(self, activations: torch.Tensor, labels: torch.Tensor, pred_lens: torch.Tensor):
        """train.py:222-225 (`method_2_eval`): (accuracy, n, loss).  Eval mode: on a CUDA device the forward
        runs through the fused kernels."""
        self.prober.eval()
        x = tokens_mean_inputs(activations.to(self.device), pred_lens)
        loss, probs = make_loss(self.prober, x, labels.to(self.device))
        return accuracy(probs, labels), int(labels.numel()), float(loss.item())

    def save(self, path: str) -> None:
        """train.py:344-345: the 12-tensor state_dict on the CPU (what utils.load_prober loads)."""
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        sd = {k: v.detach().to("cpu") for k, v in self.prober.state_dict().items()}
        torch.save(sd, path)
