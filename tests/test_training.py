"""Prober training step against a literal restatement of /root/reference/train.py."""
import copy
import os

import pytest
import torch

from oracle import prober_oracle as po
from probing_rag_b200.prober import STATE_KEYS, ImprovedProbe
from probing_rag_b200.training import ProberTrainer, checkpoint_name, tokens_mean_inputs


def batch(seed, B=6, T=40, d=256):
    g = torch.Generator().manual_seed(seed)
    acts = torch.randn(B, T, d, generator=g)
    pred_lens = torch.randint(1, T + 1, (B,), generator=g)
    pred_lens[0], pred_lens[1] = 1, T              # edge cases: one token, the whole sequence
    labels = torch.randint(0, 2, (B,), generator=g)
    return acts, labels, pred_lens


def test_tokens_mean_inputs_match_the_slice_concat_split_mean_loop():
    acts, labels, pred_lens = batch(1)
    ref = po.tokens_mean_reference(acts, labels, pred_lens)
    got = tokens_mean_inputs(acts, pred_lens)
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        tokens_mean_inputs(acts, torch.zeros(acts.shape[0], dtype=torch.long))


def test_train_steps_follow_the_reference_loop(tmp_path):
    torch.manual_seed(16)                           # train.py:30 seeds by layer
    mine = ImprovedProbe(256, 2)
    ref = po.OracleImprovedProbe(256, 2)
    ref.load_state_dict(copy.deepcopy(mine.state_dict()))
    mine.dropout.p = ref.dropout.p = 0.0            # make the two runs comparable (dropout draws differ otherwise)
    tr = ProberTrainer(prober=mine, lr=1e-3, device="cpu")
    opt = torch.optim.AdamW(ref.parameters(), lr=1e-3)
    sch = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.995)
    ref.train()
    for step in range(4):
        acts, labels, pred_lens = batch(100 + step)
        l_ref, lr_ref = po.train_step_reference(ref, opt, sch, acts, labels, pred_lens)
        l_mine, lr_mine = tr.train_step(acts, labels, pred_lens)
        assert abs(l_mine - l_ref) < 2e-4 and lr_mine == pytest.approx(lr_ref)
    # AdamW normalises every gradient component to ~lr, so ulp-level differences of near-zero gradients (the masked
    # reduction vs the slice/mean loop) move single weights by up to lr per step: compare at that scale, and the
    # gradients themselves tightly below
    for k in STATE_KEYS:
        assert torch.allclose(mine.state_dict()[k], ref.state_dict()[k], rtol=0, atol=4 * 1e-3 + 1e-6), k
    acts, labels, pred_lens = batch(55)
    ref.load_state_dict(copy.deepcopy(mine.state_dict()))
    from probing_rag_b200.training import make_loss
    mine.train()
    make_loss(mine, tokens_mean_inputs(acts, pred_lens), labels)[0].backward()
    x = po.tokens_mean_reference(acts, labels, pred_lens)
    torch.nn.CrossEntropyLoss()(torch.nn.Softmax(dim=-1)(ref(x)), labels).backward()
    for (n1, p1), (n2, p2) in zip(mine.named_parameters(), ref.named_parameters()):
        assert n1 == n2 and torch.allclose(p1.grad, p2.grad, rtol=1e-3, atol=1e-6), n1
    tr.optim.zero_grad()
    acc, n, loss = tr.eval_step(*batch(7))
    assert 0.0 <= acc <= 1.0 and n == 6 and loss > 0
    # checkpoint: the reference's file name pattern and the 12-tensor state_dict its loader expects
    name = checkpoint_name(1.0, "google/gemma-2b", "tokens_mean", 2, 16, "resid_post", 1, root=str(tmp_path))
    assert os.path.basename(name) == "in3_1.0_gemma-2b_tokens_mean_2_l16_resid_post_ep1.pt"
    tr.save(name)
    sd = torch.load(name)
    assert set(sd) == set(STATE_KEYS) and all(v.device.type == "cpu" for v in sd.values())
    fresh = po.OracleImprovedProbe(256, 2)
    fresh.load_state_dict(sd)                       # utils.py:302-326 path: load_state_dict(torch.load(path))


@pytest.mark.gpu
def test_train_and_eval_steps_on_the_gpu_follow_the_reference_loop(tmp_path):
    """SURVEY 8f-4 on the device the hot path runs on: the training step (tokens_mean inputs as one masked
    reduction, CE-on-softmax, AdamW + ExponentialLR per batch) with the gemma-2b prober shape on cuda:0 tracks
    the literal restatement of train.py:199-220 run on the CPU, the eval step goes through the fused tcgen05
    forward, and the checkpoint it writes is the file utils.load_prober reads (train.py:344-345, utils.py:316)."""
    torch.manual_seed(16)
    d = 2048
    mine = ImprovedProbe(d, 2)
    ref = po.OracleImprovedProbe(d, 2)
    ref.load_state_dict(copy.deepcopy(mine.state_dict()))
    mine.dropout.p = ref.dropout.p = 0.0
    tr = ProberTrainer(prober=mine, lr=1e-4, device="cuda")
    opt = torch.optim.AdamW(ref.parameters(), lr=1e-4)
    sch = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.995)
    ref.train()
    for step in range(3):
        acts, labels, pred_lens = batch(200 + step, B=16, T=24, d=d)
        l_ref, lr_ref = po.train_step_reference(ref, opt, sch, acts, labels, pred_lens)
        l_gpu, lr_gpu = tr.train_step(acts.cuda(), labels.cuda(), pred_lens)
        assert abs(l_gpu - l_ref) < 5e-4 and lr_gpu == pytest.approx(lr_ref)
    assert next(tr.prober.parameters()).is_cuda
    for k in STATE_KEYS:                             # AdamW moves a weight by <= lr per step (see the CPU test)
        assert torch.allclose(tr.prober.state_dict()[k].cpu(), ref.state_dict()[k], rtol=0, atol=3 * 1e-4 + 1e-6), k
    # eval step: fused forward on the device vs the plain module on the same weights
    acts, labels, pred_lens = batch(77, B=40, T=24, d=d)
    acc, n, loss = tr.eval_step(acts.cuda(), labels.cuda(), pred_lens)
    ref.load_state_dict({k: v.cpu() for k, v in tr.prober.state_dict().items()})
    ref.eval()
    with torch.no_grad():
        x = po.tokens_mean_reference(acts, labels, pred_lens)
        probs = torch.softmax(ref(x), -1)
        loss_ref = float(torch.nn.functional.cross_entropy(probs, labels).item())
    assert n == 40 and abs(loss - loss_ref) < 1e-3
    margin = (probs[:, 0] - probs[:, 1]).abs()
    assert abs(acc - float((probs.argmax(-1) == labels).double().mean())) <= float((margin < 2e-3).sum()) / 40 + 1e-9
    name = checkpoint_name(1.0, "google/gemma-2b", "tokens_mean", 2, 16, "resid_post", 1, root=str(tmp_path))
    tr.save(name)
    sd = torch.load(name)
    assert set(sd) == set(STATE_KEYS) and all(v.device.type == "cpu" for v in sd.values())
