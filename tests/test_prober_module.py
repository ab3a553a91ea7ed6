"""ImprovedProbe drop-in on CPU: same state_dict, same eager arithmetic as the oracle."""
import torch

from oracle import prober_oracle as po
from probing_rag_b200.prober import STATE_KEYS, ImprovedProbe, split_bf16


def test_state_dict_and_eager_forward_match_oracle():
    sd = po.make_prober_state(8)
    m = ImprovedProbe(input_size=2048, output_size=2)
    assert tuple(sorted(m.state_dict().keys())) == tuple(sorted(STATE_KEYS)) == tuple(sorted(po.STATE_KEYS))
    m.load_state_dict(sd)
    m.eval()
    o = po.OracleImprovedProbe(2048, 2)
    o.load_state_dict(sd)
    o.eval()
    x = po.make_hidden_states(9, seed=1)[:, 0]
    with torch.no_grad():
        assert torch.equal(m(x), o(x))          # CPU tensors take the plain-module path


def test_split_bf16_reconstructs_to_16_bits():
    w = torch.randn(512, 2048)
    hi, lo = split_bf16(w)
    err = (hi.float() + lo.float() - w).abs().max() / w.abs().max()
    assert err < 2 ** -15
