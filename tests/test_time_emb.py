"""models/time_emb.py mirrored as PyTorch modules (odevit_b200/time_emb.py): against the golden
vectors the unmodified reference produced (tests/golden/time_emb.npz) and the oracle."""
import pytest
import torch

import odevit_oracle as orc
from _util import Golden, max_rel


def test_modules_match_reference_goldens():
    from odevit_b200 import time_emb as te
    g = Golden("time_emb")
    t = g.get("t")
    assert max_rel(te.SinusoidalPosEmb(16)(t), g.get("fourier")) < 1e-6
    emb = te.TimeEmbedding(sinusoidal_dim=16, embed_dim=64, multiplier=2, dropout=0.1, learnable_sinusoidal=True).eval()
    emb.load_state_dict(g.group("sd_emb"), strict=True)
    ss = te.ScaleShift(embed_dim=64, out_dim=64)
    ss.load_state_dict(g.group("sd_ss"), strict=True)
    with torch.no_grad():
        e = emb(t)
        scale, shift = ss(e)
    assert max_rel(e, g.get("emb")) < 1e-6
    assert max_rel(scale, g.get("scale")) < 1e-6
    assert max_rel(shift, g.get("shift")) < 1e-6
    assert scale.shape == shift.shape == (7, 64)


def test_non_learnable_width_mismatch_is_kept():
    """The reference's TimeEmbedding(learnable_sinusoidal=False) fails: lin1 expects 2*sd+1 inputs,
    SinusoidalPosEmb(sd) yields sd+1 (time_emb.py:92 vs :20-39)."""
    from odevit_b200 import time_emb as te
    emb = te.TimeEmbedding(sinusoidal_dim=16, embed_dim=8)
    assert te.SinusoidalPosEmb(16)(torch.zeros(3)).shape == (3, 17)
    with pytest.raises(RuntimeError):
        emb(torch.zeros(3))


def test_oracle_time_embedding_agrees():
    from odevit_b200 import time_emb as te
    g = Golden("time_emb")
    t = torch.linspace(0, 2, 5)
    assert max_rel(te.SinusoidalPosEmb(16)(t), orc.sinusoidal_pos_emb(t, 16)) < 1e-6
