"""Live cross-check of the restatement against the UNMODIFIED reference modules, where
/root/reference is present (the build container).  Skipped on the GPU box."""
import pytest
import torch

import odevit_oracle as orc
from _util import max_rel
from ref_import import import_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def mods():
    return import_reference()


@pytest.mark.parametrize("solver,T", [("euler", 7), ("rk4", 4)])
@pytest.mark.parametrize("posreg", [False, True])
def test_vit_live(mods, solver, T, posreg):
    ctor = dict(img_size=16, patch_size=4, num_classes=5, embed_dim=48, num_heads=3, mlp_ratio=1.0,
                emulate_depth=12, time_interval=1.0, num_eval_steps=T, solver=solver,
                register_tokens=3, pos_embed_register_tokens=posreg)
    torch.manual_seed(11)
    ref = mods["ode"].ViTNeuralODE(**ctor)
    px = torch.randn(2, 3, 16, 16)
    labels = torch.tensor([1, 4])
    want = ref(px, labels=labels, output_hidden_states=True, output_attentions=True,
               output_attention_trajectory=True, jasmin_k=2)
    got = orc.vit_ode_forward(dict(ref.state_dict()), ctor, px, labels=labels, output_hidden_states=True,
                              output_attentions=True, output_attention_trajectory=True, jasmin_k=2)
    for k in ("logits", "loss", "states", "attentions", "attention_trajectory", "jasmin_loss"):
        assert got[k].shape == want[k].shape, k
        if want[k].numel():
            assert max_rel(got[k], want[k]) < 2e-6, k


def test_macaron_live(mods):
    ctor = dict(img_size=16, patch_size=4, num_classes=5, embed_dim=48, num_heads=3, mlp_ratio=2.0,
                emulate_depth=12, time_interval=12.0, num_eval_steps=4, solver="rk4")
    torch.manual_seed(12)
    ref = mods["macaron"].ViTMacaron(**ctor)
    px = torch.randn(2, 3, 16, 16)
    want = ref(px, output_hidden_states=True)
    got = orc.macaron_forward(dict(ref.state_dict()), ctor, px, output_hidden_states=True)
    assert max_rel(got["states"], want["states"]) < 2e-6
    assert max_rel(got["logits"], want["logits"]) < 2e-6
