"""odevit_b200.ViTTeacher (odevit_encoder_fwd: the distillation teacher's encoder through the library, SURVEY
section 8 row (f)3) against the HF model it wraps, run the way the reference runs it: eager attention, fp32,
no grad (loss_trainer.py:318-321)."""
import pytest
import torch

from _util import max_rel

pytestmark = pytest.mark.gpu


def _hf(layers=3, dim=768, heads=12, inter=3072, img=224):
    tr = pytest.importorskip("transformers")
    cfg = tr.ViTConfig(num_labels=100, num_hidden_layers=layers, hidden_size=dim, num_attention_heads=heads,
                       intermediate_size=inter, image_size=img, attn_implementation="eager")
    torch.manual_seed(1)
    return tr.ViTForImageClassification(cfg).cuda().eval()


@pytest.mark.parametrize("precision,tol,qk_gain", [("fp32", 2e-4, 6.0), ("bf16", 2e-2, 2.5)])
def test_teacher_matches_hf_vit(precision, tol, qk_gain):
    import odevit_b200 as ob
    hf = _hf()
    # a random-init ViT has near-uniform attention and tiny residual updates: scale the projections so that the
    # softmax and the GELU are exercised away from their linear regime
    with torch.no_grad():
        for lyr in hf.vit.encoder.layer:
            lyr.attention.attention.query.weight.mul_(qk_gain)   # logits of O(qk_gain^2): bf16 operands put an
            lyr.attention.attention.key.weight.mul_(qk_gain)     # absolute error of ~2^-8 |logit| on them
            lyr.intermediate.dense.weight.mul_(3.0)
            lyr.intermediate.dense.bias.normal_(0, 0.5)
            lyr.attention.attention.query.bias.normal_(0, 0.5)
    px = torch.randn(5, 3, 224, 224, generator=torch.Generator().manual_seed(2)).cuda()
    with torch.no_grad():
        want = hf(pixel_values=px, output_hidden_states=True, output_attentions=True)
    teacher = ob.ViTTeacher(hf, precision=precision, attention_maps="all")
    got = teacher(pixel_values=px, output_hidden_states=True, output_attentions=True)
    assert len(got["hidden_states"]) == len(want.hidden_states) == 4
    for a, b in zip(got["hidden_states"], want.hidden_states):
        assert a.shape == b.shape
        assert max_rel(a, b) < tol
    for a, b in zip(got["attentions"], want.attentions):
        assert a.shape == b.shape == (5, 12, 197, 197)
        assert max_rel(a, b) < 5 * tol
    assert max_rel(got["logits"], want.logits) < 5 * tol
    # what the reference's losses read (loss_trainer.py:169-171, :259), also through attribute access
    last = ob.ViTTeacher(hf, precision=precision, attention_maps="last")(pixel_values=px)
    # (the exporting attention kernel rounds the normalised map to bf16, the other one the un-normalised exponentials:
    # the two modes agree to bf16 rounding, not bitwise)
    assert max_rel(torch.stack(last.attentions, dim=0)[-1], want.attentions[-1]) < 5 * tol
    assert max_rel(torch.stack(last["hidden_states"], dim=0)[1:], torch.stack(want.hidden_states, dim=0)[1:]) < tol
    # second call reuses the prepared weights and gives the same result
    again = teacher(pixel_values=px)
    assert torch.equal(again["hidden_states"][-1], got["hidden_states"][-1])
