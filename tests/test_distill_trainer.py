"""The loss consumers against the reference's REAL trainer.

`tests/golden/distill_trainer_tiny.npz` was produced by the UNMODIFIED `loss_trainer.ImageDistilTrainer.forward`
(oracle/make_golden_distill.py).  Here:
  * CPU (`not gpu`): `odevit_b200.loss_trainer.ImageDistilTrainer` driving the ORACLE student and the HF teacher
    reproduces every loss term and the post-clip gradients -- pins the trainer restatement itself;
  * GPU: the same trainer driving `odevit_b200.ViTNeuralODE` (libodevit) + the HF teacher / `ViTTeacher`
    reproduces them through the CUDA path (fp32 mode: loss <= 1e-4, gradients <= 2e-3; bf16 mode: <= 2e-2 / 5e-2).
"""
import json

import pytest
import torch
from torch import nn

import odevit_oracle as orc
from _util import Golden, max_rel

SCALARS = ("loss", "mse_loss", "kl_loss", "jasmin_loss", "supervision_loss") + tuple(f"mse_loss_t@{i}" for i in range(12))


def _teacher(g: Golden):
    from transformers import ViTConfig, ViTForImageClassification
    t = ViTForImageClassification(ViTConfig(attn_implementation="eager", **g.meta["teacher"]))
    t.load_state_dict(g.group("tsd"), strict=True)
    return t.eval()


class OracleStudent(nn.Module):
    """The oracle's functional forward behind the student's call signature (CPU checker, test only)."""

    def __init__(self, sd, cfg):
        super().__init__()
        self.cfg = cfg
        self.names = list(sd.keys())
        self.ps = nn.ParameterList([nn.Parameter(v.clone()) for v in sd.values()])

    def forward(self, pixel_values, labels=None, **kw):
        sd = dict(zip(self.names, self.ps))
        return orc.vit_ode_forward(sd, self.cfg, pixel_values, labels=labels, **kw)


def _check(out, student_named_grads, g: Golden, tag, tol_loss, tol_grad, cos_min=None):
    for k in SCALARS:
        want = float(g.get(f"{tag}/{k}"))
        assert float(out[k]) == pytest.approx(want, rel=tol_loss, abs=tol_loss * 1e-2), (tag, k)
    worst = 0.0
    for k, grad in student_named_grads:
        want = g.get(f"{tag}/grad/{k}")
        if float(want.abs().max()) == 0.0:
            continue
        assert grad is not None, k
        worst = max(worst, max_rel(grad, want))
        assert max_rel(grad, want) < tol_grad, (tag, k)
        if cos_min is not None and want.numel() > 1:
            cos = torch.nn.functional.cosine_similarity(grad.detach().double().cpu().flatten(), want.double().flatten(), dim=0)
            assert float(cos) > cos_min, (tag, k, float(cos))
    return worst


def test_blur_matches_torchvision():
    from torchvision.transforms.functional import gaussian_blur
    from odevit_b200.loss_trainer import _blur3x3
    x = torch.randn(3, 5, 14, 14, generator=torch.Generator().manual_seed(0))
    assert torch.allclose(_blur3x3(x, 0.5), gaussian_blur(x, kernel_size=(3, 3), sigma=0.5), atol=1e-6)


@pytest.mark.parametrize("tag,epoch", [("e0", 0), ("e201", 201)])
def test_trainer_restatement_cpu_vs_reference_trainer(tag, epoch):
    from odevit_b200.loss_trainer import ImageDistilTrainer
    g = Golden("distill_trainer_tiny")
    student = OracleStudent(g.group("sd"), g.meta["student"])
    opt = torch.optim.SGD(student.parameters(), lr=0.0)
    tr = ImageDistilTrainer(teacher_model=_teacher(g), student_model=student, optimizer=opt, **g.meta["trainer"])
    out = tr({"pixel_values": g.get("in/pixel_values")}, g.get("in/labels"), epoch=epoch)
    _check(out, [(n, p.grad) for n, p in zip(student.names, student.ps)], g, tag, 2e-5, 2e-4)
    assert max_rel(out["student_output"]["control_points"], g.get(f"{tag}/control_points")) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("teacher_kind", ["hf", "library"])
@pytest.mark.parametrize("precision,tol_loss,tol_grad", [("fp32", 1e-4, 2e-3), ("bf16", 2e-2, 6e-2)])
def test_trainer_gpu_vs_reference_trainer(precision, tol_loss, tol_grad, teacher_kind):
    import odevit_b200 as ob
    from odevit_b200.loss_trainer import ImageDistilTrainer
    g = Golden("distill_trainer_tiny")
    student = ob.ViTNeuralODE(**g.meta["student"])
    student.load_state_dict(g.group("sd"), strict=True)
    student = student.cuda()
    student.precision = precision
    teacher = _teacher(g).cuda()
    if teacher_kind == "library":
        teacher = ob.ViTTeacher(teacher, precision=precision, attention_maps="last")
    opt = torch.optim.SGD(student.parameters(), lr=0.0)
    tr = ImageDistilTrainer(teacher_model=teacher, student_model=student, optimizer=opt, **g.meta["trainer"])
    px, lb = g.get("in/pixel_values").cuda(), g.get("in/labels").cuda()
    for tag, epoch in (("e0", 0), ("e201", 201)):
        out = tr({"pixel_values": px}, lb, epoch=epoch)
        worst = _check(out, [(n, p.grad) for n, p in student.named_parameters()], g, tag, tol_loss, tol_grad,
                       cos_min=0.999 if precision == "bf16" else 0.999999)
        print(json.dumps({"case": tag, "precision": precision, "teacher": teacher_kind, "loss": float(out["loss"]),
                          "worst_grad_max_rel": worst}))
    assert out["student_output"]["logits"].argmax(-1).cpu().tolist() == g.get("e201/logits").argmax(-1).tolist()
