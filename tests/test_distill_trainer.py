"""The loss consumers against the reference's REAL trainer.

`tests/golden/distill_trainer_tiny.npz` was produced by the UNMODIFIED `loss_trainer.ImageDistilTrainer.forward`
(oracle/make_golden_distill.py).  Here:
  * CPU (`not gpu`): `odevit_b200.loss_trainer.ImageDistilTrainer` driving the ORACLE student and the HF teacher
    reproduces every loss term and the post-clip gradients -- pins the trainer restatement itself;
  * GPU: the same trainer driving `odevit_b200.ViTNeuralODE` (libodevit) + the HF teacher / `ViTTeacher`
    reproduces them through the CUDA path (fp32 mode: loss <= 1e-4, gradients <= 2e-3; bf16 mode: loss <= 2e-2, gradients
    bounded by the fixture's own conditioning -- see the test).
"""
import json

import pytest
import torch
from torch import nn

import odevit_oracle as orc
from _util import Golden, cosine, max_rel

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


def _check(out, student_named_grads, g: Golden, tag, tol_loss, tol_grad, cos_min=None, tol_term=None):
    """Every scalar loss term and every (post-clip) parameter gradient against the golden; returns a report.
    `tol_term`: per-term overrides -- JaSMin is ill-conditioned on near-one-hot rows (g_1 = x_1 (1 - x_1 + x_2)
    cancels: a 1e-7 change of x_1 = 1 - 1e-4 is a 1e-3 change of g_1), and the L1 attention-mass term sits behind a
    slope-40 sigmoid on the cumulative mass of the exported map (bf16 mode: the map itself is held to 5e-2)."""
    tol_term = tol_term or {}
    report = {"terms": {}, "grads": {}}
    bad = []
    for k in SCALARS:
        want, got = float(g.get(f"{tag}/{k}")), float(out[k].detach())
        tol = tol_term.get(k, tol_loss)
        report["terms"][k] = abs(got - want) / max(abs(want), 1e-30)
        if not got == pytest.approx(want, rel=tol, abs=tol * 1e-2):
            bad.append((tag, k, got, want))
    for k, grad in student_named_grads:
        want = g.get(f"{tag}/grad/{k}")
        if float(want.abs().max()) == 0.0:
            continue
        assert grad is not None, k
        rel, cos = max_rel(grad, want), (cosine(grad, want) if want.numel() > 1 else 1.0)
        report["grads"][k] = [round(rel, 6), round(cos, 7)]
        if rel >= tol_grad or (cos_min is not None and cos <= cos_min):
            bad.append((tag, k, rel, cos))
    report["bad"] = bad
    return report


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
    rep = _check(out, [(n, p.grad) for n, p in zip(student.names, student.ps)], g, tag, 2e-5, 2e-4)
    assert not rep["bad"], rep["bad"]
    assert max_rel(out["student_output"]["control_points"], g.get(f"{tag}/control_points")) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("teacher_kind", ["hf", "library"])
@pytest.mark.parametrize("precision,tol_loss,tol_grad", [("fp32", 1e-4, 2e-3), ("bf16", 2e-2, 0.2)])
def test_trainer_gpu_vs_reference_trainer(precision, tol_loss, tol_grad, teacher_kind):
    """bf16 gradient bound: this fixture (random-init D=128 student against a random teacher, the L1 term behind a
    slope-40 sigmoid) is badly conditioned -- rounding ONLY the field's four weight matrices to bf16 in the fp32 CPU
    oracle already moves its gradients by 4.5e-2 (cosine 0.99916); the bf16 mode also rounds the operands of every GEMM
    of 23 evaluations forward and backward.  Measured: max-rel 0.12-0.17, cosine >= 0.989.  The bf16 VJP kernels
    themselves are held to 3e-2 at this shape (test_gpu_parity.py::test_field_bf16_small_shapes_vs_oracle) and the
    training gradients at bench depth to 3e-2 / cosine 0.999 (test_gpu_depth.py)."""
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
    fp32 = precision == "fp32"
    tol_term = {"jasmin_loss": 2e-3} if fp32 else {"jasmin_loss": 2e-2, "kl_loss": 8e-2}
    bad = []
    for tag, epoch in (("e0", 0), ("e201", 201)):
        out = tr({"pixel_values": px}, lb, epoch=epoch)
        rep = _check(out, [(n, p.grad) for n, p in student.named_parameters()], g, tag, tol_loss, tol_grad,
                     cos_min=0.999999 if fp32 else 0.985, tol_term=tol_term)
        print(json.dumps({"case": tag, "precision": precision, "teacher": teacher_kind, "loss": float(out["loss"].detach()),
                          "terms": {k: round(v, 6) for k, v in rep["terms"].items() if not k.startswith("mse_loss_t@")},
                          "worst_grad": max(rep["grads"].values()), "min_cos": min(v[1] for v in rep["grads"].values())}))
        bad += rep["bad"]
    assert not bad, bad
    assert out["student_output"]["logits"].argmax(-1).cpu().tolist() == g.get("e201/logits").argmax(-1).tolist()
