"""Parity AT BENCH DEPTH: the configurations `bench.py` and the sweeps actually time,
not shortened stand-ins.

  * the C100 bench model (224 px, D 768, H 12, r 1, R 10 -> N = 207), Euler over T = 24 grid points, bf16 mode,
    TRAINING mode: final state, logits, loss and every parameter gradient against the oracle's autograd;
  * the S7M distillation student (r = 4), Euler T = 36, control points + attentions + in-kernel JaSMin (k = 2)
    against the oracle's `jasmin_loss` / `control_points` (the whole 30-map window);
  * the on-chip-state solver against the oracle at 16 / 32 / 64 steps, Euler and RK4 (BASELINE config 5);
  * top-1 on the fixed 64-image batch: the agreement count is printed and asserted.

Tolerances are north_star's: bf16 max-rel <= 2e-2 on final state and logits; gradients are held to max-rel
<= 3e-2 AND cosine >= 0.999 per parameter tensor (what a single field evaluation achieves)."""
import json

import pytest
import torch

import odevit_oracle as orc
from _util import cosine, max_rel

pytestmark = pytest.mark.gpu

BF16_TOL = 2e-2
GRAD_REL, GRAD_COS = 3e-2, 0.999

C100 = dict(img_size=224, patch_size=16, num_classes=100, embed_dim=768, num_heads=12, mlp_ratio=1.0,
            emulate_depth=12, time_interval=1.0, num_eval_steps=24, solver="euler", register_tokens=10)
C10 = dict(img_size=32, patch_size=4, num_classes=10, embed_dim=192, num_heads=3, mlp_ratio=4.0, emulate_depth=12,
           time_interval=1.0, num_eval_steps=5, solver="rk4", register_tokens=4)


def _pair(cfg, seed, train=True, precision="bf16"):
    import odevit_b200 as ob
    sd = orc.reference_like_init(cfg, cfg["num_classes"], seed=seed)
    model = ob.ViTNeuralODE(**cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda()
    model = model.train() if train else model.eval()
    model.precision = precision
    return model, sd


def _grad_report(model, sdr):
    rows = {}
    for k, p in model.named_parameters():
        ref = sdr[k].grad
        if ref is None or float(ref.abs().max()) == 0.0:
            continue
        assert p.grad is not None, k
        rows[k] = (max_rel(p.grad, ref), cosine(p.grad, ref))
    return rows


@pytest.mark.parametrize("backward_mode", ["tape", "recompute"])
def test_c100_bench_config_bf16_train(backward_mode):
    """Exactly what bench.py times (config 2), at B = 2: Euler T = 24, bf16 mode, training mode, CE loss."""
    model, sd = _pair(C100, seed=3)
    model.odefunc.block.backward_mode = backward_mode
    px = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(1234))
    lb = torch.tensor([3, 77])
    out = model(px.cuda(), labels=lb.cuda(), output_hidden_states=True)
    out["loss"].backward()
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    want = orc.vit_ode_forward(sdr, C100, px, labels=lb, output_hidden_states=True)
    want["loss"].backward()
    e_state = max_rel(out["states"][-1], want["states"][-1])
    e_traj = max_rel(out["states"], want["states"])
    e_logit = max_rel(out["logits"], want["logits"])
    rows = _grad_report(model, sdr)
    print(json.dumps({"case": "c100_T24_train_" + backward_mode, "final_state": e_state, "trajectory": e_traj, "logits": e_logit,
                      "loss": [float(out["loss"]), float(want["loss"])],
                      "grads": {k: [round(a, 5), round(b, 6)] for k, (a, b) in rows.items()}}))
    assert e_state < BF16_TOL and e_traj < BF16_TOL and e_logit < BF16_TOL
    assert float(out["loss"]) == pytest.approx(float(want["loss"]), rel=BF16_TOL)
    assert out["logits"].argmax(-1).cpu().tolist() == want["logits"].argmax(-1).tolist()
    for k, (rel, cos) in rows.items():
        assert rel < GRAD_REL and cos > GRAD_COS, (k, rel, cos)


def test_s7m_distill_config_bf16():
    """BASELINE config 3's student: r = 4, Euler T = 36, control points + attentions + JaSMin k = 2 (formed inside the
    attention kernel over the 30-evaluation window) against the oracle's sort-based `jasmin_loss` on its own maps."""
    cfg = dict(C100, mlp_ratio=4.0, num_eval_steps=36)
    model, sd = _pair(cfg, seed=5)
    with torch.no_grad():   # random-init attention is near-uniform (statistic ~ 0): sharpen it
        model.odefunc.block.attn.mha.in_proj_weight[: 2 * 768].mul_(2.0)
        sd["odefunc.block.attn.mha.in_proj_weight"][: 2 * 768].mul_(2.0)
    px = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(1234))
    lb = torch.tensor([5, 42])
    kw = dict(output_hidden_states=True, output_control_points=True, output_attentions=True, jasmin_k=2)
    out = model(px.cuda(), labels=lb.cuda(), **kw)
    obj = out["loss"] + 1e-3 * (out["control_points"][:, :, 0] ** 2).mean() + out["attentions"][:, :, 0, 1:].square().mean()
    obj.backward()
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    want = orc.vit_ode_forward(sdr, cfg, px, labels=lb, **kw)
    objr = want["loss"] + 1e-3 * (want["control_points"][:, :, 0] ** 2).mean() + want["attentions"][:, :, 0, 1:].square().mean()
    objr.backward()
    rows = _grad_report(model, sdr)
    rep = {"case": "s7m_T36", "final_state": max_rel(out["states"][-1], want["states"][-1]),
           "logits": max_rel(out["logits"], want["logits"]),
           "control_points": max_rel(out["control_points"], want["control_points"]),
           "attentions": max_rel(out["attentions"], want["attentions"]),
           "jasmin": [float(out["jasmin_loss"]), float(want["jasmin_loss"])],
           "grads": {k: [round(a, 5), round(b, 6)] for k, (a, b) in rows.items()}}
    print(json.dumps(rep))
    assert rep["final_state"] < BF16_TOL and rep["logits"] < BF16_TOL and rep["control_points"] < BF16_TOL
    assert rep["attentions"] < 5e-2
    assert out["control_points"].shape == want["control_points"].shape
    assert float(out["jasmin_loss"]) == pytest.approx(float(want["jasmin_loss"]), rel=BF16_TOL, abs=1e-3)
    assert not out["jasmin_loss"].requires_grad
    for k, (rel, cos) in rows.items():
        assert rel < GRAD_REL and cos > GRAD_COS, (k, rel, cos)


@pytest.mark.parametrize("solver", ["euler", "rk4"])
@pytest.mark.parametrize("steps", [16, 32, 64])
def test_resident_solver_at_sweep_depth(solver, steps):
    """BASELINE config 5 (C10 student, Euler / RK4, up to 64 steps over t in [0, 1]) on the on-chip-state solver
    against the oracle: the whole trajectory, bf16 mode."""
    import odevit_b200 as ob
    cfg = dict(C10, solver=solver, num_eval_steps=steps + 1)
    model, sd = _pair(cfg, seed=1, train=False)
    px = torch.randn(4, 3, 32, 32, generator=torch.Generator().manual_seed(1234))
    ob.reset_launch_count()
    with torch.no_grad():
        out = model(px.cuda(), output_hidden_states=True)
        torch.cuda.synchronize()
        n_launch = ob.launch_count()
        want = orc.vit_ode_forward(sd, cfg, px, output_hidden_states=True)
    stages = 4 if solver == "rk4" else 1
    assert n_launch < steps * stages          # not 4 launches per evaluation: the resident kernel ran
    e_traj, e_final, e_logit = (max_rel(out["states"], want["states"]), max_rel(out["states"][-1], want["states"][-1]),
                                max_rel(out["logits"], want["logits"]))
    print(json.dumps({"case": f"resident_{solver}_{steps}", "trajectory": e_traj, "final": e_final, "logits": e_logit}))
    assert e_traj < BF16_TOL and e_final < BF16_TOL and e_logit < BF16_TOL
    assert out["logits"].argmax(-1).cpu().tolist() == want["logits"].argmax(-1).tolist()


@pytest.mark.parametrize("train_mode", [False, True])
@pytest.mark.parametrize("shape", ["c10", "c100"])
def test_top1_identical_fixed_batch(shape, train_mode):
    """north_star: top-1 predictions identical on a fixed synthetic batch, bf16 mode -- 64 images, C10 (RK4 T = 5;
    on-chip-state solver in eval mode, multi-kernel + tape in train mode) and the C100 bench model (Euler T = 24)."""
    cfg = C10 if shape == "c10" else C100
    model, sd = _pair(cfg, seed=1, train=train_mode)
    img = cfg["img_size"]
    px = torch.randn(64, 3, img, img, generator=torch.Generator().manual_seed(1234))
    with torch.set_grad_enabled(train_mode):
        got = torch.cat([model(px[i:i + 16].cuda())["logits"].detach().cpu() for i in range(0, 64, 16)])
    with torch.no_grad():
        want = torch.cat([orc.vit_ode_forward(sd, cfg, px[i:i + 16])["logits"] for i in range(0, 64, 16)])
    agree = got.argmax(-1) == want.argmax(-1)
    top2 = want.topk(2, dim=-1).values
    print(json.dumps({"case": f"top1_{shape}_{'train' if train_mode else 'eval'}", "agree": int(agree.sum()), "of": 64,
                      "logits_max_rel": max_rel(got, want),
                      "smallest_margin_rel": float(((top2[:, 0] - top2[:, 1]) / want.abs().max()).min())}))
    assert max_rel(got, want) < BF16_TOL
    assert int(agree.sum()) == 64
