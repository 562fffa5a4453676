"""The standalone restatement (oracle/odevit_oracle.py) against the golden vectors that the
UNMODIFIED reference produced (oracle/make_golden.py).  This is what pins the oracle."""
import pytest
import torch

import odevit_oracle as orc
from _util import Golden, MACARON_CASES, VIT_CASES, max_rel

TOL = 2e-6  # fp32 reassociation noise between two CPU evaluations of the same graph


def _objective(out, attn_w, ctrl_w):
    obj = out["loss"]
    if "control_points" in out:
        obj = obj + ctrl_w * (out["control_points"][:, :, 0] ** 2).mean()
    if "attentions" in out:
        a = out["attentions"][:, :, 0, 1:]
        obj = obj + (a * attn_w[: a.shape[-1]]).sum(-1).mean()
    if "jasmin_loss" in out:
        obj = obj + out["jasmin_loss"]
    return obj


@pytest.mark.parametrize("name", VIT_CASES)
def test_vit_forward_and_backward(name):
    g = Golden(name)
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in g.group("sd").items()}
    px = g.get("in/pixel_values").clone().requires_grad_(True)
    out = orc.vit_ode_forward(sd, g.ctor, px, labels=g.get("in/labels"), **g.call)
    want = g.group("out")
    for key in ("logits", "loss", "states", "attentions", "attentions_register_tokens",
                "control_points", "second_derivative_upper_bound", "logits_dist"):
        if key in want:
            assert max_rel(out[key], want[key]) < TOL, key
    if "jasmin_loss" in want:
        assert float(out["jasmin_loss"]) == pytest.approx(float(want["jasmin_loss"]), rel=1e-5, abs=1e-6)
    for key in ("global_upper_bound", "batched_upper_bound", "batched_upper_bound_per_seq"):
        got = out["finite_difference_upper_bound"][key]
        assert max_rel(torch.as_tensor(got), want["finite_difference_upper_bound." + key]) < 1e-4, key
    obj = _objective(out, g.get("in/attn_w"), g.meta["ctrl_w"])
    assert float(obj) == pytest.approx(float(want["objective"]), rel=1e-5)
    obj.backward()
    grads = g.group("grad")
    assert max_rel(px.grad, grads["pixel_values"]) < 2e-4
    for k, v in sd.items():
        if k in grads and grads[k].abs().max() > 0:
            gk = v.grad if v.grad is not None else torch.zeros_like(v)
            assert max_rel(gk, grads[k]) < 2e-4, k


@pytest.mark.parametrize("name", VIT_CASES)
def test_fp32_reference_error_budget(name):
    """How far the fp32 reference itself sits from its fp64 re-run: the floor under the
    1e-4 tolerance of BASELINE.json's fp32 mode."""
    g = Golden(name)
    want = g.group("out")
    w64 = g.group("out64")
    assert max_rel(want["states"][-1], w64["final"]) < 1e-4
    assert max_rel(want["logits"], w64["logits"]) < 1e-4


@pytest.mark.parametrize("name", MACARON_CASES)
def test_macaron_forward_and_backward(name):
    g = Golden(name)
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in g.group("sd").items()}
    px = g.get("in/pixel_values").clone().requires_grad_(True)
    out = orc.macaron_forward(sd, g.ctor, px, labels=g.get("in/labels"), **g.meta["call"])
    want = g.group("out")
    for key in ("logits", "loss", "states", "control_points"):
        if key in want:
            assert max_rel(out[key], want[key]) < TOL, key
    obj = out["loss"]
    if "control_points" in out:
        obj = obj + 1e-3 * (out["control_points"][:, :, 0] ** 2).mean()
    obj.backward()
    grads = g.group("grad")
    assert max_rel(px.grad, grads["pixel_values"]) < 2e-4
    for k, v in sd.items():
        if k in grads and grads[k].abs().max() > 0 and v.grad is not None:
            assert max_rel(v.grad, grads[k]) < 2e-4, k


@pytest.mark.parametrize("tag", ["mha", "l2", "macaron"])
def test_field_level(tag):
    g = Golden("fields_d64")
    sd = {k: v.clone().requires_grad_(True) for k, v in g.group(tag + "/sd").items()}
    x = g.get(tag + "/x").clone().requires_grad_(True)
    H, scaler = g.meta["num_heads"], g.meta["scaler"]
    if tag == "macaron":
        dx = orc.field_macaron(x, sd, H, scaler, prefix="block.")
    else:
        dx, p = orc.field_parallel(x, sd, H, scaler, prefix="block.", l2=(tag == "l2"))
        assert max_rel(p, g.get(tag + "/P")) < TOL
    assert max_rel(dx, g.get(tag + "/dx")) < TOL
    (dx * g.get(tag + "/w")).sum().backward()
    assert max_rel(x.grad, g.get(tag + "/grad_x")) < 1e-5
    for k, v in sd.items():
        assert max_rel(v.grad, g.get(f"{tag}/grad/{k}")) < 1e-5, k


def test_time_embedding_pieces():
    g = Golden("time_emb")
    t = g.get("t")
    assert max_rel(orc.sinusoidal_pos_emb(t, 16), g.get("fourier")) < 1e-6
    scale, shift = orc.scale_shift(g.get("emb"), g.group("sd_ss"))
    assert max_rel(scale, g.get("scale")) < 1e-6
    assert max_rel(shift, g.get("shift")) < 1e-6


def test_control_point_indices_match_reference_probe():
    # SURVEY 2.3 quirk 11 (probed on the reference)
    assert orc.control_point_indices(30, 36).tolist() == [3, 6, 8, 10, 12, 14, 16, 18, 20, 22, 24, 35]
    assert orc.control_point_indices(30, 24).tolist() == [2, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 23]
    assert orc.control_point_indices(30, 5).tolist() == [0] * 11 + [4]
    for name in VIT_CASES + MACARON_CASES:
        g = Golden(name)
        if "out/control_point_indices" in g.z.files:
            T = g.ctor["num_eval_steps"]
            if g.meta["kind"] == "vit":
                got = orc.control_point_indices(g.meta["call"].get("temperature", 30), T)
            else:
                got = orc.control_point_indices(g.meta["call"].get("temperature", 100.0), T,
                                                clamp_last=False, distances=orc.MACARON_AVG_DISTANCES)
            assert got.tolist() == g.get("out/control_point_indices").tolist()
