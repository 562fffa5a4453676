"""GPU parity of the other two vector fields of the reference -- the L2-attention parallel block
(ode_transformer_gpt.py:12-63, :259-262) and the Macaron block (macaron.py:78-150, model :157-352) --
through the module surface -> C ABI, against the reference's golden vectors and the oracle."""
import pytest
import torch

import odevit_oracle as orc
from _util import Golden, MACARON_CASES, max_rel

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 2e-2
GRAD_TOL = 2e-3


# ---- L2 attention -------------------------------------------------------------------------------

def _l2_field(g, precision):
    import odevit_b200 as ob
    f = ob.ViT_ODEFunc(dim=64, num_heads=2, mlp_ratio=2.0, emulate_depth=12, time_interval=1.0, l2_attention=True)
    f.load_state_dict(g.group("l2/sd"), strict=True)
    f = f.cuda()
    f.block.precision = precision
    return f


def test_l2_field_golden_fp32():
    g = Golden("fields_d64")
    f = _l2_field(g, "fp32")
    x = g.get("l2/x").cuda().requires_grad_(True)
    dx = f(torch.tensor(0.25), x)
    assert max_rel(dx, g.get("l2/dx")) < 1e-5
    assert max_rel(f.block.attentions, g.get("l2/P")) < 1e-5
    (dx * g.get("l2/w").cuda()).sum().backward()
    assert max_rel(x.grad, g.get("l2/grad_x")) < 1e-4
    for k, p in f.named_parameters():
        assert max_rel(p.grad, g.get(f"l2/grad/{k}")) < 1e-4, k


def test_l2_field_golden_bf16():
    g = Golden("fields_d64")
    f = _l2_field(g, "bf16")
    x = g.get("l2/x").cuda().requires_grad_(True)
    dx = f(torch.tensor(0.25), x)
    assert max_rel(dx, g.get("l2/dx")) < BF16_TOL
    (dx * g.get("l2/w").cuda()).sum().backward()
    assert max_rel(x.grad, g.get("l2/grad_x")) < 5e-2


def test_l2_attention_cotangent_and_solve_vs_oracle():
    """A loss on the L2 attention map plus a short RK4 solve, against the oracle's autograd."""
    import odevit_b200 as ob
    g = Golden("fields_d64")
    sd = g.group("l2/sd")
    f = _l2_field(g, "fp32")
    x = g.get("l2/x").cuda().requires_grad_(True)
    wp = torch.randn(2, 2, 19, 19, generator=torch.Generator().manual_seed(5))
    t = torch.tensor([0.0, 0.02, 0.05])
    states = ob.odeint(f, x, t, method="rk4")
    ((f.block.attentions * wp.cuda()).sum() + (states[-1] ** 2).mean()).backward()

    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = g.get("l2/x").clone().requires_grad_(True)
    maps = []

    def fr(y):
        dy, p = orc.field_parallel(y, sdr, 2, 12.0, prefix="block.", l2=True)
        maps.append(p)
        return dy

    sr = orc.odeint_fixed(fr, xr, t, "rk4")
    ((maps[-1] * wp).sum() + (sr[-1] ** 2).mean()).backward()
    assert max_rel(states, sr) < FP32_TOL
    assert max_rel(x.grad, xr.grad) < GRAD_TOL
    for k, p in f.named_parameters():
        assert max_rel(p.grad, sdr[k].grad) < GRAD_TOL, k


def test_l2_model_forward_fails_like_the_reference():
    """with l2_attention=True the model's forward dies on `.attn.mha` (:516)."""
    import odevit_b200 as ob
    m = ob.ViTNeuralODE(img_size=16, patch_size=4, num_classes=7, embed_dim=64, num_heads=2, mlp_ratio=2.0,
                        num_eval_steps=4, solver="euler", register_tokens=2, l2_attention=True).cuda()
    with pytest.raises(AttributeError, match="mha"):
        m(torch.randn(1, 3, 16, 16, device="cuda"))


# ---- Macaron -------------------------------------------------------------------------------------

def test_macaron_field_golden_fp32():
    import odevit_b200 as ob
    g = Golden("fields_d64")
    f = ob.macaron.ViT_ODEFunc(dim=64, num_heads=2, mlp_ratio=2.0, emulate_depth=12, time_interval=1.0)
    f.load_state_dict(g.group("macaron/sd"), strict=True)
    f = f.cuda()
    x = g.get("macaron/x").cuda().requires_grad_(True)
    dx = f(torch.tensor(0.25), x)
    assert max_rel(dx, g.get("macaron/dx")) < 1e-5
    (dx * g.get("macaron/w").cuda()).sum().backward()
    assert max_rel(x.grad, g.get("macaron/grad_x")) < 1e-4
    for k, p in f.named_parameters():
        assert max_rel(p.grad, g.get(f"macaron/grad/{k}")) < 1e-4, k


@pytest.mark.parametrize("mode", ["tape", "recompute"])
@pytest.mark.parametrize("name", MACARON_CASES)
def test_macaron_model_golden_fp32(name, mode):
    import odevit_b200 as ob
    g = Golden(name)
    model = ob.ViTMacaron(**g.ctor)
    model.load_state_dict(g.group("sd"), strict=True)
    model = model.cuda().train()
    model.precision = "fp32"
    model.odefunc.block.backward_mode = mode
    px = g.get("in/pixel_values").cuda().requires_grad_(True)
    out = model(px, labels=g.get("in/labels").cuda(), **g.meta["call"])
    want = g.group("out")
    for key in ("logits", "loss", "states", "control_points"):
        if key in want:
            assert out[key].shape == want[key].shape, key
            assert max_rel(out[key], want[key]) < FP32_TOL, key
    assert out["logits"].argmax(-1).cpu().tolist() == want["logits"].argmax(-1).tolist()
    obj = out["loss"]
    if "control_points" in out:
        obj = obj + 1e-3 * (out["control_points"][:, :, 0] ** 2).mean()
    obj.backward()
    grads = g.group("grad")
    assert max_rel(px.grad, grads["pixel_values"]) < GRAD_TOL
    for k, p in model.named_parameters():
        if grads[k].abs().max() > 0:
            assert p.grad is not None, k
            assert max_rel(p.grad, grads[k]) < GRAD_TOL, k


@pytest.mark.parametrize("solver,T", [("euler", 3), ("rk4", 2)])
def test_macaron_c10_shape_bf16_vs_oracle(solver, T):
    """CIFAR shape (D=192, H=3 -> head dim 64: tcgen05 GEMMs + the fused attention kernels), bf16 mode,
    forward and gradients against the oracle.  Few steps: the field returns a state (quirk 14)."""
    import odevit_b200 as ob
    cfg = dict(img_size=32, patch_size=4, num_classes=10, embed_dim=192, num_heads=3, mlp_ratio=4.0,
               emulate_depth=1, time_interval=1.0, num_eval_steps=T, solver=solver)
    torch.manual_seed(11)
    model = ob.ViTMacaron(**cfg)
    with torch.no_grad():
        gen = torch.Generator().manual_seed(12)
        for n, p in model.named_parameters():
            if "ffn" in n and n.endswith("weight"):
                p.add_(0.05 * torch.randn(p.shape, generator=gen))
            if "norm" in n or n.endswith("bias") or "res_scale" in n:
                p.add_(0.1 * torch.randn(p.shape, generator=gen))
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.cuda().train()
    model.precision = "bf16"
    px = torch.randn(3, 3, 32, 32, generator=torch.Generator().manual_seed(1234))
    lb = torch.tensor([1, 7, 3])
    out = model(px.cuda(), labels=lb.cuda(), output_hidden_states=True)
    out["loss"].backward()
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    want = orc.macaron_forward(sdr, cfg, px, labels=lb, output_hidden_states=True)
    want["loss"].backward()
    assert max_rel(out["states"][-1], want["states"][-1]) < BF16_TOL
    assert max_rel(out["logits"], want["logits"]) < BF16_TOL
    for k, p in model.named_parameters():
        if sdr[k].grad is not None and float(sdr[k].grad.abs().max()) > 0:
            assert max_rel(p.grad, sdr[k].grad) < 0.1, k


def test_macaron_has_no_attention_outputs():
    import odevit_b200 as ob
    f = ob.macaron.ViT_ODEFunc(dim=64, num_heads=2, mlp_ratio=2.0).cuda()
    x = torch.randn(1, 5, 64, device="cuda")
    s = ob.odeint(f, x, torch.linspace(0, 0.1, 3), method="euler")
    assert s.shape == (3, 1, 5, 64)
    assert not hasattr(f, "attention_trajectory") or len(f.attention_trajectory) == 0


# ---- time-embedding modulation -----------------------------------------------------------------

def test_time_modulation_matches_composition():
    """ScaleShift vectors folded into the CenterNorm prologue == the PyTorch composition
    n*(1+scale)+shift on the oracle's field (our documented choice)."""
    import odevit_b200 as ob
    g = Golden("fields_d64")
    sd = g.group("mha/sd")
    f = ob.ViT_ODEFunc(dim=64, num_heads=2, mlp_ratio=2.0, emulate_depth=12, time_interval=1.0, l2_attention=False)
    f.load_state_dict(sd, strict=True)
    f = f.cuda()
    torch.manual_seed(3)
    emb = ob.TimeEmbedding(sinusoidal_dim=16, embed_dim=64, multiplier=2, dropout=0.0, learnable_sinusoidal=True).cuda()
    ssa, ssm = ob.ScaleShift(64, 64).cuda(), ob.ScaleShift(64, 64).cuda()
    e = emb(torch.tensor(0.3, device="cuda"))
    ob.attach_time_modulation(f.block, e, ssa, ssm)
    x = g.get("mha/x").cuda().requires_grad_(True)
    dx = f(torch.tensor(0.3), x)
    (dx * g.get("mha/w").cuda()).sum().backward()

    sa, ba = [v.detach().cpu() for v in ssa(e)]
    sm, bm = [v.detach().cpu() for v in ssm(e)]
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = g.get("mha/x").clone().requires_grad_(True)
    na = orc.center_norm(xr, sdr["block.norm_attn.weight"], sdr["block.norm_attn.bias"]) * (1 + sa) + ba
    nm = orc.center_norm(xr, sdr["block.norm_mlp.weight"], sdr["block.norm_mlp.bias"]) * (1 + sm) + bm
    a, _ = orc.mha_explicit(na, sdr["block.attn.mha.in_proj_weight"], sdr["block.attn.mha.out_proj.weight"], 2)
    dxr = (orc.mlp(nm, sdr["block.mlp.fc1.weight"], sdr["block.mlp.fc2.weight"]) + a) * 12.0
    (dxr * g.get("mha/w")).sum().backward()
    assert max_rel(dx, dxr) < 1e-5
    assert max_rel(x.grad, xr.grad) < 1e-4
    for k, p in f.named_parameters():
        assert max_rel(p.grad, sdr[k].grad) < 1e-4, k
    ob.attach_time_modulation(f.block, None)
    assert max_rel(f(torch.tensor(0.3), x.detach()), g.get("mha/dx")) < 1e-5


# ---- patch projection as a GEMM (bf16 mode) ------------------------------------------------------

@pytest.mark.parametrize("img,patch,D,B", [(32, 4, 192, 5), (224, 16, 768, 2), (16, 4, 64, 3)])
def test_patch_project_matches_conv(img, patch, D, B):
    """ops.patch_project (im2col + bf16 tensor-core GEMM) against Conv2d(kernel=stride=patch) in fp32,
    forward and all three gradients, at bf16 tolerance."""
    from odevit_b200 import ops
    torch.manual_seed(1)
    conv = torch.nn.Conv2d(3, D, kernel_size=patch, stride=patch).cuda()
    x = torch.randn(B, 3, img, img, device="cuda", requires_grad=True)
    w = torch.randn(B, (img // patch) ** 2, D, device="cuda")
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = conv(x).flatten(2).transpose(1, 2)
        (ref * w).sum().backward()
    finally:
        torch.backends.cudnn.allow_tf32 = True
    gx, gw, gb = x.grad.clone(), conv.weight.grad.clone(), conv.bias.grad.clone()
    x.grad = None
    conv.zero_grad()
    got = ops.patch_project(x, conv.weight, conv.bias, patch)
    (got * w).sum().backward()
    assert got.shape == ref.shape
    assert max_rel(got, ref) < 1e-2
    assert max_rel(x.grad, gx) < 2e-2
    assert max_rel(conv.weight.grad, gw) < 2e-2
    assert max_rel(conv.bias.grad, gb) < 1e-4
