"""Training-mode dropout of the parallel field (attention map, after out_proj, after GELU, after fc2;
ode_transformer_gpt.py:56, :61, :196-199, :217-231), re-drawn at every field evaluation.

PyTorch's Philox stream cannot be reproduced, so parity is checked against a
PyTorch composition that applies the SAME masks: the counter-based generator is restated here in
numpy (`_mask`), keyed exactly like csrc/api.cu::make_drop / epilogue.cuh::drop_factor."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import odevit_oracle as orc
from _util import Golden, max_rel

pytestmark = pytest.mark.gpu
M32 = 0xFFFFFFFF
SITE_ATTN, SITE_PROJ, SITE_MLP_H, SITE_MLP_OUT, SITE_MLP_H2, SITE_MLP_OUT2 = 0, 1, 2, 3, 4, 5
N_SITES = 8   # csrc/internal.h::DS_SITES


def _mix(x):
    x = np.asarray(x, dtype=np.uint64) & M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & M32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & M32
    x ^= x >> np.uint64(16)
    return x


def _mask(seed, e, site, p, rows, cols):
    """[rows, cols] float32 multipliers: 0 or 1/(1-p)."""
    if p <= 0:
        return torch.ones(rows, cols)
    lo, hi = seed & M32, (seed >> 32) & M32
    key = _mix(np.uint64(lo) ^ _mix(np.uint64(hi) ^ np.uint64(0x632BE5AB)) ^ np.uint64(((e * N_SITES + site + 1) * 0x27D4EB2F) & M32))
    thresh = max(1, min(M32, int(np.float32(p).astype(np.float64) * 4294967296.0)))
    r = (np.arange(rows, dtype=np.uint64)[:, None] * np.uint64(0x9E3779B1)) & M32
    c = (np.arange(cols, dtype=np.uint64)[None, :] * np.uint64(0x85EBCA77)) & M32
    h = _mix(r ^ c ^ key)
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return torch.from_numpy(np.where(h >= thresh, scale, np.float32(0)).astype(np.float32))


def _field_with_masks(x, sd, heads, scaler, seed, e, drops, prefix="block."):
    """F(x) + G(x) of the parallel block with explicit dropout masks (same sites as the reference)."""
    p_attn, p_proj, p_mlp = drops
    B, N, D = x.shape
    d = D // heads
    na = orc.center_norm(x, sd[prefix + "norm_attn.weight"], sd[prefix + "norm_attn.bias"])
    nm = orc.center_norm(x, sd[prefix + "norm_mlp.weight"], sd[prefix + "norm_mlp.bias"])
    qkv = F.linear(na, sd[prefix + "attn.mha.in_proj_weight"]).view(B, N, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * d ** -0.5, qkv[1], qkv[2]
    P = torch.softmax(q @ k.transpose(-1, -2), -1)
    P = P * _mask(seed, e, SITE_ATTN, p_attn, B * heads * N, N).view(B, heads, N, N)
    o = (P @ v).transpose(1, 2).reshape(B, N, D)
    attn = F.linear(o, sd[prefix + "attn.mha.out_proj.weight"]) * _mask(seed, e, SITE_PROJ, p_proj, B * N, D).view(B, N, D)
    hid = sd[prefix + "mlp.fc1.weight"].shape[0]
    h = F.gelu(F.linear(nm, sd[prefix + "mlp.fc1.weight"])) * _mask(seed, e, SITE_MLP_H, p_mlp, B * N, hid).view(B, N, hid)
    mlp = F.linear(h, sd[prefix + "mlp.fc2.weight"]) * _mask(seed, e, SITE_MLP_OUT, p_mlp, B * N, D).view(B, N, D)
    return (mlp + attn) * scaler, P


DROPS = (0.3, 0.2, 0.1)
SEED = (0x1234567 << 32) | 0x89ABCDE


def _module(g, precision, drops=DROPS):
    import odevit_b200 as ob
    f = ob.ViT_ODEFunc(dim=64, num_heads=2, mlp_ratio=2.0, attn_drop=drops[0], proj_drop=drops[1], mlp_drop=drops[2],
                       emulate_depth=12, time_interval=1.0, l2_attention=False)
    f.load_state_dict(g.group("mha/sd"), strict=True)
    f = f.cuda().train()
    f.block.precision = precision
    return f


@pytest.mark.parametrize("drops", [DROPS, (0.25, 0.0, 0.0), (0.0, 0.0, 0.4), (0.0, 0.5, 0.0)])
def test_field_dropout_matches_masked_composition_fp32(monkeypatch, drops):
    from odevit_b200 import ops
    monkeypatch.setattr(ops, "draw_seed", lambda: SEED)
    g = Golden("fields_d64")
    sd = g.group("mha/sd")
    f = _module(g, "fp32", drops)
    x = g.get("mha/x").cuda().requires_grad_(True)
    wp = torch.randn(2, 2, 19, 19, generator=torch.Generator().manual_seed(5))
    dx = f(torch.tensor(0.0), x)
    ((dx * g.get("mha/w").cuda()).sum() + (f.block.attentions * wp.cuda()).sum()).backward()

    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = g.get("mha/x").clone().requires_grad_(True)
    dxr, pr = _field_with_masks(xr, sdr, 2, 12.0, SEED, 0, drops)
    ((dxr * g.get("mha/w")).sum() + (pr * wp).sum()).backward()
    assert max_rel(dx, dxr) < 1e-5
    assert max_rel(f.block.attentions, pr) < 1e-5          # the returned map is post-dropout
    assert max_rel(x.grad, xr.grad) < 1e-4
    for k, p in f.named_parameters():
        assert max_rel(p.grad, sdr[k].grad) < 1e-4, k


@pytest.mark.parametrize("mode", ["tape", "recompute"])
def test_solve_redraws_masks_every_evaluation_fp32(monkeypatch, mode):
    """Midpoint over 3 grid points = 4 evaluations, each with its own masks (key = evaluation index),
    against autograd through the same masked composition."""
    import odevit_b200 as ob
    from odevit_b200 import ops
    monkeypatch.setattr(ops, "draw_seed", lambda: SEED)
    g = Golden("fields_d64")
    sd = g.group("mha/sd")
    f = _module(g, "fp32")
    f.block.backward_mode = mode
    x = g.get("mha/x").cuda().requires_grad_(True)
    t = torch.tensor([0.0, 0.02, 0.05])
    states = ob.odeint(f, x, t, method="midpoint")
    (states[-1] ** 2).mean().backward()

    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = g.get("mha/x").clone().requires_grad_(True)
    counter = [0]

    def fr(y):
        dy, _ = _field_with_masks(y, sdr, 2, 12.0, SEED, counter[0], DROPS)
        counter[0] += 1
        return dy

    sr = orc.odeint_fixed(fr, xr, t, "midpoint")
    (sr[-1] ** 2).mean().backward()
    assert counter[0] == 4
    assert max_rel(states, sr) < 1e-4
    assert max_rel(x.grad, xr.grad) < 2e-3
    for k, p in f.named_parameters():
        assert max_rel(p.grad, sdr[k].grad) < 2e-3, k


def _c10_block(drops, precision):
    import odevit_b200 as ob
    torch.manual_seed(3)
    f = ob.ViT_ODEFunc(dim=192, num_heads=3, mlp_ratio=4.0, attn_drop=drops[0], proj_drop=drops[1], mlp_drop=drops[2],
                       emulate_depth=12, time_interval=1.0, l2_attention=False)
    sd = {k: v.clone() for k, v in f.state_dict().items()}
    f = f.cuda().train()
    f.block.precision = precision
    return f, sd


@pytest.mark.parametrize("N", [69, 207])
def test_field_dropout_bf16_fused_kernels(monkeypatch, N):
    """Head dim 64: tcgen05 GEMM epilogues and the fused attention forward / VJP kernels generate the
    masks themselves; against the masked fp32 composition at bf16 tolerance."""
    from odevit_b200 import ops
    monkeypatch.setattr(ops, "draw_seed", lambda: SEED)
    f, sd = _c10_block(DROPS, "bf16")
    x = torch.randn(3, N, 192, generator=torch.Generator().manual_seed(6)) * 2
    w = torch.randn(3, N, 192, generator=torch.Generator().manual_seed(7))
    xg = x.cuda().requires_grad_(True)
    dx = f(torch.tensor(0.0), xg)
    (dx * w.cuda()).sum().backward()
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    dxr, pr = _field_with_masks(xr, sdr, 3, 12.0, SEED, 0, DROPS)
    (dxr * w).sum().backward()
    assert max_rel(dx, dxr) < 2e-2
    assert max_rel(f.block.attentions, pr) < 5e-2
    kept = (f.block.attentions > 0).float().mean().item()
    assert abs(kept - (1 - DROPS[0])) < 0.02
    assert max_rel(xg.grad, xr.grad) < 4e-2
    for k, p in f.named_parameters():
        assert max_rel(p.grad, sdr[k].grad) < 4e-2, k


def test_training_solve_bf16_tape_equals_recompute_and_is_seeded():
    """Whole model, the reference's training dropout (0.3 everywhere, experiment_vit_edo.yaml:53-55):
    torch.manual_seed makes a step reproducible, different seeds differ, the reverse sweep regenerates
    the masks (tape == recompute), eval() switches dropout off."""
    import odevit_b200 as ob
    cfg = dict(img_size=32, patch_size=4, num_classes=10, embed_dim=192, num_heads=3, mlp_ratio=4.0, attn_drop=0.3,
               proj_drop=0.3, mlp_drop=0.3, emulate_depth=12, time_interval=1.0, num_eval_steps=4, solver="rk4",
               register_tokens=4)
    torch.manual_seed(0)
    model = ob.ViTNeuralODE(**cfg).cuda().train()
    model.precision = "bf16"
    px = torch.randn(4, 3, 32, 32, device="cuda")
    lb = torch.tensor([1, 2, 3, 4], device="cuda")

    def run(seed, mode):
        torch.manual_seed(seed)
        model.odefunc.block.backward_mode = mode
        model.zero_grad(set_to_none=True)
        out = model(px, labels=lb, output_hidden_states=True)
        out["loss"].backward()
        return out["states"][-1].clone(), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}

    s1, g1 = run(11, "tape")
    s2, g2 = run(11, "tape")
    s3, g3 = run(11, "recompute")
    s4, _ = run(12, "tape")
    assert torch.equal(s1, s2)
    assert torch.equal(s1, s3)
    assert not torch.equal(s1, s4)
    for k in g1:
        assert max_rel(g1[k], g3[k]) < 3e-2, k      # bf16 accumulation order only
        assert max_rel(g1[k], g2[k]) < 1e-3, k
    model.eval()
    with torch.no_grad():
        e1 = model(px)["logits"]
        e2 = model(px)["logits"]
    assert torch.equal(e1, e2)


def test_l2_field_dropout_fp32(monkeypatch):
    """The L2 block takes attention-map and projection dropout too (ode_transformer_gpt.py:56, :61)."""
    import odevit_b200 as ob
    from odevit_b200 import ops
    monkeypatch.setattr(ops, "draw_seed", lambda: SEED)
    g = Golden("fields_d64")
    f = ob.ViT_ODEFunc(dim=64, num_heads=2, mlp_ratio=2.0, attn_drop=0.3, proj_drop=0.2, emulate_depth=12,
                       time_interval=1.0, l2_attention=True)
    f.load_state_dict(g.group("l2/sd"), strict=True)
    f = f.cuda().train()
    x = g.get("l2/x").cuda().requires_grad_(True)
    dx = f(torch.tensor(0.0), x)
    (dx * g.get("l2/w").cuda()).sum().backward()
    sd = {k: v.clone().requires_grad_(True) for k, v in g.group("l2/sd").items()}
    xr = g.get("l2/x").clone().requires_grad_(True)
    B, N, D, H, d = 2, 19, 64, 2, 32
    na = orc.center_norm(xr, sd["block.norm_attn.weight"], sd["block.norm_attn.bias"])
    nm = orc.center_norm(xr, sd["block.norm_mlp.weight"], sd["block.norm_mlp.bias"])
    lin = lambda n, z: F.linear(z, sd[f"block.attn.{n}.weight"], sd[f"block.attn.{n}.bias"])
    q, k, v = [lin(n, na).view(B, N, H, d).transpose(1, 2) for n in ("q_proj", "k_proj", "v_proj")]
    dist2 = (q ** 2).sum(-1, keepdim=True) + (k ** 2).sum(-1).unsqueeze(-2) - 2 * q @ k.transpose(-1, -2)
    a = torch.exp(-dist2 * d ** -0.5)
    a = a / (a.sum(-1, keepdim=True) + 1e-8)
    a = a * _mask(SEED, 0, SITE_ATTN, 0.3, B * H * N, N).view(B, H, N, N)
    o = (a @ v).transpose(1, 2).reshape(B, N, D)
    attn = lin("out_proj", o) * _mask(SEED, 0, SITE_PROJ, 0.2, B * N, D).view(B, N, D)
    mlp = F.linear(F.gelu(F.linear(nm, sd["block.mlp.fc1.weight"])), sd["block.mlp.fc2.weight"])
    dxr = (mlp + attn) * 12.0
    (dxr * g.get("l2/w")).sum().backward()
    assert max_rel(dx, dxr) < 1e-5
    assert max_rel(f.block.attentions, a) < 1e-5
    assert max_rel(x.grad, xr.grad) < 1e-4
    for kk, p in f.named_parameters():
        assert max_rel(p.grad, sd[kk].grad) < 1e-4, kk


def test_persistent_attention_forward_with_dropout_equals_per_unit_kernel(monkeypatch):
    """attn_fwd_pp_kernel<DROP=true> (taken when a launch has >= 2 units per SM) against the CTA-per-unit kernel:
    same mask generator, same arithmetic -> bitwise equal trajectories, in training mode with attention dropout."""
    import os
    import odevit_b200 as ob
    from odevit_b200 import ops
    monkeypatch.setattr(ops, "draw_seed", lambda: SEED)
    cfg = dict(img_size=32, patch_size=4, num_classes=10, embed_dim=192, num_heads=3, mlp_ratio=2.0, emulate_depth=12,
               time_interval=1.0, num_eval_steps=3, solver="euler", register_tokens=4, attn_drop=0.25)
    torch.manual_seed(2)
    model = ob.ViTNeuralODE(**cfg).cuda().train()
    model.precision = "bf16"
    px = torch.randn(128, 3, 32, 32, generator=torch.Generator().manual_seed(3)).cuda()

    def run(persist):
        os.environ["ODEVIT_ATTN_PERSIST"] = "1" if persist else "0"
        try:
            with torch.no_grad():
                return model(px, output_hidden_states=True)["states"].clone()
        finally:
            os.environ.pop("ODEVIT_ATTN_PERSIST", None)
    a, b = run(True), run(False)
    assert torch.equal(a, b)
    model.eval()
    with torch.no_grad():
        c = model(px, output_hidden_states=True)["states"]
    assert not torch.equal(a[-1], c[-1])          # the masks did something


# ------------------------------------------------------------------------------------------------------
# device-resident seed (CUDA-graph replay) and the Macaron block
# ------------------------------------------------------------------------------------------------------
def _u64(t) -> int:
    return int(t.item()) & 0xFFFFFFFFFFFFFFFF


def test_device_seed_path_matches_masked_composition_fp32():
    """`block.device_seed = True`: the kernels resolve their mask keys from a seed in device memory
    (odevit_desc.drop_seed_dev, rows.cu::resolve_drop_keys).  Same masks as the host-seeded path for the same 64-bit
    value, and the seed advances by itself from call to call."""
    g = Golden("fields_d64")
    sd = g.group("mha/sd")
    f = _module(g, "fp32")
    f.block.device_seed = True
    t = torch.tensor([0.0, 0.02, 0.05])
    seeds = []
    for _ in range(2):
        x = g.get("mha/x").cuda().requires_grad_(True)
        f.zero_grad(set_to_none=True)
        import odevit_b200 as ob
        states = ob.odeint(f, x, t, method="midpoint")
        (states[-1] ** 2).mean().backward()
        seed = _u64(f.block.last_drop_seed)
        seeds.append(seed)
        sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        xr = g.get("mha/x").clone().requires_grad_(True)
        counter = [0]

        def fr(y):
            dy, _ = _field_with_masks(y, sdr, 2, 12.0, seed, counter[0], DROPS)
            counter[0] += 1
            return dy

        sr = orc.odeint_fixed(fr, xr, t, "midpoint")
        (sr[-1] ** 2).mean().backward()
        assert max_rel(states, sr) < 1e-4
        assert max_rel(x.grad, xr.grad) < 2e-3
        for k, p in f.named_parameters():
            assert max_rel(p.grad, sdr[k].grad) < 2e-3, k
    assert seeds[0] != seeds[1]


def test_graph_replay_draws_new_masks_every_step():
    """GraphedTrainStep with the reference's training dropout (0.3, experiment_vit_edo.yaml:53-55): every replay of
    the captured step must draw new masks.  lr = 0 keeps the weights fixed, so the loss changes from replay to replay
    only through the masks; with dropout 0 the same replays give one loss."""
    import odevit_b200 as ob
    from odevit_b200.graphs import GraphedTrainStep

    def losses(drop):
        cfg = dict(img_size=32, patch_size=4, num_classes=10, embed_dim=192, num_heads=3, mlp_ratio=4.0, attn_drop=drop,
                   proj_drop=drop, mlp_drop=drop, emulate_depth=12, time_interval=1.0, num_eval_steps=4, solver="rk4",
                   register_tokens=4)
        torch.manual_seed(0)
        model = ob.ViTNeuralODE(**cfg).cuda().train()
        model.precision = "bf16"
        opt = torch.optim.AdamW(model.parameters(), lr=0.0, weight_decay=0.0, fused=True, capturable=True)
        px = torch.randn(8, 3, 32, 32, device="cuda")
        lb = torch.arange(8, device="cuda") % 10
        stepper = GraphedTrainStep(model, opt, (px, lb), clip=1.0)
        out = [float(stepper(px, lb).item()) for _ in range(4)]
        grads = [p.grad.clone() for p in model.parameters() if p.grad is not None]
        return out, grads, model

    with_drop, _, model = losses(0.3)
    assert len(set(with_drop)) == 4, with_drop
    st = model.odefunc.block._drop_state.state.cpu()
    assert int(st[1]) >= 4 + 1       # the counter advanced inside every replay (and once in the capture)
    without, _, _ = losses(0.0)
    assert len(set(without)) == 1, without


def _macaron_field_with_masks(x, sd, heads, scaler, seed, e, drops, prefix="block."):
    """macaron.py:106-123 with explicit masks: after GELU and after ffn.3 of BOTH half steps (separate masks), on the
    attention map, after out_proj (macaron.py:58-61, :88-94)."""
    p_attn, p_proj, p_mlp = drops
    B, N, D = x.shape
    d = D // heads
    rs = sd[prefix + "res_scale"]
    hid = sd[prefix + "ffn.0.weight"].shape[0]

    def ln(i, z):
        return F.layer_norm(z, (D,), sd[f"{prefix}norm{i}.weight"], sd[f"{prefix}norm{i}.bias"], 1e-5)

    def ffn(z, site_h, site_out):
        h = F.gelu(F.linear(z, sd[prefix + "ffn.0.weight"], sd[prefix + "ffn.0.bias"]))
        h = h * _mask(seed, e, site_h, p_mlp, B * N, hid).view(B, N, hid)
        o = F.linear(h, sd[prefix + "ffn.3.weight"], sd[prefix + "ffn.3.bias"])
        return o * _mask(seed, e, site_out, p_mlp, B * N, D).view(B, N, D)

    x1 = x + 0.5 * rs * ffn(ln(1, x), SITE_MLP_H, SITE_MLP_OUT)
    qkv = F.linear(ln(2, x1), sd[prefix + "attn.mha.in_proj_weight"], sd[prefix + "attn.mha.in_proj_bias"])
    qkv = qkv.view(B, N, 3, heads, d).permute(2, 0, 3, 1, 4)
    P = torch.softmax((qkv[0] * d ** -0.5) @ qkv[1].transpose(-1, -2), -1)
    P = P * _mask(seed, e, SITE_ATTN, p_attn, B * heads * N, N).view(B, heads, N, N)
    o = (P @ qkv[2]).transpose(1, 2).reshape(B, N, D)
    a = F.linear(o, sd[prefix + "attn.mha.out_proj.weight"], sd[prefix + "attn.mha.out_proj.bias"])
    a = a * _mask(seed, e, SITE_PROJ, p_proj, B * N, D).view(B, N, D)
    x2 = x1 + rs * a
    x3 = x2 + 0.5 * rs * ffn(ln(3, x2), SITE_MLP_H2, SITE_MLP_OUT2)
    return x3 * scaler


@pytest.mark.parametrize("drops", [DROPS, (0.0, 0.0, 0.4), (0.25, 0.5, 0.0)])
def test_macaron_field_dropout_matches_masked_composition_fp32(monkeypatch, drops):
    import odevit_b200 as ob
    from odevit_b200 import ops
    monkeypatch.setattr(ops, "draw_seed", lambda: SEED)
    g = Golden("fields_d64")
    sd = g.group("macaron/sd")
    f = ob.macaron.ViT_ODEFunc(dim=64, num_heads=2, mlp_ratio=2.0, attn_drop=drops[0], proj_drop=drops[1],
                               mlp_drop=drops[2], emulate_depth=12, time_interval=1.0)
    f.load_state_dict(sd, strict=True)
    f = f.cuda().train()
    f.block.precision = "fp32"
    with torch.no_grad():
        f.block.res_scale.fill_(0.7)
    x = g.get("macaron/x").cuda().requires_grad_(True)
    dx = f(torch.tensor(0.0), x)
    (dx * g.get("macaron/w").cuda()).sum().backward()
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    with torch.no_grad():
        sdr["block.res_scale"].fill_(0.7)
    xr = g.get("macaron/x").clone().requires_grad_(True)
    dxr = _macaron_field_with_masks(xr, sdr, 2, 12.0, SEED, 0, drops)
    (dxr * g.get("macaron/w")).sum().backward()
    assert max_rel(dx, dxr) < 1e-5
    assert max_rel(x.grad, xr.grad) < 1e-4
    for k, p in f.named_parameters():
        assert max_rel(p.grad, sdr[k].grad) < 1e-4, k


@pytest.mark.parametrize("mode", ["tape", "recompute"])
def test_macaron_solve_dropout_bf16(monkeypatch, mode):
    """Macaron at the CIFAR shape in bf16 (tcgen05 GEMM epilogues + fused attention generate the masks), Euler over 3
    steps with masks re-drawn per evaluation, against autograd through the masked fp32 composition."""
    import odevit_b200 as ob
    from odevit_b200 import ops
    monkeypatch.setattr(ops, "draw_seed", lambda: SEED)
    torch.manual_seed(4)
    f = ob.macaron.ViT_ODEFunc(dim=192, num_heads=3, mlp_ratio=4.0, attn_drop=0.2, proj_drop=0.1, mlp_drop=0.1,
                               emulate_depth=12, time_interval=1.0)
    with torch.no_grad():
        for q in f.block.ffn.parameters():      # the reference's 1e-3 init would leave the FFN masks untested
            if q.dim() == 2:
                q.normal_(0, 0.05)
    sd = {k: v.clone() for k, v in f.state_dict().items()}
    f = f.cuda().train()
    f.block.precision = "bf16"
    f.block.backward_mode = mode
    x = torch.randn(2, 69, 192, generator=torch.Generator().manual_seed(6))
    t = torch.linspace(0, 0.02, 4)
    xg = x.cuda().requires_grad_(True)
    states = ob.odeint(f, xg, t, method="euler")
    (states[-1] ** 2).mean().backward()
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    counter = [0]

    def fr(y):
        dy = _macaron_field_with_masks(y, sdr, 3, 12.0, SEED, counter[0], (0.2, 0.1, 0.1))
        counter[0] += 1
        return dy

    sr = orc.odeint_fixed(fr, xr, t, "euler")
    (sr[-1] ** 2).mean().backward()
    assert max_rel(states[-1], sr[-1]) < 2e-2
    assert max_rel(xg.grad, xr.grad) < 4e-2
    for k, p in f.named_parameters():
        assert max_rel(p.grad, sdr[k].grad) < 5e-2, k
