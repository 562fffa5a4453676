"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python oracle/make_golden.py            # writes tests/golden/<case>.npz

Each fixture holds, for one seeded case: the reference model's `state_dict` (`sd/<key>`), the
inputs (`in/...`), the reference's forward outputs (`out/...`), the gradients of a scalar
objective that touches every cotangent-injection point of the hot path (`grad/<key>`, plus
`grad/pixel_values`), fp64 re-runs of the final state and logits (`out64/...`, the error
budget of the fp32 reference itself) and a JSON `meta` string (ctor kwargs, call flags).

The reference modules are imported from where they lie through `oracle/ref_import.py`; the
only non-reference arithmetic on the path is the fixed-grid solver restatement in
`oracle/shims/torchdiffeq` (third-party, un-vendored -- parity unpinned there, see its header).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_import import import_reference  # noqa: E402

OUT_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _jitter_(model, seed):
    """The reference initialises every norm to (1, 0) and every bias to 0, which would leave
    those code paths untested; jitter them deterministically (recorded in the fixture)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "norm" in name and name.endswith("weight"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("bias") or name.endswith("res_scale"):
                p.add_(0.05 * torch.randn(p.shape, generator=g))


def _objective(out, attn_w, ctrl_w):
    """CE loss + a control-point CLS-row term (the MSE consumer, loss_trainer.py:135-158) + an
    attention CLS-row term (the L1 consumer, loss_trainer.py:169-172) + jasmin (no grad)."""
    obj = out["loss"]
    if "control_points" in out:
        obj = obj + ctrl_w * (out["control_points"][:, :, 0] ** 2).mean()
    if "attentions" in out:
        a = out["attentions"][:, :, 0, 1:]
        obj = obj + (a * attn_w[: a.shape[-1]]).sum(-1).mean()
    if "jasmin_loss" in out:
        obj = obj + out["jasmin_loss"]
    return obj


def _flatten_out(out, store):
    for k, v in out.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                store[f"out/{k}.{kk}"] = np.asarray(vv.detach().numpy() if torch.is_tensor(vv) else vv)
        else:
            store[f"out/{k}"] = v.detach().numpy()


def run_vit_case(mods, name, ctor, call, B, img, seed_model=0, jitter=True, ctrl_w=1e-3):
    ode = mods["ode"]
    torch.manual_seed(seed_model)
    model = ode.ViTNeuralODE(**ctor)
    if jitter:
        _jitter_(model, 99)
    model.train()  # dropout p=0 in every golden case
    C = ctor.get("in_chans", 3)
    px = torch.randn(B, C, img, img, generator=torch.Generator().manual_seed(1234))
    labels = torch.randint(0, ctor["num_classes"], (B,), generator=torch.Generator().manual_seed(1235))
    attn_w = torch.randn(4096, generator=torch.Generator().manual_seed(1236))
    px.requires_grad_(True)
    out = model(px, labels=labels, **call)
    obj = _objective(out, attn_w, ctrl_w)
    obj.backward()
    store = {"in/pixel_values": px.detach().numpy(), "in/labels": labels.numpy(),
             "in/attn_w": attn_w.numpy(), "out/objective": obj.detach().numpy(),
             "grad/pixel_values": px.grad.numpy()}
    for k, v in model.state_dict().items():
        store[f"sd/{k}"] = v.numpy()
    for k, p in model.named_parameters():
        store[f"grad/{k}"] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    _flatten_out(out, store)
    if call.get("output_control_points"):
        T = len(call["t_grid"]) if call.get("t_grid") is not None else ctor.get("num_eval_steps", 24)
        store["out/control_point_indices"] = model.get_proportional_control_points_with_temperature(
            temperature=call.get("temperature", 30), num_eval_steps=T).numpy()
    # fp64 rerun: how far the fp32 reference is from exact arithmetic
    m64 = ode.ViTNeuralODE(**ctor).double()
    m64.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    m64.t_grid = torch.linspace(0.0, ctor.get("time_interval", 12.0), ctor.get("num_eval_steps", 24)).double()
    call64 = dict(call)
    if call64.get("t_grid") is not None:
        call64["t_grid"] = call64["t_grid"].double()
    with torch.no_grad():
        o64 = m64(px.detach().double(), labels=labels, output_hidden_states=True,
                  **{k: v for k, v in call64.items() if k != "output_hidden_states"})
    store["out64/logits"] = o64["logits"].numpy()
    store["out64/final"] = o64["states"][-1].numpy()
    meta = {"kind": "vit", "ctor": ctor, "B": B, "img": img, "ctrl_w": ctrl_w,
            "call": {k: (v.tolist() if torch.is_tensor(v) else v) for k, v in call.items()}}
    store["meta"] = np.asarray(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT_DIR, name + ".npz"), **store)
    print(f"{name}: objective={float(obj):.6f} |final|max={float(out['states'][-1].abs().max()) if 'states' in out else -1:.3f}")


def run_macaron_case(mods, name, ctor, call, B, img):
    mac = mods["macaron"]
    torch.manual_seed(0)
    model = mac.ViTMacaron(**ctor)
    _jitter_(model, 98)
    with torch.no_grad():  # ffn is initialised at std 1e-3; make it matter
        g = torch.Generator().manual_seed(97)
        for n, p in model.named_parameters():
            if "ffn" in n and n.endswith("weight"):
                p.add_(0.05 * torch.randn(p.shape, generator=g))
    px = torch.randn(B, 3, img, img, generator=torch.Generator().manual_seed(1234))
    labels = torch.randint(0, ctor["num_classes"], (B,), generator=torch.Generator().manual_seed(1235))
    px.requires_grad_(True)
    out = model(px, labels=labels, **call)
    obj = out["loss"]
    if "control_points" in out:
        obj = obj + 1e-3 * (out["control_points"][:, :, 0] ** 2).mean()
    obj.backward()
    store = {"in/pixel_values": px.detach().numpy(), "in/labels": labels.numpy(),
             "out/objective": obj.detach().numpy(), "grad/pixel_values": px.grad.numpy()}
    for k, v in model.state_dict().items():
        store[f"sd/{k}"] = v.numpy()
    for k, p in model.named_parameters():
        store[f"grad/{k}"] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    _flatten_out(out, store)
    if call.get("output_control_points"):
        store["out/control_point_indices"] = model.get_proportional_control_points_with_temperature(
            temperature=call.get("temperature", 100.0), num_eval_steps=ctor["num_eval_steps"]).numpy()
    meta = {"kind": "macaron", "ctor": ctor, "B": B, "img": img, "call": call}
    store["meta"] = np.asarray(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT_DIR, name + ".npz"), **store)
    print(f"{name}: objective={float(obj):.6f}")


def run_field_cases(mods):
    """Direct f(t, x) calls: the MHA field, the L2 field (unreachable through the model's
    forward -- SURVEY 2.3 quirk 13) and the Macaron field."""
    ode, mac = mods["ode"], mods["macaron"]
    store = {}
    x = torch.randn(2, 19, 64, generator=torch.Generator().manual_seed(7))
    t = torch.tensor(0.25)
    for tag, l2 in (("mha", False), ("l2", True)):
        torch.manual_seed(3)
        f = ode.ViT_ODEFunc(dim=64, num_heads=2, mlp_ratio=2.0, emulate_depth=12,
                            time_interval=1.0, l2_attention=l2)
        _jitter_(f, 96)
        xx = x.clone().requires_grad_(True)
        dx = f(t, xx)
        w = torch.randn(dx.shape, generator=torch.Generator().manual_seed(8))
        (dx * w).sum().backward()
        store[f"{tag}/x"] = x.numpy()
        store[f"{tag}/w"] = w.numpy()
        store[f"{tag}/dx"] = dx.detach().numpy()
        store[f"{tag}/P"] = f.block.attentions.detach().numpy()
        store[f"{tag}/grad_x"] = xx.grad.numpy()
        for k, v in f.state_dict().items():
            store[f"{tag}/sd/{k}"] = v.numpy()
        for k, p in f.named_parameters():
            store[f"{tag}/grad/{k}"] = p.grad.numpy()
    torch.manual_seed(4)
    f = mac.ViT_ODEFunc(dim=64, num_heads=2, mlp_ratio=2.0, emulate_depth=12, time_interval=1.0)
    _jitter_(f, 95)
    with torch.no_grad():
        g = torch.Generator().manual_seed(94)
        for n, p in f.named_parameters():
            if "ffn" in n and n.endswith("weight"):
                p.add_(0.05 * torch.randn(p.shape, generator=g))
    xx = x.clone().requires_grad_(True)
    dx = f(t, xx)
    w = torch.randn(dx.shape, generator=torch.Generator().manual_seed(8))
    (dx * w).sum().backward()
    store["macaron/x"] = x.numpy()
    store["macaron/w"] = w.numpy()
    store["macaron/dx"] = dx.detach().numpy()
    store["macaron/grad_x"] = xx.grad.numpy()
    for k, v in f.state_dict().items():
        store[f"macaron/sd/{k}"] = v.numpy()
    for k, p in f.named_parameters():
        store[f"macaron/grad/{k}"] = p.grad.numpy()
    store["meta"] = np.asarray(json.dumps({"kind": "fields", "dim": 64, "num_heads": 2,
                                           "mlp_ratio": 2.0, "scaler": 12.0}))
    np.savez_compressed(os.path.join(OUT_DIR, "fields_d64.npz"), **store)
    print("fields_d64 done")


def run_time_emb_case(mods):
    """models/time_emb.py is dead code in the reference and only partly runnable:
    `TimeEmbedding(learnable_sinusoidal=False)` crashes on a shape mismatch (lin1 expects
    2*sd+1 features, SinusoidalPosEmb yields sd+1, :92 vs :20-39) and the learnable path stops in
    a stray `pdb.set_trace()` (:66).  We record what does run: SinusoidalPosEmb, ScaleShift, and
    the learnable TimeEmbedding with `pdb.set_trace` patched to a no-op (module code unmodified)."""
    import pdb
    te = mods["time_emb"]
    torch.manual_seed(5)
    emb = te.TimeEmbedding(sinusoidal_dim=16, embed_dim=64, multiplier=2, dropout=0.1,
                           learnable_sinusoidal=True).eval()
    ss = te.ScaleShift(embed_dim=64, out_dim=64)
    t = torch.linspace(0, 1, 7)
    real_set_trace = pdb.set_trace
    pdb.set_trace = lambda *a, **k: None
    try:
        with torch.no_grad():
            e = emb(t)
            scale, shift = ss(e)
            four = te.SinusoidalPosEmb(16)(t)
    finally:
        pdb.set_trace = real_set_trace
    store = {"t": t.numpy(), "fourier": four.numpy(), "emb": e.numpy(),
             "scale": scale.numpy(), "shift": shift.numpy()}
    for k, v in emb.state_dict().items():
        store[f"sd_emb/{k}"] = v.numpy()
    for k, v in ss.state_dict().items():
        store[f"sd_ss/{k}"] = v.numpy()
    store["meta"] = np.asarray(json.dumps({"kind": "time_emb", "sinusoidal_dim": 16,
                                           "embed_dim": 64, "multiplier": 2,
                                           "learnable_sinusoidal": True}))
    np.savez_compressed(os.path.join(OUT_DIR, "time_emb.npz"), **store)
    print("time_emb done")


def main():
    os.makedirs(OUT_DIR, exist_ok=True)
    torch.set_num_threads(4)
    mods = import_reference()
    full = dict(output_hidden_states=True, output_control_points=True, output_attentions=True)
    # BASELINE.json configs[0]: CIFAR-10 model (0.5 M params), 32 px patch 4, RK4, fwd+bwd
    c10 = dict(img_size=32, patch_size=4, num_classes=10, embed_dim=192, num_heads=3,
               mlp_ratio=4.0, emulate_depth=12, time_interval=1.0, num_eval_steps=5,
               solver="rk4", register_tokens=4)
    run_vit_case(mods, "c10_rk4_T5_B2", c10, full, B=2, img=32)
    run_vit_case(mods, "c10_euler_T13_B2", dict(c10, solver="euler", num_eval_steps=13), full, B=2, img=32)
    tiny = dict(img_size=16, patch_size=4, num_classes=7, embed_dim=64, num_heads=2,
                mlp_ratio=2.0, emulate_depth=12, time_interval=1.0, num_eval_steps=6,
                solver="euler", register_tokens=2, pos_embed_register_tokens=True)
    run_vit_case(mods, "tiny_euler_T6_B3", tiny, dict(full, jasmin_k=2), B=3, img=16)
    run_vit_case(mods, "tiny_rk4_T4_B3", dict(tiny, solver="rk4", num_eval_steps=4,
                                               pos_embed_register_tokens=False),
                 dict(full, temperature=10, jasmin_k=3), B=3, img=16)
    run_vit_case(mods, "tiny_midpoint_T5_B2", dict(tiny, solver="midpoint", num_eval_steps=5,
                                                    time_interval=12.0),
                 dict(full), B=2, img=16)
    # non-uniform user grid (dt differs per step)
    tg = torch.tensor([0.0, 0.1, 0.25, 0.6, 1.0])
    run_vit_case(mods, "tiny_rk4_tgrid_B2", dict(tiny, solver="rk4"),
                 dict(output_hidden_states=True, output_attentions=True, t_grid=tg), B=2, img=16)
    # (the reference crashes at :172 with a distillation token AND pos_embed_register_tokens=True)
    run_vit_case(mods, "tiny_dist_token_B2", dict(tiny, add_distillation_token=True,
                                                   pos_embed_register_tokens=False),
                 dict(output_hidden_states=True), B=2, img=16)
    mac = dict(img_size=16, patch_size=4, num_classes=7, embed_dim=64, num_heads=2,
               mlp_ratio=2.0, emulate_depth=12, time_interval=12.0, num_eval_steps=4, solver="rk4")
    run_macaron_case(mods, "macaron_rk4_T4_B2", mac, dict(output_hidden_states=True), B=2, img=16)
    run_macaron_case(mods, "macaron_euler_T13_B2", dict(mac, solver="euler", num_eval_steps=13,
                                                         time_interval=1.0, emulate_depth=1),
                     dict(output_hidden_states=True, output_control_points=True, temperature=30), B=2, img=16)
    run_field_cases(mods)
    run_time_emb_case(mods)


if __name__ == "__main__":
    main()
