"""Per-kernel-class device times (CUDA event pairs inside libodevit, `odevit_profile_*`) of one field evaluation +
VJP at the bench shape: the quick measurement used while a single kernel is being reworked (a tool, not a bench value).
    python tools/kernel_times.py [--batch 64] [--tokens 207] [--dim 768] [--heads 12] [--ratio 1.0] [--iters 10]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import odevit_b200 as ob  # noqa: E402
from odevit_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--tokens", type=int, default=207)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--heads", type=int, default=12)
ap.add_argument("--ratio", type=float, default=1.0)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--steps", type=int, default=3, help="Euler steps per solve")
ap.add_argument("--precision", default="bf16")
ap.add_argument("--drop", type=float, default=0.0, help="attention / projection / MLP dropout (training mode)")
a = ap.parse_args()
torch.manual_seed(0)
f = ob.ViT_ODEFunc(dim=a.dim, num_heads=a.heads, mlp_ratio=a.ratio, emulate_depth=12, time_interval=1.0,
                   l2_attention=False, attn_drop=a.drop, proj_drop=a.drop, mlp_drop=a.drop).cuda().train()
f.block.precision = a.precision
x = torch.randn(a.batch, a.tokens, a.dim, device="cuda", requires_grad=True)
t = torch.linspace(0, 1, a.steps + 1)


def run():
    s = ob.odeint(f, x, t, method="euler", record_attention=False)
    s[-1].square().mean().backward()
    return s


for _ in range(3):
    run()
torch.cuda.synchronize()
_lib.profile_reserve(4096)
_lib.profile_enable(True)
for _ in range(a.iters):
    s = run()
torch.cuda.synchronize()
prof = _lib.profile_read()
_lib.profile_enable(False)
B, N, D = a.batch, a.tokens, a.dim
att = 2.0 * B * N * N * D
flops = {"fused_attn": 2 * att, "fused_attn_bwd": 4 * att}
out = {k: {"avg_us": round(ms / n * 1e3, 2), "launches": n} for k, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0])}
for k, fl in flops.items():
    if k in out:
        out[k]["tflops_algorithmic"] = round(fl / (out[k]["avg_us"] * 1e-6) / 1e12, 1)
print(json.dumps({"shape": [B, N, D, a.heads], "finite": bool(torch.isfinite(s[-1]).all()), "classes": out}))
