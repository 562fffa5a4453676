"""One vector-field evaluation + VJP at the bench shape, a few times: the short command to put under
ncu when a single kernel of the field is being studied (tools, not a bench value).
    python tools/field_probe.py [--batch 64] [--tokens 207] [--dim 768] [--heads 12] [--ratio 1.0] [--iters 3]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import odevit_b200 as ob  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--tokens", type=int, default=207)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--heads", type=int, default=12)
ap.add_argument("--ratio", type=float, default=1.0)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
torch.manual_seed(0)
f = ob.ViT_ODEFunc(dim=a.dim, num_heads=a.heads, mlp_ratio=a.ratio, emulate_depth=12, time_interval=1.0,
                   l2_attention=False).cuda()
f.block.precision = a.precision
x = torch.randn(a.batch, a.tokens, a.dim, device="cuda", requires_grad=True)
t = torch.linspace(0, 1, 3)
for i in range(a.iters):
    s = ob.odeint(f, x, t, method="euler", record_attention=False)
    s[-1].square().mean().backward()
torch.cuda.synchronize()
print("ok", float(s[-1].abs().max()))
