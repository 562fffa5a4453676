"""Per-step wall / device times of the bench step and of an inference call (diagnosis tool: looks for
allocator or host stalls that a mean hides).  python tools/step_times.py [--steps 12]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import odevit_b200 as ob  # noqa: E402
from odevit_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--profile", type=int, default=0)
a = ap.parse_args()
wl = bench.WORKLOADS["c100"]
cfg, B = wl["cfg"], wl["batch"]
torch.manual_seed(0)
model = ob.ViTNeuralODE(**cfg).cuda().train()
model.precision = "bf16"
params = [p for p in model.parameters() if p.requires_grad]
opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=5e-2, fused=True)
px, lb = bench.synthetic_batch(cfg, B)
px, lb = px.cuda(), lb.cuda()


def step():
    opt.zero_grad(set_to_none=True)
    out = model(px, labels=lb)
    out["loss"].backward()
    torch.nn.utils.clip_grad_norm_(params, 1.0, foreach=True)
    opt.step()


if a.profile:
    _lib.profile_enable(True)
rows = []
for i in range(a.steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    step()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    rows.append((1e3 * (t1 - t0), 1e3 * (t2 - t0), e0.elapsed_time(e1)))
print("train  host-enqueue ms / wall ms / device ms per step")
for r in rows:
    print("  %.2f  %.2f  %.2f" % r)
print("reserved GB %.2f allocated GB %.2f" % (torch.cuda.memory_reserved() / 1e9, torch.cuda.memory_allocated() / 1e9))
model.eval()
rows = []
with torch.no_grad():
    for i in range(a.steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model(px)
        torch.cuda.synchronize()
        rows.append(1e3 * (time.perf_counter() - t0))
print("inference wall ms per call:", " ".join("%.2f" % r for r in rows))
