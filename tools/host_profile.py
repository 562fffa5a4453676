"""cProfile of the host side of the bench step (where do the CPU milliseconds of a step go)."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import odevit_b200 as ob  # noqa: E402

print("cpus:", os.cpu_count(), "affinity:", len(os.sched_getaffinity(0)), "torch threads:", torch.get_num_threads())
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


for _ in range(4):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
