import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import odevit_b200 as ob
cfg = dict(img_size=32, patch_size=4, num_classes=10, embed_dim=192, num_heads=3, mlp_ratio=4.0, emulate_depth=12,
           time_interval=1.0, num_eval_steps=33, solver="euler", register_tokens=4)
torch.manual_seed(0)
m = ob.ViTNeuralODE(**cfg).cuda().eval(); m.precision = "bf16"
for B in (1, 8, 16, 37, 74, 111, 148, 296, 592):
    px = torch.randn(B, 3, 32, 32, device="cuda")
    with torch.no_grad():
        for _ in range(2): m(px)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): m(px)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    rounds = (B + 147) // 148
    print(f"B={B:4d} {ms:.3f} ms  per eval per round {ms*1e3/32/rounds:.2f} us = {ms*1e3/32/rounds*1.965:.1f} kcycles")
