import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import odevit_b200 as ob
cfg = dict(img_size=32, patch_size=4, num_classes=10, embed_dim=192, num_heads=3, mlp_ratio=4.0, emulate_depth=12,
           time_interval=1.0, num_eval_steps=int(os.environ.get("T","4")), solver=os.environ.get("SOLVER","euler"), register_tokens=4)
torch.manual_seed(0)
m = ob.ViTNeuralODE(**cfg).cuda().eval(); m.precision = "bf16"
px = torch.randn(512, 3, 32, 32, device="cuda")
with torch.no_grad():
    m(px); torch.cuda.synchronize()
