"""Forward error of the bf16 mode against the fp32 CPU oracle on the distillation fixture's student, with the GEMM operand
written by center_rows (default) or by the previous stage-combine epilogue (ODEVIT_OPERAND_FROM_EPILOGUE=1).  A probe."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import odevit_b200 as ob
import odevit_oracle as orc
from _util import Golden, max_rel

g = Golden("distill_trainer_tiny")
sd, cfg = g.group("sd"), g.meta["student"]
px = g.get("in/pixel_values")
with torch.no_grad():
    want = orc.vit_ode_forward(sd, cfg, px, output_hidden_states=True)
m = ob.ViTNeuralODE(**cfg)
m.load_state_dict(sd, strict=True)
m = m.cuda().train()            # training mode: the multi-kernel path with its tape (eval takes the on-chip-state solver)
m.precision = "bf16"
got = m(px.cuda(), output_hidden_states=True)
ws, gs = want["states"], got["states"].detach().cpu()
print(json.dumps({"env": os.environ.get("ODEVIT_OPERAND_FROM_EPILOGUE", "0"), "final_state": max_rel(gs[-1], ws[-1]), "state_5": max_rel(gs[5], ws[5]),
                  "state_12": max_rel(gs[12], ws[12]), "logits": max_rel(got["logits"].detach().cpu(), want["logits"])}))
