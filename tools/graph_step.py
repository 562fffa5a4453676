"""C10-shape training step (BASELINE config 1 model; the YAML's batch 64 by default), eager launches vs CUDA-graph
replay (odevit_b200.graphs.GraphedTrainStep): ms per step, img/s, and the agreement of the two after a few steps.
    python tools/graph_step.py [--batch 64] [--T 24] [--solver euler]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import odevit_b200 as ob  # noqa: E402
from odevit_b200.graphs import GraphedTrainStep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--T", type=int, default=24)
    ap.add_argument("--solver", default="euler")
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    cfg = dict(img_size=32, patch_size=4, num_classes=10, embed_dim=192, num_heads=3, mlp_ratio=4.0, emulate_depth=12,
               time_interval=1.0, num_eval_steps=a.T, solver=a.solver, register_tokens=4)

    def build(capturable=False):
        torch.manual_seed(0)
        m = ob.ViTNeuralODE(**cfg).cuda().train()
        m.precision = "bf16"
        params = [p for p in m.parameters() if p.requires_grad]
        return m, params, torch.optim.AdamW(params, lr=1e-4, weight_decay=5e-2, fused=True, capturable=capturable)
    px = torch.randn(a.batch, 3, 32, 32, generator=torch.Generator().manual_seed(1234)).cuda()
    lb = torch.randint(0, 10, (a.batch,), generator=torch.Generator().manual_seed(1235)).cuda()

    def timed(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    m1, p1, o1 = build()

    def eager():
        o1.zero_grad(set_to_none=True)
        loss = m1(px, labels=lb)["loss"]
        loss.backward()
        torch.nn.utils.clip_grad_norm_(p1, 1.0, foreach=True)
        o1.step()
        return loss
    ms_eager = timed(eager, a.steps)
    m2, p2, o2 = build(capturable=True)
    stepper = GraphedTrainStep(m2, o2, (px, lb), clip=1.0)
    ms_graph = timed(lambda: stepper(px, lb), a.steps)
    # same number of optimizer steps on both: 3 + steps eager; 3 (warm-up) + 1 (capture) + 3 + steps graphed
    eager()
    torch.cuda.synchronize()
    w1 = torch.cat([p.detach().flatten() for p in p1])
    w2 = torch.cat([p.detach().flatten() for p in p2])
    print(json.dumps({"workload": f"C10 model train step, {a.solver} T={a.T}, batch {a.batch}, bf16", "ms_eager": ms_eager,
                      "ms_graph_replay": ms_graph, "img_per_s_eager": a.batch / ms_eager * 1e3,
                      "img_per_s_graph": a.batch / ms_graph * 1e3,
                      "weights_max_abs_diff_after_equal_steps": float((w1 - w2).abs().max()),
                      "weights_max_abs": float(w1.abs().max())}))


if __name__ == "__main__":
    main()
