"""On-chip-state solver (csrc/solve_resident.cu) alone: C10-shape inference over a batch sweep, one image
per CTA, so the time per evaluation per round of 148 images shows whether anything outside the SM (L2,
weight streaming) limits it -- it does not: B=1 and B=148 take the same time.

    python tools/resident_probe.py                    # batch sweep, Euler, 32 steps
    SOLVER=rk4 T=17 B=592 python tools/resident_probe.py

Per-job clock trace of CTA 0 (second evaluation of its first image): build the library with
`-DRES_TRACE` on solve_resident.cu, run this with B=512 T=5 and pipe stdout into
tools/resident_trace_view.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import odevit_b200 as ob  # noqa: E402

STAGES = {"euler": 1, "midpoint": 2, "rk4": 4}


def main():
    solver = os.environ.get("SOLVER", "euler")
    T = int(os.environ.get("T", "33"))
    cfg = dict(img_size=32, patch_size=4, num_classes=10, embed_dim=192, num_heads=3, mlp_ratio=4.0, emulate_depth=12,
               time_interval=1.0, num_eval_steps=T, solver=solver, register_tokens=4)
    torch.manual_seed(0)
    m = ob.ViTNeuralODE(**cfg).cuda().eval()
    m.precision = "bf16"
    batches = [int(os.environ["B"])] if "B" in os.environ else [1, 8, 37, 74, 148, 296, 592]
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    clk_ghz = 1.965
    for B in batches:
        px = torch.randn(B, 3, 32, 32, device="cuda")
        with torch.no_grad():
            for _ in range(2):
                m(px)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                m(px)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        rounds = (B + sms - 1) // sms
        evals = (T - 1) * STAGES[solver]
        us = ms * 1e3 / evals / rounds
        print(f"B={B:5d} {solver} T={T}: {ms:.3f} ms/call, {us:.2f} us per evaluation per round "
              f"(~{us * clk_ghz:.1f} kcycles), {B * evals / ms * 1e3 / 1e6:.2f} M field evaluations/s")


if __name__ == "__main__":
    main()
