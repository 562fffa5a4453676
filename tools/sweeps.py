"""The other BASELINE.json configurations, measured on one GPU (parity cases in tests/, numbers here):

  config 1  C10 model, RK4 (3/8) T in {5, 13}, fwd+bwd, batch 8            -> ms / step, img/s
  config 4  S3.8M-shape model (224 px, D 768, r 1, R 10), inference, Euler T=36, batch sweep
  config 5  C10 model, Euler and RK4, steps in {4, 8, 16, 32, 64}, inference -> NFE/s

    python tools/sweeps.py [--which 1,4,5] [--precision bf16] > profiles/rNN_sweeps.jsonl

One JSON line per point: device time by CUDA events (max of 3 repeats' median is not taken: mean of
`reps` timed calls after 2 warm-up calls), inputs resident on the device; L2 note as in bench.py."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import odevit_b200 as ob  # noqa: E402

STAGES = {"euler": 1, "midpoint": 2, "rk4": 4}
C10 = dict(img_size=32, patch_size=4, num_classes=10, embed_dim=192, num_heads=3, mlp_ratio=4.0, emulate_depth=12,
           time_interval=1.0, num_eval_steps=5, solver="rk4", register_tokens=4)
S38 = dict(img_size=224, patch_size=16, num_classes=100, embed_dim=768, num_heads=12, mlp_ratio=1.0, emulate_depth=12,
           time_interval=1.0, num_eval_steps=36, solver="euler", register_tokens=10)


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def flops_fwd(cfg):
    D = cfg["embed_dim"]
    N = (cfg["img_size"] // cfg["patch_size"]) ** 2 + 1 + cfg["register_tokens"]
    return 8.0 * N * D * D + 4.0 * N * D * int(D * cfg["mlp_ratio"]) + 4.0 * N * N * D


def build(cfg, precision, train=False):
    torch.manual_seed(0)
    m = ob.ViTNeuralODE(**cfg).cuda()
    m.precision = precision
    return m.train() if train else m.eval()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="1,4,5")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    which = set(a.which.split(","))
    out = []
    if "1" in which:
        for T in (5, 13):
            cfg = dict(C10, num_eval_steps=T)
            for prec in ("fp32", a.precision):
                m = build(cfg, prec, train=True)
                px = torch.randn(8, 3, 32, 32, device="cuda")
                lb = torch.randint(0, 10, (8,), device="cuda")

                def step():
                    m.zero_grad(set_to_none=True)
                    m(px, labels=lb)["loss"].backward()
                ms = timed(step, a.reps * 4)
                out.append({"config": 1, "workload": f"C10 RK4 T={T} fwd+bwd batch 8", "precision": prec,
                            "ms_per_step": ms, "img_per_s": 8 / ms * 1e3, "nfe_per_s": 8 * (T - 1) * 4 / ms * 1e3})
    if "4" in which:
        m = build(S38, a.precision)
        nfe = (S38["num_eval_steps"] - 1)
        for B in (64, 256, 1024, 2048, 4096, 8192):
            px = torch.randn(B, 3, 224, 224, device="cuda")
            with torch.no_grad():
                ms = timed(lambda: m(px), max(2, a.reps if B <= 256 else 2))
            tf = B * nfe * flops_fwd(S38) / (ms * 1e-3) / 1e12
            lean = 4 * S38["num_eval_steps"] * B * 207 * 768 > 0.05 * torch.cuda.mem_get_info()[0]
            out.append({"config": 4, "workload": "S3.8M-shape inference, Euler T=36", "trajectory": "none (FD bound formed inside "
                        "the solve, odevit_solve_fwd_lean)" if lean else "materialised (one FD pass over it)",
                        "precision": a.precision, "batch": B, "ms_per_call": ms,
                        "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
                        "img_per_s": B / ms * 1e3, "nfe_per_s": B * nfe / ms * 1e3, "algorithmic_tflops": tf})
            del px
            ob.ops.free_workspaces()
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()
    if "5" in which:
        B = 512
        px = torch.randn(B, 3, 32, 32, device="cuda")
        for solver in ("euler", "rk4"):
            for steps in (4, 8, 16, 32, 64):
                cfg = dict(C10, solver=solver, num_eval_steps=steps + 1)
                m = build(cfg, a.precision)
                with torch.no_grad():
                    ms = timed(lambda: m(px), a.reps)
                nfe = steps * STAGES[solver]
                tf = B * nfe * flops_fwd(cfg) / (ms * 1e-3) / 1e12
                out.append({"config": 5, "workload": f"C10 inference, {solver}, {steps} steps over [0,1], batch {B}",
                            "precision": a.precision, "ms_per_call": ms, "img_per_s": B / ms * 1e3,
                            "nfe_per_s": B * nfe / ms * 1e3, "algorithmic_tflops": tf})
    for r in out:
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
