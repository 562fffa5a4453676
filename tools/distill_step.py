"""BASELINE.json config 3 on one GPU: teacher-student distillation step, random-init DINO-shape ViT-B/16
teacher (HF `ViTForImageClassification`, eager attention, no grad) + ODE-ViT student (S7M: 224 px, patch 16,
D 768, H 12, r 4.0, R 10 -> 207 tokens; Euler over T=36 grid points), 64 images, synthetic data.

The student is this repo's drop-in module; its outputs (`control_points`, `attentions`, `jasmin_loss`) are
consumed UNCHANGED by the reference's trajectory losses, restated below because /root/reference does not
exist on the GPU box (the loss consumers stay PyTorch):

  loss = lambda * (sum_i mean MSE(teacher_hidden[i+1][:, 0], control_points[i][:, 0])            loss_trainer.py:119-160
                   + lambda * sum | mass(student CLS attention row) - mass(teacher CLS row) |)     :162-183, :80-117
         + jasmin_loss                                                                              :345-346
  (`use_supervision` adds the CE term only after epoch 200, :348: not in this step), then backward,
  clip_grad_norm_ 1.0 and AdamW (lr 1e-4, wd 5e-2), as `ImageDistilTrainer.forward` does (:305-374).
  YAML values: configs/classification/experiment_classification_edo_distillation.yaml:9-23.

    python tools/distill_step.py [--batch 64] [--steps 5] [--ratio 4.0]  ->  one JSON line"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import odevit_b200 as ob  # noqa: E402


def extract_mass(attn, threshold, scale_factor=40.0):
    """loss_trainer.py:80-117 (smooth=True): soft mask of the rows' top-`threshold` mass, 3x3 Gaussian blur
    (sigma 0.5), mean over heads."""
    from torchvision.transforms.functional import gaussian_blur
    B, nh, n = attn.shape
    side = int(n ** 0.5 + 0.5)
    val, idx = torch.sort(attn, dim=-1)
    val = val / (val.sum(dim=-1, keepdim=True) + 1e-8)
    mask_soft = torch.sigmoid((torch.cumsum(val, dim=-1) - (1 - threshold)) * scale_factor)
    th = torch.gather(mask_soft, dim=-1, index=torch.argsort(idx, dim=-1)).view(B, nh, side, side).float()
    filt = gaussian_blur(attn.view(B, nh, side, side) * th, kernel_size=(3, 3), sigma=0.5)
    return filt.mean(dim=1)


def distillation_loss(student_out, teacher_out, lambda_param=0.5):
    teacher_states = torch.stack(teacher_out["hidden_states"], dim=0)[1:]              # :259
    cps = student_out["control_points"]                                                 # :274
    mse = sum(F.mse_loss(t[:, 0], c[:, 0], reduction="none").mean() for t, c in zip(teacher_states, cps))   # :136-143
    attn_t = teacher_out["attentions"][-1][:, :, 0, 1:]                                 # :169-171
    attn_s = student_out["attentions"][:, :, 0, 1:]
    l1 = (extract_mass(attn_s, 0.5) - extract_mass(attn_t, 0.7)).abs().sum() * lambda_param      # :174-183
    if os.environ.get("DISTILL_NO_L1"):   # diagnosis: how much of the backward is the cotangent on `attentions`
        l1 = l1.detach()
    return (mse + l1) * lambda_param + student_out["jasmin_loss"]                       # :297, :345-346


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--ratio", type=float, default=4.0)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--teacher", default="library", choices=["library", "hf"],
                    help="library: odevit_b200.ViTTeacher (encoder through libodevit.so, same precision mode); hf: the wrapped HF model in eager fp32 PyTorch, as the reference runs it")
    ap.add_argument("--graph", action="store_true", help="replay student forward + teacher + losses + backward as one CUDA graph (odevit_b200.graphs)")
    a = ap.parse_args()
    from transformers import ViTConfig, ViTForImageClassification
    dev = torch.device("cuda", 0)
    torch.manual_seed(1)
    teacher = ViTForImageClassification(ViTConfig(num_labels=100, attn_implementation="eager")).to(dev).eval()
    if a.teacher == "library":
        teacher = ob.ViTTeacher(teacher, precision=a.precision, attention_maps="last")
    cfg = dict(img_size=224, patch_size=16, num_classes=100, embed_dim=768, num_heads=12, mlp_ratio=a.ratio, emulate_depth=12,
               time_interval=1.0, num_eval_steps=36, solver="euler", register_tokens=10)
    torch.manual_seed(0)
    student = ob.ViTNeuralODE(**cfg).to(dev).train()
    student.precision = a.precision
    params = [p for p in student.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=5e-2, fused=True)
    B = a.batch
    px = torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(1234)).to(dev)
    lb = torch.randint(0, 100, (B,), generator=torch.Generator().manual_seed(1235)).to(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    split = {"teacher": 0.0, "student_fwd_loss": 0.0, "backward_opt": 0.0}

    def step(record=False):
        opt.zero_grad(set_to_none=True)
        if record: ev[0].record()
        s_out = student(px, labels=lb, output_hidden_states=True, output_control_points=True, output_attentions=True, jasmin_k=2)
        if record: ev[1].record()
        with torch.no_grad():
            t_out = teacher(pixel_values=px, output_hidden_states=True, output_attentions=True)
        if record: ev[2].record()
        loss = distillation_loss(s_out, t_out)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0, foreach=True)
        opt.step()
        if record:
            ev[3].record()
            torch.cuda.synchronize()
            split["student_fwd_loss"] += ev[0].elapsed_time(ev[1])
            split["teacher"] += ev[1].elapsed_time(ev[2])
            split["backward_opt"] += ev[2].elapsed_time(ev[3])
        return loss

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    run = step
    if a.graph:
        from odevit_b200.graphs import GraphedTrainStep

        def loss_fn(model, px_, lb_):
            s_out = model(px_, labels=lb_, output_hidden_states=True, output_control_points=True, output_attentions=True, jasmin_k=2)
            with torch.no_grad():
                t_out = teacher(pixel_values=px_, output_hidden_states=True, output_attentions=True)
            return distillation_loss(s_out, t_out)
        stepper = GraphedTrainStep(student, opt, (px, lb), clip=1.0, loss_fn=loss_fn, warmup=1, capture_optimizer=False)
        run = lambda: stepper(px, lb)   # noqa: E731
        run()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    for _ in range(3):
        step(record=True)
    nfe = cfg["num_eval_steps"] - 1
    print(json.dumps({
        "config": 3, "teacher": a.teacher, "launch_mode": "cuda_graph_replay" if a.graph else "eager", "workload": f"distillation step: ViT-B/16 teacher ({'odevit_b200.ViTTeacher' if a.teacher == 'library' else 'eager fp32 PyTorch'}, no grad) + ODE-ViT student r={a.ratio} "
        f"(N=207, Euler T=36), MSE full path + L1 attention mass + JaSMin k=2, batch {B}",
        "precision": a.precision, "ms_per_step": ms, "img_per_s": B / ms * 1e3, "student_field_evals_per_s": B * nfe / ms * 1e3,
        "loss": float(loss), "split_ms": {k: round(v / 3, 3) for k, v in split.items()},
        "student_params_M": round(sum(p.numel() for p in params) / 1e6, 3)}))


if __name__ == "__main__":
    main()
