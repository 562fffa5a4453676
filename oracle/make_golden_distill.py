"""TEST INFRASTRUCTURE ONLY -- golden vectors of the reference's REAL distillation trainer.

Runs the UNMODIFIED `loss_trainer.ImageDistilTrainer.forward` (loss_trainer.py:305-372; imported from
/root/reference through oracle/ref_import.py, `turtle` shimmed) on a small random-init HF ViT teacher and the
UNMODIFIED reference `ViTNeuralODE` student, and stores everything a replacement must reproduce:

    tests/golden/distill_trainer_tiny.npz
      sd/<key>      student state_dict            tsd/<key>   teacher state_dict
      in/...        pixel_values, labels
      e0/...        epoch 0   (no CE term):  loss, mse_loss, kl_loss, jasmin_loss, supervision_loss,
                                             mse_loss_t@i, grad/<param> (AFTER clip_grad_norm_ 1.0, as the
                                             trainer leaves them), gnorm (pre-clip total norm)
      e201/...      epoch 201 (CE added, loss_trainer.py:348-349): same keys
      meta          JSON: student ctor, teacher ViTConfig kwargs, trainer kwargs

The optimizer is SGD with lr = 0 so the trainer's own `optimizer.step()` leaves the weights where they were
(the two epochs see the same model).  Shapes: the YAML's structure at toy size -- 12 teacher layers <-> 12
control points, teacher width == student width (MSE on CLS rows), teacher patches == student patches (L1 on
the 4 x 4 attention mass), Euler, `mse_full_path=True`, `lambda_param=0.5`, `jasmin_k=2`
(configs/classification/experiment_classification_edo_distillation.yaml:9-23).

    python oracle/make_golden_distill.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_import import import_reference  # noqa: E402

OUT_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")

STUDENT = dict(img_size=32, patch_size=8, num_classes=10, embed_dim=128, num_heads=2, mlp_ratio=1.0,
               emulate_depth=12, time_interval=1.0, num_eval_steps=24, solver="euler", register_tokens=2)
TEACHER = dict(hidden_size=128, num_hidden_layers=12, num_attention_heads=2, intermediate_size=128,
               image_size=32, patch_size=8, num_labels=10, hidden_act="gelu")
TRAINER = dict(mse_full_path=True, use_distillation=True, use_supervision=True, use_mse_loss=True,
               temperature=2.0, jasmin_k=2, lambda_param=0.5)


def build_teacher(seed=1):
    from transformers import ViTConfig, ViTForImageClassification
    torch.manual_seed(seed)
    return ViTForImageClassification(ViTConfig(attn_implementation="eager", **TEACHER)).eval()


def main():
    torch.set_num_threads(4)
    mods = import_reference(with_loss_trainer=True)
    ode, lt = mods["ode"], mods["loss_trainer"]
    teacher = build_teacher()
    torch.manual_seed(0)
    student = ode.ViTNeuralODE(**STUDENT)
    g = torch.Generator().manual_seed(99)
    with torch.no_grad():   # the reference's init leaves norms at (1, 0): jitter so those paths carry signal
        for n, p in student.named_parameters():
            if "norm" in n and n.endswith("weight"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif n.endswith("bias"):
                p.add_(0.05 * torch.randn(p.shape, generator=g))
        # sharpen the student's attention a little: near-uniform rows make extract_mass's sort order a coin toss
        student.odefunc.block.attn.mha.in_proj_weight[: 2 * STUDENT["embed_dim"]].mul_(2.0)
    B = 3
    px = torch.randn(B, 3, 32, 32, generator=torch.Generator().manual_seed(1234))
    labels = torch.randint(0, 10, (B,), generator=torch.Generator().manual_seed(1235))
    store = {"in/pixel_values": px.numpy(), "in/labels": labels.numpy()}
    for k, v in student.state_dict().items():
        store[f"sd/{k}"] = v.numpy().copy()
    for k, v in teacher.state_dict().items():
        store[f"tsd/{k}"] = v.numpy().copy()
    opt = torch.optim.SGD(student.parameters(), lr=0.0)
    trainer = lt.ImageDistilTrainer(teacher_model=teacher, student_model=student, optimizer=opt, scheduler=None, **TRAINER)
    for tag, epoch in (("e0", 0), ("e201", 201)):
        out = trainer({"pixel_values": px}, labels, epoch=epoch)          # the unmodified forward: backward + clip + step
        for k, v in out.items():
            if torch.is_tensor(v) and v.ndim == 0:
                store[f"{tag}/{k}"] = v.detach().numpy().copy()
        for k, p in student.named_parameters():
            store[f"{tag}/grad/{k}"] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().copy()
        so = out["student_output"]
        store[f"{tag}/logits"] = so["logits"].detach().numpy().copy()
        store[f"{tag}/control_points"] = so["control_points"].detach().numpy().copy()
        store[f"{tag}/attentions"] = so["attentions"].detach().numpy().copy()
        # pre-clip total norm: one more backward of the same loss on fresh grads
        opt.zero_grad(set_to_none=True)
        s_out = student(pixel_values=px, labels=labels, output_hidden_states=True, output_control_points=True,
                        output_attentions=True, jasmin_k=TRAINER["jasmin_k"])
        with torch.no_grad():
            t_out = teacher(pixel_values=px, output_hidden_states=True, output_attentions=True)
        trainer.epoch = epoch
        rep = trainer.train_batch_representation(s_out, t_out)
        loss = rep["loss"] + s_out["jasmin_loss"] + (s_out["loss"] if epoch > 200 else 0.0)
        loss.backward()
        gn = torch.norm(torch.stack([p.grad.norm(2) for p in student.parameters() if p.grad is not None]), 2)
        store[f"{tag}/gnorm"] = gn.numpy().copy()
        print(tag, {k: float(v) for k, v in out.items() if torch.is_tensor(v) and v.ndim == 0}, "gnorm", float(gn))
    store["meta"] = np.asarray(json.dumps({"kind": "distill_trainer", "student": STUDENT, "teacher": TEACHER,
                                           "trainer": TRAINER, "B": B}))
    np.savez_compressed(os.path.join(OUT_DIR, "distill_trainer_tiny.npz"), **store)


if __name__ == "__main__":
    main()
