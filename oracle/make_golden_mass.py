"""TEST INFRASTRUCTURE ONLY -- golden vectors of the reference's `ImageDistilTrainer.extract_mass`
(loss_trainer.py:80-117), produced by the UNMODIFIED method (imported from /root/reference through ref_import.py):

    tests/golden/extract_mass.npz     <case>/rows [B,H,n], mean, heads, mask, w_mean, w_heads (random cotangents),
                                      grad_rows = d(<mean,w_mean> + <heads,w_heads>)/d rows;  meta: per-case arguments

    python oracle/make_golden_mass.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_import import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "extract_mass.npz")
CASES = {"smooth_t05": dict(threshold=0.5, smooth=True, scale_factor=40, shape=(4, 12, 196)),
         "smooth_t07": dict(threshold=0.7, smooth=True, scale_factor=40, shape=(3, 2, 16)),
         "hard_t08": dict(threshold=0.8, smooth=False, scale_factor=40, shape=(2, 3, 64))}


def main():
    lt = import_reference(with_loss_trainer=True)["loss_trainer"]
    self = lt.ImageDistilTrainer.__new__(lt.ImageDistilTrainer)      # extract_mass reads no attribute of the trainer
    store, meta = {}, {}
    for i, (name, c) in enumerate(CASES.items()):
        g = torch.Generator().manual_seed(100 + i)
        B, H, n = c["shape"]
        rows = torch.softmax(torch.randn(B, H, n + 1, generator=g) * 2.5, -1)[..., 1:].contiguous().requires_grad_(True)
        mean, heads, mask = lt.ImageDistilTrainer.extract_mass(self, rows, threshold=c["threshold"], smooth=c["smooth"],
                                                               scale_factor=c["scale_factor"], return_mask=True)
        w_mean, w_heads = torch.randn(mean.shape, generator=g), torch.randn(heads.shape, generator=g)
        ((mean * w_mean).sum() + (heads * w_heads).sum()).backward()
        for k, v in dict(rows=rows, mean=mean, heads=heads, mask=mask, w_mean=w_mean, w_heads=w_heads, grad_rows=rows.grad).items():
            store[f"{name}/{k}"] = v.detach().numpy()
        meta[name] = {k: v for k, v in c.items() if k != "shape"}
    store["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
