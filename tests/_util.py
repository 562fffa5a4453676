"""Shared helpers for the test-suite: golden fixture loading and error metrics."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
    """One tests/golden/<name>.npz fixture (see oracle/make_golden.py for the layout)."""

    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.meta = json.loads(str(self.z["meta"]))

    def group(self, prefix, dtype=None, device=None):
        out = {}
        for k in self.z.files:
            if k.startswith(prefix + "/"):
                t = torch.from_numpy(np.array(self.z[k]))
                if dtype is not None and t.is_floating_point():
                    t = t.to(dtype)
                if device is not None:
                    t = t.to(device)
                out[k[len(prefix) + 1:]] = t
        return out

    def get(self, key, device=None):
        t = torch.from_numpy(np.array(self.z[key]))
        return t.to(device) if device is not None else t

    @property
    def ctor(self):
        return self.meta["ctor"]

    @property
    def call(self):
        call = dict(self.meta["call"])
        if call.get("t_grid") is not None:
            call["t_grid"] = torch.tensor(call["t_grid"], dtype=torch.float32)
        return call


def max_rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def cosine(a, b):
    a = torch.as_tensor(a).detach().double().cpu().flatten()
    b = torch.as_tensor(b).detach().double().cpu().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))


VIT_CASES = ["c10_rk4_T5_B2", "c10_euler_T13_B2", "tiny_euler_T6_B3", "tiny_rk4_T4_B3",
             "tiny_midpoint_T5_B2", "tiny_rk4_tgrid_B2", "tiny_dist_token_B2"]
MACARON_CASES = ["macaron_rk4_T4_B2", "macaron_euler_T13_B2"]
