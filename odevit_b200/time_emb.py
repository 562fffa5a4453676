"""Time embedding of the reference (models/time_emb.py), kept as plain PyTorch modules with the same
class names, constructor arguments and `state_dict` keys:

    SinusoidalPosEmb :7-41        LearnedSinusoidalPosEmb :44-68
    TimeEmbedding :71-111         ScaleShift :114-132

The reference never calls them from the vector field; they produce the
optional per-call `(scale, shift)` vectors that libodevit.so folds into the CenterNorm prologue as
n * (1 + scale) + shift (include/odevit.h `mod_*`; `attach_time_modulation` below).  How the
reference would have applied them is unspecified, so that composition is our documented choice.

Known defects of the reference kept out: `LearnedSinusoidalPosEmb.forward` stops in a stray
`pdb.set_trace()` (:66) -- not reproduced; `TimeEmbedding(learnable_sinusoidal=False)` fails on a
width mismatch (lin1 expects 2*sd+1 features, SinusoidalPosEmb yields sd+1) -- reproduced, since it
follows from the documented shapes.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


class SinusoidalPosEmb(nn.Module):
    """t -> [t, sin(scale*t*w_i), cos(scale*t*w_i)], w_i = max_period^(-i/half), half = dim // 2."""

    def __init__(self, dim: int):
        super().__init__()
        assert dim % 2 == 0
        self.dim = dim

    def forward(self, x: torch.Tensor, max_period: float = 10000, scale: float = 1000) -> torch.Tensor:
        half = self.dim // 2
        xs = x * scale
        w = torch.exp(torch.arange(half, dtype=x.dtype, device=x.device) * (-math.log(max_period) / half))
        ang = xs[..., None] * w
        return torch.cat([(xs / scale)[..., None], ang.sin(), ang.cos()], dim=-1)


class LearnedSinusoidalPosEmb(nn.Module):
    """t -> [t, sin(2 pi t w), cos(2 pi t w)] with learnable frequencies `weights` [dim]."""

    def __init__(self, dim: int):
        super().__init__()
        assert dim % 2 == 0
        self.dim = dim
        self.weights = nn.Parameter(torch.randn(dim))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        ang = x[..., None] * self.weights * 2 * math.pi
        return torch.cat([x[..., None], ang.sin(), ang.cos()], dim=-1)


class TimeEmbedding(nn.Module):
    """Fourier features (width 2*sinusoidal_dim + 1) -> Linear -> SiLU -> Dropout -> Linear(embed_dim)."""

    def __init__(self, sinusoidal_dim: int, embed_dim: int, multiplier: int = 1, dropout: float = 0.1,
                 learnable_sinusoidal: bool = False):
        super().__init__()
        self.sinusoidal = (LearnedSinusoidalPosEmb if learnable_sinusoidal else SinusoidalPosEmb)(sinusoidal_dim)
        self.lin1 = nn.Linear(2 * sinusoidal_dim + 1, embed_dim * multiplier)
        self.lin2 = nn.Linear(embed_dim * multiplier, embed_dim)
        self.dropout = nn.Dropout(dropout)

    def forward(self, t: torch.Tensor) -> torch.Tensor:
        return self.lin2(self.dropout(F.silu(self.lin1(self.sinusoidal(t)))))


class ScaleShift(nn.Module):
    """SiLU -> Linear(2*out_dim); the output interleaves (scale_0, shift_0, scale_1, shift_1, ...)."""

    def __init__(self, embed_dim: int, out_dim: int):
        super().__init__()
        self.lin = nn.Linear(embed_dim, out_dim * 2)

    def forward(self, x: torch.Tensor):
        y = self.lin(F.silu(x))
        y = y.view(*y.shape[:-1], y.shape[-1] // 2, 2)
        return y[..., 0], y[..., 1]


def attach_time_modulation(block, emb: Optional[torch.Tensor], attn: Optional[ScaleShift] = None,
                           mlp: Optional[ScaleShift] = None) -> None:
    """Give the parallel block (`ParallelAttentionMLP`) per-call modulation vectors computed from ONE
    time embedding `emb` [embed_dim] (detached: the fold treats them as constants of the call).
    `emb=None` switches the modulation off (the reference's behaviour)."""
    if emb is None:
        block.time_mod = None
        return
    mod = {}
    if attn is not None:
        s, b = attn(emb)
        mod["mod_attn_scale"], mod["mod_attn_shift"] = s.detach().float().contiguous(), b.detach().float().contiguous()
    if mlp is not None:
        s, b = mlp(emb)
        mod["mod_mlp_scale"], mod["mod_mlp_shift"] = s.detach().float().contiguous(), b.detach().float().contiguous()
    block.time_mod = mod
