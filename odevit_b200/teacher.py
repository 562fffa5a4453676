"""The frozen teacher of the distillation step through libodevit.so.

The reference's teacher is a HF `ViTForImageClassification` run under `torch.no_grad()` in eager fp32 PyTorch
(main_classification_ode_distillation.py:64-99, :170; loss_trainer.py:318-321).  `ViTTeacher` wraps such a model:
HF's own embedding module produces the token sequence, the 12 pre-LayerNorm layers run in the library
(`odevit_encoder_fwd`: one GEMM per projection with bias / GELU / residual in its epilogue, the fused
attention kernel), and the result comes back under the keys the loss code reads --
`out["hidden_states"]` (tuple of L+1 tensors [B,N,D]), `out["attentions"]`, `out["logits"]` -- by item or
attribute like a HF ModelOutput.  `precision="bf16"` (tensor cores, fp32 accumulation and residual stream;
<= 2e-2 of the fp32 teacher) or `"fp32"`."""
from __future__ import annotations

import torch
from torch import nn

from . import ops


class TeacherOutput(dict):
    __getattr__ = dict.__getitem__


class ViTTeacher(nn.Module):
    def __init__(self, hf_model: nn.Module, precision: str = "bf16", attention_maps: str = "last"):
        """attention_maps: "last" returns a 1-tuple holding the last layer's map (all the reference's losses read:
        `torch.stack(attentions)[-1]`, loss_trainer.py:169); "all" returns every layer's (HF's output_attentions)."""
        super().__init__()
        self.hf = hf_model.eval()
        self.vit = hf_model.vit if hasattr(hf_model, "vit") else hf_model
        self.precision = precision
        self.attention_maps = attention_maps
        self._weights = None
        for p in self.hf.parameters():
            p.requires_grad_(False)

    def _layer_weights(self):
        layers = []
        for lyr in self.vit.encoder.layer:
            att = lyr.attention.attention
            layers.append(dict(
                norm_a_w=lyr.layernorm_before.weight, norm_a_b=lyr.layernorm_before.bias,
                norm_b_w=lyr.layernorm_after.weight, norm_b_b=lyr.layernorm_after.bias,
                in_proj_w=torch.cat([att.query.weight, att.key.weight, att.value.weight], 0),
                in_proj_b=torch.cat([att.query.bias, att.key.bias, att.value.bias], 0),
                out_proj_w=lyr.attention.output.dense.weight, out_proj_b=lyr.attention.output.dense.bias,
                fc1_w=lyr.intermediate.dense.weight, fc1_b=lyr.intermediate.dense.bias,
                fc2_w=lyr.output.dense.weight, fc2_b=lyr.output.dense.bias))
        return ops.EncoderWeights(layers)

    def refresh_weights(self) -> None:
        """Re-read the HF parameters (after loading a checkpoint into the wrapped model)."""
        self._weights = None

    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor, output_hidden_states: bool = True, output_attentions: bool = True,
                **_unused) -> TeacherOutput:
        cfg = self.vit.config
        if cfg.hidden_act != "gelu":
            raise ValueError(f"ViTTeacher: activation {cfg.hidden_act!r} is not the erf GELU the kernels implement")
        if self._weights is None or self._weights._keep[0].device != pixel_values.device:
            self._weights = self._layer_weights()
        emb = self.vit.embeddings(pixel_values)
        maps_mode = self.attention_maps if output_attentions else "none"
        hidden, maps = ops.encoder_forward(emb, self._weights, cfg.num_attention_heads, cfg.intermediate_size,
                                           cfg.layer_norm_eps, self.precision, maps_mode)
        out = TeacherOutput()
        out["last_hidden_state"] = self.vit.layernorm(hidden[-1])
        if hasattr(self.hf, "classifier"):
            out["logits"] = self.hf.classifier(out["last_hidden_state"][:, 0])
        if output_hidden_states:
            out["hidden_states"] = (emb,) + tuple(hidden.unbind(0))
        if output_attentions:
            out["attentions"] = (maps,) if maps_mode == "last" else tuple(maps.unbind(0))
        return out
