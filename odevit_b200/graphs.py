"""CUDA-graph replay of a whole training step (forward, backward, gradient clipping, optimizer).

At CIFAR shapes the step is host-bound: ~300 kernel launches of 5-10 us each are enqueued through Python and
ctypes at ~8 ms per step while the GPU needs ~2 ms.  libodevit never allocates, never synchronises and reads its
weights at call time from fixed parameter storage, so the step can be captured once into a CUDA graph (static
input buffers, torch.cuda.graph's private memory pool for the tape and the autograd intermediates) and replayed
with one launch.  Eager and replayed steps run the same kernels on the same data: results are identical up to
the fp32 atomics of the weight-gradient accumulation.

    stepper = GraphedTrainStep(model, optimizer, example_inputs=(px, labels), clip=1.0)
    for px, labels in loader:
        loss = stepper(px, labels)          # device tensor; .item() it when you need the number

Restrictions: fixed input shapes; no data-dependent Python control flow in the step (the reference's step has
none: train.py:40-67); the optimizer must be capturable (`torch.optim.AdamW(..., fused=True, capturable=True)`) when
it is part of the graph.  Dropout (the shipped YAMLs use 0.1-0.3) works under replay: the mask seed lives in device
memory and a 1-thread kernel inside the captured step advances it (ops.DropState), so every replay draws new masks."""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch
from torch import nn


class GraphedTrainStep:
    def __init__(self, model: nn.Module, optimizer: torch.optim.Optimizer, example_inputs: Sequence[torch.Tensor],
                 clip: Optional[float] = 1.0, loss_fn: Optional[Callable] = None, warmup: int = 3,
                 grad_hook: Optional[Callable[[], None]] = None, capture_optimizer: bool = True):
        """loss_fn(model, *inputs) -> scalar loss (default: model(px, labels=labels)["loss"]); grad_hook runs
        between backward and clipping (the data-parallel FlatGradAllReduce).  capture_optimizer=False captures
        forward + backward only and runs grad_hook / clipping / optimizer.step() eagerly after each replay (a
        few dozen launches): required with a grad_hook that issues NCCL collectives -- capturing them hung a
        2-GPU run here -- and it lifts the `capturable=True` requirement on the optimizer."""
        # dropout under replay: the mask seed must live in device memory BEFORE the capture starts (ops.DropState is
        # created by the warm-up steps below; vit_ode._next_drop_seed reads this flag)
        for m in model.modules():
            if getattr(m, "_drops", None) is not None:
                m.device_seed = True
        # `block.attentions` (ode_transformer_gpt.py:276) keeps the last forward's map WITH its autograd graph, hence
        # the AccumulateGrad nodes of earlier eager steps on the legacy stream: a capture must not depend on that stream
        for m in model.modules():
            t = getattr(m, "attentions", None)
            if torch.is_tensor(t) and t.grad_fn is not None:
                m.attentions = t.detach()
        self.model, self.opt, self.clip = model, optimizer, clip
        self.loss_fn = loss_fn or (lambda m, px, lb: m(px, labels=lb)["loss"])
        self.grad_hook = grad_hook
        self.capture_optimizer = capture_optimizer
        if grad_hook is not None and capture_optimizer:
            raise ValueError("GraphedTrainStep: a grad_hook needs capture_optimizer=False (collectives stay eager)")
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.static_in = [t.clone() for t in example_inputs]
        # warm-up on a side stream (lazy initialisations, workspace growth, optimizer state) as torch prescribes
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream().wait_stream(s)
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.static_loss = self._eager_step(zero=False, tail=self.capture_optimizer)
        # The captured backward writes these very tensors at every replay.  Anything that re-binds `p.grad`
        # between replays (`zero_grad(set_to_none=True)`, an eager step on the same model) would leave the eager
        # tail reading stale or missing gradients: __call__ binds them back before the tail runs.
        self.static_grads = [p.grad for p in self.params]
        # device scratch whose addresses are baked into the graph: keep it alive as long as the graph is
        from . import ops
        self._pinned_workspaces = ops.live_workspaces()

    def _tail(self) -> None:
        if self.grad_hook is not None:
            self.grad_hook()
        if self.clip is not None:
            torch.nn.utils.clip_grad_norm_(self.params, self.clip, foreach=True)
        self.opt.step()

    def _eager_step(self, zero: bool = True, tail: bool = True) -> torch.Tensor:
        if zero:
            self.opt.zero_grad(set_to_none=True)
        loss = self.loss_fn(self.model, *self.static_in)
        loss.backward()
        if tail:
            self._tail()
        return loss.detach()

    def __call__(self, *inputs: torch.Tensor) -> torch.Tensor:
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()          # the captured backward (re)writes every static gradient tensor in place
        for p, g in zip(self.params, self.static_grads):
            if p.grad is not g:
                p.grad = g
        if not self.capture_optimizer:
            self._tail()
        return self.static_loss
