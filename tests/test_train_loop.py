"""`odevit_b200.train.train_classification_task` against the UNMODIFIED reference loop (train.py:18-110) on a CPU
stand-in model: same weights after an epoch with gradient accumulation, clipping, AdamW and a scheduler -- including
the reference's two quirks (JaSMin added twice; the clip only bites at the first optimizer step of a call because
`params` is a generator).  Needs /root/reference (build container only)."""
import copy
import importlib
import os
import sys

import pytest
import torch
from torch import nn

import ref_import

pytestmark = pytest.mark.skipif(not ref_import.reference_available(), reason="reference sources not present")


class Stub(nn.Module):
    """Returns what the train loop reads: logits, loss, jasmin_loss (a differentiable function of the weights)."""

    def __init__(self):
        super().__init__()
        self.fc = nn.Linear(12, 5)

    def forward(self, pixel_values, labels=None, **kw):
        logits = self.fc(pixel_values.flatten(1))
        return {"logits": logits, "loss": nn.functional.cross_entropy(logits, labels) * 3.0,
                "jasmin_loss": logits.square().mean()}


def _loader(n=7):
    from odevit_b200.data import PixelBatch
    g = torch.Generator().manual_seed(0)
    return [{"pixel_values": PixelBatch({"pixel_values": torch.randn(4, 3, 2, 2, generator=g) * 3}), "labels": torch.randint(0, 5, (4,), generator=g)}
            for _ in range(n)]


def _reference_train():
    ref_import.import_reference()
    shims = os.path.join(os.path.dirname(ref_import.__file__), "shims")
    if shims not in sys.path:
        sys.path.insert(0, shims)
    sys.modules.pop("train", None)
    return importlib.import_module("train")


@pytest.mark.parametrize("accum", [1, 2, 3])
def test_train_loop_matches_reference(accum):
    ref = _reference_train()
    from odevit_b200 import train as ours
    torch.manual_seed(0)
    base = Stub()
    results = []
    for fn in (ref.train_classification_task, ours.train_classification_task):
        m = copy.deepcopy(base)
        opt = torch.optim.AdamW(m.parameters(), lr=1e-2, weight_decay=5e-2)
        sch = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.7)
        _, loss = fn(_loader(), m, opt, None, sch, wandb_logger=None, epoch=1, num_accumulation_steps=accum, log_every=3)
        results.append((loss, [p.detach().clone() for p in m.parameters()], [None if p.grad is None else p.grad.clone() for p in m.parameters()],
                        opt.param_groups[0]["lr"]))
    (l0, p0, g0, lr0), (l1, p1, g1, lr1) = results
    assert l0 == pytest.approx(l1, rel=1e-6)
    assert lr0 == lr1
    for a, b in zip(p0, p1):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-7)
    for a, b in zip(g0, g1):                     # the partial accumulation group left in .grad
        assert (a is None) == (b is None) and (a is None or torch.allclose(a, b, rtol=1e-6, atol=1e-7))


def test_config_and_schedule_follow_the_shipped_yaml():
    from odevit_b200 import train as ours
    path = os.path.join(ref_import.REFERENCE_ROOT, "configs", "classification", "experiment_vit_edo.yaml")
    cfg = ours.load_config(path, ["setup.dict.epochs=200", "modeling.inputs.solver=rk4"])
    assert cfg["modeling"]["inputs"]["embed_dim"] == 768 and cfg["modeling"]["inputs"]["solver"] == "rk4"
    assert cfg["setup"]["dict"]["epochs"] == 200 and cfg["data"]["collator"]["train"]["batch_size"] == 64
    opt = torch.optim.AdamW([nn.Parameter(torch.zeros(1))], lr=1e-4)
    sch = ours.restart_schedule(opt, 200, 10)
    lrs = []
    for _ in range(2000):
        opt.step()
        sch.step()
        lrs.append(opt.param_groups[0]["lr"])
    assert max(lrs) == pytest.approx(1e-4, rel=1e-3) and lrs[199] == pytest.approx(1e-4, rel=1e-2)   # 10 % warm-up
    assert lrs[1098] < 1e-6 and lrs[1100] > 9e-5                                                        # hard restart (2 cycles)
