"""The fused L1-attention front-end (odevit_extract_mass_fwd / _bwd, csrc/mass.cu) against
the reference's `ImageDistilTrainer.extract_mass` (loss_trainer.py:80-117): the golden holds outputs and input
gradients produced by the UNMODIFIED method (oracle/make_golden_distill.py::mass_golden); the plain PyTorch
composition of odevit_b200/loss_trainer.py is compared at more shapes."""
import numpy as np
import pytest
import torch

from _util import Golden, max_rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["smooth_t05", "smooth_t07", "hard_t08"])
def test_extract_mass_matches_reference_golden(case):
    from odevit_b200 import ops
    g = Golden("extract_mass")
    meta = g.meta[case]
    rows = g.get(f"{case}/rows").cuda().requires_grad_(True)
    mean, heads, mask = ops.extract_mass(rows, threshold=meta["threshold"], smooth=meta["smooth"], scale_factor=meta["scale_factor"],
                                         return_mask=True)
    assert max_rel(mean, g.get(f"{case}/mean")) < 1e-5
    assert max_rel(heads, g.get(f"{case}/heads")) < 1e-5
    assert max_rel(mask, g.get(f"{case}/mask")) < 1e-5
    ((mean * g.get(f"{case}/w_mean").cuda()).sum() + (heads * g.get(f"{case}/w_heads").cuda()).sum()).backward()
    assert max_rel(rows.grad, g.get(f"{case}/grad_rows")) < 1e-4


@pytest.mark.parametrize("B,H,n", [(3, 12, 196), (2, 3, 64), (1, 2, 576), (5, 1, 4), (2, 4, 1024)])
def test_extract_mass_matches_composition(B, H, n):
    from odevit_b200 import loss_trainer, ops
    x = torch.softmax(torch.randn(B, H, n, generator=torch.Generator().manual_seed(n)) * 2, -1)
    wm = torch.randn(B, int(n ** 0.5), int(n ** 0.5), generator=torch.Generator().manual_seed(1))
    xr = x.clone().requires_grad_(True)
    mr, hr, kr = loss_trainer.extract_mass(xr, threshold=0.5, return_mask=True)          # CPU composition
    (mr * wm).sum().backward()
    xg = x.cuda().requires_grad_(True)
    mg, hg, kg = ops.extract_mass(xg, threshold=0.5, return_mask=True)
    (mg * wm.cuda()).sum().backward()
    assert max_rel(mg, mr) < 1e-5 and max_rel(hg, hr) < 1e-5 and max_rel(kg, kr) < 1e-5
    assert max_rel(xg.grad, xr.grad) < 1e-4
