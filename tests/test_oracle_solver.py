"""Known-answer tests for the fixed-grid solver restatement (oracle/shims/torchdiffeq and
oracle/odevit_oracle.odeint_fixed).  The reference holds no test for this boundary and the real
torchdiffeq is not installable here (PARITY UNPINNED), so these pin the published
algorithm: exact linear-ODE answers, orders of convergence, 3/8-rule vs classic RK4, per-step dt,
and row layout."""
import math
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "shims"))
import torchdiffeq as shim  # noqa: E402
from odevit_oracle import odeint_fixed  # noqa: E402


def _lin_system():
    g = torch.Generator().manual_seed(0)
    A = torch.randn(5, 5, generator=g, dtype=torch.float64) * 0.5
    y0 = torch.randn(3, 5, generator=g, dtype=torch.float64)
    return A, y0


@pytest.mark.parametrize("method,order", [("euler", 1), ("midpoint", 2), ("rk4", 4)])
def test_order_of_convergence(method, order):
    A, y0 = _lin_system()
    exact = y0 @ torch.matrix_exp(A).T
    errs = []
    for n in (8, 16, 32):
        t = torch.linspace(0, 1, n + 1, dtype=torch.float64)
        ys = shim.odeint(lambda tt, y: y @ A.T, y0, t, method=method)
        errs.append(float((ys[-1] - exact).abs().max()))
    for e0, e1 in zip(errs, errs[1:]):
        assert math.log2(e0 / e1) == pytest.approx(order, abs=0.35)


def test_rk4_is_three_eighths_rule_not_classic():
    # one step of size h on y' = y^2 (nonlinear, so the two 4th-order tableaux differ at O(h^5))
    y0 = torch.tensor([0.7], dtype=torch.float64)
    h = 0.5
    f = lambda y: y * y
    ys = shim.odeint(lambda t, y: f(y), y0, torch.tensor([0.0, h], dtype=torch.float64), method="rk4")
    k1 = f(y0); k2 = f(y0 + h * k1 / 3); k3 = f(y0 + h * (k2 - k1 / 3)); k4 = f(y0 + h * (k1 - k2 + k3))
    three_eighths = y0 + h * (k1 + 3 * (k2 + k3) + k4) / 8
    c1 = f(y0); c2 = f(y0 + h * c1 / 2); c3 = f(y0 + h * c2 / 2); c4 = f(y0 + h * c3)
    classic = y0 + h * (c1 + 2 * c2 + 2 * c3 + c4) / 6
    assert torch.allclose(ys[-1], three_eighths, rtol=0, atol=1e-15)
    assert (ys[-1] - classic).abs().item() > 1e-7   # the two tableaux differ at O(h^5)


def test_rows_and_per_step_dt():
    # non-uniform grid: row j is the state at t[j]; dt is per step, not constant
    t = torch.tensor([0.0, 0.1, 0.4, 1.0], dtype=torch.float64)
    y0 = torch.tensor([[2.0]], dtype=torch.float64)
    ys = shim.odeint(lambda tt, y: -y, y0, t, method="euler")
    assert ys.shape == (4, 1, 1)
    expect = [2.0, 2.0 * 0.9, 2.0 * 0.9 * 0.7, 2.0 * 0.9 * 0.7 * 0.4]
    assert torch.allclose(ys.flatten(), torch.tensor(expect, dtype=torch.float64), atol=1e-15)
    assert torch.equal(ys[0], y0)


def test_time_is_cast_to_state_dtype_and_stage_times():
    seen = []
    def f(tt, y):
        seen.append((tt.dtype, float(tt)))
        return torch.zeros_like(y)
    shim.odeint(f, torch.zeros(1, dtype=torch.float32), torch.tensor([0.0, 0.3], dtype=torch.float64), method="rk4")
    assert all(d == torch.float32 for d, _ in seen)
    assert [round(v, 6) for _, v in seen] == [0.0, 0.1, 0.2, 0.3]


@pytest.mark.parametrize("method", ["euler", "midpoint", "rk4"])
def test_functional_restatement_matches_shim_bitwise(method):
    g = torch.Generator().manual_seed(1)
    W = torch.randn(6, 6, generator=g) * 0.3
    y0 = torch.randn(4, 6, generator=g)
    t = torch.linspace(0.0, 1.0, 7)
    f = lambda y: torch.tanh(y @ W.T)
    a = shim.odeint(lambda tt, y: f(y), y0, t, method=method)
    b = odeint_fixed(f, y0, t, method)
    assert torch.equal(a, b)


def test_backprop_through_solver():
    # gradient mode of the reference = plain autograd through the unrolled steps
    y0 = torch.tensor([1.5], dtype=torch.float64, requires_grad=True)
    t = torch.linspace(0, 1, 5, dtype=torch.float64)
    ys = shim.odeint(lambda tt, y: -0.5 * y, y0, t, method="euler")
    ys[-1].sum().backward()
    assert y0.grad.item() == pytest.approx((1 - 0.5 * 0.25) ** 4, abs=1e-14)


def test_rejects_unsupported():
    y0 = torch.zeros(1)
    with pytest.raises(ValueError):
        shim.odeint(lambda t, y: y, y0, torch.tensor([0.0, 1.0]), method="dopri8x")
    with pytest.raises(ValueError):
        shim.odeint(lambda t, y: y, y0, torch.tensor([0.0, 1.0, 0.5]), method="euler")
