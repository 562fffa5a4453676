"""TEST INFRASTRUCTURE ONLY -- CPU restatement of torchdiffeq's fixed-grid solvers.

PARITY UNPINNED at this boundary: the reference (`/root/reference/models/
ode_transformer_gpt.py:9`, `models/macaron.py:7`) imports `odeint` from the third-party
package `torchdiffeq` (version unpinned by the reference: no requirements/lock file; latest
public release 0.2.5).  That package is neither vendored under /root/reference nor installable
here (no network), and the reference holds no test or golden vector for it.  This module
restates the *published* fixed-grid algorithm (torchdiffeq `_impl/solvers.py::
FixedGridODESolver.integrate`, `_impl/fixed_grid.py::{Euler,Midpoint,RK4}`, `_impl/rk_common.py
::rk4_alt_step_func`) so that the reference's own model files import UNMODIFIED, and is pinned
by the known-answer tests in tests/test_oracle_solver.py (matrix exponential, order of
convergence, 3/8-rule vs classic RK4 discrimination).

Only the call shape the reference uses is supported (call sites
`ode_transformer_gpt.py:571-578`, `macaron.py:323,326`):
    odeint(func, y0, t, method="euler"|"rk4"|"midpoint")   # no options, no adjoint

Semantics restated:
  * the solver grid IS `t` (no `step_size` option);  `dt = t[i+1] - t[i]` is computed per
    step in the dtype of `t` -- it is NOT a constant h;
  * `solution[0] = y0`; after every step `solution[j] = y1` (the linear interpolation of the
    original degenerates to y1 because the output time equals the step end);
  * `func` receives `t` cast to `y`'s dtype;
  * "rk4" is the 3/8-rule ("rk4_alt"), not the classic tableau;
  * rtol/atol are accepted and ignored by fixed-grid methods.

Nothing under odevit_b200/ may import this file.
"""
import torch

__all__ = ["odeint", "FIXED_GRID_METHODS"]
__version__ = "0.0-oracle-restatement"

_THIRD = 1.0 / 3.0
_TWO_THIRDS = 2.0 / 3.0


def _euler_increment(f, t0, dt, t1, y):
    # fixed_grid.py::Euler._step_func -> dt * f(t0, y)
    return dt * f(t0, y)


def _midpoint_increment(f, t0, dt, t1, y):
    # fixed_grid.py::Midpoint._step_func
    half = 0.5 * dt
    y_mid = y + f(t0, y) * half
    return dt * f(t0 + half, y_mid)


def _rk4_38_increment(f, t0, dt, t1, y):
    # rk_common.py::rk4_alt_step_func (Kutta's 3/8 rule), same association order
    k1 = f(t0, y)
    k2 = f(t0 + dt * _THIRD, y + dt * k1 * _THIRD)
    k3 = f(t0 + dt * _TWO_THIRDS, y + dt * (k2 - k1 * _THIRD))
    k4 = f(t1, y + dt * (k1 - k2 + k3))
    return (k1 + 3 * (k2 + k3) + k4) * dt * 0.125


FIXED_GRID_METHODS = {
    "euler": _euler_increment,
    "midpoint": _midpoint_increment,
    "rk4": _rk4_38_increment,
}


def odeint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None):
    if event_fn is not None or options:
        raise NotImplementedError("oracle restatement: only the reference's call shape is supported")
    if method is None:
        raise NotImplementedError("oracle restatement: adaptive default (dopri5) not restated")
    if method not in FIXED_GRID_METHODS:
        raise ValueError(f"oracle restatement: unsupported method {method!r}")
    if not torch.is_tensor(y0):
        raise NotImplementedError("oracle restatement: tuple states not restated")
    if t.ndim != 1 or len(t) < 1:
        raise ValueError("t must be one dimensional")
    dts = t[1:] - t[:-1]
    if len(dts) and not (bool((dts > 0).all()) or bool((dts < 0).all())):
        raise ValueError("t must be strictly increasing or decreasing")
    if t.device != y0.device:
        t = t.to(y0.device)

    increment = FIXED_GRID_METHODS[method]

    def f(tt, yy):  # _PerturbFunc: time is cast to the state's dtype
        return func(tt.to(yy.dtype), yy)

    rows = [y0]
    y = y0
    for i in range(len(t) - 1):
        t0, t1 = t[i], t[i + 1]
        dt = t1 - t0
        y = y + increment(f, t0, dt, t1, y)
        rows.append(y)
    return torch.stack(rows, dim=0)
