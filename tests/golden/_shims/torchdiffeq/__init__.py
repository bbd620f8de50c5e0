"""Stand-in for torchdiffeq==0.2.3 (un-vendored dependency of the reference,
requirement.txt:44).  For method='euler' on the grid t=[0, dt] the fixed-grid
solver takes exactly one explicit Euler step y1 = y0 + (t1 - t0) * f(t0, y0)
and returns the stack [y0, y1] (call sites: envs/models/F16_model.py:64-67)."""
import torch


def odeint_adjoint(func, y0, t, method="euler", **kw):
    assert method == "euler" and t.numel() == 2
    with torch.no_grad():
        y1 = y0 + (t[1] - t[0]) * func(t[0], y0)
    return torch.stack((y0, y1))


odeint = odeint_adjoint
