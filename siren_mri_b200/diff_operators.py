"""Differential operators on ``model_out`` w.r.t. ``model_in`` -- same call signatures as the
reference's diff_operators.py (gradient :39-43, divergence :32-36, laplace :27-29,
jacobian :46-59, hessian :5-24).  They are plain autograd queries; on outputs of the native path
the queries are answered by the forward-mode jets the kernels computed (functional.py), so no
second or third autograd graph through the MLP is built.  The reference's own diff_operators
module works unchanged on the same outputs; this file exists so that users of this package do
not need the reference on their path.
"""
import torch

from .functional import composed_of


def gradient(y, x, grad_outputs=None):
    if grad_outputs is None:
        grad_outputs = torch.ones_like(y)
    return torch.autograd.grad(y, [x], grad_outputs=grad_outputs, create_graph=True)[0]


def divergence(y, x):
    div = 0.0
    for i in range(y.shape[-1]):
        gi = torch.autograd.grad(y[..., i], x, torch.ones_like(y[..., i]), create_graph=True)[0]
        div = div + gi[..., i:i + 1]
    return div


def laplace(y, x):
    return divergence(gradient(y, x), x)


def jacobian(y, x):
    """[B, N, o, d] Jacobian and a status flag (-1 if any NaN), like the reference."""
    B, N = y.shape[:2]
    jac = torch.zeros(B, N, y.shape[-1], x.shape[-1], device=y.device, dtype=y.dtype)
    for i in range(y.shape[-1]):
        yi = y[..., i].reshape(-1, 1)
        jac[:, :, i, :] = torch.autograd.grad(yi, x, torch.ones_like(yi), create_graph=True)[0]
    status = -1 if torch.any(torch.isnan(jac)) else 0
    return jac, status


def hessian(y, x):
    """[B, N, o, d, d] Hessian and a status flag.

    Mixed second derivatives are not in the native path's jets (they carry d/dx_k and d2/dx_k^2): when ``y`` came
    from the jet path (``coord_derivs`` 1 or 2) it is re-evaluated as the composed PyTorch graph on the same
    coordinate leaf first, so the result is the reference's (diff_operators.py:5-24), never a Hessian with zeroed
    off-diagonals.  With ``coord_derivs=0`` the query falls back to the composed graph by itself."""
    y = composed_of(y)
    B, N = y.shape[:2]
    ones = torch.ones_like(y[..., 0])
    h = torch.zeros(B, N, y.shape[-1], x.shape[-1], x.shape[-1], device=y.device, dtype=y.dtype)
    for i in range(y.shape[-1]):
        dydx = torch.autograd.grad(y[..., i], x, ones, create_graph=True)[0]
        for j in range(x.shape[-1]):
            h[..., i, j, :] = torch.autograd.grad(dydx[..., j], x, ones, create_graph=True)[0]
    status = -1 if torch.any(torch.isnan(h)) else 0
    return h, status
