"""CPU oracle for the SIREN hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the arithmetic the reference performs on the
hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it; the product package
(``siren_mri_b200``) never does.

Parity status: PINNED against the live reference.  ``tests/golden/make_golden.py``
imports the unmodified reference from ``/root/reference`` in the dev container and
writes small fixtures; ``tests/test_oracle_golden.py`` checks every function in
this file against them (the reference holds no golden vectors of its own for this
path).

What each function follows (all paths relative to /root/reference):

* ``linear``            modules.py:16-27    BatchLinear.forward (matmul with the
                                            weight's last two dims swapped, then
                                            ``+= bias.unsqueeze(-2)``)
* ``sine``              modules.py:35-38    Sine.forward  sin(w0 * x)
* ``siren_forward``     modules.py:92-97, 146-164 and
                        torchmeta/modules/container.py:9-19  (the sequential
                        BatchLinear/Sine chain, outermost layer linear)
                        plus the coordinate derivatives that
                        diff_operators.py:27-43 obtain by double backward: here
                        they are propagated in forward mode (value, d/dx_k,
                        d2/dx_k^2), SURVEY.md appendix A.
* ``siren_backward``    the autograd backward of the above that training.py:91
                        triggers.
* ``gradient`` / ``laplace``   diff_operators.py:39-43 / 27-36 expressed on jets.
* ``adam_step``         torch.optim.Adam defaults as constructed at
                        training.py:23, with the global-norm clip of
                        training.py:93-97.
* ``fourier_features``  features.py:31-41 (GaussianFourierFeatureTransform.forward:
                        ``x @ B``, times 2 pi, ``cat[sin, cos]`` on the last axis).
* ``data_consistency``  data_consistency.py:7-20 with the k0 / mask layout change of
                        DataConsistencyInKspace.forward (data_consistency.py:32-47).
* ``image_mse_grad`` / ``mse``  loss_functions.py:66-96 (high_freq=False) and
                        loss_functions.py:326-327.

Shapes.  ``x`` is ``[T, N, d]`` (T tasks, N coordinates).  Shared weights are
``W[l]: [out, in]``, ``b[l]: [out]``; per-task weights (hypernetwork case,
meta_modules.py:50-54) are ``W[l]: [T, out, in]``, ``b[l]: [T, out]``.
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------- #
# layer primitives
# --------------------------------------------------------------------------- #
def linear(h, W, b):
    """modules.py:25-26: ``input.matmul(weight^T) + bias.unsqueeze(-2)``."""
    out = np.matmul(h, np.swapaxes(W, -1, -2))
    return out + np.expand_dims(b, -2)


def sine(z, w0):
    """modules.py:38."""
    return np.sin(w0 * z)


def _wT(W):
    return np.swapaxes(W, -1, -2)


# --------------------------------------------------------------------------- #
# forward with forward-mode jets
# --------------------------------------------------------------------------- #
def siren_forward(x, Ws, bs, w0=30.0, order=0):
    """Value (and coordinate jets) of the sine MLP.

    Returns ``(y, J, D, cache)`` with ``y: [T,N,o]``, ``J: [T,N,o,d]`` =
    dy_o/dx_k (order>=1), ``D: [T,N,o,d]`` = d2y_o/dx_k^2 (order==2).
    """
    x = np.asarray(x)
    T, N, d = x.shape
    L = len(Ws) - 1                       # index of the outermost (linear) layer
    dt = x.dtype
    h = x
    Jp = Dp = None                        # jets of the layer input: [d, T, N, feat]
    if order >= 1:
        Jp = np.zeros((d, T, N, d), dt)
        for k in range(d):
            Jp[k, :, :, k] = 1.0
    if order >= 2:
        Dp = np.zeros((d, T, N, d), dt)
    cache = {"x": x, "h": [], "J": [], "D": [], "s": [], "c": [], "Jz": [], "Dz": [],
             "order": order, "w0": w0}
    for l in range(L + 1):
        cache["h"].append(h)
        cache["J"].append(Jp)
        cache["D"].append(Dp)
        z = linear(h, Ws[l], bs[l])
        Jz = np.matmul(Jp, _wT(Ws[l])) if order >= 1 else None
        Dz = np.matmul(Dp, _wT(Ws[l])) if order >= 2 else None
        if l == L:
            y = z
            J = np.moveaxis(Jz, 0, -1) if order >= 1 else None     # [T,N,o,d]
            D = np.moveaxis(Dz, 0, -1) if order >= 2 else None
            break
        s = np.sin(w0 * z)
        c = np.cos(w0 * z)
        cache["s"].append(s)
        cache["c"].append(c)
        cache["Jz"].append(Jz)
        cache["Dz"].append(Dz)
        h = s
        if order >= 2:
            Dp = w0 * c * Dz - (w0 * w0) * s * Jz * Jz
        if order >= 1:
            Jp = w0 * c * Jz
    return y, J, D, cache


# --------------------------------------------------------------------------- #
# reverse of the above
# --------------------------------------------------------------------------- #
def _sum_tasks(a, per_task):
    return a if per_task else a.sum(axis=0)


def siren_backward(cache, Ws, gy, gJ=None, gD=None):
    """Adjoint of ``siren_forward``.

    ``gy: [T,N,o]``, ``gJ, gD: [T,N,o,d]`` are the adjoints of y, J, D.  Returns
    ``(dWs, dbs, gx)`` where dW/db have the shape of the weights (summed over
    tasks when the weights are shared) and ``gx: [T,N,d]`` is the adjoint that
    reaches the coordinates through the first pre-activation.
    """
    order = cache["order"]
    w0 = cache["w0"]
    L = len(Ws) - 1
    per_task = Ws[0].ndim == 3
    dWs = [None] * (L + 1)
    dbs = [None] * (L + 1)

    # adjoints of the outermost pre-activation streams
    zb = gy
    d = cache["x"].shape[-1]
    if order >= 1 and gJ is None:
        gJ = np.zeros(gy.shape + (d,), gy.dtype)
    if order >= 2 and gD is None:
        gD = np.zeros(gy.shape + (d,), gy.dtype)
    Jzb = np.moveaxis(gJ, -1, 0) if order >= 1 else None   # [d,T,N,o]
    Dzb = np.moveaxis(gD, -1, 0) if order >= 2 else None
    for l in range(L, -1, -1):
        h = cache["h"][l]
        # linear layer: weight / bias gradients
        dW = np.matmul(_wT(zb), h)                                    # [T,out,in]
        if Jzb is not None:
            dW = dW + np.matmul(_wT(Jzb), cache["J"][l]).sum(axis=0)
        if Dzb is not None:
            dW = dW + np.matmul(_wT(Dzb), cache["D"][l]).sum(axis=0)
        dWs[l] = _sum_tasks(dW, per_task)
        dbs[l] = _sum_tasks(zb.sum(axis=-2), per_task)
        # adjoints of the layer inputs
        W = Ws[l]
        hb = np.matmul(zb, W)
        Jb = np.matmul(Jzb, W) if Jzb is not None else None
        Db = np.matmul(Dzb, W) if Dzb is not None else None
        if l == 0:
            gx = hb
            break
        # sine of layer l-1
        s, c = cache["s"][l - 1], cache["c"][l - 1]
        Jz, Dz = cache["Jz"][l - 1], cache["Dz"][l - 1]
        zb = w0 * c * hb
        if Jb is not None:
            zb = zb - (w0 * w0) * s * (Jz * Jb).sum(axis=0)
        if Db is not None:
            zb = zb - (w0 * w0) * s * (Dz * Db).sum(axis=0) - (w0 ** 3) * c * (Jz * Jz * Db).sum(axis=0)
        Jzb_new = None
        if Jb is not None or Db is not None:
            Jzb_new = 0.0
            if Jb is not None:
                Jzb_new = Jzb_new + w0 * c * Jb
            if Db is not None:
                Jzb_new = Jzb_new - 2.0 * (w0 * w0) * s * Jz * Db
        Dzb = (w0 * c * Db) if Db is not None else None
        Jzb = Jzb_new
    return dWs, dbs, gx


# --------------------------------------------------------------------------- #
# the queries diff_operators makes, expressed on jets
# --------------------------------------------------------------------------- #
def gradient(J, grad_outputs=None):
    """diff_operators.py:39-43: vector-Jacobian product with ``ones_like(y)``."""
    if grad_outputs is None:
        return J.sum(axis=-2)
    return (grad_outputs[..., None] * J).sum(axis=-2)


def laplace(D):
    """diff_operators.py:27-36 (divergence of the gradient; diagonal terms only)."""
    return D.sum(axis=-2).sum(axis=-1, keepdims=True)


# --------------------------------------------------------------------------- #
# losses used by the five configurations (value and dL/dy)
# --------------------------------------------------------------------------- #
def image_mse(y, gt, weight=1.0 / (128 * 128)):
    """loss_functions.py:66-96 with high_freq=False: sum((y-gt)^2) * 1/16384."""
    diff = y - gt
    return float((diff * diff).sum() * weight), 2.0 * weight * diff


def mse(y, gt):
    """loss_functions.py:326-327 style mean squared error."""
    diff = y - gt
    return float((diff * diff).mean()), 2.0 * diff / diff.size


# --------------------------------------------------------------------------- #
# optimizer tail
# --------------------------------------------------------------------------- #
def clip_coef(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ (training.py:93-97): global L2 norm."""
    total = np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads))
    return min(1.0, max_norm / (total + 1e-6)), total


def adam_step(p, g, m, v, step, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults (training.py:23); ``step`` counts from 1."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = np.sqrt(v) / np.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v


# --------------------------------------------------------------------------- #
# emulation of the operand rounding of the tensor-core modes (documentation aid)
# --------------------------------------------------------------------------- #
def bf16_round(a):
    """Round-to-nearest-even of fp32 values to bfloat16, returned as fp32."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


# --------------------------------------------------------------------------- #
# deterministic synthetic parameters / inputs shared by golden script and tests
# --------------------------------------------------------------------------- #
def make_params(d_in, hidden, n_hidden, d_out, seed=0, tasks=0, w0=30.0, perturb=0.0):
    """Weights with the distributions of modules.py:641-654 (sine_init /
    first_layer_sine_init) and nn.Linear's default bias init, drawn from a numpy
    PCG64 stream so fixtures do not depend on torch's RNG.  ``tasks>0`` adds a
    leading task axis (each task an independent draw)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    dims = [d_in] + [hidden] * (n_hidden + 1) + [d_out]
    Ws, bs = [], []
    lead = (tasks,) if tasks else ()
    for l in range(len(dims) - 1):
        fi, fo = dims[l], dims[l + 1]
        bound = (1.0 / fi) if l == 0 else (np.sqrt(6.0 / fi) / 30.0)
        W = rng.uniform(-bound, bound, size=lead + (fo, fi))
        b = rng.uniform(-1.0 / np.sqrt(fi), 1.0 / np.sqrt(fi), size=lead + (fo,))
        if perturb:
            W = W + perturb * rng.standard_normal(W.shape)
        Ws.append(W.astype(np.float32))
        bs.append(b.astype(np.float32))
    return Ws, bs


def make_coords(tasks, n, d, seed=1):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.uniform(-1.0, 1.0, size=(tasks, n, d)).astype(np.float32)


# --------------------------------------------------------------------------- #
# MRI prologue / epilogue (neural-process models)
# --------------------------------------------------------------------------- #
def fourier_features(x, B):
    """features.py:35-41: ``x @ B`` -> ``2 pi x`` -> ``cat([sin, cos], dim=2)``."""
    u = 2.0 * np.pi * np.matmul(x, B)
    return np.concatenate([np.sin(u), np.cos(u)], axis=-1)


def data_consistency(pred, k0, mask, noise_lvl=None):
    """data_consistency.py:32-47 + :7-20.  ``pred`` [B, N, 2]; ``k0`` / ``mask`` [B, 2, nx, ny] are permuted to
    [B, nx, ny, 2] and flattened to [B, N, 2] first."""
    Bn = k0.shape[0]
    k0r = np.transpose(k0, (0, 2, 3, 1)).reshape(Bn, -1, 2)
    mr = np.transpose(mask, (0, 2, 3, 1)).reshape(Bn, -1, 2)
    if noise_lvl:
        return (1 - mr) * pred + mr * (pred + noise_lvl * k0r) / (1 + noise_lvl)
    return (1 - mr) * pred + mr * k0r
