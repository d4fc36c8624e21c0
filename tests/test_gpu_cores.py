"""GPU tests of the bare tensor-core cores through the C ABI test hooks
(siren_b200_debug_linear / siren_b200_debug_wgrad), against numpy in fp64."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import siren_oracle as so
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu


def _lib():
    from siren_mri_b200 import _lib
    return _lib, _lib.load()


def _block_report(err, br=32, bc=64):
    R, C = err.shape
    e = err[: R // br * br].reshape(R // br, br, C // bc, bc).max(axis=(1, 3))
    return "max abs err per (row-block %d, col-block %d):\n%s" % (br, bc, np.array2string(e[:8], precision=3))


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
@pytest.mark.parametrize("R", [128, 512, 128 * 301])
def test_linear_core(prec, R):
    L, lib = _lib()
    rng = np.random.default_rng(R)
    A = rng.uniform(-1, 1, size=(R, 256)).astype(np.float32)
    W = (0.05 * rng.standard_normal((256, 256))).astype(np.float32)
    dA, dW = torch.from_numpy(A).cuda(), torch.from_numpy(W).cuda()
    out = torch.full((R, 256), float("nan"), device="cuda")
    scratch = torch.empty(8 * R * 256 + (1 << 20), dtype=torch.uint8, device="cuda")
    rc = lib.siren_b200_debug_linear(L.dptr(dA), L.dptr(dW), L.dptr(out), R, L.PRECISIONS[prec], L.dptr(scratch),
                                     torch.cuda.current_stream().cuda_stream)
    L.check(rc, "debug_linear")
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    if prec == "bf16":
        ref = so.bf16_round(A).astype(np.float64) @ so.bf16_round(W).astype(np.float64).T
        tol = 2e-6
    else:
        ref = A.astype(np.float64) @ W.astype(np.float64).T
        tol = 3e-5
    assert np.isfinite(got).all(), "non-finite output (rows never written?)\n" + _block_report(np.isnan(got) * 1.0)
    e = rel_l2(got, ref)
    assert e < tol, "rel err %.3e\n%s" % (e, _block_report(np.abs(got - ref)))


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
@pytest.mark.parametrize("R", [128, 1024, 128 * 301])
def test_wgrad_core(prec, R):
    L, lib = _lib()
    rng = np.random.default_rng(R + 1)
    A = rng.uniform(-1, 1, size=(R, 256)).astype(np.float32)
    B = rng.uniform(-1, 1, size=(R, 256)).astype(np.float32)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    out = torch.full((256, 256), float("nan"), device="cuda")
    scratch = torch.empty(8 * R * 256 + (1 << 20), dtype=torch.uint8, device="cuda")
    rc = lib.siren_b200_debug_wgrad(L.dptr(dA), L.dptr(dB), L.dptr(out), R, L.PRECISIONS[prec], L.dptr(scratch),
                                    torch.cuda.current_stream().cuda_stream)
    L.check(rc, "debug_wgrad")
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    if prec == "bf16":
        ref = so.bf16_round(A).astype(np.float64).T @ so.bf16_round(B).astype(np.float64)
        tol = 5e-6
    else:
        ref = A.astype(np.float64).T @ B.astype(np.float64)
        tol = 3e-5
    assert np.isfinite(got).all()
    e = rel_l2(got, ref)
    assert e < tol, "rel err %.3e\n%s" % (e, _block_report(np.abs(got - ref)))


def test_linear_core_is_linear():
    """Size-independent property at a large size: f(a x1 + x2) = a f(x1) + f(x2) up to rounding."""
    L, lib = _lib()
    R = 128 * 1024
    g = torch.Generator(device="cuda").manual_seed(0)
    x1 = torch.rand((R, 256), device="cuda", generator=g) - 0.5
    x2 = torch.rand((R, 256), device="cuda", generator=g) - 0.5
    W = 0.05 * torch.randn((256, 256), device="cuda", generator=g)
    scratch = torch.empty(8 * R * 256 + (1 << 20), dtype=torch.uint8, device="cuda")

    def f(x):
        out = torch.empty((R, 256), device="cuda")
        rc = lib.siren_b200_debug_linear(L.dptr(x), L.dptr(W), L.dptr(out), R, L.PREC_FP32, L.dptr(scratch),
                                         torch.cuda.current_stream().cuda_stream)
        L.check(rc, "debug_linear")
        return out

    lhs = f(0.5 * x1 + x2)
    rhs = 0.5 * f(x1) + f(x2)
    err = ((lhs - rhs).norm() / rhs.norm()).item()
    assert err < 5e-5, err
