"""Pins oracle/siren_oracle.py against fixtures produced by the live reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import siren_oracle as so
from tests.helpers import (case_inputs, check_grads, load_golden, loss_adjoints_gradmse,
                           loss_adjoints_lapmse, rel_l2)

TOL64 = 1e-9


@pytest.mark.parametrize("name", ["img_d2_o1", "sdf_d3_o1", "vec_d2_o3"])
def test_forward_and_jets_match_reference(name):
    g = load_golden(name, "f64")
    d, o, n, Ws, bs, x = case_inputs(g)
    W64 = [w.astype(np.float64) for w in Ws]
    b64 = [b.astype(np.float64) for b in bs]
    y, J, D, _ = so.siren_forward(x.astype(np.float64), W64, b64, 30.0, order=2)
    assert rel_l2(y, g["y"]) < TOL64
    assert rel_l2(so.gradient(J), g["grad"]) < TOL64
    if o == 1:
        assert rel_l2(so.laplace(D), g["lap"]) < TOL64


@pytest.mark.parametrize("name", ["img_d2_o1", "sdf_d3_o1", "vec_d2_o3"])
def test_backward_matches_reference(name):
    g = load_golden(name, "f64")
    d, o, n, Ws, bs, x = case_inputs(g)
    W64 = [w.astype(np.float64) for w in Ws]
    b64 = [b.astype(np.float64) for b in bs]
    x64 = x.astype(np.float64)
    # value loss
    y, _, _, cache = so.siren_forward(x64, W64, b64, 30.0, order=0)
    loss, gy = so.image_mse(y, g["gt"].astype(np.float64))
    assert abs(loss - float(g["mse_loss"])) < 1e-9 * max(1.0, abs(loss))
    dWs, dbs, gx = so.siren_backward(cache, W64, gy)
    check_grads("mse", g, dWs, dbs, TOL64)
    assert rel_l2(gx, g["mse_gx"]) < TOL64
    # first-order loss
    y, J, _, cache = so.siren_forward(x64, W64, b64, 30.0, order=1)
    gJ = loss_adjoints_gradmse(J, g["gt_grad"].astype(np.float64))
    dWs, dbs, _ = so.siren_backward(cache, W64, np.zeros_like(y), gJ)
    check_grads("gradmse", g, dWs, dbs, TOL64)
    # second-order loss
    if o == 1:
        y, J, D, cache = so.siren_forward(x64, W64, b64, 30.0, order=2)
        gD = loss_adjoints_lapmse(D, g["gt_lap"].astype(np.float64))
        dWs, dbs, _ = so.siren_backward(cache, W64, np.zeros_like(y), np.zeros_like(J), gD)
        check_grads("lapmse", g, dWs, dbs, TOL64)


def sdf_loss_torch(y, grad, gt_sdf, gt_normals):
    """loss_functions.py:460-484 restated on (value, gradient) tensors."""
    F = torch.nn.functional
    sdf_c = torch.where(gt_sdf != -1, y, torch.zeros_like(y))
    inter = torch.where(gt_sdf != -1, torch.zeros_like(y), torch.exp(-1e2 * torch.abs(y)))
    normal = torch.where(gt_sdf != -1, 1 - F.cosine_similarity(grad, gt_normals, dim=-1)[..., None],
                         torch.zeros_like(grad[..., :1]))
    gc = torch.abs(grad.norm(dim=-1) - 1)
    return (torch.abs(sdf_c).mean() * 3e3 + inter.mean() * 1e2 + normal.mean() * 1e2 + gc.mean() * 5e1)


def test_sdf_loss_gradients_match_reference():
    g = load_golden("sdf_d3_o1", "f64")
    d, o, n, Ws, bs, x = case_inputs(g)
    W64 = [w.astype(np.float64) for w in Ws]
    b64 = [b.astype(np.float64) for b in bs]
    y, J, _, cache = so.siren_forward(x.astype(np.float64), W64, b64, 30.0, order=1)
    ty = torch.from_numpy(y).requires_grad_(True)
    tg = torch.from_numpy(so.gradient(J)).requires_grad_(True)
    loss = sdf_loss_torch(ty, tg, torch.from_numpy(g["sdf_gt"]).double(), torch.from_numpy(g["sdf_normals"]).double())
    assert abs(loss.item() - float(g["sdf_loss"])) < 1e-9 * abs(loss.item())
    loss.backward()
    gJ = tg.grad.numpy()[..., None, :]
    dWs, dbs, _ = so.siren_backward(cache, W64, ty.grad.numpy(), gJ)
    check_grads("sdf", g, dWs, dbs, TOL64)


def test_per_task_weights_match_reference():
    g = load_golden("mri_t3_d16_o2", "f64")
    d, o, n, Ws, bs, x = case_inputs(g, tasks=int(g["tasks"]))
    W64 = [w.astype(np.float64) for w in Ws]
    b64 = [b.astype(np.float64) for b in bs]
    y, _, _, cache = so.siren_forward(x.astype(np.float64), W64, b64, 30.0, order=0)
    assert rel_l2(y, g["y"]) < TOL64
    loss, gy = so.image_mse(y, g["gt"].astype(np.float64))
    dWs, dbs, _ = so.siren_backward(cache, W64, gy)
    assert dWs[1].shape == (3, 256, 256)
    check_grads("mse", g, dWs, dbs, TOL64)


def test_fp32_reference_noise_floor():
    """The fp32 fixtures differ from fp64 only at rounding level."""
    a, b = load_golden("img_d2_o1", "f32"), load_golden("img_d2_o1", "f64")
    assert rel_l2(a["y"], b["y"]) < 1e-5
    assert rel_l2(a["grad"], b["grad"]) < 1e-4
    assert rel_l2(a["lap"], b["lap"]) < 1e-4


@pytest.mark.parametrize("clip", [0, 1])
def test_adam_matches_torch(clip, golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "adam_clip%d.npz" % clip))
    p = g["p0"].astype(np.float32)
    m = np.zeros_like(p)
    v = np.zeros_like(p)
    for step, grad in enumerate(g["grads"], start=1):
        grad = grad.astype(np.float32)
        if clip:
            coef, _ = so.clip_coef([grad], float(g["clip"]))
            grad = (grad * np.float32(coef)).astype(np.float32)
        p, m, v = so.adam_step(p, grad, m, v, step, lr=float(g["lr"]))
        p, m, v = p.astype(np.float32), m.astype(np.float32), v.astype(np.float32)
        assert np.abs(p - g["traj"][step - 1]).max() < 2e-6


def test_bf16_round_matches_torch():
    a = np.random.default_rng(0).standard_normal(4096).astype(np.float32)
    ref = torch.from_numpy(a).to(torch.bfloat16).float().numpy()
    assert np.array_equal(so.bf16_round(a), ref)


def test_ref_port_matches_reference():
    """oracle/siren_ref_port.py (the CPU port bench.py times) against the live reference's record."""
    from oracle.siren_ref_port import RefPortSiren
    g = load_golden("img_d2_o1", "f64")
    d, o, n, Ws, bs, x = case_inputs(g)
    m = RefPortSiren(d_in=d, d_out=o).double()
    m.load_numpy(Ws, bs)
    out = m({"coords": torch.from_numpy(x).double()})
    assert rel_l2(out["model_out"].detach().numpy(), g["y"]) < 1e-12
    loss = ((out["model_out"] - torch.from_numpy(g["gt"]).double()) ** 2).sum() / 16384.0
    loss.backward()
    dWs = [lin.weight.grad.numpy() for lin in m.lin]
    dbs = [lin.bias.grad.numpy() for lin in m.lin]
    check_grads("mse", g, dWs, dbs, 1e-10)


def _fourier_case(tag):
    g = load_golden("fourier_t2_f8_o2", tag)
    tasks, F, o = int(g["tasks"]), int(g["F"]), int(g["o"])
    Ws, bs = so.make_params(2 * F, 256, 3, o, seed=int(g["seed"]), tasks=tasks)
    return g, tasks, F, o, Ws, bs


def test_fourier_features_and_data_consistency_match_reference():
    """features.py:31-41 and data_consistency.py:32-47 (the MRI prologue / epilogue), plus the whole chain
    features -> per-task SIREN -> DC -> MSE with its parameter gradients."""
    g, tasks, F, o, Ws, bs = _fourier_case("f64")
    x64, B64 = g["x"].astype(np.float64), g["B"].astype(np.float64)
    feat = so.fourier_features(x64, B64)
    assert rel_l2(feat, g["feat"]) < TOL64
    W64 = [w.astype(np.float64) for w in Ws]
    b64 = [b.astype(np.float64) for b in bs]
    y, _, _, cache = so.siren_forward(feat, W64, b64, 30.0, order=0)
    assert rel_l2(y, g["y"]) < TOL64
    k0, mask = g["k0"].astype(np.float64), g["mask"].astype(np.float64)
    y_dc = so.data_consistency(y, k0, mask)
    assert rel_l2(y_dc, g["y_dc"]) < TOL64
    # loss = sum (y_dc - gt)^2 / 16384: the DC passes (1 - mask) of the adjoint on to the network output
    mr = np.transpose(mask, (0, 2, 3, 1)).reshape(tasks, -1, 2)
    gy = 2.0 * (y_dc - g["gt"].astype(np.float64)) / 16384.0 * (1 - mr)
    dWs, dbs, _ = so.siren_backward(cache, W64, gy)
    check_grads("mse", g, dWs, dbs, TOL64)
