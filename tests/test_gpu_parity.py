"""GPU parity tests: the native path (through the public modules / C ABI) against the numpy
oracle on seeded inputs and directly against the golden fixtures of the live reference.

Tolerances (north star): fp32-parity mode rel-L2 <= 1e-4 on outputs and gradients.  The bf16
fast mode is checked against its own documented bound: <= 2e-2 on outputs/gradients.
"""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import siren_oracle as so
from tests.helpers import (case_inputs, check_grads, load_golden, loss_adjoints_gradmse, loss_adjoints_lapmse,
                           rel_l2)

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16": 2e-2}


def native_model(d, o, Ws, bs, precision, coord_derivs=0, tasks=0):
    from siren_mri_b200 import modules
    m = modules.SingleBVPNet(out_features=o, type="sine", in_features=d, hidden_features=256, num_hidden_layers=3,
                             precision=precision, coord_derivs=coord_derivs).cuda()
    if not tasks:
        sd = OrderedDict()
        for l, (W, b) in enumerate(zip(Ws, bs)):
            sd["net.net.%d.0.weight" % l] = torch.from_numpy(W)
            sd["net.net.%d.0.bias" % l] = torch.from_numpy(b)
        m.load_state_dict(sd)
    return m


def oracle64(x, Ws, bs, order):
    W64 = [w.astype(np.float64) for w in Ws]
    b64 = [b.astype(np.float64) for b in bs]
    return so.siren_forward(x.astype(np.float64), W64, b64, 30.0, order=order), W64


def model_grads(m):
    dWs = [m.net.net[l][0].weight.grad.detach().cpu().numpy() for l in range(5)]
    dbs = [m.net.net[l][0].bias.grad.detach().cpu().numpy() if m.net.net[l][0].bias.grad is not None
           else np.zeros(m.net.net[l][0].bias.shape, np.float32) for l in range(5)]
    return dWs, dbs


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["img_d2_o1", "sdf_d3_o1", "vec_d2_o3"])
def test_forward_value(name, prec):
    g = load_golden(name, "f32")
    d, o, n, Ws, bs, x = case_inputs(g)
    m = native_model(d, o, Ws, bs, prec)
    with torch.no_grad():
        out = m({"coords": torch.from_numpy(x).cuda()})
    y = out["model_out"].cpu().numpy()
    (yo, _, _, _), _ = oracle64(x, Ws, bs, 0)
    assert y.shape == yo.shape
    assert rel_l2(y, yo) < TOL[prec], rel_l2(y, yo)
    assert rel_l2(y, g["y"]) < TOL[prec]          # the live reference's own fp32 output


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["img_d2_o1", "sdf_d3_o1", "vec_d2_o3"])
def test_value_loss_backward(name, prec):
    g = load_golden(name, "f32")
    d, o, n, Ws, bs, x = case_inputs(g)
    m = native_model(d, o, Ws, bs, prec)
    m.net.coords_grad = True
    out = m({"coords": torch.from_numpy(x).cuda()})
    loss = ((out["model_out"] - torch.from_numpy(g["gt"]).cuda()) ** 2).sum() / 16384.0
    loss.backward()
    dWs, dbs = model_grads(m)
    (yo, _, _, cache), W64 = oracle64(x, Ws, bs, 0)
    _, gy = so.image_mse(yo, g["gt"].astype(np.float64))
    oW, ob, ogx = so.siren_backward(cache, W64, gy)
    for l in range(5):
        assert rel_l2(dWs[l], oW[l]) < TOL[prec], (l, rel_l2(dWs[l], oW[l]))
        assert rel_l2(dbs[l], ob[l]) < TOL[prec], (l, rel_l2(dbs[l], ob[l]))
    assert rel_l2(out["model_in"].grad.cpu().numpy(), ogx) < TOL[prec]
    check_grads("mse", g, dWs, dbs, TOL[prec])     # against the live reference's record


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["img_d2_o1", "sdf_d3_o1", "vec_d2_o3"])
def test_gradient_and_laplace_queries(name, prec):
    """diff_operators.gradient / laplace (unchanged call pattern) answered by the jets."""
    from siren_mri_b200 import diff_operators
    g = load_golden(name, "f32")
    d, o, n, Ws, bs, x = case_inputs(g)
    m = native_model(d, o, Ws, bs, prec, coord_derivs=2)
    out = m({"coords": torch.from_numpy(x).cuda()})
    grad = diff_operators.gradient(out["model_out"], out["model_in"])
    assert rel_l2(grad.detach().cpu().numpy(), g["grad"]) < TOL[prec], rel_l2(grad.detach().cpu().numpy(), g["grad"])
    if o == 1:
        lap = diff_operators.laplace(out["model_out"], out["model_in"])
        tol = TOL[prec] if prec == "fp32" else 5e-2
        assert rel_l2(lap.detach().cpu().numpy(), g["lap"]) < tol, rel_l2(lap.detach().cpu().numpy(), g["lap"])


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["img_d2_o1", "sdf_d3_o1", "vec_d2_o3"])
def test_first_order_loss_backward(name, prec):
    """loss_functions.gradients_mse pattern: loss on diff_operators.gradient, then backward."""
    from siren_mri_b200 import diff_operators
    g = load_golden(name, "f32")
    d, o, n, Ws, bs, x = case_inputs(g)
    m = native_model(d, o, Ws, bs, prec, coord_derivs=1)
    out = m({"coords": torch.from_numpy(x).cuda()})
    grad = diff_operators.gradient(out["model_out"], out["model_in"])
    loss = torch.mean((grad - torch.from_numpy(g["gt_grad"]).cuda()).pow(2).sum(-1))
    loss.backward()
    dWs, dbs = model_grads(m)
    assert abs(loss.item() - float(g["gradmse_loss"])) < 5 * TOL[prec] * abs(float(g["gradmse_loss"]))
    check_grads("gradmse", g, dWs, dbs, TOL[prec] if prec == "fp32" else 5e-2)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["img_d2_o1", "sdf_d3_o1"])
def test_laplace_loss_backward(name, prec):
    """loss_functions.laplace_mse pattern (second order)."""
    from siren_mri_b200 import diff_operators
    g = load_golden(name, "f32")
    d, o, n, Ws, bs, x = case_inputs(g)
    m = native_model(d, o, Ws, bs, prec, coord_derivs=2)
    out = m({"coords": torch.from_numpy(x).cuda()})
    lap = diff_operators.laplace(out["model_out"], out["model_in"])
    loss = torch.mean((lap - torch.from_numpy(g["gt_lap"]).cuda()) ** 2)
    loss.backward()
    dWs, dbs = model_grads(m)
    check_grads("lapmse", g, dWs, dbs, TOL[prec] if prec == "fp32" else 1e-1)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_sdf_loss_backward(prec):
    from siren_mri_b200 import diff_operators
    from tests.test_oracle_golden import sdf_loss_torch
    g = load_golden("sdf_d3_o1", "f32")
    d, o, n, Ws, bs, x = case_inputs(g)
    m = native_model(d, o, Ws, bs, prec, coord_derivs=1)
    out = m({"coords": torch.from_numpy(x).cuda()})
    grad = diff_operators.gradient(out["model_out"], out["model_in"])
    loss = sdf_loss_torch(out["model_out"], grad, torch.from_numpy(g["sdf_gt"]).cuda(),
                          torch.from_numpy(g["sdf_normals"]).cuda())
    loss.backward()
    dWs, dbs = model_grads(m)
    check_grads("sdf", g, dWs, dbs, TOL[prec] if prec == "fp32" else 5e-2)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_per_task_weights(prec):
    """BatchLinear with [B, out, in] weights, as produced by HyperNetwork (cfg5 shape family)."""
    g = load_golden("mri_t3_d16_o2", "f32")
    T = int(g["tasks"])
    d, o, n, Ws, bs, x = case_inputs(g, tasks=T)
    m = native_model(d, o, None, None, prec, tasks=T)
    params = OrderedDict()
    for l, (W, b) in enumerate(zip(Ws, bs)):
        params["net.net.%d.0.weight" % l] = torch.from_numpy(W).cuda().requires_grad_(True)
        params["net.net.%d.0.bias" % l] = torch.from_numpy(b).cuda().requires_grad_(True)
    out = m({"coords": torch.from_numpy(x).cuda()}, params=params)
    y = out["model_out"]
    assert rel_l2(y.detach().cpu().numpy(), g["y"]) < TOL[prec]
    loss = ((y - torch.from_numpy(g["gt"]).cuda()) ** 2).sum() / 16384.0
    loss.backward()
    dWs = [params["net.net.%d.0.weight" % l].grad.cpu().numpy() for l in range(5)]
    dbs = [params["net.net.%d.0.bias" % l].grad.cpu().numpy() for l in range(5)]
    assert dWs[1].shape == (T, 256, 256)
    check_grads("mse", g, dWs, dbs, TOL[prec])


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_fourier_prologue_in_kernel(prec):
    """MRI prologue / epilogue: GaussianFourierFeatureTransform(lazy=True) hands the RAW coordinates to
    the model, the kernels build the 2 F features of the first layer on chip (forward AND the dW_0 pass); the
    data-consistency epilogue and the loss are the reference's ops.  Compared with the fp64 run of the unmodified
    reference (features.py:31-41 -> SingleBVPNet with per-sample weights -> data_consistency.py:32-47 -> MSE)."""
    from siren_mri_b200 import data_consistency, features, functional
    g = load_golden("fourier_t2_f8_o2", "f64")
    T, F, o = int(g["tasks"]), int(g["F"]), int(g["o"])
    Ws, bs = so.make_params(2 * F, 256, 3, o, seed=int(g["seed"]), tasks=T)
    tr = features.GaussianFourierFeatureTransform(num_input_channels=2, mapping_size_spatial=F, scale=21, lazy=True)
    tr.set_B(torch.from_numpy(g["B"]))
    coords = tr(torch.from_numpy(g["x"]).cuda())
    assert tuple(coords.shape) == (T, 300, 2) and coords._siren_fourier is not None      # nothing was materialised
    m = native_model(2 * F, o, None, None, prec, tasks=T)
    params = OrderedDict()
    for l, (W, b) in enumerate(zip(Ws, bs)):
        params["net.net.%d.0.weight" % l] = torch.from_numpy(W.astype(np.float32)).cuda().requires_grad_(True)
        params["net.net.%d.0.bias" % l] = torch.from_numpy(b.astype(np.float32)).cuda().requires_grad_(True)
    calls = []
    orig = functional._SirenFourierFn.apply
    functional._SirenFourierFn.apply = lambda *a: (calls.append(1), orig(*a))[1]
    try:
        out = m({"coords": coords}, params=params)
    finally:
        functional._SirenFourierFn.apply = orig
    assert calls, "the Fourier kernel path did not run"
    y = out["model_out"]
    assert rel_l2(y.detach().cpu().numpy(), g["y"]) < TOL[prec]
    dc = data_consistency.DataConsistencyInKspace(noise_lvl=None)
    y_dc = dc(y, torch.from_numpy(g["k0"]).cuda(), torch.from_numpy(g["mask"]).cuda())
    assert rel_l2(y_dc.detach().cpu().numpy(), g["y_dc"]) < TOL[prec]
    loss = ((y_dc - torch.from_numpy(g["gt"]).cuda()) ** 2).sum() / 16384.0
    loss.backward()
    dWs = [params["net.net.%d.0.weight" % l].grad.cpu().numpy() for l in range(5)]
    dbs = [params["net.net.%d.0.bias" % l].grad.cpu().numpy() for l in range(5)]
    assert dWs[0].shape == (T, 256, 2 * F)
    check_grads("mse", g, dWs, dbs, TOL[prec])
    # inference launch (no stash) gives the same numbers as the training forward
    with torch.no_grad():
        y_inf = m({"coords": tr(torch.from_numpy(g["x"]).cuda())}, params=params)["model_out"]
    assert rel_l2(y_inf.cpu().numpy(), y.detach().cpu().numpy()) < (1e-6 if prec == "fp32" else 5e-3)


@pytest.mark.parametrize("F,raw,tasks,per_task,n", [(5, 2, 2, True, 700), (3, 3, 1, False, 1000), (7, 1, 3, True, 130),
                                                    (8, 2, 1, False, 33000)])
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_fourier_prologue_general_shapes(prec, F, raw, tasks, per_task, n):
    """The Fourier prologue outside the reference's F = 8 / two-channel configuration: F = 3..8 features, 1..3 raw
    channels, shared and per-task weights, ragged n -- forward, every parameter gradient (dW_0 comes from the
    tensor-core items of the weight-gradient kernel in bf16 mode, from first_bwd in fp32-parity mode) and the
    inference launch, against the fp64 oracle on oracle-built features."""
    from siren_mri_b200 import functional as Fn
    d, o = 2 * F, 2
    Ws, bs = so.make_params(d, 256, 3, o, seed=40 + F, tasks=tasks if per_task else 0)
    rng = np.random.default_rng(50 + F)
    x = rng.uniform(-1, 1, size=(tasks, n, raw)).astype(np.float32)
    B = (21.0 * rng.standard_normal((raw, F))).astype(np.float32)
    gy = (rng.standard_normal((tasks, n, o)) / n).astype(np.float32)
    feat = so.fourier_features(x.astype(np.float64), B.astype(np.float64))
    W64 = [w.astype(np.float64) for w in Ws]
    b64 = [b.astype(np.float64) for b in bs]
    yo, _, _, cache = so.siren_forward(feat, W64, b64, 30.0, order=0)
    dWo, dbo, _ = so.siren_backward(cache, W64, gy.astype(np.float64))
    Wt = [torch.from_numpy(w.astype(np.float32)).cuda().requires_grad_(True) for w in Ws]
    bt = [torch.from_numpy(b.astype(np.float32)).cuda().requires_grad_(True) for b in bs]
    xt, Bt = torch.from_numpy(x).cuda(), torch.from_numpy(B).cuda()
    assert Fn.native_supported(xt, Wt, bt, 0, fourier=Bt)
    y = Fn.siren_mlp(xt, Wt, bt, w0=30.0, precision=prec, fourier=Bt)
    assert rel_l2(y.detach().cpu().numpy(), yo) < TOL[prec]
    y.backward(torch.from_numpy(gy).cuda())
    for l in range(5):
        assert rel_l2(Wt[l].grad.cpu().numpy(), dWo[l]) < TOL[prec], ("dW", l)
        assert rel_l2(bt[l].grad.cpu().numpy(), dbo[l]) < TOL[prec], ("db", l)
    with torch.no_grad():
        y_inf = Fn.siren_mlp(xt, Wt, bt, w0=30.0, precision=prec, fourier=Bt)
    assert rel_l2(y_inf.cpu().numpy(), yo) < TOL[prec]


@pytest.mark.parametrize("d,F,raw,tasks,per_task,n", [(60, 30, 2, 2, True, 1500),      # train_mri_neural_process_ddp.py:54
                                                     (64, 32, 2, 1, False, 20000), (40, 0, 0, 3, True, 900),
                                                     (17, 0, 0, 1, False, 4097),
                                                     (120, 60, 2, 2, True, 1300),      # :65 (two K chunks)
                                                     (256, 128, 2, 1, False, 40000),   # :77 (four K chunks, many units per pair)
                                                     (206, 103, 2, 3, True, 700),      # :130 (ragged last chunk)
                                                     (65, 0, 0, 1, False, 2000), (200, 0, 0, 2, True, 515)])
@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_wide_first_layer(prec, d, F, raw, tasks, per_task, n):
    """17..256 first-layer inputs (the MRI scripts' larger Fourier blocks), materialised (F = 0) or built on chip from
    raw coordinates.  bf16 mode: the fused value path -- one to four 64-wide K chunks of plain bf16 inputs on the tensor
    core, dW_0 as regular weight-gradient items on the input plane the forward leaves.  fp32-parity mode: the layer
    runs as one more hidden layer of the split-operand kernels on padded hi + lo input planes.  Against the fp64
    oracle, each mode to its tolerance."""
    from siren_mri_b200 import functional as Fn
    o = 2
    Ws, bs = so.make_params(d, 256, 3, o, seed=60 + d, tasks=tasks if per_task else 0)
    rng = np.random.default_rng(70 + d)
    gy = (rng.standard_normal((tasks, n, o)) / n).astype(np.float32)
    if F:
        x = rng.uniform(-1, 1, size=(tasks, n, raw)).astype(np.float32)
        B = (21.0 * rng.standard_normal((raw, F))).astype(np.float32)
        feat = so.fourier_features(x.astype(np.float64), B.astype(np.float64))
    else:
        x = rng.uniform(-1, 1, size=(tasks, n, d)).astype(np.float32)
        B, feat = None, x.astype(np.float64)
    W64 = [w.astype(np.float64) for w in Ws]
    b64 = [b.astype(np.float64) for b in bs]
    yo, _, _, cache = so.siren_forward(feat, W64, b64, 30.0, order=0)
    dWo, dbo, _ = so.siren_backward(cache, W64, gy.astype(np.float64))
    Wt = [torch.from_numpy(w.astype(np.float32)).cuda().requires_grad_(True) for w in Ws]
    bt = [torch.from_numpy(b.astype(np.float32)).cuda().requires_grad_(True) for b in bs]
    xt = torch.from_numpy(x).cuda()
    Bt = torch.from_numpy(B).cuda() if F else None
    assert Fn.native_supported(xt, Wt, bt, 0, fourier=Bt, precision=prec)
    assert not Fn.native_supported(xt, Wt, bt, 1, fourier=Bt, precision=prec)      # no jets above 16 inputs
    y = Fn.siren_mlp(xt, Wt, bt, w0=30.0, precision=prec, fourier=Bt)
    assert rel_l2(y.detach().cpu().numpy(), yo) < TOL[prec]
    y.backward(torch.from_numpy(gy).cuda())
    for l in range(5):
        assert rel_l2(Wt[l].grad.cpu().numpy(), dWo[l]) < TOL[prec], ("dW", l, rel_l2(Wt[l].grad.cpu().numpy(), dWo[l]))
        assert rel_l2(bt[l].grad.cpu().numpy(), dbo[l]) < TOL[prec], ("db", l, rel_l2(bt[l].grad.cpu().numpy(), dbo[l]))
    with torch.no_grad():
        y_inf = Fn.siren_mlp(xt, Wt, bt, w0=30.0, precision=prec, fourier=Bt)
    assert rel_l2(y_inf.cpu().numpy(), yo) < TOL[prec]


@pytest.mark.parametrize("hidden,tasks", [(64, 0), (128, 2), (200, 0)])
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_narrow_hidden_widths_run_zero_padded(prec, hidden, tasks):
    """hidden_features < 256 (train_mri_neural_process_ddp.py:98 uses 64): the module pads the weights with zeros onto
    the 256-wide kernels; outputs and gradients (in the net's own shapes) are those of the narrow net."""
    from siren_mri_b200 import functional as Fn, modules
    d, o, n = 2, 2, 3000
    T = max(tasks, 1)
    Ws, bs = so.make_params(d, hidden, 3, o, seed=80 + hidden, tasks=tasks)
    x = so.make_coords(T, n, d, seed=81)
    gy = (np.random.default_rng(82).standard_normal((T, n, o)) / n).astype(np.float32)
    W64 = [w.astype(np.float64) for w in Ws]
    b64 = [b.astype(np.float64) for b in bs]
    yo, _, _, cache = so.siren_forward(x.astype(np.float64), W64, b64, 30.0, order=0)
    dWo, dbo, _ = so.siren_backward(cache, W64, gy.astype(np.float64))
    m = modules.SingleBVPNet(out_features=o, type="sine", in_features=d, hidden_features=hidden, num_hidden_layers=3,
                             precision=prec).cuda()
    params = OrderedDict()
    for l, (W, b) in enumerate(zip(Ws, bs)):
        params["net.net.%d.0.weight" % l] = torch.from_numpy(W.astype(np.float32)).cuda().requires_grad_(True)
        params["net.net.%d.0.bias" % l] = torch.from_numpy(b.astype(np.float32)).cuda().requires_grad_(True)
    calls = []
    orig = Fn._SirenKernelFn.apply
    Fn._SirenKernelFn.apply = lambda *a: (calls.append(1), orig(*a))[1]
    try:
        y = m({"coords": torch.from_numpy(x).cuda()}, params=params)["model_out"]
    finally:
        Fn._SirenKernelFn.apply = orig
    assert calls, "the native kernels did not run"
    assert rel_l2(y.detach().cpu().numpy(), yo) < TOL[prec]
    y.backward(torch.from_numpy(gy).cuda())
    for l in range(5):
        gW = params["net.net.%d.0.weight" % l].grad
        assert tuple(gW.shape) == tuple(Ws[l].shape)
        assert rel_l2(gW.cpu().numpy(), dWo[l]) < TOL[prec], ("dW", l)
        assert rel_l2(params["net.net.%d.0.bias" % l].grad.cpu().numpy(), dbo[l]) < TOL[prec], ("db", l)


def test_sdf_grid_sampling_on_the_inference_kernel():
    """sdf_meshing.sample_sdf_grid (the sampling half of sdf_meshing.create_mesh, sdf_meshing.py:13-61): chunks of a
    dense grid through the stash-free inference launch, against the fp64 oracle on the reference's enumeration."""
    from siren_mri_b200 import sdf_meshing
    N = 40
    Ws, bs = so.make_params(3, 256, 3, 1, seed=90)
    m = native_model(3, 1, Ws, bs, "fp32")
    decoder = lambda c: m.net(c)      # noqa: E731   [M, 3] -> [M, 1], as the reference's SDFDecoder (test_sdf.py:27-45)
    vol = sdf_meshing.sample_sdf_grid(decoder, N=N, max_batch=17000)
    assert vol.shape == (N, N, N) and vol.is_cuda
    coords = sdf_meshing.grid_coords(0, N ** 3, N, "cpu").numpy()[None]
    (yo, _, _, _), _ = oracle64(coords.astype(np.float32), Ws, bs, 0)
    assert rel_l2(vol.cpu().numpy().reshape(-1), yo.reshape(-1)) < 1e-4


def test_lazy_higher_order_fallback_is_exact():
    """coord_derivs=0: a create_graph query falls back to the composed graph (any order)."""
    from siren_mri_b200 import diff_operators
    g = load_golden("img_d2_o1", "f32")
    d, o, n, Ws, bs, x = case_inputs(g)
    m = native_model(d, o, Ws, bs, "fp32", coord_derivs=0)
    out = m({"coords": torch.from_numpy(x).cuda()})
    with pytest.warns(UserWarning):
        lap = diff_operators.laplace(out["model_out"], out["model_in"])
    assert rel_l2(lap.detach().cpu().numpy(), g["lap"]) < 1e-4


def test_ragged_and_2d_inputs():
    from siren_mri_b200 import modules
    Ws, bs = so.make_params(3, 256, 3, 1, seed=5)
    m = native_model(3, 1, Ws, bs, "fp32")
    for n in (1, 127, 129, 1000):
        x = so.make_coords(1, n, 3, seed=n)
        with torch.no_grad():
            y2 = m.net(torch.from_numpy(x[0]).cuda())            # 2-D [N, 3] input (sdf_meshing.py:48-52)
        (yo, _, _, _), _ = oracle64(x, Ws, bs, 0)
        assert y2.shape == (n, 1)
        assert rel_l2(y2.cpu().numpy(), yo[0]) < 1e-4


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_full_size_properties(prec):
    """cfg2 size (262144 coords): outputs finite, deterministic, and row-independent -- a
    permutation of the coordinates permutes the outputs (size-independent property)."""
    Ws, bs = so.make_params(2, 256, 3, 1, seed=0)
    m = native_model(2, 1, Ws, bs, prec)
    n = 262144
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand((1, n, 2), device="cuda", generator=g) * 2 - 1
    perm = torch.randperm(n, device="cuda", generator=g)
    with torch.no_grad():
        y1 = m({"coords": x})["model_out"]
        y1b = m({"coords": x})["model_out"]
        y2 = m({"coords": x[:, perm]})["model_out"]
    assert torch.isfinite(y1).all()
    assert torch.equal(y1, y1b)
    assert torch.equal(y1[:, perm], y2)
    # spot check 4096 rows against the oracle
    idx = torch.arange(0, n, 64, device="cuda")
    (yo, _, _, _), _ = oracle64(x[:, idx].cpu().numpy(), Ws, bs, 0)
    assert rel_l2(y1[:, idx].cpu().numpy(), yo) < TOL[prec]


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("d,o,nh,w0,tasks", [(1, 1, 1, 30.0, 0), (2, 8, 5, 10.0, 0), (3, 2, 2, 30.0, 0),
                                              (5, 1, 3, 30.0, 0), (16, 2, 1, 30.0, 2), (4, 3, 4, 5.0, 3)])
def test_envelope_shapes_forward_backward(d, o, nh, w0, tasks, prec):
    """Edges of the native envelope: depth 1..5, in 1..16, out 1..8, other w0, shared and per-task."""
    from siren_mri_b200 import modules
    n = 700
    Ws, bs = so.make_params(d, 256, nh, o, seed=31, tasks=tasks)
    x = so.make_coords(max(tasks, 1), n, d, seed=32)
    gt = np.random.default_rng(33).uniform(-1, 1, size=(max(tasks, 1), n, o)).astype(np.float32)
    m = modules.SingleBVPNet(out_features=o, in_features=d, num_hidden_layers=nh, w0=w0, precision=prec).cuda()
    params = OrderedDict()
    for l, (W, b) in enumerate(zip(Ws, bs)):
        params["net.net.%d.0.weight" % l] = torch.from_numpy(W).cuda().requires_grad_(True)
        params["net.net.%d.0.bias" % l] = torch.from_numpy(b).cuda().requires_grad_(True)
    out = m({"coords": torch.from_numpy(x).cuda()}, params=params)
    W64 = [w.astype(np.float64) for w in Ws]
    b64 = [b.astype(np.float64) for b in bs]
    yo, _, _, cache = so.siren_forward(x.astype(np.float64), W64, b64, w0, 0)
    tol = TOL[prec] * (2.0 if nh >= 4 else 1.0)
    assert rel_l2(out["model_out"].detach().cpu().numpy(), yo) < tol
    loss = ((out["model_out"] - torch.from_numpy(gt).cuda()) ** 2).sum() / 16384.0
    loss.backward()
    _, gy = so.image_mse(yo, gt.astype(np.float64))
    oW, ob, _ = so.siren_backward(cache, W64, gy)
    for l in range(nh + 2):
        gW = params["net.net.%d.0.weight" % l].grad.cpu().numpy()
        gb = params["net.net.%d.0.bias" % l].grad.cpu().numpy()
        assert gW.shape == oW[l].shape
        assert rel_l2(gW, oW[l]) < tol, (l, rel_l2(gW, oW[l]))
        assert rel_l2(gb, ob[l]) < tol, (l, rel_l2(gb, ob[l]))


def test_unsupported_width_takes_composed_path_and_matches():
    """hidden_features != 256 is outside the native envelope: same results through composed ops."""
    from siren_mri_b200 import modules
    torch.manual_seed(0)
    m = modules.SingleBVPNet(in_features=2, out_features=1, hidden_features=64).cuda()
    x = torch.rand(1, 300, 2, device="cuda")
    y = m({"coords": x})["model_out"]
    ref = functional_composed(m, x)
    assert torch.allclose(y, ref, atol=1e-6)


def functional_composed(m, x):
    from siren_mri_b200 import functional
    ws = [m.net.net[l][0].weight for l in range(5)]
    bs = [m.net.net[l][0].bias for l in range(5)]
    return functional.composed_mlp(x, ws, bs, 30.0)


def test_retain_graph_allows_second_backward():
    from siren_mri_b200 import modules
    m = modules.SingleBVPNet(in_features=2, out_features=1, precision="fp32").cuda()
    x = torch.rand(1, 500, 2, device="cuda")
    loss = m({"coords": x})["model_out"].pow(2).mean()
    loss.backward(retain_graph=True)
    g1 = m.net.net[1][0].weight.grad.clone()
    m.zero_grad()
    loss.backward()
    assert torch.allclose(g1, m.net.net[1][0].weight.grad, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("order", [1, 2])
def test_per_task_weights_with_coordinate_jets(order, prec):
    """Hypernetwork-style per-task weights combined with coordinate derivatives (jets + their reverse)."""
    from siren_mri_b200 import functional
    T, n, d, o = 2, 300, 2, 1
    Ws, bs = so.make_params(d, 256, 3, o, seed=41, tasks=T)
    x = so.make_coords(T, n, d, seed=42)
    weights = [torch.from_numpy(w).cuda().requires_grad_(True) for w in Ws]
    biases = [torch.from_numpy(b).cuda().requires_grad_(True) for b in bs]
    coords = torch.from_numpy(x).cuda().requires_grad_(True)
    y = functional.siren_mlp(coords, weights, biases, 30.0, prec, coord_derivs=order)
    grad = torch.autograd.grad(y, [coords], torch.ones_like(y), create_graph=True)[0]
    W64 = [w.astype(np.float64) for w in Ws]
    b64 = [b.astype(np.float64) for b in bs]
    yo, J, D, cache = so.siren_forward(x.astype(np.float64), W64, b64, 30.0, order=order)
    tol = TOL[prec] if prec == "fp32" else 5e-2
    assert rel_l2(y.detach().cpu().numpy(), yo) < TOL[prec]
    assert rel_l2(grad.detach().cpu().numpy(), so.gradient(J)) < tol
    loss = (grad ** 2).mean()
    gJ = np.broadcast_to((2.0 * so.gradient(J) / so.gradient(J).size)[..., None, :], J.shape).copy()
    gD = None
    if order == 2:
        lap = sum(torch.autograd.grad(grad[..., i], coords, torch.ones_like(grad[..., i]), create_graph=True)[0][..., i:i + 1]
                  for i in range(d))
        assert rel_l2(lap.detach().cpu().numpy(), so.laplace(D)) < (tol if prec == "fp32" else 1e-1)
        loss = loss + 1e-4 * (lap ** 2).mean()
        gD = np.broadcast_to((1e-4 * 2.0 * so.laplace(D) / so.laplace(D).size)[..., None], D.shape).copy()
    loss.backward()
    oW, ob, _ = so.siren_backward(cache, W64, np.zeros_like(yo), gJ, gD)
    for l in range(5):
        e = rel_l2(weights[l].grad.cpu().numpy(), oW[l])
        assert e < (tol if prec == "fp32" else 1e-1), (l, e)


@pytest.mark.parametrize("coord_derivs", [0, 1, 2])
@pytest.mark.parametrize("name", ["img_d2_o1", "vec_d2_o3", "sdf_d3_o1"])
def test_jacobian_and_hessian_queries(name, coord_derivs):
    """diff_operators.jacobian / hessian (reference diff_operators.py:46-59, 5-24) on native outputs against the live
    reference's record: first derivatives from the jets (or the lazy composed fallback for coord_derivs=0); the full
    Hessian -- mixed terms are not in the jets -- through the composed re-evaluation hessian() switches to."""
    import warnings
    from siren_mri_b200 import diff_operators
    g = load_golden(name, "f32")
    d, o, n, Ws, bs, x = case_inputs(g)
    m = native_model(d, o, Ws, bs, "fp32", coord_derivs=coord_derivs)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = m({"coords": torch.from_numpy(x).cuda()})
        jac, status = diff_operators.jacobian(out["model_out"], out["model_in"])
        assert status == 0 and rel_l2(jac.detach().cpu().numpy(), g["jac"]) < 1e-4
        hes, status = diff_operators.hessian(out["model_out"], out["model_in"])
        assert status == 0 and rel_l2(hes.detach().cpu().numpy(), g["hess"]) < 1e-4
    assert float(hes[..., 0, 1].abs().max()) > 0          # the mixed second derivatives are there
    hes.pow(2).mean().backward()                           # and differentiable w.r.t. the parameters
    assert torch.isfinite(m.net.net[1][0].weight.grad).all()


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("outer,w_first", [(True, 30.0), (False, 30.0), (True, 12.0)])
def test_notebook_siren_alias(outer, w_first, prec):
    """modules.Siren (explore_siren.ipynb cell 3) on the native kernels: outputs and parameter gradients against the
    same layers evaluated as written (backend='composed'), for both outermost settings and unequal omegas."""
    from siren_mri_b200 import modules
    torch.manual_seed(5)
    m = modules.Siren(2, 256, 3, 1, outermost_linear=outer, first_omega_0=w_first, hidden_omega_0=30.0,
                      precision=prec).cuda()
    x = torch.rand(3000, 2, device="cuda") * 2 - 1
    gt = torch.rand(3000, 1, device="cuda")
    y, c = m(x)
    assert c.requires_grad and y.shape == (3000, 1)
    ((y - gt) ** 2).mean().backward()
    got = [p.grad.clone() for p in m.parameters()]
    m.zero_grad()
    m.backend = "composed"
    y_ref, _ = m(x)
    ((y_ref - gt) ** 2).mean().backward()
    tol = TOL[prec]
    assert rel_l2(y.detach().cpu().numpy(), y_ref.detach().cpu().numpy()) < tol
    for a, p in zip(got, m.parameters()):
        assert rel_l2(a.cpu().numpy(), p.grad.cpu().numpy()) < tol
    # coordinate derivatives through the notebook's own operators (cell 5 = diff_operators.gradient / laplace)
    from siren_mri_b200 import diff_operators
    m.backend, m.coord_derivs = "auto", 2
    y2, c2 = m(x)
    lap = diff_operators.laplace(y2, c2)
    m.backend = "composed"
    y3, c3 = m(x)
    lap_ref = diff_operators.laplace(y3, c3)
    if outer:      # a sine on top of the jets is ordinary autograd on [N, 1]: covered by the linear case
        assert rel_l2(lap.detach().cpu().numpy(), lap_ref.detach().cpu().numpy()) < (1e-4 if prec == "fp32" else 1e-1)
