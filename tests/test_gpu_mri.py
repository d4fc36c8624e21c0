"""GPU parity tests of the MRI neural-process pieces around the sine MLP: the k-space
data-consistency epilogue applied inside the kernels (siren_b200_forward_dc / _backward_dc / _forward_dc_mse) and the
hypernetwork head that emits the per-task weights.  Against the golden record of the unmodified reference's
features -> per-task SIREN -> data consistency -> MSE chain and against the fp64 oracle.

Tolerances: fp32-parity mode rel-L2 <= 1e-4; bf16 mode its documented bound.
"""
import ctypes
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import siren_oracle as so
from tests.helpers import check_grads, load_golden, rel_l2

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16": 2e-2}


def _golden_model(prec, fuse_dc):
    from siren_mri_b200 import features, modules
    g = load_golden("fourier_t2_f8_o2", "f64")
    T, F, o = int(g["tasks"]), int(g["F"]), int(g["o"])
    Ws, bs = so.make_params(2 * F, 256, 3, o, seed=int(g["seed"]), tasks=T)
    tr = features.GaussianFourierFeatureTransform(num_input_channels=2, mapping_size_spatial=F, scale=21, lazy=True)
    tr.set_B(torch.from_numpy(g["B"]))
    m = modules.SingleBVPNet(out_features=o, type="sine", in_features=2 * F, hidden_features=256, num_hidden_layers=3,
                             precision=prec).cuda()
    m.fuse_dc = fuse_dc
    params = OrderedDict()
    for l, (W, b) in enumerate(zip(Ws, bs)):
        params["net.net.%d.0.weight" % l] = torch.from_numpy(W.astype(np.float32)).cuda().requires_grad_(True)
        params["net.net.%d.0.bias" % l] = torch.from_numpy(b.astype(np.float32)).cuda().requires_grad_(True)
    return g, tr, m, params


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_data_consistency_epilogue_in_kernel(prec):
    """The reference's chain features.py:31-41 -> per-task SingleBVPNet -> data_consistency.py:32-47 -> MSE with BOTH
    ends inside the kernels: Fourier features built on chip, the blend applied where a row's output is completed, the
    output adjoint masked where the input-gradient chain picks it up.  DataConsistencyInKspace passes the tagged
    prediction on.  Against the fp64 run of the unmodified reference."""
    from siren_mri_b200 import data_consistency, functional
    g, tr, m, params = _golden_model(prec, True)
    k0, mask = torch.from_numpy(g["k0"]).cuda(), torch.from_numpy(g["mask"]).cuda()
    calls = []
    orig = functional._SirenDCFn.apply
    functional._SirenDCFn.apply = lambda *a: (calls.append(1), orig(*a))[1]
    try:
        out = m({"coords": tr(torch.from_numpy(g["x"]).cuda()), "img_sparse": k0, "dc_mask": mask}, params=params)
    finally:
        functional._SirenDCFn.apply = orig
    assert calls, "the data-consistency kernel path did not run"
    y_dc = out["model_out"]
    assert y_dc._siren_dc_done == 0.0
    dc = data_consistency.DataConsistencyInKspace(noise_lvl=None)
    again = dc(y_dc, k0, mask)
    assert again is y_dc                                   # nothing left to do for the module
    assert rel_l2(y_dc.detach().cpu().numpy(), g["y_dc"]) < TOL[prec]
    loss = ((again - torch.from_numpy(g["gt"]).cuda()) ** 2).sum() / 16384.0
    assert abs(float(loss.detach()) - float(g["mse_loss"])) <= (1e-4 if prec == "fp32" else 3e-2) * float(g["mse_loss"])
    loss.backward()
    dWs = [params["net.net.%d.0.weight" % l].grad.cpu().numpy() for l in range(5)]
    dbs = [params["net.net.%d.0.bias" % l].grad.cpu().numpy() for l in range(5)]
    check_grads("mse", g, dWs, dbs, TOL[prec])
    # a module configured for another noise level must not silently accept the tagged prediction
    with pytest.raises(RuntimeError):
        data_consistency.DataConsistencyInKspace(noise_lvl=0.5)(y_dc, k0, mask)
    # the unfused mirror gives the same numbers
    g2, tr2, m2, params2 = _golden_model(prec, False)
    y2 = dc(m2({"coords": tr2(torch.from_numpy(g["x"]).cuda()), "img_sparse": k0, "dc_mask": mask}, params=params2)["model_out"],
            k0, mask)
    assert rel_l2(y2.detach().cpu().numpy(), y_dc.detach().cpu().numpy()) < 1e-6
    with torch.no_grad():      # inference launch (no stash)
        y_inf = m({"coords": tr(torch.from_numpy(g["x"]).cuda()), "img_sparse": k0, "dc_mask": mask}, params=params)["model_out"]
    assert rel_l2(y_inf.cpu().numpy(), y_dc.detach().cpu().numpy()) < (1e-6 if prec == "fp32" else 5e-3)


CASES = [  # d, F, raw, o, tasks, per_task, n, channels_first, noise
    (2, 0, 0, 1, 1, False, 70000, False, None),       # cfg2-like single-output net, 4-5 units per CTA pair
    (16, 8, 2, 2, 3, True, 20000, True, None),        # the MRI configuration: Fourier prologue + DC, per-task weights
    (16, 0, 0, 2, 2, False, 1300, True, 0.25),        # noisy blend, shared weights, ragged n
    (3, 0, 0, 3, 2, True, 900, True, None),           # three outputs: outermost linear off the fused kernels
    (60, 30, 2, 2, 2, True, 1500, False, 2.0),        # wide first layer
]


@pytest.mark.parametrize("d,F,raw,o,tasks,per_task,n,cf,noise", CASES)
@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_dc_epilogue_general_shapes(prec, d, F, raw, o, tasks, per_task, n, cf, noise):
    """y_dc and every parameter gradient of  L = sum gy * DC(siren(x))  against the fp64 oracle: fused top / bottom
    steps (d_out <= 2, bf16), the elementwise kernels elsewhere (fp32-parity mode, d_out = 3), both layouts of
    k0 / mask, the noisy blend, many units per CTA pair."""
    from siren_mri_b200 import functional as Fn
    Ws, bs = so.make_params(d, 256, 3, o, seed=80 + d + o, tasks=tasks if per_task else 0)
    rng = np.random.default_rng(90 + d + o)
    gy = (rng.standard_normal((tasks, n, o)) / n).astype(np.float32)
    if F:
        x = rng.uniform(-1, 1, size=(tasks, n, raw)).astype(np.float32)
        B = (21.0 * rng.standard_normal((raw, F))).astype(np.float32)
        feat = so.fourier_features(x.astype(np.float64), B.astype(np.float64))
    else:
        x = rng.uniform(-1, 1, size=(tasks, n, d)).astype(np.float32)
        B, feat = None, x.astype(np.float64)
    k0 = rng.standard_normal((tasks, n, o)).astype(np.float32)             # [B, N, o] view of the samples / mask
    mask = (rng.uniform(size=(tasks, n, o)) < 0.3).astype(np.float32)
    pull = 1.0 if not noise else noise / (1.0 + noise)
    W64 = [w.astype(np.float64) for w in Ws]
    b64 = [b.astype(np.float64) for b in bs]
    yo, _, _, cache = so.siren_forward(feat, W64, b64, 30.0, order=0)
    ydc_o = yo + mask * pull * (k0 - yo)
    dWo, dbo, _ = so.siren_backward(cache, W64, gy.astype(np.float64) * (1.0 - mask * pull))
    Wt = [torch.from_numpy(w.astype(np.float32)).cuda().requires_grad_(True) for w in Ws]
    bt = [torch.from_numpy(b.astype(np.float32)).cuda().requires_grad_(True) for b in bs]
    xt = torch.from_numpy(x).cuda()
    Bt = None if B is None else torch.from_numpy(B).cuda()
    lay = (lambda a: np.ascontiguousarray(np.transpose(a, (0, 2, 1)))) if cf else (lambda a: a)
    k0t, mt = torch.from_numpy(lay(k0)).cuda(), torch.from_numpy(lay(mask)).cuda()
    y = Fn.siren_mlp(xt, Wt, bt, w0=30.0, precision=prec, fourier=Bt, dc=(k0t, mt, noise, cf))
    assert rel_l2(y.detach().cpu().numpy(), ydc_o) < TOL[prec]
    sampled = mask > 0
    if not noise:      # sampled entries are the samples themselves, exactly
        assert np.array_equal(y.detach().cpu().numpy()[sampled], k0[sampled])
    y.backward(torch.from_numpy(gy).cuda())
    for l in range(5):
        assert rel_l2(Wt[l].grad.cpu().numpy(), dWo[l]) < TOL[prec], ("dW", l)
        assert rel_l2(bt[l].grad.cpu().numpy(), dbo[l]) < TOL[prec], ("db", l)
    with torch.no_grad():
        y_inf = Fn.siren_mlp(xt, Wt, bt, w0=30.0, precision=prec, fourier=Bt, dc=(k0t, mt, noise, cf))
    assert rel_l2(y_inf.cpu().numpy(), ydc_o) < TOL[prec]


@pytest.mark.parametrize("prec,o", [("bf16", 2), ("bf16", 1), ("fp32", 2)])
def test_forward_dc_mse_c_abi(prec, o):
    """siren_b200_forward_dc_mse straight through the C ABI: the data-consistent y, the k-space MSE summed into
    loss4[1] and its gradient w.r.t. the NETWORK output (mask factor included), so that the plain backward follows."""
    from siren_mri_b200 import _lib
    from siren_mri_b200.functional import _make_desc
    lib = _lib.load()
    tasks, n, d = 2, 5000, 2
    Ws, bs = so.make_params(d, 256, 3, o, seed=7, tasks=0)
    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, size=(tasks, n, d)).astype(np.float32)
    k0 = rng.standard_normal((tasks, o, n)).astype(np.float32)
    mask = (rng.uniform(size=(tasks, o, n)) < 0.4).astype(np.float32)
    gt = rng.standard_normal((tasks, n, o)).astype(np.float32)
    Wt = [torch.from_numpy(w).cuda() for w in Ws]
    bt = [torch.from_numpy(b).cuda() for b in bs]
    xt, k0t, mt, gtt = (torch.from_numpy(a).cuda() for a in (x, k0, mask, gt))
    desc = _make_desc(xt, Wt, 30.0, prec, 0)
    ws = torch.empty(lib.siren_b200_workspace_bytes_ex(desc, 0), dtype=torch.uint8, device="cuda")
    y = torch.empty((tasks, n, o), device="cuda")
    gy = torch.empty_like(y)
    loss4 = torch.zeros(4, device="cuda")
    dc = _lib.SirenDC()
    dc.k0, dc.mask, dc.noise_lvl, dc.channels_first = _lib.dptr(k0t), _lib.dptr(mt), 0.0, 1
    w = 1.0 / 16384.0
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.siren_b200_forward_dc_mse(desc, None, dc, _lib.dptr(xt), _lib.ptr_array(Wt), _lib.ptr_array(bt), _lib.dptr(y),
                                       _lib.dptr(gtt), ctypes.c_float(w), _lib.dptr(gy), _lib.dptr(loss4), _lib.dptr(ws), st)
    _lib.check(rc, "siren_b200_forward_dc_mse")
    yo, _, _, _ = so.siren_forward(x.astype(np.float64), [a.astype(np.float64) for a in Ws], [a.astype(np.float64) for a in bs],
                                   30.0, order=0)
    m_l, k_l = np.transpose(mask, (0, 2, 1)), np.transpose(k0, (0, 2, 1))
    ydc = (1 - m_l) * yo + m_l * k_l
    assert rel_l2(y.cpu().numpy(), ydc) < TOL[prec]
    ref_gy = 2 * w * (y.cpu().numpy().astype(np.float64) - gt) * (1 - m_l)
    assert rel_l2(gy.cpu().numpy(), ref_gy) < 1e-6
    ref_loss = w * ((y.cpu().numpy().astype(np.float64) - gt) ** 2).sum()
    assert abs(float(loss4[1]) - ref_loss) < 1e-5 * ref_loss


# ---------------------------------------------------------------------------------------------------------------
# hypernetwork head in the consumer's layout
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tasks,k_h", [(8, 256), (1, 128), (11, 512), (40, 36)])
def test_hyper_head_kernel(tasks, k_h):
    """siren_b200_hyper_head against the reference's last linear of a hypernetwork head (meta_modules.py:32-35, 50-54:
    net(z).reshape(-1, 256, 256)) in fp64; the two 16-bit operands are the roundings of ITS fp32 result to the bit
    (what prep_weights would have produced from it), the sum of squares is the hypo_weight_loss term; the backward
    equals autograd through the plain linear."""
    from siren_mri_b200 import functional as Fn
    rng = np.random.default_rng(k_h + tasks)
    h = rng.standard_normal((tasks, k_h)).astype(np.float32)
    Wl = (rng.standard_normal((65536, k_h)) * 0.02).astype(np.float32)
    bl = (rng.standard_normal(65536) * 0.01).astype(np.float32)
    G = rng.standard_normal((tasks, 256, 256)).astype(np.float32)
    ht, Wt, bt = (torch.from_numpy(a).cuda().requires_grad_(True) for a in (h, Wl, bl))
    assert Fn.hyper_head_supported(ht, Wt, bt)
    W, wk, wt, ss = Fn._HyperHeadFn.apply(ht, Wt, bt, 30.0)
    ref = (h.astype(np.float64) @ Wl.astype(np.float64).T + bl).reshape(tasks, 256, 256)
    assert rel_l2(W.detach().cpu().numpy(), ref) < 1e-6
    assert torch.equal(wk, W.detach().half())
    assert torch.equal(wt, (W.detach() * 30.0).transpose(1, 2).contiguous().bfloat16())
    assert abs(float(ss) - (ref ** 2).sum()) < 1e-5 * (ref ** 2).sum()
    Gt = torch.from_numpy(G).cuda()
    ((W * Gt).sum() + 0.1 * ss).backward()
    h2, W2, b2 = (torch.from_numpy(a).cuda().requires_grad_(True) for a in (h, Wl, bl))
    Wr = torch.nn.functional.linear(h2, W2, b2).reshape(tasks, 256, 256)
    ((Wr * Gt).sum() + 0.1 * (Wr ** 2).sum()).backward()
    for a, b in ((ht, h2), (Wt, W2), (bt, b2)):
        assert rel_l2(a.grad.cpu().numpy(), b.grad.cpu().numpy()) < 1e-5


def test_hypernetwork_native_heads_feed_the_fused_kernels():
    """meta_modules.HyperNetwork with native heads -> per-task hypo_params whose hidden weights carry ready-made
    operands -> SingleBVPNet (bf16 mode, lazy Fourier features, fused data consistency): the call converts no weights.
    Same numbers (to the bf16 mode's rounding) as the reference flow (plain heads), for the output, the
    hypo_weight_loss term and the gradients that reach the hypernetwork."""
    from siren_mri_b200 import features, functional, meta_modules, modules
    torch.manual_seed(0)
    T, N, F = 3, 2500, 8
    hypo = modules.SingleBVPNet(out_features=2, type="sine", in_features=2 * F, hidden_features=256, num_hidden_layers=3,
                                precision="bf16").cuda()
    hypo.fuse_dc = True
    hyper = meta_modules.HyperNetwork(hyper_in_features=32, hyper_hidden_layers=1, hyper_hidden_features=64,
                                      hypo_module=hypo).cuda()
    tr = features.GaussianFourierFeatureTransform(num_input_channels=2, mapping_size_spatial=F, scale=21, lazy=True)
    z = torch.randn(T, 32, device="cuda")
    x = torch.rand(T, N, 2, device="cuda") * 2 - 1
    k0 = torch.randn(T, 2, 50, 50, device="cuda")
    mask = (torch.rand(T, 2, 50, 50, device="cuda") < 0.3).float()
    gt = torch.randn(T, N, 2, device="cuda")

    def run(native):
        hyper.native_heads = native
        hyper.zero_grad()
        used = []
        orig = functional._prepared_ops
        functional._prepared_ops = lambda *a: (used.append(orig(*a)), used[-1])[1]
        try:
            hp = hyper(z)
            out = hypo({"coords": tr(x), "img_sparse": k0, "dc_mask": mask}, params=hp)["model_out"]
        finally:
            functional._prepared_ops = orig
        reg = meta_modules.hypo_weight_loss({"hypo_params": hp})
        loss = ((out - gt) ** 2).mean() + 1e2 * reg
        loss.backward()
        grads = [p.grad.detach().clone() for p in hyper.parameters()]
        return out.detach(), float(reg), grads, used

    y1, r1, g1, used1 = run(True)
    y0, r0, g0, used0 = run(False)
    assert used1 and used1[0] is not None and len(used1[0][0]) == 3      # three hidden weights arrived as operands
    assert used0 and used0[0] is None
    assert rel_l2(y1.cpu().numpy(), y0.cpu().numpy()) < 5e-3
    assert abs(r1 - r0) < 1e-5 * abs(r0)
    num = sum(float(((a - b) ** 2).sum()) for a, b in zip(g1, g0))
    den = sum(float((b ** 2).sum()) for b in g0)
    assert (num / den) ** 0.5 < 2e-2


def test_graphed_step_replays_the_neural_process_step():
    """training.GraphedStep: the hypo-path step of the neural-process model (hypernetwork with native heads -> fused
    hypo path with Fourier prologue and data consistency -> losses -> backward) captured into one CUDA graph gives, on
    replay with a NEW batch copied into the static inputs, the loss and hypernetwork gradients of the same step
    launched from Python."""
    from siren_mri_b200 import features, meta_modules
    from siren_mri_b200.training import GraphedStep
    torch.manual_seed(0)
    T, side, F = 2, 48, 8
    model = meta_modules.ConvolutionalNeuralProcessImplicit2DHypernetFourierFeatures(
        in_features=2 * F, out_features=2, image_resolution=(side, side), fourier_features_size=2 * F, latent_dim=32,
        num_hidden_layers=3, hyper_hidden_features=64, num_conv_res_blocks=1, precision="bf16").cuda()
    tr = features.GaussianFourierFeatureTransform(2, F, 21, lazy=True)
    params = list(model.hyper_net.parameters())

    def batch(seed):
        g = torch.Generator(device="cuda").manual_seed(seed)
        return (torch.rand((T, side * side, 2), device="cuda", generator=g) * 2 - 1,
                torch.randn((T, 2, side, side), device="cuda", generator=g),
                (torch.rand((T, 2, side, side), device="cuda", generator=g) < 0.3).float(),
                torch.randn((T, side * side, 2), device="cuda", generator=g),
                torch.randn((T, 32), device="cuda", generator=g))

    def step(x, k0, mask, gt, z):
        out = model({"coords": tr(x), "img_sparse": k0, "dc_mask": mask, "embedding": z})
        loss = ((out["model_out"] - gt) ** 2).mean() + 1e2 * meta_modules.hypo_weight_loss(out)
        for p in params:
            p.grad = None
        loss.backward()
        return loss

    gs = GraphedStep(step, batch(1))
    new = batch(2)
    loss_g = float(gs(*new).detach())
    grads_g = [p.grad.detach().clone() for p in params]
    loss_e = float(step(*new).detach())
    grads_e = [p.grad.detach().clone() for p in params]
    assert abs(loss_g - loss_e) <= 1e-5 * abs(loss_e)
    num = sum(float(((a - b) ** 2).sum()) for a, b in zip(grads_g, grads_e))
    den = sum(float((b ** 2).sum()) for b in grads_e)
    assert (num / den) ** 0.5 < 1e-3      # the weight-gradient sums are atomics: last bits differ run to run
