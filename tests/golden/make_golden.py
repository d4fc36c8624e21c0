"""Generate golden fixtures from the UNMODIFIED reference (dev container only).

Run:  python tests/golden/make_golden.py
Needs /root/reference (read-only).  Writes tests/golden/*.npz.  The fixtures are
what pins oracle/siren_oracle.py (tests/test_oracle_golden.py) and what the GPU
parity tests compare against directly.  Nothing here runs on the GPU box.

Import recipe: bypass torchmeta/__init__.py (it pulls h5py)
and stub the plotting / IO modules that loss_functions.py imports but never uses
on this path.
"""
import os
import sys
import types
from collections import OrderedDict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def import_reference():
    sys.path.insert(0, REF)
    pkg = types.ModuleType("torchmeta")
    pkg.__path__ = [os.path.join(REF, "torchmeta")]
    sys.modules["torchmeta"] = pkg
    for name in ["h5py", "matplotlib", "matplotlib.colors", "matplotlib.pyplot", "skimage",
                 "skimage.filters", "skimage.measure", "skvideo", "skvideo.io", "cmapy"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    import modules, diff_operators, loss_functions, meta_modules  # noqa: E401
    return modules, diff_operators, loss_functions, meta_modules


def sample(a, stride=97):
    return np.ascontiguousarray(a.reshape(-1)[::stride])


def load_weights(model, Ws, bs, dtype):
    sd = OrderedDict()
    for l, (W, b) in enumerate(zip(Ws, bs)):
        sd["net.net.%d.0.weight" % l] = torch.from_numpy(W).to(dtype)
        sd["net.net.%d.0.bias" % l] = torch.from_numpy(b).to(dtype)
    model.load_state_dict(sd)


def param_grads(model):
    out = {}
    for l in range(5):
        lin = model.net.net[l][0]
        for nm, prm in (("dW%d" % l, lin.weight), ("db%d" % l, lin.bias)):
            # a parameter the loss does not depend on keeps grad None (e.g. the last bias
            # under a pure derivative loss); record that as zeros
            out[nm] = (prm.grad.detach().numpy().copy() if prm.grad is not None
                       else np.zeros(tuple(prm.shape), prm.detach().numpy().dtype))
    return out


def pack_grads(prefix, grads, store):
    for k, g in grads.items():
        if k.startswith("dW") and g.size > 4096:
            store["%s_%s_sample" % (prefix, k)] = sample(g)
            store["%s_%s_sum" % (prefix, k)] = np.array(g.astype(np.float64).sum())
            store["%s_%s_l2" % (prefix, k)] = np.array(np.sqrt((g.astype(np.float64) ** 2).sum()))
        else:
            store["%s_%s" % (prefix, k)] = g


def shared_case(modules, diff_operators, loss_functions, name, d, o, n, seed, dtype):
    from oracle import siren_oracle as so
    Ws, bs = so.make_params(d, 256, 3, o, seed=seed)
    x = so.make_coords(1, n, d, seed=seed + 100)
    rng = np.random.Generator(np.random.PCG64(seed + 200))
    gt = rng.uniform(-1, 1, size=(1, n, o)).astype(np.float32)
    gt_grad = rng.standard_normal((1, n, d)).astype(np.float32)
    gt_lap = (50.0 * rng.standard_normal((1, n, 1))).astype(np.float32)
    model = modules.SingleBVPNet(out_features=o, type="sine", in_features=d, mode="mlp",
                                 hidden_features=256, num_hidden_layers=3)
    model = model.to(dtype)
    load_weights(model, Ws, bs, dtype)
    tx = torch.from_numpy(x).to(dtype)
    store = {"d": d, "o": o, "n": n, "seed": seed}
    tag = "f64" if dtype == torch.float64 else "f32"

    out = model({"coords": tx})
    y, xin = out["model_out"], out["model_in"]
    store["y"] = y.detach().numpy()
    g = diff_operators.gradient(y, xin)
    store["grad"] = g.detach().numpy()
    if o == 1:
        lap = diff_operators.laplace(y, xin)
        store["lap"] = lap.detach().numpy()
    # diff_operators.jacobian (:46-59) and hessian (:5-24): full [B, N, o, d] / [B, N, o, d, d] tensors
    # (the reference allocates them in fp32 whatever the model's dtype)
    out = model({"coords": tx})
    jac, _ = diff_operators.jacobian(out["model_out"], out["model_in"])
    store["jac"] = jac.detach().numpy()
    hes, _ = diff_operators.hessian(out["model_out"], out["model_in"])
    store["hess"] = hes.detach().numpy()

    # loss A: plain MSE on the value (cfg1/cfg2 use image_mse == sum/16384)
    model.zero_grad()
    out = model({"coords": tx})
    loss = ((out["model_out"] - torch.from_numpy(gt).to(dtype)) ** 2).sum() / 16384.0
    loss.backward()
    store["mse_loss"] = np.array(loss.item())
    pack_grads("mse", param_grads(model), store)
    store["mse_gx"] = out["model_in"].grad.numpy().copy()

    # loss B: gradients_mse (loss_functions.py:330-335), first-order coordinate derivatives
    model.zero_grad()
    out = model({"coords": tx})
    loss = loss_functions.gradients_mse(out, {"gradients": torch.from_numpy(gt_grad).to(dtype)})["gradients_loss"]
    loss.backward()
    store["gradmse_loss"] = np.array(loss.item())
    pack_grads("gradmse", param_grads(model), store)

    if o == 1:
        # loss C: laplace_mse (loss_functions.py:350-355), second order
        model.zero_grad()
        out = model({"coords": tx})
        loss = loss_functions.laplace_mse(out, {"laplace": torch.from_numpy(gt_lap).to(dtype)})["laplace_loss"]
        loss.backward()
        store["lapmse_loss"] = np.array(loss.item())
        pack_grads("lapmse", param_grads(model), store)

    if d == 3 and o == 1:
        # loss D: sdf (loss_functions.py:460-484)
        sdf = np.where(np.arange(n) < n // 2, 0.0, -1.0).astype(np.float32).reshape(1, n, 1)
        normals = x / np.maximum(np.linalg.norm(x, axis=-1, keepdims=True), 1e-6)
        normals = np.where(sdf != -1, normals, -1.0).astype(np.float32)
        model.zero_grad()
        out = model({"coords": tx})
        losses = loss_functions.sdf(out, {"sdf": torch.from_numpy(sdf).to(dtype),
                                          "normals": torch.from_numpy(normals).to(dtype)})
        loss = sum(v.mean() for v in losses.values())
        loss.backward()
        store["sdf_loss"] = np.array(loss.item())
        store["sdf_gt"] = sdf
        store["sdf_normals"] = normals
        pack_grads("sdf", param_grads(model), store)

    store["gt"] = gt
    store["gt_grad"] = gt_grad
    store["gt_lap"] = gt_lap
    np.savez_compressed(os.path.join(HERE, "%s_%s.npz" % (name, tag)), **store)
    print("wrote", name, tag)


def per_task_case(modules, loss_functions, name, tasks, d, o, n, seed, dtype):
    """BatchLinear with per-sample weights (modules.py:25, weight [B,out,in])."""
    from oracle import siren_oracle as so
    Ws, bs = so.make_params(d, 256, 3, o, seed=seed, tasks=tasks)
    x = so.make_coords(tasks, n, d, seed=seed + 100)
    rng = np.random.Generator(np.random.PCG64(seed + 200))
    gt = rng.uniform(-1, 1, size=(tasks, n, o)).astype(np.float32)
    model = modules.SingleBVPNet(out_features=o, type="sine", in_features=d, mode="mlp",
                                 hidden_features=256, num_hidden_layers=3).to(dtype)
    params = OrderedDict()
    for l, (W, b) in enumerate(zip(Ws, bs)):
        params["net.net.%d.0.weight" % l] = torch.from_numpy(W).to(dtype).requires_grad_(True)
        params["net.net.%d.0.bias" % l] = torch.from_numpy(b).to(dtype).requires_grad_(True)
    out = model({"coords": torch.from_numpy(x).to(dtype)}, params=params)
    y = out["model_out"]
    loss = ((y - torch.from_numpy(gt).to(dtype)) ** 2).sum() / 16384.0
    loss.backward()
    tag = "f64" if dtype == torch.float64 else "f32"
    store = {"d": d, "o": o, "n": n, "seed": seed, "tasks": tasks, "y": y.detach().numpy(),
             "gt": gt, "mse_loss": np.array(loss.item())}
    grads = {}
    for l in range(5):
        grads["dW%d" % l] = params["net.net.%d.0.weight" % l].grad.numpy()
        grads["db%d" % l] = params["net.net.%d.0.bias" % l].grad.numpy()
    pack_grads("mse", grads, store)
    np.savez_compressed(os.path.join(HERE, "%s_%s.npz" % (name, tag)), **store)
    print("wrote", name, tag)


def fourier_case(modules, name, tasks, F, o, nx, ny, seed, dtype):
    """MRI prologue / epilogue of the neural-process models: GaussianFourierFeatureTransform (features.py:31-41) on the
    raw coordinates, per-sample-weight SingleBVPNet (meta_modules.py:213), DataConsistencyInKspace
    (data_consistency.py:32-47, as called at meta_modules.py:217-219), MSE on the result."""
    import features, data_consistency      # the reference's own modules
    from oracle import siren_oracle as so
    n, d = nx * ny, 2 * F
    Ws, bs = so.make_params(d, 256, 3, o, seed=seed, tasks=tasks)
    rng = np.random.Generator(np.random.PCG64(seed + 300))
    x = rng.uniform(-1, 1, size=(tasks, n, 2)).astype(np.float32)
    B = (21.0 * rng.standard_normal((2, F))).astype(np.float32)
    gt = rng.uniform(-1, 1, size=(tasks, n, o)).astype(np.float32)
    k0 = rng.uniform(-1, 1, size=(tasks, 2, nx, ny)).astype(np.float32)
    mask = (rng.uniform(0, 1, size=(tasks, 1, nx, ny)) < 0.3).astype(np.float32).repeat(2, axis=1)
    tr = features.GaussianFourierFeatureTransform(num_input_channels=2, mapping_size_spatial=F, scale=21, device="cpu")
    tr.set_B(torch.from_numpy(B).to(dtype))
    model = modules.SingleBVPNet(out_features=o, type="sine", in_features=d, mode="mlp",
                                 hidden_features=256, num_hidden_layers=3).to(dtype)
    params = OrderedDict()
    for l, (W, b) in enumerate(zip(Ws, bs)):
        params["net.net.%d.0.weight" % l] = torch.from_numpy(W).to(dtype).requires_grad_(True)
        params["net.net.%d.0.bias" % l] = torch.from_numpy(b).to(dtype).requires_grad_(True)
    feat = tr(torch.from_numpy(x).to(dtype))
    y = model({"coords": feat}, params=params)["model_out"]
    dc = data_consistency.DataConsistencyInKspace(noise_lvl=None)
    y_dc = dc(y, torch.from_numpy(k0).to(dtype), torch.from_numpy(mask).to(dtype))
    loss = ((y_dc - torch.from_numpy(gt).to(dtype)) ** 2).sum() / 16384.0
    loss.backward()
    tag = "f64" if dtype == torch.float64 else "f32"
    store = {"F": F, "o": o, "nx": nx, "ny": ny, "seed": seed, "tasks": tasks, "x": x, "B": B, "gt": gt, "k0": k0,
             "mask": mask, "feat": feat.detach().numpy(), "y": y.detach().numpy(), "y_dc": y_dc.detach().numpy(),
             "mse_loss": np.array(loss.item())}
    grads = {}
    for l in range(5):
        grads["dW%d" % l] = params["net.net.%d.0.weight" % l].grad.numpy()
        grads["db%d" % l] = params["net.net.%d.0.bias" % l].grad.numpy()
    pack_grads("mse", grads, store)
    np.savez_compressed(os.path.join(HERE, "%s_%s.npz" % (name, tag)), **store)
    print("wrote", name, tag)


def adam_case():
    """torch.optim.Adam (training.py:23) + clip_grad_norm_ (training.py:93-97)."""
    rng = np.random.Generator(np.random.PCG64(7))
    n = 1000
    p0 = rng.standard_normal(n).astype(np.float32)
    grads = [(10.0 ** (i - 2)) * rng.standard_normal(n).astype(np.float32) for i in range(5)]
    for clip in (0.0, 1.0):
        p = torch.nn.Parameter(torch.from_numpy(p0.copy()))
        opt = torch.optim.Adam([p], lr=1e-4)
        traj = []
        for g in grads:
            p.grad = torch.from_numpy(g.copy())
            if clip:
                torch.nn.utils.clip_grad_norm_([p], max_norm=clip)
            opt.step()
            traj.append(p.detach().numpy().copy())
        np.savez_compressed(os.path.join(HERE, "adam_clip%d.npz" % int(clip)), p0=p0,
                            grads=np.stack(grads), traj=np.stack(traj), lr=np.array(1e-4),
                            clip=np.array(clip))
    print("wrote adam")


def hypernet_case(modules, meta_modules):
    """HyperNetwork output contract (meta_modules.py:42-54): names and shapes."""
    torch.manual_seed(0)
    hypo = modules.SingleBVPNet(out_features=2, type="sine", in_features=16, hidden_features=256,
                                num_hidden_layers=3)
    hyper = meta_modules.HyperNetwork(hyper_in_features=8, hyper_hidden_layers=1,
                                      hyper_hidden_features=16, hypo_module=hypo)
    z = torch.randn(3, 8)
    params = hyper(z)
    names = list(params.keys())
    shapes = [tuple(v.shape) for v in params.values()]
    np.savez_compressed(os.path.join(HERE, "hypernet_contract.npz"), names=np.array(names),
                        shapes=np.array([str(s) for s in shapes]),
                        state_keys=np.array(list(hypo.state_dict().keys())))
    print("wrote hypernet contract", names[:2], shapes[:2])


if __name__ == "__main__":
    torch.manual_seed(0)
    modules, diff_operators, loss_functions, meta_modules = import_reference()
    only = sys.argv[1:]      # e.g. `make_golden.py fourier` regenerates that family alone
    if only == ["fourier"]:
        for dtype in (torch.float64, torch.float32):
            fourier_case(modules, "fourier_t2_f8_o2", 2, 8, 2, 20, 15, 15, dtype)
        sys.exit(0)
    for dtype in (torch.float64, torch.float32):
        fourier_case(modules, "fourier_t2_f8_o2", 2, 8, 2, 20, 15, 15, dtype)
        shared_case(modules, diff_operators, loss_functions, "img_d2_o1", 2, 1, 200, 11, dtype)
        shared_case(modules, diff_operators, loss_functions, "sdf_d3_o1", 3, 1, 131, 12, dtype)
        shared_case(modules, diff_operators, loss_functions, "vec_d2_o3", 2, 3, 77, 13, dtype)
        per_task_case(modules, loss_functions, "mri_t3_d16_o2", 3, 16, 2, 150, 14, dtype)
    adam_case()
    hypernet_case(modules, meta_modules)
