#!/usr/bin/env python
"""Time the five BASELINE.json configurations through the PUBLIC module API (autograd path):
model(model_input) -> unchanged-style loss -> backward -> Adam.  One JSON line per config.

  python tools/bench_configs.py [--precision bf16|fp32] [--configs 1,2,3,4,5] [--steps 10]

cfg1  SIREN 3x256 image fit, 256x256 = 65,536 coords, MSE
cfg2  same at 512x512 = 262,144 coords
cfg3  SDF, d=3, 250,000 coords, loss_functions.sdf (first-order derivatives), clip_grad
cfg4  Poisson from Laplacian, 262,144 coords, loss_functions.laplace_mse (second order)
cfg5  hypernetwork SIREN, B tasks x 65,536 coords, d=16, o=2, per-task weights, MSE
      (B = 8 = one GPU's share of the 64-task batch sharded over 8 GPUs; --tasks to change)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from siren_mri_b200 import diff_operators, modules  # noqa: E402

U = lambda d, o: 2 * (d * 256 + 3 * 256 * 256 + 256 * o)  # noqa: E731
PREC = lambda p: p if p in ("bf16", "fp32") else "fp32"  # noqa: E731


def sdf_loss(model_output, gt):
    """loss_functions.py:460-484 (restated: the reference file is not on the GPU box)."""
    gt_sdf, gt_normals = gt["sdf"], gt["normals"]
    coords, pred = model_output["model_in"], model_output["model_out"]
    gradient = diff_operators.gradient(pred, coords)
    sdf_c = torch.where(gt_sdf != -1, pred, torch.zeros_like(pred))
    inter = torch.where(gt_sdf != -1, torch.zeros_like(pred), torch.exp(-1e2 * torch.abs(pred)))
    normal = torch.where(gt_sdf != -1, 1 - F.cosine_similarity(gradient, gt_normals, dim=-1)[..., None],
                         torch.zeros_like(gradient[..., :1]))
    gc = torch.abs(gradient.norm(dim=-1) - 1)
    return torch.abs(sdf_c).mean() * 3e3 + inter.mean() * 1e2 + normal.mean() * 1e2 + gc.mean() * 5e1


def run(cfg, precision, steps, warmup, tasks):
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(cfg)
    torch.manual_seed(cfg)
    if cfg in (1, 2, 4):
        n = 65536 if cfg == 1 else 262144
        d, o, derivs = 2, 1, (2 if cfg == 4 else 0)
        model = modules.SingleBVPNet(in_features=d, out_features=o, precision=PREC(precision), coord_derivs=derivs).to(dev)
        x = torch.rand((1, n, d), device=dev, generator=g) * 2 - 1
        gt = torch.rand((1, n, 1), device=dev, generator=g) * 2 - 1
        if cfg == 4:
            def loss_fn(out):
                lap = diff_operators.laplace(out["model_out"], out["model_in"])
                return torch.mean((lap - 1e3 * gt) ** 2)
            flop = 12 * U(d, o)
        else:
            def loss_fn(out):
                return ((out["model_out"] - gt) ** 2).sum() / 16384.0
            flop = 3 * U(d, o)
        params, clip, coords_total = None, False, n
    elif cfg == 3:
        n, d, o = 250000, 3, 1
        model = modules.SingleBVPNet(in_features=d, out_features=o, precision=PREC(precision), coord_derivs=1).to(dev)
        p = torch.randn((1, n, 3), device=dev, generator=g)
        on = 0.5 * p[:, : n // 2] / p[:, : n // 2].norm(dim=-1, keepdim=True)
        off = torch.rand((1, n - n // 2, 3), device=dev, generator=g) * 2 - 1
        x = torch.cat([on, off], dim=1)
        sdf = torch.cat([torch.zeros(1, n // 2, 1, device=dev), -torch.ones(1, n - n // 2, 1, device=dev)], dim=1)
        normals = torch.cat([on / 0.5, -torch.ones(1, n - n // 2, 3, device=dev)], dim=1)
        gtd = {"sdf": sdf, "normals": normals}

        def loss_fn(out):
            return sdf_loss(out, gtd)
        flop = 6 * U(d, o)
        params, clip, coords_total = None, True, n
    else:
        n, d, o = 65536, 16, 2
        model = modules.SingleBVPNet(in_features=d, out_features=o, precision=PREC(precision)).to(dev)
        x = torch.rand((tasks, n, d), device=dev, generator=g) * 2 - 1
        gt = torch.rand((tasks, n, o), device=dev, generator=g) * 2 - 1
        from collections import OrderedDict
        params = OrderedDict()
        for name, p in model.named_parameters():          # hypernetwork-style per-task weights
            pt = (p.detach().unsqueeze(0) + 1e-3 * torch.randn((tasks,) + tuple(p.shape), device=dev, generator=g))
            params[name] = pt.requires_grad_(True)

        def loss_fn(out):
            return ((out["model_out"] - gt) ** 2).sum() / 16384.0
        flop = 3 * U(d, o)
        clip, coords_total = False, tasks * n

    leaves = list(params.values()) if params is not None else list(model.parameters())
    opt = torch.optim.Adam(leaves, lr=1e-4)

    def step():
        out = model({"coords": x}, params=params) if params is not None else model({"coords": x})
        loss = loss_fn(out)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if clip:
            torch.nn.utils.clip_grad_norm_(leaves, max_norm=1.0)
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    value = coords_total / (ms * 1e-3)
    kernels = None
    if os.environ.get("SIREN_PROFILE"):
        import ctypes
        from siren_mri_b200 import _lib
        lib = _lib.load()
        lib.siren_b200_profile_begin()
        for _ in range(3):
            step()
        buf = ctypes.create_string_buffer(1 << 16)
        lib.siren_b200_profile_end(buf, len(buf))
        kernels = {}
        for ln in buf.value.decode().strip().splitlines():
            name, cnt, tot = ln.split()
            kernels[name] = round(1e3 * float(tot) / 3, 1)      # us per step
    print(json.dumps({"config": "cfg%d" % cfg, "precision": precision, "kernels_us_per_step": kernels, "coords_per_step": coords_total,
                      "ms_per_step": ms, "coords_per_sec": value, "flop_per_coord_algorithmic": flop,
                      "frac_of_bf16_peak_1656.6TF": flop * value / 1656.6e12, "loss": float(loss.item()),
                      "path": "public modules + torch autograd + torch.optim.Adam",
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tasks", type=int, default=8)
    ap.add_argument("--backend", default="auto", choices=["auto", "composed"],
                    help="composed = the reference's own PyTorch ops in eager mode on the GPU (fp32, TF32 off)")
    a = ap.parse_args()
    import siren_mri_b200
    siren_mri_b200.set_defaults(backend=a.backend)
    if a.backend == "composed":
        a.precision = "eager-fp32"
    for c in [int(v) for v in a.configs.split(",")]:
        run(c, a.precision, a.steps, a.warmup, a.tasks)
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
