#!/usr/bin/env python
"""Developer probe: one training step of the MRI neural-process HYPO PATH on one GPU's share of cfg5 -- 8 slices x
65,536 k-space coordinates, F = 8 Fourier features, hypo-network 3 x 256, hypernetwork latent 128 -> 128 -> heads,
k-space data consistency, image MSE + latent + hypo-weight terms (loss_functions.py:290-293), backward into the
hypernetwork.  The ConvImgEncoder is skipped through model_input['embedding'] (meta_modules.py:198-209) so that the
numbers isolate what this package builds.

  native      bf16 mode, Fourier features in the kernels, data consistency in the kernels, hypernetwork heads that
              emit the kernels' operands
  native-f    the same without those three fusions (features materialised, elementwise data consistency, plain heads
              + weight conversion in the call)
  native-graph  "native" with the whole step (forward, losses, backward) captured into ONE CUDA graph and replayed:
              the step is ~150 small launches around three large ones, i.e. host-bound when launched from Python
  reference   the unmodified reference classes (baseline/_ref) in eager PyTorch, fp32, TF32 off
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import workloads  # noqa: E402

TASKS, SIDE, F, LATENT, HYPER = 8, 256, 8, 128, 128
KW = dict(in_features=2 * F, out_features=2, image_resolution=(SIDE, SIDE), fourier_features_size=2 * F, latent_dim=LATENT,
          hidden_features=256, num_hidden_layers=3, hyper_hidden_features=HYPER, hyper_hidden_layers=1,
          conv_kernel_size=3, num_conv_res_blocks=1)


def inputs(dev):
    g = torch.Generator().manual_seed(0)
    x = workloads.mgrid(SIDE).unsqueeze(0).expand(TASKS, -1, -1).contiguous().to(dev)
    k0 = torch.randn((TASKS, 2, SIDE, SIDE), generator=g).to(dev)
    mask = (torch.rand((TASKS, 2, SIDE, SIDE), generator=g) < 0.25).float().to(dev)
    gt = torch.randn((TASKS, SIDE * SIDE, 2), generator=g).to(dev)
    z = torch.randn((TASKS, LATENT), generator=g).to(dev)
    B = (torch.randn((2, F), generator=g) * 21.0).to(dev)
    return x, k0, mask, gt, z, B


def loss_of(out, gt, hwl):
    img = ((out["model_out"] - gt) ** 2).mean()
    lat = torch.mean(out["latent_vec"] ** 2)
    return img + 1e-1 * lat + 1e2 * hwl(out)


def ref_hwl(out):      # loss_functions.hypo_weight_loss (loss_functions.py:279-287)
    s, n = 0, 0
    for w in out["hypo_params"].values():
        s = s + torch.sum(w ** 2)
        n += w.numel()
    return s * (1 / n)


def timed(step, steps):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run(mode, steps=10):
    dev = torch.device("cuda")
    torch.manual_seed(0)
    x, k0, mask, gt, z, B = inputs(dev)
    if mode == "reference":
        ref = workloads.reference_modules()
        if ref is None:
            return {"mode": mode, "unavailable": "baseline/_ref not staged"}
        import features as ref_features
        import meta_modules as ref_meta
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            model = ref_meta.ConvolutionalNeuralProcessImplicit2DHypernetFourierFeatures(**KW).to(dev)
        tr = ref_features.GaussianFourierFeatureTransform(2, F, 21, device=dev)
        tr.set_B(B)
        hwl = ref_hwl
    else:
        from siren_mri_b200 import features, meta_modules
        fused = mode in ("native", "native-graph")
        model = meta_modules.ConvolutionalNeuralProcessImplicit2DHypernetFourierFeatures(
            precision="bf16", fuse_dc=fused, native_heads=fused, **KW).to(dev)
        tr = features.GaussianFourierFeatureTransform(2, F, 21, lazy=fused)
        tr.set_B(B)
        hwl = meta_modules.hypo_weight_loss if fused else ref_hwl
    params = [p for p in model.hyper_net.parameters()]

    def step():
        out = model({"coords": tr(x), "img_sparse": k0, "dc_mask": mask, "embedding": z})
        loss = loss_of(out, gt, hwl)
        for p in params:
            p.grad = None
        loss.backward()
        return loss

    if mode == "native-graph":
        from siren_mri_b200.training import GraphedStep
        gs = GraphedStep(lambda: step(), [])
        graph, static_loss = gs.graph, gs.static_outputs
        ms = timed(graph.replay, steps)
        graph.replay()
        torch.cuda.synchronize()
        loss = float(static_loss.detach())
        gn = sum(float((p.grad ** 2).sum()) for p in params) ** 0.5
        return {"mode": mode, "tasks": TASKS, "coords_per_step": TASKS * SIDE * SIDE, "ms_per_step": round(ms, 3),
                "Mcoord_per_s": round(TASKS * SIDE * SIDE / ms / 1e3, 1), "loss": loss, "grad_norm": gn}
    ms = timed(step, steps)
    loss = float(step().detach())
    gn = sum(float((p.grad ** 2).sum()) for p in params) ** 0.5
    return {"mode": mode, "tasks": TASKS, "coords_per_step": TASKS * SIDE * SIDE, "ms_per_step": round(ms, 3),
            "Mcoord_per_s": round(TASKS * SIDE * SIDE / ms / 1e3, 1), "loss": loss, "grad_norm": gn}


if __name__ == "__main__":
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    for mode in ("native", "native-graph", "native-f", "reference"):
        print(json.dumps(run(mode, steps=10 if mode != "reference" else 3)), flush=True)
        torch.cuda.empty_cache()
