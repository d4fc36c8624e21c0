#!/usr/bin/env python
"""Developer probe: one training step (forward, MSE, backward; per-task weights, 8 tasks x 65,536 coordinates) for the
Fourier blocks of train_mri_neural_process_ddp.py:54-130 -- features materialised by the reference's ops every step vs
built inside the kernels (lazy) vs the reference's ops in eager PyTorch."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import siren_mri_b200  # noqa: E402
from siren_mri_b200 import features, modules  # noqa: E402
import workloads  # noqa: E402


def run(F, mode, tasks=8, steps=10):
    dev = torch.device("cuda")
    torch.manual_seed(0)
    backend = "composed" if mode == "eager" else "auto"
    model = modules.SingleBVPNet(in_features=2 * F, out_features=2, precision="bf16", backend=backend).to(dev)
    params = workloads.per_task_params(model, tasks, dev)
    tr = features.GaussianFourierFeatureTransform(2, F, 21, lazy=(mode == "lazy"))
    tr.set_B((torch.randn((2, F), generator=torch.Generator().manual_seed(0)) * 21.0).to(dev))
    x = workloads.mgrid(256).unsqueeze(0).expand(tasks, -1, -1).contiguous().to(dev)
    gt = (torch.rand((tasks, 65536, 2), generator=torch.Generator().manual_seed(1)) * 2 - 1).to(dev)
    leaves = list(params.values())

    def step():
        out = model({"coords": tr(x)}, params=params)
        loss = ((out["model_out"] - gt) ** 2).mean()
        for p in leaves:
            p.grad = None
        loss.backward()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    res = {"F": F, "in_features": 2 * F, "tasks": tasks, "coords_per_step": tasks * 65536, "mode": mode,
           "ms_per_step": round(ms, 3), "Mcoord_per_s": round(tasks * 65536 / ms / 1e3, 1)}
    del model, params, leaves
    siren_mri_b200.functional.clear_workspace_cache()
    torch.cuda.empty_cache()
    return res


if __name__ == "__main__":
    for F in (8, 30, 60, 128):
        for mode in ("materialised", "lazy", "eager"):
            print(json.dumps(run(F, mode, steps=10 if mode != "eager" else 3)), flush=True)
