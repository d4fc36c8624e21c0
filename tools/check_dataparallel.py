#!/usr/bin/env python
"""nn.DataParallel check (train_mri_neural_process.py:165-166 wraps the model that way): one process, one thread per
GPU, every replica calling the C ABI concurrently.  A per-task-weight SIREN is scattered over the visible GPUs along the
task axis; outputs and weight gradients must equal the single-GPU run."""
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

from siren_mri_b200 import modules  # noqa: E402


class Wrapper(nn.Module):
    """hypo-net with per-task weights handed in as tensors (DataParallel scatters tensors along dim 0)."""

    def __init__(self, prec):
        super().__init__()
        self.hypo = modules.SingleBVPNet(in_features=16, out_features=2, precision=prec)
        self.names = [k for k, _ in self.hypo.named_parameters()]

    def forward(self, coords, *params):
        p = OrderedDict(zip(self.names, params))
        return self.hypo({"coords": coords}, params=p)["model_out"]


def main():
    n_gpu = torch.cuda.device_count()
    assert n_gpu >= 2, "needs two GPUs"
    for prec in ("bf16", "fp32"):
        torch.manual_seed(0)
        tasks, n = 2 * n_gpu, 5000
        w = Wrapper(prec).cuda(0)
        coords = torch.rand(tasks, n, 16, device="cuda:0") * 2 - 1
        params = [(p.detach().unsqueeze(0) + 0.01 * p.detach().abs().mean() * torch.randn((tasks,) + tuple(p.shape), device="cuda:0"))
                  .requires_grad_(True) for p in w.hypo.parameters()]
        y1 = w(coords, *params)
        y1.square().sum().backward()
        g1 = [p.grad.clone() for p in params]
        for p in params:
            p.grad = None
        dp = nn.DataParallel(w, device_ids=list(range(n_gpu)))
        for it in range(3):                      # several rounds: the replicas' threads race into the library
            for p in params:
                p.grad = None
            y2 = dp(coords, *params)
            y2.square().sum().backward()
        err_y = float((y2 - y1).norm() / y1.norm())
        err_g = max(float((p.grad - g).norm() / g.norm()) for p, g in zip(params, g1))
        tol = 1e-6 if prec == "fp32" else 2e-3      # bf16: split-K atomics order differs between runs
        print("DataParallel over %d GPUs, %s: |y - y1| / |y1| = %.2e, worst gradient %.2e" % (n_gpu, prec, err_y, err_g), flush=True)
        assert err_y < tol and err_g < max(tol, 1e-5), (err_y, err_g)
    print("DataParallel check OK", flush=True)


if __name__ == "__main__":
    main()
