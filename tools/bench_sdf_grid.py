#!/usr/bin/env python
"""Throughput of siren_mri_b200.sdf_meshing.sample_sdf_grid (sdf_meshing.create_mesh's sampling loop) at N = 256 and
N = 512, bf16 and fp32-parity, next to the reference's loop shape (host-built samples, a copy each way per chunk) on
the composed ops."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import siren_mri_b200  # noqa: E402
from siren_mri_b200 import modules, sdf_meshing  # noqa: E402


def reference_loop(decoder, N, max_batch):
    """sdf_meshing.py:25-56 as written (floor divisions), values gathered on the host."""
    idx = torch.arange(0, N ** 3)
    samples = torch.zeros(N ** 3, 4)
    vs = 2.0 / (N - 1)
    samples[:, 2] = (idx % N) * vs - 1
    samples[:, 1] = ((idx // N) % N) * vs - 1
    samples[:, 0] = ((idx // N // N) % N) * vs - 1
    head = 0
    with torch.no_grad():
        while head < N ** 3:
            sub = samples[head:head + max_batch, 0:3].cuda()
            samples[head:head + max_batch, 3] = decoder(sub).squeeze().detach().cpu()
            head += max_batch
    return samples[:, 3].reshape(N, N, N)


for N in (256, 512):
    for prec, backend in (("bf16", "auto"), ("fp32", "auto"), ("fp32", "composed")):
        siren_mri_b200.set_defaults(backend=backend)
        torch.manual_seed(0)
        m = modules.SingleBVPNet(in_features=3, out_features=1, precision=prec).cuda()
        dec = lambda c: m.net(c)      # noqa: E731
        sdf_meshing.sample_sdf_grid(dec, N=64, max_batch=64 ** 3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if backend == "composed":
            vol = reference_loop(dec, N, 64 ** 3)
        else:
            vol = sdf_meshing.sample_sdf_grid(dec, N=N, max_batch=16 * 64 ** 3)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps({"N": N, "mode": prec if backend == "auto" else "reference loop, eager fp32", "seconds": round(dt, 4),
                          "Mcoord_per_s": round(N ** 3 / dt / 1e6, 1), "finite": bool(torch.isfinite(vol).all())}), flush=True)
siren_mri_b200.set_defaults(backend="auto")
