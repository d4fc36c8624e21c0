#!/usr/bin/env python
"""Developer probe: one small training forward + backward through every fused-path variant
(narrow first layer, wide d = 16 per task, Fourier prologue F = 8 and F = 30, d = 40 materialised, d_out = 3)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import siren_oracle as so  # noqa: E402
from siren_mri_b200 import functional as F  # noqa: E402

CASES = [  # d, o, tasks, per_task, n, fourier F (0: none), raw
    (2, 1, 1, False, 1300, 0, 0), (16, 2, 2, True, 700, 0, 0), (16, 2, 2, True, 700, 8, 2), (60, 2, 1, False, 900, 30, 2),
    (40, 2, 2, True, 500, 0, 0), (3, 3, 1, False, 600, 0, 0),
]
for d, o, tasks, per_task, n, ff, raw in CASES:
    Ws, bs = so.make_params(d, 256, 3, o, seed=d, tasks=tasks if per_task else 0)
    rng = np.random.default_rng(d)
    x = rng.uniform(-1, 1, (tasks, n, raw if ff else d)).astype(np.float32)
    B = torch.from_numpy((21 * rng.standard_normal((raw, ff))).astype(np.float32)).cuda() if ff else None
    Wt = [torch.from_numpy(w.astype(np.float32)).cuda().requires_grad_(True) for w in Ws]
    bt = [torch.from_numpy(b.astype(np.float32)).cuda().requires_grad_(True) for b in bs]
    y = F.siren_mlp(torch.from_numpy(x).cuda(), Wt, bt, w0=30.0, precision="bf16", fourier=B)
    y.backward(torch.ones_like(y) / n)
    torch.cuda.synchronize()
    print("case", (d, o, tasks, per_task, n, ff), "ok", float(y.abs().mean()), float(Wt[0].grad.abs().mean()), flush=True)
