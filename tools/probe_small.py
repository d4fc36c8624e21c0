#!/usr/bin/env python
"""Developer probe: the captured training step and its per-kernel times at small batches (the per-GPU share of the
strong-scaling leg: 32,768 coordinates at 8 GPUs; cfg1: 65,536)."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from siren_mri_b200 import _lib, modules  # noqa: E402
from siren_mri_b200.trainer import SirenTrainer  # noqa: E402

lib = _lib.load()
for n in (int(a) for a in (sys.argv[1:] or ["32768", "65536", "262144"])):
    torch.manual_seed(0)
    m = modules.SingleBVPNet(in_features=2, out_features=1, precision="bf16").cuda()
    tr = SirenTrainer(m, n, lr=1e-4, loss_weight=1.0 / n, distributed=False)
    tr.coords.copy_(torch.rand((1, n, 2), device="cuda") * 2 - 1)
    tr.gt.copy_(torch.rand((1, n, 1), device="cuda") * 2 - 1)
    for _ in range(10):
        tr.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        tr.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 200
    tr.use_graph = False
    tr.step()
    torch.cuda.synchronize()
    lib.siren_b200_profile_begin()
    for _ in range(10):
        tr.step()
    torch.cuda.synchronize()
    buf = ctypes.create_string_buffer(1 << 16)
    lib.siren_b200_profile_end(buf, len(buf))
    k = {ln.split()[0]: round(1e3 * float(ln.split()[2]) / int(ln.split()[1]), 1) for ln in buf.value.decode().strip().splitlines()}
    print(json.dumps({"n": n, "graph_step_us": round(1e3 * ms, 1), "kernels_us": k}), flush=True)
    del tr, m
