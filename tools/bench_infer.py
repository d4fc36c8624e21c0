#!/usr/bin/env python
"""Forward-only (torch.no_grad) throughput, the sdf_meshing.create_mesh / summary-writer use of the path
(sdf_meshing.py:46-56: N^3 samples in 64^3 chunks through decoder(sample_subset))."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import siren_mri_b200  # noqa: E402
from siren_mri_b200 import modules  # noqa: E402


def run(n, d, precision, backend):
    siren_mri_b200.set_defaults(backend=backend)
    torch.manual_seed(0)
    m = modules.SingleBVPNet(in_features=d, out_features=1, precision=precision if backend == "auto" else "fp32").cuda()
    x = torch.rand((n, d), device="cuda") * 2 - 1          # 2-D input like sdf_meshing
    with torch.no_grad():
        for _ in range(3):
            m.net(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            y = m.net(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"n": n, "d": d, "mode": precision if backend == "auto" else "eager-fp32", "ms": ms,
                      "Mcoord_per_s": n / ms / 1e3, "finite": bool(torch.isfinite(y).all())}), flush=True)


if __name__ == "__main__":
    for n in (262144, 64 ** 3 * 16):
        run(n, 3, "bf16", "auto")
        run(n, 3, "fp32", "auto")
        run(n, 3, "fp32", "composed")
