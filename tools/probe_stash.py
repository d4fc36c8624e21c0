#!/usr/bin/env python
"""Developer probe: one fused training forward + backward with a synchronize after each, errors against the fp64 oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import siren_oracle as so  # noqa: E402
from siren_mri_b200 import functional as F  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def main():
    d, nh, o, n = 2, 3, 1, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    Ws, bs = so.make_params(d, 256, nh, o, seed=0, tasks=0)
    Ws = [w.astype(np.float32) for w in Ws]
    bs = [b.astype(np.float32) for b in bs]
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, (1, n, d)).astype(np.float32)
    gy = rng.standard_normal((1, n, o)).astype(np.float32) / n
    xt = torch.from_numpy(x).cuda()
    Wt = [torch.from_numpy(w).cuda().requires_grad_(True) for w in Ws]
    bt = [torch.from_numpy(b).cuda().requires_grad_(True) for b in bs]
    y = F.siren_mlp(xt, Wt, bt, w0=30.0, precision="bf16")
    torch.cuda.synchronize()
    print("forward ok", flush=True)
    W64 = [w.astype(np.float64) for w in Ws]
    b64 = [b.astype(np.float64) for b in bs]
    y64, _, _, cache = so.siren_forward(x.astype(np.float64), W64, b64, 30.0, order=0)
    print("y rel", rel(y.detach().cpu().numpy(), y64), flush=True)
    y.backward(torch.from_numpy(gy).cuda())
    torch.cuda.synchronize()
    print("backward ok", flush=True)
    dWs, dbs, _ = so.siren_backward(cache, W64, gy.astype(np.float64))
    for l in range(len(Ws)):
        print("layer %d dW %.2e db %.2e" % (l, rel(Wt[l].grad.cpu().numpy(), dWs[l]), rel(bt[l].grad.cpu().numpy(), dbs[l])))


if __name__ == "__main__":
    main()
