#!/usr/bin/env python
"""Developer probe: SIREN_WGRAD_DBG=1 makes the weight-gradient launch print, per item kind (layer), when its CTAs
started and finished (globaltimer): shows how evenly the split-K items of the layers end."""
import os
import sys

os.environ["SIREN_WGRAD_DBG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from siren_mri_b200 import modules  # noqa: E402

torch.manual_seed(0)
if len(sys.argv) > 1 and sys.argv[1] == "cfg5":      # 8 tasks x 65,536, 16 inputs, per-task weights
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import workloads
    m = modules.SingleBVPNet(in_features=16, out_features=2, precision="bf16").cuda()
    params = workloads.per_task_params(m, 8, torch.device("cuda"))
    x = torch.rand((8, 65536, 16), device="cuda") * 2 - 1
    run = lambda: m({"coords": x}, params=params)["model_out"]      # noqa: E731
else:
    m = modules.SingleBVPNet(in_features=2, out_features=1, precision="bf16").cuda()
    x = torch.rand((1, 262144, 2), device="cuda") * 2 - 1
    run = lambda: m.net(x)      # noqa: E731
for _ in range(3):
    y = run()
    y.sum().backward()
    torch.cuda.synchronize()
    print("--", file=sys.stderr, flush=True)
