#!/usr/bin/env python
"""Debug aid: SIREN_FUSED_DBG=1 makes the fused forward dump a clock64 trace of CTA 0 (api.cu)."""
import os
import sys

os.environ["SIREN_FUSED_DBG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from siren_mri_b200 import modules  # noqa: E402

torch.manual_seed(0)
m = modules.SingleBVPNet(in_features=2, out_features=1, precision="bf16").cuda()
x = torch.rand((1, 262144, 2), device="cuda") * 2 - 1
print("== inference", file=sys.stderr, flush=True)
with torch.no_grad():
    m.net(x)
    m.net(x)
torch.cuda.synchronize()
print("== training forward", file=sys.stderr, flush=True)
y = m.net(x)
torch.cuda.synchronize()
print("== backward", file=sys.stderr, flush=True)
y.sum().backward()
torch.cuda.synchronize()
