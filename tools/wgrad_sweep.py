#!/usr/bin/env python
"""Developer aid: time the weight-gradient kernel of cfg2 (bf16) for several split-K factors (SIREN_WGRAD_SLICES)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from siren_mri_b200 import _lib, modules  # noqa: E402
from siren_mri_b200.trainer import SirenTrainer  # noqa: E402

torch.manual_seed(0)
N = 262144
lib = _lib.load()
m = modules.SingleBVPNet(in_features=2, out_features=1, precision="bf16").cuda()
tr = SirenTrainer(m, N, use_graph=False)
tr.coords.copy_(torch.rand(1, N, 2) * 2 - 1)
tr.gt.copy_(torch.rand(1, N, 1))
for s in [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "0,37,49,74,98,148,197,296").split(",")]:
    if s:
        os.environ["SIREN_WGRAD_SLICES"] = str(s)
    else:
        os.environ.pop("SIREN_WGRAD_SLICES", None)
    for _ in range(3):
        tr.step()
    torch.cuda.synchronize()
    lib.siren_b200_profile_begin()
    for _ in range(20):
        tr.step()
    buf = ctypes.create_string_buffer(1 << 16)
    lib.siren_b200_profile_end(buf, len(buf))
    t = {ln.split()[0]: 1e3 * float(ln.split()[2]) / int(ln.split()[1]) for ln in buf.value.decode().strip().splitlines()}
    print("slices %4s  wgrad %.1f us  fwd %.1f  chain %.1f  adam %.1f" % (s or "auto", t["wgrad"], t["mlp_fused_fwd"],
                                                                          t["mlp_fused_bwd"], t["adam_step"]), flush=True)
