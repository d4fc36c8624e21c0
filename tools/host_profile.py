#!/usr/bin/env python
"""Host-side profile (cProfile) of one cfg2 training step through the PUBLIC module API: where the Python time of
the host-bound path goes (model(...) -> loss -> backward -> torch.optim.Adam)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from siren_mri_b200 import modules  # noqa: E402

torch.manual_seed(0)
dev = torch.device("cuda")
n = 262144
model = modules.SingleBVPNet(in_features=2, out_features=1, precision="bf16").to(dev)
x = torch.rand((1, n, 2), device=dev) * 2 - 1
gt = torch.rand((1, n, 1), device=dev) * 2 - 1
opt = torch.optim.Adam(model.parameters(), lr=1e-4)


def step():
    out = model({"coords": x})
    loss = ((out["model_out"] - gt) ** 2).sum() / 16384.0
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss


for _ in range(20):
    step()
torch.cuda.synchronize()
K = 300
t = time.perf_counter()
for _ in range(K):
    step()
t_host = time.perf_counter() - t
torch.cuda.synchronize()
t_all = time.perf_counter() - t
print("host enqueue %.1f us/step, with GPU drain %.1f us/step" % (t_host / K * 1e6, t_all / K * 1e6))


def part(fn, label):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(K):
        fn()
    print("  %-28s %.1f us" % (label, (time.perf_counter() - t) / K * 1e6))
    torch.cuda.synchronize()


out = model({"coords": x})
part(lambda: model({"coords": x}), "forward (model call)")
part(lambda: ((out["model_out"] - gt) ** 2).sum() / 16384.0, "loss")
part(lambda: opt.step(), "optimizer step")


def fb():
    o = model({"coords": x})
    (((o["model_out"] - gt) ** 2).sum() / 16384.0).backward()


part(fb, "forward + loss + backward")
pr = cProfile.Profile()
pr.enable()
for _ in range(K):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
