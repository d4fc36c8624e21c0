#!/usr/bin/env python
"""Where does the end-to-end step lose time against the device-only step?  Times variants of the pipelined loop
(SirenTrainer.submit_from_host) with pieces switched off.  cfg2, one GPU."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from siren_mri_b200 import modules  # noqa: E402
from siren_mri_b200.trainer import SirenTrainer  # noqa: E402

torch.manual_seed(0)
N = 262144
m = modules.SingleBVPNet(in_features=2, out_features=1, precision="bf16").cuda()
tr = SirenTrainer(m, N)
xh = (torch.rand((1, N, 2)) * 2 - 1).pin_memory()
gh = torch.rand((1, N, 1)).pin_memory()
K = 300


def timed(fn, label):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(K):
        fn()
    torch.cuda.synchronize()
    print("%-58s %.1f us/step" % (label, (time.perf_counter() - t) / K * 1e6), flush=True)


timed(tr.step, "step() (graph replay only)")
evs = [torch.cuda.Event() for _ in range(4)]
cnt = [0]


def step_paced(lag):
    def f():
        i = cnt[0]
        cnt[0] += 1
        tr.step()
        evs[i % 4].record()
        if i >= lag:
            evs[(i - lag) % 4].synchronize()
    return f


timed(step_paced(1), "step() + host waits for the previous step (depth 2)")
timed(step_paced(2), "step() + host waits for step k-2 (depth 3)")
timed(tr.step, "step() again (unpaced)")
prev = [None]


def pipe():
    h = tr.submit_from_host(xh, gh)
    if prev[0] is not None:
        prev[0].result()
    prev[0] = h


timed(pipe, "submit_from_host, loss read one step late")
prev[0] = None


def pipe_nowait():
    tr.submit_from_host(xh, gh)


timed(pipe_nowait, "submit_from_host, loss never read")

# graph of a slot + events, no upload
s = tr._slots[0]


def slot_graph():
    s["graphs"][(True, 1, 0)].replay()


timed(slot_graph, "slot graph replay (loss published by a kernel), no upload")
ev = torch.cuda.Event()


def slot_graph_ev():
    s["graphs"][(True, 1, 0)].replay()
    ev.record()


timed(slot_graph_ev, "  + event record per step")
cs = torch.cuda.Stream()


def upload_only():
    with torch.cuda.stream(cs):
        s["coords"].copy_(xh, non_blocking=True)
        s["gt"][0].copy_(gh, non_blocking=True)


timed(upload_only, "upload only (copy stream)")


def both_unsynced():
    with torch.cuda.stream(cs):
        tr._slots[1]["coords"].copy_(xh, non_blocking=True)
        tr._slots[1]["gt"][0].copy_(gh, non_blocking=True)
    s["graphs"][(True, 1, 0)].replay()


timed(both_unsynced, "slot graph + upload into another slot, no dependencies")
