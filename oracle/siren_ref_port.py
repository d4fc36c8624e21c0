"""CPU port of the reference's training step for timing.  TEST / BENCH INFRASTRUCTURE ONLY.

The reference executes this path as PyTorch CPU ops (it has no native code): BatchLinear is
``input.matmul(weight^T) + bias`` (modules.py:25-26), Sine is ``torch.sin(w0 * x)`` (modules.py:38),
the backward is torch autograd (training.py:91) and the update is ``torch.optim.Adam``
(training.py:23, 101-103).  /root/reference does not exist on the GPU box, so this file restates
exactly those ops for ``bench.py --impl reference`` and the ``cpu_baseline`` leg (kind "port").
It is pinned to the live reference by tests/test_oracle_golden.py::test_ref_port_matches_reference.
Only tests/ and bench.py import it.
"""
import time

import torch


class RefPortSiren(torch.nn.Module):
    """SingleBVPNet(type='sine', mode='mlp') restated with stock torch.nn.Linear layers."""

    def __init__(self, d_in=2, hidden=256, n_hidden=3, d_out=1, w0=30.0):
        super().__init__()
        dims = [d_in] + [hidden] * (n_hidden + 1) + [d_out]
        self.lin = torch.nn.ModuleList(torch.nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1))
        self.w0 = w0

    def load_numpy(self, Ws, bs):
        with torch.no_grad():
            for lin, W, b in zip(self.lin, Ws, bs):
                lin.weight.copy_(torch.as_tensor(W))
                lin.bias.copy_(torch.as_tensor(b))

    def forward(self, model_input):
        coords = model_input["coords"].clone().detach().requires_grad_(True)     # modules.py:151
        h = coords
        for i, lin in enumerate(self.lin):
            h = h.matmul(lin.weight.transpose(-1, -2))                           # modules.py:25
            h = h + lin.bias.unsqueeze(-2)                                       # modules.py:26
            if i != len(self.lin) - 1:
                h = torch.sin(self.w0 * h)                                       # modules.py:38
        return {"model_in": coords, "model_out": h}


def train_step(model, optim, coords, gt, weight=1.0 / 16384.0):
    """training.py:66-103 for image_mse(high_freq=False)."""
    out = model({"coords": coords})
    loss = ((out["model_out"] - gt) ** 2).sum() * weight
    optim.zero_grad()
    loss.backward()
    optim.step()
    return loss


def time_steps(n_coords, steps, warmup, d_in=2, d_out=1, threads=None, seed=0):
    """Seconds per step of the CPU port on ``n_coords`` coordinates (all host threads)."""
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(seed)
    model = RefPortSiren(d_in=d_in, d_out=d_out)
    optim = torch.optim.Adam(lr=1e-4, params=model.parameters())
    g = torch.Generator().manual_seed(seed)
    coords = torch.rand((1, n_coords, d_in), generator=g) * 2 - 1
    gt = torch.rand((1, n_coords, d_out), generator=g) * 2 - 1
    for _ in range(warmup):
        train_step(model, optim, coords, gt)
    t0 = time.perf_counter()
    for _ in range(steps):
        train_step(model, optim, coords, gt)
    return (time.perf_counter() - t0) / max(steps, 1)
