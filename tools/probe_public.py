"""Developer probe: per-step wall times of the public path after the legs bench.py runs before it."""
import sys, time
sys.path.insert(0, '.')
import torch
from tools import workloads
import siren_mri_b200
from siren_mri_b200 import modules, diff_operators
dev = torch.device('cuda')

def public(cfg, prec, n=14):
    torch.manual_seed(cfg)
    x, gt, d, o, derivs, clip = workloads.make_inputs(cfg, dev, 8)
    model = modules.SingleBVPNet(in_features=d, out_features=o, precision=prec, coord_derivs=derivs).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    times = []
    for i in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = model({'coords': x})
        loss = workloads.laplace_mse(out, gt, diff_operators.laplace) if cfg == 4 else workloads.image_mse(out, gt)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        torch.cuda.synchronize(); times.append(round((time.perf_counter() - t0) * 1e3, 2))
    print('public', cfg, prec, times, flush=True)
    del model, opt, x, gt, out, loss
    siren_mri_b200.functional.clear_workspace_cache(); torch.cuda.empty_cache()

public(2, 'fp32')
print(workloads.run_config(2, 'eager', 'fp32', steps=3, warmup=2, dev=dev)['ms_per_step'], flush=True)
public(2, 'fp32')
print(workloads.run_trainer_config(1, 'bf16', dev=dev)['ms_per_step'], flush=True)
public(2, 'fp32')
print(workloads.run_trainer_config(1, 'fp32', dev=dev)['ms_per_step'], flush=True)
public(2, 'fp32')
public(4, 'fp32', 8)
