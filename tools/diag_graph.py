"""Diagnostic: graph-captured trainer at small n, syncing after every replay."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import siren_oracle as so
from siren_mri_b200 import modules
from siren_mri_b200.trainer import SirenTrainer

prec = sys.argv[1]; n = int(sys.argv[2]); sync = int(sys.argv[3])
Ws, bs = so.make_params(2, 256, 3, 1, seed=21)
x = so.make_coords(1, n, 2, seed=22)
m = modules.SingleBVPNet(in_features=2, out_features=1, precision=prec).cuda()
tr = SirenTrainer(m, n, lr=1e-4, use_graph=True)
tr.coords.copy_(torch.from_numpy(x)); tr.gt.copy_(torch.from_numpy(x[..., :1]))
for i in range(12):
    tr.step()
    if sync:
        torch.cuda.synchronize()
    print("replay", i, "ok", flush=True)
torch.cuda.synchronize()
print("done", prec, n, sync, float(tr.loss.item()), flush=True)
