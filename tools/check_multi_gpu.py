#!/usr/bin/env python
"""Multi-GPU self-check (run under torch.distributed.run, one rank per GPU):
  1. siren_b200_allreduce (C ABI, NCCL) equals torch.distributed.all_reduce;
  2. after 5 graph-captured steps on different coordinate shards the replicas' weights are bit-identical;
  3. the sharded step equals the single-GPU full-batch step (rank 0 recomputes it) to fp32-mode tolerance.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from siren_mri_b200 import modules, parallel  # noqa: E402
from siren_mri_b200.trainer import SirenTrainer  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 40000
    g = torch.Generator().manual_seed(0)
    x = torch.rand((1, n, 2), generator=g) * 2 - 1
    gt = torch.rand((1, n, 1), generator=g) * 2 - 1
    b, e = parallel.shard_bounds(n, rank, world)

    def make(nloc, pg_comm, distributed=True):
        torch.manual_seed(0)
        m = modules.SingleBVPNet(in_features=2, out_features=1, precision="fp32").to(dev)
        return m, SirenTrainer(m, nloc, lr=1e-4, loss_weight=1.0 / n, comm=pg_comm, distributed=distributed)

    import sys as _sys
    mode = _sys.argv[1] if len(_sys.argv) > 1 else "c_abi"
    m, tr = make(e - b, mode)
    if tr.comm is None:          # fused all-reduce path: the C-ABI communicator only exists for check 1
        tr.comm = tr._make_comm()
    # 1. C-ABI all-reduce vs torch.distributed
    a = torch.full((1000,), float(rank + 1), device=dev)
    ref = a.clone()
    dist.all_reduce(ref)
    from siren_mri_b200 import _lib
    _lib.check(tr.lib.siren_b200_allreduce(tr.comm, _lib.dptr(a), a.numel(), torch.cuda.current_stream().cuda_stream),
               "allreduce")
    torch.cuda.synchronize()
    assert torch.equal(a, ref), "C-ABI all-reduce differs from torch.distributed"
    # 2. replicas stay identical
    tr.coords.copy_(x[:, b:e])
    tr.gt.copy_(gt[:, b:e])
    for _ in range(5):
        tr.step()
    torch.cuda.synchronize()
    flat = tr.flat.clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    for r in range(1, world):
        assert torch.equal(gathered[0], gathered[r]), "replica %d diverged" % r
    # 3. equals the full-batch single-GPU step
    if rank == 0:
        m1, tr1 = make(n, None, distributed=False)
        tr1.coords.copy_(x)
        tr1.gt.copy_(gt)
        for _ in range(5):
            tr1.step()
        torch.cuda.synchronize()
        du = (flat - tr1.flat).norm() / (tr1.flat.norm())
        assert du < 1e-4, float(du)
        print("multi-GPU check OK (%s, all-reduce %s): world=%d, replicas identical, |sharded - full| / |full| = %.2e"
              % (mode, "fused over peer memory" if tr.p2p is not None else "NCCL", world, float(du)), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
