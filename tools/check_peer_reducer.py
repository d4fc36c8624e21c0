#!/usr/bin/env python
"""Multi-GPU check + timing of parallel.PeerGradientReducer (run under torch.distributed.run, one rank per GPU):
the peer-memory all-reduce kernel against torch.distributed.all_reduce (NCCL) on a 31.3 M-parameter flat gradient --
the size of the neural-process models' DDP exchange (train_mri_neural_process_ddp.py:238)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from siren_mri_b200 import parallel  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(rank)
    # 31.3 M parameters in tensors of assorted sizes (ragged ends exercise the padded tail)
    sizes = [256 * 256 * 25] * 8 + [65536 * 128] * 2 + [128 * 128] * 88 + [4099, 7, 1]
    params = [torch.nn.Parameter(torch.zeros(s, device=dev)) for s in sizes]
    red = parallel.PeerGradientReducer(params, average=True)
    n = red.n
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    local_grad = torch.randn(red.n_pad, device=dev, generator=g)
    local_grad[n:] = 0
    ref = local_grad.clone()
    dist.all_reduce(ref)
    ref /= world
    red.flat.copy_(local_grad)
    red.reduce()
    torch.cuda.synchronize()
    err = float((red.flat - ref).abs().max())
    # every replica holds the same bits
    gathered = [torch.empty_like(red.flat) for _ in range(world)]
    dist.all_gather(gathered, red.flat)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    assert red.fused, "symmetric memory not available: " + getattr(red, "error", "?")
    assert err < 1e-5 and same, (err, same)
    assert all(p.grad.data_ptr() >= red.flat.data_ptr() for p in params)

    def timed(fn, steps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    scratch = local_grad.clone()
    t_auto = timed(red.reduce)
    t_nccl = timed(lambda: (dist.all_reduce(scratch), scratch.mul_(1.0 / world)))
    res = {"world": world, "floats": n, "multicast": bool(red.mc), "max_abs_err_vs_nccl": err, "replicas_identical": same,
           ("multicast_kernel_ms" if red.mc else "peer_kernel_ms"): round(t_auto, 4),
           "nccl_allreduce_plus_scale_ms": round(t_nccl, 4)}
    if red.mc:      # the plain peer-memory kernel on a second buffer, for comparison (and its own correctness check)
        params2 = [torch.nn.Parameter(torch.zeros(s, device=dev)) for s in sizes]
        red2 = parallel.PeerGradientReducer(params2, average=True, multicast=False)
        red2.flat.copy_(local_grad)
        red2.reduce()
        torch.cuda.synchronize()
        res["peer_max_abs_err_vs_nccl"] = float((red2.flat - ref).abs().max())
        res["peer_kernel_ms"] = round(timed(red2.reduce), 4)
    if rank == 0:
        gb = n * 4 / 1e9
        for k in ("multicast_kernel_ms", "peer_kernel_ms", "nccl_allreduce_plus_scale_ms"):
            if k in res:
                res[k.replace("_ms", "_algbw_GBs")] = round(gb / res[k] * 1e3, 1)
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
