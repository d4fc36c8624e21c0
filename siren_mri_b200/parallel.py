"""Coordinate sharding across the GPUs of one box.

Every coordinate row is independent through forward, dgrad and the jets; only dW/db (sums over
coordinates) couple them.  So: replicate the weights, give each rank a contiguous, equal shard of
the coordinate batch, fold the GLOBAL loss normalisation into the local loss, and sum the flat
gradient buffer with one all-reduce.  This replaces the DDP-Reducer path of
train_mri_neural_process_ddp.py:238 / training_ddp.py:155-164 for the single-scene configs.
For the neural-process config the unit of sharding is the task (the reference's own
DistributedSampler semantics, train_mri_neural_process_ddp.py:188-189) and the hot path needs no
collective at all.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous [begin, end) of ``n`` units for ``rank``; sizes differ by at most one."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_coords(coords, *others, rank=None, world=None, dim=1):
    """Slice ``coords`` ([B, N, d]) and companions along the coordinate axis for this rank."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    b, e = shard_bounds(coords.shape[dim], rank, world)
    idx = [slice(None)] * coords.dim()
    idx[dim] = slice(b, e)
    out = [coords[tuple(idx)].contiguous()]
    for t in others:
        out.append(t[tuple(idx)].contiguous())
    return out if others else out[0]


def shard_tasks(n_tasks, rank=None, world=None):
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    return shard_bounds(n_tasks, rank, world)


def allreduce_gradients(params, group=None, flat=None):
    """Sum gradients over ranks with ONE collective.  ``flat``: an existing flat grad buffer
    (siren_mri_b200.optim.flatten_parameters); otherwise grads are packed, reduced and unpacked."""
    if flat is not None:
        dist.all_reduce(flat, group=group)
        return flat
    grads = [p.grad for p in params if p.grad is not None]
    buf = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(buf, group=group)
    off = 0
    for g in grads:
        k = g.numel()
        g.copy_(buf[off:off + k].view_as(g))
        off += k
    return buf
