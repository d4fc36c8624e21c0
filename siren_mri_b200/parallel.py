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
import ctypes

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


class PeerGradientReducer:
    """Gradient exchange of a data-parallel model whose gradients are too many for the fused Adam kernel's replicated
    read (the 31.3 M parameters of the neural-process models: ConvImgEncoder + HyperNetwork,
    train_mri_neural_process_ddp.py:238): the DDP Reducer's bucketed ``ncclAllReduce`` becomes ONE kernel per rank over
    NVLink peer memory.

    Every parameter's ``.grad`` is a view into one flat buffer allocated in symmetric memory
    (``torch.distributed._symmetric_memory``: peer-mapped on every rank of the box), so ``backward()`` accumulates
    straight into it -- nothing is packed or unpacked.  ``reduce()`` = barrier, ``siren_b200_allreduce_peers`` (rank r
    sums slice r over all ranks' buffers and stores the total into every rank's slice r: reduce-scatter + all-gather in
    one pass, all replicas receive the same bits), barrier.  On an NVSwitch fabric with a multicast mapping (NVLS) the
    kernel is ``siren_b200_allreduce_multicast`` instead: ``multimem.ld_reduce`` sums an element inside the switch and
    ``multimem.st`` broadcasts the total, so every rank moves its slice once per direction.  Where symmetric memory is not available (other backends,
    CPU, no peer access) the same object keeps a plain flat buffer and calls ``torch.distributed.all_reduce`` on it.

    Use ``reducer.zero_grad()`` instead of ``optimizer.zero_grad()`` (which would drop the views)."""

    def __init__(self, params, group=None, average=True, force_fallback=False, multicast=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.scale = 1.0 / self.world if average else 1.0
        n = sum(p.numel() for p in self.params)
        self.n = n
        self.n_pad = (n + 1023) // 1024 * 1024
        dev = self.params[0].device
        self.device = dev
        self.hdl, self.peers, self.mc = None, None, 0
        self.flat = None
        if dev.type == "cuda" and self.world > 1 and not force_fallback:
            ok = 1
            try:
                import torch.distributed._symmetric_memory as symm
                buf = symm.empty(self.n_pad, dtype=torch.float32, device=dev)
                hdl = symm.rendezvous(buf, group if group is not None else dist.group.WORLD)
                ptrs = [int(p) for p in hdl.buffer_ptrs]
                if len(ptrs) != self.world or any(p == 0 for p in ptrs):
                    ok = 0
            except Exception as e:      # pragma: no cover  (no NVLink peer access, old driver, ...)
                ok = 0
                self.error = repr(e)
            flag = torch.tensor([ok], device=dev, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)      # the decision is the same on every rank
            if int(flag.item()):
                self.flat, self.hdl = buf, hdl
                self.peers = torch.tensor(ptrs, dtype=torch.int64, device=dev)
                # NVSwitch multicast mapping of the same allocation (NVLS), when the fabric has one: the reduction
                # then happens inside the switch (siren_b200_allreduce_multicast); multicast=False keeps the peer kernel
                mc = 0
                try:
                    # (default: from three ranks up -- with two the link traffic is the same either way and the peer
                    #  kernel measured faster: 0.21 against 0.33 ms for 125 MB)
                    want = multicast if multicast is not None else self.world > 2
                    mc = int(hdl.multicast_ptr) if want else 0
                except Exception:      # pragma: no cover
                    mc = 0
                mflag = torch.tensor([1 if mc else 0], device=dev, dtype=torch.int32)
                dist.all_reduce(mflag, op=dist.ReduceOp.MIN, group=group)
                self.mc = mc if int(mflag.item()) else 0
        if self.flat is None:
            self.flat = torch.empty(self.n_pad, dtype=torch.float32, device=dev)
        self.flat.zero_()
        off = 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise TypeError("PeerGradientReducer: fp32 parameters only")
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        if self.hdl is not None:
            torch.cuda.synchronize(dev)
            dist.barrier(group=group)      # every rank's buffer is zero before anyone's first reduce can read it

    @property
    def fused(self):
        return self.hdl is not None

    def zero_grad(self):
        self.flat.zero_()

    def reduce(self):
        """Sum (``average``: mean) the flat gradient over the ranks, in place on every rank."""
        if self.world == 1:
            return self.flat
        if self.hdl is None:
            dist.all_reduce(self.flat, group=self.group)
            if self.scale != 1.0:
                self.flat.mul_(self.scale)
            return self.flat
        from . import _lib
        lib = _lib.load()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.hdl.barrier(channel=0)      # every rank's backward has finished writing its buffer
        with torch.cuda.device(self.device):
            if self.mc:
                rc = lib.siren_b200_allreduce_multicast(ctypes.c_void_p(self.mc), self.world, self.rank, self.n_pad,
                                                        self.scale, stream)
            else:
                rc = lib.siren_b200_allreduce_peers(_lib.dptr(self.peers), self.world, self.rank, self.n_pad,
                                                    self.scale, stream)
        _lib.check(rc, "siren_b200_allreduce_peers")
        self.hdl.barrier(channel=0)      # every rank's slice has landed in every buffer
        return self.flat
