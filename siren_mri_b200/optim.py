"""Fused (clip_grad_norm_ +) Adam over one flat fp32 buffer.

Mirrors ``torch.optim.Adam(lr=lr, params=model.parameters())`` as constructed at
training.py:23 (betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad) and the optional
``torch.nn.utils.clip_grad_norm_`` of training.py:93-97, in two kernel launches instead of a
foreach pass over ten small tensors.
"""
import torch

from . import _lib


def flatten_parameters(params, grad=None):
    """Re-home ``params`` as views of one flat buffer; returns (flat_param, flat_grad).  ``grad``: a preallocated
    (zeroed) flat gradient buffer to use, e.g. one half of a symmetric-memory allocation peers can read."""
    params = list(params)
    n = sum(p.numel() for p in params)
    dev, dt = params[0].device, params[0].dtype
    flat = torch.empty(n, device=dev, dtype=dt)
    if grad is None:
        grad = torch.zeros(n, device=dev, dtype=dt)
    off = 0
    with torch.no_grad():
        for p in params:
            k = p.numel()
            flat[off:off + k].copy_(p.reshape(-1))
            p.data = flat[off:off + k].view_as(p)
            p.grad = grad[off:off + k].view_as(p)
            off += k
    return flat, grad


class FusedAdam:
    """Adam on a flat buffer.  ``step()`` consumes ``flat_grad`` (optionally clipping its global
    L2 norm to ``max_grad_norm`` and scaling by ``grad_scale`` first)."""

    def __init__(self, flat_param, flat_grad, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=0.0):
        if not flat_param.is_cuda:
            raise _lib.NativeError("FusedAdam needs CUDA tensors")
        self.p, self.g = flat_param, flat_grad
        self.m = torch.zeros_like(flat_param)
        self.v = torch.zeros_like(flat_param)
        self.state = torch.zeros(64, device=flat_param.device, dtype=torch.uint8)   # AdamState (csrc/simt.h)
        self.lr, self.betas, self.eps = lr, betas, eps
        self.max_grad_norm = float(max_grad_norm)
        self.steps = 0

    def step(self, grad_scale=1.0):
        lib = _lib.load()
        self.steps += 1          # host mirror; the authoritative counter lives in self.state
        stream = torch.cuda.current_stream(self.p.device).cuda_stream
        with torch.cuda.device(self.p.device):
            rc = lib.siren_b200_adam(_lib.dptr(self.p), _lib.dptr(self.g), _lib.dptr(self.m), _lib.dptr(self.v),
                                     self.p.numel(), self.lr, self.betas[0], self.betas[1], self.eps,
                                     self.max_grad_norm, float(grad_scale), _lib.dptr(self.state), stream)
        _lib.check(rc, "siren_b200_adam")

    def step_fused(self, grad_scale=1.0, zero_grad=True, loss4=None, desc=None, w_ptrs=None, ws=None, clip=True):
        """The same update in ONE launch (siren_b200_adam_step): step tick, clip, Adam, the consumed gradient cleared,
        ``loss4[1] -> loss4[0]``, and (``desc`` / ``w_ptrs`` / ``ws`` given) the workspace's bf16 weight copies
        refreshed from the updated parameters -- the optimizer tail of SirenTrainer's captured step."""
        lib = _lib.load()
        self.steps += 1
        stream = torch.cuda.current_stream(self.p.device).cuda_stream
        with torch.cuda.device(self.p.device):
            rc = lib.siren_b200_adam_step(_lib.dptr(self.p), _lib.dptr(self.g), _lib.dptr(self.m), _lib.dptr(self.v),
                                          self.p.numel(), self.lr, self.betas[0], self.betas[1], self.eps,
                                          self.max_grad_norm if clip else 0.0, float(grad_scale), _lib.dptr(self.state),
                                          1 if zero_grad else 0, _lib.dptr(loss4), desc, w_ptrs, _lib.dptr(ws), stream)
        _lib.check(rc, "siren_b200_adam_step")

    def step_peers(self, peers_dev, world, zero_buf, grad_scale=1.0, loss4=None, desc=None, w_ptrs=None, ws=None,
                   clip=True):
        """step_fused with the all-reduce fused in (siren_b200_adam_step_peers): the gradient is summed over the ranks'
        buffers read from peer memory; ``zero_buf`` (the buffer the next step accumulates into) is cleared."""
        lib = _lib.load()
        self.steps += 1
        stream = torch.cuda.current_stream(self.p.device).cuda_stream
        with torch.cuda.device(self.p.device):
            rc = lib.siren_b200_adam_step_peers(_lib.dptr(self.p), _lib.dptr(self.g), _lib.dptr(self.m), _lib.dptr(self.v),
                                                self.p.numel(), self.lr, self.betas[0], self.betas[1], self.eps,
                                                self.max_grad_norm if clip else 0.0, float(grad_scale),
                                                _lib.dptr(self.state), _lib.dptr(loss4), desc, w_ptrs, _lib.dptr(ws),
                                                _lib.dptr(peers_dev), int(world), _lib.dptr(zero_buf), stream)
        _lib.check(rc, "siren_b200_adam_step_peers")

    def zero_grad(self):
        self.g.zero_()
