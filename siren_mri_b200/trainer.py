"""Fast training step for the image-fitting configurations: forward, MSE loss gradient, backward,
gradient all-reduce, clip + Adam -- one CUDA graph, no host synchronisation.

It is the sibling of the reference loop body at training.py:66-103
(``model(model_input)`` -> ``image_mse`` -> ``backward`` -> ``clip_grad_norm_`` -> ``Adam.step``)
with identical arithmetic, for callers that do not need the per-step tensorboard scalars
(training.py:77-81 force a device->host sync every step).  The model is an ordinary
``modules.SingleBVPNet`` / ``FCBlock``; its parameters are re-homed into one flat buffer so that
the all-reduce is a single collective and Adam a single kernel.

Multi-GPU: every rank holds the full weights and a contiguous shard of the coordinates; the
loss weight carries the GLOBAL normalisation, gradients are summed with one flat all-reduce
(``torch.distributed``, NCCL over NVLink) and every rank applies the same Adam update, so the
replicas stay bit-identical without a broadcast.
"""
import os

import torch

from . import _lib
from .optim import FusedAdam, flatten_parameters


def _fcblock_of(model):
    net = model
    while not hasattr(net, "_n_layers"):
        net = net.net
    return net


class SirenTrainer:
    def __init__(self, model, n_coords, lr=1e-4, loss_weight=None, max_grad_norm=0.0, precision=None,
                 process_group=None, use_graph=True, comm="c_abi"):
        self.block = _fcblock_of(model)
        if not self.block._sine:
            raise ValueError("SirenTrainer needs a sine FCBlock")
        self.lib = _lib.load()
        params = list(self.block.parameters())
        self.device = params[0].device
        if self.device.type != "cuda":
            raise _lib.NativeError("SirenTrainer needs the model on a CUDA device")
        self.flat, self.grad = flatten_parameters(params)
        self.weights = [self.block.net[l][0].weight for l in range(self.block._n_layers)]
        self.biases = [self.block.net[l][0].bias for l in range(self.block._n_layers)]
        self.opt = FusedAdam(self.flat, self.grad, lr=lr, max_grad_norm=max_grad_norm)
        self.pg = process_group
        self.world = 1
        self.comm = None
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        if self.world > 1 and comm == "c_abi":
            self.comm = self._make_comm()
        d_in = self.weights[0].shape[1]
        d_out = self.weights[-1].shape[0]
        self.n = int(n_coords)
        self.precision = precision or self.block._opt("precision")
        desc = _lib.SirenDesc()
        desc.d_in, desc.hidden, desc.n_hidden, desc.d_out = d_in, self.weights[0].shape[0], self.block._n_layers - 2, d_out
        desc.w0, desc.tasks, desc.per_task, desc.n_coords = self.block._w0, 1, 0, self.n
        desc.precision, desc.deriv_order = _lib.PRECISIONS[self.precision], 0
        self.desc = desc
        nbytes = self.lib.siren_b200_workspace_bytes(desc)
        if nbytes == 0:
            _lib.check(1, "siren_b200_workspace_bytes")
        dev = self.device
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.coords = torch.zeros((1, self.n, d_in), device=dev)
        self.gt = torch.zeros((1, self.n, d_out), device=dev)
        self.y = torch.empty((1, self.n, d_out), device=dev)
        self.gy = torch.empty_like(self.y)
        self.loss = torch.zeros(1, device=dev)
        # image_mse (loss_functions.py:88): sum of squares / 16384 regardless of the image size
        self.loss_weight = (1.0 / 16384.0) if loss_weight is None else float(loss_weight)
        self._w_ptrs = _lib.ptr_array(self.weights)
        self._b_ptrs = _lib.ptr_array(self.biases)
        self._dw_ptrs = _lib.ptr_array([w.grad for w in self.weights])
        self._db_ptrs = _lib.ptr_array([b.grad for b in self.biases])
        self.use_graph = use_graph
        self.graph = None
        self.steps = 0            # completed optimizer steps (host count; the device counter is in opt.state)
        # kernels of this library launched per step (see csrc/api.cu):
        #   hidden_fwd, hidden_dgrad per hidden layer; prep_weights, first_fwd, mse_grad, last_bwd, wgrad,
        #   adam_tick, adam; plus (generic path) colsum per hidden layer below the top, last_fwd, first_bwd
        #   fused path (bf16, <= 4 hidden layers): prep_weights, mlp_fused_fwd, mse_grad, mlp_fused_bwd,
        #   wgrad, adam_tick, adam; plus last_fwd / last_bwd when the outermost linear is not fused
        nh = desc.n_hidden
        fast = self.precision == "bf16"
        clip = 1 if max_grad_norm > 0 else 0
        fused = fast and nh <= 4 and os.environ.get("SIREN_FUSED", "1")[:1] != "0"
        if fused:
            fuse_top = d_out <= 2
            self.kernels_per_step = 7 + clip + (0 if d_out <= 2 else 1) + (0 if fuse_top else 1)
            if d_in > 4:
                self.kernels_per_step += 2        # prep_first, first_bwd
        else:
            self.kernels_per_step = 2 * nh + 7 + clip
            if not fast:
                self.kernels_per_step += (nh - 1) + 2
            else:
                self.kernels_per_step += (0 if d_out <= 2 else 1) + (0 if d_in <= 3 else 1)

    def _make_comm(self):
        """NCCL communicator of the C ABI: rank 0 draws the id, torch.distributed hands it round."""
        import ctypes
        dist = torch.distributed
        rank = dist.get_rank(self.pg)
        idbuf = (ctypes.c_char * 128)()
        if rank == 0:
            _lib.check(self.lib.siren_b200_comm_unique_id(ctypes.cast(idbuf, ctypes.c_void_p)), "comm_unique_id")
        t = torch.frombuffer(bytearray(bytes(idbuf)), dtype=torch.uint8).clone().to(self.device)
        dist.broadcast(t, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0, group=self.pg)
        raw = bytes(t.cpu().numpy().tobytes())
        comm = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.siren_b200_comm_init(rank, self.world, ctypes.c_char_p(raw), ctypes.byref(comm))
        _lib.check(rc, "comm_init")
        return comm

    # one step, enqueued on the current stream (batch buffers other than self.coords / self.gt: the pipelined entry)
    def _enqueue(self, coords=None, gt=None, loss_out=None):
        lib, d = self.lib, self.desc
        coords = self.coords if coords is None else coords
        gt = self.gt if gt is None else gt
        stream = torch.cuda.current_stream(self.device).cuda_stream
        P = _lib.dptr
        _lib.check(lib.siren_b200_forward(d, P(coords), self._w_ptrs, self._b_ptrs, P(self.y), None, None,
                                          P(self.ws), stream), "forward")
        self.loss.zero_()
        _lib.check(lib.siren_b200_mse_grad(P(self.y), P(gt), P(self.gy), self.y.numel(), self.loss_weight,
                                           P(self.loss), stream), "mse_grad")
        # every gradient is a view of one flat buffer: clear it with ONE fill and let the kernels accumulate
        # (accumulate = 0 would clear the ten tensors one by one: ten more nodes in the step's graph)
        self.grad.zero_()
        _lib.check(lib.siren_b200_backward(d, P(coords), self._w_ptrs, self._b_ptrs, P(self.ws), P(self.gy),
                                           None, None, self._dw_ptrs, self._db_ptrs, None, 1, stream), "backward")
        if self.world > 1:
            if self.comm is not None:
                _lib.check(lib.siren_b200_allreduce(self.comm, P(self.grad), self.grad.numel(), stream), "allreduce")
            else:
                torch.distributed.all_reduce(self.grad, group=self.pg)
        self.opt.step()
        if loss_out is not None:
            # pinned host word, written by a kernel (a copy-engine node here costs ~35 us per step of hand-over)
            _lib.check(lib.siren_b200_publish(P(self.loss), loss_out.data_ptr(), 1, stream), "publish")

    def _warm_up(self):
        """Two steps outside capture (function attributes, NCCL) with the optimizer state restored afterwards."""
        if getattr(self, "_warm", False):
            return
        state = [t.clone() for t in (self.flat, self.opt.m, self.opt.v, self.opt.state)]
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(2):
                self._enqueue()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        for dst, src in zip((self.flat, self.opt.m, self.opt.v, self.opt.state), state):
            dst.copy_(src)
        self._warm = True

    def _capture(self, *bufs):
        self._warm_up()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._enqueue(*bufs)
        return graph              # capture does not execute: state is untouched

    def step(self):
        """Run one training step on the data currently in ``self.coords`` / ``self.gt``."""
        self.steps += 1
        with torch.cuda.device(self.device):
            if not self.use_graph:
                self._enqueue()
                return
            if self.graph is None:
                self.graph = self._capture()
            self.graph.replay()

    def step_from_host(self, coords_host, gt_host):
        """Public end-to-end step: pinned host batch in, loss value out."""
        self.coords.copy_(coords_host.view_as(self.coords), non_blocking=True)
        self.gt.copy_(gt_host.view_as(self.gt), non_blocking=True)
        self.step()
        return float(self.loss.item())

    def submit_from_host(self, coords_host, gt_host):
        """Pipelined end-to-end step: enqueue (pinned host batch -> device -> step -> loss to pinned host memory)
        and return a handle at once; ``handle.result()`` blocks for that step's loss.  Reading a step's loss
        after submitting the next step lets the upload of step k+1 run under the kernels of step k.

        Four batch slots (device coords + gt, one pinned loss word, one captured graph each) are used in turn: the
        upload goes straight into the slot its graph reads, on a copy stream, and the loss leaves through a kernel
        at the end of the graph that stores it to pinned host memory (siren_b200_publish) -- the compute stream
        carries nothing but graph launches, and no copy-engine hand-over.

        (The training loop of the reference logs the loss every step, training.py:83-104; a loop built on this
        call logs it one step late.)"""
        dev = self.device
        if getattr(self, "_slots", None) is None:
            self._copy_stream = torch.cuda.Stream(dev)
            self._slots = [dict(coords=torch.empty_like(self.coords), gt=torch.empty_like(self.gt),
                                loss=torch.zeros(1, dtype=torch.float32).pin_memory(), staged=torch.cuda.Event(),
                                done=torch.cuda.Event(), graph=None) for _ in range(4)]
            self._submitted = 0
            if self.use_graph:                    # all four graphs now: no capture (it synchronises) in later steps
                with torch.cuda.device(dev):
                    for sl in self._slots:
                        sl["graph"] = self._capture(sl["coords"], sl["gt"], sl["loss"])
        s = self._slots[self._submitted % len(self._slots)]
        self._submitted += 1
        self.steps += 1
        cur = torch.cuda.current_stream(dev)
        cs = self._copy_stream
        cs.wait_event(s["done"])                  # the step that last read this slot has finished (no-op the first time)
        with torch.cuda.stream(cs):
            s["coords"].copy_(coords_host.view_as(self.coords), non_blocking=True)
            s["gt"].copy_(gt_host.view_as(self.gt), non_blocking=True)
            s["staged"].record(cs)
        with torch.cuda.device(dev):
            cur.wait_event(s["staged"])
            if self.use_graph:
                s["graph"].replay()
            else:
                self._enqueue(s["coords"], s["gt"], s["loss"])
        s["done"].record(cur)
        return _LossHandle(s["loss"], s["done"])


class _LossHandle:
    """Result of ``SirenTrainer.submit_from_host``: the step's loss once its device-to-host copy has landed.
    Valid until three further steps have been submitted (the pinned slots are a ring of four)."""

    def __init__(self, slot, event):
        self._slot, self._event = slot, event

    def result(self):
        self._event.synchronize()
        return float(self._slot[0])
