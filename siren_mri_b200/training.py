"""``train_fast``: the image-fitting training loop of the reference on the graph-captured step.

Sibling of ``training.train`` (reference training.py:19-146) for the case the fast step covers: a sine
``SingleBVPNet`` / ``FCBlock`` fitted with ``loss_functions.image_mse`` (mask None / high_freq False:
sum of squared differences / 16384, loss_functions.py:66-96), Adam, optional ``clip_grad``.  Same
arguments, same files:

    model_dir/summaries/                            (tensorboard, when a writer can be made)
    model_dir/checkpoints/model_epoch_%04d.pth      every ``epochs_til_checkpoint`` epochs   (training.py:47-52)
    model_dir/checkpoints/train_losses_epoch_%04d.txt
    model_dir/checkpoints/model_current.pth         every ``steps_til_summary`` steps       (training.py:83-86)
    model_dir/checkpoints/model_final.pth, train_losses_final.txt                           (training.py:140-143)

with ``model.state_dict()`` in the reference's own key layout, so checkpoints load into either implementation.

What differs, on purpose: the step is one CUDA graph fed from pinned host batches (SirenTrainer.submit_from_host),
so the per-step ``train_loss.item()`` / ``writer.add_scalar`` synchronisations of training.py:77-81 are gone.  Each
step's loss reaches the host one step late; it is logged under its own step number.  The reference prompts on an
existing ``model_dir`` (training.py:25-31); this function takes ``overwrite`` instead.
"""
import os
import shutil
import time

import numpy as np
import torch

IMAGE_MSE_WEIGHT = 1.0 / 16384.0          # loss_functions.py:88-96


def _make_writer(summaries_dir):
    try:
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(summaries_dir)
    except Exception:                      # tensorboard is optional here; the loss files are always written
        return None


def _default_trainer(model, n_coords, lr, loss_weight, max_grad_norm, loss="image_mse"):
    from .trainer import SirenTrainer
    return SirenTrainer(model, n_coords, lr=lr, loss_weight=loss_weight, max_grad_norm=max_grad_norm, loss=loss)


def _batch(t):
    """fp32, contiguous; pinned batches (DataLoader(pin_memory=True)) upload asynchronously, pageable ones through
    the driver's staging buffer, device tensors by a device-to-device copy."""
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def train_fast(model, train_dataloader, epochs, lr, steps_til_summary, epochs_til_checkpoint, model_dir,
               loss_weight=IMAGE_MSE_WEIGHT, summary_fn=None, clip_grad=False, overwrite=False, loss_name="img_loss",
               trainer_factory=None, progress=None, accumulation_steps=1, loss="image_mse"):
    """Fit ``model`` to the ``(model_input, gt)`` batches of ``train_dataloader``.

    ``model_input['coords']`` is ``[1, N, d]`` and ``gt['img']`` is ``[1, N, o]`` with the same N every step (the
    reference's image datasets yield the whole image as one batch, dataio.py:754-770).  ``clip_grad`` is False,
    True (max norm 1) or the max norm, as at training.py:93-97.  ``summary_fn(model, model_input, gt, model_output,
    writer, total_steps)`` is called every ``steps_til_summary`` steps like the reference's (training.py:83-86),
    with tensors on the model's device and a model output evaluated with gradients enabled, as in the reference, so
    that summary functions which differentiate it (utils.write_image_summary calls diff_operators.gradient / laplace on
    ``model_output['model_out']`` w.r.t. ``['model_in']``) work unchanged.  ``accumulation_steps`` is the gradient
    accumulation of training.py:90, 99-103: every batch adds the gradient of ``loss / accumulation_steps``, the
    optimizer steps after every ``accumulation_steps``-th batch and after the last batch of an epoch; the logged loss
    is the undivided one.  ``loss`` selects the captured loss tail: ``'image_mse'`` (``gt['img']``; ``loss_weight``
    applies), ``'sdf'`` (``gt['sdf']``, ``gt['normals']``; loss_functions.py:460-484, the reference runs it with
    ``clip_grad=True``, train_sdf.py:57-60) or ``'laplace_mse'`` (``gt['laplace']``; loss_functions.py:350-355).
    Returns the list of per-step training losses (what ``train_losses_final.txt`` holds)."""
    if os.path.exists(model_dir):
        if not overwrite:
            raise FileExistsError("model directory %s exists (pass overwrite=True to replace it)" % model_dir)
        shutil.rmtree(model_dir)
    os.makedirs(model_dir)
    summaries_dir = os.path.join(model_dir, "summaries")
    checkpoints_dir = os.path.join(model_dir, "checkpoints")
    os.makedirs(summaries_dir, exist_ok=True)
    os.makedirs(checkpoints_dir, exist_ok=True)
    writer = _make_writer(summaries_dir)

    max_norm = 0.0
    if clip_grad:
        max_norm = 1.0 if isinstance(clip_grad, bool) else float(clip_grad)

    trainer = None
    train_losses = []
    pending = []                           # (step number, loss handle) of submitted steps not read yet

    def drain(keep):
        while len(pending) > keep:
            k, h = pending.pop(0)
            v = h.result()
            assert k == len(train_losses)
            train_losses.append(v)
            if writer is not None:
                writer.add_scalar(loss_name, v, k)
                writer.add_scalar("total_train_loss", v, k)

    total_steps = 0
    t_last = time.time()
    for epoch in range(epochs):
        if not epoch % epochs_til_checkpoint and epoch:
            drain(0)
            torch.save(model.state_dict(), os.path.join(checkpoints_dir, "model_epoch_%04d.pth" % epoch))
            np.savetxt(os.path.join(checkpoints_dir, "train_losses_epoch_%04d.txt" % epoch), np.array(train_losses))
        n_batches = len(train_dataloader) if hasattr(train_dataloader, "__len__") else None
        for step, (model_input, gt) in enumerate(train_dataloader):
            coords = model_input["coords"]
            keys = {"image_mse": ("img",), "sdf": ("sdf", "normals"), "laplace_mse": ("laplace",)}[loss]
            img = _batch(gt[keys[0]]) if len(keys) == 1 else [_batch(gt[k]) for k in keys]
            if trainer is None:
                make = trainer_factory or _default_trainer
                if loss == "image_mse":
                    trainer = make(model, coords.shape[-2], lr, loss_weight, max_norm)
                else:
                    trainer = make(model, coords.shape[-2], lr, None, max_norm, loss)
            summary = not total_steps % steps_til_summary
            if summary:
                # as in the reference, the checkpoint and the summary see the weights BEFORE this step's update
                drain(0)
                torch.save(model.state_dict(), os.path.join(checkpoints_dir, "model_current.pth"))
                if summary_fn is not None:
                    dev = next(model.parameters()).device
                    mi = {k: v.to(dev) for k, v in model_input.items()}
                    g = {k: v.to(dev) for k, v in gt.items()}
                    out = model(mi)        # grad enabled (training.py:66, 83-86)
                    summary_fn(model, mi, g, out, writer, total_steps)
                    del out
            if accumulation_steps > 1:
                update = (step + 1) % accumulation_steps == 0 or (step + 1 == n_batches)
                handle = trainer.submit_from_host(_batch(coords), img, update=update,
                                                  accumulation_steps=accumulation_steps)
            else:
                handle = trainer.submit_from_host(_batch(coords), img)
            pending.append((total_steps, handle))
            drain(0 if summary else 1)     # normally the previous step's loss, while this step runs
            if summary:
                msg = "Epoch %d, Total loss %0.6f, iteration time %0.6f" % (epoch, train_losses[-1], time.time() - t_last)
                (progress or print)(msg)
            t_last = time.time()
            total_steps += 1
    drain(0)
    torch.save(model.state_dict(), os.path.join(checkpoints_dir, "model_final.pth"))
    np.savetxt(os.path.join(checkpoints_dir, "train_losses_final.txt"), np.array(train_losses))
    if writer is not None:
        writer.close()
    return train_losses


class GraphedStep:
    """A training step with fixed shapes captured into ONE CUDA graph and replayed.

    For the steps the hand-written trainer does not cover -- above all the neural-process step of
    training.py:61-103 / training_ddp.py:66-118 (encoder -> hypernetwork -> per-sample hypo-network -> data consistency
    -> losses -> backward), which launched from Python is host-bound: ~100 small kernels around three large ones (2.9 ms per
    step against 2.0 ms replayed, tools/probe_mri_step.py).  Every entry point of the C ABI is capturable (no
    allocation, no synchronisation; workspaces come from the stream-ordered free list, which the capture draws from the
    graph's private pool).

    ``step_fn(*static_inputs)`` runs forward, losses and ``backward()`` and returns the tensor(s) to read back (e.g. the
    loss).  It must not synchronise (no ``.item()``).  Gradients: either drop them INSIDE the step (``p.grad = None``
    before ``backward()``: every replay then re-creates them at the captured addresses, read them after the call) or
    clear them in place inside the step (``grad.zero_()``, ``PeerGradientReducer.zero_grad()``) -- not from outside with
    ``set_to_none`` (the replay writes into the tensors it captured).  The optimizer step may be part of
    ``step_fn`` when the optimizer is capturable (``torch.optim.Adam(capturable=True)``, ``optim.FusedAdam``).

        gs = GraphedStep(step_fn, example_inputs)        # warm-up on a side stream, then capture
        loss = gs(coords, img_sparse, dc_mask, gt)       # copies the batch into the static inputs, replays
    """

    def __init__(self, step_fn, example_inputs, warmup=3):
        self.static_inputs = [t.clone() if torch.is_tensor(t) else t for t in example_inputs]
        dev = next((t.device for t in self.static_inputs if torch.is_tensor(t)), None)
        if dev is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step_fn(*self.static_inputs)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_outputs = step_fn(*self.static_inputs)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_inputs, inputs):
            if torch.is_tensor(dst) and src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_outputs
