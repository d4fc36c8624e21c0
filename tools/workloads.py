"""The five BASELINE.json configurations as synthetic workloads, shared by bench.py and
tools/bench_configs.py: inputs, the reference's loss for each configuration, and one timed
``model -> loss -> backward -> (clip) -> Adam`` loop through the PUBLIC module API.

  cfg1  SIREN 3x256 image fit, 256x256 grid = 65,536 coords, image_mse
  cfg2  same at 512x512 = 262,144 coords
  cfg3  SDF, d = 3, 250,000 point-cloud coords, loss_functions.sdf (first-order derivatives), clip_grad
  cfg4  Poisson from the Laplacian, 512x512 grid, loss_functions.laplace_mse (second order)
  cfg5  hypernetwork SIREN: B tasks x 65,536 coords, Fourier-feature input d = 16, o = 2, per-task weights,
        image_mse (B = 8 = one GPU's share of the 64-task batch sharded over 8 GPUs)

The model can be this package's (backend 'auto' = native kernels, 'composed' = the reference's ops in eager
PyTorch) or, when ``baseline/_ref`` holds the staged unmodified reference (``__graft_entry__.build()`` copies it
from /root/reference in the dev container; git-ignored), the reference's own classes and loss functions.
"""
import contextlib
import io
import math
import os
import sys
import types

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_STAGE = os.path.join(ROOT, "baseline", "_ref")

U = lambda d, o: 2 * (d * 256 + 3 * 256 * 256 + 256 * o)      # noqa: E731  forward GEMM FLOP per coordinate
# algorithmic FLOP per coordinate: value path 3U; first-order loss 6U; Laplacian loss 12U
FLOP = {1: 3 * U(2, 1), 2: 3 * U(2, 1), 3: 6 * U(3, 1), 4: 12 * U(2, 1), 5: 3 * U(16, 2)}
NAMES = {1: "cfg1 image 256^2", 2: "cfg2 image 512^2", 3: "cfg3 SDF first-order", 4: "cfg4 Poisson second-order",
         5: "cfg5 per-task weights"}


# ------------------------------------------------------------------------------------------------------------
# the staged reference (optional)
# ------------------------------------------------------------------------------------------------------------
_ref = None


def reference_modules():
    """(modules, loss_functions, diff_operators) of the UNMODIFIED reference staged in baseline/_ref, or None.
    Import recipe of SURVEY.md 8c: bypass torchmeta/__init__ (it pulls h5py), stub the plotting / IO imports."""
    global _ref
    if _ref is not None:
        return _ref or None
    if not os.path.isfile(os.path.join(REF_STAGE, "modules.py")):
        _ref = False
        return None
    try:
        sys.path.insert(0, REF_STAGE)
        pkg = types.ModuleType("torchmeta")
        pkg.__path__ = [os.path.join(REF_STAGE, "torchmeta")]
        sys.modules["torchmeta"] = pkg
        for name in ["h5py", "matplotlib", "matplotlib.colors", "matplotlib.pyplot", "skimage", "skimage.filters",
                     "skimage.measure", "skvideo", "skvideo.io", "cmapy"]:
            sys.modules.setdefault(name, types.ModuleType(name))
        import modules as ref_modules
        import loss_functions as ref_losses
        import diff_operators as ref_diff
        _ref = (ref_modules, ref_losses, ref_diff)
    except Exception as e:  # pragma: no cover
        print("[workloads] staged reference not importable: %r" % (e,), file=sys.stderr)
        _ref = False
        return None
    return _ref


def reference_model(d, o):
    ref = reference_modules()
    with contextlib.redirect_stdout(io.StringIO()):          # SingleBVPNet.__init__ prints the module (modules.py:144)
        return ref[0].SingleBVPNet(out_features=o, type="sine", in_features=d, mode="mlp", hidden_features=256,
                                   num_hidden_layers=3)


# ------------------------------------------------------------------------------------------------------------
# losses (restated; used when the reference files are not staged, and pinned against them in tests)
# ------------------------------------------------------------------------------------------------------------
def image_mse(model_output, gt):
    """loss_functions.py:66-96 with mask None, high_freq False: sum of squares / 16384."""
    return ((model_output["model_out"] - gt["img"]) ** 2).sum() / 16384.0


def sdf_loss(model_output, gt, gradient_fn):
    """loss_functions.py:460-484."""
    gt_sdf, gt_normals = gt["sdf"], gt["normals"]
    coords, pred = model_output["model_in"], model_output["model_out"]
    gradient = gradient_fn(pred, coords)
    sdf_c = torch.where(gt_sdf != -1, pred, torch.zeros_like(pred))
    inter = torch.where(gt_sdf != -1, torch.zeros_like(pred), torch.exp(-1e2 * torch.abs(pred)))
    normal = torch.where(gt_sdf != -1, 1 - F.cosine_similarity(gradient, gt_normals, dim=-1)[..., None],
                         torch.zeros_like(gradient[..., :1]))
    gc = torch.abs(gradient.norm(dim=-1) - 1)
    return torch.abs(sdf_c).mean() * 3e3 + inter.mean() * 1e2 + normal.mean() * 1e2 + gc.mean() * 5e1


def laplace_mse(model_output, gt, laplace_fn):
    """loss_functions.py:350-355."""
    lap = laplace_fn(model_output["model_out"], model_output["model_in"])
    return torch.mean((lap - gt["laplace"]) ** 2)


# ------------------------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------------------------
def mgrid(side, dev="cpu"):
    """dataio.get_mgrid (dataio.py:28-48) for dim 2: [side^2, 2] in [-1, 1]."""
    lin = torch.linspace(-1, 1, side, device=dev)
    return torch.stack(torch.meshgrid(lin, lin, indexing="ij"), dim=-1).reshape(-1, 2)


def synthetic_image(grid, seed=1234):
    """Smooth stand-in for the cameraman image (no skimage / network): eight random sinusoids, in [-1, 1]."""
    g = torch.Generator().manual_seed(seed)
    img = torch.zeros(grid.shape[0], 1)
    for _ in range(8):
        f = torch.randn(2, generator=g) * 6.0
        ph = torch.rand(1, generator=g) * 6.28
        img += torch.sin(grid.cpu() @ f.view(2, 1) + ph)
    return (img / 8.0).to(grid.device)


def make_inputs(cfg, dev, tasks=8):
    """(coords [B, N, d], gt dict, d, o, coord_derivs, clip) for configuration ``cfg``."""
    g = torch.Generator().manual_seed(cfg)
    if cfg in (1, 2, 4):
        side = 256 if cfg == 1 else 512
        grid = mgrid(side)
        img = synthetic_image(grid)
        x = grid.unsqueeze(0).to(dev)
        if cfg == 4:
            import numpy as np
            import scipy.ndimage
            lap = scipy.ndimage.laplace(1e4 * img.reshape(side, side).numpy())           # dataio.py:783-785
            gt = {"laplace": torch.from_numpy(np.ascontiguousarray(lap)).float().reshape(1, -1, 1).to(dev)}
            return x, gt, 2, 1, 2, False
        return x, {"img": img.unsqueeze(0).to(dev)}, 2, 1, 0, False
    if cfg == 3:
        n = 250000
        p = torch.randn((1, n // 2, 3), generator=g)
        on = 0.5 * p / p.norm(dim=-1, keepdim=True)                                       # dataio.py:431-453
        off = torch.rand((1, n - n // 2, 3), generator=g) * 2 - 1
        x = torch.cat([on, off], dim=1).to(dev)
        sdf = torch.cat([torch.zeros(1, n // 2, 1), -torch.ones(1, n - n // 2, 1)], dim=1).to(dev)
        normals = torch.cat([on / 0.5, -torch.ones(1, n - n // 2, 3)], dim=1).to(dev)
        return x, {"sdf": sdf, "normals": normals}, 3, 1, 1, True
    # cfg5: Gaussian Fourier features of the 256^2 grid (features.py:31-41: x @ B, 2 pi, sin || cos), B ~ N(0, 21^2)
    grid = mgrid(256)
    Bm = torch.randn((2, 8), generator=torch.Generator().manual_seed(0)) * 21.0
    proj = 2 * math.pi * (grid @ Bm)
    feat = torch.cat([torch.sin(proj), torch.cos(proj)], dim=-1)                          # [65536, 16]
    x = feat.unsqueeze(0).expand(tasks, -1, -1).contiguous().to(dev)
    gt = {"img": (torch.rand((tasks, 65536, 2), generator=g) * 2 - 1).to(dev)}
    return x, gt, 16, 2, 0, False


def per_task_params(model, tasks, dev):
    """Hypernetwork-style per-task weights: the hypo-net's init plus N(0, 1e-2)."""
    from collections import OrderedDict
    g = torch.Generator().manual_seed(5)
    params = OrderedDict()
    for name, p in model.named_parameters():
        noise = 1e-2 * torch.randn((tasks,) + tuple(p.shape), generator=g)
        params[name] = (p.detach().cpu().unsqueeze(0) + noise * p.detach().cpu().abs().mean()).to(dev).requires_grad_(True)
    return params


# ------------------------------------------------------------------------------------------------------------
# one configuration through the public API
# ------------------------------------------------------------------------------------------------------------
def run_config(cfg, impl, precision="bf16", steps=10, warmup=3, tasks=8, dev=None, want_profile=False, lazy_fourier=False):
    """Time ``steps`` training steps of configuration ``cfg``.

    impl: 'native'    this package's modules on the native kernels (precision 'bf16' | 'fp32')
          'eager'     the reference's ops in eager PyTorch on the same device (fp32, TF32 off): the staged
                      reference's own classes + loss functions when present, else this package's composed path
    """
    import siren_mri_b200
    from siren_mri_b200 import diff_operators, modules
    dev = dev or torch.device("cuda")
    is_cuda = torch.device(dev).type == "cuda"
    torch.manual_seed(cfg)
    x, gt, d, o, derivs, clip = make_inputs(cfg, dev, tasks)
    transform = None
    if cfg == 5 and impl == "native":
        # the training loops apply the Fourier-feature transform to the raw grid EVERY step (training.py:61-64,
        # training_ddp.py:66-69): so does this step -- materialised (the reference flow: x @ B, 2 pi, sin | cos, cat) or
        # lazy (the raw grid goes to the model, the kernels build the features)
        from siren_mri_b200 import features
        transform = features.GaussianFourierFeatureTransform(2, 8, 21, lazy=lazy_fourier)
        transform.set_B((torch.randn((2, 8), generator=torch.Generator().manual_seed(0)) * 21.0).to(dev))
        x = mgrid(256).unsqueeze(0).expand(tasks, -1, -1).contiguous().to(dev)
    ref = reference_modules() if impl == "eager" else None
    if impl == "eager":
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
    if ref is not None:
        model = reference_model(d, o).to(dev)
        grad_fn, lap_fn, used = ref[2].gradient, ref[2].laplace, "reference classes (baseline/_ref)"
    else:
        model = modules.SingleBVPNet(in_features=d, out_features=o, precision=precision if impl == "native" else "fp32",
                                     coord_derivs=derivs if impl == "native" else 0,
                                     backend="auto" if impl == "native" else "composed").to(dev)
        grad_fn, lap_fn = diff_operators.gradient, diff_operators.laplace
        used = "siren_mri_b200 native kernels" if impl == "native" else "siren_mri_b200 composed ops (reference not staged)"
    params = per_task_params(model, tasks, dev) if cfg == 5 else None
    if cfg == 3:
        loss_fn = lambda out: sdf_loss(out, gt, grad_fn)               # noqa: E731
        if ref is not None:
            loss_fn = lambda out: sum(v.mean() for v in ref[1].sdf(out, gt).values())     # noqa: E731
    elif cfg == 4:
        loss_fn = lambda out: laplace_mse(out, gt, lap_fn)             # noqa: E731
        if ref is not None:
            loss_fn = lambda out: ref[1].laplace_mse(out, gt)["laplace_loss"]             # noqa: E731
    else:
        loss_fn = lambda out: image_mse(out, gt)                       # noqa: E731
        if ref is not None:
            loss_fn = lambda out: ref[1].image_mse(None, out, gt, high_freq=False)["img_loss"]   # noqa: E731
    leaves = list(params.values()) if params is not None else list(model.parameters())
    opt = torch.optim.Adam(leaves, lr=1e-4)

    def step():
        xin = transform(x) if transform is not None else x
        out = model({"coords": xin}, params=params) if params is not None else model({"coords": xin})
        loss = loss_fn(out)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if clip:
            torch.nn.utils.clip_grad_norm_(leaves, max_norm=1.0)
        opt.step()
        return loss

    import time
    for _ in range(warmup):
        step()
    if is_cuda:
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        evs[0].record()
    t0 = time.perf_counter()
    for i in range(steps):
        loss = step()
        if is_cuda:
            evs[i + 1].record()
    if is_cuda:
        torch.cuda.synchronize()
        per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(steps))
        ms, ms_mean = per[len(per) // 2], sum(per) / len(per)      # median: one allocator stall does not define the step
    else:
        ms = ms_mean = (time.perf_counter() - t0) * 1e3 / steps
    n_coords = x.shape[0] * x.shape[1]
    res = {"config": NAMES[cfg] + (" (Fourier prologue in the kernels)" if lazy_fourier else ""), "impl": impl,
           "precision": precision if impl == "native" else "fp32 (TF32 off)",
           "model": used, "coords_per_step": n_coords, "ms_per_step": ms, "ms_per_step_mean": ms_mean,
           "coords_per_sec": n_coords / (ms * 1e-3),
           "flop_per_coord": FLOP[cfg], "loss": float(loss.detach()),
           "path": "public modules + reference-style loss + torch autograd + torch.optim.Adam"}
    if is_cuda:
        res["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
    if want_profile and impl == "native":
        import ctypes
        from siren_mri_b200 import _lib
        lib = _lib.load()
        lib.siren_b200_profile_begin()
        for _ in range(3):
            step()
        buf = ctypes.create_string_buffer(1 << 16)
        lib.siren_b200_profile_end(buf, len(buf))
        res["kernels_us_per_step"] = {ln.split()[0]: round(1e3 * float(ln.split()[2]) / 3, 1)
                                      for ln in buf.value.decode().strip().splitlines()}
    del model, opt, leaves, params, x, gt
    siren_mri_b200.functional.clear_workspace_cache()
    if is_cuda:
        torch.cuda.empty_cache()
    return res


def run_trainer_config(cfg, precision="bf16", steps=20, warmup=5, dev=None):
    """cfg1..cfg4 through SirenTrainer: the whole step (forward, loss tail, backward, clip, Adam) as ONE CUDA graph
    of this library's kernels -- what training.train_fast runs."""
    import siren_mri_b200
    from siren_mri_b200 import modules
    from siren_mri_b200.trainer import SirenTrainer
    dev = dev or torch.device("cuda")
    torch.manual_seed(cfg)
    x, gt, d, o, derivs, clip = make_inputs(cfg, dev)
    loss = {1: "image_mse", 2: "image_mse", 3: "sdf", 4: "laplace_mse"}[cfg]
    model = modules.SingleBVPNet(in_features=d, out_features=o, precision=precision).to(dev)
    tr = SirenTrainer(model, x.shape[1], lr=1e-4, max_grad_norm=1.0 if clip else 0.0, precision=precision, loss=loss,
                      distributed=False)
    tr.coords.copy_(x)
    for dst, k in zip(tr.gts, tr.gt_keys):
        dst.copy_(gt[k])
    for _ in range(warmup):
        tr.step()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for i in range(steps):
        tr.step()
        evs[i + 1].record()
    torch.cuda.synchronize()
    per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(steps))
    ms, ms_mean = per[len(per) // 2], sum(per) / len(per)
    n = x.shape[1]
    res = {"config": NAMES[cfg], "impl": "native-trainer", "precision": precision, "model": "siren_mri_b200 SirenTrainer (one CUDA graph)",
           "coords_per_step": n, "ms_per_step": ms, "ms_per_step_mean": ms_mean, "coords_per_sec": n / (ms * 1e-3),
           "flop_per_coord": FLOP[cfg],
           "loss": float(tr.loss.item()), "path": "SirenTrainer: forward + %s tail + backward + %sAdam, captured" % (loss, "clip + " if clip else ""),
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    del tr, model
    siren_mri_b200.functional.clear_workspace_cache()
    torch.cuda.empty_cache()
    return res
