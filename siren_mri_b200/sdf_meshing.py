"""Bulk forward-only evaluation of an SDF network on a dense grid: the sampling half of the reference's
``sdf_meshing.create_mesh`` (sdf_meshing.py:13-61, from DeepSDF).

The reference builds all ``N^3`` sample points on the host (``N^3 x 4`` floats: 65 GB at N = 1600), ships ``max_batch``
of them to the GPU per iteration and copies every chunk of values back.  Here the coordinates of a chunk are generated
on the device from the flat index (same enumeration: x slowest, z fastest, ``voxel_origin = -1``,
``voxel_size = 2 / (N - 1)``; the integer divisions are floor divisions, as under the PyTorch 1.5 the reference pins),
the decoder runs under ``torch.no_grad()`` -- a native ``SingleBVPNet`` / ``FCBlock`` then takes the stash-free
inference kernel (``siren_b200_forward_infer``) -- and the values stay on the device until the volume is complete.

``create_mesh`` keeps the reference's signature; the surface extraction itself needs ``skimage`` and ``plyfile``
exactly as in the reference (``convert_sdf_samples_to_ply``, sdf_meshing.py:64-128) and is delegated to a callable.
"""
import torch


def grid_coords(head, tail, N, device):
    """Sample points ``head .. tail - 1`` of the ``N^3`` grid as ``[tail - head, 3]`` fp32 (sdf_meshing.py:25-38)."""
    idx = torch.arange(head, tail, device=device, dtype=torch.int64)
    voxel_size = 2.0 / (N - 1)
    z = (idx % N).to(torch.float32)
    y = (torch.div(idx, N, rounding_mode="floor") % N).to(torch.float32)
    x = (torch.div(idx, N * N, rounding_mode="floor") % N).to(torch.float32)
    return torch.stack([x * voxel_size - 1.0, y * voxel_size - 1.0, z * voxel_size - 1.0], dim=-1)


def sample_sdf_grid(decoder, N=256, max_batch=64 ** 3, device=None):
    """``[N, N, N]`` fp32 tensor of decoder values on the grid (on ``device``), evaluated ``max_batch`` points at a
    time.  ``decoder`` maps ``[M, 3]`` coordinates to ``[M, 1]`` (or ``[M]``) values, like the reference's SDFDecoder
    (test_sdf.py:27-45)."""
    if device is None:
        try:
            device = next(decoder.parameters()).device
        except (StopIteration, AttributeError):
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    if hasattr(decoder, "eval"):
        decoder.eval()
    num = N ** 3
    out = torch.empty(num, dtype=torch.float32, device=device)
    with torch.no_grad():
        for head in range(0, num, max_batch):
            tail = min(head + max_batch, num)
            vals = decoder(grid_coords(head, tail, N, device))
            out[head:tail] = vals.reshape(-1).to(torch.float32)
    return out.reshape(N, N, N)


def create_mesh(decoder, filename, N=256, max_batch=64 ** 3, offset=None, scale=None, convert=None):
    """sdf_meshing.py:13-61.  ``convert(sdf_cpu, voxel_origin, voxel_size, ply_filename, offset, scale)`` defaults to
    the reference's ``convert_sdf_samples_to_ply`` when ``sdf_meshing`` (and with it skimage / plyfile) is importable."""
    sdf = sample_sdf_grid(decoder, N=N, max_batch=max_batch)
    if convert is None:
        try:
            import sdf_meshing as _ref      # the reference's module, if its checkout is on sys.path
            convert = _ref.convert_sdf_samples_to_ply
        except Exception as e:      # noqa: BLE001
            raise ImportError("create_mesh needs a surface extractor: pass convert=..., or put the reference's "
                              "sdf_meshing.py (skimage, plyfile) on sys.path") from e
    convert(sdf.cpu(), [-1, -1, -1], 2.0 / (N - 1), filename + ".ply", offset, scale)
    return sdf
