"""Host-side mirror of the reference's features.py (GaussianFourierFeatureTransform, AsinhTransform).

``GaussianFourierFeatureTransform`` keeps the reference's constructor, ``forward`` and B accessors
(features.py:21-53).  With ``lazy=True`` its forward does NOT write the ``[B, N, 2 F]`` feature tensor: it hands the
raw coordinates on, tagged with the projection matrix, and ``modules.SingleBVPNet`` / ``FCBlock`` give both to the
kernels, whose first-layer operand producer builds the features of a row on chip (siren_b200_forward_ff).  Anything
that is not the native path (CPU tensors, other widths, F outside 3..8) materialises the features from the tag with
the reference's own ops, so the result is the same either way.  ``lazy`` needs this package's ``SingleBVPNet`` or the
reference's own class after ``integration.patch_reference`` (which carries the tag through the clone of the
coordinates): it is off by default.
"""
import torch

from . import functional


class GaussianFourierFeatureTransform(torch.nn.Module):
    """features.py:6-53.  ``[batches, n_coords, num_input_channels]`` -> ``[batches, n_coords, 2 * mapping_size]``."""

    def __init__(self, num_input_channels, mapping_size_spatial=256, scale=10, loaded_B=None, device="cuda:0", lazy=False):
        super().__init__()
        self._num_input_channels = num_input_channels
        self._mapping_size = mapping_size_spatial
        self._spatial_dims = [0, 1]
        self._B_spatial = torch.randn((num_input_channels, mapping_size_spatial)) * scale
        self.lazy = bool(lazy)
        self._B_dev = {}

    def _B_on(self, device):
        key = (str(device), self._B_spatial.data_ptr(), self._B_spatial._version)
        hit = self._B_dev.get("B")
        if hit is None or hit[0] != key:
            self._B_dev["B"] = (key, self._B_spatial.detach().to(device=device, dtype=torch.float32).contiguous())
        return self._B_dev["B"][1]

    def forward(self, x):
        if self.lazy and torch.is_tensor(x) and x.is_cuda and x.dim() == 3:
            out = x.detach().view_as(x)
            out._siren_fourier = self._B_on(x.device)
            return out
        return functional.fourier_features(x, self._B_spatial)

    # the reference's accessors of the projection matrix (features.py:43-53), kept by name
    def get_B(self):
        return self._B_spatial

    def set_B(self, B):
        self._B_spatial = B

    def save_B(self, filename):
        torch.save(self.get_B(), filename)

    def load_B(self, filename):
        self.set_B(torch.load(filename))


class AsinhTransform(torch.nn.Module):
    """features.py:56-60."""

    def forward(self, img):
        return torch.asinh(40 * img)
