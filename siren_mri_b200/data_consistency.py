"""Host-side mirror of the reference's data_consistency.py (k-space data consistency of the MRI neural-process
models, meta_modules.py:217-219): same function and module, plain PyTorch ops -- an elementwise pass over the
``[B, N, 2]`` network output, which autograd differentiates as in the reference."""
import torch.nn as nn


def data_consistency(pred, k0, mask, noise_lvl=None):
    """data_consistency.py:7-20: keep the prediction where nothing was sampled, the sampled value elsewhere."""
    v = noise_lvl
    if v:
        return (1 - mask) * pred + mask * (pred + v * k0) / (1 + v)
    return (1 - mask) * pred + mask * k0


class DataConsistencyInKspace(nn.Module):
    """data_consistency.py:23-47.  ``prediction`` [B, nspatial, 2]; ``k0`` / ``mask`` [B, 2, nx, ny]."""

    def __init__(self, noise_lvl=None):
        super().__init__()
        self.noise_lvl = noise_lvl

    def forward(self, prediction, k0, mask):
        batch = k0.shape[0]
        k0 = k0.permute(0, 2, 3, 1).reshape(batch, -1, 2)
        mask = mask.permute(0, 2, 3, 1).reshape(batch, -1, 2)
        return data_consistency(prediction, k0, mask, self.noise_lvl)
