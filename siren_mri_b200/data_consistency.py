"""Host-side mirror of the reference's data_consistency.py (k-space data consistency of the MRI neural-process
models, meta_modules.py:217-219): same function and module, plain PyTorch ops -- an elementwise pass over the
``[B, N, 2]`` network output, which autograd differentiates as in the reference."""
import torch.nn as nn


def data_consistency(pred, k0, mask, noise_lvl=None):
    """data_consistency.py:7-20 as one blend: where ``mask`` is set the prediction moves to the sampled value ``k0``
    (noiseless) or to the noise-weighted mean ``(pred + v k0) / (1 + v)``; elsewhere it stays."""
    pull = 1.0 if not noise_lvl else noise_lvl / (1.0 + noise_lvl)
    return pred + (mask * pull) * (k0 - pred)


class DataConsistencyInKspace(nn.Module):
    """data_consistency.py:23-47.  ``prediction`` [B, nspatial, 2]; ``k0`` / ``mask`` [B, 2, nx, ny] (channel first, as
    the datasets deliver them) are brought to the prediction's [B, nspatial, 2] layout."""

    def __init__(self, noise_lvl=None):
        super().__init__()
        self.noise_lvl = noise_lvl

    @staticmethod
    def _channels_last(t):
        return t.permute(0, 2, 3, 1).reshape(t.shape[0], -1, t.shape[1])

    def forward(self, prediction, k0, mask):
        return data_consistency(prediction, self._channels_last(k0), self._channels_last(mask), self.noise_lvl)
