"""Host-side mirror of the reference's data_consistency.py (k-space data consistency of the MRI neural-process
models, meta_modules.py:217-219): same function and module.  With ``fuse_dc`` (config / ``SingleBVPNet.fuse_dc``) the
blend is applied by the kernels where they complete a row's output (siren_b200_forward_dc / _backward_dc) and the
module passes the tagged prediction on; otherwise it is the reference's elementwise pass over the ``[B, N, 2]`` output,
which autograd differentiates."""
import torch.nn as nn

from . import functional


def data_consistency(pred, k0, mask, noise_lvl=None):
    """data_consistency.py:7-20 as one blend: where ``mask`` is set the prediction moves to the sampled value ``k0``
    (noiseless) or to the noise-weighted mean ``(pred + v k0) / (1 + v)``; elsewhere it stays."""
    return functional.data_consistency_blend(pred, k0, mask, noise_lvl)


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
        done = getattr(prediction, "_siren_dc_done", None)
        if done is not None:      # the kernels' output epilogue has applied the blend (modules.FCBlock.forward, fuse_dc)
            check_fused_noise(done, self.noise_lvl)
            return prediction
        return data_consistency(prediction, self._channels_last(k0), self._channels_last(mask), self.noise_lvl)


def check_fused_noise(done, noise_lvl):
    if abs(float(done) - float(noise_lvl or 0.0)) > 0.0:
        raise RuntimeError("siren_mri_b200: the fused data-consistency epilogue ran with noise_lvl=%g but this "
                           "DataConsistencyInKspace has noise_lvl=%r; set hypo_net.dc_noise_lvl to match or turn "
                           "fuse_dc off" % (done, noise_lvl))
