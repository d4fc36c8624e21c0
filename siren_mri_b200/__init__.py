"""siren_mri_b200 -- B200-native SIREN hot path behind the reference's nn.Module API.

Public surface (mirrors /root/reference/modules.py for the path named in BASELINE.json):

    from siren_mri_b200 import modules            # BatchLinear, Sine, FCBlock, SingleBVPNet
    from siren_mri_b200 import meta_modules       # HyperNetwork (per-sample params producer)
    from siren_mri_b200 import diff_operators     # gradient / divergence / laplace / jacobian / hessian
    from siren_mri_b200.optim import FusedAdam    # clip + Adam over one flat buffer
    from siren_mri_b200.trainer import SirenTrainer   # graph-captured fwd+loss+bwd+allreduce+Adam step
    from siren_mri_b200.training import train_fast    # training.train (training.py:19-146) on that step

The arithmetic runs in csrc/libsiren_b200.so (hand-written sm_100a kernels, C ABI in
include/siren_b200.h).  On a CUDA device the library MUST load: there is no silent fallback.
"""
from . import config  # noqa: F401
from .config import set_defaults, get_defaults  # noqa: F401

__all__ = ["config", "set_defaults", "get_defaults"]
__version__ = "0.1.0"
