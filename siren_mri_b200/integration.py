"""Drop-in hook for an unmodified checkout of the reference.

``patch_reference(modules)`` rebinds ``modules.BatchLinear`` and ``modules.FCBlock`` to the native
classes, built on the reference's OWN torchmeta base classes so that every isinstance check and
``meta_named_parameters()`` walk in the reference (torchmeta/modules/container.py:11,
meta_modules.py:23) keeps working.  ``modules.SingleBVPNet`` (modules.py:122-170) is left alone:
it looks ``FCBlock`` up by name at construction time, so models built after the patch get the
native block, including its rbf / nerf / downsampling front ends.  training.py, training_ddp.py,
loss_functions.py and diff_operators.py need no change.  See INTEGRATION.md.
"""
from . import modules as _native


def patch_reference(ref_modules, ref_diff_operators=None):
    """``ref_diff_operators`` (optional): the reference's diff_operators module; its ``hessian`` (diff_operators.py:5-24)
    is wrapped so that outputs of the native JET path (coord_derivs 1 / 2: first and diagonal second derivatives
    only) are re-evaluated as the composed graph before mixed second derivatives are taken."""
    from torchmeta.modules import MetaModule, MetaSequential          # the reference's vendored copy
    from torchmeta.modules.utils import get_subdict
    BatchLinear, FCBlock, _ = _native.build_classes(MetaModule, MetaSequential, get_subdict)
    ref_modules._reference_BatchLinear = ref_modules.BatchLinear
    ref_modules._reference_FCBlock = ref_modules.FCBlock
    ref_modules.BatchLinear = BatchLinear
    ref_modules.FCBlock = FCBlock
    if ref_diff_operators is not None and not hasattr(ref_diff_operators, "_reference_hessian"):
        from .functional import composed_of
        orig = ref_diff_operators.hessian
        ref_diff_operators._reference_hessian = orig
        ref_diff_operators.hessian = lambda y, x: orig(composed_of(y), x)
    return ref_modules


def unpatch_reference(ref_modules):
    if hasattr(ref_modules, "_reference_FCBlock"):
        ref_modules.FCBlock = ref_modules._reference_FCBlock
        ref_modules.BatchLinear = ref_modules._reference_BatchLinear
    return ref_modules
