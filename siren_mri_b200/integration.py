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


def patch_reference(ref_modules):
    from torchmeta.modules import MetaModule, MetaSequential          # the reference's vendored copy
    from torchmeta.modules.utils import get_subdict
    BatchLinear, FCBlock, _ = _native.build_classes(MetaModule, MetaSequential, get_subdict)
    ref_modules._reference_BatchLinear = ref_modules.BatchLinear
    ref_modules._reference_FCBlock = ref_modules.FCBlock
    ref_modules.BatchLinear = BatchLinear
    ref_modules.FCBlock = FCBlock
    return ref_modules


def unpatch_reference(ref_modules):
    if hasattr(ref_modules, "_reference_FCBlock"):
        ref_modules.FCBlock = ref_modules._reference_FCBlock
        ref_modules.BatchLinear = ref_modules._reference_BatchLinear
    return ref_modules
