"""Drop-in hook for an unmodified checkout of the reference.

``patch_reference(modules)`` rebinds ``modules.BatchLinear`` and ``modules.FCBlock`` to the native
classes, built on the reference's OWN torchmeta base classes so that every isinstance check and
``meta_named_parameters()`` walk in the reference (torchmeta/modules/container.py:11,
meta_modules.py:23) keeps working.  ``modules.SingleBVPNet`` (modules.py:122-170) is left alone:
it looks ``FCBlock`` up by name at construction time, so models built after the patch get the
native block, including its rbf / nerf / downsampling front ends; only its ``forward`` is wrapped so that the tag of a
lazy Fourier-feature transform (siren_mri_b200.features) survives the clone of the coordinates and (``fuse_dc``) the
k-space samples / mask of the data consistency that follows the network reach the kernels' output epilogue.  training.py,
training_ddp.py, loss_functions.py and diff_operators.py need no change.  See INTEGRATION.md.
"""
from . import modules as _native


def patch_reference(ref_modules, ref_diff_operators=None, ref_data_consistency=None):
    """``ref_diff_operators`` (optional): the reference's diff_operators module; its ``hessian`` (diff_operators.py:5-24)
    is wrapped so that outputs of the native JET path (coord_derivs 1 / 2: first and diagonal second derivatives
    only) are re-evaluated as the composed graph before mixed second derivatives are taken.
    ``ref_data_consistency`` (optional): the reference's data_consistency module; its
    ``DataConsistencyInKspace.forward`` (data_consistency.py:32-47) is wrapped to pass on a prediction the kernels'
    output epilogue has already made data-consistent, and that fusion is switched on (``fuse_dc``): the neural-process
    models (meta_modules.py:173-232) then run hypo-network + data consistency as one kernel pass."""
    from torchmeta.modules import MetaModule, MetaSequential          # the reference's vendored copy
    from torchmeta.modules.utils import get_subdict
    BatchLinear, FCBlock, _ = _native.build_classes(MetaModule, MetaSequential, get_subdict)
    ref_modules._reference_BatchLinear = ref_modules.BatchLinear
    ref_modules._reference_FCBlock = ref_modules.FCBlock
    ref_modules.BatchLinear = BatchLinear
    ref_modules.FCBlock = FCBlock
    _carry_fourier_tag(ref_modules.SingleBVPNet)
    if ref_diff_operators is not None and not hasattr(ref_diff_operators, "_reference_hessian"):
        from .functional import composed_of
        orig = ref_diff_operators.hessian
        ref_diff_operators._reference_hessian = orig
        ref_diff_operators.hessian = lambda y, x: orig(composed_of(y), x)
    if ref_data_consistency is not None:
        cls = ref_data_consistency.DataConsistencyInKspace
        if not hasattr(cls, "_reference_forward"):
            from .data_consistency import check_fused_noise
            orig_dc = cls.forward

            def dc_forward(self, prediction, k0, mask):
                done = getattr(prediction, "_siren_dc_done", None)
                if done is None:
                    return orig_dc(self, prediction, k0, mask)
                check_fused_noise(done, self.noise_lvl)
                return prediction

            cls._reference_forward = orig_dc
            cls.forward = dc_forward
        from . import config
        config.set_defaults(fuse_dc=True)
    return ref_modules


def _carry_fourier_tag(bvp_cls):
    """The reference's ``SingleBVPNet.forward`` clones the coordinates (modules.py:151), which drops the tag a lazy
    ``features.GaussianFourierFeatureTransform`` put on them; this wrapper hands the tag to the (native) FCBlock for the
    duration of the call, so the reference's own class -- and the neural-process models built on it
    (meta_modules.py:173-232) -- take the in-kernel Fourier prologue too."""
    if hasattr(bvp_cls, "_reference_forward"):
        return
    orig = bvp_cls.forward

    def forward(self, model_input, params=None):
        from . import config
        coords = model_input.get("coords", None) if hasattr(model_input, "get") else None
        tag = getattr(coords, "_siren_fourier", None)
        dc = None      # fuse_dc: the k-space data consistency that follows this network (meta_modules.py:217-219)
        fuse = getattr(self, "fuse_dc", None)
        if (config.get_defaults()["fuse_dc"] if fuse is None else fuse) and hasattr(model_input, "get") \
                and "img_sparse" in model_input and "dc_mask" in model_input:
            dc = (model_input["img_sparse"], model_input["dc_mask"], getattr(self, "dc_noise_lvl", None))
        if (tag is None and dc is None) or getattr(self, "mode", "mlp") != "mlp":
            return orig(self, model_input, params)
        self.net._siren_pending_fourier = tag
        self.net._siren_pending_dc = dc
        try:
            return orig(self, model_input, params)
        finally:
            self.net._siren_pending_fourier = None
            self.net._siren_pending_dc = None

    bvp_cls._reference_forward = orig
    bvp_cls.forward = forward


def unpatch_reference(ref_modules):
    if hasattr(ref_modules, "_reference_FCBlock"):
        ref_modules.FCBlock = ref_modules._reference_FCBlock
        ref_modules.BatchLinear = ref_modules._reference_BatchLinear
    if hasattr(ref_modules.SingleBVPNet, "_reference_forward"):
        ref_modules.SingleBVPNet.forward = ref_modules.SingleBVPNet._reference_forward
        del ref_modules.SingleBVPNet._reference_forward
    return ref_modules


def unpatch_data_consistency(ref_data_consistency):
    cls = ref_data_consistency.DataConsistencyInKspace
    if hasattr(cls, "_reference_forward"):
        cls.forward = cls._reference_forward
        del cls._reference_forward
    from . import config
    config.set_defaults(fuse_dc=False)
    return ref_data_consistency
