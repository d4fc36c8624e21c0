"""Process-wide defaults for the native path (precision and coordinate-derivative order)."""
import threading

_lock = threading.Lock()
_defaults = {
    # 'fp32': bf16 hi/lo split operands (3 MMAs per product) + fp32 stash -> rel err <= 1e-4 vs reference
    # 'bf16': single bf16 operands + bf16 stash (fast mode; documented bound in DESIGN.md)
    "precision": "fp32",
    # 0: value only (coordinate derivatives, if requested through autograd, take the composed
    #    PyTorch path); 1 / 2: forward-mode jets dy/dx_k (and d2y/dx_k^2) come from the kernels
    "coord_derivs": 0,
    # also produce d loss / d coords in plain backward() (the reference fills model_in.grad; nothing reads it)
    "coords_grad": False,
    # 'auto': native kernels for CUDA tensors inside the envelope, composed PyTorch otherwise
    # 'composed': never use the native kernels
    "backend": "auto",
    # SingleBVPNet.forward hands model_input['img_sparse'] / ['dc_mask'] to the kernels, whose output epilogue applies the
    # k-space data consistency of data_consistency.py:7-20; DataConsistencyInKspace then passes the tagged result on
    "fuse_dc": False,
}


def set_defaults(**kw):
    with _lock:
        for k, v in kw.items():
            if k not in _defaults:
                raise KeyError("unknown option %r (have %s)" % (k, sorted(_defaults)))
            _defaults[k] = v


def get_defaults():
    with _lock:
        return dict(_defaults)
