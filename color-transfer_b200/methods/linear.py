"""Linear Color Transfer Methods - B200 kernels behind the reference's function names.

Mirror of ref: methods/linear.py.  Same names, positional signatures, array conventions
(``[H,W,3]`` float RGB in, new ``[H,W,3]`` array out, inputs untouched), return dtypes and
exceptions; the arithmetic runs in libct_b200.so (moments -> one-warp 3x3 solve -> fused remap).
There is no CPU path: without the library or a CUDA device these functions raise.
"""

import numpy as np

from .. import _cabi

__all__ = ["color_transfer_between_images", "color_transfer_in_correlated_color_space",
           "monge_kantorovitch_color_transfer"]

_DECOMPOSITIONS = {"MK": _cabi.CT_MKL_MK, "sqrt": _cabi.CT_MKL_SQRT, "cholesky": _cabi.CT_MKL_CHOLESKY}


def _as_image(x, name):
    a = np.asarray(x)
    if a.ndim != 3 or a.shape[-1] != 3:
        raise ValueError(f"{name} must have shape [H, W, 3], got {a.shape}")
    if a.dtype != np.float32 and a.dtype != np.float64:
        # the reference's float math promotes integer / half inputs to float64
        a = a.astype(np.float64)
    return a


def _run(method, target, reference, out_dtype, out=None, handle=None):
    t = _as_image(target, "target")
    r = _as_image(reference, "reference")
    h = handle or _cabi.default_handle()
    tb, t_keep = _cabi.batch_from_numpy(t)
    rb, r_keep = _cabi.batch_from_numpy(r)
    if out is None:
        out = np.empty(t.shape, dtype=out_dtype)
    elif out.shape != t.shape or out.dtype != out_dtype or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous array of the target's shape and the result dtype")
    ob, _ = _cabi.batch_from_numpy(out)
    rc = h.lib.ct_linear_transfer_host(h.h, method, tb, rb, ob)
    if rc in (_cabi.CT_E_NOT_PD, _cabi.CT_E_SINGULAR):
        # np.linalg.cholesky / np.linalg.inv raise LinAlgError on such covariances
        raise np.linalg.LinAlgError(h.lib.ct_last_error(h.h).decode())
    h.check(rc)
    del t_keep, r_keep
    return out


def color_transfer_between_images(target, reference, *, out=None, handle=None):
    """Color Transfer between Images (Reinhard et al., 2001) - ref: methods/linear.py:8-42.

    Per-channel mean / standard-deviation matching in CIE-Lab.  Returns RGB clipped to [0, 1]
    with the float dtype of ``target`` (float32 stays float32 as in scikit-image >= 0.19).
    """
    t = np.asarray(target)
    out_dtype = np.float32 if t.dtype == np.float32 else np.float64
    return _run(_cabi.CT_REINHARD, target, reference, out_dtype, out, handle)


def color_transfer_in_correlated_color_space(target, reference, *, out=None, handle=None):
    """Color Transfer in Correlated Color Space (Xiao & Ma, 2006) - ref: methods/linear.py:45-82.

    Returns float64, not clipped.  The singular-vector signs that LAPACK leaves arbitrary are
    fixed by orienting each reference axis along its target axis.
    """
    return _run(_cabi.CT_CCS, target, reference, np.float64, out, handle)


def monge_kantorovitch_color_transfer(target, reference, decomposition="MK", *, out=None, handle=None):
    """Linear Monge-Kantorovitch colour mapping (Pitie & Kokaram, 2007) - ref: methods/linear.py:85-124.

    ``decomposition`` is "MK", "sqrt" or "cholesky"; anything else raises the reference's
    ValueError.  Returns float64, not clipped.
    """
    if decomposition not in _DECOMPOSITIONS:
        raise ValueError("Unknown decomposition, use either 'cholesky', 'sqrt', or 'MK'")
    return _run(_DECOMPOSITIONS[decomposition], target, reference, np.float64, out, handle)


def _device_impl(method):
    def run(target, reference):
        """[B,H,W,3] CUDA tensors -> [B,H,W,3] CUDA tensor; same kernels, no host round trip."""
        from .. import device
        return device.linear_transfer(method, target, reference)
    return run


color_transfer_between_images.device_impl = _device_impl(_cabi.CT_REINHARD)
color_transfer_in_correlated_color_space.device_impl = _device_impl(_cabi.CT_CCS)
monge_kantorovitch_color_transfer.device_impl = _device_impl(_cabi.CT_MKL_MK)
