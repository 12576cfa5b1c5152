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


def _img_as_float(x, name):
    """What ``skimage.color.rgb2lab`` does to its argument first (``img_as_float`` via
    ``_prepare_colorarray``, ref: methods/linear.py:25-26): unsigned integers are divided by their
    maximum, float16 becomes float32, float32 / float64 are kept."""
    a = np.asarray(x)
    if a.ndim != 3 or a.shape[-1] != 3:
        raise ValueError(f"{name} must have shape [H, W, 3], got {a.shape}")
    if a.dtype == np.float32 or a.dtype == np.float64:
        return a
    if a.dtype == np.float16:
        return a.astype(np.float32)
    if a.dtype.kind == "u":
        return a.astype(np.float64) / np.iinfo(a.dtype).max
    if a.dtype.kind == "i":   # signed integers map to [-1, 1]
        return a.astype(np.float64) / np.iinfo(a.dtype).max
    if a.dtype == np.bool_:
        return a.astype(np.float64)
    return a.astype(np.float64)


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
    if rc == _cabi.CT_E_SINGULAR and method == _cabi.CT_CCS:
        # 1 / sqrt(0) in the reference (linear.py:73-78): inf / NaN propagate, nothing is raised
        rc = _cabi.CT_OK
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
    t = _img_as_float(target, "target")
    r = _img_as_float(reference, "reference")
    # (target_lab - mean_t) * std_r / std_t + mean_r promotes to the wider of the two float dtypes
    out_dtype = np.result_type(t.dtype, r.dtype)
    if t.dtype != out_dtype:
        t = t.astype(out_dtype)
    return _run(_cabi.CT_REINHARD, t, r, out_dtype, out, handle)


def _mean_cov(sums):
    """np.mean(axis=0) and np.cov(X.T) (ddof 1) from the device's raw moments about 0.5
    {n, S(x-K), S(x-K)(x-K)^T as 00,01,02,11,12,22} (ref: methods/linear.py:64-67)."""
    n = sums[0]
    m1 = sums[1:4] / n
    sxx = np.array([[sums[4], sums[5], sums[6]], [sums[5], sums[7], sums[8]], [sums[6], sums[8], sums[9]]])
    return m1 + 0.5, (sxx - n * np.outer(m1, m1)) / (n - 1)


def color_transfer_in_correlated_color_space(target, reference, *, out=None, handle=None, lapack_signs=True):
    """Color Transfer in Correlated Color Space (Xiao & Ma, 2006) - ref: methods/linear.py:45-82.

    Returns float64, not clipped.  The transform depends on the signs LAPACK happens to give the
    singular vectors of the two 3x3 covariances, which no device-side convention can predict (on random
    covariance pairs the natural rule "orient u_r along u_t" disagrees in more than half the cases).
    The drop-in therefore brings the covariances (reduced on the device) back to the host, runs the
    reference's own three lines on them - np.linalg.svd, i.e. dgesdd (linear.py:69-78) - and applies
    the resulting matrix on the device.  ``lapack_signs=False`` keeps everything on the device with the
    orientation rule (what the batched / device-resident entry points do).
    """
    if not lapack_signs:
        return _run(_cabi.CT_CCS, target, reference, np.float64, out, handle)
    t = _as_image(target, "target")
    r = _as_image(reference, "reference")
    h = handle or _cabi.default_handle()
    tb, t_keep = _cabi.batch_from_numpy(t)
    rb, r_keep = _cabi.batch_from_numpy(r)
    if out is None:
        out = np.empty(t.shape, dtype=np.float64)
    elif out.shape != t.shape or out.dtype != np.float64 or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous float64 array of the target's shape")
    ob, _ = _cabi.batch_from_numpy(out)
    sums_t, sums_r = np.empty(_cabi.CT_MOMENT_DOUBLES), np.empty(_cabi.CT_MOMENT_DOUBLES)
    h.check(h.lib.ct_linear_stats_host(h.h, 0, tb, rb, sums_t.ctypes.data, sums_r.ctypes.data))
    mean_t, cov_t = _mean_cov(sums_t)
    mean_r, cov_r = _mean_cov(sums_r)
    with np.errstate(divide="ignore", invalid="ignore"):      # a singular covariance propagates inf / NaN like the reference
        u_t, s_t, _ = np.linalg.svd(cov_t)                    # linear.py:69
        u_r, s_r, _ = np.linalg.svd(cov_r)                    # linear.py:70
        transform = u_t @ np.diag(1 / np.sqrt(s_t)) @ np.diag(np.sqrt(s_r)) @ np.linalg.inv(u_r)   # linear.py:72-78
    xform = np.zeros(_cabi.CT_XFORM_DOUBLES)
    xform[:9] = transform.T.reshape(-1)       # out = (x - mean_t) @ M + mean_r with M = T.T (linear.py:80)
    xform[9:12] = mean_t
    xform[12:15] = mean_r
    h.check(h.lib.ct_linear_apply_staged_host(h.h, _cabi.CT_CCS, xform.ctypes.data, ob))
    del t_keep, r_keep
    return out


def monge_kantorovitch_color_transfer(target, reference, decomposition="MK", *, out=None, handle=None):
    """Linear Monge-Kantorovitch colour mapping (Pitie & Kokaram, 2007) - ref: methods/linear.py:85-124.

    ``decomposition`` is "MK", "sqrt" or "cholesky"; anything else raises the reference's
    ValueError.  Returns float64, not clipped.
    """
    if decomposition not in _DECOMPOSITIONS:
        raise ValueError("Unknown decomposition, use either 'cholesky', 'sqrt', or 'MK'")
    return _run(_DECOMPOSITIONS[decomposition], target, reference, np.float64, out, handle)


def _device_impl(method):
    def run(target, reference, out_dtype=None, clamp=False):
        """[B,H,W,3] CUDA tensors (float, or uint8 frames decoded as float32 like the reference's dataset
        loader) -> [B,H,W,3] CUDA tensor; same kernels, no host round trip."""
        from .. import device
        return device.linear_transfer(method, target, reference, out_dtype=out_dtype, clamp=clamp)
    return run


color_transfer_between_images.device_impl = _device_impl(_cabi.CT_REINHARD)
color_transfer_in_correlated_color_space.device_impl = _device_impl(_cabi.CT_CCS)
monge_kantorovitch_color_transfer.device_impl = _device_impl(_cabi.CT_MKL_MK)
