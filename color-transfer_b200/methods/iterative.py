"""Iterative Color Transfer Methods - B200 kernels behind the reference's function name.

Mirror of ref: methods/iterative.py:8-59 (``iterative_distribution_transfer``).  The random
rotations are drawn here, on the host, with the very call the reference makes
(``scipy.stats.special_ortho_group.rvs(3)`` once per iteration, in order), so the global numpy
RNG advances exactly as in the reference and ``np.random.seed(s)`` reproduces its matrices.
``automated_color_grading`` = IDT + regrain (ref: methods/iterative.py:62-138, SURVEY.md
section 8f-2) runs on the device in one call as well.
"""

import ctypes

import numpy as np
import scipy.stats

from .. import _cabi

__all__ = ["iterative_distribution_transfer", "automated_color_grading", "draw_rotations"]


def draw_rotations(n_iter, n_dims=3):
    """n_iter Haar-random rotations from the global numpy RNG, in the reference's order
    (ref: methods/iterative.py:32)."""
    return np.stack([scipy.stats.special_ortho_group.rvs(n_dims) for _ in range(n_iter)]) \
        if n_iter > 0 else np.empty((0, n_dims, n_dims))


def iterative_distribution_transfer(target, reference, bins=255, n_iter=4, *, rotations=None, out=None,
                                    trace=None, handle=None):
    """Iterative Distribution Transfer (Pitie, Kokaram & Dahyot, 2007) - ref: methods/iterative.py:8-59.

    Returns float64 ``[H,W,3]``, not clipped.  Keyword-only extras (not in the reference):
    ``rotations`` ([n_iter,3,3], skips the RNG draws), ``out`` (preallocated result), ``trace``
    (a dict that receives lo / hi / counts_t / counts_r / lut per iteration, for parity tests).
    """
    t = np.asarray(target)
    r = np.asarray(reference)
    if t.ndim != 3 or r.ndim != 3:
        raise ValueError("target and reference must have shape [H, W, 3]")
    if t.shape[-1] != 3 or r.shape[-1] != 3:
        # the reference reshapes to (-1, 3) and draws rotations of dimension shape[-1]
        raise ValueError("only 3-channel images are supported")
    if t.dtype != np.float32 and t.dtype != np.float64:
        t = t.astype(np.float64)
    if r.dtype != np.float32 and r.dtype != np.float64:
        r = r.astype(np.float64)
    bins = int(bins)
    n_iter = int(n_iter)
    if rotations is None:
        rotations = draw_rotations(n_iter)
    rot = np.ascontiguousarray(rotations, dtype=np.float64).reshape(-1, 3, 3)
    if rot.shape[0] != n_iter:
        raise ValueError("rotations must hold n_iter matrices")
    if out is None:
        out = np.empty(t.shape, dtype=np.float64)
    elif out.shape != t.shape or out.dtype != np.float64 or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous float64 array of the target's shape")
    if n_iter <= 0:  # the reference's loop body never runs: it returns the (reshaped) input
        out[...] = t
        return out
    if bins < 1:
        raise ValueError("`bins` must be positive, when an integer")  # np.histogram's message
    h = handle or _cabi.default_handle()
    tb, t_keep = _cabi.batch_from_numpy(t)
    rb, r_keep = _cabi.batch_from_numpy(r)
    ob, _ = _cabi.batch_from_numpy(out)
    tr = None
    if trace is not None:
        trace["rot"] = rot
        trace["lo"] = np.empty((n_iter, 3))
        trace["hi"] = np.empty((n_iter, 3))
        trace["counts_t"] = np.empty((n_iter, 3, bins), dtype=np.int64)
        trace["counts_r"] = np.empty((n_iter, 3, bins), dtype=np.int64)
        trace["lut"] = np.empty((n_iter, 3, bins))
        tr = _cabi.IdtTrace(*(ctypes.c_void_p(trace[k].ctypes.data) for k in ("lo", "hi", "counts_t", "counts_r", "lut")))
    rc = h.lib.ct_idt_transfer_host(h.h, tb, rb, ob, ctypes.c_void_p(rot.ctypes.data), bins, n_iter,
                                    ctypes.byref(tr) if tr is not None else None)
    if rc == _cabi.CT_E_NONFINITE:
        raise ValueError("supplied range of projected values is not finite")  # np.histogram, _get_outer_edges
    if rc == _cabi.CT_E_UNSUPPORTED:
        raise NotImplementedError(h.lib.ct_last_error(h.h).decode())
    h.check(rc)
    del t_keep, r_keep
    return out


def automated_color_grading(target, reference, *, rotations=None, out=None, handle=None):
    """Automated Colour Grading using Colour Distribution Transfer (Pitie et al., 2007) -
    ref: methods/iterative.py:118-138: ``iterative_distribution_transfer(target, reference)``
    followed by the gradient-preserving regrain of the original target.  float64 ``[H,W,3]``."""
    t = np.asarray(target)
    r = np.asarray(reference)
    if t.ndim != 3 or r.ndim != 3 or t.shape[-1] != 3 or r.shape[-1] != 3:
        raise ValueError("target and reference must have shape [H, W, 3]")
    if t.dtype != np.float32 and t.dtype != np.float64:
        t = t.astype(np.float64)
    if r.dtype != np.float32 and r.dtype != np.float64:
        r = r.astype(np.float64)
    n_iter, bins = 4, 255                       # the reference calls IDT with its defaults
    if rotations is None:
        rotations = draw_rotations(n_iter)
    rot = np.ascontiguousarray(rotations, dtype=np.float64).reshape(n_iter, 3, 3)
    if out is None:
        out = np.empty(t.shape, dtype=np.float64)
    elif out.shape != t.shape or out.dtype != np.float64 or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous float64 array of the target's shape")
    h = handle or _cabi.default_handle()
    tb, t_keep = _cabi.batch_from_numpy(t)
    rb, r_keep = _cabi.batch_from_numpy(r)
    ob, _ = _cabi.batch_from_numpy(out)
    rc = h.lib.ct_acg_transfer_host(h.h, tb, rb, ob, t.shape[0], t.shape[1], ctypes.c_void_p(rot.ctypes.data), bins, n_iter)
    if rc == _cabi.CT_E_NONFINITE:
        raise ValueError("supplied range of projected values is not finite")
    h.check(rc)
    del t_keep, r_keep
    return out


def _idt_device_impl(target, reference, bins=255, n_iter=4, out_dtype=None, clamp=False):
    """[B,H,W,3] CUDA tensors (float, or uint8 frames decoded as float32) -> [B,H,W,3] CUDA tensor
    (float64 unless ``out_dtype`` says otherwise).  Rotations are drawn pair by pair from the global
    numpy RNG, n_iter per pair - the order a loop over the reference function sees."""
    import torch

    from .. import device
    b = target.shape[0]
    rot = np.stack([draw_rotations(n_iter) for _ in range(b)])
    if out_dtype is None:
        out_dtype = torch.float64
    return device.idt_transfer(target, reference, torch.from_numpy(rot).to(target.device), bins, n_iter,
                               out_dtype=out_dtype, clamp=clamp)


iterative_distribution_transfer.device_impl = _idt_device_impl
