"""Batched host-memory entry points: many pairs / video frames per call.

The reference processes a batch with a Python loop of single-pair calls
(ref: methods/__init__.py:20-27) and its dataset tool calls the MKL transfer once per video
frame (ref: utils/postprocess.py:120-144).  These functions take the whole stack
``[B,H,W,3]`` in one C-ABI call so that host-to-device copies, kernels and device-to-host
copies of consecutive pairs overlap on three streams.  Pass pinned arrays (e.g. from
``pinned_empty``) to get full PCIe bandwidth.
"""

import ctypes

import numpy as np

from . import _cabi
from .methods.iterative import draw_rotations

_METHODS = {"reinhard": _cabi.CT_REINHARD, "ccs": _cabi.CT_CCS, "mkl": _cabi.CT_MKL_MK,
            "mkl_sqrt": _cabi.CT_MKL_SQRT, "mkl_cholesky": _cabi.CT_MKL_CHOLESKY}


def pinned_empty(shape, dtype):
    """A numpy array backed by page-locked memory (allocated through torch)."""
    import torch
    t = torch.empty(tuple(shape), dtype={np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
                                         np.dtype(np.uint8): torch.uint8}[np.dtype(dtype)]).pin_memory()
    a = t.numpy()
    return a


def _stack_u8(x, name):
    a = np.asarray(x)
    if a.ndim != 4 or a.shape[-1] != 3 or a.dtype != np.uint8:
        raise ValueError(f"{name} must be a uint8 array of shape [B, H, W, 3]")
    return np.ascontiguousarray(a)


def linear_transfer_frames_u8(method, targets, references, out=None, as_float32=True, handle=None):
    """uint8 frames in, uint8 frames out: decode k/255 (float32 like the reference's dataset loader,
    or float64 like skimage.img_as_float), transfer, clip to [0,1], round (img_as_ubyte)."""
    t, r = _stack_u8(targets, "targets"), _stack_u8(references, "references")
    if t.shape[0] != r.shape[0]:
        raise ValueError("targets and references must hold the same number of pairs")
    if out is None:
        out = np.empty(t.shape, dtype=np.uint8)
    elif out.shape != t.shape or out.dtype != np.uint8 or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous uint8 array of the targets' shape")
    h = handle or _cabi.default_handle()
    rc = h.lib.ct_linear_transfer_host_u8(h.h, _METHODS[method], ctypes.c_void_p(t.ctypes.data), ctypes.c_void_p(r.ctypes.data),
                                          ctypes.c_void_p(out.ctypes.data), t.shape[0], t.shape[1] * t.shape[2],
                                          r.shape[1] * r.shape[2], 1 if as_float32 else 0)
    if rc in (_cabi.CT_E_NOT_PD, _cabi.CT_E_SINGULAR):
        raise np.linalg.LinAlgError(h.lib.ct_last_error(h.h).decode())
    h.check(rc)
    return out


def idt_frames_u8(targets, references, bins=255, n_iter=4, rotations=None, out=None, as_float32=True, handle=None):
    """IDT on uint8 frames (see linear_transfer_frames_u8 for the decode / encode convention)."""
    t, r = _stack_u8(targets, "targets"), _stack_u8(references, "references")
    b = t.shape[0]
    if r.shape[0] != b:
        raise ValueError("targets and references must hold the same number of pairs")
    if rotations is None:
        rotations = np.stack([draw_rotations(n_iter) for _ in range(b)])
    rot = np.ascontiguousarray(rotations, dtype=np.float64).reshape(b, n_iter, 3, 3)
    if out is None:
        out = np.empty(t.shape, dtype=np.uint8)
    elif out.shape != t.shape or out.dtype != np.uint8 or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous uint8 array of the targets' shape")
    h = handle or _cabi.default_handle()
    rc = h.lib.ct_idt_transfer_host_u8(h.h, ctypes.c_void_p(t.ctypes.data), ctypes.c_void_p(r.ctypes.data),
                                       ctypes.c_void_p(out.ctypes.data), b, t.shape[1] * t.shape[2], r.shape[1] * r.shape[2],
                                       1 if as_float32 else 0, ctypes.c_void_p(rot.ctypes.data), int(bins), int(n_iter))
    if rc == _cabi.CT_E_NONFINITE:
        raise ValueError("supplied range of projected values is not finite")
    h.check(rc)
    return out


def _stack(x, name):
    a = np.asarray(x)
    if a.ndim != 4 or a.shape[-1] != 3:
        raise ValueError(f"{name} must have shape [B, H, W, 3]")
    if a.dtype != np.float32 and a.dtype != np.float64:
        a = a.astype(np.float64)
    return a


def linear_transfer_frames(method, targets, references, out=None, handle=None):
    """``method`` in {"reinhard","ccs","mkl","mkl_sqrt","mkl_cholesky"}; one transfer per pair."""
    t, r = _stack(targets, "targets"), _stack(references, "references")
    if t.shape[0] != r.shape[0]:
        raise ValueError("targets and references must hold the same number of pairs")
    code = _METHODS[method]
    out_dtype = t.dtype if code == _cabi.CT_REINHARD else np.float64
    if out is None:
        out = np.empty(t.shape, dtype=out_dtype)
    h = handle or _cabi.default_handle()
    tb, k1 = _cabi.batch_from_numpy(t)
    rb, k2 = _cabi.batch_from_numpy(r)
    ob, k3 = _cabi.batch_from_numpy(out)
    if k3 is not out or out.dtype != out_dtype:
        raise ValueError("out must be C-contiguous with the result dtype")
    rc = h.lib.ct_linear_transfer_host(h.h, code, tb, rb, ob)
    if rc in (_cabi.CT_E_NOT_PD, _cabi.CT_E_SINGULAR):
        raise np.linalg.LinAlgError(h.lib.ct_last_error(h.h).decode())
    h.check(rc)
    return out


def idt_frames(targets, references, bins=255, n_iter=4, rotations=None, out=None, handle=None):
    """IDT of every pair of the stack.  Rotations are drawn frame by frame, ``n_iter`` per frame,
    from the global numpy RNG - the order a sequential loop over the reference function sees."""
    t, r = _stack(targets, "targets"), _stack(references, "references")
    b = t.shape[0]
    if r.shape[0] != b:
        raise ValueError("targets and references must hold the same number of pairs")
    if rotations is None:
        rotations = np.stack([draw_rotations(n_iter) for _ in range(b)])
    rot = np.ascontiguousarray(rotations, dtype=np.float64).reshape(b, n_iter, 3, 3)
    if out is None:
        out = np.empty(t.shape, dtype=np.float64)
    h = handle or _cabi.default_handle()
    tb, k1 = _cabi.batch_from_numpy(t)
    rb, k2 = _cabi.batch_from_numpy(r)
    ob, k3 = _cabi.batch_from_numpy(out)
    if k3 is not out or out.dtype != np.float64:
        raise ValueError("out must be a C-contiguous float64 array")
    rc = h.lib.ct_idt_transfer_host(h.h, tb, rb, ob, ctypes.c_void_p(rot.ctypes.data), int(bins), int(n_iter), None)
    if rc == _cabi.CT_E_NONFINITE:
        raise ValueError("supplied range of projected values is not finite")
    h.check(rc)
    return out
