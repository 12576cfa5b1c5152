"""Quality metrics on the device: host mirror of the metric calls of the reference's
test loop (ref: methods/__init__.py:32-40) for CUDA tensors, so that evaluation never leaves the GPU.

``icid`` keeps the signature of ref: utils/icid.py:28; ``psnr`` / ``ssim`` are piq.psnr / piq.ssim
with their defaults.
Both take float image batches [B,3,H,W] (or one [3,H,W] image) in [0,1] on a CUDA device and
return a 0-dim float32 tensor on that device, like the functions they replace.  There is no CPU
path: tensors that are not on a CUDA device raise.
"""

import ctypes

import torch

from . import _cabi
from .device import _handle_for

_INTENTS = {"perceptual": 0, "hue-preserving": 1, "chromatic": 2}


def _planar_f32(x, name):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (there is no CPU fallback)")
    if x.dim() == 3:
        x = x[None]
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError(f"{name} must have shape [B,3,H,W] or [3,H,W], got {tuple(x.shape)}")
    return x.detach().to(torch.float32).contiguous()


def icid(img1, img2, intent="perceptual", omit_maps67=False, downsampling=True):
    """improved colour-image-difference of two image batches (ref: utils/icid.py:28-152)."""
    if intent not in _INTENTS:
        raise ValueError("Intent should be either 'perceptual', 'hue-preserving', or 'chromatic'")
    a, b = _planar_f32(img1, "img1"), _planar_f32(img2, "img2")
    if a.shape != b.shape:
        raise ValueError(f"img1 and img2 differ in shape: {tuple(a.shape)} vs {tuple(b.shape)}")
    h = _handle_for(a, None)
    out = ctypes.c_double()
    n, _, hh, ww = a.shape
    h.check(h.lib.ct_icid(h.h, ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()), n, hh, ww,
                          _INTENTS[intent], int(bool(omit_maps67)), int(bool(downsampling)), ctypes.byref(out)))
    return torch.tensor(out.value, dtype=torch.float32, device=a.device)


def psnr(x, y):
    """piq.psnr(x, y) (data_range 1, mean over the batch; ref: methods/__init__.py:35)."""
    a, b = _planar_f32(x, "x"), _planar_f32(y, "y")
    if a.shape != b.shape:
        raise ValueError(f"x and y differ in shape: {tuple(a.shape)} vs {tuple(b.shape)}")
    h = _handle_for(a, None)
    out = ctypes.c_double()
    h.check(h.lib.ct_psnr(h.h, ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()), a.shape[0],
                          a[0].numel(), ctypes.byref(out)))
    return torch.tensor(out.value, dtype=torch.float32, device=a.device)


def ssim(x, y, downsample=True):
    """piq.ssim(x, y) with its defaults (11x11 Gaussian, sigma 1.5, k1 = 0.01, k2 = 0.03,
    data_range 1, mean over the batch; ref: methods/__init__.py:36)."""
    a, b = _planar_f32(x, "x"), _planar_f32(y, "y")
    if a.shape != b.shape:
        raise ValueError(f"x and y differ in shape: {tuple(a.shape)} vs {tuple(b.shape)}")
    h = _handle_for(a, None)
    out = ctypes.c_double()
    n, _, hh, ww = a.shape
    h.check(h.lib.ct_ssim(h.h, ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()), n, hh, ww,
                          int(bool(downsample)), ctypes.byref(out)))
    return torch.tensor(out.value, dtype=torch.float32, device=a.device)
