"""Test-set side of the hot path: host mirror of ref: utils/data.py for CUDA tensors.

``setup_grid_distortions`` keeps the reference's name, arguments and order (ref: utils/data.py:12-22):
identity, then for each of ``num`` magnitudes brightness, contrast, saturation, hue and gamma.  Each
entry is a callable on a uint8 ``[3,H,W]`` (or ``[B,3,H,W]``) CUDA tensor like the
``functools.partial(F.adjust_*)`` objects it replaces; ``distort_grid`` applies a whole list in ONE
pass over the image (libct_b200.so ``ct_distort``: 3 bytes per pixel in, 3 per distorted copy out),
which is how ``ArtificialTestDataset`` produces the 31 targets of one ground-truth image.

Images stay uint8 on the device: the transfers decode them in their kernels (``device.py``), so the
``/ 255`` of ref: utils/data.py:106 never materialises unless ``as_float`` asks for it.
There is no CPU path: tensors that are not on a CUDA device raise.
"""

import ctypes
from pathlib import Path

import numpy as np
import torch

from . import _cabi
from .device import _handle_for

IDENTITY, BRIGHTNESS, CONTRAST, SATURATION, HUE, GAMMA = range(6)
_NAMES = {IDENTITY: "identity", BRIGHTNESS: "adjust_brightness", CONTRAST: "adjust_contrast",
          SATURATION: "adjust_saturation", HUE: "adjust_hue", GAMMA: "adjust_gamma"}
MAX_OPS = 32


class Distortion(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("reserved", ctypes.c_int32), ("factor", ctypes.c_double)]


def _check(img):
    if not isinstance(img, torch.Tensor):
        raise TypeError("Input img should be Tensor image")
    if not img.is_cuda:
        raise TypeError("img must be a CUDA tensor (there is no CPU fallback)")
    if img.dtype != torch.uint8:
        raise TypeError("the distortion generator takes uint8 images (ref: utils/data.py:101-104)")
    x = img[None] if img.dim() == 3 else img
    if x.dim() != 4 or x.shape[1] != 3:
        raise TypeError(f"img must be [3,H,W] or [B,3,H,W], got {tuple(img.shape)}")
    return x.contiguous()


def distort_grid(img, specs, handle=None):
    """Every distortion of ``specs`` (``Distortion`` callables or (kind, factor) tuples) applied to ``img``
    (uint8 [3,H,W] or [B,3,H,W], CUDA) in one pass.  Returns uint8 [len(specs),3,H,W] ([B,len(specs),3,H,W]
    for a batch).  Asynchronous on the current stream."""
    x = _check(img)
    specs = [s.spec if hasattr(s, "spec") else tuple(s) for s in specs]
    if not 1 <= len(specs) <= MAX_OPS:
        raise ValueError(f"between 1 and {MAX_OPS} distortions per call, got {len(specs)}")
    ops = (Distortion * len(specs))(*[Distortion(int(k), 0, float(f)) for k, f in specs])
    b, _, hh, ww = x.shape
    npix = hh * ww
    out = torch.empty((b, len(specs), 3, hh, ww), dtype=torch.uint8, device=x.device)
    h = _handle_for(x, handle)
    src = _cabi.Batch(ctypes.c_void_p(x.data_ptr()), npix, 3 * npix, 0, b, _cabi.CT_U8, _cabi.CT_CHW, 0)
    dst = _cabi.Batch(ctypes.c_void_p(out.data_ptr()), npix, 3 * npix, 0, b * len(specs), _cabi.CT_U8, _cabi.CT_CHW, 0)
    rc = h.lib.ct_distort(h.h, ctypes.byref(src), ctypes.cast(ops, ctypes.c_void_p), len(specs), ctypes.byref(dst))
    if rc == _cabi.CT_E_INVALID:
        raise ValueError(h.lib.ct_last_error(h.h).decode())     # torchvision's argument errors
    h.check(rc)
    return out[0] if img.dim() == 3 else out


class GridDistortion:
    """One entry of the grid: callable like ``partial(F.adjust_*, factor)``; ``spec`` = (kind, factor)."""

    def __init__(self, kind, factor):
        self.spec = (kind, float(factor))

    def __call__(self, img):
        out = distort_grid(img, [self.spec])
        return out[0] if img.dim() == 3 else out[:, 0]

    def __repr__(self):
        return f"{_NAMES[self.spec[0]]}({self.spec[1]:g})"


def setup_grid_distortions(max_magnitude=0.5, num=6):
    """ref: utils/data.py:12-22 - the same 1 + 5 * num functions in the same order."""
    fns = [GridDistortion(IDENTITY, 0.0)]
    for magnitude in np.linspace(-max_magnitude, max_magnitude, num):
        fns.append(GridDistortion(BRIGHTNESS, 1 + magnitude))
        fns.append(GridDistortion(CONTRAST, 1 + magnitude))
        fns.append(GridDistortion(SATURATION, 1 + magnitude))
        fns.append(GridDistortion(HUE, magnitude))
        fns.append(GridDistortion(GAMMA, 1 + magnitude))
    return fns


def read_image(path, device="cuda"):
    """torchvision.io.read_image semantics (uint8 [3,H,W]) with the pixels landing on the device:
    decoded on the host (PIL), copied through pinned memory."""
    from PIL import Image
    with Image.open(path) as im:
        a = np.asarray(im.convert("RGB"))
    t = torch.from_numpy(np.ascontiguousarray(a.transpose(2, 0, 1)))
    return t.pin_memory().to(device, non_blocking=True)


class ArtificialTestDataset(torch.utils.data.Dataset):
    """ref: utils/data.py:88-106.  Item ``index`` = ground truth ``index // 31`` under distortion
    ``index % 31``; ``targets(i)`` gives all 31 targets of ground truth ``i`` from one pass.
    Tensors are uint8 on ``device`` unless ``as_float`` (then ``/ 255`` float32 like the reference)."""

    def __init__(self, image_dir, device="cuda", as_float=False):
        image_dir = Path(image_dir)
        self.gts = sorted(image_dir.glob("*_L.*"))
        self.references = sorted(image_dir.glob("*_R.*"))
        assert len(self.gts) == len(self.references)
        self.distortion_fns = setup_grid_distortions()
        self.device = device
        self.as_float = as_float

    def __len__(self):
        return len(self.gts) * len(self.distortion_fns)

    def _out(self, x):
        return x / 255 if self.as_float else x

    def targets(self, image_index):
        gt = read_image(str(self.gts[image_index]), self.device)
        reference = read_image(str(self.references[image_index]), self.device)
        return {"gt": self._out(gt), "reference": self._out(reference), "target": self._out(distort_grid(gt, self.distortion_fns))}

    def __getitem__(self, index):
        n = len(self.distortion_fns)
        gt = read_image(str(self.gts[index // n]), self.device)
        reference = read_image(str(self.references[index // n]), self.device)
        target = self.distortion_fns[index % n](gt)
        return {"gt": self._out(gt), "reference": self._out(reference), "target": self._out(target)}


class RealWorldTestDataset(torch.utils.data.Dataset):
    """ref: utils/data.py:109-126: (gt, distorted target, reference) triples read to the device."""

    def __init__(self, image_dir, device="cuda", as_float=False):
        image_dir = Path(image_dir)
        self.gts = sorted(image_dir.glob("*/*_L.*"))
        self.targets = sorted(image_dir.glob("*/*_LD.*"))
        self.references = sorted(image_dir.glob("*/*_R.*"))
        assert len(self.gts) == len(self.targets) == len(self.references)
        self.device = device
        self.as_float = as_float

    def __len__(self):
        return len(self.gts)

    def __getitem__(self, index):
        out = {"gt": read_image(str(self.gts[index]), self.device),
               "reference": read_image(str(self.references[index]), self.device),
               "target": read_image(str(self.targets[index]), self.device)}
        return {k: (v / 255 if self.as_float else v) for k, v in out.items()}
