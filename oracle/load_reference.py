"""TEST INFRASTRUCTURE — load the UNMODIFIED reference hot path by file path.

``/root/reference`` exists only in the build container, never on the GPU box, so this
loader is used by ``oracle/gen_golden.py`` (fixture generation) and by CPU tests that
skip when the directory is absent.  Nothing under color-transfer_b200/ imports it.

``import methods.linear`` on the reference fails here (ref: methods/__init__.py:3-7 pulls
pytorch_lightning / piq / kornia) and loading the two files directly fails on
``import skimage`` (ref: methods/linear.py:5, methods/iterative.py:5).  We therefore put
a stub ``skimage`` on ``sys.modules`` for the duration of the load:

* ``skimage.color.rgb2lab / lab2rgb`` -> oracle/skimage_color.py (restatement, so the
  Reinhard function is "reference data flow + restated Lab"; Xiao / MKL / IDT do not touch
  skimage at all and run as pure reference code);
* ``skimage.transform.resize`` -> oracle/skimage_resize.py (restated wrapper over the real
  scipy.ndimage), used by the regrain of ``automated_color_grading``.
"""

import importlib.util
import os
import sys
import types

from . import skimage_color

REFERENCE_ROOT = os.environ.get("CT_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "methods", "linear.py"))


def _stub_skimage():
    if skimage_color.have_real_skimage:
        return {}
    pkg = types.ModuleType("skimage")
    color = types.ModuleType("skimage.color")
    color.rgb2lab = skimage_color.rgb2lab
    color.lab2rgb = skimage_color.lab2rgb
    transform = types.ModuleType("skimage.transform")

    from . import skimage_resize
    transform.resize = skimage_resize.resize
    pkg.color, pkg.transform = color, transform
    return {"skimage": pkg, "skimage.color": color, "skimage.transform": transform}


def _load(name, relpath):
    stubs = _stub_skimage()
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, relpath))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def linear():
    """The reference's methods/linear.py as a module object."""
    return _load("_ct_reference_linear", os.path.join("methods", "linear.py"))


def iterative():
    """The reference's methods/iterative.py as a module object."""
    return _load("_ct_reference_iterative", os.path.join("methods", "iterative.py"))


def icid():
    """The reference's utils/icid.py ``icid`` function.  kornia is not installed, so
    ``kornia.color.rgb_to_lab`` is stubbed by a torch restatement of kornia's published
    algorithm (the same constants as scikit-image's rgb2lab); torch and torchvision are real."""
    import torch

    def rgb_to_lab(image):
        lin = torch.where(image > 0.04045, torch.pow((image + 0.055) / 1.055, 2.4), image / 12.92)
        r, g, b = lin[..., 0, :, :], lin[..., 1, :, :], lin[..., 2, :, :]
        x = (0.412453 * r + 0.357580 * g + 0.180423 * b) / 0.95047
        y = (0.212671 * r + 0.715160 * g + 0.072169 * b) / 1.0
        z = (0.019334 * r + 0.119193 * g + 0.950227 * b) / 1.08883
        xyz = torch.stack([x, y, z], dim=-3)
        f = torch.where(xyz > 0.008856, torch.pow(xyz.clamp(min=0.008856), 1.0 / 3.0), 7.787 * xyz + 4.0 / 29.0)
        fx, fy, fz = f[..., 0, :, :], f[..., 1, :, :], f[..., 2, :, :]
        return torch.stack([116.0 * fy - 16.0, 500.0 * (fx - fy), 200.0 * (fy - fz)], dim=-3)

    kornia = types.ModuleType("kornia")
    color = types.ModuleType("kornia.color")
    color.rgb_to_lab = rgb_to_lab
    kornia.color = color
    saved = {k: sys.modules.get(k) for k in ("kornia", "kornia.color")}
    sys.modules.update({"kornia": kornia, "kornia.color": color})
    try:
        spec = importlib.util.spec_from_file_location("_ct_reference_icid", os.path.join(REFERENCE_ROOT, "utils", "icid.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod.icid
