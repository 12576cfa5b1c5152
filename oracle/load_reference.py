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
