"""TEST INFRASTRUCTURE - restatement of ``skimage.transform.resize`` as the reference's regrain
uses it (ref: methods/iterative.py:68-72: ``resize(img, (h2, w2))`` with every default).

scikit-image is not installed here and is un-pinned in the reference (requirements.txt:2), so this
follows the published implementation of scikit-image >= 0.19 (``skimage/transform/_warps.py``):
defaults order=1, mode='reflect' (-> scipy.ndimage 'mirror'), anti_aliasing=True when any output
dimension shrinks with sigma = max(0, (factor - 1) / 2) per axis, then
``scipy.ndimage.zoom(filtered, 1/factors, order=1, mode='mirror', grid_mode=True)`` and a clip to
the input's [min, max].  The heavy lifting is the REAL scipy.ndimage (installed, so that part is
pinned by scipy itself); only this wrapper logic is restated.  PARITY UNPINNED for the wrapper.
"""

import numpy as np
import scipy.ndimage as ndi

try:  # pragma: no cover
    from skimage.transform import resize as _real_resize
    have_real_skimage = True
except Exception:  # noqa: BLE001
    _real_resize = None
    have_real_skimage = False


def resize(image, output_shape):
    image = np.asarray(image)
    if image.dtype not in (np.float32, np.float64):
        image = image.astype(np.float64)
    output_shape = tuple(output_shape)
    if len(output_shape) < image.ndim:                      # channels are kept
        output_shape = output_shape + image.shape[len(output_shape):]
    factors = np.divide(image.shape, output_shape)
    anti_aliasing = any(o < i for o, i in zip(output_shape, image.shape))
    if anti_aliasing:
        sigma = np.maximum(0, (factors - 1) / 2)
        filtered = ndi.gaussian_filter(image, sigma, cval=0, mode="mirror")
    else:
        filtered = image
    out = ndi.zoom(filtered, [1 / f for f in factors], order=1, mode="mirror", cval=0, grid_mode=True)
    np.clip(out, image.min(), image.max(), out=out)
    return out
