"""TEST INFRASTRUCTURE — CPU restatement of scikit-image's sRGB<->CIE-Lab conversion.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (color-transfer_b200/) never does.

Why a restatement: the reference's Reinhard transfer calls
``skimage.color.rgb2lab`` / ``lab2rgb`` (ref: methods/linear.py:5,25,26,40).
scikit-image is a third-party dependency that is NOT vendored under
/root/reference, is listed un-pinned in ref: requirements.txt:2, and is not
installed in this image (no network).  The arithmetic below restates the
published algorithm of ``skimage/color/colorconv.py`` (scikit-image >= 0.19:
``rgb2xyz``, ``xyz2lab``, ``_lab2xyz``, ``xyz2rgb``; D65 illuminant, 2 degree
observer; float32 stays float32, float64 stays float64).

PARITY UNPINNED for this file: the reference repository holds no golden vector
for the Lab conversion and the real scikit-image cannot be executed here, so
these functions are pinned only by (a) round-trip identity, (b) agreement with
OpenCV's approximate ``COLOR_RGB2Lab`` (tests/test_oracle_cpu.py) and (c) the
CIE definitions.  If scikit-image is importable on a box, ``have_real_skimage``
is True and tests compare this restatement with it.
"""

import numpy as np

# sRGB (D65) primaries -> CIE XYZ, as tabulated in skimage.color.colorconv.xyz_from_rgb
XYZ_FROM_RGB = np.array(
    [[0.412453, 0.357580, 0.180423],
     [0.212671, 0.715160, 0.072169],
     [0.019334, 0.119193, 0.950227]], dtype=np.float64)
# skimage computes the inverse once at import with scipy.linalg.inv
RGB_FROM_XYZ = np.linalg.inv(XYZ_FROM_RGB)
# xyz_tristimulus_values(illuminant="D65", observer="2")
WHITE_D65_2 = np.array([0.95047, 1.0, 1.08883], dtype=np.float64)

try:  # pragma: no cover - not installed in the build image
    import skimage.color as _real
    have_real_skimage = True
except Exception:  # noqa: BLE001
    _real = None
    have_real_skimage = False


def _as_float(img):
    img = np.asarray(img)
    if img.dtype == np.float32:
        return img.astype(np.float32, copy=True)
    return img.astype(np.float64, copy=True)


def rgb2lab(rgb):
    """sRGB in [0,1] -> CIE-Lab.  Restates rgb2xyz + xyz2lab."""
    c = _as_float(rgb)
    dt = c.dtype
    hi = c > 0.04045
    c[hi] = np.power((c[hi] + 0.055) / 1.055, 2.4)
    c[~hi] /= 12.92
    xyz = c @ XYZ_FROM_RGB.T.astype(dt)
    xyz = xyz / WHITE_D65_2.astype(dt)
    big = xyz > 0.008856
    xyz[big] = np.cbrt(xyz[big])
    xyz[~big] = 7.787 * xyz[~big] + 16.0 / 116.0
    fx, fy, fz = xyz[..., 0], xyz[..., 1], xyz[..., 2]
    L = 116.0 * fy - 16.0
    a = 500.0 * (fx - fy)
    b = 200.0 * (fy - fz)
    return np.stack([L, a, b], axis=-1).astype(dt, copy=False)


def lab2rgb(lab):
    """CIE-Lab -> sRGB clipped to [0,1].  Restates _lab2xyz + xyz2rgb."""
    lab = _as_float(lab)
    dt = lab.dtype
    L, a, b = lab[..., 0], lab[..., 1], lab[..., 2]
    fy = (L + 16.0) / 116.0
    fx = a / 500.0 + fy
    fz = fy - b / 200.0
    fz = np.where(fz < 0, 0, fz)          # skimage zeroes invalid z (and warns)
    f = np.stack([fx, fy, fz], axis=-1)
    big = f > 0.2068966
    f[big] = np.power(f[big], 3.0)
    f[~big] = (f[~big] - 16.0 / 116.0) / 7.787
    xyz = f * WHITE_D65_2.astype(dt)
    rgb = xyz @ RGB_FROM_XYZ.T.astype(dt)
    hi = rgb > 0.0031308
    rgb[hi] = 1.055 * np.power(rgb[hi], 1.0 / 2.4) - 0.055
    rgb[~hi] *= 12.92
    np.clip(rgb, 0, 1, out=rgb)
    return rgb.astype(dt, copy=False)
