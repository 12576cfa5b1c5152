"""TEST INFRASTRUCTURE - CPU restatement of the artificial-distortion generator.

The reference builds its test-set distortions from torchvision.transforms.functional.adjust_*
(ref: utils/data.py:12-22) applied to uint8 CHW tensors (ref: utils/data.py:100-104).  torchvision is a
third-party dependency of the reference (not vendored under /root/reference); the version in this image
is 0.26.0.  This file restates the uint8 branch of its tensor backend
(torchvision/transforms/_functional_tensor.py: _blend, rgb_to_grayscale, adjust_brightness / contrast /
saturation / hue / gamma, _rgb2hsv, _hsv2rgb, convert_image_dtype) in numpy float32, one numpy call per
torch call.

Pinned by tests/test_oracle_cpu.py against torchvision itself where it is importable (this image) and
against tests/golden/distort_grid.npz (generated from torchvision by oracle/gen_golden_distort.py).
Two steps are not bit-reproducible across implementations and are pinned to "at most one level":
the float32 mean of the contrast blend (summation order) and the float32 power of the gamma curve.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""

import numpy as np

F32 = np.float32
IDENTITY, BRIGHTNESS, CONTRAST, SATURATION, HUE, GAMMA = range(6)


def grid_specs(max_magnitude=0.5, num=6):
    """(kind, factor) of ref: utils/data.py:12-22 in order."""
    specs = [(IDENTITY, 0.0)]
    for magnitude in np.linspace(-max_magnitude, max_magnitude, num):
        specs += [(BRIGHTNESS, 1 + magnitude), (CONTRAST, 1 + magnitude), (SATURATION, 1 + magnitude),
                  (HUE, magnitude), (GAMMA, 1 + magnitude)]
    return specs


def _gray(img):
    r, g, b = (img[c].astype(F32) for c in range(3))
    l_img = F32(0.2989) * r + F32(0.587) * g + F32(0.114) * b
    return l_img.astype(np.uint8)       # .to(uint8): truncation


def _blend(img1, img2, ratio):
    ratio = float(ratio)
    v = F32(ratio) * img1.astype(F32) + F32(1.0 - ratio) * np.asarray(img2, dtype=F32)
    return np.clip(v, F32(0), F32(255)).astype(np.uint8)


def _to_unit(img):
    return img.astype(F32) / F32(255.0)


def _to_u8(x):
    return (x * F32(255 + 1.0 - 1e-3)).astype(np.uint8)


def adjust_brightness(img, factor):
    if factor < 0:
        raise ValueError(f"brightness_factor ({factor}) is not non-negative.")
    return _blend(img, np.zeros_like(img), factor)


def adjust_contrast(img, factor):
    if factor < 0:
        raise ValueError(f"contrast_factor ({factor}) is not non-negative.")
    g = _gray(img)
    mean = F32(F32(np.sum(g, dtype=np.int64)) / F32(g.size))     # float32 sum (as if exact, rounded once), / n in float32
    return _blend(img, mean, factor)


def adjust_saturation(img, factor):
    if factor < 0:
        raise ValueError(f"saturation_factor ({factor}) is not non-negative.")
    return _blend(img, _gray(img)[None], factor)


def _rgb2hsv(img):
    r, g, b = img[0], img[1], img[2]
    maxc = img.max(axis=0)
    minc = img.min(axis=0)
    eqc = maxc == minc
    cr = maxc - minc
    ones = np.ones_like(maxc)
    s = cr / np.where(eqc, ones, maxc)
    cr_divisor = np.where(eqc, ones, cr)
    rc = (maxc - r) / cr_divisor
    gc = (maxc - g) / cr_divisor
    bc = (maxc - b) / cr_divisor
    hr = (maxc == r).astype(F32) * (bc - gc)
    hg = ((maxc == g) & (maxc != r)).astype(F32) * (F32(2.0) + rc - bc)
    hb = ((maxc != g) & (maxc != r)).astype(F32) * (F32(4.0) + gc - rc)
    h = hr + hg + hb
    h = np.fmod(h / F32(6.0) + F32(1.0), F32(1.0))
    return h, s, maxc


def _hsv2rgb(h, s, v):
    h6 = h * F32(6.0)
    i = np.floor(h6)
    f = h6 - i
    i = i.astype(np.int32)
    one = F32(1.0)
    p = np.clip(v * (one - s), F32(0), one)
    q = np.clip(v * (one - s * f), F32(0), one)
    t = np.clip(v * (one - s * (one - f)), F32(0), one)
    i = i % 6
    a1 = np.stack((v, q, p, p, t, v))
    a2 = np.stack((t, v, v, q, p, p))
    a3 = np.stack((p, p, t, v, v, q))
    pick = lambda a: np.take_along_axis(a, i[None], axis=0)[0]   # the einsum with the one-hot mask
    return np.stack((pick(a1), pick(a2), pick(a3)))


def adjust_hue(img, factor):
    if not (-0.5 <= factor <= 0.5):
        raise ValueError(f"hue_factor ({factor}) is not in [-0.5, 0.5].")
    h, s, v = _rgb2hsv(_to_unit(img))
    m = np.fmod(h + F32(factor), F32(1.0))          # torch.remainder: fmod, then the divisor's sign
    m = np.where((m != 0) & (m < 0), m + F32(1.0), m).astype(F32)
    return _to_u8(_hsv2rgb(m, s, v))


def adjust_gamma(img, gamma, gain=1):
    if gamma < 0:
        raise ValueError("Gamma should be a non-negative real number")
    x = _to_unit(img)
    if gamma == 0.5:
        y = np.sqrt(x)                               # torch.pow's exact special cases
    elif gamma == 2.0:
        y = x * x
    elif gamma == 3.0:
        y = x * x * x
    else:
        y = np.power(x.astype(np.float64), float(F32(gamma))).astype(F32)
    return _to_u8(np.clip(F32(gain) * y, F32(0), F32(1)))


_FUNCS = {BRIGHTNESS: adjust_brightness, CONTRAST: adjust_contrast, SATURATION: adjust_saturation,
          HUE: adjust_hue, GAMMA: adjust_gamma}


def distort(img, kind, factor):
    """img: uint8 [3,H,W]."""
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[0] == 3
    if kind == IDENTITY:
        return img.copy()
    return _FUNCS[kind](img, factor)


def distort_grid(img, specs=None):
    specs = grid_specs() if specs is None else specs
    return np.stack([distort(img, k, f) for k, f in specs])
