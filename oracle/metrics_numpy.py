"""TEST INFRASTRUCTURE — CPU restatement (numpy, float64) of the quality metrics the reference
computes right after the hot path: the iCID of ref: utils/icid.py:28-152 and the
PSNR of ref: methods/__init__.py:35 (piq.psnr).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.

Pinning: ``icid`` is checked against the UNMODIFIED ref: utils/icid.py executed in the build
container (oracle/load_reference.py: real torch + torchvision, kornia.color.rgb_to_lab stubbed by
the restatement below because kornia is not installed) — tests/golden/metrics.npz holds those
values; the reference computes in float32, this restatement in float64, so they agree to ~1e-6,
not bit for bit.  PARITY UNPINNED for ``psnr`` and ``ssim``: piq is neither vendored, pinned nor
installed; the formulas below are piq.psnr's / piq.ssim's published definitions with their defaults
(data_range=1, reduction='mean', no greyscale; SSIM: 11x11 Gaussian of sigma 1.5, valid
convolution, k1=0.01, k2=0.03, average-pool downscale by max(1, round(min(H,W)/256))).
"""

import numpy as np

from . import skimage_color as _sk

_ICID_WEIGHTS = {
    # ref: utils/icid.py:42-49
    "perceptual": (0.002, 10, 10, 0.002, 0.002, 10, 10),
    "hue-preserving": (0.002, 10, 10, 0.002, 0.02, 10, 10),
    "chromatic": (0.002, 10, 10, 0.02, 0.02, 10, 10),
}


def bilinear_downscale(img, f):
    """torch.nn.functional.interpolate(img, scale_factor=1/f, mode='bilinear') for [.., H, W]
    (align_corners=False, no antialias): out = floor(size / f), src = (dst + 0.5) * f - 0.5."""
    h, w = img.shape[-2:]
    oh, ow = int(np.floor(h * (1.0 / f))), int(np.floor(w * (1.0 / f)))

    def taps(n_out, n_in):
        src = np.maximum((np.arange(n_out) + 0.5) * f - 0.5, 0.0)
        i0 = np.minimum(np.floor(src).astype(np.int64), n_in - 1)
        i1 = np.minimum(i0 + 1, n_in - 1)
        lam = src - i0
        return i0, i1, lam

    y0, y1, ly = taps(oh, h)
    x0, x1, lx = taps(ow, w)
    top = img[..., y0, :][..., :, x0] * (1 - lx) + img[..., y0, :][..., :, x1] * lx
    bot = img[..., y1, :][..., :, x0] * (1 - lx) + img[..., y1, :][..., :, x1] * lx
    return top * (1 - ly)[:, None] + bot * ly[:, None]


def rgb_to_lab_planar(img):
    """kornia.color.rgb_to_lab on [.., 3, H, W] (D65 / 2 degrees; the same constants as
    scikit-image's rgb2lab, oracle/skimage_color.py)."""
    hwc = np.moveaxis(np.asarray(img, dtype=np.float64), -3, -1)
    return np.moveaxis(_sk.rgb2lab(hwc), -1, -3)


def gaussian_blur_11(x, sigma=2.0, ksize=11):
    """torchvision.transforms.functional.gaussian_blur(x, [11, 11], [2, 2]) on [.., H, W]:
    normalised sampled Gaussian, reflect padding (no edge repeat), outer-product kernel."""
    half = (ksize - 1) * 0.5
    t = np.linspace(-half, half, ksize)
    k = np.exp(-0.5 * (t / sigma) ** 2)
    k /= k.sum()
    r = ksize // 2
    pad = [(0, 0)] * (x.ndim - 2) + [(r, r), (r, r)]
    p = np.pad(x, pad, mode="reflect")
    h, w = x.shape[-2:]
    tmp = sum(k[i] * p[..., :, i:i + w] for i in range(ksize))
    return sum(k[i] * tmp[..., i:i + h, :] for i in range(ksize))


def icid(img1, img2, intent="perceptual", omit_maps67=False, downsampling=True):
    """iCID of two image batches [B, 3, H, W] in [0, 1] (one number, like the reference: the
    mean runs over the batch too).  Follows ref: utils/icid.py:28-152 stage by stage."""
    if intent not in _ICID_WEIGHTS:
        raise ValueError("Intent should be either 'perceptual', 'hue-preserving', or 'chromatic'")
    w = _ICID_WEIGHTS[intent]
    a, b = np.asarray(img1, dtype=np.float64), np.asarray(img2, dtype=np.float64)
    if downsampling:                                                    # icid.py:60-65
        f = max(1, round(min(a.shape[-2:]) / 256))
        if f > 1:
            a, b = bilinear_downscale(a, f), bilinear_downscale(b, f)
    lab1, lab2 = rgb_to_lab_planar(a), rgb_to_lab_planar(b)            # icid.py:68-69
    L1, A1, B1 = lab1[..., 0, :, :], lab1[..., 1, :, :], lab1[..., 2, :, :]
    L2, A2, B2 = lab2[..., 0, :, :], lab2[..., 1, :, :], lab2[..., 2, :, :]
    C1, C2 = np.sqrt(A1 ** 2 + B1 ** 2), np.sqrt(A2 ** 2 + B2 ** 2)
    g = gaussian_blur_11
    muL1, muL2, muC1, muC2 = g(L1), g(L2), g(C1), g(C2)                # icid.py:89-92

    def spread(x, mu):                                                   # icid.py:95-107
        v = np.maximum(g(x ** 2) - mu ** 2, 0.0)
        return v, np.sqrt(v)

    vL1, sL1 = spread(L1, muL1)
    vL2, sL2 = spread(L2, muL2)
    vC1, sC1 = spread(C1, muC1)
    vC2, sC2 = spread(C2, muC2)
    dL = (muL1 - muL2) ** 2                                              # icid.py:110-116
    dC = (muC1 - muC2) ** 2
    hue = np.maximum((A1 - A2) ** 2 + (B1 - B2) ** 2 - (C1 - C2) ** 2, 0.0)
    dH = g(np.sqrt(hue)) ** 2
    sL12 = g(L1 * L2) - muL1 * muL2
    sC12 = g(C1 * C2) - muC1 * muC2
    maps = [1 / (w[0] * dL + 1),                                         # icid.py:119-140
            (w[1] + 2 * sL1 * sL2) / (w[1] + vL1 + vL2),
            (w[2] + np.abs(sL12)) / (w[2] + sL1 * sL2),
            1 / (w[3] * dC + 1),
            1 / (w[4] * dH + 1),
            (w[5] + 2 * sC1 * sC2) / (w[5] + sC1 ** 2 + sC2 ** 2),
            (w[6] + np.abs(sC12)) / (w[6] + sC1 * sC2)]
    expo = [1, 1, 3, 1, 1, 0, 0] if omit_maps67 else [1, 1, 3, 1, 1, 1, 1]   # icid.py:51-54, alpha = 3
    prod = np.ones_like(maps[0])
    for m, e in zip(maps, expo):
        prod = prod * m ** e
    return float(1.0 - prod.mean())                                      # icid.py:146


def psnr(x, y, data_range=1.0):
    """piq.psnr(x, y) with its defaults on [B, 3, H, W]: per image -10 log10(mse + 1e-8) of the
    images scaled by data_range, averaged over the batch."""
    a, b = np.asarray(x, dtype=np.float64) / data_range, np.asarray(y, dtype=np.float64) / data_range
    mse = ((a - b) ** 2).reshape(a.shape[0], -1).mean(axis=1)
    return float(np.mean(-10.0 * np.log10(mse + 1e-8)))


def ssim(x, y, kernel_size=11, kernel_sigma=1.5, k1=0.01, k2=0.03, downsample=True):
    """piq.ssim(x, y) with its defaults on [B, 3, H, W] in [0, 1] (Wang et al. 2004): per channel
    Gaussian-weighted local means / variances / covariance by VALID convolution, the mean of the
    SSIM map per channel, then the mean over channels and over the batch."""
    a, b = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    f = max(1, round(min(a.shape[-2:]) / 256))
    if f > 1 and downsample:                      # torch.nn.functional.avg_pool2d(kernel_size=f)
        oh, ow = a.shape[-2] // f, a.shape[-1] // f
        a = a[..., :oh * f, :ow * f].reshape(*a.shape[:-2], oh, f, ow, f).mean(axis=(-3, -1))
        b = b[..., :oh * f, :ow * f].reshape(*b.shape[:-2], oh, f, ow, f).mean(axis=(-3, -1))
    t = np.arange(kernel_size) - (kernel_size - 1) / 2.0
    g = np.exp(-(t[:, None] ** 2 + t[None, :] ** 2) / (2.0 * kernel_sigma ** 2))
    g /= g.sum()
    k1d = g.sum(axis=0)                            # the kernel is separable: rows of g sum to it
    k1d_col = g.sum(axis=1)
    h, w = a.shape[-2] - kernel_size + 1, a.shape[-1] - kernel_size + 1

    def blur(v):
        tmp = sum(k1d[i] * v[..., :, i:i + w] for i in range(kernel_size))
        return sum(k1d_col[i] * tmp[..., i:i + h, :] for i in range(kernel_size))

    c1, c2 = k1 ** 2, k2 ** 2
    mu_x, mu_y = blur(a), blur(b)
    s_xx, s_yy, s_xy = blur(a * a) - mu_x ** 2, blur(b * b) - mu_y ** 2, blur(a * b) - mu_x * mu_y
    cs = (2 * s_xy + c2) / (s_xx + s_yy + c2)
    ss = (2 * mu_x * mu_y + c1) / (mu_x ** 2 + mu_y ** 2 + c1) * cs
    return float(ss.mean(axis=(-1, -2)).mean(axis=1).mean())
