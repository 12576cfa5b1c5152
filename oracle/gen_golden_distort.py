"""TEST INFRASTRUCTURE - generate tests/golden/distort_grid.npz from torchvision itself.

    python -m oracle.gen_golden_distort

Runs the reference's own `setup_grid_distortions` recipe (ref: utils/data.py:12-22: functools.partial
objects over torchvision.transforms.functional.adjust_*) on seeded uint8 images on the CPU and stores the
31 outputs per image, together with torchvision's version.  The oracle restatement
(oracle/distort_numpy.py) is required to agree with every output before the file is written
(identical except the contrast / gamma entries, which may differ by one level).
"""

import os
from functools import partial

import numpy as np

from . import distort_numpy as oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "..", "tests", "golden")


def torchvision_grid(max_magnitude=0.5, num=6):
    import torchvision.transforms.functional as F
    fns = [lambda x: x]
    for magnitude in np.linspace(-max_magnitude, max_magnitude, num):
        fns.append(partial(F.adjust_brightness, brightness_factor=1 + magnitude))
        fns.append(partial(F.adjust_contrast, contrast_factor=1 + magnitude))
        fns.append(partial(F.adjust_saturation, saturation_factor=1 + magnitude))
        fns.append(partial(F.adjust_hue, hue_factor=magnitude))
        fns.append(partial(F.adjust_gamma, gamma=1 + magnitude))
    return fns


def images():
    """Seeded uint8 [3,H,W] images: smooth field + noise, uniform noise with saturated / gray / black
    patches (hue's maxc == minc branch, clamps), and every gray level once."""
    rng = np.random.default_rng(20240607)
    out = {}
    yy, xx = np.mgrid[0:61, 0:83]
    base = np.stack([0.5 + 0.4 * np.sin(xx / 9.0 + c) * np.cos(yy / 7.0 - c) for c in range(3)])
    out["smooth"] = np.clip((base + rng.normal(0, 0.05, base.shape)) * 255, 0, 255).astype(np.uint8)
    noise = rng.integers(0, 256, (3, 48, 64), dtype=np.uint8)
    noise[:, :6, :6] = 255
    noise[:, 6:12, :6] = 0
    noise[:, 12:18, :6] = 128
    noise[0, 18:24, :6] = 255
    noise[1:, 18:24, :6] = 0
    out["noise"] = noise
    ramp = np.arange(256, dtype=np.uint8).reshape(16, 16)
    out["ramp"] = np.stack([ramp, ramp[::-1], ramp.T])
    return out


def compare(kind, a, b):
    """a, b uint8: (identical fraction, max abs difference)."""
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    return float((d == 0).mean()), int(d.max())


def main():
    import torch
    import torchvision
    specs = oracle.grid_specs()
    fns = torchvision_grid()
    assert len(specs) == len(fns) == 31
    store = {"torchvision_version": np.array(torchvision.__version__), "kinds": np.array([k for k, _ in specs]),
             "factors": np.array([f for _, f in specs])}
    for name, img in images().items():
        want = np.stack([fn(torch.from_numpy(img)).numpy() for fn in fns])
        got = oracle.distort_grid(img, specs)
        for k, (kind, factor) in enumerate(specs):
            same, worst = compare(kind, want[k], got[k])
            loose = kind in (oracle.CONTRAST, oracle.GAMMA)
            assert (worst <= 1 and same >= 0.99) if loose else worst == 0, (name, k, kind, factor, same, worst)
            if worst:
                print(f"{name}: distortion {k} (kind {kind}, factor {factor:g}): {100 * (1 - same):.3f} % of the values differ by one level")
        store[f"{name}_image"] = img
        store[f"{name}_grid"] = want
    path = os.path.join(GOLDEN, "distort_grid.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
