import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pair0964():
    from PIL import Image
    left = np.asarray(Image.open(os.path.join(GOLDEN, "0964_L.png")).convert("RGB")) / 255.0
    right = np.asarray(Image.open(os.path.join(GOLDEN, "0964_R.png")).convert("RGB")) / 255.0
    return left, right


@pytest.fixture(scope="session")
def golden():
    return {name: np.load(os.path.join(GOLDEN, name + ".npz")) for name in ("small_f64", "small_f32", "pair0964")}


def synthetic_pair(h, w, seed, dtype=np.float64, ref_shape=None):
    """Smooth-ish uint8-quantised stereo pair in [0,1] (same generator family as oracle/gen_golden.py)."""
    rng = np.random.default_rng(seed)

    def img(hh, ww):
        yy, xx = np.mgrid[0:hh, 0:ww]
        base = np.zeros((hh, ww, 3))
        for c in range(3):
            f, g, p = rng.uniform(0, 3), rng.uniform(0, 3), rng.uniform(0, 2 * np.pi)
            base[..., c] = 128 + 90 * np.sin(2 * np.pi * (f * xx / ww + g * yy / hh) + p) + 12 * rng.standard_normal((hh, ww))
        return np.clip(base, 0, 255).astype(np.uint8)

    ref = img(*(ref_shape or (h, w)))
    src = img(h, w) if ref_shape else ref
    gain, gamma = rng.uniform(0.7, 1.3, 3), rng.uniform(0.7, 1.3, 3)
    tgt = np.clip(255 * gain * (src / 255.0) ** gamma, 0, 255).astype(np.uint8)
    return (tgt / 255.0).astype(dtype), (ref / 255.0).astype(dtype)


def u8_identical_fraction(a, b):
    qa = np.rint(np.clip(a, 0, 1) * 255).astype(np.uint8)
    qb = np.rint(np.clip(b, 0, 1) * 255).astype(np.uint8)
    return float(np.mean(qa == qb))
