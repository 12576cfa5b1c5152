"""CPU tests: the oracle against the committed golden vectors (generated from the unmodified
reference by oracle/gen_golden.py), against the reference itself when /root/reference is
present, and sanity of the restated scikit-image Lab conversion."""

import hashlib
import os

import numpy as np
import pytest

from oracle import load_reference, skimage_color
from oracle import reference_numpy as oracle

IDT_SEED = 42


@pytest.mark.parametrize("case", ["small_f64", "small_f32"])
def test_oracle_matches_golden_small(golden, case):
    g = golden[case]
    t, r = g["target"], g["reference"]
    assert np.array_equal(oracle.color_transfer_in_correlated_color_space(t, r), g["ccs"])
    for dec in ("MK", "sqrt", "cholesky"):
        assert np.array_equal(oracle.monge_kantorovitch_color_transfer(t, r, dec), g["mkl_" + dec])
        assert np.array_equal(oracle.mkl_matrix(g["cov_t"], g["cov_r"], dec), g["T_" + dec])
    out = oracle.color_transfer_between_images(t, r)
    assert out.dtype == g["reinhard"].dtype == t.dtype
    assert np.array_equal(out, g["reinhard"])
    np.random.seed(IDT_SEED)
    out, traces = oracle.idt_instrumented(t, r)
    assert out.dtype == np.float64
    assert np.array_equal(out, g["idt"])
    assert np.array_equal(np.stack([x["rot"] for x in traces]), g["idt_rot"])
    assert np.array_equal(np.stack([x["counts_t"] for x in traces]), g["idt_counts_t"])
    assert np.array_equal(np.stack([x["counts_r"] for x in traces]), g["idt_counts_r"])
    assert np.array_equal(np.stack([x["lut"] for x in traces]), g["idt_lut"])
    assert np.array_equal(np.stack([x["lo"] for x in traces]), g["idt_lo"])
    np.random.seed(IDT_SEED + 1)
    assert np.array_equal(oracle.iterative_distribution_transfer(t, r, 64, 2), g["idt_b64_n2"])
    np.random.seed(IDT_SEED)
    assert np.array_equal(oracle.automated_color_grading(t, r), g["acg"])        # IDT + regrain


def test_oracle_matches_golden_pair0964(golden, pair0964):
    g = golden["pair0964"]
    left, right = pair0964
    mu_t, cov_t = oracle.mean_and_cov(left)
    assert np.array_equal(cov_t, g["cov_t"]) and np.array_equal(mu_t, g["mean_t"])
    # the survey's smoke values
    np.testing.assert_allclose(cov_t[0], [0.02677205, 0.02357668, 0.00893746], atol=1e-8)
    np.testing.assert_allclose(g["T_MK"][0], [0.98420107, -0.00395838, 0.00630462], atol=1e-8)
    np.testing.assert_allclose(g["lab_mean_t"], [48.14974039, -2.70486872, 17.63913478], atol=1e-7)
    np.testing.assert_allclose(g["lab_std_r"], [15.69873659, 4.07359346, 22.39241709], atol=1e-7)
    for name, fn in (("ccs", oracle.color_transfer_in_correlated_color_space),
                     ("mkl_MK", oracle.monge_kantorovitch_color_transfer),
                     ("reinhard", oracle.color_transfer_between_images)):
        out = fn(left, right)
        assert np.array_equal(out.reshape(-1)[::997], g[name + "_sample"])
        digest = hashlib.sha256(np.rint(np.clip(out, 0, 1) * 255).astype(np.uint8).tobytes()).digest()
        assert np.array_equal(np.frombuffer(digest, dtype=np.uint8), g[name + "_u8_sha256"])
    np.random.seed(IDT_SEED)
    out, traces = oracle.idt_instrumented(left, right, keep_arrays=False)
    assert np.array_equal(out.reshape(-1)[::997], g["idt_sample"])
    assert np.array_equal(np.stack([x["counts_t"] for x in traces]), g["idt_counts_t"])
    np.testing.assert_array_equal(np.array([out.min(), out.max(), out.mean()]), g["idt_stats"])


@pytest.mark.skipif(not load_reference.available(), reason="/root/reference only exists in the build container")
def test_oracle_equals_unmodified_reference():
    ref_lin, ref_it = load_reference.linear(), load_reference.iterative()
    rng = np.random.default_rng(3)
    for dtype in (np.float64, np.float32):
        t = rng.random((19, 23, 3)).astype(dtype)
        r = rng.random((17, 31, 3)).astype(dtype)
        assert np.array_equal(ref_lin.color_transfer_in_correlated_color_space(t, r),
                              oracle.color_transfer_in_correlated_color_space(t, r))
        for dec in ("MK", "sqrt", "cholesky"):
            assert np.array_equal(ref_lin.monge_kantorovitch_color_transfer(t, r, decomposition=dec),
                                  oracle.monge_kantorovitch_color_transfer(t, r, dec))
        np.random.seed(11)
        a = ref_it.iterative_distribution_transfer(t, r, bins=100, n_iter=3)
        np.random.seed(11)
        assert np.array_equal(a, oracle.iterative_distribution_transfer(t, r, 100, 3))
    with pytest.raises(ValueError, match="Unknown decomposition"):
        ref_lin.monge_kantorovitch_color_transfer(t, r, decomposition="qr")
    with pytest.raises(ValueError, match="Unknown decomposition"):
        oracle.monge_kantorovitch_color_transfer(t, r, "qr")


def test_bin_index_restates_np_histogram():
    rng = np.random.default_rng(5)
    for n, bins in ((1000, 255), (5000, 7), (300, 1)):
        x = rng.normal(size=n)
        x[:3] = [x.min(), x.max(), x.max()]
        counts, edges = np.histogram(x, bins=bins, range=[x.min(), x.max()])
        k = oracle.bin_index(x, edges)
        assert np.array_equal(np.bincount(k, minlength=bins), counts)
        assert np.all((edges[k] <= x) & ((x < edges[k + 1]) | (k == bins - 1)))


# ------------------------------------------------------------------ restated scikit-image Lab
def test_lab_known_values_and_round_trip():
    white = skimage_color.rgb2lab(np.ones((1, 1, 3)))
    np.testing.assert_allclose(white[0, 0], [100.0, 0.0, 0.0], atol=5e-3)
    black = skimage_color.rgb2lab(np.zeros((1, 1, 3)))
    np.testing.assert_allclose(black[0, 0], [0.0, 0.0, 0.0], atol=1e-12)
    red = skimage_color.rgb2lab(np.array([[[1.0, 0.0, 0.0]]]))
    np.testing.assert_allclose(red[0, 0], [53.24, 80.09, 67.20], atol=0.02)     # CIE values of sRGB red
    rng = np.random.default_rng(0)
    rgb = rng.random((64, 64, 3))
    back = skimage_color.lab2rgb(skimage_color.rgb2lab(rgb))
    assert np.max(np.abs(back - rgb)) < 1e-12
    assert skimage_color.rgb2lab(rgb.astype(np.float32)).dtype == np.float32
    out = skimage_color.lab2rgb(np.array([[[150.0, 200.0, -300.0]]]))             # out of gamut: clipped
    assert out.min() >= 0.0 and out.max() <= 1.0


def test_lab_agrees_with_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    rgb = rng.random((32, 32, 3)).astype(np.float32)
    ours = skimage_color.rgb2lab(rgb.astype(np.float64))
    theirs = cv2.cvtColor(rgb, cv2.COLOR_RGB2Lab)       # OpenCV's own (approximate) implementation
    assert np.max(np.abs(ours - theirs)) < 0.5


def test_resize_restatement_shapes_and_identities():
    """oracle/skimage_resize.py: defaults of skimage.transform.resize over the real scipy.ndimage."""
    from oracle import skimage_resize
    rng = np.random.default_rng(4)
    img = rng.random((45, 62, 3))
    small = skimage_resize.resize(img, (23, 31))
    assert small.shape == (23, 31, 3) and small.dtype == np.float64
    assert img.min() <= small.min() and small.max() <= img.max()              # clip=True
    assert np.array_equal(skimage_resize.resize(img, (45, 62)), img)          # same size: identity
    flat = np.full((40, 40, 3), 0.37)
    assert np.allclose(skimage_resize.resize(flat, (20, 20)), 0.37, atol=1e-15)
    up = skimage_resize.resize(small, (45, 62))
    assert up.shape == img.shape and np.abs(up - img).mean() < 0.3
    ramp = np.repeat(np.linspace(0, 1, 64)[None, :, None], 8, axis=0).repeat(3, axis=2)
    half = skimage_resize.resize(ramp, (8, 32))                               # linear ramps survive away from the border
    assert np.allclose(half[:, 4:-4, 0], (ramp[:, 8:-8:2, 0] + ramp[:, 9:-7:2, 0]) / 2, atol=1e-12)


@pytest.mark.skipif(not skimage_color.have_real_skimage, reason="scikit-image is not installed")
def test_resize_restatement_equals_scikit_image():
    from skimage.transform import resize

    from oracle import skimage_resize
    rng = np.random.default_rng(5)
    img = rng.random((45, 62, 3))
    assert np.max(np.abs(resize(img, (23, 31)) - skimage_resize.resize(img, (23, 31)))) < 1e-12
    assert np.max(np.abs(resize(img[:23, :31], (45, 62)) - skimage_resize.resize(img[:23, :31], (45, 62)))) < 1e-12


@pytest.mark.skipif(not skimage_color.have_real_skimage, reason="scikit-image is not installed")
def test_lab_restatement_equals_scikit_image():
    from skimage.color import lab2rgb, rgb2lab
    rng = np.random.default_rng(2)
    rgb = rng.random((48, 48, 3))
    assert np.max(np.abs(rgb2lab(rgb) - skimage_color.rgb2lab(rgb))) < 1e-12
    lab = rgb2lab(rgb)
    assert np.max(np.abs(lab2rgb(lab) - skimage_color.lab2rgb(lab))) < 1e-12


# ---- artificial-distortion generator ------------------------------------------
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _distort_close(kind, a, b):
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    if kind in (2, 5):      # contrast, gamma: at most one level on at most 1 % of the values
        return d.max() <= 1 and (d == 0).mean() >= 0.99
    return d.max() == 0


def test_distort_oracle_equals_torchvision_golden():
    """oracle/distort_numpy.py against the outputs torchvision produced (oracle/gen_golden_distort.py)"""
    from oracle import distort_numpy
    g = np.load(os.path.join(GOLDEN, "distort_grid.npz"))
    specs = distort_numpy.grid_specs()
    assert [k for k, _ in specs] == list(g["kinds"])
    np.testing.assert_array_equal([f for _, f in specs], g["factors"])
    for name in ("smooth", "noise", "ramp"):
        got = distort_numpy.distort_grid(g[f"{name}_image"], specs)
        for k, (kind, factor) in enumerate(specs):
            assert _distort_close(kind, got[k], g[f"{name}_grid"][k]), (name, k, kind, factor)


def test_distort_oracle_equals_torchvision_live():
    """the same on fresh random images where torchvision is importable (it is in the build image)"""
    tvf = pytest.importorskip("torchvision.transforms.functional")
    import torch
    from oracle import distort_numpy
    names = {1: "adjust_brightness", 2: "adjust_contrast", 3: "adjust_saturation", 4: "adjust_hue", 5: "adjust_gamma"}
    rng = np.random.default_rng(99)
    img = rng.integers(0, 256, (3, 37, 53), dtype=np.uint8)
    for kind, factor in distort_numpy.grid_specs()[1:] + [(5, 2.0), (5, 3.0), (1, 0.0), (4, 0.5), (4, -0.5)]:
        want = getattr(tvf, names[kind])(torch.from_numpy(img), factor).numpy()
        assert _distort_close(kind, distort_numpy.distort(img, kind, factor), want), (kind, factor)
