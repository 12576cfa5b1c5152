"""K4 (projected ranges) known-answer tests: the fp32-screened kernel must return EXACTLY the fp64
minimum / maximum of `rot @ X.T` (ref: methods/iterative.py:34-35, 39-40) whatever the screen does -
ties at the extremes, saturated regions, values beyond the screen's validity bound, tiny ranges,
float64 images, ragged sizes (scalar tails), unaligned planes."""

import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch
    import color_transfer_b200  # noqa: F401
    from color_transfer_b200 import _cabi, device, sharded
    return torch, _cabi, device, sharded


def _device_ranges(mods, img, rots):
    """[n_rot, 6] (lo[3], hi[3]) of one image under rots [n_rot,3,3] through ct_idt_ranges."""
    torch, _cabi, device, _ = mods
    n_rot = rots.shape[0]
    x = torch.from_numpy(np.ascontiguousarray(img)).cuda()
    x4 = device._check_images(x, "img")
    h = device._handle_for(x4, None)
    xb, _keep = device.batch_of(x4)
    rot = torch.from_numpy(np.ascontiguousarray(rots, dtype=np.float64).reshape(1, n_rot * 9)).cuda()
    keys = torch.empty((n_rot, 6), dtype=torch.int64, device="cuda")
    status = torch.zeros((1,), dtype=torch.int32, device="cuda")
    h.check(h.lib.ct_idt_keys_init(h.h, ctypes.c_void_p(keys.data_ptr()), keys.numel()))
    h.check(h.lib.ct_idt_ranges(h.h, xb, ctypes.c_void_p(rot.data_ptr()), n_rot * 9, n_rot,
                                ctypes.c_void_p(keys.data_ptr()), 6 * n_rot, ctypes.c_void_p(status.data_ptr())))
    k = keys.cpu().numpy()
    vals = np.array([[h.lib.ct_idt_value_of(int(v)) for v in row] for row in k])
    vals[:, 3:] *= -1.0
    return vals, int(status.item())


def _numpy_ranges(img, rots):
    x = img.reshape(-1, 3)
    out = []
    for r in rots:
        p = r @ x.T          # float64 @ float32 -> float64, as in the reference
        out.append(np.concatenate([p.min(axis=1), p.max(axis=1)]))
    return np.array(out)


def _rots(mods, n, seed):
    return mods[3].predraw_rotations(1, n, seed=seed)[0]


CASES = {
    "smooth_u8_f32": lambda rng: (np.clip(128 + 90 * np.sin(np.linspace(0, 40, 300 * 200 * 3)).reshape(300, 200, 3)
                                          + 12 * rng.standard_normal((300, 200, 3)), 0, 255).astype(np.uint8) / 255.0).astype(np.float32),
    "saturated_white_and_black": lambda rng: np.where(rng.random((257, 129, 1)) < 0.4, 1.0, np.where(
        rng.random((257, 129, 1)) < 0.5, 0.0, rng.random((257, 129, 3)))).astype(np.float32),
    "constant": lambda rng: np.full((64, 64, 3), 0.25, dtype=np.float32),
    "two_values_tie": lambda rng: np.where(rng.random((128, 96, 1)) < 0.5, np.float32(1.0), np.float32(254 / 255)).repeat(3, axis=2).astype(np.float32),
    "beyond_screen_bound": lambda rng: (rng.random((96, 160, 3)) * 255.0).astype(np.float32),       # un-normalised frame
    "one_outlier": lambda rng: np.concatenate([rng.random((90, 64, 3)), np.full((1, 64, 3), 1000.0)]).astype(np.float32),
    "tiny_range": lambda rng: (0.5 + 1e-7 * rng.standard_normal((80, 128, 3))).astype(np.float32),
    "negative_and_large": lambda rng: ((rng.random((70, 90, 3)) - 0.5) * 7.0).astype(np.float32),
    "float64_image": lambda rng: rng.random((150, 130, 3)),
    "float64_close_values": lambda rng: 0.3 + 1e-12 * rng.standard_normal((64, 50, 3)),
    "ragged_tail": lambda rng: rng.random((7, 13, 3)).astype(np.float32),
    "single_pixel": lambda rng: rng.random((1, 1, 3)).astype(np.float32),
    "gradient_records": lambda rng: np.linspace(0, 1, 512 * 128 * 3, dtype=np.float32).reshape(512, 128, 3),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("n_rot", [1, 4, 6])
def test_ranges_exact(mods, name, n_rot):
    import zlib
    rng = np.random.default_rng(zlib.crc32(name.encode()) % 1000)
    img = CASES[name](rng)
    rots = _rots(mods, n_rot, seed=5 + n_rot)
    got, status = _device_ranges(mods, img, rots)
    want = _numpy_ranges(img, rots)
    assert status == 0
    if name == "single_pixel":   # numpy multiplies a 3x3 by a 3x1 with a different (gemv) kernel: 1 ulp
        assert np.allclose(got, want, rtol=0, atol=1e-16)
        return
    assert np.array_equal(got, want), f"{name}: ranges differ by {np.max(np.abs(got - want)):.3e}"


def test_ranges_planar_view_and_large(mods):
    """The Runner's CHW-memory view (planar kernel path) and a frame large enough that every CTA streams
    many tiles (4K): exact ranges under 4 rotations."""
    torch, _cabi, device, _ = mods
    from color_transfer_b200 import synth
    t, _r = synth.frame_pair(2160, 3840, 2000, np.float32)
    rots = _rots(mods, 4, seed=42)
    got, status = _device_ranges(mods, t, rots)
    assert status == 0 and np.array_equal(got, _numpy_ranges(t, rots))
    chw = np.ascontiguousarray(t[:301, :203].transpose(2, 0, 1))      # planes of 61103 floats: not 16-byte multiples
    x = torch.from_numpy(chw).cuda().permute(1, 2, 0)                 # HWC view of CHW memory
    x4 = device._check_images(x, "img")
    h = device._handle_for(x4, None)
    xb, _keep = device.batch_of(x4)
    assert xb.layout == _cabi.CT_CHW
    rot = torch.from_numpy(rots.reshape(1, 36)).cuda()
    keys = torch.empty((4, 6), dtype=torch.int64, device="cuda")
    h.check(h.lib.ct_idt_keys_init(h.h, ctypes.c_void_p(keys.data_ptr()), keys.numel()))
    h.check(h.lib.ct_idt_ranges(h.h, xb, ctypes.c_void_p(rot.data_ptr()), 36, 4, ctypes.c_void_p(keys.data_ptr()), 24, None))
    vals = np.array([[h.lib.ct_idt_value_of(int(v)) for v in row] for row in keys.cpu().numpy()])
    vals[:, 3:] *= -1.0
    assert np.array_equal(vals, _numpy_ranges(t[:301, :203], rots))


@pytest.mark.parametrize("bad", [np.nan, np.inf, -np.inf])
def test_ranges_nonfinite_sets_status(mods, bad):
    _cabi = mods[1]
    rng = np.random.default_rng(3)
    img = rng.random((64, 100, 3)).astype(np.float32)
    img[37, 61, 1] = bad
    _vals, status = _device_ranges(mods, img, _rots(mods, 4, seed=9))
    assert status == _cabi.CT_E_NONFINITE
