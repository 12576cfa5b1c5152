"""uint8 frames decoded / encoded INSIDE the kernels (ref: utils/data.py:99-106,
utils/postprocess.py:138, methods/__init__.py:18-30): the fused path must be byte-identical to
"decode on the host -> float transfer -> quantise", for every source layout the callers produce
(interleaved video frames, planar read_image tensors), odd sizes (scalar tails, unaligned planes),
both decode dtypes, and the float32 / clamped outputs the Runner needs."""

import numpy as np
import pytest

from conftest import synthetic_pair, u8_identical_fraction

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch
    import color_transfer_b200  # noqa: F401
    from color_transfer_b200 import _cabi, batch, device, sharded
    from oracle import reference_numpy as oracle
    return torch, _cabi, batch, device, sharded, oracle


def _frames(b, h, w, seed):
    pairs = [synthetic_pair(h, w, seed + i) for i in range(b)]
    t8 = np.stack([np.rint(p[0] * 255).astype(np.uint8) for p in pairs])
    r8 = np.stack([np.rint(p[1] * 255).astype(np.uint8) for p in pairs])
    return t8, r8


def _quantise(x):
    return np.rint(np.clip(np.asarray(x, dtype=np.float64), 0, 1) * 255).astype(np.uint8)


def _decode(x8, as_float32):
    dt = np.float32 if as_float32 else np.float64
    return (x8 / dt(255)).astype(dt)


SHAPES = [(2, 64, 64), (3, 37, 45), (1, 128, 96), (2, 33, 129), (1, 1, 5), (1, 270, 480)]


@pytest.mark.parametrize("b,h,w", SHAPES)
@pytest.mark.parametrize("as_float32", [True, False])
@pytest.mark.parametrize("planar", [False, True])
def test_linear_u8_equals_decoded_float_path(mods, b, h, w, as_float32, planar):
    torch, _cabi, batch, device, sharded, oracle = mods
    t8, r8 = _frames(b, h, w, 700)
    tf, rf = _decode(t8, as_float32), _decode(r8, as_float32)

    def dev8(x):
        if planar:   # read_image layout: CHW memory, handed over as an HWC view
            return torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2))).cuda().permute(0, 2, 3, 1)
        return torch.from_numpy(x).cuda()

    for code in (_cabi.CT_REINHARD, _cabi.CT_MKL_MK, _cabi.CT_CCS, _cabi.CT_MKL_CHOLESKY):
        got = device.linear_transfer(code, dev8(t8), dev8(r8), as_float32=as_float32)
        assert got.dtype == torch.uint8 and got.shape == (b, h, w, 3)
        ref = device.linear_transfer(code, torch.from_numpy(tf).cuda(), torch.from_numpy(rf).cuda())
        assert np.array_equal(got.cpu().numpy(), _quantise(ref.cpu().numpy())), f"method {code}"
        # float results straight from uint8 frames: what the float path computes.  The statistics are summed
        # in another thread order (16 instead of 4 or 2 pixels per thread): last-ulp differences, and for
        # Reinhard the fp32 group sums of the Lab statistics pass differ at 1e-8.
        f = device.linear_transfer(code, dev8(t8), dev8(r8), as_float32=as_float32, out_dtype=ref.dtype)
        tol = 2e-6 if (ref.dtype == torch.float32 or code == _cabi.CT_REINHARD) else 1e-12
        assert float((f.double() - ref.double()).abs().max()) <= tol


@pytest.mark.parametrize("b,h,w", SHAPES)
@pytest.mark.parametrize("as_float32", [True, False])
@pytest.mark.parametrize("planar", [False, True])
def test_idt_u8_equals_decoded_float_path(mods, b, h, w, as_float32, planar):
    torch, _cabi, batch, device, sharded, oracle = mods
    t8, r8 = _frames(b, h, w, 900)
    tf, rf = _decode(t8, as_float32), _decode(r8, as_float32)
    rots = sharded.predraw_rotations(b, 4, seed=21)
    drot = torch.from_numpy(rots).cuda()

    def dev8(x):
        if planar:
            return torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2))).cuda().permute(0, 2, 3, 1)
        return torch.from_numpy(x).cuda()

    ref = device.idt_transfer(torch.from_numpy(tf).cuda(), torch.from_numpy(rf).cuda(), drot)      # float64
    got = device.idt_transfer(dev8(t8), dev8(r8), drot, as_float32=as_float32)
    assert got.dtype == torch.uint8
    assert np.array_equal(got.cpu().numpy(), _quantise(ref.cpu().numpy()))
    f64 = device.idt_transfer(dev8(t8), dev8(r8), drot, as_float32=as_float32, out_dtype=torch.float64)
    assert torch.equal(f64, ref)
    f32 = device.idt_transfer(dev8(t8), dev8(r8), drot, as_float32=as_float32, out_dtype=torch.float32)
    assert torch.equal(f32, ref.float())                                    # torch.from_numpy(out).float()
    f32c = device.idt_transfer(dev8(t8), dev8(r8), drot, as_float32=as_float32, out_dtype=torch.float32, clamp=True)
    assert torch.equal(f32c, ref.float().clamp(0, 1))                       # .clamp(0, 1), methods/__init__.py:30
    # and against the CPU oracle on the decoded frames
    for i in range(b):
        want = oracle.iterative_distribution_transfer(tf[i], rf[i], rotations=rots[i])
        assert np.max(np.abs(ref[i].cpu().numpy() - want)) < 1e-9
        if h * w >= 10000:
            assert u8_identical_fraction(got[i].cpu().numpy() / 255.0, want) >= 0.9999


def test_idt_u8_stage_counts_bit_exact(mods):
    """Histogram counts of uint8 frames through the stage API == the oracle's on the decoded frames."""
    torch, _cabi, batch, device, sharded, oracle = mods
    t8, r8 = _frames(1, 101, 77, 950)
    tf, rf = _decode(t8, True), _decode(r8, True)
    rots = sharded.predraw_rotations(1, 4, seed=5)
    st = device.IdtStages(torch.from_numpy(t8).cuda(), torch.from_numpy(r8).cuda(), torch.from_numpy(rots).cuda())
    counts = []
    out = st.run(between=lambda n, x: counts.append(x.clone()) if n == "counts" else None, fuse_lut=False)
    st.raise_for_status()
    want, traces = oracle.idt_instrumented(tf[0], rf[0], rotations=rots[0], keep_arrays=False)
    for i, c in enumerate(counts):
        c = c.cpu().numpy().reshape(2, 3, 255)
        assert np.array_equal(c[0], traces[i]["counts_t"]) and np.array_equal(c[1], traces[i]["counts_r"])
    assert np.max(np.abs(out[0].cpu().numpy() - want)) < 1e-9


def test_host_u8_api_single_iteration_and_batches(mods):
    """ct_idt_transfer_host_u8 with n_iter = 1 (no float64 state to convert from) and with more pairs than
    pipeline slots; ct_linear_transfer_host_u8."""
    torch, _cabi, batch, device, sharded, oracle = mods
    t8, r8 = _frames(5, 48, 80, 1200)
    tf, rf = _decode(t8, True), _decode(r8, True)
    for n_iter in (1, 2, 4):
        rots = sharded.predraw_rotations(5, n_iter, seed=8)
        got = batch.idt_frames_u8(t8, r8, 255, n_iter, rotations=rots)
        ref = device.idt_transfer(torch.from_numpy(tf).cuda(), torch.from_numpy(rf).cuda(), torch.from_numpy(rots).cuda(), 255, n_iter)
        assert np.array_equal(got, _quantise(ref.cpu().numpy())), f"n_iter={n_iter}"
    got = batch.linear_transfer_frames_u8("mkl", t8, r8)
    ref = device.linear_transfer(_cabi.CT_MKL_MK, torch.from_numpy(tf).cuda(), torch.from_numpy(rf).cuda())
    assert np.array_equal(got, _quantise(ref.cpu().numpy()))


def test_runner_uint8_tensors(mods):
    """Runner.forward on uint8 CHW CUDA tensors (read_image output) == Runner on the same frames / 255
    (the reference's `target / 255` in the dataset, utils/data.py:106), result float32 CHW."""
    torch, _cabi, batch, device, sharded, oracle = mods
    import methods
    t8, r8 = _frames(3, 40, 56, 1300)
    u8 = {"target": torch.from_numpy(np.ascontiguousarray(t8.transpose(0, 3, 1, 2))).cuda(),
          "reference": torch.from_numpy(np.ascontiguousarray(r8.transpose(0, 3, 1, 2))).cuda()}
    # the reference divides on the CPU (correctly rounded); torch's CUDA `/ 255` multiplies by 1/255
    fl = {"target": torch.from_numpy(np.ascontiguousarray((t8 / np.float32(255)).astype(np.float32).transpose(0, 3, 1, 2))).cuda(),
          "reference": torch.from_numpy(np.ascontiguousarray((r8 / np.float32(255)).astype(np.float32).transpose(0, 3, 1, 2))).cuda()}
    for spec in ("methods.linear.color_transfer_between_images", "methods.linear.monge_kantorovitch_color_transfer",
                 "methods.iterative.iterative_distribution_transfer"):
        runner = methods.Runner(spec)
        np.random.seed(3)
        a = runner(u8)
        np.random.seed(3)
        b = runner(fl)
        assert a.dtype == torch.float32 and a.shape == (3, 3, 40, 56)
        exact = "iterative" in spec      # IDT has no floating-point reduction; the linear statistics are summed
        tol = 0.0 if exact else 1e-6     # in another thread order for uint8 groups (last-ulp differences)
        assert float((a - b).abs().max()) <= tol, spec
        np.random.seed(3)
        runner._clamp_fused = True
        c = runner(u8)
        runner._clamp_fused = False
        assert float((c - b.clamp(0, 1)).abs().max()) <= tol, spec


def test_reinhard_numpy_wrapper_integer_inputs(mods):
    """skimage.rgb2lab runs img_as_float first: a uint8 array is divided by 255 (ref: methods/linear.py:25-26);
    a float32 target with a float64 reference promotes to float64."""
    torch, _cabi, batch, device, sharded, oracle = mods
    import methods.linear as lin
    t8, r8 = _frames(1, 60, 70, 1400)
    a = lin.color_transfer_between_images(t8[0], r8[0])
    b = lin.color_transfer_between_images(t8[0] / 255.0, r8[0] / 255.0)
    assert a.dtype == np.float64 and np.array_equal(a, b)
    c = lin.color_transfer_between_images((t8[0] / 255.0).astype(np.float32), r8[0] / 255.0)
    assert c.dtype == np.float64
