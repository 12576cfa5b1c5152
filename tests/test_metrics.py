"""Quality metrics: iCID and PSNR.

CPU: the float64 restatement (oracle/metrics_numpy.py) against the golden values the UNMODIFIED
ref: utils/icid.py produced (tests/golden/metrics.npz, oracle/gen_golden_metrics.py) and, in the
build container, against the reference itself.  GPU: the device implementation
(color_transfer_b200.metrics -> ct_icid / ct_psnr) against the restatement on the same inputs.
Tolerances: the reference computes in float32, so golden-vs-float64 agree to ~2e-6 on a value of
order 1; the device (fp32 maps, fp64 sums) is held to 2e-5."""

import numpy as np
import pytest

from conftest import GOLDEN  # noqa: F401
from oracle import load_reference
from oracle import metrics_numpy as M
from oracle.gen_golden_metrics import CASES, VARIANTS, metric_case

REF_TOL = 5e-6
DEV_TOL = 2e-5


@pytest.fixture(scope="module")
def golden_metrics():
    import os
    return np.load(os.path.join(GOLDEN, "metrics.npz"))


@pytest.mark.parametrize("name,spec", CASES[:3])
def test_oracle_icid_matches_reference_golden(golden_metrics, name, spec):
    a, b = metric_case(spec)
    for (intent, omit, down), want in zip(VARIANTS, golden_metrics["icid_" + name]):
        if np.isnan(want):
            continue
        assert abs(M.icid(a, b, intent=intent, omit_maps67=omit, downsampling=down) - want) <= REF_TOL


def test_oracle_icid_pair0964_and_errors(golden_metrics):
    a, b = metric_case("0964")
    assert abs(M.icid(a, b) - golden_metrics["icid_pair0964"][0]) <= REF_TOL
    assert M.icid(a, a) == pytest.approx(0.0, abs=1e-12)
    with pytest.raises(ValueError, match="Intent"):
        M.icid(a, b, intent="vivid")


@pytest.mark.skipif(not load_reference.available(), reason="needs /root/reference (build container)")
def test_oracle_icid_against_the_reference_itself():
    import torch
    ref_icid = load_reference.icid()
    rng = np.random.default_rng(3)
    a = rng.random((2, 3, 70, 90), dtype=np.float32)
    b = np.clip(a * 0.9 + 0.05 * rng.random(a.shape, dtype=np.float32), 0, 1)
    for intent in ("perceptual", "hue-preserving", "chromatic"):
        want = float(ref_icid(torch.from_numpy(a), torch.from_numpy(b), intent=intent))
        assert abs(M.icid(a, b, intent=intent) - want) <= REF_TOL


def test_oracle_psnr_definition():
    rng = np.random.default_rng(4)
    x = rng.random((3, 3, 20, 24))
    y = np.clip(x + 0.1, 0, 1)
    mse = ((x - y) ** 2).reshape(3, -1).mean(1)
    assert M.psnr(x, y) == pytest.approx(np.mean(10 * np.log10(1 / (mse + 1e-8))), rel=1e-12)
    assert M.psnr(x, x) == pytest.approx(80.0, rel=1e-12)        # -10 log10(1e-8)


def test_oracle_ssim_definition():
    """The restatement against a direct conv2d transcription of the SSIM definition (Wang et al.
    2004 with piq's defaults) and its basic properties."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(5)
    x = rng.random((2, 3, 300, 340))
    y = np.clip(x + 0.08 * rng.standard_normal(x.shape), 0, 1)
    tx, ty = F.avg_pool2d(torch.from_numpy(x), 1), F.avg_pool2d(torch.from_numpy(y), 1)   # min side 300 -> factor 1
    c = torch.arange(11, dtype=torch.float64) - 5
    g = (-(c[None] ** 2 + c[:, None] ** 2) / (2 * 1.5 ** 2)).exp()
    k = (g / g.sum())[None, None].repeat(3, 1, 1, 1)
    mx, my = F.conv2d(tx, k, groups=3), F.conv2d(ty, k, groups=3)
    sxx, syy = F.conv2d(tx * tx, k, groups=3) - mx ** 2, F.conv2d(ty * ty, k, groups=3) - my ** 2
    sxy = F.conv2d(tx * ty, k, groups=3) - mx * my
    ss = (2 * mx * my + 1e-4) / (mx ** 2 + my ** 2 + 1e-4) * (2 * sxy + 9e-4) / (sxx + syy + 9e-4)
    assert M.ssim(x, y) == pytest.approx(float(ss.mean((-1, -2)).mean(1).mean()), abs=1e-12)
    assert M.ssim(x, x) == pytest.approx(1.0, abs=1e-12)
    assert M.ssim(x, y) == pytest.approx(M.ssim(y, x), abs=1e-12)
    assert M.ssim(x, 1.0 - x) < 0.2


# ------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def dev_metrics():
    import torch

    import color_transfer_b200  # noqa: F401
    from color_transfer_b200 import metrics
    return torch, metrics


@pytest.mark.gpu
@pytest.mark.parametrize("name,spec", CASES)
def test_device_icid_matches_oracle_and_golden(dev_metrics, golden_metrics, name, spec):
    torch, metrics = dev_metrics
    a, b = metric_case(spec)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    for (intent, omit, down), want in zip(VARIANTS, golden_metrics["icid_" + name]):
        if np.isnan(want):
            continue
        got = metrics.icid(da, db, intent=intent, omit_maps67=omit, downsampling=down)
        assert got.is_cuda and got.dtype == torch.float32 and got.dim() == 0
        assert abs(float(got) - want) <= DEV_TOL, (name, intent, omit, down)
        assert abs(float(got) - M.icid(a, b, intent=intent, omit_maps67=omit, downsampling=down)) <= DEV_TOL


@pytest.mark.gpu
def test_device_icid_shapes_and_errors(dev_metrics):
    torch, metrics = dev_metrics
    rng = np.random.default_rng(6)
    for h, w in ((37, 53), (6, 200), (515, 260), (768, 1030)):     # ragged tiles; factors 1, 2, 3
        a = rng.random((1, 3, h, w), dtype=np.float32)
        b = np.clip(a + 0.2 * (rng.random(a.shape, dtype=np.float32) - 0.5), 0, 1)
        got = float(metrics.icid(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()))
        assert abs(got - M.icid(a, b)) <= DEV_TOL, (h, w)
    x = torch.rand(3, 64, 64, device="cuda")
    assert float(metrics.icid(x, x)) == pytest.approx(0.0, abs=1e-6)       # [3,H,W] accepted
    with pytest.raises(ValueError, match="Intent"):
        metrics.icid(x, x, intent="vivid")
    with pytest.raises(TypeError):
        metrics.icid(x.cpu(), x.cpu())                                      # no CPU fallback
    with pytest.raises(Exception):
        metrics.icid(x[:, :4, :4], x[:, :4, :4])                            # smaller than the blur radius


@pytest.mark.gpu
def test_device_psnr_and_runner_test_step(dev_metrics):
    torch, metrics = dev_metrics
    rng = np.random.default_rng(8)
    x = rng.random((4, 3, 90, 131), dtype=np.float32)
    y = np.clip(x + 0.05 * rng.standard_normal(x.shape).astype(np.float32), 0, 1)
    got = float(metrics.psnr(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()))
    assert got == pytest.approx(M.psnr(x, y), abs=1e-4)
    assert float(metrics.psnr(torch.from_numpy(x).cuda(), torch.from_numpy(x).cuda())) == pytest.approx(80.0, abs=1e-4)
    # the Runner's test_step (ref: methods/__init__.py:29-40) on a CUDA batch
    import methods
    runner = methods.Runner("methods.linear.monge_kantorovitch_color_transfer")
    batch = {"target": torch.from_numpy(y).cuda(), "reference": torch.from_numpy(x).cuda(), "gt": torch.from_numpy(x).cuda()}
    vals = runner.test_step(batch, 0)
    out = runner(batch).clamp(0, 1).cpu().numpy()
    assert float(vals["Test PSNR"]) == pytest.approx(M.psnr(out, x), abs=1e-3)
    assert float(vals["Test iCID"]) == pytest.approx(M.icid(out, x), abs=DEV_TOL)
    assert float(vals["Test SSIM"]) == pytest.approx(M.ssim(out, x), abs=DEV_TOL)


@pytest.mark.gpu
def test_device_ssim_matches_oracle(dev_metrics):
    torch, metrics = dev_metrics
    rng = np.random.default_rng(11)
    for b, h, w in ((2, 64, 80), (1, 300, 420), (1, 540, 960), (1, 11, 45), (1, 700, 1029)):   # factors 1, 1, 2, 1, 3
        x = rng.random((b, 3, h, w), dtype=np.float32)
        y = np.clip(x + 0.1 * rng.standard_normal(x.shape).astype(np.float32), 0, 1)
        got = metrics.ssim(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda())
        assert got.is_cuda and got.dtype == torch.float32
        assert float(got) == pytest.approx(M.ssim(x, y), abs=2e-5), (b, h, w)
    assert float(metrics.ssim(torch.from_numpy(x).cuda(), torch.from_numpy(x).cuda())) == pytest.approx(1.0, abs=1e-6)
    assert float(metrics.ssim(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), downsample=False)) == \
        pytest.approx(M.ssim(x, y, downsample=False), abs=2e-5)
    with pytest.raises(Exception):
        metrics.ssim(torch.rand(1, 3, 8, 40, device="cuda"), torch.rand(1, 3, 8, 40, device="cuda"))
