"""GPU parity at BASELINE.json's full sizes: the synthetic frames of configs[2] (960x540) and
configs[3] (3840x2160) against the oracle, plus size-independent properties (permutation
invariance, histogram checksums, moment matching) that hold for any size."""

import numpy as np
import pytest

from conftest import u8_identical_fraction

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch

    import color_transfer_b200  # noqa: F401
    import methods.iterative
    import methods.linear
    from color_transfer_b200 import _cabi, device, sharded, synth
    from oracle import reference_numpy as oracle
    return torch, _cabi, device, sharded, synth, oracle, methods.linear, methods.iterative


@pytest.mark.parametrize("index,stress", [(0, False), (1, False), (517, False), (1034, False), (0, True)])
def test_config3_pairs_linear(mods, index, stress):
    """configs[2]: pair i of the 1035 synthetic 960x540 float32 pairs (seed 1000+i), Reinhard and MKL;
    primary (smooth field) distribution and, for pair 0, the i.i.d. uniform stress distribution."""
    torch, _cabi, device, sharded, synth, oracle, lin, it = mods
    t, r = synth.frame_pair(540, 960, 1000 + index, np.float32, stress=stress)
    t64, r64 = t.astype(np.float64), r.astype(np.float64)
    out = lin.color_transfer_between_images(t, r)
    want = oracle.color_transfer_between_images(t64, r64)
    assert out.dtype == np.float32
    assert np.max(np.abs(out - want)) <= 1e-4
    assert u8_identical_fraction(out, want) >= 0.9999
    out = lin.monge_kantorovitch_color_transfer(t, r)
    assert np.max(np.abs(out - oracle.monge_kantorovitch_color_transfer(t64, r64))) <= 1e-9


@pytest.mark.parametrize("frame", [0, 599])
def test_config4_frame_idt_against_oracle(mods, frame):
    """configs[3]: frames 0 and 599 of the synthetic 4K stereo video (seed 2000 + k), float32, full
    oracle run; the rotations of frame k are draws 4k..4k+3 after np.random.seed(42) (frame order)."""
    torch, _cabi, device, sharded, synth, oracle, lin, it = mods
    t, r = synth.frame_pair(2160, 3840, 2000 + frame, np.float32)
    np.random.seed(42)
    rot = sharded.predraw_rotations(frame + 1, 4)[frame]
    want, traces = oracle.idt_instrumented(t, r, rotations=rot, keep_arrays=False)
    trace = {}
    out = it.iterative_distribution_transfer(t, r, rotations=rot, trace=trace)
    for i in range(4):
        assert np.array_equal(trace["counts_t"][i], traces[i]["counts_t"]), f"target counts differ at iteration {i}"
        assert np.array_equal(trace["counts_r"][i], traces[i]["counts_r"]), f"reference counts differ at iteration {i}"
        assert trace["counts_t"][i].sum() == 3 * t.shape[0] * t.shape[1]           # checksum of checksums
    assert np.array_equal(trace["lo"][0], traces[0]["lo"]) and np.array_equal(trace["lut"][0], traces[0]["lut"])
    assert np.max(np.abs(out - want)) <= 1e-9
    assert u8_identical_fraction(out, want) >= 0.9999


def test_permutation_invariance_at_4k(mods):
    """The statistics are global, so permuting the target's pixels must permute the output and
    nothing else - bit for bit for IDT (integer counts, order-independent ranges) and for the
    deterministic moment reduction up to its fixed summation order."""
    torch, _cabi, device, sharded, synth, oracle, lin, it = mods
    dev = torch.device("cuda", 0)
    t, r = synth.frame_pairs_cuda(1, 2160, 3840, 77, dev)
    n = 2160 * 3840
    perm = torch.randperm(n, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    tp = t.view(n, 3)[perm].view(1, 2160, 3840, 3).contiguous()
    np.random.seed(5)
    rot = torch.from_numpy(sharded.predraw_rotations(1, 4)).to(dev)
    a = device.idt_transfer(t, r, rot).view(n, 3)
    b = device.idt_transfer(tp, r, rot).view(n, 3)
    assert torch.equal(a[perm], b)
    a = device.linear_transfer(_cabi.CT_MKL_MK, t, r).view(n, 3)
    b = device.linear_transfer(_cabi.CT_MKL_MK, tp, r).view(n, 3)
    assert float((a[perm] - b).abs().max()) < 1e-12


def test_moment_matching_properties_at_4k(mods):
    """MKL maps the target's mean / covariance onto the reference's (Pitie & Kokaram 2007); the
    Reinhard output, taken back to Lab, has the reference's Lab mean wherever nothing clipped."""
    torch, _cabi, device, sharded, synth, oracle, lin, it = mods
    dev = torch.device("cuda", 0)
    t, r = synth.frame_pairs_cuda(1, 2160, 3840, 78, dev, dtype=torch.float64)
    for code in (_cabi.CT_MKL_MK, _cabi.CT_MKL_SQRT, _cabi.CT_MKL_CHOLESKY, _cabi.CT_CCS):
        out = device.linear_transfer(code, t, r).view(-1, 3)
        ref = r.view(-1, 3)
        assert float((out.mean(0) - ref.mean(0)).abs().max()) < 1e-12
        if code == _cabi.CT_MKL_MK:
            # only the symmetric MK map reproduces the covariance: the reference applies `@ T` (not
            # `@ T.T`, linear.py:122), which for "sqrt" / "cholesky" gives T^T cov_t T != cov_r
            assert float((torch.cov(out.T) - torch.cov(ref.T)).abs().max()) < 1e-12
    # transferring an image onto itself is the identity
    same = device.linear_transfer(_cabi.CT_MKL_MK, t, t)
    assert float((same - t).abs().max()) < 1e-12
    same = device.linear_transfer(_cabi.CT_REINHARD, t, t)
    assert float((same - t).abs().max()) < 1e-9
