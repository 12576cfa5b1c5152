"""GPU tests of the device-resident API, the batched host API, the stage-by-stage C-ABI calls and
the multi-GPU drivers (the 2-GPU cases skip on a single-GPU box)."""

import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, synthetic_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch

    import color_transfer_b200  # noqa: F401
    from color_transfer_b200 import _cabi, batch, device, sharded, synth
    from oracle import reference_numpy as oracle
    return torch, _cabi, batch, device, sharded, synth, oracle


def _stack(n, h, w, dtype, seed=100):
    pairs = [synthetic_pair(h, w, seed + i, dtype) for i in range(n)]
    return np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_batched_device_and_host_apis(mods, dtype):
    torch, _cabi, batch, device, sharded, synth, oracle = mods
    B, H, W = 5, 33, 47                                   # odd sizes: unaligned image strides
    t, r = _stack(B, H, W, dtype)
    dt, dr = torch.from_numpy(t).cuda(), torch.from_numpy(r).cuda()
    t64, r64 = t.astype(np.float64), r.astype(np.float64)
    for name, code, fn in (("mkl", _cabi.CT_MKL_MK, oracle.monge_kantorovitch_color_transfer),
                           ("reinhard", _cabi.CT_REINHARD, oracle.color_transfer_between_images)):
        dev = device.linear_transfer(code, dt, dr).cpu().numpy()
        host = batch.linear_transfer_frames(name, t, r)
        assert np.array_equal(dev, host)                  # same kernels either way
        for i in range(B):
            assert np.max(np.abs(dev[i] - fn(t64[i], r64[i]))) < 1e-4
    rot = sharded.predraw_rotations(B, 4, seed=42)
    dev = device.idt_transfer(dt, dr, torch.from_numpy(rot).cuda()).cpu().numpy()
    host = batch.idt_frames(t, r, rotations=rot)
    assert np.array_equal(dev, host)
    for i in range(B):
        want = oracle.iterative_distribution_transfer(t[i], r[i], rotations=rot[i])
        assert np.max(np.abs(dev[i] - want)) < 1e-10
    # CHW-memory views (the Runner's layout), batched
    dt_chw = torch.from_numpy(np.ascontiguousarray(t.transpose(0, 3, 1, 2))).cuda().permute(0, 2, 3, 1)
    dr_chw = torch.from_numpy(np.ascontiguousarray(r.transpose(0, 3, 1, 2))).cuda().permute(0, 2, 3, 1)
    assert np.array_equal(device.idt_transfer(dt_chw, dr_chw, torch.from_numpy(rot).cuda()).cpu().numpy(), dev)


def test_stage_api_unfused_equals_fused(mods):
    torch, _cabi, batch, device, sharded, synth, oracle = mods
    t, r = _stack(3, 64, 80, np.float32, seed=200)
    dt, dr = torch.from_numpy(t).cuda(), torch.from_numpy(r).cuda()
    rot = torch.from_numpy(sharded.predraw_rotations(3, 4, seed=1)).cuda()
    fused = device.idt_transfer(dt, dr, rot).cpu().numpy()
    seen = []
    st = device.IdtStages(dt, dr, rot)
    out = st.run(between=lambda name, tensor: seen.append((name, tuple(tensor.shape))), fuse_lut=False).cpu().numpy()
    st.raise_for_status()
    assert np.array_equal(out, fused)
    assert [s[0] for s in seen] == ["keys"] + ["counts", "keys"] * 3 + ["counts"]
    assert np.array_equal(device.IdtStages(dt, dr, rot).run().cpu().numpy(), fused)   # fused LUT through the stage API


def test_synthetic_frames_and_linear_device_parity(mods):
    torch, _cabi, batch, device, sharded, synth, oracle = mods
    t, r = synth.frame_pairs_cuda(2, 135, 240, 7, torch.device("cuda", 0))
    assert t.shape == (2, 135, 240, 3) and t.dtype == torch.float32 and 0.0 <= float(t.min()) and float(t.max()) <= 1.0
    out = device.linear_transfer(_cabi.CT_REINHARD, t, r).cpu().numpy()
    for i in range(2):
        want = oracle.color_transfer_between_images(t[i].cpu().numpy().astype(np.float64), r[i].cpu().numpy().astype(np.float64))
        assert np.max(np.abs(out[i] - want)) < 1e-4
    tn, rn = synth.frame_pair(54, 96, 1000, np.float32)
    assert tn.shape == (54, 96, 3) and tn.dtype == np.float32


def test_row_sharded_drivers_world1(mods):
    """With one rank the sharded drivers must reproduce the unsharded result bit for bit."""
    torch, _cabi, batch, device, sharded, synth, oracle = mods
    t, r = synthetic_pair(70, 90, 41, np.float32)
    dt, dr = torch.from_numpy(t).cuda(), torch.from_numpy(r).cuda()
    rot = sharded.predraw_rotations(1, 4, seed=3)[0]
    out = sharded.idt_transfer_sharded(dt, dr, rot).cpu().numpy()
    assert np.array_equal(out, device.idt_transfer(dt, dr, torch.from_numpy(rot[None]).cuda()).cpu().numpy())
    for code in (_cabi.CT_MKL_MK, _cabi.CT_REINHARD, _cabi.CT_CCS):
        a = sharded.linear_transfer_sharded(code, dt, dr).cpu().numpy()
        b = device.linear_transfer(code, dt, dr).cpu().numpy()
        assert np.max(np.abs(a - b)) < 1e-12


def _run_dist_check(nproc, port, same_device):
    env = dict(os.environ)
    if same_device:
        env["CT_DIST_SAME_DEVICE"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-6000:]
    assert "DIST_GPU_CHECK_OK" in res.stdout
    return res.stdout


@pytest.mark.parametrize("world", [2, 3])
def test_row_sharding_ranks_on_one_gpu_vs_oracle(mods, world):
    """SURVEY 8e / 8d config-5 protocol on a ONE-GPU box: `world` ranks share cuda:0 and all-reduce over
    gloo, so the row-sharded drivers, their kernels and the collective schedule run where the driver's
    pytest runs.  Inside: a 2048x2048 float32 pair of the config-5 generator, rows sharded, bit-identical
    (counts and output) to the single-GPU run and bit-exact counts / <=1e-9 output against the CPU oracle;
    world = 3 gives ragged row blocks (683/683/682 rows)."""
    _run_dist_check(world, 29520 + world, same_device=True)


def test_two_gpu_row_sharding_and_frame_parallel(mods):
    """The same body, one GPU per rank over NCCL (the product configuration)."""
    torch = mods[0]
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2); the one-GPU gloo variant above covers the same checks")
    _run_dist_check(2, 29517, same_device=False)


def test_runner_contract_host_and_device_paths(mods):
    """methods.Runner(func_spec).forward(batch): the reference's CHW float32 contract (ref:
    methods/__init__.py:18-27) through the numpy path, and the device-resident fast path."""
    torch, _cabi, batch, device, sharded, synth, oracle = mods
    import methods
    t, r = _stack(3, 40, 56, np.float32, seed=300)
    bt = {"target": torch.from_numpy(np.ascontiguousarray(t.transpose(0, 3, 1, 2))),
          "reference": torch.from_numpy(np.ascontiguousarray(r.transpose(0, 3, 1, 2)))}
    for spec, fn in (("methods.linear.color_transfer_between_images", oracle.color_transfer_between_images),
                     ("methods.linear.monge_kantorovitch_color_transfer", oracle.monge_kantorovitch_color_transfer)):
        runner = methods.Runner(spec)
        host = runner(bt)
        assert host.shape == (3, 3, 40, 56) and host.dtype == torch.float32
        dev = runner({k: v.cuda() for k, v in bt.items()})
        assert dev.is_cuda and dev.dtype == torch.float32
        assert torch.equal(dev.cpu(), host)
        for i in range(3):
            want = fn(t[i].astype(np.float64), r[i].astype(np.float64))
            assert np.max(np.abs(host[i].permute(1, 2, 0).numpy() - want)) < 1e-4
    runner = methods.Runner("methods.iterative.iterative_distribution_transfer")
    np.random.seed(11)
    host = runner(bt)
    np.random.seed(11)
    dev = runner({k: v.cuda() for k, v in bt.items()})
    assert torch.equal(dev.cpu(), host)          # same rotations, same kernels
    np.random.seed(11)
    for i in range(3):
        want = oracle.iterative_distribution_transfer(t[i], r[i])
        assert np.max(np.abs(host[i].permute(1, 2, 0).numpy() - want)) < 1e-6


@pytest.mark.parametrize("as_float32", [True, False])
def test_uint8_frame_io(mods, as_float32):
    """uint8 frames in / out: decode k/255 as the reference's loaders do, transfer,
    clip + round like img_as_ubyte.  Compared with the oracle run on the decoded float frames."""
    torch, _cabi, batch, device, sharded, synth, oracle = mods
    rng = np.random.default_rng(9)
    B, H, W = 3, 37, 45
    pairs = [synthetic_pair(H, W, 400 + i) for i in range(B)]
    t8 = np.stack([np.rint(p[0] * 255).astype(np.uint8) for p in pairs])
    r8 = np.stack([np.rint(p[1] * 255).astype(np.uint8) for p in pairs])
    dt = np.float32 if as_float32 else np.float64
    tf, rf = (t8 / dt(255)).astype(dt), (r8 / dt(255)).astype(dt)

    def quantise(x):
        return np.rint(np.clip(x, 0, 1) * 255).astype(np.uint8)

    rot = sharded.predraw_rotations(B, 4, seed=8)
    out = batch.idt_frames_u8(t8, r8, rotations=rot, as_float32=as_float32)
    assert out.dtype == np.uint8 and out.shape == t8.shape
    for i in range(B):
        want = quantise(oracle.iterative_distribution_transfer(tf[i], rf[i], rotations=rot[i]))
        assert np.mean(out[i] == want) >= 0.9999
    for name, fn in (("mkl", oracle.monge_kantorovitch_color_transfer), ("reinhard", oracle.color_transfer_between_images)):
        out = batch.linear_transfer_frames_u8(name, t8, r8, as_float32=as_float32)
        for i in range(B):
            want = quantise(fn(tf[i].astype(np.float64), rf[i].astype(np.float64)))
            assert np.mean(out[i] == want) >= (0.999 if H * W < 5000 else 0.9999)   # 37 x 45 frames: one flip is 2e-4
    # the float path on the decoded frames gives the same bytes
    again = quantise(batch.idt_frames(tf, rf, rotations=rot))
    assert np.array_equal(again, batch.idt_frames_u8(t8, r8, rotations=rot, as_float32=as_float32))


def test_large_batches_run_chunked_on_two_streams(mods):
    """Batches of >= 32 Mpix go through the two-stream chunked driver (ct_linear_transfer): the
    result must match the same pairs transferred one call at a time and the oracle."""
    torch, _cabi, batch, device, sharded, synth, oracle = mods
    B, H, W = 71, 540, 960                                  # 36.8 Mpix: chunks of 36 and 35 pairs
    dt, dr = synth.frame_pairs_cuda(B, H, W, 4000, "cuda")
    for code, fn, tol in ((_cabi.CT_REINHARD, oracle.color_transfer_between_images, 1e-4),
                          (_cabi.CT_MKL_MK, oracle.monge_kantorovitch_color_transfer, 1e-9)):
        whole = device.linear_transfer(code, dt, dr)
        for b in (0, 35, 36, 70):
            single = device.linear_transfer(code, dt[b:b + 1], dr[b:b + 1])[0]
            assert float((whole[b].double() - single.double()).abs().max()) <= 1e-6
        for b in (35, 70):
            ref = fn(dt[b].cpu().numpy().astype(np.float64), dr[b].cpu().numpy().astype(np.float64))
            assert np.max(np.abs(whole[b].cpu().numpy() - ref)) <= tol
    torch.cuda.synchronize()


def test_many_small_pairs_in_one_chunked_call(mods):
    """7000 pairs of 96x96 in one device-resident call: the two-stream chunks hold ~3500 pairs each, more
    than one partial-sum slot per resident CTA - the regions of the two streams must not overlap
    (each chunk's moments are checked against single-pair calls)."""
    torch, _cabi, batch, device, sharded, synth, oracle = mods
    gen = torch.Generator(device="cuda").manual_seed(5)
    B, H, W = 7000, 96, 96
    t = torch.rand((B, H, W, 3), generator=gen, device="cuda")
    r = torch.rand((B, H, W, 3), generator=gen, device="cuda") * 0.8 + 0.1
    out = device.linear_transfer(_cabi.CT_MKL_MK, t, r)
    torch.cuda.synchronize()
    for i in (0, 1, 3471, 3472, 3473, 5000, 6999):
        one = device.linear_transfer(_cabi.CT_MKL_MK, t[i], r[i])
        assert float((out[i] - one).abs().max()) < 1e-12, f"pair {i}"


def test_reinhard_float32_toes_and_nan(mods):
    """float32 Reinhard: images that live entirely in the linear toes of the gamma / Lab curves
    (the patched rare branches of ct_lab.cuh), and NaN propagation like numpy (a NaN pixel turns the
    statistics, hence every output, into NaN)."""
    torch, _cabi, batch, device, sharded, synth, oracle = mods
    rng = np.random.default_rng(5)
    dark_t = (rng.integers(0, 12, (64, 80, 3)) / 255.0).astype(np.float32)      # all below 0.04045
    dark_r = (rng.integers(0, 30, (64, 80, 3)) / 255.0).astype(np.float32)
    mixed_t = dark_t.copy()
    mixed_t[::7, ::5] = (rng.integers(0, 256, mixed_t[::7, ::5].shape) / 255.0).astype(np.float32)
    for t, r in ((dark_t, dark_r), (mixed_t, dark_r), (dark_r, mixed_t)):
        out = device.linear_transfer(_cabi.CT_REINHARD, torch.from_numpy(t).cuda(), torch.from_numpy(r).cuda()).cpu().numpy()
        ref = oracle.color_transfer_between_images(t.astype(np.float64), r.astype(np.float64))
        assert out.dtype == np.float32 and out.min() >= 0.0 and out.max() <= 1.0
        assert np.max(np.abs(out - ref)) <= 1e-4
    bad = mixed_t.copy()
    bad[3, 4, 1] = np.nan
    for t, r in ((bad, dark_r), (dark_r, bad)):
        out = device.linear_transfer(_cabi.CT_REINHARD, torch.from_numpy(t).cuda(), torch.from_numpy(r).cuda()).cpu().numpy()
        assert np.isnan(out).all()


def test_bench_prints_one_contract_line(mods):
    """bench.py (our arm, N=1): exactly one JSON line on stdout with the keys the driver reads."""
    import json
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--frames", "2",
                        "--e2e-frames", "2", "--no-extras"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout[-2000:]
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["unit"] == "Mpix/s" and d["n_gpus"] == 1 and d["steps"] == 3 and d["scaling"] == "weak"
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["config"]["workload"].startswith("configs[3]")
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and 0 < d["e2e"]["value"] < d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] == 1 and c["value"] > 0 and "sample" in c
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
