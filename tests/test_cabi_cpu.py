"""CPU tests of the drop-in boundary: the shared library loads and exports every symbol that
include/ct_b200.h declares, the Python mirror keeps the reference's names / signatures / errors,
and everything fails loudly (no CPU fallback) when no CUDA device is present."""

import ctypes
import inspect
import os
import re

import numpy as np
import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "ct_b200.h")


@pytest.fixture(scope="module")
def cabi():
    import color_transfer_b200  # noqa: F401
    from color_transfer_b200 import _cabi
    return _cabi


def _declared_functions():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(ct_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(cabi):
    lib = cabi.load_library()
    declared = _declared_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"libct_b200.so does not export {name}"
    assert sorted(cabi.SIGNATURES) == declared, "ctypes prototypes and header disagree"
    assert lib.ct_abi_version() == int(re.search(r"#define CT_ABI_VERSION (\d+)", open(HEADER).read()).group(1))


def test_header_constants_match_binding(cabi):
    text = open(HEADER).read()
    assert int(re.search(r"#define CT_IDT_MAX_BINS (\d+)", text).group(1)) == cabi.CT_IDT_MAX_BINS
    assert int(re.search(r"#define CT_XFORM_DOUBLES (\d+)", text).group(1)) == cabi.CT_XFORM_DOUBLES
    assert int(re.search(r"#define CT_MOMENT_DOUBLES (\d+)", text).group(1)) == cabi.CT_MOMENT_DOUBLES
    assert ctypes.sizeof(cabi.Batch) == 48 and ctypes.sizeof(cabi.IdtStage) == 96 and ctypes.sizeof(cabi.IdtTrace) == 40
    assert cabi.lut_doubles(255) == 3 * (3 * 256 + 4) and cabi.lut_doubles(64) == 3 * (3 * 66 + 4)
    assert cabi.load_library().ct_idt_workspace_bytes(1000, 2, 255, 4) > 2 * 3 * 1000 * 8


def test_range_keys_are_monotone(cabi):
    lib = cabi.load_library()
    vals = np.concatenate([[-np.inf, -1e300, -2.5, -1e-300, -0.0, 0.0, 1e-300, 0.1, 2.5, 1e300, np.inf],
                           np.random.default_rng(0).normal(size=200)])
    keys = np.array([lib.ct_idt_key_of(float(v)) for v in vals], dtype=np.int64)
    order = np.argsort(vals, kind="stable")
    assert np.all(np.diff(keys[order]) >= 0)
    for v, k in zip(vals, keys):
        assert lib.ct_idt_value_of(int(k)) == v
    assert lib.ct_idt_key_of(float("inf")) == 0x7FF0000000000000


def test_no_cpu_fallback(cabi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    assert cabi.load_library().ct_create(0, ctypes.byref(h)) == cabi.CT_E_CUDA
    with pytest.raises(cabi.CtError):
        cabi.Handle(0)
    import methods.iterative
    import methods.linear
    t = np.random.default_rng(0).random((8, 8, 3))
    for fn in (methods.linear.color_transfer_between_images, methods.linear.color_transfer_in_correlated_color_space,
               methods.linear.monge_kantorovitch_color_transfer, methods.iterative.iterative_distribution_transfer):
        with pytest.raises(cabi.CtError):
            fn(t, t)


def test_reference_signatures_and_errors():
    import methods
    import methods.iterative
    import methods.linear

    def positional(fn):
        return [(p.name, p.default) for p in inspect.signature(fn).parameters.values()
                if p.kind == inspect.Parameter.POSITIONAL_OR_KEYWORD]

    e = inspect.Parameter.empty
    # ref: methods/linear.py:8, :45, :85 and methods/iterative.py:8
    assert positional(methods.linear.color_transfer_between_images) == [("target", e), ("reference", e)]
    assert positional(methods.linear.color_transfer_in_correlated_color_space) == [("target", e), ("reference", e)]
    assert positional(methods.linear.monge_kantorovitch_color_transfer) == [("target", e), ("reference", e), ("decomposition", "MK")]
    assert positional(methods.iterative.iterative_distribution_transfer) == [("target", e), ("reference", e), ("bins", 255), ("n_iter", 4)]
    # the plugin mechanism: a dotted path resolved with importlib (ref: methods/__init__.py:14-16)
    f = methods.resolve("methods.linear.color_transfer_between_images")
    assert f is methods.linear.color_transfer_between_images
    assert methods.resolve("color_transfer_b200.methods.iterative.iterative_distribution_transfer") is \
        methods.iterative.iterative_distribution_transfer
    t = np.zeros((4, 4, 3))
    with pytest.raises(ValueError, match="Unknown decomposition, use either 'cholesky', 'sqrt', or 'MK'"):
        methods.linear.monge_kantorovitch_color_transfer(t, t, decomposition="svd")     # before any GPU work
    with pytest.raises(ValueError):
        methods.linear.color_transfer_between_images(np.zeros((4, 4)), t)
    out = methods.iterative.iterative_distribution_transfer(t + 0.25, t, n_iter=0)           # loop body never runs
    assert out.dtype == np.float64 and np.array_equal(out, t + 0.25)


def test_batch_descriptors(cabi):
    a = np.zeros((5, 7, 3), dtype=np.float32)
    b, keep = cabi.batch_from_numpy(a)
    assert (b.npix, b.count, b.dtype, b.layout) == (35, 1, cabi.CT_F32, cabi.CT_HWC) and keep is a
    chw = np.zeros((3, 5, 7), dtype=np.float32)
    view = chw.transpose(1, 2, 0)                      # what the reference Runner passes
    assert view.strides == (28, 4, 140)
    b, keep = cabi.batch_from_numpy(view)
    assert b.layout == cabi.CT_CHW and b.data == chw.ctypes.data
    b, keep = cabi.batch_from_numpy(np.zeros((10, 14, 3))[::2, ::2])
    assert b.layout == cabi.CT_HWC and keep.flags.c_contiguous and b.npix == 35 and b.dtype == cabi.CT_F64
    b, _ = cabi.batch_from_numpy(np.zeros((4, 5, 7, 3)))
    assert (b.count, b.image_stride) == (4, 105)
    b, _ = cabi.batch_from_numpy(np.zeros((5, 7, 3), dtype=np.uint8))
    assert b.dtype == cabi.CT_F64                      # integer inputs are promoted like the reference's float math
    with pytest.raises(ValueError):
        cabi.batch_from_numpy(np.zeros((5, 7, 4)))


def test_rotation_draws_follow_the_reference_rng_stream():
    import scipy.stats

    from color_transfer_b200.methods.iterative import draw_rotations
    from color_transfer_b200.sharded import predraw_rotations
    np.random.seed(42)
    want = np.stack([scipy.stats.special_ortho_group.rvs(3) for _ in range(8)])
    np.random.seed(42)
    got = draw_rotations(8)
    assert np.array_equal(got, want)
    assert np.array_equal(predraw_rotations(2, 4, seed=42).reshape(8, 3, 3), want)
    np.testing.assert_allclose(np.einsum("nij,nkj->nik", got, got), np.broadcast_to(np.eye(3), (8, 3, 3)), atol=1e-12)
    np.testing.assert_allclose(np.linalg.det(got), 1.0, atol=1e-12)


def test_header_is_plain_c_and_library_links_without_python(tmp_path):
    """gcc (C, not C++) compiles a consumer of include/ct_b200.h, links libct_b200.so and runs it."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    libdir = os.path.join(ROOT, "color-transfer_b200")
    exe = str(tmp_path / "abi_smoke")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", exe, "-L", libdir, "-lct_b200",
                    "-Wl,-rpath," + libdir, "-lm"], check=True, capture_output=True, text=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, f"abi_smoke exit code {res.returncode}: {res.stdout}{res.stderr}"
    assert "abi_smoke ok" in res.stdout


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) needs no GPU: it must
    print exactly one JSON line on stdout carrying the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--height", "96", "--width", "128"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "Mpix/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["metric"].startswith("stereopair Mpix/s")
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("configs[3]")
