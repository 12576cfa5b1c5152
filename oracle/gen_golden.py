"""TEST INFRASTRUCTURE — generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):

    python -m oracle.gen_golden

For every case the real reference functions (loaded by path, oracle/load_reference.py)
are executed, the numpy restatement (oracle/reference_numpy.py) is required to return
bit-identical arrays for Xiao / MKL / IDT, and the reference's outputs are stored.  The
committed fixtures are what pins the oracle on machines without /root/reference.

Cases
  small_f64 / small_f32 : seeded 48x56 synthetic stereo pair, all functions, full outputs
                          and full IDT traces (rotations, ranges, counts, LUTs).
  pair0964              : graphics/0964_{L,R}.png (copied to tests/golden/), float64;
                          covariances, the three MKL matrices, IDT ranges/counts/LUTs after
                          np.random.seed(42), and a strided sample (every 997th value) of
                          each function's output plus min/max/mean.
"""

import hashlib
import os

import numpy as np
from PIL import Image

from . import load_reference
from . import reference_numpy as oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
IDT_SEED = 42
SAMPLE_STRIDE = 997


def synthetic_pair(h, w, seed, dtype):
    """Small smooth-ish stereo pair quantised to uint8 then scaled to [0,1]."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.zeros((h, w, 3))
    for c in range(3):
        f, g, p = rng.uniform(0, 3), rng.uniform(0, 3), rng.uniform(0, 2 * np.pi)
        base[..., c] = 128 + 90 * np.sin(2 * np.pi * (f * xx / w + g * yy / h) + p) + 12 * rng.standard_normal((h, w))
    ref = np.clip(base, 0, 255).astype(np.uint8)
    gain, gamma = rng.uniform(0.7, 1.3, 3), rng.uniform(0.7, 1.3, 3)
    tgt = np.clip(255 * gain * (ref / 255.0) ** gamma, 0, 255).astype(np.uint8)
    tgt = np.roll(tgt, 3, axis=1)
    return (tgt / 255.0).astype(dtype), (ref / 255.0).astype(dtype)


def load_0964():
    left = np.asarray(Image.open(os.path.join(GOLDEN, "0964_L.png")).convert("RGB")) / 255.0
    right = np.asarray(Image.open(os.path.join(GOLDEN, "0964_R.png")).convert("RGB")) / 255.0
    return left, right


def _same(a, b, what):
    if not np.array_equal(a, b):
        raise SystemExit(f"oracle restatement differs from the reference for {what}: "
                         f"max-abs {np.max(np.abs(np.asarray(a, float) - np.asarray(b, float)))}")


def run_case(ref_lin, ref_it, tgt, ref, full):
    out = {}
    # --- linear, real reference code -----------------------------------------------
    out["reinhard"] = ref_lin.color_transfer_between_images(tgt, ref)
    _same(out["reinhard"], oracle.color_transfer_between_images(tgt, ref), "reinhard")
    out["ccs"] = ref_lin.color_transfer_in_correlated_color_space(tgt, ref)
    _same(out["ccs"], oracle.color_transfer_in_correlated_color_space(tgt, ref), "ccs")
    for dec in ("MK", "sqrt", "cholesky"):
        out["mkl_" + dec] = ref_lin.monge_kantorovitch_color_transfer(tgt, ref, decomposition=dec)
        _same(out["mkl_" + dec], oracle.monge_kantorovitch_color_transfer(tgt, ref, dec), "mkl " + dec)
    mu_t, cov_t = oracle.mean_and_cov(tgt)
    mu_r, cov_r = oracle.mean_and_cov(ref)
    out.update(mean_t=mu_t, mean_r=mu_r, cov_t=cov_t, cov_r=cov_r)
    for dec in ("MK", "sqrt", "cholesky"):
        out["T_" + dec] = oracle.mkl_matrix(cov_t, cov_r, dec)
    out["T_ccs"] = oracle.ccs_matrix(cov_t, cov_r)
    lm_t, ls_t = oracle.lab_statistics(tgt)
    lm_r, ls_r = oracle.lab_statistics(ref)
    out.update(lab_mean_t=lm_t, lab_std_t=ls_t, lab_mean_r=lm_r, lab_std_r=ls_r)
    # --- IDT, real reference code, then the instrumented restatement with the same seed
    np.random.seed(IDT_SEED)
    out["idt"] = ref_it.iterative_distribution_transfer(tgt, ref)
    np.random.seed(IDT_SEED)
    idt_o, traces = oracle.idt_instrumented(tgt, ref)
    _same(out["idt"], idt_o, "idt")
    out["idt_rot"] = np.stack([t["rot"] for t in traces])
    out["idt_lo"] = np.stack([t["lo"] for t in traces])
    out["idt_hi"] = np.stack([t["hi"] for t in traces])
    out["idt_counts_t"] = np.stack([t["counts_t"] for t in traces])
    out["idt_counts_r"] = np.stack([t["counts_r"] for t in traces])
    out["idt_lut"] = np.stack([t["lut"] for t in traces])
    # non-default arguments
    np.random.seed(IDT_SEED + 1)
    out["idt_b64_n2"] = ref_it.iterative_distribution_transfer(tgt, ref, bins=64, n_iter=2)
    np.random.seed(IDT_SEED + 1)
    _same(out["idt_b64_n2"], oracle.iterative_distribution_transfer(tgt, ref, 64, 2), "idt b64")
    # automated colour grading = IDT + regrain (the reference code with the restated resize)
    np.random.seed(IDT_SEED)
    out["acg"] = ref_it.automated_color_grading(tgt, ref)
    np.random.seed(IDT_SEED)
    _same(out["acg"], oracle.automated_color_grading(tgt, ref), "automated_color_grading")
    if not full:
        for k in ("reinhard", "ccs", "mkl_MK", "mkl_sqrt", "mkl_cholesky", "idt", "idt_b64_n2", "acg"):
            v = out.pop(k)
            out[k + "_sample"] = v.reshape(-1)[::SAMPLE_STRIDE].copy()
            out[k + "_stats"] = np.array([v.min(), v.max(), v.mean()])
            out[k + "_u8_sha256"] = np.frombuffer(
                hashlib.sha256(np.rint(np.clip(v, 0, 1) * 255).astype(np.uint8).tobytes()).digest(),
                dtype=np.uint8)
    return out


def main():
    if not load_reference.available():
        raise SystemExit("needs /root/reference (build container only)")
    ref_lin, ref_it = load_reference.linear(), load_reference.iterative()
    os.makedirs(GOLDEN, exist_ok=True)
    for name, dtype in (("small_f64", np.float64), ("small_f32", np.float32)):
        tgt, ref = synthetic_pair(48, 56, 7, dtype)          # > 40 px per side: the regrain pyramid has 2 levels
        res = run_case(ref_lin, ref_it, tgt, ref, full=True)
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), target=tgt, reference=ref, **res)
        print(name, {k: (v.shape, str(v.dtype)) for k, v in res.items() if k in ("reinhard", "idt", "ccs")})
    left, right = load_0964()
    res = run_case(ref_lin, ref_it, left, right, full=False)
    np.savez_compressed(os.path.join(GOLDEN, "pair0964.npz"), **res)
    print("pair0964 idt stats", res["idt_stats"], "T_MK", res["T_MK"])


if __name__ == "__main__":
    main()
