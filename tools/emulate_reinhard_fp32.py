"""CPU emulation of candidate mixed-precision Reinhard remap chains (design aid for csrc/ct_lab.cuh).

Models MUFU lg2/ex2 as the exact function with a uniform relative error of 2^-22 and evaluates
the max-abs error and the uint8 flip rate against the float64 oracle for
  A: everything fp32
  B: fp32 decode / encode, fp64 island (matrix, cbrt polish, affine, cube, inverse matrix)
  C: as B with the forward matrix in fp32
Run: python tools/emulate_reinhard_fp32.py
"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from conftest import synthetic_pair, u8_identical_fraction
from oracle import reference_numpy as oracle
from oracle import skimage_color as sk

f32 = np.float32
rng = np.random.default_rng(0)

def mufu(fn, x):
    y = fn(x.astype(np.float64))
    y = y * (1 + rng.uniform(-1, 1, y.shape) * 2.0 ** -22)
    return y.astype(f32)

def lg2(x): return mufu(np.log2, x)
def ex2(x): return mufu(np.exp2, x)

M = (sk.XYZ_FROM_RGB / sk.WHITE_D65_2[:, None])
Minv = sk.RGB_FROM_XYZ * sk.WHITE_D65_2[None, :]

def decode32(v):
    u = ((v + f32(0.055)) * f32(1 / 1.055)).astype(f32)
    p = ex2((f32(0.4) * lg2(u)).astype(f32))
    hi = ((u * u).astype(f32) * p).astype(f32)
    return np.where(v > f32(0.04045), hi, (v * f32(1 / 12.92)).astype(f32)).astype(f32)

def encode32(c):
    c = c.astype(f32)
    cs = np.maximum(c, f32(1e-30))
    p = ex2((f32(1 / 2.4) * lg2(cs)).astype(f32))
    hi = (f32(1.055) * p - f32(0.055)).astype(f32)
    s = np.where(c > f32(0.0031308), hi, (f32(12.92) * c).astype(f32))
    return np.clip(s, 0, 1).astype(f32)

def chain(t, st, mode):
    v = t.reshape(-1, 3).astype(f32)
    l = decode32(v)
    if mode == "A":
        xyz = (l @ M.T.astype(f32)).astype(f32)
        z = ex2((f32(-1 / 3) * lg2(np.maximum(xyz, f32(1e-30)))).astype(f32))
        z = (z * (f32(4 / 3) - f32(1 / 3) * ((xyz * z).astype(f32) * (z * z).astype(f32)).astype(f32)).astype(f32)).astype(f32)
        cb = (xyz * (z * z).astype(f32)).astype(f32)
        f = np.where(xyz > f32(0.008856), cb, (f32(7.787) * xyz + f32(16 / 116)).astype(f32)).astype(f32)
        dt = f32
    else:
        if mode == "B":
            xyz = l.astype(np.float64) @ M.T
        else:
            xyz = (l @ M.T.astype(f32)).astype(f32).astype(np.float64)
        z = ex2((f32(-1 / 3) * lg2(np.maximum(xyz, 1e-30).astype(f32))).astype(f32)).astype(np.float64)
        z = z * (4 / 3 - (xyz * z) * (z * z) / 3)
        cb = xyz * z * z
        f = np.where(xyz > 0.008856, cb, 7.787 * xyz + 16 / 116)
        dt = np.float64
    fx, fy, fz = f[:, 0], f[:, 1], f[:, 2]
    mu_t, sd_t, mu_r, sd_r = [a.astype(dt) for a in st]
    L = (dt(116) * fy - dt(16)); a = dt(500) * (fx - fy); b = dt(200) * (fy - fz)
    L = ((L - mu_t[0]) * (sd_r[0] / sd_t[0]) + mu_r[0]).astype(dt)
    a = ((a - mu_t[1]) * (sd_r[1] / sd_t[1]) + mu_r[1]).astype(dt)
    b = ((b - mu_t[2]) * (sd_r[2] / sd_t[2]) + mu_r[2]).astype(dt)
    fy = ((L + dt(16)) * dt(1 / 116)).astype(dt); fx = (a * dt(1 / 500) + fy).astype(dt); fz = np.maximum(fy - b * dt(1 / 200), 0).astype(dt)
    g = np.stack([fx, fy, fz], -1)
    g = np.where(g > dt(0.2068966), g * g * g, (g - dt(16 / 116)) * dt(1 / 7.787)).astype(dt)
    rgb = (g @ Minv.T.astype(dt)).astype(dt)
    return encode32(rgb).reshape(t.shape)

for seed, (h, w) in enumerate([(512, 512), (540, 960), (300, 400)]):
    t, r = synthetic_pair(h, w, 12 + seed, np.float32)
    t64, r64 = t.astype(np.float64), r.astype(np.float64)
    ref = oracle.color_transfer_between_images(t64, r64)
    lt, lr = sk.rgb2lab(t64).reshape(-1, 3), sk.rgb2lab(r64).reshape(-1, 3)
    st = (lt.mean(0), lt.std(0), lr.mean(0), lr.std(0))
    for mode in "ABC":
        out = chain(t, st, mode)
        err = np.abs(out.astype(np.float64) - ref)
        print(h, w, mode, "max %.3e mean %.3e u8 %.6f" % (err.max(), err.mean(), u8_identical_fraction(out, ref)))
    print("  f32-rounded oracle: u8 %.6f" % u8_identical_fraction(ref.astype(f32), ref))
