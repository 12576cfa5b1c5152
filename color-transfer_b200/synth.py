"""Synthetic stereo pairs of the shapes BASELINE.json names.

``frame_pair`` is the numpy definition (used for parity cases and the CPU baseline);
``frame_pairs_cuda`` draws pairs of the same family directly on the device for the benchmark
(same construction, torch's RNG - the benchmark needs shape and statistics, not identical bits).

primary: per channel a smooth field of four sinusoids plus sensor-like noise, quantised to
uint8; reference = the field shifted 16 px (disparity); target = per-channel gain and gamma.
stress: i.i.d. uniform uint8 (worst case for histogram spread).
"""

import numpy as np


def frame_pair(h, w, seed, dtype=np.float32, stress=False):
    """(target, reference) as k/255 in ``dtype``."""
    target, reference = frame_pair_u8(h, w, seed, stress)
    return (target / 255.0).astype(dtype), (reference / 255.0).astype(dtype)


def frame_pair_u8(h, w, seed, stress=False):
    """(target, reference) as uint8 video frames [H,W,3]."""
    rng = np.random.default_rng(seed)
    if stress:
        g = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    else:
        yy, xx = np.mgrid[0:h, 0:w]
        g = np.empty((h, w, 3))
        for c in range(3):
            field = np.zeros((h, w))
            for k in range(1, 5):
                a = rng.uniform(0, 1) / k
                f, gg, p = rng.uniform(0, 8), rng.uniform(0, 8), rng.uniform(0, 2 * np.pi)
                field += a * np.sin(2 * np.pi * (f * xx / w + gg * yy / h) + p)
            g[..., c] = 128 + 96 * field + 8 * rng.standard_normal((h, w))
        g = np.clip(g, 0, 255).astype(np.uint8)
    reference = np.roll(g, 16, axis=1)
    gain, gamma = rng.uniform(0.7, 1.3, 3), rng.uniform(0.7, 1.3, 3)
    target = np.clip(255.0 * gain * (g / 255.0) ** gamma, 0, 255).astype(np.uint8)
    return target, reference


def frame_pairs_cuda(count, h, w, seed, device, dtype=None, stress=False):
    """[count,H,W,3] target and reference tensors on ``device`` (float32 by default)."""
    import math
    import torch
    dtype = dtype or torch.float32
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    tgt = torch.empty((count, h, w, 3), dtype=dtype, device=device)
    ref = torch.empty((count, h, w, 3), dtype=dtype, device=device)
    yy = torch.arange(h, device=device, dtype=torch.float32).view(h, 1) / h
    xx = torch.arange(w, device=device, dtype=torch.float32).view(1, w) / w
    for i in range(count):
        if stress:
            g = torch.randint(0, 256, (h, w, 3), generator=gen, device=device).float()
        else:
            g = torch.empty((h, w, 3), device=device)
            for c in range(3):
                field = torch.zeros((h, w), device=device)
                for k in range(1, 5):
                    u = torch.rand(4, generator=gen, device=device)
                    field += (u[0] / k) * torch.sin(2 * math.pi * (8 * u[1] * xx + 8 * u[2] * yy) + 2 * math.pi * u[3])
                g[..., c] = 128 + 96 * field + 8 * torch.randn((h, w), generator=gen, device=device)
            g = g.clamp_(0, 255).floor_()
        # k/255 through float64: torch divides float32 by a scalar as a multiply by 1/255, which is
        # not the correctly rounded quotient a CPU loader (uint8 / 255) produces
        ref[i] = (torch.roll(g, 16, dims=1).double() / 255.0).to(dtype)
        gg = torch.rand(6, generator=gen, device=device) * 0.6 + 0.7
        t = (255.0 * gg[:3] * (g / 255.0) ** gg[3:]).clamp_(0, 255).floor_()
        tgt[i] = (t.double() / 255.0).to(dtype)
    return tgt, ref
