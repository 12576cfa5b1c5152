"""TEST INFRASTRUCTURE — generate tests/golden/metrics.npz from the UNMODIFIED ref: utils/icid.py.

Run in the build container (needs /root/reference, torch, torchvision):

    python -m oracle.gen_golden_metrics

The reference function (oracle/load_reference.icid: real torch / torchvision, kornia's rgb_to_lab
stubbed) is executed on seeded synthetic batches and on the 0964 stereo pair; the float64
restatement oracle/metrics_numpy.py must agree with every value to 5e-6 before the file is
written.  Inputs are regenerated from the seeds by tests/ (metric_case below), only the values
are stored."""

import os

import numpy as np

from . import load_reference
from . import metrics_numpy as M
from .gen_golden import GOLDEN, load_0964, synthetic_pair

CASES = [  # name, (h, w, seed) or "0964"; shapes hit downscale factors 1, 2 and 3
    ("s96x128", (96, 128, 31)),
    ("s300x420", (300, 420, 32)),
    ("s540x960", (540, 960, 33)),
    ("pair0964", "0964"),
]
VARIANTS = [("perceptual", False, True), ("perceptual", True, True), ("hue-preserving", False, True),
            ("chromatic", False, True), ("chromatic", True, True), ("perceptual", False, False)]


def metric_case(spec):
    """[B,3,H,W] float32 batches (a, b) of a case."""
    if spec == "0964":
        left, right = load_0964()
        return (left.transpose(2, 0, 1)[None].astype(np.float32), right.transpose(2, 0, 1)[None].astype(np.float32))
    h, w, seed = spec
    t, r = synthetic_pair(h, w, seed, np.float32)
    a = np.stack([t.transpose(2, 0, 1), np.roll(t, 5, axis=1).transpose(2, 0, 1)])
    b = np.stack([r.transpose(2, 0, 1), r.transpose(2, 0, 1)])
    return np.ascontiguousarray(a), np.ascontiguousarray(b)


def main():
    if not load_reference.available():
        raise SystemExit("needs /root/reference (build container only)")
    import torch
    ref_icid = load_reference.icid()
    out = {}
    for name, spec in CASES:
        a, b = metric_case(spec)
        vals = []
        for intent, omit, down in VARIANTS:
            if not down and min(a.shape[-2:]) > 600:
                vals.append(np.nan)        # full-resolution variant only on the small cases
                continue
            v = float(ref_icid(torch.from_numpy(a), torch.from_numpy(b), intent=intent, omit_maps67=omit, downsampling=down))
            mine = M.icid(a, b, intent=intent, omit_maps67=omit, downsampling=down)
            if abs(v - mine) > 5e-6:   # the reference computes in float32
                raise SystemExit(f"oracle iCID differs from the reference on {name} {intent} {omit} {down}: {v} vs {mine}")
            vals.append(v)
        out["icid_" + name] = np.array(vals)
        print(name, vals)
    np.savez(os.path.join(GOLDEN, "metrics.npz"), **out)


if __name__ == "__main__":
    main()
