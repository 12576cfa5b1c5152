"""Kernel-level known-answer tests of the linear path (SURVEY section 4 tier 1, section 7.2):
K1 `ct_moments` against np.mean / np.cov, K2 `ct_linear_solve` against the golden transforms the
UNMODIFIED reference produced (tests/golden/*.npz: T_MK, T_sqrt, T_cholesky, T_ccs), and a census of
how often the device's CCS sign convention differs from LAPACK's on random pairs.
ref: methods/linear.py:33-36, 64-78, 103-118."""

import ctypes

import numpy as np
import pytest

from conftest import synthetic_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch
    import color_transfer_b200  # noqa: F401
    from color_transfer_b200 import _cabi, device
    from oracle import reference_numpy as oracle
    return torch, _cabi, device, oracle


def _moments(mods, x, lab=0):
    """Raw sums [10] of one image tensor through ct_moments."""
    torch, _cabi, device, _ = mods
    x4 = device._check_images(x, "img")
    h = device._handle_for(x4, None)
    xb, _keep = device.batch_of(x4)
    sums = torch.empty((x4.shape[0], _cabi.CT_MOMENT_DOUBLES), dtype=torch.float64, device=x.device)
    h.check(h.lib.ct_moments(h.h, xb, lab, ctypes.c_void_p(sums.data_ptr())))
    return sums.cpu().numpy(), xb.layout


def _mean_cov_from_sums(s, shift=0.5):
    """np.mean / np.cov (ddof 1) from {n, S(x-K), S(x-K)(x-K)^T as 00,01,02,11,12,22}."""
    n = s[0]
    m1 = s[1:4] / n
    sxx = np.array([[s[4], s[5], s[6]], [s[5], s[7], s[8]], [s[6], s[8], s[9]]])
    cov = (sxx - n * np.outer(m1, m1)) / (n - 1)
    return m1 + shift, cov


def _solve(mods, method, sums_t, sums_r):
    torch, _cabi, device, _ = mods
    h = _cabi.default_handle(0)
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    st = torch.from_numpy(np.ascontiguousarray(sums_t)).cuda()
    sr = torch.from_numpy(np.ascontiguousarray(sums_r)).cuda()
    count = st.shape[0]
    xform = torch.empty((count, _cabi.CT_XFORM_DOUBLES), dtype=torch.float64, device="cuda")
    status = torch.zeros((count,), dtype=torch.int32, device="cuda")
    h.check(h.lib.ct_linear_solve(h.h, method, ctypes.c_void_p(st.data_ptr()), ctypes.c_void_p(sr.data_ptr()), count,
                                  ctypes.c_void_p(xform.data_ptr()), ctypes.c_void_p(status.data_ptr())))
    return xform.cpu().numpy(), status.cpu().numpy()


SHAPES = [(48, 56), (37, 53), (1, 7), (255, 129), (540, 960)]


@pytest.mark.parametrize("h,w", SHAPES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("planar", [False, True])
def test_moments_equal_numpy_mean_and_cov(mods, h, w, dtype, planar):
    """Device mean / covariance vs np.mean / np.cov <= 1e-12 relative (float32 and float64 images,
    interleaved and planar memory, odd sizes with scalar tails)."""
    torch = mods[0]
    t, _ = synthetic_pair(h, w, 17, dtype)
    if planar:   # CHW memory viewed as HWC: what the reference Runner hands over (methods/__init__.py:21-22)
        x = torch.from_numpy(np.ascontiguousarray(t.transpose(2, 0, 1))).cuda().permute(1, 2, 0)
    else:
        x = torch.from_numpy(t).cuda()
    sums, layout = _moments(mods, x)
    assert layout == (1 if planar and h * w > 1 else layout)
    mean, cov = _mean_cov_from_sums(sums[0])
    x64 = t.reshape(-1, 3).astype(np.float64)
    want_mean, want_cov = x64.mean(axis=0), np.cov(x64.T)
    assert sums[0][0] == h * w
    assert np.max(np.abs(mean - want_mean)) <= 1e-12 * max(1.0, np.max(np.abs(want_mean)))
    assert np.max(np.abs(cov - want_cov)) <= 1e-12 * max(1e-3, np.max(np.abs(want_cov)))


def test_moments_are_additive_over_row_shards(mods):
    """The property the row-sharded mode rests on: shard sums add up to the whole image's sums exactly
    enough that mean / cov agree to 1e-13 (the all-gather + fixed-order sum of sharded.py)."""
    torch = mods[0]
    t, _ = synthetic_pair(301, 203, 5, np.float32)
    whole, _ = _moments(mods, torch.from_numpy(t).cuda())
    parts = [_moments(mods, torch.from_numpy(np.ascontiguousarray(t[a:b])).cuda())[0][0] for a, b in ((0, 100), (100, 217), (217, 301))]
    acc = parts[0] + parts[1] + parts[2]
    m0, c0 = _mean_cov_from_sums(whole[0])
    m1, c1 = _mean_cov_from_sums(acc)
    assert acc[0] == whole[0][0]
    assert np.max(np.abs(m0 - m1)) < 1e-13 and np.max(np.abs(c0 - c1)) < 1e-13


@pytest.mark.parametrize("name", ["small_f64", "small_f32", "pair0964"])
def test_solve_equals_golden_transforms(mods, golden, pair0964, name):
    """K2 from the device's own moments against the matrices the unmodified reference computed
    (golden T_MK / T_sqrt / T_cholesky / T_ccs, mean_t / mean_r): <= 1e-12 (MKL), CCS up to the
    documented sign convention, which must agree with LAPACK on the 0964 pair."""
    torch, _cabi, device, oracle = mods
    g = golden[name]
    if name == "pair0964":
        t, r = pair0964
    else:
        t, r = g["target"], g["reference"]
    st, _ = _moments(mods, torch.from_numpy(np.ascontiguousarray(t)).cuda())
    sr, _ = _moments(mods, torch.from_numpy(np.ascontiguousarray(r)).cuda())
    mean_t, cov_t = _mean_cov_from_sums(st[0])
    assert np.max(np.abs(cov_t - g["cov_t"])) <= 1e-12 * np.max(np.abs(g["cov_t"])) + 1e-15
    for method, key, transpose in ((_cabi.CT_MKL_MK, "T_MK", False), (_cabi.CT_MKL_SQRT, "T_sqrt", False),
                                   (_cabi.CT_MKL_CHOLESKY, "T_cholesky", False), (_cabi.CT_CCS, "T_ccs", True)):
        xf, status = _solve(mods, method, st, sr)
        assert status[0] == 0
        m = xf[0][:9].reshape(3, 3)            # out = (x - mu_t) @ M + mu_r
        want = g[key].T if transpose else g[key]   # linear.py:80 applies T.T for CCS, :122 T for MKL
        if name == "small_f32":   # the reference's float32 np.mean differs from the float64 gate
            t64, r64 = t.astype(np.float64), r.astype(np.float64)
            _mu, ct = oracle.mean_and_cov(t64)
            _mu, cr = oracle.mean_and_cov(r64)
            want = oracle.ccs_matrix(ct, cr).T if transpose else oracle.mkl_matrix(ct, cr, key[2:])
        err = np.max(np.abs(m - want))
        assert err <= 1e-11, f"{name} {key}: {err:.3e}"
        assert np.max(np.abs(xf[0][9:12] - t.reshape(-1, 3).astype(np.float64).mean(axis=0))) <= 1e-12
        assert np.max(np.abs(xf[0][12:15] - r.reshape(-1, 3).astype(np.float64).mean(axis=0))) <= 1e-12


def test_ccs_sign_census(mods, pair0964, capsys):
    """count on how many of N random covariance pairs the device's rule
    (dot(u_r,i, u_t,i) >= 0) gives a different matrix than LAPACK's SVD signs, and require that every
    difference is a pure sign choice.  The 0964 pair must need no alignment."""
    torch, _cabi, device, oracle = mods
    rng = np.random.default_rng(7)
    n = 200
    sums_t, sums_r, covs = [], [], []
    for _ in range(n):
        def rand_sums():
            a = rng.standard_normal((3, 3)) * rng.uniform(0.02, 0.3)
            cov = a @ a.T + 1e-4 * np.eye(3)
            mean = rng.uniform(0.2, 0.8, 3)
            npx = 5000.0
            m1 = mean - 0.5
            sxx = cov * (npx - 1) + npx * np.outer(m1, m1)
            return np.array([npx, *(npx * m1), sxx[0, 0], sxx[0, 1], sxx[0, 2], sxx[1, 1], sxx[1, 2], sxx[2, 2]]), cov
        s_t, c_t = rand_sums()
        s_r, c_r = rand_sums()
        sums_t.append(s_t); sums_r.append(s_r); covs.append((c_t, c_r))
    xf, status = _solve(mods, _cabi.CT_CCS, np.array(sums_t), np.array(sums_r))
    assert (status == 0).all()
    needed = 0
    signs = [(a, b, c) for a in (1, -1) for b in (1, -1) for c in (1, -1)]
    for i, (c_t, c_r) in enumerate(covs):
        m = xf[i][:9].reshape(3, 3)
        if np.max(np.abs(m - oracle.ccs_matrix(c_t, c_r).T)) <= 1e-9:
            continue
        needed += 1
        best = min(np.max(np.abs(m - oracle.ccs_matrix(c_t, c_r, s).T)) for s in signs)
        assert best <= 1e-9, f"pair {i}: not a sign choice of the reference's matrix ({best:.3e})"
    with capsys.disabled():
        print(f"\n[ccs sign census] {needed} of {n} random covariance pairs needed sign alignment against LAPACK dgesdd")
    t, r = pair0964
    st, _ = _moments(mods, torch.from_numpy(t).cuda())
    sr, _ = _moments(mods, torch.from_numpy(r).cuda())
    xf, _ = _solve(mods, _cabi.CT_CCS, st, sr)
    mu_t, c_t = oracle.mean_and_cov(t)
    mu_r, c_r = oracle.mean_and_cov(r)
    assert np.max(np.abs(xf[0][:9].reshape(3, 3) - oracle.ccs_matrix(c_t, c_r).T)) <= 1e-10, "the 0964 pair must need no sign alignment"
