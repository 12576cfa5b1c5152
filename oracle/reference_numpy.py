"""TEST INFRASTRUCTURE — CPU oracle for the global statistical colour-transfer path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (color-transfer_b200/) never does and
fails loudly when its CUDA library is missing.

A numpy restatement of the reference's hot path, written stage by stage so that
every intermediate the CUDA kernels must reproduce bit-for-bit (rotations,
projected ranges, bin edges, integer counts, CDFs, inverse-CDF tables, bin
indices) can be inspected:

* Reinhard Lab mean/std matching      ref: methods/linear.py:8-42
* Xiao correlated colour space        ref: methods/linear.py:45-82
* Pitie Monge-Kantorovitch            ref: methods/linear.py:85-124
* Pitie iterative distribution transfer  ref: methods/iterative.py:8-59
* regrain / automated colour grading     ref: methods/iterative.py:62-138 (its
  ``skimage.transform.resize`` is the restated wrapper over scipy.ndimage, oracle/skimage_resize.py)

Pinning: ``oracle/gen_golden.py`` executes the UNMODIFIED reference files
(loaded by path from /root/reference, see oracle/load_reference.py) on seeded
inputs and on the 0964 stereo pair, checks that this restatement returns
bit-identical arrays for Xiao / MKL / IDT, and writes the golden vectors under
tests/golden/.  The reference itself has no tests or golden vectors
(SURVEY.md section 4), so those generated vectors are the pin.  The Lab
conversion inside the Reinhard transfer is the one part that cannot be pinned
(scikit-image is absent; see oracle/skimage_color.py: "parity unpinned").
"""

import numpy as np
import scipy.linalg
import scipy.stats

from . import skimage_color, skimage_resize

# --------------------------------------------------------------------------- linear


def lab_statistics(img):
    """Per-channel Lab mean and population std of one RGB image.
    ref: methods/linear.py:25-36 (np.mean / np.std over axis 0, ddof=0)."""
    lab = skimage_color.rgb2lab(img).reshape(-1, 3)
    return np.mean(lab, axis=0), np.std(lab, axis=0)


def color_transfer_between_images(target, reference):
    """Reinhard et al. 2001.  ref: methods/linear.py:8-42."""
    t_lab = skimage_color.rgb2lab(target)
    r_lab = skimage_color.rgb2lab(reference)
    hw3 = t_lab.shape
    t = t_lab.reshape(-1, 3)
    r = r_lab.reshape(-1, 3)
    mu_t, mu_r = np.mean(t, axis=0), np.mean(r, axis=0)
    sd_t, sd_r = np.std(t, axis=0), np.std(r, axis=0)
    moved = (t - mu_t) * sd_r / sd_t + mu_r              # :38
    return skimage_color.lab2rgb(moved.reshape(hw3))     # :40 (clips to [0,1])


def mean_and_cov(img):
    """np.mean(axis=0) and np.cov(X.T) (ddof=1) of the flattened pixels.
    ref: methods/linear.py:64-67 and :103-106."""
    x = np.asarray(img).reshape(-1, 3)
    return np.mean(x, axis=0), np.cov(x.T)


def ccs_matrix(cov_t, cov_r, signs=None):
    """T of Xiao & Ma: U_t S_t^-1/2 S_r^1/2 U_r^-1.  ref: methods/linear.py:69-78.
    ``signs`` (length 3, +-1) optionally flips the columns of U_r; it exists only so
    tests can quantify the LAPACK sign ambiguity."""
    u_t, s_t, _ = np.linalg.svd(cov_t)
    u_r, s_r, _ = np.linalg.svd(cov_r)
    if signs is not None:
        u_r = u_r * np.asarray(signs, dtype=np.float64)[None, :]
    return u_t @ np.diag(1 / np.sqrt(s_t)) @ np.diag(np.sqrt(s_r)) @ np.linalg.inv(u_r)


def color_transfer_in_correlated_color_space(target, reference):
    """Xiao & Ma 2006.  ref: methods/linear.py:45-82 (applies ``@ T.T``)."""
    hw3 = target.shape
    t = target.reshape(-1, 3)
    r = reference.reshape(-1, 3)
    mu_t, cov_t = mean_and_cov(t)
    mu_r, cov_r = mean_and_cov(r)
    T = ccs_matrix(cov_t, cov_r)
    return ((t - mu_t) @ T.T + mu_r).reshape(hw3)


def mkl_matrix(cov_t, cov_r, decomposition="MK"):
    """The three closed forms of Pitie & Kokaram 2007.  ref: methods/linear.py:108-120."""
    if decomposition == "cholesky":
        return np.linalg.cholesky(cov_r) @ np.linalg.inv(np.linalg.cholesky(cov_t))
    if decomposition == "sqrt":
        return scipy.linalg.sqrtm(cov_r) @ np.linalg.inv(scipy.linalg.sqrtm(cov_t))
    if decomposition == "MK":
        root_t = scipy.linalg.sqrtm(cov_t)
        inv_root_t = np.linalg.inv(root_t)
        return inv_root_t @ scipy.linalg.sqrtm(root_t @ cov_r @ root_t) @ inv_root_t
    raise ValueError("Unknown decomposition, use either 'cholesky', 'sqrt', or 'MK'")


def monge_kantorovitch_color_transfer(target, reference, decomposition="MK"):
    """Pitie & Kokaram 2007.  ref: methods/linear.py:85-124 (applies ``@ T``)."""
    hw3 = target.shape
    t = target.reshape(-1, 3)
    r = reference.reshape(-1, 3)
    mu_t, cov_t = mean_and_cov(t)
    mu_r, cov_r = mean_and_cov(r)
    T = mkl_matrix(cov_t, cov_r, decomposition)
    return ((t - mu_t) @ T + mu_r).reshape(hw3)


# --------------------------------------------------------------------------- IDT stages


def draw_rotation(n_dims=3):
    """One Haar-random rotation from the GLOBAL numpy RNG.  ref: methods/iterative.py:32."""
    return scipy.stats.special_ortho_group.rvs(n_dims)


def project(rot, pixels):
    """rot @ pixels.T -> [3, N] float64.  ref: methods/iterative.py:34-35."""
    return rot @ pixels.T


def axis_range(p_t, p_r):
    """Shared range of one projected axis.  ref: methods/iterative.py:39-40."""
    return min(p_t.min(), p_r.min()), max(p_t.max(), p_r.max())


def axis_histograms(p_t, p_r, lo, hi, bins):
    """Integer counts of both images on the shared uniform grid.
    ref: methods/iterative.py:42-43."""
    c_t, edges = np.histogram(p_t, bins=bins, range=[lo, hi])
    c_r, _ = np.histogram(p_r, bins=bins, range=[lo, hi])
    return c_t, c_r, edges


def cdf(counts):
    """Normalised cumulative histogram.  ref: methods/iterative.py:45-49."""
    c = counts.cumsum().astype(float)
    c /= c[-1]
    return c


def inverse_cdf_lut(cdf_t, cdf_r, edges):
    """f[i] = value whose reference-CDF equals the target-CDF at the right edge of
    bin i.  ref: methods/iterative.py:51."""
    return np.interp(cdf_t, cdf_r, edges[1:])


def remap_axis(p_t, edges, lut, bins):
    """Per-sample lookup + linear interpolation; values left of edges[1] map to 0
    (the reference's first-bin behaviour).  ref: methods/iterative.py:53."""
    return np.interp(p_t, edges[1:], lut, left=0, right=bins)


def bin_index(p, edges):
    """The bin np.histogram puts each sample in (unique k with
    edges[k] <= x < edges[k+1], last bin closed).  Restates the uniform fast path of
    numpy/lib/_histograms_impl.py (estimate, clamp, -1/+1 correction)."""
    bins = len(edges) - 1
    lo, hi = edges[0], edges[-1]
    k = (((p - lo) / (hi - lo)) * bins).astype(np.intp)
    k[k == bins] -= 1
    k[p < edges[k]] -= 1
    k[(p >= edges[k + 1]) & (k != bins - 1)] += 1
    return k


def back_rotate(rot, moved, projected, state):
    """ref: methods/iterative.py:55."""
    return np.linalg.solve(rot, moved - projected).T + state


def idt_iteration(state, reference, rot, bins, moved_dtype, keep_arrays=True):
    """One full iteration; returns the new state and a dict of every intermediate
    (the per-pixel arrays only when ``keep_arrays``)."""
    p_t = project(rot, state)
    p_r = project(rot, reference)
    moved = np.empty(p_t.shape, dtype=moved_dtype)      # np.empty_like(target.T), :36
    trace = {"rot": rot, "lo": [], "hi": [], "edges": [], "counts_t": [], "counts_r": [],
             "lut": [], "proj_t": p_t}
    for j in range(p_t.shape[0]):
        lo, hi = axis_range(p_t[j], p_r[j])
        c_t, c_r, edges = axis_histograms(p_t[j], p_r[j], lo, hi, bins)
        lut = inverse_cdf_lut(cdf(c_t), cdf(c_r), edges)
        moved[j] = remap_axis(p_t[j], edges, lut, bins)
        for key, val in (("lo", lo), ("hi", hi), ("edges", edges), ("counts_t", c_t),
                         ("counts_r", c_r), ("lut", lut)):
            trace[key].append(val)
    new_state = back_rotate(rot, moved, p_t, state)
    for key in ("lo", "hi", "edges", "counts_t", "counts_r", "lut"):
        trace[key] = np.asarray(trace[key])
    if keep_arrays:
        trace["moved"] = moved
        trace["state"] = new_state
    else:
        del trace["proj_t"]
    return new_state, trace


def idt_instrumented(target, reference, bins=255, n_iter=4, rotations=None, keep_arrays=True):
    """IDT that also returns the per-iteration traces.  ``rotations`` (n_iter x 3 x 3)
    replaces the RNG draws when given (used when the matrices were pre-drawn)."""
    hw3 = target.shape
    state = target.reshape(-1, 3)
    ref = reference.reshape(-1, 3)
    traces = []
    for it in range(n_iter):
        rot = draw_rotation(hw3[-1]) if rotations is None else np.asarray(rotations[it])
        # the reference allocates the remapped values with the dtype of the CURRENT
        # state: float32 in iteration 0 for float32 input, float64 afterwards.
        state, tr = idt_iteration(state, ref, rot, bins, state.dtype, keep_arrays)
        traces.append(tr)
    return state.reshape(hw3), traces


def iterative_distribution_transfer(target, reference, bins=255, n_iter=4, rotations=None):
    """Pitie, Kokaram & Dahyot 2007.  ref: methods/iterative.py:8-59."""
    return idt_instrumented(target, reference, bins, n_iter, rotations, keep_arrays=False)[0]


# --------------------------------------------------------------------------- regrain (ACG)

REGRAIN_SWEEPS = (4, 16, 32, 64, 64, 64)


def _neighbours(a):
    """(next column, next row, previous column, previous row) with the edge replicated - the
    reference's last_pad_1 / last_pad_0 / first_pad_1 / first_pad_0 (iterative.py:87-90)."""
    p = np.pad(a, ((1, 1), (1, 1), (0, 0)), mode="edge")
    return p[1:-1, 2:], p[2:, 1:-1], p[1:-1, :-2], p[:-2, 1:-1]


def regrain_relax(out, src, col, sweeps, level, eps=1e-6):
    """`sweeps` Jacobi sweeps of the gradient-preserving relaxation.  ref: methods/iterative.py:80-115."""
    s_e, s_s, s_w, s_n = _neighbours(src)
    grad = np.sqrt(((s_e - s_w) ** 2 + (s_s - s_n) ** 2).sum(axis=2, keepdims=True))
    psi = 256 * grad / 5
    psi[psi > 1] = 1
    phi = 30 * 2 ** (-level) / (1 + 10 * grad)
    p_e, p_s, p_w, p_n = _neighbours(phi)
    phi1, phi2, phi3, phi4 = (p_e + phi) / 2, (p_s + phi) / 2, (p_w + phi) / 2, (p_n + phi) / 2
    rho = 1 / 5.0
    for _ in range(sweeps):
        o_e, o_s, o_w, o_n = _neighbours(out)
        den = psi + phi1 + phi2 + phi3 + phi4
        num = (psi * col + phi1 * (o_e - s_e + src) + phi2 * (o_s - s_s + src)
               + phi3 * (o_w - s_w + src) + phi4 * (o_n - s_n + src))
        out = num / (den + eps) * (1 - rho) + rho * out
    return out


def regrain(src, col, sweeps=REGRAIN_SWEEPS, level=0):
    """Multigrid regrain.  ref: methods/iterative.py:62-77."""
    h, w, _ = src.shape
    h2, w2 = (h + 1) // 2, (w + 1) // 2
    if len(sweeps) > 1 and h2 > 20 and w2 > 20:
        coarse = regrain(skimage_resize.resize(src, (h2, w2)), skimage_resize.resize(col, (h2, w2)), sweeps[1:], level + 1)
        start = skimage_resize.resize(coarse, (h, w))
    else:
        start = src
    return regrain_relax(start, src, col, sweeps[0], level)


def automated_color_grading(target, reference, rotations=None):
    """Pitie, Kokaram & Dahyot 2007.  ref: methods/iterative.py:118-138."""
    return regrain(target, iterative_distribution_transfer(target, reference, rotations=rotations))
