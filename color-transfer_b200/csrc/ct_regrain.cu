// Regrain (Pitie et al. 2007, section 5): the gradient-preserving multigrid relaxation that turns
// the IDT result into `automated_color_grading` (ref: methods/iterative.py:62-138).  SURVEY.md
// section 8f-2 ("next" row): every step is a streaming fp64 kernel over [H,W,3] images.
//
//   _regrain  (iterative.py:62-77)   pyramid: resize both images to ceil(h/2) x ceil(w/2) while that
//                                    is > 20 x 20 (at most 6 levels), recurse, resize the result back
//   resize    skimage.transform.resize defaults = scipy.ndimage.gaussian_filter (sigma = (f-1)/2,
//             truncate 4, mode 'mirror', axis 0 then axis 1) when shrinking, then
//             scipy.ndimage.zoom(order=1, mode='mirror', grid_mode=True), then clip to the input's
//             [min, max]   (restated in oracle/skimage_resize.py; formulas checked against scipy to 3e-16)
//   _solve    (iterative.py:80-115)  nbit Jacobi sweeps of the 5-point relaxation with
//             psi = min(256 |grad in| / 5, 1), phi = 30 * 2^-level / (1 + 10 |grad in|)
#include <vector>

#include "ct_context.h"

namespace ct {

namespace {

__device__ __forceinline__ int mirror_index(int i, int n) {  // scipy 'mirror': d c b | a b c d | c b a
    if (n == 1) return 0;
    const int p = 2 * (n - 1);
    i = i < 0 ? -i : i;
    i %= p;
    return i > n - 1 ? p - i : i;
}

// ---- scipy.ndimage.gaussian_filter1d along one axis of an interleaved [H][W][3] image
__global__ void __launch_bounds__(256) gauss_axis_kernel(const double *__restrict__ in, double *__restrict__ out, int H, int W,
                                                         int axis, double sigma) {
    const int64_t n = (int64_t)H * W;
    const int lw = (int)(4.0 * sigma + 0.5);
    double wsum = 0.0;
    for (int k = -lw; k <= lw; ++k) wsum += exp(-0.5 / (sigma * sigma) * (double)(k * k));
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
        double acc[3] = {0.0, 0.0, 0.0};
        for (int k = -lw; k <= lw; ++k) {
            const double w = exp(-0.5 / (sigma * sigma) * (double)(k * k)) / wsum;
            const int64_t q = axis == 0 ? (int64_t)mirror_index(y + k, H) * W + x : (int64_t)y * W + mirror_index(x + k, W);
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[c] = fma(w, in[3 * q + c], acc[c]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) out[3 * p + c] = acc[c];
    }
}

// ---- min / max of all elements (the clip range of resize), folded into two monotone keys
__global__ void __launch_bounds__(256) minmax_kernel(const double *__restrict__ in, int64_t n, int64_t *keys) {
    double lo = INFINITY, nhi = INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = in[i];
        lo = v < lo ? v : lo;
        nhi = -v < nhi ? -v : nhi;
    }
    lo = warp_min(lo);
    nhi = warp_min(nhi);
    if ((threadIdx.x & 31) == 0) {
        if (lo < INFINITY) atomicMin(reinterpret_cast<long long *>(keys), (long long)key_of(lo));
        if (nhi < INFINITY) atomicMin(reinterpret_cast<long long *>(keys) + 1, (long long)key_of(nhi));
    }
}
__global__ void keys2_init_kernel(int64_t *keys) {
    if (threadIdx.x < 2) keys[threadIdx.x] = kKeyPlusInf;
}

// ---- scipy.ndimage.zoom(order=1, mode='mirror', grid_mode=True) + clip to [lo, hi]
__device__ __forceinline__ void zoom_coord(int o, int n_in, int n_out, int &i0, int &i1, double &w) {
    double c = ((double)o + 0.5) * ((double)n_in / (double)n_out) - 0.5;
    if (n_in > 1) {
        const double p = 2.0 * (double)(n_in - 1);
        c = fabs(c);
        c = fmod(c, p);
        c = c > (double)(n_in - 1) ? p - c : c;
    } else {
        c = 0.0;
    }
    const double f = floor(c);
    i0 = (int)f;
    w = c - f;
    i1 = mirror_index(i0 + 1, n_in);
}
__global__ void __launch_bounds__(256) zoom_kernel(const double *__restrict__ in, int H, int W, double *__restrict__ out, int h, int w,
                                                   const int64_t *clip_keys) {
    const double lo = value_of(clip_keys[0]), hi = -value_of(clip_keys[1]);
    const int64_t n = (int64_t)h * w;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(p / w), x = (int)(p - (int64_t)y * w);
        int i0, i1, j0, j1;
        double a, b;
        zoom_coord(y, H, h, i0, i1, a);
        zoom_coord(x, W, w, j0, j1, b);
        const double w00 = (1.0 - a) * (1.0 - b), w01 = (1.0 - a) * b, w10 = a * (1.0 - b), w11 = a * b;
        const double *p00 = in + 3 * ((int64_t)i0 * W + j0), *p01 = in + 3 * ((int64_t)i0 * W + j1);
        const double *p10 = in + 3 * ((int64_t)i1 * W + j0), *p11 = in + 3 * ((int64_t)i1 * W + j1);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double v = w00 * p00[c] + w01 * p01[c] + w10 * p10[c] + w11 * p11[c];
            v = v < lo ? lo : (v > hi ? hi : v);
            out[3 * p + c] = v;
        }
    }
}

// ---- _solve, part 1: psi and phi from the gradient of `in` (iterative.py:91-96)
__global__ void __launch_bounds__(256) solve_setup_kernel(const double *__restrict__ in, int H, int W, double level_scale,
                                                          double *__restrict__ psi, double *__restrict__ phi) {
    const int64_t n = (int64_t)H * W;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
        const int64_t xm = (int64_t)y * W + max(x - 1, 0), xp = (int64_t)y * W + min(x + 1, W - 1);   // first_pad_1 / last_pad_1
        const int64_t ym = (int64_t)max(y - 1, 0) * W + x, yp = (int64_t)min(y + 1, H - 1) * W + x;   // first_pad_0 / last_pad_0
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double dx = in[3 * xp + c] - in[3 * xm + c], dy = in[3 * yp + c] - in[3 * ym + c];
            s += dx * dx + dy * dy;
        }
        const double delta = sqrt(s);
        const double ps = 256.0 * delta / 5.0;
        psi[p] = ps > 1.0 ? 1.0 : ps;
        phi[p] = level_scale / (1.0 + 10.0 * delta);
    }
}

// ---- _solve, part 2: one Jacobi sweep (iterative.py:103-113)
__global__ void __launch_bounds__(256) solve_iter_kernel(const double *__restrict__ prev, const double *__restrict__ in,
                                                         const double *__restrict__ col, const double *__restrict__ psi,
                                                         const double *__restrict__ phi, int H, int W, double *__restrict__ next) {
    const int64_t n = (int64_t)H * W;
    const double rho = 1.0 / 5.0, eps = 1e-6;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
        const int64_t q1 = (int64_t)y * W + min(x + 1, W - 1);       // last_pad_1
        const int64_t q2 = (int64_t)min(y + 1, H - 1) * W + x;       // last_pad_0
        const int64_t q3 = (int64_t)y * W + max(x - 1, 0);           // first_pad_1
        const int64_t q4 = (int64_t)max(y - 1, 0) * W + x;           // first_pad_0
        const double ph = phi[p], ps = psi[p];
        const double phi1 = (phi[q1] + ph) / 2, phi2 = (phi[q2] + ph) / 2, phi3 = (phi[q3] + ph) / 2, phi4 = (phi[q4] + ph) / 2;
        const double den = ps + phi1 + phi2 + phi3 + phi4;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double ic = in[3 * p + c];
            const double num = ps * col[3 * p + c] + phi1 * (prev[3 * q1 + c] - in[3 * q1 + c] + ic) +
                               phi2 * (prev[3 * q2 + c] - in[3 * q2 + c] + ic) + phi3 * (prev[3 * q3 + c] - in[3 * q3 + c] + ic) +
                               phi4 * (prev[3 * q4 + c] - in[3 * q4 + c] + ic);
            next[3 * p + c] = num / (den + eps) * (1 - rho) + rho * prev[3 * p + c];
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) to_f64_hwc_kernel(const T *__restrict__ in, int64_t npix, int64_t plane, int chw, double *__restrict__ out) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x)
#pragma unroll
        for (int c = 0; c < 3; ++c) out[3 * p + c] = (double)(chw ? in[c * plane + p] : in[3 * p + c]);
}

struct Level {
    int H, W;
    double *in, *col, *out, *tmp, *psi, *phi;  // out/tmp ping-pong during the sweeps
};

int grid_for(const ct_context *h, int64_t n) {
    int64_t g = (n + 255) / 256;
    const int64_t cap = (int64_t)h->sm_count * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int resize_dev(ct_context *h, const double *src, int H, int W, double *dst, int hh, int ww, double *scratch_a, double *scratch_b,
               int64_t *clip_keys) {
    // clip range = min / max of the ORIGINAL input of this resize call
    keys2_init_kernel<<<1, 32, 0, h->stream>>>(clip_keys);
    minmax_kernel<<<grid_for(h, (int64_t)H * W * 3), 256, 0, h->stream>>>(src, (int64_t)H * W * 3, clip_keys);
    h->launches += 2;
    const double *cur = src;
    if (hh < H || ww < W) {  // anti-aliasing only when an axis shrinks; sigma = max(0, (factor - 1) / 2)
        const double s0 = ((double)H / hh - 1.0) / 2.0, s1 = ((double)W / ww - 1.0) / 2.0;
        if (s0 > 1e-15) {
            gauss_axis_kernel<<<grid_for(h, (int64_t)H * W), 256, 0, h->stream>>>(cur, scratch_a, H, W, 0, s0);
            h->launches++;
            cur = scratch_a;
        }
        if (s1 > 1e-15) {
            gauss_axis_kernel<<<grid_for(h, (int64_t)H * W), 256, 0, h->stream>>>(cur, scratch_b, H, W, 1, s1);
            h->launches++;
            cur = scratch_b;
        }
    }
    zoom_kernel<<<grid_for(h, (int64_t)hh * ww), 256, 0, h->stream>>>(cur, H, W, dst, hh, ww, clip_keys);
    h->launches++;
    CT_CUDA(h, cudaGetLastError());
    return CT_OK;
}

}  // namespace

size_t regrain_workspace_bytes(int H, int W) {
    size_t total = 0;
    int hh = H, ww = W;
    for (int lvl = 0; lvl < 6; ++lvl) {
        const size_t px = (size_t)hh * ww;
        total += (4 * 3 * px + 2 * px) * sizeof(double) + 6 * 256;   // in, col, out, tmp (x3) + psi, phi
        const int h2 = (hh + 1) / 2, w2 = (ww + 1) / 2;
        if (!(lvl + 1 < 6 && h2 > 20 && w2 > 20)) break;
        hh = h2;
        ww = w2;
    }
    total += 2 * (size_t)H * W * 3 * sizeof(double) + 1024;  // resize scratch a / b + clip keys
    return total;
}

// in / col: device fp64 [H][W][3]; result written to `out` (fp64 [H][W][3]); asynchronous.
int launch_regrain(ct_context *h, const double *in, const double *col, double *out, int H, int W, void *workspace,
                   size_t workspace_bytes) {
    if (!in || !col || !out || H <= 0 || W <= 0) return fail(h, CT_E_INVALID, "bad regrain arguments");
    if (workspace_bytes < regrain_workspace_bytes(H, W)) return fail(h, CT_E_NOMEM, "regrain workspace too small");
    static const int nbits[6] = {4, 16, 32, 64, 64, 64};
    unsigned char *base = static_cast<unsigned char *>(workspace);
    size_t off = 0;
    auto take = [&](size_t doubles) {
        double *p = reinterpret_cast<double *>(base + off);
        off += (doubles * sizeof(double) + 255) / 256 * 256;
        return p;
    };
    std::vector<Level> lv;
    int hh = H, ww = W;
    for (int lvl = 0; lvl < 6; ++lvl) {
        const size_t px = (size_t)hh * ww;
        Level L{hh, ww, take(3 * px), take(3 * px), take(3 * px), take(3 * px), take(px), take(px)};
        lv.push_back(L);
        const int h2 = (hh + 1) / 2, w2 = (ww + 1) / 2;
        if (!(lvl + 1 < 6 && h2 > 20 && w2 > 20)) break;
        hh = h2;
        ww = w2;
    }
    double *sa = take((size_t)H * W * 3), *sb = take((size_t)H * W * 3);
    int64_t *clip = reinterpret_cast<int64_t *>(take(32));
    const int n_levels = (int)lv.size();
    // level 0 holds the caller's images
    CT_CUDA(h, cudaMemcpyAsync(lv[0].in, in, sizeof(double) * 3 * (size_t)H * W, cudaMemcpyDeviceToDevice, h->stream));
    CT_CUDA(h, cudaMemcpyAsync(lv[0].col, col, sizeof(double) * 3 * (size_t)H * W, cudaMemcpyDeviceToDevice, h->stream));
    for (int l = 1; l < n_levels; ++l) {  // build the pyramids top-down (iterative.py:68-69)
        CT_TRY(resize_dev(h, lv[l - 1].in, lv[l - 1].H, lv[l - 1].W, lv[l].in, lv[l].H, lv[l].W, sa, sb, clip));
        CT_TRY(resize_dev(h, lv[l - 1].col, lv[l - 1].H, lv[l - 1].W, lv[l].col, lv[l].H, lv[l].W, sa, sb, clip));
    }
    for (int l = n_levels - 1; l >= 0; --l) {  // solve bottom-up
        Level &L = lv[l];
        const size_t px = (size_t)L.H * L.W;
        if (l == n_levels - 1) {  // coarsest level starts from `in` (iterative.py:73-74)
            CT_CUDA(h, cudaMemcpyAsync(L.out, L.in, sizeof(double) * 3 * px, cudaMemcpyDeviceToDevice, h->stream));
        } else {                  // otherwise from the resized coarser result (iterative.py:71)
            CT_TRY(resize_dev(h, lv[l + 1].out, lv[l + 1].H, lv[l + 1].W, L.out, L.H, L.W, sa, sb, clip));
        }
        const int g = grid_for(h, (int64_t)px);
        solve_setup_kernel<<<g, 256, 0, h->stream>>>(L.in, L.H, L.W, 30.0 * exp2(-(double)l), L.psi, L.phi);
        h->launches++;
        double *prev = L.out, *next = L.tmp;
        for (int it = 0; it < nbits[l]; ++it) {
            solve_iter_kernel<<<g, 256, 0, h->stream>>>(prev, L.in, L.col, L.psi, L.phi, L.H, L.W, next);
            h->launches++;
            double *t = prev;
            prev = next;
            next = t;
        }
        if (prev != L.out) {  // keep the level's result in L.out for the next resize
            double *t = L.out;
            L.out = prev;
            L.tmp = t;
        }
    }
    CT_CUDA(h, cudaMemcpyAsync(out, lv[0].out, sizeof(double) * 3 * (size_t)H * W, cudaMemcpyDeviceToDevice, h->stream));
    CT_CUDA(h, cudaGetLastError());
    return CT_OK;
}

int launch_to_f64_hwc(ct_context *h, const ct_batch *img, double *out) {
    CT_TRY(check_batch(h, img, "image"));
    const int g = grid_for(h, img->npix);
    if (img->dtype == CT_F32)
        to_f64_hwc_kernel<float><<<g, 256, 0, h->stream>>>(static_cast<const float *>(img->data), img->npix, plane_of(img), img->layout == CT_CHW, out);
    else
        to_f64_hwc_kernel<double><<<g, 256, 0, h->stream>>>(static_cast<const double *>(img->data), img->npix, plane_of(img), img->layout == CT_CHW, out);
    h->launches++;
    CT_CUDA(h, cudaGetLastError());
    return CT_OK;
}

}  // namespace ct
