// Quality metrics on the device: the step after the hot path in the reference's
// test loop (methods/__init__.py:32-40), so that evaluation never leaves the GPU.
//   iCID  ref: utils/icid.py:28-152   (bilinear downscale, Lab, seven SSIM-like maps from 11x11
//                                      Gaussian statistics, 1 - mean of their product)
//   PSNR  ref: methods/__init__.py:35 (piq.psnr defaults)
//   SSIM  ref: methods/__init__.py:36 (piq.ssim defaults; kernels K11-K12 further down)
// Images are planar float32 [B,3,H,W] (the tensors the reference hands to these metrics).
//
//   K8  icid_premaps_kernel   downscale + Lab of both images -> 11 planes L1 L2 C1 C2 L1^2 L2^2
//                             C1^2 C2^2 sqrt(H) L1L2 C1C2            (24 B read, 44 B written / px)
//   K9  icid_maps_kernel      per 32x32 tile: separable 11-tap Gaussian (reflect padding) of the
//                             11 planes through shared memory, the seven maps, their product,
//                             one partial sum per block                (44 B read / px)
//   K10 finish kernels        fixed-order sums of the partials -> the scalar
// Arithmetic is fp32 like the reference's (torch float32); partial sums are fp64.
#include "ct_context.h"

namespace ct {

namespace {

__device__ __forceinline__ float lg2f_(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2f_(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// kornia.color.rgb_to_lab on one pixel (the constants of scikit-image's rgb2lab)
__device__ __forceinline__ void rgb2lab_px(float r, float g, float b, float &L, float &A, float &Bv) {
    float lin[3] = {r, g, b};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v = lin[c];
        const float u = fmaf(v, 1.0f / 1.055f, 0.055f / 1.055f);
        lin[c] = v > 0.04045f ? (u * u) * ex2f_(0.4f * lg2f_(u)) : v * (1.0f / 12.92f);
    }
    float t[3];
    t[0] = (0.412453f * lin[0] + 0.357580f * lin[1] + 0.180423f * lin[2]) * (1.0f / 0.95047f);
    t[1] = 0.212671f * lin[0] + 0.715160f * lin[1] + 0.072169f * lin[2];
    t[2] = (0.019334f * lin[0] + 0.119193f * lin[1] + 0.950227f * lin[2]) * (1.0f / 1.08883f);
    float f[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float x = t[c];
        const float z = ex2f_(-0.33333334f * lg2f_(fmaxf(x, 0.008856f)));
        const float zz = z * z;
        float y = x * zz;
        y = fmaf(fmaf(-y * y, y, x), zz * 0.33333334f, y);   // one Newton step on cbrt
        f[c] = x > 0.008856f ? y : fmaf(7.787f, x, 4.0f / 29.0f);
    }
    L = fmaf(116.0f, f[1], -16.0f);
    A = 500.0f * (f[0] - f[1]);
    Bv = 200.0f * (f[1] - f[2]);
}

struct PremapArgs {
    const float *img[2];   // [B][3][H][W]
    float *planes;         // [B][11][oh][ow]
    int H, W, oh, ow, f;
};

// torch.nn.functional.interpolate(scale_factor=1/f, mode="bilinear"): src = (dst + 0.5) f - 0.5
__device__ __forceinline__ float sample(const float *plane, int H, int W, int f, int oy, int ox) {
    if (f == 1) return plane[(int64_t)oy * W + ox];
    const float sy = (oy + 0.5f) * f - 0.5f, sx = (ox + 0.5f) * f - 0.5f;
    const int y0 = min((int)sy, H - 1), x0 = min((int)sx, W - 1);
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = sy - y0, lx = sx - x0;
    const float top = plane[(int64_t)y0 * W + x0] * (1.0f - lx) + plane[(int64_t)y0 * W + x1] * lx;
    const float bot = plane[(int64_t)y1 * W + x0] * (1.0f - lx) + plane[(int64_t)y1 * W + x1] * lx;
    return top * (1.0f - ly) + bot * ly;
}

__global__ void __launch_bounds__(256) icid_premaps_kernel(PremapArgs a) {
    const int ox = blockIdx.x * 32 + (threadIdx.x & 31), oy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (ox >= a.ow || oy >= a.oh) return;
    const int64_t b = blockIdx.z, in_plane = (int64_t)a.H * a.W, out_plane = (int64_t)a.oh * a.ow;
    float lab[2][3];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float *im = a.img[i] + b * 3 * in_plane;
        const float r = sample(im, a.H, a.W, a.f, oy, ox);
        const float g = sample(im + in_plane, a.H, a.W, a.f, oy, ox);
        const float bl = sample(im + 2 * in_plane, a.H, a.W, a.f, oy, ox);
        rgb2lab_px(r, g, bl, lab[i][0], lab[i][1], lab[i][2]);
    }
    const float L1 = lab[0][0], A1 = lab[0][1], B1 = lab[0][2], L2 = lab[1][0], A2 = lab[1][1], B2 = lab[1][2];
    const float c1sq = A1 * A1 + B1 * B1, c2sq = A2 * A2 + B2 * B2;
    const float C1 = sqrtf(c1sq), C2 = sqrtf(c2sq);
    const float dA = A1 - A2, dB = B1 - B2, dC = C1 - C2;
    const float hue = fmaxf(dA * dA + dB * dB - dC * dC, 0.0f);
    float *p = a.planes + b * 11 * out_plane + (int64_t)oy * a.ow + ox;
    p[0 * out_plane] = L1;
    p[1 * out_plane] = L2;
    p[2 * out_plane] = C1;
    p[3 * out_plane] = C2;
    p[4 * out_plane] = L1 * L1;
    p[5 * out_plane] = L2 * L2;
    p[6 * out_plane] = C1 * C1;
    p[7 * out_plane] = C2 * C2;
    p[8 * out_plane] = sqrtf(hue);
    p[9 * out_plane] = L1 * L2;
    p[10 * out_plane] = C1 * C2;
}

constexpr int kTile = 32, kRad = 5, kIn = kTile + 2 * kRad;   // 42

struct MapsArgs {
    const float *planes;   // [B][11][oh][ow]
    double *partials;      // [B * gridDim.y * gridDim.x]
    int oh, ow;
    float w[7];
    int expo67;            // exponent of maps 6 and 7 (0 when omitted)
    float k[11];           // normalised Gaussian taps
};

__device__ __forceinline__ int reflect(int i, int n) {   // torch "reflect" padding (no edge repeat)
    i = i < 0 ? -i : i;
    return i >= n ? 2 * n - 2 - i : i;
}

__global__ void __launch_bounds__(256) icid_maps_kernel(MapsArgs a) {
    __shared__ float tile[kIn][kIn + 1];
    __shared__ float mid[kIn][kTile + 1];
    __shared__ double red[8];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8 threads, 4 rows each
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
    const int64_t plane = (int64_t)a.oh * a.ow;
    const float *base = a.planes + (int64_t)blockIdx.z * 11 * plane;
    float v[4][11];
    for (int q = 0; q < 11; ++q) {
        const float *src = base + q * plane;
        for (int i = threadIdx.x; i < kIn * kIn; i += 256) {
            const int r = i / kIn, c = i % kIn;
            const int yy = reflect(min(y0 + r - kRad, a.oh - 1 + kRad), a.oh);
            const int xx = reflect(min(x0 + c - kRad, a.ow - 1 + kRad), a.ow);
            tile[r][c] = src[(int64_t)yy * a.ow + xx];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < kIn * kTile; i += 256) {   // horizontal pass
            const int r = i / kTile, c = i % kTile;
            float s = 0.0f;
#pragma unroll
            for (int j = 0; j < 11; ++j) s = fmaf(a.k[j], tile[r][c + j], s);
            mid[r][c] = s;
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < 4; ++m) {                            // vertical pass
            const int r = ty + 8 * m;
            float s = 0.0f;
#pragma unroll
            for (int j = 0; j < 11; ++j) s = fmaf(a.k[j], mid[r + j][tx], s);
            v[m][q] = s;
        }
        __syncthreads();
    }
    double acc = 0.0;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int y = y0 + ty + 8 * m, x = x0 + tx;
        if (y >= a.oh || x >= a.ow) continue;
        const float muL1 = v[m][0], muL2 = v[m][1], muC1 = v[m][2], muC2 = v[m][3];
        const float vL1 = fmaxf(v[m][4] - muL1 * muL1, 0.0f), vL2 = fmaxf(v[m][5] - muL2 * muL2, 0.0f);
        const float vC1 = fmaxf(v[m][6] - muC1 * muC1, 0.0f), vC2 = fmaxf(v[m][7] - muC2 * muC2, 0.0f);
        const float sL1 = sqrtf(vL1), sL2 = sqrtf(vL2), sC1 = sqrtf(vC1), sC2 = sqrtf(vC2);
        const float dL = (muL1 - muL2) * (muL1 - muL2), dC = (muC1 - muC2) * (muC1 - muC2);
        const float dH = v[m][8] * v[m][8];
        const float sL12 = v[m][9] - muL1 * muL2, sC12 = v[m][10] - muC1 * muC2;
        const float m1 = 1.0f / (a.w[0] * dL + 1.0f);
        const float m2 = (a.w[1] + 2.0f * sL1 * sL2) / (a.w[1] + vL1 + vL2);
        const float m3 = (a.w[2] + fabsf(sL12)) / (a.w[2] + sL1 * sL2);
        const float m4 = 1.0f / (a.w[3] * dC + 1.0f);
        const float m5 = 1.0f / (a.w[4] * dH + 1.0f);
        float prod = m1 * m2 * (m3 * m3 * m3) * m4 * m5;        // alpha = 3 on the structure map
        if (a.expo67) {
            const float m6 = (a.w[5] + 2.0f * sC1 * sC2) / (a.w[5] + sC1 * sC1 + sC2 * sC2);
            const float m7 = (a.w[6] + fabsf(sC12)) / (a.w[6] + sC1 * sC2);
            prod *= m6 * m7;
        }
        acc += (double)prod;
    }
    acc = warp_sum(acc);
    if (tx == 0) red[ty] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i];
        a.partials[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
    }
}

// out[0] = 1 - sum(partials) / count, summed in a fixed order by one warp
__global__ void icid_finish_kernel(const double *partials, int64_t n, double count, double *out) {
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 32) s += partials[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) out[0] = 1.0 - s / count;
}

__global__ void __launch_bounds__(256) sqdiff_kernel(const float *x, const float *y, int64_t n, double *partials) {
    const float *px = x + (int64_t)blockIdx.y * n, *py = y + (int64_t)blockIdx.y * n;
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const float d = px[i] - py[i];
        s += (double)(d * d);
    }
    __shared__ double red[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i];
        partials[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}

// piq.psnr: mean over images of -10 log10(mse + 1e-8)
__global__ void psnr_finish_kernel(const double *partials, int nblk, int B, double n, double *out) {
    double total = 0.0;
    for (int b = 0; b < B; ++b) {
        double s = 0.0;
        for (int i = threadIdx.x; i < nblk; i += 32) s += partials[(int64_t)b * nblk + i];
        s = warp_sum(s);
        total += -10.0 * log10(s / n + 1e-8);
    }
    if (threadIdx.x == 0) out[0] = total / B;
}

// ---------------------------------------------------------------------------------------------
// SSIM (piq.ssim defaults, ref: methods/__init__.py:36): average-pool downscale, then per channel
// the 11x11 Gaussian (sigma 1.5) statistics by VALID convolution and the mean of the SSIM map.
//   K11 avgpool_kernel     [B,3,H,W] -> [B,3,oh,ow], f x f blocks (only when f > 1)
//   K12 ssim_maps_kernel   per 32x32 output tile and plane: x and y tiles in shared memory, the
//                          five statistics (x, y, xx, yy, xy) through one separable pass each,
//                          SSIM per pixel, one partial sum per block
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) avgpool_kernel(const float *in, float *out, int H, int W, int oh, int ow, int f) {
    const int ox = blockIdx.x * 32 + (threadIdx.x & 31), oy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (ox >= ow || oy >= oh) return;
    const float *p = in + (int64_t)blockIdx.z * H * W + (int64_t)oy * f * W + (int64_t)ox * f;
    float s = 0.0f;
    for (int r = 0; r < f; ++r)
        for (int c = 0; c < f; ++c) s += p[(int64_t)r * W + c];
    out[(int64_t)blockIdx.z * oh * ow + (int64_t)oy * ow + ox] = s / (float)(f * f);
}

struct SsimArgs {
    const float *x, *y;    // [planes][h][w]
    double *partials;      // [planes * gridDim.y * gridDim.x]
    int h, w;              // plane size; outputs are (h - 10) x (w - 10)
    float c1, c2;
    float k[11];
};

__global__ void __launch_bounds__(256) ssim_maps_kernel(SsimArgs a) {
    __shared__ float tx[kIn][kIn + 1], ty[kIn][kIn + 1];
    __shared__ float mid[kIn][kTile + 1];
    __shared__ double red[8];
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
    const int64_t plane = (int64_t)a.h * a.w;
    const float *px = a.x + blockIdx.z * plane, *py = a.y + blockIdx.z * plane;
    for (int i = threadIdx.x; i < kIn * kIn; i += 256) {
        const int r = i / kIn, c = i % kIn;
        const int yy = min(y0 + r, a.h - 1), xx = min(x0 + c, a.w - 1);   // beyond the image: unused outputs
        tx[r][c] = px[(int64_t)yy * a.w + xx];
        ty[r][c] = py[(int64_t)yy * a.w + xx];
    }
    __syncthreads();
    float v[4][5];
    for (int q = 0; q < 5; ++q) {
        for (int i = threadIdx.x; i < kIn * kTile; i += 256) {
            const int r = i / kTile, c = i % kTile;
            float s = 0.0f;
#pragma unroll
            for (int j = 0; j < 11; ++j) {
                const float xv = tx[r][c + j], yv = ty[r][c + j];
                const float t = q == 0 ? xv : q == 1 ? yv : q == 2 ? xv * xv : q == 3 ? yv * yv : xv * yv;
                s = fmaf(a.k[j], t, s);
            }
            mid[r][c] = s;
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int r = ly + 8 * m;
            float s = 0.0f;
#pragma unroll
            for (int j = 0; j < 11; ++j) s = fmaf(a.k[j], mid[r + j][lx], s);
            v[m][q] = s;
        }
        __syncthreads();
    }
    double acc = 0.0;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int y = y0 + ly + 8 * m, x = x0 + lx;
        if (y >= a.h - 10 || x >= a.w - 10) continue;
        const float mx = v[m][0], my = v[m][1];
        const float sxx = v[m][2] - mx * mx, syy = v[m][3] - my * my, sxy = v[m][4] - mx * my;
        const float cs = (2.0f * sxy + a.c2) / (sxx + syy + a.c2);
        acc += (double)((2.0f * mx * my + a.c1) / (mx * mx + my * my + a.c1) * cs);
    }
    acc = warp_sum(acc);
    if (lx == 0) red[ly] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i];
        a.partials[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
    }
}

// out[0] = sum(partials) / count (every channel and image has the same number of map pixels, so
// the mean of per-channel means is the overall mean)
__global__ void mean_finish_kernel(const double *partials, int64_t n, double count, double *out) {
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 32) s += partials[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) out[0] = s / count;
}

}  // namespace

int launch_icid(ct_context *h, const float *img1, const float *img2, int B, int H, int W, int intent,
                int omit_maps67, int downsampling, double *out_dev) {
    if (!img1 || !img2 || !out_dev) return fail(h, CT_E_INVALID, "icid: NULL argument");
    if (B < 1 || H < 1 || W < 1 || B > 65535) return fail(h, CT_E_INVALID, "icid: bad shape");
    if (intent < 0 || intent > 2) return fail(h, CT_E_INVALID, "Intent should be either 'perceptual', 'hue-preserving', or 'chromatic'");
    int f = 1;
    if (downsampling) {
        f = (int)nearbyint((double)(H < W ? H : W) / 256.0);   // Python round(): half to even
        if (f < 1) f = 1;
    }
    const int oh = f > 1 ? (int)floor((double)H * (1.0 / f)) : H, ow = f > 1 ? (int)floor((double)W * (1.0 / f)) : W;
    if (oh <= kRad || ow <= kRad) return fail(h, CT_E_UNSUPPORTED, "icid: image smaller than the blur radius after downscaling");
    const dim3 gmaps((ow + kTile - 1) / kTile, (oh + kTile - 1) / kTile, B);
    const int64_t nparts = (int64_t)gmaps.x * gmaps.y * gmaps.z;
    const size_t planes_bytes = (size_t)B * 11 * oh * ow * sizeof(float);
    CT_TRY(ensure_ws(h, planes_bytes + (size_t)nparts * sizeof(double) + 256));
    float *planes = static_cast<float *>(h->ws);
    double *partials = reinterpret_cast<double *>(static_cast<unsigned char *>(h->ws) + ((planes_bytes + 255) & ~(size_t)255));
    PremapArgs p{{img1, img2}, planes, H, W, oh, ow, f};
    icid_premaps_kernel<<<dim3((ow + 31) / 32, (oh + 7) / 8, B), 256, 0, h->stream>>>(p);
    MapsArgs m{};
    m.planes = planes;
    m.partials = partials;
    m.oh = oh;
    m.ow = ow;
    const float wts[3][7] = {{0.002f, 10, 10, 0.002f, 0.002f, 10, 10},
                             {0.002f, 10, 10, 0.002f, 0.02f, 10, 10},
                             {0.002f, 10, 10, 0.02f, 0.02f, 10, 10}};
    for (int i = 0; i < 7; ++i) m.w[i] = wts[intent][i];
    m.expo67 = omit_maps67 ? 0 : 1;
    double k[11], ksum = 0.0;   // torchvision _get_gaussian_kernel1d(11, 2.0)
    for (int i = 0; i < 11; ++i) {
        const double t = (i - 5) / 2.0;
        k[i] = exp(-0.5 * t * t);
        ksum += k[i];
    }
    for (int i = 0; i < 11; ++i) m.k[i] = (float)(k[i] / ksum);
    icid_maps_kernel<<<gmaps, 256, 0, h->stream>>>(m);
    icid_finish_kernel<<<1, 32, 0, h->stream>>>(partials, nparts, (double)B * oh * ow, out_dev);
    h->launches += 3;
    CT_CUDA(h, cudaGetLastError());
    return CT_OK;
}

int launch_psnr(ct_context *h, const float *x, const float *y, int B, int64_t n, double *out_dev) {
    if (!x || !y || !out_dev) return fail(h, CT_E_INVALID, "psnr: NULL argument");
    if (B < 1 || n < 1 || B > 65535) return fail(h, CT_E_INVALID, "psnr: bad shape");
    int64_t nblk = (n + 256 * 8 - 1) / (256 * 8);
    const int64_t cap = (int64_t)h->sm_count * 8 / B > 0 ? (int64_t)h->sm_count * 8 / B : 1;
    if (nblk > cap) nblk = cap;
    CT_TRY(ensure_ws(h, (size_t)B * nblk * sizeof(double)));
    double *partials = static_cast<double *>(h->ws);
    sqdiff_kernel<<<dim3((unsigned)nblk, B), 256, 0, h->stream>>>(x, y, n, partials);
    psnr_finish_kernel<<<1, 32, 0, h->stream>>>(partials, (int)nblk, B, (double)n, out_dev);
    h->launches += 2;
    CT_CUDA(h, cudaGetLastError());
    return CT_OK;
}

int launch_ssim(ct_context *h, const float *x, const float *y, int B, int H, int W, int downsample, double *out_dev) {
    if (!x || !y || !out_dev) return fail(h, CT_E_INVALID, "ssim: NULL argument");
    if (B < 1 || H < 1 || W < 1 || (int64_t)B * 3 > 65535) return fail(h, CT_E_INVALID, "ssim: bad shape");
    int f = (int)nearbyint((double)(H < W ? H : W) / 256.0);   // Python round(): half to even
    if (f < 1 || !downsample) f = 1;
    const int oh = H / f, ow = W / f, planes = 3 * B;
    if (oh < 11 || ow < 11) return fail(h, CT_E_UNSUPPORTED, "ssim: image smaller than the 11x11 window after downscaling");
    const dim3 grid((ow - 10 + kTile - 1) / kTile, (oh - 10 + kTile - 1) / kTile, planes);
    const int64_t nparts = (int64_t)grid.x * grid.y * grid.z;
    const size_t pooled = f > 1 ? (((size_t)planes * oh * ow * sizeof(float) + 255) & ~(size_t)255) : 0;
    CT_TRY(ensure_ws(h, 2 * pooled + (size_t)nparts * sizeof(double)));
    unsigned char *ws = static_cast<unsigned char *>(h->ws);
    const float *sx = x, *sy = y;
    int launches = 2;
    if (f > 1) {
        float *qx = reinterpret_cast<float *>(ws), *qy = reinterpret_cast<float *>(ws + pooled);
        const dim3 g((ow + 31) / 32, (oh + 7) / 8, planes);
        avgpool_kernel<<<g, 256, 0, h->stream>>>(x, qx, H, W, oh, ow, f);
        avgpool_kernel<<<g, 256, 0, h->stream>>>(y, qy, H, W, oh, ow, f);
        sx = qx;
        sy = qy;
        launches += 2;
    }
    SsimArgs a{};
    a.x = sx;
    a.y = sy;
    a.partials = reinterpret_cast<double *>(ws + 2 * pooled);
    a.h = oh;
    a.w = ow;
    a.c1 = 0.01f * 0.01f;
    a.c2 = 0.03f * 0.03f;
    double k[11], ksum = 0.0;   // piq gaussian_filter(11, 1.5): separable, normalised
    for (int i = 0; i < 11; ++i) {
        k[i] = exp(-(double)((i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5));
        ksum += k[i];
    }
    for (int i = 0; i < 11; ++i) a.k[i] = (float)(k[i] / ksum);
    ssim_maps_kernel<<<grid, 256, 0, h->stream>>>(a);
    mean_finish_kernel<<<1, 32, 0, h->stream>>>(a.partials, nparts, (double)planes * (oh - 10) * (ow - 10), out_dev);
    h->launches += launches;
    CT_CUDA(h, cudaGetLastError());
    return CT_OK;
}

}  // namespace ct
