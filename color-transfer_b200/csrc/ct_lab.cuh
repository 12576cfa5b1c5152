// sRGB <-> CIE-Lab (D65, 2 degree) per-pixel math, the published scikit-image algorithm that
// the reference calls at methods/linear.py:25,26,40 (rgb2xyz, xyz2lab, _lab2xyz, xyz2rgb).
//
// The three fractional powers (x^2.4, cbrt, x^(1/2.4)) are the cost of this path.  Each one is
// seeded with MUFU lg2/ex2 in fp32 (relative error ~3e-7) and polished by ONE Newton step whose
// residual is evaluated in fp64 with an FMA, which squares the error (~1e-12): fp64-grade
// results for ~10 DP instructions instead of the ~150 of pow().  The division inside the
// Newton step only scales an already tiny correction, so it is an fp32 reciprocal.
#pragma once

#include "ct_common.cuh"

namespace ct {
namespace lab {

__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// u^2.4 for u > 0:  u^2 * w,  w = u^0.4 solves w^5 = u^2.
__device__ __forceinline__ double pow_2p4(double u) {
    const double u2 = u * u;
    const float w0 = ex2_approx(0.4f * lg2_approx((float)u));
    const double w = (double)w0;
    const double w2 = w * w, w4 = w2 * w2;
    const double resid = fma(w4, w, -u2);                // w^5 - u^2
    const float inv = 0.2f * rcp_approx((float)w4);      // 1 / (5 w^4)
    return u2 * fma(-resid, (double)inv, w);
}

// t^(1/3) for t > 0:  c solves c^3 = t.
__device__ __forceinline__ double cbrt_pos(double t) {
    const float c0 = ex2_approx(0.33333334f * lg2_approx((float)t));
    const double c = (double)c0;
    const double c2 = c * c;
    const double resid = fma(c2, c, -t);                 // c^3 - t
    const float inv = 0.33333334f * rcp_approx(c0 * c0); // 1 / (3 c^2)
    return fma(-resid, (double)inv, c);
}

// c^(1/2.4) = c^(5/12) for c > 0:  v solves v^12 = c^5.
__device__ __forceinline__ double pow_5_12(double c) {
    const float v0 = ex2_approx(0.41666666f * lg2_approx((float)c));
    const double v = (double)v0;
    const double v2 = v * v, v4 = v2 * v2, v8 = v4 * v4;
    const double c2 = c * c, c5 = c2 * c2 * c;
    const double resid = fma(v8, v4, -c5);               // v^12 - c^5
    const float v2f = v0 * v0, v4f = v2f * v2f, v8f = v4f * v4f;
    const float inv = 0.083333336f * rcp_approx(v8f * v2f * v0);  // 1 / (12 v^11)
    return fma(-resid, (double)inv, v);
}

// xyz_from_rgb with each row divided by the D65 white (0.95047, 1, 1.08883)
__device__ __forceinline__ void rgb2lab(const double (&rgb)[3], double (&out)[3]) {
    double lin[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double v = rgb[c];
        lin[c] = v > 0.04045 ? pow_2p4((v + 0.055) * (1.0 / 1.055)) : v * (1.0 / 12.92);
    }
    const double x = (0.412453 / 0.95047) * lin[0] + (0.357580 / 0.95047) * lin[1] + (0.180423 / 0.95047) * lin[2];
    const double y = 0.212671 * lin[0] + 0.715160 * lin[1] + 0.072169 * lin[2];
    const double z = (0.019334 / 1.08883) * lin[0] + (0.119193 / 1.08883) * lin[1] + (0.950227 / 1.08883) * lin[2];
    const double fx = x > 0.008856 ? cbrt_pos(x) : fma(7.787, x, 16.0 / 116.0);
    const double fy = y > 0.008856 ? cbrt_pos(y) : fma(7.787, y, 16.0 / 116.0);
    const double fz = z > 0.008856 ? cbrt_pos(z) : fma(7.787, z, 16.0 / 116.0);
    out[0] = fma(116.0, fy, -16.0);
    out[1] = 500.0 * (fx - fy);
    out[2] = 200.0 * (fy - fz);
}

// scipy.linalg.inv(xyz_from_rgb) (what skimage computes at import), columns pre-multiplied by
// the D65 white so that rgb = m @ (finv(fx), finv(fy), finv(fz)).
__device__ __forceinline__ double rgb_from_f(int c, double X, double Y, double Z) {
    constexpr double m[9] = {3.079980302271805,   -1.5371515162713183,  -0.5428213080224701,
                             -0.9212477523232383, 1.8759900014898907,   0.045247339514465995,
                             0.05289046109881184, -0.20404133836651123, 1.1512320119619401};
    return m[3 * c + 0] * X + m[3 * c + 1] * Y + m[3 * c + 2] * Z;
}

__device__ __forceinline__ double finv(double f) {
    return f > 0.2068966 ? f * f * f : (f - 16.0 / 116.0) * (1.0 / 7.787);
}
__device__ __forceinline__ double gamma_encode(double c) {
    const double s = c > 0.0031308 ? fma(1.055, pow_5_12(c), -0.055) : 12.92 * c;
    double s2 = s < 0.0 ? 0.0 : s;  // np.clip(arr, 0, 1); NaN stays NaN
    return s2 > 1.0 ? 1.0 : s2;
}

__device__ __forceinline__ void lab2rgb(const double (&labv)[3], double (&rgb)[3]) {
    const double fy = (labv[0] + 16.0) * (1.0 / 116.0);
    const double fx = fma(labv[1], 1.0 / 500.0, fy);
    double fz = fma(labv[2], -1.0 / 200.0, fy);
    fz = fz < 0.0 ? 0.0 : fz;  // skimage zeroes invalid z (and warns)
    const double X = finv(fx), Y = finv(fy), Z = finv(fz);
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb[c] = gamma_encode(rgb_from_f(c, X, Y, Z));
}

}  // namespace lab
}  // namespace ct
