// sRGB <-> CIE-Lab (D65, 2 degree) per-pixel math, the published scikit-image algorithm that
// the reference calls at methods/linear.py:25,26,40 (rgb2xyz, xyz2lab, _lab2xyz, xyz2rgb).
//
// The three fractional powers (x^2.4, cbrt, x^(1/2.4)) are the cost of this path, and on B200
// the scarce unit is the XU pipe (MUFU and every F2F/I2F conversion: 16 lanes/clk/SM, a quarter
// of the fp64 rate).  Each power is therefore computed as
//     seed   y0 ~ x^(-1/n) from MUFU lg2/ex2 on the operand rounded to fp32             (2-3 XU ops)
//     polish one DIVISION-FREE Newton step for y = x^(-1/n) in fp64:
//            y = y0 * ((n+1) - x*y0^n) / n          error (n+1)/2 * e0^2 ~ 1e-12  (1 XU op: y0 -> double)
//     result x^2.4 = (x*y)^3 (n=5),  cbrt x = x*y^2 (n=3),  x^(5/12) = x*y^7 (n=12)
// i.e. 3-4 XU operations per power instead of 7 (pow() would be ~150 fp64 instructions).
//
// The statistics pass (mean / std over millions of pixels) uses the fp32 chain alone: its
// per-pixel error (~3e-7 relative) shifts the Lab means by ~1e-5 of a Lab unit, 1e-7 in RGB.
#pragma once

#include "ct_common.cuh"

namespace ct {
namespace lab {

// fp64 literals live in the constant bank: DMUL/DFMA take a c[bank][offset] operand directly,
// whereas a 64-bit immediate costs two MOVs every time the compiler re-materialises it.
struct LabConstants {
    double m0[3], m1[3], m2[3];     // xyz_from_rgb rows / white
    double r0[3], r1[3], r2[3];     // inv(xyz_from_rgb) columns * white
    double inv1055, c0055, inv1292, thr_dec, thr_f, k7787, k16_116, third, four_thirds;
    double c116, c500, c200, inv116, inv500, ninv200, thr_finv, inv7787, thr_enc, c1055, c1292;
    double fifth, six_fifths, twelfth, thirteen_twelfths;
};
__constant__ LabConstants kL = {
    {0.412453 / 0.95047, 0.357580 / 0.95047, 0.180423 / 0.95047},
    {0.212671, 0.715160, 0.072169},
    {0.019334 / 1.08883, 0.119193 / 1.08883, 0.950227 / 1.08883},
    {3.079980302271805, -1.5371515162713183, -0.5428213080224701},
    {-0.9212477523232383, 1.8759900014898907, 0.045247339514465995},
    {0.05289046109881184, -0.20404133836651123, 1.1512320119619401},
    1.0 / 1.055, 0.055, 1.0 / 12.92, 0.04045, 0.008856, 7.787, 16.0 / 116.0, 1.0 / 3.0, 4.0 / 3.0,
    116.0, 500.0, 200.0, 1.0 / 116.0, 1.0 / 500.0, -1.0 / 200.0, 0.2068966, 1.0 / 7.787, 0.0031308, 1.055, 12.92,
    0.2, 1.2, 1.0 / 12.0, 13.0 / 12.0};

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

// xyz_from_rgb with each row divided by the D65 white (0.95047, 1, 1.08883)
#define CT_XYZ_ROW0(T, a, b, c) ((T)(0.412453 / 0.95047) * (a) + (T)(0.357580 / 0.95047) * (b) + (T)(0.180423 / 0.95047) * (c))
#define CT_XYZ_ROW1(T, a, b, c) ((T)0.212671 * (a) + (T)0.715160 * (b) + (T)0.072169 * (c))
#define CT_XYZ_ROW2(T, a, b, c) ((T)(0.019334 / 1.08883) * (a) + (T)(0.119193 / 1.08883) * (b) + (T)(0.950227 / 1.08883) * (c))

// ------------------------------------------------------------------ fp32 chain (statistics, seeds)
__device__ __forceinline__ float srgb_decode_f(float v) {
    return v > 0.04045f ? ex2_approx(2.4f * lg2_approx((v + 0.055f) * (1.0f / 1.055f))) : v * (1.0f / 12.92f);
}
__device__ __forceinline__ float lab_f_f(float t) {
    return t > 0.008856f ? ex2_approx(0.33333334f * lg2_approx(t)) : fmaf(7.787f, t, 16.0f / 116.0f);
}
__device__ __forceinline__ void rgb2lab_f32(const float (&rgb)[3], float (&out)[3]) {
    const float l0 = srgb_decode_f(rgb[0]), l1 = srgb_decode_f(rgb[1]), l2 = srgb_decode_f(rgb[2]);
    const float fx = lab_f_f(CT_XYZ_ROW0(float, l0, l1, l2));
    const float fy = lab_f_f(CT_XYZ_ROW1(float, l0, l1, l2));
    const float fz = lab_f_f(CT_XYZ_ROW2(float, l0, l1, l2));
    out[0] = fmaf(116.0f, fy, -16.0f);
    out[1] = 500.0f * (fx - fy);
    out[2] = 200.0f * (fy - fz);
}

// ------------------------------------------------------------------ fp64 chain, fp32 seeds
// The seed of each power needs its operand as a float: one F2F (XU) per power, except for the
// gamma decode of float32 images whose operand already is one (`vf`).

// ((v + 0.055) / 1.055)^2.4 for v > 0.04045, else v / 12.92
__device__ __forceinline__ double srgb_decode(double v, float vf) {
    if (v > kL.thr_dec) {
        const double u = (v + kL.c0055) * kL.inv1055;
        const float uf = (vf + 0.055f) * (1.0f / 1.055f);
        double y = (double)ex2_approx(-0.2f * lg2_approx(uf));   // u^(-1/5)
        const double y2 = y * y, y4 = y2 * y2;
        y *= fma(-(u * y4) * y, kL.fifth, kL.six_fifths);        // y (6 - u y^5) / 5
        const double uy = u * y;                                 // u^(4/5)
        return uy * uy * uy;
    }
    return v * kL.inv1292;
}

// cbrt(t) for t > 0.008856, else 7.787 t + 16/116
__device__ __forceinline__ double lab_f(double t) {
    if (t > kL.thr_f) {
        double y = (double)ex2_approx(-0.33333334f * lg2_approx((float)t));  // t^(-1/3)
        y *= fma(-(t * y) * (y * y), kL.third, kL.four_thirds);               // y (4 - t y^3) / 3
        return t * (y * y);
    }
    return fma(kL.k7787, t, kL.k16_116);
}

__device__ __forceinline__ void rgb2lab(const double (&rgb)[3], const float (&rgbf)[3], double (&out)[3]) {
    double l[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) l[c] = srgb_decode(rgb[c], rgbf[c]);
    const double fx = lab_f(fma(kL.m0[2], l[2], fma(kL.m0[1], l[1], kL.m0[0] * l[0])));
    const double fy = lab_f(fma(kL.m1[2], l[2], fma(kL.m1[1], l[1], kL.m1[0] * l[0])));
    const double fz = lab_f(fma(kL.m2[2], l[2], fma(kL.m2[1], l[1], kL.m2[0] * l[0])));
    out[0] = fma(kL.c116, fy, -16.0);
    out[1] = kL.c500 * (fx - fy);
    out[2] = kL.c200 * (fy - fz);
}

__device__ __forceinline__ double finv(double f) {
    return f > kL.thr_finv ? f * f * f : (f - kL.k16_116) * kL.inv7787;
}

// 1.055 c^(1/2.4) - 0.055 for c > 0.0031308, else 12.92 c; then np.clip(., 0, 1)
__device__ __forceinline__ double srgb_encode(double c) {
    double s;
    if (c > kL.thr_enc) {
        double y = (double)ex2_approx(-0.083333336f * lg2_approx((float)c));  // c^(-1/12)
        double y2 = y * y, y4 = y2 * y2;
        const double y12 = (y4 * y4) * y4;
        y *= fma(-c * y12, kL.twelfth, kL.thirteen_twelfths);           // y (13 - c y^12) / 12
        y2 = y * y;
        y4 = y2 * y2;
        s = fma(kL.c1055, c * ((y4 * y2) * y), -kL.c0055);              // c y^7 = c^(5/12)
    } else {
        s = kL.c1292 * c;
    }
    s = s < 0.0 ? 0.0 : s;  // NaN stays NaN, like np.clip
    return s > 1.0 ? 1.0 : s;
}

__device__ __forceinline__ void lab2rgb(const double (&labv)[3], double (&rgb)[3]) {
    const double fy = (labv[0] + 16.0) * kL.inv116;
    const double fx = fma(labv[1], kL.inv500, fy);
    double fz = fma(labv[2], kL.ninv200, fy);
    fz = fz < 0.0 ? 0.0 : fz;  // skimage zeroes invalid z (and warns)
    const double X = finv(fx), Y = finv(fy), Z = finv(fz);
    rgb[0] = srgb_encode(fma(kL.r0[2], Z, fma(kL.r0[1], Y, kL.r0[0] * X)));
    rgb[1] = srgb_encode(fma(kL.r1[2], Z, fma(kL.r1[1], Y, kL.r1[0] * X)));
    rgb[2] = srgb_encode(fma(kL.r2[2], Z, fma(kL.r2[1], Y, kL.r2[0] * X)));
}

}  // namespace lab
}  // namespace ct
