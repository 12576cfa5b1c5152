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
// That fp64 chain serves float64 images.  The statistics pass (mean / std over millions of
// pixels) and the whole float32-image remap use the fp32 chain at the end of this file.
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


// ------------------------------------------------------------------ float32 images: fp32 chain
// Reinhard on float32 images (the dtype of the reference's CLI path) returns float32 and only has
// to land within 1e-4 / 99.99 % uint8 of the float64 oracle (the reference's own float32 path is
// 1e-4 away from it), so the whole remap runs in fp32: 126 instructions per pixel
// instead of 340, no F2F conversions.  What decides the accuracy is the cube root: its error is
// amplified ~4.4x by the cancelling rows of rgb_from_xyz, so the MUFU seed (2.5 ulp as t z^2) is
// polished by one Newton step ON y = cbrt t (residual by FMA): 0.7 ulp.  Emulated on the CPU
// (tools/emulate_reinhard_fp32.py): mean |err| 1.2e-7, uint8 flips 2.8e-5 - the same as with a
// correctly rounded fp32 cbrt; without the polish 5.4e-7 / 1.2e-4 (gate missed); with fp64
// between the two matrices 0.9e-7 / 2.3e-5 at twice the instructions.

// float thresholds are the doubles rounded DOWN, so that `xf > c_f` decides like `(double)xf > c`
#define CT_THR_DEC_F 0.040449999272823334f
#define CT_THR_F_F 0.008855999447405338f
#define CT_THR_FINV_F 0.2068965882062912f
#define CT_THR_ENC_F 0.0031307998578995466f

// t^(-1/3), t > 0, on the FMA pipe alone: integer seed (3.4 %), then one degree-5 correction
// y (1 + r P(r)), r = 1 - t y^3, P a Chebyshev fit of ((1-r)^(-1/3) - 1) / r over |r| < 0.107:
// 2.5e-7 max, bias 3e-9 (11 instructions, no XU) - the XU pipe is what the Lab kernels run out of
__device__ __forceinline__ float rcbrt_fma(float t) {
    const float y = __uint_as_float(0x54a23000u - __umulhi(__float_as_uint(t), 0x55555556u));
    const float r = fmaf(-t, (y * y) * y, 1.0f);
    float p = fmaf(r, 0.1244070650f, 0.1453286727f);
    p = fmaf(r, p, 0.1728475290f);
    p = fmaf(r, p, 0.2222193078f);
    p = fmaf(r, p, 0.3333333260f);
    return fmaf(y * r, p, y);
}
__device__ __forceinline__ float rcbrt_xu(float t) { return ex2_approx(-0.33333334f * lg2_approx(t)); }

// np.clip(s, 0, 1) in two instructions; the .NaN forms keep a NaN a NaN, like np.clip
__device__ __forceinline__ float clip01_nan(float s) {
    float r;
    asm("min.NaN.f32 %0, %1, 0f3F800000;\n\tmax.NaN.f32 %0, %0, 0f00000000;" : "=f"(r) : "f"(s));
    return r;
}

// ---- groups of N pixels, rare branches patched
// Each piecewise curve (gamma decode, f(), its inverse, gamma encode) has a linear toe that only
// very dark values take.  Evaluating both pieces and selecting costs a compare and a predicated
// instruction per value; instead the N pixels of a thread evaluate the power piece only, track the
// minimum of the arguments (one 3-input min per 2-3 values), and a thread whose minimum is in the
// toe patches its values afterwards.  Identical results, ~12 % fewer instructions where no toe is
// hit, ~4 % more where one is.  (A NaN argument gives NaN on the power piece too; FMNMX skips it.)
__device__ __forceinline__ float min3(float a, float b, float c) { return fminf(fminf(a, b), c); }

template <int N>
__device__ __forceinline__ float min_of(const float (&x)[N][3]) {
    float m = min3(x[0][0], x[0][1], x[0][2]);
#pragma unroll
    for (int i = 1; i < N; ++i) m = fminf(m, min3(x[i][0], x[i][1], x[i][2]));
    return m;
}

// linear-light rgb of N pixels (FAST: the statistics pass, see srgb_decode_h)
template <int N, bool FAST>
__device__ __forceinline__ void decode_group(const float (&v)[N][3], float (&l)[N][3]) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float u = fmaf(v[i][c], 1.0f / 1.055f, 0.055f / 1.055f);
            const float lg = lg2_approx(u);
            l[i][c] = FAST ? ex2_approx(2.4f * lg) : (u * u) * ex2_approx(0.4f * lg);
        }
    if (min_of<N>(v) <= CT_THR_DEC_F) {
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) l[i][c] = v[i][c] > CT_THR_DEC_F ? l[i][c] : v[i][c] * (1.0f / 12.92f);
    }
}

// (fx, fy, fz) of N pixels from linear-light rgb
template <int N, int FMA_SEEDS, bool POLISH>
__device__ __forceinline__ void xyzf_group(const float (&l)[N][3], float (&f)[N][3]) {
    float t[N][3];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        t[i][0] = CT_XYZ_ROW0(float, l[i][0], l[i][1], l[i][2]);
        t[i][1] = CT_XYZ_ROW1(float, l[i][0], l[i][1], l[i][2]);
        t[i][2] = CT_XYZ_ROW2(float, l[i][0], l[i][1], l[i][2]);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            // seeds: fy first from the FMA pipe, then fx, then fz
            const bool fma_seed = c == 1 ? FMA_SEEDS > 0 : (c == 0 ? FMA_SEEDS > 1 : FMA_SEEDS > 2);
            const float z = fma_seed ? rcbrt_fma(t[i][c]) : rcbrt_xu(t[i][c]);
            const float zz = z * z;
            float y = t[i][c] * zz;
            if (POLISH) y = fmaf(fmaf(-y * y, y, t[i][c]), zz * 0.33333334f, y);  // y + (t - y^3) / (3 y^2)
            f[i][c] = y;
        }
    }
    if (min_of<N>(t) <= CT_THR_F_F) {
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) f[i][c] = t[i][c] > CT_THR_F_F ? f[i][c] : fmaf(7.787f, t[i][c], 16.0f / 116.0f);
    }
}

// Statistics pass: N pixels as w = (fy - 66/116, fx - fy, fy - fz), i.e. Lab = (50 + 116 w0,
// 500 w1, 200 w2); the caller scales the SUMS instead of every pixel.  FMA_SEEDS of the three
// cube roots avoid the XU pipe.
#define CT_LAB_W0_SHIFT (66.0f / 116.0f)
template <int N, int FMA_SEEDS>
__device__ __forceinline__ void rgb2labw_group(const float (&rgb)[N][3], float (&w)[N][3]) {
    float l[N][3], f[N][3];
    decode_group<N, true>(rgb, l);
    xyzf_group<N, FMA_SEEDS, false>(l, f);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        w[i][0] = f[i][1] - CT_LAB_W0_SHIFT;
        w[i][1] = f[i][0] - f[i][1];
        w[i][2] = f[i][1] - f[i][2];
    }
}

// Per-pair constants of the fused Lab affine, folded so that the 116/500/200 scalings cancel:
//   fy' = sL fy + cL,  fx' = fy' + sa (fx - fy) + ca,  fz' = fy' - sb (fy - fz) - cb
struct ReinhardFold {
    float sL, cL, sa, ca, sb, cb;
};
// xf: the CT_XFORM_DOUBLES block of the pair (scales at 0,4,8; mean_t at 9..11; mean_r at 12..14)
__device__ __forceinline__ ReinhardFold fold_reinhard(const double *xf) {
    ReinhardFold k;
    k.sL = (float)xf[0];
    k.cL = (float)((xf[12] - (16.0 + xf[9]) * xf[0] + 16.0) / 116.0);
    k.sa = (float)xf[4];
    k.ca = (float)((xf[13] - xf[10] * xf[4]) / 500.0);
    k.sb = (float)xf[8];
    k.cb = (float)((xf[14] - xf[11] * xf[8]) / 200.0);
    return k;
}

// rgb (float) -> Reinhard-transferred rgb (float), clipped to [0,1], N pixels  (linear.py:25-40 fused)
template <int N, int FMA_SEEDS>
__device__ __forceinline__ void reinhard_group_h(const ReinhardFold &k, const float (&v)[N][3], float (&out)[N][3]) {
    float l[N][3], f[N][3], g[N][3], x[N][3], c[N][3];
    decode_group<N, false>(v, l);
    xyzf_group<N, FMA_SEEDS, true>(l, f);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const float gy = fmaf(k.sL, f[i][1], k.cL);
        float gz = gy - fmaf(k.sb, f[i][1] - f[i][2], k.cb);
        asm("max.NaN.f32 %0, %0, 0f00000000;" : "+f"(gz));  // skimage zeroes invalid z (and warns)
        g[i][0] = fmaf(k.sa, f[i][0] - f[i][1], k.ca) + gy;
        g[i][1] = gy;
        g[i][2] = gz;
#pragma unroll
        for (int j = 0; j < 3; ++j) x[i][j] = (g[i][j] * g[i][j]) * g[i][j];
    }
    if (min_of<N>(g) <= CT_THR_FINV_F) {
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                x[i][j] = g[i][j] > CT_THR_FINV_F ? x[i][j] : fmaf(g[i][j], 1.0f / 7.787f, -(16.0f / 116.0f) / 7.787f);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const float X = x[i][0], Y = x[i][1], Z = x[i][2];
        c[i][0] = fmaf(-0.5428213080224701f, Z, fmaf(-1.5371515162713183f, Y, 3.079980302271805f * X));
        c[i][1] = fmaf(0.045247339514465995f, Z, fmaf(1.8759900014898907f, Y, -0.9212477523232383f * X));
        c[i][2] = fmaf(1.1512320119619401f, Z, fmaf(-0.20404133836651123f, Y, 0.05289046109881184f * X));
#pragma unroll
        for (int j = 0; j < 3; ++j) out[i][j] = fmaf(1.055f, ex2_approx(0.41666666f * lg2_approx(c[i][j])), -0.055f);
    }
    if (min_of<N>(c) <= CT_THR_ENC_F) {
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) out[i][j] = c[i][j] > CT_THR_ENC_F ? out[i][j] : 12.92f * c[i][j];
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) out[i][j] = clip01_nan(out[i][j]);
}


// ------------------------------------------------------------------ the fp32 chain on pixel PAIRS
// Both Lab kernels were bound by instruction issue and the XU pipe together (round 1: XU 62-68 %,
// issue 61-67 %, 0.66-0.70 of the HBM roofline).  sm_100 has packed fp32 arithmetic (FFMA2 / FMUL2 /
// FADD2: two IEEE operations per instruction, same lane throughput), so the chain above is written
// once more on float2 = (pixel 2i, pixel 2i+1): the same operations in the same order on every lane
// - results are bit-identical to the scalar functions, which stay for the one-pixel tails - at half
// the issue slots.  MUFU (lg2 / ex2) and the 3-input minima of the toe tests remain scalar.  With the
// issue pressure gone all three cube-root seeds can come from the FMA pipe, which leaves the XU pipe
// with the 12 MUFU of the two gamma curves per pixel.
using f2 = float2;
__device__ __forceinline__ f2 splat(float v) { return make_float2(v, v); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 neg2(f2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ f2 lg2_2(f2 a) { return make_float2(lg2_approx(a.x), lg2_approx(a.y)); }
__device__ __forceinline__ f2 ex2_2(f2 a) { return make_float2(ex2_approx(a.x), ex2_approx(a.y)); }
__device__ __forceinline__ f2 select_gt(f2 x, float thr, f2 a, f2 b) {   // x > thr ? a : b, per lane
    return make_float2(x.x > thr ? a.x : b.x, x.y > thr ? a.y : b.y);
}

template <int P>
__device__ __forceinline__ float min_of2(const f2 (&x)[P][3]) {
    float m = min3(x[0][0].x, x[0][0].y, x[0][1].x);
    m = min3(m, x[0][1].y, x[0][2].x);
    m = fminf(m, x[0][2].y);
#pragma unroll
    for (int i = 1; i < P; ++i) {
        m = min3(m, x[i][0].x, x[i][0].y);
        m = min3(m, x[i][1].x, x[i][1].y);
        m = min3(m, x[i][2].x, x[i][2].y);
    }
    return m;
}

__device__ __forceinline__ f2 rcbrt_fma2(f2 t) {
    const f2 y = make_float2(__uint_as_float(0x54a23000u - __umulhi(__float_as_uint(t.x), 0x55555556u)),
                             __uint_as_float(0x54a23000u - __umulhi(__float_as_uint(t.y), 0x55555556u)));
    const f2 r = fma2(neg2(t), mul2(mul2(y, y), y), splat(1.0f));
    f2 p = fma2(r, splat(0.1244070650f), splat(0.1453286727f));
    p = fma2(r, p, splat(0.1728475290f));
    p = fma2(r, p, splat(0.2222193078f));
    p = fma2(r, p, splat(0.3333333260f));
    return fma2(mul2(y, r), p, y);
}
__device__ __forceinline__ f2 rcbrt_xu2(f2 t) { return ex2_2(mul2(splat(-0.33333334f), lg2_2(t))); }

template <int P, bool FAST>
__device__ __forceinline__ void decode_pairs(const f2 (&v)[P][3], f2 (&l)[P][3]) {
#pragma unroll
    for (int i = 0; i < P; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const f2 u = fma2(v[i][c], splat(1.0f / 1.055f), splat(0.055f / 1.055f));
            const f2 lg = lg2_2(u);
            l[i][c] = FAST ? ex2_2(mul2(splat(2.4f), lg)) : mul2(mul2(u, u), ex2_2(mul2(splat(0.4f), lg)));
        }
    if (min_of2<P>(v) <= CT_THR_DEC_F) {
#pragma unroll
        for (int i = 0; i < P; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) l[i][c] = select_gt(v[i][c], CT_THR_DEC_F, l[i][c], mul2(v[i][c], splat(1.0f / 12.92f)));
    }
}

// the scalar CT_XYZ_ROW* evaluate (m0 a + m1 b) + m2 c with plain multiplies and adds that the compiler
// contracts to fma(m2, c, fma(m1, b, m0 a)); written out here so that both versions round identically
template <int P, int FMA_SEEDS, bool POLISH>
__device__ __forceinline__ void xyzf_pairs(const f2 (&l)[P][3], f2 (&f)[P][3]) {
    f2 t[P][3];
#pragma unroll
    for (int i = 0; i < P; ++i) {
        t[i][0] = fma2(splat((float)(0.180423 / 0.95047)), l[i][2], fma2(splat((float)(0.357580 / 0.95047)), l[i][1], mul2(splat((float)(0.412453 / 0.95047)), l[i][0])));
        t[i][1] = fma2(splat(0.072169f), l[i][2], fma2(splat(0.715160f), l[i][1], mul2(splat(0.212671f), l[i][0])));
        t[i][2] = fma2(splat((float)(0.950227 / 1.08883)), l[i][2], fma2(splat((float)(0.119193 / 1.08883)), l[i][1], mul2(splat((float)(0.019334 / 1.08883)), l[i][0])));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const bool fma_seed = c == 1 ? FMA_SEEDS > 0 : (c == 0 ? FMA_SEEDS > 1 : FMA_SEEDS > 2);
            const f2 z = fma_seed ? rcbrt_fma2(t[i][c]) : rcbrt_xu2(t[i][c]);
            const f2 zz = mul2(z, z);
            f2 y = mul2(t[i][c], zz);
            if (POLISH) y = fma2(fma2(neg2(mul2(y, y)), y, t[i][c]), mul2(zz, splat(0.33333334f)), y);
            f[i][c] = y;
        }
    }
    if (min_of2<P>(t) <= CT_THR_F_F) {
#pragma unroll
        for (int i = 0; i < P; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) f[i][c] = select_gt(t[i][c], CT_THR_F_F, f[i][c], fma2(splat(7.787f), t[i][c], splat(16.0f / 116.0f)));
    }
}

// pixels [N][3] <-> pairs [N/2][3]
template <int N>
__device__ __forceinline__ void to_pairs(const float (&x)[N][3], f2 (&p)[N / 2][3]) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) p[i][c] = make_float2(x[2 * i][c], x[2 * i + 1][c]);
}
template <int N>
__device__ __forceinline__ void from_pairs(const f2 (&p)[N / 2][3], float (&x)[N][3]) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            x[2 * i][c] = p[i][c].x;
            x[2 * i + 1][c] = p[i][c].y;
        }
}

// rgb2labw_group on pairs (N even)
template <int N, int FMA_SEEDS>
__device__ __forceinline__ void rgb2labw_pairs(const float (&rgb)[N][3], float (&w)[N][3]) {
    constexpr int P = N / 2;
    f2 v[P][3], l[P][3], f[P][3], o[P][3];
    to_pairs<N>(rgb, v);
    decode_pairs<P, true>(v, l);
    xyzf_pairs<P, FMA_SEEDS, false>(l, f);
#pragma unroll
    for (int i = 0; i < P; ++i) {
        o[i][0] = add2(f[i][1], splat(-CT_LAB_W0_SHIFT));
        o[i][1] = add2(f[i][0], neg2(f[i][1]));
        o[i][2] = add2(f[i][1], neg2(f[i][2]));
    }
    from_pairs<N>(o, w);
}

// reinhard_group_h on pairs (N even)
template <int N, int FMA_SEEDS>
__device__ __forceinline__ void reinhard_pairs_h(const ReinhardFold &k, const float (&vin)[N][3], float (&out)[N][3]) {
    constexpr int P = N / 2;
    f2 v[P][3], l[P][3], f[P][3], g[P][3], x[P][3], c[P][3], o[P][3];
    to_pairs<N>(vin, v);
    decode_pairs<P, false>(v, l);
    xyzf_pairs<P, FMA_SEEDS, true>(l, f);
#pragma unroll
    for (int i = 0; i < P; ++i) {
        const f2 gy = fma2(splat(k.sL), f[i][1], splat(k.cL));
        f2 gz = add2(gy, neg2(fma2(splat(k.sb), add2(f[i][1], neg2(f[i][2])), splat(k.cb))));
        asm("max.NaN.f32 %0, %0, 0f00000000;" : "+f"(gz.x));  // skimage zeroes invalid z (and warns)
        asm("max.NaN.f32 %0, %0, 0f00000000;" : "+f"(gz.y));
        g[i][0] = add2(fma2(splat(k.sa), add2(f[i][0], neg2(f[i][1])), splat(k.ca)), gy);
        g[i][1] = gy;
        g[i][2] = gz;
#pragma unroll
        for (int j = 0; j < 3; ++j) x[i][j] = mul2(mul2(g[i][j], g[i][j]), g[i][j]);
    }
    if (min_of2<P>(g) <= CT_THR_FINV_F) {
#pragma unroll
        for (int i = 0; i < P; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                x[i][j] = select_gt(g[i][j], CT_THR_FINV_F, x[i][j], fma2(g[i][j], splat(1.0f / 7.787f), splat(-(16.0f / 116.0f) / 7.787f)));
    }
#pragma unroll
    for (int i = 0; i < P; ++i) {
        const f2 X = x[i][0], Y = x[i][1], Z = x[i][2];
        c[i][0] = fma2(splat(-0.5428213080224701f), Z, fma2(splat(-1.5371515162713183f), Y, mul2(splat(3.079980302271805f), X)));
        c[i][1] = fma2(splat(0.045247339514465995f), Z, fma2(splat(1.8759900014898907f), Y, mul2(splat(-0.9212477523232383f), X)));
        c[i][2] = fma2(splat(1.1512320119619401f), Z, fma2(splat(-0.20404133836651123f), Y, mul2(splat(0.05289046109881184f), X)));
#pragma unroll
        for (int j = 0; j < 3; ++j) o[i][j] = fma2(splat(1.055f), ex2_2(mul2(splat(0.41666666f), lg2_2(c[i][j]))), splat(-0.055f));
    }
    if (min_of2<P>(c) <= CT_THR_ENC_F) {
#pragma unroll
        for (int i = 0; i < P; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) o[i][j] = select_gt(c[i][j], CT_THR_ENC_F, o[i][j], mul2(splat(12.92f), c[i][j]));
    }
#pragma unroll
    for (int i = 0; i < P; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) o[i][j] = make_float2(clip01_nan(o[i][j].x), clip01_nan(o[i][j].y));
    from_pairs<N>(o, out);
}

}  // namespace lab
}  // namespace ct
