// Shared device helpers: image descriptors, vectorised pixel-group loads/stores, reductions,
// monotone range keys.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/ct_b200.h"

namespace ct {

constexpr int kThreads = 256;         // every streaming kernel uses 8 warps per CTA
constexpr int kWarps = kThreads / 32;

// Device-side view of one ct_batch (passed by value in kernel parameters).
struct Img {
    const void *data;
    int64_t npix;
    int64_t image_stride;  // elements
    int64_t plane_stride;  // elements (CHW)
};

struct ImgOut {
    void *data;
    int64_t npix;
    int64_t image_stride;
    int64_t plane_stride;
};

// ---------------------------------------------------------------------------------------------
// Pixel groups.  A thread moves one group of G = 16 B / sizeof(T) pixels per step so that
// every global access is a 128-bit LDG/STG: planar images take one vector per plane (fully
// coalesced, 512 B per warp instruction); interleaved images take three consecutive vectors
// per thread (48 B per thread, 1536 contiguous bytes per warp, the unused halves of each
// 32 B sector are served from L1 by the sibling instruction).
// ---------------------------------------------------------------------------------------------
template <typename T> struct Vec;
template <> struct Vec<float> { using type = float4; static constexpr int G = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int G = 2; };

// output vectors are written once and not read again by the same kernel
#ifndef CT_STORE_CS
#define CT_STORE_CS 0
#endif
template <typename V>
__device__ __forceinline__ void st_vec(V *p, const V &v) {
#if CT_STORE_CS
    __stcs(p, v);
#else
    *p = v;
#endif
}

template <typename T, int LAYOUT>
struct PixelIO {
    static constexpr int G = Vec<T>::G;
    using V = typename Vec<T>::type;
    using elem_t = T;
    static constexpr int kLayout = LAYOUT;

    // scalar access to one pixel (tails, unaligned batches)
    __device__ __forceinline__ static void load1(const T *img, int64_t plane, int64_t p,
                                                 double (&x)[3]) {
        if (LAYOUT == CT_HWC) {
            x[0] = (double)img[3 * p + 0];
            x[1] = (double)img[3 * p + 1];
            x[2] = (double)img[3 * p + 2];
        } else {
            x[0] = (double)img[p];
            x[1] = (double)img[plane + p];
            x[2] = (double)img[2 * plane + p];
        }
    }
    template <typename X>
    __device__ __forceinline__ static void store1(T *img, int64_t plane, int64_t p,
                                                  const X (&x)[3]) {
        if (LAYOUT == CT_HWC) {
            img[3 * p + 0] = (T)x[0];
            img[3 * p + 1] = (T)x[1];
            img[3 * p + 2] = (T)x[2];
        } else {
            img[p] = (T)x[0];
            img[plane + p] = (T)x[1];
            img[2 * plane + p] = (T)x[2];
        }
    }

    // The raw vectors of one group: pixels [g*G, g*G+G).  Kept packed (12 registers) so that a
    // kernel can have the next group's loads in flight while it works on the current one.
    struct Raw {
        T e[3 * G];
    };
    template <bool VEC>
    __device__ __forceinline__ static Raw load_raw(const T *img, int64_t plane, int g) {
        Raw r;
        if (VEC) {
            if (LAYOUT == CT_HWC) {
                const V *v = reinterpret_cast<const V *>(img) + 3 * (int64_t)g;
#pragma unroll
                for (int k = 0; k < 3; ++k) *reinterpret_cast<V *>(&r.e[k * G]) = v[k];
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    *reinterpret_cast<V *>(&r.e[c * G]) = *(reinterpret_cast<const V *>(img + c * plane) + g);
            }
        } else {
            if (LAYOUT == CT_HWC) {
                const T *q = img + 3 * G * (int64_t)g;
#pragma unroll
                for (int k = 0; k < 3 * G; ++k) r.e[k] = q[k];
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int i = 0; i < G; ++i) r.e[c * G + i] = img[c * plane + (int64_t)g * G + i];
            }
        }
        return r;
    }
    __device__ __forceinline__ static void unpack(const Raw &r, double (&x)[G][3]) {
#pragma unroll
        for (int i = 0; i < G; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) x[i][c] = (double)(LAYOUT == CT_HWC ? r.e[3 * i + c] : r.e[c * G + i]);
    }
    // the same pixels as floats (exact for float images): seeds / statistics of the Lab path
    __device__ __forceinline__ static void unpack_f(const Raw &r, float (&x)[G][3]) {
#pragma unroll
        for (int i = 0; i < G; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) x[i][c] = (float)(LAYOUT == CT_HWC ? r.e[3 * i + c] : r.e[c * G + i]);
    }
    template <bool VEC>
    __device__ __forceinline__ static void load(const T *img, int64_t plane, int64_t g, double (&x)[G][3]) {
        unpack(load_raw<VEC>(img, plane, (int)g), x);
    }

    // store GS pixels starting at pixel p0 (GS is the SOURCE group size; p0 % GS == 0)
    template <bool VEC, int GS, typename X>
    __device__ __forceinline__ static void store(T *img, int64_t plane, int64_t p0,
                                                 const X (&x)[GS][3]) {
        if (!VEC || GS % G != 0) {  // narrower source group than one destination vector
#pragma unroll
            for (int i = 0; i < GS; ++i) store1(img, plane, p0 + i, x[i]);
            return;
        }
        constexpr int NV = GS / G > 0 ? GS / G : 1;  // destination vectors per plane / triple-set
        if (LAYOUT == CT_HWC) {
            T raw[3 * GS];
#pragma unroll
            for (int i = 0; i < GS; ++i)
#pragma unroll
                for (int c = 0; c < 3; ++c) raw[3 * i + c] = (T)x[i][c];
            V *v = reinterpret_cast<V *>(img + 3 * p0);
#pragma unroll
            for (int k = 0; k < 3 * NV; ++k) st_vec(v + k, *reinterpret_cast<V *>(&raw[k * G]));
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                T raw[GS];
#pragma unroll
                for (int i = 0; i < GS; ++i) raw[i] = (T)x[i][c];
                V *v = reinterpret_cast<V *>(img + c * plane + p0);
#pragma unroll
                for (int k = 0; k < NV; ++k) st_vec(v + k, *reinterpret_cast<V *>(&raw[k * G]));
            }
        }
    }
};

// in-kernel dispatch over (dtype, layout, vectorised): ID = (dtype*2 + layout)*2 + vec
#define CT_FOR_EACH_SRC(CALL)                                    \
    CALL(0, float, CT_HWC, false) CALL(1, float, CT_HWC, true)   \
    CALL(2, float, CT_CHW, false) CALL(3, float, CT_CHW, true)   \
    CALL(4, double, CT_HWC, false) CALL(5, double, CT_HWC, true) \
    CALL(6, double, CT_CHW, false) CALL(7, double, CT_CHW, true)

// ---------------------------------------------------------------------------------------------
// Monotone int64 keys for doubles: signed integer order == floating-point order, so ranges can
// be folded with atomicMin and all-reduced with an integer MIN (exact, order independent).
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int64_t key_of(double x) {
#ifdef __CUDA_ARCH__
    int64_t b = __double_as_longlong(x);
#else
    int64_t b;
    memcpy(&b, &x, 8);
#endif
    return b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
}
__host__ __device__ __forceinline__ double value_of(int64_t k) {
    int64_t b = k >= 0 ? k : (k ^ 0x7fffffffffffffffLL);
#ifdef __CUDA_ARCH__
    return __longlong_as_double(b);
#else
    double x;
    memcpy(&x, &b, 8);
    return x;
#endif
}
constexpr int64_t kKeyPlusInf = 0x7ff0000000000000LL;

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// IEEE operations that must not be contracted into FMAs (numpy evaluates them separately).
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

// r @ x with the accumulation order of a K=3 dgemm micro-kernel: ((r0*x0) + r1*x1) + r2*x2,
// each step fused (iterative.py:34-35).
__device__ __forceinline__ double dot3(const double *r, const double (&x)[3]) {
    return fma(r[2], x[2], fma(r[1], x[1], mul_rn(r[0], x[0])));
}

}  // namespace ct
