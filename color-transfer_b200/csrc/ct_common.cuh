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
//
// A group is worked on in SUB-GROUPS of GS pixels (what a kernel holds unpacked in registers):
// GS = G for float (4) and double (2) images; uint8 images (G = 16: 48 bytes per thread) have four
// sub-groups of 4 pixels.  uint8 samples are decoded through a 256-entry table in shared memory
// (Decode) to exactly the value the reference's loaders produce: k/255 in float32 (torch `/ 255`,
// ref: utils/data.py:106) or float64 (skimage.img_as_float, ref: utils/postprocess.py:138).
// ---------------------------------------------------------------------------------------------
template <typename T> struct Vec;
template <> struct Vec<float> { using type = float4; static constexpr int G = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int G = 2; };
template <> struct Vec<uint8_t> { using type = uint4; static constexpr int G = 16; };

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

// 256-bit store (sm_100: STG.E.256) of four doubles to a 32-byte aligned address: a thread that holds four
// consecutive fp64 values writes whole 32-byte sectors in one instruction (two 128-bit stores at a lane stride
// of 32 bytes each touch every sector half-filled)
__device__ __forceinline__ void st_f64x4(double *p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// decode tables of a uint8 image (shared memory, 256 entries each); unused for float images
struct Decode {
    const double *d;  // the decoded sample as the kernel's fp64 working value
    const float *f;   // the same value rounded to float (screens, fp32 Lab chain)
};
// fills the tables: as_float32 != 0 -> float32 k/255 (widened exactly), else float64 k/255
__device__ __forceinline__ void fill_decode(double *d, float *f, int as_float32) {
    for (int k = threadIdx.x; k < 256; k += blockDim.x) {
        const double v = (double)k / 255.0;
        const float vf = (float)v;               // == float32(k) / float32(255) for every k (checked on the CPU)
        d[k] = as_float32 ? (double)vf : v;
        f[k] = vf;
    }
}

// np.rint(np.clip(y, 0, 1) * 255) as uint8 (img_as_ubyte of the clipped result, ref:
// utils/postprocess.py:138); NaN -> 0 like the C cast numpy performs
template <typename X>
__device__ __forceinline__ uint32_t quantize_u8(X y) {
    double c = (double)y;
    c = c < 0.0 ? 0.0 : (c > 1.0 ? 1.0 : c);
    return (uint32_t)__double2int_rn(c * 255.0) & 0xffu;
}
// torch.clamp(y, 0, 1) (ref: methods/__init__.py:30): NaN stays NaN
template <typename X>
__device__ __forceinline__ X clamp01(X y) { return y < (X)0 ? (X)0 : (y > (X)1 ? (X)1 : y); }

template <typename T, int LAYOUT>
struct PixelIO {
    static constexpr int G = Vec<T>::G;
    static constexpr bool kU8 = sizeof(T) == 1;
    static constexpr int GS = kU8 ? 4 : G;       // pixels per sub-group
    static constexpr int NSUB = G / GS;          // sub-groups per group
    using V = typename Vec<T>::type;
    using elem_t = T;
    static constexpr int kLayout = LAYOUT;

    // scalar access to one pixel (tails, unaligned batches)
    __device__ __forceinline__ static void load1(const T *img, int64_t plane, int64_t p, const Decode &dec,
                                                 double (&x)[3]) {
        if (kU8) {
            if (LAYOUT == CT_HWC) {
                x[0] = dec.d[(int)img[3 * p + 0]];
                x[1] = dec.d[(int)img[3 * p + 1]];
                x[2] = dec.d[(int)img[3 * p + 2]];
            } else {
                x[0] = dec.d[(int)img[p]];
                x[1] = dec.d[(int)img[plane + p]];
                x[2] = dec.d[(int)img[2 * plane + p]];
            }
        } else if (LAYOUT == CT_HWC) {
            x[0] = (double)img[3 * p + 0];
            x[1] = (double)img[3 * p + 1];
            x[2] = (double)img[3 * p + 2];
        } else {
            x[0] = (double)img[p];
            x[1] = (double)img[plane + p];
            x[2] = (double)img[2 * plane + p];
        }
    }
    // `clamp`: torch.clamp(y, 0, 1) first - float32 outputs only (the Runner's tensors); float64
    // outputs are the reference functions' own unclipped results, uint8 outputs always clip
    template <typename X>
    __device__ __forceinline__ static T encode(X v, bool clamp) {
        if (kU8) return (T)quantize_u8(v);
        if (sizeof(T) == 4) return (T)(clamp ? clamp01(v) : v);
        return (T)v;
    }
    template <typename X>
    __device__ __forceinline__ static void store1(T *img, int64_t plane, int64_t p,
                                                  const X (&x)[3], bool clamp = false) {
        if (LAYOUT == CT_HWC) {
            img[3 * p + 0] = encode(x[0], clamp);
            img[3 * p + 1] = encode(x[1], clamp);
            img[3 * p + 2] = encode(x[2], clamp);
        } else {
            img[p] = encode(x[0], clamp);
            img[plane + p] = encode(x[1], clamp);
            img[2 * plane + p] = encode(x[2], clamp);
        }
    }

    // The raw vectors of one group: pixels [g*G, g*G+G).  Kept packed (12 registers) so that a
    // kernel can have the next group's loads in flight while it works on the current one.
    // uint8: 12 words; sub-group s is words 3s..3s+2 (interleaved: 12 consecutive bytes = 4 pixels)
    // or words s, 4+s, 8+s (planar: 4 pixels of each plane).
    struct alignas(16) Raw {
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
    // sample (pixel i of sub-group s, channel c) of a raw group, as stored
    __device__ __forceinline__ static T raw_at(const Raw &r, int s, int i, int c) {
        return LAYOUT == CT_HWC ? r.e[3 * (s * GS + i) + c] : r.e[c * G + s * GS + i];
    }
    __device__ __forceinline__ static void unpack_sub(const Raw &r, int s, const Decode &dec, double (&x)[GS][3]) {
#pragma unroll
        for (int i = 0; i < GS; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (kU8) x[i][c] = dec.d[(int)raw_at(r, s, i, c)];
                else x[i][c] = (double)raw_at(r, s, i, c);
            }
    }
    // the same pixels as floats (exact for float images): screens / statistics of the Lab path
    __device__ __forceinline__ static void unpack_sub_f(const Raw &r, int s, const Decode &dec, float (&x)[GS][3]) {
#pragma unroll
        for (int i = 0; i < GS; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (kU8) x[i][c] = dec.f[(int)raw_at(r, s, i, c)];
                else x[i][c] = (float)raw_at(r, s, i, c);
            }
    }
    // first pixel of sub-group s of group g when groups are consecutive per thread (load_raw)
    __device__ __forceinline__ static int64_t sub_pixel0(int64_t g, int s) { return g * G + s * GS; }

    // store NS pixels starting at pixel p0 (NS is the SOURCE sub-group size; p0 % NS == 0)
    // wide: float64 destination, 32-byte aligned image (and planes): 256-bit stores where a thread holds 32 contiguous bytes
    template <bool VEC, int NS, typename X>
    __device__ __forceinline__ static void store(T *img, int64_t plane, int64_t p0,
                                                 const X (&x)[NS][3], bool clamp = false, bool wide = false) {
        if constexpr (VEC && sizeof(T) == 8 && LAYOUT == CT_CHW && NS == 4) {
            if (wide) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    st_f64x4(reinterpret_cast<double *>(img) + c * plane + p0, (double)encode(x[0][c], clamp), (double)encode(x[1][c], clamp),
                             (double)encode(x[2][c], clamp), (double)encode(x[3][c], clamp));
                return;
            }
        }
        if constexpr (VEC && sizeof(T) == 8 && LAYOUT == CT_HWC && NS == 2) {
            // 48 contiguous bytes per thread, 32-byte aligned for every other thread: one 256-bit and one 128-bit
            // store, arranged (without divergence) so that the two lanes that share a sector write it in the
            // same instruction - three 128-bit stores at a lane stride of 48 bytes touch every sector twice
            if (wide) {
                double *o = reinterpret_cast<double *>(img) + 3 * p0;
                const bool odd = ((p0 >> 1) & 1) != 0;
                const double e[6] = {(double)encode(x[0][0], clamp), (double)encode(x[0][1], clamp), (double)encode(x[0][2], clamp),
                                     (double)encode(x[1][0], clamp), (double)encode(x[1][1], clamp), (double)encode(x[1][2], clamp)};
                st_f64x4(o + (odd ? 2 : 0), odd ? e[2] : e[0], odd ? e[3] : e[1], odd ? e[4] : e[2], odd ? e[5] : e[3]);
                st_vec(reinterpret_cast<double2 *>(o + (odd ? 0 : 4)), make_double2(odd ? e[0] : e[4], odd ? e[1] : e[5]));
                return;
            }
        }
        if constexpr (VEC && sizeof(T) == 8 && LAYOUT == CT_HWC && NS == 4) {   // 96 contiguous bytes per thread
            if (wide) {
                double *o = reinterpret_cast<double *>(img) + 3 * p0;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    double e[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) e[i] = (double)encode(x[(4 * k + i) / 3][(4 * k + i) % 3], clamp);
                    st_f64x4(o + 4 * k, e[0], e[1], e[2], e[3]);
                }
                return;
            }
        }
        // (float32 interleaved output, 48 bytes per thread: the same 256 + 128-bit arrangement was measured on the
        // Reinhard remap - 0.79 -> 0.77 of the roofline, its 12 selects cost more issue slots than the stores save)
        if (kU8) {
            // 3 bytes per pixel: NS = 4 -> three 32-bit words (interleaved) or one per plane;
            // NS = 2 (float64 state) -> 16-bit stores
            if (LAYOUT == CT_HWC) {
                uint8_t q[3 * NS];
#pragma unroll
                for (int i = 0; i < NS; ++i)
#pragma unroll
                    for (int c = 0; c < 3; ++c) q[3 * i + c] = (uint8_t)quantize_u8(x[i][c]);
                uint8_t *o = reinterpret_cast<uint8_t *>(img) + 3 * p0;
                if (VEC && NS == 4) {
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        reinterpret_cast<uint32_t *>(o)[k] = (uint32_t)q[4 * k] | ((uint32_t)q[4 * k + 1] << 8) |
                                                             ((uint32_t)q[4 * k + 2] << 16) | ((uint32_t)q[4 * k + 3] << 24);
                } else if (VEC && NS == 2) {
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        reinterpret_cast<uint16_t *>(o)[k] = (uint16_t)((uint32_t)q[2 * k] | ((uint32_t)q[2 * k + 1] << 8));
                } else {
#pragma unroll
                    for (int k = 0; k < 3 * NS; ++k) o[k] = q[k];
                }
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    uint8_t *o = reinterpret_cast<uint8_t *>(img) + c * plane + p0;
                    if (VEC && NS == 4) {
                        *reinterpret_cast<uint32_t *>(o) = quantize_u8(x[0][c]) | (quantize_u8(x[1][c]) << 8) |
                                                           (quantize_u8(x[2][c]) << 16) | (quantize_u8(x[3][c]) << 24);
                    } else {
#pragma unroll
                        for (int i = 0; i < NS; ++i) o[i] = (uint8_t)quantize_u8(x[i][c]);
                    }
                }
            }
            return;
        }
        if (VEC && LAYOUT == CT_HWC && sizeof(T) == 4 && NS == 2) {   // float32 output of a float64 state group: 24 bytes
            float2 *v = reinterpret_cast<float2 *>(img + 3 * p0);
            const float r[6] = {(float)encode(x[0][0], clamp), (float)encode(x[0][1], clamp), (float)encode(x[0][2], clamp),
                                (float)encode(x[1][0], clamp), (float)encode(x[1][1], clamp), (float)encode(x[1][2], clamp)};
#pragma unroll
            for (int k = 0; k < 3; ++k) v[k] = make_float2(r[2 * k], r[2 * k + 1]);
            return;
        }
        if (!VEC || NS % G != 0) {  // narrower source group than one destination vector
#pragma unroll
            for (int i = 0; i < NS; ++i) store1(img, plane, p0 + i, x[i], clamp);
            return;
        }
        constexpr int NV = NS / G > 0 ? NS / G : 1;  // destination vectors per plane / triple-set
        if (LAYOUT == CT_HWC) {
            T raw[3 * NS];
#pragma unroll
            for (int i = 0; i < NS; ++i)
#pragma unroll
                for (int c = 0; c < 3; ++c) raw[3 * i + c] = encode(x[i][c], clamp);
            V *v = reinterpret_cast<V *>(img + 3 * p0);
#pragma unroll
            for (int k = 0; k < 3 * NV; ++k) st_vec(v + k, *reinterpret_cast<V *>(&raw[k * G]));
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                T raw[NS];
#pragma unroll
                for (int i = 0; i < NS; ++i) raw[i] = encode(x[i][c], clamp);
                V *v = reinterpret_cast<V *>(img + c * plane + p0);
#pragma unroll
                for (int k = 0; k < NV; ++k) st_vec(v + k, *reinterpret_cast<V *>(&raw[k * G]));
            }
        }
    }
};

template <typename A, typename B> struct same_io { static constexpr bool value = false; };
template <typename A> struct same_io<A, A> { static constexpr bool value = true; };

// in-kernel dispatch over (dtype, layout, vectorised): ID = (dtype*2 + layout)*2 + vec
#define CT_FOR_EACH_SRC(CALL)                                    \
    CALL(0, float, CT_HWC, false) CALL(1, float, CT_HWC, true)   \
    CALL(2, float, CT_CHW, false) CALL(3, float, CT_CHW, true)   \
    CALL(4, double, CT_HWC, false) CALL(5, double, CT_HWC, true) \
    CALL(6, double, CT_CHW, false) CALL(7, double, CT_CHW, true) \
    CALL(8, uint8_t, CT_HWC, false) CALL(9, uint8_t, CT_HWC, true) \
    CALL(10, uint8_t, CT_CHW, false) CALL(11, uint8_t, CT_CHW, true)
// the float / double kinds only (kernels that keep a separate instantiation for uint8 images so that
// the float paths' register allocation is not disturbed)
#define CT_FOR_EACH_FLOAT_SRC(CALL)                              \
    CALL(0, float, CT_HWC, false) CALL(1, float, CT_HWC, true)   \
    CALL(2, float, CT_CHW, false) CALL(3, float, CT_CHW, true)   \
    CALL(4, double, CT_HWC, false) CALL(5, double, CT_HWC, true) \
    CALL(6, double, CT_CHW, false) CALL(7, double, CT_CHW, true)
#define CT_FOR_EACH_U8_SRC(CALL)                                 \
    CALL(8, uint8_t, CT_HWC, false) CALL(9, uint8_t, CT_HWC, true) \
    CALL(10, uint8_t, CT_CHW, false) CALL(11, uint8_t, CT_CHW, true)

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
