// Artificial-distortion generator of the reference's test set:
// ref: utils/data.py:12-22 `setup_grid_distortions` = identity + {brightness, contrast, saturation, hue,
// gamma} x 6 magnitudes, each a torchvision.transforms.functional.adjust_* call on the uint8 ground-truth
// image (ref: utils/data.py:101-104).  One pass over the source writes every distorted copy: 3 bytes per
// pixel in, 3 bytes per pixel and distortion out.
//
// The arithmetic restates torchvision 0.26's tensor backend (torchvision/transforms/_functional_tensor.py)
// for uint8 images, operation by operation in float32 with IEEE rounding (no FMA contraction):
//   _blend(a, b, ratio)   = (ratio * a + (1 - ratio) * b).clamp(0, 255).to(uint8)       (truncation)
//   brightness            = _blend(img, 0, f)
//   contrast              = _blend(img, mean(gray(img)), f)        gray = (0.2989 r + 0.587 g + 0.114 b).to(uint8)
//   saturation            = _blend(img, gray(img), f)
//   hue                   = img / 255 -> _rgb2hsv -> h = (h + f) % 1 -> _hsv2rgb -> (x * 255.999).to(uint8)
//   gamma                 = img / 255 -> x ** g -> clamp(0, 1) -> (x * 255.999).to(uint8)
// Brightness, contrast and gamma depend on one channel value only: 256-entry tables per distortion, built in
// the prologue.  Two places cannot be bit-identical to a CPU torch by construction and are within one
// level there (tests/test_gpu_distort.py): the float32 mean of the contrast blend (torch sums 8 M float32
// values in its own order; here the integer sum is exact and rounded once) and x ** g (torch: Sleef's
// 1-ulp powf; here the float64 pow rounded to float32).
#include "ct_context.h"

namespace ct {

constexpr int kMaxDistortions = CT_DISTORT_MAX_OPS;

struct DistortArgs {
    const uint8_t *src;
    uint8_t *dst;
    int64_t npix;
    int64_t src_image_stride, src_plane, dst_image_stride, dst_plane;   // in bytes (= elements)
    int layout;   // CT_HWC / CT_CHW, the same for source and copies
    int vec;      // 8-pixel groups through 64-bit accesses
    int any_hue, any_sat;
    int n_ops;
    const unsigned long long *gray_sum;   // [count] sum of gray(img) over the image (contrast), or NULL
    int kind[kMaxDistortions];
    float ratio[kMaxDistortions];   // blend ratio / hue shift / gamma, as float32 (what the float32 tensor op sees)
    float omr[kMaxDistortions];     // float32(1.0 - ratio) computed in double like the Python expression
    int special[kMaxDistortions];   // gamma: torch.pow's exact special cases (1: sqrt, 2: square, 3: cube)
};

__device__ __forceinline__ uint32_t gray_u8(uint32_t r, uint32_t g, uint32_t b) {
    const float l = __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, (float)r), __fmul_rn(0.587f, (float)g)), __fmul_rn(0.114f, (float)b));
    return (uint32_t)l;   // .to(uint8): truncation, l <= 254.97
}

// (ratio * a + (1 - ratio) * b).clamp(0, 255).to(uint8)
__device__ __forceinline__ uint32_t blend_u8(float ratio, float omr, float a, float b) {
    float v = __fadd_rn(__fmul_rn(ratio, a), __fmul_rn(omr, b));
    v = fminf(fmaxf(v, 0.0f), 255.0f);   // NaN cannot occur (finite factors, finite images)
    return (uint32_t)v;
}

// (x * (255 + 1 - 1e-3)).to(uint8) of convert_image_dtype, x in [0, 1]
__device__ __forceinline__ uint32_t unit_to_u8(float x) { return (uint32_t)__fmul_rn(x, 255.999f) & 0xffu; }

__device__ __forceinline__ float clamp01f(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

// adjust_hue, first half: _rgb2hsv of one pixel (x = value / 255 comes from a table), shared by every hue distortion
struct Hsv {
    float h, s, v;
};
__device__ __forceinline__ Hsv rgb2hsv(float r, float g, float b) {
    const float maxc = fmaxf(fmaxf(r, g), b), minc = fminf(fminf(r, g), b);
    const bool eqc = maxc == minc;
    const float cr = __fsub_rn(maxc, minc);
    const float s = __fdiv_rn(cr, eqc ? 1.0f : maxc);
    const float div = eqc ? 1.0f : cr;
    const float rc = __fdiv_rn(__fsub_rn(maxc, r), div), gc = __fdiv_rn(__fsub_rn(maxc, g), div), bc = __fdiv_rn(__fsub_rn(maxc, b), div);
    const float hr = maxc == r ? __fsub_rn(bc, gc) : 0.0f;
    const float hg = (maxc == g && maxc != r) ? __fsub_rn(__fadd_rn(2.0f, rc), bc) : 0.0f;
    const float hb = (maxc != g && maxc != r) ? __fsub_rn(__fadd_rn(4.0f, gc), rc) : 0.0f;
    float h = __fadd_rn(__fadd_rn(hr, hg), hb);
    h = __fadd_rn(__fdiv_rn(h, 6.0f), 1.0f);
    h = __fsub_rn(h, truncf(h));   // torch.fmod(h, 1.0): exact, |h| < 2^23
    return Hsv{h, s, maxc};
}

// second half: h = (h + hue_factor) % 1.0, _hsv2rgb, convert_image_dtype
__device__ __forceinline__ void hue_shift(const Hsv &c, float shift, uint32_t (&out)[3]) {
    // torch.remainder: fmod (exact: x - trunc(x)), then the divisor's sign
    float m = __fadd_rn(c.h, shift);
    m = __fsub_rn(m, truncf(m));
    if (m < 0.0f) m = __fadd_rn(m, 1.0f);
    const float v = c.v, s = c.s;
    const float h6 = __fmul_rn(m, 6.0f);
    const float fl = floorf(h6);
    const float f = __fsub_rn(h6, fl);
    int i = (int)fl;   // 0 .. 6
    const float p = clamp01f(__fmul_rn(v, __fsub_rn(1.0f, s)));
    const float q = clamp01f(__fmul_rn(v, __fsub_rn(1.0f, __fmul_rn(s, f))));
    const float t = clamp01f(__fmul_rn(v, __fsub_rn(1.0f, __fmul_rn(s, __fsub_rn(1.0f, f)))));
    if (i >= 6) i -= 6;   // i % 6
    float ro, go, bo;
    switch (i) {
        case 0: ro = v; go = t; bo = p; break;
        case 1: ro = q; go = v; bo = p; break;
        case 2: ro = p; go = v; bo = t; break;
        case 3: ro = p; go = q; bo = v; break;
        case 4: ro = t; go = p; bo = v; break;
        default: ro = v; go = p; bo = q; break;
    }
    out[0] = unit_to_u8(ro);
    out[1] = unit_to_u8(go);
    out[2] = unit_to_u8(bo);
}

// sum of gray(img) per image: exact integer sum, one 64-bit atomic per warp
__global__ void __launch_bounds__(256) gray_sum_kernel(const uint8_t *__restrict__ src, int64_t npix, int64_t image_stride, int64_t plane,
                                                       int layout, unsigned long long *sums) {
    const uint8_t *img = src + (int64_t)blockIdx.y * image_stride;
    unsigned long long s = 0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t r = layout == CT_HWC ? img[3 * p] : img[p];
        const uint32_t g = layout == CT_HWC ? img[3 * p + 1] : img[plane + p];
        const uint32_t b = layout == CT_HWC ? img[3 * p + 2] : img[2 * plane + p];
        s += gray_u8(r, g, b);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(sums + blockIdx.y, s);
}

// One thread takes 8 consecutive pixels per step (64-bit accesses: 3 loads, 3 stores per distortion), or single
// pixels when the buffers are not 8-byte aligned.  What depends on the pixel only - its gray value, its HSV
// triple - is computed once and reused by every distortion.
#ifndef CT_DISTORT_MINB
#define CT_DISTORT_MINB 2
#endif
template <int LAYOUT>
__global__ void __launch_bounds__(256, CT_DISTORT_MINB) distort_kernel(DistortArgs a) {
    __shared__ uint8_t lut[kMaxDistortions][256];
    __shared__ float unit[256];   // value / 255 (convert_image_dtype)
    const int64_t image = blockIdx.y;
    const uint8_t *src = a.src + image * a.src_image_stride;
    {
        const uint32_t v8 = threadIdx.x;   // blockDim.x == 256
        const float v = (float)v8;
        unit[v8] = __fdiv_rn(v, 255.0f);
        for (int k = 0; k < a.n_ops; ++k) {
            uint32_t o = v8;
            if (a.kind[k] == CT_DISTORT_BRIGHTNESS) {
                o = blend_u8(a.ratio[k], a.omr[k], v, 0.0f);
            } else if (a.kind[k] == CT_DISTORT_CONTRAST) {
                // torch.mean(gray.to(float32)): float32 sum, then / n in float32
                const float mean = __fdiv_rn((float)(double)a.gray_sum[image], (float)a.npix);
                o = blend_u8(a.ratio[k], a.omr[k], v, mean);
            } else if (a.kind[k] == CT_DISTORT_GAMMA) {
                const float x = __fdiv_rn(v, 255.0f);
                float y;
                if (a.special[k] == 1) y = __fsqrt_rn(x);
                else if (a.special[k] == 2) y = __fmul_rn(x, x);
                else if (a.special[k] == 3) y = __fmul_rn(__fmul_rn(x, x), x);
                else y = (float)pow((double)x, (double)a.ratio[k]);
                o = unit_to_u8(clamp01f(y));
            }
            lut[k][v8] = (uint8_t)o;
        }
    }
    __syncthreads();
    const bool any_hue = a.any_hue != 0, any_sat = a.any_sat != 0;

    // one pixel under distortion k; gr / hsv are its shared precomputations
    auto pixel = [&](int k, uint32_t r, uint32_t g, uint32_t b, float gr, const Hsv &hsv, uint32_t (&o)[3]) {
        const int kind = a.kind[k];
        if (kind == CT_DISTORT_SATURATION) {
            o[0] = blend_u8(a.ratio[k], a.omr[k], (float)r, gr);
            o[1] = blend_u8(a.ratio[k], a.omr[k], (float)g, gr);
            o[2] = blend_u8(a.ratio[k], a.omr[k], (float)b, gr);
        } else if (kind == CT_DISTORT_HUE) {
            hue_shift(hsv, a.ratio[k], o);
        } else {
            o[0] = lut[k][r];
            o[1] = lut[k][g];
            o[2] = lut[k][b];
        }
    };

    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t first_scalar = 0;
    if (a.vec) {
        constexpr int W = 8;   // pixels per step
        const int64_t ngroups = a.npix / W;
        for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < ngroups; q += stride) {
            uint32_t w[6];   // HWC: 24 interleaved bytes; CHW: 8 bytes of each plane
            if (LAYOUT == CT_HWC) {
                const uint2 *s64 = reinterpret_cast<const uint2 *>(src) + 3 * q;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const uint2 t = s64[j];
                    w[2 * j] = t.x;
                    w[2 * j + 1] = t.y;
                }
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const uint2 t = reinterpret_cast<const uint2 *>(src + c * a.src_plane)[q];
                    w[2 * c] = t.x;
                    w[2 * c + 1] = t.y;
                }
            }
            auto byte_of = [&](int i, int c) { return LAYOUT == CT_HWC ? 3 * i + c : W * c + i; };
            float gr[W];
            Hsv hsv[W];
#pragma unroll
            for (int i = 0; i < W; ++i) {
                uint32_t v[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int byte = byte_of(i, c);
                    v[c] = (w[byte >> 2] >> (8 * (byte & 3))) & 0xffu;
                }
                gr[i] = any_sat ? (float)gray_u8(v[0], v[1], v[2]) : 0.0f;
                hsv[i] = any_hue ? rgb2hsv(unit[v[0]], unit[v[1]], unit[v[2]]) : Hsv{0.0f, 0.0f, 0.0f};
            }
            for (int k = 0; k < a.n_ops; ++k) {
                uint32_t ow[6] = {0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
                for (int i = 0; i < W; ++i) {
                    uint32_t v[3], o[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int byte = byte_of(i, c);
                        v[c] = (w[byte >> 2] >> (8 * (byte & 3))) & 0xffu;
                    }
                    pixel(k, v[0], v[1], v[2], gr[i], hsv[i], o);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int byte = byte_of(i, c);
                        ow[byte >> 2] |= o[c] << (8 * (byte & 3));
                    }
                }
                uint8_t *d = a.dst + (image * a.n_ops + k) * a.dst_image_stride;
                if (LAYOUT == CT_HWC) {
                    uint2 *d64 = reinterpret_cast<uint2 *>(d) + 3 * q;
#pragma unroll
                    for (int j = 0; j < 3; ++j) d64[j] = make_uint2(ow[2 * j], ow[2 * j + 1]);
                } else {
#pragma unroll
                    for (int c = 0; c < 3; ++c) reinterpret_cast<uint2 *>(d + c * a.dst_plane)[q] = make_uint2(ow[2 * c], ow[2 * c + 1]);
                }
            }
        }
        first_scalar = ngroups * W;
    }
    for (int64_t p = first_scalar + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.npix; p += stride) {
        const uint32_t r = LAYOUT == CT_HWC ? src[3 * p] : src[p];
        const uint32_t g = LAYOUT == CT_HWC ? src[3 * p + 1] : src[a.src_plane + p];
        const uint32_t b = LAYOUT == CT_HWC ? src[3 * p + 2] : src[2 * a.src_plane + p];
        const float gr = any_sat ? (float)gray_u8(r, g, b) : 0.0f;
        const Hsv hsv = any_hue ? rgb2hsv(unit[r], unit[g], unit[b]) : Hsv{0.0f, 0.0f, 0.0f};
        for (int k = 0; k < a.n_ops; ++k) {
            uint32_t o[3];
            pixel(k, r, g, b, gr, hsv, o);
            uint8_t *d = a.dst + (image * a.n_ops + k) * a.dst_image_stride;
            if (LAYOUT == CT_HWC) {
                d[3 * p] = (uint8_t)o[0]; d[3 * p + 1] = (uint8_t)o[1]; d[3 * p + 2] = (uint8_t)o[2];
            } else {
                d[p] = (uint8_t)o[0]; d[a.dst_plane + p] = (uint8_t)o[1]; d[2 * a.dst_plane + p] = (uint8_t)o[2];
            }
        }
    }
}

int launch_distort(ct_context *h, const ct_batch *src, const ct_distortion *ops, int n_ops, const ct_batch *dst) {
    CT_TRY(check_batch(h, src, "src"));
    CT_TRY(check_batch(h, dst, "dst"));
    if (!ops || n_ops < 1 || n_ops > kMaxDistortions) return fail(h, CT_E_INVALID, "n_ops must be in [1, %d]", kMaxDistortions);
    if (src->dtype != CT_U8 || dst->dtype != CT_U8) return fail(h, CT_E_UNSUPPORTED, "the distortion generator takes uint8 images (ref: utils/data.py:101-104)");
    if (src->layout != dst->layout || src->npix != dst->npix) return fail(h, CT_E_INVALID, "src / dst layouts or sizes differ");
    if (dst->count != src->count * n_ops) return fail(h, CT_E_INVALID, "dst must hold count * n_ops images");
    DistortArgs a{};
    a.src = static_cast<const uint8_t *>(src->data);
    a.dst = static_cast<uint8_t *>(dst->data);
    a.npix = src->npix;
    a.src_image_stride = src->count > 1 ? src->image_stride : 0;
    a.dst_image_stride = dst->count > 1 ? dst->image_stride : 0;
    a.src_plane = plane_of(src);
    a.dst_plane = plane_of(dst);
    a.layout = src->layout;
    a.n_ops = n_ops;
    auto aligned8 = [](const ct_batch *b, int64_t image_stride) {
        if (((uintptr_t)b->data) & 7) return false;
        if (image_stride & 7) return false;
        if (b->layout == CT_CHW && (plane_of(b) & 7)) return false;
        return true;
    };
    a.vec = aligned8(src, a.src_image_stride) && aligned8(dst, a.dst_image_stride) ? 1 : 0;
    bool contrast = false;
    for (int k = 0; k < n_ops; ++k) {
        const int kind = ops[k].kind;
        const double f = ops[k].factor;
        if (kind < CT_DISTORT_IDENTITY || kind > CT_DISTORT_GAMMA) return fail(h, CT_E_INVALID, "distortion %d: unknown kind %d", k, kind);
        if (!(f == f) || f > 1e30 || f < -1e30) return fail(h, CT_E_INVALID, "distortion %d: factor is not finite", k);
        // torchvision's argument checks
        if (kind == CT_DISTORT_HUE && !(f >= -0.5 && f <= 0.5)) return fail(h, CT_E_INVALID, "hue_factor (%g) is not in [-0.5, 0.5].", f);
        if (kind != CT_DISTORT_HUE && kind != CT_DISTORT_IDENTITY && f < 0) return fail(h, CT_E_INVALID, "distortion %d: factor (%g) is not non-negative.", k, f);
        a.kind[k] = kind;
        a.ratio[k] = (float)f;
        a.omr[k] = (float)(1.0 - f);
        a.special[k] = 0;
        if (kind == CT_DISTORT_GAMMA) a.special[k] = f == 0.5 ? 1 : (f == 2.0 ? 2 : (f == 3.0 ? 3 : 0));
        contrast |= kind == CT_DISTORT_CONTRAST;
        a.any_hue |= kind == CT_DISTORT_HUE;
        a.any_sat |= kind == CT_DISTORT_SATURATION;
    }
    const int64_t units = a.vec ? (a.npix + 7) / 8 : a.npix;
    int64_t blocks = (units + 255) / 256;
    const int64_t cap = ((int64_t)h->sm_count * 8 + src->count - 1) / src->count;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (contrast) {
        CT_TRY(ensure_scratch(h, src->count));
        unsigned long long *sums = reinterpret_cast<unsigned long long *>(h->sums);   // >= 20 doubles per pair
        CT_CUDA(h, cudaMemsetAsync(sums, 0, sizeof(unsigned long long) * (size_t)src->count, h->stream));
        int64_t gb = (a.npix + 255) / 256;
        if (gb > cap) gb = cap;
        gray_sum_kernel<<<dim3((unsigned)gb, (unsigned)src->count), 256, 0, h->stream>>>(a.src, a.npix, a.src_image_stride, a.src_plane,
                                                                                      a.layout, sums);
        h->launches++;
        CT_CUDA(h, cudaGetLastError());
        a.gray_sum = sums;
    }
    if (a.layout == CT_HWC) distort_kernel<CT_HWC><<<dim3((unsigned)blocks, (unsigned)src->count), 256, 0, h->stream>>>(a);
    else distort_kernel<CT_CHW><<<dim3((unsigned)blocks, (unsigned)src->count), 256, 0, h->stream>>>(a);
    h->launches++;
    CT_CUDA(h, cudaGetLastError());
    return CT_OK;
}

}  // namespace ct
