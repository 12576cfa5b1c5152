// 8-bit frame I/O: uint8 -> float on the way in, clip + round to uint8 on
// the way out, so that a video pipeline moves 3 bytes per pixel over PCIe instead of 12-24.
//   in : x = k / 255 exactly as the reference's loaders produce it - float32 (torch `/ 255`,
//        ref: utils/data.py:106) or float64 (skimage.img_as_float, ref: utils/postprocess.py:138)
//   out: np.rint(np.clip(y, 0, 1) * 255).astype(uint8)  (img_as_ubyte of the clipped result,
//        ref: utils/postprocess.py:138; Runner clamps the same way, ref: methods/__init__.py:30)
// Since round 2 the transfer kernels decode and encode uint8 images themselves (PixelIO<uint8_t, .>,
// ct_common.cuh); these streaming conversions remain for the one case without a float64 state to
// convert from (a one-iteration IDT on uint8 frames through the host API).
#include "ct_context.h"

namespace ct {

template <typename T>
__global__ void __launch_bounds__(256) u8_to_float_kernel(const uint8_t *__restrict__ in, T *__restrict__ out, int64_t n) {
    __shared__ T lut[256];
    lut[threadIdx.x] = (T)((double)threadIdx.x / 255.0);   // (float)(k/255.0) == float32 k/255 for every k
    __syncthreads();
    const int64_t n16 = n / 16, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 v = reinterpret_cast<const uint4 *>(in)[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        T r[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) r[k] = lut[(w[k >> 2] >> (8 * (k & 3))) & 0xff];
        constexpr int per = 16 / sizeof(T);   // elements per 16-byte store
#pragma unroll
        for (int k = 0; k < 16 / per; ++k)
            reinterpret_cast<uint4 *>(out + 16 * i)[k] = *reinterpret_cast<uint4 *>(&r[k * per]);
    }
    for (int64_t i = n16 * 16 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = lut[in[i]];
}

template <typename T>
__global__ void __launch_bounds__(256) float_to_u8_kernel(const T *__restrict__ in, uint8_t *__restrict__ out, int64_t n) {
    const int64_t n16 = n / 16, stride = (int64_t)gridDim.x * blockDim.x;
    auto q = [](T y) -> uint32_t {
        double c = (double)y;
        c = c < 0.0 ? 0.0 : (c > 1.0 ? 1.0 : c);            // np.clip; NaN -> 0 after the cast below
        return (uint32_t)__double2int_rn(c * 255.0) & 0xffu;   // np.rint: round half to even
    };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 16; ++k) w[k >> 2] |= q(in[16 * i + k]) << (8 * (k & 3));
        reinterpret_cast<uint4 *>(out)[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    for (int64_t i = n16 * 16 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (uint8_t)q(in[i]);
}

int launch_u8_to_float(ct_context *h, const uint8_t *in, void *out, int dtype, int64_t n) {
    if (!in || !out || n <= 0) return fail(h, CT_E_INVALID, "bad u8 conversion arguments");
    if ((((uintptr_t)in) | ((uintptr_t)out)) & 15) return fail(h, CT_E_INVALID, "u8 conversion buffers must be 16-byte aligned");
    const int grid = h->sm_count * 8;
    if (dtype == CT_F32) u8_to_float_kernel<float><<<grid, 256, 0, h->stream>>>(in, static_cast<float *>(out), n);
    else u8_to_float_kernel<double><<<grid, 256, 0, h->stream>>>(in, static_cast<double *>(out), n);
    h->launches++;
    CT_CUDA(h, cudaGetLastError());
    return CT_OK;
}

int launch_float_to_u8(ct_context *h, const void *in, int dtype, uint8_t *out, int64_t n) {
    if (!in || !out || n <= 0) return fail(h, CT_E_INVALID, "bad u8 conversion arguments");
    if ((((uintptr_t)in) | ((uintptr_t)out)) & 15) return fail(h, CT_E_INVALID, "u8 conversion buffers must be 16-byte aligned");
    const int grid = h->sm_count * 8;
    if (dtype == CT_F32) float_to_u8_kernel<float><<<grid, 256, 0, h->stream>>>(static_cast<const float *>(in), out, n);
    else float_to_u8_kernel<double><<<grid, 256, 0, h->stream>>>(static_cast<const double *>(in), out, n);
    h->launches++;
    CT_CUDA(h, cudaGetLastError());
    return CT_OK;
}

}  // namespace ct
