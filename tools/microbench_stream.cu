// Streaming-read ceilings on the B200: plain LDG.128 grid-stride vs the TMA bulk-copy pipeline
// of color-transfer_b200/csrc/ct_pipe.cuh, for several CTA counts / pipeline depths.
#include <cstdio>
#include <cuda_runtime.h>
#include "../color-transfer_b200/csrc/ct_pipe.cuh"
using namespace ct;

__global__ void __launch_bounds__(256) ldg_read(const double2* __restrict__ p, size_t n, double* out) {
    double s = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * 4) {
        double2 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) if (i + k * stride < n) v[k] = p[i + k * stride]; else v[k] = make_double2(0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) s += v[k].x + v[k].y;
    }
    if (s == 1.2345) out[0] = s;
}

template <int S>
__global__ void __launch_bounds__(256) tma_read(const double* img, int64_t plane, int ntiles, double* out) {
    extern __shared__ __align__(16) unsigned char sm[];
    Pipe<S> pipe(sm);
    if (threadIdx.x == 0) pipe.init();
    __syncthreads();
    double s = 0;
    using IO = PixelIO<double, CT_CHW>;
    pipe_for_each_group<IO>(pipe, img, plane, ntiles, blockIdx.x, gridDim.x, [&](const IO::Raw& raw, int64_t) {
        s += raw.e[0] + raw.e[1] + raw.e[2] + raw.e[3] + raw.e[4] + raw.e[5];
    });
    if (s == 1.2345) out[0] = s;
}

template <class F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f(); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int i = 0; i < 5; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5;
}

template <int S> void run_tma(const double* buf, int64_t plane, int ntiles, double* out, double gb, int sm) {
    cudaFuncSetAttribute(tma_read<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int per : {1, 2, 3, 4, 6}) {
        if ((size_t)per * pipe_bytes(S) > 220 * 1024) continue;
        float ms = timeit([&] { tma_read<S><<<sm * per, 256, pipe_bytes(S)>>>(buf, plane, ntiles, out); });
        printf("TMA  stages=%d ctas/SM=%d  in flight/SM=%3d KB  %.0f GB/s  (%s)\n", S, per, per * S * 12, gb / (ms / 1e3),
               cudaGetErrorString(cudaGetLastError()));
    }
}

int main() {
    int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    const int64_t plane = 100 * 1000 * 1000 / 2 * 2;          // 3 planes x 100 M doubles = 2.4 GB
    double* buf; cudaMalloc(&buf, 3 * plane * 8); cudaMemset(buf, 0, 3 * plane * 8);
    double* out; cudaMalloc(&out, 64);
    const double gb = 3.0 * plane * 8 / 1e9;
    for (int per : {2, 4, 8}) {
        float ms = timeit([&] { ldg_read<<<sm * per, 256>>>((const double2*)buf, (size_t)3 * plane / 2, out); });
        printf("LDG.128 x4 grid-stride, %d CTAs/SM: %.0f GB/s\n", per, gb / (ms / 1e3));
    }
    const int ntiles = (int)(plane / (kThreads * 2));
    run_tma<2>(buf, plane, ntiles, out, gb, sm);
    run_tma<3>(buf, plane, ntiles, out, gb, sm);
    run_tma<4>(buf, plane, ntiles, out, gb, sm);
    run_tma<6>(buf, plane, ntiles, out, gb, sm);
    // copy ceiling for reference
    double* dst; cudaMalloc(&dst, 3 * plane * 8);
    float ms = timeit([&] { cudaMemcpyAsync(dst, buf, 3 * plane * 8, cudaMemcpyDeviceToDevice); });
    printf("cudaMemcpy D2D: %.0f GB/s (read+write)\n", 2 * gb / (ms / 1e3));
    return 0;
}
