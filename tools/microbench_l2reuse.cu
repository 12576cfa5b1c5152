// Does a second streaming pass over a buffer larger than L2 hit in L2 if it runs in the opposite
// direction (the tail of the first pass is still resident)?  199 MB = the fp64 state of one 4K frame.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) pass(const double2* __restrict__ p, size_t n, int reverse, double* out) {
    double s = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t chunks = (n + stride - 1) / stride;
    for (size_t c = 0; c < chunks; ++c) {
        const size_t cc = reverse ? chunks - 1 - c : c;
        const size_t i = cc * stride + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i < n) { const double2 v = p[i]; s += v.x + v.y; }
    }
    if (s == 1.2345) out[0] = s;
}
__global__ void __launch_bounds__(256) rw(double2* p, size_t n, int reverse) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t chunks = (n + stride - 1) / stride;
    for (size_t c = 0; c < chunks; ++c) {
        const size_t cc = reverse ? chunks - 1 - c : c;
        const size_t i = cc * stride + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i < n) { double2 v = p[i]; v.x += 1.0; p[i] = v; }
    }
}
int main() {
    int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    for (double mb : {60.0, 100.0, 199.0, 300.0}) {
        const size_t n = (size_t)(mb * 1e6 / 16);
        double2* buf; cudaMalloc(&buf, n * 16); cudaMemset(buf, 0, n * 16);
        double2* other; cudaMalloc(&other, 400u << 20); cudaMemset(other, 0, 400u << 20);
        double* out; cudaMalloc(&out, 64);
        cudaEvent_t e[4]; for (auto& x : e) cudaEventCreate(&x);
        float fwd_fwd = 0, fwd_rev = 0, w_rev_then_fwd = 0, cold = 0;
        const int grid = sm * 8;
        for (int rep = 0; rep < 6; ++rep) {
            // cold: flush with another buffer, then read
            pass<<<grid, 256>>>(other, (400u << 20) / 16, 0, out);
            cudaEventRecord(e[0]); pass<<<grid, 256>>>(buf, n, 0, out); cudaEventRecord(e[1]);
            pass<<<grid, 256>>>(buf, n, 0, out); cudaEventRecord(e[2]);             // forward after forward
            pass<<<grid, 256>>>(buf, n, 1, out); cudaEventRecord(e[3]);             // reverse after forward
            cudaEventSynchronize(e[3]);
            float a, b, c; cudaEventElapsedTime(&a, e[0], e[1]); cudaEventElapsedTime(&b, e[1], e[2]); cudaEventElapsedTime(&c, e[2], e[3]);
            if (rep) { cold += a; fwd_fwd += b; fwd_rev += c; }
            // read-modify-write in reverse, then a forward read (remap -> hist)
            pass<<<grid, 256>>>(other, (400u << 20) / 16, 0, out);
            rw<<<grid, 256>>>(buf, n, 1);
            cudaEventRecord(e[0]); pass<<<grid, 256>>>(buf, n, 0, out); cudaEventRecord(e[1]); cudaEventSynchronize(e[1]);
            cudaEventElapsedTime(&a, e[0], e[1]); if (rep) w_rev_then_fwd += a;
        }
        const double gb = n * 16 / 1e9;
        printf("%5.0f MB: cold read %.0f GB/s | fwd after fwd %.0f GB/s | REVERSE after fwd %.0f GB/s | fwd read after reverse RMW %.0f GB/s\n",
               mb, gb / (cold / 5e3), gb / (fwd_fwd / 5e3), gb / (fwd_rev / 5e3), gb / (w_rev_then_fwd / 5e3));
        cudaFree(buf); cudaFree(other); cudaFree(out);
    }
    return 0;
}
