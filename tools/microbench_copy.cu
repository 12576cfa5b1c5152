// What can K7's access pattern stream?  K7 reads the fp64 state (24 B/px) through the TMA tile pipeline and
// writes the new state (24 B/px) with 128-bit stores: a copy.  This microbenchmark runs that copy without any
// arithmetic, for the layouts and store paths K7 could use, against cudaMemcpyAsync D2D on the same buffers:
//   planar  : 3 bulk loads of 4 KB per 512-pixel tile (one per channel plane), STG.128 per plane   (K7 today)
//   packed  : 1 bulk load of 12 KB per tile (interleaved state), STG.128 at stride 48 B
//   planar+T: planar loads, result staged in shared memory and written by 3 TMA bulk stores of 4 KB
//   packed+T: packed load, one TMA bulk store of 12 KB
// 4 images of 3840x2160 per launch, grid (blocks, 4), 256 threads, like the product's launches.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o microbench_copy tools/microbench_copy.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kThreads = 256, kWarps = 8, kTile = kThreads * 48;   // 12 KB: 512 fp64 pixels

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_arrive_last(uint32_t bar) {
    uint64_t state;
    uint32_t pending;
    asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(state) : "r"(bar) : "memory");
    asm volatile("mbarrier.pending_count.b64 %0, %1;" : "=r"(pending) : "l"(state));
    return pending == 1u;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// PLANAR: image = 3 planes of npix doubles; else npix x 3 doubles interleaved.  TSTORE: write through shared memory + TMA.
template <int STAGES, bool PLANAR, bool TSTORE>
__global__ void __launch_bounds__(kThreads, 2) copy_kernel(const double *src, double *dst, int64_t npix, uint32_t zero) {
    extern __shared__ __align__(128) unsigned char sm[];
    const uint32_t stage = smem_u32(sm), out = stage + STAGES * kTile, full = out + (TSTORE ? 2 * kTile : 0), empty = full + 8 * STAGES;
    const double *s = src + (int64_t)blockIdx.y * 3 * npix;
    double *d = dst + (int64_t)blockIdx.y * 3 * npix;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(full + 8 * i, 1);
            mbar_init(empty + 8 * i, kWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int ntiles = (int)(npix / 512);
    const int mine = (int)blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto issue = [&](int i, int st) {
        const int64_t p0 = (int64_t)(blockIdx.x + (int64_t)i * gridDim.x) * 512;
        const uint32_t dsts = stage + st * kTile, bar = full + 8 * st;
        mbar_expect_tx(bar, kTile);
        if (PLANAR) {
            for (int c = 0; c < 3; ++c) bulk_g2s(dsts + c * (kTile / 3), s + c * npix + p0, kTile / 3, bar);
        } else {
            bulk_g2s(dsts, s + 3 * p0, kTile, bar);
        }
    };
    if (threadIdx.x == 0)
        for (int i = 0; i < STAGES && i < mine; ++i) issue(i, i);
    int st = 0;
    uint32_t parity = 0;
    for (int i = 0; i < mine; ++i) {
        mbar_wait(full + 8 * st, parity);
        const uint32_t base = stage + st * kTile + (PLANAR ? 16u : 48u) * threadIdx.x;
        uint4 v[3];
        for (int k = 0; k < 3; ++k) v[k] = lds128(base + (PLANAR ? k * (kTile / 3) : 16 * k));
        const uint32_t dep = v[0].x ^ v[1].x ^ v[2].x;
        __syncwarp();
        if ((threadIdx.x & 31) == 0 && mbar_arrive_last(empty + 8 * st + (dep & zero)) && i + STAGES < mine) issue(i + STAGES, st);
        const int64_t p0 = (int64_t)(blockIdx.x + (int64_t)i * gridDim.x) * 512;
        if (!TSTORE) {
            if (PLANAR) {
                for (int k = 0; k < 3; ++k) reinterpret_cast<uint4 *>(d + k * npix + p0)[threadIdx.x] = v[k];
            } else {
                for (int k = 0; k < 3; ++k) reinterpret_cast<uint4 *>(d + 3 * p0)[3 * threadIdx.x + k] = v[k];
            }
        } else {
            // two output buffers: wait until the bulk store that last read this one has finished reading
            const uint32_t ob = out + (i & 1) * kTile;
            if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncthreads();
            for (int k = 0; k < 3; ++k) sts128(ob + (PLANAR ? 16u : 48u) * threadIdx.x + (PLANAR ? k * (kTile / 3) : 16 * k), v[k]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (threadIdx.x == 0) {
                if (PLANAR) {
                    for (int c = 0; c < 3; ++c) bulk_s2g(d + c * npix + p0, ob + c * (kTile / 3), kTile / 3);
                } else {
                    bulk_s2g(d + 3 * p0, ob, kTile);
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (++st == STAGES) {
            st = 0;
            parity ^= 1u;
        }
    }
    if (TSTORE && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int STAGES, bool PLANAR, bool TSTORE>
static void run(const char *name, const double *src, double *dst, int64_t npix, int images, int sms, int cap = 2) {
    const size_t smem = (size_t)STAGES * kTile + (TSTORE ? 2 * kTile : 0) + 16 * STAGES;
    cudaFuncSetAttribute(copy_kernel<STAGES, PLANAR, TSTORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, copy_kernel<STAGES, PLANAR, TSTORE>, kThreads, smem);
    if (occ > cap) occ = cap;   // K7 keeps 2 CTAs per SM (128 registers)
    const int blocks = sms * occ / images;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) copy_kernel<STAGES, PLANAR, TSTORE><<<dim3(blocks, images), kThreads, smem>>>(src, dst, npix, 0u);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) copy_kernel<STAGES, PLANAR, TSTORE><<<dim3(blocks, images), kThreads, smem>>>(src, dst, npix, 0u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const cudaError_t err = cudaGetLastError();
    const double gb = 2.0 * images * npix * 24 / 1e9;
    printf("%-28s stages %d  %d CTAs/SM  %7.1f us  %6.0f GB/s%s\n", name, STAGES, occ, ms / reps * 1e3, gb / (ms / reps) * 1e3,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int images = 4;
    const int64_t npix = 3840LL * 2160;
    const size_t bytes = (size_t)images * npix * 24;
    double *a, *b;
    cudaMalloc(&a, bytes);
    cudaMalloc(&b, bytes);
    cudaMemset(a, 1, bytes);
    cudaMemset(b, 0, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-28s %7.1f us  %6.0f GB/s\n", "cudaMemcpyAsync D2D", ms / 20 * 1e3, 2.0 * bytes / 1e9 / (ms / 20) * 1e3);
    run<3, true, false>("planar, STG", a, b, npix, images, sms);
    run<5, true, false>("planar, STG", a, b, npix, images, sms);
    run<3, false, false>("packed, STG stride 48", a, b, npix, images, sms);
    run<5, false, false>("packed, STG stride 48", a, b, npix, images, sms);
    run<3, true, true>("planar, TMA store", a, b, npix, images, sms);
    run<3, false, true>("packed, TMA store", a, b, npix, images, sms);
    run<5, false, true>("packed, TMA store", a, b, npix, images, sms);
    // in place (what the state update does): dst == src
    run<3, true, false>("planar, STG, 4 CTAs/SM", a, b, npix, images, sms, 4);
    run<3, false, true>("packed, TMA store, 4 CTAs/SM", a, b, npix, images, sms, 4);
    run<3, true, false>("planar, STG, in place", a, a, npix, images, sms);
    run<3, false, true>("packed, TMA store, in place", a, a, npix, images, sms);
    cudaDeviceSynchronize();
    return 0;
}
