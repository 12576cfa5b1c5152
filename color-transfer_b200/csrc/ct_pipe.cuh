// Streaming tile pipeline: TMA bulk copies (cp.async.bulk, SASS UBLKCP) fill a ring of
// shared-memory stages, mbarriers (complete_tx) signal arrival, consumers hand a stage back
// through a second mbarrier.  The bytes in flight live in shared memory instead of registers:
// kStages x 12 KB per CTA are outstanding while every thread computes on 12 registers of raw
// pixels, which is what lets these fp64-heavy kernels hide HBM latency at 2-3 CTAs per SM.
//
// A tile is kThreads pixel groups (1024 f32 pixels, 512 f64 pixels or 4096 uint8 pixels = 12 KB).
// Interleaved images need one bulk copy per tile, planar images three (one per channel plane).
// Thread i then reads its group with conflict-free 128-bit LDS (stride 48 B -> the 16 B bank groups
// 3i mod 8 of a quarter warp are distinct; planar: stride 16 B).  A uint8 tile is four sub-tiles of
// 1024 pixels; thread i takes pixels 4i..4i+3 of each (32-bit LDS at stride 12 B / 4 B: conflict
// free), so that what it writes per sub-group is contiguous across the warp like a float tile.
#pragma once

#include "ct_common.cuh"

#ifndef CT_PIPE_UNROLL
#define CT_PIPE_UNROLL 0   // 1: tile loop unrolled over the stages (compile-time stage index: ~6 % fewer instructions per tile; measured: no change, 4.18 ms per pass either way - the loops are not issue-bound)
#endif
#ifndef CT_PIPE_LAST_REFILLS
#define CT_PIPE_LAST_REFILLS 1   // 0: thread 0 waits for the stage to be released and refills it (round 1)
#endif

namespace ct {

constexpr int kTileBytes = kThreads * 48;  // 12 KB for every (dtype, layout)
constexpr int pipe_bytes(int stages) { return stages * kTileBytes + 2 * stages * 8; }  // a multiple of 16 B

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// All helpers take 32-bit shared-window addresses: converting a generic pointer costs ~5
// instructions (S2UR CgaCtaId, LEA, ...) and the tile loop would redo it for every stage.
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive, and tell whether this arrival completed the phase (the returned state is the one BEFORE the arrive)
__device__ __forceinline__ bool mbar_arrive_last(uint32_t bar) {
    uint64_t state;
    uint32_t pending;
    asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(state) : "r"(bar) : "memory");
    asm volatile("mbarrier.pending_count.b64 %0, %1;" : "=r"(pending) : "l"(state));
    return pending == 1u;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// 128-bit / 64-bit loads from a 32-bit shared-window address
template <typename V>
__device__ __forceinline__ V lds_vec(uint32_t addr);
template <>
__device__ __forceinline__ float4 lds_vec<float4>(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <>
__device__ __forceinline__ uint4 lds_vec<uint4>(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
template <>
__device__ __forceinline__ double2 lds_vec<double2>(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

// Shared-memory view of one CTA's pipeline (carved from dynamic shared memory, 16 B aligned):
// kStages tiles, then per barrier set kStages "full" and kStages "empty" mbarriers.
template <int kStages>
struct Pipe {
    static constexpr int kNumStages = kStages;
    uint32_t stage;  // shared-window address of kStages * kTileBytes
    uint32_t full;   // [kStages] x 8 B
    uint32_t empty;  // [kStages] x 8 B
    // `set` selects one of several barrier sets laid out after the stages, so that a kernel can run
    // several pipelines one after the other over the same stage buffers (each starts at phase 0)
    __device__ __forceinline__ explicit Pipe(void *base, int set = 0)
        : stage(smem_u32(base)), full(stage + kStages * kTileBytes + 16 * kStages * set), empty(full + 8 * kStages) {}
    // one thread; followed by __syncthreads() in the caller
    __device__ __forceinline__ void init() {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full + 8 * s, 1);
            mbar_init(empty + 8 * s, kWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
};

// first pixel of sub-group s of this thread's group in a pipelined tile that starts at pixel t0
template <typename IO>
__device__ __forceinline__ int64_t pipe_sub_pixel0(int64_t t0, int s) {
    return IO::kU8 ? t0 + s * (kThreads * IO::GS) + (int64_t)threadIdx.x * IO::GS : t0 + (int64_t)threadIdx.x * IO::G;
}

// byte offset of group gi (in [0, kThreads)) inside a stage
template <typename IO>
__device__ __forceinline__ uint32_t pipe_group_offset(uint32_t gi) {
    if (IO::kU8) return IO::kLayout == CT_HWC ? 12u * gi : 4u * gi;
    return IO::kLayout == CT_HWC ? 48u * gi : 16u * gi;
}
// one group of a stage -> registers; `src` = stage address + pipe_group_offset
template <typename IO>
__device__ __forceinline__ typename IO::Raw pipe_load_group(uint32_t src) {
    constexpr int G = IO::G;
    typename IO::Raw raw;
    if (IO::kU8) {
        uint32_t *w = reinterpret_cast<uint32_t *>(raw.e);
        if (IO::kLayout == CT_HWC) {
#pragma unroll
            for (int q = 0; q < 4; ++q)      // sub-tile q: 3072 bytes, this thread's 12 bytes
#pragma unroll
                for (int k = 0; k < 3; ++k) w[3 * q + k] = lds_u32(src + q * (kTileBytes / 4) + 4 * k);
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c)      // plane c: 4096 bytes, sub-tile q: 1024 bytes
#pragma unroll
                for (int q = 0; q < 4; ++q) w[4 * c + q] = lds_u32(src + c * (kTileBytes / 3) + q * (kTileBytes / 12));
        }
    } else {
        using V = typename IO::V;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            *reinterpret_cast<V *>(&raw.e[k * G]) = lds_vec<V>(src + (IO::kLayout == CT_HWC ? 16 * k : k * (kTileBytes / 3)));
    }
    return raw;
}

// Runs f(raw, first_pixel_of_tile) on this thread's pixel group of the FULL tiles first_tile,
// first_tile + tile_stride, ... of one image (pipe_sub_pixel0 locates its sub-groups); every thread of
// the CTA must call it (block-uniform trip count).  `pipe` must be freshly initialised (phase 0) for
// each call.
// SPLIT = 2: the two halves of the CTA do DIFFERENT work on the SAME pixels (K4 gives each half two of
// the four rotations): every thread takes groups t % 128 and t % 128 + 128 of each tile and f is
// called as f(raw, first_pixel_of_tile, u) for u = 0, 1.
//
// `zero` must be 0 and must come from a kernel argument (a value the compiler cannot know).  A warp
// hands a stage back with an mbarrier arrive right after its shared-memory loads were ISSUED; the
// arrive is not ordered behind those loads in hardware, so when the refill is fast (an L2-resident
// image, an idle copy engine) the bulk copy of tile i + stages could land in the stage while a
// quarter-warp wavefront of the load of tile i was still pending: a few pixels of tile i + stages
// processed as tile i (seen as ~5 % wrong calls on 1080x860 float64 pairs).  The arrive's address is
// therefore computed from the loaded registers (`& zero`), which makes it wait for the loads.
template <typename IO, int SPLIT = 1, typename P, typename F>
__device__ __forceinline__ void pipe_for_each_group(P &pipe, const typename IO::elem_t *img, int64_t plane,
                                                    int ntiles, int first_tile, int tile_stride, uint32_t zero, F &&f) {
    using T = typename IO::elem_t;
    constexpr int kStages = P::kNumStages;
    constexpr int G = IO::G;
    constexpr int kTilePx = kThreads * G;
    const int mine = first_tile < ntiles ? (ntiles - first_tile + tile_stride - 1) / tile_stride : 0;
    auto issue = [&](int i, int s) {  // one thread; s == i % kStages
        const int64_t p0 = (int64_t)(first_tile + (int64_t)i * tile_stride) * kTilePx;
        const uint32_t dst = pipe.stage + s * kTileBytes, bar = pipe.full + 8 * s;
        mbar_expect_tx(bar, kTileBytes);
        if (IO::kLayout == CT_HWC) {
            bulk_g2s(dst, img + 3 * p0, kTileBytes, bar);
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) bulk_g2s(dst + c * (kTileBytes / 3), img + c * plane + p0, kTileBytes / 3, bar);
        }
    };
    if (threadIdx.x == 0)
        for (int i = 0; i < kStages && i < mine; ++i) issue(i, i);
    // this thread's group(s) inside a stage
    uint32_t my_off[SPLIT];
#pragma unroll
    for (int u = 0; u < SPLIT; ++u)
        my_off[u] = pipe_group_offset<IO>(SPLIT == 1 ? threadIdx.x : (threadIdx.x % (kThreads / SPLIT)) + u * (kThreads / SPLIT));
#if CT_PIPE_UNROLL
    // The tile loop is unrolled over the stages: the stage index is a compile-time constant inside the body, so the
    // stage and barrier addresses are immediates off one base register and the phase parity flips once per round
    // (the rolled loop spent ~10 of its ~40 bookkeeping instructions per tile on them).
    uint32_t parity = 0;
    for (int i0 = 0; i0 < mine; i0 += kStages) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            const int i = i0 + s;
            if (i >= mine) break;
            const uint32_t full = pipe.full + 8 * s, empty = pipe.empty + 8 * s;
            mbar_wait(full, parity);
            typename IO::Raw raw[SPLIT];
#pragma unroll
            for (int u = 0; u < SPLIT; ++u) raw[u] = pipe_load_group<IO>(pipe.stage + s * kTileBytes + my_off[u]);
            uint32_t dep = 0;   // one word of every load instruction's result
#pragma unroll
            for (int u = 0; u < SPLIT; ++u) {
                const uint32_t *w = reinterpret_cast<const uint32_t *>(raw[u].e);
                if (IO::kU8) {
#pragma unroll
                    for (int k = 0; k < 12; ++k) dep ^= w[k];
                } else {
#pragma unroll
                    for (int k = 0; k < 3; ++k) dep ^= w[4 * k];
                }
            }
            __syncwarp();
#if CT_PIPE_LAST_REFILLS
            // this warp has copied its groups out; the warp whose arrival completes the "empty" phase - the last of
            // the 8 to let go of the stage - refills it at once, so nobody ever waits on that barrier
            if ((threadIdx.x & 31) == 0 && mbar_arrive_last(empty + (dep & zero)) && i + kStages < mine) issue(i + kStages, s);
#else
            if ((threadIdx.x & 31) == 0) mbar_arrive(empty + (dep & zero));  // this warp has copied its groups out
            if (threadIdx.x == 0 && i + kStages < mine) {
                mbar_wait(empty, parity);  // all 8 warps are done with the stage
                issue(i + kStages, s);
            }
#endif
            const int64_t tile0 = (int64_t)(first_tile + (int64_t)i * tile_stride) * kTilePx;
            if constexpr (SPLIT == 1) {
                f(raw[0], tile0);
            } else {
#pragma unroll
                for (int u = 0; u < SPLIT; ++u) f(raw[u], tile0, u);
            }
        }
        parity ^= 1u;
    }
#else
    int s = 0;
    uint32_t parity = 0;
    for (int i = 0; i < mine; ++i) {
        const uint32_t full = pipe.full + 8 * s, empty = pipe.empty + 8 * s;
        mbar_wait(full, parity);
        typename IO::Raw raw[SPLIT];
#pragma unroll
        for (int u = 0; u < SPLIT; ++u) raw[u] = pipe_load_group<IO>(pipe.stage + s * kTileBytes + my_off[u]);
        uint32_t dep = 0;   // one word of every load instruction's result
#pragma unroll
        for (int u = 0; u < SPLIT; ++u) {
            const uint32_t *w = reinterpret_cast<const uint32_t *>(raw[u].e);
            if (IO::kU8) {
#pragma unroll
                for (int k = 0; k < 12; ++k) dep ^= w[k];
            } else {
#pragma unroll
                for (int k = 0; k < 3; ++k) dep ^= w[4 * k];
            }
        }
        __syncwarp();
#if CT_PIPE_LAST_REFILLS
        // this warp has copied its groups out; the warp whose arrival completes the "empty" phase - the last of
        // the 8 to let go of the stage - refills it at once, so nobody ever waits on that barrier
        if ((threadIdx.x & 31) == 0 && mbar_arrive_last(empty + (dep & zero)) && i + kStages < mine) issue(i + kStages, s);
#else
        if ((threadIdx.x & 31) == 0) mbar_arrive(empty + (dep & zero));  // this warp has copied its groups out
        if (threadIdx.x == 0 && i + kStages < mine) {
            mbar_wait(empty, parity);  // all 8 warps are done with the stage
            issue(i + kStages, s);
        }
#endif
        const int64_t tile0 = (int64_t)(first_tile + (int64_t)i * tile_stride) * kTilePx;
        if constexpr (SPLIT == 1) {
            f(raw[0], tile0);
        } else {
#pragma unroll
            for (int u = 0; u < SPLIT; ++u) f(raw[u], tile0, u);
        }
        if (++s == kStages) {
            s = 0;
            parity ^= 1u;
        }
    }
#endif
    (void)sizeof(T);
}

}  // namespace ct
