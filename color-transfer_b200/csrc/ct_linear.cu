// Linear (closed-form) transfers: K1 moments (+K1L Lab-fused), K2 one-warp 3x3 solve,
// K3 affine remap (+K3L Lab-fused).  ref: methods/linear.py:8-124.
#include "ct_context.h"
#include "ct_lab.cuh"
#include "ct_solve3.cuh"

namespace ct {

// ---------------------------------------------------------------------------------------------
// K1 / K1L: single-pass raw moments about a fixed shift, fp64 accumulators, warp-shuffle tree,
// one partial per block, last block of each pair combines all partials in block-index order
// (deterministic: no floating-point atomics) and, when asked, runs the 3x3 solve in its first
// warp.  Replaces np.mean/np.std/np.cov (linear.py:33-36, 64-67, 103-106).
// ---------------------------------------------------------------------------------------------
struct MomentsArgs {
    Img img[2];
    int kind[2];       // dtype*2 + layout
    int vec[2];
    int u8_as_f32[2];
    int nimg;          // images per pair in this launch (1 or 2)
    int lab;
    double *partials;  // [B][nimg][gridDim.x][9]
    unsigned int *tickets;
    double *sums;      // [B][nimg][10]
    int method;        // >= 0: fused solve (needs nimg == 2)
    double *xform;     // [B][16]
    int *status;       // [B]
};

// Cube-root seeds of the Lab kernels taken from the FMA pipe instead of MUFU (0..3 of the three
// per pixel): with 0 both kernels are bound by the XU pipe, with 3 by instruction issue.
#ifndef CT_LAB_PACKED   // 1: the fp32 Lab chain on pixel pairs (FFMA2), 0: the scalar chain (A/B, identical results)
#define CT_LAB_PACKED 1
#endif
#ifndef CT_STATS_FMA_SEEDS
#define CT_STATS_FMA_SEEDS 2
#endif
#ifndef CT_REINHARD_FMA_SEEDS
#define CT_REINHARD_FMA_SEEDS 1
#endif
// Minimum waves of a batched remap launch.  CTAs are scheduled pair by pair (blockIdx.y), so more,
// smaller CTAs per pair keep fewer images open at once: the write-heavy plain remap (12 B in,
// 24 B out per pixel) gains 5 % from 4 -> 12+ waves, the compute-heavy Lab remap prefers 4-8.
#ifndef CT_APPLY_MIN_WAVES
#define CT_APPLY_MIN_WAVES 12
#endif
#ifndef CT_APPLY_LAB_MIN_WAVES
#define CT_APPLY_LAB_MIN_WAVES 4
#endif
#ifndef CT_LAB_CTAS_PER_SM   // resident CTAs per SM the Lab kernels are compiled for (register cap)
#define CT_LAB_CTAS_PER_SM 4
#endif
#ifndef CT_APPLY_CTAS_PER_SM   // the plain remap: 3 measured best (2: -2 %, 4: -5 %, 5: -25 %, spills)
#define CT_APPLY_CTAS_PER_SM 3
#endif
#ifndef CT_LAB_APPLY_CTAS_PER_SM   // the Lab remap alone: fits 48 registers without spills, 5 measured best (4: -3 %, 6: -2 %)
#define CT_LAB_APPLY_CTAS_PER_SM 5
#endif

__device__ __forceinline__ void accumulate_rgb(const double (&rgb)[3], double (&acc)[9]) {
    const double v0 = rgb[0] - 0.5, v1 = rgb[1] - 0.5, v2 = rgb[2] - 0.5;
    acc[0] += v0;
    acc[1] += v1;
    acc[2] += v2;
    acc[3] = fma(v0, v0, acc[3]);
    acc[4] = fma(v0, v1, acc[4]);
    acc[5] = fma(v0, v2, acc[5]);
    acc[6] = fma(v1, v1, acc[6]);
    acc[7] = fma(v1, v2, acc[7]);
    acc[8] = fma(v2, v2, acc[8]);
}

// Lab statistics (Reinhard needs only per-channel mean and variance): the fp32 Lab chain of
// ct_lab.cuh, summed in fp32 over the N pixels a thread holds and only then folded into the
// fp64 accumulators - 6 F2F conversions (XU pipe) per group instead of per pixel.
template <int N>
__device__ __forceinline__ void accumulate_lab(const float (&rgbf)[N][3], double (&acc)[9]) {
    float s[3] = {0.0f, 0.0f, 0.0f}, q[3] = {0.0f, 0.0f, 0.0f};
    float w[N][3];  // (L - 50) / 116, a / 500, b / 200: scaled back when the sums are combined
    if constexpr (N % 2 == 0 && CT_LAB_PACKED) lab::rgb2labw_pairs<N, CT_STATS_FMA_SEEDS>(rgbf, w);
    else lab::rgb2labw_group<N, CT_STATS_FMA_SEEDS>(rgbf, w);
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            s[c] += w[i][c];
            q[c] = fmaf(w[i][c], w[i][c], q[c]);
        }
    }
    acc[0] += (double)s[0];
    acc[1] += (double)s[1];
    acc[2] += (double)s[2];
    acc[3] += (double)q[0];
    acc[6] += (double)q[1];
    acc[8] += (double)q[2];
}

template <typename IO, bool VEC, bool LAB>
__device__ __forceinline__ void moments_image(const Img &im, int64_t pair, const Decode &dec, double (&acc)[9]) {
    using T = typename IO::elem_t;
    const T *base = reinterpret_cast<const T *>(im.data) + pair * im.image_stride;
    constexpr int G = IO::G, GS = IO::GS;
    const int64_t ngroups = im.npix / G;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (LAB) {
        if (g < ngroups) {
            // compute-heavy pass: the next group's three vectors are in flight while this one is worked on
            typename IO::Raw raw = IO::template load_raw<VEC>(base, im.plane_stride, (int)g);
            for (;;) {
                g += stride;
                const bool more = g < ngroups;
                typename IO::Raw next;
                if (more) next = IO::template load_raw<VEC>(base, im.plane_stride, (int)g);
#pragma unroll
                for (int q = 0; q < IO::NSUB; ++q) {
                    float xf[GS][3];
                    IO::unpack_sub_f(raw, q, dec, xf);
                    accumulate_lab<GS>(xf, acc);
                }
                if (!more) break;
                raw = next;
            }
        }
    } else {
        for (; g < ngroups; g += stride) {
            const typename IO::Raw raw = IO::template load_raw<VEC>(base, im.plane_stride, (int)g);
#pragma unroll
            for (int q = 0; q < IO::NSUB; ++q) {
                double x[GS][3];
                IO::unpack_sub(raw, q, dec, x);
#pragma unroll
                for (int i = 0; i < GS; ++i) accumulate_rgb(x[i], acc);
            }
        }
    }
    if (blockIdx.x == 0) {
        for (int64_t p = ngroups * G + threadIdx.x; p < im.npix; p += kThreads) {
            double x[3];
            IO::load1(base, im.plane_stride, p, dec, x);
            if (LAB) {
                const float xf[1][3] = {{(float)x[0], (float)x[1], (float)x[2]}};
                accumulate_lab<1>(xf, acc);
            } else {
                accumulate_rgb(x, acc);
            }
        }
    }
}

// ANY_U8: the instantiation that also holds the uint8 variants (launched when an image is uint8), so that
// the float paths keep their own register allocation
template <bool LAB, bool ANY_U8>
__global__ void __launch_bounds__(kThreads, LAB ? CT_LAB_CTAS_PER_SM : 3) moments_kernel(MomentsArgs a) {
    const int z = blockIdx.z;
    const int64_t pair = blockIdx.y;
    double acc[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) acc[i] = 0.0;

    const Img im = a.img[z];
    const int sel = a.kind[z] * 2 + a.vec[z];
    __shared__ double dec_d[ANY_U8 ? 256 : 1];
    __shared__ float dec_f[ANY_U8 ? 256 : 1];
    if (ANY_U8 && a.kind[z] >= 4) {   // uint8 image (block-uniform)
        fill_decode(dec_d, dec_f, a.u8_as_f32[z]);
        __syncthreads();
    }
    const Decode dec{dec_d, dec_f};
    switch (sel) {
#define CT_CASE(ID, T, L, V) case ID: moments_image<PixelIO<T, L>, V, LAB>(im, pair, dec, acc); break;
        CT_FOR_EACH_FLOAT_SRC(CT_CASE)
        default:
            if constexpr (ANY_U8) {
                switch (sel) { CT_FOR_EACH_U8_SRC(CT_CASE) }
            }
#undef CT_CASE
    }

    __shared__ double red[kWarps][9];
    __shared__ double total[2][10];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const double s = warp_sum(acc[i]);
        if (lane == 0) red[warp][i] = s;
    }
    __syncthreads();
    const int64_t nblk = gridDim.x;
    double *mine = a.partials + (((int64_t)pair * a.nimg + z) * nblk + blockIdx.x) * 9;
    if (threadIdx.x < 9) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += red[w][threadIdx.x];
        mine[threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&a.tickets[pair], 1u);
        is_last = (t == (unsigned int)(nblk * a.nimg) - 1u);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();

    // fixed-order combine, one warp per quantity (img, k): lane l adds the partials of blocks
    // l, l+32, ... in that order, then a fixed shuffle tree - deterministic for a given grid, and
    // ~30x shorter than one thread walking all blocks (this tail is most of a small launch).
    for (int q = warp; q < 9 * a.nimg; q += kWarps) {
        const int img = q / 9, k = q % 9;
        const double *p = a.partials + ((int64_t)pair * a.nimg + img) * nblk * 9 + k;
        double s = 0.0;
        for (int64_t b = lane; b < nblk; b += 32) s += __ldcg(p + b * 9);
        s = warp_sum(s);
        if (lane == 0) {
            if (LAB) {  // the Lab pass sums w = ((L - 50) / 116, a / 500, b / 200) and their squares
                const double f = (k == 0 || k == 3) ? 116.0 : (k == 1 || k == 6) ? 500.0 : 200.0;
                s *= k < 3 ? f : f * f;
            }
            total[img][1 + k] = s;
            if (k == 0) total[img][0] = (double)a.img[img].npix;
        }
    }
    __syncthreads();
    if (threadIdx.x < 10 * a.nimg) {
        const int img = threadIdx.x / 10, k = threadIdx.x % 10;
        a.sums[((int64_t)pair * a.nimg + img) * 10 + k] = total[img][k];
    }
    if (threadIdx.x == 0) {
        a.tickets[pair] = 0;  // ready for the next launch on this stream
        if (a.method >= 0) {
            const int st = solve3::solve(a.method, total[0], total[1], a.xform + pair * CT_XFORM_DOUBLES);
            if (a.status) a.status[pair] = st;
        }
    }
}

// K2 standalone: one warp, lane = pair (used after an all-reduce of the sums, and by tests).
__global__ void solve_kernel(int method, const double *sums_t, const double *sums_r,
                             int64_t sums_stride, int count, double *xform, int *status) {
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= count) return;
    const int st = solve3::solve(method, sums_t + pair * sums_stride, sums_r + pair * sums_stride,
                                 xform + (int64_t)pair * CT_XFORM_DOUBLES);
    if (status) status[pair] = st;
}

// ---------------------------------------------------------------------------------------------
// K3 / K3L: out = (x - mu_t) @ M + mu_r per pixel; Reinhard wraps it in rgb2lab / lab2rgb+clip.
// Pure streaming: 128-bit loads and stores, grid-stride, 2 x S bytes per pixel.
// ---------------------------------------------------------------------------------------------
struct ApplyArgs {
    Img src;
    ImgOut dst;
    const double *xform;  // [B][16]
    int u8_as_f32;        // uint8 source: decode k/255 in float32 (else float64)
    int clamp;            // float output: clamp to [0,1]
};

template <bool LAB>
__device__ __forceinline__ void apply_pixel(const double *xf, const float *xff, const double (&x)[3],
                                            const float (&xs)[3], double (&y)[3]) {
    if (LAB) {  // xs: the pixel as floats (seeds of the gamma decode)
        double l[3], m[3];
        lab::rgb2lab(x, xs, l);
        // (lab - mean_t) * std_r / std_t + mean_r   (linear.py:38)
        m[0] = fma(l[0] - xf[9], xf[0], xf[12]);
        m[1] = fma(l[1] - xf[10], xf[4], xf[13]);
        m[2] = fma(l[2] - xf[11], xf[8], xf[14]);
        lab::lab2rgb(m, y);
    } else {
        const double d0 = x[0] - xf[9], d1 = x[1] - xf[10], d2 = x[2] - xf[11];
#pragma unroll
        for (int c = 0; c < 3; ++c) y[c] = fma(d2, xf[6 + c], fma(d1, xf[3 + c], d0 * xf[c])) + xf[12 + c];
    }
}

template <typename A, typename B> struct same_t { static constexpr bool value = false; };
template <typename A> struct same_t<A, A> { static constexpr bool value = true; };

// HYB32: the whole Lab chain in fp32 (float32 image in and out, or a uint8 frame decoded as float32:
// Reinhard keeps the input dtype, linear.py:25-40)
template <typename SIO, typename DIO, bool VEC, bool LAB, bool HYB32>
__global__ void __launch_bounds__(kThreads, LAB ? CT_LAB_APPLY_CTAS_PER_SM : CT_APPLY_CTAS_PER_SM) apply_kernel(ApplyArgs a) {
    using TS = typename SIO::elem_t;
    using TD = typename DIO::elem_t;
    constexpr bool HYB = LAB && HYB32;
    const int64_t pair = blockIdx.y;
    __shared__ double xf[CT_XFORM_DOUBLES];
    __shared__ float xff[CT_XFORM_DOUBLES];
    __shared__ lab::ReinhardFold fold;
    __shared__ double dec_d[SIO::kU8 ? 256 : 1];
    __shared__ float dec_f[SIO::kU8 ? 256 : 1];
    if (threadIdx.x < CT_XFORM_DOUBLES) {
        xf[threadIdx.x] = a.xform[pair * CT_XFORM_DOUBLES + threadIdx.x];
        xff[threadIdx.x] = (float)xf[threadIdx.x];
    }
    if (HYB && threadIdx.x == 32) fold = lab::fold_reinhard(a.xform + pair * CT_XFORM_DOUBLES);
    if (SIO::kU8) fill_decode(dec_d, dec_f, a.u8_as_f32);
    __syncthreads();
    const Decode dec{dec_d, dec_f};
    const bool clamp = a.clamp != 0;
    const TS *src = reinterpret_cast<const TS *>(a.src.data) + pair * a.src.image_stride;
    TD *dst = reinterpret_cast<TD *>(a.dst.data) + pair * a.dst.image_stride;
    constexpr int G = SIO::G, GS = SIO::GS;
    const int64_t ngroups = a.src.npix / G;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    // 256-bit stores of a 32-byte aligned destination (ct_common.cuh: store)
    const bool wide = VEC && (reinterpret_cast<uintptr_t>(dst) & 31) == 0 && (DIO::kLayout == CT_HWC || (a.dst.plane_stride & 3) == 0);
    for (int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x; g < ngroups; g += stride) {
        const typename SIO::Raw raw = SIO::template load_raw<VEC>(src, a.src.plane_stride, (int)g);
#pragma unroll
        for (int q = 0; q < SIO::NSUB; ++q) {
            if (HYB) {
                float xs[GS][3], ys[GS][3];
                SIO::unpack_sub_f(raw, q, dec, xs);
                if constexpr (GS % 2 == 0 && CT_LAB_PACKED) lab::reinhard_pairs_h<GS, CT_REINHARD_FMA_SEEDS>(fold, xs, ys);
                else lab::reinhard_group_h<GS, CT_REINHARD_FMA_SEEDS>(fold, xs, ys);
                DIO::template store<VEC, GS>(dst, a.dst.plane_stride, SIO::sub_pixel0(g, q), ys, clamp, wide);
            } else {
                double x[GS][3], y[GS][3];
                float xs[GS][3];
                SIO::unpack_sub(raw, q, dec, x);
                if (LAB) SIO::unpack_sub_f(raw, q, dec, xs);
#pragma unroll
                for (int i = 0; i < GS; ++i) apply_pixel<LAB>(xf, xff, x[i], xs[i], y[i]);
                DIO::template store<VEC, GS>(dst, a.dst.plane_stride, SIO::sub_pixel0(g, q), y, clamp, wide);
            }
        }
    }
    if (blockIdx.x == 0) {
        for (int64_t p = ngroups * G + threadIdx.x; p < a.src.npix; p += kThreads) {
            double x[3], y[3];
            SIO::load1(src, a.src.plane_stride, p, dec, x);
            const float xs[3] = {(float)x[0], (float)x[1], (float)x[2]};
            if (HYB) {
                const float x1[1][3] = {{xs[0], xs[1], xs[2]}};
                float y1[1][3];
                lab::reinhard_group_h<1, CT_REINHARD_FMA_SEEDS>(fold, x1, y1);
                DIO::store1(dst, a.dst.plane_stride, p, y1[0], clamp);
            } else {
                apply_pixel<LAB>(xf, xff, x, xs, y);
                DIO::store1(dst, a.dst.plane_stride, p, y, clamp);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
// CTAs per image for a launch of `units` images that keeps `ctas_per_sm` CTAs resident per SM
// (the kernel's __launch_bounds__): every CTA does the same work, so the grid is sized to fill
// WHOLE waves - a 5.3-wave grid runs as long as a 6-wave one.  Picks the smallest wave count whose
// last wave is at least 97 % full (or the fullest of eight candidates), starting at `min_waves`:
// the remap kernels run best with >= 4 waves, the moments pass with few (short combine).
static int blocks_for(const ct_context *h, int64_t npix, int group, int64_t units, int ctas_per_sm, int min_waves) {
    int64_t want = (npix / group + kThreads - 1) / kThreads;
    if (want < 1) want = 1;
    if (units < 1) units = 1;
    const int64_t resident = (int64_t)h->sm_count * ctas_per_sm;
    int64_t best = 1;
    double best_fill = 0.0;
    for (int waves = min_waves; waves <= min_waves + 7; ++waves) {
        int64_t n = resident * waves / units;
        if (n < 1) continue;
        if (n >= want) return (int)want;   // small images: one CTA per kThreads groups
        const double fill = (double)(n * units) / (double)(resident * waves);
        if (fill > best_fill) {
            best_fill = fill;
            best = n;
        }
        if (fill >= 0.97) break;
    }
    return (int)best;
}

int launch_moments(ct_context *h, const ct_batch *a, const ct_batch *b, int lab, double *sums,
                   int method, double *xform, int *status) {
    CT_TRY(check_batch(h, a, "images"));
    const int nimg = b ? 2 : 1;
    if (b) {
        CT_TRY(check_batch(h, b, "reference"));
        if (b->count != a->count) return fail(h, CT_E_INVALID, "target/reference batch counts differ");
    }
    if (!sums) return fail(h, CT_E_INVALID, "sums is NULL");
    const int B = a->count;
    int64_t npix_max = a->npix;
    if (b && b->npix > npix_max) npix_max = b->npix;
    // a single pair gets exactly one resident wave (short fixed-order combine); batches whole waves
    const int nblk = blocks_for(h, npix_max, 2, (int64_t)B * nimg, lab ? CT_LAB_CTAS_PER_SM : 3, 1);
    // (a chunked batch pre-sizes both buffers: growing them here would free memory the other side stream is using)
    if (h->partials_region && (size_t)B * nimg * nblk * 9 > h->partials_region)
        return fail(h, CT_E_NOMEM, "moments partials region too small for %d pairs x %d blocks", B, nblk);
    CT_TRY(ensure_partials(h, h->partials_base + (size_t)B * nimg * nblk * 9));
    CT_TRY(ensure_scratch(h, h->ticket_base + B));
    MomentsArgs m{};
    m.img[0] = img_of(a);
    m.kind[0] = src_kind(a);
    m.vec[0] = vec_ok(a);
    m.u8_as_f32[0] = (a->flags & CT_BATCH_U8_AS_F32) != 0;
    if (b) {
        m.img[1] = img_of(b);
        m.kind[1] = src_kind(b);
        m.vec[1] = vec_ok(b);
        m.u8_as_f32[1] = (b->flags & CT_BATCH_U8_AS_F32) != 0;
    }
    m.nimg = nimg;
    m.lab = lab;
    m.partials = h->partials + h->partials_base;
    m.tickets = h->tickets + h->ticket_base;
    m.sums = sums;
    m.method = (b && xform) ? method : -1;
    m.xform = xform;
    m.status = status;
    const bool any_u8 = a->dtype == CT_U8 || (b && b->dtype == CT_U8);
    const dim3 grid(nblk, B, nimg);
    if (lab) {
        if (any_u8) moments_kernel<true, true><<<grid, kThreads, 0, h->stream>>>(m);
        else moments_kernel<true, false><<<grid, kThreads, 0, h->stream>>>(m);
    } else {
        if (any_u8) moments_kernel<false, true><<<grid, kThreads, 0, h->stream>>>(m);
        else moments_kernel<false, false><<<grid, kThreads, 0, h->stream>>>(m);
    }
    h->launches++;
    CT_CUDA(h, cudaGetLastError());
    return CT_OK;
}

int launch_solve(ct_context *h, int method, const double *sums_t, const double *sums_r,
                 int64_t sums_stride, int count, double *xform, int *status) {
    if (!sums_t || !sums_r || !xform || count <= 0) return fail(h, CT_E_INVALID, "bad solve arguments");
    if (method < CT_REINHARD || method > CT_MKL_CHOLESKY) return fail(h, CT_E_INVALID, "unknown method %d", method);
    solve_kernel<<<(count + 31) / 32, 32, 0, h->stream>>>(method, sums_t, sums_r, sums_stride, count, xform, status);
    h->launches++;
    CT_CUDA(h, cudaGetLastError());
    return CT_OK;
}

// HYB32 (all-fp32 Lab chain) when the image the reference would see is float32: a float32 source
// with a float32 or uint8 destination, or a uint8 source decoded as float32.
template <typename SIO, typename DIO, bool LAB>
static void launch_apply_one(bool vec, bool hyb32, dim3 grid, cudaStream_t st, const ApplyArgs &a) {
    if (LAB && hyb32) {
        if (vec) apply_kernel<SIO, DIO, true, LAB, true><<<grid, kThreads, 0, st>>>(a);
        else apply_kernel<SIO, DIO, false, LAB, true><<<grid, kThreads, 0, st>>>(a);
    } else {
        if (vec) apply_kernel<SIO, DIO, true, LAB, false><<<grid, kThreads, 0, st>>>(a);
        else apply_kernel<SIO, DIO, false, LAB, false><<<grid, kThreads, 0, st>>>(a);
    }
}

template <typename SIO, bool LAB>
static int launch_apply_dst(ct_context *h, const ct_batch *out, bool vec, bool hyb32, dim3 grid, const ApplyArgs &a) {
    cudaStream_t st = h->stream;
    if (out->dtype == CT_F32 && out->layout == CT_HWC) launch_apply_one<SIO, PixelIO<float, CT_HWC>, LAB>(vec, hyb32, grid, st, a);
    else if (out->dtype == CT_F64 && out->layout == CT_HWC) launch_apply_one<SIO, PixelIO<double, CT_HWC>, LAB>(vec, false, grid, st, a);
    else if (out->dtype == CT_U8 && out->layout == CT_HWC) launch_apply_one<SIO, PixelIO<uint8_t, CT_HWC>, LAB>(vec, hyb32, grid, st, a);
    else return fail(h, CT_E_UNSUPPORTED, "linear output must be float32 / float64 / uint8 CT_HWC");
    return CT_OK;
}

int launch_apply(ct_context *h, int method, const ct_batch *target, const double *xform,
                 const ct_batch *out) {
    CT_TRY(check_batch(h, target, "target"));
    CT_TRY(check_batch(h, out, "out"));
    if (!xform) return fail(h, CT_E_INVALID, "xform is NULL");
    if (out->npix != target->npix || out->count != target->count)
        return fail(h, CT_E_INVALID, "out must have the target's npix and count");
    const bool vec = vec_ok(target) && vec_ok(out);
    const int group = target->dtype == CT_F32 ? 4 : (target->dtype == CT_U8 ? 16 : 2);
    const int nblk = blocks_for(h, target->npix, group, target->count, method == CT_REINHARD ? CT_LAB_APPLY_CTAS_PER_SM : CT_APPLY_CTAS_PER_SM,
                                target->count == 1 ? 1 : (method == CT_REINHARD ? CT_APPLY_LAB_MIN_WAVES : CT_APPLY_MIN_WAVES));
    const dim3 grid(nblk, target->count);
    const bool u8f32 = (target->flags & CT_BATCH_U8_AS_F32) != 0;
    ApplyArgs a{img_of(target), imgout_of(out), xform, u8f32 ? 1 : 0, (out->flags & CT_BATCH_CLAMP01) ? 1 : 0};
    const bool labm = method == CT_REINHARD;
    // the dtype the reference would compute Reinhard in (linear.py:25-40 keeps the input float dtype)
    const bool hyb32 = (target->dtype == CT_F32 || (target->dtype == CT_U8 && u8f32)) && out->dtype != CT_F64;
    int rc = CT_OK;
    switch (src_kind(target)) {
#define CT_APPLY(T, L)                                                                        \
    rc = labm ? launch_apply_dst<PixelIO<T, L>, true>(h, out, vec, hyb32, grid, a)            \
              : launch_apply_dst<PixelIO<T, L>, false>(h, out, vec, hyb32, grid, a);          \
    break;
        case 0: CT_APPLY(float, CT_HWC)
        case 1: CT_APPLY(float, CT_CHW)
        case 2: CT_APPLY(double, CT_HWC)
        case 3: CT_APPLY(double, CT_CHW)
        case 4: CT_APPLY(uint8_t, CT_HWC)
        case 5: CT_APPLY(uint8_t, CT_CHW)
#undef CT_APPLY
    }
    CT_TRY(rc);
    h->launches++;
    CT_CUDA(h, cudaGetLastError());
    return CT_OK;
}

}  // namespace ct
