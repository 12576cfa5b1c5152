// Iterative distribution transfer (Pitie 2007): K4 projected ranges, K5 projection + privatised
// histograms, K6 CDF / inverse-CDF table, K7 remap + back-rotation + next range.
// ref: methods/iterative.py:31-55.  Numpy semantics restated in oracle/reference_numpy.py.
//
// Bit-exactness contract: projections use the FMA chain of a K=3 dgemm micro-kernel (checked
// against numpy's `r @ X.T`), the bin grid is np.linspace's `i*step + lo` with separate IEEE
// multiply and add, a sample's bin is the unique k with edges[k] <= x < edges[k+1] (last bin
// closed) exactly as np.histogram's corrected estimate, counts are integers, CDFs and the
// table use un-contracted IEEE ops in np.interp's order.
#include "ct_context.h"
#include "ct_pipe.cuh"

#ifndef CT_MINB
#define CT_MINB 2
#endif
#ifndef CT_RANGES_CHUNK
#define CT_RANGES_CHUNK 4   // rotations evaluated per pass over a static image (1, 2 or 4)
#endif
#ifndef CT_HIST_COPIES_LOG2
#define CT_HIST_COPIES_LOG2 3   // lane-interleaved copies of the three histograms for bins <= 256
#endif
#ifndef CT_STREAM_ONLY
#define CT_STREAM_ONLY 0   // 1: K5 / K7 only move their tiles (tools/README.md: how the streaming ceilings in DESIGN.md were measured)
#endif
#ifndef CT_HIST_STAGES
#define CT_HIST_STAGES 3
#endif
#ifndef CT_REMAP_STAGES
#define CT_REMAP_STAGES 3   // 3 / 4 / 5 / 6 measured with the last-releaser refill: 2.09 / 2.16 / 2.18 / 2.21 ms per 8-frame pass
#endif
namespace ct {

#ifndef CT_RANGES_STAGES1
#define CT_RANGES_STAGES1 3   // tile stages of the one-rotation pass (4 CTAs per SM)
#endif
#ifndef CT_RANGES_STAGES4
#define CT_RANGES_STAGES4 4   // ... of the two- and four-rotation passes (3 CTAs per SM)
#endif
constexpr int ranges_stages(int nrot) { return nrot == 1 ? CT_RANGES_STAGES1 : CT_RANGES_STAGES4; }
constexpr int kHistStages = CT_HIST_STAGES, kRemapStages = CT_REMAP_STAGES;
using HistPipe = Pipe<kHistStages>;
// hist kernel dynamic shared memory: stages + two barrier sets (one per image), then the histograms
constexpr int kHistSmemFront = kHistStages * kTileBytes + 2 * (2 * kHistStages * 8);
using RemapPipe = Pipe<kRemapStages>;

// ---------------------------------------------------------------------------------------------
// the shared uniform grid of one axis (np.histogram(bins, range=[lo, hi]))
// ---------------------------------------------------------------------------------------------
struct AxisGrid {
    double lo, hi, step, inv;
};

__device__ __forceinline__ AxisGrid grid_from_keys(const int64_t *keys, int j, int bins, bool &finite) {
    AxisGrid g;
    g.lo = value_of(__ldcg(keys + j));
    g.hi = -value_of(__ldcg(keys + 3 + j));
    finite = isfinite(g.lo) && isfinite(g.hi);
    if (g.lo == g.hi) {  // numpy/lib/_histograms_impl.py:321-324
        g.lo -= 0.5;
        g.hi += 0.5;
    }
    g.step = div_rn(sub_rn(g.hi, g.lo), (double)bins);  // np.linspace: delta / div
    g.inv = (double)bins / (g.hi - g.lo);
    return g;
}

// np.linspace(lo, hi, bins + 1)[k]
__device__ __forceinline__ double edge(const AxisGrid &g, int k, int bins) {
    return k >= bins ? g.hi : add_rn(mul_rn((double)k, g.step), g.lo);
}

// Adding 2^52 + 2^51 leaves round-to-nearest(t) in the low mantissa word: no F2I / I2F.
constexpr double kMagic = 6755399441055744.0;

// Nearest-integer estimate k' of (p - lo) / step, clamped to [0, bins] so that edges[k'] is a
// valid table read.  The sample's bin is k' - (p < edges[k']), see bin_from.
__device__ __forceinline__ int bin_estimate(double p, double lo, double inv, int bins) {
    const double u = fma(p - lo, inv, kMagic);
    return (int)min((unsigned)__double2loint(u), (unsigned)bins);
}
// Unique k in [0, bins-1] with edges[k] <= p < edges[k+1] (last bin closed) from the estimate
// and the exact table value e = edges[k'].  The estimate is within 1/2 of the true quotient, so
// the answer is k' or k' - 1 and one comparison against the exact edge decides - the same
// invariant np.histogram enforces with its -1/+1 correction.
__device__ __forceinline__ int bin_from(int kest, double p, double e, int bins) {
    const int k = kest - (p < e ? 1 : 0);
    return (int)min((unsigned)k, (unsigned)(bins - 1));
}
// The same decision with edges[k'] recomputed as np.linspace does (k'*step + lo, separate IEEE
// multiply and add; k' as a double falls out of the magic-number trick for free).  k' = bins is
// not special-cased to `hi`: either outcome of the comparison folds into the last bin.
__device__ __forceinline__ int bin_exact(double p, double lo, double inv, double step, int bins) {
    const double u = fma(p - lo, inv, kMagic);
    const double e = add_rn(mul_rn(u - kMagic, step), lo);
    const int k = __double2loint(u) - (p < e ? 1 : 0);
    return (int)min((unsigned)k, (unsigned)(bins - 1));
}

__device__ __forceinline__ bool not_finite(double p) {
    return (__double2hiint(p) & 0x7ff00000) == 0x7ff00000;
}

// fold per-thread (min p0, min p1, min p2, max p0, max p1, max p2) into the pair's keys, which hold
// the minima of (p, -p)
__device__ __forceinline__ void fold_range(double (&mn)[6], int64_t *keys, double (*red)[6]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const double v = warp_min(i < 3 ? mn[i] : -mn[i]);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = red[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) v = fmin(v, red[w][threadIdx.x]);
        if (v < INFINITY) atomicMin(reinterpret_cast<long long *>(keys) + threadIdx.x, (long long)key_of(v));
    }
}

// Running (min, max) of p: mn[0..2] minima, mn[3..5] MAXIMA (tracking -p instead would cost a DADD
// per sample on the fp64 pipe).  Plain compare-select: fmin()'s NaN handling costs twice as many
// instructions, and non-finite samples are caught once, by K4 on the inputs (CHECK).
template <bool CHECK>
__device__ __forceinline__ void track_range(const double *rot, const double (&x)[3], double (&mn)[6], bool &bad) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double p = dot3(rot + 3 * j, x);
        if (CHECK) bad |= not_finite(p);
        mn[j] = p < mn[j] ? p : mn[j];
        mn[3 + j] = p > mn[3 + j] ? p : mn[3 + j];
    }
}

// ---------------------------------------------------------------------------------------------
// keys init
// ---------------------------------------------------------------------------------------------
__global__ void keys_init_kernel(int64_t *keys, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = kKeyPlusInf;
}

// ---------------------------------------------------------------------------------------------
// K4: projected ranges of an image under up to kMaxRot rotations in one pass (iterative.py:34-35,
// 39-40).
//
// The result must be the exact fp64 minimum / maximum of p = rot_k[j] . x (it defines the bin grid),
// but evaluating 3 DFMA-chains and 6 fp64 compare-selects per pixel and rotation made the pass
// instruction-bound (round 1: 36 instructions per pixel and rotation, 0.27 of the HBM roofline).
// Almost no pixel can move an extreme, so the work is split:
//   K4a `ranges_seed_kernel`: exact fp64 ranges of a 1/64 subsample spread over the whole image
//        (runs of 8 pixels), one partial result per seed CTA - a few microseconds.
//   K4b `ranges_kernel`: every pixel is SCREENED in packed fp32 (FFMA2: two directions per
//        instruction).  With c, h the centre and half width of an interval safely inside the seed's
//        exact extremes, the rotation rows are pre-scaled by 1/h, so a pixel whose fp32 values
//        d = (r/h) . x - c/h all satisfy |d| < 1 cannot change any extreme and is done after 3 FFMA2
//        per direction pair and one 3-input maximum of the absolute values.  Only flagged pixels
//        (beyond the sample's range, or tying an extreme) take the exact fp64 path, which folds them
//        into the CTA's exact extremes (monotone int64 keys in shared memory, seeded with the sample's).
// The answer is exactly the fp64 minimum / maximum whatever the screen does: the seed's extremes are
// projections of real pixels, and the screen can only send too many pixels to the exact path, never
// too few (error bound below).
//
// Error bound of the screen.  d = fma(r2,x2, fma(r1,x1, fma(r0,x0,c'))) in fp32 with r = fl32(rot/h),
// c' = fl32(-c/h) and x = fl32(pixel): |d - (p - c)/h| <= 2^-24 (2 S + |c| + 3 (S + |c|)) / h (1 + o(1))
// with S = |x0|+|x1|+|x2| (|rot| <= 1).  With S <= bound this is below delta / h,
// delta = 2^-21 (bound + |c|); pixels with S > bound (or NaN) are always flagged.  h is the half width
// of [lo + delta, hi - delta] about c, so |d| < 1 implies lo <= p <= hi.
// ---------------------------------------------------------------------------------------------
#ifndef CT_SEED_FRACTION
#define CT_SEED_FRACTION 128   // K4a samples one pixel in this many
#endif
constexpr int kMaxRot = 4;     // most rotations one pass can take
constexpr int kSeedRun = 8;    // consecutive pixels per sampled run

struct RangesArgs {
    Img img[2];      // up to two images per pair in one launch (target, then reference)
    int kind[2], vec[2];
    int n_rot[2];    // 0: image absent; else rotations rot + 9k -> key slots keys + 6k, k < n_rot
    int u8_as_f32[2];  // uint8 images: decode k/255 in float32 (else float64)
    const double *rot;
    int64_t rot_stride;
    int64_t *keys;
    int64_t keys_stride;
    int32_t *status;
    float bound;     // screen validity bound on |x0|+|x1|+|x2| (pixels beyond it take the exact path)
    long long *seed; // [B][2][6 * kMaxRot] keys of the subsample's extremes (K4a -> K4b), kSeedEmpty where none
    int64_t *init_keys;   // K4a also sets these n_init keys to "+inf" (what keys_init_kernel does)
    int64_t n_init;
    unsigned long long *stats;  // optional diagnostics: [0] pixels sent to the exact path, [1] flagged repeats skipped
    int which;       // K4b: the image slot this launch streams
    uint32_t zero;   // always 0 (ct_pipe.cuh: pipe_for_each_group)
};

__device__ __forceinline__ int64_t seed_samples(int64_t npix) {
    const int64_t want = npix / CT_SEED_FRACTION > 4096 ? npix / CT_SEED_FRACTION : 4096;
    return want < npix ? want : npix;
}

// K4a.  grid (seed CTAs, B, 2): exact ranges of the subsample of image z of pair y, folded into
// seed[pair][z][6 * kMaxRot] with integer atomic minima (the host presets the array to kSeedEmpty);
// the grid also initialises the pair's range keys.  Every thread loads four samples once (all in
// flight together) and evaluates every rotation on them.
constexpr long long kSeedEmpty = 0x7f7f7f7f7f7f7f7fLL;   // what cudaMemsetAsync(0x7f) leaves: the key of 1.38e306; a sample
                                                         // that really had it would only make the pass run unseeded

__global__ void __launch_bounds__(kThreads) ranges_seed_kernel(RangesArgs a) {
    __shared__ double rot[9 * kMaxRot];
    __shared__ long long best[6 * kMaxRot];
    __shared__ double dec_d[256];
    __shared__ float dec_f[256];
    const int z = blockIdx.z;
    const int64_t pair = blockIdx.y;
    if (a.init_keys) {
        const int64_t ncta = (int64_t)gridDim.x * gridDim.y * gridDim.z;
        const int64_t cta = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        for (int64_t i = cta * kThreads + threadIdx.x; i < a.n_init; i += ncta * kThreads) a.init_keys[i] = kKeyPlusInf;
    }
    const int n_rot = a.n_rot[z];
    if (!n_rot) return;
    if (threadIdx.x < 9 * n_rot) rot[threadIdx.x] = a.rot[pair * a.rot_stride + threadIdx.x];
    if (threadIdx.x < 6 * kMaxRot) best[threadIdx.x] = kSeedEmpty;
    if (a.kind[z] >= 4) fill_decode(dec_d, dec_f, a.u8_as_f32[z]);
    __syncthreads();
    const Decode dec{dec_d, dec_f};
    const Img &im = a.img[z];
    const int64_t nsamp = seed_samples(im.npix), step = im.npix / nsamp;
    const int lane = threadIdx.x & 31;
    // runs of kSeedRun pixels, one per cell of kSeedRun * step pixels, at a hashed offset inside the
    // cell (a fixed offset would alias with the row length and sample only a few image columns)
    const int64_t cell = kSeedRun * step, slack = cell - kSeedRun + 1;
    // block-uniform trip count: the warp shuffles below need every lane
    for (int64_t b0 = (int64_t)blockIdx.x * (4 * kThreads); b0 < nsamp; b0 += (int64_t)gridDim.x * (4 * kThreads)) {
        const int64_t i0 = b0 + threadIdx.x;
        double x[4][3];
        bool have[4];
        auto load = [&](auto io) {
            using IO = decltype(io);
            using T = typename IO::elem_t;
            const T *base = reinterpret_cast<const T *>(im.data) + pair * im.image_stride;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t i = i0 + u * kThreads;
                const int64_t run = i / kSeedRun;
                const int64_t jitter = step == 1 ? 0 : (int64_t)(((uint32_t)run * 2654435761u) >> 8) % slack;
                const int64_t p = run * cell + jitter + (i % kSeedRun);
                have[u] = i < nsamp && p < im.npix;
                if (have[u]) IO::load1(base, im.plane_stride, p, dec, x[u]);
            }
        };
        switch (a.kind[z]) {
            case 0: load(PixelIO<float, CT_HWC>{}); break;
            case 1: load(PixelIO<float, CT_CHW>{}); break;
            case 2: load(PixelIO<double, CT_HWC>{}); break;
            case 3: load(PixelIO<double, CT_CHW>{}); break;
            case 4: load(PixelIO<uint8_t, CT_HWC>{}); break;
            default: load(PixelIO<uint8_t, CT_CHW>{}); break;
        }
        for (int k = 0; k < n_rot; ++k) {
            double mn[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) mn[i] = i < 3 ? INFINITY : -INFINITY;
            bool bad = false;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (have[u]) track_range<false>(rot + 9 * k, x[u], mn, bad);
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const double v = warp_min(i < 3 ? mn[i] : -mn[i]);   // min p, min -p
                // a NaN sample is never smaller; an infinite one leaves +-inf: the main pass then flags every pixel
                if (lane == 0 && v < INFINITY) atomicMin(&best[6 * k + i], (long long)key_of(v));
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 6 * n_rot && best[threadIdx.x] != kSeedEmpty)
        atomicMin(a.seed + (pair * 2 + z) * (6 * kMaxRot) + threadIdx.x, best[threadIdx.x]);
}

struct RangesShared {
    double rot[9 * kMaxRot];
    long long key[kMaxRot][6];        // exact extremes so far: keys of min p [3], min -p [3], seeded by K4a
    float scr[3 * kMaxRot + 1][4];    // per direction d = 3k + j: rot row / h [3], -c / h   ((0,0,0,2) = "flag everything")
    int bad;
    unsigned int n_exact, n_repeat;
};

template <int NP>
struct Screen {
    float2 r0[NP], r1[NP], r2[NP];  // rotation rows / h of directions (2i, 2i+1), fp32
    float2 c[NP];                   // -c / h of the two directions
};

// screen row of one direction from its exact extremes so far: (rot row / h, -c / h)
__device__ __forceinline__ void screen_row(const double *rot3, double lo, double hi, float bound, float (&out)[4]) {
    out[0] = out[1] = out[2] = 0.0f;
    out[3] = 2.0f;   // |d| = 2: flag every pixel
    if (!(lo < hi)) return;
    const float c32 = (float)(0.5 * (lo + hi));
    const double c = (double)c32;
    const double delta = ((double)bound + fabs(c)) * 4.76837158203125e-07 + 7.888609052210118e-31;  // 2^-21, 2^-100
    const double h = fmin((hi - delta) - c, c - (lo + delta));
    if (!(h > 1e-30) || !(h < 1e18) || !(fabs(c) < 1e18)) return;
    const double inv = 1.0 / h;
    out[0] = (float)(rot3[0] * inv);
    out[1] = (float)(rot3[1] * inv);
    out[2] = (float)(rot3[2] * inv);
    out[3] = (float)(-c * inv);
    if (!(fabsf(out[0]) < 1e30f && fabsf(out[1]) < 1e30f && fabsf(out[2]) < 1e30f && fabsf(out[3]) < 1e30f)) {
        out[0] = out[1] = out[2] = 0.0f;
        out[3] = 2.0f;
    }
}

// true: the pixel may move an extreme (or cannot be screened) and must take the exact path
template <int NP>
__device__ __forceinline__ bool screen_flag(const Screen<NP> &s, const float (&x)[3], float bound) {
    const float2 x0 = make_float2(x[0], x[0]), x1 = make_float2(x[1], x[1]), x2 = make_float2(x[2], x[2]);
    float m = 0.0f;   // max |d| over the directions (a NaN d is dropped here; only a non-finite pixel makes one)
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        float2 d = __ffma2_rn(s.r0[i], x0, s.c[i]);
        d = __ffma2_rn(s.r1[i], x1, d);
        d = __ffma2_rn(s.r2[i], x2, d);
        m = fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), m);
    }
    const float sum = fabsf(x[0]) + fabsf(x[1]) + fabsf(x[2]);
    return !(sum <= bound) || !(m < 1.0f);
}

// exact path of one pixel: fold p and -p of rotations [k0, k1) into the CTA's keys
__device__ __noinline__ void ranges_exact_pixel(RangesShared &sh, double x0, double x1, double x2, int k0, int k1) {
    const double x[3] = {x0, x1, x2};
    for (int k = k0; k < k1; ++k) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double p = dot3(sh.rot + 9 * k + 3 * j, x);
            if (not_finite(p)) sh.bad = 1;
            const long long klo = (long long)key_of(p), khi = (long long)key_of(-p);
            if (klo < *(volatile long long *)&sh.key[k][j]) atomicMin(&sh.key[k][j], klo);
            if (khi < *(volatile long long *)&sh.key[k][3 + j]) atomicMin(&sh.key[k][3 + j], khi);
        }
    }
}

// A pixel that ties an extreme (saturated white, letterbox black) stays inside the screen's error
// band for ever: each thread remembers the last two distinct pixels it folded exactly and skips
// their repeats.
struct KnownPixels {
    double a[3] = {NAN, NAN, NAN}, b[3] = {NAN, NAN, NAN};
    __device__ __forceinline__ bool seen(const double (&x)[3]) const {
        return (x[0] == a[0] && x[1] == a[1] && x[2] == a[2]) || (x[0] == b[0] && x[1] == b[1] && x[2] == b[2]);
    }
    __device__ __forceinline__ void push(const double (&x)[3]) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            b[c] = a[c];
            a[c] = x[c];
        }
    }
};

// NROT = rotations one thread screens.  SPLIT = 2: the CTA's two halves (warps 0-3 / 4-7) take rotations
// [0, NROT) / [NROT, 2 NROT) of the pass and every thread visits two groups per tile - half the screen
// state per thread (24 instead of 48 registers), which is what lets three CTAs share an SM.
template <typename IO, bool VEC, int NROT, int SPLIT, typename RangesPipe>
__device__ __forceinline__ void ranges_image(const Img &im, int64_t pair, RangesShared &sh, const Decode &dec, int n_rot,
                                             float bound, RangesPipe &pipe, int first_block, int nblocks, bool stats, uint32_t zero) {
    using T = typename IO::elem_t;
    constexpr int NP = (3 * NROT + 1) / 2;
    const T *base = reinterpret_cast<const T *>(im.data) + pair * im.image_stride;
    constexpr int G = IO::G, GS = IO::GS;
    const int half = SPLIT == 1 ? 0 : threadIdx.x / (kThreads / SPLIT);
    const int d0 = 3 * NROT * half;   // first direction of this thread
    Screen<NP> s;   // fixed for the whole pass; unused directions (k >= n_rot) are all-zero rows: d = 0, always inside
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int da = d0 + 2 * i, db = d0 + 2 * i + 1;  // directions d = 3k + j <-> rot[3d .. 3d+2]
        const bool ha = 2 * i < 3 * NROT && da < 3 * n_rot, hb = 2 * i + 1 < 3 * NROT && db < 3 * n_rot;
        s.r0[i] = make_float2(ha ? sh.scr[da][0] : 0.0f, hb ? sh.scr[db][0] : 0.0f);
        s.r1[i] = make_float2(ha ? sh.scr[da][1] : 0.0f, hb ? sh.scr[db][1] : 0.0f);
        s.r2[i] = make_float2(ha ? sh.scr[da][2] : 0.0f, hb ? sh.scr[db][2] : 0.0f);
        s.c[i] = make_float2(ha ? sh.scr[da][3] : 0.0f, hb ? sh.scr[db][3] : 0.0f);
    }
    // rotations the exact path of this thread folds: [k0, k1)
    const int k0 = NROT * half, k1 = min(n_rot, NROT * (half + 1));
    unsigned int n_exact = 0, n_repeat = 0;
    KnownPixels known;
    auto exact = [&](const double (&x)[3]) {
        if (known.seen(x)) {
            ++n_repeat;
            return;
        }
        ranges_exact_pixel(sh, x[0], x[1], x[2], k0, k1);
        known.push(x);
        ++n_exact;
    };
    auto group = [&](const typename IO::Raw &raw) {
#pragma unroll
        for (int q = 0; q < IO::NSUB; ++q) {
            float xf[GS][3];
            IO::unpack_sub_f(raw, q, dec, xf);
            bool flag[GS], any = false;
#pragma unroll
            for (int i = 0; i < GS; ++i) {
                flag[i] = screen_flag<NP>(s, xf[i], bound);
                any |= flag[i];
            }
            if (any) {
                double x[GS][3];
                IO::unpack_sub(raw, q, dec, x);
#pragma unroll
                for (int i = 0; i < GS; ++i)
                    if (flag[i]) exact(x[i]);
            }
        }
    };
    int64_t p = (int64_t)first_block * kThreads + threadIdx.x, step = (int64_t)nblocks * kThreads;
    if (VEC) {
        const int ntiles = (int)(im.npix / (kThreads * G));
        if constexpr (SPLIT == 1)
            pipe_for_each_group<IO>(pipe, base, im.plane_stride, ntiles, first_block, nblocks, zero,
                                    [&](const typename IO::Raw &raw, int64_t) { group(raw); });
        else
            pipe_for_each_group<IO, SPLIT>(pipe, base, im.plane_stride, ntiles, first_block, nblocks, zero,
                                           [&](const typename IO::Raw &raw, int64_t, int) { group(raw); });
        p = (int64_t)ntiles * kThreads * G + threadIdx.x;   // the tail belongs to block 0
        step = kThreads;
        if (first_block != 0) p = im.npix;
    }
    // scalar pixels (tails, unaligned images): each half of a split CTA must see every one of them,
    // so a half acts as a block of kThreads / SPLIT threads
    if (SPLIT > 1) {
        const int vt = threadIdx.x % (kThreads / SPLIT);
        if (VEC) {
            p = p - threadIdx.x + vt;
            step = kThreads / SPLIT;
        } else {
            p = (int64_t)first_block * (kThreads / SPLIT) + vt;
            step = (int64_t)nblocks * (kThreads / SPLIT);
        }
    }
    for (; p < im.npix; p += step) {
        double x[3];
        IO::load1(base, im.plane_stride, p, dec, x);
        const float xf[3] = {(float)x[0], (float)x[1], (float)x[2]};
        if (screen_flag<NP>(s, xf, bound)) exact(x);
    }
    if (stats) {   // diagnostics only (CT_RANGES_STATS)
        atomicAdd(&sh.n_exact, n_exact);
        atomicAdd(&sh.n_repeat, n_repeat);
    }
}

#ifndef CT_RANGES_MINB
#define CT_RANGES_MINB 3    // resident CTAs per SM the two- / four-rotation pass is compiled for
#endif
#ifndef CT_RANGES_MINB1
#define CT_RANGES_MINB1 4   // ... the one-rotation pass
#endif
// One image (slot a.which) per launch, one instantiation per (source kind, vectorised, rotations per
// pass) so that every variant gets its own register allocation: NROT = 1 (the target's iteration-0
// range), 2, or 4 (split over the CTA's halves).
template <typename IO, bool VEC, int NROT>
__global__ void __launch_bounds__(kThreads, NROT == 1 ? CT_RANGES_MINB1 : CT_RANGES_MINB) ranges_kernel(RangesArgs a) {
    extern __shared__ __align__(16) unsigned char sm_pipe[];
    __shared__ RangesShared sh;
    __shared__ double dec_d[IO::kU8 ? 256 : 1];
    __shared__ float dec_f[IO::kU8 ? 256 : 1];
    const Decode dec{dec_d, dec_f};
    const int64_t pair = blockIdx.y;
    const int z = a.which;
    const int n_rot = a.n_rot[z];
    Pipe<ranges_stages(NROT)> pipe(sm_pipe);
    if (threadIdx.x == 0) {
        pipe.init();
        sh.bad = 0;
        sh.n_exact = sh.n_repeat = 0;
    }
    if (IO::kU8) fill_decode(dec_d, dec_f, a.u8_as_f32[z]);
    if (threadIdx.x < 9 * n_rot) sh.rot[threadIdx.x] = a.rot[pair * a.rot_stride + threadIdx.x];
    if (threadIdx.x < 6 * kMaxRot) {   // the subsample's extremes
        long long v = (long long)kKeyPlusInf;
        if (threadIdx.x < 6 * n_rot) {
            const long long w = __ldcg(a.seed + (pair * 2 + z) * (6 * kMaxRot) + threadIdx.x);
            if (w != kSeedEmpty) v = w;
        }
        (&sh.key[0][0])[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3 * n_rot) {
        const int d = threadIdx.x, k = d / 3, j = d - 3 * k;
        screen_row(sh.rot + 3 * d, value_of(sh.key[k][j]), -value_of(sh.key[k][3 + j]), a.bound, sh.scr[d]);
    }
    __syncthreads();
    ranges_image<IO, VEC, (NROT > 2 ? NROT / 2 : NROT), (NROT > 2 ? 2 : 1)>(a.img[z], pair, sh, dec, n_rot, a.bound, pipe, blockIdx.x,
                                                                            gridDim.x, a.stats != nullptr, a.zero);
    __syncthreads();
    if (threadIdx.x < 6 * n_rot) {
        const long long v = (&sh.key[0][0])[threadIdx.x];
        if (v < (long long)kKeyPlusInf) atomicMin(reinterpret_cast<long long *>(a.keys + pair * a.keys_stride) + threadIdx.x, v);
    }
    if (a.status && threadIdx.x == 0 && sh.bad) a.status[pair] = CT_E_NONFINITE;
    if (a.stats && threadIdx.x == 0) {
        atomicAdd(a.stats + 0, (unsigned long long)sh.n_exact);
        atomicAdd(a.stats + 1, (unsigned long long)sh.n_repeat);
    }
}

// ---------------------------------------------------------------------------------------------
// K6: CDFs and inverse-CDF table of one pair, run by one whole block (iterative.py:45-51)
//
// LUT block of one pair (CT_IDT_LUT_DOUBLES(bins) doubles), per axis j:
//   edges[E]              np.linspace(lo, hi, bins+1), E = bins+1 rounded up to even
//   entry[E][2] = {fp, slope}  what np.interp(x, edges[1:], f, left=0) needs, besides xp = edges[k],
//                      for a sample whose bin index (before folding x == hi into the last bin) is k:
//                      k = 0 -> {0,0} (left=0), 1 <= k < bins -> {f[k-1], slope[k-1]},
//                      k = bins (x == hi) -> {f[bins-1], 0};  m = slope*(x - edges[k]) + fp.
// then {lo, hi, step, inv} per axis.
// ---------------------------------------------------------------------------------------------
struct LutArgs {
    const int64_t *keys;
    int64_t keys_stride;
    uint64_t *counts;  // [B][2][3][bins]
    double *lut;
    int32_t *status;
    int bins;
    int keep_counts;
    // optional trace
    double *tr_lo, *tr_hi, *tr_lut;
    int64_t *tr_ct, *tr_cr;
    int tr_iter, tr_niter;
};

// smem: cdf_t[bins], cdf_r[bins], f[bins]  (doubles)
__device__ void build_lut(const LutArgs &a, int64_t pair, double *sm) {
    const int bins = a.bins;
    double *cdf_t = sm, *cdf_r = sm + bins, *f = sm + 2 * bins;
    uint64_t *cnt = a.counts + pair * 6 * (int64_t)bins;
    double *lut = a.lut + pair * CT_IDT_LUT_DOUBLES(bins);
    const int64_t tr_base = ((int64_t)pair * a.tr_niter + a.tr_iter) * 3;
    for (int j = 0; j < 3; ++j) {
        bool finite;
        const AxisGrid g = grid_from_keys(a.keys + pair * a.keys_stride, j, bins, finite);
        if (!finite && a.status && threadIdx.x == 0 && !CT_STREAM_ONLY) a.status[pair] = CT_E_NONFINITE;
        for (int k = threadIdx.x; k < bins; k += kThreads) {
            const uint64_t ct_ = __ldcg(cnt + (0 * 3 + j) * bins + k);
            const uint64_t cr = __ldcg(cnt + (1 * 3 + j) * bins + k);
            cdf_t[k] = (double)ct_;
            cdf_r[k] = (double)cr;
            if (a.tr_ct) a.tr_ct[(tr_base + j) * bins + k] = (int64_t)ct_;
            if (a.tr_cr) a.tr_cr[(tr_base + j) * bins + k] = (int64_t)cr;
        }
        __syncthreads();
        // p.cumsum().astype(float); cp /= cp[-1].  The counts are integers below 2^53, so the
        // running sums are exact in any order: warp 0 scans the target, warp 1 the reference,
        // each lane owning a contiguous chunk; the IEEE division then runs on all threads.
        if (threadIdx.x < 64) {
            double *c = threadIdx.x < 32 ? cdf_t : cdf_r;
            const int lane = threadIdx.x & 31;
            const int chunk = (bins + 31) / 32;
            const int k0 = lane * chunk, k1 = min(k0 + chunk, bins);
            double mine = 0.0;
            for (int k = k0; k < k1; ++k) mine += c[k];
            double incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            double run = incl - mine;
            for (int k = k0; k < k1; ++k) {
                run += c[k];
                c[k] = run;
            }
        }
        __syncthreads();
        {
            const double total_t = cdf_t[bins - 1], total_r = cdf_r[bins - 1];
            __syncthreads();
            for (int k = threadIdx.x; k < bins; k += kThreads) {
                cdf_t[k] = div_rn(cdf_t[k], total_t);
                cdf_r[k] = div_rn(cdf_r[k], total_r);
            }
        }
        __syncthreads();
        // f = np.interp(cdf_t, cdf_r, edges[1:])
        for (int i = threadIdx.x; i < bins; i += kThreads) {
            const double x = cdf_t[i];
            int lo = 0, hi = bins;  // number of xp <= x
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (cdf_r[mid] <= x) lo = mid + 1; else hi = mid;
            }
            const int jx = lo - 1;
            double v;
            if (x != x) v = x;
            else if (jx < 0) v = edge(g, 1, bins);            // left = fp[0]
            else if (jx >= bins - 1) v = edge(g, bins, bins);  // fp[-1]
            else if (cdf_r[jx] == x) v = edge(g, jx + 1, bins);
            else {
                const double y0 = edge(g, jx + 1, bins), y1 = edge(g, jx + 2, bins);
                const double slope = div_rn(sub_rn(y1, y0), sub_rn(cdf_r[jx + 1], cdf_r[jx]));
                v = add_rn(mul_rn(slope, sub_rn(x, cdf_r[jx])), y0);
            }
            f[i] = v;
        }
        __syncthreads();
        const int E = CT_IDT_EDGE_STRIDE(bins);
        double *le = lut + (int64_t)j * E, *lt = lut + (int64_t)(3 + 2 * j) * E;
        for (int k = threadIdx.x; k <= bins; k += kThreads) {
            le[k] = edge(g, k, bins);
            double xp = 0.0, fp = 0.0, sl = 0.0;
            if (k == bins) {
                xp = g.hi;
                fp = f[bins - 1];
            } else if (k >= 1) {
                xp = edge(g, k, bins);
                fp = f[k - 1];
                // slopes of np.interp(x, edges[1:], f): (f[i+1]-f[i]) / (edges[i+2]-edges[i+1]), i = k-1
                sl = div_rn(sub_rn(f[k], f[k - 1]), sub_rn(edge(g, k + 1, bins), xp));
            }
            lt[2 * k + 0] = fp;
            lt[2 * k + 1] = sl;
            if (a.tr_lut && k < bins) a.tr_lut[(tr_base + j) * bins + k] = f[k];
        }
        if (threadIdx.x == 0) {
            double *tail = lut + (int64_t)9 * CT_IDT_EDGE_STRIDE(bins) + 4 * j;
            tail[0] = g.lo; tail[1] = g.hi; tail[2] = g.step; tail[3] = g.inv;
            // the trace reports the data range itself (before np.histogram widens lo == hi)
            if (a.tr_lo) a.tr_lo[tr_base + j] = value_of(__ldcg(a.keys + pair * a.keys_stride + j));
            if (a.tr_hi) a.tr_hi[tr_base + j] = -value_of(__ldcg(a.keys + pair * a.keys_stride + 3 + j));
        }
        __syncthreads();
    }
    if (!a.keep_counts)
        for (int k = threadIdx.x; k < 6 * bins; k += kThreads) cnt[k] = 0;
}

__global__ void __launch_bounds__(kThreads) lut_kernel(LutArgs a) {
    extern __shared__ double sm_lut[];
    build_lut(a, blockIdx.x, sm_lut);
}

// ---------------------------------------------------------------------------------------------
// K5: projection + shared-memory-privatised histograms (iterative.py:34-35, 42-43)
// Each block serves one image of one pair.  Nothing but the projection rows, the grid and the
// pipeline state lives in registers (72), so 3 CTAs = 24 warps fit per SM.  The block's three 1-D histograms are replicated
// R times, copy = lane % R, copies interleaved (index = bin*R + copy) so that the lanes of a
// warp that hit the same or neighbouring bins (smooth images) land in different banks.
// The exact edges of the three axes sit in shared memory next to them.
// ---------------------------------------------------------------------------------------------
struct HistArgs {
    Img img[2];
    int kind[2], vec[2];
    int u8_as_f32[2];
    int nblk[2];        // blocks of the target, blocks of the reference (either may be 0)
    const double *rot, *rot_next;
    int64_t rot_stride;
    const int64_t *keys;
    int64_t *keys_next;
    int64_t keys_stride;
    uint64_t *counts;
    int32_t *status;
    unsigned int *tickets;
    int bins, copies_log2;
    int fuse_lut;
    uint32_t zero;   // always 0 (ct_pipe.cuh: pipe_for_each_group)
    LutArgs lut;
};

template <bool U8>
struct HistShared {
    double rot[9];
    AxisGrid grid[3];
    bool is_last;
    double dec_d[U8 ? 256 : 1];
    float dec_f[U8 ? 256 : 1];
};

// exact bin of one sample on axis j (the slow path of the screen below)
template <typename Shared>
__device__ __noinline__ int hist_exact_bin(const Shared &sh, int j, double x0, double x1, double x2, int bins) {
    const double x[3] = {x0, x1, x2};
    const double p = dot3(sh.rot + 3 * j, x);
    return bin_exact(p, sh.grid[j].lo, sh.grid[j].inv, sh.grid[j].step, bins);
}

// CT_HIST_SCREEN: bin in packed fp32 first, exactly only where that cannot decide.
//
// The exact rule (np.histogram) needs p = rot_j . x in fp64 and a comparison with the fp64 edge: 9 fp64
// instructions per sample and axis, which made K5 instruction / latency bound (round 1: 0.77 of the HBM
// roofline, fp64 pipe 50 % busy).  But t = (p - lo) / (hi - lo) * bins only has to be known well enough to
// say on which side of an integer it lies.  So t is evaluated in fp32 with the rotation row pre-scaled
// (two axes per FFMA2), rounded to the nearest integer k' with the magic-number add, and the sample is
// SAFE when |t - k'| > eps: the bin is then k' - (t < k').  eps bounds the fp32 error of t,
//     |t32 - t| <= 2^-24 inv (5 S + 4 |lo|),  S = |x0| + |x1| + |x2|
// (fl32 of the scaled row and offset, fl32 of the pixel, three fma roundings; the grid's own edges sit
// within 1e-13 of the integers in t), evaluated per pixel.  Unsafe samples - about 1 in 1000 - take the
// exact fp64 path above, so the counts stay bit-exact.
#ifndef CT_HIST_SCREEN
#define CT_HIST_SCREEN 1
#endif

template <typename IO, bool VEC, int CL2, typename Shared>  // CL2: log2(copies) when known at compile time, else -1
__device__ __forceinline__ void hist_image(const Img &im, int64_t pair, const Shared &sh,
                                           int bins, int copies_log2_rt, unsigned int *hist, HistPipe &pipe,
                                           int first_block, int nblocks, uint32_t zero) {
    const int copies_log2 = CL2 >= 0 ? CL2 : copies_log2_rt;
    using T = typename IO::elem_t;
    const T *base = reinterpret_cast<const T *>(im.data) + pair * im.image_stride;
    constexpr int G = IO::G, GS = IO::GS;
    const Decode dec{sh.dec_d, sh.dec_f};
    const int copy = threadIdx.x & ((1 << copies_log2) - 1);
    // this lane's copy of each axis histogram as a 32-bit shared-window address: the slot of bin k
    // is one shift-add away (the generic-pointer form cost four integer instructions per sample)
    uint32_t hb[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) hb[j] = smem_u32(hist) + 4u * (uint32_t)(((j * bins) << copies_log2) + copy);
    auto count = [&](int j, int k) {
        atomicAdd(reinterpret_cast<unsigned int *>(__cvta_shared_to_generic(hb[j] + ((uint32_t)k << (copies_log2 + 2)))), 1u);
    };
    int64_t p = (int64_t)first_block * kThreads + threadIdx.x, step = (int64_t)nblocks * kThreads;
#if CT_HIST_SCREEN
    // screen rows: (axis 0, axis 1) and (axis 2, unused)
    float2 s0[2], s1[2], s2[2], sc[2];
    float ea[3], eb[3];
    {
        float r[3][3], c[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double inv = sh.grid[j].inv, lo = sh.grid[j].lo;
            const bool ok = inv > 0.0 && inv < 1e30 && fabs(lo) < 1e30;
#pragma unroll
            for (int i = 0; i < 3; ++i) r[j][i] = ok ? (float)(sh.rot[3 * j + i] * inv) : 0.0f;
            c[j] = ok ? (float)(-lo * inv) : 0.0f;                       // not ok: t = 0, every sample unsafe
            ea[j] = ok ? (float)(inv * 3.7252902984619141e-07) : 0.0f;   // 6.25 * 2^-24 inv
            eb[j] = ok ? (float)(inv * fabs(lo) * 2.9802322387695312e-07) + 1e-7f : 1.0f;   // 5 * 2^-24 inv |lo|
        }
        s0[0] = make_float2(r[0][0], r[1][0]); s1[0] = make_float2(r[0][1], r[1][1]); s2[0] = make_float2(r[0][2], r[1][2]);
        s0[1] = make_float2(r[2][0], 0.0f); s1[1] = make_float2(r[2][1], 0.0f); s2[1] = make_float2(r[2][2], 0.0f);
        sc[0] = make_float2(c[0], c[1]);
        sc[1] = make_float2(c[2], 0.0f);
    }
    const unsigned last = (unsigned)(bins - 1);
    const float2 magic = make_float2(12582912.0f, 12582912.0f);   // 1.5 * 2^23: adding it rounds to the nearest integer
    // bins of one pixel on the three axes; false: some axis could not be decided in fp32
    auto screen = [&](const float (&x)[3], int (&k)[3]) -> bool {
        const float2 x0 = make_float2(x[0], x[0]), x1 = make_float2(x[1], x[1]), x2 = make_float2(x[2], x[2]);
        const float sum = fabsf(x[0]) + fabsf(x[1]) + fabsf(x[2]);
        bool safe = true;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float2 t = __ffma2_rn(s2[h], x2, __ffma2_rn(s1[h], x1, __ffma2_rn(s0[h], x0, sc[h])));
            const float2 u = __fadd2_rn(t, magic);
            const float2 fr = __fadd2_rn(t, make_float2(-(u.x - 12582912.0f), -(u.y - 12582912.0f)));
            const float uu[2] = {u.x, u.y}, ff[2] = {fr.x, fr.y};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = 2 * h + e;
                if (j < 3) {
                    safe &= fabsf(ff[e]) > fmaf(sum, ea[j], eb[j]);     // false for NaN
                    const int kk = (__float_as_int(uu[e]) - 0x4B400000) + (__float_as_int(ff[e]) >> 31);
                    k[j] = (int)min((unsigned)kk, last);
                }
            }
        }
        return safe;
    };
    if (VEC) {
        const int ntiles = (int)(im.npix / (kThreads * G));
        pipe_for_each_group<IO>(pipe, base, im.plane_stride, ntiles, first_block, nblocks, zero,
                                [&](const typename IO::Raw &raw, int64_t) {
#if CT_STREAM_ONLY   // measurement aid: the tile pipeline alone (what this access pattern can stream), no binning
                                    const uint32_t *w32 = reinterpret_cast<const uint32_t *>(raw.e);
                                    if (w32[0] == 0x12345678u && w32[1] == 0x9abcdef0u) count(0, 0);
                                    return;
#endif
#pragma unroll
                                    for (int q = 0; q < IO::NSUB; ++q) {
                                        float xf[GS][3];
                                        IO::unpack_sub_f(raw, q, dec, xf);
                                        int k[GS][3];
                                        bool safe[GS], all = true;
#pragma unroll
                                        for (int i = 0; i < GS; ++i) {
                                            safe[i] = screen(xf[i], k[i]);
                                            all &= safe[i];
                                        }
                                        if (!all) {
                                            double x[GS][3];
                                            IO::unpack_sub(raw, q, dec, x);
#pragma unroll
                                            for (int i = 0; i < GS; ++i)
                                                if (!safe[i]) {
#pragma unroll
                                                    for (int j = 0; j < 3; ++j) k[i][j] = hist_exact_bin(sh, j, x[i][0], x[i][1], x[i][2], bins);
                                                }
                                        }
#pragma unroll
                                        for (int i = 0; i < GS; ++i)
#pragma unroll
                                            for (int j = 0; j < 3; ++j) count(j, k[i][j]);
                                    }
                                });
        p = (int64_t)ntiles * kThreads * G + threadIdx.x;   // the tail belongs to block 0
        step = kThreads;
        if (first_block != 0) return;
    }
    for (; p < im.npix; p += step) {
        double x[3];
        IO::load1(base, im.plane_stride, p, dec, x);
#pragma unroll
        for (int j = 0; j < 3; ++j) count(j, hist_exact_bin(sh, j, x[0], x[1], x[2], bins));
    }
#else
    double r[9], lo[3], inv[3], stp[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) r[i] = sh.rot[i];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        lo[j] = sh.grid[j].lo;
        inv[j] = sh.grid[j].inv;
        stp[j] = sh.grid[j].step;
    }
    auto one = [&](const double(&x)[3]) {
#pragma unroll
        for (int j = 0; j < 3; ++j) count(j, bin_exact(dot3(r + 3 * j, x), lo[j], inv[j], stp[j], bins));
    };
    if (VEC) {
        const int ntiles = (int)(im.npix / (kThreads * G));
        pipe_for_each_group<IO>(pipe, base, im.plane_stride, ntiles, first_block, nblocks, zero,
                                [&](const typename IO::Raw &raw, int64_t) {
#pragma unroll
                                    for (int q = 0; q < IO::NSUB; ++q) {
                                        double x[GS][3];
                                        IO::unpack_sub(raw, q, dec, x);
#pragma unroll
                                        for (int i = 0; i < GS; ++i) one(x[i]);
                                    }
                                });
        p = (int64_t)ntiles * kThreads * G + threadIdx.x;   // the tail belongs to block 0
        step = kThreads;
        if (first_block != 0) return;
    }
    for (; p < im.npix; p += step) {
        double x[3];
        IO::load1(base, im.plane_stride, p, dec, x);
        one(x);
    }
#endif
}

// ANY_U8: the instantiation that also holds the uint8 variants (launched when either image is uint8);
// float / double images run the instantiation without them.
template <bool ANY_U8>
#ifndef CT_HIST_MINB
#define CT_HIST_MINB 3
#endif
__global__ void __launch_bounds__(kThreads, CT_HIST_MINB) hist_kernel(HistArgs a) {
    // tile pipeline (stages, then one barrier set per image) | histograms; the front is later
    // reused by the LUT build
    extern __shared__ __align__(16) double sm_dyn[];
    __shared__ HistShared<ANY_U8> sh;
    const int64_t pair = blockIdx.y;
    const int bins = a.bins;
    unsigned int *hist = reinterpret_cast<unsigned int *>(sm_dyn + kHistSmemFront / 8);
    if (threadIdx.x < 9) sh.rot[threadIdx.x] = a.rot[pair * a.rot_stride + threadIdx.x];
    if (threadIdx.x >= 32 && threadIdx.x < 35) {
        bool finite;
        sh.grid[threadIdx.x - 32] = grid_from_keys(a.keys + pair * a.keys_stride, threadIdx.x - 32, bins, finite);
    }
    const int nslots = (3 * bins) << a.copies_log2;
    const int copies = 1 << a.copies_log2;

    // Every CTA takes its 1/gridDim.x share of the target tiles and then of the reference tiles:
    // identical work per CTA whatever the two images' sizes and dtypes are (a static split of the
    // CTAs between the images is unbalanced as soon as their bytes per pixel differ).
    for (int z = 0; z < 2; ++z) {
        if (!a.nblk[z]) continue;  // this image is absent (stage API)
        HistPipe pipe(sm_dyn, z);
        if (threadIdx.x == 0) pipe.init();
        if (ANY_U8 && a.kind[z] >= 4) fill_decode(sh.dec_d, sh.dec_f, a.u8_as_f32[z]);
        for (int i = threadIdx.x; i < nslots; i += kThreads) hist[i] = 0u;
        __syncthreads();
        const int sel = a.kind[z] * 2 + a.vec[z];
#define CT_HIST_CALL(T, L, V, CL2V) \
    hist_image<PixelIO<T, L>, V, CL2V>(a.img[z], pair, sh, bins, a.copies_log2, hist, pipe, blockIdx.x, gridDim.x, a.zero)
#define CT_CASE_3(ID, T, L, V) case ID: CT_HIST_CALL(T, L, V, CT_HIST_COPIES_LOG2); break;
#define CT_CASE_g(ID, T, L, V) case ID: CT_HIST_CALL(T, L, V, -1); break;
        if (a.copies_log2 == CT_HIST_COPIES_LOG2) {  // bins <= 256: the default 255
            switch (sel) {
                CT_FOR_EACH_FLOAT_SRC(CT_CASE_3)
                default:
                    if constexpr (ANY_U8) {
                        switch (sel) { CT_FOR_EACH_U8_SRC(CT_CASE_3) }
                    }
            }
        } else {
            switch (sel) {
                CT_FOR_EACH_FLOAT_SRC(CT_CASE_g)
                default:
                    if constexpr (ANY_U8) {
                        switch (sel) { CT_FOR_EACH_U8_SRC(CT_CASE_g) }
                    }
            }
        }
#undef CT_HIST_CALL
#undef CT_CASE_3
#undef CT_CASE_g
        __syncthreads();
        // flush: sum the copies of each bin, one 64-bit integer atomic per non-empty bin
        uint64_t *cnt = a.counts + (pair * 2 + z) * 3 * (int64_t)bins;
        for (int i = threadIdx.x; i < 3 * bins; i += kThreads) {
            unsigned int s = 0;
            for (int c = 0; c < copies; ++c) s += hist[(i << a.copies_log2) + c];
            if (s) atomicAdd(reinterpret_cast<unsigned long long *>(cnt + i), (unsigned long long)s);
        }
        __syncthreads();
    }
    if (!a.fuse_lut) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&a.tickets[pair], 1u);
        sh.is_last = (t == gridDim.x - 1u);
        if (sh.is_last) a.tickets[pair] = 0;
    }
    __syncthreads();
    if (!sh.is_last) return;
    __threadfence();
    build_lut(a.lut, pair, sm_dyn);
}

// ---------------------------------------------------------------------------------------------
// K7: remap + back-rotation + state update + next range (iterative.py:53, 55)
// ---------------------------------------------------------------------------------------------
struct RemapArgs {
    Img src;
    ImgOut dst;
    int kind, vec;
    int u8_as_f32;   // uint8 source: decode k/255 in float32 (else float64)
    int dst_kind;    // kDstState (planar fp64 state), kDstF64 (final fp64 HWC), or a fused final conversion:
                     // kDstU8 / kDstU8Planar (clip + round to uint8), kDstF32 (float32 HWC, optional clamp)
    int clamp;
    int dst_vec;     // the destination allows vector stores (else element stores; the source still streams through TMA)
    const double *rot, *rot_next;
    int64_t rot_stride;
    int64_t *keys_next;
    int64_t keys_stride;
    const double *lut;
    int32_t *status;
    int bins;
    int round_f32;
    uint32_t zero;   // always 0 (ct_pipe.cuh: pipe_for_each_group)
};
enum { kDstState = 0, kDstF64 = 1, kDstU8 = 2, kDstF32 = 3, kDstU8Planar = 4 };

template <bool U8>
struct RemapShared {
    double rot[18];
    AxisGrid grid[3];
    double red[kWarps][6];
    unsigned amb[3];
    double dec_d[U8 ? 256 : 1];
    float dec_f[U8 ? 256 : 1];
};

// K7 takes the bin of a sample as floor(t~), t~ = (p - lo) * inv rounded to a multiple of 2^-20, unless t~ is an
// integer or one step (2^-20) beside one; those samples - about 3 in 2^20 - compare against the tabulated edge
// as before.  An unflagged t~ is at least 2^-19 away from every integer, so floor(t~) is the searchsorted answer
// against the tabulated edges as long as t~ and the edges' positions in t-space are together off by less than
// that: the rounding of t~ contributes 2^-21, the arithmetic of t~ (p - lo, inv, the product) at most
// 3 * 2^-53 * bins, and the edges, computed as k * step + lo with three roundings, sit within
// 7 * 2^-53 * max(|lo|, |hi|) of lo + k (hi - lo) / bins, i.e. within 7 * 2^-53 * max(|lo|, |hi|) * inv in t-space.
// remap_ambiguity() demands 3 * 2^-53 * bins + 2^-51 * max(|lo|, |hi|) * inv < 2^-22, which keeps the total below
// 2^-21 + 2^-21 - a quarter of the 2^-19 available - and holds for every grid whose width is not below ~1e-9 of
// its magnitude; otherwise every sample is flagged.
#ifndef CT_REMAP_FLOOR
#define CT_REMAP_FLOOR 1
#endif
constexpr int kFracBits = 20;
constexpr double kMagicFrac = 6442450944.0;   // 1.5 * 2^32: ulp 2^-20
__device__ __forceinline__ unsigned remap_ambiguity(const AxisGrid &g, int bins) {
    const double slack = 3.0 * 1.1102230246251565e-16 * bins + 4.440892098500626e-16 * fmax(fabs(g.lo), fabs(g.hi)) * g.inv;
    return slack < 2.3841857910156250e-07 /* 2^-22 */ ? 3u : (1u << kFracBits);
}

template <typename SIO, typename DIO, bool VEC, bool NEXT, bool ROUND32>
__device__ __forceinline__ void remap_image(const RemapArgs &a, int64_t pair, const RemapShared<SIO::kU8> &sh,
                                            const double *tab, RemapPipe &pipe, double (&mn)[6], bool &bad) {
    using TS = typename SIO::elem_t;
    using TD = typename DIO::elem_t;
    const TS *src = reinterpret_cast<const TS *>(a.src.data) + pair * a.src.image_stride;
    TD *dst = reinterpret_cast<TD *>(a.dst.data) + pair * a.dst.image_stride;
    constexpr int G = SIO::G, GS = SIO::GS;
    const Decode dec{sh.dec_d, sh.dec_f};
    const bool clamp = a.clamp != 0, dst_vec = a.dst_vec != 0;
    // 256-bit stores of a float64 destination (ct_common.cuh: store)
    const bool wide = dst_vec && (reinterpret_cast<uintptr_t>(dst) & 31) == 0 && (DIO::kLayout == CT_HWC || (a.dst.plane_stride & 3) == 0);
    const int bins = a.bins;
    const int E = CT_IDT_EDGE_STRIDE(bins);
    double r[9], rn[9], lo[3], inv[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        r[i] = sh.rot[i];
        rn[i] = NEXT ? sh.rot[9 + i] : 0.0;
    }
    unsigned amb[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        lo[j] = sh.grid[j].lo;
        inv[j] = sh.grid[j].inv;
        amb[j] = sh.amb[j];
    }
    auto one = [&](const double(&x)[3], double(&y)[3]) {
#if CT_STREAM_ONLY   // measurement aid: read the state, write it back (what this access pattern can stream)
        y[0] = x[0]; y[1] = x[1]; y[2] = x[2];
        return;
#endif
        double d[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double p = dot3(r + 3 * j, x);
            const double *edges = tab + j * E;
            const double2 *ent = reinterpret_cast<const double2 *>(tab + (3 + 2 * j) * E);
#if CT_REMAP_FLOOR
            // floor((p - lo) * inv) straight from a fixed-point magic number (20 fractional bits in the low
            // mantissa word); only a sample within 2^-20 of a grid line - where the estimate and the exact
            // edge could disagree - goes through the comparison against the tabulated edge.  amb[j] is 3,
            // or 2^20 ("always compare") on a grid too narrow for its magnitude (remap_kernel prologue).
            const unsigned w = (unsigned)__double2loint(fma(p - lo[j], inv[j], kMagicFrac));
            int k = (int)(w >> kFracBits);
            if (((w + 1u) & ((1u << kFracBits) - 1u)) < amb[j]) {
                const int ke = bin_estimate(p, lo[j], inv[j], bins);
                k = ke - (p < edges[ke] ? 1 : 0);
            }
            k = (int)min((unsigned)k, (unsigned)bins);
#else
            const int ke = bin_estimate(p, lo[j], inv[j], bins);
            // bin index BEFORE folding p == hi into the last bin: ke - (p < edges[ke]) in [0, bins]
            const int k = (int)min((unsigned)(ke - (p < edges[ke] ? 1 : 0)), (unsigned)bins);
#endif
            const double xp = edges[k];
            const double2 fs = ent[k];
            const double fp = fs.x, sl = fs.y;
            double m = add_rn(mul_rn(sl, sub_rn(p, xp)), fp);   // np.interp: slope*(x - xp[j]) + fp[j]
            if (ROUND32) m = (double)(float)m;                  // float32 d_r buffer, iterative.py:36
            d[j] = sub_rn(m, p);
        }
        // solve(r, d) for orthogonal r is r^T d; then "+ target"
#pragma unroll
        for (int c = 0; c < 3; ++c)
            y[c] = add_rn(fma(r[6 + c], d[2], fma(r[3 + c], d[1], mul_rn(r[c], d[0]))), x[c]);
        if (NEXT) track_range<false>(rn, y, mn, bad);
    };
    int64_t p = (int64_t)blockIdx.x * kThreads + threadIdx.x, step = (int64_t)gridDim.x * kThreads;
    if (VEC) {
        const int ntiles = (int)(a.src.npix / (kThreads * G));
        pipe_for_each_group<SIO>(pipe, src, a.src.plane_stride, ntiles, blockIdx.x, gridDim.x, a.zero,
                                 [&](const typename SIO::Raw &raw, int64_t tile0) {
#pragma unroll
                                     for (int q = 0; q < SIO::NSUB; ++q) {
                                         double x[GS][3], y[GS][3];
                                         SIO::unpack_sub(raw, q, dec, x);
#pragma unroll
                                         for (int i = 0; i < GS; ++i) one(x[i], y[i]);
                                         if (dst_vec) DIO::template store<true, GS>(dst, a.dst.plane_stride, pipe_sub_pixel0<SIO>(tile0, q), y, clamp, wide);
                                         else DIO::template store<false, GS>(dst, a.dst.plane_stride, pipe_sub_pixel0<SIO>(tile0, q), y, clamp);
                                     }
                                 });
        p = (int64_t)ntiles * kThreads * G + threadIdx.x;   // the tail belongs to block 0
        step = kThreads;
        if (blockIdx.x != 0) return;
    }
    for (; p < a.src.npix; p += step) {
        double x[3], y[3];
        SIO::load1(src, a.src.plane_stride, p, dec, x);
        one(x, y);
        DIO::store1(dst, a.dst.plane_stride, p, y, clamp);
    }
}

// one instantiation per (source kind, vectorised); destination, NEXT and ROUND32 are block-uniform
// runtime switches inside.  The fused final conversions exist for the planar fp64 state as the
// source only (the last of n_iter >= 2 iterations).
template <typename SIO, bool VEC>
__device__ __forceinline__ void remap_dispatch(const RemapArgs &a, int64_t pair, const RemapShared<SIO::kU8> &sh,
                                               const double *tab, RemapPipe &pipe, bool next, double (&mn)[6], bool &bad) {
    using StateIO = PixelIO<double, CT_CHW>;
    using FinalIO = PixelIO<double, CT_HWC>;
    constexpr bool kFromState = same_io<SIO, StateIO>::value;
    if (kFromState && a.dst_kind >= kDstU8) {
        if (a.dst_kind == kDstU8) remap_image<SIO, PixelIO<uint8_t, CT_HWC>, VEC, false, false>(a, pair, sh, tab, pipe, mn, bad);
        else if (a.dst_kind == kDstU8Planar) remap_image<SIO, PixelIO<uint8_t, CT_CHW>, VEC, false, false>(a, pair, sh, tab, pipe, mn, bad);
        else remap_image<SIO, PixelIO<float, CT_HWC>, VEC, false, false>(a, pair, sh, tab, pipe, mn, bad);
        return;
    }
    if (a.round_f32) {  // iteration 0 of float32 input: the state buffer is always the destination or n_iter == 1
        if (a.dst_kind == kDstState) {
            if (next) remap_image<SIO, StateIO, VEC, true, true>(a, pair, sh, tab, pipe, mn, bad);
            else remap_image<SIO, StateIO, VEC, false, true>(a, pair, sh, tab, pipe, mn, bad);
        } else {
            if (next) remap_image<SIO, FinalIO, VEC, true, true>(a, pair, sh, tab, pipe, mn, bad);
            else remap_image<SIO, FinalIO, VEC, false, true>(a, pair, sh, tab, pipe, mn, bad);
        }
    } else if (a.dst_kind == kDstState) {
        if (next) remap_image<SIO, StateIO, VEC, true, false>(a, pair, sh, tab, pipe, mn, bad);
        else remap_image<SIO, StateIO, VEC, false, false>(a, pair, sh, tab, pipe, mn, bad);
    } else {
        if (next) remap_image<SIO, FinalIO, VEC, true, false>(a, pair, sh, tab, pipe, mn, bad);
        else remap_image<SIO, FinalIO, VEC, false, false>(a, pair, sh, tab, pipe, mn, bad);
    }
}

template <typename SIO, bool VEC>
__global__ void __launch_bounds__(kThreads, CT_MINB) remap_kernel(RemapArgs a) {
    extern __shared__ __align__(16) double sm_remap[];  // tile pipeline | edges + {fp, slope} entries of the three axes
    __shared__ RemapShared<SIO::kU8> sh;
    RemapPipe pipe(sm_remap);
    double *sm_tab = sm_remap + pipe_bytes(kRemapStages) / 8;
    if (threadIdx.x == 0) pipe.init();
    const int64_t pair = blockIdx.y;
    const int bins = a.bins;
    const bool next = a.rot_next != nullptr && a.keys_next != nullptr;
    const double *lut = a.lut + pair * CT_IDT_LUT_DOUBLES(bins);
    if (threadIdx.x < 9) sh.rot[threadIdx.x] = a.rot[pair * a.rot_stride + threadIdx.x];
    else if (threadIdx.x < 18 && next) sh.rot[threadIdx.x] = a.rot_next[pair * a.rot_stride + threadIdx.x - 9];
    if (SIO::kU8) fill_decode(sh.dec_d, sh.dec_f, a.u8_as_f32);
    if (threadIdx.x >= 32 && threadIdx.x < 35) {
        const double *tail = lut + 9 * CT_IDT_EDGE_STRIDE(bins) + 4 * (threadIdx.x - 32);
        const AxisGrid g{tail[0], tail[1], tail[2], tail[3]};
        sh.grid[threadIdx.x - 32] = g;
        sh.amb[threadIdx.x - 32] = remap_ambiguity(g, bins);
    }
    for (int i = threadIdx.x; i < 9 * CT_IDT_EDGE_STRIDE(bins); i += kThreads) sm_tab[i] = lut[i];
    __syncthreads();

    double mn[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) mn[i] = i < 3 ? INFINITY : -INFINITY;
    bool bad = false;
    remap_dispatch<SIO, VEC>(a, pair, sh, sm_tab, pipe, next, mn, bad);
    if (next) fold_range(mn, a.keys_next + pair * a.keys_stride, sh.red);
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
// blocks per image for a persistent, tile-strided launch: one wave of `occ` resident CTAs per SM
// shared by `units` images, never more than one CTA per tile
template <typename K>
static int resident_blocks(const ct_context *h, K kernel, size_t smem, int64_t npix, int64_t units) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, smem) != cudaSuccess || occ < 1) occ = 1;
    const int64_t want = (npix / 2 + kThreads - 1) / kThreads;
    int64_t cap = ((int64_t)h->sm_count * occ) / (units > 0 ? units : 1);
    if (cap < 1) cap = 1;
    const int64_t n = want < cap ? want : cap;
    return (int)(n < 1 ? 1 : n);
}


int launch_keys_init(ct_context *h, int64_t *keys, int64_t n) {
    if (!keys || n <= 0) return fail(h, CT_E_INVALID, "bad keys_init arguments");
    keys_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(keys, n);
    h->launches++;
    CT_CUDA(h, cudaGetLastError());
    return CT_OK;
}

template <typename IO, bool V, int N>
static int launch_ranges_one(ct_context *h, const RangesArgs &a, int64_t npix, int count) {
    constexpr int kRangesSmem = pipe_bytes(ranges_stages(N));
    if (kRangesSmem > 48 * 1024) {
        static bool raised[64] = {};
        if (!raised[h->device & 63]) {
            CT_CUDA(h, cudaFuncSetAttribute(ranges_kernel<IO, V, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRangesSmem));
            raised[h->device & 63] = true;
        }
    }
    ranges_kernel<IO, V, N><<<dim3(resident_blocks(h, ranges_kernel<IO, V, N>, kRangesSmem, npix, count), count), kThreads,
                              kRangesSmem, h->stream>>>(a);
    return CT_OK;
}

// target (one rotation: iteration 0) and / or reference (n_rot_r rotations): per chunk of kMaxRot
// reference rotations one K4a seed launch and one K4b launch; either image may be NULL.  With
// init_keys the first seed launch also sets those keys to "+inf" (saves the keys_init launch).
int launch_ranges_pair(ct_context *h, const ct_batch *target, const ct_batch *reference, int n_rot_r,
                       const double *rot, int64_t rot_stride, int64_t *keys, int64_t keys_stride, int32_t *status,
                       int64_t *init_keys, int64_t n_init) {
    if (!target && !reference) return fail(h, CT_E_INVALID, "ranges needs at least one image");
    if (target) CT_TRY(check_batch(h, target, "target"));
    if (reference) CT_TRY(check_batch(h, reference, "reference"));
    if (!rot || !keys) return fail(h, CT_E_INVALID, "rot/keys is NULL");
    if (reference && n_rot_r < 1) return fail(h, CT_E_INVALID, "n_rot must be >= 1");
    if (target && reference && target->count != reference->count) return fail(h, CT_E_INVALID, "batch counts differ");
    const int count = target ? target->count : reference->count;
    CT_TRY(ensure_seed(h, (size_t)count * 2 * 6 * kMaxRot));
    for (int k0 = 0; k0 == 0 || (reference && k0 < n_rot_r); k0 += kMaxRot) {
        RangesArgs a{};
        const ct_batch *imgs[2] = {target && k0 == 0 ? target : nullptr, reference};
        for (int z = 0; z < 2; ++z) {
            if (!imgs[z]) continue;
            a.img[z] = img_of(imgs[z]);
            a.kind[z] = src_kind(imgs[z]);
            a.vec[z] = vec_ok(imgs[z]);
            a.u8_as_f32[z] = (imgs[z]->flags & CT_BATCH_U8_AS_F32) != 0;
            a.n_rot[z] = z == 0 ? 1 : (n_rot_r - k0 < kMaxRot ? n_rot_r - k0 : kMaxRot);
        }
        a.rot = rot + 9 * k0;
        a.rot_stride = rot_stride;
        a.keys = keys + CT_IDT_KEYS * k0;
        a.keys_stride = keys_stride;
        a.status = status;
        a.bound = h->ranges_bound;
        a.seed = h->seed;
        a.stats = h->ranges_stats;
        if (k0 == 0) {
            a.init_keys = init_keys;
            a.n_init = n_init;
        }
        int64_t npix_max = 0;
        for (int z = 0; z < 2; ++z)
            if (imgs[z] && imgs[z]->npix > npix_max) npix_max = imgs[z]->npix;
        int64_t seed_ctas = (npix_max / CT_SEED_FRACTION + 4 * kThreads - 1) / (4 * kThreads);
        seed_ctas = seed_ctas < 4 ? 4 : (seed_ctas > 1024 ? 1024 : seed_ctas);
        CT_CUDA(h, cudaMemsetAsync(h->seed, 0x7f, sizeof(long long) * (size_t)count * 2 * 6 * kMaxRot, h->stream));
        ranges_seed_kernel<<<dim3((unsigned)seed_ctas, count, 2), kThreads, 0, h->stream>>>(a);
        h->launches++;
        CT_CUDA(h, cudaGetLastError());
        prof_mark(h, CT_PROF_SEED);
        for (int z = 0; z < 2; ++z) {
            if (!imgs[z]) continue;
            a.which = z;
            const int nr = a.n_rot[z] > 2 ? 4 : a.n_rot[z];   // instantiated widths: 1, 2, 4
            const int64_t npix = imgs[z]->npix;
            int rc = CT_OK;
            switch (a.kind[z] * 2 + a.vec[z]) {
#define CT_CASE(ID, T, L, V)                                                                  \
    case ID:                                                                                  \
        rc = nr == 1   ? launch_ranges_one<PixelIO<T, L>, V, 1>(h, a, npix, count)            \
             : nr == 2 ? launch_ranges_one<PixelIO<T, L>, V, 2>(h, a, npix, count)            \
                       : launch_ranges_one<PixelIO<T, L>, V, 4>(h, a, npix, count);           \
        break;
                CT_FOR_EACH_SRC(CT_CASE)
#undef CT_CASE
            }
            CT_TRY(rc);
            h->launches++;
            CT_CUDA(h, cudaGetLastError());
            prof_mark(h, z == 0 ? CT_PROF_RANGES_TARGET : CT_PROF_RANGES_REFERENCE);
        }
    }
    return CT_OK;
}

int launch_ranges(ct_context *h, const ct_batch *img, const double *rot, int64_t rot_stride, int n_rot,
                  int64_t *keys, int64_t keys_stride, int32_t *status) {
    if (n_rot < 1) return fail(h, CT_E_INVALID, "n_rot must be >= 1");
    if (n_rot == 1) return launch_ranges_pair(h, img, nullptr, 0, rot, rot_stride, keys, keys_stride, status, nullptr, 0);
    return launch_ranges_pair(h, nullptr, img, n_rot, rot, rot_stride, keys, keys_stride, status, nullptr, 0);
}

static int check_stage(ct_context *h, const ct_idt_stage *s) {
    if (!s) return fail(h, CT_E_INVALID, "stage is NULL");
    if (s->bins < 1) return fail(h, CT_E_INVALID, "bins must be >= 1");
    if (s->bins > CT_IDT_MAX_BINS) return fail(h, CT_E_UNSUPPORTED, "bins=%d exceeds CT_IDT_MAX_BINS=%d", s->bins, CT_IDT_MAX_BINS);
    return CT_OK;
}

static LutArgs lut_args(const ct_idt_stage *s, int keep_counts, const ct_idt_trace *tr, int it, int niter) {
    LutArgs l{};
    l.keys = s->keys;
    l.keys_stride = s->keys_stride;
    l.counts = s->counts;
    l.lut = s->lut;
    l.status = s->status;
    l.bins = s->bins;
    l.keep_counts = keep_counts;
    if (tr) {
        l.tr_lo = tr->lo; l.tr_hi = tr->hi; l.tr_lut = tr->lut;
        l.tr_ct = tr->counts_t; l.tr_cr = tr->counts_r;
    }
    l.tr_iter = it;
    l.tr_niter = niter > 0 ? niter : 1;
    return l;
}

static int copies_log2_for(int bins) { return bins <= 256 ? CT_HIST_COPIES_LOG2 : (bins <= 512 ? (CT_HIST_COPIES_LOG2 < 2 ? CT_HIST_COPIES_LOG2 : 2) : 0); }

int launch_hist(ct_context *h, const ct_idt_stage *s, int fuse_lut, const ct_idt_trace *trace,
                int trace_iter, int trace_niter) {
    CT_TRY(check_stage(h, s));
    if (!s->target && !s->reference) return fail(h, CT_E_INVALID, "hist needs at least one image");
    if (!s->rot || !s->keys || !s->counts) return fail(h, CT_E_INVALID, "rot/keys/counts is NULL");
    if (fuse_lut && (!s->lut || !s->target || !s->reference)) return fail(h, CT_E_INVALID, "fused LUT needs both images and lut");
    HistArgs a{};
    int B = 0;
    int64_t npix[2] = {0, 0};
    const ct_batch *imgs[2] = {s->target, s->reference};
    for (int z = 0; z < 2; ++z) {
        if (!imgs[z]) continue;
        CT_TRY(check_batch(h, imgs[z], z ? "reference" : "target"));
        if (B && imgs[z]->count != B) return fail(h, CT_E_INVALID, "target/reference batch counts differ");
        B = imgs[z]->count;
        a.img[z] = img_of(imgs[z]);
        a.kind[z] = src_kind(imgs[z]);
        a.vec[z] = vec_ok(imgs[z]);
        a.u8_as_f32[z] = (imgs[z]->flags & CT_BATCH_U8_AS_F32) != 0;
        npix[z] = imgs[z]->npix;
    }
    const size_t smem = (size_t)kHistSmemFront + (size_t)((3 * s->bins) << copies_log2_for(s->bins)) * sizeof(unsigned int);
    // (the pipeline region alone is >= the 3*bins doubles the fused LUT build reuses)
    const bool any_u8 = (imgs[0] && imgs[0]->dtype == CT_U8) || (imgs[1] && imgs[1]->dtype == CT_U8);
    if (smem > 48 * 1024 && !h->hist_smem_raised) {
        CT_CUDA(h, cudaFuncSetAttribute(hist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
        CT_CUDA(h, cudaFuncSetAttribute(hist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
        h->hist_smem_raised = true;
    }
    // one wave of resident CTAs per launch (persistent, tile-strided); every CTA serves both images
    int occ = 0;
    if (any_u8) CT_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hist_kernel<true>, kThreads, smem));
    else CT_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hist_kernel<false>, kThreads, smem));
    if (occ < 1) occ = 1;
    int64_t budget = ((int64_t)h->sm_count * occ) / B;
    const int64_t want = ((npix[0] > npix[1] ? npix[0] : npix[1]) / 2 + kThreads - 1) / kThreads;
    if (budget > want) budget = want;
    if (budget < 1) budget = 1;
    const int nblk = (int)budget;
    for (int z = 0; z < 2; ++z) a.nblk[z] = imgs[z] ? nblk : 0;
    a.rot = s->rot;
    a.rot_next = s->rot_next;
    a.rot_stride = s->rot_stride;
    a.keys = s->keys;
    a.keys_next = s->keys_next;
    a.keys_stride = s->keys_stride;
    a.counts = s->counts;
    a.status = s->status;
    a.bins = s->bins;
    a.copies_log2 = copies_log2_for(s->bins);
    a.fuse_lut = fuse_lut;
    CT_TRY(ensure_scratch(h, B));
    a.tickets = h->tickets;
    a.lut = lut_args(s, 0, trace, trace_iter, trace_niter);
    if (any_u8) hist_kernel<true><<<dim3(nblk, B), kThreads, smem, h->stream>>>(a);
    else hist_kernel<false><<<dim3(nblk, B), kThreads, smem, h->stream>>>(a);
    h->launches++;
    CT_CUDA(h, cudaGetLastError());
    prof_mark(h, CT_PROF_HIST);
    return CT_OK;
}

int launch_lut(ct_context *h, const ct_idt_stage *s, int keep_counts, const ct_idt_trace *trace,
               int trace_iter, int trace_niter) {
    CT_TRY(check_stage(h, s));
    if (!s->keys || !s->counts || !s->lut || !s->target) return fail(h, CT_E_INVALID, "keys/counts/lut/target is NULL");
    const LutArgs l = lut_args(s, keep_counts, trace, trace_iter, trace_niter);
    lut_kernel<<<s->target->count, kThreads, (size_t)3 * s->bins * sizeof(double), h->stream>>>(l);
    h->launches++;
    CT_CUDA(h, cudaGetLastError());
    return CT_OK;
}

int launch_remap(ct_context *h, const ct_idt_stage *s, const ct_batch *dst, int round_f32) {
    CT_TRY(check_stage(h, s));
    CT_TRY(check_batch(h, s->target, "target"));
    CT_TRY(check_batch(h, dst, "dst"));
    if (!s->rot || !s->lut) return fail(h, CT_E_INVALID, "rot/lut is NULL");
    if (dst->npix != s->target->npix || dst->count != s->target->count)
        return fail(h, CT_E_INVALID, "dst must have the target's npix and count");
    RemapArgs a{};
    a.src = img_of(s->target);
    a.dst = imgout_of(dst);
    a.kind = src_kind(s->target);
    a.vec = vec_ok(s->target);
    a.dst_vec = vec_ok(dst);
    a.u8_as_f32 = (s->target->flags & CT_BATCH_U8_AS_F32) != 0;
    a.clamp = (dst->flags & CT_BATCH_CLAMP01) != 0;
    if (dst->dtype == CT_F64) {
        a.dst_kind = dst->layout == CT_CHW ? kDstState : kDstF64;
    } else {
        // fused final conversion: only from the planar fp64 state (the last of n_iter >= 2 iterations)
        const bool from_state = s->target->dtype == CT_F64 && s->target->layout == CT_CHW;
        if (!from_state) return fail(h, CT_E_UNSUPPORTED, "uint8 / float32 IDT output needs the planar float64 state as the source");
        if (dst->dtype == CT_U8) a.dst_kind = dst->layout == CT_HWC ? kDstU8 : kDstU8Planar;
        else if (dst->dtype == CT_F32 && dst->layout == CT_HWC) a.dst_kind = kDstF32;
        else return fail(h, CT_E_UNSUPPORTED, "IDT output must be float64, uint8, or float32 CT_HWC");
        if (s->rot_next || s->keys_next) return fail(h, CT_E_INVALID, "a converted output is final: rot_next / keys_next must be NULL");
    }
    a.rot = s->rot;
    a.rot_next = s->rot_next;
    a.rot_stride = s->rot_stride;
    a.keys_next = s->keys_next;
    a.keys_stride = s->keys_stride;
    a.lut = s->lut;
    a.status = s->status;
    a.bins = s->bins;
    a.round_f32 = round_f32;
    const size_t smem = (size_t)pipe_bytes(kRemapStages) + (size_t)9 * CT_IDT_EDGE_STRIDE(s->bins) * sizeof(double);
    switch (a.kind * 2 + a.vec) {
#define CT_CASE(ID, T, L, V)                                                                                   \
    case ID:                                                                                                   \
        if (smem > 40 * 1024 && !h->remap_smem_raised[ID]) {                                                   \
            CT_CUDA(h, cudaFuncSetAttribute(remap_kernel<PixelIO<T, L>, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                            pipe_bytes(kRemapStages) + 9 * CT_IDT_EDGE_STRIDE(CT_IDT_MAX_BINS) * (int)sizeof(double)));                \
            h->remap_smem_raised[ID] = true;                                                                   \
        }                                                                                                      \
        remap_kernel<PixelIO<T, L>, V><<<dim3(resident_blocks(h, remap_kernel<PixelIO<T, L>, V>, smem, s->target->npix, \
                                                                s->target->count), s->target->count),          \
                                         kThreads, smem, h->stream>>>(a);                                      \
        break;
        CT_FOR_EACH_SRC(CT_CASE)
#undef CT_CASE
    }
    h->launches++;
    CT_CUDA(h, cudaGetLastError());
    prof_mark(h, CT_PROF_REMAP);
    return CT_OK;
}

}  // namespace ct
