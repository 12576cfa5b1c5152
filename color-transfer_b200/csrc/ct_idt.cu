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
#ifndef CT_HIST_STAGES
#define CT_HIST_STAGES 3
#endif
#ifndef CT_REMAP_STAGES
#define CT_REMAP_STAGES 5
#endif
namespace ct {

constexpr int kRangesStages = 3, kHistStages = CT_HIST_STAGES, kRemapStages = CT_REMAP_STAGES;
using RangesPipe = Pipe<kRangesStages>;
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
// K4: projected range of one image under one rotation (iterative.py:34-35, 39-40)
// ---------------------------------------------------------------------------------------------
constexpr int kMaxRot = 4;  // most rotations one pass can take

struct RangesArgs {
    Img img;
    int kind, vec;
    int n_rot;  // 1..kMaxRot consecutive rotations (rot + 9k) folded into consecutive key slots (keys + 6k)
    const double *rot;
    int64_t rot_stride;
    int64_t *keys;
    int64_t keys_stride;
    int32_t *status;
};

template <typename IO, bool VEC, int NROT>
__device__ __forceinline__ void ranges_image(const Img &im, int64_t pair, const double *rot, int n_rot, RangesPipe &pipe,
                                             int first_block, int nblocks, double (&mn)[NROT][6], bool &bad) {
    using T = typename IO::elem_t;
    const T *base = reinterpret_cast<const T *>(im.data) + pair * im.image_stride;
    constexpr int G = IO::G;
    auto group = [&](const double(*x)[3], int n) {
#pragma unroll
        for (int k = 0; k < NROT; ++k) {
            if (NROT == 1 || k < n_rot) {
                double r[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) r[i] = rot[9 * k + i];   // shared-memory broadcast
#pragma unroll
                for (int i = 0; i < G; ++i)
                    if (i < n) {
                        if (k == 0) track_range<true>(r, x[i], mn[k], bad);
                        else track_range<false>(r, x[i], mn[k], bad);
                    }
            }
        }
    };
    int64_t p = (int64_t)first_block * kThreads + threadIdx.x, step = (int64_t)nblocks * kThreads;
    if (VEC) {
        const int ntiles = (int)(im.npix / (kThreads * G));
        pipe_for_each_group<IO>(pipe, base, im.plane_stride, ntiles, first_block, nblocks,
                                [&](const typename IO::Raw &raw, int64_t) {
                                    double x[G][3];
                                    IO::unpack(raw, x);
                                    group(x, G);
                                });
        p = (int64_t)ntiles * kThreads * G + threadIdx.x;   // the tail belongs to block 0
        step = kThreads;
        if (first_block != 0) return;
    }
    for (; p < im.npix; p += step) {
        double x[G][3];
        IO::load1(base, im.plane_stride, p, x[0]);
        group(x, 1);
    }
}

template <int NROT>  // rotations per pass: registers hold NROT x 6 running minima
__global__ void __launch_bounds__(kThreads, NROT == 1 ? 3 : 2) ranges_kernel(RangesArgs a) {
    extern __shared__ __align__(16) unsigned char sm_pipe[];
    const int64_t pair = blockIdx.y;
    __shared__ double rot[9 * kMaxRot];
    __shared__ double red[kWarps][6];
    RangesPipe pipe(sm_pipe);
    if (threadIdx.x == 0) pipe.init();
    if (threadIdx.x < 9 * a.n_rot) rot[threadIdx.x] = a.rot[pair * a.rot_stride + threadIdx.x];
    __syncthreads();
    double mn[NROT][6];
#pragma unroll
    for (int k = 0; k < NROT; ++k)
#pragma unroll
        for (int i = 0; i < 6; ++i) mn[k][i] = i < 3 ? INFINITY : -INFINITY;
    bool bad = false;
    switch (a.kind * 2 + a.vec) {
#define CT_CASE(ID, T, L, V) \
    case ID: ranges_image<PixelIO<T, L>, V, NROT>(a.img, pair, rot, a.n_rot, pipe, blockIdx.x, gridDim.x, mn, bad); break;
        CT_FOR_EACH_SRC(CT_CASE)
#undef CT_CASE
    }
#pragma unroll
    for (int k = 0; k < NROT; ++k) {
        if (k < a.n_rot) {
            fold_range(mn[k], a.keys + pair * a.keys_stride + CT_IDT_KEYS * k, red);
            __syncthreads();
        }
    }
    if (a.status && __syncthreads_or(bad) && threadIdx.x == 0) a.status[pair] = CT_E_NONFINITE;
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
        if (!finite && a.status && threadIdx.x == 0) a.status[pair] = CT_E_NONFINITE;
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
    LutArgs lut;
};

struct HistShared {
    double rot[9];
    AxisGrid grid[3];
    bool is_last;
};

template <typename IO, bool VEC, int CL2>  // CL2: log2(copies) when known at compile time, else -1
__device__ __forceinline__ void hist_image(const Img &im, int64_t pair, const HistShared &sh,
                                           int bins, int copies_log2_rt, unsigned int *hist, HistPipe &pipe,
                                           int first_block, int nblocks) {
    const int copies_log2 = CL2 >= 0 ? CL2 : copies_log2_rt;
    using T = typename IO::elem_t;
    const T *base = reinterpret_cast<const T *>(im.data) + pair * im.image_stride;
    constexpr int G = IO::G;
    const int copy = threadIdx.x & ((1 << copies_log2) - 1);
    double r[9], lo[3], inv[3], stp[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) r[i] = sh.rot[i];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        lo[j] = sh.grid[j].lo;
        inv[j] = sh.grid[j].inv;
        stp[j] = sh.grid[j].step;
    }
    // this lane's copy of each axis histogram as a 32-bit shared-window address: the slot of bin k
    // is one shift-add away (the generic-pointer form cost four integer instructions per sample)
    uint32_t hb[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) hb[j] = smem_u32(hist) + 4u * (uint32_t)(((j * bins) << copies_log2) + copy);
    auto one = [&](const double(&x)[3]) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double p = dot3(r + 3 * j, x);
            const int k = bin_exact(p, lo[j], inv[j], stp[j], bins);
            atomicAdd(reinterpret_cast<unsigned int *>(__cvta_shared_to_generic(hb[j] + ((uint32_t)k << (copies_log2 + 2)))), 1u);
        }
    };
    int64_t p = (int64_t)first_block * kThreads + threadIdx.x, step = (int64_t)nblocks * kThreads;
    if (VEC) {
        const int ntiles = (int)(im.npix / (kThreads * G));
        pipe_for_each_group<IO>(pipe, base, im.plane_stride, ntiles, first_block, nblocks,
                                [&](const typename IO::Raw &raw, int64_t) {
                                    double x[G][3];
                                    IO::unpack(raw, x);
#pragma unroll
                                    for (int i = 0; i < G; ++i) one(x[i]);
                                });
        p = (int64_t)ntiles * kThreads * G + threadIdx.x;   // the tail belongs to block 0
        step = kThreads;
        if (first_block != 0) return;
    }
    for (; p < im.npix; p += step) {
        double x[3];
        IO::load1(base, im.plane_stride, p, x);
        one(x);
    }
}

__global__ void __launch_bounds__(kThreads, 3) hist_kernel(HistArgs a) {
    // tile pipeline (stages, then one barrier set per image) | histograms; the front is later
    // reused by the LUT build
    extern __shared__ __align__(16) double sm_dyn[];
    __shared__ HistShared sh;
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
        for (int i = threadIdx.x; i < nslots; i += kThreads) hist[i] = 0u;
        __syncthreads();
        const int sel = a.kind[z] * 2 + a.vec[z];
#define CT_HIST_CALL(T, L, V, CL2V) \
    hist_image<PixelIO<T, L>, V, CL2V>(a.img[z], pair, sh, bins, a.copies_log2, hist, pipe, blockIdx.x, gridDim.x)
#define CT_CASE_3(ID, T, L, V) case ID: CT_HIST_CALL(T, L, V, 3); break;
#define CT_CASE_g(ID, T, L, V) case ID: CT_HIST_CALL(T, L, V, -1); break;
        if (a.copies_log2 == 3) {  // bins <= 256: the default 255
            switch (sel) { CT_FOR_EACH_SRC(CT_CASE_3) }
        } else {
            switch (sel) { CT_FOR_EACH_SRC(CT_CASE_g) }
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
    int dst_layout;  // CT_CHW (planar fp64 state) or CT_HWC (final fp64 output)
    const double *rot, *rot_next;
    int64_t rot_stride;
    int64_t *keys_next;
    int64_t keys_stride;
    const double *lut;
    int32_t *status;
    int bins;
    int round_f32;
};

struct RemapShared {
    double rot[18];
    AxisGrid grid[3];
    double red[kWarps][6];
};

template <typename SIO, typename DIO, bool VEC, bool NEXT, bool ROUND32>
__device__ __forceinline__ void remap_image(const RemapArgs &a, int64_t pair, const RemapShared &sh,
                                            const double *tab, RemapPipe &pipe, double (&mn)[6], bool &bad) {
    using TS = typename SIO::elem_t;
    const TS *src = reinterpret_cast<const TS *>(a.src.data) + pair * a.src.image_stride;
    double *dst = reinterpret_cast<double *>(a.dst.data) + pair * a.dst.image_stride;
    constexpr int G = SIO::G;
    const int bins = a.bins;
    const int E = CT_IDT_EDGE_STRIDE(bins);
    double r[9], rn[9], lo[3], inv[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        r[i] = sh.rot[i];
        rn[i] = NEXT ? sh.rot[9 + i] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        lo[j] = sh.grid[j].lo;
        inv[j] = sh.grid[j].inv;
    }
    auto one = [&](const double(&x)[3], double(&y)[3]) {
        double d[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double p = dot3(r + 3 * j, x);
            const double *edges = tab + j * E;
            const double2 *ent = reinterpret_cast<const double2 *>(tab + (3 + 2 * j) * E);
            const int ke = bin_estimate(p, lo[j], inv[j], bins);
            // bin index BEFORE folding p == hi into the last bin: ke - (p < edges[ke]) in [0, bins]
            const int k = (int)min((unsigned)(ke - (p < edges[ke] ? 1 : 0)), (unsigned)bins);
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
        pipe_for_each_group<SIO>(pipe, src, a.src.plane_stride, ntiles, blockIdx.x, gridDim.x,
                                 [&](const typename SIO::Raw &raw, int64_t pix0) {
                                     double x[G][3], y[G][3];
                                     SIO::unpack(raw, x);
#pragma unroll
                                     for (int i = 0; i < G; ++i) one(x[i], y[i]);
                                     DIO::template store<true, G>(dst, a.dst.plane_stride, pix0, y);
                                 });
        p = (int64_t)ntiles * kThreads * G + threadIdx.x;   // the tail belongs to block 0
        step = kThreads;
        if (blockIdx.x != 0) return;
    }
    for (; p < a.src.npix; p += step) {
        double x[3], y[3];
        SIO::load1(src, a.src.plane_stride, p, x);
        one(x, y);
        DIO::store1(dst, a.dst.plane_stride, p, y);
    }
}

// one instantiation per (source kind, vectorised); destination layout, NEXT and ROUND32 are
// block-uniform runtime switches inside
template <typename SIO, bool VEC>
__device__ __forceinline__ void remap_dispatch(const RemapArgs &a, int64_t pair, const RemapShared &sh,
                                               const double *tab, RemapPipe &pipe, bool next, double (&mn)[6], bool &bad) {
    using StateIO = PixelIO<double, CT_CHW>;
    using FinalIO = PixelIO<double, CT_HWC>;
    if (a.round_f32) {  // iteration 0 of float32 input: the state buffer is always the destination or n_iter == 1
        if (a.dst_layout == CT_CHW) {
            if (next) remap_image<SIO, StateIO, VEC, true, true>(a, pair, sh, tab, pipe, mn, bad);
            else remap_image<SIO, StateIO, VEC, false, true>(a, pair, sh, tab, pipe, mn, bad);
        } else {
            if (next) remap_image<SIO, FinalIO, VEC, true, true>(a, pair, sh, tab, pipe, mn, bad);
            else remap_image<SIO, FinalIO, VEC, false, true>(a, pair, sh, tab, pipe, mn, bad);
        }
    } else if (a.dst_layout == CT_CHW) {
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
    __shared__ RemapShared sh;
    RemapPipe pipe(sm_remap);
    double *sm_tab = sm_remap + pipe_bytes(kRemapStages) / 8;
    if (threadIdx.x == 0) pipe.init();
    const int64_t pair = blockIdx.y;
    const int bins = a.bins;
    const bool next = a.rot_next != nullptr && a.keys_next != nullptr;
    const double *lut = a.lut + pair * CT_IDT_LUT_DOUBLES(bins);
    if (threadIdx.x < 9) sh.rot[threadIdx.x] = a.rot[pair * a.rot_stride + threadIdx.x];
    else if (threadIdx.x < 18 && next) sh.rot[threadIdx.x] = a.rot_next[pair * a.rot_stride + threadIdx.x - 9];
    if (threadIdx.x >= 32 && threadIdx.x < 35) {
        const double *tail = lut + 9 * CT_IDT_EDGE_STRIDE(bins) + 4 * (threadIdx.x - 32);
        sh.grid[threadIdx.x - 32] = AxisGrid{tail[0], tail[1], tail[2], tail[3]};
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

int launch_ranges(ct_context *h, const ct_batch *img, const double *rot, int64_t rot_stride, int n_rot,
                  int64_t *keys, int64_t keys_stride, int32_t *status) {
    CT_TRY(check_batch(h, img, "images"));
    if (!rot || !keys) return fail(h, CT_E_INVALID, "rot/keys is NULL");
    if (n_rot < 1) return fail(h, CT_E_INVALID, "n_rot must be >= 1");
    const int chunk = n_rot == 1 ? 1 : CT_RANGES_CHUNK;
    for (int k0 = 0; k0 < n_rot; k0 += chunk) {  // `chunk` rotations per pass over the image
        const int n = n_rot - k0 < chunk ? n_rot - k0 : chunk;
        RangesArgs a{img_of(img), src_kind(img), vec_ok(img), n, rot + 9 * k0, rot_stride, keys + CT_IDT_KEYS * k0, keys_stride, status};
        if (chunk == 1) {
            const int nblk = resident_blocks(h, ranges_kernel<1>, pipe_bytes(kRangesStages), img->npix, img->count);
            ranges_kernel<1><<<dim3(nblk, img->count), kThreads, pipe_bytes(kRangesStages), h->stream>>>(a);
        } else {
            const int nblk = resident_blocks(h, ranges_kernel<CT_RANGES_CHUNK>, pipe_bytes(kRangesStages), img->npix, img->count);
            ranges_kernel<CT_RANGES_CHUNK><<<dim3(nblk, img->count), kThreads, pipe_bytes(kRangesStages), h->stream>>>(a);
        }
        h->launches++;
        CT_CUDA(h, cudaGetLastError());
    }
    return CT_OK;
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

static int copies_log2_for(int bins) { return bins <= 256 ? 3 : (bins <= 512 ? 2 : 0); }

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
        npix[z] = imgs[z]->npix;
    }
    const size_t smem = (size_t)kHistSmemFront + (size_t)((3 * s->bins) << copies_log2_for(s->bins)) * sizeof(unsigned int);
    // (the pipeline region alone is >= the 3*bins doubles the fused LUT build reuses)
    if (smem > 48 * 1024 && !h->hist_smem_raised) {
        CT_CUDA(h, cudaFuncSetAttribute(hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
        h->hist_smem_raised = true;
    }
    // one wave of resident CTAs per launch (persistent, tile-strided); every CTA serves both images
    int occ = 0;
    CT_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hist_kernel, kThreads, smem));
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
    hist_kernel<<<dim3(nblk, B), kThreads, smem, h->stream>>>(a);
    h->launches++;
    CT_CUDA(h, cudaGetLastError());
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
    if (dst->dtype != CT_F64) return fail(h, CT_E_INVALID, "IDT state/output must be float64");
    if (dst->npix != s->target->npix || dst->count != s->target->count)
        return fail(h, CT_E_INVALID, "dst must have the target's npix and count");
    RemapArgs a{};
    a.src = img_of(s->target);
    a.dst = imgout_of(dst);
    a.kind = src_kind(s->target);
    a.vec = vec_ok(s->target) && vec_ok(dst);
    a.dst_layout = dst->layout;
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
    return CT_OK;
}

}  // namespace ct
