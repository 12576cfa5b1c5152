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

namespace ct {

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

__device__ __forceinline__ double edge(const AxisGrid &g, int k, int bins) {
    return k >= bins ? g.hi : add_rn(mul_rn((double)k, g.step), g.lo);
}

// Unique k in [0, bins-1] with edges[k] <= x < edges[k+1] (last bin closed); x in [lo, hi].
// Also returns edges[k].
__device__ __forceinline__ int bin_of(const AxisGrid &g, int bins, double x, double &e_lo) {
    int k = (int)((x - g.lo) * g.inv);
    k = k < 0 ? 0 : (k > bins - 1 ? bins - 1 : k);
    double e0 = edge(g, k, bins);
    if (x < e0) {
        --k;
        e0 = edge(g, k, bins);
    } else if (k != bins - 1) {
        const double e1 = edge(g, k + 1, bins);
        if (x >= e1) {
            ++k;
            e0 = e1;
        }
    }
    e_lo = e0;
    return k;
}

__device__ __forceinline__ bool not_finite(double p) {
    return (__double2hiint(p) & 0x7ff00000) == 0x7ff00000;
}

// fold per-thread minima of (p0,p1,p2,-p0,-p1,-p2) into the pair's keys
__device__ __forceinline__ void fold_range(double (&mn)[6], int64_t *keys, double (*red)[6]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const double v = warp_min(mn[i]);
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

__device__ __forceinline__ void track_range(const double *rot, const double (&x)[3], double (&mn)[6], bool &bad) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double p = dot3(rot + 3 * j, x);
        bad |= not_finite(p);
        mn[j] = fmin(mn[j], p);
        mn[3 + j] = fmin(mn[3 + j], -p);
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
struct RangesArgs {
    Img img;
    int kind, vec;
    const double *rot;
    int64_t rot_stride;
    int64_t *keys;
    int64_t keys_stride;
    int32_t *status;
};

template <typename IO, bool VEC>
__device__ __forceinline__ void ranges_image(const Img &im, int64_t pair, const double *rot,
                                             int64_t first_block, int64_t nblocks, double (&mn)[6], bool &bad) {
    using T = typename IO::elem_t;
    const T *base = reinterpret_cast<const T *>(im.data) + pair * im.image_stride;
    constexpr int G = IO::G;
    const int64_t ngroups = im.npix / G;
    const int64_t stride = nblocks * kThreads;
    for (int64_t g = first_block * kThreads + threadIdx.x; g < ngroups; g += stride) {
        double x[G][3];
        IO::template load<VEC>(base, im.plane_stride, g, x);
#pragma unroll
        for (int i = 0; i < G; ++i) track_range(rot, x[i], mn, bad);
    }
    if (first_block == 0 && threadIdx.x == 0) {
        for (int64_t p = ngroups * G; p < im.npix; ++p) {
            double x[3];
            IO::load1(base, im.plane_stride, p, x);
            track_range(rot, x, mn, bad);
        }
    }
}

__global__ void __launch_bounds__(kThreads) ranges_kernel(RangesArgs a) {
    const int64_t pair = blockIdx.y;
    __shared__ double rot[9];
    __shared__ double red[kWarps][6];
    if (threadIdx.x < 9) rot[threadIdx.x] = a.rot[pair * a.rot_stride + threadIdx.x];
    __syncthreads();
    double mn[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) mn[i] = INFINITY;
    bool bad = false;
    switch (a.kind * 2 + a.vec) {
#define CT_CASE(ID, T, L, V) \
    case ID: ranges_image<PixelIO<T, L>, V>(a.img, pair, rot, blockIdx.x, gridDim.x, mn, bad); break;
        CT_FOR_EACH_SRC(CT_CASE)
#undef CT_CASE
    }
    fold_range(mn, a.keys + pair * a.keys_stride, red);
    if (a.status && __syncthreads_or(bad) && threadIdx.x == 0) a.status[pair] = CT_E_NONFINITE;
}

// ---------------------------------------------------------------------------------------------
// K6: CDFs and inverse-CDF table of one pair, run by one whole block (iterative.py:45-51)
// ---------------------------------------------------------------------------------------------
struct LutArgs {
    const int64_t *keys;
    int64_t keys_stride;
    uint64_t *counts;  // [B][2][3][bins]
    double *lut;       // [B][3*(2*bins+4)]
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
        // integer counts -> running sums kept as doubles (exact below 2^53), one thread per image
        for (int k = threadIdx.x; k < bins; k += kThreads) {
            const uint64_t ct_ = __ldcg(cnt + (0 * 3 + j) * bins + k);
            const uint64_t cr = __ldcg(cnt + (1 * 3 + j) * bins + k);
            cdf_t[k] = (double)ct_;
            cdf_r[k] = (double)cr;
            if (a.tr_ct) a.tr_ct[(tr_base + j) * bins + k] = (int64_t)ct_;
            if (a.tr_cr) a.tr_cr[(tr_base + j) * bins + k] = (int64_t)cr;
        }
        __syncthreads();
        if (threadIdx.x < 2) {  // p.cumsum().astype(float); cp /= cp[-1]
            double *c = threadIdx.x == 0 ? cdf_t : cdf_r;
            double run = 0.0;
            for (int k = 0; k < bins; ++k) {
                run += c[k];
                c[k] = run;
            }
            const double total = run;
            for (int k = 0; k < bins; ++k) c[k] = div_rn(c[k], total);
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
        double *lf = lut + j * 2 * bins, *ls = lf + bins;
        for (int i = threadIdx.x; i < bins; i += kThreads) {
            lf[i] = f[i];
            // slopes of np.interp(x, edges[1:], f): (f[i+1]-f[i]) / (edges[i+2]-edges[i+1])
            ls[i] = i < bins - 1 ? div_rn(sub_rn(f[i + 1], f[i]), sub_rn(edge(g, i + 2, bins), edge(g, i + 1, bins))) : 0.0;
            if (a.tr_lut) a.tr_lut[(tr_base + j) * bins + i] = f[i];
        }
        if (threadIdx.x == 0) {
            double *tail = lut + 3 * 2 * bins + 4 * j;
            tail[0] = g.lo; tail[1] = g.hi; tail[2] = g.step; tail[3] = g.inv;
            if (a.tr_lo) a.tr_lo[tr_base + j] = g.lo;
            if (a.tr_hi) a.tr_hi[tr_base + j] = g.hi;
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
// Each block serves one image of one pair.  The block's three 1-D histograms are replicated
// R times, copy = lane % R, copies interleaved (index = bin*R + copy) so that the lanes of a
// warp that hit the same or neighbouring bins (smooth images) land in different banks.
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

template <typename IO, bool VEC, bool NEXT>
__device__ __forceinline__ void hist_image(const Img &im, int64_t pair, const double *rot, const double *rot_next,
                                           const AxisGrid *grid, int bins, int copies_log2, unsigned int *hist,
                                           int64_t first_block, int64_t nblocks, double (&mn)[6], bool &bad) {
    using T = typename IO::elem_t;
    const T *base = reinterpret_cast<const T *>(im.data) + pair * im.image_stride;
    constexpr int G = IO::G;
    const int64_t ngroups = im.npix / G;
    const int64_t stride = nblocks * kThreads;
    const int copy = threadIdx.x & ((1 << copies_log2) - 1);
    auto one = [&](const double(&x)[3]) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double p = dot3(rot + 3 * j, x);
            if (p >= grid[j].lo && p <= grid[j].hi) {  // np.histogram's `keep`
                double e;
                const int k = bin_of(grid[j], bins, p, e);
                atomicAdd(&hist[((j * bins + k) << copies_log2) + copy], 1u);
            }
        }
        if (NEXT) track_range(rot_next, x, mn, bad);
    };
    for (int64_t g = first_block * kThreads + threadIdx.x; g < ngroups; g += stride) {
        double x[G][3];
        IO::template load<VEC>(base, im.plane_stride, g, x);
#pragma unroll
        for (int i = 0; i < G; ++i) one(x[i]);
    }
    if (first_block == 0 && threadIdx.x == 0) {
        for (int64_t p = ngroups * G; p < im.npix; ++p) {
            double x[3];
            IO::load1(base, im.plane_stride, p, x);
            one(x);
        }
    }
}

__global__ void __launch_bounds__(kThreads) hist_kernel(HistArgs a) {
    extern __shared__ double sm_dyn[];  // histograms, later reused by the LUT build
    unsigned int *hist = reinterpret_cast<unsigned int *>(sm_dyn);
    __shared__ double rot[18];
    __shared__ AxisGrid grid[3];
    __shared__ double red[kWarps][6];
    __shared__ bool is_last;
    const int64_t pair = blockIdx.y;
    const int bins = a.bins;
    const int z = blockIdx.x < a.nblk[0] ? 0 : 1;
    const int64_t first_block = z == 0 ? blockIdx.x : blockIdx.x - a.nblk[0];
    const bool next = (z == 1) && a.rot_next != nullptr && a.keys_next != nullptr;

    if (threadIdx.x < 9) rot[threadIdx.x] = a.rot[pair * a.rot_stride + threadIdx.x];
    else if (threadIdx.x < 18 && next) rot[threadIdx.x] = a.rot_next[pair * a.rot_stride + threadIdx.x - 9];
    if (threadIdx.x >= 32 && threadIdx.x < 35) {
        bool finite;
        grid[threadIdx.x - 32] = grid_from_keys(a.keys + pair * a.keys_stride, threadIdx.x - 32, bins, finite);
    }
    const int nslots = (3 * bins) << a.copies_log2;
    for (int i = threadIdx.x; i < nslots; i += kThreads) hist[i] = 0u;
    __syncthreads();

    double mn[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) mn[i] = INFINITY;
    bool bad = false;
    const int sel = a.kind[z] * 2 + a.vec[z];
    if (next) {
        switch (sel) {
#define CT_CASE(ID, T, L, V)                                                                              \
    case ID: hist_image<PixelIO<T, L>, V, true>(a.img[z], pair, rot, rot + 9, grid, bins, a.copies_log2,  \
                                                hist, first_block, a.nblk[z], mn, bad); break;
            CT_FOR_EACH_SRC(CT_CASE)
#undef CT_CASE
        }
    } else {
        switch (sel) {
#define CT_CASE(ID, T, L, V)                                                                              \
    case ID: hist_image<PixelIO<T, L>, V, false>(a.img[z], pair, rot, rot + 9, grid, bins, a.copies_log2, \
                                                 hist, first_block, a.nblk[z], mn, bad); break;
            CT_FOR_EACH_SRC(CT_CASE)
#undef CT_CASE
        }
    }
    __syncthreads();
    // flush: sum the copies of each bin, one 64-bit integer atomic per non-empty bin
    uint64_t *cnt = a.counts + (pair * 2 + z) * 3 * (int64_t)bins;
    const int copies = 1 << a.copies_log2;
    for (int i = threadIdx.x; i < 3 * bins; i += kThreads) {
        unsigned int s = 0;
        for (int c = 0; c < copies; ++c) s += hist[(i << a.copies_log2) + c];
        if (s) atomicAdd(reinterpret_cast<unsigned long long *>(cnt + i), (unsigned long long)s);
    }
    if (next) {
        fold_range(mn, a.keys_next + pair * a.keys_stride, red);
        if (a.status && __syncthreads_or(bad) && threadIdx.x == 0) a.status[pair] = CT_E_NONFINITE;
    }
    if (!a.fuse_lut) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&a.tickets[pair], 1u);
        is_last = (t == gridDim.x - 1u);
        if (is_last) a.tickets[pair] = 0;
    }
    __syncthreads();
    if (!is_last) return;
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

template <typename SIO, typename DIO, bool VEC, bool NEXT>
__device__ __forceinline__ void remap_image(const RemapArgs &a, int64_t pair, const double *rot, const double *rot_next,
                                            const AxisGrid *grid, const double *lut, double (&mn)[6], bool &bad) {
    using TS = typename SIO::elem_t;
    const TS *src = reinterpret_cast<const TS *>(a.src.data) + pair * a.src.image_stride;
    double *dst = reinterpret_cast<double *>(a.dst.data) + pair * a.dst.image_stride;
    constexpr int G = SIO::G;
    const int bins = a.bins;
    const bool round_f32 = a.round_f32 != 0;
    const int64_t ngroups = a.src.npix / G;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    auto one = [&](const double(&x)[3], double(&y)[3]) {
        double d[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double p = dot3(rot + 3 * j, x);
            const double *f = lut + j * 2 * bins, *sl = f + bins;
            double e;
            const int k = bin_of(grid[j], bins, p, e);
            double m;
            if (p >= grid[j].hi) m = f[bins - 1];   // x == xp[-1]
            else if (k == 0) m = 0.0;                // x < edges[1]: np.interp(..., left=0)
            else m = add_rn(mul_rn(sl[k - 1], sub_rn(p, e)), f[k - 1]);
            if (round_f32) m = (double)(float)m;     // float32 d_r buffer, iterative.py:36
            d[j] = sub_rn(m, p);
        }
        // solve(r, d) for orthogonal r is r^T d; then "+ target"
#pragma unroll
        for (int c = 0; c < 3; ++c)
            y[c] = add_rn(fma(rot[6 + c], d[2], fma(rot[3 + c], d[1], mul_rn(rot[c], d[0]))), x[c]);
        if (NEXT) track_range(rot_next, y, mn, bad);
    };
    for (int64_t g = (int64_t)blockIdx.x * kThreads + threadIdx.x; g < ngroups; g += stride) {
        double x[G][3], y[G][3];
        SIO::template load<VEC>(src, a.src.plane_stride, g, x);
#pragma unroll
        for (int i = 0; i < G; ++i) one(x[i], y[i]);
        DIO::template store<VEC, G>(dst, a.dst.plane_stride, g * G, y);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t p = ngroups * G; p < a.src.npix; ++p) {
            double x[3], y[3];
            SIO::load1(src, a.src.plane_stride, p, x);
            one(x, y);
            DIO::store1(dst, a.dst.plane_stride, p, y);
        }
    }
}

__global__ void __launch_bounds__(kThreads) remap_kernel(RemapArgs a) {
    extern __shared__ double sm_tab[];  // f and slope of the three axes
    __shared__ double rot[18];
    __shared__ AxisGrid grid[3];
    __shared__ double red[kWarps][6];
    const int64_t pair = blockIdx.y;
    const int bins = a.bins;
    const bool next = a.rot_next != nullptr && a.keys_next != nullptr;
    const double *lut = a.lut + pair * CT_IDT_LUT_DOUBLES(bins);
    if (threadIdx.x < 9) rot[threadIdx.x] = a.rot[pair * a.rot_stride + threadIdx.x];
    else if (threadIdx.x < 18 && next) rot[threadIdx.x] = a.rot_next[pair * a.rot_stride + threadIdx.x - 9];
    if (threadIdx.x >= 32 && threadIdx.x < 35) {
        const double *tail = lut + 6 * bins + 4 * (threadIdx.x - 32);
        grid[threadIdx.x - 32] = AxisGrid{tail[0], tail[1], tail[2], tail[3]};
    }
    for (int i = threadIdx.x; i < 6 * bins; i += kThreads) sm_tab[i] = lut[i];
    __syncthreads();

    double mn[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) mn[i] = INFINITY;
    bool bad = false;
    const int sel = a.kind * 2 + a.vec;
#define CT_REMAP_SWITCH(DIO, NEXTV)                                                                      \
    switch (sel) {                                                                                       \
        case 0: remap_image<PixelIO<float, CT_HWC>, DIO, false, NEXTV>(a, pair, rot, rot + 9, grid, sm_tab, mn, bad); break;  \
        case 1: remap_image<PixelIO<float, CT_HWC>, DIO, true, NEXTV>(a, pair, rot, rot + 9, grid, sm_tab, mn, bad); break;   \
        case 2: remap_image<PixelIO<float, CT_CHW>, DIO, false, NEXTV>(a, pair, rot, rot + 9, grid, sm_tab, mn, bad); break;  \
        case 3: remap_image<PixelIO<float, CT_CHW>, DIO, true, NEXTV>(a, pair, rot, rot + 9, grid, sm_tab, mn, bad); break;   \
        case 4: remap_image<PixelIO<double, CT_HWC>, DIO, false, NEXTV>(a, pair, rot, rot + 9, grid, sm_tab, mn, bad); break; \
        case 5: remap_image<PixelIO<double, CT_HWC>, DIO, true, NEXTV>(a, pair, rot, rot + 9, grid, sm_tab, mn, bad); break;  \
        case 6: remap_image<PixelIO<double, CT_CHW>, DIO, false, NEXTV>(a, pair, rot, rot + 9, grid, sm_tab, mn, bad); break; \
        case 7: remap_image<PixelIO<double, CT_CHW>, DIO, true, NEXTV>(a, pair, rot, rot + 9, grid, sm_tab, mn, bad); break;  \
    }
    using StateIO = PixelIO<double, CT_CHW>;
    using FinalIO = PixelIO<double, CT_HWC>;
    if (a.dst_layout == CT_CHW) {
        if (next) { CT_REMAP_SWITCH(StateIO, true) } else { CT_REMAP_SWITCH(StateIO, false) }
    } else {
        if (next) { CT_REMAP_SWITCH(FinalIO, true) } else { CT_REMAP_SWITCH(FinalIO, false) }
    }
#undef CT_REMAP_SWITCH
    if (next) {
        fold_range(mn, a.keys_next + pair * a.keys_stride, red);
        if (a.status && __syncthreads_or(bad) && threadIdx.x == 0) a.status[pair] = CT_E_NONFINITE;
    }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static int stream_blocks(const ct_context *h, int64_t npix, int64_t units, int per_sm) {
    const int64_t want = (npix / 2 + kThreads - 1) / kThreads;
    int64_t cap = ((int64_t)h->sm_count * per_sm) / (units > 0 ? units : 1);
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

int launch_ranges(ct_context *h, const ct_batch *img, const double *rot, int64_t rot_stride,
                  int64_t *keys, int64_t keys_stride, int32_t *status) {
    CT_TRY(check_batch(h, img, "images"));
    if (!rot || !keys) return fail(h, CT_E_INVALID, "rot/keys is NULL");
    RangesArgs a{img_of(img), src_kind(img), vec_ok(img), rot, rot_stride, keys, keys_stride, status};
    const int nblk = stream_blocks(h, img->npix, img->count, 8);
    ranges_kernel<<<dim3(nblk, img->count), kThreads, 0, h->stream>>>(a);
    h->launches++;
    CT_CUDA(h, cudaGetLastError());
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

static int copies_log2_for(int bins) { return bins <= 256 ? 3 : (bins <= 512 ? 2 : 1); }

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
    // split a budget of ~4 blocks per SM between the two images in proportion to their pixels
    int64_t budget = ((int64_t)h->sm_count * 4) / B;
    if (budget < 2) budget = 2;
    for (int z = 0; z < 2; ++z) {
        if (!imgs[z]) continue;
        int64_t share = (int64_t)((double)budget * (double)npix[z] / (double)(npix[0] + npix[1]) + 0.5);
        const int64_t want = (npix[z] / 2 + kThreads - 1) / kThreads;
        if (share > want) share = want;
        a.nblk[z] = (int)(share < 1 ? 1 : share);
    }
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
    size_t smem = (size_t)((3 * s->bins) << a.copies_log2) * sizeof(unsigned int);
    const size_t lut_smem = (size_t)3 * s->bins * sizeof(double);
    if (fuse_lut && lut_smem > smem) smem = lut_smem;
    hist_kernel<<<dim3(a.nblk[0] + a.nblk[1], B), kThreads, smem, h->stream>>>(a);
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
    const int nblk = stream_blocks(h, s->target->npix, s->target->count, 8);
    if ((size_t)6 * s->bins * sizeof(double) > 40 * 1024) {  // above the default 48 KB with the static part
        if (!h->remap_smem_raised) {
            CT_CUDA(h, cudaFuncSetAttribute(remap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            6 * CT_IDT_MAX_BINS * (int)sizeof(double)));
            h->remap_smem_raised = true;
        }
    }
    remap_kernel<<<dim3(nblk, s->target->count), kThreads, (size_t)6 * s->bins * sizeof(double), h->stream>>>(a);
    h->launches++;
    CT_CUDA(h, cudaGetLastError());
    return CT_OK;
}

}  // namespace ct
