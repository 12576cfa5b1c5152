// C ABI of libct_b200.so (include/ct_b200.h): handle management, argument checking, the fused
// device drivers and the host-buffer pipelines.
#include <chrono>
#include <new>
#include <vector>

#include <stdlib.h>

#include "ct_context.h"
#include "ct_host_copy.h"

namespace ct {

int check_batch(ct_context *h, const ct_batch *b, const char *name) {
    if (!b) return fail(h, CT_E_INVALID, "%s is NULL", name);
    if (!b->data) return fail(h, CT_E_INVALID, "%s.data is NULL", name);
    if (b->npix <= 0) return fail(h, CT_E_INVALID, "%s.npix must be positive (got %lld)", name, (long long)b->npix);
    if (b->count <= 0) return fail(h, CT_E_INVALID, "%s.count must be positive", name);
    if (b->dtype != CT_F32 && b->dtype != CT_F64 && b->dtype != CT_U8) return fail(h, CT_E_INVALID, "%s.dtype unknown", name);
    if (b->layout != CT_HWC && b->layout != CT_CHW) return fail(h, CT_E_INVALID, "%s.layout unknown", name);
    if (b->count > 65535) return fail(h, CT_E_UNSUPPORTED, "%s.count above 65535 pairs per call", name);
    if (b->npix >= ((int64_t)1 << 31)) return fail(h, CT_E_UNSUPPORTED, "%s.npix must be below 2^31 pixels per image", name);
    if (((uintptr_t)b->data) % elem_size(b->dtype)) return fail(h, CT_E_INVALID, "%s.data is not element aligned", name);
    return CT_OK;
}

template <typename T>
static int grow(ct_context *h, T **ptr, size_t *have, size_t want, bool zero) {
    if (*have >= want && *ptr) return CT_OK;
    if (*ptr) {
        CT_CUDA(h, cudaStreamSynchronize(h->stream));
        CT_CUDA(h, cudaFree(*ptr));
        *ptr = nullptr;
        *have = 0;
    }
    size_t n = want + want / 4;
    void *p = nullptr;
    if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) {
        cudaGetLastError();
        n = want;
        if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess)
            return fail(h, CT_E_NOMEM, "cudaMalloc of %zu bytes failed", n * sizeof(T));
    }
    if (zero) CT_CUDA(h, cudaMemsetAsync(p, 0, n * sizeof(T), h->stream));
    *ptr = static_cast<T *>(p);
    *have = n;
    return CT_OK;
}

int ensure_partials(ct_context *h, size_t doubles) { return grow(h, &h->partials, &h->partials_doubles, doubles, false); }

int ensure_scratch(ct_context *h, int pairs) {
    if (pairs <= h->scratch_pairs) return CT_OK;
    int cap = h->scratch_pairs * 2;
    if (cap < pairs) cap = pairs;
    if (cap < 64) cap = 64;
    size_t have_t = 0, have_x = 0, have_s = 0, have_st = 0;
    if (h->scratch_pairs) CT_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->tickets); cudaFree(h->xform); cudaFree(h->sums); cudaFree(h->status);
    h->tickets = nullptr; h->xform = nullptr; h->sums = nullptr; h->status = nullptr;
    h->scratch_pairs = 0;
    CT_TRY(grow(h, &h->tickets, &have_t, (size_t)cap, true));
    CT_TRY(grow(h, &h->xform, &have_x, (size_t)cap * CT_XFORM_DOUBLES, false));
    CT_TRY(grow(h, &h->sums, &have_s, (size_t)cap * 2 * CT_MOMENT_DOUBLES, false));
    CT_TRY(grow(h, &h->status, &have_st, (size_t)cap, true));
    // (the pinned bounce buffers of ct_host_copy.h do not depend on the pair count: they stay)
    if (h->host_status) cudaFreeHost(h->host_status);
    h->host_status = nullptr;
    CT_CUDA(h, cudaMallocHost(&h->host_status, sizeof(int) * (size_t)cap));
    h->scratch_pairs = cap;
    return CT_OK;
}

int ensure_ws(ct_context *h, size_t bytes) {
    unsigned char *p = static_cast<unsigned char *>(h->ws);
    CT_TRY(grow(h, &p, &h->ws_bytes, bytes, false));
    h->ws = p;
    return CT_OK;
}
int ensure_stage(ct_context *h, size_t bytes) {
    unsigned char *p = static_cast<unsigned char *>(h->stage);
    CT_TRY(grow(h, &p, &h->stage_bytes, bytes, false));
    h->stage = p;
    return CT_OK;
}

int ensure_seed(ct_context *h, size_t words) { return grow(h, &h->seed, &h->seed_words, words, false); }

// the float type the reference would see: a uint8 frame is what its loader decodes it to
static int decoded_dtype(const ct_batch *b) {
    if (b->dtype == CT_U8) return (b->flags & CT_BATCH_U8_AS_F32) ? CT_F32 : CT_F64;
    return b->dtype;
}

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct Carver {
    unsigned char *base;
    size_t off = 0;
    explicit Carver(void *b) : base(static_cast<unsigned char *>(b)) {}
    template <typename T> T *take(size_t n) {
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off = align_up(off + n * sizeof(T));
        return p;
    }
};

// ---------------------------------------------------------------------------------------------
// IDT fused driver
// ---------------------------------------------------------------------------------------------
struct IdtLayout {
    double *state;
    int64_t plane, state_stride;
    int64_t *keys;
    uint64_t *counts;
    double *lut;
    int32_t *status;
    size_t bytes;
};

static IdtLayout idt_layout(void *base, int64_t npix, int count, int bins, int n_iter) {
    IdtLayout L{};
    Carver c(base);
    L.plane = (npix + 3) / 4 * 4;   // 32-byte aligned planes: K7 writes 256-bit vectors when the source comes in groups of four pixels
    L.state_stride = 3 * L.plane;
    L.state = n_iter >= 2 ? c.take<double>((size_t)count * L.state_stride) : nullptr;
    L.keys = c.take<int64_t>((size_t)count * (n_iter + 1) * CT_IDT_KEYS);
    L.counts = c.take<uint64_t>((size_t)count * 6 * bins);
    L.lut = c.take<double>((size_t)count * CT_IDT_LUT_DOUBLES(bins));
    L.status = c.take<int32_t>((size_t)count);
    L.bytes = c.off;
    return L;
}

static int idt_run(ct_context *h, const ct_batch *target, const ct_batch *reference, const ct_batch *out,
                   const double *rotations, int bins, int n_iter, void *workspace, size_t workspace_bytes,
                   const ct_idt_trace *trace, int32_t *status) {
    CT_TRY(check_batch(h, target, "target"));
    CT_TRY(check_batch(h, reference, "reference"));
    CT_TRY(check_batch(h, out, "out"));
    if (!rotations) return fail(h, CT_E_INVALID, "rotations is NULL");
    if (n_iter < 1) return fail(h, CT_E_INVALID, "n_iter must be >= 1 (n_iter = 0 is a host-side copy)");
    if (bins < 1) return fail(h, CT_E_INVALID, "bins must be >= 1");
    if (bins > CT_IDT_MAX_BINS) return fail(h, CT_E_UNSUPPORTED, "bins=%d exceeds CT_IDT_MAX_BINS=%d", bins, CT_IDT_MAX_BINS);
    if (reference->count != target->count || out->count != target->count) return fail(h, CT_E_INVALID, "batch counts differ");
    if (out->npix != target->npix) return fail(h, CT_E_INVALID, "out.npix != target.npix");
    const bool out_f64 = out->dtype == CT_F64 && out->layout == CT_HWC;
    const bool out_fused = (out->dtype == CT_U8) || (out->dtype == CT_F32 && out->layout == CT_HWC);   // converted inside the last K7
    if (!out_f64 && !out_fused) return fail(h, CT_E_INVALID, "IDT output must be float64 CT_HWC, float32 CT_HWC or uint8");
    if (out_fused && n_iter < 2)
        return fail(h, CT_E_UNSUPPORTED, "uint8 / float32 IDT output needs n_iter >= 2 (it is written by the last iteration from the float64 state)");
    const int B = target->count;
    const size_t need = idt_layout(nullptr, target->npix, B, bins, n_iter).bytes;
    if (!workspace) {
        CT_TRY(ensure_ws(h, need));
        workspace = h->ws;
    } else if (workspace_bytes < need) {
        return fail(h, CT_E_NOMEM, "IDT workspace too small: %zu < %zu", workspace_bytes, need);
    }
    const IdtLayout L = idt_layout(workspace, target->npix, B, bins, n_iter);
    int32_t *st = status ? status : L.status;
    CT_CUDA(h, cudaMemsetAsync(L.counts, 0, sizeof(uint64_t) * (size_t)B * 6 * bins, h->stream));
    CT_CUDA(h, cudaMemsetAsync(st, 0, sizeof(int32_t) * (size_t)B, h->stream));
    const int64_t keys_stride = (int64_t)(n_iter + 1) * CT_IDT_KEYS, rot_stride = (int64_t)n_iter * 9;
    prof_mark(h, CT_PROF_START);
    // iteration 0 needs the target's range; the reference never changes, so its range under
    // EVERY rotation is taken in the same single pass over it.  One seed launch (which also sets the
    // keys to +inf) and one screened pass over both images.
    CT_TRY(launch_ranges_pair(h, target, reference, n_iter, rotations, rot_stride, L.keys, keys_stride, st, L.keys,
                              (int64_t)B * keys_stride));

    ct_batch state{};
    state.data = L.state;
    state.npix = target->npix;
    state.image_stride = L.state_stride;
    state.plane_stride = L.plane;
    state.count = B;
    state.dtype = CT_F64;
    state.layout = CT_CHW;
    for (int it = 0; it < n_iter; ++it) {
        const bool last = it == n_iter - 1;
        ct_idt_stage s{};
        s.target = it == 0 ? target : &state;
        s.reference = reference;
        s.rot = rotations + it * 9;
        s.rot_next = last ? nullptr : rotations + (it + 1) * 9;
        s.rot_stride = rot_stride;
        s.keys = L.keys + it * CT_IDT_KEYS;
        s.keys_next = last ? nullptr : L.keys + (it + 1) * CT_IDT_KEYS;
        s.keys_stride = keys_stride;
        s.counts = L.counts;
        s.lut = L.lut;
        s.status = st;
        s.bins = bins;
        CT_TRY(launch_hist(h, &s, 1, trace, it, n_iter));
        CT_TRY(launch_remap(h, &s, last ? out : &state, it == 0 && decoded_dtype(target) == CT_F32));
    }
    return CT_OK;
}

}  // namespace ct

using namespace ct;

// =============================================================================================
extern "C" {

int ct_abi_version(void) { return CT_ABI_VERSION; }

int ct_create(int device, ct_handle *out) {
    if (!out) return CT_E_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return CT_E_CUDA;
    ct_context *h = new (std::nothrow) ct_context();
    if (!h) return CT_E_NOMEM;
    h->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete h; return CT_E_CUDA; }
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (const char *e = getenv("CT_RANGES_BOUND")) h->ranges_bound = (float)atof(e);
    if (getenv("CT_RANGES_STATS") && cudaMalloc(&h->ranges_stats, 16) == cudaSuccess) cudaMemset(h->ranges_stats, 0, 16);
    if (cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return CT_E_CUDA;
    }
    *out = h;
    return CT_OK;
}

void ct_destroy(ct_handle h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->ranges_stats) {   // diagnostics: how many pixels the K4 screen sent to the exact path
        unsigned long long st[2] = {0, 0};
        cudaMemcpy(st, h->ranges_stats, 16, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[ct] ranges screen: %llu pixels exact, %llu flagged repeats skipped\n", st[0], st[1]);
        cudaFree(h->ranges_stats);
    }
    cudaFree(h->seed);
    cudaFree(h->partials);
    cudaFree(h->tickets);
    cudaFree(h->xform);
    cudaFree(h->sums);
    cudaFree(h->status);
    cudaFree(h->ws);
    cudaFree(h->stage);
    if (h->host_status) cudaFreeHost(h->host_status);
    for (int d = 0; d < 2; ++d)
        for (int i = 0; i < 2; ++i) {
            if (h->bounce[d][i]) cudaFreeHost(h->bounce[d][i]);
            if (h->bounce_done[d][i]) cudaEventDestroy(h->bounce_done[d][i]);
        }
    for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        if (h->side[i]) cudaStreamDestroy(h->side[i]);
        if (h->join[i]) cudaEventDestroy(h->join[i]);
    }
    if (h->fork) cudaEventDestroy(h->fork);
    if (h->copy_in) cudaStreamDestroy(h->copy_in);
    if (h->copy_out) cudaStreamDestroy(h->copy_out);
    delete h;
}

const char *ct_last_error(ct_handle h) { return h ? h->err : "null handle"; }

int ct_set_stream(ct_handle h, void *cuda_stream) {
    if (!h) return CT_E_INVALID;
    h->stream = static_cast<cudaStream_t>(cuda_stream);
    return CT_OK;
}

int ct_synchronize(ct_handle h) {
    if (!h) return CT_E_INVALID;
    CT_CUDA(h, cudaStreamSynchronize(h->stream));
    return CT_OK;
}

int ct_profile_enable(ct_handle h, int on) {
    if (!h) return CT_E_INVALID;
    h->prof_on = on != 0;
    h->prof_n = 0;
    return CT_OK;
}

int ct_profile_read(ct_handle h, int32_t *ids, float *ms, int32_t max) {
    if (!h || !ids || !ms || max < 0) return CT_E_INVALID;
    int n = 0;
    if (h->prof_n >= 2) {
        CT_CUDA(h, cudaEventSynchronize(h->prof_ev[h->prof_n - 1]));
        for (size_t i = 1; i < h->prof_n && n < max; ++i) {
            if (h->prof_id[i] == CT_PROF_START) continue;   // the gap between two driver calls
            float t = 0.0f;
            CT_CUDA(h, cudaEventElapsedTime(&t, h->prof_ev[i - 1], h->prof_ev[i]));
            ids[n] = h->prof_id[i];
            ms[n++] = t;
        }
    }
    h->prof_n = 0;
    return n;
}

int ct_sm_count(ct_handle h) { return h ? h->sm_count : 0; }
int64_t ct_launch_count(ct_handle h) { return h ? h->launches : 0; }

#define CT_ENTER(h)                             \
    if (!(h)) return CT_E_INVALID;              \
    (h)->err[0] = 0;                            \
    CT_CUDA((h), cudaSetDevice((h)->device))

// ------------------------------------------------------------------ linear
int ct_moments(ct_handle h, const ct_batch *images, int lab, double *sums) {
    CT_ENTER(h);
    return launch_moments(h, images, nullptr, lab, sums, -1, nullptr, nullptr);
}

int ct_linear_solve(ct_handle h, int method, const double *sums_t, const double *sums_r, int count,
                    double *xform, int *status) {
    CT_ENTER(h);
    return launch_solve(h, method, sums_t, sums_r, CT_MOMENT_DOUBLES, count, xform, status);
}

int ct_linear_apply(ct_handle h, int method, const ct_batch *target, const double *xform, const ct_batch *out) {
    CT_ENTER(h);
    if (method < CT_REINHARD || method > CT_MKL_CHOLESKY) return fail(h, CT_E_INVALID, "unknown method %d", method);
    return launch_apply(h, method, target, xform, out);
}

int ct_linear_transfer(ct_handle h, int method, const ct_batch *target, const ct_batch *reference,
                       const ct_batch *out, double *xform, int *status) {
    CT_ENTER(h);
    if (method < CT_REINHARD || method > CT_MKL_CHOLESKY) return fail(h, CT_E_INVALID, "unknown method %d", method);
    CT_TRY(check_batch(h, target, "target"));
    CT_TRY(ensure_scratch(h, target->count));
    if (!xform) xform = h->xform;
    if (!status) status = h->status;
    // Large batches run as chunks of ~32-100 Mpix alternating between two side streams (at least two
    // chunks of >= 16 Mpix).  Measured on 960x540 float32 pairs: MKL 0.80 -> 0.85 of the HBM roofline
    // at 64 pairs and 0.83 -> 0.90 at 1035 pairs (one chunk's read-only statistics pass overlaps the
    // other's write-heavy remap and its serial tail), Reinhard +4 % at 64 pairs.
    // CT_LINEAR_CHUNK_PAIRS overrides the chunk size for experiments (0 = off).
    static const int chunk_env = [] {
        const char *e = getenv("CT_LINEAR_CHUNK_PAIRS");
        return e ? atoi(e) : -1;
    }();
    int chunk = 0;
    if (chunk_env >= 0) {
        chunk = chunk_env;
    } else if ((int64_t)target->count * target->npix >= 32000000 && target->count >= 2) {
        // compute-bound Reinhard prefers ~100 Mpix chunks (0.69 -> 0.70 over 1035 pairs), the others ~32 Mpix
        const int64_t chunk_px = method == CT_REINHARD ? 100000000 : 32000000;
        const int64_t by_size = (chunk_px + target->npix / 2) / target->npix;
        chunk = (target->count + 1) / 2;
        if (by_size < chunk) chunk = (int)(by_size < 1 ? 1 : by_size);
    }
    if (chunk <= 0 || 2 * chunk > target->count) {
        CT_TRY(launch_moments(h, target, reference, method == CT_REINHARD, h->sums, method, xform, status));
        return launch_apply(h, method, target, xform, out);
    }
    // Chunks of pairs alternate between two side streams: statistics pass and remap of one chunk run
    // back to back (its target is still in L2 for the remap), and the serial tail of one chunk's
    // statistics pass (fixed-order combine + solve) overlaps the other stream's streaming.
    CT_TRY(check_batch(h, reference, "reference"));
    CT_TRY(check_batch(h, out, "out"));
    if (reference->count != target->count || out->count != target->count)
        return fail(h, CT_E_INVALID, "target/reference/out batch counts differ");
    if (!h->fork) {
        CT_CUDA(h, cudaEventCreateWithFlags(&h->fork, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) {
            CT_CUDA(h, cudaStreamCreateWithFlags(&h->side[i], cudaStreamNonBlocking));
            CT_CUDA(h, cudaEventCreateWithFlags(&h->join[i], cudaEventDisableTiming));
        }
    }
    // one partials region per side stream: blocks_for() never launches more than 8 waves of 4 CTAs per SM
    // for both images of a chunk, and never fewer than one CTA per image (tiny images, many pairs)
    size_t region = (size_t)h->sm_count * 4 * 8 * 9;
    if ((size_t)chunk * 2 * 9 > region) region = (size_t)chunk * 2 * 9;
    CT_TRY(ensure_partials(h, 2 * region));
    h->partials_region = region;
    CT_TRY(ensure_scratch(h, target->count > 2 * chunk ? target->count : 2 * chunk));
    cudaStream_t user = h->stream;
    CT_CUDA(h, cudaEventRecord(h->fork, user));
    for (int i = 0; i < 2; ++i) CT_CUDA(h, cudaStreamWaitEvent(h->side[i], h->fork, 0));
    const int64_t esz_t = elem_size(target->dtype), esz_r = elem_size(reference->dtype), esz_o = elem_size(out->dtype);
    int rc = CT_OK, k = 0;
    for (int b0 = 0; b0 < target->count && rc == CT_OK; b0 += chunk, ++k) {
        const int n = target->count - b0 < chunk ? target->count - b0 : chunk;
        ct_batch t = *target, r = *reference, o = *out;
        t.data = (char *)target->data + (int64_t)b0 * target->image_stride * esz_t;
        r.data = (char *)reference->data + (int64_t)b0 * reference->image_stride * esz_r;
        o.data = (char *)out->data + (int64_t)b0 * out->image_stride * esz_o;
        t.count = r.count = o.count = n;
        h->stream = h->side[k & 1];
        h->ticket_base = (k & 1) * chunk;
        h->partials_base = (k & 1) * region;
        rc = launch_moments(h, &t, &r, method == CT_REINHARD, h->sums + (size_t)b0 * 2 * CT_MOMENT_DOUBLES, method,
                            xform + (size_t)b0 * CT_XFORM_DOUBLES, status + b0);
        if (rc == CT_OK) rc = launch_apply(h, method, &t, xform + (size_t)b0 * CT_XFORM_DOUBLES, &o);
    }
    h->stream = user;
    h->ticket_base = 0;
    h->partials_base = 0;
    h->partials_region = 0;
    for (int i = 0; i < 2; ++i) {
        CT_CUDA(h, cudaEventRecord(h->join[i], h->side[i]));
        CT_CUDA(h, cudaStreamWaitEvent(user, h->join[i], 0));
    }
    return rc;
}

// ------------------------------------------------------------------ quality metrics
static int metric_result(ct_context *h, double *result) {
    CT_CUDA(h, cudaMemcpyAsync(result, h->sums, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CT_CUDA(h, cudaStreamSynchronize(h->stream));
    return CT_OK;
}

int ct_icid(ct_handle h, const float *img1, const float *img2, int32_t count, int32_t height, int32_t width,
            int32_t intent, int32_t omit_maps67, int32_t downsampling, double *result) {
    CT_ENTER(h);
    if (!result) return fail(h, CT_E_INVALID, "result is NULL");
    CT_TRY(ensure_scratch(h, 1));
    CT_TRY(launch_icid(h, img1, img2, count, height, width, intent, omit_maps67, downsampling, h->sums));
    return metric_result(h, result);
}

int ct_psnr(ct_handle h, const float *x, const float *y, int32_t count, int64_t elems_per_image, double *result) {
    CT_ENTER(h);
    if (!result) return fail(h, CT_E_INVALID, "result is NULL");
    CT_TRY(ensure_scratch(h, 1));
    CT_TRY(launch_psnr(h, x, y, count, elems_per_image, h->sums));
    return metric_result(h, result);
}

int ct_distort(ct_handle h, const ct_batch *src, const ct_distortion *ops, int32_t n_ops, const ct_batch *dst) {
    CT_ENTER(h);
    return launch_distort(h, src, ops, n_ops, dst);
}

int ct_ssim(ct_handle h, const float *x, const float *y, int32_t count, int32_t height, int32_t width,
            int32_t downsample, double *result) {
    CT_ENTER(h);
    if (!result) return fail(h, CT_E_INVALID, "result is NULL");
    CT_TRY(ensure_scratch(h, 1));
    CT_TRY(launch_ssim(h, x, y, count, height, width, downsample, h->sums));
    return metric_result(h, result);
}

}  // extern "C"

// Host pipeline shared by the two *_host entry points: per pair H2D on copy_in, kernels on the
// handle's stream, D2H on copy_out, two slots in flight.
namespace {
constexpr int kSlots = 3;   // pairs in flight: H2D of pair b+1 / b+2 and D2H of pair b-1 overlap the kernels of pair b
struct Pipeline {
    cudaEvent_t in_ready[kSlots], done[kSlots], out_free[kSlots];
    bool ok = false;
    int init() {
        for (int i = 0; i < kSlots; ++i) {
            if (cudaEventCreateWithFlags(&in_ready[i], cudaEventDisableTiming) != cudaSuccess) return -1;
            if (cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) != cudaSuccess) return -1;
            if (cudaEventCreateWithFlags(&out_free[i], cudaEventDisableTiming) != cudaSuccess) return -1;
        }
        ok = true;
        return 0;
    }
    ~Pipeline() {
        if (!ok) return;
        for (int i = 0; i < kSlots; ++i) {
            cudaEventDestroy(in_ready[i]);
            cudaEventDestroy(done[i]);
            cudaEventDestroy(out_free[i]);
        }
    }
};

ct_batch single(const ct_batch *b, void *data) {
    ct_batch s = *b;
    s.data = data;
    s.count = 1;
    s.image_stride = 0;
    return s;
}
void *host_image(const ct_batch *b, int i) {
    return static_cast<unsigned char *>(b->data) + (size_t)i * (size_t)b->image_stride * elem_size(b->dtype);
}
size_t image_bytes(const ct_batch *b) {
    // CHW with padded planes is copied as one block including the padding
    const int64_t elems = b->layout == CT_CHW ? 2 * plane_of(b) + b->npix : 3 * b->npix;
    return (size_t)elems * elem_size(b->dtype);
}

template <typename Launch>
int run_host_pipeline(ct_context *h, const ct_batch *target, const ct_batch *reference, const ct_batch *out,
                      size_t extra_ws, Launch &&launch) {
    const int B = target->count;
    const size_t tb = align_up(image_bytes(target)), rb = align_up(image_bytes(reference)), ob = align_up(image_bytes(out));
    const int slots = B > 1 ? (B < kSlots ? B : kSlots) : 1;
    CT_TRY(ensure_stage(h, (tb + rb + ob) * slots + extra_ws));
    Pipeline pl;
    if (pl.init() != 0) return fail(h, CT_E_CUDA, "event creation failed");
    unsigned char *base = static_cast<unsigned char *>(h->stage);
    // A single pair in pageable memory (the drop-in numpy call): optional bounce-buffered parallel
    // copies (CT_STAGED_COPY=1).  Off by default: 3.3-4.5 ms instead of 4.2-4.6 ms per 0964-size call
    // on idle host cores, but 6 ms when other threads of the process are busy (e.g. BLAS workers
    // still spinning after a numpy call) - the driver's own pageable path does not depend on that.
    static const bool staged = getenv("CT_STAGED_COPY") != nullptr && atoi(getenv("CT_STAGED_COPY")) != 0;
    const size_t big = 2u << 20;
    if (staged && B == 1 && image_bytes(target) >= big && is_pageable(target->data) && is_pageable(reference->data) &&
        is_pageable(out->data)) {
        unsigned char *dt = base, *dr = dt + tb, *dout = dr + rb;
        static const bool prof = getenv("CT_PROFILE_HOST") != nullptr;
        auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        const double t0 = now();
        CT_TRY(staged_h2d(h, dt, target->data, image_bytes(target), h->stream));
        CT_TRY(staged_h2d(h, dr, reference->data, image_bytes(reference), h->stream));
        const double t1_ = now();
        const ct_batch t1 = single(target, dt), r1 = single(reference, dr), o1 = single(out, dout);
        CT_TRY(launch(0, &t1, &r1, &o1));
        const double t2 = now();
        CT_TRY(staged_d2h(h, out->data, dout, image_bytes(out), h->stream));
        CT_CUDA(h, cudaStreamSynchronize(h->stream));
        if (prof) fprintf(stderr, "[ct host] h2d %.3f ms, launch %.3f ms, d2h(+kernels) %.3f ms\n", t1_ - t0, t2 - t1_, now() - t2);
        return CT_OK;
    }
    for (int b = 0; b < B; ++b) {
        const int s = b % slots;
        unsigned char *dt = base + (size_t)s * (tb + rb + ob), *dr = dt + tb, *dout = dr + rb;
        if (b >= slots) CT_CUDA(h, cudaStreamWaitEvent(h->copy_in, pl.done[s], 0));
        CT_CUDA(h, cudaMemcpyAsync(dt, host_image(target, b), image_bytes(target), cudaMemcpyHostToDevice, h->copy_in));
        CT_CUDA(h, cudaMemcpyAsync(dr, host_image(reference, b), image_bytes(reference), cudaMemcpyHostToDevice, h->copy_in));
        CT_CUDA(h, cudaEventRecord(pl.in_ready[s], h->copy_in));
        CT_CUDA(h, cudaStreamWaitEvent(h->stream, pl.in_ready[s], 0));
        if (b >= slots) CT_CUDA(h, cudaStreamWaitEvent(h->stream, pl.out_free[s], 0));
        const ct_batch t1 = single(target, dt), r1 = single(reference, dr), o1 = single(out, dout);
        CT_TRY(launch(b, &t1, &r1, &o1));
        CT_CUDA(h, cudaEventRecord(pl.done[s], h->stream));
        CT_CUDA(h, cudaStreamWaitEvent(h->copy_out, pl.done[s], 0));
        CT_CUDA(h, cudaMemcpyAsync(host_image(out, b), dout, image_bytes(out), cudaMemcpyDeviceToHost, h->copy_out));
        CT_CUDA(h, cudaEventRecord(pl.out_free[s], h->copy_out));
    }
    CT_CUDA(h, cudaStreamSynchronize(h->copy_out));
    CT_CUDA(h, cudaStreamSynchronize(h->stream));
    return CT_OK;
}

// uint8 frames: [H2D u8 pair] -> launch(u8 pair -> u8 out) -> [D2H u8] with the same three-stream overlap.
// The kernels decode the frames as they read them and encode the result as they write it
// (CT_U8 batches): 3 bytes per pixel cross PCIe each way and there is no conversion pass.
template <typename Launch>
int run_host_pipeline_u8(ct_context *h, const uint8_t *target, const uint8_t *reference, uint8_t *out, int count,
                         int64_t npix_t, int64_t npix_r, int as_float32, Launch &&launch) {
    const size_t nt = (size_t)npix_t * 3, nr = (size_t)npix_r * 3;
    const size_t ut = align_up(nt), ur = align_up(nr), uo = align_up(nt);
    const size_t slot = ut + ur + uo;
    const int slots = count > 1 ? (count < kSlots ? count : kSlots) : 1;
    CT_TRY(ensure_stage(h, slot * slots));
    Pipeline pl;
    if (pl.init() != 0) return fail(h, CT_E_CUDA, "event creation failed");
    unsigned char *base = static_cast<unsigned char *>(h->stage);
    const int flags = as_float32 ? CT_BATCH_U8_AS_F32 : 0;
    for (int b = 0; b < count; ++b) {
        const int s = b % slots;
        unsigned char *p = base + (size_t)s * slot;
        uint8_t *d_ut = p, *d_ur = p + ut, *d_uo = p + ut + ur;
        if (b >= slots) CT_CUDA(h, cudaStreamWaitEvent(h->copy_in, pl.done[s], 0));
        CT_CUDA(h, cudaMemcpyAsync(d_ut, target + (size_t)b * nt, nt, cudaMemcpyHostToDevice, h->copy_in));
        CT_CUDA(h, cudaMemcpyAsync(d_ur, reference + (size_t)b * nr, nr, cudaMemcpyHostToDevice, h->copy_in));
        CT_CUDA(h, cudaEventRecord(pl.in_ready[s], h->copy_in));
        CT_CUDA(h, cudaStreamWaitEvent(h->stream, pl.in_ready[s], 0));
        if (b >= slots) CT_CUDA(h, cudaStreamWaitEvent(h->stream, pl.out_free[s], 0));
        ct_batch t1{d_ut, npix_t, 0, 0, 1, CT_U8, CT_HWC, flags}, r1{d_ur, npix_r, 0, 0, 1, CT_U8, CT_HWC, flags};
        ct_batch o1{d_uo, npix_t, 0, 0, 1, CT_U8, CT_HWC, 0};
        CT_TRY(launch(b, &t1, &r1, &o1));
        CT_CUDA(h, cudaEventRecord(pl.done[s], h->stream));
        CT_CUDA(h, cudaStreamWaitEvent(h->copy_out, pl.done[s], 0));
        CT_CUDA(h, cudaMemcpyAsync(out + (size_t)b * nt, d_uo, nt, cudaMemcpyDeviceToHost, h->copy_out));
        CT_CUDA(h, cudaEventRecord(pl.out_free[s], h->copy_out));
    }
    CT_CUDA(h, cudaStreamSynchronize(h->copy_out));
    CT_CUDA(h, cudaStreamSynchronize(h->stream));
    return CT_OK;
}

int first_bad_status(ct_context *h, const int *dev_status, int B) {
    if (cudaMemcpyAsync(h->host_status, dev_status, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
        cudaStreamSynchronize(h->stream) != cudaSuccess)
        return fail(h, CT_E_CUDA, "status read-back failed");
    for (int b = 0; b < B; ++b)
        if (h->host_status[b] != CT_OK) {
            const int st = h->host_status[b];
            const char *what = st == CT_E_NONFINITE ? "projected range is not finite"
                               : st == CT_E_NOT_PD  ? "Matrix is not positive definite"
                               : st == CT_E_SINGULAR ? "Singular matrix"
                                                     : "kernel reported an error";
            return fail(h, st, "pair %d: %s", b, what);
        }
    return CT_OK;
}
}  // namespace

extern "C" {

int ct_linear_transfer_host(ct_handle h, int method, const ct_batch *target, const ct_batch *reference,
                            const ct_batch *out) {
    CT_ENTER(h);
    if (method < CT_REINHARD || method > CT_MKL_CHOLESKY) return fail(h, CT_E_INVALID, "unknown method %d", method);
    CT_TRY(check_batch(h, target, "target"));
    CT_TRY(check_batch(h, reference, "reference"));
    CT_TRY(check_batch(h, out, "out"));
    if (reference->count != target->count || out->count != target->count) return fail(h, CT_E_INVALID, "batch counts differ");
    CT_TRY(ensure_scratch(h, target->count));
    CT_CUDA(h, cudaMemsetAsync(h->status, 0, sizeof(int) * (size_t)target->count, h->stream));
    CT_TRY(run_host_pipeline(h, target, reference, out, 0,
                             [&](int b, const ct_batch *t, const ct_batch *r, const ct_batch *o) {
                                 CT_TRY(launch_moments(h, t, r, method == CT_REINHARD, h->sums + (size_t)b * 2 * CT_MOMENT_DOUBLES,
                                                       method, h->xform + (size_t)b * CT_XFORM_DOUBLES, h->status + b));
                                 return launch_apply(h, method, t, h->xform + (size_t)b * CT_XFORM_DOUBLES, o);
                             }));
    return first_bad_status(h, h->status, target->count);
}

int ct_linear_stats_host(ct_handle h, int lab, const ct_batch *target, const ct_batch *reference, double *sums_t,
                         double *sums_r) {
    CT_ENTER(h);
    h->staged_valid = false;
    CT_TRY(check_batch(h, target, "target"));
    CT_TRY(check_batch(h, reference, "reference"));
    if (!sums_t || !sums_r) return fail(h, CT_E_INVALID, "sums is NULL");
    if (target->count != 1 || reference->count != 1) return fail(h, CT_E_UNSUPPORTED, "the two-phase transfer takes one pair per call");
    CT_TRY(ensure_scratch(h, 1));
    const size_t tb = align_up(image_bytes(target)), rb = align_up(image_bytes(reference));
    const size_t ob = align_up(sizeof(double) * 3 * (size_t)target->npix);   // room for a float64 result
    CT_TRY(ensure_stage(h, tb + rb + ob));
    unsigned char *base = static_cast<unsigned char *>(h->stage);
    CT_CUDA(h, cudaMemcpyAsync(base, target->data, image_bytes(target), cudaMemcpyHostToDevice, h->stream));
    CT_CUDA(h, cudaMemcpyAsync(base + tb, reference->data, image_bytes(reference), cudaMemcpyHostToDevice, h->stream));
    const ct_batch t1 = single(target, base), r1 = single(reference, base + tb);
    CT_TRY(launch_moments(h, &t1, &r1, lab, h->sums, -1, nullptr, nullptr));
    CT_CUDA(h, cudaMemcpyAsync(sums_t, h->sums, sizeof(double) * CT_MOMENT_DOUBLES, cudaMemcpyDeviceToHost, h->stream));
    CT_CUDA(h, cudaMemcpyAsync(sums_r, h->sums + CT_MOMENT_DOUBLES, sizeof(double) * CT_MOMENT_DOUBLES, cudaMemcpyDeviceToHost, h->stream));
    CT_CUDA(h, cudaStreamSynchronize(h->stream));
    h->staged_target = t1;
    h->staged_out = base + tb + rb;
    h->staged_valid = true;
    return CT_OK;
}

int ct_linear_apply_staged_host(ct_handle h, int method, const double *xform, const ct_batch *out) {
    CT_ENTER(h);
    if (method < CT_REINHARD || method > CT_MKL_CHOLESKY) return fail(h, CT_E_INVALID, "unknown method %d", method);
    if (!h->staged_valid) return fail(h, CT_E_INVALID, "no staged pair: call ct_linear_stats_host first");
    h->staged_valid = false;
    CT_TRY(check_batch(h, out, "out"));
    if (!xform) return fail(h, CT_E_INVALID, "xform is NULL");
    if (out->count != 1 || out->npix != h->staged_target.npix || out->layout != CT_HWC)
        return fail(h, CT_E_INVALID, "out must be one CT_HWC image of the staged target's size");
    CT_CUDA(h, cudaMemcpyAsync(h->xform, xform, sizeof(double) * CT_XFORM_DOUBLES, cudaMemcpyHostToDevice, h->stream));
    const ct_batch o1 = single(out, h->staged_out);
    CT_TRY(launch_apply(h, method, &h->staged_target, h->xform, &o1));
    CT_CUDA(h, cudaMemcpyAsync(out->data, h->staged_out, image_bytes(out), cudaMemcpyDeviceToHost, h->stream));
    CT_CUDA(h, cudaStreamSynchronize(h->stream));
    return CT_OK;
}

int ct_linear_transfer_host_u8(ct_handle h, int method, const uint8_t *target, const uint8_t *reference, uint8_t *out,
                               int32_t count, int64_t npix_target, int64_t npix_reference, int32_t as_float32) {
    CT_ENTER(h);
    if (method < CT_REINHARD || method > CT_MKL_CHOLESKY) return fail(h, CT_E_INVALID, "unknown method %d", method);
    if (!target || !reference || !out || count <= 0 || npix_target <= 0 || npix_reference <= 0)
        return fail(h, CT_E_INVALID, "bad uint8 batch arguments");
    CT_TRY(ensure_scratch(h, count));
    CT_CUDA(h, cudaMemsetAsync(h->status, 0, sizeof(int) * (size_t)count, h->stream));
    CT_TRY(run_host_pipeline_u8(h, target, reference, out, count, npix_target, npix_reference, as_float32,
                                [&](int b, const ct_batch *t, const ct_batch *r, const ct_batch *o) {
                                    CT_TRY(launch_moments(h, t, r, method == CT_REINHARD, h->sums + (size_t)b * 2 * CT_MOMENT_DOUBLES,
                                                          method, h->xform + (size_t)b * CT_XFORM_DOUBLES, h->status + b));
                                    return launch_apply(h, method, t, h->xform + (size_t)b * CT_XFORM_DOUBLES, o);
                                }));
    return first_bad_status(h, h->status, count);
}

int ct_idt_transfer_host_u8(ct_handle h, const uint8_t *target, const uint8_t *reference, uint8_t *out, int32_t count,
                            int64_t npix_target, int64_t npix_reference, int32_t as_float32, const double *rotations,
                            int32_t bins, int32_t n_iter) {
    CT_ENTER(h);
    if (!target || !reference || !out || !rotations || count <= 0 || npix_target <= 0 || npix_reference <= 0)
        return fail(h, CT_E_INVALID, "bad uint8 batch arguments");
    if (n_iter < 1 || bins < 1) return fail(h, CT_E_INVALID, "n_iter and bins must be >= 1");
    if (bins > CT_IDT_MAX_BINS) return fail(h, CT_E_UNSUPPORTED, "bins=%d exceeds CT_IDT_MAX_BINS=%d", bins, CT_IDT_MAX_BINS);
    CT_TRY(ensure_scratch(h, count));
    const size_t rot_bytes = align_up(sizeof(double) * (size_t)count * n_iter * 9);
    const size_t idt_ws = idt_layout(nullptr, npix_target, 1, bins, n_iter).bytes;
    const size_t tmp_bytes = n_iter < 2 ? align_up(sizeof(double) * 3 * (size_t)npix_target) : 0;
    CT_TRY(ensure_ws(h, rot_bytes + idt_ws + tmp_bytes));
    double *d_rot = static_cast<double *>(h->ws);
    void *idt_base = static_cast<unsigned char *>(h->ws) + rot_bytes;
    h->u8_tmp = reinterpret_cast<double *>(static_cast<unsigned char *>(h->ws) + rot_bytes + idt_ws);
    CT_CUDA(h, cudaMemcpyAsync(d_rot, rotations, sizeof(double) * (size_t)count * n_iter * 9, cudaMemcpyHostToDevice, h->stream));
    CT_CUDA(h, cudaMemsetAsync(h->status, 0, sizeof(int) * (size_t)count, h->stream));
    CT_TRY(run_host_pipeline_u8(h, target, reference, out, count, npix_target, npix_reference, as_float32,
                                [&](int b, const ct_batch *t, const ct_batch *r, const ct_batch *o) {
                                    if (n_iter >= 2)
                                        return idt_run(h, t, r, o, d_rot + (size_t)b * n_iter * 9, bins, n_iter, idt_base, idt_ws,
                                                       nullptr, h->status + b);
                                    // a single iteration has no float64 state to convert from: float64 result, then encode
                                    ct_batch f{h->u8_tmp, t->npix, 0, 0, 1, CT_F64, CT_HWC, 0};
                                    CT_TRY(idt_run(h, t, r, &f, d_rot + (size_t)b * n_iter * 9, bins, n_iter, idt_base, idt_ws,
                                                   nullptr, h->status + b));
                                    return launch_float_to_u8(h, h->u8_tmp, CT_F64, static_cast<uint8_t *>(o->data), 3 * t->npix);
                                }));
    return first_bad_status(h, h->status, count);
}

// ------------------------------------------------------------------ regrain / automated colour grading
size_t ct_regrain_workspace_bytes(int32_t height, int32_t width) {
    if (height <= 0 || width <= 0) return 0;
    return regrain_workspace_bytes(height, width);
}

int ct_regrain(ct_handle h, const double *in, const double *col, double *out, int32_t height, int32_t width,
               void *workspace, size_t workspace_bytes) {
    CT_ENTER(h);
    if (!workspace) {
        const size_t need = regrain_workspace_bytes(height, width);
        CT_TRY(ensure_ws(h, need));
        workspace = h->ws;
        workspace_bytes = h->ws_bytes;
    }
    return launch_regrain(h, in, col, out, height, width, workspace, workspace_bytes);
}

int ct_acg_transfer_host(ct_handle h, const ct_batch *target, const ct_batch *reference, const ct_batch *out,
                         int32_t height, int32_t width, const double *rotations, int32_t bins, int32_t n_iter) {
    CT_ENTER(h);
    CT_TRY(check_batch(h, target, "target"));
    CT_TRY(check_batch(h, reference, "reference"));
    CT_TRY(check_batch(h, out, "out"));
    if (target->count != 1 || reference->count != 1 || out->count != 1) return fail(h, CT_E_UNSUPPORTED, "automated colour grading takes one pair per call");
    if ((int64_t)height * width != target->npix || out->npix != target->npix) return fail(h, CT_E_INVALID, "height * width must equal target.npix and out.npix");
    if (out->dtype != CT_F64 || out->layout != CT_HWC) return fail(h, CT_E_INVALID, "output must be float64 CT_HWC");
    if (!rotations) return fail(h, CT_E_INVALID, "rotations is NULL");
    if (n_iter < 1 || bins < 1) return fail(h, CT_E_INVALID, "n_iter and bins must be >= 1");
    if (bins > CT_IDT_MAX_BINS) return fail(h, CT_E_UNSUPPORTED, "bins=%d exceeds CT_IDT_MAX_BINS=%d", bins, CT_IDT_MAX_BINS);
    CT_TRY(ensure_scratch(h, 1));
    const size_t tb = align_up(image_bytes(target)), rb = align_up(image_bytes(reference));
    const size_t fb = align_up(sizeof(double) * 3 * (size_t)target->npix);
    CT_TRY(ensure_stage(h, tb + rb + 3 * fb));
    unsigned char *st = static_cast<unsigned char *>(h->stage);
    unsigned char *d_t = st, *d_r = st + tb;
    double *d_col = reinterpret_cast<double *>(st + tb + rb), *d_in = d_col + fb / 8, *d_res = d_in + fb / 8;
    const size_t rot_bytes = align_up(sizeof(double) * (size_t)n_iter * 9);
    const size_t idt_ws = idt_layout(nullptr, target->npix, 1, bins, n_iter).bytes;
    const size_t rg_ws = regrain_workspace_bytes(height, width);
    const size_t ws_need = rot_bytes + (idt_ws > rg_ws ? idt_ws : rg_ws);
    CT_TRY(ensure_ws(h, ws_need));
    double *d_rot = static_cast<double *>(h->ws);
    void *ws = static_cast<unsigned char *>(h->ws) + rot_bytes;
    CT_CUDA(h, cudaMemcpyAsync(d_t, target->data, image_bytes(target), cudaMemcpyHostToDevice, h->stream));
    CT_CUDA(h, cudaMemcpyAsync(d_r, reference->data, image_bytes(reference), cudaMemcpyHostToDevice, h->stream));
    CT_CUDA(h, cudaMemcpyAsync(d_rot, rotations, sizeof(double) * (size_t)n_iter * 9, cudaMemcpyHostToDevice, h->stream));
    CT_CUDA(h, cudaMemsetAsync(h->status, 0, sizeof(int), h->stream));
    const ct_batch t1 = single(target, d_t), r1 = single(reference, d_r);
    ct_batch c1{d_col, target->npix, 0, 0, 1, CT_F64, CT_HWC, 0};
    CT_TRY(idt_run(h, &t1, &r1, &c1, d_rot, bins, n_iter, ws, ws_need - rot_bytes, nullptr, h->status));   // iterative.py:135
    CT_TRY(launch_to_f64_hwc(h, &t1, d_in));
    CT_TRY(launch_regrain(h, d_in, d_col, d_res, height, width, ws, ws_need - rot_bytes));                  // iterative.py:136
    CT_CUDA(h, cudaMemcpyAsync(out->data, d_res, sizeof(double) * 3 * (size_t)target->npix, cudaMemcpyDeviceToHost, h->stream));
    return first_bad_status(h, h->status, 1);
}

// ------------------------------------------------------------------ IDT
int64_t ct_idt_key_of(double value) { return key_of(value); }
double ct_idt_value_of(int64_t key) { return value_of(key); }

int ct_idt_keys_init(ct_handle h, int64_t *keys, int64_t n) {
    CT_ENTER(h);
    return launch_keys_init(h, keys, n);
}
int ct_idt_ranges(ct_handle h, const ct_batch *images, const double *rot, int64_t rot_stride, int32_t n_rot,
                  int64_t *keys, int64_t keys_stride, int32_t *status) {
    CT_ENTER(h);
    return launch_ranges(h, images, rot, rot_stride, n_rot, keys, keys_stride, status);
}
int ct_idt_hist(ct_handle h, const ct_idt_stage *s, int fuse_lut) {
    CT_ENTER(h);
    return launch_hist(h, s, fuse_lut, nullptr, 0, 1);
}
int ct_idt_lut(ct_handle h, const ct_idt_stage *s, int keep_counts) {
    CT_ENTER(h);
    return launch_lut(h, s, keep_counts, nullptr, 0, 1);
}
int ct_idt_remap(ct_handle h, const ct_idt_stage *s, const ct_batch *dst, int round_f32) {
    CT_ENTER(h);
    return launch_remap(h, s, dst, round_f32);
}

size_t ct_idt_workspace_bytes(int64_t npix_target, int32_t count, int32_t bins, int32_t n_iter) {
    if (npix_target <= 0 || count <= 0 || bins <= 0 || n_iter <= 0) return 0;
    return idt_layout(nullptr, npix_target, count, bins, n_iter).bytes;
}

int ct_idt_transfer(ct_handle h, const ct_batch *target, const ct_batch *reference, const ct_batch *out,
                    const double *rotations, int32_t bins, int32_t n_iter, void *workspace,
                    size_t workspace_bytes, const ct_idt_trace *trace, int32_t *status) {
    CT_ENTER(h);
    return idt_run(h, target, reference, out, rotations, bins, n_iter, workspace, workspace_bytes, trace, status);
}

int ct_idt_transfer_host(ct_handle h, const ct_batch *target, const ct_batch *reference, const ct_batch *out,
                         const double *rotations, int32_t bins, int32_t n_iter, const ct_idt_trace *trace) {
    CT_ENTER(h);
    CT_TRY(check_batch(h, target, "target"));
    CT_TRY(check_batch(h, reference, "reference"));
    CT_TRY(check_batch(h, out, "out"));
    if (!rotations) return fail(h, CT_E_INVALID, "rotations is NULL");
    if (n_iter < 1 || bins < 1) return fail(h, CT_E_INVALID, "n_iter and bins must be >= 1");
    if (bins > CT_IDT_MAX_BINS) return fail(h, CT_E_UNSUPPORTED, "bins=%d exceeds CT_IDT_MAX_BINS=%d", bins, CT_IDT_MAX_BINS);
    if (reference->count != target->count || out->count != target->count) return fail(h, CT_E_INVALID, "batch counts differ");
    const int B = target->count;
    CT_TRY(ensure_scratch(h, B));
    // device-side small buffers: rotations, per-pair status, optional trace
    const size_t per_axis = (size_t)B * n_iter * 3;
    Carver probe(nullptr);
    probe.take<double>((size_t)B * n_iter * 9);
    if (trace) {
        probe.take<double>(per_axis); probe.take<double>(per_axis);
        probe.take<int64_t>(per_axis * bins); probe.take<int64_t>(per_axis * bins); probe.take<double>(per_axis * bins);
    }
    const size_t small_bytes = probe.off;
    const size_t idt_ws = idt_layout(nullptr, target->npix, 1, bins, n_iter).bytes;
    CT_TRY(ensure_ws(h, small_bytes + idt_ws));
    Carver c(h->ws);
    double *d_rot = c.take<double>((size_t)B * n_iter * 9);
    ct_idt_trace dtr{};
    if (trace) {
        dtr.lo = c.take<double>(per_axis); dtr.hi = c.take<double>(per_axis);
        dtr.counts_t = c.take<int64_t>(per_axis * bins); dtr.counts_r = c.take<int64_t>(per_axis * bins);
        dtr.lut = c.take<double>(per_axis * bins);
    }
    void *idt_base = static_cast<unsigned char *>(h->ws) + small_bytes;
    CT_CUDA(h, cudaMemcpyAsync(d_rot, rotations, sizeof(double) * (size_t)B * n_iter * 9, cudaMemcpyHostToDevice, h->stream));
    CT_CUDA(h, cudaMemsetAsync(h->status, 0, sizeof(int) * (size_t)B, h->stream));
    CT_TRY(run_host_pipeline(h, target, reference, out, 0,
                             [&](int b, const ct_batch *t, const ct_batch *r, const ct_batch *o) {
                                 ct_idt_trace tb = dtr;
                                 if (trace) {
                                     const size_t o3 = (size_t)b * n_iter * 3;
                                     tb.lo += o3; tb.hi += o3;
                                     tb.counts_t += o3 * bins; tb.counts_r += o3 * bins; tb.lut += o3 * bins;
                                 }
                                 return idt_run(h, t, r, o, d_rot + (size_t)b * n_iter * 9, bins, n_iter, idt_base, idt_ws,
                                                trace ? &tb : nullptr, h->status + b);
                             }));
    if (trace) {
        if (trace->lo) CT_CUDA(h, cudaMemcpyAsync(trace->lo, dtr.lo, sizeof(double) * per_axis, cudaMemcpyDeviceToHost, h->stream));
        if (trace->hi) CT_CUDA(h, cudaMemcpyAsync(trace->hi, dtr.hi, sizeof(double) * per_axis, cudaMemcpyDeviceToHost, h->stream));
        if (trace->counts_t) CT_CUDA(h, cudaMemcpyAsync(trace->counts_t, dtr.counts_t, sizeof(int64_t) * per_axis * bins, cudaMemcpyDeviceToHost, h->stream));
        if (trace->counts_r) CT_CUDA(h, cudaMemcpyAsync(trace->counts_r, dtr.counts_r, sizeof(int64_t) * per_axis * bins, cudaMemcpyDeviceToHost, h->stream));
        if (trace->lut) CT_CUDA(h, cudaMemcpyAsync(trace->lut, dtr.lut, sizeof(double) * per_axis * bins, cudaMemcpyDeviceToHost, h->stream));
    }
    return first_bad_status(h, h->status, B);
}

}  // extern "C"
