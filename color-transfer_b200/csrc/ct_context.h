// Host-side context behind ct_handle and the internal launcher prototypes.
#pragma once

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "ct_common.cuh"

struct ct_context {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    char err[512] = {0};
    int64_t launches = 0;

    // moments scratch: per-block partial sums + one ticket per pair
    double *partials = nullptr;
    size_t partials_doubles = 0;
    unsigned int *tickets = nullptr;
    // per-pair transform / status / raw sums when the caller passes NULL
    double *xform = nullptr;
    double *sums = nullptr;
    int *status = nullptr;
    int scratch_pairs = 0;
    int *host_status = nullptr;  // pinned

    // grow-only device workspace (IDT state, staging of host images)
    void *ws = nullptr;
    size_t ws_bytes = 0;
    void *stage = nullptr;
    size_t stage_bytes = 0;
    // optional per-launch timing of the fused IDT driver (ct_profile_enable / ct_profile_read)
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;
    std::vector<int> prof_id;
    size_t prof_n = 0;

    // ct_linear_stats_host -> ct_linear_apply_staged_host: the target staged on the device, where its result goes
    ct_batch staged_target = {};
    void *staged_out = nullptr;
    bool staged_valid = false;

    double *u8_tmp = nullptr;   // inside ws: float64 result of a one-iteration IDT on uint8 frames before it is encoded

    // K4 screen: pixels with |x0|+|x1|+|x2| above this take the exact fp64 path (CT_RANGES_BOUND)
    float ranges_bound = 16.0f;
    long long *seed = nullptr;                    // K4a -> K4b: subsample extremes [pairs][2][kSeedCtas][24]
    size_t seed_words = 0;
    unsigned long long *ranges_stats = nullptr;   // device [2], allocated when CT_RANGES_STATS is set

    bool hist_smem_raised = false;
    bool remap_smem_raised[12] = {false, false, false, false, false, false, false, false, false, false, false, false};

    // host pipeline
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    // pinned bounce buffers for pageable caller memory (ct_host_copy.h): 2 chunks per direction
    unsigned char *bounce[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    cudaEvent_t bounce_done[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};

    // batched linear transfers: chunks of pairs alternate between two side streams so that the
    // serial tail of one chunk's statistics pass overlaps the next chunk's streaming
    cudaStream_t side[2] = {nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
    int ticket_base = 0;        // first ticket and first partials double a moments launch may use
    size_t partials_base = 0;   // (one region per side stream)
    size_t partials_region = 0; // doubles per region while a chunked batch is in flight, else 0
};

namespace ct {

inline int fail(ct_context *h, int code, const char *fmt, ...) {
    if (h) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(h->err, sizeof(h->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

#define CT_CUDA(h, expr)                                                                    \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess)                                                             \
            return ct::fail((h), CT_E_CUDA, "%s failed: %s (%s:%d)", #expr,                 \
                            cudaGetErrorString(e__), __FILE__, __LINE__);                   \
    } while (0)

#define CT_TRY(expr)                  \
    do {                              \
        int rc__ = (expr);            \
        if (rc__ != CT_OK) return rc__; \
    } while (0)

// record a timing mark on the handle's stream: the span since the previous mark belongs to launch `id`
inline void prof_mark(ct_context *h, int id) {
    if (!h->prof_on) return;
    if (h->prof_n == h->prof_ev.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        h->prof_ev.push_back(e);
        h->prof_id.push_back(0);
    }
    h->prof_id[h->prof_n] = id;
    cudaEventRecord(h->prof_ev[h->prof_n++], h->stream);
}

inline int src_kind(const ct_batch *b) { return b->dtype * 2 + b->layout; }
inline size_t elem_size(int dtype) { return dtype == CT_F32 ? 4 : (dtype == CT_U8 ? 1 : 8); }

inline int64_t plane_of(const ct_batch *b) { return b->plane_stride ? b->plane_stride : b->npix; }

// 128-bit accesses need: 16 B aligned base, image stride and (planar) plane stride multiples of
// one vector.  HWC groups are 3 vectors, so any 16 B aligned image start works.
inline bool vec_ok(const ct_batch *b) {
    const size_t es = elem_size(b->dtype);
    const int64_t per_vec = 16 / (int64_t)es;
    if (((uintptr_t)b->data) & 15) return false;
    if (b->count > 1 && (b->image_stride % per_vec)) return false;
    if (b->layout == CT_CHW && (plane_of(b) % per_vec)) return false;
    return true;
}

inline Img img_of(const ct_batch *b) {
    return Img{b->data, b->npix, b->count > 1 ? b->image_stride : 0, plane_of(b)};
}
inline ImgOut imgout_of(const ct_batch *b) {
    return ImgOut{b->data, b->npix, b->count > 1 ? b->image_stride : 0, plane_of(b)};
}

int check_batch(ct_context *h, const ct_batch *b, const char *name);
int ensure_scratch(ct_context *h, int pairs);
int ensure_partials(ct_context *h, size_t doubles);
int ensure_ws(ct_context *h, size_t bytes);
int ensure_stage(ct_context *h, size_t bytes);
int ensure_seed(ct_context *h, size_t words);

// ct_linear.cu
int launch_moments(ct_context *h, const ct_batch *a, const ct_batch *b, int lab, double *sums,
                   int method, double *xform, int *status);
int launch_solve(ct_context *h, int method, const double *sums_t, const double *sums_r,
                 int64_t sums_stride, int count, double *xform, int *status);
int launch_apply(ct_context *h, int method, const ct_batch *target, const double *xform,
                 const ct_batch *out);

// ct_u8.cu
int launch_u8_to_float(ct_context *h, const uint8_t *in, void *out, int dtype, int64_t n);
int launch_float_to_u8(ct_context *h, const void *in, int dtype, uint8_t *out, int64_t n);

// ct_distort.cu
int launch_distort(ct_context *h, const ct_batch *src, const ct_distortion *ops, int n_ops, const ct_batch *dst);

// ct_regrain.cu
size_t regrain_workspace_bytes(int H, int W);
int launch_regrain(ct_context *h, const double *in, const double *col, double *out, int H, int W, void *workspace,
                   size_t workspace_bytes);
int launch_to_f64_hwc(ct_context *h, const ct_batch *img, double *out);

// ct_metrics.cu
int launch_icid(ct_context *h, const float *img1, const float *img2, int B, int H, int W, int intent,
                int omit_maps67, int downsampling, double *out_dev);
int launch_psnr(ct_context *h, const float *x, const float *y, int B, int64_t n, double *out_dev);
int launch_ssim(ct_context *h, const float *x, const float *y, int B, int H, int W, int downsample, double *out_dev);

// ct_idt.cu
int launch_keys_init(ct_context *h, int64_t *keys, int64_t n);
int launch_ranges(ct_context *h, const ct_batch *img, const double *rot, int64_t rot_stride, int n_rot,
                  int64_t *keys, int64_t keys_stride, int32_t *status);
int launch_ranges_pair(ct_context *h, const ct_batch *target, const ct_batch *reference, int n_rot_r,
                       const double *rot, int64_t rot_stride, int64_t *keys, int64_t keys_stride, int32_t *status,
                       int64_t *init_keys, int64_t n_init);
int launch_hist(ct_context *h, const ct_idt_stage *s, int fuse_lut, const ct_idt_trace *trace,
                int trace_iter, int trace_niter);
int launch_lut(ct_context *h, const ct_idt_stage *s, int keep_counts, const ct_idt_trace *trace,
               int trace_iter, int trace_niter);
int launch_remap(ct_context *h, const ct_idt_stage *s, const ct_batch *dst, int round_f32);

}  // namespace ct
