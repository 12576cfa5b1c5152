/* ct_b200.h - C ABI of libct_b200.so: B200 (sm_100a) kernels for the global statistical
 * colour-transfer path of egorchistov/color-transfer.
 *
 * The reference has no FFI: its plugin boundary is a Python dotted path (`func_spec`) resolved
 * by importlib (ref: methods/__init__.py:14-16, configs/others.yaml:5) to a callable
 * `f(target[H,W,3], reference[H',W',3]) -> ndarray[H,W,3]`.  This header is what the Python
 * modules that keep that contract (color-transfer_b200/methods/linear.py, iterative.py) bind
 * through ctypes; every entry point cites the reference lines it replaces.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no exceptions: every call returns CT_OK (0) or a
 *    negative ct_status; ct_last_error(h) gives the text.
 *  - One handle per device and host thread.  Device-pointer calls are asynchronous and ordered
 *    on the handle's stream (ct_set_stream); *_host calls return when the output is in host
 *    memory.
 *  - The caller owns every data buffer.  Images are never modified.
 *  - An image is N = H*W pixels of 3 channels, float32 or float64, either interleaved
 *    (CT_HWC: [N,3], what a C-contiguous numpy [H,W,3] array is) or planar (CT_CHW: [3,N],
 *    what the reference Runner's permuted CHW tensor views are, ref: methods/__init__.py:21-22).
 */
#ifndef CT_B200_H
#define CT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CT_ABI_VERSION 2

typedef struct ct_context *ct_handle;

typedef enum ct_status {
    CT_OK = 0,
    CT_E_INVALID = -1,     /* bad argument (null pointer, unknown enum, npix <= 0, ...)            */
    CT_E_CUDA = -2,        /* a CUDA runtime call failed                                          */
    CT_E_NONFINITE = -3,   /* IDT: projected range not finite -> numpy's ValueError               */
    CT_E_NOT_PD = -4,      /* MKL "cholesky": covariance not positive definite -> LinAlgError     */
    CT_E_SINGULAR = -5,    /* singular covariance in an inverse -> LinAlgError                    */
    CT_E_UNSUPPORTED = -6, /* e.g. bins above CT_IDT_MAX_BINS                                     */
    CT_E_NOMEM = -7        /* workspace too small / allocation failed                             */
} ct_status;

enum { CT_F32 = 0, CT_F64 = 1, CT_U8 = 2 };
enum { CT_HWC = 0, CT_CHW = 1 };

/* Which closed-form transfer (ref: methods/linear.py). */
enum {
    CT_REINHARD = 0,     /* color_transfer_between_images,            linear.py:8-42   */
    CT_CCS = 1,          /* color_transfer_in_correlated_color_space, linear.py:45-82  */
    CT_MKL_MK = 2,       /* monge_kantorovitch_color_transfer "MK",       linear.py:116-118 */
    CT_MKL_SQRT = 3,     /* ... decomposition="sqrt",                     linear.py:112-115 */
    CT_MKL_CHOLESKY = 4  /* ... decomposition="cholesky",                 linear.py:108-111 */
};

/* A uniform batch of `count` images (count = 1 for a single call of the reference API). */
typedef struct ct_batch {
    void *data;           /* element 0 of image 0                                             */
    int64_t npix;         /* H*W of every image (> 0)                                          */
    int64_t image_stride; /* elements between consecutive images (ignored when count == 1)     */
    int64_t plane_stride; /* CT_CHW: elements between channel planes (0 means npix)            */
    int32_t count;        /* B                                                                 */
    int32_t dtype;        /* CT_F32 | CT_F64 | CT_U8                                           */
    int32_t layout;       /* CT_HWC | CT_CHW                                                   */
    int32_t flags;        /* CT_BATCH_* bits, 0 by default                                     */
} ct_batch;

/* uint8 images (video frames, torchvision read_image tensors) are decoded INSIDE the kernels that
 * read them and, as outputs, encoded inside the kernel that writes them - no conversion pass:
 *   input : x = k / 255 exactly as the reference's loaders produce it - float64 (skimage.img_as_float,
 *           ref: utils/postprocess.py:138) or, with CT_BATCH_U8_AS_F32, float32 (torch `/ 255`,
 *           ref: utils/data.py:106).  The transfer then runs as if the caller had passed that float image.
 *   output: np.rint(np.clip(y, 0, 1) * 255) (img_as_ubyte of the clipped result, ref:
 *           utils/postprocess.py:138).
 * CT_BATCH_CLAMP01 on a float output applies torch.clamp(y, 0, 1) (ref: methods/__init__.py:30). */
#define CT_BATCH_U8_AS_F32 1
#define CT_BATCH_CLAMP01 2

/* ------------------------------------------------------------------ handle */
int ct_abi_version(void);
int ct_create(int device, ct_handle *out);
void ct_destroy(ct_handle h);
const char *ct_last_error(ct_handle h);
int ct_set_stream(ct_handle h, void *cuda_stream); /* cudaStream_t; NULL = default stream */
int ct_synchronize(ct_handle h);
int ct_sm_count(ct_handle h);
/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t ct_launch_count(ct_handle h);

/* Per-launch device times of the fused IDT driver (bench.py's roofline): while enabled, every launch of
 * ct_idt_transfer is followed by a CUDA event on the handle's stream.  ct_profile_read waits for the last
 * one and returns up to `max` (id, milliseconds) pairs in launch order - the time between consecutive
 * events - and clears the list.  ids: CT_PROF_*. */
enum { CT_PROF_START = 0, CT_PROF_SEED = 1, CT_PROF_RANGES_TARGET = 2, CT_PROF_RANGES_REFERENCE = 3,
       CT_PROF_HIST = 4, CT_PROF_REMAP = 5 };
int ct_profile_enable(ct_handle h, int on);
int ct_profile_read(ct_handle h, int32_t *ids, float *ms, int32_t max);

/* ------------------------------------------------------------------ linear transfers */
/* Raw moments of one image about a fixed shift K (0.5 per RGB channel, (50,0,0) in Lab):
 * sums[b] = { n, S(x-K)[3], S(x-K)(x-K)^T as 00,01,02,11,12,22 }.  They are exactly additive
 * over row shards.  Replaces np.mean / np.std / np.cov at linear.py:33-36, 64-67, 103-106;
 * with lab != 0 the skimage rgb2lab of linear.py:25-26 is fused in front. */
#define CT_MOMENT_DOUBLES 10
int ct_moments(ct_handle h, const ct_batch *images, int lab, double *sums /* dev [B][10] */);

/* One warp: 3x3 algebra of the chosen method from the two moment sets (linear.py:38, 69-78,
 * 108-118).  xform[b] = { M[9] row-major, mu_t[3], mu_r[3], 0 } with
 * out[c] = sum_k (x[k]-mu_t[k]) * M[k][c] + mu_r[c]   (Reinhard: M diagonal, applied in Lab).
 * status[b] (device ints, may be NULL) gets CT_OK / CT_E_NOT_PD / CT_E_SINGULAR. */
#define CT_XFORM_DOUBLES 16
int ct_linear_solve(ct_handle h, int method, const double *sums_t, const double *sums_r,
                    int count, double *xform, int *status);

/* Fused per-pixel remap (linear.py:38-40, 80, 122).  CT_REINHARD: rgb2lab -> affine ->
 * lab2rgb -> clip[0,1].  `out` has the target's npix/count and its own dtype/layout. */
int ct_linear_apply(ct_handle h, int method, const ct_batch *target, const double *xform,
                    const ct_batch *out);

/* Whole transfer in two launches: moments of both images with the solve fused into the last
 * block, then the remap.  xform/status may be NULL (handle scratch is used). */
int ct_linear_transfer(ct_handle h, int method, const ct_batch *target,
                       const ct_batch *reference, const ct_batch *out, double *xform,
                       int *status);

/* Host buffers (numpy memory): H2D, the two launches, D2H; pairs of a batch are pipelined
 * over copy and compute streams.  Returns the first non-OK per-pair status. */
int ct_linear_transfer_host(ct_handle h, int method, const ct_batch *target,
                            const ct_batch *reference, const ct_batch *out);

/* Two-phase host transfer of ONE pair: the statistics come back to the host between the two device passes so
 * that the caller can do the 3x3 algebra itself.  The Python wrapper of color_transfer_in_correlated_color_space
 * uses it to run numpy's SVD (LAPACK dgesdd, the reference's own call at linear.py:69-70) on the two
 * covariances and so reproduce the singular-vector SIGNS the reference gets, which no convention can predict.
 * ct_linear_stats_host copies both images to the device (they stay staged inside the handle until its next
 * host call) and returns their raw moments; ct_linear_apply_staged_host applies `xform` (host, CT_XFORM_DOUBLES,
 * layout of ct_linear_solve) to the staged target and copies the result to `out` (host). */
int ct_linear_stats_host(ct_handle h, int lab, const ct_batch *target, const ct_batch *reference,
                         double *sums_t /* host [10] */, double *sums_r /* host [10] */);
int ct_linear_apply_staged_host(ct_handle h, int method, const double *xform, const ct_batch *out);

/* ------------------------------------------------------------------ IDT (iterative.py:8-59) */
#define CT_IDT_MAX_BINS 1024
#define CT_IDT_KEYS 6 /* per pair and iteration: monotone int64 keys of lo[3], -hi[3] */
/* doubles per pair in a LUT block: edges[3][E], then {fp,slope}[E] per axis (E = CT_IDT_EDGE_STRIDE), then
 * lo,hi,step,inv per axis (layout documented at K6 in csrc/ct_idt.cu) */
#define CT_IDT_EDGE_STRIDE(bins) (((int)(bins) + 2) / 2 * 2) /* bins + 1 rounded up to even */
#define CT_IDT_LUT_DOUBLES(bins) (3 * (3 * (int64_t)CT_IDT_EDGE_STRIDE(bins) + 4))

typedef struct ct_idt_stage {
    const ct_batch *target;    /* current target state (the input images on iteration 0)   */
    const ct_batch *reference;
    const double *rot;         /* dev: this iteration's rotation of pair 0, row-major 3x3  */
    const double *rot_next;    /* dev: next iteration's rotation or NULL                   */
    int64_t rot_stride;        /* doubles between consecutive pairs in rot / rot_next      */
    int64_t *keys;             /* dev: this iteration's CT_IDT_KEYS range keys of pair 0   */
    int64_t *keys_next;        /* dev: next iteration's keys or NULL                       */
    int64_t keys_stride;       /* int64 between consecutive pairs                          */
    uint64_t *counts;          /* dev [B][2][3][bins]: target then reference counts        */
    double *lut;               /* dev [B][CT_IDT_LUT_DOUBLES(bins)]                        */
    int32_t *status;           /* dev [B] or NULL                                          */
    int32_t bins;
    int32_t reserved;
} ct_idt_stage;

/* Host helpers (no device needed): the monotone int64 key of a double and its inverse.  Signed
 * integer order of keys == floating-point order of values, so per-shard ranges combine with an
 * integer MIN all-reduce.  Slot layout per pair and iteration: keys of lo[0..2], then -hi[0..2]. */
int64_t ct_idt_key_of(double value);
double ct_idt_value_of(int64_t key);

/* Set n keys to "+inf" so that atomic minima can be folded in (the fused driver lets K4a do it). */
int ct_idt_keys_init(ct_handle h, int64_t *keys, int64_t n);
/* K4: fold min(p), min(-p) of the projections p = rot_k @ x (iterative.py:34-35, 39-40) of every
 * pixel of `images` into keys, for n_rot consecutive rotations k (rot + 9k -> keys + 6k) in one
 * pass over the image (per 4 rotations: K4a seeds exact extremes from a subsample, K4b screens every
 * pixel in packed fp32 and evaluates the exact fp64 projections only where needed - the result is
 * the exact fp64 minimum / maximum).  The target needs n_rot = 1 (iteration 0); the reference never changes,
 * so its range under all n_iter rotations is taken up front. */
int ct_idt_ranges(ct_handle h, const ct_batch *images, const double *rot, int64_t rot_stride,
                  int32_t n_rot, int64_t *keys, int64_t keys_stride, int32_t *status);
/* K5: projection + shared-memory-privatised histograms of target and reference on the
 * np.histogram grid (iterative.py:42-43); either image may be NULL.  With fuse_lut the last
 * block of each pair runs K6 and clears the counts. */
int ct_idt_hist(ct_handle h, const ct_idt_stage *s, int fuse_lut);
/* K6: CDFs + inverse-CDF table (iterative.py:45-51); clears counts unless keep_counts. */
int ct_idt_lut(ct_handle h, const ct_idt_stage *s, int keep_counts);
/* K7: projection + bin lookup + lerp + back-rotation + state update (iterative.py:53, 55),
 * folding the new state's range under rot_next into keys_next.  round_f32 reproduces the
 * reference's float32 `np.empty_like(target.T)` buffer on iteration 0 of float32 inputs
 * (iterative.py:36). */
int ct_idt_remap(ct_handle h, const ct_idt_stage *s, const ct_batch *dst, int round_f32);

/* Per-iteration intermediates of the fused driver, all device pointers, any may be NULL. */
typedef struct ct_idt_trace {
    double *lo;        /* [B][n_iter][3]        */
    double *hi;        /* [B][n_iter][3]        */
    int64_t *counts_t; /* [B][n_iter][3][bins]  */
    int64_t *counts_r; /* [B][n_iter][3][bins]  */
    double *lut;       /* [B][n_iter][3][bins]  */
} ct_idt_trace;

size_t ct_idt_workspace_bytes(int64_t npix_target, int32_t count, int32_t bins, int32_t n_iter);
/* Whole IDT: 3 + 2*n_iter launches for n_iter <= 4 (one more K4 pair per further 4 rotations of the reference's
 * up-front pass).  rotations: dev [B][n_iter][9], drawn by the caller (the
 * Python wrapper calls scipy.stats.special_ortho_group.rvs once per iteration, in order, so the
 * global numpy RNG advances exactly as at iterative.py:32).  `out`: float64 CT_HWC (the reference's result), or - written
 * by the last iteration's kernel, n_iter >= 2 - uint8 (clip + round) / float32 CT_HWC (optional CT_BATCH_CLAMP01).
 * target / reference may be CT_U8 frames.  workspace may be NULL (the handle then grows its own). */
int ct_idt_transfer(ct_handle h, const ct_batch *target, const ct_batch *reference,
                    const ct_batch *out, const double *rotations, int32_t bins, int32_t n_iter,
                    void *workspace, size_t workspace_bytes, const ct_idt_trace *trace,
                    int32_t *status);
/* Host buffers; rotations and trace members are host pointers. */
int ct_idt_transfer_host(ct_handle h, const ct_batch *target, const ct_batch *reference,
                         const ct_batch *out, const double *rotations, int32_t bins,
                         int32_t n_iter, const ct_idt_trace *trace);

/* ------------------------------------------------------------------ regrain
 * _regrain(img_arr_in, img_arr_col) of ref methods/iterative.py:62-115 on device fp64 [H][W][3]
 * images (multigrid pyramid by skimage.transform.resize semantics + Jacobi relaxation), and
 * automated_color_grading = IDT + regrain (iterative.py:118-138) from host buffers. */
size_t ct_regrain_workspace_bytes(int32_t height, int32_t width);
int ct_regrain(ct_handle h, const double *in, const double *col, double *out, int32_t height, int32_t width,
               void *workspace, size_t workspace_bytes);
int ct_acg_transfer_host(ct_handle h, const ct_batch *target, const ct_batch *reference, const ct_batch *out,
                         int32_t height, int32_t width, const double *rotations, int32_t bins, int32_t n_iter);

/* ------------------------------------------------------------------ uint8 frames
 * Stacks of `count` interleaved uint8 [npix,3] frames in host memory; 3 bytes per pixel cross
 * PCIe in each direction and the kernels decode / encode them as they read / write (CT_U8 batches).  Frames are decoded exactly as the reference's loaders do
 * (as_float32 != 0: float32 k/255, ref utils/data.py:106; else float64 k/255.0 as
 * skimage.img_as_float, ref utils/postprocess.py:138), transferred with the kernels above, and the
 * result is clipped to [0,1] and rounded to uint8 (np.rint, i.e. img_as_ubyte of the clipped
 * image, ref utils/postprocess.py:138). */
int ct_linear_transfer_host_u8(ct_handle h, int method, const uint8_t *target, const uint8_t *reference,
                               uint8_t *out, int32_t count, int64_t npix_target,
                               int64_t npix_reference, int32_t as_float32);
int ct_idt_transfer_host_u8(ct_handle h, const uint8_t *target, const uint8_t *reference, uint8_t *out,
                            int32_t count, int64_t npix_target, int64_t npix_reference,
                            int32_t as_float32, const double *rotations, int32_t bins, int32_t n_iter);

/* ------------------------------------------------------------------ quality metrics
 * The step after the hot path in the reference's test loop (ref: methods/__init__.py:32-40), on
 * planar float32 image batches [count,3,height,width] in DEVICE memory.  Both calls are ordered on
 * the handle's stream, write one double to `result` (HOST memory) and return after it is there.
 *
 * ct_icid replaces ref: utils/icid.py:28 `icid(img1, img2, intent, omit_maps67, downsampling)`:
 * intent 0 = "perceptual", 1 = "hue-preserving", 2 = "chromatic" (anything else: CT_E_INVALID, the
 * reference's ValueError); the result is 1 - mean over batch and pixels of the map product.
 * ct_psnr replaces piq.psnr(x, y) with its defaults (ref: methods/__init__.py:35): the mean over
 * the batch of -10 log10(mse + 1e-8), data_range 1. */
#define CT_ICID_PERCEPTUAL 0
#define CT_ICID_HUE_PRESERVING 1
#define CT_ICID_CHROMATIC 2
int ct_icid(ct_handle h, const float *img1, const float *img2, int32_t count, int32_t height, int32_t width,
            int32_t intent, int32_t omit_maps67, int32_t downsampling, double *result);
int ct_psnr(ct_handle h, const float *x, const float *y, int32_t count, int64_t elems_per_image, double *result);
/* ct_ssim replaces piq.ssim(x, y) with its defaults (ref: methods/__init__.py:36): 11x11 Gaussian of
 * sigma 1.5, valid convolution, k1 = 0.01, k2 = 0.03, average-pool downscale by
 * max(1, round(min(H,W)/256)) when `downsample` != 0; mean over channels and batch. */
int ct_ssim(ct_handle h, const float *x, const float *y, int32_t count, int32_t height, int32_t width,
            int32_t downsample, double *result);

/* ---- artificial-distortion generator ------------------------------
 * Replaces the torchvision.transforms.functional.adjust_* calls behind ref: utils/data.py:12-22
 * `setup_grid_distortions` (identity + {brightness, contrast, saturation, hue, gamma} x 6 magnitudes,
 * applied to the uint8 ground-truth image at ref: utils/data.py:101-104).  One pass over `src`
 * (CT_U8, device, CT_HWC or CT_CHW) writes every distorted copy: image b * n_ops + k of `dst` (same
 * dtype / layout / npix, count = src->count * n_ops) is distortion k of image b.  `ops` is HOST memory.
 * factor: brightness / contrast / saturation factor (>= 0), hue shift in [-0.5, 0.5], gamma (>= 0);
 * out-of-range factors return CT_E_INVALID (torchvision's ValueError). */
#define CT_DISTORT_IDENTITY 0
#define CT_DISTORT_BRIGHTNESS 1
#define CT_DISTORT_CONTRAST 2
#define CT_DISTORT_SATURATION 3
#define CT_DISTORT_HUE 4
#define CT_DISTORT_GAMMA 5
#define CT_DISTORT_MAX_OPS 32
typedef struct ct_distortion {
    int32_t kind; /* CT_DISTORT_* */
    int32_t reserved;
    double factor;
} ct_distortion;
int ct_distort(ct_handle h, const ct_batch *src, const ct_distortion *ops, int32_t n_ops, const ct_batch *dst);

#ifdef __cplusplus
}
#endif
#endif /* CT_B200_H */
