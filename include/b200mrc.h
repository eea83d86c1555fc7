/*
 * b200mrc.h -- C ABI of libb200mrc.so, the B200 (sm_100a) MRC page-decomposition engine.
 *
 * This is the drop-in boundary for the per-page pixel path of internetarchive/archive-pdf-tools
 * (reference citations are relative to the reference checkout):
 *
 *   reference native surface (Cython `def`s imported at internetarchivepdf/mrc.py:36-37)
 *     binarise_sauvola(in, out, width, height, window_width, window_height, k, R)  cython/sauvola.pyx:29
 *     fast_mask_denoise(mask, width, height, mincnt, n_size)                      cython/optimiser.pyx:436
 *     optimise_gray / optimise_gray2(mask, img, width, height, n_size)            cython/optimiser.pyx:22 / :153
 *     optimise_rgb  / optimise_rgb2 (mask, img, width, height, n_size)            cython/optimiser.pyx:83 / :280
 *   reference numeric glue (third-party calls inside mrc.py / grayconvert.py)
 *     PIL convert('L')            mrc.py:361        scipy gaussian_filter   mrc.py:311
 *     skimage estimate_sigma      mrc.py:52-55,294  PIL Image.thumbnail     mrc.py:427, 462
 *     special_gray_convert        grayconvert.py:38-66
 *
 * Conventions
 *   - plain C: device pointers, sizes, a cudaStream_t passed as void*; no torch / C++ types.
 *   - every image argument is a batch of `n_pages` equally-shaped pages: `ptr` = page 0 row 0,
 *     `pitch` = bytes between rows, `page_stride` = bytes between pages.  Pixels are uint8,
 *     interleaved (channels = 1 or 3).  Masks are one byte per pixel holding 0 or 1 (numpy bool).
 *   - all calls are asynchronous on `stream` unless stated otherwise; they return 0 on success,
 *     a negative B200MRC_ERR_* for argument errors, or a positive cudaError_t.
 *   - no CPU fallback exists: without a CUDA device every compute entry point fails.
 */
#ifndef B200MRC_H
#define B200MRC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200MRC_API __attribute__((visibility("default")))
#else
#define B200MRC_API
#endif

#define B200MRC_VERSION 100            /* 0.1.0 */

#define B200MRC_OK                 0
#define B200MRC_ERR_INVALID       -1   /* bad argument (null pointer, non-positive size, ...)      */
#define B200MRC_ERR_UNSUPPORTED   -2   /* parameter outside the implemented range                 */
#define B200MRC_ERR_WORKSPACE     -3   /* workspace too small                                     */
#define B200MRC_ERR_ALIGNMENT     -4   /* pointer / pitch alignment requirement violated          */

#define B200MRC_MAX_WINDOW       255   /* Sauvola window side (reference uses 33..151)            */
#define B200MRC_MAX_OPT_N         16   /* optimise n_size (reference uses 3 and 10)               */

/* flags for b200mrc_sauvola */
#define B200MRC_SAUVOLA_OR_INTO      1 /* out |= fg   (mask_arr |= thres_arr, mrc.py:329)           */
#define B200MRC_SAUVOLA_RAW_INVERTED 2 /* store !fg: exactly what binarise_sauvola writes (sauvola.pyx:153) */
#define B200MRC_SAUVOLA_INVERT_INPUT 4 /* threshold 255 - p: the inverted line crop of create_hocr_mask (mrc.py:226, 235) */

/* flags for b200mrc_decompose */
#define B200MRC_DECOMPOSE_DENOISE_FAST 1 /* denoise_mask == 'fast' (mrc.py:384-390)                 */
#define B200MRC_DECOMPOSE_MASK_ONLY    2 /* stop after the first yield (recode.py:398-407, --bw-pdf) */
#define B200MRC_DECOMPOSE_NO_NOISE_EST 4 /* skip estimate_noise; use sigma_in (or 0 => no blur)     */
#define B200MRC_DECOMPOSE_OR_INTO_MASK 8 /* `mask` already holds the hOCR line masks: mask |= thres (mrc.py:329) */

B200MRC_API int         b200mrc_version(void);
B200MRC_API const char *b200mrc_error_string(int status);
/* number of CUDA kernels this library has launched in the calling process (bench bookkeeping) */
B200MRC_API uint64_t    b200mrc_launch_count(void);
/* Tuning knobs for A/B runs and tests (kernel forms, band sizes, page groups ...; INTEGRATION.md lists the names).  Each
 * knob is initialised once per process from the environment variable B200MRC_<NAME>; set_tuning changes it for the
 * launches that follow (process-wide, atomic).  Unknown names return B200MRC_ERR_INVALID. */
B200MRC_API int         b200mrc_set_tuning(const char *name, int value);
B200MRC_API int         b200mrc_get_tuning(const char *name, int *value);
/* Per-kernel device timing: while enabled, every kernel launch of this library is bracketed by a CUDA
 * event pair on its own stream.  b200mrc_profile_report synchronises and writes "kernel,launches,total_ms"
 * lines into buf (returns the full length); enable(0/1) also clears the records. */
B200MRC_API int         b200mrc_profile_enable(int on);
B200MRC_API int         b200mrc_profile_report(char *buf, size_t cap);

/* Page transfer between HOST memory (pinned for true asynchrony) and pitched device planes: what
 * np.array(image) / returning a numpy array is at the reference's boundary (mrc.py:372, 399, 436,
 * 470).  `rows` rows of `row_bytes` bytes each; a whole batch is one call when page_stride ==
 * height * pitch (rows = n_pages * height).  kind: 1 = host to device, 2 = device to host,
 * 3 = device to device.  One DMA (cudaMemcpy2DAsync) on `stream`: no staging copy, no kernel. */
#define B200MRC_COPY_H2D 1
#define B200MRC_COPY_D2H 2
#define B200MRC_COPY_D2D 3
B200MRC_API int b200mrc_copy2d(void *dst, int64_t dst_pitch, const void *src, int64_t src_pitch,
                   int64_t row_bytes, int64_t rows, int kind, void *stream);

/* A1  PIL convert('L') (mrc.py:361): L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16. */
B200MRC_API int b200mrc_rgb2gray(const uint8_t *rgb, int64_t rgb_pitch, int64_t rgb_page_stride,
                     uint8_t *gray, int64_t gray_pitch, int64_t gray_page_stride,
                     int width, int height, int n_pages, void *stream);

/* A6 (+A5, A7)  binarise_sauvola (sauvola.pyx:29-222) as used by threshold_image (mrc.py:58-87).
 * `in` is a 1-channel uint8 plane.  By default writes fg (== threshold_image's return value);
 * see the B200MRC_SAUVOLA_* flags.  window sides 1..255, k >= 0 or k < 0 (both reference
 * branches), any R > 0.  Requires in/out pitch % 4 == 0 and 4-byte aligned base pointers. */
B200MRC_API int b200mrc_sauvola(const uint8_t *in, int64_t in_pitch, int64_t in_page_stride,
                    uint8_t *out, int64_t out_pitch, int64_t out_page_stride,
                    int width, int height, int n_pages,
                    int window_width, int window_height, double k, double R,
                    int flags, void *stream);

/* A3  estimate_noise (mrc.py:273-296): sigma of the centre crop, one double per page written to
 * the device array `sigma_out`.  `in` may have 1 or 3 channels (gray conversion fused).
 * workspace: b200mrc_noise_workspace_bytes(). */
B200MRC_API size_t b200mrc_noise_workspace_bytes(int width, int height, int n_pages);
B200MRC_API int b200mrc_estimate_noise(const uint8_t *in, int64_t in_pitch, int64_t in_page_stride, int channels,
                           int width, int height, int n_pages, double *sigma_out,
                           void *workspace, size_t workspace_bytes, void *stream);

/* A1+A2+A4  gray conversion and the conditional Gaussian pre-blur of create_threshold_mask
 * (mrc.py:305-325): for each page p, if sigma[p] > 1.0 the float32 gray image is filtered with
 * scipy.ndimage.gaussian_filter(sigma = 0.1*sigma[p]) semantics and truncated to uint8; otherwise
 * (or when `sigma` is NULL) the plain gray plane is written.  `sigma` is a DEVICE array. */
B200MRC_API int b200mrc_gray_blur(const uint8_t *in, int64_t in_pitch, int64_t in_page_stride, int channels,
                      uint8_t *gray_out, int64_t gray_pitch, int64_t gray_page_stride,
                      int width, int height, int n_pages, const double *sigma, void *stream);

/* A1+A2+A4+A5+A6(+A7)  create_threshold_mask (mrc.py:300-329) in one pass: convert('L') -> conditional Gaussian
 * pre-blur (sigma = 0.1 * sigma[p] where sigma[p] > 1.0; `sigma` is a DEVICE array or NULL) -> uint8 truncation ->
 * binarise_sauvola + invert -> out (or out |= with B200MRC_SAUVOLA_OR_INTO).  `in` has 1 or 3 channels.  With
 * 16-byte aligned rows this is a single fused kernel (TMA-fed, no gray plane pass); otherwise it runs
 * b200mrc_gray_blur + b200mrc_sauvola.  workspace: b200mrc_threshold_workspace_bytes() (a gray delay-line plane). */
B200MRC_API size_t b200mrc_threshold_workspace_bytes(int width, int height, int n_pages);
B200MRC_API int b200mrc_threshold_mask(const uint8_t *in, int64_t in_pitch, int64_t in_page_stride, int channels,
                           uint8_t *out, int64_t out_pitch, int64_t out_page_stride,
                           int width, int height, int n_pages,
                           int window_width, int window_height, double k, double R,
                           const double *sigma, int flags,
                           void *workspace, size_t workspace_bytes, void *stream);

/* A8  fast_mask_denoise (optimiser.pyx:436-472), in place, exact raster-order semantics.
 * mincnt = 4, n_size = 2 -- the only configuration the reference uses (mrc.py:388) -- is one persistent bit-sliced
 * kernel on `workspace` (b200mrc_denoise_workspace_bytes()).  Any other (mincnt, 0 <= n_size <= 64) takes a general
 * one-thread-per-pixel form of the same fixed point: stream-ordered temporaries (cudaMallocAsync, one byte per pixel), one
 * launch per pass and a stream synchronisation after each -- it completes the Cython signature, it is not a fast path;
 * `workspace` is not used by it. */
B200MRC_API size_t b200mrc_denoise_workspace_bytes(int width, int height, int n_pages);
B200MRC_API int b200mrc_denoise(uint8_t *mask, int64_t pitch, int64_t page_stride,
                    int width, int height, int n_pages, int mincnt, int n_size,
                    void *workspace, size_t workspace_bytes, void *stream);

/* A9  optimise_{gray,rgb}[2] (optimiser.pyx:22-429).  One sweep produces both layers of
 * create_mrc_hocr_components: fg = optimise(mask, img, n_fg) (mrc.py:412-415) and
 * bg = optimise(mask ^ 1, img, n_bg) (mrc.py:439-449).  Either output may be NULL.  A single
 * reference call optimise_*(mask, img, w, h, n) is (out_fg, n_fg = n, out_bg = NULL).
 * 1 <= n <= 16.  workspace: b200mrc_optimise_workspace_bytes(). */
B200MRC_API size_t b200mrc_optimise_workspace_bytes(int width, int height, int n_pages);
B200MRC_API int b200mrc_optimise(const uint8_t *mask, int64_t mask_pitch, int64_t mask_page_stride,
                     const uint8_t *img, int64_t img_pitch, int64_t img_page_stride, int channels,
                     uint8_t *out_fg, int64_t fg_pitch, int64_t fg_page_stride, int n_fg,
                     uint8_t *out_bg, int64_t bg_pitch, int64_t bg_page_stride, int n_bg,
                     int width, int height, int n_pages,
                     void *workspace, size_t workspace_bytes, void *stream);

/* A10  PIL Image.thumbnail((int(w/f), int(h/f))) (mrc.py:420-434, 454-468): optional box
 * `reduce` followed by the two-pass 8-bit fixed-point BICUBIC (filter 0) or LANCZOS (filter 1)
 * resample.  A plan holds the size logic and the device coefficient tables.
 * plan_create returns NULL when Pillow would leave the image untouched (nothing to do) or on
 * failure (*status tells which). */
typedef struct b200mrc_resample_plan b200mrc_resample_plan;
B200MRC_API b200mrc_resample_plan *b200mrc_thumbnail_plan_create(int width, int height, int channels,
                                                     double req_width, double req_height,
                                                     double reducing_gap /* <=0: none */, int filter,
                                                     int *status);
B200MRC_API void   b200mrc_resample_plan_destroy(b200mrc_resample_plan *plan);
B200MRC_API void   b200mrc_resample_plan_out_size(const b200mrc_resample_plan *plan, int *out_width, int *out_height);
B200MRC_API size_t b200mrc_resample_workspace_bytes(const b200mrc_resample_plan *plan, int n_pages);
B200MRC_API int    b200mrc_resample(const b200mrc_resample_plan *plan,
                        const uint8_t *in, int64_t in_pitch, int64_t in_page_stride,
                        uint8_t *out, int64_t out_pitch, int64_t out_page_stride, int n_pages,
                        void *workspace, size_t workspace_bytes, void *stream);

/* f1  create_hocr_mask (mrc.py:188-270): per text line, Sauvola (k = 0.1) on the crop and on the inverted crop
 * (b200mrc_sauvola, B200MRC_SAUVOLA_INVERT_INPUT), np.count_nonzero of both (mrc.py:231, 236) and, for the
 * undecided lines, mean_estimate_sigma of the two BOOLEAN results (mrc.py:253-254).  The two measurements run
 * over a DEVICE array of crop descriptors, one CTA per crop:
 *   ptr/pitch/width/height : a 0/1 byte plane (a Sauvola output);
 *   keys                   : scratch of ((height+3)/2) * ((width+3)/2) uint64 (sigma only). */
typedef struct {
    const uint8_t *ptr; int64_t pitch; int32_t width, height; uint64_t *keys;
} b200mrc_rect;
/* The line crops of a page as ONE launch each (mrc.py:188-270 walks them one by one):
 *   b200mrc_rects_copy    : n 2-D byte copies described by a DEVICE array (crop gather into aligned scratch, and the
 *                           pastes `mask_arr[top:bottom, left:right] = th`, mrc.py:266); the rectangles of one call must
 *                           not overlap in their destinations.
 *   b200mrc_sauvola_items : b200mrc_sauvola on n independent single-channel images (each item is thresholded as an
 *                           image of its own: windows clamp at the item's borders); per-item flags B200MRC_SAUVOLA_*.
 * max_width / max_height: the largest item (sizes the launch grid). */
typedef struct {
    const uint8_t *src; int64_t src_pitch; uint8_t *dst; int64_t dst_pitch; int32_t width, height;
} b200mrc_copy_rect;
typedef struct {
    const uint8_t *in; int64_t in_pitch; uint8_t *out; int64_t out_pitch; int32_t width, height, flags, reserved;
} b200mrc_sauvola_item;
B200MRC_API int b200mrc_rects_copy(const b200mrc_copy_rect *rects_dev, int n_rects, int max_width, int max_height, void *stream);
B200MRC_API int b200mrc_sauvola_items(const b200mrc_sauvola_item *items_dev, int n_items, int max_width, int max_height,
                          int window_width, int window_height, double k, double R, void *stream);
B200MRC_API int b200mrc_rects_count_nonzero(const b200mrc_rect *rects_dev, int n_rects, uint32_t *counts_dev, void *stream);
B200MRC_API int b200mrc_rects_sigma_bool(const b200mrc_rect *rects_dev, int n_rects, double *sigma_dev, void *stream);

/* f2  The mask as the encoder wants it (encode_mrc_mask, mrc.py:474-520 saves Image.fromarray(np_mask), a PIL
 * mode-'1' image): rows of ceil(width/8) bytes, 8 pixels per byte, first pixel in the most significant bit
 * (== np.packbits(mask, axis=1)); `invert` applies recode.py:408's np_mask ^ True on the way. */
B200MRC_API int b200mrc_pack_mask(const uint8_t *mask, int64_t pitch, int64_t page_stride,
                      uint8_t *packed, int64_t packed_pitch, int64_t packed_page_stride,
                      int width, int height, int n_pages, int invert, void *stream);

/* ... and back on the HOST: packed rows that crossed the bus -> the bool (0/1 byte) plane the reference's callers index
 * (the first yield of create_mrc_hocr_components, mrc.py:399; == np.unpackbits(packed, axis=1)[:, :width]).  `packed`
 * and `mask` are HOST pointers; plain single-threaded C++ (call it from worker threads, one call per chunk of pages);
 * little-endian hosts. */
B200MRC_API int b200mrc_host_unpack_mask(const uint8_t *packed, int64_t packed_pitch, int64_t packed_page_stride,
                             uint8_t *mask, int64_t pitch, int64_t page_stride,
                             int width, int height, int n_pages);

/* A12  special_gray_convert (grayconvert.py:38-66).  Two steps with a host decision between them,
 * exactly like the reference: (1) per-channel min / max / sum / sum-of-squares of each page
 * (`stats_out`: n_pages x 3 x 4 uint64 on the DEVICE: min, max, sum, sumsq), (2) per-pixel level
 * stretch + HSL lightness with per-page `minv`/`maxv` (DEVICE arrays, n_pages x 3 doubles). */
B200MRC_API int b200mrc_channel_stats(const uint8_t *rgb, int64_t pitch, int64_t page_stride,
                          int width, int height, int n_pages, uint64_t *stats_out, void *stream);
B200MRC_API int b200mrc_special_gray(const uint8_t *rgb, int64_t pitch, int64_t page_stride,
                         uint8_t *gray, int64_t gray_pitch, int64_t gray_page_stride,
                         int width, int height, int n_pages,
                         const double *minv, const double *maxv, void *stream);

/* A11  The whole of create_mrc_hocr_components (mrc.py:334-471) for hocr_word_data == [] on a
 * device-resident batch: gray -> noise estimate -> conditional blur -> Sauvola -> denoise ->
 * fg/bg optimise -> bg (and fg) thumbnail.  All stages are enqueued on `stream`; nothing
 * synchronises with the host.
 *   sigma_in  : optional DEVICE array (n_pages doubles) used instead of the estimate when
 *               B200MRC_DECOMPOSE_NO_NOISE_EST is set (NULL => no blur).
 *   sigma_out : optional DEVICE array receiving the sigma used per page.
 *   bg_plan / fg_plan : thumbnail plans or NULL (no downsample => out_* is full resolution).
 */
typedef struct {
    const uint8_t *img;  int64_t img_pitch, img_page_stride;  int channels;
    int width, height, n_pages;
    int window;                      /* Sauvola window side (mrc.py:70-75: 51 or odd(int(dpi/4))) */
    double k, R;                     /* 0.34, 128 (mrc.py:58, 82)                                 */
    int flags;                       /* B200MRC_DECOMPOSE_*                                       */
    const double *sigma_in;  double *sigma_out;
    uint8_t *mask;   int64_t mask_pitch, mask_page_stride;
    uint8_t *fg;     int64_t fg_pitch, fg_page_stride;     const b200mrc_resample_plan *fg_plan;
    uint8_t *bg;     int64_t bg_pitch, bg_page_stride;     const b200mrc_resample_plan *bg_plan;
    void *workspace; size_t workspace_bytes;
} b200mrc_decompose_args;

B200MRC_API size_t b200mrc_decompose_workspace_bytes(const b200mrc_decompose_args *args);
B200MRC_API int    b200mrc_decompose(const b200mrc_decompose_args *args, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200MRC_H */
