// special_gray.cu -- k_channel_stats + k_special_gray: special_gray_convert
// (internetarchivepdf/grayconvert.py:38-66, level_arr :24-31), used by recode.py:360-364 for
// --grayscale-pdf.  The reference takes per-channel min/max/mean/std of the whole page, derives
// three percentage thresholds on the host, level-stretches each channel in uint8 and takes the
// HSL lightness through scikit-image's rgb2hsv (third-party, not installed: "parity unpinned",
// restated in oracle/mrc_oracle.c orc_special_gray_pixels).
//   k_channel_stats : exact integer reductions (min, max, sum, sum of squares as uint64) -- the
//                     host turns them into mean/std/thresholds exactly like grayconvert.py:41-55;
//   k_special_gray  : per pixel, FP64 with individually rounded operations:
//                     v = (uint8)((x - minv) / interval) (0 below minv, 255 above maxv),
//                     a = v * (1/255.), V = max a, S = (max-min)/max, l = V*(1 - S/2),
//                     out = (uint8)(l*255).
#include "common.cuh"

namespace b200mrc {
namespace {

__global__ void __launch_bounds__(256) k_channel_stats(const uint8_t *rgb, int64_t pitch, int64_t stride,
                                                       int W, int H, unsigned long long *stats)
{
    // grid: (chunks, N).  stats[page][c][4] = min, max, sum, sumsq  (min pre-set to 255 by the host)
    const int page = blockIdx.y;
    const uint8_t *img = rgb + (int64_t)page * stride;
    const int64_t npx = (int64_t)W * H;
    unsigned mn[3] = {255, 255, 255}, mx[3] = {0, 0, 0};
    unsigned long long sm[3] = {0, 0, 0}, sq[3] = {0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < npx; i += (int64_t)gridDim.x * 256) {
        const int y = (int)(i / W), x = (int)(i - (int64_t)y * W);
        const uint8_t *px = img + (int64_t)y * pitch + 3 * (int64_t)x;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const unsigned v = px[c];
            mn[c] = min(mn[c], v); mx[c] = max(mx[c], v); sm[c] += v; sq[c] += v * v;
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            mn[c] = min(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], d));
            mx[c] = max(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], d));
            sm[c] += __shfl_xor_sync(0xffffffffu, sm[c], d);
            sq[c] += __shfl_xor_sync(0xffffffffu, sq[c], d);
        }
        if ((threadIdx.x & 31) == 0) {
            unsigned long long *s = stats + ((int64_t)page * 3 + c) * 4;
            atomicMin(&s[0], (unsigned long long)mn[c]);
            atomicMax(&s[1], (unsigned long long)mx[c]);
            atomicAdd(&s[2], sm[c]);
            atomicAdd(&s[3], sq[c]);
        }
    }
}

__global__ void k_stats_init(unsigned long long *stats, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) stats[i] = (i & 3) == 0 ? 255ull : 0ull;
}

__global__ void __launch_bounds__(256) k_special_gray(const uint8_t *rgb, int64_t pitch, int64_t stride,
                                                      uint8_t *gray, int64_t gpitch, int64_t gstride,
                                                      int W, int H, const double *minv, const double *maxv)
{
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y, page = blockIdx.z;
    if (x >= W) return;
    const uint8_t *px = rgb + (int64_t)page * stride + (int64_t)y * pitch + 3 * (int64_t)x;
    const double inv255 = 1.0 / 255.0;
    double a[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double lo = minv[page * 3 + c], hi = maxv[page * 3 + c];
        const double xv = (double)px[c];
        const double interval = __dsub_rn(__ddiv_rn(hi, 255.), __ddiv_rn(lo, 255.));
        int v;
        if (xv < lo) v = 0;
        else if (xv > hi) v = 255;
        else v = (int)__ddiv_rn(__dsub_rn(xv, lo), interval) & 0xff;
        a[c] = __dmul_rn((double)v, inv255);
    }
    const double mx = fmax(a[0], fmax(a[1], a[2])), mn = fmin(a[0], fmin(a[1], a[2]));
    const double delta = __dsub_rn(mx, mn);
    const double s = delta == 0.0 ? 0.0 : __ddiv_rn(delta, mx);
    const double l = __dmul_rn(mx, __dsub_rn(1.0, __ddiv_rn(s, 2.0)));
    gray[(int64_t)page * gstride + (int64_t)y * gpitch + x] = (uint8_t)(int)__dmul_rn(l, 255.0);
}

}  // namespace
}  // namespace b200mrc

using namespace b200mrc;

extern "C" int b200mrc_channel_stats(const uint8_t *rgb, int64_t pitch, int64_t page_stride,
                                     int width, int height, int n_pages, uint64_t *stats_out, void *stream)
{
    if (!rgb || !stats_out || width <= 0 || height <= 0 || n_pages <= 0) return B200MRC_ERR_INVALID;
    if (n_pages > 65535) return B200MRC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int n = n_pages * 12;
    k_stats_init<<<cdiv(n, 256), 256, 0, st>>>((unsigned long long *)stats_out, n);
    B200MRC_LAUNCH_CHECK();
    const int64_t npx = (int64_t)width * height;
    int chunks = (int)((npx + 256 * 16 - 1) / (256 * 16));
    if (chunks > 1024) chunks = 1024;
    if (chunks < 1) chunks = 1;
    k_channel_stats<<<dim3(chunks, n_pages), 256, 0, st>>>(rgb, pitch, page_stride, width, height,
                                                          (unsigned long long *)stats_out);
    B200MRC_LAUNCH_CHECK();
    return B200MRC_OK;
}

extern "C" int b200mrc_special_gray(const uint8_t *rgb, int64_t pitch, int64_t page_stride,
                                    uint8_t *gray, int64_t gray_pitch, int64_t gray_page_stride,
                                    int width, int height, int n_pages,
                                    const double *minv, const double *maxv, void *stream)
{
    if (!rgb || !gray || !minv || !maxv || width <= 0 || height <= 0 || n_pages <= 0) return B200MRC_ERR_INVALID;
    if (n_pages > 65535 || height > 65535) return B200MRC_ERR_UNSUPPORTED;
    k_special_gray<<<dim3(cdiv(width, 256), height, n_pages), 256, 0, (cudaStream_t)stream>>>(
        rgb, pitch, page_stride, gray, gray_pitch, gray_page_stride, width, height, minv, maxv);
    B200MRC_LAUNCH_CHECK();
    return B200MRC_OK;
}
