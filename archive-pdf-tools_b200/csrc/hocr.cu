// hocr.cu -- the per-text-line measurements of create_hocr_mask (internetarchivepdf/mrc.py:188-270) for a whole
// page's worth of line crops at once:
//   k_rects_count_nonzero : np.count_nonzero(thres) of every crop (mrc.py:231, 236)
//   k_rects_sigma_bool    : mean_estimate_sigma(thres) (mrc.py:253-254) = scikit-image estimate_sigma of a BOOLEAN
//                           array: PyWavelets promotes bool to float64, so the db2 'dd' band is computed in double
//                           (True = 1.0, PyWavelets accumulation order, no FMA) and the median of |dd != 0| is a
//                           float64 median: exact order statistics by an 8 x 8-bit radix select over the 64-bit
//                           patterns, both middle ranks, np.median's (a + b) / 2.
// Third-party algorithm restated from its published form ("parity unpinned"); the CPU
// restatement it is tested against is oracle/mrc_oracle.c orc_estimate_sigma_bool.
// The Sauvola passes on the crops (k = 0.1, plain and inverted input) are b200mrc_sauvola calls.
#include "common.cuh"
#include <cstring>
#include <mutex>
#include <math_constants.h>

namespace b200mrc {
namespace {

__device__ __forceinline__ int sym_idx(int i, int n)
{
    if (n == 1) return 0;
    const int per = 2 * n;
    i %= per; if (i < 0) i += per;
    return i < n ? i : per - 1 - i;
}

__global__ void __launch_bounds__(256) k_rects_count_nonzero(const b200mrc_rect *rects, uint32_t *counts)
{
    const b200mrc_rect r = rects[blockIdx.x];
    uint32_t c = 0;
    for (int y = threadIdx.x >> 5; y < r.height; y += 8) {
        const uint8_t *row = r.ptr + (int64_t)y * r.pitch;
        for (int x = threadIdx.x & 31; x < r.width; x += 32) c += row[x] != 0;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    __shared__ uint32_t s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) counts[blockIdx.x] = s[0] + s[1] + s[2] + s[3] + s[4] + s[5] + s[6] + s[7];
}

// n independent 2-D byte copies (line-crop gather / mask pastes): grid.z = rectangle, a CTA copies 8 rows x 1024 bytes
__global__ void __launch_bounds__(256) k_rects_copy(const b200mrc_copy_rect *rects)
{
    const b200mrc_copy_rect r = rects[blockIdx.z];
    const int y0 = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (y0 >= r.height) return;
    const uint8_t *s = r.src + (int64_t)y0 * r.src_pitch;
    uint8_t *d = r.dst + (int64_t)y0 * r.dst_pitch;
    for (int x = blockIdx.x * 1024 + (threadIdx.x & 31); x < min(r.width, (int)(blockIdx.x + 1) * 1024); x += 32) d[x] = s[x];
}

// f2  mask hand-off to the encoder (encode_mrc_mask, mrc.py:474-520: Image.fromarray(np_mask).save(png)): the
// boolean mask as PIL mode-'1' rows, 8 pixels per byte, most significant bit first (== np.packbits(mask, axis=1)),
// optionally inverted (recode.py:408: np_mask ^ True for --bw-pdf).  One thread per output byte.
__global__ void __launch_bounds__(256) k_pack_mask(const uint8_t *mask, int64_t pitch, int64_t stride,
                                                   uint8_t *out, int64_t opitch, int64_t ostride, int W, int H, int invert)
{
    const int bx = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y, page = blockIdx.z;
    const int nb = (W + 7) >> 3;
    if (bx >= nb) return;
    const uint8_t *row = mask + (int64_t)page * stride + (int64_t)y * pitch + 8 * bx;
    uint32_t w0, w1;
    if (8 * bx + 8 <= W && ((pitch | stride) & 3) == 0 && (((uintptr_t)mask) & 3) == 0) {
        w0 = *reinterpret_cast<const uint32_t *>(row); w1 = *reinterpret_cast<const uint32_t *>(row + 4);
    } else {
        w0 = w1 = 0;
        for (int j = 0; j < 8 && 8 * bx + j < W; j++) {
            if (j < 4) w0 |= (uint32_t)row[j] << (8 * j); else w1 |= (uint32_t)row[j] << (8 * (j - 4));
        }
    }
    // any non-zero byte = set; 4 bytes holding 0/1 -> one nibble, first pixel = highest bit
    w0 = ((((w0 & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w0) >> 7) & 0x01010101u;
    w1 = ((((w1 & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w1) >> 7) & 0x01010101u;
    uint32_t b = (((w0 * 0x08040201u) >> 24) << 4) | ((w1 * 0x08040201u) >> 24);
    if (invert) {
        const int valid = min(8, W - 8 * bx);                // bits beyond the row stay 0, like np.packbits
        b = (~b) & (0xffu << (8 - valid)) & 0xffu;
    }
    out[(int64_t)page * ostride + (int64_t)y * opitch + bx] = (uint8_t)b;
}

constexpr int SG_T = 1024;

// 0-based rank `rank` among the non-zero keys of keys[0..n): 8 passes of 8 bits, most significant first
__device__ uint64_t select_rank(const uint64_t *keys, int n, uint32_t rank, uint32_t *hist, uint32_t *s_res)
{
    uint64_t prefix = 0;
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 56 - 8 * pass;
        for (int i = threadIdx.x; i < 256; i += SG_T) hist[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += SG_T) {
            const uint64_t k = keys[i];
            if (k != 0 && (pass == 0 || (k >> (shift + 8)) == (prefix >> (shift + 8)))) atomicAdd(&hist[(k >> shift) & 0xff], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t acc = 0;
            int b = 0;
            for (; b < 256; b++) { if (rank < acc + hist[b]) break; acc += hist[b]; }
            s_res[0] = (uint32_t)b; s_res[1] = rank - acc;
        }
        __syncthreads();
        prefix |= (uint64_t)s_res[0] << shift;
        rank = s_res[1];
        __syncthreads();
    }
    return prefix;
}

__global__ void __launch_bounds__(SG_T) k_rects_sigma_bool(const b200mrc_rect *rects, double *sigma)
{
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_res[2];
    __shared__ uint32_t s_cnt;
    const b200mrc_rect r = rects[blockIdx.x];
    const int h = r.height, w = r.width, oh = (h + 3) / 2, ow = (w + 3) / 2, n = oh * ow;
    const double f[4] = {-0.48296291314469025, 0.836516303737469, -0.22414386804185735, -0.12940952255092145};
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    uint32_t mine = 0;
    for (int i = threadIdx.x; i < n; i += SG_T) {
        const int yo = i / ow, xo = i - yo * ow;
        double dd = 0.0;
#pragma unroll
        for (int j2 = 0; j2 < 4; j2++) {
            const int col = sym_idx(2 * xo + 1 - j2, w);
            double d0 = 0.0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const double px = r.ptr[(int64_t)sym_idx(2 * yo + 1 - j, h) * r.pitch + col] ? 1.0 : 0.0;
                d0 = __dadd_rn(d0, __dmul_rn(f[j], px));
            }
            dd = __dadd_rn(dd, __dmul_rn(f[j2], d0));
        }
        const uint64_t key = dd == 0.0 ? 0ull : ((uint64_t)__double_as_longlong(dd) & 0x7fffffffffffffffull);
        r.keys[i] = key;
        mine += key != 0;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, d);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_cnt, mine);
    __syncthreads();
    const uint32_t cnt = s_cnt;
    if (cnt == 0) {                                          // np.median([]) -> nan
        if (threadIdx.x == 0) sigma[blockIdx.x] = CUDART_NAN;
        return;
    }
    const uint64_t a = select_rank(r.keys, n, (cnt - 1) / 2, hist, s_res);
    const uint64_t b = (cnt & 1) ? a : select_rank(r.keys, n, cnt / 2, hist, s_res);
    if (threadIdx.x == 0) {
        const double med = (cnt & 1) ? __longlong_as_double((long long)a)
                                     : __ddiv_rn(__dadd_rn(__longlong_as_double((long long)a), __longlong_as_double((long long)b)), 2.0);
        sigma[blockIdx.x] = __ddiv_rn(med, 0.6744897501960817);
    }
}

}  // namespace
}  // namespace b200mrc

using namespace b200mrc;

extern "C" int b200mrc_rects_count_nonzero(const b200mrc_rect *rects_dev, int n_rects, uint32_t *counts_dev, void *stream)
{
    if (n_rects < 0 || (n_rects && (!rects_dev || !counts_dev))) return B200MRC_ERR_INVALID;
    if (n_rects == 0) return B200MRC_OK;
    { ProfScope _ps("k_rects_count_nonzero", (cudaStream_t)stream); k_rects_count_nonzero<<<n_rects, 256, 0, (cudaStream_t)stream>>>(rects_dev, counts_dev); }
    B200MRC_LAUNCH_CHECK();
    return B200MRC_OK;
}

extern "C" int b200mrc_rects_copy(const b200mrc_copy_rect *rects_dev, int n_rects, int max_width, int max_height, void *stream)
{
    if (n_rects < 0 || (n_rects && (!rects_dev || max_width <= 0 || max_height <= 0))) return B200MRC_ERR_INVALID;
    if (n_rects == 0) return B200MRC_OK;
    if (n_rects > 65535 || cdiv(max_height, 8) > 65535) return B200MRC_ERR_UNSUPPORTED;
    dim3 grid(cdiv(max_width, 1024), cdiv(max_height, 8), n_rects);
    { ProfScope _ps("k_rects_copy", (cudaStream_t)stream); k_rects_copy<<<grid, 256, 0, (cudaStream_t)stream>>>(rects_dev); }
    B200MRC_LAUNCH_CHECK();
    return B200MRC_OK;
}

extern "C" int b200mrc_pack_mask(const uint8_t *mask, int64_t pitch, int64_t page_stride,
                                 uint8_t *packed, int64_t packed_pitch, int64_t packed_page_stride,
                                 int width, int height, int n_pages, int invert, void *stream)
{
    if (!mask || !packed || width <= 0 || height <= 0 || n_pages <= 0) return B200MRC_ERR_INVALID;
    if (packed_pitch < (width + 7) / 8 || pitch < width) return B200MRC_ERR_INVALID;
    if (height > 65535 || n_pages > 65535) return B200MRC_ERR_UNSUPPORTED;
    dim3 grid(cdiv((width + 7) / 8, 256), height, n_pages);
    { ProfScope _ps("k_pack_mask", (cudaStream_t)stream);
      k_pack_mask<<<grid, 256, 0, (cudaStream_t)stream>>>(mask, pitch, page_stride, packed, packed_pitch, packed_page_stride, width, height, invert ? 1 : 0); }
    B200MRC_LAUNCH_CHECK();
    return B200MRC_OK;
}

// Host side of the packed hand-off: mode-'1' rows that crossed the bus (1 bit per pixel) back into the bool plane the
// reference's callers index (mrc.py:399 yields a bool H x W array).  Plain C++ on HOST memory, one thread per call
// (callers run it from worker threads, one call per chunk of pages); no CUDA call inside.
extern "C" int b200mrc_host_unpack_mask(const uint8_t *packed, int64_t packed_pitch, int64_t packed_page_stride,
                                        uint8_t *mask, int64_t pitch, int64_t page_stride,
                                        int width, int height, int n_pages)
{
    if (!packed || !mask || width <= 0 || height <= 0 || n_pages <= 0) return B200MRC_ERR_INVALID;
    if (packed_pitch < (width + 7) / 8 || pitch < width) return B200MRC_ERR_INVALID;
    static uint64_t lut[256];
    static std::once_flag once;
    std::call_once(once, [] {
        for (int b = 0; b < 256; b++) {
            uint64_t v = 0;
            for (int i = 0; i < 8; i++) v |= (uint64_t)((b >> (7 - i)) & 1) << (8 * i);      // first pixel = most significant bit
            lut[b] = v;
        }
    });
    const int full = width / 8, rest = width - 8 * full;
    for (int n = 0; n < n_pages; n++)
        for (int y = 0; y < height; y++) {
            const uint8_t *src = packed + (int64_t)n * packed_page_stride + (int64_t)y * packed_pitch;
            uint8_t *dst = mask + (int64_t)n * page_stride + (int64_t)y * pitch;
            for (int i = 0; i < full; i++) memcpy(dst + 8 * i, &lut[src[i]], 8);
            if (rest) {
                const uint64_t v = lut[src[full]];
                memcpy(dst + 8 * full, &v, (size_t)rest);
            }
        }
    return B200MRC_OK;
}

extern "C" int b200mrc_rects_sigma_bool(const b200mrc_rect *rects_dev, int n_rects, double *sigma_dev, void *stream)
{
    if (n_rects < 0 || (n_rects && (!rects_dev || !sigma_dev))) return B200MRC_ERR_INVALID;
    if (n_rects == 0) return B200MRC_OK;
    { ProfScope _ps("k_rects_sigma_bool", (cudaStream_t)stream); k_rects_sigma_bool<<<n_rects, SG_T, 0, (cudaStream_t)stream>>>(rects_dev, sigma_dev); }
    B200MRC_LAUNCH_CHECK();
    return B200MRC_OK;
}
