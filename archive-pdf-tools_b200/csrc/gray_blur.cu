// gray_blur.cu -- k_gray_blur: PIL convert('L') (mrc.py:358-363) fused with the conditional
// Gaussian pre-blur of create_threshold_mask (mrc.py:305-325):
//     if sigma_est > 1.0: imgf = scipy.ndimage.gaussian_filter(imgf, sigma=sigma_est*0.1)
//     ... imgf.astype(np.uint8)
// scipy semantics restated (see oracle/mrc_oracle.c orc_gauss_blur, pinned bit-exact against
// scipy 1.18): radius = int(4*sigma+0.5); weights exp(-0.5/sigma^2 * j^2) / sum in double;
// correlate1d axis 0 then axis 1, symmetric fast path acc = x*w0; for j=radius..1:
// acc += (x[i-j] + x[i+j]) * w_j in double (no FMA), float32 store after each axis; 'reflect'.
//
// One CTA = one output tile of one page: gray tile + halo -> smem (u8), vertical pass -> smem
// (f32), horizontal pass -> uint8 truncation -> global.  sigma lives on the device, so the
// per-page decision (blur or plain gray copy, and the radius) is taken by the kernel without a
// host round trip; two tile configurations cover radius 0..16 and 17..128.
#include "common.cuh"
#include "blur.cuh"
#include <cstdlib>

namespace b200mrc {
namespace {

struct GrayBlurParams {
    const uint8_t *in; int64_t in_pitch, in_stride; int C;
    uint8_t *out; int64_t out_pitch, out_stride;
    int W, H;
    const double *sigma;     // device, per page, may be null
    int *err;                // device error flag (radius out of range), may be null
    int rlo;                 // tiled kernel: smallest radius it handles (5 when the fast kernel ran, else 0)
    int fast_rmin;           // fast kernel: smallest radius it handles (0; 3 behind the fused threshold kernel)
    int vec4;                // planes, pitches and strides are 4-byte aligned: the tiled kernels load 4 pixels per thread
};

__device__ __forceinline__ uint32_t load_gray(const uint8_t *page, int64_t pitch, int C, int y, int x)
{
    const uint8_t *px = page + (int64_t)y * pitch + (int64_t)x * C;
    if (C == 1) return px[0];
    return luma_l24(px[0], px[1], px[2]);
}

// One tile of one page.  256 threads as 32 x 8: no integer division in the loops.
// Conversions stay off the (quarter-rate) conversion unit: integer sums become doubles as 2^52 + x - 2^52 on the FP64 pipe,
// and for radius <= 16 the vertical pass leaves its float32-rounded results in shared memory AS doubles (TMP = double:
// exactly the same values), so the horizontal pass loads its operands ready to use.
template <int C>
__device__ __forceinline__ uint32_t load_gray4(const GrayBlurParams &p, const uint8_t *in, int y, int x);

// The gray tile (+ halo) is fetched as aligned groups of 4 pixels (3 words of RGB -> one word of gray, load_gray4; only
// groups that touch the page edge take the per-pixel reflected form): the tile's rows in shared memory start at the
// 4-aligned column xs <= x0 - radius, `off` columns before the first one the passes use.
template <int RHI, int TH, int TW, typename TMP>
__device__ __forceinline__ void blur_tile(const GrayBlurParams &p, const uint8_t *in, uint8_t *out, int x0, int y0,
                                          int radius, const double *sw, TMP *stmp, uint8_t *sg)
{
    constexpr int GW = TW + 2 * RHI, SGW = GW + 8;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int gw = TW + 2 * radius, gh = TH + 2 * radius;
    const int gx0 = x0 - radius, xs = (gx0 >> 2) << 2, off = gx0 - xs;
    if (p.vec4) {
        const int ngrp = (off + gw + 3) >> 2;
        for (int ly = ty; ly < gh; ly += 8) {
            const int y = reflect_idx(y0 - radius + ly, p.H);
            uint32_t *row = reinterpret_cast<uint32_t *>(sg + ly * SGW);
            for (int g = tx; g < ngrp; g += 32)
                row[g] = p.C == 3 ? load_gray4<3>(p, in, y, xs + 4 * g) : load_gray4<1>(p, in, y, xs + 4 * g);
        }
    } else {
        for (int ly = ty; ly < gh; ly += 8) {
            const int y = reflect_idx(y0 - radius + ly, p.H);
            for (int lx = tx; lx < gw; lx += 32) {
                const int x = reflect_idx(gx0 + lx, p.W);
                sg[ly * SGW + off + lx] = (uint8_t)load_gray(in, p.in_pitch, p.C, y, x);
            }
        }
    }
    __syncthreads();
    // axis 0 (vertical), for every column incl. the horizontal halo
    for (int ly = ty; ly < TH; ly += 8)
        for (int lx = tx; lx < gw; lx += 32) {
            const uint8_t *c = sg + (ly + radius) * SGW + off + lx;
            double acc = __dmul_rn(u2d(c[0]), sw[0]);
            for (int j = radius; j >= 1; j--)
                acc = __dadd_rn(acc, __dmul_rn(u2d((uint32_t)c[-j * SGW] + (uint32_t)c[j * SGW]), sw[j]));
            stmp[ly * GW + lx] = (TMP)__double2float_rn(acc);            // float32 store between the axes (scipy)
        }
    __syncthreads();
    // axis 1 (horizontal) + uint8 truncation
    for (int ly = ty; ly < TH; ly += 8) {
        const int y = y0 + ly;
        if (y >= p.H) break;
        for (int lx = tx; lx < TW; lx += 32) {
            const int x = x0 + lx;
            if (x >= p.W) break;
            const TMP *c = stmp + ly * GW + lx + radius;
            double acc = __dmul_rn((double)c[0], sw[0]);
            for (int j = radius; j >= 1; j--)
                acc = __dadd_rn(acc, __dmul_rn(__dadd_rn((double)c[-j], (double)c[j]), sw[j]));
            const float f = __double2float_rn(acc);
            out[(int64_t)y * p.out_pitch + x] = (uint8_t)(int)f;          // astype(uint8): truncation
        }
    }
}

template <int RHI> struct BlurTmp { using type = float; };      // radius up to 128: the tile's halo needs the shared memory
template <> struct BlurTmp<16> { using type = double; };

__device__ __forceinline__ int page_radius(const GrayBlurParams &p, int page, double &sigma)
{
    return blur_radius_of(p.sigma, page, sigma);
}

__device__ __forceinline__ void blur_weights(int radius, double sigma, double *sw, double *sphi)
{
    blur_weights_cta(radius, sigma, sw, sphi);
}

// Small-radius configuration (0..RHI): one CTA per tile, grid over (tiles, pages).
template <int RHI, int TH, int TW>
__global__ void __launch_bounds__(256) k_gray_blur(const GrayBlurParams p)
{
    constexpr int GW = TW + 2 * RHI;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    double *sw = reinterpret_cast<double *>(smem_raw);                    // RHI+1 weights
    double *sphi = sw + (RHI + 1);                                        // 2*RHI+1 scratch
    using TMP = typename BlurTmp<RHI>::type;
    TMP *stmp = reinterpret_cast<TMP *>(sphi + (2 * RHI + 1));            // TH x GW
    uint8_t *sg = reinterpret_cast<uint8_t *>(stmp + TH * GW);            // (TH + 2 RHI) x GW

    const int page = blockIdx.z;
    double sigma;
    const int radius = page_radius(p, page, sigma);
    if (radius > RHI || radius < p.rlo) return;   // the large-radius / fast kernel owns this page
    const uint8_t *in = p.in + (int64_t)page * p.in_stride;
    uint8_t *out = p.out + (int64_t)page * p.out_stride;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    if (radius == 0) {
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
        for (int ly = ty; ly < TH; ly += 8) {
            const int y = y0 + ly;
            if (y >= p.H) break;
            for (int lx = tx; lx < TW; lx += 32) {
                const int x = x0 + lx;
                if (x < p.W) out[(int64_t)y * p.out_pitch + x] = (uint8_t)load_gray(in, p.in_pitch, p.C, y, x);
            }
        }
        return;
    }
    blur_weights(radius, sigma, sw, sphi);
    blur_tile<RHI, TH, TW, TMP>(p, in, out, x0, y0, radius, sw, stmp, sg);
}

// Large-radius configuration (RLO..RHI): persistent CTAs walk the pages, skip those the small
// kernel handled, and share the tiles of the rest -- no grid of a million early-exit CTAs.
template <int RLO, int RHI, int TH, int TW>
__global__ void __launch_bounds__(256) k_gray_blur_large(const GrayBlurParams p, int n_pages)
{
    constexpr int GW = TW + 2 * RHI;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    double *sw = reinterpret_cast<double *>(smem_raw);
    double *sphi = sw + (RHI + 1);
    using TMP = typename BlurTmp<RHI>::type;
    TMP *stmp = reinterpret_cast<TMP *>(sphi + (2 * RHI + 1));
    uint8_t *sg = reinterpret_cast<uint8_t *>(stmp + TH * GW);
    const int tiles_x = (p.W + TW - 1) / TW, tiles_y = (p.H + TH - 1) / TH;
    for (int page = 0; page < n_pages; page++) {
        double sigma;
        const int radius = page_radius(p, page, sigma);
        if (radius < RLO) continue;
        if (radius > RHI) {
            if (RHI >= 128 && p.err && threadIdx.x == 0 && blockIdx.x == 0) atomicExch(p.err, 1);
            continue;
        }
        __syncthreads();
        blur_weights(radius, sigma, sw, sphi);
        const uint8_t *in = p.in + (int64_t)page * p.in_stride;
        uint8_t *out = p.out + (int64_t)page * p.out_stride;
        if constexpr (RHI > 32 && RLO <= 32) {
            // radius <= 32 inside the wide-radius launch: a 32 x 64 tile with a 32-pixel halo budget fits the same shared
            // memory and fetches 4.4x its pixels at radius 24, where the 16 x 32 tile sized for radius 128 fetches 10x
            if (radius <= 32) {
                constexpr int MTH = 32, MTW = 64, MR = 32;
                float *mtmp = reinterpret_cast<float *>(stmp);
                uint8_t *msg = reinterpret_cast<uint8_t *>(mtmp + MTH * (MTW + 2 * MR));
                const int mx = (p.W + MTW - 1) / MTW, my = (p.H + MTH - 1) / MTH;
                for (int t = blockIdx.x; t < mx * my; t += gridDim.x) {
                    __syncthreads();
                    blur_tile<MR, MTH, MTW, float>(p, in, out, (t % mx) * MTW, (t / mx) * MTH, radius, sw, mtmp, msg);
                }
                continue;
            }
        }
        for (int t = blockIdx.x; t < tiles_x * tiles_y; t += gridDim.x) {
            __syncthreads();
            blur_tile<RHI, TH, TW, TMP>(p, in, out, (t % tiles_x) * TW, (t / tiles_x) * TH, radius, sw, stmp, sg);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fast path for radius 0..4 (sigma_est < 11.25, i.e. every ordinary scan): a warp marches a
// 128-px-wide column strip down a band of rows, 4 adjacent pixels per lane.  The (2R+1)-row gray
// window lives in registers (one u32 = 4 gray bytes per row), the vertical pass needs no
// communication, the horizontal pass takes its R neighbours from the adjacent lanes by shuffle;
// lanes 0 and 31 only supply halo.  Loads are 32-bit (3 per 4 RGB pixels), stores one u32.
constexpr int FB_BAND = 64;            // output rows per warp

template <int C>
__device__ __forceinline__ uint32_t load_gray4(const GrayBlurParams &p, const uint8_t *in, int y, int x)
{
    // 4 gray pixels x..x+3 of row y (x % 4 == 0); columns outside [0, W) are reflected
    if (x >= 0 && x + 3 < p.W) {
        const uint8_t *q = in + (int64_t)y * p.in_pitch + (int64_t)x * C;
        if (C == 1) return *reinterpret_cast<const uint32_t *>(q);
        const uint32_t a = *reinterpret_cast<const uint32_t *>(q), b = *reinterpret_cast<const uint32_t *>(q + 4),
                       c = *reinterpret_cast<const uint32_t *>(q + 8);
        // a = r0 g0 b0 r1 | b = g1 b1 r2 g2 | c = b2 r3 g3 b3
        const uint32_t g0 = luma_l24(a & 0xff, (a >> 8) & 0xff, (a >> 16) & 0xff);
        const uint32_t g1 = luma_l24(a >> 24, b & 0xff, (b >> 8) & 0xff);
        const uint32_t g2 = luma_l24((b >> 16) & 0xff, b >> 24, c & 0xff);
        const uint32_t g3 = luma_l24((c >> 8) & 0xff, (c >> 16) & 0xff, c >> 24);
        return g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
    }
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) r |= load_gray(in, p.in_pitch, C, y, reflect_idx(x + k, p.W)) << (8 * k);
    return r;
}

template <int R, int C>
__device__ __forceinline__ void blur_march(const GrayBlurParams &p, const uint8_t *in, uint8_t *out,
                                           int x, int by0, int by1, const double *sw)
{
    const int lane = threadIdx.x & 31;
    double w[R + 1];
#pragma unroll
    for (int j = 0; j <= R; j++) w[j] = sw[j];
    uint32_t win[2 * R + 1];                       // gray rows y-R .. y+R (4 px each)
#pragma unroll
    for (int j = 0; j < 2 * R; j++) win[j + 1] = load_gray4<C>(p, in, reflect_idx(by0 - R + j, p.H), x);
    for (int y = by0; y < by1; y++) {
#pragma unroll
        for (int j = 0; j < 2 * R; j++) win[j] = win[j + 1];
        win[2 * R] = load_gray4<C>(p, in, reflect_idx(y + R, p.H), x);
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            double acc = __dmul_rn(u2d((win[R] >> (8 * k)) & 0xff), w[0]);
#pragma unroll
            for (int j = R; j >= 1; j--)
                acc = __dadd_rn(acc, __dmul_rn(u2d(((win[R - j] >> (8 * k)) & 0xff) + ((win[R + j] >> (8 * k)) & 0xff)), w[j]));
            v[k] = __double2float_rn(acc);
        }
        uint32_t o = 0;
        if (R == 0) {
            o = win[0];
        } else {
            // horizontal neighbours: ext[R + k] = v[k]; ext[0..R) from the left lane, ext[R+4..) from the right lane
            double ext[4 + 2 * R];
#pragma unroll
            for (int k = 0; k < 4; k++) ext[R + k] = (double)v[k];
#pragma unroll
            for (int j = 0; j < R; j++) {      // neighbours' values arrive already widened (shuffle, not convert)
                ext[j] = __shfl_up_sync(0xffffffffu, ext[R + 4 - R + j], 1);
                ext[R + 4 + j] = __shfl_down_sync(0xffffffffu, ext[R + j], 1);
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                double acc = __dmul_rn(ext[R + k], w[0]);
#pragma unroll
                for (int j = R; j >= 1; j--) acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(ext[R + k - j], ext[R + k + j]), w[j]));
                o |= ((uint32_t)(int)__double2float_rn(acc) & 0xffu) << (8 * k);      // astype(uint8): truncation
            }
        }
        if (lane >= 1 && lane <= 30 && x < p.W) {
            uint8_t *q = out + (int64_t)y * p.out_pitch + x;
            if (x + 3 < p.W) *reinterpret_cast<uint32_t *>(q) = o;
            else for (int k = 0; k < 4 && x + k < p.W; k++) q[k] = (uint8_t)(o >> (8 * k));
        }
    }
}

template <int C, int MINB>
__global__ void __launch_bounds__(256, MINB) k_gray_blur_fast(const GrayBlurParams p)
{
    __shared__ double sw[8];
    __shared__ double sphi[16];
    const int page = blockIdx.z;
    double sigma;
    const int radius = page_radius(p, page, sigma);
    if (radius > 4 || radius < p.fast_rmin) return;   // the tiled kernels / the fused threshold kernel own this page
    if (radius > 0) blur_weights(radius, sigma, sw, sphi);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int by0 = (blockIdx.y * 8 + warp) * FB_BAND;
    if (by0 >= p.H) return;
    const int by1 = min(p.H, by0 + FB_BAND);
    const int x = blockIdx.x * 120 - 4 + 4 * lane;
    const uint8_t *in = p.in + (int64_t)page * p.in_stride;
    uint8_t *out = p.out + (int64_t)page * p.out_stride;
    switch (radius) {
    case 0: blur_march<0, C>(p, in, out, x, by0, by1, sw); break;
    case 1: blur_march<1, C>(p, in, out, x, by0, by1, sw); break;
    case 2: blur_march<2, C>(p, in, out, x, by0, by1, sw); break;
    case 3: blur_march<3, C>(p, in, out, x, by0, by1, sw); break;
    default: blur_march<4, C>(p, in, out, x, by0, by1, sw); break;
    }
}

template <int RHI, int TH, int TW>
constexpr size_t gray_blur_smem()
{
    return sizeof(double) * (RHI + 1) + sizeof(double) * (2 * RHI + 1) + sizeof(typename BlurTmp<RHI>::type) * TH * (TW + 2 * RHI) +
           (size_t)(TH + 2 * RHI) * (TW + 2 * RHI + 8);
}

}  // namespace

int launch_gray_blur(const uint8_t *in, int64_t in_pitch, int64_t in_stride, int C,
                     uint8_t *out, int64_t out_pitch, int64_t out_stride,
                     int W, int H, int N, const double *sigma, int *err_flag, cudaStream_t st)
{
    GrayBlurParams p{in, in_pitch, in_stride, C, out, out_pitch, out_stride, W, H, sigma, err_flag, 0, 0, 0};
    // fast marching kernel: needs 4-byte aligned rows and an image at least as large as its reflect reach
    const bool aligned4 = !(in_pitch & 3) && !(out_pitch & 3) && !((uintptr_t)in & 3) && !((uintptr_t)out & 3) &&
                          !(in_stride & 3) && !(out_stride & 3);
    const bool fast_ok = aligned4 && W >= 8 && H >= 8;
    p.vec4 = aligned4 ? 1 : 0;
    if (fast_ok) {
        dim3 grid(cdiv(W, 120), cdiv(cdiv(H, FB_BAND), 8), N);
        // 6 CTAs / SM (40 registers, a few spills) beats 4 (60 registers) by 10 %
        { ProfScope _ps("k_gray_blur_fast", st);
          if (C == 1) k_gray_blur_fast<1, 4><<<grid, 256, 0, st>>>(p);
          else k_gray_blur_fast<3, 6><<<grid, 256, 0, st>>>(p); }
        B200MRC_LAUNCH_CHECK();
        p.rlo = 5;
    }
    if (!fast_ok) {
        // tiny images: the tiled kernel does everything up to radius 16 (one CTA per tile)
        constexpr int TH = 32, TW = 128, RHI = 16;
        constexpr size_t smem = gray_blur_smem<RHI, TH, TW>();
        B200MRC_CUDA_TRY(cudaFuncSetAttribute(k_gray_blur<RHI, TH, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(cdiv(W, TW), cdiv(H, TH), N);
        { ProfScope _ps("k_gray_blur", st); k_gray_blur<RHI, TH, TW><<<grid, 256, smem, st>>>(p); }
        B200MRC_LAUNCH_CHECK();
    } else if (sigma) {
        // radius 5..16: persistent CTAs walk the pages and share the tiles of those that need it
        constexpr int TH = 32, TW = 128, RHI = 16;
        constexpr size_t smem = gray_blur_smem<RHI, TH, TW>();
        B200MRC_CUDA_TRY(cudaFuncSetAttribute(k_gray_blur_large<5, RHI, TH, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int tiles = cdiv(W, TW) * cdiv(H, TH);
        int grid = dev_info().sm_count * 4;
        if (grid > tiles) grid = tiles;
        { ProfScope _ps("k_gray_blur_large<5..16>", st); k_gray_blur_large<5, RHI, TH, TW><<<grid, 256, smem, st>>>(p, N); }
        B200MRC_LAUNCH_CHECK();
    }
    if (sigma) {
        constexpr int TH = 16, TW = 32, RHI = 128;
        constexpr size_t smem = gray_blur_smem<RHI, TH, TW>();
        B200MRC_CUDA_TRY(cudaFuncSetAttribute(k_gray_blur_large<17, RHI, TH, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int tiles = cdiv(W, TW) * cdiv(H, TH);
        int grid = dev_info().sm_count * 2;
        if (grid > tiles) grid = tiles;
        { ProfScope _ps("k_gray_blur_large<17..128>", st); k_gray_blur_large<17, RHI, TH, TW><<<grid, 256, smem, st>>>(p, N); }
        B200MRC_LAUNCH_CHECK();
    }
    return B200MRC_OK;
}

// Only the pages whose blur radius is >= rmin (3 <= rmin <= 5), through the tiled persistent kernels: every other page is
// skipped on the device.  Used by the fused threshold path (sauvola_fused.cu), which blurs radius <= 2 itself.
int launch_gray_blur_min_radius(const uint8_t *in, int64_t in_pitch, int64_t in_stride, int C,
                                uint8_t *out, int64_t out_pitch, int64_t out_stride,
                                int W, int H, int N, const double *sigma, int rmin, int *err_flag, cudaStream_t st)
{
    if (!sigma) return B200MRC_OK;
    if (rmin != 3) return B200MRC_ERR_UNSUPPORTED;
    GrayBlurParams p{in, in_pitch, in_stride, C, out, out_pitch, out_stride, W, H, sigma, err_flag, 5, rmin, 1};   // vec4: fast_ok is required below
    const bool fast_ok = !(in_pitch & 3) && !(out_pitch & 3) && !((uintptr_t)in & 3) && !((uintptr_t)out & 3) &&
                         !(in_stride & 3) && !(out_stride & 3) && W >= 8 && H >= 8;
    if (!fast_ok) return B200MRC_ERR_ALIGNMENT;
    {
        // radius 3..4: the register-window marching kernel (pages of any other radius leave at once)
        dim3 grid(cdiv(W, 120), cdiv(cdiv(H, FB_BAND), 8), N);
        { ProfScope _ps("k_gray_blur_fast<3..4>", st);
          if (C == 1) k_gray_blur_fast<1, 4><<<grid, 256, 0, st>>>(p);
          else k_gray_blur_fast<3, 6><<<grid, 256, 0, st>>>(p); }
        B200MRC_LAUNCH_CHECK();
    }
    {
        constexpr int TH = 32, TW = 128, RHI = 16;
        constexpr size_t smem = gray_blur_smem<RHI, TH, TW>();
        B200MRC_CUDA_TRY(cudaFuncSetAttribute(k_gray_blur_large<5, RHI, TH, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int tiles = cdiv(W, TW) * cdiv(H, TH);
        int grid = dev_info().sm_count * 4;
        if (grid > tiles) grid = tiles;
        { ProfScope _ps("k_gray_blur_large<5..16>", st); k_gray_blur_large<5, RHI, TH, TW><<<grid, 256, smem, st>>>(p, N); }
        B200MRC_LAUNCH_CHECK();
    }
    {
        constexpr int TH = 16, TW = 32, RHI = 128;
        constexpr size_t smem = gray_blur_smem<RHI, TH, TW>();
        B200MRC_CUDA_TRY(cudaFuncSetAttribute(k_gray_blur_large<17, RHI, TH, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int tiles = cdiv(W, TW) * cdiv(H, TH);
        int grid = dev_info().sm_count * 2;
        if (grid > tiles) grid = tiles;
        { ProfScope _ps("k_gray_blur_large<17..128>", st); k_gray_blur_large<17, RHI, TH, TW><<<grid, 256, smem, st>>>(p, N); }
        B200MRC_LAUNCH_CHECK();
    }
    return B200MRC_OK;
}

}  // namespace b200mrc

using namespace b200mrc;

extern "C" int b200mrc_gray_blur(const uint8_t *in, int64_t in_pitch, int64_t in_page_stride, int channels,
                                 uint8_t *gray_out, int64_t gray_pitch, int64_t gray_page_stride,
                                 int width, int height, int n_pages, const double *sigma, void *stream)
{
    if (!in || !gray_out || width <= 0 || height <= 0 || n_pages <= 0) return B200MRC_ERR_INVALID;
    if (channels != 1 && channels != 3) return B200MRC_ERR_UNSUPPORTED;
    if (n_pages > 65535) return B200MRC_ERR_UNSUPPORTED;
    return launch_gray_blur(in, in_pitch, in_page_stride, channels, gray_out, gray_pitch, gray_page_stride,
                            width, height, n_pages, sigma, nullptr, (cudaStream_t)stream);
}

extern "C" int b200mrc_rgb2gray(const uint8_t *rgb, int64_t rgb_pitch, int64_t rgb_page_stride,
                                uint8_t *gray, int64_t gray_pitch, int64_t gray_page_stride,
                                int width, int height, int n_pages, void *stream)
{
    return b200mrc_gray_blur(rgb, rgb_pitch, rgb_page_stride, 3, gray, gray_pitch, gray_page_stride,
                             width, height, n_pages, nullptr, stream);
}
