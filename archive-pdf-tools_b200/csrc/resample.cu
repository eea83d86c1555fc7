// resample.cu -- k_resample_tile / k_reduce_box / k_resample_h / k_resample_v: PIL Image.thumbnail at
// internetarchivepdf/mrc.py:420-434 (fg) and :454-468 (bg).
//
// Pillow 12.2 semantics restated (oracle/mrc_oracle.c orc_resample / orc_reduce and
// oracle/oracle.py thumbnail_plan, pinned bit-exact against the installed Pillow):
//   Image.thumbnail(size): aspect-preserving target size; Image.resize(reducing_gap=2.0): box
//   `reduce` by int(scale/2) per axis when > 1 (ImagingReduce: ((sum + cells/2) * mult) >> 24),
//   then ImagingResample: per-output coefficient windows (precompute_coeffs, BICUBIC a=-0.5,
//   support 2*max(scale,1)) normalised in double, quantised to 22-bit fixed point
//   (normalize_coeffs_8bpc), horizontal pass then vertical pass with a uint8 intermediate,
//   out = clip8((2^21 + sum p*k) >> 22).
// The size logic and the coefficient tables are computed on the host once per plan (they depend
// only on the shapes); the kernels are pure gather-multiply-accumulate byte kernels.
// Integer shrink factors without a box reduce (the bg/3 case of the hot path) take k_resample_tile: both passes of a
// 32 x 32 output tile in one CTA, the uint8 intermediate in shared memory.  Its FOLLOW form is launched behind the
// optimise sweep as a programmatic dependent and consumes bg rows while the sweep still writes later ones
// (launch_resample_follow; the sweep publishes per-strip row progress, optimise_warp.cu).
#include "common.cuh"
#include "tma.cuh"
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

struct b200mrc_resample_plan {
    int W, H, C;
    int fx, fy;                 // box reduce factors (1 = none)
    int rx0, ry0, rx1, ry1;     // reduce box on the source
    int SW, SH;                 // size after reduce (== W,H when no reduce)
    float bx0, by0, bx1, by1;   // resample box on the (reduced) source
    int OW, OH, filter;
    int need_h, need_v, ksize_h, ksize_v;
    int *d_bounds_h, *d_kk_h, *d_bounds_v, *d_kk_v;
    // fused tile path (k_resample_tile): both axes shrink by the same integer factor F with T taps, so every interior
    // output uses one coefficient row; [ux0,ux1) x [uy0,uy1) are the outputs that do
    int tile_F, tile_T, offx, offy, ux0, ux1, uy0, uy1;
    int kh[16], kv[16];
};

namespace b200mrc {
namespace {

double bicubic_filter(double x)
{
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

double lanczos_filter(double x)
{
    if (x < 0.0) x = -x;
    if (x >= 3.0) return 0.0;
    if (x == 0.0) return 1.0;
    const double a = x * 3.14159265358979323846, b = a / 3.0;
    return (std::sin(a) / a) * (std::sin(b) / b);
}

// Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc
int build_coeffs(int inSize, float in0, float in1, int outSize, int filter, std::vector<int> &bounds, std::vector<int> &kk)
{
    const double support0 = filter == 1 ? 3.0 : 2.0;
    double scale, filterscale;
    filterscale = scale = (double)(in1 - in0) / outSize;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = support0 * filterscale;
    const int ksize = (int)std::ceil(support) * 2 + 1;
    bounds.assign((size_t)outSize * 2, 0);
    kk.assign((size_t)outSize * ksize, 0);
    std::vector<double> k(ksize);
    for (int xx = 0; xx < outSize; xx++) {
        const double center = in0 + (xx + 0.5) * scale;
        double ww = 0.0;
        const double ss = 1.0 / filterscale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > inSize) xmax = inSize;
        xmax -= xmin;
        int x;
        for (x = 0; x < xmax; x++) {
            const double arg = (x + xmin - center + 0.5) * ss;
            const double w = filter == 1 ? lanczos_filter(arg) : bicubic_filter(arg);
            k[x] = w; ww += w;
        }
        for (x = 0; x < xmax; x++) if (ww != 0.0) k[x] /= ww;
        for (; x < ksize; x++) k[x] = 0;
        bounds[2 * xx] = xmin; bounds[2 * xx + 1] = xmax;
        for (x = 0; x < ksize; x++)
            kk[(size_t)xx * ksize + x] = k[x] < 0 ? (int)(-0.5 + k[x] * (1 << 22)) : (int)(0.5 + k[x] * (1 << 22));
    }
    return ksize;
}

struct ResampleParams {
    const uint8_t *in; int64_t in_pitch, in_stride;
    uint8_t *out; int64_t out_pitch, out_stride;
    int C, in_w, in_h, out_w, out_h, ksize;
    const int *bounds, *kk;
};

__device__ __forceinline__ uint8_t clip8(int v)
{
    v >>= 22;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// out[y][xx][c] = clip8(2^21 + sum_k in[y][xmin+k][c] * kk[xx][k]).  One thread per output PIXEL
// (all channels share the coefficient fetch and the address arithmetic) and RPT consecutive rows
// (the coefficient window is loaded once per thread).
constexpr int H_RPT = 4;
template <int C>
__global__ void __launch_bounds__(128) k_resample_h(const ResampleParams p)
{
    const int xx = blockIdx.x * 128 + threadIdx.x;
    const int y0 = blockIdx.y * H_RPT, page = blockIdx.z;
    if (xx >= p.out_w) return;
    const int xmin = p.bounds[2 * xx], cnt = p.bounds[2 * xx + 1];
    const int *k = p.kk + (size_t)xx * p.ksize;
    const uint8_t *src = p.in + (int64_t)page * p.in_stride + (int64_t)y0 * p.in_pitch + (int64_t)xmin * C;
    uint8_t *dst = p.out + (int64_t)page * p.out_stride + (int64_t)y0 * p.out_pitch + (int64_t)xx * C;
    const int rows = min(H_RPT, p.in_h - y0);
    int acc[H_RPT][C];
#pragma unroll
    for (int r = 0; r < H_RPT; r++)
#pragma unroll
        for (int c = 0; c < C; c++) acc[r][c] = 1 << 21;
    for (int i = 0; i < cnt; i++) {
        const int kv = __ldg(k + i);
#pragma unroll
        for (int r = 0; r < H_RPT; r++) {
            if (r < rows) {
                const uint8_t *q = src + (int64_t)r * p.in_pitch + i * C;
#pragma unroll
                for (int c = 0; c < C; c++) acc[r][c] += (int)q[c] * kv;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < H_RPT; r++)
        if (r < rows) {
#pragma unroll
            for (int c = 0; c < C; c++) dst[(int64_t)r * p.out_pitch + c] = clip8(acc[r][c]);
        }
}

// out[yy][b] = clip8(2^21 + sum_k in[ymin+k][b] * kk[yy][k]); one thread per 4 consecutive bytes
// (rows of both planes are 16-byte aligned workspace / pitched planes)
__global__ void __launch_bounds__(256) k_resample_v(const ResampleParams p)
{
    const int b4 = (blockIdx.x * 256 + threadIdx.x) * 4;
    const int yy = blockIdx.y, page = blockIdx.z;
    const int nb = p.out_w * p.C;
    if (b4 >= nb) return;
    const int ymin = p.bounds[2 * yy], cnt = p.bounds[2 * yy + 1];
    const int *k = p.kk + (size_t)yy * p.ksize;
    const uint8_t *src = p.in + (int64_t)page * p.in_stride + (int64_t)ymin * p.in_pitch + b4;
    int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21, s3 = 1 << 21;
    for (int i = 0; i < cnt; i++) {
        const int kv = __ldg(k + i);
        const uint32_t w = *reinterpret_cast<const uint32_t *>(src + (int64_t)i * p.in_pitch);
        s0 += (int)(w & 0xffu) * kv; s1 += (int)((w >> 8) & 0xffu) * kv;
        s2 += (int)((w >> 16) & 0xffu) * kv; s3 += (int)(w >> 24) * kv;
    }
    uint8_t *dst = p.out + (int64_t)page * p.out_stride + (int64_t)yy * p.out_pitch + b4;
    const uint32_t o = (uint32_t)clip8(s0) | ((uint32_t)clip8(s1) << 8) | ((uint32_t)clip8(s2) << 16) | ((uint32_t)clip8(s3) << 24);
    if (b4 + 3 < nb && (((uintptr_t)dst) & 3) == 0) *reinterpret_cast<uint32_t *>(dst) = o;
    else
        for (int j = 0; j < 4 && b4 + j < nb; j++) dst[j] = (uint8_t)(o >> (8 * j));
}

// ---- fused horizontal + vertical pass for integer shrink factors (the production bg/3 thumbnail) ----------------
// One CTA produces a TOW x TOH output tile: the horizontal pass of the F*TOH+T-F input rows it needs goes into
// shared memory (uint8, exactly Pillow's intermediate), the vertical pass reads it back.  Interior outputs all use
// the same T coefficients, which sit in the kernel's constant bank; a thread of the horizontal pass loads the
// (3F+T)*C input bytes of 4 adjacent outputs as aligned words once and extracts every byte once.  Tiles that
// touch the image border (clipped, renormalised windows) take the table-driven form of the same arithmetic.
constexpr int TOW = 32;                 // tile width; the tile height is a template parameter (32 in production)

struct TileParams {
    const uint8_t *in; int64_t in_pitch, in_stride;
    uint8_t *out; int64_t out_pitch, out_stride;
    int in_w, in_h, out_w, out_h, ksize_h, ksize_v;
    const int *bounds_h, *kk_h, *bounds_v, *kk_v;
    int offx, offy, ux0, ux1, uy0, uy1;
    int kh[16], kv[16];
    // FOLLOW form: the input is still being written by the kernel this one was launched behind (programmatic dependent
    // launch); prog[page * prog_s + x / prog_w] = rows of column strip x / prog_w that are globally visible
    const int *prog; int prog_s, prog_w, ntx, npages;
};

// FOLLOW: a 1-D grid ordered (tile row, page, tile column) -- the order in which a row-sequential producer of all pages
// finishes the input -- whose CTAs wait for their input rows on the producer's progress counters.
template <int C, int F, int T, int TOH, bool FOLLOW>
__global__ void __launch_bounds__(256) k_resample_tile(const TileParams p)
{
    constexpr int RMAX = F * TOH + T, HROW = TOW * C;
    __shared__ __align__(16) uint8_t hbuf[RMAX * HROW];
    const int tid = threadIdx.x;
    int bx = blockIdx.x, by = blockIdx.y, page = blockIdx.z;
    if (FOLLOW) {
        const int per_row = p.ntx * p.npages;
        by = bx / per_row;
        const int r = bx - by * per_row;
        page = r / p.ntx;
        bx = r - page * p.ntx;
    }
    const int ox0 = bx * TOW, oy0 = by * TOH;
    const int ow = min(TOW, p.out_w - ox0), oh = min(TOH, p.out_h - oy0);
    const int ylast = oy0 + oh - 1;
    const int ry0 = p.bounds_v[2 * oy0], rows_in = p.bounds_v[2 * ylast] + p.bounds_v[2 * ylast + 1] - ry0;
    const bool ux = ox0 >= p.ux0 && ox0 + ow <= p.ux1, uy = oy0 >= p.uy0 && oy0 + oh <= p.uy1;
    const uint8_t *in = p.in + (int64_t)page * p.in_stride + (int64_t)ry0 * p.in_pitch;
    uint8_t *out = p.out + (int64_t)page * p.out_stride + (int64_t)oy0 * p.out_pitch + (int64_t)ox0 * C;
    if (FOLLOW) {
        if (tid == 0) {
            const int xl = ox0 + ow - 1, c0 = p.bounds_h[2 * ox0], c1 = p.bounds_h[2 * xl] + p.bounds_h[2 * xl + 1] - 1;
            const int need = ry0 + rows_in;
            for (int s = c0 / p.prog_w; s <= c1 / p.prog_w; s++) {
                const int *q = p.prog + (int64_t)page * p.prog_s + s;
                while (ld_acquire(q) < need) __nanosleep(400);
            }
        }
        __syncthreads();                                      // thread 0's acquire orders every thread's reads of the input rows
    }

    auto generic_h = [&](int r, int oxl) {                    // one output pixel of input row ry0 + r
        const int ox = ox0 + oxl, xmin = p.bounds_h[2 * ox], cnt = p.bounds_h[2 * ox + 1];
        const int *k = p.kk_h + (size_t)ox * p.ksize_h;
        const uint8_t *q = in + (int64_t)r * p.in_pitch + (int64_t)xmin * C;
        int acc[C];
#pragma unroll
        for (int c = 0; c < C; c++) acc[c] = 1 << 21;
        for (int i = 0; i < cnt; i++) {
            const int kv = __ldg(k + i);
#pragma unroll
            for (int c = 0; c < C; c++) acc[c] += (int)q[i * C + c] * kv;
        }
#pragma unroll
        for (int c = 0; c < C; c++) hbuf[r * HROW + oxl * C + c] = clip8(acc[c]);
    };
    // ---- horizontal pass
    const int ngrp = ux ? ow / 4 : 0;
    if (ngrp > 0) {
        constexpr int NB = (3 * F + T) * C, NW = (NB + 3) / 4;
        for (int task = tid; task < rows_in * ngrp; task += 256) {
            const int r = task / ngrp, g = task - r * ngrp;
            const uint32_t *src = reinterpret_cast<const uint32_t *>(in + (int64_t)r * p.in_pitch + (int64_t)(F * (ox0 + 4 * g) + p.offx) * C);
            uint32_t w[NW];
#pragma unroll
            for (int i = 0; i < NW; i++) w[i] = FOLLOW ? src[i] : __ldg(src + i);   // FOLLOW: written during this kernel's lifetime, no non-coherent loads
            int acc[4][C];
#pragma unroll
            for (int px = 0; px < 4; px++)
#pragma unroll
                for (int c = 0; c < C; c++) acc[px][c] = 1 << 21;
#pragma unroll
            for (int px = 0; px < 4; px++)
#pragma unroll
                for (int t = 0; t < T; t++)
#pragma unroll
                    for (int c = 0; c < C; c++) {
                        const int b = (px * F + t) * C + c;
                        acc[px][c] += (int)__byte_perm(w[b >> 2], 0, 0x4440 | (b & 3)) * p.kh[t];   // one PRMT per byte
                    }
            uint32_t ow_[C];                                  // 4*C output bytes = C words
#pragma unroll
            for (int i = 0; i < C; i++) {
                uint32_t v = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) { const int b = 4 * i + j; v |= (uint32_t)clip8(acc[b / C][b % C]) << (8 * j); }
                ow_[i] = v;
            }
            uint32_t *dst = reinterpret_cast<uint32_t *>(hbuf + r * HROW + 4 * g * C);
#pragma unroll
            for (int i = 0; i < C; i++) dst[i] = ow_[i];
        }
    }
    {
        const int rest = ow - 4 * ngrp;                       // pixels not covered by whole groups
        for (int task = tid; task < rows_in * rest; task += 256) {
            const int r = task / rest;
            generic_h(r, 4 * ngrp + (task - r * rest));
        }
    }
    __syncthreads();
    // ---- vertical pass: one thread per 4 output bytes
    const int nb = ow * C, nwords = (nb + 3) / 4;
    for (int task = tid; task < oh * nwords; task += 256) {
        const int oy = task / nwords, b4 = 4 * (task - oy * nwords);
        int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21, a3 = 1 << 21;
        if (uy) {
            const uint8_t *q = hbuf + (F * (oy0 + oy) + p.offy - ry0) * HROW + b4;
#pragma unroll
            for (int t = 0; t < T; t++) {
                const uint32_t w = *reinterpret_cast<const uint32_t *>(q + t * HROW);
                const int kv = p.kv[t];
                a0 += (int)__byte_perm(w, 0, 0x4440) * kv; a1 += (int)__byte_perm(w, 0, 0x4441) * kv;
                a2 += (int)__byte_perm(w, 0, 0x4442) * kv; a3 += (int)(w >> 24) * kv;
            }
        } else {
            const int yy = oy0 + oy, cnt = p.bounds_v[2 * yy + 1];
            const int *k = p.kk_v + (size_t)yy * p.ksize_v;
            const uint8_t *q = hbuf + (p.bounds_v[2 * yy] - ry0) * HROW + b4;
            for (int i = 0; i < cnt; i++) {
                const uint32_t w = *reinterpret_cast<const uint32_t *>(q + i * HROW);
                const int kv = __ldg(k + i);
                a0 += (int)(w & 0xffu) * kv; a1 += (int)((w >> 8) & 0xffu) * kv; a2 += (int)((w >> 16) & 0xffu) * kv; a3 += (int)(w >> 24) * kv;
            }
        }
        const uint32_t o = (uint32_t)clip8(a0) | ((uint32_t)clip8(a1) << 8) | ((uint32_t)clip8(a2) << 16) | ((uint32_t)clip8(a3) << 24);
        uint8_t *dst = out + (int64_t)oy * p.out_pitch + b4;
        if (b4 + 3 < nb) *reinterpret_cast<uint32_t *>(dst) = o;
        else
            for (int j = 0; b4 + j < nb; j++) dst[j] = (uint8_t)(o >> (8 * j));
    }
}

// the interior of a coefficient table: outputs [lo, hi) whose window is bounds = F*i + off with the same T weights
void find_uniform(const std::vector<int> &bounds, const std::vector<int> &kk, int ksize, int outSize,
                  int &F, int &T, int &off, int &lo, int &hi, int *kout)
{
    F = T = off = lo = hi = 0;
    if (outSize < 4) return;
    const int mid = outSize / 2;
    T = bounds[2 * mid + 1];
    F = bounds[2 * (mid + 1)] - bounds[2 * mid];
    off = bounds[2 * mid] - F * mid;
    if (T < 1 || T > 16 || F < 1) { F = 0; return; }
    auto same = [&](int i) {
        if (bounds[2 * i] != F * i + off || bounds[2 * i + 1] != T) return false;
        for (int t = 0; t < T; t++) if (kk[(size_t)i * ksize + t] != kk[(size_t)mid * ksize + t]) return false;
        return true;
    };
    lo = mid; hi = mid + 1;
    while (lo > 0 && same(lo - 1)) lo--;
    while (hi < outSize && same(hi)) hi++;
    for (int t = 0; t < 16; t++) kout[t] = t < T ? kk[(size_t)mid * ksize + t] : 0;
}

struct ReduceParams {
    const uint8_t *in; int64_t in_pitch, in_stride;
    uint8_t *out; int64_t out_pitch, out_stride;
    int C, bx, by, bw, bh, fx, fy, out_w, out_h;
};

__device__ __forceinline__ uint32_t division_u32(int divider, int result_bits)
{
    // Pillow Reduce.c division_UINT32
    const uint32_t max_dividend = (1u << result_bits) * (uint32_t)divider;
    const float max_int = (1 << 30) * 4.0f;
    return (uint32_t)(max_int / max_dividend);
}

__global__ void __launch_bounds__(256) k_reduce_box(const ReduceParams p)
{
    const int b = blockIdx.x * 256 + threadIdx.x;
    const int oy = blockIdx.y, page = blockIdx.z;
    if (b >= p.out_w * p.C) return;
    const int ox = b / p.C, c = b - ox * p.C;
    const int x0 = ox * p.fx, x1 = min(p.bw, x0 + p.fx), y0 = oy * p.fy, y1 = min(p.bh, y0 + p.fy);
    const int cells = (y1 - y0) * (x1 - x0);
    const uint32_t mult = division_u32(cells, 8);
    uint32_t ss = (uint32_t)cells / 2;
    const uint8_t *src = p.in + (int64_t)page * p.in_stride;
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++)
            ss += src[(int64_t)(p.by + y) * p.in_pitch + (int64_t)(p.bx + x) * p.C + c];
    p.out[(int64_t)page * p.out_stride + (int64_t)oy * p.out_pitch + b] = (uint8_t)((ss * mult) >> 24);
}

int upload(const std::vector<int> &v, int **d)
{
    *d = nullptr;
    if (v.empty()) return 0;
    cudaError_t e = cudaMalloc((void **)d, v.size() * sizeof(int));
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpy(*d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice);
    return (int)e;
}

}  // namespace
}  // namespace b200mrc

namespace b200mrc {

static bool tile_form_ok(const b200mrc_resample_plan *pl, const uint8_t *src, int64_t src_pitch, int64_t src_stride,
                         const uint8_t *out, int64_t out_pitch, int64_t out_page_stride)
{
    return pl->tile_F == 3 && pl->tile_T == 12 && !(src_pitch & 3) && !(src_stride & 3) && !((uintptr_t)src & 3) &&
           !(out_pitch & 3) && !(out_page_stride & 3) && !((uintptr_t)out & 3) && cdiv(pl->OH, 16) <= 65535 && !tune(T_RESAMPLE_2PASS);
}

template <int C, bool FOLLOW>
static const void *tile_kernel(int toh)
{
    return toh == 64 ? (const void *)k_resample_tile<C, 3, 12, 64, FOLLOW>
         : toh == 16 ? (const void *)k_resample_tile<C, 3, 12, 16, FOLLOW> : (const void *)k_resample_tile<C, 3, 12, 32, FOLLOW>;
}

// prog != null: the FOLLOW form, launched as the programmatic dependent of the kernel before it in `st`
static int launch_tile(const b200mrc_resample_plan *pl, const uint8_t *src, int64_t src_pitch, int64_t src_stride,
                       uint8_t *out, int64_t out_pitch, int64_t out_page_stride, int n_pages,
                       const int *prog, int prog_s, int prog_w, cudaStream_t st)
{
    TileParams t;
    t.in = src; t.in_pitch = src_pitch; t.in_stride = src_stride; t.out = out; t.out_pitch = out_pitch; t.out_stride = out_page_stride;
    t.in_w = pl->SW; t.in_h = pl->SH; t.out_w = pl->OW; t.out_h = pl->OH; t.ksize_h = pl->ksize_h; t.ksize_v = pl->ksize_v;
    t.bounds_h = pl->d_bounds_h; t.kk_h = pl->d_kk_h; t.bounds_v = pl->d_bounds_v; t.kk_v = pl->d_kk_v;
    t.offx = pl->offx; t.offy = pl->offy; t.ux0 = pl->ux0; t.ux1 = pl->ux1; t.uy0 = pl->uy0; t.uy1 = pl->uy1;
    memcpy(t.kh, pl->kh, sizeof(t.kh)); memcpy(t.kv, pl->kv, sizeof(t.kv));
    int toh = tune(T_TILE_H);
    toh = toh == 64 ? 64 : (toh == 16 ? 16 : 32);
    const int ntx = cdiv(pl->OW, TOW), nty = cdiv(pl->OH, toh);
    t.prog = prog; t.prog_s = prog_s; t.prog_w = prog_w; t.ntx = ntx; t.npages = n_pages;
    void *args[] = {(void *)&t};
    ProfScope _ps("k_resample_tile", st);
    if (!prog) {
        const void *kern = pl->C == 1 ? tile_kernel<1, false>(toh) : tile_kernel<3, false>(toh);
        B200MRC_CUDA_TRY(cudaLaunchKernel(kern, dim3(ntx, nty, n_pages), dim3(256), args, 0, st));
    } else {
        if ((int64_t)ntx * nty * n_pages > 0x7fffffffll) return B200MRC_ERR_UNSUPPORTED;
        const void *kern = pl->C == 1 ? tile_kernel<1, true>(toh) : tile_kernel<3, true>(toh);
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = 1;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(ntx * nty * n_pages)); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        B200MRC_CUDA_TRY(cudaLaunchKernelExC(&cfg, kern, args));
    }
    count_launch();
    return B200MRC_OK;
}

// The thumbnail of images a row-sequential kernel is still writing: that kernel (the one launched right before this call
// in `stream`) publishes per page and per strip of prog_w columns the rows that are complete, and lets its dependents
// start early (pdl_launch_dependents).  UNSUPPORTED when the plan does not take the tile form (call b200mrc_resample).
int launch_resample_follow(const b200mrc_resample_plan *pl, const uint8_t *in, int64_t in_pitch, int64_t in_page_stride,
                           uint8_t *out, int64_t out_pitch, int64_t out_page_stride, int n_pages,
                           const int *prog, int prog_s, int prog_w, cudaStream_t st)
{
    if (!pl || !in || !out || !prog || n_pages <= 0 || prog_w <= 0) return B200MRC_ERR_INVALID;
    if (pl->fx > 1 || pl->fy > 1 || !pl->need_h || !pl->need_v) return B200MRC_ERR_UNSUPPORTED;
    if (!tile_form_ok(pl, in, in_pitch, in_page_stride, out, out_pitch, out_page_stride)) return B200MRC_ERR_UNSUPPORTED;
    return launch_tile(pl, in, in_pitch, in_page_stride, out, out_pitch, out_page_stride, n_pages, prog, prog_s, prog_w, st);
}

}  // namespace b200mrc

using namespace b200mrc;

extern "C" b200mrc_resample_plan *b200mrc_thumbnail_plan_create(int width, int height, int channels,
                                                                double req_width, double req_height,
                                                                double reducing_gap, int filter, int *status)
{
    int dummy;
    if (!status) status = &dummy;
    *status = B200MRC_OK;
    if (width <= 0 || height <= 0 || (channels != 1 && channels != 3) || (filter != 0 && filter != 1)) {
        *status = B200MRC_ERR_INVALID; return nullptr;
    }
    // PIL Image.thumbnail (Image.py): provided_size = floor(size); preserve_aspect_ratio()
    const double px = std::floor(req_width), py = std::floor(req_height);
    if (!(px >= 1) || !(py >= 1)) { *status = B200MRC_ERR_INVALID; return nullptr; }
    if (px >= width && py >= height) return nullptr;                 // nothing to do
    if ((double)height > (double)width * 100) { *status = B200MRC_ERR_UNSUPPORTED; return nullptr; }  // Pillow's tall-image 2-step path
    auto round_aspect = [](double number, auto key) {
        const double f = std::floor(number), c = std::ceil(number);
        // python: max(min(floor, ceil, key=key), 1) -- min returns the first on ties
        double best = key(f) <= key(c) ? f : c;
        return best < 1 ? 1.0 : best;
    };
    const double aspect = (double)width / height;
    double x = px, y = py;
    if (x / y >= aspect) x = round_aspect(y * aspect, [&](double n) { return std::fabs(aspect - n / y); });
    else y = round_aspect(x / aspect, [&](double n) { return n == 0 ? 0.0 : std::fabs(aspect - x / n); });
    const int OW = (int)x, OH = (int)y;
    if (OW == width && OH == height) return nullptr;

    b200mrc_resample_plan *pl = new b200mrc_resample_plan();
    pl->W = width; pl->H = height; pl->C = channels; pl->OW = OW; pl->OH = OH; pl->filter = filter;
    pl->fx = pl->fy = 1; pl->rx0 = pl->ry0 = 0; pl->rx1 = width; pl->ry1 = height;
    double b0 = 0, b1 = 0, b2 = width, b3 = height;
    if (reducing_gap > 0) {
        // Image.resize (Image.py): factor = int(box_size / size / reducing_gap) or 1
        int fx = (int)((b2 - b0) / OW / reducing_gap); if (fx < 1) fx = 1;
        int fy = (int)((b3 - b1) / OH / reducing_gap); if (fy < 1) fy = 1;
        if (fx > 1 || fy > 1) {
            // _get_safe_box
            const double support = (filter == 1 ? 3.0 : 2.0) - 0.5;
            const double sx = (b2 - b0) / OW * support, sy = (b3 - b1) / OH * support;
            pl->rx0 = std::max(0, (int)(b0 - sx)); pl->ry0 = std::max(0, (int)(b1 - sy));
            pl->rx1 = std::min(width, (int)std::ceil(b2 + sx)); pl->ry1 = std::min(height, (int)std::ceil(b3 + sy));
            pl->fx = fx; pl->fy = fy;
            b0 = (b0 - pl->rx0) / fx; b1 = (b1 - pl->ry0) / fy;
            b2 = (b2 - pl->rx0) / fx; b3 = (b3 - pl->ry0) / fy;
        }
    }
    pl->SW = (pl->rx1 - pl->rx0 + pl->fx - 1) / pl->fx;
    pl->SH = (pl->ry1 - pl->ry0 + pl->fy - 1) / pl->fy;
    if (pl->fx == 1 && pl->fy == 1) { pl->SW = width; pl->SH = height; }
    pl->bx0 = (float)b0; pl->by0 = (float)b1; pl->bx1 = (float)b2; pl->by1 = (float)b3;
    // ImagingResampleInner
    pl->need_h = OW != pl->SW || pl->bx0 != 0 || pl->bx1 != (float)pl->SW;
    pl->need_v = OH != pl->SH || pl->by0 != 0 || pl->by1 != (float)pl->SH;
    std::vector<int> bh, kh, bv, kv;
    pl->ksize_h = build_coeffs(pl->SW, pl->bx0, pl->bx1, OW, filter, bh, kh);
    pl->ksize_v = build_coeffs(pl->SH, pl->by0, pl->by1, OH, filter, bv, kv);
    {
        int Fh, Th, Fv, Tv;
        find_uniform(bh, kh, pl->ksize_h, OW, Fh, Th, pl->offx, pl->ux0, pl->ux1, pl->kh);
        find_uniform(bv, kv, pl->ksize_v, OH, Fv, Tv, pl->offy, pl->uy0, pl->uy1, pl->kv);
        const bool ok = pl->fx == 1 && pl->fy == 1 && pl->need_h && pl->need_v && Fh == Fv && Th == Tv && Fh == 3 && Th == 12 &&
                        ((pl->offx * channels) & 3) == 0 && pl->ux1 - pl->ux0 >= TOW && pl->uy1 - pl->uy0 >= 16;
        pl->tile_F = ok ? Fh : 0; pl->tile_T = ok ? Th : 0;
    }
    int rc = upload(bh, &pl->d_bounds_h);
    if (!rc) rc = upload(kh, &pl->d_kk_h);
    if (!rc) rc = upload(bv, &pl->d_bounds_v);
    if (!rc) rc = upload(kv, &pl->d_kk_v);
    if (rc) { *status = rc; b200mrc_resample_plan_destroy(pl); return nullptr; }
    return pl;
}

extern "C" void b200mrc_resample_plan_destroy(b200mrc_resample_plan *pl)
{
    if (!pl) return;
    cudaFree(pl->d_bounds_h); cudaFree(pl->d_kk_h); cudaFree(pl->d_bounds_v); cudaFree(pl->d_kk_v);
    delete pl;
}

extern "C" void b200mrc_resample_plan_out_size(const b200mrc_resample_plan *pl, int *out_width, int *out_height)
{
    if (out_width) *out_width = pl ? pl->OW : 0;
    if (out_height) *out_height = pl ? pl->OH : 0;
}

static size_t resample_red_bytes(const b200mrc_resample_plan *pl)
{
    return (pl->fx > 1 || pl->fy > 1) ? align_up((size_t)pl->SW * pl->C, 16) * pl->SH : 0;
}
static size_t resample_tmp_bytes(const b200mrc_resample_plan *pl)
{
    return (pl->need_h && pl->need_v) ? align_up((size_t)pl->OW * pl->C, 16) * pl->SH : 0;
}

extern "C" size_t b200mrc_resample_workspace_bytes(const b200mrc_resample_plan *pl, int n_pages)
{
    if (!pl || n_pages <= 0) return 0;
    return align_up(resample_red_bytes(pl) * n_pages, 256) + align_up(resample_tmp_bytes(pl) * n_pages, 256) + 256;
}

extern "C" int b200mrc_resample(const b200mrc_resample_plan *pl,
                                const uint8_t *in, int64_t in_pitch, int64_t in_page_stride,
                                uint8_t *out, int64_t out_pitch, int64_t out_page_stride, int n_pages,
                                void *workspace, size_t workspace_bytes, void *stream)
{
    if (!pl || !in || !out || n_pages <= 0) return B200MRC_ERR_INVALID;
    if (n_pages > 65535 || pl->SH > 65535 * H_RPT || pl->OH > 65535) return B200MRC_ERR_UNSUPPORTED;
    if (workspace_bytes < b200mrc_resample_workspace_bytes(pl, n_pages) || (!workspace && workspace_bytes)) return B200MRC_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *ws = (uint8_t *)workspace;
    const size_t red_page = resample_red_bytes(pl), tmp_page = resample_tmp_bytes(pl);
    uint8_t *red = ws, *tmp = ws + align_up(red_page * n_pages, 256);

    const uint8_t *src = in; int64_t src_pitch = in_pitch, src_stride = in_page_stride;
    if (red_page) {
        ReduceParams r;
        r.in = in; r.in_pitch = in_pitch; r.in_stride = in_page_stride;
        r.out = red; r.out_pitch = (int64_t)align_up((size_t)pl->SW * pl->C, 16); r.out_stride = (int64_t)red_page;
        r.C = pl->C; r.bx = pl->rx0; r.by = pl->ry0; r.bw = pl->rx1 - pl->rx0; r.bh = pl->ry1 - pl->ry0;
        r.fx = pl->fx; r.fy = pl->fy; r.out_w = pl->SW; r.out_h = pl->SH;
        dim3 grid(cdiv(pl->SW * pl->C, 256), pl->SH, n_pages);
        { ProfScope _ps("k_reduce_box", st); k_reduce_box<<<grid, 256, 0, st>>>(r); }
        B200MRC_LAUNCH_CHECK();
        src = red; src_pitch = r.out_pitch; src_stride = r.out_stride;
    }
    if (!pl->need_h && !pl->need_v) {
        for (int n = 0; n < n_pages; n++)
            B200MRC_CUDA_TRY(cudaMemcpy2DAsync(out + (int64_t)n * out_page_stride, out_pitch, src + (int64_t)n * src_stride, src_pitch,
                                               (size_t)pl->OW * pl->C, (size_t)pl->OH, cudaMemcpyDeviceToDevice, st));
        return B200MRC_OK;
    }
    if (tile_form_ok(pl, src, src_pitch, src_stride, out, out_pitch, out_page_stride))
        return launch_tile(pl, src, src_pitch, src_stride, out, out_pitch, out_page_stride, n_pages, nullptr, 0, 0, st);
    if (pl->need_h) {
        ResampleParams h;
        h.in = src; h.in_pitch = src_pitch; h.in_stride = src_stride;
        if (pl->need_v) { h.out = tmp; h.out_pitch = (int64_t)align_up((size_t)pl->OW * pl->C, 16); h.out_stride = (int64_t)tmp_page; }
        else { h.out = out; h.out_pitch = out_pitch; h.out_stride = out_page_stride; }
        h.C = pl->C; h.in_w = pl->SW; h.in_h = pl->SH; h.out_w = pl->OW; h.out_h = pl->SH; h.ksize = pl->ksize_h;
        h.bounds = pl->d_bounds_h; h.kk = pl->d_kk_h;
        dim3 grid(cdiv(pl->OW, 128), cdiv(pl->SH, H_RPT), n_pages);
        { ProfScope _ps("k_resample_h", st);
          if (pl->C == 1) k_resample_h<1><<<grid, 128, 0, st>>>(h);
          else k_resample_h<3><<<grid, 128, 0, st>>>(h); }
        B200MRC_LAUNCH_CHECK();
        src = h.out; src_pitch = h.out_pitch; src_stride = h.out_stride;
    }
    if (pl->need_v) {
        ResampleParams v;
        v.in = src; v.in_pitch = src_pitch; v.in_stride = src_stride;
        v.out = out; v.out_pitch = out_pitch; v.out_stride = out_page_stride;
        v.C = pl->C; v.in_w = pl->OW; v.in_h = pl->SH; v.out_w = pl->OW; v.out_h = pl->OH; v.ksize = pl->ksize_v;
        v.bounds = pl->d_bounds_v; v.kk = pl->d_kk_v;
        if ((src_pitch & 3) || (src_stride & 3) || ((uintptr_t)src & 3)) return B200MRC_ERR_ALIGNMENT;
        dim3 grid(cdiv(cdiv(pl->OW * pl->C, 4), 256), pl->OH, n_pages);
        { ProfScope _ps("k_resample_v", st); k_resample_v<<<grid, 256, 0, st>>>(v); }
        B200MRC_LAUNCH_CHECK();
    }
    return B200MRC_OK;
}
