// sauvola.cu -- k_sauvola_mask_w: Sauvola local-adaptive threshold on a gray plane, any 4-byte aligned rows (the general
// form: small crops, k < 0, inverted input; 16-byte aligned planes take the fused kernel of sauvola_fused.cu).  Sauvola local-adaptive threshold (reference: binarise_sauvola,
// cython/sauvola.pyx:29-222, called through threshold_image, internetarchivepdf/mrc.py:58-87).
//
// Semantics (closed form of the reference's running sums, see oracle/mrc_oracle.c orc_sauvola):
//   window rows [max(0,y-o+1), min(H,y+u+1)), cols [max(0,x-l+1), min(W,x+r+1)),
//   l=(ww+1)/2, r=ww/2, o=(wh+1)/2, u=wh/2;  n = rows*cols
//   m = (double)(S/n), v = (double)(Q/n) - m*m  (integer divisions, sauvola.pyx:144-145)
//   t = p + m*(k-1);  fg = t<=0 || t*t <= ((m*m)*k2)*v          (sauvola.pyx:146-147)
//   every double operation individually rounded (no FMA): the reference build is x86-64 SSE2.
//
// B200 mapping: one CTA marches a (column strip x row band) of one page top to bottom.
//   * each of the 256 threads owns 4 adjacent input columns and keeps their running column sums
//     (sum, sum of squares over the window's rows) in registers; per output row it adds the
//     entering row and subtracts the leaving row (32-bit vector loads, L2-resident re-reads);
//   * the horizontal window sum is a difference of two entries of the per-row prefix of the
//     column sums: thread-local prefix -> warp-shuffle inclusive scan; prefixes are published in a
//     bank-conflict-free SoA layout (column c -> [c&3][c>>2]); the published prefixes are warp-local and a
//     pixel adds the totals of the warps between its two entries -> one CTA barrier per row;
//   * uint32 wrap-around arithmetic is exact because every window sum is < 2^32 for w <= 255;
//   * S/n and Q/n are exact floors computed on the FP64 pipe (floor((a+0.5)*(1/n))); the test runs in
//     FP64 with __dmul_rn/__dadd_rn so nvcc cannot contract to FMA.
// Algorithmic HBM bytes: 1 B/px read + 1 B/px written.
#include "common.cuh"
#include <cstdlib>
#include <cstring>

namespace b200mrc {

namespace {

constexpr int ST = 256;          // threads per CTA
constexpr int SK = 4;            // columns per thread
constexpr int SE = ST * SK;      // input columns covered by one CTA (strip + window halo)

struct SauvolaParams {
    const uint8_t *in; int64_t in_pitch, in_stride;
    uint8_t *out;      int64_t out_pitch, out_stride;
    int W, H;
    int l, r, o, u;
    int n_strips, strip_w, ext_left, n_bands, band_h;
    double km1, k2;
    int kneg, flags;
};

__device__ __forceinline__ uint32_t load_word_clamped(const uint8_t *row, int gx, int W, uint32_t inv)
{
    // 4 pixels gx..gx+3 (gx % 4 == 0), optionally inverted (255 - p); pixels outside [0, W) read as 0
    if (gx < 0 || gx >= W) return 0u;
    uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(row + gx)) ^ inv;
    int valid = W - gx;                  // >= 1
    if (valid < 4) w &= (1u << (8 * valid)) - 1u;
    return w;
}

// ---- warp-local prefixes, ONE barrier per row ---------------------------------------------------------------
// A CTA-wide prefix would cost two barriers and 8 warp totals folded in by every thread per row.  Here every
// warp publishes the prefix of ITS OWN 128 columns (L) plus its total (WT); a window [clo, chi) spans at most three
// warps (window <= 255), so   S = L[chi] - L[clo] + WT[wlo] (+ WT[wlo+1]),   the WT indices being row-invariant per
// pixel (an all-zero entry stands for "same warp").  The buffers alternate with the row, the row loop is unrolled
// by two so that the buffer is a compile-time offset, and the scan steps use the shuffle's own range predicate.
constexpr int SPN = SK * (ST + 1);       // prefix entries per buffer
constexpr int WTN = 16;                  // warp-total entries per buffer: [0..7] totals, [8] = 0
template <int V> struct IntC { static constexpr int value = V; };

__device__ __forceinline__ void scan_up2(uint32_t &a, uint32_t &b, int d)
{
    // a += shfl_up(a, d), b += shfl_up(b, d) for lanes >= d (the shuffle's predicate says whether the source lane exists)
    asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 t, u;\n\t"
        "shfl.sync.up.b32 t|p, %0, %2, 0, 0xffffffff;\n\t"
        "shfl.sync.up.b32 u, %1, %2, 0, 0xffffffff;\n\t"
        "@p add.u32 %0, %0, t;\n\t"
        "@p add.u32 %1, %1, u;\n\t}"
        : "+r"(a), "+r"(b) : "r"(d));
}

template <bool KNEG, bool WIDE>
__device__ __forceinline__ void sauvola_mask_body(const SauvolaParams &p, const int bx, const int page, uint2 *sP, uint2 *sWT)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int strip = bx % p.n_strips, band = bx / p.n_strips;
    const int sx0 = strip * p.strip_w;
    const int ex0 = sx0 - p.ext_left;             // first input column of this CTA (multiple of 4)
    const int by0 = band * p.band_h;
    const int by1 = min(p.H, by0 + p.band_h);
    const int W = p.W, H = p.H;
    const uint8_t *in = p.in + (int64_t)page * p.in_stride;
    uint8_t *out = p.out + (int64_t)page * p.out_stride;

    const int i0 = tid * SK;                      // local column of this thread's first pixel
    const int gx = ex0 + i0;                      // global column (multiple of 4, may be < 0 or >= W)
    const bool is_out = (gx >= sx0) && (gx < sx0 + p.strip_w) && (gx < W);

    const uint32_t inv = (p.flags & B200MRC_SAUVOLA_INVERT_INPUT) ? 0xffffffffu : 0u;   // threshold 255 - p (mrc.py:226)
    uint32_t cs[SK], cq[SK];
#pragma unroll
    for (int j = 0; j < SK; j++) { cs[j] = 0; cq[j] = 0; }
    // row-invariant per-thread constants: smem slots of the two prefix entries of each pixel's window and of the
    // warp totals between them, the window width, and where this thread publishes its own prefixes
    int slot_hi[SK], slot_lo[SK], slot_st[SK];
    uint32_t wtA = 0, wtB = 0;                    // byte offsets of the 4 pixels' warp-total entries, one per byte
    auto nx_of = [&](const int j) { const int x = gx + j; return (x < W) ? min(W, x + p.r + 1) - max(0, x - p.l + 1) : 0; };
#pragma unroll
    for (int j = 0; j < SK; j++) {
        const int i = i0 + j;
        const int chi = i + p.r + 1, clo = max(i - p.l + 1, 0), cst = i + 1;
        slot_hi[j] = (chi & 3) * (ST + 1) + (chi >> 2);
        slot_lo[j] = (clo & 3) * (ST + 1) + (clo >> 2);
        slot_st[j] = (cst & 3) * (ST + 1) + (cst >> 2);
        const int whi = (chi - 1) >> 7, wlo = clo > 0 ? (clo - 1) >> 7 : 0;      // warps holding columns chi-1 / clo-1
        wtA |= (uint32_t)(min(whi > wlo ? wlo : 8, 8) * 8) << (8 * j);
        wtB |= (uint32_t)(min(whi > wlo + 1 ? wlo + 1 : 8, 8) * 8) << (8 * j);
    }
    const int nx0 = nx_of(0);
    const bool nx_uniform = nx0 == nx_of(1) && nx0 == nx_of(2) && nx0 == nx_of(3);

    // ---- band warm-up: column sums over the window rows of the first output row
    {
        const int r0 = max(0, by0 - p.o + 1), r1 = min(H - 1, by0 + p.u);
        for (int yy = r0; yy <= r1; yy++) {
            uint32_t w = load_word_clamped(in + (int64_t)yy * p.in_pitch, gx, W, inv);
#pragma unroll
            for (int j = 0; j < SK; j++) {
                uint32_t v = (w >> (8 * j)) & 0xFFu;
                cs[j] += v; cq[j] += v * v;
            }
        }
    }
    if (tid < 2) {
        // L[0] = 0 (slot of column boundary 0, written by nobody else) and the all-zero warp total
        sP[tid * SPN] = make_uint2(0u, 0u);
        sWT[tid * WTN + 8] = make_uint2(0u, 0u);
    }

    // loads run two rows ahead of their use (L2/HBM latency >> one row step): stage A feeds the next iteration, stage B
    // the one after.  Words are loaded raw through three running row pointers and masked (columns outside the page read
    // as 0) only when they are consumed, so nothing waits on a load in the iteration that issues it.
    const bool lok = gx >= 0 && gx < W;
    const uint32_t lmask = !lok ? 0u : (W - gx >= 4 ? 0xffffffffu : (1u << (8 * (W - gx))) - 1u);
    const uint8_t *pe = in + (int64_t)(by0 + 1 + p.u) * p.in_pitch + gx;      // entering row of the next load_upd
    const uint8_t *pl = in + (int64_t)(by0 + 1 - p.o) * p.in_pitch + gx;      // leaving row
    const uint8_t *pc = in + (int64_t)(by0 + 1) * p.in_pitch + gx;            // its own pixels
    auto load_upd = [&](int yu, uint32_t &we, uint32_t &wl, uint32_t &wc) {
        we = 0; wl = 0; wc = 0;
        if (yu < by1 && lok) {
            if (yu + p.u < H) we = __ldg(reinterpret_cast<const uint32_t *>(pe));
            if (yu - p.o >= 0) wl = __ldg(reinterpret_cast<const uint32_t *>(pl));
            wc = __ldg(reinterpret_cast<const uint32_t *>(pc));
        }
        pe += p.in_pitch; pl += p.in_pitch; pc += p.in_pitch;
    };
    uint32_t wcur = load_word_clamped(in + (int64_t)by0 * p.in_pitch, gx, W, inv);
    uint32_t weA, wlA, wcA, weB, wlB, wcB;
    load_upd(by0 + 1, weA, wlA, wcA);
    load_upd(by0 + 2, weB, wlB, wcB);
    int ny_cached = -1;
    double rn_u = 0.0;
    uint8_t *orow = out + (int64_t)by0 * p.out_pitch + gx;

    auto step = [&](auto BUFC, const int y) {
        constexpr int B = decltype(BUFC)::value;
        uint2 *P = sP + B * SPN, *T = sWT + B * WTN;
        // rows outside the page were not loaded (0) and must stay 0 under inversion
        const uint32_t inv_e = (y + 1 + p.u < H && y + 1 < by1) ? inv : 0u, inv_l = (y + 1 - p.o >= 0 && y + 1 < by1) ? inv : 0u,
                       inv_c = (y + 1 < by1) ? inv : 0u;
        const uint32_t wenter = (weA ^ inv_e) & lmask, wleave = (wlA ^ inv_l) & lmask, wnext = (wcA ^ inv_c) & lmask;
        weA = weB; wlA = wlB; wcA = wcB;
        load_upd(y + 3, weB, wlB, wcB);

        // ---- prefix of the column sums across the warp
        uint32_t ps[SK], pq[SK];
        ps[0] = cs[0]; pq[0] = cq[0];
#pragma unroll
        for (int j = 1; j < SK; j++) { ps[j] = ps[j - 1] + cs[j]; pq[j] = pq[j - 1] + cq[j]; }
        uint32_t ws = ps[SK - 1], wq = pq[SK - 1];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) scan_up2(ws, wq, d);
        if (lane == 31) T[warp] = make_uint2(ws, wq);
        const uint32_t bs = ws - ps[SK - 1], bq = wq - pq[SK - 1];     // exclusive within the warp
        // inclusive warp-local prefix up to local column c-1 is stored at "c": column c -> [c&3][c>>2]
#pragma unroll
        for (int j = 0; j < SK; j++) P[slot_st[j]] = make_uint2(bs + ps[j], bq + pq[j]);
        __syncthreads();

        // ---- per-pixel test
        if (is_out) {
            const int ny = min(H, y + p.u + 1) - max(0, y - p.o + 1);
            uint32_t bits = 0;
            // S/n and Q/n on the (otherwise idle) FP64 pipe: floor((a + 0.5) * (1/n)) == a / n exactly, because (a + 0.5)/n is
            // at least 1/(2n) away from every integer while the double product is accurate to 2^-52 relative (a < 2^32,
            // n <= 65025).  The quotients come out as the doubles the test needs.
            if (nx_uniform && ny != ny_cached) { rn_u = 1.0 / (double)(nx0 * ny); ny_cached = ny; }
            uint32_t S[SK], Q[SK];
#pragma unroll
            for (int j = 0; j < SK; j++) {
                const uint2 hi = P[slot_hi[j]];
                const uint2 lo = P[slot_lo[j]];
                const uint2 ta = *reinterpret_cast<const uint2 *>(reinterpret_cast<const uint8_t *>(T) + ((wtA >> (8 * j)) & 0xFFu));
                S[j] = hi.x - lo.x + ta.x; Q[j] = hi.y - lo.y + ta.y;
                if (WIDE) {
                    const uint2 tb = *reinterpret_cast<const uint2 *>(reinterpret_cast<const uint8_t *>(T) + ((wtB >> (8 * j)) & 0xFFu));
                    S[j] += tb.x; Q[j] += tb.y;
                }
            }
            auto test = [&](const uint32_t s, const uint32_t q, const int j, const double rn) -> uint32_t {
                const double md = floor(__dmul_rn(__dadd_rn((double)s, 0.5), rn));      // (double)(S / n)
                const double qd = floor(__dmul_rn(__dadd_rn((double)q, 0.5), rn));      // (double)(Q / n)
                const double mm = __dmul_rn(md, md);                                    // exact (integers < 2^53)
                const double v = __dadd_rn(qd, -mm);
                const double pix = (double)((wcur >> (8 * j)) & 0xFFu);
                const double t = __dadd_rn(pix, __dmul_rn(md, p.km1));
                const double rhs = __dmul_rn(__dmul_rn(mm, p.k2), v);
                const double lhs = __dmul_rn(t, t);
                return KNEG ? ((t <= 0.0) && (lhs >= rhs)) : ((t <= 0.0) || (lhs <= rhs));
            };
            if (nx_uniform) {
                // interior thread: one window area for its 4 pixels (n > 0 because the thread is inside the page)
#pragma unroll
                for (int j = 0; j < SK; j++) bits |= test(S[j], Q[j], j, rn_u) << (8 * j);
            } else {
#pragma unroll
                for (int j = 0; j < SK; j++) {
                    const int n = nx_of(j) * ny;
                    if (n > 0) bits |= test(S[j], Q[j], j, 1.0 / (double)n) << (8 * j);
                }
            }
            if (p.flags & B200MRC_SAUVOLA_RAW_INVERTED) bits ^= 0x01010101u;
            if (gx + 3 < W) {
                uint32_t *o32 = reinterpret_cast<uint32_t *>(orow);
                if (p.flags & B200MRC_SAUVOLA_OR_INTO) bits |= *o32;
                *o32 = bits;
            } else {
                for (int j = 0; j < SK && gx + j < W; j++) {
                    uint8_t b = (uint8_t)((bits >> (8 * j)) & 0xFFu);
                    if (p.flags & B200MRC_SAUVOLA_OR_INTO) b |= orow[j];
                    orow[j] = b;
                }
            }
        }
        orow += p.out_pitch;

        // ---- slide the window rows for the next output row
#pragma unroll
        for (int j = 0; j < SK; j++) {
            const uint32_t a = (wenter >> (8 * j)) & 0xFFu, b = (wleave >> (8 * j)) & 0xFFu;
            cs[j] += a - b;
            cq[j] += a * a - b * b;
        }
        wcur = wnext;
    };

    int y = by0;
    for (; y + 1 < by1; y += 2) { step(IntC<0>(), y); step(IntC<1>(), y + 1); }
    if (y < by1) step(IntC<0>(), y);
}

template <bool KNEG, bool WIDE, int MINB>
__global__ void __launch_bounds__(ST, MINB) k_sauvola_mask_w(const SauvolaParams p)
{
    __shared__ uint2 sP[2 * SPN];
    __shared__ uint2 sWT[2 * WTN];
    sauvola_mask_body<KNEG, WIDE>(p, (int)blockIdx.x, (int)blockIdx.y, sP, sWT);
}

// strip / band geometry of an image of the given width (shared by the launcher and the items kernel)
__host__ __device__ __forceinline__ void sauvola_geometry(SauvolaParams &p, int width, int height)
{
    p.W = width; p.H = height;
    p.ext_left = (p.l - 1 + 3) / 4 * 4;
    const int sw_max = (SE - p.ext_left - p.r) / 4 * 4;          // >= 768
    p.n_strips = (width + sw_max - 1) / sw_max;
    p.strip_w = ((width + p.n_strips - 1) / p.n_strips + 3) / 4 * 4;
    p.band_h = 128;
    p.n_bands = (height + p.band_h - 1) / p.band_h;
}

// One launch over a device array of independent images (the text-line crops of create_hocr_mask, mrc.py:226-236):
// blockIdx.y = item, blockIdx.x = (strip, band) of that item; CTAs beyond an item's own grid leave at once.
template <bool KNEG, bool WIDE>
__global__ void __launch_bounds__(ST, 4) k_sauvola_items(const b200mrc_sauvola_item *items, const SauvolaParams base)
{
    __shared__ uint2 sP[2 * SPN];
    __shared__ uint2 sWT[2 * WTN];
    const b200mrc_sauvola_item it = items[blockIdx.y];
    SauvolaParams p = base;
    p.in = it.in; p.in_pitch = it.in_pitch; p.in_stride = 0;
    p.out = it.out; p.out_pitch = it.out_pitch; p.out_stride = 0;
    p.flags = it.flags;
    sauvola_geometry(p, it.width, it.height);
    if ((int)blockIdx.x >= p.n_strips * p.n_bands) return;
    sauvola_mask_body<KNEG, WIDE>(p, (int)blockIdx.x, 0, sP, sWT);
}

}  // namespace

bool sauvola_fused_ok(const uint8_t *src, int64_t src_pitch, int64_t src_stride, int C, const uint8_t *out, int64_t out_pitch,
                      int64_t out_stride, int W, int H, int ww, int wh, double k, int flags);
int launch_sauvola_fused(const uint8_t *src, int64_t src_pitch, int64_t src_stride, int C,
                         uint8_t *gray, int64_t gray_pitch, int64_t gray_stride,
                         uint8_t *out, int64_t out_pitch, int64_t out_stride,
                         int W, int H, int N, int ww, int wh, double k, double Rr, const double *sigma, int flags, cudaStream_t st);
bool threshold_path_legacy();

}  // namespace b200mrc

using namespace b200mrc;

extern "C" int b200mrc_sauvola(const uint8_t *in, int64_t in_pitch, int64_t in_page_stride,
                               uint8_t *out, int64_t out_pitch, int64_t out_page_stride,
                               int width, int height, int n_pages,
                               int window_width, int window_height, double k, double R,
                               int flags, void *stream)
{
    if (!in || !out || width <= 0 || height <= 0 || n_pages <= 0 || !(R > 0)) return B200MRC_ERR_INVALID;
    if (window_width < 1 || window_height < 1 || window_width > B200MRC_MAX_WINDOW ||
        window_height > B200MRC_MAX_WINDOW)
        return B200MRC_ERR_UNSUPPORTED;
    if (n_pages > 65535) return B200MRC_ERR_UNSUPPORTED;
    if ((in_pitch & 3) || (out_pitch & 3) || ((uintptr_t)in & 3) || ((uintptr_t)out & 3) ||
        (in_page_stride & 3) || (out_page_stride & 3) || in_pitch < width || out_pitch < width)
        return B200MRC_ERR_ALIGNMENT;

    // rows that TMA can move: the 8-columns-per-thread kernel of sauvola_fused.cu in its direct (gray in) form
    if (!threshold_path_legacy() && n_pages <= 65535 &&
        sauvola_fused_ok(in, in_pitch, in_page_stride, 1, out, out_pitch, out_page_stride, width, height, window_width, window_height, k, flags))
        return launch_sauvola_fused(in, in_pitch, in_page_stride, 1, nullptr, 0, 0, out, out_pitch, out_page_stride, width, height, n_pages,
                                    window_width, window_height, k, R, nullptr, flags, (cudaStream_t)stream);

    SauvolaParams p;
    p.in = in; p.in_pitch = in_pitch; p.in_stride = in_page_stride;
    p.out = out; p.out_pitch = out_pitch; p.out_stride = out_page_stride;
    p.l = (window_width + 1) / 2;  p.r = window_width / 2;
    p.o = (window_height + 1) / 2; p.u = window_height / 2;
    sauvola_geometry(p, width, height);
    p.km1 = k - 1.0;
    p.k2 = k * k / R / R;                                         // sauvola.pyx:60
    p.kneg = k < 0;
    p.flags = flags;

    dim3 grid((unsigned)(p.n_strips * p.n_bands), (unsigned)n_pages);
    { ProfScope _ps("k_sauvola_mask", (cudaStream_t)stream);
      const bool wide = window_width > 128;                        // a window may then span three warps
      if (p.kneg) {
          if (wide) k_sauvola_mask_w<true, true, 4><<<grid, ST, 0, (cudaStream_t)stream>>>(p);
          else k_sauvola_mask_w<true, false, 4><<<grid, ST, 0, (cudaStream_t)stream>>>(p);
      } else if (wide) {
          k_sauvola_mask_w<false, true, 4><<<grid, ST, 0, (cudaStream_t)stream>>>(p);
      } else k_sauvola_mask_w<false, false, 4><<<grid, ST, 0, (cudaStream_t)stream>>>(p); }
    B200MRC_LAUNCH_CHECK();
    return B200MRC_OK;
}

extern "C" int b200mrc_sauvola_items(const b200mrc_sauvola_item *items_dev, int n_items, int max_width, int max_height,
                                     int window_width, int window_height, double k, double R, void *stream)
{
    if (n_items < 0 || (n_items && (!items_dev || max_width <= 0 || max_height <= 0)) || !(R > 0)) return B200MRC_ERR_INVALID;
    if (window_width < 1 || window_height < 1 || window_width > B200MRC_MAX_WINDOW || window_height > B200MRC_MAX_WINDOW)
        return B200MRC_ERR_UNSUPPORTED;
    if (n_items == 0) return B200MRC_OK;
    if (n_items > 65535) return B200MRC_ERR_UNSUPPORTED;
    SauvolaParams p;
    memset(&p, 0, sizeof(p));
    p.l = (window_width + 1) / 2;  p.r = window_width / 2;
    p.o = (window_height + 1) / 2; p.u = window_height / 2;
    sauvola_geometry(p, max_width, max_height);                    // the largest item sizes the grid
    p.km1 = k - 1.0;
    p.k2 = k * k / R / R;
    p.kneg = k < 0;
    dim3 grid((unsigned)(p.n_strips * p.n_bands), (unsigned)n_items);
    const bool wide = window_width > 128;
    { ProfScope _ps("k_sauvola_items", (cudaStream_t)stream);
      if (p.kneg) {
          if (wide) k_sauvola_items<true, true><<<grid, ST, 0, (cudaStream_t)stream>>>(items_dev, p);
          else k_sauvola_items<true, false><<<grid, ST, 0, (cudaStream_t)stream>>>(items_dev, p);
      } else if (wide) k_sauvola_items<false, true><<<grid, ST, 0, (cudaStream_t)stream>>>(items_dev, p);
      else k_sauvola_items<false, false><<<grid, ST, 0, (cudaStream_t)stream>>>(items_dev, p); }
    B200MRC_LAUNCH_CHECK();
    return B200MRC_OK;
}
