// optimise_split.cu -- the production optimise path for n_fg = 3 / n_bg = 10
// (internetarchive/archive-pdf-tools internetarchivepdf/mrc.py:412-415, 439-449;
// semantics cython/optimiser.pyx:153-429, restated in optimise.cu / oracle orc_optimise).
//
//   out[y,x] = (FIR(y,x) + IIR(y,x)) / den(y,x)        for every pixel not in the layer's mask
//   FIR = sum of mask*img over the 2n x 2n box (depends only on the inputs: fully parallel)
//   IIR = sum of `out` over the n x n box above-left (sequential over rows)
//   den = #mask in the FIR box + (y-ys)(x-xs)
// Exactly one layer is computed per pixel (fg where mask==0, bg where mask==1), so one 64-bit
// record per pixel carries everything the sequential part needs from the parallel part:
//       r[0,17) | g[17,34) | b[34,51) | den[51,60) | mask bit 63
//
//   k_opt_fir : parallel.  CTA = column strip x row band, one barrier per row; packed RGBM input
//               ring in smem fed by TMA bulk loads; fg sums in 16-bit lanes, bg sums in 3 words;
//               the 20-column bg window only in warps that hold a mask pixel; writes the record
//               plane by TMA bulk stores.
//   k_opt_iir : the row-sequential sweep, now short: per row a thread adds out[y-1] / drops
//               out[y-n-1] in its IIR column sums (16-bit lanes), publishes them, one barrier, a
//               3- (fg) or 10-column (bg) left window, one exact multiply-high division, done.
//               Strips of a page are pipelined through the global mailbox like optimise.cu (start
//               lag + sparse release/acquire keep the hand-off off the critical path); the small
//               smem footprint keeps a whole 64-page batch resident in one wave.
// Each thread owns K adjacent columns (K = 2 in production, K = 4 selectable: B200MRC_FIR_K /
// B200MRC_IIR_K): window sums slide across the K columns and smem accesses are 64/128-bit.  Measured on
// B200 both kernels sit at the same time for K = 2 and 4: the FIR kernel is bound by shared-memory
// bandwidth (window reads) and the IIR kernel by the per-row dependency latency.  The record plane costs 8 B/px written + 8 B/px read of HBM traffic, hidden behind the
// arithmetic.
#include "common.cuh"
#include "tma.cuh"
#include <cstring>
#include <cstdlib>

namespace b200mrc {
namespace {

constexpr int FH = 16;                  // halo columns (>= 10, multiple of 16 for TMA alignment)
constexpr int NFG = 3, NBG = 10;
constexpr int RIN = 2 * NBG + 2;        // input ring rows (k_opt_fir)
constexpr int RFG = NFG + 1, RBG = NBG + 1;
constexpr int STAGES = 3;               // k_opt_fir input stage depth
constexpr int ISTAGES = 4;              // k_opt_iir input stage depth (rows in flight)
constexpr int L2_AHEAD = 12;            // rows prefetched into L2 ahead of the bulk loads
constexpr int LAG = 8;                  // a strip starts once its left neighbour is this many rows ahead
constexpr int PUB = 4;                  // progress is published every PUB rows
constexpr int MAXDEN = 4 * NBG * NBG + NBG * NBG;

__device__ __forceinline__ uint32_t lane_rb(uint32_t px) { return __byte_perm(px, 0, 0x4240); }   // r | b<<16
__device__ __forceinline__ uint32_t lane_gm(uint32_t px) { return __byte_perm(px, 0, 0x4341); }   // g | m<<16
__device__ __forceinline__ uint32_t byte_g(uint32_t px) { return __byte_perm(px, 0, 0x4441); }    // g

template <int K> struct VecK;
template <> struct VecK<1> { using type = uint32_t; };
template <> struct VecK<2> { using type = uint2; };
template <> struct VecK<4> { using type = uint4; };

template <int K> __device__ __forceinline__ void ldk(const uint32_t *p, uint32_t *v)
{
    const typename VecK<K>::type t = *reinterpret_cast<const typename VecK<K>::type *>(p);
    const uint32_t *s = reinterpret_cast<const uint32_t *>(&t);
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = s[k];
}
template <int K> __device__ __forceinline__ void stk(uint32_t *p, const uint32_t *v)
{
    typename VecK<K>::type t;
    uint32_t *s = reinterpret_cast<uint32_t *>(&t);
#pragma unroll
    for (int k = 0; k < K; k++) s[k] = v[k];
    *reinterpret_cast<typename VecK<K>::type *>(p) = t;
}
template <int K> __device__ __forceinline__ void ldk_cg(const uint32_t *p, uint32_t *v)
{
    const typename VecK<K>::type t = __ldcg(reinterpret_cast<const typename VecK<K>::type *>(p));
    const uint32_t *s = reinterpret_cast<const uint32_t *>(&t);
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = s[k];
}
template <int K> __device__ __forceinline__ void stk_cg(uint32_t *p, const uint32_t *v)
{
    typename VecK<K>::type t;
    uint32_t *s = reinterpret_cast<uint32_t *>(&t);
#pragma unroll
    for (int k = 0; k < K; k++) s[k] = v[k];
    __stcg(reinterpret_cast<typename VecK<K>::type *>(p), t);
}

// K interleaved pixels (C bytes each) -> K packed r | g<<8 | b<<16 words; byte k of `mbytes`
// becomes the top byte of pixel k
template <int K, int C>
__device__ __forceinline__ void unpack_px(const uint8_t *rgb, uint32_t mbytes, uint32_t *px)
{
    if (K == 1) {
        const uint32_t v = C == 1 ? (uint32_t)rgb[0] * 0x010101u
                                  : ((uint32_t)rgb[0] | ((uint32_t)rgb[1] << 8) | ((uint32_t)rgb[2] << 16));
        px[0] = v | (mbytes << 24);
    } else if (C == 3 && K == 4) {
        const uint32_t w0 = *reinterpret_cast<const uint32_t *>(rgb), w1 = *reinterpret_cast<const uint32_t *>(rgb + 4),
                       w2 = *reinterpret_cast<const uint32_t *>(rgb + 8);
        px[0] = __byte_perm(w0, mbytes, 0x4210);
        px[1] = __byte_perm(__byte_perm(w0, w1, 0x0543), mbytes, 0x5210);
        px[2] = __byte_perm(__byte_perm(w1, w2, 0x0432), mbytes, 0x6210);
        px[3] = __byte_perm(w2, mbytes, 0x7321);
    } else if (C == 3 && K == 2) {
        const uint16_t *h = reinterpret_cast<const uint16_t *>(rgb);
        const uint32_t w = (uint32_t)h[0] | ((uint32_t)h[1] << 16), x = h[2];      // r0 g0 b0 r1 | g1 b1
        px[0] = __byte_perm(w, mbytes, 0x4210);
        px[1] = __byte_perm(__byte_perm(w, x, 0x0543), mbytes, 0x5210);
    } else if (C == 1 && K == 4) {
        const uint32_t w = *reinterpret_cast<const uint32_t *>(rgb);
        px[0] = __byte_perm(w, mbytes, 0x4000); px[1] = __byte_perm(w, mbytes, 0x5111);
        px[2] = __byte_perm(w, mbytes, 0x6222); px[3] = __byte_perm(w, mbytes, 0x7333);
    } else {
        const uint32_t w = *reinterpret_cast<const uint16_t *>(rgb);
        px[0] = __byte_perm(w, mbytes, 0x4000); px[1] = __byte_perm(w, mbytes, 0x5111);
    }
}

// K packed RGB words -> K*C interleaved bytes at `dst` (4-byte aligned for K=4, 2-byte for K=2)
template <int K, int C>
__device__ __forceinline__ void pack_px(uint8_t *dst, const uint32_t *px)
{
    if (K == 1) {
        dst[0] = (uint8_t)px[0];
        if (C == 3) { dst[1] = (uint8_t)(px[0] >> 8); dst[2] = (uint8_t)(px[0] >> 16); }
    } else if (C == 3 && K == 4) {
        uint32_t *d = reinterpret_cast<uint32_t *>(dst);
        d[0] = __byte_perm(px[0], px[1], 0x4210); d[1] = __byte_perm(px[1], px[2], 0x5421); d[2] = __byte_perm(px[2], px[3], 0x6542);
    } else if (C == 3 && K == 2) {
        uint16_t *d = reinterpret_cast<uint16_t *>(dst);
        d[0] = (uint16_t)px[0]; d[1] = (uint16_t)(((px[0] >> 16) & 0xffu) | ((px[1] & 0xffu) << 8)); d[2] = (uint16_t)(px[1] >> 8);
    } else if (C == 1 && K == 4) {
        *reinterpret_cast<uint32_t *>(dst) = (px[0] & 0xffu) | ((px[1] & 0xffu) << 8) | ((px[2] & 0xffu) << 16) | (px[3] << 24);
    } else {
        *reinterpret_cast<uint16_t *>(dst) = (uint16_t)((px[0] & 0xffu) | ((px[1] & 0xffu) << 8));
    }
}

// mask bytes (any non-zero = set) of K pixels -> K bytes holding 0/1
template <int K> __device__ __forceinline__ uint32_t load_mask01(const uint8_t *m)
{
    if (K == 1) return m[0] != 0;
    uint32_t w = K == 4 ? *reinterpret_cast<const uint32_t *>(m) : (uint32_t)*reinterpret_cast<const uint16_t *>(m);
    w = (((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) >> 7;
    return w & 0x01010101u;
}

// =================================================================================== k_opt_fir
struct FirParams {
    const uint8_t *mask; int64_t mpitch, mstride;
    const uint8_t *img;  int64_t ipitch, istride;
    uint8_t *rec; int64_t rpitch, rstride;         // 8 B / pixel
    int W, H, S, SW, n_bands, band_h;
    int fmt;                                       // 0: legacy records; 1: 16-bit-lane fg records (optimise_warp.cu)
};

template <int C, int K, int T>
__global__ void __launch_bounds__(T) k_opt_fir(const FirParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    constexpr int E = T * K, SW = E - 2 * FH;               // compile-time geometry: address arithmetic folds away
    constexpr int rowRGB = (E * C + 15) & ~15, rowM = (E + 15) & ~15, rowOut = SW * 8;

    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem);
    uint8_t *rawRGB = smem + 64;
    uint8_t *rawM = rawRGB + STAGES * rowRGB;
    uint8_t *outSt = rawM + STAGES * rowM;                                   // [3][rowOut]
    uint32_t *ringIn = reinterpret_cast<uint32_t *>(outSt + 3 * rowOut);     // [RIN][E]
    uint32_t *ex = ringIn + RIN * E;                                         // [2][5][E]

    const int strip = blockIdx.x % p.S, band = blockIdx.x / p.S, page = blockIdx.y;
    const int W = p.W, H = p.H;
    const int x0 = strip * SW, i0 = tid * K, xg = x0 - FH + i0;
    const int by0 = band * p.band_h, by1 = min(H, by0 + p.band_h);
    const int rmin = max(0, by0 - NBG);                   // first row that can matter to this band
    const bool in_core = i0 >= FH && i0 < FH + SW;
    uint32_t cvm[K];                                      // all-ones for columns inside the page
#pragma unroll
    for (int k = 0; k < K; k++) cvm[k] = ((xg + k) >= 0 && (xg + k) < W) ? 0xffffffffu : 0u;

    const uint8_t *img = p.img + (int64_t)page * p.istride;
    const uint8_t *mask = p.mask + (int64_t)page * p.mstride;
    uint8_t *rec = p.rec + (int64_t)page * p.rstride;

    const int cs = max(0, x0 - FH), ce = min(W, x0 + SW + FH), dcol = cs - (x0 - FH);
    const uint32_t bytesRGB = (uint32_t)(((ce - cs) * C + 15) & ~15), bytesM = (uint32_t)(((ce - cs) + 15) & ~15);
    const int ocols = min(W, x0 + SW) - x0;
    const uint32_t bytesOut = (uint32_t)(((ocols * 8) + 15) & ~15);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) mbar_init(&mbar[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto issue_row = [&](int row) {                          // thread 0; slot = (row - rmin) % STAGES
        const int s = (row - rmin) % STAGES;
        mbar_expect_tx(&mbar[s], bytesRGB + bytesM);
        tma_load(rawRGB + s * rowRGB + dcol * C, img + (int64_t)row * p.ipitch + (int64_t)cs * C, bytesRGB, &mbar[s]);
        tma_load(rawM + s * rowM + dcol, mask + (int64_t)row * p.mpitch + cs, bytesM, &mbar[s]);
    };
    const int rlast = min(H, by1 + NBG - 1);                 // rows [rmin, rlast) are streamed
    if (tid == 0)
        for (int r = rmin; r < rmin + STAGES && r < rlast; r++) issue_row(r);

    uint32_t Ffg_rb[K], Ffg_gm[K], Fbg0[K], Fbg1[K], Fbg2[K];
#pragma unroll
    for (int k = 0; k < K; k++) { Ffg_rb[k] = Ffg_gm[k] = Fbg0[k] = Fbg1[k] = Fbg2[k] = 0; }

    // ring words: valid pixel = r|g<<8|b<<16 with top byte 0/1 (mask); column outside the page = 0x80000000
    // (top byte 0x80: belongs to neither layer)
    auto fir_fg = [&](uint32_t px, int k, bool add) {
        if ((px >> 24) == 1u) {
            if (add) { Ffg_rb[k] += lane_rb(px); Ffg_gm[k] += lane_gm(px); }
            else     { Ffg_rb[k] -= lane_rb(px); Ffg_gm[k] -= lane_gm(px); }
        }
    };
    auto fir_bg = [&](uint32_t px, int k, bool add) {
        if ((px >> 24) == 0u) {
            const uint32_t w0 = (px & 0xffu) | (1u << 20), g = byte_g(px), b = (px >> 16) & 0xffu;
            if (add) { Fbg0[k] += w0; Fbg1[k] += g; Fbg2[k] += b; }
            else     { Fbg0[k] -= w0; Fbg1[k] -= g; Fbg2[k] -= b; }
        }
    };

    // virtual row loop: starts early enough that the running sums are complete at y = by0
    const int ys0 = rmin - (NBG - 1);
    auto slot_of = [&](int r) { int s = (r - rmin) % RIN; return s < 0 ? s + RIN : s; };
    // ring cursors as word offsets (slot * E), advanced by E and wrapped by compare
    int o_e9 = slot_of(ys0 + NBG - 1) * E, o_e2 = slot_of(ys0 + NFG - 1) * E, o_cur = slot_of(ys0) * E,
        o_l4 = slot_of(ys0 - NFG - 1) * E, o_l11 = slot_of(ys0 - NBG - 1) * E;
    int st_slot = 0, st_par = 0, ob = 0;                     // ob: staging buffer of row y (cycles 0,1,2 from y = by0)

    for (int y = ys0; y < by1; y++) {
        const int re9 = y + NBG - 1, re2 = y + NFG - 1, rl4 = y - NFG - 1, rl11 = y - NBG - 1;
        const bool emit = y >= by0;
        uint32_t t[K];

        const bool have9 = re9 >= rmin && re9 < rlast;       // == the rows that were issued
        if (have9) {
            mbar_wait(&mbar[st_slot], (uint32_t)st_par);
            const uint32_t mb = load_mask01<K>(rawM + st_slot * rowM + i0);
            unpack_px<K, C>(rawRGB + st_slot * rowRGB + i0 * C, mb, t);
#pragma unroll
            for (int k = 0; k < K; k++) t[k] = (t[k] & cvm[k]) | (~cvm[k] & 0x80000000u);
            stk<K>(ringIn + o_e9 + i0, t);
#pragma unroll
            for (int k = 0; k < K; k++) fir_bg(t[k], k, true);
        }
        if (re2 >= rmin && re2 < H) {
            ldk<K>(ringIn + o_e2 + i0, t);
#pragma unroll
            for (int k = 0; k < K; k++) fir_fg(t[k], k, true);
        }
        if (rl4 >= rmin) {
            ldk<K>(ringIn + o_l4 + i0, t);
#pragma unroll
            for (int k = 0; k < K; k++) fir_fg(t[k], k, false);
        }
        if (rl11 >= rmin) {
            ldk<K>(ringIn + o_l11 + i0, t);
#pragma unroll
            for (int k = 0; k < K; k++) fir_bg(t[k], k, false);
        }
        uint32_t *exb = ex + (y & 1) * 5 * E;
        if (emit) {
            stk<K>(exb + 0 * E + i0, Ffg_rb);
            stk<K>(exb + 1 * E + i0, Ffg_gm);
            stk<K>(exb + 2 * E + i0, Fbg0);
            stk<K>(exb + 3 * E + i0, Fbg1);
            stk<K>(exb + 4 * E + i0, Fbg2);
        }
        if (tid == 0) tma_wait_read<1>();                    // staging buffer `ob` is free again
        __syncthreads();
        if (tid == 0) {
            if (y - 1 >= by0) {                              // row y-1 is completely staged now
                const int pb = ob == 0 ? 2 : ob - 1;
                tma_store(rec + (int64_t)(y - 1) * p.rpitch + (int64_t)x0 * 8, outSt + pb * rowOut, bytesOut);
                tma_commit();
            }
            const int nr = re9 + STAGES;
            if (have9 && nr < rlast) issue_row(nr);
        }

        if (emit && in_core) {
            uint32_t cur[K];
            ldk<K>(ringIn + o_cur + i0, cur);
            uint32_t any_m = 0;
#pragma unroll
            for (int k = 0; k < K; k++) any_m |= cur[k] & 0x01000000u;
            const bool need_bg = __any_sync(__activemask(), any_m != 0);
            uint32_t num_r[K], num_g[K], num_b[K], den[K];
            uint32_t s[2][K];
            {   // fg: sum of F over [c-3, c+3), slid across the K columns (16-bit lanes)
                constexpr int NL = (NFG + K - 1) / K * K;
#pragma unroll
                for (int w = 0; w < 2; w++) {
                    uint32_t a[NL + 2 * K];
                    const uint32_t *src = exb + w * E + i0;
#pragma unroll
                    for (int g = 0; g < NL / K; g++) ldk<K>(src - NL + g * K, &a[g * K]);
#pragma unroll
                    for (int k = 0; k < K; k++) a[NL + k] = w == 0 ? Ffg_rb[k] : Ffg_gm[k];
                    ldk<K>(src + K, &a[NL + K]);
                    uint32_t acc = 0;
#pragma unroll
                    for (int d = -NFG; d < NFG; d++) acc += a[NL + d];
                    s[w][0] = acc;
#pragma unroll
                    for (int k = 1; k < K; k++) { acc += a[NL + k - 1 + NFG] - a[NL + k - 1 - NFG]; s[w][k] = acc; }
                }
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const int x = xg + k;
                    num_r[k] = s[0][k] & 0xffffu; num_b[k] = s[0][k] >> 16; num_g[k] = s[1][k] & 0xffffu;
                    den[k] = (s[1][k] >> 16) + (uint32_t)((y - max(0, y - NFG)) * (x - max(0, x - NFG)));
                }
            }
            if (need_bg) {   // bg: sum of F over [c-10, c+10), word by word
                constexpr int NL = (NBG + K - 1) / K * K, NR = (NBG - 1 + K - 1) / K * K;
                uint32_t res[3][K];
#pragma unroll
                for (int w = 0; w < 3; w++) {
                    uint32_t a[NL + K + NR];
                    const uint32_t *src = exb + (2 + w) * E + i0;
#pragma unroll
                    for (int g = 0; g < NL / K; g++) ldk<K>(src - NL + g * K, &a[g * K]);
#pragma unroll
                    for (int k = 0; k < K; k++) a[NL + k] = w == 0 ? Fbg0[k] : (w == 1 ? Fbg1[k] : Fbg2[k]);
#pragma unroll
                    for (int g = 0; g < NR / K; g++) ldk<K>(src + K + g * K, &a[NL + K + g * K]);
                    uint32_t acc = 0;
#pragma unroll
                    for (int d = -NBG; d < NBG; d++) acc += a[NL + d];
                    res[w][0] = acc;
#pragma unroll
                    for (int k = 1; k < K; k++) { acc += a[NL + k - 1 + NBG] - a[NL + k - 1 - NBG]; res[w][k] = acc; }
                }
#pragma unroll
                for (int k = 0; k < K; k++) {
                    if (cur[k] & 0x01000000u) {
                        const int x = xg + k;
                        num_r[k] = res[0][k] & 0xfffffu; num_g[k] = res[1][k]; num_b[k] = res[2][k];
                        den[k] = (res[0][k] >> 20) + (uint32_t)((y - max(0, y - NBG)) * (x - max(0, x - NBG)));
                    }
                }
            }
            uint32_t o[2 * K];
#pragma unroll
            for (int k = 0; k < K; k++) {
                o[2 * k] = num_r[k] | (num_g[k] << 17);
                o[2 * k + 1] = (num_g[k] >> 15) | (num_b[k] << 2) | (den[k] << 19) | ((cur[k] & 0x01000000u) << 7);
                if (p.fmt == 1 && !(cur[k] & 0x01000000u)) {
                    // fg-type pixel, lanes form: 2*Fr | 2*Fb << 16 ; 2*Fg | 4*den << 16  (bit 31 stays clear: den <= 45)
                    o[2 * k] = s[0][k] << 1;
                    o[2 * k + 1] = ((s[1][k] & 0xffffu) << 1) | (den[k] << 18);
                }
            }
            uint32_t *dst = reinterpret_cast<uint32_t *>(outSt + ob * rowOut + (i0 - FH) * 8);
            if (K == 1) *reinterpret_cast<uint2 *>(dst) = make_uint2(o[0], o[1]);
            else {
#pragma unroll
                for (int g = 0; g + 3 < 2 * K; g += 4) *reinterpret_cast<uint4 *>(dst + g) = make_uint4(o[g], o[g + 1], o[g + 2], o[g + 3]);
            }
            fence_proxy_async();
        }
        o_e9 += E; if (o_e9 == RIN * E) o_e9 = 0;
        o_e2 += E; if (o_e2 == RIN * E) o_e2 = 0;
        o_cur += E; if (o_cur == RIN * E) o_cur = 0;
        o_l4 += E; if (o_l4 == RIN * E) o_l4 = 0;
        o_l11 += E; if (o_l11 == RIN * E) o_l11 = 0;
        if (emit && ++ob == 3) ob = 0;
        if (have9 && ++st_slot == STAGES) { st_slot = 0; st_par ^= 1; }
    }
    __syncthreads();
    if (tid == 0) {
        const int pb = ob == 0 ? 2 : ob - 1;
        tma_store(rec + (int64_t)(by1 - 1) * p.rpitch + (int64_t)x0 * 8, outSt + pb * rowOut, bytesOut);
        tma_commit();
        tma_wait_all<0>();
    }
}

size_t fir_smem_bytes(int T, int K, int C)
{
    const size_t E = (size_t)T * K, SW = E - 2 * FH;
    const size_t rowRGB = (E * C + 15) & ~(size_t)15, rowM = (E + 15) & ~(size_t)15;
    return 64 + STAGES * (rowRGB + rowM) + 3 * SW * 8 + (size_t)(RIN + 10) * E * 4 + 64;
}

// =================================================================================== k_opt_iir
struct IirParams {
    const uint8_t *img; int64_t ipitch, istride;
    const uint8_t *rec; int64_t rpitch, rstride;
    uint8_t *ofg; int64_t fpitch, fstride;
    uint8_t *obg; int64_t bpitch, bstride;
    int W, H, N, S, SW;
    uint32_t *mailbox;                  // [N][S][H][2][FH]
    int *prog;                          // [N][S]
    unsigned *ticket;
};

__device__ __forceinline__ uint32_t div31(uint32_t num, uint32_t m31)
{
    return (uint32_t)(((unsigned long long)num * m31) >> 31);
}

template <int C, int K, int T>
__global__ void __launch_bounds__(T) k_opt_iir(const IirParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    constexpr int E = T * K, SW = E - FH;                   // E = FH + SW : left halo + interior (compile-time)
    constexpr int rowRGB = (SW * C + 15) & ~15, rowRec = SW * 8, rowOut = rowRGB;

    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem);
    int *s_misc = reinterpret_cast<int *>(smem + 32);
    uint8_t *rawRGB = smem + 64;                                             // [ISTAGES][rowRGB]
    uint8_t *rawRec = rawRGB + ISTAGES * rowRGB;                             // [ISTAGES][rowRec]
    uint8_t *outSt = rawRec + ISTAGES * rowRec;                              // [3][2][rowOut]
    uint32_t *ringFg = reinterpret_cast<uint32_t *>(outSt + 6 * rowOut);     // [RFG][E]
    uint32_t *ringBg = ringFg + RFG * E;                                     // [RBG][E]
    uint32_t *ex = ringBg + RBG * E;                                         // [2][4][E]
    uint32_t *Mtab = ex + 8 * E;                                             // [MAXDEN + 1]

    if (tid == 0) {
        s_misc[0] = (int)atomicAdd(p.ticket, 1u);
        s_misc[1] = 0;
#pragma unroll
        for (int s = 0; s < ISTAGES; s++) mbar_init(&mbar[s], 1);
        fence_mbar_init();
    }
    // den -> ceil(2^31 / den): floor(num/den) == (num * M) >> 31 for num <= 255*den, den <= 500
    for (int d = 1 + tid; d <= MAXDEN; d += T) Mtab[d] = (uint32_t)((0x80000000ull + d - 1) / (unsigned long long)d);
    if (tid == 0) Mtab[0] = 0;
    __syncthreads();
    const int job = s_misc[0];
    const int page = job / p.S, strip = job - page * p.S;
    const int W = p.W, H = p.H;
    const int x0 = strip * SW, i0 = tid * K, xg = x0 - FH + i0;
    const bool haloL = i0 < FH && strip > 0;
    const bool in_core = i0 >= FH;
    const bool has_right = strip + 1 < p.S;
    uint32_t cvm[K];
#pragma unroll
    for (int k = 0; k < K; k++) cvm[k] = ((xg + k) >= 0 && (xg + k) < W) ? 0xffffffffu : 0u;

    const uint8_t *img = p.img + (int64_t)page * p.istride;
    const uint8_t *rec = p.rec + (int64_t)page * p.rstride;
    uint8_t *ofg = p.ofg + (int64_t)page * p.fstride;
    uint8_t *obg = p.obg + (int64_t)page * p.bstride;
    uint32_t *mb_out = p.mailbox + ((int64_t)page * p.S + strip) * (int64_t)H * 2 * FH;
    const uint32_t *mb_in = strip > 0 ? p.mailbox + ((int64_t)page * p.S + strip - 1) * (int64_t)H * 2 * FH : nullptr;
    int *prog_out = p.prog + (int64_t)page * p.S + strip;
    const int *prog_in = strip > 0 ? p.prog + (int64_t)page * p.S + strip - 1 : nullptr;
    int known = 0;

    const int ocols = min(W, x0 + SW) - x0;
    const uint32_t bytesRGB = (uint32_t)((ocols * C + 15) & ~15), bytesRec = (uint32_t)((ocols * 8 + 15) & ~15);

    auto issue_row = [&](int row) {
        const int s = row % ISTAGES;
        const int pr = row + L2_AHEAD;
        if (pr < H) {
            tma_prefetch_l2(img + (int64_t)pr * p.ipitch + (int64_t)x0 * C, bytesRGB);
            tma_prefetch_l2(rec + (int64_t)pr * p.rpitch + (int64_t)x0 * 8, bytesRec);
        }
        mbar_expect_tx(&mbar[s], bytesRGB + bytesRec);
        tma_load(rawRGB + s * rowRGB, img + (int64_t)row * p.ipitch + (int64_t)x0 * C, bytesRGB, &mbar[s]);
        tma_load(rawRec + s * rowRec, rec + (int64_t)row * p.rpitch + (int64_t)x0 * 8, bytesRec, &mbar[s]);
    };
    if (tid == 0) {
        for (int r = 0; r < L2_AHEAD && r < H; r++) {
            tma_prefetch_l2(img + (int64_t)r * p.ipitch + (int64_t)x0 * C, bytesRGB);
            tma_prefetch_l2(rec + (int64_t)r * p.rpitch + (int64_t)x0 * 8, bytesRec);
        }
        for (int r = 0; r < ISTAGES && r < H; r++) issue_row(r);
        // start lag: run LAG rows behind the left neighbour so that the per-row hand-off (release -> acquire
        // through L2) is never on the critical path; the cached progress then covers several rows per poll
        if (strip > 0) {
            const int want = min(H, LAG);
            int v = ld_acquire(prog_in);
            while (v < want) { __nanosleep(100); v = ld_acquire(prog_in); }
            known = v;
            s_misc[1] = v;
        }
    }
    __syncthreads();

    // IIR column sums, both layers in 16-bit lanes (r | b<<16, g): a column sum is <= 10*255 and a
    // 10-column window sum <= 25500, so nothing can carry across a lane
    uint32_t Cfg_rb[K], Cfg_g[K], Cbg_rb[K], Cbg_g[K], prev_fg[K], prev_bg[K], pf_fg[K], pf_bg[K];
#pragma unroll
    for (int k = 0; k < K; k++) { Cfg_rb[k] = Cfg_g[k] = Cbg_rb[k] = Cbg_g[k] = prev_fg[k] = prev_bg[k] = pf_fg[k] = pf_bg[k] = 0; }
    int pf_row = -1;
    int of_new = (RFG - 1) * E, of_old = ((RFG - 1 - NFG + RFG) % RFG) * E;   // word offsets of rows y-1 (write) / y-4 (read)
    int ob_new = (RBG - 1) * E, ob_old = ((RBG - 1 - NBG + RBG) % RBG) * E;
    int st_slot = 0, st_par = 0, ob = 0;                     // ob = y % 3: output staging buffer of row y

    for (int y = 0; y < H; y++) {
        // ---- left halo: out[y-1] of the neighbour strip's last columns
        if (haloL) {
            if (y >= 1) {
                if (pf_row == y - 1) {
#pragma unroll
                    for (int k = 0; k < K; k++) { prev_fg[k] = pf_fg[k]; prev_bg[k] = pf_bg[k]; }
                } else {                                     // published: waited for before the previous barrier
                    ldk_cg<K>(mb_in + ((int64_t)(y - 1) * 2 + 0) * FH + i0, prev_fg);
                    ldk_cg<K>(mb_in + ((int64_t)(y - 1) * 2 + 1) * FH + i0, prev_bg);
                }
            }
            if (s_misc[1] >= y + 1) {
                ldk_cg<K>(mb_in + ((int64_t)y * 2 + 0) * FH + i0, pf_fg);
                ldk_cg<K>(mb_in + ((int64_t)y * 2 + 1) * FH + i0, pf_bg);
                pf_row = y;
            }
        }
        // ---- IIR column sums: + out[y-1], - out[y-n-1]
        if (y >= 1) {
            uint32_t t[K];
#pragma unroll
            for (int k = 0; k < K; k++) {
                const uint32_t pf = prev_fg[k] & cvm[k], pb = prev_bg[k] & cvm[k];      // columns outside the page hold zeros
                prev_fg[k] = pf; prev_bg[k] = pb;
                Cfg_rb[k] += lane_rb(pf); Cfg_g[k] += byte_g(pf);
                Cbg_rb[k] += lane_rb(pb); Cbg_g[k] += byte_g(pb);
            }
            if (y - NFG - 1 >= 0) {
                ldk<K>(ringFg + of_old + i0, t);
#pragma unroll
                for (int k = 0; k < K; k++) { Cfg_rb[k] -= lane_rb(t[k]); Cfg_g[k] -= byte_g(t[k]); }
            }
            if (y - NBG - 1 >= 0) {
                ldk<K>(ringBg + ob_old + i0, t);
#pragma unroll
                for (int k = 0; k < K; k++) { Cbg_rb[k] -= lane_rb(t[k]); Cbg_g[k] -= byte_g(t[k]); }
            }
            stk<K>(ringFg + of_new + i0, prev_fg);
            stk<K>(ringBg + ob_new + i0, prev_bg);
        }
        uint32_t *exb = ex + (y & 1) * 4 * E;
        stk<K>(exb + 0 * E + i0, Cfg_rb);
        stk<K>(exb + 1 * E + i0, Cfg_g);
        stk<K>(exb + 2 * E + i0, Cbg_rb);
        stk<K>(exb + 3 * E + i0, Cbg_g);

        if (tid == 0) {
            tma_wait_read<1>();                              // output staging buffer `ob` is free again
            // the left neighbour must have published row y before row y+1 starts
            if (strip > 0 && y + 1 < H && known < y + 1) {
                int v = ld_acquire(prog_in);
                while (v < y + 1) { __nanosleep(20); v = ld_acquire(prog_in); }
                known = v;
                s_misc[1] = v;
            }
        }
        mbar_wait(&mbar[st_slot], (uint32_t)st_par);         // this row's RGB + records have landed
        __syncthreads();
        if (tid == 0 && y >= 1) {
            const int pb = ob == 0 ? 2 : ob - 1;             // row y-1 is completely staged now
            tma_store(ofg + (int64_t)(y - 1) * p.fpitch + (int64_t)x0 * C, outSt + (pb * 2 + 0) * rowOut, bytesRGB);
            tma_store(obg + (int64_t)(y - 1) * p.bpitch + (int64_t)x0 * C, outSt + (pb * 2 + 1) * rowOut, bytesRGB);
            tma_commit();
            const int nr = (y - 1) + ISTAGES;                // stage slot of row y-1: every reader passed this barrier
            if (nr < H) issue_row(nr);
        }

        if (in_core) {
            const int li = i0 - FH;                          // interior column index of this thread's first pixel
            uint32_t rc[2 * K], rgb[K];
            if (K == 1) {
                const uint2 v = *reinterpret_cast<const uint2 *>(rawRec + st_slot * rowRec + li * 8);
                rc[0] = v.x; rc[1] = v.y;
            } else {
#pragma unroll
                for (int g = 0; g + 3 < 2 * K; g += 4) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(rawRec + st_slot * rowRec + li * 8 + g * 4);
                    rc[g] = v.x; rc[g + 1] = v.y; rc[g + 2] = v.z; rc[g + 3] = v.w;
                }
            }
            unpack_px<K, C>(rawRGB + st_slot * rowRGB + li * C, 0u, rgb);
            uint32_t any_m = 0;
#pragma unroll
            for (int k = 0; k < K; k++) { rgb[k] &= 0x00ffffffu; any_m |= rc[2 * k + 1]; }
            const bool need_bg = __any_sync(__activemask(), (any_m >> 31) != 0);
            // IIR windows [c-n, c), slid across the K columns
            uint32_t ifg[2][K], ibg[2][K];
            {
                constexpr int NL = (NFG + K - 1) / K * K;
#pragma unroll
                for (int w = 0; w < 2; w++) {
                    uint32_t a[NL + K];
                    const uint32_t *src = exb + w * E + i0;
#pragma unroll
                    for (int g = 0; g < NL / K; g++) ldk<K>(src - NL + g * K, &a[g * K]);
#pragma unroll
                    for (int k = 0; k < K; k++) a[NL + k] = w == 0 ? Cfg_rb[k] : Cfg_g[k];
                    uint32_t acc = 0;
#pragma unroll
                    for (int d = 1; d <= NFG; d++) acc += a[NL - d];
                    ifg[w][0] = acc;
#pragma unroll
                    for (int k = 1; k < K; k++) { acc += a[NL + k - 1] - a[NL + k - 1 - NFG]; ifg[w][k] = acc; }
                }
            }
            if (need_bg) {
                constexpr int NL = (NBG + K - 1) / K * K;
#pragma unroll
                for (int w = 0; w < 2; w++) {
                    uint32_t a[NL + K];
                    const uint32_t *src = exb + (2 + w) * E + i0;
#pragma unroll
                    for (int g = 0; g < NL / K; g++) ldk<K>(src - NL + g * K, &a[g * K]);
#pragma unroll
                    for (int k = 0; k < K; k++) a[NL + k] = w == 0 ? Cbg_rb[k] : Cbg_g[k];
                    uint32_t acc = 0;
#pragma unroll
                    for (int d = 1; d <= NBG; d++) acc += a[NL - d];
                    ibg[w][0] = acc;
#pragma unroll
                    for (int k = 1; k < K; k++) { acc += a[NL + k - 1] - a[NL + k - 1 - NBG]; ibg[w][k] = acc; }
                }
            } else {
#pragma unroll
                for (int k = 0; k < K; k++) { ibg[0][k] = 0; ibg[1][k] = 0; }
            }
            uint32_t ofg_px[K], obg_px[K];
#pragma unroll
            for (int k = 0; k < K; k++) {
                const uint32_t lo = rc[2 * k], hi = rc[2 * k + 1];
                const bool m = (hi >> 31) != 0;
                uint32_t nr = lo & 0x1ffffu, ng = (lo >> 17) | ((hi & 3u) << 15), nb = (hi >> 2) & 0x1ffffu;
                const uint32_t den = min((hi >> 19) & 0xfffu, (uint32_t)MAXDEN);   // clamp: padding columns hold garbage
                const uint32_t srb = m ? ibg[0][k] : ifg[0][k], sg = m ? ibg[1][k] : ifg[1][k];
                nr += srb & 0xffffu; nb += srb >> 16; ng += sg;
                const uint32_t m31 = Mtab[den];
                const uint32_t q = div31(nr, m31) | (div31(ng, m31) << 8) | (div31(nb, m31) << 16);
                ofg_px[k] = m ? rgb[k] : q;
                obg_px[k] = m ? q : rgb[k];
                prev_fg[k] = ofg_px[k]; prev_bg[k] = obg_px[k];
            }
            pack_px<K, C>(outSt + (ob * 2 + 0) * rowOut + li * C, ofg_px);
            pack_px<K, C>(outSt + (ob * 2 + 1) * rowOut + li * C, obg_px);
            if (has_right && i0 >= SW) {                     // last FH interior columns feed the right neighbour
                // rows <= y-1 were stored a full row-step (one CTA barrier) ago: the release has nothing to wait for
                if (i0 == SW && (y % PUB) == 0) st_release(prog_out, y);
                uint32_t a[K], b[K];
#pragma unroll
                for (int k = 0; k < K; k++) { a[k] = ofg_px[k] & cvm[k]; b[k] = obg_px[k] & cvm[k]; }
                stk_cg<K>(mb_out + ((int64_t)y * 2 + 0) * FH + (i0 - SW), a);
                stk_cg<K>(mb_out + ((int64_t)y * 2 + 1) * FH + (i0 - SW), b);
            }
            fence_proxy_async();
        }
        of_new += E; if (of_new == RFG * E) of_new = 0;
        of_old += E; if (of_old == RFG * E) of_old = 0;
        ob_new += E; if (ob_new == RBG * E) ob_new = 0;
        ob_old += E; if (ob_old == RBG * E) ob_old = 0;
        if (++ob == 3) ob = 0;
        if (++st_slot == ISTAGES) { st_slot = 0; st_par ^= 1; }
    }
    __syncthreads();
    if (has_right && i0 == SW) { __threadfence(); st_release(prog_out, H); }
    if (tid == 0) {
        const int pb = ob == 0 ? 2 : ob - 1;
        tma_store(ofg + (int64_t)(H - 1) * p.fpitch + (int64_t)x0 * C, outSt + (pb * 2 + 0) * rowOut, bytesRGB);
        tma_store(obg + (int64_t)(H - 1) * p.bpitch + (int64_t)x0 * C, outSt + (pb * 2 + 1) * rowOut, bytesRGB);
        tma_commit();
        tma_wait_all<0>();
    }
}

size_t iir_smem_bytes(int T, int K, int C)
{
    const size_t E = (size_t)T * K, SW = E - FH;
    const size_t rowRGB = (SW * C + 15) & ~(size_t)15;
    return 64 + ISTAGES * (rowRGB + SW * 8) + 6 * rowRGB + (size_t)(RFG + RBG + 8) * E * 4 + (MAXDEN + 1) * 4 + 64;
}

// ---- kernel tables: (K, T) instantiations.  FIR: E = K*T in {64, 128, 256}; IIR: E in 64..384 step 64
template <int C, int K> const void *fir_kernel(int T)
{
    switch (T * K) {
    case 64:  return (const void *)k_opt_fir<C, K, 64 / K>;
    case 128: return (const void *)k_opt_fir<C, K, 128 / K>;
    case 256: return (const void *)k_opt_fir<C, K, 256 / K>;
    default: return nullptr;
    }
}
template <int C, int K> const void *iir_kernel(int T)
{
    switch (T * K) {
    case 64:  return (const void *)k_opt_iir<C, K, 64 / K>;
    case 128: return (const void *)k_opt_iir<C, K, 128 / K>;
    case 192: return (const void *)k_opt_iir<C, K, 192 / K>;
    case 256: return (const void *)k_opt_iir<C, K, 256 / K>;
    case 320: return (const void *)k_opt_iir<C, K, 320 / K>;
    case 384: return (const void *)k_opt_iir<C, K, 384 / K>;
    default: return nullptr;
    }
}
const void *get_fir(int C, int K, int T)
{
    if (K == 2) return C == 1 ? fir_kernel<1, 2>(T) : fir_kernel<3, 2>(T);
    return C == 1 ? fir_kernel<1, 4>(T) : fir_kernel<3, 4>(T);
}
const void *get_iir(int C, int K, int T)
{
    if (K == 2) return C == 1 ? iir_kernel<1, 2>(T) : iir_kernel<3, 2>(T);
    return C == 1 ? iir_kernel<1, 4>(T) : iir_kernel<3, 4>(T);
}

struct SplitPlan { int fK, fS, fSW, fT, bands, band_h; size_t fsmem; int iK, iS, iSW, iT; size_t ismem; };

int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

int plan_split(int W, int H, int N, int C, SplitPlan &out)
{
    static int cW = -1, cH = -1, cN = -1, cC = -1;
    static SplitPlan cp;
    if (cW == W && cH == H && cN == N && cC == C) { out = cp; return B200MRC_OK; }
    const DevInfo &di = dev_info();
    SplitPlan pl;
    // ---- FIR kernel: throughput kernel; E = 256 columns per CTA (224 useful), 256-row bands
    {
        pl.fK = env_int("B200MRC_FIR_K", 2);
        const int E = env_int("B200MRC_FIR_E", W > 96 ? 256 : (W > 32 ? 128 : 64));
        if ((pl.fK != 2 && pl.fK != 4) || (E != 64 && E != 128 && E != 256)) return B200MRC_ERR_UNSUPPORTED;
        pl.fSW = E - 2 * FH; pl.fS = cdiv(W, pl.fSW); pl.fT = E / pl.fK;
        pl.band_h = env_int("B200MRC_FIR_BAND", 256);
        pl.bands = cdiv(H, pl.band_h);
        pl.fsmem = fir_smem_bytes(pl.fT, pl.fK, C);
        if (!get_fir(C, pl.fK, pl.fT) || pl.fsmem > (size_t)di.max_smem_optin) return B200MRC_ERR_UNSUPPORTED;
    }
    // ---- IIR kernel: strips such that every CTA of the batch is resident, widest first
    {
        pl.iK = env_int("B200MRC_IIR_K", 2);
        if (pl.iK != 2 && pl.iK != 4) return B200MRC_ERR_UNSUPPORTED;
        const int forced = env_int("B200MRC_IIR_E", 0);
        int pick = 0; double pick_score = -1;
        for (int E = 64; E <= 384; E += 64) {              // whole warps for K = 2; mailbox writers share a warp
            if (forced && E != forced) continue;
            const int T = E / pl.iK, SW = E - FH;
            const void *kern = get_iir(C, pl.iK, T);
            if (!kern) continue;
            const size_t smem = iir_smem_bytes(T, pl.iK, C);
            if (smem > (size_t)di.max_smem_optin) continue;
            cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (err != cudaSuccess) return (int)err;
            int per_sm = 0;
            err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T, smem);
            if (err != cudaSuccess) return (int)err;
            if (per_sm < 1) continue;
            const double ctas = (double)cdiv(W, SW) * N, cap = (double)per_sm * di.sm_count;
            const bool fits = ctas <= cap;
            const double warps = (fits ? ctas : cap) * ((T + 31) / 32) / di.sm_count;
            // resident work, discounted by halo overhead; a batch that does not fit pays a second wave
            double score = (fits ? ctas : cap) * SW * ((double)SW / (SW + FH)) * (fits ? 1.0 : 0.6);
            if (warps < 8) score *= warps / 8;
            if (score > pick_score) { pick = E; pick_score = score; }
        }
        if (!pick) return B200MRC_ERR_UNSUPPORTED;
        pl.iSW = pick - FH; pl.iS = cdiv(W, pl.iSW); pl.iT = pick / pl.iK; pl.ismem = iir_smem_bytes(pl.iT, pl.iK, C);
    }
    out = pl;
    cW = W; cH = H; cN = N; cC = C; cp = pl;
    return B200MRC_OK;
}

}  // namespace

int launch_opt_iir_warp(const uint8_t *img, int64_t ipitch, int64_t istride, int C,
                        const uint8_t *rec, int64_t rpitch, int64_t rstride,
                        uint8_t *ofg, int64_t fpitch, int64_t fstride,
                        uint8_t *obg, int64_t bpitch, int64_t bstride,
                        int W, int H, int N, uint32_t *mailbox, unsigned *ticket, int wpc, cudaStream_t st);

int launch_opt_iir_ghost(const uint8_t *img, int64_t ipitch, int64_t istride, int C,
                         const uint8_t *rec, int64_t rpitch, int64_t rstride,
                         uint8_t *ofg, int64_t fpitch, int64_t fstride,
                         uint8_t *obg, int64_t bpitch, int64_t bstride,
                         int W, int H, int N, uint32_t *mailbox, unsigned *ticket, int K, int wpc, cudaStream_t st);
void iirg_forget(const void *mailbox);
int64_t iirg_rec_pitch(int W);
int launch_opt_fir_warp(const uint8_t *mask, int64_t mpitch, int64_t mstride,
                        const uint8_t *img, int64_t ipitch, int64_t istride, int C,
                        uint8_t *rec, int64_t rpitch, int64_t rstride,
                        int W, int H, int N, int band_h, int wpc, cudaStream_t st);

// record plane: 8 B / pixel, rows padded to whole 4-pixel groups (optimise_ghost.cu reads whole groups)
size_t optimise_split_rec_bytes(int W, int H, int N)
{
    return (size_t)iirg_rec_pitch(W) * (size_t)H * (size_t)N;
}

// Returns B200MRC_ERR_UNSUPPORTED when this path does not apply (the caller falls back).
int launch_optimise_split(const uint8_t *mask, int64_t mpitch, int64_t mstride,
                          const uint8_t *img, int64_t ipitch, int64_t istride, int C,
                          uint8_t *ofg, int64_t fpitch, int64_t fstride,
                          uint8_t *obg, int64_t bpitch, int64_t bstride,
                          int W, int H, int N, uint8_t *rec, uint32_t *mailbox, int *prog, unsigned *ticket, cudaStream_t st)
{
    auto al16 = [](const void *q) { return ((uintptr_t)q & 15) == 0; };
    const int64_t need_i = ((int64_t)W * C + 15) & ~15ll, need_m = ((int64_t)W + 15) & ~15ll;
    if (!ofg || !obg || !rec) return B200MRC_ERR_UNSUPPORTED;
    if (!al16(mask) || !al16(img) || !al16(ofg) || !al16(obg) || !al16(rec)) return B200MRC_ERR_UNSUPPORTED;
    if ((mpitch | mstride | ipitch | istride | fpitch | fstride | bpitch | bstride) & 15) return B200MRC_ERR_UNSUPPORTED;
    if (mpitch < need_m || ipitch < need_i || fpitch < need_i || bpitch < need_i) return B200MRC_ERR_UNSUPPORTED;
    if (N > 65535) return B200MRC_ERR_UNSUPPORTED;
    SplitPlan pl;
    int rc = plan_split(W, H, N, C, pl);
    if (rc != B200MRC_OK) return rc;
    const int64_t rpitch = iirg_rec_pitch(W), rstride = rpitch * H;
    // IIR sweep: "warp" (default: optimise_warp.cu) | "ghost" (optimise_ghost.cu, experimental) | "cta" (k_opt_iir below)
    const char *iir_sel = getenv("B200MRC_IIR");
    const bool ghost_iir = iir_sel && !strcmp(iir_sel, "ghost");
    const bool warp_iir = ghost_iir || !iir_sel || strcmp(iir_sel, "cta") != 0;   // both read fmt-1 records
    if (!ghost_iir) iirg_forget(mailbox);
    // FIR records: "warp" (default: optimise_firw.cu, fmt 1 only) | "cta" (k_opt_fir below)
    const char *fir_sel = getenv("B200MRC_FIR");
    if (warp_iir && (!fir_sel || strcmp(fir_sel, "cta") != 0)) {
        rc = launch_opt_fir_warp(mask, mpitch, mstride, img, ipitch, istride, C, rec, rpitch, rstride, W, H, N,
                                 env_int("B200MRC_FIRW_BAND", 256), env_int("B200MRC_FIRW_WPC", 4), st);
        if (rc != B200MRC_OK) return rc;
    } else {
        FirParams p;
        p.fmt = warp_iir ? 1 : 0;
        p.mask = mask; p.mpitch = mpitch; p.mstride = mstride; p.img = img; p.ipitch = ipitch; p.istride = istride;
        p.rec = rec; p.rpitch = rpitch; p.rstride = rstride;
        p.W = W; p.H = H; p.S = pl.fS; p.SW = pl.fSW; p.n_bands = pl.bands; p.band_h = pl.band_h;
        const void *kern = get_fir(C, pl.fK, pl.fT);
        B200MRC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.fsmem));
        void *args[] = {(void *)&p};
        { ProfScope _ps("k_opt_fir", st); B200MRC_CUDA_TRY(cudaLaunchKernel(kern, dim3((unsigned)(pl.fS * pl.bands), (unsigned)N), dim3(pl.fT), args, pl.fsmem, st)); }
        count_launch();
    }
    if (ghost_iir)
        return launch_opt_iir_ghost(img, ipitch, istride, C, rec, rpitch, rstride, ofg, fpitch, fstride, obg, bpitch, bstride,
                                    W, H, N, mailbox, ticket, env_int("B200MRC_IIRG_K", 4), env_int("B200MRC_IIRG_WPC", 2), st);
    if (warp_iir)
        return launch_opt_iir_warp(img, ipitch, istride, C, rec, rpitch, rstride, ofg, fpitch, fstride, obg, bpitch, bstride,
                                   W, H, N, mailbox, ticket, env_int("B200MRC_IIRW_WPC", 2), st);
    {
        IirParams p;
        p.img = img; p.ipitch = ipitch; p.istride = istride; p.rec = rec; p.rpitch = rpitch; p.rstride = rstride;
        p.ofg = ofg; p.fpitch = fpitch; p.fstride = fstride; p.obg = obg; p.bpitch = bpitch; p.bstride = bstride;
        p.W = W; p.H = H; p.N = N; p.S = pl.iS; p.SW = pl.iSW;
        p.mailbox = mailbox; p.prog = prog; p.ticket = ticket;
        B200MRC_CUDA_TRY(cudaMemsetAsync(prog, 0, sizeof(int) * (size_t)N * pl.iS, st));
        B200MRC_CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned) * 4, st));
        const void *kern = get_iir(C, pl.iK, pl.iT);
        B200MRC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.ismem));
        void *args[] = {(void *)&p};
        { ProfScope _ps("k_opt_iir", st); B200MRC_CUDA_TRY(cudaLaunchKernel(kern, dim3((unsigned)(pl.iS * N)), dim3(pl.iT), args, pl.ismem, st)); }
        count_launch();
    }
    return B200MRC_OK;
}

}  // namespace b200mrc
