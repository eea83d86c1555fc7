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
//               out[y-n-1] in its IIR column sums, publishes them, one barrier, a 3- (fg) or
//               10-column (bg) left window, one exact multiply-high division, done.  Strips of a
//               page are pipelined through the global mailbox exactly like optimise.cu; the small
//               smem footprint keeps a whole 64-page batch resident in one wave.
// The record plane costs 8 B/px written + 8 B/px read of HBM traffic; both kernels are
// instruction-issue bound, so that traffic is hidden.
#include "common.cuh"
#include "tma.cuh"

namespace b200mrc {
namespace {

constexpr int FH = 16;                  // halo columns (>= 10, multiple of 16 for TMA alignment)
constexpr int NFG = 3, NBG = 10;
constexpr int RIN = 2 * NBG + 2;        // input ring rows (k_opt_fir)
constexpr int RFG = NFG + 1, RBG = NBG + 1;
constexpr int STAGES = 3;              // k_opt_fir input stage depth
constexpr int ISTAGES = 4;             // k_opt_iir input stage depth (rows in flight)
constexpr int L2_AHEAD = 12;           // rows prefetched into L2 ahead of the bulk loads
constexpr int LAG = 8;                 // a strip starts once its left neighbour is this many rows ahead
constexpr int PUB = 4;                 // progress is published every PUB rows
constexpr int MAXDEN = 4 * NBG * NBG + NBG * NBG;
constexpr int K = 2;                    // columns per thread

__device__ __forceinline__ uint32_t lane_rb(uint32_t px) { return __byte_perm(px, 0, 0x4240); }   // r | b<<16
__device__ __forceinline__ uint32_t lane_gm(uint32_t px) { return __byte_perm(px, 0, 0x4341); }   // g | m<<16
__device__ __forceinline__ uint32_t byte_g(uint32_t px) { return __byte_perm(px, 0, 0x4441); }    // g

__device__ __forceinline__ uint2 ld2(const uint32_t *p) { return *reinterpret_cast<const uint2 *>(p); }
__device__ __forceinline__ void st2(uint32_t *p, uint32_t a, uint32_t b) { *reinterpret_cast<uint2 *>(p) = make_uint2(a, b); }

// =================================================================================== k_opt_fir
struct FirParams {
    const uint8_t *mask; int64_t mpitch, mstride;
    const uint8_t *img;  int64_t ipitch, istride;
    uint8_t *rec; int64_t rpitch, rstride;         // 8 B / pixel
    int W, H, S, SW, n_bands, band_h;
};

template <int C, int T>
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
    bool cv[K];
#pragma unroll
    for (int k = 0; k < K; k++) cv[k] = (xg + k) >= 0 && (xg + k) < W;

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

    uint32_t Ffg_rb[K] = {0, 0}, Ffg_gm[K] = {0, 0}, Fbg0[K] = {0, 0}, Fbg1[K] = {0, 0}, Fbg2[K] = {0, 0};

    auto fir_fg = [&](uint32_t px, int k, bool add) {
        if (cv[k] && (px >> 24)) {
            if (add) { Ffg_rb[k] += lane_rb(px); Ffg_gm[k] += lane_gm(px); }
            else     { Ffg_rb[k] -= lane_rb(px); Ffg_gm[k] -= lane_gm(px); }
        }
    };
    auto fir_bg = [&](uint32_t px, int k, bool add) {
        if (cv[k] && !(px >> 24)) {
            const uint32_t w0 = (px & 0xffu) | (1u << 20), g = byte_g(px), b = (px >> 16) & 0xffu;
            if (add) { Fbg0[k] += w0; Fbg1[k] += g; Fbg2[k] += b; }
            else     { Fbg0[k] -= w0; Fbg1[k] -= g; Fbg2[k] -= b; }
        }
    };

    // virtual row loop: starts early enough that the running sums are complete at y = by0
    const int ys0 = rmin - (NBG - 1);
    // ring slot of row r is (r - rmin) % RIN; cursors for rows y+9, y+2, y, y-4, y-11 (valid only when the row is live)
    auto slot_of = [&](int r) { int s = (r - rmin) % RIN; return s < 0 ? s + RIN : s; };
    int s_e9 = slot_of(ys0 + NBG - 1), s_e2 = slot_of(ys0 + NFG - 1), s_cur = slot_of(ys0), s_l4 = slot_of(ys0 - NFG - 1), s_l11 = slot_of(ys0 - NBG - 1);
    int st_slot = 0, st_par = 0;

    for (int y = ys0; y < by1; y++) {
        const int re9 = y + NBG - 1, re2 = y + NFG - 1, rl4 = y - NFG - 1, rl11 = y - NBG - 1;
        const bool emit = y >= by0;

        uint32_t pe9[K] = {0, 0};
        const bool have9 = re9 >= rmin && re9 < rlast;       // == the rows that were issued
        if (have9) {
            mbar_wait(&mbar[st_slot], (uint32_t)st_par);
            const uint8_t *rr = rawRGB + st_slot * rowRGB + i0 * C;
            const uint8_t *rm = rawM + st_slot * rowM + i0;
#pragma unroll
            for (int k = 0; k < K; k++) {
                uint32_t v;
                if (C == 1) { v = rr[k]; v |= (v << 8) | (v << 16); }
                else v = rr[3 * k] | ((uint32_t)rr[3 * k + 1] << 8) | ((uint32_t)rr[3 * k + 2] << 16);
                pe9[k] = cv[k] ? (v | ((uint32_t)(rm[k] != 0) << 24)) : 0u;
            }
            st2(ringIn + s_e9 * E + i0, pe9[0], pe9[1]);
#pragma unroll
            for (int k = 0; k < K; k++) fir_bg(pe9[k], k, true);
        }
        if (re2 >= rmin && re2 < H) {
            const uint2 t = ld2(ringIn + s_e2 * E + i0);
            fir_fg(t.x, 0, true); fir_fg(t.y, 1, true);
        }
        if (rl4 >= rmin) {
            const uint2 t = ld2(ringIn + s_l4 * E + i0);
            fir_fg(t.x, 0, false); fir_fg(t.y, 1, false);
        }
        if (rl11 >= rmin) {
            const uint2 t = ld2(ringIn + s_l11 * E + i0);
            fir_bg(t.x, 0, false); fir_bg(t.y, 1, false);
        }
        uint32_t *exb = ex + (y & 1) * 5 * E;
        if (emit) {
            st2(exb + 0 * E + i0, Ffg_rb[0], Ffg_rb[1]);
            st2(exb + 1 * E + i0, Ffg_gm[0], Ffg_gm[1]);
            st2(exb + 2 * E + i0, Fbg0[0], Fbg0[1]);
            st2(exb + 3 * E + i0, Fbg1[0], Fbg1[1]);
            st2(exb + 4 * E + i0, Fbg2[0], Fbg2[1]);
        }
        if (tid == 0) tma_wait_read<1>();                    // staging buffer (y % 3) is free again
        __syncthreads();
        if (tid == 0) {
            if (y - 1 >= by0) {                              // row y-1 is completely staged now
                tma_store(rec + (int64_t)(y - 1) * p.rpitch + (int64_t)x0 * 8, outSt + ((y - 1) % 3) * rowOut, bytesOut);
                tma_commit();
            }
            const int nr = re9 + STAGES;
            if (have9 && nr < rlast) issue_row(nr);
        }

        if (emit && in_core) {
            const uint2 cur = ld2(ringIn + s_cur * E + i0);
            const uint32_t curk[K] = {cur.x, cur.y};
            const bool need_bg = __any_sync(__activemask(), ((cur.x | cur.y) >> 24) != 0);
            uint32_t num_r[K], num_g[K], num_b[K], den[K];
            {   // fg: sum of F over [c-3, c+3)
                uint32_t s[2][K];
#pragma unroll
                for (int w = 0; w < 2; w++) {
                    const uint32_t *a = exb + w * E + i0;
                    const uint2 m4 = ld2(a - 4), m2 = ld2(a - 2), p2 = ld2(a + 2);
                    const uint32_t o0 = w == 0 ? Ffg_rb[0] : Ffg_gm[0], o1 = w == 0 ? Ffg_rb[1] : Ffg_gm[1];
                    const uint32_t s0 = m4.y + m2.x + m2.y + o0 + o1 + p2.x;       // c0-3 .. c0+2
                    s[w][0] = s0;
                    s[w][1] = s0 - m4.y + p2.y;                                    // c0-2 .. c0+3
                }
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const int x = xg + k;
                    num_r[k] = s[0][k] & 0xffffu; num_b[k] = s[0][k] >> 16; num_g[k] = s[1][k] & 0xffffu;
                    den[k] = (s[1][k] >> 16) + (uint32_t)((y - max(0, y - NFG)) * (x - max(0, x - NFG)));
                }
            }
            if (need_bg) {   // bg: sum of F over [c-10, c+10)
                uint32_t res[3][K];
#pragma unroll
                for (int w = 0; w < 3; w++) {
                    const uint32_t *a = exb + (2 + w) * E + i0;
                    uint32_t acc = 0, first = 0, last = 0;
#pragma unroll
                    for (int d = -NBG; d < NBG; d += 2) {
                        const uint2 v = ld2(a + d);
                        acc += v.x + v.y;
                        if (d == -NBG) first = v.x;
                    }
                    last = ld2(a + NBG).x;
                    res[w][0] = acc;
                    res[w][1] = acc - first + last;
                }
#pragma unroll
                for (int k = 0; k < K; k++) {
                    if (curk[k] >> 24) {
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
                o[2 * k + 1] = (num_g[k] >> 15) | (num_b[k] << 2) | (den[k] << 19) | ((curk[k] >> 24) << 31);
            }
            *reinterpret_cast<uint4 *>(outSt + (y % 3) * rowOut + (i0 - FH) * 8) = make_uint4(o[0], o[1], o[2], o[3]);
            fence_proxy_async();
        }
        if (++s_e9 == RIN) s_e9 = 0;
        if (++s_e2 == RIN) s_e2 = 0;
        if (++s_cur == RIN) s_cur = 0;
        if (++s_l4 == RIN) s_l4 = 0;
        if (++s_l11 == RIN) s_l11 = 0;
        if (have9 && ++st_slot == STAGES) { st_slot = 0; st_par ^= 1; }
    }
    __syncthreads();
    if (tid == 0) {
        tma_store(rec + (int64_t)(by1 - 1) * p.rpitch + (int64_t)x0 * 8, outSt + ((by1 - 1) % 3) * rowOut, bytesOut);
        tma_commit();
        tma_wait_all<0>();
    }
}

size_t fir_smem_bytes(int T, int C, int SW)
{
    const size_t E = (size_t)T * K;
    const size_t rowRGB = (E * C + 15) & ~(size_t)15, rowM = (E + 15) & ~(size_t)15;
    return 64 + STAGES * (rowRGB + rowM) + 3 * (size_t)SW * 8 + (size_t)(RIN + 10) * E * 4 + 64;
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

template <int C, int T>
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
    uint8_t *outSt = rawRec + ISTAGES * rowRec;                               // [3][2][rowOut]
    uint32_t *ringFg = reinterpret_cast<uint32_t *>(outSt + 6 * rowOut);     // [RFG][E]
    uint32_t *ringBg = ringFg + RFG * E;                                     // [RBG][E]
    uint32_t *ex = ringBg + RBG * E;                                         // [2][4][E]
    uint32_t *Mtab = ex + 8 * E;                                            // [MAXDEN + 1]

    if (tid == 0) {
        s_misc[0] = (int)atomicAdd(p.ticket, 1u);
        s_misc[1] = 0;
#pragma unroll
        for (int s = 0; s < ISTAGES; s++) mbar_init(&mbar[s], 1);
        fence_mbar_init();
    }
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
    bool cv[K];
#pragma unroll
    for (int k = 0; k < K; k++) cv[k] = (xg + k) >= 0 && (xg + k) < W;

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
    uint32_t Cfg_rb[K] = {0, 0}, Cfg_g[K] = {0, 0}, Cbg_rb[K] = {0, 0}, Cbg_g[K] = {0, 0};
    const uint32_t cvm[K] = {cv[0] ? 0xffffffffu : 0u, cv[1] ? 0xffffffffu : 0u};
    uint32_t prev_fg[K] = {0, 0}, prev_bg[K] = {0, 0}, pf_fg[K] = {0, 0}, pf_bg[K] = {0, 0};
    int pf_row = -1;
    int f_new = RFG - 1, f_old = (RFG - 1 - NFG + RFG) % RFG;     // slots of rows y-1 (write) and y-4 (read)
    int b_new = RBG - 1, b_old = (RBG - 1 - NBG + RBG) % RBG;
    int st_slot = 0, st_par = 0;

    for (int y = 0; y < H; y++) {
        // ---- left halo: out[y-1] of the neighbour strip's last columns
        if (haloL) {
            if (y >= 1) {
                if (pf_row == y - 1) { prev_fg[0] = pf_fg[0]; prev_fg[1] = pf_fg[1]; prev_bg[0] = pf_bg[0]; prev_bg[1] = pf_bg[1]; }
                else {
                    const uint2 a = __ldcg(reinterpret_cast<const uint2 *>(mb_in + ((int64_t)(y - 1) * 2 + 0) * FH + i0));
                    const uint2 b = __ldcg(reinterpret_cast<const uint2 *>(mb_in + ((int64_t)(y - 1) * 2 + 1) * FH + i0));
                    prev_fg[0] = a.x; prev_fg[1] = a.y; prev_bg[0] = b.x; prev_bg[1] = b.y;
                }
            }
            if (s_misc[1] >= y + 1) {
                const uint2 a = __ldcg(reinterpret_cast<const uint2 *>(mb_in + ((int64_t)y * 2 + 0) * FH + i0));
                const uint2 b = __ldcg(reinterpret_cast<const uint2 *>(mb_in + ((int64_t)y * 2 + 1) * FH + i0));
                pf_fg[0] = a.x; pf_fg[1] = a.y; pf_bg[0] = b.x; pf_bg[1] = b.y;
                pf_row = y;
            }
        }
        // ---- IIR column sums: + out[y-1], - out[y-n-1]
        if (y >= 1) {
#pragma unroll
            for (int k = 0; k < K; k++) {
                const uint32_t pf = prev_fg[k] & cvm[k], pb = prev_bg[k] & cvm[k];      // columns outside the page hold zeros
                prev_fg[k] = pf; prev_bg[k] = pb;
                Cfg_rb[k] += lane_rb(pf); Cfg_g[k] += byte_g(pf);
                Cbg_rb[k] += lane_rb(pb); Cbg_g[k] += byte_g(pb);
            }
            if (y - NFG - 1 >= 0) {
                const uint2 t = ld2(ringFg + f_old * E + i0);
                Cfg_rb[0] -= lane_rb(t.x); Cfg_g[0] -= byte_g(t.x); Cfg_rb[1] -= lane_rb(t.y); Cfg_g[1] -= byte_g(t.y);
            }
            if (y - NBG - 1 >= 0) {
                const uint2 t = ld2(ringBg + b_old * E + i0);
                Cbg_rb[0] -= lane_rb(t.x); Cbg_g[0] -= byte_g(t.x); Cbg_rb[1] -= lane_rb(t.y); Cbg_g[1] -= byte_g(t.y);
            }
            st2(ringFg + f_new * E + i0, prev_fg[0], prev_fg[1]);
            st2(ringBg + b_new * E + i0, prev_bg[0], prev_bg[1]);
        }
        uint32_t *exb = ex + (y & 1) * 4 * E;
        st2(exb + 0 * E + i0, Cfg_rb[0], Cfg_rb[1]);
        st2(exb + 1 * E + i0, Cfg_g[0], Cfg_g[1]);
        st2(exb + 2 * E + i0, Cbg_rb[0], Cbg_rb[1]);
        st2(exb + 3 * E + i0, Cbg_g[0], Cbg_g[1]);

        if (tid == 0) {
            tma_wait_read<1>();                              // staging buffer (y % 3) is free again
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
        if (tid == 0) {
            if (y >= 1) {
                tma_store(ofg + (int64_t)(y - 1) * p.fpitch + (int64_t)x0 * C, outSt + (((y - 1) % 3) * 2 + 0) * rowOut, bytesRGB);
                tma_store(obg + (int64_t)(y - 1) * p.bpitch + (int64_t)x0 * C, outSt + (((y - 1) % 3) * 2 + 1) * rowOut, bytesRGB);
                tma_commit();
            }
        }

        if (in_core) {
            const int li = i0 - FH;                          // interior column index of this thread's first pixel
            const uint4 rc = *reinterpret_cast<const uint4 *>(rawRec + st_slot * rowRec + li * 8);
            const uint8_t *rr = rawRGB + st_slot * rowRGB + li * C;
            uint32_t rgb[K];
#pragma unroll
            for (int k = 0; k < K; k++) {
                if (C == 1) { rgb[k] = rr[k]; rgb[k] |= (rgb[k] << 8) | (rgb[k] << 16); }
                else rgb[k] = rr[3 * k] | ((uint32_t)rr[3 * k + 1] << 8) | ((uint32_t)rr[3 * k + 2] << 16);
            }
            const uint32_t lo[K] = {rc.x, rc.z}, hi[K] = {rc.y, rc.w};
            const bool need_bg = __any_sync(__activemask(), ((hi[0] | hi[1]) >> 31) != 0);
            // fg: IIR = sum of C over [c-3, c)
            uint32_t irb[K], ig[K];
            {
                const uint32_t *a = exb + 0 * E + i0, *b = exb + 1 * E + i0;
                const uint2 a4 = ld2(a - 4), a2 = ld2(a - 2), b4 = ld2(b - 4), b2 = ld2(b - 2);
                irb[0] = a4.y + a2.x + a2.y; irb[1] = a2.x + a2.y + Cfg_rb[0];
                ig[0] = b4.y + b2.x + b2.y;  ig[1] = b2.x + b2.y + Cfg_g[0];
            }
            uint32_t ib[2][K];
            if (need_bg) {   // bg: IIR = sum of C over [c-10, c)
#pragma unroll
                for (int w = 0; w < 2; w++) {
                    const uint32_t *a = exb + (2 + w) * E + i0;
                    uint32_t acc = 0, first = 0;
#pragma unroll
                    for (int d = -NBG; d < 0; d += 2) {
                        const uint2 v = ld2(a + d);
                        acc += v.x + v.y;
                        if (d == -NBG) first = v.x;
                    }
                    const uint32_t own = w == 0 ? Cbg_rb[0] : Cbg_g[0];
                    ib[w][0] = acc;
                    ib[w][1] = acc - first + own;
                }
            }
            uint32_t ofg_px[K], obg_px[K];
#pragma unroll
            for (int k = 0; k < K; k++) {
                const uint32_t m = hi[k] >> 31;
                uint32_t nr = lo[k] & 0x1ffffu, ng = (lo[k] >> 17) | ((hi[k] & 3u) << 15), nb = (hi[k] >> 2) & 0x1ffffu;
                const uint32_t den = min((hi[k] >> 19) & 0xfffu, (uint32_t)MAXDEN);   // clamp: padding columns hold garbage
                const uint32_t srb = m ? ib[0][k] : irb[k], sg = m ? ib[1][k] : ig[k];
                nr += srb & 0xffffu; nb += srb >> 16; ng += sg;
                const uint32_t m31 = Mtab[den];
                const uint32_t q = div31(nr, m31) | (div31(ng, m31) << 8) | (div31(nb, m31) << 16);
                ofg_px[k] = m ? rgb[k] : q;
                obg_px[k] = m ? q : rgb[k];
                prev_fg[k] = ofg_px[k]; prev_bg[k] = obg_px[k];
            }
            {
                uint8_t *sf = outSt + ((y % 3) * 2 + 0) * rowOut + li * C;
                uint8_t *sb = outSt + ((y % 3) * 2 + 1) * rowOut + li * C;
                if (C == 1) {
                    sf[0] = (uint8_t)ofg_px[0]; sf[1] = (uint8_t)ofg_px[1]; sb[0] = (uint8_t)obg_px[0]; sb[1] = (uint8_t)obg_px[1];
                } else {
                    uint16_t *f16 = reinterpret_cast<uint16_t *>(sf), *b16 = reinterpret_cast<uint16_t *>(sb);
                    f16[0] = (uint16_t)ofg_px[0]; f16[1] = (uint16_t)(((ofg_px[0] >> 16) & 0xffu) | ((ofg_px[1] & 0xffu) << 8)); f16[2] = (uint16_t)(ofg_px[1] >> 8);
                    b16[0] = (uint16_t)obg_px[0]; b16[1] = (uint16_t)(((obg_px[0] >> 16) & 0xffu) | ((obg_px[1] & 0xffu) << 8)); b16[2] = (uint16_t)(obg_px[1] >> 8);
                }
            }
            if (has_right && i0 >= SW) {                     // last FH interior columns feed the right neighbour
                // rows <= y-1 were stored a full row-step (one CTA barrier) ago: the release has nothing to wait for
                if (i0 == SW && (y % PUB) == 0) st_release(prog_out, y);
                __stcg(reinterpret_cast<uint2 *>(mb_out + ((int64_t)y * 2 + 0) * FH + (i0 - SW)), make_uint2(cv[0] ? ofg_px[0] : 0u, cv[1] ? ofg_px[1] : 0u));
                __stcg(reinterpret_cast<uint2 *>(mb_out + ((int64_t)y * 2 + 1) * FH + (i0 - SW)), make_uint2(cv[0] ? obg_px[0] : 0u, cv[1] ? obg_px[1] : 0u));
            }
            fence_proxy_async();
        }
        if (tid == 0) {                                      // refill the stage slot consumed by this row
            // (its readers are the in_core threads above: the refill is issued after the NEXT barrier)
        }
        if (++f_new == RFG) f_new = 0;
        if (++f_old == RFG) f_old = 0;
        if (++b_new == RBG) b_new = 0;
        if (++b_old == RBG) b_old = 0;
        // stage slot of row y is re-armed for row y+STAGES only after every thread has consumed it:
        // that is one barrier later, so thread 0 does it at the top of the next iteration's barrier
        if (tid == 0 && y >= 1) {
            const int nr = (y - 1) + ISTAGES;
            if (nr < H) issue_row(nr);                       // slot (y-1) % ISTAGES was consumed before this row's barrier
        }
        if (++st_slot == ISTAGES) { st_slot = 0; st_par ^= 1; }
    }
    __syncthreads();
    if (has_right && i0 == SW) { __threadfence(); st_release(prog_out, H); }
    if (tid == 0) {
        tma_store(ofg + (int64_t)(H - 1) * p.fpitch + (int64_t)x0 * C, outSt + (((H - 1) % 3) * 2 + 0) * rowOut, bytesRGB);
        tma_store(obg + (int64_t)(H - 1) * p.bpitch + (int64_t)x0 * C, outSt + (((H - 1) % 3) * 2 + 1) * rowOut, bytesRGB);
        tma_commit();
        tma_wait_all<0>();
    }
}

size_t iir_smem_bytes(int T, int C, int SW)
{
    const size_t E = (size_t)T * K;
    const size_t rowRGB = ((size_t)SW * C + 15) & ~(size_t)15;
    return 64 + ISTAGES * (rowRGB + (size_t)SW * 8) + 6 * rowRGB + (size_t)(RFG + RBG + 8) * E * 4 + (MAXDEN + 1) * 4 + 64;
}

template <int C> const void *fir_kernel(int T)
{
    switch (T) {
    case 32: return (const void *)k_opt_fir<C, 32>;
    case 64: return (const void *)k_opt_fir<C, 64>;
    case 128: return (const void *)k_opt_fir<C, 128>;
    default: return nullptr;
    }
}
template <int C> const void *iir_kernel(int T)
{
    switch (T) {
    case 32: return (const void *)k_opt_iir<C, 32>;
    case 64: return (const void *)k_opt_iir<C, 64>;
    case 96: return (const void *)k_opt_iir<C, 96>;
    case 128: return (const void *)k_opt_iir<C, 128>;
    case 160: return (const void *)k_opt_iir<C, 160>;
    case 192: return (const void *)k_opt_iir<C, 192>;
    default: return nullptr;
    }
}

struct SplitPlan { int fS, fSW, fT, bands, band_h; size_t fsmem; int iS, iSW, iT; size_t ismem; };

int plan_split(int W, int H, int N, int C, SplitPlan &out)
{
    static int cW = -1, cH = -1, cN = -1, cC = -1;
    static SplitPlan cp;
    if (cW == W && cH == H && cN == N && cC == C) { out = cp; return B200MRC_OK; }
    const DevInfo &di = dev_info();
    SplitPlan pl;
    // ---- FIR kernel: throughput kernel; 224-column strips (E = 256, 128 threads), 256-row bands
    {
        const char *e = getenv("B200MRC_FIR_SW");
        pl.fSW = e ? atoi(e) : (W > 96 ? 224 : (W > 32 ? 96 : 32));
        if (pl.fSW != 224 && pl.fSW != 96 && pl.fSW != 32) return B200MRC_ERR_UNSUPPORTED;
        pl.fS = cdiv(W, pl.fSW); pl.fT = (pl.fSW + 2 * FH) / K;
        const char *eb = getenv("B200MRC_FIR_BAND");
        pl.band_h = eb ? atoi(eb) : 256;
        pl.bands = cdiv(H, pl.band_h);
        pl.fsmem = fir_smem_bytes(pl.fT, C, pl.fSW);
        if (pl.fT > 256 || pl.fsmem > (size_t)di.max_smem_optin) return B200MRC_ERR_UNSUPPORTED;
    }
    // ---- IIR kernel: widest strips that keep every CTA of the batch resident with >= 16 warps/SM
    {
        const char *e = getenv("B200MRC_IIR_SW");
        int pick = 0; double pick_score = -1;
        for (int SW = 48; SW <= 368; SW += 64) {           // FH + SW multiple of 64: whole warps, mailbox writers in one warp
            if (e && SW != atoi(e)) continue;
            const int T = (SW + FH) / K;
            const void *kern = C == 1 ? iir_kernel<1>(T) : iir_kernel<3>(T);
            if (!kern) continue;
            const size_t smem = iir_smem_bytes(T, C, SW);
            if (T > 256 || smem > (size_t)di.max_smem_optin) continue;
            cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (err != cudaSuccess) return (int)err;
            int per_sm = 0;
            err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T, smem);
            if (err != cudaSuccess) return (int)err;
            if (per_sm < 1) continue;
            const double ctas = (double)cdiv(W, SW) * N, cap = (double)per_sm * di.sm_count;
            const bool fits = ctas <= cap;
            const double warps = (fits ? ctas : cap) * (T / 32) / di.sm_count;
            // resident work, discounted by halo overhead; a batch that does not fit pays a second wave
            double score = (fits ? ctas : cap) * SW * ((double)SW / (SW + FH)) * (fits ? 1.0 : 0.6);
            if (warps < 8) score *= warps / 8;
            if (score > pick_score) { pick = SW; pick_score = score; }
        }
        if (!pick) return B200MRC_ERR_UNSUPPORTED;
        pl.iSW = pick; pl.iS = cdiv(W, pick); pl.iT = (pick + FH) / K; pl.ismem = iir_smem_bytes(pl.iT, C, pick);
    }
    out = pl;
    cW = W; cH = H; cN = N; cC = C; cp = pl;
    return B200MRC_OK;
}

}  // namespace

size_t optimise_split_rec_bytes(int W, int H, int N)
{
    return align_up((size_t)W * 8, 16) * (size_t)H * (size_t)N;
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
    const int64_t rpitch = (int64_t)align_up((size_t)W * 8, 16), rstride = rpitch * H;
    {
        FirParams p;
        p.mask = mask; p.mpitch = mpitch; p.mstride = mstride; p.img = img; p.ipitch = ipitch; p.istride = istride;
        p.rec = rec; p.rpitch = rpitch; p.rstride = rstride;
        p.W = W; p.H = H; p.S = pl.fS; p.SW = pl.fSW; p.n_bands = pl.bands; p.band_h = pl.band_h;
        const void *kern = C == 1 ? fir_kernel<1>(pl.fT) : fir_kernel<3>(pl.fT);
        if (!kern) return B200MRC_ERR_UNSUPPORTED;
        B200MRC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.fsmem));
        void *args[] = {(void *)&p};
        B200MRC_CUDA_TRY(cudaLaunchKernel(kern, dim3((unsigned)(pl.fS * pl.bands), (unsigned)N), dim3(pl.fT), args, pl.fsmem, st));
        count_launch();
    }
    {
        IirParams p;
        p.img = img; p.ipitch = ipitch; p.istride = istride; p.rec = rec; p.rpitch = rpitch; p.rstride = rstride;
        p.ofg = ofg; p.fpitch = fpitch; p.fstride = fstride; p.obg = obg; p.bpitch = bpitch; p.bstride = bstride;
        p.W = W; p.H = H; p.N = N; p.S = pl.iS; p.SW = pl.iSW;
        p.mailbox = mailbox; p.prog = prog; p.ticket = ticket;
        B200MRC_CUDA_TRY(cudaMemsetAsync(prog, 0, sizeof(int) * (size_t)N * pl.iS, st));
        B200MRC_CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned) * 4, st));
        const void *kern = C == 1 ? iir_kernel<1>(pl.iT) : iir_kernel<3>(pl.iT);
        if (!kern) return B200MRC_ERR_UNSUPPORTED;
        B200MRC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.ismem));
        void *args[] = {(void *)&p};
        B200MRC_CUDA_TRY(cudaLaunchKernel(kern, dim3((unsigned)(pl.iS * N)), dim3(pl.iT), args, pl.ismem, st));
        count_launch();
    }
    return B200MRC_OK;
}

}  // namespace b200mrc
