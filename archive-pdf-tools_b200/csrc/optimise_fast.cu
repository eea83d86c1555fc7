// optimise_fast.cu -- k_optimise_3_10: the production form of the fused fg/bg optimise sweep for
// the reference's only configuration, n_fg = 3 / n_bg = 10 (internetarchivepdf/mrc.py:412-415,
// 446-449; semantics: cython/optimiser.pyx:153-429, restated in optimise.cu / oracle orc_optimise).
//
// Same strip pipeline as optimise.cu (ticketed (page, strip) jobs, mailbox + progress counter
// between neighbouring strips), rebuilt around the instruction budget -- the sweep is
// issue-bound, not HBM-bound:
//   * every global byte moves through the TMA: input rows (RGB + mask segments) arrive by
//     cp.async.bulk into a 3-slot smem stage signalled by mbarriers, 3 rows ahead of use;
//     output rows are staged in smem and leave by cp.async.bulk stores (bulk groups);
//   * K (2 or 4) adjacent columns per thread, vector LDS/STS on thread-private ring rows
//     (22 packed RGBM input rows, 4 + 11 packed output rows), ring slots advanced by counters
//     (no div/mod in the row loop);
//   * fg sums live in two 16-bit lanes per word (r|b, g|count: every fg window sum < 2^16);
//     bg sums in three words (r + count<<20, g, b); window sums slide across the thread's K
//     columns; the 20-column bg window is evaluated only by warps that contain a mask pixel;
//   * one division path per pixel: the layer that is computed (fg off the mask, bg on it)
//     selects its numerators/denominator, the truncating division is an exact multiply-high.
#include "common.cuh"
#include "tma.cuh"

namespace b200mrc {
namespace {

constexpr int FH = 16;                  // halo columns each side
constexpr int NFG = 3, NBG = 10;
constexpr int RIN = 2 * NBG + 2;        // input ring rows
constexpr int RFG = NFG + 1, RBG = NBG + 1;
constexpr int STAGES = 3;               // TMA input stage depth (rows in flight)
constexpr int MAXDEN = 4 * NBG * NBG + NBG * NBG;

struct FastParams {
    const uint8_t *mask; int64_t mpitch, mstride;
    const uint8_t *img;  int64_t ipitch, istride;
    uint8_t *ofg; int64_t fpitch, fstride;
    uint8_t *obg; int64_t bpitch, bstride;
    int W, H, N, S, SW;
    uint32_t *mailbox;                  // [N][S][H][2][FH]
    int *prog;                          // [N][S]
    unsigned *ticket;
};

// ---- small helpers ------------------------------------------------------------------------------
template <int K> struct VecK;
template <> struct VecK<2> { using type = uint2; };
template <> struct VecK<4> { using type = uint4; };

template <int K> __device__ __forceinline__ void ldv(const uint32_t *p, uint32_t (&v)[K])
{
    const typename VecK<K>::type t = *reinterpret_cast<const typename VecK<K>::type *>(p);
    const uint32_t *s = reinterpret_cast<const uint32_t *>(&t);
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = s[k];
}
template <int K> __device__ __forceinline__ void stv(uint32_t *p, const uint32_t (&v)[K])
{
    typename VecK<K>::type t;
    uint32_t *s = reinterpret_cast<uint32_t *>(&t);
#pragma unroll
    for (int k = 0; k < K; k++) s[k] = v[k];
    *reinterpret_cast<typename VecK<K>::type *>(p) = t;
}

__device__ __forceinline__ uint32_t lane_rb(uint32_t px) { return __byte_perm(px, 0, 0x4240); }   // r | b<<16
__device__ __forceinline__ uint32_t lane_gm(uint32_t px) { return __byte_perm(px, 0, 0x4341); }   // g | m<<16
__device__ __forceinline__ uint32_t byte_g(uint32_t px) { return __byte_perm(px, 0, 0x4441); }    // g

__device__ __forceinline__ uint32_t div31(uint32_t num, uint32_t m31)
{
    return (uint32_t)(((unsigned long long)num * m31) >> 31);           // exact floor(num/den), see host note
}

template <int K, int C>
__global__ void __launch_bounds__(256) k_optimise_3_10(const FastParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int T = blockDim.x, tid = threadIdx.x;
    const int E = T * K;                                   // columns incl. both halos
    const int SW = p.SW;
    const int rowRGB = (E * C + 15) & ~15, rowM = (E + 15) & ~15, rowOut = (SW * C + 15) & ~15;

    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem);                  // [STAGES]
    int *s_misc = reinterpret_cast<int *>(smem + 32);                     // [0] job [1] known progress
    uint8_t *rawRGB = smem + 64;                                          // [STAGES][rowRGB]
    uint8_t *rawM = rawRGB + STAGES * rowRGB;                             // [STAGES][rowM]
    uint8_t *outSt = rawM + STAGES * rowM;                                // [2][2][rowOut]
    uint32_t *ringIn = reinterpret_cast<uint32_t *>(outSt + 4 * rowOut);  // [RIN][E]
    uint32_t *ringFg = ringIn + RIN * E;                                  // [RFG][E]
    uint32_t *ringBg = ringFg + RFG * E;                                  // [RBG][E]
    uint32_t *ex = ringBg + RBG * E;                                      // [10][E]
    uint32_t *Mtab = ex + 10 * E;                                         // [MAXDEN + 1]

    if (tid == 0) {
        s_misc[0] = (int)atomicAdd(p.ticket, 1u);
        s_misc[1] = 0;
#pragma unroll
        for (int s = 0; s < STAGES; s++) mbar_init(&mbar[s], 1);
        fence_mbar_init();
    }
    // den -> ceil(2^31 / den): floor(num/den) == (num * M) >> 31 for num <= 255*den, den <= 500
    for (int d = 1 + tid; d <= MAXDEN; d += T) Mtab[d] = (uint32_t)((0x80000000ull + d - 1) / (unsigned long long)d);
    if (tid == 0) Mtab[0] = 0;
    __syncthreads();
    const int job = s_misc[0];
    const int page = job / p.S, strip = job - page * p.S;
    const int W = p.W, H = p.H;
    const int x0 = strip * SW;
    const int i0 = tid * K;                                // first local column of this thread
    const int xg = x0 - FH + i0;                           // its global column
    const bool haloL = i0 < FH && strip > 0;
    const bool in_core = i0 >= FH && i0 < FH + SW;          // interior columns (may still be >= W)
    const bool wants_c = i0 < FH + SW;

    const uint8_t *img = p.img + (int64_t)page * p.istride;
    const uint8_t *mask = p.mask + (int64_t)page * p.mstride;
    uint8_t *ofg = p.ofg + (int64_t)page * p.fstride;
    uint8_t *obg = p.obg + (int64_t)page * p.bstride;
    uint32_t *mb_out = p.mailbox + ((int64_t)page * p.S + strip) * (int64_t)H * 2 * FH;
    const uint32_t *mb_in = strip > 0 ? p.mailbox + ((int64_t)page * p.S + strip - 1) * (int64_t)H * 2 * FH : nullptr;
    int *prog_out = p.prog + (int64_t)page * p.S + strip;
    const int *prog_in = strip > 0 ? p.prog + (int64_t)page * p.S + strip - 1 : nullptr;
    const bool has_right = strip + 1 < p.S;
    int known = 0;                                          // thread 0: last observed progress of the left strip

    // TMA geometry of one input row: columns [cs, ce) land at smem column offset (cs - (x0-FH))
    const int cs = max(0, x0 - FH), ce = min(W, x0 + SW + FH);
    const int dcol = cs - (x0 - FH);
    const uint32_t bytesRGB = (uint32_t)(((ce - cs) * C + 15) & ~15), bytesM = (uint32_t)(((ce - cs) + 15) & ~15);
    const int ocols = min(W, x0 + SW) - x0;                                // interior columns that exist
    const uint32_t bytesOut = (uint32_t)((ocols * C + 15) & ~15);

    auto issue_row = [&](int row) {                          // thread 0 only
        const int s = row % STAGES;
        mbar_expect_tx(&mbar[s], bytesRGB + bytesM);
        tma_load(rawRGB + s * rowRGB + dcol * C, img + (int64_t)row * p.ipitch + (int64_t)cs * C, bytesRGB, &mbar[s]);
        tma_load(rawM + s * rowM + dcol, mask + (int64_t)row * p.mpitch + cs, bytesM, &mbar[s]);
    };
    if (tid == 0)
        for (int r = NBG - 1; r < NBG - 1 + STAGES && r < H; r++) issue_row(r);

    bool cv[K];
#pragma unroll
    for (int k = 0; k < K; k++) cv[k] = (xg + k) >= 0 && (xg + k) < W;

    // per-column running sums
    uint32_t Ffg_rb[K], Ffg_gm[K], Cfg_rb[K], Cfg_g[K];
    uint32_t Fbg0[K], Fbg1[K], Fbg2[K], Cbg0[K], Cbg1[K], Cbg2[K];
    uint32_t prev_fg[K], prev_bg[K], pf_fg[K], pf_bg[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        Ffg_rb[k] = Ffg_gm[k] = Cfg_rb[k] = Cfg_g[k] = 0;
        Fbg0[k] = Fbg1[k] = Fbg2[k] = Cbg0[k] = Cbg1[k] = Cbg2[k] = 0;
        prev_fg[k] = prev_bg[k] = pf_fg[k] = pf_bg[k] = 0;
    }
    int pf_row = -1;

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

    // ---- warm-up: rows 0 .. NBG-2 straight from global (once per strip); FIR sums start as the
    //      state "after row -1": fg rows [0, NFG-1), bg rows [0, NBG-1)
    for (int ry = 0; ry < NBG - 1 && ry < H; ry++) {
        uint32_t px[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            px[k] = 0;
            if (cv[k]) {
                const uint8_t *q = img + (int64_t)ry * p.ipitch + (int64_t)(xg + k) * C;
                const uint32_t m = mask[(int64_t)ry * p.mpitch + xg + k] != 0;
                uint32_t v;
                if (C == 1) { v = q[0]; v |= (v << 8) | (v << 16); }
                else v = q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16);
                px[k] = v | (m << 24);
                if (ry < NFG - 1) fir_fg(px[k], k, true);
                fir_bg(px[k], k, true);
            }
        }
        stv<K>(ringIn + ry * E + i0, px);
    }

    // ring cursors (advanced by one each row, wrapped by compare: no div/mod in the loop)
    int s_e9 = (NBG - 1) % RIN, s_e2 = (NFG - 1) % RIN, s_cur = 0, s_l4 = (RIN - NFG - 1) % RIN, s_l11 = (RIN - NBG - 1) % RIN;
    int f_new = 0, f_old = (RFG - NFG) % RFG;               // ringFg: slot of row y-1 (write) and y-4 (read)
    int b_new = 0, b_old = (RBG - NBG) % RBG;
    // row y-1 goes to slot (y-1) % R ; at y = 0 nothing is written/read (guards below)
    f_new = RFG - 1; f_old = (RFG - 1 - NFG + RFG) % RFG;    // y-1 = -1 -> slot R-1 ; y-4 = -4 -> (R-4) % R
    b_new = RBG - 1; b_old = (RBG - 1 - NBG + RBG) % RBG;
    int st_slot = (NBG - 1) % STAGES, st_par = 0;

    for (int y = 0; y < H; y++) {
        const int re9 = y + NBG - 1, re2 = y + NFG - 1, rl4 = y - NFG - 1, rl11 = y - NBG - 1;

        // ---- entering bg row: raw bytes from the TMA stage -> packed RGBM, kept in the input ring
        uint32_t pe9[K];
#pragma unroll
        for (int k = 0; k < K; k++) pe9[k] = 0;
        if (re9 < H) {
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
            stv<K>(ringIn + s_e9 * E + i0, pe9);
        }

        // ---- left halo: out[y-1] of the neighbour strip's last columns
        if (haloL) {
            if (y >= 1) {
                if (pf_row == y - 1) {
#pragma unroll
                    for (int k = 0; k < K; k++) { prev_fg[k] = pf_fg[k]; prev_bg[k] = pf_bg[k]; }
                } else {
                    // published: waited for at the end of row y-1 (L2 loads: L1 is not coherent)
                    const typename VecK<K>::type a = __ldcg(reinterpret_cast<const typename VecK<K>::type *>(mb_in + ((int64_t)(y - 1) * 2 + 0) * FH + i0));
                    const typename VecK<K>::type b = __ldcg(reinterpret_cast<const typename VecK<K>::type *>(mb_in + ((int64_t)(y - 1) * 2 + 1) * FH + i0));
#pragma unroll
                    for (int k = 0; k < K; k++) { prev_fg[k] = reinterpret_cast<const uint32_t *>(&a)[k]; prev_bg[k] = reinterpret_cast<const uint32_t *>(&b)[k]; }
                }
            }
            if (s_misc[1] >= y + 1) {
                const typename VecK<K>::type a = __ldcg(reinterpret_cast<const typename VecK<K>::type *>(mb_in + ((int64_t)y * 2 + 0) * FH + i0));
                const typename VecK<K>::type b = __ldcg(reinterpret_cast<const typename VecK<K>::type *>(mb_in + ((int64_t)y * 2 + 1) * FH + i0));
#pragma unroll
                for (int k = 0; k < K; k++) { pf_fg[k] = reinterpret_cast<const uint32_t *>(&a)[k]; pf_bg[k] = reinterpret_cast<const uint32_t *>(&b)[k]; }
                pf_row = y;
            }
        }

        // ---- slide the column sums
        uint32_t cur[K];
        ldv<K>(ringIn + s_cur * E + i0, cur);
        {
            uint32_t t[K];
            if (re9 < H) {
#pragma unroll
                for (int k = 0; k < K; k++) fir_bg(pe9[k], k, true);
            }
            if (re2 < H) {
                ldv<K>(ringIn + s_e2 * E + i0, t);
#pragma unroll
                for (int k = 0; k < K; k++) fir_fg(t[k], k, true);
            }
            if (rl4 >= 0) {
                ldv<K>(ringIn + s_l4 * E + i0, t);
#pragma unroll
                for (int k = 0; k < K; k++) fir_fg(t[k], k, false);
            }
            if (rl11 >= 0) {
                ldv<K>(ringIn + s_l11 * E + i0, t);
#pragma unroll
                for (int k = 0; k < K; k++) fir_bg(t[k], k, false);
            }
            if (wants_c && y >= 1) {
#pragma unroll
                for (int k = 0; k < K; k++) {
                    if (cv[k]) {
                        Cfg_rb[k] += lane_rb(prev_fg[k]); Cfg_g[k] += byte_g(prev_fg[k]);
                        Cbg0[k] += prev_bg[k] & 0xffu; Cbg1[k] += byte_g(prev_bg[k]); Cbg2[k] += (prev_bg[k] >> 16) & 0xffu;
                    }
                }
                if (rl4 >= 0) {
                    ldv<K>(ringFg + f_old * E + i0, t);
#pragma unroll
                    for (int k = 0; k < K; k++) { Cfg_rb[k] -= lane_rb(t[k]); Cfg_g[k] -= byte_g(t[k]); }
                }
                if (rl11 >= 0) {
                    ldv<K>(ringBg + b_old * E + i0, t);
#pragma unroll
                    for (int k = 0; k < K; k++) { Cbg0[k] -= t[k] & 0xffu; Cbg1[k] -= byte_g(t[k]); Cbg2[k] -= (t[k] >> 16) & 0xffu; }
                }
                // row y-1 of out enters the output rings (columns outside the page hold zeros)
                uint32_t z[K];
#pragma unroll
                for (int k = 0; k < K; k++) z[k] = cv[k] ? prev_fg[k] : 0u;
                stv<K>(ringFg + f_new * E + i0, z);
#pragma unroll
                for (int k = 0; k < K; k++) z[k] = cv[k] ? prev_bg[k] : 0u;
                stv<K>(ringBg + b_new * E + i0, z);
            }
        }

        // ---- publish this column's sums for the neighbours' windows
        {
            uint32_t v[K];
#pragma unroll
            for (int k = 0; k < K; k++) v[k] = Ffg_rb[k] + Cfg_rb[k];
            stv<K>(ex + 0 * E + i0, v);
#pragma unroll
            for (int k = 0; k < K; k++) v[k] = Ffg_gm[k] + Cfg_g[k];
            stv<K>(ex + 1 * E + i0, v);
            stv<K>(ex + 2 * E + i0, Ffg_rb);
            stv<K>(ex + 3 * E + i0, Ffg_gm);
#pragma unroll
            for (int k = 0; k < K; k++) v[k] = Fbg0[k] + Cbg0[k];
            stv<K>(ex + 4 * E + i0, v);
#pragma unroll
            for (int k = 0; k < K; k++) v[k] = Fbg1[k] + Cbg1[k];
            stv<K>(ex + 5 * E + i0, v);
#pragma unroll
            for (int k = 0; k < K; k++) v[k] = Fbg2[k] + Cbg2[k];
            stv<K>(ex + 6 * E + i0, v);
            stv<K>(ex + 7 * E + i0, Fbg0);
            stv<K>(ex + 8 * E + i0, Fbg1);
            stv<K>(ex + 9 * E + i0, Fbg2);
        }
        // the staging row buffer written below was last read by the bulk store of row y-2
        if (tid == 0) tma_wait_read<1>();
        __syncthreads();

        // ---- this row's outputs
        if (in_core) {
            uint32_t any_m = 0;
#pragma unroll
            for (int k = 0; k < K; k++) any_m |= cur[k] >> 24;
            const bool need_bg = __any_sync(__activemask(), any_m != 0);

            // fg windows: L = A[c-3..c-1], R = F[c..c+2]  (16-bit lanes)
            uint32_t nrb[K], ngm[K];
            {
                constexpr int NL = (NFG + K - 1) / K * K;           // columns fetched to the left (multiple of K)
                uint32_t a[NL + K], f[2 * K];
#pragma unroll
                for (int w = 0; w < 2; w++) {
#pragma unroll
                    for (int g = 0; g < NL / K; g++) ldv<K>(ex + w * E + i0 - NL + g * K, *reinterpret_cast<uint32_t(*)[K]>(&a[g * K]));
#pragma unroll
                    for (int k = 0; k < K; k++) a[NL + k] = (w == 0 ? Ffg_rb[k] + Cfg_rb[k] : Ffg_gm[k] + Cfg_g[k]);
#pragma unroll
                    for (int k = 0; k < K; k++) f[k] = (w == 0 ? Ffg_rb[k] : Ffg_gm[k]);
                    ldv<K>(ex + (2 + w) * E + i0 + K, *reinterpret_cast<uint32_t(*)[K]>(&f[K]));
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        const uint32_t s = a[NL + k - 3] + a[NL + k - 2] + a[NL + k - 1] + f[k] + f[k + 1] + f[k + 2];
                        if (w == 0) nrb[k] = s; else ngm[k] = s;
                    }
                }
            }
            uint32_t num_r[K], num_g[K], num_b[K], den[K];
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int x = xg + k;
                num_r[k] = nrb[k] & 0xffffu; num_b[k] = nrb[k] >> 16; num_g[k] = ngm[k] & 0xffffu;
                den[k] = (ngm[k] >> 16) + (uint32_t)((y - max(0, y - NFG)) * (x - max(0, x - NFG)));
            }
            if (need_bg) {
                // bg windows: L = A[c-10..c-1], R = F[c..c+9], slid across the K columns, word by word
                constexpr int NL = (NBG + K - 1) / K * K;           // 10 (K=2) / 12 (K=4)
                constexpr int NR = (NBG - 1 + K - 1) / K * K;       // columns fetched to the right of the own K
                uint32_t res[3][K];
#pragma unroll
                for (int w = 0; w < 3; w++) {
                    uint32_t a[NL + K], f[K + NR];
#pragma unroll
                    for (int g = 0; g < NL / K; g++) ldv<K>(ex + (4 + w) * E + i0 - NL + g * K, *reinterpret_cast<uint32_t(*)[K]>(&a[g * K]));
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        const uint32_t fv = w == 0 ? Fbg0[k] : (w == 1 ? Fbg1[k] : Fbg2[k]);
                        const uint32_t cvv = w == 0 ? Cbg0[k] : (w == 1 ? Cbg1[k] : Cbg2[k]);
                        a[NL + k] = fv + cvv; f[k] = fv;
                    }
#pragma unroll
                    for (int g = 0; g < NR / K; g++) ldv<K>(ex + (7 + w) * E + i0 + K + g * K, *reinterpret_cast<uint32_t(*)[K]>(&f[K + g * K]));
                    uint32_t L = 0, R = 0;
#pragma unroll
                    for (int d = 1; d <= NBG; d++) L += a[NL - d];
#pragma unroll
                    for (int d = 0; d < NBG; d++) R += f[d];
                    res[w][0] = L + R;
#pragma unroll
                    for (int k = 1; k < K; k++) {
                        L += a[NL + k - 1] - a[NL + k - 1 - NBG];
                        R += f[k + NBG - 1] - f[k - 1];
                        res[w][k] = L + R;
                    }
                }
#pragma unroll
                for (int k = 0; k < K; k++) {
                    if (cur[k] >> 24) {
                        const int x = xg + k;
                        num_r[k] = res[0][k] & 0xfffffu; num_g[k] = res[1][k]; num_b[k] = res[2][k];
                        den[k] = (res[0][k] >> 20) + (uint32_t)((y - max(0, y - NBG)) * (x - max(0, x - NBG)));
                    }
                }
            }
            uint32_t ofg_px[K], obg_px[K];
#pragma unroll
            for (int k = 0; k < K; k++) {
                const uint32_t m31 = Mtab[den[k]];                  // den == 0 -> 0 -> result 0
                const uint32_t q = div31(num_r[k], m31) | (div31(num_g[k], m31) << 8) | (div31(num_b[k], m31) << 16);
                const uint32_t rgb = cur[k] & 0xffffffu;
                const bool m = (cur[k] >> 24) != 0;
                ofg_px[k] = m ? rgb : q;
                obg_px[k] = m ? q : rgb;
                prev_fg[k] = ofg_px[k]; prev_bg[k] = obg_px[k];
            }
            // stage the output bytes of this row (interior columns only)
            {
                uint8_t *sf = outSt + ((y & 1) * 2 + 0) * rowOut + (i0 - FH) * C;
                uint8_t *sb = outSt + ((y & 1) * 2 + 1) * rowOut + (i0 - FH) * C;
                if (C == 1) {
#pragma unroll
                    for (int k = 0; k < K; k++) { sf[k] = (uint8_t)ofg_px[k]; sb[k] = (uint8_t)obg_px[k]; }
                } else if (K == 4) {
                    uint32_t *f32 = reinterpret_cast<uint32_t *>(sf), *b32 = reinterpret_cast<uint32_t *>(sb);
                    f32[0] = __byte_perm(ofg_px[0], ofg_px[1], 0x4210); f32[1] = __byte_perm(ofg_px[1], ofg_px[2], 0x5421); f32[2] = __byte_perm(ofg_px[2], ofg_px[3], 0x6542);
                    b32[0] = __byte_perm(obg_px[0], obg_px[1], 0x4210); b32[1] = __byte_perm(obg_px[1], obg_px[2], 0x5421); b32[2] = __byte_perm(obg_px[2], obg_px[3], 0x6542);
                } else {
                    uint16_t *f16 = reinterpret_cast<uint16_t *>(sf), *b16 = reinterpret_cast<uint16_t *>(sb);
                    f16[0] = (uint16_t)ofg_px[0]; f16[1] = (uint16_t)(((ofg_px[0] >> 16) & 0xffu) | ((ofg_px[1] & 0xffu) << 8)); f16[2] = (uint16_t)(ofg_px[1] >> 8);
                    b16[0] = (uint16_t)obg_px[0]; b16[1] = (uint16_t)(((obg_px[0] >> 16) & 0xffu) | ((obg_px[1] & 0xffu) << 8)); b16[2] = (uint16_t)(obg_px[1] >> 8);
                }
            }
            if (has_right && i0 >= SW) {                             // last FH interior columns feed the right neighbour
                // rows <= y-1 were stored one row-step (and two CTA barriers) ago: publishing them now is a
                // release that has nothing left to wait for; the neighbour runs >= 1 row behind
                if (i0 == SW) st_release(prog_out, y);
                typename VecK<K>::type a, b;
#pragma unroll
                for (int k = 0; k < K; k++) { reinterpret_cast<uint32_t *>(&a)[k] = cv[k] ? ofg_px[k] : 0u; reinterpret_cast<uint32_t *>(&b)[k] = cv[k] ? obg_px[k] : 0u; }
                __stcg(reinterpret_cast<typename VecK<K>::type *>(mb_out + ((int64_t)y * 2 + 0) * FH + (i0 - SW)), a);
                __stcg(reinterpret_cast<typename VecK<K>::type *>(mb_out + ((int64_t)y * 2 + 1) * FH + (i0 - SW)), b);
            }
            fence_proxy_async();                                     // staged bytes -> visible to the TMA store
        }

        // ---- the left neighbour must have published row y before row y+1 starts (poll only when the
        //      cached progress is not enough: in steady state the neighbour is rows ahead)
        if (tid == 0 && strip > 0 && y + 1 < H && known < y + 1) {
            int v = ld_acquire(prog_in);
            while (v < y + 1) { __nanosleep(20); v = ld_acquire(prog_in); }
            known = v;
            s_misc[1] = v;
        }
        __syncthreads();
        if (tid == 0) {
            tma_store(ofg + (int64_t)y * p.fpitch + (int64_t)x0 * C, outSt + ((y & 1) * 2 + 0) * rowOut, bytesOut);
            tma_store(obg + (int64_t)y * p.bpitch + (int64_t)x0 * C, outSt + ((y & 1) * 2 + 1) * rowOut, bytesOut);
            tma_commit();
            const int nr = re9 + STAGES;                              // refill the stage slot just consumed
            if (re9 < H && nr < H) issue_row(nr);
        }
        // advance ring cursors
        if (++s_e9 == RIN) s_e9 = 0;
        if (++s_e2 == RIN) s_e2 = 0;
        if (++s_cur == RIN) s_cur = 0;
        if (++s_l4 == RIN) s_l4 = 0;
        if (++s_l11 == RIN) s_l11 = 0;
        if (++f_new == RFG) f_new = 0;
        if (++f_old == RFG) f_old = 0;
        if (++b_new == RBG) b_new = 0;
        if (++b_old == RBG) b_old = 0;
        if (++st_slot == STAGES) { st_slot = 0; st_par ^= 1; }
    }
    if (has_right && i0 == SW) { __threadfence(); st_release(prog_out, H); }   // last row (stored before the final barrier)
    if (tid == 0) tma_wait_all<0>();                                   // smem must outlive the last bulk stores
}

size_t fast_smem_bytes(int T, int K, int C, int SW)
{
    const size_t E = (size_t)T * K;
    const size_t rowRGB = (E * C + 15) & ~(size_t)15, rowM = (E + 15) & ~(size_t)15, rowOut = ((size_t)SW * C + 15) & ~(size_t)15;
    return 64 + STAGES * (rowRGB + rowM) + 4 * rowOut + (size_t)(RIN + RFG + RBG + 10) * E * 4 + (MAXDEN + 1) * 4 + 64;
}

template <int K, int C> const void *fast_kernel() { return (const void *)k_optimise_3_10<K, C>; }

const void *pick_kernel(int K, int C)
{
    if (K == 2) return C == 1 ? fast_kernel<2, 1>() : fast_kernel<2, 3>();
    return C == 1 ? fast_kernel<4, 1>() : fast_kernel<4, 3>();
}

struct FastPlan { int S, SW, T, K; size_t smem; };

int plan_fast(int W, int N, int C, FastPlan &best)
{
    static int cW = -1, cN = -1, cC = -1;
    static FastPlan cplan;
    if (cW == W && cN == N && cC == C) { best = cplan; return B200MRC_OK; }
    const DevInfo &di = dev_info();
    const char *env_k = getenv("B200MRC_OPT_K");
    const char *env_sw = getenv("B200MRC_OPT_SW");
    const int K = env_k ? atoi(env_k) : 2;
    if (K != 2 && K != 4) return B200MRC_ERR_UNSUPPORTED;
    FastPlan pick{0, 0, 0, K, 0};
    double pick_score = -1.0;
    for (int SW = 32; SW <= 480; SW += 16) {
        if (env_sw && SW != atoi(env_sw)) continue;
        FastPlan c;
        c.K = K; c.SW = SW; c.S = cdiv(W, SW); c.T = (SW + 2 * FH) / K;
        if ((SW + 2 * FH) % (32 * K) != 0) continue;                 // whole warps; mailbox writers share a warp
        if (c.T > 256) continue;
        c.smem = fast_smem_bytes(c.T, K, C, SW);
        if (c.smem > (size_t)di.max_smem_optin) continue;
        const void *kern = pick_kernel(K, C);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
        if (e != cudaSuccess) return (int)e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, c.T, c.smem);
        if (e != cudaSuccess) return (int)e;
        if (per_sm < 1) continue;
        // resident useful columns per SM (work in flight), discounted by the halo overhead
        const double ctas = (double)c.S * N, cap = (double)per_sm * di.sm_count;
        const double resident = ctas < cap ? ctas : cap;
        const double score = resident * SW * ((double)SW / (SW + 2 * FH));
        if (score > pick_score) { pick = c; pick_score = score; }
    }
    if (pick.T == 0) return B200MRC_ERR_UNSUPPORTED;
    best = pick;
    cW = W; cN = N; cC = C; cplan = pick;
    return B200MRC_OK;
}

}  // namespace

// Returns B200MRC_ERR_UNSUPPORTED when the fast path does not apply (caller falls back).
int launch_optimise_fast(const uint8_t *mask, int64_t mpitch, int64_t mstride,
                         const uint8_t *img, int64_t ipitch, int64_t istride, int C,
                         uint8_t *ofg, int64_t fpitch, int64_t fstride,
                         uint8_t *obg, int64_t bpitch, int64_t bstride,
                         int W, int H, int N, uint32_t *mailbox, int *prog, unsigned *ticket, cudaStream_t st)
{
    auto al16 = [](const void *q) { return ((uintptr_t)q & 15) == 0; };
    const int64_t need_i = ((int64_t)W * C + 15) & ~15ll, need_m = ((int64_t)W + 15) & ~15ll;
    if (!ofg || !obg) return B200MRC_ERR_UNSUPPORTED;
    if (!al16(mask) || !al16(img) || !al16(ofg) || !al16(obg)) return B200MRC_ERR_UNSUPPORTED;
    if ((mpitch | mstride | ipitch | istride | fpitch | fstride | bpitch | bstride) & 15) return B200MRC_ERR_UNSUPPORTED;
    if (mpitch < need_m || ipitch < need_i || fpitch < need_i || bpitch < need_i) return B200MRC_ERR_UNSUPPORTED;
    if (getenv("B200MRC_OPT_GENERIC")) return B200MRC_ERR_UNSUPPORTED;
    FastPlan plan;
    int rc = plan_fast(W, N, C, plan);
    if (rc != B200MRC_OK) return rc;
    FastParams p;
    p.mask = mask; p.mpitch = mpitch; p.mstride = mstride;
    p.img = img; p.ipitch = ipitch; p.istride = istride;
    p.ofg = ofg; p.fpitch = fpitch; p.fstride = fstride;
    p.obg = obg; p.bpitch = bpitch; p.bstride = bstride;
    p.W = W; p.H = H; p.N = N; p.S = plan.S; p.SW = plan.SW;
    p.mailbox = mailbox; p.prog = prog; p.ticket = ticket;
    B200MRC_CUDA_TRY(cudaMemsetAsync(prog, 0, sizeof(int) * (size_t)N * plan.S, st));
    B200MRC_CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned) * 4, st));
    const void *kern = pick_kernel(plan.K, C);
    B200MRC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    void *args[] = {(void *)&p};
    { ProfScope _ps("k_optimise_3_10", st); B200MRC_CUDA_TRY(cudaLaunchKernel(kern, dim3((unsigned)(plan.S * N)), dim3(plan.T), args, plan.smem, st)); }
    count_launch();
    return B200MRC_OK;
}

}  // namespace b200mrc
