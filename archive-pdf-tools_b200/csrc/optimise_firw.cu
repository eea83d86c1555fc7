// optimise_firw.cu -- k_opt_fir_w: the parallel half of the production optimise path (n_fg = 3 /
// n_bg = 10; internetarchivepdf/mrc.py:412-415, 439-449; semantics cython/optimiser.pyx:153-429).
//
//   out[y,x] = (FIR(y,x) + IIR(y,x)) / den(y,x)        for every pixel not in the layer's mask
//   FIR = sum of mask*img over the 2n x 2n box [y-n,y+n) x [x-n,x+n)  (inputs only: fully parallel)
//   den = #mask in the FIR box + (y-ys)(x-xs)
// Exactly one layer is computed per pixel (fg where mask==0, bg where mask==1), so one 64-bit record
// per pixel carries what the sequential sweep (optimise_warp.cu) needs:
//   fg pixel: 2*Fr | 2*Fb << 16 ; 2*Fg | 4*den << 16          (16-bit lanes, pre-doubled for umulhi)
//   bg pixel: r[0,17) g[17,34) b[34,51) den[51,60) 1 << 63
//
// A warp owns a strip of 104 columns (26 lanes x 4 columns; 3 ghost lanes each side keep the sums of
// the 12 halo columns) x a band of rows of one page, and marches down the band:
//   * rows arrive by lane-private cp.async copies (pixels + mask bytes, DEPTH-1 rows ahead), are packed
//     once (r | g<<8 | b<<16 | layer flags) into a 22-row smem ring the lane alone reads back: no
//     barrier of any kind in the row loop;
//   * column sums of both boxes live in registers in 16-bit lanes (r | b<<16, g | 2*count<<16); a row
//     enters / leaves a box as one multiply-add per word with its 0/1 layer flag;
//   * the 6-column fg window and the two 10-column halves of the bg window are sums of warp shuffles;
//     the bg window is evaluated only in rows where the strip holds a mask pixel.
// The kernel also clears the sweep's mailbox rows (128 B per strip row, one coalesced store per warp and row): the sweep
// validates a hand-off word by its launch tag, so the rows must not hold another stage's bytes that happen to carry the
// tag -- the workspace is shared with other stages and with earlier launches.  Clearing here costs no extra pass.
#include "common.cuh"

namespace b200mrc {
namespace {

constexpr int NFG = 3, NBG = 10;
constexpr int K = 4, GL = 3, RL = 32 - 2 * GL, RW = RL * K;      // 26 real lanes, 104 columns per strip
constexpr int RIN = 2 * NBG + 2;                                  // ring rows
constexpr int DEPTH = 4;                                          // staged rows in flight
constexpr unsigned FULL = 0xffffffffu;

struct FirWParams {
    const uint8_t *mask; int64_t mpitch, mstride;
    const uint8_t *img;  int64_t ipitch, istride;
    uint8_t *rec; int64_t rpitch, rstride;
    int W, H, N, S, n_bands, band_h;
    int64_t jobs;
    uint32_t *mailbox; int S_sweep;      // the sweep's hand-off rows [N][S_sweep][H][32], cleared here (see below)
};

template <int C> struct FirSmem {
    static constexpr int slot = C == 3 ? 16 : 8;                  // staged bytes per lane and row: pixels + 4 mask bytes
    static constexpr int stage_bytes = DEPTH * 32 * slot;
    static constexpr int ring_bytes = RIN * 32 * K * 4;
    static constexpr int warp_bytes = ring_bytes + stage_bytes;
};

__device__ __forceinline__ uint32_t perm(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
__device__ __forceinline__ uint32_t up(uint32_t v, int d) { return __shfl_up_sync(FULL, v, d); }
__device__ __forceinline__ uint32_t dn(uint32_t v, int d) { return __shfl_down_sync(FULL, v, d); }
__device__ __forceinline__ void cp_async4(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int C>
__global__ void __launch_bounds__(128) k_opt_fir_w(const FirWParams p)
{
    using SM = FirSmem<C>;
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t job = (int64_t)blockIdx.x * (blockDim.x >> 5) + wid;
    if (job >= p.jobs) return;
    uint32_t *ring = reinterpret_cast<uint32_t *>(smem + (size_t)wid * SM::warp_bytes);     // [RIN][32*K]
    uint8_t *stage = reinterpret_cast<uint8_t *>(ring) + SM::ring_bytes;                    // [DEPTH][32][slot]

    const int strip = (int)(job % p.S);
    const int band = (int)((job / p.S) % p.n_bands);
    const int page = (int)(job / ((int64_t)p.S * p.n_bands));
    const int W = p.W, H = p.H;
    const int col0 = strip * RW + (lane - GL) * K;
    const int by0 = band * p.band_h, by1 = min(H, by0 + p.band_h);
    const int rmin = max(0, by0 - NBG), rlast = min(H, by1 + NBG - 1), nrows = rlast - rmin;   // rows [rmin, rlast) are streamed

    // per-lane column geometry
    uint32_t vw = 0;                                         // validity bytes of the 4 columns
    uint32_t gx3[K], gx10[K];                                // (x - xs) << 18 for the fg record / plain for the bg record
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int x = col0 + k;
        if (x >= 0 && x < W) vw |= 0xffu << (8 * k);
        gx3[k] = (uint32_t)max(0, min(x, NFG)) << 18;
        gx10[k] = (uint32_t)max(0, min(x, NBG));
    }
    const bool real = lane >= GL && lane < 32 - GL && col0 < W;
    // whole-group copies are legal while the 4-column group stays inside the row pitches
    const bool inside = col0 >= 0 && col0 < W;
    const bool acopy = inside && (int64_t)(col0 + K) * C <= p.ipitch && (int64_t)(col0 + K) <= p.mpitch;
    const bool scopy = inside && !acopy;                     // ragged group that would leave the pitch: guarded byte loads
    const int64_t colc = inside ? col0 : 0;
    const uint8_t *ip = p.img + (int64_t)page * p.istride + colc * C;
    const uint8_t *mp = p.mask + (int64_t)page * p.mstride + colc;
    uint8_t *rp = p.rec + (int64_t)page * p.rstride + colc * 8;
    uint8_t *myslot = stage + lane * SM::slot;
    uint32_t *mbz = (p.mailbox && strip < p.S_sweep)
                        ? p.mailbox + (((int64_t)page * p.S_sweep + strip) * H + by0) * 32 + lane : nullptr;

    // lanes outside the page never copy: their slots stay zero (mask 0, pixels 0) and vw clears the flags
    for (int s = 0; s < DEPTH; s++) {
        if (C == 3) *reinterpret_cast<uint4 *>(myslot + s * 32 * SM::slot) = make_uint4(0, 0, 0, 0);
        else *reinterpret_cast<uint2 *>(myslot + s * 32 * SM::slot) = make_uint2(0, 0);
    }
    const uint8_t *ii = ip + (int64_t)rmin * p.ipitch, *mm = mp + (int64_t)rmin * p.mpitch;   // next row to copy
    int ist = 0;                                             // its stage
    auto issue = [&](int i) {                                // streamed row i -> stage i % DEPTH; one commit group per call
        if (i < nrows) {
            uint8_t *dst = myslot + ist * 32 * SM::slot;
            if (acopy) {
                if (C == 3) { cp_async4(dst, ii); cp_async4(dst + 4, ii + 4); cp_async4(dst + 8, ii + 8); cp_async4(dst + 12, mm); }
                else { cp_async4(dst, ii); cp_async4(dst + 4, mm); }
            } else if (scopy) {
                for (int k = 0; k < K; k++) {
                    const bool v = col0 + k < W;
                    for (int c = 0; c < C; c++) dst[k * C + c] = v ? ii[k * C + c] : 0;
                    dst[K * C + k] = v ? mm[k] : 0;
                }
            }
            ii += p.ipitch; mm += p.mpitch;
            if (++ist == DEPTH) ist = 0;
        }
        cp_async_commit();
    };
    for (int i = 0; i < DEPTH - 1; i++) issue(i);

    // column sums, 16-bit lanes: (r | b << 16, g | 2*count << 16)
    uint32_t Ff_rb[K], Ff_gm[K], Fb_rb[K], Fb_gm[K];
#pragma unroll
    for (int k = 0; k < K; k++) Ff_rb[k] = Ff_gm[k] = Fb_rb[k] = Fb_gm[k] = 0;
    const uint32_t two = 2u;
    auto lanes_rb = [](uint32_t px) { return perm(px, 0, 0x4240); };
    auto lanes_gm = [&](uint32_t px) { return perm(px, two, 0x5451); };       // g | 2 << 16

    // virtual row loop: starts early enough that the running sums are complete at y = by0
    const int ys0 = rmin - (NBG - 1);
    auto slot_of = [&](int r) { int s = (r - rmin) % RIN; return (s < 0 ? s + RIN : s) * 32 * K; };
    int o_e9 = slot_of(ys0 + NBG - 1), o_e2 = slot_of(ys0 + NFG - 1), o_cur = slot_of(ys0),
        o_l4 = slot_of(ys0 - NFG - 1), o_l11 = slot_of(ys0 - NBG - 1);
    constexpr int RINW = RIN * 32 * K;
    int i9 = 0, cst = 0;                                     // streamed index of row y + 9 and its stage
    // warp-uniform row history: bit j = the strip's segment of row (y + 9 - j) holds a mask pixel /
    // holds a pixel that is not a plain bg-layer pixel (mask or outside the page)
    uint32_t hist_m = 0, hist_x = 0;

    for (int y = ys0; y < by1; y++) {
        const int re2 = y + NFG - 1, rl4 = y - NFG - 1, rl11 = y - NBG - 1;
        uint32_t t[K];
        if (i9 < nrows) {
            // ---- row y+9 enters: unpack, flag, keep in the ring, add to the bg sums
            issue(i9 + DEPTH - 1);
            cp_async_wait<DEPTH - 1>();
            const uint8_t *src = myslot + cst * 32 * SM::slot;
            if (++cst == DEPTH) cst = 0;
            uint32_t mb;
            if (C == 3) {
                const uint4 v = *reinterpret_cast<const uint4 *>(src);
                mb = v.w;
                mb = ((((mb & 0x7f7f7f7fu) + 0x7f7f7f7fu) | mb) >> 7) & 0x01010101u;
                const uint32_t fb = (0x40404040u + (mb << 6)) & vw;                 // 0x40: bg layer, 0x80: fg layer, 0: outside
                t[0] = perm(v.x, fb, 0x4210);
                t[1] = perm(perm(v.x, v.y, 0x0543), fb, 0x5210);
                t[2] = perm(perm(v.y, v.z, 0x0432), fb, 0x6210);
                t[3] = perm(v.z, fb, 0x7321);
            } else {
                const uint2 v = *reinterpret_cast<const uint2 *>(src);
                mb = v.y;
                mb = ((((mb & 0x7f7f7f7fu) + 0x7f7f7f7fu) | mb) >> 7) & 0x01010101u;
                const uint32_t fb = (0x40404040u + (mb << 6)) & vw;
                t[0] = perm(v.x, fb, 0x4000); t[1] = perm(v.x, fb, 0x5111); t[2] = perm(v.x, fb, 0x6222); t[3] = perm(v.x, fb, 0x7333);
            }
            *reinterpret_cast<uint4 *>(ring + o_e9 + lane * K) = make_uint4(t[0], t[1], t[2], t[3]);
            const uint32_t tor = t[0] | t[1] | t[2] | t[3], tand = t[0] & t[1] & t[2] & t[3];
            hist_m = (hist_m << 1) | (__any_sync(FULL, (int)tor < 0) ? 1u : 0u);
            hist_x = (hist_x << 1) | (__all_sync(FULL, (tor >> 30) == 1u && ((tand >> 30) & 1u)) ? 0u : 1u);
            if (!(hist_x & 1u)) {                            // every pixel of the segment is a bg-layer pixel: no flag arithmetic
#pragma unroll
                for (int k = 0; k < K; k++) { Fb_rb[k] += lanes_rb(t[k]); Fb_gm[k] += lanes_gm(t[k]); }
            } else {
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const uint32_t m = (t[k] >> 30) & 1u;
                    Fb_rb[k] += lanes_rb(t[k]) * m; Fb_gm[k] += lanes_gm(t[k]) * m;
                }
            }
            i9++;
        } else {
            hist_m <<= 1; hist_x <<= 1;
        }
        // rows y+2 / y-4 entered 7 / 13 steps ago; rows y-11 entered 20 steps ago
        if (re2 >= rmin && re2 < H && (hist_m & (1u << 7))) {
            const uint4 v = *reinterpret_cast<const uint4 *>(ring + o_e2 + lane * K);
            t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
#pragma unroll
            for (int k = 0; k < K; k++) {
                const uint32_t m = t[k] >> 31;
                Ff_rb[k] += lanes_rb(t[k]) * m; Ff_gm[k] += lanes_gm(t[k]) * m;
            }
        }
        if (rl4 >= rmin && (hist_m & (1u << 13))) {
            const uint4 v = *reinterpret_cast<const uint4 *>(ring + o_l4 + lane * K);
            t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
#pragma unroll
            for (int k = 0; k < K; k++) {
                const uint32_t m = t[k] >> 31;
                Ff_rb[k] -= lanes_rb(t[k]) * m; Ff_gm[k] -= lanes_gm(t[k]) * m;
            }
        }
        if (rl11 >= rmin) {
            const uint4 v = *reinterpret_cast<const uint4 *>(ring + o_l11 + lane * K);
            t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
            if (!(hist_x & (1u << 20))) {
#pragma unroll
                for (int k = 0; k < K; k++) { Fb_rb[k] -= lanes_rb(t[k]); Fb_gm[k] -= lanes_gm(t[k]); }
            } else {
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const uint32_t m = (t[k] >> 30) & 1u;
                    Fb_rb[k] -= lanes_rb(t[k]) * m; Fb_gm[k] -= lanes_gm(t[k]) * m;
                }
            }
        }

        if (y >= by0) {
            // ---- emit the records of row y
            uint32_t cur[K] = {0, 0, 0, 0};
            bool need_bg = false;
            if (hist_m & (1u << 9)) {                        // row y entered 9 steps ago with a mask pixel in the segment
                const uint4 cv = *reinterpret_cast<const uint4 *>(ring + o_cur + lane * K);
                cur[0] = cv.x; cur[1] = cv.y; cur[2] = cv.z; cur[3] = cv.w;
                need_bg = __any_sync(FULL, real && ((int)(cur[0] | cur[1] | cur[2] | cur[3]) < 0));
            }
            // fg: sum of Ff over [c-3, c+3)
            uint32_t s_rb[K], s_gm[K];
            {
                const uint32_t Tr = (Ff_rb[0] + Ff_rb[1]) + (Ff_rb[2] + Ff_rb[3]), Tg = (Ff_gm[0] + Ff_gm[1]) + (Ff_gm[2] + Ff_gm[3]);
                if (__any_sync(FULL, (Tr | Tg) != 0)) {
                    const uint32_t l1r = up(Ff_rb[3], 1), l2r = up(Ff_rb[2], 1), l3r = up(Ff_rb[1], 1), r0r = dn(Ff_rb[0], 1), r1r = dn(Ff_rb[1], 1);
                    const uint32_t l1g = up(Ff_gm[3], 1), l2g = up(Ff_gm[2], 1), l3g = up(Ff_gm[1], 1), r0g = dn(Ff_gm[0], 1), r1g = dn(Ff_gm[1], 1);
                    const uint32_t ar = l2r + l1r, ag = l2g + l1g;
                    s_rb[0] = ar + l3r + (Tr - Ff_rb[3]); s_gm[0] = ag + l3g + (Tg - Ff_gm[3]);
                    s_rb[1] = ar + Tr;                    s_gm[1] = ag + Tg;
                    s_rb[2] = l1r + Tr + r0r;             s_gm[2] = l1g + Tg + r0g;
                    s_rb[3] = Tr + r0r + r1r;             s_gm[3] = Tg + r0g + r1g;
                } else {
#pragma unroll
                    for (int k = 0; k < K; k++) s_rb[k] = s_gm[k] = 0;
                }
            }
            const uint32_t gy3 = (uint32_t)min(y, NFG);
            uint32_t o[2 * K];
#pragma unroll
            for (int k = 0; k < K; k++) {
                // 2*Fr | 2*Fb << 16 ; 2*Fg | 4*den << 16 with den = count + (y-ys)(x-xs): the count lane holds 2*count
                o[2 * k] = s_rb[k] + s_rb[k];
                o[2 * k + 1] = (s_gm[k] + s_gm[k]) + gy3 * gx3[k];
            }
            if (need_bg) {
                // bg: sum of Fb over [c-10, c+10) as two 10-column halves (each fits the 16-bit lanes)
                uint32_t h_rb[2][K], h_gm[2][K];
                auto halves = [&](const uint32_t (&F)[K], uint32_t (&h)[2][K]) {
                    const uint32_t T = (F[0] + F[1]) + (F[2] + F[3]);
                    const uint32_t AB = up(T, 1) + up(T, 2), U = up(F[2] + F[3], 3), V = up(F[3], 3), X = up(F[0], 2);
                    const uint32_t D1 = dn(T, 1), D2 = dn(T, 2), E = dn(F[0] + F[1], 2), Fq = dn(F[2], 2), Y = dn(F[0], 3);
                    const uint32_t c01 = F[0] + F[1];
                    h[0][0] = AB + U;         h[0][1] = AB + V + F[0];  h[0][2] = AB + c01;  h[0][3] = AB - X + c01 + F[2];
                    h[1][0] = T + D1 + E;     h[1][1] = (T - F[0]) + D1 + E + Fq;
                    h[1][2] = (F[2] + F[3]) + D1 + D2;                  h[1][3] = F[3] + D1 + D2 + Y;
                };
                halves(Fb_rb, h_rb);
                halves(Fb_gm, h_gm);
                const uint32_t gy10 = (uint32_t)min(y, NBG);
#pragma unroll
                for (int k = 0; k < K; k++) {
                    if ((int)cur[k] < 0) {
                        const uint32_t nr = (h_rb[0][k] & 0xffffu) + (h_rb[1][k] & 0xffffu), nb = (h_rb[0][k] >> 16) + (h_rb[1][k] >> 16);
                        const uint32_t ng = (h_gm[0][k] & 0xffffu) + (h_gm[1][k] & 0xffffu);
                        const uint32_t den = (((h_gm[0][k] >> 16) + (h_gm[1][k] >> 16)) >> 1) + gy10 * gx10[k];
                        o[2 * k] = nr | (ng << 17);
                        o[2 * k + 1] = (ng >> 15) | (nb << 2) | (den << 19) | 0x80000000u;
                    }
                }
            }
            if (mbz) { *mbz = 0u; mbz += 32; }
            if (real) {
                uint4 *dst = reinterpret_cast<uint4 *>(rp + (int64_t)y * p.rpitch);
                dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
                dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
            }
        }
        o_e9 += 32 * K; if (o_e9 == RINW) o_e9 = 0;
        o_e2 += 32 * K; if (o_e2 == RINW) o_e2 = 0;
        o_cur += 32 * K; if (o_cur == RINW) o_cur = 0;
        o_l4 += 32 * K; if (o_l4 == RINW) o_l4 = 0;
        o_l11 += 32 * K; if (o_l11 == RINW) o_l11 = 0;
    }
}

}  // namespace

// Writes fmt-1 records (see the header comment) with row pitch `rpitch` >= round_up(W, 4) * 8.
int launch_opt_fir_warp(const uint8_t *mask, int64_t mpitch, int64_t mstride,
                        const uint8_t *img, int64_t ipitch, int64_t istride, int C,
                        uint8_t *rec, int64_t rpitch, int64_t rstride,
                        int W, int H, int N, int band_h, int wpc, uint32_t *mailbox, int S_sweep, cudaStream_t st)
{
    if (wpc < 1 || wpc > 4 || band_h < 1) return B200MRC_ERR_UNSUPPORTED;
    if (rpitch < (int64_t)((W + 3) / 4 * 4) * 8) return B200MRC_ERR_INVALID;
    FirWParams p;
    p.mask = mask; p.mpitch = mpitch; p.mstride = mstride; p.img = img; p.ipitch = ipitch; p.istride = istride;
    p.rec = rec; p.rpitch = rpitch; p.rstride = rstride;
    p.W = W; p.H = H; p.N = N; p.S = cdiv(W, RW); p.band_h = band_h; p.n_bands = cdiv(H, band_h);
    p.jobs = (int64_t)N * p.S * p.n_bands;
    p.mailbox = mailbox; p.S_sweep = S_sweep;
    const size_t smem = (size_t)wpc * (C == 3 ? FirSmem<3>::warp_bytes : FirSmem<1>::warp_bytes);
    const void *kern = C == 3 ? (const void *)k_opt_fir_w<3> : (const void *)k_opt_fir_w<1>;
    B200MRC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void *args[] = {(void *)&p};
    const int64_t ctas = (p.jobs + wpc - 1) / wpc;
    if (ctas > 0x7fffffffll) return B200MRC_ERR_UNSUPPORTED;
    { ProfScope _ps("k_opt_fir_w", st); B200MRC_CUDA_TRY(cudaLaunchKernel(kern, dim3((unsigned)ctas), dim3(32 * wpc), args, smem, st)); }
    count_launch();
    return B200MRC_OK;
}

}  // namespace b200mrc
