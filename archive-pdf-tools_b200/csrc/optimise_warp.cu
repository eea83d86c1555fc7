// optimise_warp.cu -- k_opt_iir_w: the row-sequential half of the production optimise path
// (n_fg = 3 / n_bg = 10; internetarchivepdf/mrc.py:412-415, 439-449; semantics
// cython/optimiser.pyx:153-429) as free-running WARP strips.
//
//   out[y,x] = (FIR(y,x) + IIR(y,x)) / den(y,x)      FIR, den: record plane written by k_opt_fir_w
//   IIR      = sum of `out` over the n x n box above-left (n = 3 for fg pixels, 10 for bg pixels)
//
// One warp owns a strip of 128 columns (4 adjacent columns per lane) of one page and marches down
// its rows.  Nothing in the row loop is CTA-wide:
//   * window sums come from the left lanes by warp shuffles; only lane 0 needs the neighbouring
//     strip, whose column sums arrive through a global mailbox row (128 B) prefetched two rows ahead;
//   * strips of a page are pipelined: a strip runs a couple of rows behind its left neighbour.  The
//     hand-off needs no fence and no progress counter: every mailbox word carries an 8-bit launch
//     epoch in the spare bits of its two 16-bit lanes (sums are < 4096), so a word is either stale
//     (tag mismatch -> poll again) or complete; jobs are ticketed in dependency order, so any
//     residency is deadlock-free;
//   * every warp feeds itself: its lane 0 issues the TMA bulk loads of its input rows (RGB + 8-byte records, IST
//     rows ahead, one mbarrier per stage) and the bulk stores of its staged fg/bg rows.  (Template parameter ASYNC,
//     tuning key IIRW_FEED=async: lane-private cp.async copies and re-tiled 16-byte stores instead, no barrier and
//     no proxy fence -- kept for A/B runs);
//   * column sums live in registers in 16-bit lanes (r | b << 16, g); the last n output rows are
//     smem rings private to each lane (no synchronisation); rows without a mask pixel in the strip
//     (the common case) take a short path: fg = quotient, bg = copy of the input row.
// Two forms share the arithmetic and the mailbox protocol (launch_opt_iir_warp picks by batch size): k_opt_iir_w, one
// warp per strip as described above, and k_opt_iir_w3 (further down), where the two layers of a strip run on two warps
// fed by a producer thread -- shorter row steps for machines that are not full.
// Followers: with IirWParams::progress set, every strip publishes (st.release.gpu, every 64 rows -- by lane 0 after the
// bulk stores of those rows have completed, in the trio form by the producer thread) how many of its bg rows are
// globally visible, and every CTA executes griddepcontrol.launch_dependents once it is set up: the bg thumbnail pass
// (resample.cu, FOLLOW form), launched behind this kernel with programmatic stream serialization, becomes resident
// beside it only after every CTA of this grid has started -- it fills the issue slots this latency-bound kernel
// leaves idle and can never starve the strips it waits for.
// The truncating division is one multiply-high: floor(num/den) = umulhi(2*num, ceil(2^31/den)),
// exact for num <= 255*den, den <= 500; records of fg pixels arrive pre-doubled in 16-bit lanes
// (optimise_firw.cu) so that numerator assembly is one multiply-add per word.
#include "common.cuh"
#include "tma.cuh"
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>

namespace b200mrc {
namespace {

constexpr int NFG = 3, NBG = 10;
constexpr int K = 4, SWW = 32 * K;      // columns per lane / per warp strip
constexpr int IST = 6;                  // input stages (rows in flight per warp)
constexpr uint32_t TAGMASK = 0xF000F000u;   // mailbox words: bits 12-15 / 28-31 hold the launch epoch
constexpr int MAXDEN = 4 * NBG * NBG + NBG * NBG;
constexpr int MBW = 32;                 // mailbox words per (strip, row)
constexpr unsigned FULL = 0xffffffffu;
constexpr int MTAB_BYTES = 2048;

struct IirWParams {
    const uint8_t *img; int64_t ipitch, istride;
    const uint8_t *rec; int64_t rpitch, rstride;
    uint8_t *ofg; int64_t fpitch, fstride;
    uint8_t *obg; int64_t bpitch, bstride;
    int W, H, N, S;
    uint32_t *mailbox;                  // [N][S][H][MBW]
    unsigned *ticket;
    uint32_t tag;                       // epoch spread over TAGMASK
    unsigned psleep;                    // trio form: producer back-off (ns) when no stage is free
    int *progress;                      // optional [N][S]: rows of the bg output of every strip that are globally visible (followers: resample.cu)
};

template <int C> struct WarpSmem {
    static constexpr int ROWB = SWW * C;                    // bytes of one strip row (384 / 128)
    static constexpr int RGBS = 512;                        // stage bytes for the pixels: [ROWB] (TMA) or [32 lanes][16] (cp.async)
    static constexpr int off_mbar = 0;                      // IST x 8
    static constexpr int off_rgb = 64;                      // [IST][RGBS]
    static constexpr int off_rec = off_rgb + IST * RGBS;    // [IST][SWW * 8]  (TMA: row order; cp.async: [2][32 lanes][16])
    static constexpr int off_ost = off_rec + IST * SWW * 8; // [3][2][ROWB]
    static constexpr int off_rbg = off_ost + 6 * ROWB;      // bg ring, packed px  [NBG][SWW] u32
    static constexpr int off_rfg = off_rbg + NBG * SWW * 4; // fg ring, lanes form [NFG][2][SWW] u32
    static constexpr int off_halo = off_rfg + NFG * 2 * SWW * 4;   // [2][MBW] u32
    static constexpr int bytes = (off_halo + 2 * MBW * 4 + 127) / 128 * 128;
};

__device__ __forceinline__ uint32_t perm(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
__device__ __forceinline__ uint4 ld4(const uint32_t *p) { return *reinterpret_cast<const uint4 *>(p); }
__device__ __forceinline__ void st4(uint32_t *p, const uint32_t *v) { *reinterpret_cast<uint4 *>(p) = make_uint4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void un4(const uint4 v, uint32_t *d) { d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
__device__ __forceinline__ uint4 ld_relaxed4(const uint32_t *p)
{
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed4(uint32_t *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ bool tags_ok(const uint4 v, uint32_t tag)
{
    return ((((v.x ^ tag) | (v.y ^ tag)) | ((v.z ^ tag) | (v.w ^ tag))) & TAGMASK) == 0;
}

// (2*nr, 2*ng, 2*nb, 4*den) -> quotients in lanes form
__device__ __forceinline__ void div3(const uint32_t *Mtab, uint32_t nr2, uint32_t ng2, uint32_t nb2, uint32_t den4,
                                     uint32_t &q_rb, uint32_t &q_g)
{
    den4 = min(den4, (uint32_t)(MAXDEN * 4));
    const uint32_t m31 = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(Mtab) + den4);
    q_rb = __umulhi(nr2, m31) | (__umulhi(nb2, m31) << 16);
    q_g = __umulhi(ng2, m31);
}

// ASYNC = false: inputs by TMA bulk loads + mbarriers, outputs by bulk stores (lane 0 drives them);
// ASYNC = true : inputs by lane-private cp.async copies (no barrier, no proxy fence), outputs re-tiled through smem and
//                stored as 16-byte words by the first lanes -- fewer instructions on the per-row path.
template <int C, bool ASYNC>
__global__ void __launch_bounds__(256) k_opt_iir_w(const IirWParams p)
{
    using SM = WarpSmem<C>;
    constexpr int ROWB = SM::ROWB;
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int wid = (int)__reduce_max_sync(FULL, threadIdx.x >> 5);      // warp-uniform for the compiler: bulk-copy operands stay in uniform registers
    uint32_t *Mtab = reinterpret_cast<uint32_t *>(smem);
    uint8_t *wb = smem + MTAB_BYTES + (size_t)wid * SM::bytes;

    // den -> ceil(2^31 / den)
    for (int d = 1 + (int)threadIdx.x; d <= MAXDEN; d += blockDim.x) Mtab[d] = (uint32_t)((0x80000000ull + d - 1) / (unsigned long long)d);
    if (threadIdx.x == 0) Mtab[0] = 0;
    for (int i = lane; i < SM::bytes / 16; i += 32) reinterpret_cast<uint4 *>(wb)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();                                         // the only CTA-wide barrier
    pdl_launch_dependents();                                 // a programmatic dependent (the bg thumbnail pass) may become resident beside this grid

    uint64_t *mbar = reinterpret_cast<uint64_t *>(wb + SM::off_mbar);
    uint8_t *rgbS = wb + SM::off_rgb, *recS = wb + SM::off_rec, *ost = wb + SM::off_ost;
    uint32_t *ringB = reinterpret_cast<uint32_t *>(wb + SM::off_rbg);
    uint32_t *ringF = reinterpret_cast<uint32_t *>(wb + SM::off_rfg);
    uint32_t *halo = reinterpret_cast<uint32_t *>(wb + SM::off_halo);

    int job = 0;
    if (lane == 0) {
        job = (int)atomicAdd(p.ticket, 1u);
#pragma unroll
        for (int s = 0; s < IST; s++) mbar_init(&mbar[s], 1);
        fence_mbar_init();
    }
    fence_proxy_async();                                     // the zero fill precedes the bulk copies into the same smem
    job = (int)__reduce_max_sync(FULL, (unsigned)job);         // warp-uniform (keeps the bulk-copy operands in uniform registers)
    if (job >= p.N * p.S) return;
    const int page = job / p.S, strip = job - page * p.S;
    const int W = p.W, H = p.H;
    const int x0 = strip * SWW;
    const int ocols = min(W - x0, SWW);
    const uint32_t bytesRGB = (uint32_t)((ocols * C + 15) & ~15), bytesRec = (uint32_t)((ocols * 8 + 15) & ~15);
    const uint8_t *img = p.img + (int64_t)page * p.istride + (int64_t)x0 * C;
    const uint8_t *rec = p.rec + (int64_t)page * p.rstride + (int64_t)x0 * 8;
    uint8_t *ofg = p.ofg + (int64_t)page * p.fstride + (int64_t)x0 * C;
    uint8_t *obg = p.obg + (int64_t)page * p.bstride + (int64_t)x0 * C;
    const bool has_left = strip > 0, has_right = strip + 1 < p.S;
    uint32_t *mb_out = p.mailbox + ((int64_t)page * p.S + strip) * (int64_t)H * MBW;
    const uint32_t *mb_in = has_left ? p.mailbox + ((int64_t)page * p.S + strip - 1) * (int64_t)H * MBW : p.mailbox;
    const uint32_t tag = p.tag;

    auto issue_row = [&](int row, int s) {                   // lane 0
        mbar_expect_tx(&mbar[s], bytesRGB + bytesRec);
        tma_load(rgbS + s * SM::RGBS, img + (int64_t)row * p.ipitch, bytesRGB, &mbar[s]);
        tma_load(recS + s * (SWW * 8), rec + (int64_t)row * p.rpitch, bytesRec, &mbar[s]);
    };
    // cp.async mode: every lane copies the 4 records + 4 pixels of its own columns into its own 16-byte slots
    const int col0 = x0 + lane * K;
    const bool inside = col0 < W;
    const bool acopy = inside && (int64_t)(col0 + K) * C <= p.ipitch && (int64_t)(col0 + K) * 8 <= p.rpitch;
    const uint8_t *a_img = img + (int64_t)lane * K * C, *a_rec = rec + (int64_t)lane * K * 8;    // next row to copy
    int a_st = 0;
    auto issue_async = [&](int row) {                        // one commit group per call, stages in row order
        if (row < H) {
            uint8_t *dr = recS + a_st * (SWW * 8) + lane * 16, *dg = rgbS + a_st * SM::RGBS + lane * 16;
            if (acopy) {
                cp_async16(dr, a_rec); cp_async16(dr + 512, a_rec + 16);
                if (C == 3) { cp_async4(dg, a_img); cp_async4(dg + 4, a_img + 4); cp_async4(dg + 8, a_img + 8); }
                else cp_async4(dg, a_img);
            } else if (inside) {                             // ragged last group that would leave the row pitch: guarded loads
                for (int k = 0; k < K; k++) {
                    const bool v = col0 + k < W;
                    reinterpret_cast<uint2 *>(k < 2 ? dr : dr + 512)[k & 1] = v ? *reinterpret_cast<const uint2 *>(a_rec + 8 * k) : make_uint2(0, 0);
                    for (int c = 0; c < C; c++) dg[k * C + c] = v ? a_img[k * C + c] : 0;
                }
            }
            a_img += p.ipitch; a_rec += p.rpitch;
            if (++a_st == IST) a_st = 0;
        }
        cp_async_commit();
    };
    if (ASYNC) {
        for (int r = 0; r < IST - 1; r++) issue_async(r);
    } else if (lane == 0) {
        for (int r = 0; r < IST && r < H; r++) issue_row(r, r);
    }
    // mailbox rows of the left neighbour are prefetched two rows ahead (lanes 0-7, 16 B each)
    const bool hl = has_left && lane < 8;
    uint4 pfA = make_uint4(0, 0, 0, 0), pfB = make_uint4(0, 0, 0, 0);      // rows y+1 / y+2
    if (hl && 1 < H) pfA = ld_relaxed4(mb_in + (int64_t)1 * MBW + lane * 4);

    // column sums of the last n output rows, lanes form (r | b << 16, g)
    uint32_t Cf_rb[K], Cf_g[K], Cb_rb[K], Cb_g[K];
#pragma unroll
    for (int k = 0; k < K; k++) Cf_rb[k] = Cf_g[k] = Cb_rb[k] = Cb_g[k] = 0;
    int slot = 0, par = 0, ob = 0, rf = 0, rbg = 0, hs = 0;
    uint8_t *o_fg = ofg, *o_bg = obg;                        // output rows (cp.async mode)

    for (int y = 0; y < H; y++) {
        if (hl && y + 2 < H) pfB = ld_relaxed4(mb_in + (int64_t)(y + 2) * MBW + lane * 4);
        if (ASYNC) {
            issue_async(y + IST - 1);                        // into the stage consumed in row y-1
            cp_async_wait<IST - 1>();                        // this lane's copies of row y have landed
        } else {
            if (lane == 0) tma_wait_read<2>();               // output staging buffer `ob` is free again
            __syncwarp();
            mbar_wait(&mbar[slot], (uint32_t)par);           // this row's RGB + records have landed
        }

        // ---- inputs of the lane's 4 pixels
        uint32_t lo[K], hi[K], px[K], img_rb[K], img_g[K];
        {
            const uint32_t *rp = reinterpret_cast<const uint32_t *>(recS + slot * (SWW * 8)) + (ASYNC ? lane * 4 : lane * 8);
            const uint4 a = ld4(rp), b = ld4(rp + (ASYNC ? 128 : 4));
            lo[0] = a.x; hi[0] = a.y; lo[1] = a.z; hi[1] = a.w; lo[2] = b.x; hi[2] = b.y; lo[3] = b.z; hi[3] = b.w;
        }
        uint32_t raw[3];
        if (C == 3) {
            const uint32_t *gp = reinterpret_cast<const uint32_t *>(rgbS + slot * SM::RGBS) + (ASYNC ? lane * 4 : lane * 3);
            raw[0] = gp[0]; raw[1] = gp[1]; raw[2] = gp[2];
            px[0] = raw[0]; px[1] = perm(raw[0], raw[1], 0x5543); px[2] = perm(raw[1], raw[2], 0x4432); px[3] = raw[2] >> 8;
        } else {
            raw[0] = reinterpret_cast<const uint32_t *>(rgbS + slot * SM::RGBS)[ASYNC ? lane * 4 : lane]; raw[1] = raw[2] = 0;
            px[0] = perm(raw[0], 0, 0x4000); px[1] = perm(raw[0], 0, 0x4111); px[2] = perm(raw[0], 0, 0x4222); px[3] = perm(raw[0], 0, 0x4333);
        }
#pragma unroll
        for (int k = 0; k < K; k++) { img_rb[k] = perm(px[k], 0, 0x4240); img_g[k] = perm(px[k], 0, 0x4441); }
        const bool need_bg = __any_sync(FULL, (int)(hi[0] | hi[1] | hi[2] | hi[3]) < 0);

        // ---- fg windows: sums of Cf over the 3 columns to the left
        uint32_t sf_rb[K], sf_g[K];
        {
            uint32_t l1r = __shfl_up_sync(FULL, Cf_rb[3], 1), l2r = __shfl_up_sync(FULL, Cf_rb[2], 1), l3r = __shfl_up_sync(FULL, Cf_rb[1], 1);
            uint32_t l1g = __shfl_up_sync(FULL, Cf_g[3], 1), l2g = __shfl_up_sync(FULL, Cf_g[2], 1), l3g = __shfl_up_sync(FULL, Cf_g[1], 1);
            if (lane == 0) {
                const uint4 hr = ld4(halo + hs * MBW + 24), hg = ld4(halo + hs * MBW + 28);
                l3r = hr.y; l2r = hr.z; l1r = hr.w; l3g = hg.y; l2g = hg.z; l1g = hg.w;
            }
            const uint32_t ar = l2r + l1r, br = Cf_rb[0] + Cf_rb[1], ag = l2g + l1g, bgs = Cf_g[0] + Cf_g[1];
            sf_rb[0] = ar + l3r; sf_rb[1] = ar + Cf_rb[0]; sf_rb[2] = br + l1r; sf_rb[3] = br + Cf_rb[2];
            sf_g[0] = ag + l3g;  sf_g[1] = ag + Cf_g[0];   sf_g[2] = bgs + l1g; sf_g[3] = bgs + Cf_g[2];
        }

        uint32_t of_rb[K], of_g[K], ob_rb[K], ob_g[K], pbg[K];
        if (!need_bg) {
            // ---- no mask pixel in this strip row: fg = quotient everywhere, bg = input row
#pragma unroll
            for (int k = 0; k < K; k++) {
                const uint32_t t_rb = sf_rb[k] * 2u + lo[k], t_gd = sf_g[k] * 2u + hi[k];
                div3(Mtab, t_rb & 0xffffu, t_gd & 0xffffu, t_rb >> 16, t_gd >> 16, of_rb[k], of_g[k]);
                ob_rb[k] = img_rb[k]; ob_g[k] = img_g[k]; pbg[k] = px[k];
            }
        } else {
            // ---- bg windows: sums of Cb over the 10 columns to the left = 2 whole lanes + parts of the 3rd
            uint32_t sb_rb[K], sb_g[K];
            {
                const uint32_t Tr = (Cb_rb[0] + Cb_rb[1]) + (Cb_rb[2] + Cb_rb[3]), Tg = (Cb_g[0] + Cb_g[1]) + (Cb_g[2] + Cb_g[3]);
                const uint32_t Ur = Cb_rb[2] + Cb_rb[3], Ug = Cb_g[2] + Cb_g[3];
                uint32_t A_r = __shfl_up_sync(FULL, Tr, 1), B_r = __shfl_up_sync(FULL, Tr, 2), U_r = __shfl_up_sync(FULL, Ur, 3),
                         V_r = __shfl_up_sync(FULL, Cb_rb[3], 3), X_r = __shfl_up_sync(FULL, Cb_rb[0], 2);
                uint32_t A_g = __shfl_up_sync(FULL, Tg, 1), B_g = __shfl_up_sync(FULL, Tg, 2), U_g = __shfl_up_sync(FULL, Ug, 3),
                         V_g = __shfl_up_sync(FULL, Cb_g[3], 3), X_g = __shfl_up_sync(FULL, Cb_g[0], 2);
                if (lane < 3) {
                    // virtual lanes -1, -2, -3 = lanes 31, 30, 29 of the left strip: halo words [8..11], [4..7], [0..3]
                    const uint32_t *hb = halo + hs * MBW;
                    uint32_t vr[3][4], vg[3][4];
                    un4(ld4(hb + 8), vr[0]); un4(ld4(hb + 4), vr[1]); un4(ld4(hb + 0), vr[2]);
                    un4(ld4(hb + 20), vg[0]); un4(ld4(hb + 16), vg[1]); un4(ld4(hb + 12), vg[2]);
                    uint32_t tr[3], tg[3];
#pragma unroll
                    for (int v = 0; v < 3; v++) {
                        tr[v] = (vr[v][0] + vr[v][1]) + (vr[v][2] + vr[v][3]);
                        tg[v] = (vg[v][0] + vg[v][1]) + (vg[v][2] + vg[v][3]);
                    }
                    // a shift by d at lane j reads virtual lane -(d - j) when j < d
                    if (lane == 0) {
                        A_r = tr[0]; A_g = tg[0]; B_r = tr[1]; B_g = tg[1]; X_r = vr[1][0]; X_g = vg[1][0];
                        U_r = vr[2][2] + vr[2][3]; U_g = vg[2][2] + vg[2][3]; V_r = vr[2][3]; V_g = vg[2][3];
                    } else if (lane == 1) {
                        B_r = tr[0]; B_g = tg[0]; X_r = vr[0][0]; X_g = vg[0][0];
                        U_r = vr[1][2] + vr[1][3]; U_g = vg[1][2] + vg[1][3]; V_r = vr[1][3]; V_g = vg[1][3];
                    } else {
                        U_r = vr[0][2] + vr[0][3]; U_g = vg[0][2] + vg[0][3]; V_r = vr[0][3]; V_g = vg[0][3];
                    }
                }
                const uint32_t ABr = A_r + B_r, ABg = A_g + B_g;
                const uint32_t c01r = Cb_rb[0] + Cb_rb[1], c01g = Cb_g[0] + Cb_g[1];
                sb_rb[0] = ABr + U_r;             sb_g[0] = ABg + U_g;
                sb_rb[1] = ABr + V_r + Cb_rb[0];  sb_g[1] = ABg + V_g + Cb_g[0];
                sb_rb[2] = ABr + c01r;            sb_g[2] = ABg + c01g;
                sb_rb[3] = ABr - X_r + c01r + Cb_rb[2];  sb_g[3] = ABg - X_g + c01g + Cb_g[2];
            }
#pragma unroll
            for (int k = 0; k < K; k++) {
                const bool m = (int)hi[k] < 0;
                // fg-type record (lanes form, doubled) ...
                const uint32_t t_rb = sf_rb[k] * 2u + lo[k], t_gd = sf_g[k] * 2u + hi[k];
                uint32_t nr2 = t_rb & 0xffffu, nb2 = t_rb >> 16, ng2 = t_gd & 0xffffu, den4 = t_gd >> 16;
                // ... or bg-type record (legacy packing): r[0,17) g[17,34) b[34,51) den[51,60)
                if (m) {
                    const uint32_t Fr = lo[k] & 0x1ffffu, Fg = (lo[k] >> 17) | ((hi[k] & 3u) << 15), Fb = (hi[k] >> 2) & 0x1ffffu;
                    nr2 = (Fr + (sb_rb[k] & 0xffffu)) * 2u; nb2 = (Fb + (sb_rb[k] >> 16)) * 2u; ng2 = (Fg + sb_g[k]) * 2u;
                    den4 = ((hi[k] >> 19) & 0xfffu) * 4u;
                }
                uint32_t q_rb, q_g;
                div3(Mtab, nr2, ng2, nb2, den4, q_rb, q_g);
                of_rb[k] = m ? img_rb[k] : q_rb; of_g[k] = m ? img_g[k] : q_g;
                ob_rb[k] = m ? q_rb : img_rb[k]; ob_g[k] = m ? q_g : img_g[k];
                pbg[k] = perm(ob_rb[k], ob_g[k], 0x7240);
            }
        }

        // ---- column sums: + out[y], - out[y-n] (lane-private rings; zero-initialised = rows above the page)
        {
            uint32_t o_rb[K], o_g[K], o_px[K];
            un4(ld4(ringF + (rf * 2 + 0) * SWW + lane * 4), o_rb);
            un4(ld4(ringF + (rf * 2 + 1) * SWW + lane * 4), o_g);
            un4(ld4(ringB + rbg * SWW + lane * 4), o_px);
#pragma unroll
            for (int k = 0; k < K; k++) {
                Cf_rb[k] += of_rb[k] - o_rb[k]; Cf_g[k] += of_g[k] - o_g[k];
                Cb_rb[k] += ob_rb[k] - perm(o_px[k], 0, 0x4240); Cb_g[k] += ob_g[k] - perm(o_px[k], 0, 0x4441);
            }
            st4(ringF + (rf * 2 + 0) * SWW + lane * 4, of_rb);
            st4(ringF + (rf * 2 + 1) * SWW + lane * 4, of_g);
            st4(ringB + rbg * SWW + lane * 4, pbg);
        }

        // ---- stage the two output rows
        {
            uint8_t *sf = ost + (ob * 2 + 0) * ROWB, *sb = ost + (ob * 2 + 1) * ROWB;
            uint32_t pf_[K];
#pragma unroll
            for (int k = 0; k < K; k++) pf_[k] = perm(of_rb[k], of_g[k], 0x7240);
            if (C == 3) {
                uint32_t *d = reinterpret_cast<uint32_t *>(sf) + lane * 3;
                d[0] = perm(pf_[0], pf_[1], 0x4210); d[1] = perm(pf_[1], pf_[2], 0x5421); d[2] = perm(pf_[2], pf_[3], 0x6542);
                uint32_t *e = reinterpret_cast<uint32_t *>(sb) + lane * 3;
                if (!need_bg) { e[0] = raw[0]; e[1] = raw[1]; e[2] = raw[2]; }
                else { e[0] = perm(pbg[0], pbg[1], 0x4210); e[1] = perm(pbg[1], pbg[2], 0x5421); e[2] = perm(pbg[2], pbg[3], 0x6542); }
            } else {
                reinterpret_cast<uint32_t *>(sf)[lane] = perm(perm(pf_[0], pf_[1], 0x0040), perm(pf_[2], pf_[3], 0x0040), 0x5410);
                reinterpret_cast<uint32_t *>(sb)[lane] = need_bg ? perm(perm(pbg[0], pbg[1], 0x0040), perm(pbg[2], pbg[3], 0x0040), 0x5410) : raw[0];
            }
            if (!ASYNC) fence_proxy_async();
        }

        // ---- hand the new column sums (those of row y+1) to the right neighbour
        if (has_right && y + 1 < H && lane >= 29) {
            uint32_t *dst = mb_out + (int64_t)(y + 1) * MBW;
            st_relaxed4(dst + (lane - 29) * 4, Cb_rb[0] | tag, Cb_rb[1] | tag, Cb_rb[2] | tag, Cb_rb[3] | tag);
            st_relaxed4(dst + 12 + (lane - 29) * 4, Cb_g[0] | tag, Cb_g[1] | tag, Cb_g[2] | tag, Cb_g[3] | tag);
            if (lane == 31) {
                st_relaxed4(dst + 24, Cf_rb[0] | tag, Cf_rb[1] | tag, Cf_rb[2] | tag, Cf_rb[3] | tag);
                st_relaxed4(dst + 28, Cf_g[0] | tag, Cf_g[1] | tag, Cf_g[2] | tag, Cf_g[3] | tag);
            }
        }
        // ---- the left neighbour's sums for row y+1: valid once every word carries this launch's tag
        if (has_left && y + 1 < H) {
            bool ok = !hl || tags_ok(pfA, tag);
            while (!__all_sync(FULL, ok)) {
                __nanosleep(32);
                if (hl) { pfA = ld_relaxed4(mb_in + (int64_t)(y + 1) * MBW + lane * 4); ok = tags_ok(pfA, tag); }
            }
            if (hl)
                *reinterpret_cast<uint4 *>(halo + (hs ^ 1) * MBW + lane * 4) =
                    make_uint4(pfA.x & ~TAGMASK, pfA.y & ~TAGMASK, pfA.z & ~TAGMASK, pfA.w & ~TAGMASK);
            pfA = pfB;
        }
        __syncwarp();
        if (ASYNC) {
            // the staged rows leave as 16-byte words (row bytes rounded up to 16, like the bulk stores)
            if (lane * 16 < (int)bytesRGB) {
                const uint4 vf = *reinterpret_cast<const uint4 *>(ost + (ob * 2 + 0) * ROWB + lane * 16);
                const uint4 vb = *reinterpret_cast<const uint4 *>(ost + (ob * 2 + 1) * ROWB + lane * 16);
                *reinterpret_cast<uint4 *>(o_fg + lane * 16) = vf;
                *reinterpret_cast<uint4 *>(o_bg + lane * 16) = vb;
            }
            o_fg += p.fpitch; o_bg += p.bpitch;
        } else if (lane == 0) {
            tma_store(ofg + (int64_t)y * p.fpitch, ost + (ob * 2 + 0) * ROWB, bytesRGB);
            tma_store(obg + (int64_t)y * p.bpitch, ost + (ob * 2 + 1) * ROWB, bytesRGB);
            tma_commit();
            if (y + IST < H) issue_row(y + IST, slot);       // every lane has read this stage (syncwarp above)
        }
        if (p.progress && (y & 63) == 63 && lane == 0) {
            // rows the followers of this sweep may read (every 64 rows: 32 measured 0.05 ms slower per batch).  Bulk stores:
            // all groups but the 8 newest have completed (no stall: they are long done), the release makes them visible.
            // Lanes' 16-byte stores: those of the rows before this one precede this row's warp barrier, hence lane 0's release.
            if (ASYNC) st_release(p.progress + job, y);
            else { tma_wait_all<8>(); st_release(p.progress + job, y - 7); }
        }
        if (++slot == IST) { slot = 0; par ^= 1; }
        if (++ob == 3) ob = 0;
        if (++rf == NFG) rf = 0;
        if (++rbg == NBG) rbg = 0;
        hs ^= 1;
    }
    if (!ASYNC && lane == 0) tma_wait_all<0>();
    __syncwarp();
    if (p.progress && lane == 0) st_release(p.progress + job, H);
}


// ---------------------------------------------------------------------------------------------------------------
// k_opt_iir_w3: the same sweep with a strip served by THREE warps of one CTA.  The fg recurrence (n = 3) and the bg
// recurrence (n = 10) never read each other's outputs, so each gets its own warp and its row step is about half as
// long as the combined one (the sweep is bound by the length of that dependent chain, not by bytes); a third warp only
// feeds them.  The two compute warps share the input stages (RGB + records cross HBM once) and nothing else:
//   * producer warp: per row waits for empty[s] (two arrivals: both compute warps are done with the stage), arms
//     full[s] with the byte count and issues the two TMA bulk loads; runs up to NST rows ahead;
//   * fg / bg warp: waits for full[s], computes its layer, stages its output row in smem and stores it as 16-byte
//     words (no bulk store: no proxy fence, no single-lane section on the row path), then lane 0 arrives on empty[s];
//     each owns its ring, its mailbox words (fg: 24-31, bg: 0-23), its half of the halo, and polls only its own words
//     of the left neighbour's mailbox row (same fence-free tagged hand-off as above).
template <int C, int NST> struct TrioSmem {
    static constexpr int ROWB = SWW * C;
    static constexpr int RGBS = 512;
    static constexpr int off_full = 0;                      // NST x 8
    static constexpr int off_empty = NST * 8;               // NST x 8
    static constexpr int off_job = NST * 16;                // int
    static constexpr int off_rgb = (NST * 16 + 4 + 127) / 128 * 128;   // [NST][RGBS]
    static constexpr int off_rec = off_rgb + NST * RGBS;    // [NST][SWW * 8]
    static constexpr int off_ost = off_rec + NST * SWW * 8; // [2 buffers][2 layers][ROWB]
    static constexpr int off_rbg = off_ost + 4 * ROWB;      // bg ring, packed px  [NBG][SWW] u32
    static constexpr int off_rfg = off_rbg + NBG * SWW * 4; // fg ring, lanes form [NFG][2][SWW] u32
    static constexpr int off_halo = off_rfg + NFG * 2 * SWW * 4;   // [2][MBW] u32
    static constexpr int bytes = (off_halo + 2 * MBW * 4 + 127) / 128 * 128;
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// TPC strips per CTA: warps 2t / 2t+1 are the fg / bg warps of strip t, the last warp feeds all of them.
template <int C, int NST, int TPC>
__global__ void __launch_bounds__(32 * (2 * TPC + 1), TPC == 2 ? 5 : 1) k_opt_iir_w3(const IirWParams p)
{
    using SM = TrioSmem<C, NST>;
    constexpr int ROWB = SM::ROWB;
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int wid = (int)__reduce_max_sync(FULL, threadIdx.x >> 5);      // warp-uniform for the compiler
    const bool is_prod = wid == 2 * TPC;
    const int role = is_prod ? 0 : 1 + (wid & 1);                        // 0: producer, 1: fg warp, 2: bg warp
    uint32_t *Mtab = reinterpret_cast<uint32_t *>(smem);
    uint8_t *wb = smem + MTAB_BYTES + (size_t)(is_prod ? 0 : wid >> 1) * SM::bytes;

    for (int d = 1 + (int)threadIdx.x; d <= MAXDEN; d += blockDim.x) Mtab[d] = (uint32_t)((0x80000000ull + d - 1) / (unsigned long long)d);
    if (threadIdx.x == 0) Mtab[0] = 0;
    {
        const int n16 = TPC * SM::bytes / 16;
        uint4 *z = reinterpret_cast<uint4 *>(smem + MTAB_BYTES);
        for (int i = (int)threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();

    uint64_t *full = reinterpret_cast<uint64_t *>(wb + SM::off_full), *empty = reinterpret_cast<uint64_t *>(wb + SM::off_empty);
    volatile int *jobslot = reinterpret_cast<volatile int *>(wb + SM::off_job);
    uint8_t *rgbS = wb + SM::off_rgb, *recS = wb + SM::off_rec, *ost = wb + SM::off_ost;
    uint32_t *ringB = reinterpret_cast<uint32_t *>(wb + SM::off_rbg);
    uint32_t *ringF = reinterpret_cast<uint32_t *>(wb + SM::off_rfg);
    uint32_t *halo = reinterpret_cast<uint32_t *>(wb + SM::off_halo);

    if (is_prod && lane == 0) {
#pragma unroll
        for (int t = 0; t < TPC; t++) {                      // consecutive tickets: neighbouring strips share the CTA
            uint8_t *wt = smem + MTAB_BYTES + (size_t)t * SM::bytes;
            *reinterpret_cast<volatile int *>(wt + SM::off_job) = (int)atomicAdd(p.ticket, 1u);
            for (int s = 0; s < NST; s++) {
                mbar_init(reinterpret_cast<uint64_t *>(wt + SM::off_full) + s, 1);
                mbar_init(reinterpret_cast<uint64_t *>(wt + SM::off_empty) + s, 2);
            }
        }
        fence_mbar_init();
    }
    fence_proxy_async();                                     // the zero fill precedes the bulk copies into the same smem
    __syncthreads();                                         // tickets + barrier init visible to every warp; last CTA-wide barrier
    pdl_launch_dependents();

    if (is_prod) {
        // =========================================== producer: one thread feeds the CTA's strips ===========================================
        if (lane != 0) return;
        const uint8_t *gi[TPC], *gr[TPC];
        uint32_t nb_rgb[TPC], nb_rec[TPC];
        int yy[TPC], sl[TPC], pr[TPC], jbs[TPC];
#pragma unroll
        for (int t = 0; t < TPC; t++) {
            const int jb = *reinterpret_cast<volatile int *>(smem + MTAB_BYTES + (size_t)t * SM::bytes + SM::off_job);
            const bool valid = jb < p.N * p.S;
            jbs[t] = jb;
            const int pg = valid ? jb / p.S : 0, stp = valid ? jb - pg * p.S : 0;
            const int oc = min(p.W - stp * SWW, SWW);
            nb_rgb[t] = (uint32_t)((oc * C + 15) & ~15); nb_rec[t] = (uint32_t)((oc * 8 + 15) & ~15);
            gi[t] = p.img + (int64_t)pg * p.istride + (int64_t)stp * SWW * C;
            gr[t] = p.rec + (int64_t)pg * p.rstride + (int64_t)stp * SWW * 8;
            yy[t] = valid ? 0 : p.H; sl[t] = 0; pr[t] = 0;
        }
        for (;;) {
            bool any = false, prog = false;
#pragma unroll
            for (int t = 0; t < TPC; t++) {
                if (yy[t] < p.H) {
                    any = true;
                    uint8_t *wt = smem + MTAB_BYTES + (size_t)t * SM::bytes;
                    uint64_t *fl = reinterpret_cast<uint64_t *>(wt + SM::off_full) + sl[t];
                    // a stage is refilled once both compute warps have released the row it held
                    if (yy[t] < NST || mbar_test(reinterpret_cast<uint64_t *>(wt + SM::off_empty) + sl[t], (uint32_t)(pr[t] ^ 1))) {
                        mbar_expect_tx(fl, nb_rgb[t] + nb_rec[t]);
                        tma_load(wt + SM::off_rgb + sl[t] * SM::RGBS, gi[t], nb_rgb[t], fl);
                        tma_load(wt + SM::off_rec + sl[t] * (SWW * 8), gr[t], nb_rec[t], fl);
                        gi[t] += p.ipitch; gr[t] += p.rpitch;
                        // The bg warp arrived on this stage's empty barrier (release.cta) after the warp barrier of row
                        // yy - NST, i.e. after every lane's stores of the bg rows before it; the test above (acquire.cta)
                        // observed that, so this thread's gpu-scope release publishes those rows -- off the sweep's row path.
                        if (p.progress && yy[t] >= NST && (yy[t] & 63) == 0) st_release(p.progress + jbs[t], yy[t] - NST);
                        yy[t]++;
                        if (++sl[t] == NST) { sl[t] = 0; pr[t] ^= 1; }
                        prog = true;
                    }
                }
            }
            if (!any) break;
            if (!prog) __nanosleep(p.psleep);
        }
        return;
    }
    const int job = (int)__reduce_max_sync(FULL, (unsigned)*jobslot);
    if (job >= p.N * p.S) return;
    const int page = job / p.S, strip = job - page * p.S;
    const int W = p.W, H = p.H;
    const int x0 = strip * SWW;
    const int ocols = min(W - x0, SWW);
    const uint32_t bytesRGB = (uint32_t)((ocols * C + 15) & ~15);
    const bool has_left = strip > 0, has_right = strip + 1 < p.S;
    const uint32_t tag = p.tag;
    int slot = 0, par = 0;

    // mailbox rows: mine (row y+1 is written while row y is processed) and the left neighbour's
    uint32_t *mb_o = p.mailbox + ((int64_t)page * p.S + strip) * (int64_t)H * MBW + MBW;
    const uint32_t *mb_i = (has_left ? p.mailbox + ((int64_t)page * p.S + strip - 1) * (int64_t)H * MBW : p.mailbox) + MBW;
    int ob = 0, hs = 0;

    if (role == 1) {
        // =========================================== fg warp ===========================================
        uint8_t *o_fg = p.ofg + (int64_t)page * p.fstride + (int64_t)x0 * C + lane * 16;
        const bool hl = has_left && lane < 2;                // words 24..31 of the neighbour's mailbox row
        mb_i += 24 + lane * 4;                               // row y+1 of the neighbour (this lane's words)
        uint4 pfA = make_uint4(0, 0, 0, 0), pfB = make_uint4(0, 0, 0, 0);      // rows y+1 / y+2
        if (hl && 1 < H) pfA = ld_relaxed4(mb_i);
        uint32_t Cf_rb[K], Cf_g[K];
#pragma unroll
        for (int k = 0; k < K; k++) Cf_rb[k] = Cf_g[k] = 0;
        int rf = 0;

        for (int y = 0; y < H; y++) {
            if (hl && y + 2 < H) pfB = ld_relaxed4(mb_i + MBW);
            mbar_wait(&full[slot], (uint32_t)par);           // this row's RGB + records have landed

            uint32_t lo[K], hi[K];
            {
                const uint32_t *rp = reinterpret_cast<const uint32_t *>(recS + slot * (SWW * 8)) + lane * 8;
                const uint4 a = ld4(rp), b = ld4(rp + 4);
                lo[0] = a.x; hi[0] = a.y; lo[1] = a.z; hi[1] = a.w; lo[2] = b.x; hi[2] = b.y; lo[3] = b.z; hi[3] = b.w;
            }
            const bool any_m = __any_sync(FULL, (int)(hi[0] | hi[1] | hi[2] | hi[3]) < 0);

            // ---- fg windows: sums of Cf over the 3 columns to the left
            uint32_t sf_rb[K], sf_g[K];
            {
                uint32_t l1r = __shfl_up_sync(FULL, Cf_rb[3], 1), l2r = __shfl_up_sync(FULL, Cf_rb[2], 1), l3r = __shfl_up_sync(FULL, Cf_rb[1], 1);
                uint32_t l1g = __shfl_up_sync(FULL, Cf_g[3], 1), l2g = __shfl_up_sync(FULL, Cf_g[2], 1), l3g = __shfl_up_sync(FULL, Cf_g[1], 1);
                if (lane == 0) {
                    const uint4 hr = ld4(halo + hs * MBW + 24), hg = ld4(halo + hs * MBW + 28);
                    l3r = hr.y; l2r = hr.z; l1r = hr.w; l3g = hg.y; l2g = hg.z; l1g = hg.w;
                }
                const uint32_t ar = l2r + l1r, br = Cf_rb[0] + Cf_rb[1], ag = l2g + l1g, bgs = Cf_g[0] + Cf_g[1];
                sf_rb[0] = ar + l3r; sf_rb[1] = ar + Cf_rb[0]; sf_rb[2] = br + l1r; sf_rb[3] = br + Cf_rb[2];
                sf_g[0] = ag + l3g;  sf_g[1] = ag + Cf_g[0];   sf_g[2] = bgs + l1g; sf_g[3] = bgs + Cf_g[2];
            }

            // ---- quotients (records of mask pixels are of the other type: their quotient is discarded below; the
            //      divisor comes straight from the record -- 2 * numerator < 2^16 never carries into its lane)
            uint32_t of_rb[K], of_g[K];
#pragma unroll
            for (int k = 0; k < K; k++) {
                const uint32_t t_rb = sf_rb[k] * 2u + lo[k], t_gd = sf_g[k] * 2u + hi[k];
                div3(Mtab, t_rb & 0xffffu, t_gd & 0xffffu, t_rb >> 16, hi[k] >> 16, of_rb[k], of_g[k]);
            }
            if (any_m) {
                // mask pixels keep the image
                uint32_t px[K];
                if (C == 3) {
                    const uint32_t *gp = reinterpret_cast<const uint32_t *>(rgbS + slot * SM::RGBS) + lane * 3;
                    const uint32_t r0 = gp[0], r1 = gp[1], r2 = gp[2];
                    px[0] = r0; px[1] = perm(r0, r1, 0x5543); px[2] = perm(r1, r2, 0x4432); px[3] = r2 >> 8;
                } else {
                    const uint32_t r0 = reinterpret_cast<const uint32_t *>(rgbS + slot * SM::RGBS)[lane];
                    px[0] = perm(r0, 0, 0x4000); px[1] = perm(r0, 0, 0x4111); px[2] = perm(r0, 0, 0x4222); px[3] = perm(r0, 0, 0x4333);
                }
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const bool m = (int)hi[k] < 0;
                    of_rb[k] = m ? perm(px[k], 0, 0x4240) : of_rb[k];
                    of_g[k] = m ? perm(px[k], 0, 0x4441) : of_g[k];
                }
            }

            // ---- column sums: + out[y], - out[y-3] (lane-private ring; zero-initialised = rows above the page)
            {
                uint32_t o_rb[K], o_g[K];
                un4(ld4(ringF + (rf * 2 + 0) * SWW + lane * 4), o_rb);
                un4(ld4(ringF + (rf * 2 + 1) * SWW + lane * 4), o_g);
#pragma unroll
                for (int k = 0; k < K; k++) { Cf_rb[k] += of_rb[k] - o_rb[k]; Cf_g[k] += of_g[k] - o_g[k]; }
                st4(ringF + (rf * 2 + 0) * SWW + lane * 4, of_rb);
                st4(ringF + (rf * 2 + 1) * SWW + lane * 4, of_g);
            }
            // ---- hand the new column sums (those of row y+1) to the right neighbour
            if (has_right && y + 1 < H && lane == 31) {
                st_relaxed4(mb_o + 24, Cf_rb[0] | tag, Cf_rb[1] | tag, Cf_rb[2] | tag, Cf_rb[3] | tag);
                st_relaxed4(mb_o + 28, Cf_g[0] | tag, Cf_g[1] | tag, Cf_g[2] | tag, Cf_g[3] | tag);
            }
            mb_o += MBW;

            // ---- stage the output row
            uint8_t *sf = ost + (ob * 2 + 0) * ROWB;
            {
                uint32_t pf_[K];
#pragma unroll
                for (int k = 0; k < K; k++) pf_[k] = perm(of_rb[k], of_g[k], 0x7240);
                if (C == 3) {
                    uint32_t *d = reinterpret_cast<uint32_t *>(sf) + lane * 3;
                    d[0] = perm(pf_[0], pf_[1], 0x4210); d[1] = perm(pf_[1], pf_[2], 0x5421); d[2] = perm(pf_[2], pf_[3], 0x6542);
                } else {
                    reinterpret_cast<uint32_t *>(sf)[lane] = perm(perm(pf_[0], pf_[1], 0x0040), perm(pf_[2], pf_[3], 0x0040), 0x5410);
                }
            }
            // ---- the left neighbour's sums for row y+1: valid once every word carries this launch's tag
            if (has_left && y + 1 < H) {
                bool ok = !hl || tags_ok(pfA, tag);
                while (!__all_sync(FULL, ok)) {
                    __nanosleep(32);
                    if (hl) { pfA = ld_relaxed4(mb_i); ok = tags_ok(pfA, tag); }
                }
                if (hl)
                    *reinterpret_cast<uint4 *>(halo + (hs ^ 1) * MBW + 24 + lane * 4) =
                        make_uint4(pfA.x & ~TAGMASK, pfA.y & ~TAGMASK, pfA.z & ~TAGMASK, pfA.w & ~TAGMASK);
                pfA = pfB;
            }
            mb_i += MBW;
            __syncwarp();
            if (lane * 16 < (int)bytesRGB) *reinterpret_cast<uint4 *>(o_fg) = *reinterpret_cast<const uint4 *>(sf + lane * 16);
            o_fg += p.fpitch;
            if (lane == 0) mbar_arrive(&empty[slot]);        // every lane has read this stage (syncwarp above)
            if (++slot == NST) { slot = 0; par ^= 1; }
            ob ^= 1;
            if (++rf == NFG) rf = 0;
            hs ^= 1;
        }
    } else {
        // =========================================== bg warp ===========================================
        uint8_t *o_bg = p.obg + (int64_t)page * p.bstride + (int64_t)x0 * C + lane * 16;
        const bool hl = has_left && lane < 6;                // words 0..23 of the neighbour's mailbox row
        mb_i += lane * 4;
        uint4 pfA = make_uint4(0, 0, 0, 0), pfB = make_uint4(0, 0, 0, 0);
        if (hl && 1 < H) pfA = ld_relaxed4(mb_i);
        uint32_t Cb_rb[K], Cb_g[K];
#pragma unroll
        for (int k = 0; k < K; k++) Cb_rb[k] = Cb_g[k] = 0;
        int rbg = 0;

        for (int y = 0; y < H; y++) {
            if (hl && y + 2 < H) pfB = ld_relaxed4(mb_i + MBW);
            mbar_wait(&full[slot], (uint32_t)par);

            uint32_t lo[K], hi[K], px[K], raw[3];
            {
                const uint32_t *rp = reinterpret_cast<const uint32_t *>(recS + slot * (SWW * 8)) + lane * 8;
                const uint4 a = ld4(rp), b = ld4(rp + 4);
                lo[0] = a.x; hi[0] = a.y; lo[1] = a.z; hi[1] = a.w; lo[2] = b.x; hi[2] = b.y; lo[3] = b.z; hi[3] = b.w;
            }
            if (C == 3) {
                const uint32_t *gp = reinterpret_cast<const uint32_t *>(rgbS + slot * SM::RGBS) + lane * 3;
                raw[0] = gp[0]; raw[1] = gp[1]; raw[2] = gp[2];
                px[0] = raw[0]; px[1] = perm(raw[0], raw[1], 0x5543); px[2] = perm(raw[1], raw[2], 0x4432); px[3] = raw[2] >> 8;
            } else {
                raw[0] = reinterpret_cast<const uint32_t *>(rgbS + slot * SM::RGBS)[lane]; raw[1] = raw[2] = 0;
                px[0] = perm(raw[0], 0, 0x4000); px[1] = perm(raw[0], 0, 0x4111); px[2] = perm(raw[0], 0, 0x4222); px[3] = perm(raw[0], 0, 0x4333);
            }
            const bool need_bg = __any_sync(FULL, (int)(hi[0] | hi[1] | hi[2] | hi[3]) < 0);

            uint32_t ob_rb[K], ob_g[K], pbg[K];
#pragma unroll
            for (int k = 0; k < K; k++) { ob_rb[k] = perm(px[k], 0, 0x4240); ob_g[k] = perm(px[k], 0, 0x4441); pbg[k] = px[k]; }
            if (need_bg) {
                // ---- bg windows: sums of Cb over the 10 columns to the left = 2 whole lanes + parts of the 3rd
                uint32_t sb_rb[K], sb_g[K];
                {
                    const uint32_t Tr = (Cb_rb[0] + Cb_rb[1]) + (Cb_rb[2] + Cb_rb[3]), Tg = (Cb_g[0] + Cb_g[1]) + (Cb_g[2] + Cb_g[3]);
                    const uint32_t Ur = Cb_rb[2] + Cb_rb[3], Ug = Cb_g[2] + Cb_g[3];
                    uint32_t A_r = __shfl_up_sync(FULL, Tr, 1), B_r = __shfl_up_sync(FULL, Tr, 2), U_r = __shfl_up_sync(FULL, Ur, 3),
                             V_r = __shfl_up_sync(FULL, Cb_rb[3], 3), X_r = __shfl_up_sync(FULL, Cb_rb[0], 2);
                    uint32_t A_g = __shfl_up_sync(FULL, Tg, 1), B_g = __shfl_up_sync(FULL, Tg, 2), U_g = __shfl_up_sync(FULL, Ug, 3),
                             V_g = __shfl_up_sync(FULL, Cb_g[3], 3), X_g = __shfl_up_sync(FULL, Cb_g[0], 2);
                    if (lane < 3) {
                        // virtual lanes -1, -2, -3 = lanes 31, 30, 29 of the left strip: halo words [8..11], [4..7], [0..3]
                        const uint32_t *hb = halo + hs * MBW;
                        uint32_t vr[3][4], vg[3][4];
                        un4(ld4(hb + 8), vr[0]); un4(ld4(hb + 4), vr[1]); un4(ld4(hb + 0), vr[2]);
                        un4(ld4(hb + 20), vg[0]); un4(ld4(hb + 16), vg[1]); un4(ld4(hb + 12), vg[2]);
                        uint32_t tr[3], tg[3];
#pragma unroll
                        for (int v = 0; v < 3; v++) {
                            tr[v] = (vr[v][0] + vr[v][1]) + (vr[v][2] + vr[v][3]);
                            tg[v] = (vg[v][0] + vg[v][1]) + (vg[v][2] + vg[v][3]);
                        }
                        if (lane == 0) {
                            A_r = tr[0]; A_g = tg[0]; B_r = tr[1]; B_g = tg[1]; X_r = vr[1][0]; X_g = vg[1][0];
                            U_r = vr[2][2] + vr[2][3]; U_g = vg[2][2] + vg[2][3]; V_r = vr[2][3]; V_g = vg[2][3];
                        } else if (lane == 1) {
                            B_r = tr[0]; B_g = tg[0]; X_r = vr[0][0]; X_g = vg[0][0];
                            U_r = vr[1][2] + vr[1][3]; U_g = vg[1][2] + vg[1][3]; V_r = vr[1][3]; V_g = vg[1][3];
                        } else {
                            U_r = vr[0][2] + vr[0][3]; U_g = vg[0][2] + vg[0][3]; V_r = vr[0][3]; V_g = vg[0][3];
                        }
                    }
                    const uint32_t ABr = A_r + B_r, ABg = A_g + B_g;
                    const uint32_t c01r = Cb_rb[0] + Cb_rb[1], c01g = Cb_g[0] + Cb_g[1];
                    sb_rb[0] = ABr + U_r;             sb_g[0] = ABg + U_g;
                    sb_rb[1] = ABr + V_r + Cb_rb[0];  sb_g[1] = ABg + V_g + Cb_g[0];
                    sb_rb[2] = ABr + c01r;            sb_g[2] = ABg + c01g;
                    sb_rb[3] = ABr - X_r + c01r + Cb_rb[2];  sb_g[3] = ABg - X_g + c01g + Cb_g[2];
                }
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const bool m = (int)hi[k] < 0;
                    // bg-type record (legacy packing): r[0,17) g[17,34) b[34,51) den[51,60); other records: quotient discarded
                    const uint32_t Fr = lo[k] & 0x1ffffu, Fg = (lo[k] >> 17) | ((hi[k] & 3u) << 15), Fb = (hi[k] >> 2) & 0x1ffffu;
                    const uint32_t nr2 = (Fr + (sb_rb[k] & 0xffffu)) * 2u, nb2 = (Fb + (sb_rb[k] >> 16)) * 2u, ng2 = (Fg + sb_g[k]) * 2u;
                    const uint32_t den4 = ((hi[k] >> 19) & 0xfffu) * 4u;
                    uint32_t q_rb, q_g;
                    div3(Mtab, nr2, ng2, nb2, den4, q_rb, q_g);
                    ob_rb[k] = m ? q_rb : ob_rb[k]; ob_g[k] = m ? q_g : ob_g[k];
                    pbg[k] = perm(ob_rb[k], ob_g[k], 0x7240);
                }
            }

            // ---- column sums: + out[y], - out[y-10]
            {
                uint32_t o_px[K];
                un4(ld4(ringB + rbg * SWW + lane * 4), o_px);
#pragma unroll
                for (int k = 0; k < K; k++) { Cb_rb[k] += ob_rb[k] - perm(o_px[k], 0, 0x4240); Cb_g[k] += ob_g[k] - perm(o_px[k], 0, 0x4441); }
                st4(ringB + rbg * SWW + lane * 4, pbg);
            }
            if (has_right && y + 1 < H && lane >= 29) {
                st_relaxed4(mb_o + (lane - 29) * 4, Cb_rb[0] | tag, Cb_rb[1] | tag, Cb_rb[2] | tag, Cb_rb[3] | tag);
                st_relaxed4(mb_o + 12 + (lane - 29) * 4, Cb_g[0] | tag, Cb_g[1] | tag, Cb_g[2] | tag, Cb_g[3] | tag);
            }
            mb_o += MBW;

            // ---- stage the output row
            uint8_t *sb = ost + (ob * 2 + 1) * ROWB;
            if (C == 3) {
                uint32_t *e = reinterpret_cast<uint32_t *>(sb) + lane * 3;
                if (!need_bg) { e[0] = raw[0]; e[1] = raw[1]; e[2] = raw[2]; }
                else { e[0] = perm(pbg[0], pbg[1], 0x4210); e[1] = perm(pbg[1], pbg[2], 0x5421); e[2] = perm(pbg[2], pbg[3], 0x6542); }
            } else {
                reinterpret_cast<uint32_t *>(sb)[lane] = need_bg ? perm(perm(pbg[0], pbg[1], 0x0040), perm(pbg[2], pbg[3], 0x0040), 0x5410) : raw[0];
            }
            if (has_left && y + 1 < H) {
                bool ok = !hl || tags_ok(pfA, tag);
                while (!__all_sync(FULL, ok)) {
                    __nanosleep(32);
                    if (hl) { pfA = ld_relaxed4(mb_i); ok = tags_ok(pfA, tag); }
                }
                if (hl)
                    *reinterpret_cast<uint4 *>(halo + (hs ^ 1) * MBW + lane * 4) =
                        make_uint4(pfA.x & ~TAGMASK, pfA.y & ~TAGMASK, pfA.z & ~TAGMASK, pfA.w & ~TAGMASK);
                pfA = pfB;
            }
            mb_i += MBW;
            __syncwarp();
            if (lane * 16 < (int)bytesRGB) *reinterpret_cast<uint4 *>(o_bg) = *reinterpret_cast<const uint4 *>(sb + lane * 16);
            o_bg += p.bpitch;
            if (lane == 0) mbar_arrive(&empty[slot]);
            if (++slot == NST) { slot = 0; par ^= 1; }
            ob ^= 1;
            if (++rbg == NBG) rbg = 0;
            hs ^= 1;
        }
        __syncwarp();
        if (p.progress && lane == 0) st_release(p.progress + job, H);
    }
}

}  // namespace

size_t iirw_smem_bytes(int C, int wpc) { return MTAB_BYTES + (size_t)wpc * (C == 3 ? WarpSmem<3>::bytes : WarpSmem<1>::bytes); }
size_t iirw3_smem_bytes(int C, int tpc) { return MTAB_BYTES + (size_t)tpc * (C == 3 ? TrioSmem<3, IST>::bytes : TrioSmem<1, IST>::bytes); }

size_t iirw_mailbox_words(int W, int H, int N) { return (size_t)N * cdiv(W, SWW) * (size_t)H * MBW; }

namespace {
// Launch epochs: mailbox words written by launch e carry tag(e) (1..255).  k_opt_fir_w clears every mailbox row of the
// batch right before the sweep (optimise_firw.cu), so a word is either zero (tag 0: poll again) or written by this launch;
// the epoch only guards against a caller who skips the FIR pass.
std::atomic<unsigned> g_epoch{0};
}  // namespace

// The record plane `rec` must hold k_opt_fir's fmt-1 records.  ticket: 1 word.
int launch_opt_iir_warp(const uint8_t *img, int64_t ipitch, int64_t istride, int C,
                        const uint8_t *rec, int64_t rpitch, int64_t rstride,
                        uint8_t *ofg, int64_t fpitch, int64_t fstride,
                        uint8_t *obg, int64_t bpitch, int64_t bstride,
                        int W, int H, int N, uint32_t *mailbox, unsigned *ticket, int wpc, int *progress, cudaStream_t st)
{
    if (wpc < 1 || wpc > 8) return B200MRC_ERR_UNSUPPORTED;
    const unsigned epoch = g_epoch.fetch_add(1u) % 255u + 1u;    // 1..255
    struct { unsigned psleep; int mode, tpc, feed; } env;
    env.psleep = (unsigned)tune(T_IIRW_PSLEEP); env.mode = tune(T_IIRW_MODE); env.feed = tune(T_IIRW_FEED);
    env.tpc = std::min(3, std::max(1, tune(T_IIRW_TPC)));
    IirWParams p;
    p.img = img; p.ipitch = ipitch; p.istride = istride; p.rec = rec; p.rpitch = rpitch; p.rstride = rstride;
    p.ofg = ofg; p.fpitch = fpitch; p.fstride = fstride; p.obg = obg; p.bpitch = bpitch; p.bstride = bstride;
    p.W = W; p.H = H; p.N = N; p.S = cdiv(W, SWW);
    p.mailbox = mailbox; p.ticket = ticket; p.progress = progress;
    p.tag = ((epoch & 0xfu) << 12) | ((epoch >> 4) << 28);
    p.psleep = env.psleep;
    const int jobs = N * p.S;
    B200MRC_CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned) * 4, st));
    if (progress) B200MRC_CUDA_TRY(cudaMemsetAsync(progress, 0, sizeof(int) * (size_t)jobs, st));
    // Which form (measured on B200, 3300x2550 RGB pages, ms per launch; profiles/r1q_ab_sweep.txt):
    //   pages        4      16     32     64
    //   one warp     2.39   2.44   2.97   3.59      (45 % of the issue slots at 64 pages)
    //   trio         1.78   2.17   2.51   3.60      (62 %: a second batch on another stream overlaps less)
    // The trio form (one warp per layer + a producer thread, shared TMA-fed stages) wins while the machine is not full;
    // from ~6 strips per SM on the one-warp form is as fast and leaves more room for concurrent kernels.
    // B200MRC_IIRW_MODE=trio|single and B200MRC_IIRW_FEED=tma|async (single form) override for A/B runs and tests.
    const bool big = jobs >= 6 * dev_info().sm_count;
    const bool use_trio = env.mode ? env.mode == 2 : !big;
    if (use_trio) {
        const int tpc = env.tpc;
        const size_t smem3 = iirw3_smem_bytes(C, tpc);
        if (smem3 > (size_t)dev_info().max_smem_optin) return B200MRC_ERR_UNSUPPORTED;
        const void *kern = tpc == 1 ? (C == 3 ? (const void *)k_opt_iir_w3<3, IST, 1> : (const void *)k_opt_iir_w3<1, IST, 1>)
                         : tpc == 2 ? (C == 3 ? (const void *)k_opt_iir_w3<3, IST, 2> : (const void *)k_opt_iir_w3<1, IST, 2>)
                                    : (C == 3 ? (const void *)k_opt_iir_w3<3, IST, 3> : (const void *)k_opt_iir_w3<1, IST, 3>);
        B200MRC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
        void *args[] = {(void *)&p};
        { ProfScope _ps("k_opt_iir_w", st); B200MRC_CUDA_TRY(cudaLaunchKernel(kern, dim3((unsigned)cdiv(jobs, tpc)), dim3(32 * (2 * tpc + 1)), args, smem3, st)); }
        count_launch();
        return B200MRC_OK;
    }
    const size_t smem = iirw_smem_bytes(C, wpc);
    if (smem > (size_t)dev_info().max_smem_optin) return B200MRC_ERR_UNSUPPORTED;
    const bool use_tma = env.feed ? env.feed == 1 : big;
    const void *kern = use_tma ? (C == 3 ? (const void *)k_opt_iir_w<3, false> : (const void *)k_opt_iir_w<1, false>)
                               : (C == 3 ? (const void *)k_opt_iir_w<3, true> : (const void *)k_opt_iir_w<1, true>);
    B200MRC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void *args[] = {(void *)&p};
    { ProfScope _ps("k_opt_iir_w", st); B200MRC_CUDA_TRY(cudaLaunchKernel(kern, dim3((unsigned)cdiv(jobs, wpc)), dim3(32 * wpc), args, smem, st)); }
    count_launch();
    return B200MRC_OK;
}

}  // namespace b200mrc
