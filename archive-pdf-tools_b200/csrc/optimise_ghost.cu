// optimise_ghost.cu -- k_opt_iir_g: the row-sequential half of the production optimise path
// (n_fg = 3 / n_bg = 10; internetarchivepdf/mrc.py:412-415, 439-449; semantics
// cython/optimiser.pyx:153-429) as free-running warp strips with ghost lanes.
//
//   out[y,x] = (FIR(y,x) + IIR(y,x)) / den(y,x)      FIR, den: record plane written by k_opt_fir (fmt 1)
//   IIR      = sum of `out` over the n x n box above-left (n = 3 for fg pixels, 10 for bg pixels)
//
// The sweep is bound by the latency of one row step (every row needs the previous one), so the
// design minimises the instructions and round trips on that path and maximises independent warps:
//   * a warp owns a strip of one page and marches down its rows; lane j holds K adjacent columns.
//     The first G lanes are GHOST lanes: they hold the column sums of the 12 (K=4) / 10 (K=2)
//     columns left of the strip, refreshed every row from the left neighbour's mailbox row, so all
//     window sums are plain warp shuffles with no edge cases;
//   * nothing is CTA-wide and nothing needs a fence: every mailbox word carries an 8-bit launch
//     epoch in the spare bits of its two 16-bit lanes (sums < 4096), so a word is either stale
//     (poll again) or complete; mailbox rows are prefetched two rows ahead;
//   * inputs (8-byte records + pixels) are prefetched into registers two rows ahead with plain
//     vector loads (and into L2 eight rows ahead), outputs leave by direct stores: no staging, no
//     barriers of any kind in the row loop;
//   * column sums live in registers in 16-bit lanes (r | b << 16, g); the last n output rows are
//     lane-private smem rings; rows without a mask pixel in the strip take a short path.
// Jobs (page, strip) are ticketed in dependency order, so any residency is deadlock-free.
// floor(num/den) = umulhi(2*num, ceil(2^31/den)), exact for num <= 255*den, den <= 500.
#include "common.cuh"
#include <map>
#include <mutex>

namespace b200mrc {
namespace {

constexpr int NFG = 3, NBG = 10;
constexpr int MAXDEN = 4 * NBG * NBG + NBG * NBG;
constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t TAGMASK = 0xF000F000u;   // mailbox words: bits 12-15 / 28-31 hold the launch epoch
constexpr int MTAB_BYTES = 2048;
constexpr int L2_AHEAD = 10;
constexpr int DEPTH = 4;                 // input rows in flight per lane (cp.async stages, K = 4)

struct IirGParams {
    const uint8_t *img; int64_t ipitch, istride;
    const uint8_t *rec; int64_t rpitch, rstride;
    uint8_t *ofg; int64_t fpitch, fstride;
    uint8_t *obg; int64_t bpitch, bstride;
    int W, H, N, S;
    uint32_t *mailbox;                  // [N][S][H][MBROW]
    unsigned *ticket;
    uint32_t tag;                       // epoch spread over TAGMASK
};

template <int K> struct Geo {
    static constexpr int G = K == 4 ? 3 : 5;            // ghost lanes: G*K >= 10 columns
    static constexpr int RL = 32 - G, RW = RL * K;      // real lanes / real columns per strip
    static constexpr int MBROW = G * 4 * K;             // mailbox words per (strip, row)
    static constexpr int ring_bytes = (NBG + 2 * NFG) * 32 * K * 4;   // bg ring (packed px) + fg ring (lanes form)
    static constexpr int stage_bytes = K == 4 ? 3 * 32 * 16 : 0;      // one input row: records (2 x 16 B) + pixels (<= 16 B) per lane
    static constexpr int warp_bytes = ring_bytes + DEPTH * stage_bytes;
};

__device__ __forceinline__ uint32_t perm(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
__device__ __forceinline__ uint32_t up(uint32_t v, int d) { return __shfl_up_sync(FULL, v, d); }

template <int K> struct VecW;
template <> struct VecW<4> { using type = uint4; };
template <> struct VecW<2> { using type = uint2; };

template <int K> __device__ __forceinline__ void lds(const uint32_t *p, uint32_t *v)
{
    const typename VecW<K>::type t = *reinterpret_cast<const typename VecW<K>::type *>(p);
    const uint32_t *s = reinterpret_cast<const uint32_t *>(&t);
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = s[k];
}
template <int K> __device__ __forceinline__ void sts(uint32_t *p, const uint32_t *v)
{
    typename VecW<K>::type t;
    uint32_t *s = reinterpret_cast<uint32_t *>(&t);
#pragma unroll
    for (int k = 0; k < K; k++) s[k] = v[k];
    *reinterpret_cast<typename VecW<K>::type *>(p) = t;
}
template <int K> __device__ __forceinline__ void ld_relaxed(const uint32_t *p, uint32_t *v)
{
    if (K == 4) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "l"(p) : "memory");
    else        asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "l"(p) : "memory");
}
template <int K> __device__ __forceinline__ void st_relaxed_tag(uint32_t *p, const uint32_t *v, uint32_t tag)
{
    if (K == 4) asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v[0] | tag), "r"(v[1] | tag), "r"(v[2] | tag), "r"(v[3] | tag) : "memory");
    else        asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v[0] | tag), "r"(v[1] | tag) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
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

// (2*nr, 2*ng, 2*nb, 4*den) -> quotients in lanes form
__device__ __forceinline__ void div3(const uint32_t *Mtab, uint32_t nr2, uint32_t ng2, uint32_t nb2, uint32_t den4,
                                     uint32_t &q_rb, uint32_t &q_g)
{
    den4 = min(den4, (uint32_t)(MAXDEN * 4));
    const uint32_t m31 = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(Mtab) + den4);
    q_rb = __umulhi(nr2, m31) | (__umulhi(nb2, m31) << 16);
    q_g = __umulhi(ng2, m31);
}

// One row of inputs for the lane's K pixels: records (lo, hi) and packed pixels r | g << 8 | b << 16
// (the top byte of a pixel word is unspecified).
template <int K> struct RowIn { uint32_t lo[K], hi[K], px[K]; };

template <int C, int K>
__device__ __forceinline__ void load_row(RowIn<K> &r, const uint8_t *ip, const uint8_t *rp, bool full, bool tail, int nvalid)
{
#pragma unroll
    for (int k = 0; k < K; k++) { r.lo[k] = r.hi[k] = r.px[k] = 0; }
    if (full) {
        if (K == 4) {
            const uint4 a = __ldg(reinterpret_cast<const uint4 *>(rp)), b = __ldg(reinterpret_cast<const uint4 *>(rp) + 1);
            r.lo[0] = a.x; r.hi[0] = a.y; r.lo[1] = a.z; r.hi[1] = a.w; r.lo[2] = b.x; r.hi[2] = b.y; r.lo[3] = b.z; r.hi[3] = b.w;
            if (C == 3) {
                const uint32_t w0 = __ldg(reinterpret_cast<const uint32_t *>(ip)), w1 = __ldg(reinterpret_cast<const uint32_t *>(ip) + 1),
                               w2 = __ldg(reinterpret_cast<const uint32_t *>(ip) + 2);
                r.px[0] = w0; r.px[1] = perm(w0, w1, 0x5543); r.px[2] = perm(w1, w2, 0x4432); r.px[3] = w2 >> 8;
            } else {
                const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(ip));
                r.px[0] = perm(w, 0, 0x4000); r.px[1] = perm(w, 0, 0x4111); r.px[2] = perm(w, 0, 0x4222); r.px[3] = perm(w, 0, 0x4333);
            }
        } else {
            const uint4 a = __ldg(reinterpret_cast<const uint4 *>(rp));
            r.lo[0] = a.x; r.hi[0] = a.y; r.lo[1] = a.z; r.hi[1] = a.w;
            if (C == 3) {
                const uint32_t h0 = __ldg(reinterpret_cast<const uint16_t *>(ip)), h1 = __ldg(reinterpret_cast<const uint16_t *>(ip) + 1),
                               h2 = __ldg(reinterpret_cast<const uint16_t *>(ip) + 2);
                r.px[0] = h0 | (h1 << 16); r.px[1] = (h1 >> 8) | (h2 << 8);
            } else {
                const uint32_t h = __ldg(reinterpret_cast<const uint16_t *>(ip));
                r.px[0] = perm(h, 0, 0x4000); r.px[1] = perm(h, 0, 0x4111);
            }
        }
    } else if (tail) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            if (k < nvalid) {
                const uint2 a = __ldg(reinterpret_cast<const uint2 *>(rp) + k);
                r.lo[k] = a.x; r.hi[k] = a.y;
                if (C == 3) r.px[k] = (uint32_t)__ldg(ip + 3 * k) | ((uint32_t)__ldg(ip + 3 * k + 1) << 8) | ((uint32_t)__ldg(ip + 3 * k + 2) << 16);
                else r.px[k] = (uint32_t)__ldg(ip + k) * 0x010101u;
            }
        }
    }
}

// K packed pixels -> the lane's K*C output bytes
template <int C, int K>
__device__ __forceinline__ void store_row(uint8_t *op, const uint32_t *px, bool full, bool tail, int nvalid)
{
    if (full) {
        if (K == 4 && C == 3) {
            uint32_t *d = reinterpret_cast<uint32_t *>(op);
            d[0] = perm(px[0], px[1], 0x4210); d[1] = perm(px[1], px[2], 0x5421); d[2] = perm(px[2], px[3], 0x6542);
        } else if (K == 4) {
            *reinterpret_cast<uint32_t *>(op) = perm(perm(px[0], px[1], 0x0040), perm(px[2], px[3], 0x0040), 0x5410);
        } else if (C == 3) {
            uint16_t *d = reinterpret_cast<uint16_t *>(op);
            d[0] = (uint16_t)px[0]; d[1] = (uint16_t)perm(px[0], px[1], 0x0042); d[2] = (uint16_t)(px[1] >> 8);
        } else {
            *reinterpret_cast<uint16_t *>(op) = (uint16_t)perm(px[0], px[1], 0x0040);
        }
    } else if (tail) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            if (k < nvalid) {
                if (C == 3) { op[3 * k] = (uint8_t)px[k]; op[3 * k + 1] = (uint8_t)(px[k] >> 8); op[3 * k + 2] = (uint8_t)(px[k] >> 16); }
                else op[k] = (uint8_t)px[k];
            }
        }
    }
}

template <int C, int K>
__global__ void __launch_bounds__(256) k_opt_iir_g(const IirGParams p)
{
    using GE = Geo<K>;
    constexpr int G = GE::G, RW = GE::RW, MBROW = GE::MBROW;
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t *Mtab = reinterpret_cast<uint32_t *>(smem);
    uint32_t *ringB = reinterpret_cast<uint32_t *>(smem + MTAB_BYTES + (size_t)wid * GE::warp_bytes);   // [NBG][32*K]
    uint32_t *ringF = ringB + NBG * 32 * K;                                                              // [NFG][2][32*K]
    uint8_t *stage = reinterpret_cast<uint8_t *>(ringB) + GE::ring_bytes;                                // [DEPTH][3][32][16]  (K = 4)

    for (int d = 1 + (int)threadIdx.x; d <= MAXDEN; d += blockDim.x) Mtab[d] = (uint32_t)((0x80000000ull + d - 1) / (unsigned long long)d);
    if (threadIdx.x == 0) Mtab[0] = 0;
    for (int i = lane; i < GE::warp_bytes / 16; i += 32) reinterpret_cast<uint4 *>(ringB)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();                                         // the only CTA-wide barrier

    int job = 0;
    if (lane == 0) job = (int)atomicAdd(p.ticket, 1u);
    job = __shfl_sync(FULL, job, 0);
    if (job >= p.N * p.S) return;
    const int page = job / p.S, strip = job - page * p.S;
    const int W = p.W, H = p.H;
    const int col0 = strip * RW + (lane - G) * K;
    const bool ghost = lane < G;
    const bool real = !ghost && col0 < W;
    const bool full = real && col0 + K <= W, tail = real && !full;
    const int nvalid = real ? min(K, W - col0) : 0;
    // K = 4: whole-group loads (cp.async, DEPTH rows ahead) are legal while the group stays inside the row pitch;
    // a ragged last lane whose group would leave the pitch falls back to guarded synchronous loads
    const bool afull = K == 4 && real && (int64_t)(col0 + K) * C <= p.ipitch && (int64_t)(col0 + K) * 8 <= p.rpitch;
    const bool async_warp = K == 4 && __all_sync(FULL, afull || !real);
    const bool has_left = strip > 0, has_right = strip + 1 < p.S;
    const bool gl = ghost && has_left;                       // ghost lane with a neighbour to listen to
    const bool pub = has_right && lane >= 32 - G;            // lane that feeds the right neighbour's ghost lanes
    const int64_t colc = real ? col0 : 0;
    const uint8_t *ip = p.img + (int64_t)page * p.istride + colc * C;
    const uint8_t *rp = p.rec + (int64_t)page * p.rstride + colc * 8;
    uint8_t *fp = p.ofg + (int64_t)page * p.fstride + colc * C;
    uint8_t *bp = p.obg + (int64_t)page * p.bstride + colc * C;
    // mailbox row r of strip s: the column sums for row r of s's last G lanes, slot g = lane - (32 - G)
    uint32_t *mb_out = p.mailbox + ((int64_t)page * p.S + strip) * (int64_t)H * MBROW + (pub ? (lane - (32 - G)) * 4 * K : 0);
    const uint32_t *mb_in = p.mailbox + ((int64_t)page * p.S + (has_left ? strip - 1 : 0)) * (int64_t)H * MBROW + (ghost ? lane * 4 * K : 0);
    const uint32_t tag = p.tag;

    // column sums of the last n output rows, lanes form (r | b << 16, g)
    uint32_t Cf_rb[K], Cf_g[K], Cb_rb[K], Cb_g[K];
#pragma unroll
    for (int k = 0; k < K; k++) Cf_rb[k] = Cf_g[k] = Cb_rb[k] = Cb_g[k] = 0;

    // inputs: K = 4 -> cp.async stages DEPTH-1 rows ahead (lane-private slots: no barrier); otherwise rows y, y+1 in registers
    RowIn<K> in0, in1;
    auto issue_async = [&](int r) {                          // row r -> stage r % DEPTH, one commit group per call
        if (afull && r < H) {
            uint8_t *s0 = stage + (r % DEPTH) * GE::stage_bytes + lane * 16;
            const uint8_t *rr = rp + (int64_t)r * p.rpitch, *ii = ip + (int64_t)r * p.ipitch;
            cp_async16(s0, rr); cp_async16(s0 + 512, rr + 16);
            if (C == 3) { cp_async4(s0 + 1024, ii); cp_async4(s0 + 1028, ii + 4); cp_async4(s0 + 1032, ii + 8); }
            else cp_async4(s0 + 1024, ii);
        }
        cp_async_commit();
    };
    if (async_warp) {
        for (int r = 0; r < DEPTH - 1; r++) issue_async(r);
    } else {
        load_row<C, K>(in0, ip, rp, full, tail, nvalid);
        if (1 < H) load_row<C, K>(in1, ip + p.ipitch, rp + p.rpitch, full, tail, nvalid); else in1 = in0;
    }
    uint32_t mA[4][K], mB[4][K];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int k = 0; k < K; k++) mA[a][k] = mB[a][k] = 0;
    if (gl && 1 < H) {
#pragma unroll
        for (int a = 0; a < 4; a++) ld_relaxed<K>(mb_in + (int64_t)1 * MBROW + a * K, mA[a]);
    }
    for (int r = 0; r < L2_AHEAD && r < H; r++)
        if (real) { prefetch_l2(ip + (int64_t)r * p.ipitch); prefetch_l2(rp + (int64_t)r * p.rpitch); }
    int rf = 0, rbg = 0;

    for (int y = 0; y < H; y++) {
        // ---- prefetch: inputs of a later row, L2 lines of row y+L2_AHEAD, mailbox row y+2
        RowIn<K> cur;
        if (async_warp) {
            issue_async(y + DEPTH - 1);                      // into the stage consumed in row y-1
            cp_async_wait<DEPTH - 1>();                      // row y has landed (this lane's own copies)
#pragma unroll
            for (int k = 0; k < K; k++) { cur.lo[k] = cur.hi[k] = cur.px[k] = 0; }
            if (K == 4 && afull) {
                const uint8_t *s0 = stage + (y % DEPTH) * GE::stage_bytes + lane * 16;
                const uint4 a = *reinterpret_cast<const uint4 *>(s0), b = *reinterpret_cast<const uint4 *>(s0 + 512);
                const uint4 w = *reinterpret_cast<const uint4 *>(s0 + 1024);
                cur.lo[0] = a.x; cur.hi[0] = a.y; cur.lo[1] = a.z; cur.hi[1] = a.w; cur.lo[2] = b.x; cur.hi[2] = b.y; cur.lo[3] = b.z; cur.hi[3] = b.w;
                if (C == 3) { cur.px[0] = w.x; cur.px[1] = perm(w.x, w.y, 0x5543); cur.px[2] = perm(w.y, w.z, 0x4432); cur.px[3] = w.z >> 8; }
                else { cur.px[0] = perm(w.x, 0, 0x4000); cur.px[1] = perm(w.x, 0, 0x4111); cur.px[2] = perm(w.x, 0, 0x4222); cur.px[3] = perm(w.x, 0, 0x4333); }
                if (tail) {                                  // ragged last lane: columns beyond the page hold row padding
#pragma unroll
                    for (int k = 0; k < K; k++)
                        if (k >= nvalid) { cur.lo[k] = cur.hi[k] = 0; }
                }
            }
        } else {
            cur = in0;
            in0 = in1;
            if (y + 2 < H) load_row<C, K>(in1, ip + (int64_t)(y + 2) * p.ipitch, rp + (int64_t)(y + 2) * p.rpitch, full, tail, nvalid);
        }
        if (real && y + L2_AHEAD < H) { prefetch_l2(ip + (int64_t)(y + L2_AHEAD) * p.ipitch); prefetch_l2(rp + (int64_t)(y + L2_AHEAD) * p.rpitch); }
        if (gl && y + 2 < H) {
#pragma unroll
            for (int a = 0; a < 4; a++) ld_relaxed<K>(mb_in + (int64_t)(y + 2) * MBROW + a * K, mB[a]);
        }

        uint32_t img_rb[K], img_g[K];
#pragma unroll
        for (int k = 0; k < K; k++) { img_rb[k] = perm(cur.px[k], 0, 0x4240); img_g[k] = perm(cur.px[k], 0, 0x4441); }
        uint32_t hor = 0;
#pragma unroll
        for (int k = 0; k < K; k++) hor |= cur.hi[k];
        const bool need_bg = __any_sync(FULL, (int)hor < 0);

        // ---- fg windows: sums of Cf over the 3 columns to the left
        uint32_t sf_rb[K], sf_g[K];
        if (K == 4) {
            const uint32_t e2r = up(Cf_rb[3], 1), e1r = up(Cf_rb[2], 1), e0r = up(Cf_rb[1], 1);
            const uint32_t e2g = up(Cf_g[3], 1), e1g = up(Cf_g[2], 1), e0g = up(Cf_g[1], 1);
            const uint32_t ar = e1r + e2r, br = Cf_rb[0] + Cf_rb[1], ag = e1g + e2g, bgs = Cf_g[0] + Cf_g[1];
            sf_rb[0] = ar + e0r; sf_rb[1] = ar + Cf_rb[0]; sf_rb[2] = br + e2r; sf_rb[3] = br + Cf_rb[2];
            sf_g[0] = ag + e0g;  sf_g[1] = ag + Cf_g[0];   sf_g[2] = bgs + e2g; sf_g[3] = bgs + Cf_g[2];
        } else {
            const uint32_t e2r = up(Cf_rb[1], 1), e1r = up(Cf_rb[0], 1), e0r = up(Cf_rb[1], 2);
            const uint32_t e2g = up(Cf_g[1], 1), e1g = up(Cf_g[0], 1), e0g = up(Cf_g[1], 2);
            sf_rb[0] = e0r + e1r + e2r; sf_rb[1] = e1r + e2r + Cf_rb[0];
            sf_g[0] = e0g + e1g + e2g;  sf_g[1] = e1g + e2g + Cf_g[0];
        }

        uint32_t of_rb[K], of_g[K], ob_rb[K], ob_g[K], pfg[K], pbg[K];
        if (!need_bg) {
            // ---- no mask pixel in this strip row: fg = quotient everywhere, bg = input row
#pragma unroll
            for (int k = 0; k < K; k++) {
                const uint32_t t_rb = sf_rb[k] * 2u + cur.lo[k], t_gd = sf_g[k] * 2u + cur.hi[k];
                div3(Mtab, t_rb & 0xffffu, t_gd & 0xffffu, t_rb >> 16, t_gd >> 16, of_rb[k], of_g[k]);
                ob_rb[k] = img_rb[k]; ob_g[k] = img_g[k]; pbg[k] = cur.px[k];
            }
        } else {
            // ---- bg windows: sums of Cb over the 10 columns to the left
            uint32_t sb_rb[K], sb_g[K];
            if (K == 4) {
                const uint32_t Tr = (Cb_rb[0] + Cb_rb[1]) + (Cb_rb[2] + Cb_rb[3]), Tg = (Cb_g[0] + Cb_g[1]) + (Cb_g[2] + Cb_g[3]);
                const uint32_t ABr = up(Tr, 1) + up(Tr, 2), ABg = up(Tg, 1) + up(Tg, 2);
                const uint32_t U_r = up(Cb_rb[2] + Cb_rb[3], 3), V_r = up(Cb_rb[3], 3), X_r = up(Cb_rb[0], 2);
                const uint32_t U_g = up(Cb_g[2] + Cb_g[3], 3), V_g = up(Cb_g[3], 3), X_g = up(Cb_g[0], 2);
                const uint32_t c01r = Cb_rb[0] + Cb_rb[1], c01g = Cb_g[0] + Cb_g[1];
                sb_rb[0] = ABr + U_r;             sb_g[0] = ABg + U_g;
                sb_rb[1] = ABr + V_r + Cb_rb[0];  sb_g[1] = ABg + V_g + Cb_g[0];
                sb_rb[2] = ABr + c01r;            sb_g[2] = ABg + c01g;
                sb_rb[3] = ABr - X_r + c01r + Cb_rb[2];  sb_g[3] = ABg - X_g + c01g + Cb_g[2];
            } else {
                // columns c-10 .. c-1 are exactly the 5 lanes to the left
                const uint32_t Tr = Cb_rb[0] + Cb_rb[1], Tg = Cb_g[0] + Cb_g[1];
                const uint32_t s2r = Tr + up(Tr, 1), s2g = Tg + up(Tg, 1);
                const uint32_t s4r = s2r + up(s2r, 2), s4g = s2g + up(s2g, 2);
                const uint32_t s5r = up(s4r, 1) + up(Tr, 5), s5g = up(s4g, 1) + up(Tg, 5);
                sb_rb[0] = s5r; sb_g[0] = s5g;
                sb_rb[1] = s5r - up(Cb_rb[0], 5) + Cb_rb[0]; sb_g[1] = s5g - up(Cb_g[0], 5) + Cb_g[0];
            }
#pragma unroll
            for (int k = 0; k < K; k++) {
                const bool m = (int)cur.hi[k] < 0;
                // fg-type record (lanes form, doubled) ...
                const uint32_t t_rb = sf_rb[k] * 2u + cur.lo[k], t_gd = sf_g[k] * 2u + cur.hi[k];
                uint32_t nr2 = t_rb & 0xffffu, nb2 = t_rb >> 16, ng2 = t_gd & 0xffffu, den4 = t_gd >> 16;
                // ... or bg-type record (legacy packing): r[0,17) g[17,34) b[34,51) den[51,60)
                if (m) {
                    const uint32_t Fr = cur.lo[k] & 0x1ffffu, Fg = (cur.lo[k] >> 17) | ((cur.hi[k] & 3u) << 15), Fb = (cur.hi[k] >> 2) & 0x1ffffu;
                    nr2 = (Fr + (sb_rb[k] & 0xffffu)) * 2u; nb2 = (Fb + (sb_rb[k] >> 16)) * 2u; ng2 = (Fg + sb_g[k]) * 2u;
                    den4 = ((cur.hi[k] >> 19) & 0xfffu) * 4u;
                }
                uint32_t q_rb, q_g;
                div3(Mtab, nr2, ng2, nb2, den4, q_rb, q_g);
                of_rb[k] = m ? img_rb[k] : q_rb; of_g[k] = m ? img_g[k] : q_g;
                ob_rb[k] = m ? q_rb : img_rb[k]; ob_g[k] = m ? q_g : img_g[k];
                pbg[k] = perm(ob_rb[k], ob_g[k], 0x7240);
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) pfg[k] = perm(of_rb[k], of_g[k], 0x7240);

        // ---- outputs
        store_row<C, K>(fp + (int64_t)y * p.fpitch, pfg, full, tail, nvalid);
        store_row<C, K>(bp + (int64_t)y * p.bpitch, pbg, full, tail, nvalid);

        // ---- column sums: + out[y], - out[y-n] (lane-private rings; zero-initialised = rows above the page)
        {
            uint32_t o_rb[K], o_g[K], o_px[K];
            lds<K>(ringF + (rf * 2 + 0) * 32 * K + lane * K, o_rb);
            lds<K>(ringF + (rf * 2 + 1) * 32 * K + lane * K, o_g);
            lds<K>(ringB + rbg * 32 * K + lane * K, o_px);
#pragma unroll
            for (int k = 0; k < K; k++) {
                Cf_rb[k] += of_rb[k] - o_rb[k]; Cf_g[k] += of_g[k] - o_g[k];
                Cb_rb[k] += ob_rb[k] - perm(o_px[k], 0, 0x4240); Cb_g[k] += ob_g[k] - perm(o_px[k], 0, 0x4441);
            }
            sts<K>(ringF + (rf * 2 + 0) * 32 * K + lane * K, of_rb);
            sts<K>(ringF + (rf * 2 + 1) * 32 * K + lane * K, of_g);
            sts<K>(ringB + rbg * 32 * K + lane * K, pbg);
        }

        if (y + 1 < H) {
            // ---- hand the new column sums (those of row y+1) to the right neighbour's ghost lanes
            if (pub) {
                uint32_t *dst = mb_out + (int64_t)(y + 1) * MBROW;
                st_relaxed_tag<K>(dst + 0 * K, Cb_rb, tag); st_relaxed_tag<K>(dst + 1 * K, Cb_g, tag);
                st_relaxed_tag<K>(dst + 2 * K, Cf_rb, tag); st_relaxed_tag<K>(dst + 3 * K, Cf_g, tag);
            }
            // ---- ghost lanes: take over the left neighbour's sums for row y+1 once every word carries this launch's tag
            if (has_left) {
                auto tags_ok = [&]() {
                    uint32_t bad = 0;
#pragma unroll
                    for (int a = 0; a < 4; a++)
#pragma unroll
                        for (int k = 0; k < K; k++) bad |= mA[a][k] ^ tag;
                    return (bad & TAGMASK) == 0;
                };
                bool ok = !ghost || tags_ok();
                while (!__all_sync(FULL, ok)) {
                    __nanosleep(32);
                    if (ghost) {
#pragma unroll
                        for (int a = 0; a < 4; a++) ld_relaxed<K>(mb_in + (int64_t)(y + 1) * MBROW + a * K, mA[a]);
                        ok = tags_ok();
                    }
                }
                if (ghost) {
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        Cb_rb[k] = mA[0][k] & ~TAGMASK; Cb_g[k] = mA[1][k] & ~TAGMASK;
                        Cf_rb[k] = mA[2][k] & ~TAGMASK; Cf_g[k] = mA[3][k] & ~TAGMASK;
                    }
                }
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int k = 0; k < K; k++) mA[a][k] = mB[a][k];
            } else if (ghost) {
#pragma unroll
                for (int k = 0; k < K; k++) Cf_rb[k] = Cf_g[k] = Cb_rb[k] = Cb_g[k] = 0;   // left page edge
            }
        }
        if (++rf == NFG) rf = 0;
        if (++rbg == NBG) rbg = 0;
    }
}

// Launch epochs: mailbox words written by launch e carry tag(e); a mailbox region is cleared whenever it is
// first used, used with another geometry, or the 8-bit epoch wraps, so a stale word can never match.
struct MbState { int W, H, N, C, K; };
std::mutex g_mb_mutex;
std::map<const void *, MbState> g_mb_state;
unsigned g_epoch = 0;

template <int K> size_t mailbox_words(int W, int H, int N) { return (size_t)N * cdiv(W, Geo<K>::RW) * (size_t)H * Geo<K>::MBROW; }

}  // namespace

size_t iirg_mailbox_bytes(int W, int H, int N) { return sizeof(uint32_t) * std::max(mailbox_words<4>(W, H, N), mailbox_words<2>(W, H, N)); }
// record plane pitch: rows are read in whole K-pixel groups
int64_t iirg_rec_pitch(int W) { return (int64_t)((W + 3) / 4 * 4) * 8; }

void iirg_forget(const void *mailbox)
{
    std::lock_guard<std::mutex> lk(g_mb_mutex);
    g_mb_state.erase(mailbox);
}

// The record plane `rec` must hold k_opt_fir's fmt-1 records with pitch iirg_rec_pitch(W).  ticket: 1 word.
int launch_opt_iir_ghost(const uint8_t *img, int64_t ipitch, int64_t istride, int C,
                         const uint8_t *rec, int64_t rpitch, int64_t rstride,
                         uint8_t *ofg, int64_t fpitch, int64_t fstride,
                         uint8_t *obg, int64_t bpitch, int64_t bstride,
                         int W, int H, int N, uint32_t *mailbox, unsigned *ticket, int K, int wpc, cudaStream_t st)
{
    if (wpc < 1 || wpc > 8 || (K != 2 && K != 4)) return B200MRC_ERR_UNSUPPORTED;
    const size_t mb_bytes = sizeof(uint32_t) * (K == 4 ? mailbox_words<4>(W, H, N) : mailbox_words<2>(W, H, N));
    unsigned epoch;
    {
        std::lock_guard<std::mutex> lk(g_mb_mutex);
        g_epoch = g_epoch % 255u + 1u;                       // 1..255
        if (g_epoch == 1u) g_mb_state.clear();               // wrapped: every region is cleared before its next use
        auto it = g_mb_state.find(mailbox);
        const bool same = it != g_mb_state.end() && it->second.W == W && it->second.H == H && it->second.N == N &&
                          it->second.C == C && it->second.K == K;
        if (!same) {
            B200MRC_CUDA_TRY(cudaMemsetAsync(mailbox, 0, mb_bytes, st));
            g_mb_state[mailbox] = MbState{W, H, N, C, K};
        }
        epoch = g_epoch;
    }
    IirGParams p;
    p.img = img; p.ipitch = ipitch; p.istride = istride; p.rec = rec; p.rpitch = rpitch; p.rstride = rstride;
    p.ofg = ofg; p.fpitch = fpitch; p.fstride = fstride; p.obg = obg; p.bpitch = bpitch; p.bstride = bstride;
    p.W = W; p.H = H; p.N = N; p.S = cdiv(W, K == 4 ? Geo<4>::RW : Geo<2>::RW);
    p.mailbox = mailbox; p.ticket = ticket;
    p.tag = ((epoch & 0xfu) << 12) | ((epoch >> 4) << 28);
    const int jobs = N * p.S;
    const size_t smem = MTAB_BYTES + (size_t)wpc * (K == 4 ? Geo<4>::warp_bytes : Geo<2>::warp_bytes);
    B200MRC_CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned) * 4, st));
    const void *kern = K == 4 ? (C == 3 ? (const void *)k_opt_iir_g<3, 4> : (const void *)k_opt_iir_g<1, 4>)
                              : (C == 3 ? (const void *)k_opt_iir_g<3, 2> : (const void *)k_opt_iir_g<1, 2>);
    B200MRC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void *args[] = {(void *)&p};
    { ProfScope _ps("k_opt_iir_g", st); B200MRC_CUDA_TRY(cudaLaunchKernel(kern, dim3((unsigned)cdiv(jobs, wpc)), dim3(32 * wpc), args, smem, st)); }
    count_launch();
    return B200MRC_OK;
}

}  // namespace b200mrc
