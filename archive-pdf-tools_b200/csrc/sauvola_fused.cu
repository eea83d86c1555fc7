// sauvola_fused.cu -- k_sauvola_fused: create_threshold_mask (internetarchivepdf/mrc.py:300-329) as ONE kernel:
//     PIL convert('L') (mrc.py:358-363)  ->  [scipy gaussian_filter(sigma = 0.1 sigma_est) if sigma_est > 1 (mrc.py:309-311)]
//     -> .astype(uint8) (mrc.py:325)  ->  binarise_sauvola (cython/sauvola.pyx:29-222) + invert (mrc.py:84)  [-> mask |= (mrc.py:329)]
// Semantics of every step are those of gray_blur.cu / sauvola.cu (oracle: orc_gray, orc_gauss_blur, orc_sauvola); what
// changes is the data flow: the gray / blurred plane is never a separate pass over HBM.
//
// B200 mapping.  A CTA (128..256 threads) marches a (column strip x row band) of one page top to bottom, 8 adjacent
// columns per thread:
//   * the source rows (RGB or gray) arrive by TMA bulk copies (cp.async.bulk, one row segment per copy, NS rows in
//     flight, one mbarrier per stage) issued by one elected thread; threads read their 24 / 8 bytes with 64-bit LDS;
//   * RGB -> L is two dp4a per pixel on the coefficient bytes (exact: 19595 = 76*256 + 139 ...); the blur's vertical pass
//     runs on a (2R+1)-row register window (R <= 2), its float32 results cross threads through a double-buffered smem
//     row, the horizontal pass + uint8 truncation give the final gray row -- FP64, scipy's operation order, no FMA;
//   * that row enters the running column sums (sum, sum of squares) straight from registers and is also written to a
//     gray "delay line" plane, from which the SAME thread re-reads it as the current row (u rows later) and as the row
//     leaving the window (window rows later): L2-resident re-reads, 64-bit loads two rows ahead of use;
//   * horizontal window sums are differences of a CTA-wide prefix of the column sums: thread-local prefix -> warp
//     shuffle scan -> warp totals through smem; the CTA-wide prefix is published AFTER the row's single barrier for the
//     row that follows, so one barrier per row serves the blur exchange, the warp totals and the prefix hand-over;
//   * S/n and Q/n are multiply-high + shift with a per-row magic number (exact, see fast_div_ok), the test runs in FP64
//     (__dmul_rn/__dadd_rn, the reference's operation order); edge columns (clamped windows) take the FP64-reciprocal
//     form of sauvola.cu.
// Pages whose blur radius exceeds 2 (sigma_est >= 22.5) are pre-blurred into the gray plane by gray_blur.cu's tiled
// kernels and thresholded from there (no production step).  Algorithmic HBM bytes: C in + 1 out per pixel.
#include "common.cuh"
#include "blur.cuh"
#include "tma.cuh"
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace b200mrc {

int launch_gray_blur_min_radius(const uint8_t *in, int64_t in_pitch, int64_t in_stride, int C,
                                uint8_t *out, int64_t out_pitch, int64_t out_stride,
                                int W, int H, int N, const double *sigma, int rmin, int *err_flag, cudaStream_t st);

namespace {

constexpr int SK8 = 8;            // columns per thread
constexpr int NS = 4;             // TMA stages (source rows in flight)
constexpr int EB = 128;           // bias of prefix entry indices (window half-width <= 128)
constexpr int MAXR = 2;           // largest blur radius produced in-kernel
constexpr unsigned FULLW = 0xffffffffu;

struct FusedParams {
    const uint8_t *src; int64_t src_pitch, src_stride;          // page image, C = 1 or 3 (template)
    uint8_t *gray; int64_t gray_pitch, gray_stride;             // delay line / pre-blurred plane (may be null when never needed)
    uint8_t *out; int64_t out_pitch, out_stride;
    const double *sigma;                                        // device, per page, may be null
    int W, H;
    int ww, wh, l, r, o, u;
    int n_strips, strip_w, ext_left, n_bands, band_h;
    int sps;                                                    // prefix plane stride (entries per residue row)
    double km1, k2;
    int flags;
    int out8;                                                   // out rows allow 8-byte stores
};

// smem carve-up (bytes), nt = threads per CTA
struct FusedSmem {
    int off_wt, off_my, off_v, off_p, off_stage, total;
    __host__ __device__ FusedSmem(int nt, int C)
    {
        const int ncols = nt * SK8;
        const int sps = (ncols + 2 * EB + 8 + 7) / 8;
        off_wt = 256;                                   // [2][8] uint2 (mbarriers at 0, weights at 64)
        off_my = off_wt + 2 * 8 * 8;                    // [256] uint2
        off_v = off_my + 256 * 8;                       // [2][ncols + 8] float
        off_p = (off_v + 2 * (ncols + 8) * 4 + 15) & ~15;   // [2][8][sps] uint2
        off_stage = (off_p + 2 * 8 * sps * 8 + 127) & ~127; // [NS][ncols * C]
        total = off_stage + NS * ncols * C;
    }
};

__device__ __forceinline__ uint32_t byte_of(uint32_t w, int k) { return __byte_perm(w, 0, 0x4440 + k); }

// a / n == umulhi(a, M) >> sh for every a < 2^31 when M = ceil(2^(32+sh) / n) with 2^sh < n <= 2^(sh+1):
// e = M n - 2^(32+sh) < n <= 2^(sh+1), so a e < 2^(32+sh).  Window sums: S <= 255 n, Q <= 65025 n, hence n < 33025.
__host__ __device__ __forceinline__ bool fast_div_ok(int n) { return n >= 2 && n < 33025; }

__device__ __forceinline__ void scan_up2f(uint32_t &a, uint32_t &b, int d)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 t, u;\n\t"
        "shfl.sync.up.b32 t|p, %0, %2, 0, 0xffffffff;\n\t"
        "shfl.sync.up.b32 u, %1, %2, 0, 0xffffffff;\n\t"
        "@p add.u32 %0, %0, t;\n\t"
        "@p add.u32 %1, %1, u;\n\t}"
        : "+r"(a), "+r"(b) : "r"(d));
}

template <int V> struct IC { static constexpr int value = V; };

// CS = channels of the source plane, R = blur radius produced here, PRODUCE = gray rows are produced (and written to the
// delay line); otherwise the source IS the gray plane and is read directly.
template <int CS, int R, bool PRODUCE>
__device__ __forceinline__ void fused_march(const FusedParams &p, uint8_t *smem, const uint8_t *src, int64_t src_pitch,
                                            uint8_t *gray, uint8_t *out, const double *sw)
{
    static_assert(PRODUCE || (CS == 1 && R == 0), "direct mode reads a gray plane");
    const int nt = blockDim.x, ncols = nt * SK8;
    const FusedSmem L(nt, CS);                       // stage region is sized by the launcher for the kernel's C >= CS
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem);
    uint2 *sWT = reinterpret_cast<uint2 *>(smem + L.off_wt);
    const uint2 *sMY = reinterpret_cast<const uint2 *>(smem + L.off_my);
    float *sV = reinterpret_cast<float *>(smem + L.off_v);
    uint2 *sP = reinterpret_cast<uint2 *>(smem + L.off_p);
    uint8_t *stage = smem + L.off_stage;
    const int sps = p.sps, pbuf = 8 * sps, vbuf = ncols + 8;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int strip = blockIdx.x % p.n_strips, band = blockIdx.x / p.n_strips;
    const int W = p.W, H = p.H;
    const int sx0 = strip * p.strip_w;
    const int ex0 = sx0 - p.ext_left;                // first column of this CTA (multiple of 16, may be negative)
    const int by0 = band * p.band_h, by1 = min(H, by0 + p.band_h);
    const int i0 = tid * SK8, gx = ex0 + i0;         // local / global first column of this thread (multiples of 8)
    const bool lok = gx >= 0 && gx < W;
    const bool is_out = lok && gx >= sx0 && gx < sx0 + p.strip_w;
    // the first and the last thread of the CTA are padding: their blurred values lack a horizontal neighbour, so they never
    // write the delay line (a neighbouring CTA owns those columns) and nothing reads their column sums (launcher geometry)
    const bool wr_ok = lok && tid != 0 && tid != nt - 1;
    uint32_t lmask[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int v = W - (gx + 4 * h);              // valid columns from this word on
        lmask[h] = (!lok || v <= 0) ? 0u : (v >= 4 ? 0xffffffffu : (1u << (8 * v)) - 1u);
    }
    // window widths: interior threads have ww for all 8 pixels
    const bool nx_uniform = gx - p.l + 1 >= 0 && gx + 7 + p.r + 1 <= W;
    const bool blur_edge = R > 0 && (gx - R < 0 || gx + 7 + R >= W);

    // ---- TMA feed of the source rows (one elected thread)
    const int cx0 = max(ex0, 0);
    const uint32_t copy_bytes = (uint32_t)(min((ex0 + ncols) * CS, ((W * CS + 15) & ~15)) - cx0 * CS);
    const int dst_off = (cx0 - ex0) * CS;
    const uint8_t *src_col = src + (int64_t)cx0 * CS;
    const int stage_bytes = ncols * CS;
    const int g0 = max(0, by0 - p.o + 1);            // first gray row this band needs
    const int y_start = g0 - 2 - p.u - 2 * R;        // iteration y consumes source row y + 2 + u + R and completes gray row y + 2 + u
    const int rs_first = g0 - R, rs_last = by1 + 1 + p.u + R;
    auto issue_row = [&](int rs, int s) {
        mbar_expect_tx(&mbar[s], copy_bytes);
        tma_load(stage + s * stage_bytes + dst_off, src_col + (int64_t)reflect_idx(rs, H) * src_pitch, copy_bytes, &mbar[s]);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++)
            if (rs_first + s <= rs_last) issue_row(rs_first + s, s);
    }
    if (tid < 2) sP[tid * pbuf + (EB & 7) * sps + (EB >> 3)] = make_uint2(0u, 0u);      // prefix entry of "no columns"

    double w0 = 0.0, w1 = 0.0, w2 = 0.0;
    if constexpr (R > 0) { w0 = sw[0]; w1 = sw[1]; if constexpr (R > 1) w2 = sw[2]; }

    uint32_t cs[SK8], cq[SK8];
#pragma unroll
    for (int k = 0; k < SK8; k++) { cs[k] = 0; cq[k] = 0; }
    uint32_t win[2 * R + 1][2];
#pragma unroll
    for (int j = 0; j < 2 * R + 1; j++) { win[j][0] = 0; win[j][1] = 0; }

    // delay-line reads (current row / leaving row), two iterations ahead
    const uint8_t *dl = PRODUCE ? gray : src;
    const int64_t dl_pitch = PRODUCE ? p.gray_pitch : src_pitch;
    auto load_row = [&](int row, uint32_t (&d)[2]) {
        d[0] = 0; d[1] = 0;
        if (lok && row >= 0 && row < H) {
            const uint2 v = __ldcg(reinterpret_cast<const uint2 *>(dl + (int64_t)row * dl_pitch + gx));
            d[0] = v.x; d[1] = v.y;
        }
    };
    uint32_t cA[2] = {0, 0}, lA[2] = {0, 0}, cB[2] = {0, 0}, lB[2] = {0, 0};

    const int it_main = (by0 - 1) - y_start;         // first main iteration
    int slot = 0, par = 0;
    int ny_cached = -1;
    uint32_t magicM = 0, magicS = 0;
    uint8_t *orow = out + (int64_t)by0 * p.out_pitch + gx;
    uint8_t *grow = PRODUCE ? gray + (int64_t)g0 * p.gray_pitch + gx : nullptr;   // next delay-line row to write (row rc)

    auto step = [&](auto PARC, const int y, const int it) {
        constexpr int PB = decltype(PARC)::value;
        const bool main = it >= it_main;
        // ================= part 1: source row -> gray words -> vertical pass; scan of the column sums
        mbar_wait(&mbar[slot], (uint32_t)par);
        uint32_t gw[2];
        {
            const uint8_t *sp = stage + slot * stage_bytes + i0 * CS;
            if (CS == 3) {
                const uint2 a = *reinterpret_cast<const uint2 *>(sp), b = *reinterpret_cast<const uint2 *>(sp + 8),
                            c = *reinterpret_cast<const uint2 *>(sp + 16);
                const uint32_t wv[6] = {a.x, a.y, b.x, b.y, c.x, c.y};
                constexpr uint32_t KH0 = 76u | (150u << 8) | (29u << 16), KL0 = 139u | (70u << 8) | (47u << 16);
                constexpr uint32_t KH1 = KH0 << 8, KL1 = KL0 << 8;
                uint32_t tv[8];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t x0 = wv[3 * h], x1 = wv[3 * h + 1], x2 = wv[3 * h + 2];
                    const uint32_t p1 = __byte_perm(x0, x1, 0x0543), p2 = __byte_perm(x1, x2, 0x0432);
                    tv[4 * h + 0] = __dp4a(x0, KH0, 0u) * 256u + __dp4a(x0, KL0, 0x8000u);
                    tv[4 * h + 1] = __dp4a(p1, KH0, 0u) * 256u + __dp4a(p1, KL0, 0x8000u);
                    tv[4 * h + 2] = __dp4a(p2, KH0, 0u) * 256u + __dp4a(p2, KL0, 0x8000u);
                    tv[4 * h + 3] = __dp4a(x2, KH1, 0u) * 256u + __dp4a(x2, KL1, 0x8000u);
                    // L = bits 16..23 of each total
                    gw[h] = __byte_perm(__byte_perm(tv[4 * h], tv[4 * h + 1], 0x0062), __byte_perm(tv[4 * h + 2], tv[4 * h + 3], 0x0062), 0x5410);
                }
            } else {
                const uint2 a = *reinterpret_cast<const uint2 *>(sp);
                gw[0] = a.x; gw[1] = a.y;
            }
        }
        if constexpr (R > 0) {
#pragma unroll
            for (int j = 0; j < 2 * R; j++) { win[j][0] = win[j + 1][0]; win[j][1] = win[j + 1][1]; }
            win[2 * R][0] = gw[0]; win[2 * R][1] = gw[1];
            float v[SK8];
#pragma unroll
            for (int k = 0; k < SK8; k++) {
                const int h = k >> 2, b = k & 3;
                double acc = __dmul_rn(u2d(byte_of(win[R][h], b)), w0);
                if constexpr (R > 1) acc = __dadd_rn(acc, __dmul_rn(u2d(byte_of(win[0][h], b) + byte_of(win[2 * R][h], b)), w2));
                acc = __dadd_rn(acc, __dmul_rn(u2d(byte_of(win[R - 1][h], b) + byte_of(win[R + 1][h], b)), w1));
                v[k] = __double2float_rn(acc);
            }
            float *vd = sV + PB * vbuf + 4 + i0;
            *reinterpret_cast<float4 *>(vd) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4 *>(vd + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
        uint32_t ps[SK8], pq[SK8], bs = 0, bq = 0;
        if (main) {
            ps[0] = cs[0]; pq[0] = cq[0];
#pragma unroll
            for (int k = 1; k < SK8; k++) { ps[k] = ps[k - 1] + cs[k]; pq[k] = pq[k - 1] + cq[k]; }
            uint32_t ws = ps[SK8 - 1], wq = pq[SK8 - 1];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) scan_up2f(ws, wq, d);
            if (lane == 31) sWT[PB * 8 + warp] = make_uint2(ws, wq);
            bs = ws - ps[SK8 - 1]; bq = wq - pq[SK8 - 1];       // exclusive within the warp
        }
        __syncthreads();
        if (tid == 0 && y + 2 + p.u + R + NS <= rs_last) issue_row(y + 2 + p.u + R + NS, slot);   // every thread has read this stage

        // ================= part 2: horizontal pass -> gray row rc; publish the prefix; test row y; slide the window
        const int rc = y + 2 + p.u;
        uint32_t G[2];
        if constexpr (R > 0) {
            const float *vs = sV + PB * vbuf + 4;
            double ext[SK8 + 2 * R];
            if (!blur_edge) {
                const float4 a = *reinterpret_cast<const float4 *>(vs + i0), b = *reinterpret_cast<const float4 *>(vs + i0 + 4);
                ext[R + 0] = (double)a.x; ext[R + 1] = (double)a.y; ext[R + 2] = (double)a.z; ext[R + 3] = (double)a.w;
                ext[R + 4] = (double)b.x; ext[R + 5] = (double)b.y; ext[R + 6] = (double)b.z; ext[R + 7] = (double)b.w;
#pragma unroll
                for (int j = 0; j < R; j++) { ext[j] = (double)vs[i0 - R + j]; ext[R + SK8 + j] = (double)vs[i0 + SK8 + j]; }
            } else {
                // columns outside the page are scipy-reflected (vertical-pass results mirror with their columns)
#pragma unroll
                for (int j = 0; j < SK8 + 2 * R; j++) {
                    int li = reflect_idx(gx - R + j, W) - ex0;
                    li = min(max(li, -4), ncols + 3);
                    ext[j] = (double)vs[li];
                }
            }
            uint32_t ob[SK8];
#pragma unroll
            for (int k = 0; k < SK8; k++) {
                double acc = __dmul_rn(ext[R + k], w0);
                if constexpr (R > 1) acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(ext[R + k - 2], ext[R + k + 2]), w2));
                acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(ext[R + k - 1], ext[R + k + 1]), w1));
                ob[k] = (uint32_t)(int)__double2float_rn(acc) & 0xffu;        // astype(uint8): truncation
            }
            G[0] = ob[0] | (ob[1] << 8) | (ob[2] << 16) | (ob[3] << 24);
            G[1] = ob[4] | (ob[5] << 8) | (ob[6] << 16) | (ob[7] << 24);
        } else {
            G[0] = gw[0]; G[1] = gw[1];
        }
        const bool rc_ok = rc >= g0 && rc < H;
        if (PRODUCE) {
            if (rc >= g0) {
                if (rc_ok && wr_ok) *reinterpret_cast<uint2 *>(grow) = make_uint2(G[0], G[1]);
                grow += p.gray_pitch;
            }
        }
        if (!rc_ok) { G[0] = 0; G[1] = 0; }
        G[0] &= lmask[0]; G[1] &= lmask[1];

        if (main) {
            // ---- CTA-wide prefix of the column sums of row y+1 -> the other buffer (read in the next iteration)
            {
                uint32_t as = bs, aq = bq;
                const uint2 *T = sWT + PB * 8;
                for (int w2i = 0; w2i < warp; w2i++) { const uint2 t = T[w2i]; as += t.x; aq += t.y; }
                uint2 *P = sP + (PB ^ 1) * pbuf + (EB >> 3) + tid;
#pragma unroll
                for (int k = 0; k < SK8; k++) P[((k + 1) & 7) * sps + ((k + 1) >> 3)] = make_uint2(as + ps[k], aq + pq[k]);
            }
            // ---- test of row y
            if (y >= by0) {
                if (is_out) {
                    const uint2 *P = sP + PB * pbuf + tid;
                    const int ny = min(H, y + p.u + 1) - max(0, y - p.o + 1);
                    if (ny != ny_cached) { const uint2 m = sMY[ny]; magicM = m.x; magicS = m.y; ny_cached = ny; }
                    uint32_t S[SK8], Q[SK8];
#pragma unroll
                    for (int k = 0; k < SK8; k++) {
                        const int qh = EB + p.r + 1 + k, ql = EB - p.l + 1 + k;
                        const uint2 hi = P[(qh & 7) * sps + (qh >> 3)];
                        const uint2 lo = P[(ql & 7) * sps + (ql >> 3)];
                        S[k] = hi.x - lo.x; Q[k] = hi.y - lo.y;
                    }
                    uint32_t bits[2] = {0, 0};
                    if (nx_uniform && magicM != 0u) {
#pragma unroll
                        for (int k = 0; k < SK8; k++) {
                            const uint32_t m = __umulhi(S[k], magicM) >> magicS, q = __umulhi(Q[k], magicM) >> magicS;
                            const double md = u2d(m), vd = u2d(q - m * m), pd = u2d(byte_of(cA[k >> 2], k & 3));
                            const double mm = __dmul_rn(md, md);
                            const double t = __dadd_rn(pd, __dmul_rn(md, p.km1));
                            const double rhs = __dmul_rn(__dmul_rn(mm, p.k2), vd);
                            const double lhs = __dmul_rn(t, t);
                            const uint32_t fg = ((t <= 0.0) || (lhs <= rhs)) ? 1u : 0u;
                            bits[k >> 2] |= fg << (8 * (k & 3));
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < SK8; k++) {
                            const int x = gx + k;
                            const int nx = (x < W) ? min(W, x + p.r + 1) - max(0, x - p.l + 1) : 0;
                            const int n = nx * ny;
                            if (n > 0) {
                                // floor((a + 0.5) * (1/n)) == a / n exactly (sauvola.cu)
                                const double rn = 1.0 / (double)n;
                                const double md = floor(__dmul_rn(__dadd_rn((double)S[k], 0.5), rn));
                                const double qd = floor(__dmul_rn(__dadd_rn((double)Q[k], 0.5), rn));
                                const double mm = __dmul_rn(md, md);
                                const double vd = __dadd_rn(qd, -mm);
                                const double pd = (double)byte_of(cA[k >> 2], k & 3);
                                const double t = __dadd_rn(pd, __dmul_rn(md, p.km1));
                                const double rhs = __dmul_rn(__dmul_rn(mm, p.k2), vd);
                                const double lhs = __dmul_rn(t, t);
                                const uint32_t fg = ((t <= 0.0) || (lhs <= rhs)) ? 1u : 0u;
                                bits[k >> 2] |= fg << (8 * (k & 3));
                            }
                        }
                    }
                    if (p.flags & B200MRC_SAUVOLA_RAW_INVERTED) { bits[0] ^= 0x01010101u; bits[1] ^= 0x01010101u; }
                    if (gx + 7 < W) {
                        if (p.out8) {
                            uint2 *o64 = reinterpret_cast<uint2 *>(orow);
                            if (p.flags & B200MRC_SAUVOLA_OR_INTO) { const uint2 old = *o64; bits[0] |= old.x; bits[1] |= old.y; }
                            *o64 = make_uint2(bits[0], bits[1]);
                        } else {
                            uint32_t *o32 = reinterpret_cast<uint32_t *>(orow);
                            if (p.flags & B200MRC_SAUVOLA_OR_INTO) { bits[0] |= o32[0]; bits[1] |= o32[1]; }
                            o32[0] = bits[0]; o32[1] = bits[1];
                        }
                    } else {
                        for (int k = 0; k < SK8 && gx + k < W; k++) {
                            uint8_t b = (uint8_t)((bits[k >> 2] >> (8 * (k & 3))) & 0xffu);
                            if (p.flags & B200MRC_SAUVOLA_OR_INTO) b |= orow[k];
                            orow[k] = b;
                        }
                    }
                }
                orow += p.out_pitch;
            }
            // ---- slide: row y+2's window = row y+1's + gray row rc - gray row y+2-o
            {
                const uint32_t l0 = lA[0] & lmask[0], l1 = lA[1] & lmask[1];
#pragma unroll
                for (int k = 0; k < SK8; k++) {
                    const uint32_t a = byte_of(G[k >> 2], k & 3), b = byte_of(k < 4 ? l0 : l1, k & 3);
                    cs[k] += a - b;
                    cq[k] += a * a - b * b;
                }
            }
            cA[0] = cB[0]; cA[1] = cB[1]; lA[0] = lB[0]; lA[1] = lB[1];
            if (y + 2 < by1) { load_row(y + 2, cB); load_row(y + 4 - p.o, lB); }
        } else {
            // warm-up: gray row rc joins the window of the band's first row
#pragma unroll
            for (int k = 0; k < SK8; k++) {
                const uint32_t a = byte_of(G[k >> 2], k & 3);
                cs[k] += a; cq[k] += a * a;
            }
            if (it + 1 == it_main) {
                // the delay line now holds every row the first two main iterations read back
                load_row(by0 + 1 - p.o, lA);
                load_row(by0, cB); load_row(by0 + 2 - p.o, lB);
            }
        }
        if (++slot == NS) { slot = 0; par ^= 1; }
    };

    __syncthreads();                                   // mbarrier init + prefix zero entry visible
    const int n_it = by1 - y_start;                    // iterations y_start .. by1-1
    if (it_main == 0) {                                // no warm-up iteration (cannot happen for window >= 3; kept for safety)
        load_row(by0 + 1 - p.o, lA); load_row(by0, cB); load_row(by0 + 2 - p.o, lB);
    }
    int it = 0;
    for (; it + 1 < n_it; it += 2) { step(IC<0>(), y_start + it, it); step(IC<1>(), y_start + it + 1, it + 1); }
    if (it < n_it) step(IC<0>(), y_start + it, it);
}

template <int C>
__global__ void __launch_bounds__(256, 2) k_sauvola_fused(const FusedParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem);
    double *sw = reinterpret_cast<double *>(smem + 64);          // 3 weights
    double *sphi = sw + 4;                                       // 5 scratch
    const int page = blockIdx.y;
    const int tid = threadIdx.x;
    const FusedSmem L(blockDim.x, C);

    double sigma;
    const int radius = blur_radius_of(p.sigma, page, sigma);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++) mbar_init(&mbar[s], 1);
        fence_mbar_init();
    }
    // per-row magic numbers: n = ww * ny  ->  (ceil(2^(32+sh) / n), sh); 0 = take the FP64 form
    {
        uint2 *my = reinterpret_cast<uint2 *>(smem + L.off_my);
        for (int ny = tid; ny < 256; ny += blockDim.x) {
            const int n = p.ww * ny;
            uint2 e = make_uint2(0u, 0u);
            if (fast_div_ok(n)) {
                const int sh = 31 - __clz(n - 1);                // 2^sh < n <= 2^(sh+1)
                e.x = (uint32_t)(((1ull << (32 + sh)) + (unsigned long long)n - 1ull) / (unsigned long long)n);
                e.y = (uint32_t)sh;
            }
            my[ny] = e;
        }
    }
    if (radius >= 1 && radius <= MAXR) blur_weights_cta(radius, sigma, sw, sphi);

    const uint8_t *src = p.src + (int64_t)page * p.src_stride;
    uint8_t *gray = p.gray ? p.gray + (int64_t)page * p.gray_stride : nullptr;
    uint8_t *out = p.out + (int64_t)page * p.out_stride;
    if (radius > MAXR) {
        fused_march<1, 0, false>(p, smem, gray, p.gray_pitch, nullptr, out, sw);      // pre-blurred by gray_blur.cu
    } else if (C == 3) {
        if (radius == 0) fused_march<3, 0, true>(p, smem, src, p.src_pitch, gray, out, sw);
        else if (radius == 1) fused_march<3, 1, true>(p, smem, src, p.src_pitch, gray, out, sw);
        else fused_march<3, 2, true>(p, smem, src, p.src_pitch, gray, out, sw);
    } else {
        if (radius == 0) fused_march<1, 0, false>(p, smem, src, p.src_pitch, nullptr, out, sw);
        else if (radius == 1) fused_march<1, 1, true>(p, smem, src, p.src_pitch, gray, out, sw);
        else fused_march<1, 2, true>(p, smem, src, p.src_pitch, gray, out, sw);
    }
}

}  // namespace

// Does the fused kernel apply?  16-byte aligned rows (TMA), a window of 3..255, no k < 0 / inverted-input variants.
bool sauvola_fused_ok(const uint8_t *src, int64_t src_pitch, int64_t src_stride, int C, const uint8_t *out, int64_t out_pitch,
                      int64_t out_stride, int W, int H, int ww, int wh, double k, int flags)
{
    if (k < 0 || (flags & B200MRC_SAUVOLA_INVERT_INPUT)) return false;
    if (ww < 3 || wh < 3 || ww > B200MRC_MAX_WINDOW || wh > B200MRC_MAX_WINDOW) return false;
    if (W < 64 || H < 8) return false;
    if (((uintptr_t)src & 15) || (src_pitch & 15) || (src_stride & 15) || src_pitch < (((int64_t)W * C + 15) & ~15ll)) return false;
    if (((uintptr_t)out & 3) || (out_pitch & 3) || (out_stride & 3) || out_pitch < W) return false;
    return true;
}

// gray: delay-line plane (pitch % 16 == 0, >= round_up(W, 16)); required when C == 3 or sigma != null.
int launch_sauvola_fused(const uint8_t *src, int64_t src_pitch, int64_t src_stride, int C,
                         uint8_t *gray, int64_t gray_pitch, int64_t gray_stride,
                         uint8_t *out, int64_t out_pitch, int64_t out_stride,
                         int W, int H, int N, int ww, int wh, double k, double Rr, const double *sigma, int flags, cudaStream_t st)
{
    if (!sauvola_fused_ok(src, src_pitch, src_stride, C, out, out_pitch, out_stride, W, H, ww, wh, k, flags)) return B200MRC_ERR_UNSUPPORTED;
    const bool need_gray = C == 3 || sigma != nullptr;
    if (need_gray && (!gray || (gray_pitch & 15) || ((uintptr_t)gray & 15) || (gray_stride & 15) || gray_pitch < (((int64_t)W + 15) & ~15ll)))
        return B200MRC_ERR_INVALID;
    if (N > 65535) return B200MRC_ERR_UNSUPPORTED;
    if (sigma) {
        // pages whose blur radius exceeds MAXR are pre-blurred by the tiled kernels (they skip every other page)
        int rc = launch_gray_blur_min_radius(src, src_pitch, src_stride, C, gray, gray_pitch, gray_stride, W, H, N, sigma, MAXR + 1, nullptr, st);
        if (rc != B200MRC_OK) return rc;
    }
    FusedParams p;
    memset(&p, 0, sizeof(p));
    p.src = src; p.src_pitch = src_pitch; p.src_stride = src_stride;
    p.gray = need_gray ? gray : nullptr; p.gray_pitch = gray_pitch; p.gray_stride = gray_stride;
    p.out = out; p.out_pitch = out_pitch; p.out_stride = out_stride;
    p.sigma = sigma; p.W = W; p.H = H;
    p.ww = ww; p.wh = wh;
    p.l = (ww + 1) / 2; p.r = ww / 2; p.o = (wh + 1) / 2; p.u = wh / 2;
    p.ext_left = (p.l - 1 + SK8 + 15) / 16 * 16;                 // >= 8 padding columns on either side (see wr_ok)
    const int right = p.r + SK8;
    // threads per CTA: fewest thread-columns per row of the page
    struct { int nt, bands; } env = {tune(T_FUSED_NT), tune(T_FUSED_BANDS)};
    int nt = 0; long best = 0;
    for (int cand : {128, 192, 256}) {
        if (env.nt && cand != env.nt) continue;
        const int sw_max = (cand * SK8 - p.ext_left - right) / 16 * 16;
        if (sw_max < 16) continue;
        const long cost = (long)cdiv(W, sw_max) * cand;
        if (!nt || cost < best) { nt = cand; best = cost; }
    }
    if (!nt) return B200MRC_ERR_UNSUPPORTED;
    const int ncols = nt * SK8;
    const int sw_max = (ncols - p.ext_left - right) / 16 * 16;
    p.n_strips = cdiv(W, sw_max);
    p.strip_w = (cdiv(W, p.n_strips) + 15) / 16 * 16;
    p.sps = (ncols + 2 * EB + 8 + 7) / 8;
    p.km1 = k - 1.0;
    p.k2 = k * k / Rr / Rr;                                       // sauvola.pyx:60
    p.flags = flags;
    p.out8 = !(((uintptr_t)out & 7) || (out_pitch & 7) || (out_stride & 7));

    const FusedSmem L(nt, C);
    const void *kern = C == 3 ? (const void *)k_sauvola_fused<3> : (const void *)k_sauvola_fused<1>;
    static std::mutex mu;
    static int occ_cache[2][3] = {{0, 0, 0}, {0, 0, 0}};
    int occ;
    {
        std::lock_guard<std::mutex> lk(mu);
        int &oc = occ_cache[C == 3][nt == 128 ? 0 : (nt == 192 ? 1 : 2)];
        if (!oc) {
            B200MRC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FusedSmem(256, C).total));
            B200MRC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oc, kern, nt, L.total));
            if (oc < 1) oc = 1;
        }
        occ = oc;
    }
    // bands: warm-up costs about half a row per window row; pick the band count with the least modelled time
    const long slots = (long)dev_info().sm_count * occ;
    const long items0 = (long)N * p.n_strips;
    int bands = 1; double best_t = 0;
    for (int b = 1; b <= cdiv(H, 32); b++) {
        const int bh = cdiv(H, b);
        if (cdiv(H, bh) != b) continue;
        const long waves = (items0 * b + slots - 1) / slots;
        const double t = (double)waves * (bh + 0.55 * wh + 12);
        if (b == 1 || t < best_t) { best_t = t; bands = b; }
    }
    if (env.bands > 0) bands = env.bands;
    p.band_h = cdiv(H, bands);
    p.n_bands = cdiv(H, p.band_h);

    dim3 grid((unsigned)(p.n_strips * p.n_bands), (unsigned)N);
    void *args[] = {(void *)&p};
    { ProfScope _ps("k_sauvola_fused", st); B200MRC_CUDA_TRY(cudaLaunchKernel(kern, grid, dim3(nt), args, L.total, st)); }
    count_launch();
    return B200MRC_OK;
}

}  // namespace b200mrc

using namespace b200mrc;

namespace b200mrc {
int launch_gray_blur(const uint8_t *in, int64_t in_pitch, int64_t in_stride, int C,
                     uint8_t *out, int64_t out_pitch, int64_t out_stride,
                     int W, int H, int N, const double *sigma, int *err_flag, cudaStream_t st);
// tuning key THRESHOLD_PATH = legacy keeps the two-pass form (A/B runs)
bool threshold_path_legacy() { return tune(T_THRESHOLD_PATH) == 1; }
}  // namespace b200mrc

extern "C" size_t b200mrc_threshold_workspace_bytes(int width, int height, int n_pages)
{
    if (width <= 0 || height <= 0 || n_pages <= 0) return 0;
    return align_up((size_t)width, 16) * (size_t)height * (size_t)n_pages + 256;
}

extern "C" int b200mrc_threshold_mask(const uint8_t *in, int64_t in_pitch, int64_t in_page_stride, int channels,
                                      uint8_t *out, int64_t out_pitch, int64_t out_page_stride,
                                      int width, int height, int n_pages,
                                      int window_width, int window_height, double k, double R,
                                      const double *sigma, int flags,
                                      void *workspace, size_t workspace_bytes, void *stream)
{
    if (!in || !out || width <= 0 || height <= 0 || n_pages <= 0 || !(R > 0)) return B200MRC_ERR_INVALID;
    if (channels != 1 && channels != 3) return B200MRC_ERR_UNSUPPORTED;
    if (window_width < 1 || window_height < 1 || window_width > B200MRC_MAX_WINDOW || window_height > B200MRC_MAX_WINDOW)
        return B200MRC_ERR_UNSUPPORTED;
    if (n_pages > 65535) return B200MRC_ERR_UNSUPPORTED;
    const bool need_gray = channels == 3 || sigma != nullptr;
    const int64_t gpitch = (int64_t)align_up((size_t)width, 16), gstride = gpitch * height;
    uint8_t *gray = reinterpret_cast<uint8_t *>(align_up((size_t)(uintptr_t)workspace, 16));      // TMA rows: 16-byte aligned
    if (need_gray) {
        if (!workspace) return B200MRC_ERR_WORKSPACE;
        const size_t slack = (size_t)((uintptr_t)gray - (uintptr_t)workspace);
        if (workspace_bytes < slack + (size_t)gstride * n_pages) return B200MRC_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (!threshold_path_legacy() &&
        sauvola_fused_ok(in, in_pitch, in_page_stride, channels, out, out_pitch, out_page_stride, width, height,
                         window_width, window_height, k, flags))
        return launch_sauvola_fused(in, in_pitch, in_page_stride, channels, need_gray ? gray : nullptr, gpitch, gstride,
                                    out, out_pitch, out_page_stride, width, height, n_pages, window_width, window_height, k, R,
                                    sigma, flags, st);
    const uint8_t *g = in; int64_t gp = in_pitch, gs = in_page_stride;
    if (need_gray) {
        int rc = launch_gray_blur(in, in_pitch, in_page_stride, channels, gray, gpitch, gstride, width, height, n_pages, sigma, nullptr, st);
        if (rc != B200MRC_OK) return rc;
        g = gray; gp = gpitch; gs = gstride;
    }
    return b200mrc_sauvola(g, gp, gs, out, out_pitch, out_page_stride, width, height, n_pages, window_width, window_height, k, R, flags, stream);
}
