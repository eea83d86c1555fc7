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
//     runs on a (2R+1)-row register window (R <= 2) over the thread's 8 columns plus R neighbour columns on either side
//     (warp shuffles; the edge lanes of a warp read them from the stage), so the horizontal pass + uint8 truncation
//     need no exchange through shared memory -- FP64, scipy's operation order, no FMA;
//   * that row enters the running column sums (sum, sum of squares) straight from registers and is also written to a
//     gray "delay line" plane, from which the SAME thread re-reads it as the current row (u rows later) and as the row
//     leaving the window (window rows later): L2-resident re-reads, 64-bit loads issued an iteration ahead of use;
//   * horizontal window sums are differences of a CTA-wide prefix of the column sums: thread-local prefix -> warp
//     shuffle scan -> warp totals through smem -> prefix published to a ring of three smem buffers.  The one CTA-wide
//     synchronisation per row is a SPLIT mbarrier: a warp arrives right after publishing its total and waits only
//     before it needs the others', with gray production and the test of the previous row in between;
//   * S/n and Q/n are multiply-high + shift with a per-row magic number (exact, see fast_div_ok; columns whose window
//     the page clamps read per-pixel magics from a row-invariant smem table), and the decision is ONE 16-bit table
//     lookup + integer compare: the reference's FP64 test is a monotone function of the integer variance for given
//     (mean, pixel), tabulated exactly once per (k, R) by k_sauvola_vmin.  Page corners (windows clamped both ways) and
//     windows too large for the 32-bit magic run the FP64 sequence itself (__dmul_rn/__dadd_rn, reference order).
// Pages whose blur radius exceeds 2 (sigma_est >= 22.5) are pre-blurred into the gray plane by gray_blur.cu's tiled
// kernels and thresholded from there (no production step).  Algorithmic HBM bytes: C in + 1 out per pixel.
#include "common.cuh"
#include <map>
#include <mutex>
#include "blur.cuh"
#include "tma.cuh"
#include <cstdlib>
#include <cstring>

namespace b200mrc {

int launch_gray_blur_min_radius(const uint8_t *in, int64_t in_pitch, int64_t in_stride, int C,
                                uint8_t *out, int64_t out_pitch, int64_t out_stride,
                                int W, int H, int N, const double *sigma, int rmin, int *err_flag, cudaStream_t st);

namespace {

constexpr int SK8 = 8;            // columns per thread
constexpr int NS = 4;             // TMA stages (source rows in flight)
constexpr int MAXR = 2;           // largest blur radius produced in-kernel
constexpr int ETAB_SLOTS = 40;    // clamped-window output threads per CTA (2 page edges x ceil(128 / 8) columns, with slack)
constexpr int STAGE_PAD = 16;     // bytes in front of / behind a staged row (neighbour reads of the CTA's edge lanes)
constexpr unsigned FULLM = 0xffffffffu;

struct FusedParams {
    const uint8_t *src; int64_t src_pitch, src_stride;          // page image, C = 1 or 3 (template)
    uint8_t *gray; int64_t gray_pitch, gray_stride;             // delay line / pre-blurred plane (may be null when never needed)
    uint8_t *out; int64_t out_pitch, out_stride;
    const double *sigma;                                        // device, per page, may be null
    int W, H;
    int ww, wh, l, r, o, u;
    int n_strips, strip_w, ext_left, n_bands, band_h;
    int sps, eb;                                                // prefix plane stride (entries per residue row); bias of the entry indices
    double km1, k2;
    const uint16_t *vmin;               // decision table (below): pixel p with window mean m is foreground iff variance >= vmin[m << 8 | p]
    int flags;
    int out8;                                                   // out rows allow 8-byte stores
    int dbg;                                                    // 1: skip the per-pixel test (timing experiments, -DB200MRC_EXPERIMENTS builds only; always 0 otherwise)
    // byte offsets, inside one prefix buffer and relative to the thread's slot, of the entries a thread publishes
    // (st) and of the two window-edge entries of each of its 8 pixels (hi, lo): uniform values, so the shared-memory
    // accesses take them from the uniform register file
    int bo_st[SK8], bo_hi[SK8], bo_lo[SK8];
};

// smem carve-up (bytes), nt = threads per CTA
struct FusedSmem {
    int off_rowbar, off_wt, off_my, off_mx, off_etab, off_p, off_stage, stage_stride, pbuf_bytes, total;
    __host__ __device__ FusedSmem(int nt, int C, int eb)
    {
        const int ncols = nt * SK8;
        const int sps = (ncols + 2 * eb + 8 + 7) / 8;
        off_rowbar = 32;                                // TMA mbarriers at 0 (NS x 8), weights at 64
        off_wt = 256;                                   // [2][8] uint2 warp totals
        off_my = off_wt + 2 * 8 * 8;                    // [256] uint2: magic of n = ww * ny
        off_mx = off_my + 256 * 8;                      // [256] uint2: magic of n = nx * wh
        off_etab = off_mx + 256 * 8;                    // [ETAB_SLOTS][8] uint2 + counter: per-pixel magics of clamped-window threads
        off_p = off_etab + ETAB_SLOTS * 64 + 16;        // [3][8][sps] uint2 prefix of the column sums
        pbuf_bytes = 8 * sps * 8;
        off_stage = (off_p + 3 * pbuf_bytes + 127) & ~127;      // [NS][stage_stride]
        stage_stride = (ncols * C + 2 * STAGE_PAD + 15) & ~15;
        total = off_stage + NS * stage_stride;
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

__device__ __forceinline__ void mbar_arrive_cta(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// CS = channels of the source plane, R = blur radius produced here, PRODUCE = gray rows are produced (and written to the
// delay line); otherwise the source IS the gray plane and is read directly.
//
// One iteration y of the row loop (every thread, 8 columns):
//   B  scan of the column sums of row y+1 (thread total -> warp shuffle scan -> warp total to smem); ARRIVE on the row barrier
//   A  source row y+2+u+R from its TMA stage -> gray -> (R > 0) blur: the vertical pass runs on a (2R+1)-row register window
//      for the thread's 8 columns plus R neighbour columns per side (their gray bytes come from the adjacent lanes by
//      shuffle; a warp's edge lanes convert them from the stage), so the horizontal pass needs nobody else -> gray row
//      rc = y+2+u, written to the delay line
//   D  test of row y-1 against the prefix published two iterations ago (sP is a ring of three buffers)
//   -- WAIT on the row barrier: everything between ARRIVE and WAIT is independent of the other warps, so a warp only
//      stalls here if it runs most of an iteration ahead of the slowest one
//   C  CTA-wide prefix of row y+1 (base = totals of the lower warps) -> third buffer
//   E  slide the window: + gray row rc, - gray row y+2-o (delay line)
template <int CS, int R, bool PRODUCE>
__device__ __forceinline__ void fused_march(const FusedParams &p, uint8_t *smem, const uint8_t *src, int64_t src_pitch,
                                            uint8_t *gray, uint8_t *out, const double *sw, const int item)
{
    static_assert(PRODUCE || (CS == 1 && R == 0), "direct mode reads a gray plane");
    constexpr int NE = SK8 + 2 * R;                  // columns of the vertical pass
    const int nt = blockDim.x, ncols = nt * SK8;
    const FusedSmem L(nt, CS, p.eb);                 // offsets up to the stage region do not depend on the channel count
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem);
    uint64_t *rowbar = reinterpret_cast<uint64_t *>(smem + L.off_rowbar);
    uint2 *sWT = reinterpret_cast<uint2 *>(smem + L.off_wt);
    const uint2 *sMY = reinterpret_cast<const uint2 *>(smem + L.off_my);
    const uint16_t *vmin = p.vmin;
    const uint2 *sMX = reinterpret_cast<const uint2 *>(smem + L.off_mx);
    uint8_t *stage = smem + L.off_stage + STAGE_PAD;
    const int stage_stride = (ncols * CS + 2 * STAGE_PAD + 15) & ~15;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int strip = item % p.n_strips, band = item / p.n_strips;
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
    const bool nx_uniform = gx - p.l + 1 >= 0 && gx + 7 + p.r + 1 <= W;     // window width ww for all 8 pixels
    // clamped-window threads keep the (magic, shift) of their 8 pixels' areas nx * wh in a small smem table
    bool edge_ok = false;
    const uint8_t *edge_tab = nullptr;
    const bool edge_l = R > 0 && gx == 0;                                     // blur taps left of the page: scipy 'reflect'
    const bool edge_r = R > 0 && lok && gx + 7 + R >= W;                      // blur taps right of the page

    // ---- TMA feed of the source rows (one elected thread)
    const int cx0 = max(ex0, 0);
    const uint32_t copy_bytes = (uint32_t)(min((ex0 + ncols) * CS, ((W * CS + 15) & ~15)) - cx0 * CS);
    const int dst_off = (cx0 - ex0) * CS;
    const uint8_t *src_col = src + (int64_t)cx0 * CS;
    const int g0 = max(0, by0 - p.o + 1);            // first gray row this band needs
    const int y_start = g0 - 2 - p.u - 2 * R;        // iteration y consumes source row y + 2 + u + R and completes gray row y + 2 + u
    const int n_it = by1 - y_start + 1;              // iterations y_start .. by1 (the test runs one row behind)
    const int rs_first = g0 - R, rs_last = rs_first + n_it - 1;
    auto issue_row = [&](int rs, int s) {
        mbar_expect_tx(&mbar[s], copy_bytes);
        const int rr = rs < 0 ? -1 - rs : (rs >= H ? 2 * H - 1 - rs : rs);      // scipy 'reflect'; |overshoot| <= window/2 + 3 < H
        tma_load(stage + s * stage_stride + dst_off, src_col + (int64_t)min(max(rr, 0), H - 1) * src_pitch, copy_bytes, &mbar[s]);
    };
    if (tid == 0) {
        mbar_init(rowbar, nt >> 5);
        fence_mbar_init();
#pragma unroll
        for (int s = 0; s < NS; s++)
            if (rs_first + s <= rs_last) issue_row(rs_first + s, s);
    }
    uint8_t *pb = smem + L.off_p + tid * 8;          // this thread's slot in prefix buffer 0
    if (tid < 3) *reinterpret_cast<uint2 *>(smem + L.off_p + tid * L.pbuf_bytes + ((p.eb & 7) * p.sps + (p.eb >> 3)) * 8) = make_uint2(0u, 0u);   // prefix of "no columns"

    double w0 = 0.0, w1 = 0.0, w2 = 0.0;
    if constexpr (R > 0) { w0 = sw[0]; w1 = sw[1]; if constexpr (R > 1) w2 = sw[2]; }

    uint32_t cs[SK8], cq[SK8];
#pragma unroll
    for (int k = 0; k < SK8; k++) { cs[k] = 0; cq[k] = 0; }
    uint32_t win[2 * R + 1][3];                      // per row: neighbour bytes (left | right), own 8 gray bytes
#pragma unroll
    for (int j = 0; j < 2 * R + 1; j++) { win[j][0] = 0; win[j][1] = 0; win[j][2] = 0; }

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
    // cA / lA: rows the current iteration consumes (pixels of the tested row y-1, row y+2-o leaving the window);
    // cN / lN: the same for the next iteration, loaded at the TOP of this one so that a whole iteration hides the latency
    uint32_t cA[2] = {0, 0}, lA[2] = {0, 0}, cN[2] = {0, 0}, lN[2] = {0, 0};

    const int it_main = (by0 - 1) - y_start;         // first main iteration (>= 2)
    int slot = 0, par = 0;
    int o_test = 0, o_mid = L.pbuf_bytes, o_pub = 2 * L.pbuf_bytes;      // prefix ring: L(y-1), L(y), L(y+1)
    int ny_cached = -1;
    uint32_t magicM = 0, magicS = 0;
    uint8_t *orow = out + (int64_t)by0 * p.out_pitch + gx;                // next mask row to write
    uint8_t *grow = PRODUCE ? gray + (int64_t)g0 * p.gray_pitch + gx : nullptr;   // next delay-line row to write (row rc)

    {
        // one 64-byte slot per clamped-window output thread (at most 2 * ceil(128 / 8) + 2 of them per CTA)
        uint2 *etab = reinterpret_cast<uint2 *>(smem + L.off_etab);
        const bool need = is_out && !nx_uniform;
        const unsigned bal = __ballot_sync(FULLM, need);
        int *ecount = reinterpret_cast<int *>(smem + L.off_etab + ETAB_SLOTS * 64);
        if (tid == 0) *ecount = 0;
        __syncthreads();
        int slot_e = -1;
        if (need) {
            int basee = 0;
            if ((bal & ((1u << lane) - 1u)) == 0) basee = atomicAdd(ecount, __popc(bal));   // first needing lane of the warp
            basee = __shfl_sync(bal, basee, __ffs(bal) - 1);
            slot_e = basee + __popc(bal & ((1u << lane) - 1u));
            if (slot_e < ETAB_SLOTS) {
                bool ok = true;
                for (int k = 0; k < SK8; k++) {
                    const int x = gx + k;
                    uint2 e = make_uint2(1u, 0u);                    // pixels beyond the page: any divisor (never stored)
                    if (x < W) {
                        e = sMX[min(W, x + p.r + 1) - max(0, x - p.l + 1)];
                        ok = ok && e.x != 0u;
                    }
                    etab[slot_e * 8 + k] = e;
                }
                edge_ok = ok;
                edge_tab = reinterpret_cast<const uint8_t *>(etab + slot_e * 8);
            }
        }
    }
    __syncthreads();                                 // barrier init + prefix zero entries + edge tables visible

    for (int it = 0; it < n_it; it++) {
        const int y = y_start + it;
        const bool main = it >= it_main;
        if (it + 1 >= it_main && it + 1 < n_it) {            // the next iteration is a main one: its delay-line rows
            cN[0] = 0; cN[1] = 0;
            if (y >= by0) load_row(y, cN);
            load_row(y + 3 - p.o, lN);
        }
        // ================= B: scan of the column sums of row y+1
        uint32_t bs = 0, bq = 0;
        if (main) {
            const uint32_t ts = ((cs[0] + cs[1]) + (cs[2] + cs[3])) + ((cs[4] + cs[5]) + (cs[6] + cs[7]));
            const uint32_t tq = ((cq[0] + cq[1]) + (cq[2] + cq[3])) + ((cq[4] + cq[5]) + (cq[6] + cq[7]));
            uint32_t ws = ts, wq = tq;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) scan_up2f(ws, wq, d);
            if (lane == 31) sWT[(it & 1) * 8 + warp] = make_uint2(ws, wq);
            bs = ws - ts; bq = wq - tq;              // exclusive within the warp
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_cta(rowbar);

        // ================= A: source row -> gray row rc
        mbar_wait(&mbar[slot], (uint32_t)par);
        uint32_t gw[2];
        const uint8_t *sp = stage + slot * stage_stride + i0 * CS;
        if (CS == 3) {
            const uint2 a = *reinterpret_cast<const uint2 *>(sp), b = *reinterpret_cast<const uint2 *>(sp + 8),
                        c = *reinterpret_cast<const uint2 *>(sp + 16);
            const uint32_t wv[6] = {a.x, a.y, b.x, b.y, c.x, c.y};
            constexpr uint32_t KH0 = 76u | (150u << 8) | (29u << 16), KL0 = 139u | (70u << 8) | (47u << 16);
            constexpr uint32_t KH1 = KH0 << 8, KL1 = KL0 << 8;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t x0 = wv[3 * h], x1 = wv[3 * h + 1], x2 = wv[3 * h + 2];
                const uint32_t p1 = __byte_perm(x0, x1, 0x0543), p2 = __byte_perm(x1, x2, 0x0432);
                const uint32_t t0 = __dp4a(x0, KH0, 0u) * 256u + __dp4a(x0, KL0, 0x8000u);
                const uint32_t t1 = __dp4a(p1, KH0, 0u) * 256u + __dp4a(p1, KL0, 0x8000u);
                const uint32_t t2 = __dp4a(p2, KH0, 0u) * 256u + __dp4a(p2, KL0, 0x8000u);
                const uint32_t t3 = __dp4a(x2, KH1, 0u) * 256u + __dp4a(x2, KL1, 0x8000u);
                gw[h] = __byte_perm(__byte_perm(t0, t1, 0x0062), __byte_perm(t2, t3, 0x0062), 0x5410);   // L = bits 16..23
            }
        } else {
            const uint2 a = *reinterpret_cast<const uint2 *>(sp);
            gw[0] = a.x; gw[1] = a.y;
        }
        uint32_t G[2];
        if constexpr (R > 0) {
            // neighbour columns: R to the left (from lane - 1), R to the right (from lane + 1)
            uint32_t left = __shfl_up_sync(FULLM, gw[1], 1), right = __shfl_down_sync(FULLM, gw[0], 1);
            if (lane == 0 || lane == 31) {
                // the neighbour lane belongs to another warp: convert its pixels from the stage
                uint32_t acc = 0;
#pragma unroll
                for (int j = 0; j < R; j++) {
                    const uint8_t *q = lane == 0 ? sp - (R - j) * CS : sp + (SK8 + j) * CS;
                    const uint32_t g = CS == 3 ? luma_l24(q[0], q[1], q[2]) : (uint32_t)q[0];
                    acc |= g << (8 * (lane == 0 ? 4 - R + j : j));
                }
                if (lane == 0) left = acc; else right = acc;
            }
#pragma unroll
            for (int j = 0; j < 2 * R; j++) { win[j][0] = win[j + 1][0]; win[j][1] = win[j + 1][1]; win[j][2] = win[j + 1][2]; }
            win[2 * R][0] = R == 1 ? __byte_perm(left, right, 0x0043) : __byte_perm(left, right, 0x5432);
            win[2 * R][1] = gw[0]; win[2 * R][2] = gw[1];
            // column e of window row j: e < R left neighbours, e >= R + 8 right neighbours
            auto colv = [&](const int j, const int e) -> uint32_t {
                if (e < R) return byte_of(win[j][0], e);
                if (e >= R + SK8) return byte_of(win[j][0], e - SK8);
                return byte_of(win[j][1 + ((e - R) >> 2)], (e - R) & 3);
            };
            float vf[NE];
#pragma unroll
            for (int e = 0; e < NE; e++) {
                double acc = __dmul_rn(u2d(colv(R, e)), w0);
                if constexpr (R > 1) acc = __dadd_rn(acc, __dmul_rn(u2d(colv(0, e) + colv(2 * R, e)), w2));
                acc = __dadd_rn(acc, __dmul_rn(u2d(colv(R - 1, e) + colv(R + 1, e)), w1));
                vf[e] = __double2float_rn(acc);
            }
            // columns outside the page are scipy-reflected; vertical-pass results mirror with their columns
            if (edge_l) {
#pragma unroll
                for (int j = 0; j < R; j++) vf[R - 1 - j] = vf[R + j];
            }
            if (edge_r) {
                // column W + j mirrors to W - 1 - j; only the R columns next to the page edge feed a valid output.
                // d = index of column W in vf[] (1 .. NE - 1), unrolled so that every access is a register
                const int d = W - gx + R;
#pragma unroll
                for (int dd = 1; dd < NE; dd++)
                    if (d == dd) {
                        vf[dd] = vf[dd - 1];
                        if (R > 1 && dd + 1 < NE && dd >= 2) vf[dd + 1] = vf[dd - 2];
                    }
            }
            double ext[NE];
#pragma unroll
            for (int e = 0; e < NE; e++) ext[e] = (double)vf[e];
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
        const int rc = y + 2 + p.u;
        const bool rc_ok = rc >= g0 && rc < H;
        if (PRODUCE) {
            if (rc >= g0) {
                if (rc_ok && wr_ok) *reinterpret_cast<uint2 *>(grow) = make_uint2(G[0], G[1]);
                grow += p.gray_pitch;
            }
        }
        if (!rc_ok) { G[0] = 0; G[1] = 0; }
        G[0] &= lmask[0]; G[1] &= lmask[1];

        // ================= D: test of row y-1 (prefix published two iterations ago, pixels from the delay line)
        if (main && y - 1 >= by0) {
            if (is_out && !(p.dbg & 1)) {
                const int yt = y - 1;
                const uint8_t *P = pb + o_test;
                const int ny = min(H, yt + p.u + 1) - max(0, yt - p.o + 1);
                if (ny != ny_cached) { const uint2 m = sMY[ny]; magicM = m.x; magicS = m.y; ny_cached = ny; }
                uint32_t S[SK8], Q[SK8];
#pragma unroll
                for (int k = 0; k < SK8; k++) {
                    const uint2 hi = *reinterpret_cast<const uint2 *>(P + p.bo_hi[k]);
                    const uint2 lo = *reinterpret_cast<const uint2 *>(P + p.bo_lo[k]);
                    S[k] = hi.x - lo.x; Q[k] = hi.y - lo.y;
                }
                // The reference's FP64 test is a function of three integers -- pixel, mean m = S / n, variance v = Q / n - m*m --
                // and monotone in v, so it is tabulated exactly: vmin[m][p] = the least v that makes the pixel foreground
                // (0: always, 0xffff: never; built once per (k, R) by k_sauvola_vmin with the reference's own operations).
                auto test_fast = [&](const int k, const uint32_t M, const uint32_t sh) -> uint32_t {
                    const uint32_t m = __umulhi(S[k], M) >> sh, q = __umulhi(Q[k], M) >> sh;
                    const uint32_t v = q - m * m;
                    const uint32_t T = __ldg(vmin + (((m << 8) | byte_of(cA[k >> 2], k & 3)) & 0xffffu));
                    return (uint32_t)(v >= T);
                };
                uint32_t bits[2] = {0, 0};
                // (magic, shift) of every pixel's window area: interior columns share the row's entry sMY[ny] (stride 0);
                // columns whose window the page clamps have their own row-invariant entries sMX[nx] (interior rows)
                const bool area_by_table = nx_uniform ? magicM != 0u : (ny == p.wh && edge_ok);
                const unsigned act = __activemask();
                if (__all_sync(act, nx_uniform && magicM != 0u)) {
                    // the common warp: one window area per row for every pixel
#pragma unroll
                    for (int k = 0; k < SK8; k++) bits[k >> 2] |= test_fast(k, magicM, magicS) << (8 * (k & 3));
                } else if (__all_sync(act, area_by_table)) {
                    // a warp at the left / right page edge: per-pixel areas through a pointer (stride 0 for its interior threads)
                    const uint8_t *mp = nx_uniform ? reinterpret_cast<const uint8_t *>(sMY + ny) : edge_tab;
                    const int mstride = nx_uniform ? 0 : 8;
#pragma unroll
                    for (int k = 0; k < SK8; k++) {
                        const uint2 mg = *reinterpret_cast<const uint2 *>(mp + k * mstride);
                        bits[k >> 2] |= test_fast(k, mg.x, mg.y) << (8 * (k & 3));
                    }
                } else {
                    // page corners (clamped in both directions), windows too large for the 32-bit magic, pixels beyond the
                    // right page edge: per-pixel area, FP64 quotient floor((a + 0.5) * (1/n)) == a / n (sauvola.cu)
#pragma unroll 1
                    for (int k = 0; k < SK8; k++) {
                        const int x = gx + k;
                        const int nx = (x < W) ? min(W, x + p.r + 1) - max(0, x - p.l + 1) : 0;
                        const int n = nx * ny;
                        if (n > 0) {
                            const uint2 hi = *reinterpret_cast<const uint2 *>(P + p.bo_hi[k]);
                            const uint2 lo = *reinterpret_cast<const uint2 *>(P + p.bo_lo[k]);
                            const double rn = 1.0 / (double)n;
                            const double md = floor(__dmul_rn(__dadd_rn((double)(hi.x - lo.x), 0.5), rn));
                            const double qd = floor(__dmul_rn(__dadd_rn((double)(hi.y - lo.y), 0.5), rn));
                            const double mm = __dmul_rn(md, md);
                            const double vd = __dadd_rn(qd, -mm);
                            const double pd = (double)(((k < 4 ? cA[0] : cA[1]) >> (8 * (k & 3))) & 0xffu);
                            const double t = __dadd_rn(pd, __dmul_rn(md, p.km1));
                            const double rhs = __dmul_rn(__dmul_rn(mm, p.k2), vd);
                            const double lhs = __dmul_rn(t, t);
                            const uint32_t fgb = ((uint32_t)(t <= 0.0) | (uint32_t)(lhs <= rhs)) << (8 * (k & 3));
                            if (k < 4) bits[0] |= fgb; else bits[1] |= fgb;
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

        // ================= WAIT: every warp has published its total of row y+1 (and is done with the previous stage)
        mbar_wait(rowbar, (uint32_t)(it & 1));
        if (tid == 0 && it > 0) {
            const int rs = rs_first + (it - 1) + NS;                 // refill the stage every thread read in iteration it-1
            if (rs <= rs_last) issue_row(rs, slot == 0 ? NS - 1 : slot - 1);
        }
        if (main) {
            // ================= C: CTA-wide prefix of the column sums of row y+1 -> ring buffer o_pub
            uint32_t as = bs, aq = bq;
            const uint2 *T = sWT + (it & 1) * 8;
            for (int w2i = 0; w2i < warp; w2i++) { const uint2 t = T[w2i]; as += t.x; aq += t.y; }
            uint8_t *P = pb + o_pub;
#pragma unroll
            for (int k = 0; k < SK8; k++) {
                as += cs[k]; aq += cq[k];
                *reinterpret_cast<uint2 *>(P + p.bo_st[k]) = make_uint2(as, aq);
            }
            // ================= E: slide: row y+2's window = row y+1's + gray row rc - gray row y+2-o
            const uint32_t l0 = lA[0] & lmask[0], l1 = lA[1] & lmask[1];
#pragma unroll
            for (int k = 0; k < SK8; k++) {
                const uint32_t a = byte_of(G[k >> 2], k & 3), b = byte_of(k < 4 ? l0 : l1, k & 3);
                cs[k] += a - b;
                cq[k] += a * a - b * b;
            }
            const int t_ = o_test; o_test = o_mid; o_mid = o_pub; o_pub = t_;
        } else {
            // warm-up: gray row rc joins the window of the band's first row
#pragma unroll
            for (int k = 0; k < SK8; k++) {
                const uint32_t a = byte_of(G[k >> 2], k & 3);
                cs[k] += a; cq[k] += a * a;
            }
        }
        cA[0] = cN[0]; cA[1] = cN[1]; lA[0] = lN[0]; lA[1] = lN[1];
        if (++slot == NS) { slot = 0; par ^= 1; }
    }
}

template <int C, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_sauvola_fused(const FusedParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem);
    double *sw = reinterpret_cast<double *>(smem + 64);          // 3 weights
    double *sphi = sw + 4;                                       // 5 scratch
    // grid: x = (band, strip) of a page, y = page.  (Placing the CTAs of one page -- one blur radius, one code variant -- on
    // the same SMs was tried and is slower for batches of mixed radii: the SMs holding the heavier variant finish last.)
    const int page = blockIdx.y, item = blockIdx.x;
    const int tid = threadIdx.x;
    const FusedSmem L(blockDim.x, C, p.eb);

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
        uint2 *mx = reinterpret_cast<uint2 *>(smem + L.off_mx);
        for (int nx = tid; nx < 256; nx += blockDim.x) {
            const int n = nx * p.wh;
            uint2 e = make_uint2(0u, 0u);
            if (fast_div_ok(n)) {
                const int sh = 31 - __clz(n - 1);
                e.x = (uint32_t)(((1ull << (32 + sh)) + (unsigned long long)n - 1ull) / (unsigned long long)n);
                e.y = (uint32_t)sh;
            }
            mx[nx] = e;
        }
    }
    if (radius >= 1 && radius <= MAXR) blur_weights_cta(radius, sigma, sw, sphi);

    const uint8_t *src = p.src + (int64_t)page * p.src_stride;
    uint8_t *gray = p.gray ? p.gray + (int64_t)page * p.gray_stride : nullptr;
    uint8_t *out = p.out + (int64_t)page * p.out_stride;
    if (radius > MAXR) {
        fused_march<1, 0, false>(p, smem, gray, p.gray_pitch, nullptr, out, sw, item);      // pre-blurred by gray_blur.cu
    } else if (C == 3) {
        if (radius == 0) fused_march<3, 0, true>(p, smem, src, p.src_pitch, gray, out, sw, item);
        else if (radius == 1) fused_march<3, 1, true>(p, smem, src, p.src_pitch, gray, out, sw, item);
        else fused_march<3, 2, true>(p, smem, src, p.src_pitch, gray, out, sw, item);
    } else {
        if (radius == 0) fused_march<1, 0, false>(p, smem, src, p.src_pitch, nullptr, out, sw, item);
        else if (radius == 1) fused_march<1, 1, true>(p, smem, src, p.src_pitch, gray, out, sw, item);
        else fused_march<1, 2, true>(p, smem, src, p.src_pitch, gray, out, sw, item);
    }
}

}  // namespace

// vmin[m << 8 | p]: least variance v (integer, 0..65025) for which the reference's test (sauvola.pyx:6-7, k >= 0)
//   t = p + m*(k - 1);  fg = t <= 0 || t*t <= ((m*m)*k2)*v        (every operation rounded on its own)
// holds for pixel p and window mean m; 0xffff when no v does.  The right-hand side is monotone in v, so a bisection with
// the reference's own operations gives the exact boundary: the kernel's integer comparison v >= vmin decides every
// pixel exactly as the FP64 expression would.
__global__ void __launch_bounds__(256) k_sauvola_vmin(uint16_t *tab, double km1, double k2)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;              // m << 8 | p
    const double md = (double)(idx >> 8), pd = (double)(idx & 255);
    const double t = __dadd_rn(pd, __dmul_rn(md, km1));
    const double lhs = __dmul_rn(t, t), c = __dmul_rn(__dmul_rn(md, md), k2);
    auto fg = [&](int v) { return t <= 0.0 || lhs <= __dmul_rn(c, (double)v); };
    int lo = 0, hi = 65026;                                      // first v in [0, 65026) with fg(v); 65026 = none
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (fg(mid)) hi = mid; else lo = mid + 1;
    }
    tab[idx] = lo > 65025 ? (uint16_t)0xffffu : (uint16_t)lo;
}

// one table per (device, k, R), built at first use on the caller's stream and kept for the life of the process
struct VminKey { int dev; double k, R; bool operator<(const VminKey &o) const { return dev != o.dev ? dev < o.dev : (k != o.k ? k < o.k : R < o.R); } };
std::mutex g_vmin_mu;
std::map<VminKey, uint16_t *> g_vmin;

int sauvola_vmin_table(double k, double Rr, const uint16_t **tab, cudaStream_t st)
{
    int dev = 0;
    B200MRC_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_vmin_mu);
    auto it = g_vmin.find(VminKey{dev, k, Rr});
    if (it != g_vmin.end()) { *tab = it->second; return B200MRC_OK; }
    if (g_vmin.size() >= 64) return B200MRC_ERR_UNSUPPORTED;     // a caller sweeping k: the two-pass kernels take over
    uint16_t *d = nullptr;
    B200MRC_CUDA_TRY(cudaMalloc((void **)&d, 65536 * sizeof(uint16_t)));
    k_sauvola_vmin<<<256, 256, 0, st>>>(d, k - 1.0, k * k / Rr / Rr);
    count_launch();
    // later launches may come on other streams: the table must be complete before this call returns (first use only)
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { cudaFree(d); return (int)e; }
    g_vmin[VminKey{dev, k, Rr}] = d;
    *tab = d;
    return B200MRC_OK;
}

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
    p.eb = (p.l + 7) / 8 * 8;                                     // entry index of column c is c + eb >= 1
    p.sps = (ncols + 2 * p.eb + 8 + 7) / 8;
    for (int k = 0; k < SK8; k++) {
        auto bo = [&](int q) { return ((q & 7) * p.sps + (q >> 3)) * 8; };
        p.bo_st[k] = bo(p.eb + k + 1); p.bo_hi[k] = bo(p.eb + p.r + 1 + k); p.bo_lo[k] = bo(p.eb - p.l + 1 + k);
    }
    { const int trc = sauvola_vmin_table(k, Rr, &p.vmin, st); if (trc != B200MRC_OK) return trc; }
    p.km1 = k - 1.0;
    p.k2 = k * k / Rr / Rr;                                       // sauvola.pyx:60
    p.flags = flags;
#ifdef B200MRC_EXPERIMENTS
    p.dbg = tune(T_FUSED_DBG);                                    // timing experiments: builds with -DB200MRC_EXPERIMENTS only
#else
    p.dbg = 0;                                                    // a shipped library never skips the test, whatever the knob says
#endif
    p.out8 = !(((uintptr_t)out & 7) || (out_pitch & 7) || (out_stride & 7));

    const FusedSmem L(nt, C, p.eb);
    // occupancy variant: CTAs of 128 threads run 4 per SM with 128 registers (default) or 5 / 3 per SM (tuning key FUSED_OCC)
    const int occ_req = nt == 128 ? tune(T_FUSED_OCC) : 0;
    const void *kern;
    if (nt == 128 && occ_req == 5) kern = C == 3 ? (const void *)k_sauvola_fused<3, 128, 5> : (const void *)k_sauvola_fused<1, 128, 5>;
    else if (nt == 128 && occ_req == 3) kern = C == 3 ? (const void *)k_sauvola_fused<3, 128, 3> : (const void *)k_sauvola_fused<1, 128, 3>;
    else kern = C == 3 ? (const void *)k_sauvola_fused<3, 256, 2> : (const void *)k_sauvola_fused<1, 256, 2>;
    int occ = 0;
    B200MRC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    B200MRC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nt, L.total));
    if (occ < 1) occ = 1;
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
// Tuning key THRESHOLD_PATH: auto / fused (one kernel, this file) or legacy (gray_blur.cu + sauvola.cu).  Measured on B200,
// 64 RGB pages 3300x2550, window 101 (profiles/r2_threshold_paths.txt): fused 2.4 / 3.3 / 3.9 ms for blur radius 0 / 1 / 2
// against 3.0 / 3.5 / 3.8 ms for the two kernels; batches that mix radii favour the two-pass form by ~15 % (the fused
// kernel runs as one wave, so its slowest CTA sets the time).  auto picks the fused kernel wherever it applies.
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
