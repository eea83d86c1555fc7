// optimise.cu -- k_optimise_fg_bg: optimise_gray/rgb[2] (cython/optimiser.pyx:22-429), both MRC
// layers of create_mrc_hocr_components in ONE sweep over the page:
//     fg = optimise(mask,     img, n_fg = 3)    internetarchivepdf/mrc.py:412-415
//     bg = optimise(mask ^ 1, img, n_bg = 10)   internetarchivepdf/mrc.py:439-449
//
// Semantics (oracle/mrc_oracle.c orc_optimise): out = copy(img); raster order; for each pixel
// NOT in `mask`:  box = [max(0,y-n),min(H,y+n)) x [max(0,x-n),min(W,x+n))
//     num = sum_{box & mask} img  +  sum_{[ys,y) x [xs,x)} out          (FIR + causal IIR)
//     den = #mask in box + (y-ys)(x-xs);  out = den > 0 ? num / den : 0  (C truncation)
// The IIR term reads rows strictly above and columns strictly left: pixels of one row are
// mutually independent, rows are sequential.  Every pixel is computed for
// exactly one of the two layers (fg where mask==0, bg where mask==1) and copied for the other.
//
// B200 mapping: a page is cut into column strips; one CTA marches one strip top to bottom, one
// thread per column (plus 16 halo columns each side).
//   * per column, registers hold the FIR column sums (mask*img, mask count over 2n rows) and the
//     IIR column sums (out over the last n rows) of both layers; each row step adds the entering
//     row and subtracts the leaving row.  Rows live in thread-private smem rings (packed RGBM
//     input ring, packed RGB output rings), so no row is fetched from HBM twice;
//   * the horizontal window sums need the neighbours' column sums: published through smem once
//     per row (2 __syncthreads per row step);
//   * truncating division by den <= (2n)^2+n^2 is an exact multiply-high with a per-CTA table;
//   * strips of one page form a software pipeline: strip s needs out[y-1] of the last columns
//     of strip s-1.  Those 16 columns are published through a global "mailbox" + a per-strip
//     progress counter (release/acquire with __threadfence); a strip runs >= 1 row behind its
//     left neighbour and prefetches the mailbox one row ahead once the neighbour has a lead.
//     CTAs take (page, strip) jobs from a ticket counter in dependency order, so a waiting CTA
//     always waits on a CTA that is already resident (no deadlock, any grid size).
// Algorithmic HBM bytes per pixel (RGB): 3 (img) + 1 (mask) read, 3 (fg) + 3 (bg) written.
#include "common.cuh"
#include <cstring>
#include <cstdlib>

namespace b200mrc {
size_t iirw_mailbox_words(int W, int H, int N);        // optimise_warp.cu
size_t optimise_split_rec_bytes(int W, int H, int N);  // optimise_split.cu
namespace {

constexpr int OH = 16;                 // halo columns each side (>= B200MRC_MAX_OPT_N)

struct OptParams {
    const uint8_t *mask; int64_t mpitch, mstride;
    const uint8_t *img;  int64_t ipitch, istride; int C;
    uint8_t *ofg; int64_t fpitch, fstride; int nfg;
    uint8_t *obg; int64_t bpitch, bstride; int nbg;
    int W, H, N, S, SW;
    uint32_t *mailbox;                 // [N][S][H][2][OH]
    int *prog;                         // [N][S] rows completed
    unsigned *ticket;
};

struct Acc { int r, g, b, m; };

__device__ __forceinline__ void acc_add(Acc &a, uint32_t px) { a.r += px & 0xff; a.g += (px >> 8) & 0xff; a.b += (px >> 16) & 0xff; }
__device__ __forceinline__ void acc_sub(Acc &a, uint32_t px) { a.r -= px & 0xff; a.g -= (px >> 8) & 0xff; a.b -= (px >> 16) & 0xff; }

__device__ __forceinline__ uint32_t load_px(const OptParams &p, const uint8_t *img, const uint8_t *mask, int y, int x)
{
    const uint32_t m = mask[(int64_t)y * p.mpitch + x] != 0;
    const uint8_t *px = img + (int64_t)y * p.ipitch + (int64_t)x * p.C;
    uint32_t v;
    if (p.C == 1) { v = px[0]; v |= (v << 8) | (v << 16); }
    else v = px[0] | ((uint32_t)px[1] << 8) | ((uint32_t)px[2] << 16);
    return v | (m << 24);
}

__device__ __forceinline__ void store_px(uint8_t *plane, int64_t pitch, int C, int y, int x, uint32_t v)
{
    uint8_t *o = plane + (int64_t)y * pitch + (int64_t)x * C;
    o[0] = (uint8_t)(v & 0xff);
    if (C == 3) { o[1] = (uint8_t)((v >> 8) & 0xff); o[2] = (uint8_t)((v >> 16) & 0xff); }
}

__device__ __forceinline__ uint32_t div_magic(uint32_t num, int den, const uint32_t *M)
{
    return den == 1 ? num : __umulhi(num, M[den]);
}

// window sums and the division for one layer at thread column `tid`
__device__ __forceinline__ uint32_t solve_layer(const uint4 *exA, const uint4 *exF, int tid, int n, int y, int x,
                                                const uint32_t *M)
{
    int r = 0, g = 0, b = 0, m = 0;
    for (int j = 1; j <= n; j++) { const uint4 a = exA[tid - j]; r += a.x; g += a.y; b += a.z; m += a.w; }
    for (int j = 0; j < n; j++)  { const uint4 a = exF[tid + j]; r += a.x; g += a.y; b += a.z; m += a.w; }
    const int ys = max(0, y - n), xs = max(0, x - n);
    const int den = m + (y - ys) * (x - xs);
    if (den <= 0) return 0u;
    return div_magic(r, den, M) | (div_magic(g, den, M) << 8) | (div_magic(b, den, M) << 16);
}

__global__ void __launch_bounds__(1024) k_optimise_fg_bg(const OptParams p)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int T = blockDim.x, tid = threadIdx.x;
    const int nfg = p.nfg, nbg = p.nbg, nmax = max(nfg, nbg);
    const int RIN = 2 * nmax + 2, RFG = nfg + 1, RBG = nbg + 1;
    const int maxden = 4 * nmax * nmax + nmax * nmax;

    uint4 *exAfg = reinterpret_cast<uint4 *>(smem_raw);
    uint4 *exFfg = exAfg + T, *exAbg = exFfg + T, *exFbg = exAbg + T;
    uint32_t *ringIn = reinterpret_cast<uint32_t *>(exFbg + T);
    uint32_t *ringFg = ringIn + (size_t)RIN * T;
    uint32_t *ringBg = ringFg + (size_t)RFG * T;
    uint32_t *M = ringBg + (size_t)RBG * T;
    int *s_misc = reinterpret_cast<int *>(M + maxden + 1);      // [0] job, [1] known progress of the left strip

    if (tid == 0) { s_misc[0] = (int)atomicAdd(p.ticket, 1u); s_misc[1] = 0; }
    for (int d = 1 + tid; d <= maxden; d += T) M[d] = (uint32_t)(0x100000000ull / (unsigned long long)d) + 1u;
    __syncthreads();
    const int job = s_misc[0];
    if (job >= p.N * p.S) return;
    const int page = job / p.S, strip = job - page * p.S;
    const int x0 = strip * p.SW, x = x0 - OH + tid;
    const int W = p.W, H = p.H;
    const bool colvalid = x >= 0 && x < W;
    const bool interior = tid >= OH && tid < OH + p.SW && x < W;
    const bool haloL = tid < OH && strip > 0;
    const bool wants_c = colvalid && tid < OH + p.SW;             // left halo + interior keep IIR sums

    const uint8_t *img = p.img + (int64_t)page * p.istride;
    const uint8_t *mask = p.mask + (int64_t)page * p.mstride;
    uint8_t *ofg = p.ofg ? p.ofg + (int64_t)page * p.fstride : nullptr;
    uint8_t *obg = p.obg ? p.obg + (int64_t)page * p.bstride : nullptr;
    uint32_t *mb_out = p.mailbox + ((int64_t)page * p.S + strip) * (int64_t)H * 2 * OH;
    const uint32_t *mb_in = strip > 0 ? p.mailbox + ((int64_t)page * p.S + strip - 1) * (int64_t)H * 2 * OH : nullptr;
    int *prog_out = p.prog + (int64_t)page * p.S + strip;
    const volatile int *prog_in = strip > 0 ? p.prog + (int64_t)page * p.S + strip - 1 : nullptr;

    Acc Ffg = {0, 0, 0, 0}, Fbg = {0, 0, 0, 0}, Cfg = {0, 0, 0, 0}, Cbg = {0, 0, 0, 0};

    // preload rows 0 .. nmax-1; FIR sums start as rows [0, n-1) (the state "after row -1")
    if (colvalid) {
        for (int ry = 0; ry < nmax && ry < H; ry++) {
            const uint32_t px = load_px(p, img, mask, ry, x);
            ringIn[(size_t)(ry % RIN) * T + tid] = px;
            if (ry < nfg - 1 && (px >> 24)) { acc_add(Ffg, px); Ffg.m++; }
            if (ry < nbg - 1 && !(px >> 24)) { acc_add(Fbg, px); Fbg.m++; }
        }
    }

    uint32_t prev_fg = 0, prev_bg = 0;        // out[y-1] of this column (own or the neighbour's)
    uint32_t pf_fg = 0, pf_bg = 0;            // mailbox prefetch registers (left halo)
    int pf_row = -1;

    for (int y = 0; y < H; y++) {
        // ---- prefetch the next input row into registers
        const int ry_pf = y + nmax;
        uint32_t pf = 0;
        const bool do_pf = colvalid && ry_pf < H;
        if (do_pf) pf = load_px(p, img, mask, ry_pf, x);

        // ---- left halo: out[y-1] of the neighbour strip's last columns
        if (haloL) {
            if (y >= 1) {
                if (pf_row == y - 1) { prev_fg = pf_fg; prev_bg = pf_bg; }
                else {
                    prev_fg = __ldcg(mb_in + ((int64_t)(y - 1) * 2 + 0) * OH + tid);
                    prev_bg = __ldcg(mb_in + ((int64_t)(y - 1) * 2 + 1) * OH + tid);
                }
            }
            if (s_misc[1] >= y + 1) {          // neighbour already finished row y: fetch it early
                pf_fg = __ldcg(mb_in + ((int64_t)y * 2 + 0) * OH + tid);
                pf_bg = __ldcg(mb_in + ((int64_t)y * 2 + 1) * OH + tid);
                pf_row = y;
            }
        }

        // ---- slide the column sums
        if (colvalid) {
            int e = y + nfg - 1, l = y - nfg - 1;
            if (e < H) { const uint32_t px = ringIn[(size_t)(e % RIN) * T + tid]; if (px >> 24) { acc_add(Ffg, px); Ffg.m++; } }
            if (l >= 0) { const uint32_t px = ringIn[(size_t)(l % RIN) * T + tid]; if (px >> 24) { acc_sub(Ffg, px); Ffg.m--; } }
            e = y + nbg - 1; l = y - nbg - 1;
            if (e < H) { const uint32_t px = ringIn[(size_t)(e % RIN) * T + tid]; if (!(px >> 24)) { acc_add(Fbg, px); Fbg.m++; } }
            if (l >= 0) { const uint32_t px = ringIn[(size_t)(l % RIN) * T + tid]; if (!(px >> 24)) { acc_sub(Fbg, px); Fbg.m--; } }
            if (wants_c) {
                if (y >= 1) {
                    acc_add(Cfg, prev_fg); acc_add(Cbg, prev_bg);
                    if (y - nfg - 1 >= 0) acc_sub(Cfg, ringFg[(size_t)((y - nfg - 1) % RFG) * T + tid]);
                    if (y - nbg - 1 >= 0) acc_sub(Cbg, ringBg[(size_t)((y - nbg - 1) % RBG) * T + tid]);
                    ringFg[(size_t)((y - 1) % RFG) * T + tid] = prev_fg;
                    ringBg[(size_t)((y - 1) % RBG) * T + tid] = prev_bg;
                }
            }
        }
        exAfg[tid] = make_uint4(Ffg.r + Cfg.r, Ffg.g + Cfg.g, Ffg.b + Cfg.b, Ffg.m);
        exFfg[tid] = make_uint4(Ffg.r, Ffg.g, Ffg.b, Ffg.m);
        exAbg[tid] = make_uint4(Fbg.r + Cbg.r, Fbg.g + Cbg.g, Fbg.b + Cbg.b, Fbg.m);
        exFbg[tid] = make_uint4(Fbg.r, Fbg.g, Fbg.b, Fbg.m);
        __syncthreads();

        // ---- this row's outputs
        if (interior) {
            const uint32_t cur = ringIn[(size_t)(y % RIN) * T + tid];
            const uint32_t rgb = cur & 0xffffffu;
            uint32_t vfg, vbg;
            if (cur >> 24) { vfg = rgb; vbg = solve_layer(exAbg, exFbg, tid, nbg, y, x, M); }
            else           { vbg = rgb; vfg = solve_layer(exAfg, exFfg, tid, nfg, y, x, M); }
            if (ofg) store_px(ofg, p.fpitch, p.C, y, x, vfg);
            if (obg) store_px(obg, p.bpitch, p.C, y, x, vbg);
            prev_fg = vfg; prev_bg = vbg;
            if (tid >= p.SW) {                 // last OH interior columns feed the right neighbour
                __stcg(mb_out + ((int64_t)y * 2 + 0) * OH + (tid - p.SW), vfg);
                __stcg(mb_out + ((int64_t)y * 2 + 1) * OH + (tid - p.SW), vbg);
                __threadfence();
            }
        }
        if (do_pf) ringIn[(size_t)(ry_pf % RIN) * T + tid] = pf;

        // ---- make sure the left neighbour has published row y before anybody needs it
        if (tid == 0 && strip > 0 && y + 1 < H) {
            int v = *prog_in;
            while (v < y + 1) { __nanosleep(32); v = *prog_in; }
            __threadfence();
            s_misc[1] = v;
        }
        __syncthreads();
        if (tid == 0) { __threadfence(); *reinterpret_cast<volatile int *>(prog_out) = y + 1; }
    }
}

size_t opt_smem_bytes(int T, int nfg, int nbg)
{
    const int nmax = nfg > nbg ? nfg : nbg;
    const size_t words = (size_t)(2 * nmax + 2) * T + (size_t)(nfg + 1) * T + (size_t)(nbg + 1) * T + (5 * nmax * nmax + 1) + 4;
    return (size_t)4 * sizeof(uint4) * T + words * 4;
}

struct OptPlan { int S, SW, T; size_t smem; };

int eval_plan(int W, int N, int nfg, int nbg, int SW, OptPlan &c, bool &fits, double &warps)
{
    const DevInfo &di = dev_info();
    c.SW = SW; c.S = cdiv(W, SW); c.T = SW + 2 * OH; c.smem = opt_smem_bytes(c.T, nfg, nbg);
    fits = false; warps = -1.0;
    if (c.smem > (size_t)di.max_smem_optin) return 1;
    int per_sm = 0;
    cudaError_t e = cudaFuncSetAttribute(k_optimise_fg_bg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_optimise_fg_bg, c.T, c.smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) return 1;
    const int64_t cap = (int64_t)per_sm * di.sm_count, ctas = (int64_t)c.S * N;
    fits = ctas <= cap;
    warps = (double)(fits ? ctas : cap) * ((c.T + 31) / 32) / di.sm_count;
    return 0;
}

int plan_optimise(int W, int N, int nfg, int nbg, OptPlan &best)
{
    // Pick the strip width.  Preference: (1) every CTA of the batch resident at once and >= 24
    // resident warps per SM to hide the row-step latency: the widest such strips (least halo
    // overhead); (2) resident at once: the most warps; (3) a batch too large to be resident runs
    // in ticket order over several waves: 256-column strips (throughput-bound regime).
    // last plan per host thread (the search calls cudaFuncSetAttribute / the occupancy query ~60 times)
    thread_local int cW = -1, cN = -1, cfg = -1, cbg = -1, cdev = -1;
    thread_local OptPlan cplan;
    int dev = 0;
    cudaGetDevice(&dev);
    if (cW == W && cN == N && cfg == nfg && cbg == nbg && cdev == dev) { best = cplan; return B200MRC_OK; }
    OptPlan pick{0, 0, 0, 0}, c;
    double pick_warps = -1.0, warps;
    bool fits, good = false;
    for (int SW = 992; SW >= 32 && !good; SW -= 16) {
        if (SW > 32 && cdiv(W, SW - 16) == cdiv(W, SW)) continue;   // a narrower strip gives the same count
        const int rc = eval_plan(W, N, nfg, nbg, SW, c, fits, warps);
        if (rc > 1) return rc;
        if (rc == 1 || !fits) continue;
        if (warps >= 24.0) { pick = c; good = true; }
        else if (warps > pick_warps) { pick = c; pick_warps = warps; }
    }
    if (pick.T == 0) {
        for (int SW = 256; SW >= 32 && pick.T == 0; SW -= 16) {
            const int rc = eval_plan(W, N, nfg, nbg, SW, c, fits, warps);
            if (rc > 1) return rc;
            if (rc == 0) pick = c;
        }
    }
    if (pick.T == 0) return B200MRC_ERR_UNSUPPORTED;
    best = pick;
    cW = W; cN = N; cfg = nfg; cbg = nbg; cdev = dev; cplan = pick;
    return B200MRC_OK;
}

struct OptLayout { size_t off_mailbox, off_prog, off_ticket, off_rec, total; };

OptLayout opt_layout(int W, int H, int N)
{
    // sized for the worst case strip count (narrowest strips)
    const size_t Smax = (size_t)cdiv(W, 32);
    Carver c;
    OptLayout L;
    L.off_mailbox = c.take<uint32_t>(std::max((size_t)N * Smax * H * 2 * OH, iirw_mailbox_words(W, H, N)));
    L.off_prog = c.take<int>((size_t)N * Smax);
    L.off_ticket = c.take<unsigned>(4);
    L.off_rec = c.take<uint8_t>(optimise_split_rec_bytes(W, H, N));   // FIR record plane (optimise_split.cu)
    L.total = c.used();
    return L;
}

}  // namespace

int launch_optimise_split(const uint8_t *mask, int64_t mpitch, int64_t mstride,
                          const uint8_t *img, int64_t ipitch, int64_t istride, int C,
                          uint8_t *ofg, int64_t fpitch, int64_t fstride,
                          uint8_t *obg, int64_t bpitch, int64_t bstride,
                          int W, int H, int N, uint8_t *rec, uint32_t *mailbox, unsigned *ticket, int *progress, cudaStream_t st);

size_t optimise_workspace_bytes(int W, int H, int N) { return opt_layout(W, H, N).total; }

int launch_optimise(const uint8_t *mask, int64_t mpitch, int64_t mstride,
                    const uint8_t *img, int64_t ipitch, int64_t istride, int C,
                    uint8_t *ofg, int64_t fpitch, int64_t fstride, int nfg,
                    uint8_t *obg, int64_t bpitch, int64_t bstride, int nbg,
                    int W, int H, int N, void *workspace, size_t workspace_bytes, cudaStream_t st, int **bg_progress)
{
    // bg_progress (optional): on return *bg_progress is null, or a device array [N][cdiv(W, 128)] in which the sweep
    // publishes, per 128-column strip, how many rows of `obg` are globally visible (release stores) -- a kernel launched
    // behind the sweep as its programmatic dependent may consume bg rows while the sweep is still running (resample.cu)
    if (bg_progress) *bg_progress = nullptr;
    const OptLayout L = opt_layout(W, H, N);
    if (!workspace || workspace_bytes < L.total) return B200MRC_ERR_WORKSPACE;
    uint8_t *ws = (uint8_t *)workspace;
    // tuning key OPT_PATH: split (default) | generic, A/B switch for profiling
    if (nfg == 3 && nbg == 10 && tune(T_OPT_PATH) == 0) {
        // production path: parallel FIR record plane + row-sequential warp-strip sweep (optimise_split.cu)
        int *prog = bg_progress ? (int *)(ws + L.off_prog) : nullptr;
        const int frc = launch_optimise_split(mask, mpitch, mstride, img, ipitch, istride, C, ofg, fpitch, fstride,
                                              obg, bpitch, bstride, W, H, N, ws + L.off_rec, (uint32_t *)(ws + L.off_mailbox),
                                              (unsigned *)(ws + L.off_ticket), prog, st);
        if (frc == B200MRC_OK && bg_progress) *bg_progress = prog;
        if (frc != B200MRC_ERR_UNSUPPORTED) return frc;
    }
    OptPlan plan;
    int rc = plan_optimise(W, N, nfg, nbg, plan);
    if (rc != B200MRC_OK) return rc;
    OptParams p;
    p.mask = mask; p.mpitch = mpitch; p.mstride = mstride;
    p.img = img; p.ipitch = ipitch; p.istride = istride; p.C = C;
    p.ofg = ofg; p.fpitch = fpitch; p.fstride = fstride; p.nfg = nfg;
    p.obg = obg; p.bpitch = bpitch; p.bstride = bstride; p.nbg = nbg;
    p.W = W; p.H = H; p.N = N; p.S = plan.S; p.SW = plan.SW;
    p.mailbox = (uint32_t *)(ws + L.off_mailbox);
    p.prog = (int *)(ws + L.off_prog);
    p.ticket = (unsigned *)(ws + L.off_ticket);
    B200MRC_CUDA_TRY(cudaMemsetAsync(p.prog, 0, sizeof(int) * (size_t)N * plan.S, st));
    B200MRC_CUDA_TRY(cudaMemsetAsync(p.ticket, 0, sizeof(unsigned) * 4, st));
    B200MRC_CUDA_TRY(cudaFuncSetAttribute(k_optimise_fg_bg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    { ProfScope _ps("k_optimise_fg_bg", st); k_optimise_fg_bg<<<(unsigned)(plan.S * N), plan.T, plan.smem, st>>>(p); }
    B200MRC_LAUNCH_CHECK();
    return B200MRC_OK;
}

}  // namespace b200mrc

using namespace b200mrc;

extern "C" size_t b200mrc_optimise_workspace_bytes(int width, int height, int n_pages)
{
    if (width <= 0 || height <= 0 || n_pages <= 0) return 0;
    return optimise_workspace_bytes(width, height, n_pages);
}

extern "C" int b200mrc_optimise(const uint8_t *mask, int64_t mask_pitch, int64_t mask_page_stride,
                                const uint8_t *img, int64_t img_pitch, int64_t img_page_stride, int channels,
                                uint8_t *out_fg, int64_t fg_pitch, int64_t fg_page_stride, int n_fg,
                                uint8_t *out_bg, int64_t bg_pitch, int64_t bg_page_stride, int n_bg,
                                int width, int height, int n_pages,
                                void *workspace, size_t workspace_bytes, void *stream)
{
    if (!mask || !img || width <= 0 || height <= 0 || n_pages <= 0) return B200MRC_ERR_INVALID;
    if (!out_fg && !out_bg) return B200MRC_ERR_INVALID;
    if (channels != 1 && channels != 3) return B200MRC_ERR_UNSUPPORTED;
    if (!out_fg) n_fg = 1;
    if (!out_bg) n_bg = 1;
    if (n_fg < 1 || n_bg < 1 || n_fg > B200MRC_MAX_OPT_N || n_bg > B200MRC_MAX_OPT_N) return B200MRC_ERR_UNSUPPORTED;
    return launch_optimise(mask, mask_pitch, mask_page_stride, img, img_pitch, img_page_stride, channels,
                           out_fg, fg_pitch, fg_page_stride, n_fg, out_bg, bg_pitch, bg_page_stride, n_bg,
                           width, height, n_pages, workspace, workspace_bytes, (cudaStream_t)stream, nullptr);
}
