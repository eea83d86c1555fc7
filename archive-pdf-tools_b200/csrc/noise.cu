// noise.cu -- k_noise_dd + k_noise_select: estimate_noise (internetarchivepdf/mrc.py:273-296)
// -> mean_estimate_sigma (mrc.py:52-55) -> scikit-image restoration.estimate_sigma:
//     sigma = median(|dd[dd != 0]|) / 0.6744897501960817,
//     dd = PyWavelets dwtn(crop, 'db2')['dd']  (mode 'symmetric', float32 for float32 input)
// on the centre crop rows [int(h/2-h/4), int(h/2+h/4)), cols likewise (mrc.py:278-292).
// scikit-image / PyWavelets are third-party code that is neither installed nor vendored by the
// reference: this follows their published algorithm ("parity unpinned"); the CPU
// restatement it is tested against is oracle/mrc_oracle.c orc_estimate_sigma_crop.
//
//   k_noise_dd     : one thread per dd coefficient; gray conversion fused into the loads; the
//                    float32 accumulation order of PyWavelets' convolution is kept with
//                    __fmul_rn/__fadd_rn (no FMA); |dd| is stored as its IEEE bit pattern
//                    (monotone as uint32 for non-negative floats), 0 for dd == 0.
//   k_noise_select : one thread-block cluster (1..8 CTAs, histograms summed through distributed shared memory) per page,
//                    exact order statistics by 3-pass radix select
//                    (12 + 10 + 10 bits) over smem histograms; both middle ranks are tracked
//                    so an even count reproduces np.median's float32 mean of two.
#include "common.cuh"
#include <cooperative_groups.h>
#include <math_constants.h>
#include <cstdlib>

namespace b200mrc {
namespace {
namespace cg = cooperative_groups;

struct NoiseParams {
    const uint8_t *in; int64_t in_pitch, in_stride; int C;
    int W, H;
    int hs, he, ws, we;          // crop
    int oh, ow;                  // dd size
    uint32_t *keys;              // N x oh*ow
    double *sigma_out;           // N
};

__device__ __forceinline__ int sym_idx(int i, int n)
{
    if (n == 1) return 0;
    const int per = 2 * n;
    i %= per; if (i < 0) i += per;
    return i < n ? i : per - 1 - i;
}

__global__ void __launch_bounds__(256) k_noise_dd(const NoiseParams p)
{
    const int xo = blockIdx.x * 32 + (threadIdx.x & 31);
    const int yo = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int page = blockIdx.z;
    if (xo >= p.ow || yo >= p.oh) return;
    const float f0 = (float)-0.48296291314469025, f1 = (float)0.836516303737469,
                f2 = (float)-0.22414386804185735, f3 = (float)-0.12940952255092145;
    const float f[4] = {f0, f1, f2, f3};
    const int h = p.he - p.hs, w = p.we - p.ws;
    const uint8_t *in = p.in + (int64_t)page * p.in_stride;
    int rows[4];
#pragma unroll
    for (int j = 0; j < 4; j++) rows[j] = p.hs + sym_idx(2 * yo + 1 - j, h);
    float dd = 0.0f;
#pragma unroll
    for (int j2 = 0; j2 < 4; j2++) {
        const int col = p.ws + sym_idx(2 * xo + 1 - j2, w);
        float d0 = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint8_t *px = in + (int64_t)rows[j] * p.in_pitch + (int64_t)col * p.C;
            const uint32_t g = p.C == 1 ? (uint32_t)px[0] : luma_l24(px[0], px[1], px[2]);
            d0 = __fadd_rn(d0, __fmul_rn(f[j], (float)g));
        }
        dd = __fadd_rn(dd, __fmul_rn(f[j2], d0));
    }
    const uint32_t key = dd == 0.0f ? 0u : (__float_as_uint(dd) & 0x7fffffffu);
    const int64_t stride = ((int64_t)p.oh * p.ow + 3) & ~3ll;
    p.keys[(int64_t)page * stride + (int64_t)yo * p.ow + xo] = key;
    if (yo == p.oh - 1 && xo == p.ow - 1)
        for (int64_t q = (int64_t)p.oh * p.ow; q < stride; q++) p.keys[(int64_t)page * stride + q] = 0u;
}


// Separable marching form of k_noise_dd: a warp owns 31 output columns (lane 0 is a ghost that only supplies the
// left neighbour's pair) and marches down a band of output rows.  Every lane keeps the gray values of its two input
// columns for the 4 rows of the vertical filter in registers (2 new rows per step, each pixel converted once),
// computes the two vertical results exactly like k_noise_dd and takes the other two from its left neighbour by
// shuffle.  Same float32 operations in the same order as k_noise_dd, ~7x fewer instructions.
constexpr int DD_BAND = 32;

__global__ void __launch_bounds__(128) k_noise_dd_march(const NoiseParams p)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int xo = (blockIdx.x * 4 + wid) * 31 + lane - 1;            // lane 0: ghost column pair of xo0 - 1
    const int yo0 = blockIdx.y * DD_BAND, yo1 = min(p.oh, yo0 + DD_BAND);
    const int page = blockIdx.z;
    if ((blockIdx.x * 4 + wid) * 31 >= p.ow) return;
    const float f0 = (float)-0.48296291314469025, f1 = (float)0.836516303737469,
                f2 = (float)-0.22414386804185735, f3 = (float)-0.12940952255092145;
    const int h = p.he - p.hs, w = p.we - p.ws;
    const uint8_t *in = p.in + (int64_t)page * p.in_stride;
    // this lane's two source columns: taps j2 = 1 (2xo) and j2 = 0 (2xo + 1); xo may be -1 (ghost) or >= ow (idle)
    const int64_t ca = (int64_t)(p.ws + sym_idx(2 * xo, w)) * p.C, cb = (int64_t)(p.ws + sym_idx(2 * xo + 1, w)) * p.C;
    auto gray2 = [&](int row, float &ga, float &gb) {
        const uint8_t *r = in + (int64_t)row * p.in_pitch;
        uint32_t a, b;
        if (p.C == 1) { a = r[ca]; b = r[cb]; }
        else { a = luma_l24(r[ca], r[ca + 1], r[ca + 2]); b = luma_l24(r[cb], r[cb + 1], r[cb + 2]); }
        ga = __int_as_float(0x4B000000 | a) - 8388608.0f;            // exact (float)a for a < 2^23, no conversion unit
        gb = __int_as_float(0x4B000000 | b) - 8388608.0f;
    };
    // rows of the vertical taps j = 0..3 are 2yo+1, 2yo, 2yo-1, 2yo-2 (symmetric extension)
    int pr0 = -1, pr1 = -1;                                           // rows held in (a0,b0) / (a1,b1) from the previous step
    float a0 = 0, b0 = 0, a1 = 0, b1 = 0, a2, b2, a3, b3;
    const int64_t stride = ((int64_t)p.oh * p.ow + 3) & ~3ll;
    for (int yo = yo0; yo < yo1; yo++) {
        const int r0 = p.hs + sym_idx(2 * yo + 1, h), r1 = p.hs + sym_idx(2 * yo, h),
                  r2 = p.hs + sym_idx(2 * yo - 1, h), r3 = p.hs + sym_idx(2 * yo - 2, h);
        if (r2 == pr0 && r3 == pr1) { a2 = a0; b2 = b0; a3 = a1; b3 = b1; }   // the usual case: slide by two rows
        else { gray2(r2, a2, b2); gray2(r3, a3, b3); }
        gray2(r0, a0, b0); gray2(r1, a1, b1);
        pr0 = r0; pr1 = r1;
        float da = 0.0f, db = 0.0f;
        da = __fadd_rn(da, __fmul_rn(f0, a0)); da = __fadd_rn(da, __fmul_rn(f1, a1)); da = __fadd_rn(da, __fmul_rn(f2, a2)); da = __fadd_rn(da, __fmul_rn(f3, a3));
        db = __fadd_rn(db, __fmul_rn(f0, b0)); db = __fadd_rn(db, __fmul_rn(f1, b1)); db = __fadd_rn(db, __fmul_rn(f2, b2)); db = __fadd_rn(db, __fmul_rn(f3, b3));
        // horizontal taps j2 = 0..3: columns 2xo+1 (db), 2xo (da), 2xo-1 (left db), 2xo-2 (left da)
        const float lb = __shfl_up_sync(0xffffffffu, db, 1), la = __shfl_up_sync(0xffffffffu, da, 1);
        float dd = 0.0f;
        dd = __fadd_rn(dd, __fmul_rn(f0, db)); dd = __fadd_rn(dd, __fmul_rn(f1, da)); dd = __fadd_rn(dd, __fmul_rn(f2, lb)); dd = __fadd_rn(dd, __fmul_rn(f3, la));
        if (lane >= 1 && xo < p.ow) {
            const uint32_t key = dd == 0.0f ? 0u : (__float_as_uint(dd) & 0x7fffffffu);
            p.keys[(int64_t)page * stride + (int64_t)yo * p.ow + xo] = key;
            if (yo == p.oh - 1 && xo == p.ow - 1)
                for (int64_t q = (int64_t)p.oh * p.ow; q < stride; q++) p.keys[(int64_t)page * stride + q] = 0u;
        }
    }
}

constexpr int SEL_T = 1024;

// Finds the bin holding 0-based rank `rank` in hist[0..nbins) (nbins <= 4*SEL_T) and the rank
// inside that bin.  All threads call; results are broadcast through smem.
__device__ void block_find_rank(const uint32_t *hist, int nbins, uint32_t rank, uint32_t *s_scan,
                                uint32_t *s_res, uint32_t &bin_out, uint32_t &rank_out)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (nbins + SEL_T - 1) / SEL_T;
    uint32_t loc = 0;
    for (int i = 0; i < per; i++) { int b = tid * per + i; if (b < nbins) loc += hist[b]; }
    uint32_t inc = loc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) s_scan[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = s_scan[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += t; }
        s_scan[lane] = v;
    }
    __syncthreads();
    const uint32_t excl = (warp ? s_scan[warp - 1] : 0u) + inc - loc;
    if (rank >= excl && rank < excl + loc) {            // exactly one thread
        uint32_t acc = excl;
        for (int i = 0; i < per; i++) {
            int b = tid * per + i;
            uint32_t c = b < nbins ? hist[b] : 0u;
            if (rank < acc + c) { s_res[0] = (uint32_t)b; s_res[1] = rank - acc; break; }
            acc += c;
        }
    }
    __syncthreads();
    bin_out = s_res[0]; rank_out = s_res[1];
    __syncthreads();
}

// A page is served by a thread-block CLUSTER of 1, 2, 4 or 8 CTAs (as many as keep every SM busy: 64 pages -> 2, 16
// pages -> 8): each CTA histograms its share of the page's keys in its own shared memory, and after every pass all CTAs
// sum the cluster's histograms through distributed shared memory -- every CTA then holds the page's histogram and takes
// the same decisions, so nothing is broadcast.
__global__ void __launch_bounds__(SEL_T) k_noise_select(const NoiseParams p)
{
    __shared__ uint32_t hist[2][4096];
    __shared__ uint32_t s_scan[32];
    __shared__ uint32_t s_res[2];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned CS = cluster.num_blocks(), cr = cluster.block_rank();
    const int page = blockIdx.x / CS, tid = threadIdx.x;
    const int64_t total = ((int64_t)p.oh * p.ow + 3) & ~3ll;          // per-page key stride: padded with zeros to 4
    const uint32_t *keys = p.keys + (int64_t)page * total;
    // this CTA's share of the keys: [k_lo, k_hi), multiples of 4
    const int64_t share = ((total / 4 + CS - 1) / CS) * 4;
    const int64_t k_lo = min(total, (int64_t)cr * share), k_hi = min(total, k_lo + share);
    // hist <- sum over the cluster's CTAs (all 2 x 4096 bins; 8 per thread): read everyone's, then replace one's own
    auto cluster_sum = [&]() {
        if (CS == 1) { __syncthreads(); return; }
        cluster.sync();                                      // every CTA's local histogram is complete
        uint32_t acc[8];
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = 0;
        for (unsigned r = 0; r < CS; r++) {
            const uint32_t *h = cluster.map_shared_rank(&hist[0][0], r);
#pragma unroll
            for (int j = 0; j < 8; j++) acc[j] += h[tid + j * SEL_T];
        }
        cluster.sync();                                      // everyone has read: the local copies may be replaced
#pragma unroll
        for (int j = 0; j < 8; j++) (&hist[0][0])[tid + j * SEL_T] = acc[j];
        __syncthreads();
    };

    for (int i = tid; i < 2 * 4096; i += SEL_T) (&hist[0][0])[i] = 0;
    __syncthreads();
    // pass 1: top 12 bits (non-zero keys only)
    for (int64_t i = k_lo + (int64_t)tid * 4; i < k_hi; i += SEL_T * 8) {
        const uint4 a = *reinterpret_cast<const uint4 *>(keys + i);
        const int64_t i2 = i + SEL_T * 4;
        const uint4 b = i2 < k_hi ? *reinterpret_cast<const uint4 *>(keys + i2) : make_uint4(0, 0, 0, 0);
        const uint32_t kk[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int q = 0; q < 8; q++) if (kk[q]) atomicAdd(&hist[0][kk[q] >> 20], 1u);
    }
    cluster_sum();
    uint32_t cnt;
    {
        // total count of non-zero keys = rank of the (virtual) end
        uint32_t loc = 0;
        for (int b = tid; b < 4096; b += SEL_T) loc += hist[0][b];
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) loc += __shfl_xor_sync(0xffffffffu, loc, d);
        if ((tid & 31) == 0) s_scan[tid >> 5] = loc;
        __syncthreads();
        uint32_t v = 0;
        for (int w2 = 0; w2 < SEL_T / 32; w2++) v += s_scan[w2];
        cnt = v;
        __syncthreads();
    }
    if (cnt == 0) {                                          // the same in every CTA of the cluster
        if (tid == 0 && cr == 0) p.sigma_out[page] = CUDART_NAN;
        return;
    }
    uint32_t rk[2] = {(cnt - 1) / 2, cnt / 2};
    uint32_t pre[2];
    for (int t = 0; t < 2; t++) {
        uint32_t b, r;
        block_find_rank(hist[0], 4096, rk[t], s_scan, s_res, b, r);
        pre[t] = b; rk[t] = r;
    }
    // pass 2: middle 10 bits under each target prefix
    for (int i = tid; i < 2 * 4096; i += SEL_T) (&hist[0][0])[i] = 0;
    __syncthreads();
    for (int64_t i = k_lo + (int64_t)tid * 4; i < k_hi; i += SEL_T * 8) {
        const uint4 a = *reinterpret_cast<const uint4 *>(keys + i);
        const int64_t i2 = i + SEL_T * 4;
        const uint4 b = i2 < k_hi ? *reinterpret_cast<const uint4 *>(keys + i2) : make_uint4(0, 0, 0, 0);
        const uint32_t kk[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const uint32_t k = kk[q];
            if (!k) continue;
            const uint32_t top = k >> 20, mid = (k >> 10) & 0x3ffu;
            if (top == pre[0]) atomicAdd(&hist[0][mid], 1u);
            if (top == pre[1]) atomicAdd(&hist[1][mid], 1u);
        }
    }
    cluster_sum();
    for (int t = 0; t < 2; t++) {
        uint32_t b, r;
        block_find_rank(hist[t], 1024, rk[t], s_scan, s_res, b, r);
        pre[t] = (pre[t] << 10) | b; rk[t] = r;
    }
    // pass 3: low 10 bits
    for (int i = tid; i < 2 * 4096; i += SEL_T) (&hist[0][0])[i] = 0;
    __syncthreads();
    for (int64_t i = k_lo + (int64_t)tid * 4; i < k_hi; i += SEL_T * 8) {
        const uint4 a = *reinterpret_cast<const uint4 *>(keys + i);
        const int64_t i2 = i + SEL_T * 4;
        const uint4 b = i2 < k_hi ? *reinterpret_cast<const uint4 *>(keys + i2) : make_uint4(0, 0, 0, 0);
        const uint32_t kk[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const uint32_t k = kk[q];
            if (!k) continue;
            const uint32_t hi = k >> 10, lo = k & 0x3ffu;
            if (hi == pre[0]) atomicAdd(&hist[0][lo], 1u);
            if (hi == pre[1]) atomicAdd(&hist[1][lo], 1u);
        }
    }
    cluster_sum();
    for (int t = 0; t < 2; t++) {
        uint32_t b, r;
        block_find_rank(hist[t], 1024, rk[t], s_scan, s_res, b, r);
        pre[t] = (pre[t] << 10) | b;
    }
    if (tid == 0 && cr == 0) {
        const float a = __uint_as_float(pre[0]), b = __uint_as_float(pre[1]);
        const float med = (cnt & 1u) ? a : __fdiv_rn(__fadd_rn(a, b), 2.0f);   // np.median (float32)
        p.sigma_out[page] = __ddiv_rn((double)med, 0.6744897501960817);
    }
}

void noise_crop(int W, int H, int &hs, int &he, int &ws, int &we)
{
    // mrc.py:278-292 -- python float arithmetic then int() truncation
    hs = (int)((double)H / 2 - (double)H / 4);
    he = (int)((double)H / 2 + (double)H / 4);
    ws = (int)((double)W / 2 - (double)W / 4);
    we = (int)((double)W / 2 + (double)W / 4);
    if (he == 0 || we == 0) { hs = 0; he = H; ws = 0; we = W; }
}

}  // namespace

size_t noise_workspace_bytes(int W, int H, int N)
{
    int hs, he, ws, we;
    noise_crop(W, H, hs, he, ws, we);
    const size_t oh = (size_t)(he - hs + 3) / 2, ow = (size_t)(we - ws + 3) / 2;
    return align_up(sizeof(uint32_t) * ((oh * ow + 3) & ~(size_t)3) * (size_t)N, 256);
}

int launch_estimate_noise(const uint8_t *in, int64_t in_pitch, int64_t in_stride, int C, int W, int H, int N,
                          double *sigma_out, void *workspace, size_t workspace_bytes, cudaStream_t st)
{
    if (workspace_bytes < noise_workspace_bytes(W, H, N) || !workspace) return B200MRC_ERR_WORKSPACE;
    NoiseParams p;
    p.in = in; p.in_pitch = in_pitch; p.in_stride = in_stride; p.C = C; p.W = W; p.H = H;
    noise_crop(W, H, p.hs, p.he, p.ws, p.we);
    p.oh = (p.he - p.hs + 3) / 2; p.ow = (p.we - p.ws + 3) / 2;
    p.keys = (uint32_t *)workspace;
    p.sigma_out = sigma_out;
    if (tune(T_NOISE_DIRECT)) {                              // A/B switch
        dim3 grid(cdiv(p.ow, 32), cdiv(p.oh, 8), N);
        { ProfScope _ps("k_noise_dd", st); k_noise_dd<<<grid, 256, 0, st>>>(p); }
    } else {
        dim3 grid(cdiv(p.ow, 4 * 31), cdiv(p.oh, DD_BAND), N);
        { ProfScope _ps("k_noise_dd_march", st); k_noise_dd_march<<<grid, 128, 0, st>>>(p); }
    }
    B200MRC_LAUNCH_CHECK();
    {
        // CTAs per page: the largest cluster (<= 8, the portable limit) that still fits every page's cluster on the GPU at once
        int cs = 1;
        while (cs < 8 && (int64_t)N * cs * 2 <= dev_info().sm_count) cs *= 2;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = (unsigned)cs; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(N * cs)); cfg.blockDim = dim3(SEL_T); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        void *args[] = {(void *)&p};
        ProfScope _ps("k_noise_select", st);
        B200MRC_CUDA_TRY(cudaLaunchKernelExC(&cfg, (const void *)k_noise_select, args));
    }
    count_launch();
    return B200MRC_OK;
}

}  // namespace b200mrc

using namespace b200mrc;

extern "C" size_t b200mrc_noise_workspace_bytes(int width, int height, int n_pages)
{
    if (width <= 0 || height <= 0 || n_pages <= 0) return 0;
    return noise_workspace_bytes(width, height, n_pages);
}

extern "C" int b200mrc_estimate_noise(const uint8_t *in, int64_t in_pitch, int64_t in_page_stride, int channels,
                                      int width, int height, int n_pages, double *sigma_out,
                                      void *workspace, size_t workspace_bytes, void *stream)
{
    if (!in || !sigma_out || width <= 0 || height <= 0 || n_pages <= 0) return B200MRC_ERR_INVALID;
    if (channels != 1 && channels != 3) return B200MRC_ERR_UNSUPPORTED;
    if (n_pages > 65535) return B200MRC_ERR_UNSUPPORTED;
    return launch_estimate_noise(in, in_pitch, in_page_stride, channels, width, height, n_pages, sigma_out,
                                 workspace, workspace_bytes, (cudaStream_t)stream);
}
