// denoise.cu -- k_mask_denoise: fast_mask_denoise (cython/optimiser.pyx:436-472, called at
// internetarchivepdf/mrc.py:388 with mincnt=4, n_size=2), exact in-place raster-order semantics.
//
// Reference: for interior pixels (2 <= y < H-2, 2 <= x < W-2) in raster order, a set pixel is
// kept iff (sum of the 5x5 window of the array being written) - 1 >= 4.  Pixels raster-earlier
// than p are read *after* their own update, later ones before.  That is the unique fixed point of
//     r[p] = m0[p] & ( sum_{q earlier} r[q] + sum_{q later} m0[q] >= 4 )
// (the dependency graph is acyclic), and the operator is monotone, so chaotic (asynchronous)
// iteration from r = m0 converges to exactly the sequential result in any update order
// (oracle/mrc_oracle.c orc_denoise).
//
// B200 mapping -- one persistent cooperative kernel, no host round trips:
//   phase 0  bytes -> 1 bit/pixel planes M0 and R (8x less traffic for every later pass);
//   phase 1  global iterations: every CTA walks its share of 64x256-pixel tiles; a tile loads
//            R and M0 (+2 px halo) into smem and iterates to *local* convergence with
//            bit-sliced arithmetic: 32 pixels per 32-bit word, 5-tap horizontal sums as 3-bit
//            planes (2 full adders), vertical combination and the ">= 4" test in ~20 LOP3 per
//            word; the m0-only half of the count is precomputed once per tile.  Tiles whose
//            3x3 neighbourhood did not change in the previous iteration are skipped.  A grid
//            barrier + global "changed" counter ends the loop when a whole pass changes nothing.
//   phase 2  cleared pixels are written back to the byte mask (sparse byte stores).
// Algorithmic HBM bytes: 1 B/px read + (few) cleared bytes written.
#include "common.cuh"

namespace b200mrc {
namespace {

constexpr int DT = 256;        // threads
constexpr int TR = 64;         // tile rows
constexpr int TWW = 8;         // tile words per row (256 px)
constexpr int LR = TR + 4;     // rows incl. halo
constexpr int LW = TWW + 2;    // words incl. halo words

struct DenoiseParams {
    uint8_t *mask; int64_t pitch, stride;
    int W, H, N, Ww;
    uint32_t *Mb, *Rb;
    uint8_t *flags;            // [2][N * tiles]
    int tiles_x, tiles_y;
    unsigned *bar;             // [0] arrivals, [1] generation
    unsigned *changed;         // [3]
    int max_iters;
};

__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned nblocks)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned gen = atomicAdd(&bar[1], 0u);
        if (atomicAdd(&bar[0], 1u) == nblocks - 1) {
            atomicExch(&bar[0], 0u);
            __threadfence();
            atomicAdd(&bar[1], 1u);
        } else {
            while (atomicAdd(&bar[1], 0u) == gen) __nanosleep(64);
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a | b)); }

// bit-sliced sum of the 5 horizontal neighbours (x-2..x+2) of a row: l | c | r words
__device__ __forceinline__ void h5(uint32_t l, uint32_t c, uint32_t r, uint32_t &h0, uint32_t &h1, uint32_t &h2)
{
    const uint32_t a = (c << 2) | (l >> 30), b = (c << 1) | (l >> 31);
    const uint32_t d = (c >> 1) | (r << 31), e = (c >> 2) | (r << 30);
    const uint32_t s1 = a ^ b ^ c, c1 = maj3(a, b, c);
    const uint32_t s0 = s1 ^ d ^ e, c2 = maj3(s1, d, e);
    h0 = s0; h1 = c1 ^ c2; h2 = c1 & c2;
}

// (a + b + e + f) for two 3-bit planes a, b and two 1-bit planes: low two bits and ">= 4"
__device__ __forceinline__ void add33_11(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b0, uint32_t b1, uint32_t b2,
                                         uint32_t e, uint32_t f, uint32_t &s0, uint32_t &s1, uint32_t &ge4)
{
    const uint32_t x0 = a0 ^ b0, k0 = a0 & b0;
    const uint32_t x1 = a1 ^ b1 ^ k0, k1 = maj3(a1, b1, k0);
    const uint32_t x2 = a2 ^ b2 ^ k1, x3 = maj3(a2, b2, k1);
    const uint32_t cc = maj3(x0, e, f);
    s0 = x0 ^ e ^ f;
    s1 = x1 ^ cc;
    ge4 = x2 | x3 | (x1 & cc);
}

__device__ __forceinline__ uint32_t col_update_mask(int j, int W)
{
    // bits of word j whose pixel x satisfies 2 <= x < W-2
    const int lo = j * 32;
    uint32_t m = 0xffffffffu;
    if (lo < 2) m &= ~((1u << (2 - lo)) - 1u);
    const int hi = W - 2 - lo;
    if (hi <= 0) return 0u;
    if (hi < 32) m &= (1u << hi) - 1u;
    return m;
}

__global__ void __launch_bounds__(DT) k_mask_denoise(const DenoiseParams p)
{
    __shared__ uint32_t sR[2][LR][LW];
    __shared__ uint32_t sM[LR][LW];
    __shared__ uint32_t sH[3][LR][TWW];
    __shared__ uint32_t sV[3][TR][TWW];

    const int tid = threadIdx.x;
    const unsigned nblocks = gridDim.x;
    const int64_t words_page = (int64_t)p.H * p.Ww;
    const int64_t total_words = words_page * p.N;

    // ---------------- phase 0: pack bytes -> bits (one thread per 32-pixel word; 32-bit index arithmetic when it fits)
    auto pack_word = [&](int page, int y, int j, int64_t g) {
        const uint8_t *row = p.mask + (int64_t)page * p.stride + (int64_t)y * p.pitch + 32 * j;
        const int valid = min(32, p.W - 32 * j);
        uint32_t word = 0;
        if (valid == 32 && (((uintptr_t)row) & 15) == 0) {
            const uint4 a = *reinterpret_cast<const uint4 *>(row), b = *reinterpret_cast<const uint4 *>(row + 16);
            const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int k = 0; k < 8; k++) word |= ((((w[k] & 0x01010101u) * 0x01020408u) >> 24) & 0xfu) << (4 * k);
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (4 * k < valid) {
                    const uint32_t w = *reinterpret_cast<const uint32_t *>(row + 4 * k);
                    const uint32_t nib = (((w & 0x01010101u) * 0x01020408u) >> 24) & 0xfu;
                    word |= nib << (4 * k);
                }
            }
            if (valid < 32) word &= (1u << valid) - 1u;
        }
        p.Mb[g] = word;
        p.Rb[g] = word;
    };
    if (total_words < 0x7fffffffll) {
        const uint32_t wp = (uint32_t)words_page, ww = (uint32_t)p.Ww, tw = (uint32_t)total_words;
        for (uint32_t g = blockIdx.x * DT + tid; g < tw; g += nblocks * DT) {
            const uint32_t page = g / wp, rem = g - page * wp, y = rem / ww;
            pack_word((int)page, (int)y, (int)(rem - y * ww), (int64_t)g);
        }
    } else {
        for (int64_t g = (int64_t)blockIdx.x * DT + tid; g < total_words; g += (int64_t)nblocks * DT) {
            const int page = (int)(g / words_page);
            const int64_t rem = g - (int64_t)page * words_page;
            const int y = (int)(rem / p.Ww);
            pack_word(page, y, (int)(rem - (int64_t)y * p.Ww), g);
        }
    }
    grid_barrier(p.bar, nblocks);

    // ---------------- phase 1: chaotic iteration to the global fixed point
    const int tiles_page = p.tiles_x * p.tiles_y;
    const int n_tiles = tiles_page * p.N;
    for (int it = 1; it <= p.max_iters; it++) {
        if (blockIdx.x == 0 && tid == 0) atomicExch(&p.changed[(it + 1) % 3], 0u);
        const uint8_t *fl_prev = p.flags + (size_t)((it - 1) & 1) * n_tiles;
        uint8_t *fl_cur = p.flags + (size_t)(it & 1) * n_tiles;
        for (int t = blockIdx.x; t < n_tiles; t += nblocks) {
            const int page = t / tiles_page;
            const int tt = t - page * tiles_page;
            const int ty = tt / p.tiles_x, tx = tt - ty * p.tiles_x;
            bool run = (it == 1);
            if (!run) {
                for (int dy = -1; dy <= 1 && !run; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        const int ny = ty + dy, nx = tx + dx;
                        if (ny >= 0 && ny < p.tiles_y && nx >= 0 && nx < p.tiles_x &&
                            __ldcg(fl_prev + page * tiles_page + ny * p.tiles_x + nx)) { run = true; break; }
                    }
            }
            if (!run) {                        // uniform across the CTA
                if (tid == 0) fl_cur[t] = 0;
                continue;
            }
            const int y0 = ty * TR, wj0 = tx * TWW;
            const uint32_t *Mb = p.Mb + (int64_t)page * words_page;
            uint32_t *Rb = p.Rb + (int64_t)page * words_page;
            __syncthreads();                   // previous tile's smem reads are done
            for (int idx = tid; idx < LR * LW; idx += DT) {
                const int ly = idx / LW, lj = idx - ly * LW;
                const int y = y0 - 2 + ly, j = wj0 - 1 + lj;
                uint32_t r = 0, m = 0;
                if (y >= 0 && y < p.H && j >= 0 && j < p.Ww) {
                    r = __ldcg(Rb + (int64_t)y * p.Ww + j);
                    m = Mb[(int64_t)y * p.Ww + j];
                }
                sR[0][ly][lj] = r; sR[1][ly][lj] = r; sM[ly][lj] = m;
            }
            __syncthreads();
            // static half: H5 of m0 rows, then V = H5m[y+1] + H5m[y+2] + m[x+1] + m[x+2]
            for (int idx = tid; idx < LR * TWW; idx += DT) {
                const int ly = idx / TWW, jj = idx - ly * TWW;
                uint32_t h0, h1, h2;
                h5(sM[ly][jj], sM[ly][jj + 1], sM[ly][jj + 2], h0, h1, h2);
                sH[0][ly][jj] = h0; sH[1][ly][jj] = h1; sH[2][ly][jj] = h2;
            }
            __syncthreads();
            for (int idx = tid; idx < TR * TWW; idx += DT) {
                const int iy = idx / TWW, jj = idx - iy * TWW, ly = iy + 2;
                const uint32_t c = sM[ly][jj + 1], rw = sM[ly][jj + 2];
                const uint32_t e = (c >> 1) | (rw << 31), f = (c >> 2) | (rw << 30);
                uint32_t v0, v1, vg;
                add33_11(sH[0][ly + 1][jj], sH[1][ly + 1][jj], sH[2][ly + 1][jj],
                         sH[0][ly + 2][jj], sH[1][ly + 2][jj], sH[2][ly + 2][jj], e, f, v0, v1, vg);
                sV[0][iy][jj] = v0; sV[1][iy][jj] = v1; sV[2][iy][jj] = vg;
            }
            __syncthreads();
            int cur = 0;
            bool tile_changed = false;
            for (;;) {
                for (int idx = tid; idx < (TR + 2) * TWW; idx += DT) {       // rows ly = 0 .. TR+1
                    const int ly = idx / TWW, jj = idx - ly * TWW;
                    uint32_t h0, h1, h2;
                    h5(sR[cur][ly][jj], sR[cur][ly][jj + 1], sR[cur][ly][jj + 2], h0, h1, h2);
                    sH[0][ly][jj] = h0; sH[1][ly][jj] = h1; sH[2][ly][jj] = h2;
                }
                __syncthreads();
                int ch = 0;
                for (int idx = tid; idx < TR * TWW; idx += DT) {
                    const int iy = idx / TWW, jj = idx - iy * TWW, ly = iy + 2;
                    const int y = y0 + iy, j = wj0 + jj;
                    const uint32_t c = sR[cur][ly][jj + 1], lw = sR[cur][ly][jj];
                    const uint32_t e = (c << 1) | (lw >> 31), f = (c << 2) | (lw >> 30);
                    uint32_t d0, d1, dg;
                    add33_11(sH[0][ly - 2][jj], sH[1][ly - 2][jj], sH[2][ly - 2][jj],
                             sH[0][ly - 1][jj], sH[1][ly - 1][jj], sH[2][ly - 1][jj], e, f, d0, d1, dg);
                    const uint32_t v0 = sV[0][iy][jj], v1 = sV[1][iy][jj], vg = sV[2][iy][jj];
                    const uint32_t ge4 = dg | vg | maj3(d1, v1, d0 & v0);
                    uint32_t upd = 0;
                    if (y >= 2 && y < p.H - 2 && j < p.Ww) upd = col_update_mask(j, p.W);
                    const uint32_t nw = c & (ge4 | ~upd);
                    sR[cur ^ 1][ly][jj + 1] = nw;
                    ch |= (nw != c);
                }
                const int any = __syncthreads_or(ch);
                cur ^= 1;
                if (!any) break;
                tile_changed = true;
            }
            if (tile_changed) {
                for (int idx = tid; idx < TR * TWW; idx += DT) {
                    const int iy = idx / TWW, jj = idx - iy * TWW;
                    const int y = y0 + iy, j = wj0 + jj;
                    if (y < p.H && j < p.Ww) __stcg(Rb + (int64_t)y * p.Ww + j, sR[cur][iy + 2][jj + 1]);
                }
            }
            if (tid == 0) {
                fl_cur[t] = tile_changed ? 1 : 0;
                if (tile_changed) atomicAdd(&p.changed[it % 3], 1u);
            }
        }
        grid_barrier(p.bar, nblocks);
        const unsigned chg = atomicAdd(&p.changed[it % 3], 0u);
        if (chg == 0) break;
    }

    // ---------------- phase 2: write the cleared pixels back to the byte mask
    for (int64_t g = (int64_t)blockIdx.x * DT + tid; g < total_words; g += (int64_t)nblocks * DT) {
        uint32_t diff = p.Mb[g] ^ __ldcg(p.Rb + g);
        if (!diff) continue;
        const int page = (int)(g / words_page);
        const int64_t rem = g - (int64_t)page * words_page;
        const int y = (int)(rem / p.Ww), j = (int)(rem - (int64_t)y * p.Ww);
        uint8_t *row = p.mask + (int64_t)page * p.stride + (int64_t)y * p.pitch + 32 * j;
        while (diff) {
            const int b = __ffs(diff) - 1;
            diff &= diff - 1;
            row[b] = 0;
        }
    }
}

// ---- any (mincnt, n_size): the same fixed point, byte planes, one thread per pixel.  The reference only ever calls
// (4, 2) (mrc.py:388), which is the bit-sliced kernel above; this form completes the Cython signature
// fast_mask_denoise(mask, w, h, mincnt, n_size) (optimiser.pyx:436).  r starts as the mask itself and pixels only ever go
// 1 -> 0, decided from counts that are never below the true ones (stale neighbours are still set), so updating in place
// in any order converges to the sequential result; a pass that clears nothing ends the iteration (host loop).
__global__ void __launch_bounds__(256) k_mask_denoise_general(uint8_t *mask, int64_t pitch, int64_t stride, const uint8_t *m0,
                                                              int W, int H, int mincnt, int n, unsigned *changed)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), page = blockIdx.z;
    if (x < n || x >= W - n || y < n || y >= H - n) return;
    uint8_t *r = mask + (int64_t)page * stride;
    const uint8_t *m = m0 + (int64_t)page * H * W;
    if (!r[(int64_t)y * pitch + x]) return;
    int cnt = 0;                                             // earlier neighbours: their current value; later ones: the input
    for (int dy = -n; dy <= n; dy++)
        for (int dx = -n; dx <= n; dx++) {
            if (dy < 0 || (dy == 0 && dx < 0)) cnt += r[(int64_t)(y + dy) * pitch + x + dx] ? 1 : 0;
            else if (dy > 0 || dx > 0) cnt += m[(int64_t)(y + dy) * W + x + dx] ? 1 : 0;
        }
    if (cnt < mincnt) { r[(int64_t)y * pitch + x] = 0; *changed = 1u; }
}

struct DenoiseLayout {
    int Ww, tiles_x, tiles_y;
    size_t off_Mb, off_Rb, off_flags, off_sync, total;
};

DenoiseLayout denoise_layout(int W, int H, int N)
{
    DenoiseLayout L;
    L.Ww = cdiv(W, 32);
    L.tiles_x = cdiv(L.Ww, TWW);
    L.tiles_y = cdiv(H, TR);
    Carver c;
    const size_t words = (size_t)H * L.Ww * N;
    L.off_Mb = c.take<uint32_t>(words);
    L.off_Rb = c.take<uint32_t>(words);
    L.off_flags = c.take<uint8_t>(2 * (size_t)L.tiles_x * L.tiles_y * N);
    L.off_sync = c.take<unsigned>(8);
    L.total = c.used();
    return L;
}

}  // namespace

size_t denoise_workspace_bytes(int W, int H, int N) { return denoise_layout(W, H, N).total; }

int launch_denoise(uint8_t *mask, int64_t pitch, int64_t stride, int W, int H, int N,
                   void *workspace, size_t workspace_bytes, cudaStream_t st)
{
    const DenoiseLayout L = denoise_layout(W, H, N);
    if (!workspace || workspace_bytes < L.total) return B200MRC_ERR_WORKSPACE;
    uint8_t *ws = (uint8_t *)workspace;
    DenoiseParams p;
    p.mask = mask; p.pitch = pitch; p.stride = stride; p.W = W; p.H = H; p.N = N; p.Ww = L.Ww;
    p.Mb = (uint32_t *)(ws + L.off_Mb); p.Rb = (uint32_t *)(ws + L.off_Rb);
    p.flags = ws + L.off_flags;
    p.tiles_x = L.tiles_x; p.tiles_y = L.tiles_y;
    p.bar = (unsigned *)(ws + L.off_sync);
    p.changed = p.bar + 2;
    p.max_iters = 1 << 30;
    B200MRC_CUDA_TRY(cudaMemsetAsync(ws + L.off_sync, 0, 8 * sizeof(unsigned), st));

    int per_sm = 0;
    B200MRC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mask_denoise, DT, 0));
    if (per_sm < 1) return B200MRC_ERR_UNSUPPORTED;
    const int n_tiles = L.tiles_x * L.tiles_y * N;
    int grid = dev_info().sm_count * per_sm;
    if (grid > n_tiles) grid = n_tiles;
    if (grid < 1) grid = 1;
    void *args[] = {(void *)&p};
    { ProfScope _ps("k_mask_denoise", st); B200MRC_CUDA_TRY(cudaLaunchCooperativeKernel((const void *)k_mask_denoise, dim3(grid), dim3(DT), args, 0, st)); }
    count_launch();
    return B200MRC_OK;
}

}  // namespace b200mrc

using namespace b200mrc;

extern "C" size_t b200mrc_denoise_workspace_bytes(int width, int height, int n_pages)
{
    if (width <= 0 || height <= 0 || n_pages <= 0) return 0;
    return denoise_workspace_bytes(width, height, n_pages);
}

extern "C" int b200mrc_denoise(uint8_t *mask, int64_t pitch, int64_t page_stride,
                               int width, int height, int n_pages, int mincnt, int n_size,
                               void *workspace, size_t workspace_bytes, void *stream)
{
    if (!mask || width <= 0 || height <= 0 || n_pages <= 0) return B200MRC_ERR_INVALID;
    if (mincnt != 4 || n_size != 2) {
        // general form: stream-ordered temporaries, one launch per pass, a host check of the "changed" word after each
        if (n_size < 0 || n_size > 64 || pitch < width || n_pages > 65535 || (height + 7) / 8 > 65535) return B200MRC_ERR_UNSUPPORTED;
        if (width <= 2 * n_size || height <= 2 * n_size) return B200MRC_OK;          // no interior pixel
        cudaStream_t st = (cudaStream_t)stream;
        uint8_t *m0 = nullptr;
        unsigned *chg = nullptr;
        const size_t page_bytes = (size_t)width * height;
        B200MRC_CUDA_TRY(cudaMallocAsync((void **)&m0, page_bytes * n_pages + 256, st));
        chg = (unsigned *)(m0 + ((page_bytes * n_pages + 15) & ~(size_t)15));
        int rc = B200MRC_OK;
        cudaError_t e = cudaMemcpy2DAsync(m0, width, mask, pitch, width, (size_t)height, cudaMemcpyDeviceToDevice, st);
        for (int pg = 1; pg < n_pages && e == cudaSuccess; pg++)
            e = cudaMemcpy2DAsync(m0 + pg * page_bytes, width, mask + (int64_t)pg * page_stride, pitch, width, (size_t)height, cudaMemcpyDeviceToDevice, st);
        const dim3 grid(cdiv(width, 32), cdiv(height, 8), n_pages);
        while (e == cudaSuccess) {
            unsigned h = 0;
            e = cudaMemsetAsync(chg, 0, sizeof(unsigned), st);
            if (e != cudaSuccess) break;
            { ProfScope _ps("k_mask_denoise_general", st);
              k_mask_denoise_general<<<grid, 256, 0, st>>>(mask, pitch, page_stride, m0, width, height, mincnt, n_size, chg); }
            count_launch();
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaMemcpyAsync(&h, chg, sizeof(unsigned), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess || !h) break;
        }
        if (e != cudaSuccess) rc = (int)e;
        cudaFreeAsync(m0, st);
        return rc;
    }
    if ((pitch & 3) || ((uintptr_t)mask & 3) || (page_stride & 3) || pitch < width) return B200MRC_ERR_ALIGNMENT;
    return launch_denoise(mask, pitch, page_stride, width, height, n_pages, workspace, workspace_bytes,
                          (cudaStream_t)stream);
}
