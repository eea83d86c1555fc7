// resample.cu -- k_reduce_box / k_resample_h / k_resample_v: PIL Image.thumbnail at
// internetarchivepdf/mrc.py:420-434 (fg) and :454-468 (bg).
//
// Pillow 12.2 semantics restated (oracle/mrc_oracle.c orc_resample / orc_reduce and
// oracle/oracle.py thumbnail_plan, pinned bit-exact against the installed Pillow):
//   Image.thumbnail(size): aspect-preserving target size; Image.resize(reducing_gap=2.0): box
//   `reduce` by int(scale/2) per axis when > 1 (ImagingReduce: ((sum + cells/2) * mult) >> 24),
//   then ImagingResample: per-output coefficient windows (precompute_coeffs, BICUBIC a=-0.5,
//   support 2*max(scale,1)) normalised in double, quantised to 22-bit fixed point
//   (normalize_coeffs_8bpc), horizontal pass then vertical pass with a uint8 intermediate,
//   out = clip8((2^21 + sum p*k) >> 22).
// The size logic and the coefficient tables are computed on the host once per plan (they depend
// only on the shapes); the kernels are pure gather-multiply-accumulate byte kernels.
#include "common.cuh"
#include <cmath>
#include <vector>

struct b200mrc_resample_plan {
    int W, H, C;
    int fx, fy;                 // box reduce factors (1 = none)
    int rx0, ry0, rx1, ry1;     // reduce box on the source
    int SW, SH;                 // size after reduce (== W,H when no reduce)
    float bx0, by0, bx1, by1;   // resample box on the (reduced) source
    int OW, OH, filter;
    int need_h, need_v, ksize_h, ksize_v;
    int *d_bounds_h, *d_kk_h, *d_bounds_v, *d_kk_v;
};

namespace b200mrc {
namespace {

double bicubic_filter(double x)
{
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

double lanczos_filter(double x)
{
    if (x < 0.0) x = -x;
    if (x >= 3.0) return 0.0;
    if (x == 0.0) return 1.0;
    const double a = x * 3.14159265358979323846, b = a / 3.0;
    return (std::sin(a) / a) * (std::sin(b) / b);
}

// Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc
int build_coeffs(int inSize, float in0, float in1, int outSize, int filter, std::vector<int> &bounds, std::vector<int> &kk)
{
    const double support0 = filter == 1 ? 3.0 : 2.0;
    double scale, filterscale;
    filterscale = scale = (double)(in1 - in0) / outSize;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = support0 * filterscale;
    const int ksize = (int)std::ceil(support) * 2 + 1;
    bounds.assign((size_t)outSize * 2, 0);
    kk.assign((size_t)outSize * ksize, 0);
    std::vector<double> k(ksize);
    for (int xx = 0; xx < outSize; xx++) {
        const double center = in0 + (xx + 0.5) * scale;
        double ww = 0.0;
        const double ss = 1.0 / filterscale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > inSize) xmax = inSize;
        xmax -= xmin;
        int x;
        for (x = 0; x < xmax; x++) {
            const double arg = (x + xmin - center + 0.5) * ss;
            const double w = filter == 1 ? lanczos_filter(arg) : bicubic_filter(arg);
            k[x] = w; ww += w;
        }
        for (x = 0; x < xmax; x++) if (ww != 0.0) k[x] /= ww;
        for (; x < ksize; x++) k[x] = 0;
        bounds[2 * xx] = xmin; bounds[2 * xx + 1] = xmax;
        for (x = 0; x < ksize; x++)
            kk[(size_t)xx * ksize + x] = k[x] < 0 ? (int)(-0.5 + k[x] * (1 << 22)) : (int)(0.5 + k[x] * (1 << 22));
    }
    return ksize;
}

struct ResampleParams {
    const uint8_t *in; int64_t in_pitch, in_stride;
    uint8_t *out; int64_t out_pitch, out_stride;
    int C, in_w, in_h, out_w, out_h, ksize;
    const int *bounds, *kk;
};

__device__ __forceinline__ uint8_t clip8(int v)
{
    v >>= 22;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// out[y][xx][c] = clip8(2^21 + sum_k in[y][xmin+k][c] * kk[xx][k]).  One thread per output PIXEL
// (all channels share the coefficient fetch and the address arithmetic) and RPT consecutive rows
// (the coefficient window is loaded once per thread).
constexpr int H_RPT = 4;
template <int C>
__global__ void __launch_bounds__(128) k_resample_h(const ResampleParams p)
{
    const int xx = blockIdx.x * 128 + threadIdx.x;
    const int y0 = blockIdx.y * H_RPT, page = blockIdx.z;
    if (xx >= p.out_w) return;
    const int xmin = p.bounds[2 * xx], cnt = p.bounds[2 * xx + 1];
    const int *k = p.kk + (size_t)xx * p.ksize;
    const uint8_t *src = p.in + (int64_t)page * p.in_stride + (int64_t)y0 * p.in_pitch + (int64_t)xmin * C;
    uint8_t *dst = p.out + (int64_t)page * p.out_stride + (int64_t)y0 * p.out_pitch + (int64_t)xx * C;
    const int rows = min(H_RPT, p.in_h - y0);
    int acc[H_RPT][C];
#pragma unroll
    for (int r = 0; r < H_RPT; r++)
#pragma unroll
        for (int c = 0; c < C; c++) acc[r][c] = 1 << 21;
    for (int i = 0; i < cnt; i++) {
        const int kv = __ldg(k + i);
#pragma unroll
        for (int r = 0; r < H_RPT; r++) {
            if (r < rows) {
                const uint8_t *q = src + (int64_t)r * p.in_pitch + i * C;
#pragma unroll
                for (int c = 0; c < C; c++) acc[r][c] += (int)q[c] * kv;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < H_RPT; r++)
        if (r < rows) {
#pragma unroll
            for (int c = 0; c < C; c++) dst[(int64_t)r * p.out_pitch + c] = clip8(acc[r][c]);
        }
}

// out[yy][b] = clip8(2^21 + sum_k in[ymin+k][b] * kk[yy][k]); one thread per 4 consecutive bytes
// (rows of both planes are 16-byte aligned workspace / pitched planes)
__global__ void __launch_bounds__(256) k_resample_v(const ResampleParams p)
{
    const int b4 = (blockIdx.x * 256 + threadIdx.x) * 4;
    const int yy = blockIdx.y, page = blockIdx.z;
    const int nb = p.out_w * p.C;
    if (b4 >= nb) return;
    const int ymin = p.bounds[2 * yy], cnt = p.bounds[2 * yy + 1];
    const int *k = p.kk + (size_t)yy * p.ksize;
    const uint8_t *src = p.in + (int64_t)page * p.in_stride + (int64_t)ymin * p.in_pitch + b4;
    int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21, s3 = 1 << 21;
    for (int i = 0; i < cnt; i++) {
        const int kv = __ldg(k + i);
        const uint32_t w = *reinterpret_cast<const uint32_t *>(src + (int64_t)i * p.in_pitch);
        s0 += (int)(w & 0xffu) * kv; s1 += (int)((w >> 8) & 0xffu) * kv;
        s2 += (int)((w >> 16) & 0xffu) * kv; s3 += (int)(w >> 24) * kv;
    }
    uint8_t *dst = p.out + (int64_t)page * p.out_stride + (int64_t)yy * p.out_pitch + b4;
    const uint32_t o = (uint32_t)clip8(s0) | ((uint32_t)clip8(s1) << 8) | ((uint32_t)clip8(s2) << 16) | ((uint32_t)clip8(s3) << 24);
    if (b4 + 3 < nb && (((uintptr_t)dst) & 3) == 0) *reinterpret_cast<uint32_t *>(dst) = o;
    else
        for (int j = 0; j < 4 && b4 + j < nb; j++) dst[j] = (uint8_t)(o >> (8 * j));
}

struct ReduceParams {
    const uint8_t *in; int64_t in_pitch, in_stride;
    uint8_t *out; int64_t out_pitch, out_stride;
    int C, bx, by, bw, bh, fx, fy, out_w, out_h;
};

__device__ __forceinline__ uint32_t division_u32(int divider, int result_bits)
{
    // Pillow Reduce.c division_UINT32
    const uint32_t max_dividend = (1u << result_bits) * (uint32_t)divider;
    const float max_int = (1 << 30) * 4.0f;
    return (uint32_t)(max_int / max_dividend);
}

__global__ void __launch_bounds__(256) k_reduce_box(const ReduceParams p)
{
    const int b = blockIdx.x * 256 + threadIdx.x;
    const int oy = blockIdx.y, page = blockIdx.z;
    if (b >= p.out_w * p.C) return;
    const int ox = b / p.C, c = b - ox * p.C;
    const int x0 = ox * p.fx, x1 = min(p.bw, x0 + p.fx), y0 = oy * p.fy, y1 = min(p.bh, y0 + p.fy);
    const int cells = (y1 - y0) * (x1 - x0);
    const uint32_t mult = division_u32(cells, 8);
    uint32_t ss = (uint32_t)cells / 2;
    const uint8_t *src = p.in + (int64_t)page * p.in_stride;
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++)
            ss += src[(int64_t)(p.by + y) * p.in_pitch + (int64_t)(p.bx + x) * p.C + c];
    p.out[(int64_t)page * p.out_stride + (int64_t)oy * p.out_pitch + b] = (uint8_t)((ss * mult) >> 24);
}

int upload(const std::vector<int> &v, int **d)
{
    *d = nullptr;
    if (v.empty()) return 0;
    cudaError_t e = cudaMalloc((void **)d, v.size() * sizeof(int));
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpy(*d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice);
    return (int)e;
}

}  // namespace
}  // namespace b200mrc

using namespace b200mrc;

extern "C" b200mrc_resample_plan *b200mrc_thumbnail_plan_create(int width, int height, int channels,
                                                                double req_width, double req_height,
                                                                double reducing_gap, int filter, int *status)
{
    int dummy;
    if (!status) status = &dummy;
    *status = B200MRC_OK;
    if (width <= 0 || height <= 0 || (channels != 1 && channels != 3) || (filter != 0 && filter != 1)) {
        *status = B200MRC_ERR_INVALID; return nullptr;
    }
    // PIL Image.thumbnail (Image.py): provided_size = floor(size); preserve_aspect_ratio()
    const double px = std::floor(req_width), py = std::floor(req_height);
    if (!(px >= 1) || !(py >= 1)) { *status = B200MRC_ERR_INVALID; return nullptr; }
    if (px >= width && py >= height) return nullptr;                 // nothing to do
    if ((double)height > (double)width * 100) { *status = B200MRC_ERR_UNSUPPORTED; return nullptr; }  // Pillow's tall-image 2-step path
    auto round_aspect = [](double number, auto key) {
        const double f = std::floor(number), c = std::ceil(number);
        // python: max(min(floor, ceil, key=key), 1) -- min returns the first on ties
        double best = key(f) <= key(c) ? f : c;
        return best < 1 ? 1.0 : best;
    };
    const double aspect = (double)width / height;
    double x = px, y = py;
    if (x / y >= aspect) x = round_aspect(y * aspect, [&](double n) { return std::fabs(aspect - n / y); });
    else y = round_aspect(x / aspect, [&](double n) { return n == 0 ? 0.0 : std::fabs(aspect - x / n); });
    const int OW = (int)x, OH = (int)y;
    if (OW == width && OH == height) return nullptr;

    b200mrc_resample_plan *pl = new b200mrc_resample_plan();
    pl->W = width; pl->H = height; pl->C = channels; pl->OW = OW; pl->OH = OH; pl->filter = filter;
    pl->fx = pl->fy = 1; pl->rx0 = pl->ry0 = 0; pl->rx1 = width; pl->ry1 = height;
    double b0 = 0, b1 = 0, b2 = width, b3 = height;
    if (reducing_gap > 0) {
        // Image.resize (Image.py): factor = int(box_size / size / reducing_gap) or 1
        int fx = (int)((b2 - b0) / OW / reducing_gap); if (fx < 1) fx = 1;
        int fy = (int)((b3 - b1) / OH / reducing_gap); if (fy < 1) fy = 1;
        if (fx > 1 || fy > 1) {
            // _get_safe_box
            const double support = (filter == 1 ? 3.0 : 2.0) - 0.5;
            const double sx = (b2 - b0) / OW * support, sy = (b3 - b1) / OH * support;
            pl->rx0 = std::max(0, (int)(b0 - sx)); pl->ry0 = std::max(0, (int)(b1 - sy));
            pl->rx1 = std::min(width, (int)std::ceil(b2 + sx)); pl->ry1 = std::min(height, (int)std::ceil(b3 + sy));
            pl->fx = fx; pl->fy = fy;
            b0 = (b0 - pl->rx0) / fx; b1 = (b1 - pl->ry0) / fy;
            b2 = (b2 - pl->rx0) / fx; b3 = (b3 - pl->ry0) / fy;
        }
    }
    pl->SW = (pl->rx1 - pl->rx0 + pl->fx - 1) / pl->fx;
    pl->SH = (pl->ry1 - pl->ry0 + pl->fy - 1) / pl->fy;
    if (pl->fx == 1 && pl->fy == 1) { pl->SW = width; pl->SH = height; }
    pl->bx0 = (float)b0; pl->by0 = (float)b1; pl->bx1 = (float)b2; pl->by1 = (float)b3;
    // ImagingResampleInner
    pl->need_h = OW != pl->SW || pl->bx0 != 0 || pl->bx1 != (float)pl->SW;
    pl->need_v = OH != pl->SH || pl->by0 != 0 || pl->by1 != (float)pl->SH;
    std::vector<int> bh, kh, bv, kv;
    pl->ksize_h = build_coeffs(pl->SW, pl->bx0, pl->bx1, OW, filter, bh, kh);
    pl->ksize_v = build_coeffs(pl->SH, pl->by0, pl->by1, OH, filter, bv, kv);
    int rc = upload(bh, &pl->d_bounds_h);
    if (!rc) rc = upload(kh, &pl->d_kk_h);
    if (!rc) rc = upload(bv, &pl->d_bounds_v);
    if (!rc) rc = upload(kv, &pl->d_kk_v);
    if (rc) { *status = rc; b200mrc_resample_plan_destroy(pl); return nullptr; }
    return pl;
}

extern "C" void b200mrc_resample_plan_destroy(b200mrc_resample_plan *pl)
{
    if (!pl) return;
    cudaFree(pl->d_bounds_h); cudaFree(pl->d_kk_h); cudaFree(pl->d_bounds_v); cudaFree(pl->d_kk_v);
    delete pl;
}

extern "C" void b200mrc_resample_plan_out_size(const b200mrc_resample_plan *pl, int *out_width, int *out_height)
{
    if (out_width) *out_width = pl ? pl->OW : 0;
    if (out_height) *out_height = pl ? pl->OH : 0;
}

static size_t resample_red_bytes(const b200mrc_resample_plan *pl)
{
    return (pl->fx > 1 || pl->fy > 1) ? align_up((size_t)pl->SW * pl->C, 16) * pl->SH : 0;
}
static size_t resample_tmp_bytes(const b200mrc_resample_plan *pl)
{
    return (pl->need_h && pl->need_v) ? align_up((size_t)pl->OW * pl->C, 16) * pl->SH : 0;
}

extern "C" size_t b200mrc_resample_workspace_bytes(const b200mrc_resample_plan *pl, int n_pages)
{
    if (!pl || n_pages <= 0) return 0;
    return align_up(resample_red_bytes(pl) * n_pages, 256) + align_up(resample_tmp_bytes(pl) * n_pages, 256) + 256;
}

extern "C" int b200mrc_resample(const b200mrc_resample_plan *pl,
                                const uint8_t *in, int64_t in_pitch, int64_t in_page_stride,
                                uint8_t *out, int64_t out_pitch, int64_t out_page_stride, int n_pages,
                                void *workspace, size_t workspace_bytes, void *stream)
{
    if (!pl || !in || !out || n_pages <= 0) return B200MRC_ERR_INVALID;
    if (n_pages > 65535 || pl->SH > 65535 * H_RPT || pl->OH > 65535) return B200MRC_ERR_UNSUPPORTED;
    if (workspace_bytes < b200mrc_resample_workspace_bytes(pl, n_pages) || (!workspace && workspace_bytes)) return B200MRC_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *ws = (uint8_t *)workspace;
    const size_t red_page = resample_red_bytes(pl), tmp_page = resample_tmp_bytes(pl);
    uint8_t *red = ws, *tmp = ws + align_up(red_page * n_pages, 256);

    const uint8_t *src = in; int64_t src_pitch = in_pitch, src_stride = in_page_stride;
    if (red_page) {
        ReduceParams r;
        r.in = in; r.in_pitch = in_pitch; r.in_stride = in_page_stride;
        r.out = red; r.out_pitch = (int64_t)align_up((size_t)pl->SW * pl->C, 16); r.out_stride = (int64_t)red_page;
        r.C = pl->C; r.bx = pl->rx0; r.by = pl->ry0; r.bw = pl->rx1 - pl->rx0; r.bh = pl->ry1 - pl->ry0;
        r.fx = pl->fx; r.fy = pl->fy; r.out_w = pl->SW; r.out_h = pl->SH;
        dim3 grid(cdiv(pl->SW * pl->C, 256), pl->SH, n_pages);
        { ProfScope _ps("k_reduce_box", st); k_reduce_box<<<grid, 256, 0, st>>>(r); }
        B200MRC_LAUNCH_CHECK();
        src = red; src_pitch = r.out_pitch; src_stride = r.out_stride;
    }
    if (!pl->need_h && !pl->need_v) {
        for (int n = 0; n < n_pages; n++)
            B200MRC_CUDA_TRY(cudaMemcpy2DAsync(out + (int64_t)n * out_page_stride, out_pitch, src + (int64_t)n * src_stride, src_pitch,
                                               (size_t)pl->OW * pl->C, (size_t)pl->OH, cudaMemcpyDeviceToDevice, st));
        return B200MRC_OK;
    }
    if (pl->need_h) {
        ResampleParams h;
        h.in = src; h.in_pitch = src_pitch; h.in_stride = src_stride;
        if (pl->need_v) { h.out = tmp; h.out_pitch = (int64_t)align_up((size_t)pl->OW * pl->C, 16); h.out_stride = (int64_t)tmp_page; }
        else { h.out = out; h.out_pitch = out_pitch; h.out_stride = out_page_stride; }
        h.C = pl->C; h.in_w = pl->SW; h.in_h = pl->SH; h.out_w = pl->OW; h.out_h = pl->SH; h.ksize = pl->ksize_h;
        h.bounds = pl->d_bounds_h; h.kk = pl->d_kk_h;
        dim3 grid(cdiv(pl->OW, 128), cdiv(pl->SH, H_RPT), n_pages);
        { ProfScope _ps("k_resample_h", st);
          if (pl->C == 1) k_resample_h<1><<<grid, 128, 0, st>>>(h);
          else k_resample_h<3><<<grid, 128, 0, st>>>(h); }
        B200MRC_LAUNCH_CHECK();
        src = h.out; src_pitch = h.out_pitch; src_stride = h.out_stride;
    }
    if (pl->need_v) {
        ResampleParams v;
        v.in = src; v.in_pitch = src_pitch; v.in_stride = src_stride;
        v.out = out; v.out_pitch = out_pitch; v.out_stride = out_page_stride;
        v.C = pl->C; v.in_w = pl->OW; v.in_h = pl->SH; v.out_w = pl->OW; v.out_h = pl->OH; v.ksize = pl->ksize_v;
        v.bounds = pl->d_bounds_v; v.kk = pl->d_kk_v;
        if ((src_pitch & 3) || (src_stride & 3) || ((uintptr_t)src & 3)) return B200MRC_ERR_ALIGNMENT;
        dim3 grid(cdiv(cdiv(pl->OW * pl->C, 4), 256), pl->OH, n_pages);
        { ProfScope _ps("k_resample_v", st); k_resample_v<<<grid, 256, 0, st>>>(v); }
        B200MRC_LAUNCH_CHECK();
    }
    return B200MRC_OK;
}
