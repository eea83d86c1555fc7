// api.cu -- library bookkeeping and b200mrc_decompose: the whole of create_mrc_hocr_components
// (internetarchivepdf/mrc.py:334-471, hocr_word_data == []) enqueued on one stream for a
// device-resident batch of pages:
//     gray (mrc.py:358-363) -> estimate_noise (:305) -> conditional Gaussian blur (:309-311)
//     -> Sauvola (:325) -> fast_mask_denoise (:388)            => yield mask   (:399)
//     -> optimise fg n=3 (:412-415) [-> thumbnail (:420-434)]  => yield fg     (:436)
//     -> optimise bg n=10 (:446-449) [-> thumbnail (:454-468)] => yield bg     (:470)
// Nothing here synchronises with the host; per-page decisions (blur or not, radius) are taken
// on the device from the sigma array.
#include "common.cuh"
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

struct b200mrc_resample_plan;

namespace b200mrc {

std::atomic<uint64_t> g_launch_count{0};

namespace {
struct ProfRec { const char *name; cudaEvent_t a, b; };
std::vector<ProfRec> g_prof;
bool g_prof_on = false;
}  // namespace

void prof_begin(const char *kernel, cudaStream_t st)
{
    if (!g_prof_on) return;
    ProfRec r{kernel, nullptr, nullptr};
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    g_prof.push_back(r);
}
void prof_end(cudaStream_t st)
{
    if (!g_prof_on || g_prof.empty()) return;
    cudaEventRecord(g_prof.back().b, st);
}

const DevInfo &dev_info()
{
    static DevInfo info[64];
    static bool have[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!have[dev]) {
        cudaDeviceGetAttribute(&info[dev].sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&info[dev].max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        have[dev] = true;
    }
    return info[dev];
}

// implemented in the kernel translation units
size_t noise_workspace_bytes(int W, int H, int N);
int launch_estimate_noise(const uint8_t *in, int64_t in_pitch, int64_t in_stride, int C, int W, int H, int N,
                          double *sigma_out, void *workspace, size_t workspace_bytes, cudaStream_t st);
int launch_gray_blur(const uint8_t *in, int64_t in_pitch, int64_t in_stride, int C,
                     uint8_t *out, int64_t out_pitch, int64_t out_stride,
                     int W, int H, int N, const double *sigma, int *err_flag, cudaStream_t st);
size_t denoise_workspace_bytes(int W, int H, int N);
int launch_denoise(uint8_t *mask, int64_t pitch, int64_t stride, int W, int H, int N,
                   void *workspace, size_t workspace_bytes, cudaStream_t st);
size_t optimise_workspace_bytes(int W, int H, int N);
int launch_optimise(const uint8_t *mask, int64_t mpitch, int64_t mstride,
                    const uint8_t *img, int64_t ipitch, int64_t istride, int C,
                    uint8_t *ofg, int64_t fpitch, int64_t fstride, int nfg,
                    uint8_t *obg, int64_t bpitch, int64_t bstride, int nbg,
                    int W, int H, int N, void *workspace, size_t workspace_bytes, cudaStream_t st);

namespace {

struct DecomposeLayout {
    size_t gray_pitch, gray_page;
    size_t full_pitch, full_page;          // full-resolution fg / bg planes when downsampled
    size_t off_sigma, off_gray, off_fgfull, off_bgfull, off_scratch, scratch_bytes, total;
};

DecomposeLayout decompose_layout(const b200mrc_decompose_args *a)
{
    DecomposeLayout L;
    memset(&L, 0, sizeof(L));
    const int W = a->width, H = a->height, N = a->n_pages, C = a->channels;
    const bool mask_only = (a->flags & B200MRC_DECOMPOSE_MASK_ONLY) != 0;
    L.gray_pitch = align_up((size_t)W, 16); L.gray_page = L.gray_pitch * H;
    L.full_pitch = align_up((size_t)W * C, 16); L.full_page = L.full_pitch * H;
    Carver c;
    L.off_sigma = c.take<double>(N);
    L.off_gray = c.take<uint8_t>(L.gray_page * N);
    if (!mask_only && a->fg_plan) L.off_fgfull = c.take<uint8_t>(L.full_page * N);
    if (!mask_only && a->bg_plan) L.off_bgfull = c.take<uint8_t>(L.full_page * N);
    // stages run one after another on the stream: they share one scratch region
    size_t s = noise_workspace_bytes(W, H, N);
    if (a->flags & B200MRC_DECOMPOSE_DENOISE_FAST) s = std::max(s, denoise_workspace_bytes(W, H, N));
    if (!mask_only) {
        s = std::max(s, optimise_workspace_bytes(W, H, N));
        if (a->fg_plan) s = std::max(s, b200mrc_resample_workspace_bytes(a->fg_plan, N));
        if (a->bg_plan) s = std::max(s, b200mrc_resample_workspace_bytes(a->bg_plan, N));
    }
    L.scratch_bytes = s;
    L.off_scratch = c.take<uint8_t>(s);
    L.total = c.used();
    return L;
}

int check_args(const b200mrc_decompose_args *a)
{
    if (!a || !a->img || !a->mask) return B200MRC_ERR_INVALID;
    if (a->width <= 0 || a->height <= 0 || a->n_pages <= 0) return B200MRC_ERR_INVALID;
    if (a->channels != 1 && a->channels != 3) return B200MRC_ERR_UNSUPPORTED;
    if (a->n_pages > 65535) return B200MRC_ERR_UNSUPPORTED;
    if (a->window < 1 || a->window > B200MRC_MAX_WINDOW) return B200MRC_ERR_UNSUPPORTED;
    if (!(a->flags & B200MRC_DECOMPOSE_MASK_ONLY) && (!a->fg || !a->bg)) return B200MRC_ERR_INVALID;
    if ((a->mask_pitch & 3) || ((uintptr_t)a->mask & 3) || (a->mask_page_stride & 3) || a->mask_pitch < a->width)
        return B200MRC_ERR_ALIGNMENT;
    return B200MRC_OK;
}

}  // namespace
}  // namespace b200mrc

using namespace b200mrc;

extern "C" int b200mrc_version(void) { return B200MRC_VERSION; }

extern "C" int b200mrc_profile_enable(int on)
{
    for (auto &r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    g_prof_on = on != 0;
    return B200MRC_OK;
}

extern "C" int b200mrc_profile_report(char *buf, size_t cap)
{
    // "kernel,launches,total_ms" lines; synchronises with every recorded event
    std::map<std::string, std::pair<int, double>> acc;
    std::vector<std::string> order;
    for (auto &r : g_prof) {
        if (cudaEventSynchronize(r.b) != cudaSuccess) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
        if (!acc.count(r.name)) order.push_back(r.name);
        acc[r.name].first += 1; acc[r.name].second += ms;
    }
    std::string out;
    for (auto &n : order) {
        char line[256];
        snprintf(line, sizeof(line), "%s,%d,%.6f\n", n.c_str(), acc[n].first, acc[n].second);
        out += line;
    }
    if (buf && cap) { strncpy(buf, out.c_str(), cap - 1); buf[cap - 1] = 0; }
    return (int)out.size();
}

extern "C" uint64_t b200mrc_launch_count(void) { return g_launch_count.load(); }

extern "C" const char *b200mrc_error_string(int status)
{
    switch (status) {
    case B200MRC_OK: return "ok";
    case B200MRC_ERR_INVALID: return "b200mrc: invalid argument";
    case B200MRC_ERR_UNSUPPORTED: return "b200mrc: parameter outside the implemented range";
    case B200MRC_ERR_WORKSPACE: return "b200mrc: workspace too small";
    case B200MRC_ERR_ALIGNMENT: return "b200mrc: pointer/pitch alignment requirement violated";
    default: break;
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "b200mrc: unknown error";
}

extern "C" int b200mrc_copy2d(void *dst, int64_t dst_pitch, const void *src, int64_t src_pitch,
                              int64_t row_bytes, int64_t rows, int kind, void *stream)
{
    if (!dst || !src || row_bytes < 0 || rows < 0 || dst_pitch < row_bytes || src_pitch < row_bytes) return B200MRC_ERR_INVALID;
    if (kind < B200MRC_COPY_H2D || kind > B200MRC_COPY_D2D) return B200MRC_ERR_INVALID;
    if (row_bytes == 0 || rows == 0) return B200MRC_OK;
    const cudaMemcpyKind k = kind == B200MRC_COPY_H2D ? cudaMemcpyHostToDevice
                           : (kind == B200MRC_COPY_D2H ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
    B200MRC_CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)row_bytes, (size_t)rows, k,
                                       (cudaStream_t)stream));
    return B200MRC_OK;
}

extern "C" size_t b200mrc_decompose_workspace_bytes(const b200mrc_decompose_args *a)
{
    if (check_args(a) != B200MRC_OK) return 0;
    return decompose_layout(a).total;
}

extern "C" int b200mrc_decompose(const b200mrc_decompose_args *a, void *stream)
{
    int rc = check_args(a);
    if (rc != B200MRC_OK) return rc;
    const DecomposeLayout L = decompose_layout(a);
    if (!a->workspace || a->workspace_bytes < L.total) return B200MRC_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *ws = (uint8_t *)a->workspace;
    const int W = a->width, H = a->height, N = a->n_pages, C = a->channels;
    double *sigma = (double *)(ws + L.off_sigma);
    uint8_t *gray = ws + L.off_gray;
    void *scratch = ws + L.off_scratch;

    // ---- noise estimate (or injected sigma)
    const double *sigma_used = sigma;
    if (a->flags & B200MRC_DECOMPOSE_NO_NOISE_EST) {
        if (a->sigma_in) B200MRC_CUDA_TRY(cudaMemcpyAsync(sigma, a->sigma_in, sizeof(double) * N, cudaMemcpyDeviceToDevice, st));
        else sigma_used = nullptr;
    } else {
        rc = launch_estimate_noise(a->img, a->img_pitch, a->img_page_stride, C, W, H, N, sigma, scratch, L.scratch_bytes, st);
        if (rc) return rc;
    }
    if (a->sigma_out) {
        if (sigma_used) B200MRC_CUDA_TRY(cudaMemcpyAsync(a->sigma_out, sigma, sizeof(double) * N, cudaMemcpyDeviceToDevice, st));
        else B200MRC_CUDA_TRY(cudaMemsetAsync(a->sigma_out, 0, sizeof(double) * N, st));
    }
    // ---- gray (+ blur where sigma > 1)
    rc = launch_gray_blur(a->img, a->img_pitch, a->img_page_stride, C, gray, (int64_t)L.gray_pitch, (int64_t)L.gray_page,
                          W, H, N, sigma_used, nullptr, st);
    if (rc) return rc;
    // ---- Sauvola: mask = threshold_image(gray), or mask |= ... when the caller filled in the hOCR line masks
    rc = b200mrc_sauvola(gray, (int64_t)L.gray_pitch, (int64_t)L.gray_page, a->mask, a->mask_pitch, a->mask_page_stride,
                         W, H, N, a->window, a->window, a->k, a->R,
                         (a->flags & B200MRC_DECOMPOSE_OR_INTO_MASK) ? B200MRC_SAUVOLA_OR_INTO : 0, st);
    if (rc) return rc;
    // ---- denoise
    if (a->flags & B200MRC_DECOMPOSE_DENOISE_FAST) {
        rc = launch_denoise(a->mask, a->mask_pitch, a->mask_page_stride, W, H, N, scratch, L.scratch_bytes, st);
        if (rc) return rc;
    }
    if (a->flags & B200MRC_DECOMPOSE_MASK_ONLY) return B200MRC_OK;

    // ---- fg / bg
    uint8_t *fg_full = a->fg_plan ? ws + L.off_fgfull : a->fg;
    uint8_t *bg_full = a->bg_plan ? ws + L.off_bgfull : a->bg;
    const int64_t fgp = a->fg_plan ? (int64_t)L.full_pitch : a->fg_pitch, fgs = a->fg_plan ? (int64_t)L.full_page : a->fg_page_stride;
    const int64_t bgp = a->bg_plan ? (int64_t)L.full_pitch : a->bg_pitch, bgs = a->bg_plan ? (int64_t)L.full_page : a->bg_page_stride;
    rc = launch_optimise(a->mask, a->mask_pitch, a->mask_page_stride, a->img, a->img_pitch, a->img_page_stride, C,
                         fg_full, fgp, fgs, 3, bg_full, bgp, bgs, 10, W, H, N, scratch, L.scratch_bytes, st);
    if (rc) return rc;
    if (a->fg_plan) {
        rc = b200mrc_resample(a->fg_plan, fg_full, fgp, fgs, a->fg, a->fg_pitch, a->fg_page_stride, N, scratch, L.scratch_bytes, st);
        if (rc) return rc;
    }
    if (a->bg_plan) {
        rc = b200mrc_resample(a->bg_plan, bg_full, bgp, bgs, a->bg, a->bg_pitch, a->bg_page_stride, N, scratch, L.scratch_bytes, st);
        if (rc) return rc;
    }
    return B200MRC_OK;
}
