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
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>

struct b200mrc_resample_plan;

namespace b200mrc {

std::atomic<uint64_t> g_launch_count{0};

namespace {
struct ProfRec { const char *name; cudaEvent_t a, b; };
std::mutex g_prof_mu;                   // guards g_prof (launches may come from several host threads)
std::vector<ProfRec> g_prof;
std::atomic<bool> g_prof_on{false};
}  // namespace

// A ProfScope holds the index of its own record, so scopes of different threads do not mix their end events.
int prof_begin(const char *kernel, cudaStream_t st)
{
    if (!g_prof_on.load(std::memory_order_relaxed)) return -1;
    ProfRec r{kernel, nullptr, nullptr};
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
    return (int)g_prof.size() - 1;
}
void prof_end(int idx, cudaStream_t st)
{
    if (idx < 0) return;
    cudaEvent_t b = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        if (idx < (int)g_prof.size()) b = g_prof[idx].b;
    }
    if (b) cudaEventRecord(b, st);
}

namespace {
struct TuneDef { const char *name; int dflt; const char *word1, *word2; };      // env strings word1 -> 1, word2 -> 2
const TuneDef k_tune_defs[T_COUNT] = {
    {"IIRW_MODE", 0, "single", "trio"}, {"IIRW_TPC", 2, nullptr, nullptr}, {"IIRW_FEED", 0, "tma", "async"},
    {"IIRW_PSLEEP", 300, nullptr, nullptr}, {"IIRW_WPC", 2, nullptr, nullptr}, {"FIRW_BAND", 256, nullptr, nullptr},
    {"FIRW_WPC", 4, nullptr, nullptr}, {"FUSED_NT", 0, nullptr, nullptr}, {"FUSED_BANDS", 0, nullptr, nullptr}, {"FUSED_DBG", 0, nullptr, nullptr}, {"FUSED_OCC", 0, nullptr, nullptr},
    {"THRESHOLD_PATH", 0, "legacy", "fused"}, {"OPT_PATH", 0, "generic", nullptr}, {"NOISE_DIRECT", 0, nullptr, nullptr},
    {"RESAMPLE_2PASS", 0, nullptr, nullptr}, {"TILE_H", 32, nullptr, nullptr}, {"DECOMPOSE_GROUPS", 0, nullptr, nullptr},
    {"DECOMPOSE_STREAMS", 0, nullptr, nullptr},
    {"BG_FOLLOW", 1, nullptr, nullptr},
};
std::atomic<int> g_tune[T_COUNT];
std::once_flag g_tune_once;
void tune_init()
{
    for (int i = 0; i < T_COUNT; i++) {
        const TuneDef &d = k_tune_defs[i];
        int v = d.dflt;
        const std::string env = std::string("B200MRC_") + d.name;
        if (const char *e = getenv(env.c_str())) {
            if (d.word1 && !strcmp(e, d.word1)) v = 1;
            else if (d.word2 && !strcmp(e, d.word2)) v = 2;
            else v = atoi(e);
        }
        g_tune[i].store(v);
    }
}
}  // namespace

int tune(TuneKey k)
{
    std::call_once(g_tune_once, tune_init);
    return g_tune[k].load(std::memory_order_relaxed);
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
                    int W, int H, int N, void *workspace, size_t workspace_bytes, cudaStream_t st, int **bg_progress);
int launch_resample_follow(const b200mrc_resample_plan *pl, const uint8_t *in, int64_t in_pitch, int64_t in_page_stride,
                           uint8_t *out, int64_t out_pitch, int64_t out_page_stride, int n_pages,
                           const int *prog, int prog_s, int prog_w, cudaStream_t st);

namespace {

// A batch is split into page groups that run on internal streams: the row-latency-bound sweep of one group overlaps the
// issue-bound stages of the others (measured: profiles/).  The tuning keys DECOMPOSE_GROUPS /
// DECOMPOSE_STREAMS override the choice.
constexpr int MAX_INTERNAL_STREAMS = 4;
struct GroupPlan { int groups, per_group, streams; };

GroupPlan group_plan(int N)
{
    const int env_groups = tune(T_DECOMPOSE_GROUPS), env_streams = tune(T_DECOMPOSE_STREAMS);
    // measured (profiles/r2_ab_groups.txt): the sweep is row-latency bound, so its time hardly depends on the number of pages;
    // splitting a 64-page batch multiplies that time and overlap does not win it back.  Groups pay off only when each
    // of them still fills the machine -- and even 2 x 128 small gray pages ran 10 % slower than one group of 256.
    int g = env_groups > 0 ? env_groups : 1;
    if (g > N) g = N;
    GroupPlan gp;
    gp.per_group = (N + g - 1) / g;
    gp.groups = (N + gp.per_group - 1) / gp.per_group;
    gp.streams = env_streams > 0 ? env_streams : (gp.groups >= 4 ? 3 : gp.groups);
    if (gp.streams > MAX_INTERNAL_STREAMS) gp.streams = MAX_INTERNAL_STREAMS;
    if (gp.streams > gp.groups) gp.streams = gp.groups;
    if (gp.groups == 1) gp.streams = 0;                  // the caller's stream
    return gp;
}

struct DevStreams {
    bool ready = false;
    cudaStream_t s[MAX_INTERNAL_STREAMS];
    cudaEvent_t ev_in, ev_done[MAX_INTERNAL_STREAMS];
};
std::mutex g_streams_mu;
DevStreams g_streams[64];

int internal_streams(DevStreams **out)
{
    int dev = 0;
    B200MRC_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return B200MRC_ERR_UNSUPPORTED;
    std::lock_guard<std::mutex> lk(g_streams_mu);
    DevStreams &d = g_streams[dev];
    if (!d.ready) {
        for (int i = 0; i < MAX_INTERNAL_STREAMS; i++) {
            B200MRC_CUDA_TRY(cudaStreamCreateWithFlags(&d.s[i], cudaStreamNonBlocking));
            B200MRC_CUDA_TRY(cudaEventCreateWithFlags(&d.ev_done[i], cudaEventDisableTiming));
        }
        B200MRC_CUDA_TRY(cudaEventCreateWithFlags(&d.ev_in, cudaEventDisableTiming));
        d.ready = true;
    }
    *out = &d;
    return B200MRC_OK;
}

struct DecomposeLayout {
    size_t gray_pitch, gray_page;
    size_t full_pitch, full_page;          // full-resolution fg / bg planes when downsampled
    GroupPlan gp;
    size_t off_sigma, off_gray, off_fgfull, off_bgfull, off_scratch, scratch_bytes, off_opt, opt_bytes, total;
};

DecomposeLayout decompose_layout(const b200mrc_decompose_args *a)
{
    DecomposeLayout L;
    memset(&L, 0, sizeof(L));
    const int W = a->width, H = a->height, N = a->n_pages, C = a->channels;
    const bool mask_only = (a->flags & B200MRC_DECOMPOSE_MASK_ONLY) != 0;
    L.gp = group_plan(N);
    const int n = L.gp.per_group;
    L.gray_pitch = align_up((size_t)W, 16); L.gray_page = L.gray_pitch * H;
    L.full_pitch = align_up((size_t)W * C, 16); L.full_page = L.full_pitch * H;
    Carver c;
    L.off_sigma = c.take<double>(N);
    L.off_gray = c.take<uint8_t>(L.gray_page * N + 256);     // the threshold's gray delay line (sauvola_fused.cu)
    if (!mask_only && a->fg_plan) L.off_fgfull = c.take<uint8_t>(L.full_page * N);
    if (!mask_only && a->bg_plan) L.off_bgfull = c.take<uint8_t>(L.full_page * N);
    // per group: one scratch region shared by the stages that run one after another on the group's stream ...
    size_t s = noise_workspace_bytes(W, H, n);
    if (a->flags & B200MRC_DECOMPOSE_DENOISE_FAST) s = std::max(s, denoise_workspace_bytes(W, H, n));
    if (!mask_only) {
        if (a->fg_plan) s = std::max(s, b200mrc_resample_workspace_bytes(a->fg_plan, n));
        if (a->bg_plan) s = std::max(s, b200mrc_resample_workspace_bytes(a->bg_plan, n));
    }
    L.scratch_bytes = align_up(s, 256);
    L.off_scratch = c.take<uint8_t>(L.scratch_bytes * L.gp.groups);
    // ... and the optimise workspace (record plane + the sweep's mailbox), which no other stage writes
    if (!mask_only) {
        L.opt_bytes = align_up(optimise_workspace_bytes(W, H, n), 256);
        L.off_opt = c.take<uint8_t>(L.opt_bytes * L.gp.groups);
    }
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

// Pages [p0, p0 + n) of the batch, every stage enqueued on `st`.
int decompose_group(const b200mrc_decompose_args *a, const DecomposeLayout &L, int g, int p0, int n, cudaStream_t st)
{
    uint8_t *ws = (uint8_t *)a->workspace;
    const int W = a->width, H = a->height, C = a->channels;
    const uint8_t *img = a->img + (int64_t)p0 * a->img_page_stride;
    uint8_t *mask = a->mask + (int64_t)p0 * a->mask_page_stride;
    double *sigma = (double *)(ws + L.off_sigma) + p0;
    uint8_t *gray = ws + L.off_gray + (size_t)p0 * L.gray_page;
    void *scratch = ws + L.off_scratch + (size_t)g * L.scratch_bytes;
    int rc;

    // ---- noise estimate (or injected sigma)
    const double *sigma_used = sigma;
    if (a->flags & B200MRC_DECOMPOSE_NO_NOISE_EST) {
        if (a->sigma_in) B200MRC_CUDA_TRY(cudaMemcpyAsync(sigma, a->sigma_in + p0, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
        else sigma_used = nullptr;
    } else {
        rc = launch_estimate_noise(img, a->img_pitch, a->img_page_stride, C, W, H, n, sigma, scratch, L.scratch_bytes, st);
        if (rc) return rc;
    }
    if (a->sigma_out) {
        if (sigma_used) B200MRC_CUDA_TRY(cudaMemcpyAsync(a->sigma_out + p0, sigma, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
        else B200MRC_CUDA_TRY(cudaMemsetAsync(a->sigma_out + p0, 0, sizeof(double) * n, st));
    }
    // ---- gray (+ blur where sigma > 1) + Sauvola: mask = threshold_image(gray), or mask |= ... when the caller filled in
    //      the hOCR line masks
    rc = b200mrc_threshold_mask(img, a->img_pitch, a->img_page_stride, C, mask, a->mask_pitch, a->mask_page_stride, W, H, n,
                                a->window, a->window, a->k, a->R, sigma_used,
                                (a->flags & B200MRC_DECOMPOSE_OR_INTO_MASK) ? B200MRC_SAUVOLA_OR_INTO : 0,
                                gray, L.gray_page * n + 256, st);
    if (rc) return rc;
    // ---- denoise
    if (a->flags & B200MRC_DECOMPOSE_DENOISE_FAST) {
        rc = launch_denoise(mask, a->mask_pitch, a->mask_page_stride, W, H, n, scratch, L.scratch_bytes, st);
        if (rc) return rc;
    }
    if (a->flags & B200MRC_DECOMPOSE_MASK_ONLY) return B200MRC_OK;

    // ---- fg / bg
    uint8_t *fg_out = a->fg + (int64_t)p0 * a->fg_page_stride, *bg_out = a->bg + (int64_t)p0 * a->bg_page_stride;
    uint8_t *fg_full = a->fg_plan ? ws + L.off_fgfull + (size_t)p0 * L.full_page : fg_out;
    uint8_t *bg_full = a->bg_plan ? ws + L.off_bgfull + (size_t)p0 * L.full_page : bg_out;
    const int64_t fgp = a->fg_plan ? (int64_t)L.full_pitch : a->fg_pitch, fgs = a->fg_plan ? (int64_t)L.full_page : a->fg_page_stride;
    const int64_t bgp = a->bg_plan ? (int64_t)L.full_pitch : a->bg_pitch, bgs = a->bg_plan ? (int64_t)L.full_page : a->bg_page_stride;
    // tuning key BG_FOLLOW: the bg thumbnail runs as a follower of the sweep.  1 (default): only when the sweep fills the GPU
    // (>= 6 strips per SM) -- below that its waiting CTAs would take the SM slots that kernels of other streams (the
    // next chunk of a streamed batch) could use, which is worth more than the follower; 2: always; 0: never
    const int follow_mode = tune(T_BG_FOLLOW);
    const bool follow = a->bg_plan && (follow_mode >= 2 || (follow_mode == 1 && (int64_t)n * cdiv(W, 128) >= 6 * (int64_t)dev_info().sm_count));
    int *bg_prog = nullptr;
    bool bg_done = false;
    rc = launch_optimise(mask, a->mask_pitch, a->mask_page_stride, img, a->img_pitch, a->img_page_stride, C,
                         fg_full, fgp, fgs, 3, bg_full, bgp, bgs, 10, W, H, n, ws + L.off_opt + (size_t)g * L.opt_bytes, L.opt_bytes, st,
                         follow ? &bg_prog : nullptr);
    if (rc) return rc;
    if (bg_prog) {
        // the thumbnail pass starts beside the sweep (its programmatic dependent) and consumes bg rows as strips publish them
        rc = launch_resample_follow(a->bg_plan, bg_full, bgp, bgs, bg_out, a->bg_pitch, a->bg_page_stride, n, bg_prog, cdiv(W, 128), 128, st);
        if (rc == B200MRC_OK) bg_done = true;
        else if (rc != B200MRC_ERR_UNSUPPORTED) return rc;
    }
    if (a->fg_plan) {
        rc = b200mrc_resample(a->fg_plan, fg_full, fgp, fgs, fg_out, a->fg_pitch, a->fg_page_stride, n, scratch, L.scratch_bytes, st);
        if (rc) return rc;
    }
    if (a->bg_plan && !bg_done) {
        rc = b200mrc_resample(a->bg_plan, bg_full, bgp, bgs, bg_out, a->bg_pitch, a->bg_page_stride, n, scratch, L.scratch_bytes, st);
        if (rc) return rc;
    }
    return B200MRC_OK;
}

}  // namespace
}  // namespace b200mrc

using namespace b200mrc;

extern "C" int b200mrc_version(void) { return B200MRC_VERSION; }

extern "C" int b200mrc_profile_enable(int on)
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto &r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    g_prof_on.store(on != 0);
    return B200MRC_OK;
}

extern "C" int b200mrc_profile_report(char *buf, size_t cap)
{
    // "kernel,launches,total_ms" lines; synchronises with every recorded event
    std::map<std::string, std::pair<int, double>> acc;
    std::vector<std::string> order;
    std::lock_guard<std::mutex> lk(g_prof_mu);
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

extern "C" int b200mrc_set_tuning(const char *name, int value)
{
    if (!name) return B200MRC_ERR_INVALID;
    std::call_once(g_tune_once, tune_init);
    for (int i = 0; i < T_COUNT; i++)
        if (!strcmp(name, k_tune_defs[i].name)) { g_tune[i].store(value); return B200MRC_OK; }
    return B200MRC_ERR_INVALID;
}

extern "C" int b200mrc_get_tuning(const char *name, int *value)
{
    if (!name || !value) return B200MRC_ERR_INVALID;
    for (int i = 0; i < T_COUNT; i++)
        if (!strcmp(name, k_tune_defs[i].name)) { *value = tune((TuneKey)i); return B200MRC_OK; }
    return B200MRC_ERR_INVALID;
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
    const int N = a->n_pages;
    if (L.gp.streams == 0) return decompose_group(a, L, 0, 0, N, st);

    // page groups round-robin over the internal streams; the caller's stream is the fork and the join point
    DevStreams *ds = nullptr;
    rc = internal_streams(&ds);
    if (rc != B200MRC_OK) return rc;
    std::lock_guard<std::mutex> lk(g_streams_mu);        // one fork/join at a time per process: the events are shared
    B200MRC_CUDA_TRY(cudaEventRecord(ds->ev_in, st));
    for (int i = 0; i < L.gp.streams; i++) B200MRC_CUDA_TRY(cudaStreamWaitEvent(ds->s[i], ds->ev_in, 0));
    int first_err = B200MRC_OK;
    for (int g = 0; g < L.gp.groups && first_err == B200MRC_OK; g++) {
        const int p0 = g * L.gp.per_group, n = std::min(L.gp.per_group, N - p0);
        first_err = decompose_group(a, L, g, p0, n, ds->s[g % L.gp.streams]);
    }
    for (int i = 0; i < L.gp.streams; i++) {             // join even after an error: nothing may run past the caller's next op
        cudaEventRecord(ds->ev_done[i], ds->s[i]);
        cudaStreamWaitEvent(st, ds->ev_done[i], 0);
    }
    return first_err;
}
