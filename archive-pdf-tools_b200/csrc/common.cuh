// common.cuh -- shared helpers for the b200mrc kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "../../include/b200mrc.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "b200mrc is written for sm_100a (Blackwell B200) only"
#endif

namespace b200mrc {

extern std::atomic<uint64_t> g_launch_count;

inline void count_launch(uint64_t n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

#define B200MRC_CUDA_TRY(expr)                                   \
    do {                                                         \
        cudaError_t _e = (expr);                                 \
        if (_e != cudaSuccess) return (int)_e;                   \
    } while (0)

#define B200MRC_LAUNCH_CHECK()                                   \
    do {                                                         \
        ::b200mrc::count_launch();                               \
        cudaError_t _e = cudaGetLastError();                     \
        if (_e != cudaSuccess) return (int)_e;                   \
    } while (0)

// Optional per-kernel CUDA-event timing (b200mrc_profile_enable / b200mrc_profile_report): an event
// pair is recorded on the launching stream around every kernel while profiling is on.
int prof_begin(const char *kernel, cudaStream_t st);      // -> record index, -1 when profiling is off
void prof_end(int record, cudaStream_t st);
struct ProfScope {
    cudaStream_t st; int rec;
    ProfScope(const char *kernel, cudaStream_t s) : st(s), rec(prof_begin(kernel, s)) {}
    ~ProfScope() { prof_end(rec, st); }
};

// Tuning knobs (A/B runs, tests).  Each is initialised ONCE per process from the environment variable B200MRC_<NAME>
// and can be changed at run time with b200mrc_set_tuning(); launchers read them as relaxed atomics -- no getenv() on a
// launch path.  INTEGRATION.md lists them.
enum TuneKey {
    T_IIRW_MODE,        // sweep form: 0 auto (by batch size), 1 single (one warp per strip), 2 trio      env: single | trio
    T_IIRW_TPC,         // trio form: strips per CTA (1..3)
    T_IIRW_FEED,        // single form input feed: 0 auto, 1 tma, 2 async                                  env: tma | async
    T_IIRW_PSLEEP,      // trio form: producer back-off in ns
    T_IIRW_WPC,         // single form: warps per CTA
    T_FIRW_BAND,        // FIR pass: rows per band
    T_FIRW_WPC,         // FIR pass: warps per CTA
    T_FUSED_NT,         // fused threshold: threads per CTA (0 auto, 128 | 192 | 256)
    T_FUSED_BANDS,      // fused threshold: row bands per strip (0 auto)
    T_FUSED_DBG,        // fused threshold: timing experiments; honoured only by builds with -DB200MRC_EXPERIMENTS (results are wrong)
    T_FUSED_OCC,        // fused threshold, 128-thread CTAs: CTAs per SM the kernel is compiled for (0 = 4; 3 | 5)
    T_THRESHOLD_PATH,   // 0 auto (fused where it applies), 1 two-pass (gray_blur + sauvola kernels), 2 fused      env: legacy | fused
    T_OPT_PATH,         // 0 split (FIR + sweep), 1 generic fused sweep                                    env: generic
    T_NOISE_DIRECT,     // 1: one-thread-per-coefficient wavelet kernel
    T_RESAMPLE_2PASS,   // 1: separate horizontal / vertical resample kernels
    T_TILE_H,           // tile resampler: output rows per CTA (16 | 32 | 64)
    T_DECOMPOSE_GROUPS, // page groups per b200mrc_decompose call (0 auto)
    T_DECOMPOSE_STREAMS,// internal streams the groups run on (0 auto)
    T_BG_FOLLOW,        // the bg thumbnail pass follows the sweep's progress counters (programmatic dependent launch): 1 when the sweep fills the GPU, 2 always, 0 never (runs after it)
    T_COUNT
};
int tune(TuneKey k);

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Device properties cached per device.
struct DevInfo {
    int sm_count;
    int max_smem_optin;
};
const DevInfo &dev_info();

// Bump allocator of byte offsets inside a caller-provided workspace (256-byte aligned).
struct Carver {
    size_t off = 0;
    template <typename T> size_t take(size_t count) {
        off = align_up(off, 256);
        const size_t r = off;
        off += count * sizeof(T);
        return r;
    }
    size_t used() const { return align_up(off, 256); }
};

// PIL L24 luma (Convert.c): (19595 R + 38470 G + 7471 B + 0x8000) >> 16
__device__ __forceinline__ uint32_t luma_l24(uint32_t r, uint32_t g, uint32_t b)
{
    return (19595u * r + 38470u * g + 7471u * b + 0x8000u) >> 16;
}

}  // namespace b200mrc
