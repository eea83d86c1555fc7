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
void prof_begin(const char *kernel, cudaStream_t st);
void prof_end(cudaStream_t st);
struct ProfScope {
    cudaStream_t st;
    ProfScope(const char *kernel, cudaStream_t s) : st(s) { prof_begin(kernel, s); }
    ~ProfScope() { prof_end(st); }
};

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
