// blur.cuh -- device helpers shared by gray_blur.cu and sauvola_fused.cu: the per-page blur decision of
// create_threshold_mask (internetarchivepdf/mrc.py:305-311) and scipy's Gaussian weights
// (scipy.ndimage._filters._gaussian_kernel1d), restated exactly (oracle: orc_gauss_blur).
#pragma once
#include <stdint.h>

namespace b200mrc {

// scipy 'reflect' boundary: d c b a | a b c d | d c b a
__device__ __forceinline__ int reflect_idx(int i, int n)
{
    if (n == 1) return 0;
    const int per = 2 * n;
    i %= per; if (i < 0) i += per;
    return i < n ? i : per - 1 - i;
}

// numpy pairwise sum (n <= 128 path): what phi_x.sum() does in scipy's _gaussian_kernel1d
__device__ inline double np_sum(const double *a, int n)
{
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; i++) res = __dadd_rn(res, a[i]);
        return res;
    }
    double r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; j++) r[j] = __dadd_rn(r[j], a[i + j]);
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; i++) res = __dadd_rn(res, a[i]);
    return res;
}

// mrc.py:309-311: blur only if sigma_est > 1.0, with sigma = 0.1 * sigma_est; scipy radius = int(4 sigma + 0.5)
__device__ __forceinline__ int blur_radius_of(const double *sigma_arr, int page, double &sigma)
{
    const double sig_est = sigma_arr ? sigma_arr[page] : 0.0;
    sigma = 0.0;
    if (!(sig_est > 1.0)) return 0;            // NaN compares false: no blur
    sigma = sig_est * 0.1;
    const double rr = 4.0 * sigma + 0.5;
    return rr > 1.0e6 ? 1000000 : (int)rr;
}

// weights w[0..radius] (double), CTA-cooperative: sphi is scratch of 2*radius+1 doubles; the sum runs on one
// thread so that its order is numpy's
__device__ inline void blur_weights_cta(int radius, double sigma, double *sw, double *sphi)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i <= 2 * radius; i += nt) {
        const int j = i - radius;
        const double sigma2 = sigma * sigma;
        sphi[i] = exp(__dmul_rn(-0.5 / sigma2, (double)(j * j)));
    }
    __syncthreads();
    if (tid == 0) {
        const double sum = np_sum(sphi, 2 * radius + 1);
        for (int j = 0; j <= radius; j++) sw[j] = sphi[radius + j] / sum;
    }
    __syncthreads();
}

// (double)x for 0 <= x < 2^31 without the conversion unit: 2^52 + x is exact, subtract 2^52 on the FP64 pipe
__device__ __forceinline__ double u2d(uint32_t x) { return __hiloint2double(0x43300000, (int)x) - 4503599627370496.0; }

}  // namespace b200mrc
