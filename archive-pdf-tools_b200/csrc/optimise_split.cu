// optimise_split.cu -- launcher of the production optimise path for n_fg = 3 / n_bg = 10
// (internetarchive/archive-pdf-tools internetarchivepdf/mrc.py:412-415, 439-449;
// semantics cython/optimiser.pyx:153-429, restated in optimise.cu / oracle orc_optimise).
//
//   out[y,x] = (FIR(y,x) + IIR(y,x)) / den(y,x)        for every pixel not in the layer's mask
//   FIR = sum of mask*img over the 2n x 2n box (depends only on the inputs: fully parallel)
//   IIR = sum of `out` over the n x n box above-left (sequential over rows)
//   den = #mask in the FIR box + (y-ys)(x-xs)
// Exactly one layer is computed per pixel (fg where mask==0, bg where mask==1), so the parallel half hands the
// sequential half one 64-bit record per pixel:
//   k_opt_fir_w (optimise_firw.cu) : FIR sums + den -> record plane (8 B / pixel, written once, read once)
//   k_opt_iir_w (optimise_warp.cu) : the row-sequential sweep, both layers, writes fg and bg
// Both are warp-strip kernels; this file only checks that the caller's planes allow them (16-byte aligned bases and
// pitches, so that rows can move as bulk / vector copies) and lays the workspace out.  Anything else (other n,
// unaligned planes) is served by the generic sweep in optimise.cu.
#include "common.cuh"
#include <cstdlib>

namespace b200mrc {

int launch_opt_fir_warp(const uint8_t *mask, int64_t mpitch, int64_t mstride,
                        const uint8_t *img, int64_t ipitch, int64_t istride, int C,
                        uint8_t *rec, int64_t rpitch, int64_t rstride,
                        int W, int H, int N, int band_h, int wpc, uint32_t *mailbox, int S_sweep, cudaStream_t st);
int launch_opt_iir_warp(const uint8_t *img, int64_t ipitch, int64_t istride, int C,
                        const uint8_t *rec, int64_t rpitch, int64_t rstride,
                        uint8_t *ofg, int64_t fpitch, int64_t fstride,
                        uint8_t *obg, int64_t bpitch, int64_t bstride,
                        int W, int H, int N, uint32_t *mailbox, unsigned *ticket, int wpc, int *progress, cudaStream_t st);

// record plane: 8 B / pixel, rows padded to whole 4-pixel groups (the sweep reads whole groups)
int64_t optimise_split_rec_pitch(int W) { return (int64_t)((W + 3) / 4 * 4) * 8; }
size_t optimise_split_rec_bytes(int W, int H, int N) { return (size_t)optimise_split_rec_pitch(W) * (size_t)H * (size_t)N; }

// Returns B200MRC_ERR_UNSUPPORTED when this path does not apply (the caller falls back to the generic sweep).
int launch_optimise_split(const uint8_t *mask, int64_t mpitch, int64_t mstride,
                          const uint8_t *img, int64_t ipitch, int64_t istride, int C,
                          uint8_t *ofg, int64_t fpitch, int64_t fstride,
                          uint8_t *obg, int64_t bpitch, int64_t bstride,
                          int W, int H, int N, uint8_t *rec, uint32_t *mailbox, unsigned *ticket, int *progress, cudaStream_t st)
{
    auto al16 = [](const void *q) { return ((uintptr_t)q & 15) == 0; };
    const int64_t need_i = ((int64_t)W * C + 15) & ~15ll, need_m = ((int64_t)W + 15) & ~15ll;
    if (!ofg || !obg || !rec) return B200MRC_ERR_UNSUPPORTED;
    if (!al16(mask) || !al16(img) || !al16(ofg) || !al16(obg) || !al16(rec)) return B200MRC_ERR_UNSUPPORTED;
    if ((mpitch | mstride | ipitch | istride | fpitch | fstride | bpitch | bstride) & 15) return B200MRC_ERR_UNSUPPORTED;
    if (mpitch < need_m || ipitch < need_i || fpitch < need_i || bpitch < need_i) return B200MRC_ERR_UNSUPPORTED;
    if (N > 65535) return B200MRC_ERR_UNSUPPORTED;
    const int64_t rpitch = optimise_split_rec_pitch(W), rstride = rpitch * H;
    const int firw_band = tune(T_FIRW_BAND), firw_wpc = tune(T_FIRW_WPC), iirw_wpc = tune(T_IIRW_WPC);
    // the FIR pass also clears the sweep's mailbox rows (128 columns per sweep strip)
    int rc = launch_opt_fir_warp(mask, mpitch, mstride, img, ipitch, istride, C, rec, rpitch, rstride, W, H, N,
                                 firw_band, firw_wpc, mailbox, cdiv(W, 128), st);
    if (rc != B200MRC_OK) return rc;
    return launch_opt_iir_warp(img, ipitch, istride, C, rec, rpitch, rstride, ofg, fpitch, fstride, obg, bpitch, bstride,
                               W, H, N, mailbox, ticket, iirw_wpc, progress, st);
}

}  // namespace b200mrc
