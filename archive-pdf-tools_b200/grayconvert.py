"""special_gray_convert (internetarchivepdf/grayconvert.py:38-66) on the GPU: integer channel
statistics on the device, the threshold arithmetic of grayconvert.py:41-55 on the host (exactly
as the reference does it, from the same statistics), per-pixel level stretch + HSL lightness on
the device."""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L
from . import engine as E

perc2val = lambda x: (x * 255) / 100            # grayconvert.py:22


def thresholds_from_stats(stats, npix):
    """stats: uint64 [3][4] = min, max, sum, sumsq per channel -> (minv[3], maxv[3]) like :41-55."""
    d = {}
    for i, k in enumerate('rgb'):
        mn, mx, sm, sq = (int(v) for v in stats[i])
        mean = sm / npix
        # np.std: sqrt(mean(|x - mean|^2)); from exact integers: (sq - sm^2/npix)/npix
        var = (sq * npix - sm * sm) / (npix * npix)
        d[k + '_min'], d[k + '_max'] = mn / 255., mx / 255.
        d[k + '_mean'], d[k + '_std'] = mean / 255., math.sqrt(var) / 255.
    bright_adjust = round(d['r_mean'] * d['g_mean'] * d['b_mean'] /
                          (d['b_max'] * (1 - d['r_std']) * (1 - d['g_std']) * (1 - d['b_std'])), 4)
    low_thres = min(int((196 * d['r_min'] + 14.5) / 1), 50)
    high = [min(int((35.66 * bright_adjust + 48.5) / 1), 95),
            min(int((39.22 * bright_adjust + 44.5) / 1), 95),
            min(int((45.16 * bright_adjust + 36.5) / 1), 95)]
    return [perc2val(low_thres)] * 3, [perc2val(h) for h in high]


def special_gray_convert(imd, engine=None):
    """imd: uint8 ndarray H x W x 3 -> uint8 ndarray H x W."""
    from . import get_engine
    eng = engine or get_engine()
    imd = np.ascontiguousarray(imd)
    h, w = imd.shape[:2]
    src = E.Plane(1, h, w, 3, eng.device).upload(imd[None], non_blocking=False)
    stats = torch.zeros((1, 3, 4), dtype=torch.int64, device=eng.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(L.lib().b200mrc_channel_stats(src.ptr, src.pitch, src.page_stride, w, h, 1, C.c_void_p(stats.data_ptr()), st),
            'b200mrc_channel_stats')
    minv, maxv = thresholds_from_stats(stats.cpu().numpy()[0].astype(np.uint64), h * w)
    mv = torch.tensor([minv], dtype=torch.float64, device=eng.device)
    xv = torch.tensor([maxv], dtype=torch.float64, device=eng.device)
    out = E.Plane(1, h, w, 1, eng.device)
    L.check(L.lib().b200mrc_special_gray(src.ptr, src.pitch, src.page_stride, out.ptr, out.pitch, out.page_stride,
                                         w, h, 1, C.c_void_p(mv.data_ptr()), C.c_void_p(xv.data_ptr()), st),
            'b200mrc_special_gray')
    return out.numpy()[0]
