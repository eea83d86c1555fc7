"""Reference-facing surface: the same two callables internetarchivepdf/mrc.py exports for the
pixel path, same names, argument meaning, yield order, dtypes, timing keys and error behaviour:

    threshold_image(img, dpi, k=0.34)                         (mrc.py:58-87)
    create_mrc_hocr_components(image, hocr_word_data, ...)    (mrc.py:334-471)

Everything numeric runs in the CUDA engine (libb200mrc.so); there is no CPU fallback.  PIL is
used only as the container type of the input image (and for the rare non-L/RGB mode
conversions, which the reference also leaves to PIL).
"""
from time import time

import numpy as np
import torch

from . import _lib as L
from . import engine as E
from .engine import DENOISE_NONE, DENOISE_FAST, DENOISE_BREGMAN

RECODE_RUNTIME_WARNING_TOO_SMALL_TO_DOWNSAMPLE = 'too-small-to-downsample'    # const.py:38

_engine = None


def get_engine():
    global _engine
    if _engine is None:
        _engine = E.MrcEngine()
    return _engine


def threshold_image(img, dpi, k=0.34):
    """Sauvola binarisation of a 2-D uint8 array; returns a bool array (mrc.py:58-87)."""
    img = np.asarray(img)
    if img.ndim != 2:
        raise ValueError('Buffer has wrong number of dimensions (expected 2, got %d)' % img.ndim)
    if img.dtype == np.bool_:
        img = img.view(np.uint8)
    if img.dtype != np.uint8:
        raise ValueError("Buffer dtype mismatch, expected 'UINT8DTYPE_t' but got '%s'" % img.dtype)
    h, w = img.shape
    if h == 0 or w == 0:
        return np.zeros(img.shape, dtype=bool)
    return get_engine().threshold_image_np(img, E.window_for_dpi(dpi), k=k)


def _sync_time(t0):
    torch.cuda.synchronize()
    return time() - t0


def create_hocr_mask(img, mask_arr, hocr_word_data, downsample=None, dpi=None, timing_data=None):
    """mrc.py:188-270.  Per text line: Sauvola (k=0.1) on the crop and on the inverted crop, pick
    the polarity by fill ratio.  The estimate_sigma tie-break on boolean crops (mrc.py:253-254)
    is the not-yet-built part of SURVEY.md section 8(f)1."""
    t = time()
    found = False
    for paragraph in hocr_word_data:
        for line in paragraph['lines']:
            found = True
    if found:
        raise NotImplementedError('hOCR line masks (mrc.py:188-270) are the next row of the hot-path scope '
                                  '(SURVEY.md section 8f); pass hocr_word_data=[] for now')
    if timing_data is not None:
        timing_data.append(('hocr_mask_gen', time() - t))


def create_mrc_hocr_components(image, hocr_word_data,
                               dpi=None,
                               downsample=None,
                               bg_downsample=None,
                               fg_downsample=None,
                               denoise_mask=None, timing_data=None,
                               errors=None):
    """Create the MRC components: mask, foreground and background (generator, mrc.py:334-471).

    image: PIL.Image.  Yields mask (bool HxW), fg (uint8 HxW[x3]), bg (uint8 H'xW'[x3]).
    """
    eng = get_engine()
    width_, height_ = image.size
    mode = image.mode

    # ---- inputs on the device.  L / RGB go up as they are (gray conversion is fused in the
    # kernels); any other mode is converted by PIL exactly like the reference does (mrc.py:361, 404)
    t = time()
    gray_host = None
    if mode == 'L':
        page = np.asarray(image)
    elif mode == 'RGB':
        page = np.asarray(image)
    else:
        gray_host = np.asarray(image.convert('L'))
        page = None
    if mode != 'L' and timing_data is not None:
        timing_data.append(('grey_conversion', time() - t))

    mask = E.Plane(1, height_, width_, 1, eng.device)
    create_hocr_mask(None, None, hocr_word_data, downsample=downsample, dpi=dpi, timing_data=timing_data)

    if page is not None:
        src = E.Plane(1, height_, width_, page.shape[2] if page.ndim == 3 else 1, eng.device).upload(page[None])
    else:
        src = E.Plane(1, height_, width_, 1, eng.device).upload(gray_host[None])

    # ---- create_threshold_mask (mrc.py:300-329)
    t = time()
    sigma_dev = eng.estimate_noise(src)
    sigma_est = float(sigma_dev.cpu()[0])              # synchronises: 'est_1' is a true stage time
    if timing_data is not None:
        timing_data.append(('est_1', time() - t))
    gray = E.Plane(1, height_, width_, 1, eng.device)
    if sigma_est > 1.0:
        t = time()
        eng.gray_blur(src, gray, sigma_dev)
        if timing_data is not None:
            timing_data.append(('blur_1', _sync_time(t)))
    else:
        eng.gray_blur(src, gray, None)
    t = time()
    eng.sauvola(gray, mask, E.window_for_dpi(dpi), k=0.34, R=128.0)
    if timing_data is not None:
        timing_data.append(('threshold', _sync_time(t)))

    if denoise_mask != DENOISE_NONE:
        t = time()
        if denoise_mask == DENOISE_FAST:
            eng.denoise(mask, 4, 2)
            if timing_data is not None:
                timing_data.append(('fast_denoise', _sync_time(t)))
        elif denoise_mask == DENOISE_BREGMAN:
            raise NotImplementedError('denoise_bregman is out of scope (SURVEY.md section 2)')
        else:
            raise ValueError('Invalid denoise option:', denoise_mask)            # mrc.py:396

    mask_arr = mask.numpy(np.bool_)[0]
    yield mask_arr

    # ---- foreground / background (mrc.py:401-470): one fused sweep produces both layers
    if mode not in ('L', 'RGB'):
        page = np.asarray(image.convert('RGB'))                                    # mrc.py:401-404
        src = E.Plane(1, height_, width_, 3, eng.device).upload(page[None])
    c = src.c
    t = time()
    fg_full = E.Plane(1, height_, width_, c, eng.device)
    bg_full = E.Plane(1, height_, width_, c, eng.device)
    eng.optimise(mask, src, fg_full, 3, bg_full, 10)
    t_opt = _sync_time(t)
    if timing_data is not None:
        timing_data.append(('fg_partial_blur', t_opt / 2))

    fg = fg_full
    if fg_downsample is not None:
        t = time()
        plan, too_small = E.downsample_plan(width_, height_, c, fg_downsample)
        if plan is not None:
            fg = E.Plane(1, plan.out_h, plan.out_w, c, eng.device)
            eng.resample(plan, fg_full, fg)
        elif too_small and errors is not None:
            errors.add(RECODE_RUNTIME_WARNING_TOO_SMALL_TO_DOWNSAMPLE)
        if timing_data is not None:
            timing_data.append(('fg_downsample', _sync_time(t)))
    yield fg.numpy()[0]

    if timing_data is not None:
        timing_data.append(('bg_partial_blur', t_opt / 2))
    bg = bg_full
    if bg_downsample is not None:
        t = time()
        plan, too_small = E.downsample_plan(width_, height_, c, bg_downsample)
        if plan is not None:
            bg = E.Plane(1, plan.out_h, plan.out_w, c, eng.device)
            eng.resample(plan, bg_full, bg)
        elif too_small and errors is not None:
            errors.add(RECODE_RUNTIME_WARNING_TOO_SMALL_TO_DOWNSAMPLE)
        if timing_data is not None:
            timing_data.append(('bg_downsample', _sync_time(t)))
    yield bg.numpy()[0]
    return


def decompose_pages(pages, dpi=None, window=None, bg_downsample=None, fg_downsample=None,
                    denoise_mask=DENOISE_FAST, mask_only=False, sigma=None, batch=None):
    """Batched form of create_mrc_hocr_components for N equally-shaped pages held in HOST memory
    (uint8 ndarray [N,H,W] or [N,H,W,3], or a pinned CPU tensor): one H2D copy, one
    b200mrc_decompose, D2H of the results.  Returns dict(mask, fg, bg, sigma, errors).
    `batch`: a DecomposeBatch to reuse (device buffers + workspace)."""
    eng = get_engine()
    shape = tuple(pages.shape)
    n, h, w = shape[:3]
    c = shape[3] if len(shape) == 4 else 1
    if batch is None:
        batch = eng.make_batch(n, h, w, c, bg_downsample, fg_downsample, mask_only)
    batch.img.upload(pages)
    batch.run(window if window is not None else E.window_for_dpi(dpi), denoise_mask=denoise_mask, sigma=sigma)
    res = dict(mask=batch.mask.numpy(np.bool_), sigma=batch.sigma.cpu().numpy(), errors=set(batch.errors))
    if not mask_only:
        res['fg'] = batch.fg.numpy()
        res['bg'] = batch.bg.numpy()
    return res
