"""ctypes front-end of oracle/libmrc_oracle.so + the restated page pipeline.

TEST INFRASTRUCTURE ONLY -- the checker for the CUDA engine.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the product package
(archive-pdf-tools_b200/) never does.

`decompose()` restates internetarchivepdf/mrc.py:334-471 (create_mrc_hocr_components) for
hocr_word_data == [] on top of the C restatements; `ref_decompose()` (oracle/ref_pipeline.py) does
the same on top of the reference's own compiled Cython (oracle/_ref) + real Pillow + real scipy.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libmrc_oracle.so')

DENOISE_NONE, DENOISE_FAST, DENOISE_BREGMAN = 'none', 'fast', 'bregman'   # const.py:31-33


def build(force=False):
    src = os.path.join(_HERE, 'mrc_oracle.c')
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-fPIC', '-shared',
                               '-fvisibility=hidden', src, '-o', _LIB_PATH, '-lm'])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p, i32p, f32p, f64p = (C.POINTER(C.c_uint8), C.POINTER(C.c_int), C.POINTER(C.c_float),
                                 C.POINTER(C.c_double))
        L.orc_rgb2gray.argtypes = [u8p, C.c_int64, u8p]
        L.orc_sauvola.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
        L.orc_denoise.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_optimise.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        L.orc_gauss_radius.argtypes = [C.c_double]
        L.orc_gauss_weights.argtypes = [C.c_double, C.c_int, f64p]
        L.orc_gauss_blur.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_double]
        L.orc_noise_crop.argtypes = [C.c_int, C.c_int, i32p, i32p, i32p, i32p]
        L.orc_estimate_sigma_crop.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_estimate_sigma_crop.restype = C.c_double
        L.orc_estimate_sigma_bool.argtypes = [u8p, C.c_int, C.c_int, C.c_int]
        L.orc_estimate_sigma_bool.restype = C.c_double
        L.orc_estimate_noise.argtypes = [u8p, C.c_int, C.c_int]
        L.orc_estimate_noise.restype = C.c_double
        L.orc_resample_ksize.argtypes = [C.c_int, C.c_float, C.c_float, C.c_int, C.c_int]
        L.orc_resample_coeffs.argtypes = [C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, i32p, i32p]
        L.orc_resample.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int,
                                   C.c_float, C.c_float, C.c_float, C.c_float, C.c_int]
        L.orc_reduce.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, u8p]
        L.orc_special_gray_pixels.argtypes = [u8p, C.c_int64, f64p, f64p, u8p]
        _lib = L
    return _lib


def _p(a, t=C.c_uint8):
    return a.ctypes.data_as(C.POINTER(t))


def _u8(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.bool_:
        a = a.view(np.uint8)
    assert a.dtype == np.uint8
    return a


# --------------------------------------------------------------------------- kernels
def rgb2gray(rgb):
    rgb = _u8(rgb)
    out = np.empty(rgb.shape[:2], np.uint8)
    lib().orc_rgb2gray(_p(rgb), rgb.shape[0] * rgb.shape[1], _p(out))
    return out


def window_for_dpi(dpi):
    """mrc.py:70-75"""
    window_size = 51
    if dpi is not None:
        window_size = int(dpi / 4)
        if window_size % 2 == 0:
            window_size += 1
    return window_size


def sauvola(gray, window_w, window_h=None, k=0.34, R=128.0):
    """Returns the fg mask (== threshold_image's return value), bool H x W."""
    gray = _u8(gray)
    H, W = gray.shape
    out = np.empty((H, W), np.uint8)
    rc = lib().orc_sauvola(_p(gray), _p(out), W, H, window_w, window_h or window_w, k, R)
    assert rc == 0
    return out.view(np.bool_)


def threshold_image(img, dpi, k=0.34):
    """mrc.py:58-87"""
    return sauvola(img, window_for_dpi(dpi), k=k)


def denoise(mask, mincnt=4, n=2):
    m = _u8(mask).copy()
    H, W = m.shape
    lib().orc_denoise(_p(m), W, H, mincnt, n)
    return m.view(np.bool_)


def optimise(mask, img, n):
    mask, img = _u8(mask), _u8(img)
    H, W = mask.shape
    Cn = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty_like(img)
    rc = lib().orc_optimise(_p(mask), _p(img), W, H, Cn, n, _p(out))
    assert rc == 0
    return out


def gauss_blur(imgf, sigma):
    imgf = np.ascontiguousarray(imgf, np.float32)
    out = np.empty_like(imgf)
    rc = lib().orc_gauss_blur(_p(imgf, C.c_float), _p(out, C.c_float), imgf.shape[1], imgf.shape[0], sigma)
    assert rc == 0
    return out


def estimate_noise(gray):
    gray = _u8(gray)
    return lib().orc_estimate_noise(_p(gray), gray.shape[1], gray.shape[0])


def estimate_sigma_full(gray):
    """estimate_sigma of a whole uint8 image held as float32 (no crop)."""
    gray = _u8(gray)
    return lib().orc_estimate_sigma_crop(_p(gray), gray.shape[1], 0, gray.shape[0], 0, gray.shape[1])


def estimate_sigma_bool(arr):
    """mean_estimate_sigma of a boolean array (mrc.py:253-254): float64 db2 'dd' median."""
    a = np.ascontiguousarray(np.asarray(arr) != 0).view(np.uint8)
    return lib().orc_estimate_sigma_bool(_p(a), a.shape[1], a.shape[0], a.shape[1])


# --------------------------------------------------------------------------- thumbnail
BICUBIC, LANCZOS = 0, 1


def thumbnail_plan(W, H, req_w, req_h, reducing_gap=2.0, filter=BICUBIC):
    """Mirror of PIL.Image.thumbnail + Image.resize size/box logic (Pillow 12.2 Image.py).
    Returns None (no-op) or dict(out_w, out_h, fx, fy, rbox=(x0,y0,x1,y1), box=(4 floats))."""
    px, py = math.floor(req_w), math.floor(req_h)
    if px >= W and py >= H:
        return None

    def round_aspect(number, key):
        return max(min(math.floor(number), math.ceil(number), key=key), 1)

    aspect = W / H
    x, y = px, py
    if x / y >= aspect:
        x = round_aspect(y * aspect, key=lambda n: abs(aspect - n / y))
    else:
        y = round_aspect(x / aspect, key=lambda n: 0 if n == 0 else abs(aspect - x / n))
    if (x, y) == (W, H):
        return None
    box = (0, 0, W, H)
    fx = fy = 1
    rbox = (0, 0, W, H)
    if reducing_gap is not None:
        fx = int((box[2] - box[0]) / x / reducing_gap) or 1
        fy = int((box[3] - box[1]) / y / reducing_gap) or 1
        if fx > 1 or fy > 1:
            support = (3.0 if filter == LANCZOS else 2.0) - 0.5
            sx = (box[2] - box[0]) / x
            sy = (box[3] - box[1]) / y
            rbox = (max(0, int(box[0] - support * sx)), max(0, int(box[1] - support * sy)),
                    min(W, math.ceil(box[2] + support * sx)), min(H, math.ceil(box[3] + support * sy)))
            box = ((box[0] - rbox[0]) / fx, (box[1] - rbox[1]) / fy,
                   (box[2] - rbox[0]) / fx, (box[3] - rbox[1]) / fy)
    return dict(out_w=x, out_h=y, fx=fx, fy=fy, rbox=rbox, box=tuple(float(b) for b in box))


def thumbnail(img, req_w, req_h, reducing_gap=2.0, filter=BICUBIC):
    """PIL Image.fromarray(img).thumbnail((req_w, req_h)) restated."""
    img = _u8(img)
    H, W = img.shape[:2]
    Cn = 1 if img.ndim == 2 else img.shape[2]
    plan = thumbnail_plan(W, H, req_w, req_h, reducing_gap, filter)
    if plan is None:
        return img.copy()
    L = lib()
    src, sw, sh = img, W, H
    if plan['fx'] > 1 or plan['fy'] > 1:
        x0, y0, x1, y1 = plan['rbox']
        sw = (x1 - x0 + plan['fx'] - 1) // plan['fx']
        sh = (y1 - y0 + plan['fy'] - 1) // plan['fy']
        red = np.empty((sh, sw) + ((Cn,) if img.ndim == 3 else ()), np.uint8)
        L.orc_reduce(_p(img), W, H, Cn, x0, y0, x1 - x0, y1 - y0, plan['fx'], plan['fy'], _p(red))
        src = red
    out = np.empty((plan['out_h'], plan['out_w']) + ((Cn,) if img.ndim == 3 else ()), np.uint8)
    b = plan['box']
    rc = L.orc_resample(_p(src), sw, sh, Cn, _p(out), plan['out_w'], plan['out_h'], b[0], b[1], b[2], b[3], filter)
    assert rc == 0
    return out


# --------------------------------------------------------------------------- special gray
def special_gray_thresholds(imd):
    """grayconvert.py:41-55, verbatim arithmetic on numpy statistics."""
    d = {}
    for i, k in enumerate('rgb'):
        for fun in ('min', 'max', 'mean', 'std'):
            d[k + '_' + fun] = getattr(np, fun)(imd[:, :, i]) / 255.
    bright_adjust = round(d['r_mean'] * d['g_mean'] * d['b_mean'] /
                          (d['b_max'] * (1 - d['r_std']) * (1 - d['g_std']) * (1 - d['b_std'])), 4)
    low_thres = min(int((196 * d['r_min'] + 14.5) / 1), 50)
    high = [min(int((35.66 * bright_adjust + 48.5) / 1), 95),
            min(int((39.22 * bright_adjust + 44.5) / 1), 95),
            min(int((45.16 * bright_adjust + 36.5) / 1), 95)]
    perc2val = lambda x: (x * 255) / 100
    return [perc2val(low_thres)] * 3, [perc2val(h) for h in high]


def special_gray_convert(imd):
    imd = _u8(imd)
    minv, maxv = special_gray_thresholds(imd)
    minv = np.asarray(minv, np.float64)
    maxv = np.asarray(maxv, np.float64)
    out = np.empty(imd.shape[:2], np.uint8)
    lib().orc_special_gray_pixels(_p(imd), imd.shape[0] * imd.shape[1], _p(minv, C.c_double),
                                  _p(maxv, C.c_double), _p(out))
    return out


# --------------------------------------------------------------------------- page pipeline
def threshold_mask(gray, dpi=None, window=None, k=0.34, sigma_est=None):
    """create_threshold_mask (mrc.py:300-329) on a uint8 gray page; returns (mask, sigma_est)."""
    if sigma_est is None:
        sigma_est = estimate_noise(gray)
    g = gray
    if sigma_est > 1.0:
        g = gauss_blur(gray.astype(np.float32), sigma_est * 0.1).astype(np.uint8)   # mrc.py:311, 325
    w = window if window is not None else window_for_dpi(dpi)
    return sauvola(g, w, k=k), sigma_est


def hocr_lines(hocr_word_data, image_width, image_height, downsample=None):
    """The text-line boxes create_hocr_mask works on, in order (mrc.py:194-222)."""
    out = []
    for paragraph in hocr_word_data:
        for line in paragraph['lines']:
            coords = line['bbox']
            line_text = ' '.join([word['text'] for word in line['words']])
            line_confs = [word['confidence'] for word in line['words']]
            line_conf = sum(line_confs) / len(line_confs) if len(line_confs) else 0
            if line_text.strip() == '' or line_conf < 20:
                continue
            if downsample is not None:
                coords = [int(x / downsample) for x in coords]
            else:
                coords = [int(x) for x in coords]
            left, top, right, bottom = coords
            if left == right or top == bottom:
                continue
            if (left >= right) or (top >= bottom):
                continue
            if (left < 0) or (right > image_width) or (top < 0) or (bottom > image_height):
                continue
            out.append((left, top, right, bottom))
    return out


def hocr_choice(ratio, inv_ratio, sigmas):
    """mrc.py:238-263: 0 = leave the mask alone, 1 = thres, 2 = thres_invert.  `sigmas` is a callable returning
    (ratio_sigma, inv_ratio_sigma); it is only called when the reference would call mean_estimate_sigma."""
    if ratio < 0.3 or inv_ratio < 0.3:
        if inv_ratio > 0.2 and ratio < 0.2:
            return 1
        ratio_sigma, inv_ratio_sigma = sigmas()
        if inv_ratio < 0.3 and inv_ratio < ratio and \
                (inv_ratio_sigma < ratio_sigma or (ratio_sigma < 0.1 and inv_ratio_sigma < 0.1)):
            return 2
        elif ratio < 0.2:
            return 1
    return 0


def hocr_mask(gray, mask, hocr_word_data, downsample=None, dpi=None):
    """create_hocr_mask (mrc.py:188-270) on a uint8 gray page; `mask` (bool) is modified in place."""
    H, W = gray.shape
    w = window_for_dpi(dpi)
    for (left, top, right, bottom) in hocr_lines(hocr_word_data, W, H, downsample):
        crop = np.ascontiguousarray(gray[top:bottom, left:right])
        thres = sauvola(crop, w, k=0.1)
        thres_inv = sauvola(255 - crop, w, k=0.1)
        ratio = np.count_nonzero(thres) / crop.size
        inv_ratio = np.count_nonzero(thres_inv) / crop.size
        c = hocr_choice(ratio, inv_ratio, lambda: (estimate_sigma_bool(thres), estimate_sigma_bool(thres_inv)))
        if c:
            mask[top:bottom, left:right] = thres if c == 1 else thres_inv
    return mask


def decompose(image, dpi=None, bg_downsample=None, fg_downsample=None, denoise_mask=None,
              window=None, sigma_est=None, mask_only=False, hocr_word_data=(), downsample=None):
    """create_mrc_hocr_components (mrc.py:334-471); image: uint8 ndarray H x W (mode L) or
    H x W x 3 (mode RGB).  Returns dict(mask, fg, bg, sigma, errors)."""
    image = _u8(image)
    gray = image if image.ndim == 2 else rgb2gray(image)
    mask, sigma = threshold_mask(gray, dpi=dpi, window=window, sigma_est=sigma_est)
    if hocr_word_data:
        hm = hocr_mask(gray, np.zeros(gray.shape, bool), hocr_word_data, downsample=downsample, dpi=dpi)
        mask = hm | mask                                                        # mrc.py:329
    if denoise_mask != DENOISE_NONE:
        if denoise_mask == DENOISE_FAST:
            mask = denoise(mask, 4, 2)
        elif denoise_mask == DENOISE_BREGMAN:
            raise NotImplementedError('denoise_bregman is out of scope (SURVEY.md section 2)')
        else:
            raise ValueError('Invalid denoise option:', denoise_mask)      # mrc.py:396
    res = dict(mask=mask, sigma=sigma, errors=set())
    if mask_only:
        return res
    H, W = mask.shape
    fg = optimise(mask, image, 3)
    if fg_downsample is not None:
        wd, hd = int(W / fg_downsample), int(H / fg_downsample)
        if wd > 0 and hd > 0:
            fg = thumbnail(fg, wd, hd)
        else:
            res['errors'].add('too-small-to-downsample')
    bg = optimise(~mask, image, 10)
    if bg_downsample is not None:
        wd, hd = int(W / bg_downsample), int(H / bg_downsample)
        if wd > 0 and hd > 0:
            bg = thumbnail(bg, wd, hd)
        else:
            res['errors'].add('too-small-to-downsample')
    res['fg'], res['bg'] = fg, bg
    return res
