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

_engines = {}


def get_engine(device=None):
    """The engine of `device` (default: the current CUDA device); one per GPU, so a process that drives several
    GPUs gets an engine for each of them."""
    E._require_cuda()
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    eng = _engines.get(idx)
    if eng is None:
        eng = _engines[idx] = E.MrcEngine('cuda:%d' % idx)
    return eng


MAX_BLUR_SIGMA_EST = (128.5 / 4) / 0.1          # blur radius int(4 * 0.1 * sigma_est + 0.5) must stay <= 128


def _check_sigma(sigmas):
    """The pre-blur kernels cover radius <= 128 (sigma_est < 321: far beyond any scan); anything larger would leave the
    gray plane unwritten, so it is an error rather than a silent garbage mask."""
    bad = [float(v) for v in np.atleast_1d(sigmas) if v >= MAX_BLUR_SIGMA_EST]
    if bad:
        raise L.B200MrcError('estimated noise sigma %.1f needs a Gaussian pre-blur of radius > 128, which the engine does '
                             'not implement' % bad[0])


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


def _hocr_choice(ratio, inv_ratio, sigmas):
    """The polarity rule of mrc.py:238-263: 0 = leave the mask alone, 1 = thres, 2 = thres_invert."""
    if ratio < 0.3 or inv_ratio < 0.3:
        if inv_ratio > 0.2 and ratio < 0.2:
            return 1
        ratio_sigma, inv_ratio_sigma = sigmas
        # Prefer ratio over inv_ratio by a bit
        if inv_ratio < 0.3 and inv_ratio < ratio and \
                (inv_ratio_sigma < ratio_sigma or (ratio_sigma < 0.1 and inv_ratio_sigma < 0.1)):
            return 2
        elif ratio < 0.2:
            return 1
    return 0


def _needs_sigma(ratio, inv_ratio):
    return (ratio < 0.3 or inv_ratio < 0.3) and not (inv_ratio > 0.2 and ratio < 0.2)


def _iter_text_lines(hocr_word_data, image_width, image_height, downsample):
    """Integer boxes (left, top, right, bottom) of the hOCR text lines create_hocr_mask looks at (mrc.py:194-221):
    lines with text and a mean word confidence of at least 20, scaled by 1/downsample, truncated to int; empty boxes are
    dropped silently, inverted boxes and boxes leaving the page with the reference's diagnostics on stderr."""
    import sys
    for paragraph in hocr_word_data or ():
        for line in paragraph['lines']:
            words = line['words']
            confs = [w['confidence'] for w in words]
            mean_conf = sum(confs) / len(confs) if confs else 0
            if mean_conf < 20 or not ' '.join(w['text'] for w in words).strip():
                continue
            box = [int(v) if downsample is None else int(v / downsample) for v in line['bbox']]
            left, top, right, bottom = box
            if left == right or top == bottom:
                continue
            if left >= right or top >= bottom:
                print('Invalid bounding box: (%d, %d, %d, %d)' % tuple(box), file=sys.stderr)
            elif left < 0 or top < 0 or right > image_width or bottom > image_height:
                print('Invalid bounding box outside image: (%d, %d, %d, %d)' % tuple(box), file=sys.stderr)
            else:
                yield tuple(box)


def _boxes_overlap(boxes):
    """True if any two (left, top, right, bottom) boxes intersect (sweep over the sorted left edges)."""
    order = sorted(range(len(boxes)), key=lambda i: boxes[i][0])
    active = []
    for i in order:
        l, t, r, b = boxes[i]
        active = [j for j in active if boxes[j][2] > l]
        for j in active:
            if boxes[j][1] < b and t < boxes[j][3]:
                return True
        active.append(i)
    return False


def create_hocr_mask(img, mask_arr, hocr_word_data, downsample=None, dpi=None, timing_data=None):
    """mrc.py:188-270 on the device.  img: gray Plane (the plain 'L' page, n = 1); mask_arr: mask Plane, modified in
    place.  Per text line: Sauvola (k = 0.1) on the crop and on the inverted crop, polarity picked by fill ratio and,
    for the undecided lines, by mean_estimate_sigma of the two boolean results; the winner is pasted into the mask.
    All line crops of the page go through the device TOGETHER: one gather launch (crops -> aligned scratch), one
    Sauvola launch over both polarities of every crop (b200mrc_sauvola_items), one count launch, one sigma launch for
    the undecided lines and one paste launch -- five launches and two small device-to-host reads per page, whatever
    the number of lines.  The host only takes the per-line decisions (the reference's rule, _hocr_choice).  Pastes of
    overlapping boxes keep the reference's loop order (they are then issued one by one)."""
    import ctypes as C
    t = time()
    lines = list(_iter_text_lines(hocr_word_data, img.w, img.h, downsample))
    if lines:
        eng = get_engine()
        lib = L.lib()
        st = E._stream_ptr()
        window = E.window_for_dpi(dpi)
        k = 0.1          # the reference's choice; its ratio / sigma limits below are tuned to it (mrc.py:228)
        # one scratch buffer: per line an aligned copy of the crop and the two threshold results
        geo, off = [], 0
        for (left, top, right, bottom) in lines:
            w, h = right - left, bottom - top
            pitch = (w + 15) // 16 * 16
            size = (pitch * h + 255) // 256 * 256
            geo.append((w, h, pitch, off, off + size, off + 2 * size))
            off += 3 * size
        scratch = torch.empty(off, dtype=torch.uint8, device=eng.device)
        base = scratch.data_ptr()
        max_w, max_h = max(g[0] for g in geo), max(g[1] for g in geo)

        def upload(structs):
            """ctypes array of descriptor structs -> device tensor (one small H2D copy)."""
            return torch.frombuffer(bytearray(bytes(structs)), dtype=torch.uint8).to(eng.device)

        # ---- gather the crops (aligned rows for the Sauvola kernel) and threshold both polarities, one launch each
        crops = (L.CopyRect * len(lines))()
        items = (L.SauvolaItem * (2 * len(lines)))()
        for i, ((left, top, right, bottom), (w, h, pitch, o_in, o_th, o_ti)) in enumerate(zip(lines, geo)):
            c = crops[i]
            c.src, c.src_pitch, c.dst, c.dst_pitch, c.width, c.height = img.t.data_ptr() + top * img.pitch + left, img.pitch, base + o_in, pitch, w, h
            for j, (o_out, flags) in enumerate(((o_th, 0), (o_ti, L.SAUVOLA_INVERT_INPUT))):
                it = items[2 * i + j]
                it.in_, it.in_pitch, it.out, it.out_pitch, it.width, it.height, it.flags = base + o_in, pitch, base + o_out, pitch, w, h, flags
        crops_d, items_d = upload(crops), upload(items)
        L.check(lib.b200mrc_rects_copy(C.c_void_p(crops_d.data_ptr()), len(lines), max_w, max_h, st), 'b200mrc_rects_copy')
        L.check(lib.b200mrc_sauvola_items(C.c_void_p(items_d.data_ptr()), 2 * len(lines), max_w, max_h, window, window, k, 128.0, st),
                'b200mrc_sauvola_items')

        def rects(entries):                                  # entries: (ptr, pitch, w, h, keys_ptr)
            arr = (L.Rect * len(entries))()
            for r, (ptr, pitch, w, h, keys) in zip(arr, entries):
                r.ptr, r.pitch, r.width, r.height, r.keys = ptr, pitch, w, h, keys
            return upload(arr)

        entries = []
        for (w, h, pitch, o_in, o_th, o_ti) in geo:
            entries += [(base + o_th, pitch, w, h, 0), (base + o_ti, pitch, w, h, 0)]
        rd = rects(entries)
        counts = torch.empty(len(entries), dtype=torch.int32, device=eng.device)
        L.check(lib.b200mrc_rects_count_nonzero(C.c_void_p(rd.data_ptr()), len(entries), C.c_void_p(counts.data_ptr()), st),
                'b200mrc_rects_count_nonzero')
        counts = counts.cpu().tolist()
        ratios = []
        for i, (w, h, *_r) in enumerate(geo):
            size = w * h
            ones, ones_i = counts[2 * i], counts[2 * i + 1]
            ratios.append((ones / ((size - ones) + ones), ones_i / ((size - ones_i) + ones_i)))
        need = [i for i, (r, ri) in enumerate(ratios) if _needs_sigma(r, ri)]
        sig = {}
        if need:
            koff, kentries = 0, []
            for i in need:
                w, h, pitch, o_in, o_th, o_ti = geo[i]
                nk = ((h + 3) // 2) * ((w + 3) // 2) * 8
                kentries += [(base + o_th, pitch, w, h, koff), (base + o_ti, pitch, w, h, koff + nk)]
                koff += 2 * nk
            keys = torch.empty(max(koff, 8), dtype=torch.uint8, device=eng.device)
            kentries = [(p_, pi, w, h, keys.data_ptr() + ko) for (p_, pi, w, h, ko) in kentries]
            kd = rects(kentries)
            sg = torch.empty(len(kentries), dtype=torch.float64, device=eng.device)
            L.check(lib.b200mrc_rects_sigma_bool(C.c_void_p(kd.data_ptr()), len(kentries), C.c_void_p(sg.data_ptr()), st),
                    'b200mrc_rects_sigma_bool')
            sg = sg.cpu().tolist()
            for j, i in enumerate(need):
                sig[i] = (sg[2 * j], sg[2 * j + 1])
        # ---- pastes: mask_arr[top:bottom, left:right] = th for the lines that got a polarity
        chosen = []
        for i, ((left, top, right, bottom), (w, h, pitch, o_in, o_th, o_ti)) in enumerate(zip(lines, geo)):
            c = _hocr_choice(ratios[i][0], ratios[i][1], sig.get(i))
            if c:
                chosen.append((mask_arr.t.data_ptr() + top * mask_arr.pitch + left, base + (o_th if c == 1 else o_ti), pitch, w, h, lines[i]))
        if chosen and not _boxes_overlap([c[5] for c in chosen]):
            pastes = (L.CopyRect * len(chosen))()
            for p_, (dst, src, pitch, w, h, _box) in zip(pastes, chosen):
                p_.src, p_.src_pitch, p_.dst, p_.dst_pitch, p_.width, p_.height = src, pitch, dst, mask_arr.pitch, w, h
            pastes_d = upload(pastes)
            L.check(lib.b200mrc_rects_copy(C.c_void_p(pastes_d.data_ptr()), len(chosen), max(c[3] for c in chosen), max(c[4] for c in chosen), st),
                    'b200mrc_rects_copy')
        else:
            for (dst, src, pitch, w, h, _box) in chosen:     # overlapping boxes: the later line wins, as in the reference's loop
                L.check(lib.b200mrc_copy2d(C.c_void_p(dst), mask_arr.pitch, C.c_void_p(src), pitch, w, h, L.COPY_D2D, st), 'b200mrc_copy2d')
        torch.cuda.current_stream().synchronize()           # scratch and descriptor buffers stay alive until the pastes are done
    if timing_data is not None:
        timing_data.append(('hocr_mask_gen', time() - t))
    return len(lines)


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
    if page is not None:
        src = E.Plane(1, height_, width_, page.shape[2] if page.ndim == 3 else 1, eng.device).upload(page[None])
    else:
        src = E.Plane(1, height_, width_, 1, eng.device).upload(gray_host[None])

    # ---- hOCR line masks (mrc.py:367-370) on the plain gray page
    n_lines = 0
    if hocr_word_data:
        mask.t.zero_()                                       # mask_arr = np.array(Image.new('1', size))
        gray_plain = src if src.c == 1 else E.Plane(1, height_, width_, 1, eng.device)
        if src.c != 1:
            eng.gray_blur(src, gray_plain, None)
        n_lines = create_hocr_mask(gray_plain, mask, hocr_word_data, downsample=downsample, dpi=dpi, timing_data=timing_data)
    elif timing_data is not None:
        timing_data.append(('hocr_mask_gen', 0.0))

    # ---- create_threshold_mask (mrc.py:300-329)
    t = time()
    sigma_dev = eng.estimate_noise(src)
    sigma_est = float(sigma_dev.cpu()[0])              # synchronises: 'est_1' is a true stage time
    if timing_data is not None:
        timing_data.append(('est_1', time() - t))
    _check_sigma(sigma_est)
    if sigma_est > 1.0 and timing_data is not None:
        timing_data.append(('blur_1', 0.0))                # the pre-blur is fused into the threshold pass below
    t = time()
    eng.threshold_mask(src, mask, E.window_for_dpi(dpi), k=0.34, R=128.0, sigma_dev=sigma_dev if sigma_est > 1.0 else None,
                       flags=L.SAUVOLA_OR_INTO if n_lines else 0)                                       # mask_arr |= thres_arr
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


def downsample_image(image, downsample):
    """The page pre-step of recode.py:368-372 -- image.thumbnail((w/downsample, h/downsample), resample=LANCZOS,
    reducing_gap=None) -- on the device.  image: PIL 'L' / 'RGB'; returns a new PIL image (Pillow resizes in place)."""
    from PIL import Image
    w, h = image.size
    if image.mode not in ('L', 'RGB'):
        # the reference thumbnails in the image's own mode (Pillow picks NEAREST for '1' / 'P'); rare: left to Pillow
        image = image.copy()
        image.thumbnail((w / downsample, h / downsample), resample=Image.LANCZOS, reducing_gap=None)
        return image
    out = get_engine().thumbnail_np(np.asarray(image), w / downsample, h / downsample, reducing_gap=None, filter=E.LANCZOS)
    return Image.fromarray(out)


def packed_mask(mask_arr, invert=False):
    """Boolean mask (H x W ndarray) -> (bytes of the PIL mode-'1' image, PIL image): what encode_mrc_mask
    (mrc.py:474-520) builds with Image.fromarray(np_mask); `invert` is recode.py:408's np_mask ^ True."""
    from PIL import Image
    m = np.ascontiguousarray(mask_arr).view(np.uint8)
    h, w = m.shape
    eng = get_engine()
    plane = E.Plane(1, h, w, 1, eng.device).upload(m[None], non_blocking=False)
    data = eng.pack_mask(plane, invert=invert).cpu().numpy()[0].tobytes()
    return data, Image.frombytes('1', (w, h), data)


def decompose_pages(pages, dpi=None, window=None, bg_downsample=None, fg_downsample=None,
                    denoise_mask=DENOISE_FAST, mask_only=False, sigma=None, batch=None,
                    hocr_word_data=None, downsample=None):
    """Batched form of create_mrc_hocr_components for N equally-shaped pages held in HOST memory
    (uint8 ndarray [N,H,W] or [N,H,W,3], or a pinned CPU tensor): one H2D copy, one
    b200mrc_decompose, D2H of the results.  Returns dict(mask, fg, bg, sigma, errors).
    `batch`: a DecomposeBatch to reuse (device buffers + workspace).  `hocr_word_data`: optional list with one
    hocr_word_data structure per page (empty / None entries for pages without text boxes); those pages get their
    hOCR line masks (create_hocr_mask) before the page threshold is OR-ed in."""
    eng = get_engine()
    shape = tuple(pages.shape)
    n, h, w = shape[:3]
    c = shape[3] if len(shape) == 4 else 1
    if batch is None:
        batch = eng.make_batch(n, h, w, c, bg_downsample, fg_downsample, mask_only)
    batch.img.upload(pages)
    or_into = False
    if hocr_word_data is not None and any(hocr_word_data):
        assert len(hocr_word_data) == n
        batch.mask.t.zero_()
        or_into = True
        gray = E.Plane(1, h, w, 1, eng.device) if c != 1 else None
        for i, hd in enumerate(hocr_word_data):
            if not hd:
                continue
            page_img = E.PageView(batch.img, i)
            if c != 1:
                eng.gray_blur(page_img, gray, None)          # the plain 'L' page (mrc.py:358-363)
            create_hocr_mask(gray if c != 1 else page_img, E.PageView(batch.mask, i), hd, downsample=downsample, dpi=dpi)
    batch.run(window if window is not None else E.window_for_dpi(dpi), denoise_mask=denoise_mask, sigma=sigma, or_into_mask=or_into)
    res = dict(mask=batch.mask.numpy(np.bool_), sigma=batch.sigma.cpu().numpy(), errors=set(batch.errors))
    _check_sigma(res['sigma'])
    if not mask_only:
        res['fg'] = batch.fg.numpy()
        res['bg'] = batch.bg.numpy()
    return res
