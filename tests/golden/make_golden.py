#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference inside the build container.

Needs /root/reference (read-only) -- it imports internetarchivepdf/mrc.py with stub fitz/skimage
(oracle/ref_pipeline.load_reference_mrc) on top of the reference's own compiled Cython
(oracle/_ref) and the real Pillow / scipy, and records, for a few small synthetic pages:
  * the three arrays create_mrc_hocr_components yields (mask, fg, bg),
  * the sigma_est the run used (scikit-image is not installed: estimate_sigma is the oracle
    restatement -- "parity unpinned" -- so sigma is stored and can be injected),
  * threshold_image outputs for the BASELINE config-1 window (33) and the hOCR k (0.1).
The pages themselves are regenerated from seeds by archive-pdf-tools_b200/synth.py.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from PIL import Image                                   # noqa: E402
import archive_pdf_tools_b200.synth as synth            # noqa: E402
from oracle import ref_pipeline as rp                   # noqa: E402
from oracle import oracle as orc                        # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, index, H, W, dpi, rgb, sigma_n, halftone, bg_downsample, fg_downsample, denoise
    ('rgb_clean_bg3', 1, 330, 255, 100, True, 0.5, False, 3, None, 'fast'),
    ('rgb_noisy_bg3', 2, 330, 255, 100, True, 6.0, False, 3, None, 'fast'),
    ('gray_noisy_bg2_fg2', 3, 297, 211, 150, False, 4.0, True, 2, 2, 'fast'),
    ('rgb_halftone_nodenoise_bg4', 4, 400, 320, 100, True, 3.0, True, 4, None, 'none'),
    ('rgb_nods', 5, 160, 200, None, True, 2.0, False, None, None, 'fast'),
    ('gray_flat_noblur_bg3', 6, 220, 180, 100, False, 0.0, False, 3, None, 'fast'),
]

# create_hocr_mask cases (mrc.py:188-270): hocr_word_data = synth.page_hocr(H, W, dpi, scale=downsample or 1)
# name, index, H, W, dpi, rgb, sigma_n, (invert_lines, noisy_dark_lines), downsample, bg_downsample, denoise
HOCR_CASES = [
    ('hocr_rgb_clean_bg3', 7, 330, 255, 100, True, 0.5, ((2, 5), (7,)), None, 3, 'fast'),
    ('hocr_gray_noisy', 8, 400, 320, 150, False, 3.0, ((1,), (3, 6)), None, None, 'fast'),
    ('hocr_rgb_downsample2', 9, 300, 240, 100, True, 1.0, ((3,), (1,)), 2, 3, 'none'),
    ('hocr_gray_dpinone', 10, 260, 300, None, False, 0.0, ((0, 4), (2, 8)), None, None, 'fast'),
]


def main():
    mrc = rp.load_reference_mrc()
    if mrc is None:
        print('reference not available; nothing generated')
        return 1
    for (name, idx, H, W, dpi, rgb, sn, ht, bgd, fgd, den) in CASES:
        page = synth.make_page(idx, H, W, dpi=dpi or 200, rgb=rgb, sigma_n=sn, halftone=ht)
        im = Image.fromarray(page)
        timing, errors = [], set()
        gen = mrc.create_mrc_hocr_components(im, [], dpi=dpi, bg_downsample=bgd, fg_downsample=fgd,
                                             denoise_mask=den, timing_data=timing, errors=errors)
        mask = next(gen).copy(); fg = next(gen).copy(); bg = next(gen).copy()
        gray = page if not rgb else np.array(im.convert('L'))
        sigma = orc.estimate_noise(gray)
        # the restated glue (oracle/ref_pipeline.ref_decompose) must agree with the imported reference
        rd = rp.ref_decompose(page, dpi=dpi, bg_downsample=bgd, fg_downsample=fgd, denoise_mask=den)
        assert np.array_equal(rd['mask'], mask) and np.array_equal(rd['fg'], fg) and np.array_equal(rd['bg'], bg), name
        # ... and so must the C restatement
        od = orc.decompose(page, dpi=dpi, bg_downsample=bgd, fg_downsample=fgd, denoise_mask=den)
        assert np.array_equal(od['mask'], mask) and np.array_equal(od['fg'], fg) and np.array_equal(od['bg'], bg), name
        t33 = mrc.threshold_image(gray, 132)             # int(132/4) = 33: BASELINE config 1 window
        t01 = mrc.threshold_image(gray, dpi, 0.1)        # hOCR line k
        np.savez_compressed(os.path.join(OUT, name + '.npz'),
                            params=np.array([idx, H, W, dpi or -1, int(rgb), ht, bgd or -1, fgd or -1], np.int64),
                            sigma_n=np.float64(sn), denoise=np.array(den), sigma=np.float64(sigma),
                            mask=np.packbits(mask), fg=fg, bg=bg, t33=np.packbits(t33), t01=np.packbits(t01),
                            timing_keys=np.array([k for k, _ in timing]))
        print(name, 'sigma=%.4f' % sigma, 'mask=%.3f' % mask.mean(), fg.shape, bg.shape, [k for k, _ in timing])
    for (name, idx, H, W, dpi, rgb, sn, inv, ds, bgd, den) in HOCR_CASES:
        page = synth.make_page(idx, H, W, dpi=dpi or 100, rgb=rgb, sigma_n=sn, invert_lines=inv[0], noisy_dark_lines=inv[1])
        hocr = synth.page_hocr(H, W, dpi=dpi or 100, scale=float(ds or 1))
        im = Image.fromarray(page)
        timing, errors = [], set()
        gen = mrc.create_mrc_hocr_components(im, hocr, dpi=dpi, downsample=ds, bg_downsample=bgd,
                                             denoise_mask=den, timing_data=timing, errors=errors)
        mask = next(gen).copy(); fg = next(gen).copy(); bg = next(gen).copy()
        gray = page if not rgb else np.array(im.convert('L'))
        # the hOCR mask on its own (before the page threshold is OR-ed in)
        hm = np.zeros(gray.shape, bool)
        mrc.create_hocr_mask(Image.fromarray(gray), hm, hocr, downsample=ds, dpi=dpi)
        plain = mrc.create_mrc_hocr_components(im, [], dpi=dpi, downsample=ds, bg_downsample=bgd, denoise_mask=den)
        mask_plain = next(plain).copy()
        od = orc.decompose(page, dpi=dpi, bg_downsample=bgd, denoise_mask=den, hocr_word_data=hocr, downsample=ds)
        assert np.array_equal(od['mask'], mask) and np.array_equal(od['fg'], fg) and np.array_equal(od['bg'], bg), name
        ohm = orc.hocr_mask(gray, np.zeros(gray.shape, bool), hocr, downsample=ds, dpi=dpi)
        assert np.array_equal(ohm, hm), name
        np.savez_compressed(os.path.join(OUT, name + '.npz'),
                            params=np.array([idx, H, W, dpi or -1, int(rgb), ds or -1, bgd or -1], np.int64),
                            invert_lines=np.array(inv[0], np.int64), noisy_dark_lines=np.array(inv[1], np.int64), sigma_n=np.float64(sn), denoise=np.array(den),
                            mask=np.packbits(mask), hocr_mask=np.packbits(hm), fg=fg, bg=bg,
                            timing_keys=np.array([k for k, _ in timing]))
        print(name, 'hocr mask=%.4f' % hm.mean(), 'mask=%.4f' % mask.mean(), 'differs from plain in %d px' % (mask != mask_plain).sum(),
              [k for k, _ in timing])
    return 0


if __name__ == '__main__':
    sys.exit(main())
