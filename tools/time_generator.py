#!/usr/bin/env python
"""The reference's own calling pattern, one page at a time: create_mrc_hocr_components(PIL image, hocr_word_data, ...)
(mrc.py:334-471) through this engine -- wall ms per 3300x2550 page from the PIL image to the three numpy arrays on the
host (H2D, all stages, three D2H yields), and for the mask-only consumer (recode.py:400-407: first yield only)."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pages', type=int, default=24)
    ap.add_argument('--gray', action='store_true')
    a = ap.parse_args()
    import torch
    from PIL import Image
    import archive_pdf_tools_b200 as pkg
    from archive_pdf_tools_b200 import synth
    imgs = [Image.fromarray(synth.make_page(i, 3300, 2550, dpi=400, rgb=not a.gray)) for i in range(4)]
    hocr = synth.page_hocr(3300, 2550, dpi=400)
    out = {}
    for label, words, mask_only in (('full', [], False), ('full_with_hocr_lines', hocr, False), ('mask_only', [], True)):
        for rep in range(a.pages + 3):
            if rep == 3:
                torch.cuda.synchronize(); t0 = time.time()
            gen = pkg.create_mrc_hocr_components(imgs[rep % 4], words, dpi=400, bg_downsample=3, denoise_mask='fast')
            mask = next(gen)
            if mask_only:
                gen.close()
            else:
                fg = next(gen); bg = next(gen)
        torch.cuda.synchronize()
        out[label] = round((time.time() - t0) / a.pages * 1e3, 2)
    out['unit'] = 'ms per 3300x2550 %s page, host PIL image -> host numpy arrays' % ('gray' if a.gray else 'RGB')
    out['hocr_lines'] = len(hocr)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
