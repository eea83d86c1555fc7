#!/usr/bin/env python
"""Mean per-kernel device time (C-ABI event profiling) of the staged pipeline on N synthetic 400-DPI pages."""
import argparse, os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pages', type=int, default=64)
    ap.add_argument('--distinct', type=int, default=4)
    ap.add_argument('--steps', type=int, default=5)
    a = ap.parse_args()
    import torch
    import archive_pdf_tools_b200 as pkg
    from archive_pdf_tools_b200 import _lib, synth
    H, W = 3300, 2550
    distinct = [synth.make_page(i, H, W, dpi=400) for i in range(a.distinct)]
    pages = np.stack([distinct[i % a.distinct] for i in range(a.pages)])
    eng = pkg.get_engine()
    b = eng.make_batch(a.pages, H, W, 3, bg_downsample=3)
    b.img.upload(pages, non_blocking=False)
    for _ in range(2):
        b.run_staged(101)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(a.steps):
        b.run_staged(101)
    torch.cuda.synchronize()
    rep = {k: round(v[1] / v[0], 3) for k, v in _lib.profile_report().items()}
    rep['total'] = round(sum(rep.values()), 3)
    rep['mask_fraction'] = round(float(b.mask.view().float().mean().item()), 4)
    rep['env'] = {k: v for k, v in os.environ.items() if k.startswith('B200MRC_')}
    print(json.dumps(rep))

if __name__ == '__main__':
    main()
