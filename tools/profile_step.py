#!/usr/bin/env python
"""Small driver for ncu: runs b200mrc_decompose (one call per step) on N synthetic 400-DPI pages (see bench.py).
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \\
        python tools/profile_step.py --pages 64 --steps 2 --warmup 1
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pages', type=int, default=8)
    ap.add_argument('--distinct', type=int, default=4)
    ap.add_argument('--steps', type=int, default=1)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--mask-only', action='store_true')
    a = ap.parse_args()
    import archive_pdf_tools_b200.synth as synth
    H, W = 3300, 2550
    distinct = [synth.make_page(i, H, W, dpi=400) for i in range(a.distinct)]
    import torch
    import archive_pdf_tools_b200 as pkg
    pages = np.stack([distinct[i % a.distinct] for i in range(a.pages)])
    eng = pkg.get_engine()
    batch = eng.make_batch(a.pages, H, W, 3, bg_downsample=3, mask_only=a.mask_only)
    batch.img.upload(pages, non_blocking=False)
    for _ in range(a.warmup + a.steps):
        batch.run(101, denoise_mask='fast')
    torch.cuda.synchronize()
    print('done', float(batch.sigma.cpu()[0]))


if __name__ == '__main__':
    main()
