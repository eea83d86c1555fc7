#!/usr/bin/env python
"""b200mrc_threshold_mask alone (create_threshold_mask: gray -> conditional blur -> Sauvola) on one batch with INJECTED
per-page sigma_est patterns: the fused kernel against the two-pass form, device ms per call (CUDA events).
  python tools/time_threshold.py [--pages 64] [--gray] [--shape H W] [--window 101]"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pages', type=int, default=64)
    ap.add_argument('--shape', type=int, nargs=2, default=[3300, 2550])
    ap.add_argument('--window', type=int, default=101)
    ap.add_argument('--gray', action='store_true')
    ap.add_argument('--patterns', type=str, default='0.5;3.4;5;8;21;3.4,5;3.4,21;3.4,5,21,0.5')
    a = ap.parse_args()
    import torch
    import archive_pdf_tools_b200 as pkg
    from archive_pdf_tools_b200 import _lib, synth, engine as E
    H, W = a.shape
    C = 1 if a.gray else 3
    eng = pkg.get_engine()
    distinct = [synth.make_page(i, H, W, dpi=400, rgb=not a.gray) for i in range(4)]
    pages = np.stack([distinct[i % 4] for i in range(a.pages)])
    src = E.Plane(a.pages, H, W, C, eng.device).upload(pages, non_blocking=False)
    dst = E.Plane(a.pages, H, W, 1, eng.device)
    for pat in a.patterns.split(';'):
        vals = [float(v) for v in pat.split(',')]
        sig = torch.tensor([vals[i % len(vals)] for i in range(a.pages)], dtype=torch.float64, device=eng.device)
        row = {'sigma_pattern': vals}
        ref = None
        for path in ('fused', 'legacy'):
            _lib.set_tuning('THRESHOLD_PATH', path)
            for _ in range(2):
                eng.threshold_mask(src, dst, a.window, sigma_dev=sig)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                eng.threshold_mask(src, dst, a.window, sigma_dev=sig)
            e1.record(); torch.cuda.synchronize()
            row[path + '_ms'] = round(e0.elapsed_time(e1) / 4, 3)
            chk = int(dst.view().to(torch.int64).sum().item())
            ref = chk if ref is None else ref
            row['same'] = chk == ref
        print(json.dumps(row), flush=True)
    _lib.set_tuning('THRESHOLD_PATH', 'auto')


if __name__ == '__main__':
    main()
