#!/usr/bin/env python
"""A/B runs of kernel / scheduling variants in ONE process.  A variant is a JSON dict of tuning knobs
(b200mrc_set_tuning names), applied before its runs and reset afterwards.  Per variant one JSON line:
device ms per step (CUDA events around the steps), mean ms per kernel launch (C-ABI event profiling; with page groups
on several streams these overlap, so they do not add up to the step) and whether mask / fg / bg checksums equal the
first variant's.

  python tools/ab_kernels.py --call decompose '{}' '{"DECOMPOSE_GROUPS":1}' '{"DECOMPOSE_GROUPS":4,"DECOMPOSE_STREAMS":2}'
  python tools/ab_kernels.py --call staged '{}' '{"THRESHOLD_PATH":"legacy"}'
"""
import argparse, os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pages', type=str, default='64', help='batch sizes, comma separated')
    ap.add_argument('--distinct', type=int, default=4)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--shape', type=int, nargs=2, default=[3300, 2550])
    ap.add_argument('--dpi', type=int, default=400)
    ap.add_argument('--gray', action='store_true')
    ap.add_argument('--mask-only', action='store_true')
    ap.add_argument('--halftone', action='store_true')
    ap.add_argument('--sigma-n', type=float, default=3.0, help='sensor noise of the synthetic pages (decides the blur radius)')
    ap.add_argument('--call', choices=['decompose', 'staged'], default='decompose',
                    help='decompose: one b200mrc_decompose per step (page groups on internal streams); staged: one C-ABI call per stage')
    ap.add_argument('variants', nargs='*', default=['{}'])
    a = ap.parse_args()
    import torch
    import archive_pdf_tools_b200 as pkg
    from archive_pdf_tools_b200 import _lib, synth
    H, W = a.shape
    C = 1 if a.gray else 3
    window = pkg.window_for_dpi(a.dpi)
    distinct = [synth.make_page(i, H, W, dpi=a.dpi, rgb=not a.gray, sigma_n=a.sigma_n, halftone=a.halftone and i % 3 == 0) for i in range(a.distinct)]
    eng = pkg.get_engine()
    for npages in [int(v) for v in a.pages.split(',')]:
        pages = np.stack([distinct[i % a.distinct] for i in range(npages)])
        ref = None
        for v in a.variants:
            knobs = json.loads(v)
            saved = {k: _lib.get_tuning(k) for k in knobs}
            for k, val in knobs.items():
                _lib.set_tuning(k, val)
            # a fresh batch per variant: the page-group layout of the workspace depends on the knobs
            b = eng.make_batch(npages, H, W, C, bg_downsample=None if a.mask_only else 3, mask_only=a.mask_only)
            b.img.upload(pages, non_blocking=False)
            step = (lambda: b.run(window)) if a.call == 'decompose' else (lambda: b.run_staged(window))
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            rep = {'ms_per_step': round(e0.elapsed_time(e1) / a.steps, 3)}
            _lib.profile_enable(True)
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            rep['kernel_ms'] = {k: round(val[1] / val[0], 3) for k, val in _lib.profile_report().items()}
            _lib.profile_enable(False)
            outs = (b.mask,) if a.mask_only else (b.mask, b.fg, b.bg)
            sums = [int((t.view().to(torch.int64) * (1 + torch.arange(t.view().shape[-1], device=t.t.device, dtype=torch.int64) % 251)).sum().item())
                    for t in outs]
            if ref is None:
                ref = sums
            rep['same_as_first'] = sums == ref
            rep['knobs'] = knobs; rep['pages'] = npages; rep['call'] = a.call
            print(json.dumps(rep), flush=True)
            for k, val in saved.items():
                _lib.set_tuning(k, val)
            del b
            torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
