#!/usr/bin/env python
"""Per-stage timing of the staged pipeline (CUDA events), for quick A/B runs:
    B200MRC_IIRW_MODE=trio python tools/time_stages.py --pages 64      (tuning knobs: INTEGRATION.md section 6)"""
import argparse, os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pages', type=int, default=64)
    ap.add_argument('--distinct', type=int, default=4)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--check', action='store_true')
    a = ap.parse_args()
    import archive_pdf_tools_b200.synth as synth
    H, W = 3300, 2550
    distinct = [synth.make_page(i, H, W, dpi=400) for i in range(a.distinct)]
    import torch
    import archive_pdf_tools_b200 as pkg
    pages = np.stack([distinct[i % a.distinct] for i in range(a.pages)])
    eng = pkg.get_engine()
    batch = eng.make_batch(a.pages, H, W, 3, bg_downsample=3)
    batch.img.upload(pages, non_blocking=False)
    for _ in range(2):
        batch.run_staged(101, denoise_mask='fast')
    torch.cuda.synchronize()
    acc = {}
    for _ in range(a.steps):
        ev = {}
        batch.run_staged(101, denoise_mask='fast', events=ev)
        torch.cuda.synchronize()
        for k, (s, e) in ev.items():
            acc.setdefault(k, []).append(s.elapsed_time(e))
    out = {k: round(float(np.mean(v)), 3) for k, v in acc.items()}
    out['total'] = round(sum(out.values()), 3)
    out['env'] = {k: v for k, v in os.environ.items() if k.startswith('B200MRC_')}
    if a.check:
        from oracle import oracle as orc
        exp = orc.decompose(pages[1], dpi=400, bg_downsample=3, denoise_mask='fast')
        out['check'] = bool(np.array_equal(batch.mask.numpy(np.bool_)[1], exp['mask']) and
                            np.array_equal(batch.fg.numpy()[1], exp['fg']) and np.array_equal(batch.bg.numpy()[1], exp['bg']))
    print(json.dumps(out))

if __name__ == '__main__':
    main()
