#!/usr/bin/env python
"""A/B runs of kernel variants in ONE process: the C-ABI reads its B200MRC_* knobs at every launch, so each
variant is an environment dict applied between runs of the staged pipeline (64 synthetic 400-DPI pages by default).
Prints one JSON line per variant: mean device ms per kernel (C-ABI event profiling) and the step total.

  python tools/ab_kernels.py '{}' '{"B200MRC_IIRW_MODE":"single"}' ...
"""
import argparse, os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pages', type=int, nargs='+', default=[64])
    ap.add_argument('--distinct', type=int, default=4)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('variants', nargs='*', default=['{}'])
    a = ap.parse_args()
    import torch
    import archive_pdf_tools_b200 as pkg
    from archive_pdf_tools_b200 import _lib, synth
    H, W = 3300, 2550
    distinct = [synth.make_page(i, H, W, dpi=400) for i in range(a.distinct)]
    eng = pkg.get_engine()
    for npages in a.pages:
      pages = np.stack([distinct[i % a.distinct] for i in range(npages)])
      b = eng.make_batch(npages, H, W, 3, bg_downsample=3)
      b.img.upload(pages, non_blocking=False)
      ref = None
      for v in a.variants:
          env = json.loads(v)
          for k in [k for k in os.environ if k.startswith('B200MRC_')]:
              del os.environ[k]
          os.environ.update(env)
          for _ in range(2):
              b.run_staged(101)
          torch.cuda.synchronize()
          _lib.profile_enable(True)
          for _ in range(a.steps):
              b.run_staged(101)
          torch.cuda.synchronize()
          rep = {k: round(val[1] / val[0], 3) for k, val in _lib.profile_report().items()}
          _lib.profile_enable(False)
          rep['total'] = round(sum(rep.values()), 3)
          # results must not depend on the variant: checksum of mask / fg / bg against the first variant
          sums = [int((t.view().to(torch.int32) * (1 + torch.arange(t.view().shape[-1], device=t.t.device, dtype=torch.int32) % 251)).sum().item()) for t in (b.mask, b.fg, b.bg)]
          if ref is None:
              ref = sums
          rep['same_as_first'] = sums == ref
          rep['env'] = env; rep['pages'] = npages
          print(json.dumps(rep), flush=True)


if __name__ == '__main__':
    main()
