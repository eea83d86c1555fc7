#!/usr/bin/env python
"""Host-to-host throughput of StreamedDecomposer over its knobs, one JSON line per setting:
chunk size, device buffers, compute streams, mask transport ('bool' plane over the bus / 'packed' rows + host unpack /
'handoff' = packed rows returned as they are), unpack workers.  Two batches in flight, like bench.py's e2e.

  python tools/e2e_sweep.py --pages 64 --steps 6 4:4:2:bool 4:4:2:packed:6 8:4:2:packed:6 4:4:2:handoff
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pages', type=int, default=64)
    ap.add_argument('--steps', type=int, default=6)
    ap.add_argument('--distinct', type=int, default=8)
    ap.add_argument('settings', nargs='+', help='chunk:buffers:streams:transport[:workers]')
    a = ap.parse_args()
    import torch
    import archive_pdf_tools_b200 as pkg
    from archive_pdf_tools_b200 import synth
    from archive_pdf_tools_b200.engine import StreamedDecomposer
    H, W, C = 3300, 2550, 3
    distinct = [synth.make_page(i, H, W, dpi=400) for i in range(a.distinct)]
    host = torch.from_numpy(np.stack([distinct[i % a.distinct] for i in range(a.pages)])).pin_memory()
    eng = pkg.get_engine()
    ref = None
    for sset in a.settings:
        f = sset.split(':')
        chunk, buffers, streams, transport = int(f[0]), int(f[1]), int(f[2]), f[3]
        workers = int(f[4]) if len(f) > 4 else 6
        sd = StreamedDecomposer(eng, a.pages, H, W, C, chunk=chunk, bg_downsample=3, buffers=buffers, compute_streams=streams,
                                packed_mask=transport == 'handoff', mask_transport='packed' if transport == 'packed' else 'bool',
                                unpack_workers=workers)
        outs = [sd.alloc_outputs() for _ in range(2)]

        def run(steps):
            pending = []
            for i in range(steps):
                pending.append(sd.run_async(host, outs[i % 2], 101, denoise_mask='fast'))
                if len(pending) >= 2:
                    pending.pop(0).synchronize()
            for p in pending:
                p.synchronize()
        run(2)
        torch.cuda.synchronize()
        t0 = time.time()
        run(a.steps)
        torch.cuda.synchronize()
        dt = (time.time() - t0) / a.steps
        o = outs[0]
        m = o['mask'].numpy()
        if transport == 'handoff':
            m = np.unpackbits(m, axis=2)[:, :, :W]
        sig = (int(m.sum()), int(o['fg'].numpy()[::7].sum()), int(o['bg'].numpy().sum()))
        ref = ref or sig
        print(json.dumps({'setting': sset, 'ms_per_step': round(dt * 1e3, 2), 'Gpx_s': round(a.pages * H * W / dt / 1e9, 2),
                          'same_as_first': sig == ref}), flush=True)
        sd.close()
        del sd, outs
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
