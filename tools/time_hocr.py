#!/usr/bin/env python
"""create_hocr_mask (mrc.py:188-270) on one 3300x2550 page with ~60 text lines: the batched device path (five launches
per page) against the per-line form of round 1 (copy2d + two Sauvola launches per line), wall ms per page incl. the
decisions' device-to-host reads, and kernel launches per page."""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def per_line_form(eng, L, E, gray, lines, window):
    """Round-1 structure: 3 launches per line (the measurements and pastes are left out: it is the floor of that form)."""
    import torch
    lib = L.lib()
    st = E._stream_ptr()
    geo, off = [], 0
    for (l, t, r, b) in lines:
        w, h = r - l, b - t
        pitch = (w + 15) // 16 * 16
        size = (pitch * h + 255) // 256 * 256
        geo.append((w, h, pitch, off, off + size, off + 2 * size)); off += 3 * size
    scratch = torch.empty(off, dtype=torch.uint8, device=eng.device)
    base = scratch.data_ptr()
    for (l, t, r, b), (w, h, pitch, o_in, o_th, o_ti) in zip(lines, geo):
        L.check(lib.b200mrc_copy2d(C.c_void_p(base + o_in), pitch, C.c_void_p(gray.t.data_ptr() + t * gray.pitch + l), gray.pitch, w, h, L.COPY_D2D, st))
        for o_out, flags in ((o_th, 0), (o_ti, L.SAUVOLA_INVERT_INPUT)):
            L.check(lib.b200mrc_sauvola(C.c_void_p(base + o_in), pitch, pitch * h, C.c_void_p(base + o_out), pitch, pitch * h,
                                        w, h, 1, window, window, 0.1, 128.0, flags, st))
    torch.cuda.synchronize()


def main():
    import torch
    import archive_pdf_tools_b200 as pkg
    from archive_pdf_tools_b200 import _lib as L, engine as E, synth, mrc
    H, W, dpi = 3300, 2550, 400
    page = synth.make_page(5, H, W, dpi=dpi, rgb=False, invert_lines=(3, 17), noisy_dark_lines=(9,))
    hocr = synth.page_hocr(H, W, dpi=dpi)
    eng = pkg.get_engine()
    gray = E.Plane(1, H, W, 1, eng.device).upload(page[None], non_blocking=False)
    mask = E.Plane(1, H, W, 1, eng.device)
    lines = list(mrc._iter_text_lines(hocr, W, H, None))
    window = E.window_for_dpi(dpi)
    out = {'page': [H, W], 'text_lines': len(lines), 'window': window}
    for name, fn in (('batched', lambda: mrc.create_hocr_mask(gray, mask, hocr, dpi=dpi)),
                     ('per_line_round1_floor', lambda: per_line_form(eng, L, E, gray, lines, window))):
        mask.t.zero_()
        fn(); torch.cuda.synchronize()
        n0 = L.lib().b200mrc_launch_count()
        t0 = time.time()
        reps = 5
        for _ in range(reps):
            mask.t.zero_()
            fn()
        torch.cuda.synchronize()
        out[name] = {'ms_per_page': (time.time() - t0) / reps * 1e3, 'kernel_launches_per_page': (L.lib().b200mrc_launch_count() - n0) / reps}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
