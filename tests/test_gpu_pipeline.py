"""GPU parity tests of the whole path: the reference-facing surface (generator, threshold_image,
drop-in modules) and the batched C-ABI pipeline, against golden reference outputs and the oracle.
Bar (BASELINE.json): masks bit-exact; fg/bg within +-1 LSB (we assert exact equality and report)."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden, hocr_cases, load_hocr_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', golden_cases())
def test_generator_matches_golden_reference(eng, synth, name):
    from PIL import Image
    import archive_pdf_tools_b200 as pkg
    g = load_golden(name, synth)
    timing, errors = [], set()
    gen = pkg.create_mrc_hocr_components(Image.fromarray(g['page']), [], dpi=g['dpi'], bg_downsample=g['bg_downsample'],
                                         fg_downsample=g['fg_downsample'], denoise_mask=g['denoise'],
                                         timing_data=timing, errors=errors)
    mask = next(gen); fg = next(gen); bg = next(gen)
    with pytest.raises(StopIteration):
        next(gen)
    assert mask.dtype == np.bool_ and fg.dtype == np.uint8 and bg.dtype == np.uint8
    assert np.array_equal(mask, g['mask']), int((mask != g['mask']).sum())
    assert fg.shape == g['fg'].shape and bg.shape == g['bg'].shape
    assert np.abs(fg.astype(int) - g['fg']).max() <= 1 and np.abs(bg.astype(int) - g['bg']).max() <= 1
    assert np.array_equal(fg, g['fg']) and np.array_equal(bg, g['bg'])
    assert [k for k, _ in timing] == g['timing_keys']
    assert errors == set()


@pytest.mark.parametrize('name', golden_cases())
def test_threshold_image_matches_golden(eng, synth, orc, name):
    import archive_pdf_tools_b200 as pkg
    g = load_golden(name, synth)
    gray = g['page'] if g['page'].ndim == 2 else orc.rgb2gray(g['page'])
    assert np.array_equal(pkg.threshold_image(gray, 132), g['t33'])
    assert np.array_equal(pkg.threshold_image(gray, g['dpi'], 0.1), g['t01'])


def test_mask_only_generator_stops_after_first_yield(eng, synth, orc):
    from PIL import Image
    import archive_pdf_tools_b200 as pkg
    page = synth.make_page(21, 240, 200, dpi=100, rgb=False)
    gen = pkg.create_mrc_hocr_components(Image.fromarray(page), [], dpi=100, denoise_mask='fast')
    mask = next(gen)                                   # recode.py:398-407 (--bw-pdf) stops here
    exp = orc.decompose(page, dpi=100, denoise_mask='fast', mask_only=True)
    assert np.array_equal(mask, exp['mask'])


def test_generator_error_behaviour(eng):
    from PIL import Image
    import archive_pdf_tools_b200 as pkg
    im = Image.fromarray(np.full((16, 16), 200, np.uint8))
    with pytest.raises(ValueError):
        next(pkg.create_mrc_hocr_components(im, [], dpi=100))                      # denoise_mask=None -> mrc.py:396
    errs = set()
    gen = pkg.create_mrc_hocr_components(Image.fromarray(np.full((2, 9), 200, np.uint8)), [], dpi=100, bg_downsample=3,
                                         denoise_mask='none', errors=errs)
    out = list(gen)
    assert errs == {'too-small-to-downsample'} and out[2].shape == (2, 9)


def test_palette_mode_goes_through_pil_like_reference(eng, synth, orc):
    from PIL import Image
    import archive_pdf_tools_b200 as pkg
    page = synth.make_page(22, 120, 100, dpi=100)
    im = Image.fromarray(page).convert('P')
    mask, fg, bg = list(pkg.create_mrc_hocr_components(im, [], dpi=100, bg_downsample=2, denoise_mask='fast'))
    gray = np.array(im.convert('L')); rgb = np.array(im.convert('RGB'))
    m, _ = orc.threshold_mask(gray, dpi=100)
    m = orc.denoise(m)
    assert np.array_equal(mask, m)
    assert np.array_equal(fg, orc.optimise(m, rgb, 3))
    assert np.array_equal(bg, orc.thumbnail(orc.optimise(~m, rgb, 10), 50, 60))


@pytest.mark.parametrize('cfg', [
    dict(n=3, h=330, w=255, rgb=True, dpi=100, bg=3, fg=None, den='fast', sn=3.0, ht=False),
    dict(n=2, h=300, w=1100, rgb=True, dpi=300, bg=3, fg=None, den='fast', sn=3.0, ht=True),
    dict(n=4, h=220, w=170, rgb=False, dpi=200, bg=None, fg=None, den='fast', sn=3.0, ht=False),
    dict(n=2, h=260, w=300, rgb=True, dpi=None, bg=4, fg=2, den='none', sn=10.0, ht=False),
])
def test_batched_decompose_matches_oracle(eng, synth, orc, cfg):
    import archive_pdf_tools_b200 as pkg
    pages = np.stack([synth.make_page(100 + i, cfg['h'], cfg['w'], dpi=cfg['dpi'] or 200, rgb=cfg['rgb'], sigma_n=cfg['sn'],
                                      halftone=cfg['ht'] and i % 2 == 0) for i in range(cfg['n'])])
    res = pkg.decompose_pages(pages, dpi=cfg['dpi'], bg_downsample=cfg['bg'], fg_downsample=cfg['fg'], denoise_mask=cfg['den'])
    for i in range(cfg['n']):
        exp = orc.decompose(pages[i], dpi=cfg['dpi'], bg_downsample=cfg['bg'], fg_downsample=cfg['fg'], denoise_mask=cfg['den'])
        assert res['sigma'][i] == exp['sigma'], (i, res['sigma'][i], exp['sigma'])
        assert np.array_equal(res['mask'][i], exp['mask']), (i, int((res['mask'][i] != exp['mask']).sum()))
        assert np.array_equal(res['fg'][i], exp['fg']), i
        assert np.array_equal(res['bg'][i], exp['bg']), i


def test_injected_sigma_and_mask_only(eng, synth, orc):
    import archive_pdf_tools_b200 as pkg
    pages = np.stack([synth.make_page(200 + i, 200, 160, dpi=100, rgb=False, sigma_n=2.0) for i in range(3)])
    sig = [0.3, 2.0, 7.5]
    res = pkg.decompose_pages(pages, dpi=100, denoise_mask='fast', mask_only=True, sigma=sig)
    for i in range(3):
        exp = orc.decompose(pages[i], dpi=100, denoise_mask='fast', mask_only=True, sigma_est=sig[i])
        assert np.array_equal(res['mask'][i], exp['mask'])
    assert 'fg' not in res


def test_full_size_400dpi_page_matches_oracle(eng, synth, orc):
    """BASELINE config 2 shape (one page of the 64): 3300x2550 RGB, dpi 400 (window 101), bg/3, denoise fast."""
    import archive_pdf_tools_b200 as pkg
    pages = np.stack([synth.make_page(i, 3300, 2550, dpi=400, halftone=(i == 1)) for i in range(2)])
    res = pkg.decompose_pages(pages, dpi=400, bg_downsample=3, denoise_mask='fast')
    for i in range(2):
        exp = orc.decompose(pages[i], dpi=400, bg_downsample=3, denoise_mask='fast')
        assert res['sigma'][i] == exp['sigma']
        assert np.array_equal(res['mask'][i], exp['mask']), int((res['mask'][i] != exp['mask']).sum())
        assert np.array_equal(res['fg'][i], exp['fg'])
        assert np.array_equal(res['bg'][i], exp['bg']) and res['bg'][i].shape == (1100, 850, 3)
    # size-independent properties: results do not depend on batch composition / order
    res1 = pkg.decompose_pages(pages[::-1].copy(), dpi=400, bg_downsample=3, denoise_mask='fast')
    assert np.array_equal(res1['mask'][1], res['mask'][0]) and np.array_equal(res1['bg'][0], res['bg'][1])
    # fg equals the page on mask pixels, bg equals the page off the mask (optimiser.pyx copies img first)
    m = res['mask'][0]
    assert np.array_equal(res['fg'][0][m], pages[0][m])


def test_config3_full_size_book_pages_match_oracle(eng, synth, orc):
    """BASELINE config 3 shape: 3300x2550 RGB @300 DPI (window 75), halftone pages mixed with text pages in one batch
    (their sigma_est puts them on the tiled pre-blur + direct-mode threshold, the others on the fused in-kernel blur)."""
    import archive_pdf_tools_b200 as pkg
    pages = np.stack([synth.make_page(20240000 + i, 3300, 2550, dpi=300, halftone=(i != 1)) for i in range(3)])
    res = pkg.decompose_pages(pages, dpi=300, bg_downsample=3, denoise_mask='fast')
    radii = set()
    for i in range(3):
        exp = orc.decompose(pages[i], dpi=300, bg_downsample=3, denoise_mask='fast')
        assert res['sigma'][i] == exp['sigma']
        radii.add(int(0.4 * exp['sigma'] + 0.5) if exp['sigma'] > 1.0 else 0)
        assert np.array_equal(res['mask'][i], exp['mask']), (i, int((res['mask'][i] != exp['mask']).sum()))
        assert np.array_equal(res['fg'][i], exp['fg']) and np.array_equal(res['bg'][i], exp['bg']), i
    assert len(radii) > 1, radii                              # the batch really mixes blur radii


def test_config1_window33_full_page(eng, synth, orc):
    import archive_pdf_tools_b200 as pkg
    page = synth.make_page(0, 3300, 2550, dpi=400, rgb=False)
    assert np.array_equal(pkg.threshold_image(page, 132), orc.sauvola(page, 33))       # BASELINE config 1
    assert np.array_equal(pkg.threshold_image(page, 600), orc.sauvola(page, 151))


def test_dropin_modules_match_reference_cython(eng, refmods):
    import archive_pdf_tools_b200 as pkg
    pkg.install(patch_reference=False)
    import sauvola, optimiser
    assert sauvola.__file__.startswith(pkg.DROPIN_DIR) and optimiser.__file__.startswith(pkg.DROPIN_DIR)
    rsau, ropt = refmods
    rng = np.random.default_rng(8)
    h, w = 90, 140
    img = np.clip(rng.normal(180, 40, (h, w)), 0, 255).astype(np.uint8)
    o1 = np.empty(h * w, np.uint8); o2 = np.ndarray(h * w, dtype=bool)
    assert rsau.binarise_sauvola(img.reshape(-1), o1, w, h, 33, 33, 0.34, 128) == 0
    assert sauvola.binarise_sauvola(img.reshape(-1), o2, w, h, 33, 33, 0.34, 128) == 0
    assert np.array_equal(o1, o2.view(np.uint8))
    mask = (rng.random((h, w)) < 0.1)
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    for n in (3, 10):
        assert np.array_equal(optimiser.optimise_rgb2(mask.view(np.uint8), rgb, w, h, n), ropt.optimise_rgb2(mask.view(np.uint8), rgb, w, h, n))
        assert np.array_equal(optimiser.optimise_gray(mask, img, w, h, n), ropt.optimise_gray(mask.view(np.uint8), img, w, h, n))
    a = mask.copy(); b = mask.copy().view(np.uint8)
    ret = optimiser.fast_mask_denoise(a, w, h, 4, 2)
    ropt.fast_mask_denoise(b, w, h, 4, 2)
    assert ret is a and np.array_equal(a.view(np.uint8), b)
    dense = rng.random((h, w)) < 0.35
    for mincnt, n_size in [(1, 1), (3, 1), (8, 3), (24, 2), (0, 2), (2, 0), (60, 5)]:   # the general form: any (mincnt, n_size)
        a = dense.copy(); b = dense.copy().view(np.uint8)
        assert optimiser.fast_mask_denoise(a, w, h, mincnt, n_size) is a
        ropt.fast_mask_denoise(b, w, h, mincnt, n_size)
        assert np.array_equal(a.view(np.uint8), b), (mincnt, n_size, int((a.view(np.uint8) != b).sum()))
    with pytest.raises(ValueError):
        optimiser.optimise_gray2(mask, img.astype(np.float32), w, h, 3)


def test_streamed_decomposer_matches_batched(eng, synth):
    """Chunked H2D / compute / D2H overlap (3 streams, double-buffered) must not change a byte."""
    import torch
    import archive_pdf_tools_b200 as pkg
    from archive_pdf_tools_b200.engine import StreamedDecomposer
    pages = np.stack([synth.make_page(300 + i, 270, 330, dpi=100, sigma_n=3.0, halftone=(i == 2)) for i in range(7)])
    ref = pkg.decompose_pages(pages, dpi=100, bg_downsample=3, denoise_mask='fast')
    host = torch.from_numpy(pages).pin_memory()
    for kw in (dict(chunk=2), dict(chunk=2, buffers=2, compute_streams=1, mask_transport='bool'), dict(chunk=3, buffers=3, compute_streams=3),
               dict(chunk=7, mask_transport='bool')):
        sd = StreamedDecomposer(eng, 7, 270, 330, 3, bg_downsample=3, **kw)
        out = sd.alloc_outputs()
        for _ in range(2):
            for v in out.values():
                v.zero_()
            sd.run(host, out, 25, denoise_mask='fast')
        assert np.array_equal(out['mask'].numpy().astype(bool), ref['mask']), kw
        assert np.array_equal(out['fg'].numpy().reshape(ref['fg'].shape), ref['fg']), kw
        assert np.array_equal(out['bg'].numpy().reshape(ref['bg'].shape), ref['bg']), kw


def test_streamed_decomposer_back_to_back_calls(eng, synth):
    """run_async: several batches in flight (no wait between calls, own result buffers each) give the same bytes
    as one call at a time; buffer slots are handed from call to call by events."""
    import torch
    import archive_pdf_tools_b200 as pkg
    from archive_pdf_tools_b200.engine import StreamedDecomposer
    sets = [np.stack([synth.make_page(400 + 10 * j + i, 270, 330, dpi=100, sigma_n=3.0, halftone=(i == 1)) for i in range(5)]) for j in range(3)]
    refs = [pkg.decompose_pages(pg, dpi=100, bg_downsample=3, denoise_mask='fast') for pg in sets]
    hosts = [torch.from_numpy(pg).pin_memory() for pg in sets]
    for kw in (dict(chunk=2), dict(chunk=1, buffers=3, compute_streams=3, mask_transport='bool'), dict(chunk=5)):
        sd = StreamedDecomposer(eng, 5, 270, 330, 3, bg_downsample=3, **kw)
        outs = [sd.alloc_outputs() for _ in range(3)]
        for rep in range(2):
            for o in outs:
                for v in o.values():
                    v.zero_()
            evs = [sd.run_async(hosts[j], outs[j], 25, denoise_mask='fast') for j in range(3)]
            for j in (2, 0, 1):
                evs[j].synchronize()
                assert np.array_equal(outs[j]['mask'].numpy().astype(bool), refs[j]['mask']), (kw, rep, j)
                assert np.array_equal(outs[j]['fg'].numpy().reshape(refs[j]['fg'].shape), refs[j]['fg']), (kw, rep, j)
                assert np.array_equal(outs[j]['bg'].numpy().reshape(refs[j]['bg'].shape), refs[j]['bg']), (kw, rep, j)


def test_copy2d_roundtrip(eng):
    """b200mrc_copy2d: pitched H2D / D2D / D2H of a batch of pages keeps every byte, touches no padding."""
    import ctypes as C
    import torch
    from archive_pdf_tools_b200 import _lib, engine as E
    n, h, w, c = 3, 37, 101, 3
    src = torch.randint(0, 256, (n, h, w * c), dtype=torch.uint8).pin_memory()
    a, b = E.Plane(n, h, w, c, eng.device), E.Plane(n, h, w, c, eng.device)
    a.t.fill_(7); b.t.fill_(9)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L = _lib.lib()
    assert L.b200mrc_copy2d(a.ptr, a.pitch, C.c_void_p(src.data_ptr()), w * c, w * c, n * h, _lib.COPY_H2D, st) == 0
    assert L.b200mrc_copy2d(b.ptr, b.pitch, a.ptr, a.pitch, w * c, n * h, _lib.COPY_D2D, st) == 0
    dst = torch.zeros_like(src).pin_memory()
    assert L.b200mrc_copy2d(C.c_void_p(dst.data_ptr()), w * c, b.ptr, b.pitch, w * c, n * h, _lib.COPY_D2H, st) == 0
    torch.cuda.synchronize()
    assert torch.equal(dst, src)
    assert bool((b.t[:, :, w * c:] == 9).all())                  # row padding untouched
    assert L.b200mrc_copy2d(None, 0, None, 0, 1, 1, 1, st) == _lib.ERR_INVALID


@pytest.mark.parametrize('name', hocr_cases())
def test_generator_with_hocr_lines_matches_golden_reference(eng, synth, orc, name):
    """create_hocr_mask on the device (mrc.py:188-270): per-line Sauvola on crop / inverted crop, fill ratios, the
    sigma tie-break and the pastes, then the usual path with mask |= thres."""
    from PIL import Image
    import archive_pdf_tools_b200 as pkg
    from archive_pdf_tools_b200 import mrc as pm, engine as E
    g = load_hocr_golden(name, synth)
    timing = []
    gen = pkg.create_mrc_hocr_components(Image.fromarray(g['page']), g['hocr'], dpi=g['dpi'], downsample=g['downsample'],
                                         bg_downsample=g['bg_downsample'], denoise_mask=g['denoise'], timing_data=timing)
    mask = next(gen); fg = next(gen); bg = next(gen)
    assert np.array_equal(mask, g['mask']), int((mask != g['mask']).sum())
    assert np.array_equal(fg, g['fg']) and np.array_equal(bg, g['bg'])
    assert [k for k, _ in timing] == g['timing_keys']
    # the hOCR mask on its own
    gray = g['page'] if g['page'].ndim == 2 else orc.rgb2gray(g['page'])
    gp = E.Plane(1, gray.shape[0], gray.shape[1], 1, eng.device).upload(gray[None])
    mp = E.Plane(1, gray.shape[0], gray.shape[1], 1, eng.device)
    mp.t.zero_()
    pm.create_hocr_mask(gp, mp, g['hocr'], downsample=g['downsample'], dpi=g['dpi'])
    assert np.array_equal(mp.numpy(np.bool_)[0], g['hocr_mask'])


def test_rect_kernels_match_oracle(eng, orc):
    """b200mrc_rects_count_nonzero / b200mrc_rects_sigma_bool against np.count_nonzero and the float64 restatement of
    mean_estimate_sigma on boolean crops of assorted shapes (incl. 1-pixel-wide, all-false, all-true)."""
    import ctypes as C
    import torch
    from archive_pdf_tools_b200 import _lib, engine as E
    rng = np.random.default_rng(11)
    shapes = [(1, 1), (1, 37), (40, 1), (2, 2), (3, 5), (10, 198), (17, 246), (33, 64), (21, 401)]
    crops = []
    for i, (h, w) in enumerate(shapes):
        p = [0.05, 0.5, 0.9, 0.0, 1.0][i % 5]
        a = (rng.random((h, w)) < p)
        if i == 6:                                           # glyph-like structure instead of noise
            a[:] = False; a[4:12, ::7] = True; a[8, :] = True
        crops.append(a)
    planes = [E.Plane(1, a.shape[0], a.shape[1], 1, eng.device).upload(a.view(np.uint8)[None]) for a in crops]
    keys = [torch.empty(((a.shape[0] + 3) // 2) * ((a.shape[1] + 3) // 2), dtype=torch.int64, device=eng.device) for a in crops]
    arr = (_lib.Rect * len(crops))()
    for r, pl, k in zip(arr, planes, keys):
        r.ptr, r.pitch, r.width, r.height, r.keys = pl.t.data_ptr(), pl.pitch, pl.w, pl.h, k.data_ptr()
    rd = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(eng.device)
    counts = torch.empty(len(crops), dtype=torch.int32, device=eng.device)
    sig = torch.empty(len(crops), dtype=torch.float64, device=eng.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L = _lib.lib()
    assert L.b200mrc_rects_count_nonzero(C.c_void_p(rd.data_ptr()), len(crops), C.c_void_p(counts.data_ptr()), st) == 0
    assert L.b200mrc_rects_sigma_bool(C.c_void_p(rd.data_ptr()), len(crops), C.c_void_p(sig.data_ptr()), st) == 0
    torch.cuda.synchronize()
    assert counts.cpu().tolist() == [int(np.count_nonzero(a)) for a in crops]
    for a, s in zip(crops, sig.cpu().tolist()):
        exp = orc.estimate_sigma_bool(a)
        assert (np.isnan(exp) and np.isnan(s)) or s == exp, (a.shape, s, exp)


def test_sauvola_invert_input_flag(eng, orc):
    from archive_pdf_tools_b200 import _lib, engine as E
    rng = np.random.default_rng(3)
    for (h, w, win) in ((10, 198, 25), (33, 70, 51), (5, 9, 25)):
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        src = E.Plane(1, h, w, 1, eng.device).upload(img[None])
        dst = E.Plane(1, h, w, 1, eng.device)
        eng.sauvola(src, dst, win, win, 0.1, 128.0, _lib.SAUVOLA_INVERT_INPUT)
        assert np.array_equal(dst.numpy(np.bool_)[0], orc.sauvola(255 - img, win, k=0.1)), (h, w, win)


def test_packed_mask_matches_packbits_and_pil(eng):
    """b200mrc_pack_mask: the PIL mode-'1' rows encode_mrc_mask builds (mrc.py:490: Image.fromarray(np_mask))."""
    from PIL import Image
    import torch
    import archive_pdf_tools_b200 as pkg
    from archive_pdf_tools_b200 import engine as E
    rng = np.random.default_rng(17)
    for (h, w) in ((1, 1), (3, 7), (5, 8), (9, 17), (40, 255), (33, 256), (21, 2550)):
        m = rng.random((h, w)) < 0.3
        for inv in (False, True):
            data, im = pkg.packed_mask(m, invert=inv)
            ref = m ^ inv
            assert data == np.packbits(ref, axis=1).tobytes(), (h, w, inv)
            assert im.mode == '1' and np.array_equal(np.array(im), np.array(Image.fromarray(ref)))
    # batched planes, arbitrary non-zero bytes count as set
    m = (rng.integers(0, 4, (3, 30, 100)) * 85).astype(np.uint8)
    pl = E.Plane(3, 30, 100, 1, eng.device).upload(m)
    out = eng.pack_mask(pl).cpu().numpy()
    assert np.array_equal(out, np.packbits(m != 0, axis=2))


def test_lanczos_pre_downsample_matches_pillow(eng, orc):
    """recode.py:368-372: image.thumbnail((w/ds, h/ds), resample=LANCZOS, reducing_gap=None)."""
    from PIL import Image
    import archive_pdf_tools_b200 as pkg
    rng = np.random.default_rng(23)
    for (h, w, c, ds) in ((120, 90, 3, 2), (201, 333, 1, 3), (97, 64, 3, 1.5), (64, 64, 1, 4), (50, 40, 3, 1)):
        arr = rng.integers(0, 256, (h, w, 3) if c == 3 else (h, w), dtype=np.uint8)
        im = Image.fromarray(arr)
        ref = im.copy()
        ref.thumbnail((w / ds, h / ds), resample=Image.LANCZOS, reducing_gap=None)
        out = pkg.downsample_image(im, ds)
        assert out.size == ref.size and out.mode == ref.mode
        assert np.array_equal(np.array(out), np.array(ref)), (h, w, c, ds)
        assert np.array_equal(orc.thumbnail(arr, w / ds, h / ds, reducing_gap=None, filter=orc.LANCZOS), np.array(ref))


def test_config4_600dpi_page_both_windows(eng, synth, orc):
    """BASELINE config 4: 600-DPI scan 6600x5100 RGB, denoise fast, bg/3 -- Sauvola window 51 as the config names it
    (explicit window) and the API-derived 151 for dpi=600."""
    import archive_pdf_tools_b200 as pkg
    page = synth.make_page(40, 6600, 5100, dpi=600)
    for kw in (dict(window=51), dict(dpi=600)):
        res = pkg.decompose_pages(page[None], bg_downsample=3, denoise_mask='fast', **kw)
        exp = orc.decompose(page, bg_downsample=3, denoise_mask='fast', **kw)
        assert res['sigma'][0] == exp['sigma']
        assert np.array_equal(res['mask'][0], exp['mask']), int((res['mask'][0] != exp['mask']).sum())
        assert np.array_equal(res['fg'][0], exp['fg']) and np.array_equal(res['bg'][0], exp['bg'])
        assert res['bg'][0].shape == (2200, 1700, 3)


def test_config5_mask_only_200dpi_gray_batch(eng, synth, orc):
    """BASELINE config 5: gray pages 2200x1700 @200 DPI (window 51), mask-only (first yield), denoise fast."""
    import archive_pdf_tools_b200 as pkg
    pages = np.stack([synth.make_page(50 + i, 2200, 1700, dpi=200, rgb=False, halftone=(i == 2)) for i in range(3)])
    res = pkg.decompose_pages(pages, dpi=200, denoise_mask='fast', mask_only=True)
    assert 'fg' not in res and 'bg' not in res
    for i in range(3):
        exp = orc.decompose(pages[i], dpi=200, denoise_mask='fast', mask_only=True)
        assert np.array_equal(res['mask'][i], exp['mask']), i


def test_batched_decompose_with_hocr_matches_oracle(eng, synth, orc):
    """decompose_pages with per-page text boxes: pages with and without hOCR data in one batch."""
    import archive_pdf_tools_b200 as pkg
    H, W, dpi = 330, 255, 100
    pages = np.stack([synth.make_page(60 + i, H, W, dpi=dpi, sigma_n=1.0, invert_lines=(2,), noisy_dark_lines=(5, 7)) for i in range(3)])
    hocr = [synth.page_hocr(H, W, dpi=dpi), [], synth.page_hocr(H, W, dpi=dpi, low_conf_every=3)]
    res = pkg.decompose_pages(pages, dpi=dpi, bg_downsample=3, denoise_mask='fast', hocr_word_data=hocr)
    for i in range(3):
        exp = orc.decompose(pages[i], dpi=dpi, bg_downsample=3, denoise_mask='fast', hocr_word_data=hocr[i])
        assert np.array_equal(res['mask'][i], exp['mask']), (i, int((res['mask'][i] != exp['mask']).sum()))
        assert np.array_equal(res['fg'][i], exp['fg']) and np.array_equal(res['bg'][i], exp['bg']), i
    plain = pkg.decompose_pages(pages, dpi=dpi, bg_downsample=3, denoise_mask='fast')
    assert not np.array_equal(plain['mask'][0], res['mask'][0]) and np.array_equal(plain['mask'][1], res['mask'][1])


@pytest.mark.parametrize('name', golden_cases()[:3] + hocr_cases()[:2])
def test_unmodified_reference_through_install(eng, synth, orc, name):
    """Drop-in smoke through the reference's own call site: install() publishes the GPU `sauvola` / `optimiser`
    modules, then the UNMODIFIED reference internetarchivepdf/mrc.py (byte-compiled into oracle/_ref by
    oracle/build_ref.py, or the checkout itself in the build container) runs create_mrc_hocr_components on top of
    them, like recode.py:400-406 / 427-433 and bin/compress-pdf-images:66-70 do.  Outputs must equal the golden
    fixtures (made by the same reference code on its own Cython)."""
    from PIL import Image
    import archive_pdf_tools_b200 as pkg
    from oracle import ref_pipeline
    pkg.install(patch_reference=False)
    ref = ref_pipeline.load_reference_mrc_on_dropin()
    if ref is None:
        pytest.skip('reference glue not available (oracle/_ref/internetarchivepdf not built)')
    assert ref.binarise_sauvola.__module__ == 'sauvola' and ref.optimise_rgb2.__module__ == 'optimiser'
    import sauvola
    assert sauvola.__file__.startswith(pkg.DROPIN_DIR)
    if name.startswith('hocr_'):
        g = load_hocr_golden(name, synth)
        kw = dict(dpi=g['dpi'], downsample=g['downsample'], bg_downsample=g['bg_downsample'], denoise_mask=g['denoise'])
        hocr = g['hocr']
    else:
        g = load_golden(name, synth)
        kw = dict(dpi=g['dpi'], bg_downsample=g['bg_downsample'], fg_downsample=g['fg_downsample'], denoise_mask=g['denoise'])
        hocr = []
    timing, errors = [], set()
    mask, fg, bg = list(ref.create_mrc_hocr_components(Image.fromarray(g['page']), hocr, timing_data=timing, errors=errors, **kw))
    assert np.array_equal(mask, g['mask']), int((mask != g['mask']).sum())
    assert np.array_equal(fg, g['fg']) and np.array_equal(bg, g['bg'])
    assert [k for k, _ in timing] == g['timing_keys']


@pytest.mark.parametrize('binding', ['dropin-modules', 'install-rebinds-mrc'])
def test_reference_recode_page_loop_on_the_dropin(eng, synth, tmp_path, binding):
    """The reference's page loop itself -- the UNMODIFIED recode.insert_images_mrc (recode.py:266-530: image loading, the
    create_mrc_hocr_components call sites at :400-406 (1-bit output) and :427-433 (full MRC), the encoder and PDF calls,
    here mocked) -- running on this engine, both ways a maintainer can switch: (a) install() only publishes the GPU
    `sauvola` / `optimiser` modules under the reference's mrc.py; (b) install() rebinds create_mrc_hocr_components in the
    imported internetarchivepdf.mrc before recode.py binds it by name (recode.py:39-40)."""
    import sys
    import archive_pdf_tools_b200 as pkg
    from conftest import drive_reference_page_loop
    from oracle import ref_pipeline
    pkg.install(patch_reference=False)
    mrc = ref_pipeline.load_reference_mrc_on_dropin()
    if mrc is None:
        pytest.skip('reference glue not available (oracle/_ref/internetarchivepdf not built)')
    if binding == 'install-rebinds-mrc':
        sys.modules['internetarchivepdf.mrc'] = mrc
        try:
            pkg.install(patch_reference=True)
        finally:
            del sys.modules['internetarchivepdf.mrc']
    recode = ref_pipeline.load_reference_recode(mrc)
    if recode is None:
        pytest.skip('reference recode glue not available')
    if binding == 'install-rebinds-mrc':
        assert recode.create_mrc_hocr_components is pkg.create_mrc_hocr_components
    else:
        assert recode.create_mrc_hocr_components.__module__ == 'internetarchivepdf.mrc' and mrc.optimise_rgb2.__module__ == 'optimiser'
    for name in ('rgb_clean_bg3', 'gray_noisy_bg2_fg2', 'rgb_halftone_nodenoise_bg4'):
        g = load_golden(name, synth)
        kw = dict(bg_downsample=g['bg_downsample'], fg_downsample=g['fg_downsample'], denoise_mask=g['denoise'])
        cap, pdf, errors = drive_reference_page_loop(recode, [g['page'], g['page']], [[], []], tmp_path, g['dpi'], **kw)
        assert len(cap) == 2 and len(pdf[1].inserted) == 2
        for c in cap:
            assert np.array_equal(c['mask'], g['mask']), (name, int((c['mask'] != g['mask']).sum()))
            assert np.array_equal(c['fg'], g['fg']) and np.array_equal(c['bg'], g['bg']), name
        cap, pdf, errors = drive_reference_page_loop(recode, [g['page']], [[]], tmp_path, g['dpi'], force_1bit=True, **kw)
        assert np.array_equal(cap[0]['mask_inverted'], ~g['mask']), name
    g = load_hocr_golden('hocr_rgb_clean_bg3', synth)
    cap, pdf, errors = drive_reference_page_loop(recode, [g['page']], [g['hocr']], tmp_path, g['dpi'], downsample=g['downsample'],
                                                 bg_downsample=g['bg_downsample'], denoise_mask=g['denoise'])
    assert np.array_equal(cap[0]['mask'], g['mask']) and np.array_equal(cap[0]['fg'], g['fg']) and np.array_equal(cap[0]['bg'], g['bg'])


def test_reference_compress_pdf_images_script_on_the_dropin(eng, synth, orc):
    """The reference's second caller, the script bin/compress-pdf-images (:66-70: create_mrc_hocr_components(image,
    hocr_word_data, denoise_mask=DENOISE_FAST, bg_downsample=3), no dpi), run UNMODIFIED as __main__ over a fake PyMuPDF
    document, on the engine's create_mrc_hocr_components (install() rebinding) -- results equal the oracle's."""
    import sys
    import archive_pdf_tools_b200 as pkg
    from conftest import run_reference_compress_script
    from oracle import ref_pipeline
    pkg.install(patch_reference=False)
    mrc = ref_pipeline.load_reference_mrc_on_dropin()
    if mrc is None:
        pytest.skip('reference glue not available')
    sys.modules['internetarchivepdf.mrc'] = mrc
    try:
        pkg.install(patch_reference=True)
    finally:
        del sys.modules['internetarchivepdf.mrc']
    assert mrc.create_mrc_hocr_components is pkg.create_mrc_hocr_components
    pages = [synth.make_page(60 + i, 300, 260, dpi=100, rgb=(i != 1)) for i in range(3)]
    out = run_reference_compress_script(mrc, pages)
    if out is None:
        pytest.skip('reference script byte code not available')
    cap, doc = out
    assert len(cap) == 3 and doc.saved == 'out.pdf' and all(len(p.inserted) == 2 for p in doc.pages)
    for pg, c in zip(pages, cap):
        exp = orc.decompose(pg, dpi=None, bg_downsample=3, denoise_mask='fast')
        assert np.array_equal(c['mask'], exp['mask']) and np.array_equal(c['fg'], exp['fg']) and np.array_equal(c['bg'], exp['bg'])


def test_install_rebinds_an_imported_reference(eng, synth):
    """install(patch_reference=True) with internetarchivepdf.mrc already imported: its pixel-path names are rebound."""
    import sys
    import archive_pdf_tools_b200 as pkg
    from oracle import ref_pipeline
    pkg.install(patch_reference=False)
    ref = ref_pipeline.load_reference_mrc_on_dropin()
    if ref is None:
        pytest.skip('reference glue not available')
    sys.modules['internetarchivepdf.mrc'] = ref
    try:
        pkg.install(patch_reference=True)
        assert ref.create_mrc_hocr_components is pkg.create_mrc_hocr_components
        assert ref.threshold_image is pkg.threshold_image
        page = synth.make_page(4, 120, 90, dpi=100, rgb=False)
        assert np.array_equal(ref.threshold_image(page, 100), pkg.threshold_image(page, 100))
    finally:
        sys.modules.pop('internetarchivepdf.mrc', None)


def test_streamed_decomposer_packed_mask(eng, synth):
    """packed_mask=True returns PIL mode-'1' rows (np.packbits) instead of the bool plane; fg / bg unchanged."""
    import torch
    import archive_pdf_tools_b200 as pkg
    from archive_pdf_tools_b200.engine import StreamedDecomposer
    pages = np.stack([synth.make_page(400 + i, 130, 203, dpi=100) for i in range(11)])
    ref = pkg.decompose_pages(pages, dpi=100, bg_downsample=3, denoise_mask='fast')
    host = torch.from_numpy(pages).pin_memory()
    sd = StreamedDecomposer(eng, 11, 130, 203, 3, chunk=2, buffers=2, bg_downsample=3, packed_mask=True)   # more chunks than buffers
    out = sd.alloc_outputs()
    sd.run(host, out, 25, denoise_mask='fast')
    assert out['mask'].shape == (11, 130, (203 + 7) // 8)
    assert np.array_equal(out['mask'].numpy(), np.packbits(ref['mask'], axis=-1))
    assert np.array_equal(out['fg'].numpy().reshape(ref['fg'].shape), ref['fg'])
    assert np.array_equal(out['bg'].numpy().reshape(ref['bg'].shape), ref['bg'])


def test_streamed_decomposer_packed_transport_returns_the_bool_plane(eng, synth):
    """mask_transport='packed': 1 bit per pixel over the bus, worker threads expand it on the host
    (b200mrc_host_unpack_mask) -- the caller gets the same bool plane, fg and bg as with the default transport."""
    import torch
    import archive_pdf_tools_b200 as pkg
    from archive_pdf_tools_b200.engine import StreamedDecomposer
    pages = np.stack([synth.make_page(500 + i, 130, 203, dpi=100) for i in range(11)])
    ref = pkg.decompose_pages(pages, dpi=100, bg_downsample=3, denoise_mask='fast')
    host = torch.from_numpy(pages).pin_memory()
    sd = StreamedDecomposer(eng, 11, 130, 203, 3, chunk=2, buffers=2, bg_downsample=3, mask_transport='packed', unpack_workers=3)
    outs = [sd.alloc_outputs() for _ in range(2)]
    for o in outs:
        o['mask'].fill_(9)
    pend = [sd.run_async(host, o, 25, denoise_mask='fast') for o in outs]            # two calls in flight
    for p, o in zip(pend, outs):
        p.synchronize()
        assert p.query()
        assert o['mask'].shape == (11, 130, 203)
        assert np.array_equal(o['mask'].numpy(), ref['mask'].view(np.uint8))
        assert np.array_equal(o['fg'].numpy().reshape(ref['fg'].shape), ref['fg'])
        assert np.array_equal(o['bg'].numpy().reshape(ref['bg'].shape), ref['bg'])
    sd.close()


@pytest.mark.parametrize('form', ['auto', 'single-tma', 'single-async', 'trio'])
def test_bg_thumbnail_following_the_sweep(eng, synth, orc, tuning, form):
    """The bg thumbnail pass runs beside the sweep (programmatic dependent launch) and reads bg rows as the sweep's
    strips publish them: same bytes as the serialized form and as the oracle, for every form of the sweep."""
    import archive_pdf_tools_b200 as pkg
    if form != 'auto':
        mode, _, feed = form.partition('-')
        tuning('IIRW_MODE', mode)
        if feed:
            tuning('IIRW_FEED', feed)
    pages = np.stack([synth.make_page(300 + i, 1210, 1000, dpi=300, sigma_n=3.0, halftone=(i == 2)) for i in range(5)])
    tuning('BG_FOLLOW', 0)
    ref = pkg.decompose_pages(pages, dpi=300, bg_downsample=3, denoise_mask='fast')
    tuning('BG_FOLLOW', 2)                       # 2 = follow at every batch size (1, the default: only when the sweep fills the GPU)
    for _ in range(4):
        res = pkg.decompose_pages(pages, dpi=300, bg_downsample=3, denoise_mask='fast')
        assert np.array_equal(res['bg'], ref['bg']) and np.array_equal(res['fg'], ref['fg'])
    exp = orc.decompose(pages[2], dpi=300, bg_downsample=3, denoise_mask='fast')
    assert np.array_equal(res['bg'][2], exp['bg'])
    # fg and bg both thumbnailed: the follower still takes the bg, the fg pass runs behind it
    res2 = pkg.decompose_pages(pages[:2], dpi=300, bg_downsample=3, fg_downsample=2, denoise_mask='fast')
    tuning('BG_FOLLOW', 0)
    ref2 = pkg.decompose_pages(pages[:2], dpi=300, bg_downsample=3, fg_downsample=2, denoise_mask='fast')
    assert np.array_equal(res2['bg'], ref2['bg']) and np.array_equal(res2['fg'], ref2['fg'])


def test_bg_follower_with_more_strips_than_fit_on_the_gpu(eng, synth, orc, tuning):
    """700 small pages x 3 strips: the sweep's CTAs do not all fit on the GPU at once, so its dependents start late
    (after the last sweep CTA has started) -- results must not change, and nothing may dead-lock."""
    import archive_pdf_tools_b200 as pkg
    distinct = [synth.make_page(400 + i, 120, 300, dpi=100, sigma_n=3.0) for i in range(7)]
    pages = np.stack([distinct[i % 7] for i in range(700)])
    tuning('IIRW_MODE', 'single')
    res = pkg.decompose_pages(pages, dpi=100, bg_downsample=3, denoise_mask='fast')
    tuning('BG_FOLLOW', 0)
    ref = pkg.decompose_pages(pages, dpi=100, bg_downsample=3, denoise_mask='fast')
    assert res['bg'].shape == (700, 40, 100, 3)
    assert np.array_equal(res['bg'], ref['bg']) and np.array_equal(res['fg'], ref['fg']) and np.array_equal(res['mask'], ref['mask'])
    for i in (0, 3, 699):
        exp = orc.decompose(pages[i], dpi=100, bg_downsample=3, denoise_mask='fast')
        assert np.array_equal(res['bg'][i], exp['bg']) and np.array_equal(res['fg'][i], exp['fg'])
