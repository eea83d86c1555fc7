"""GPU parity tests, kernel by kernel, through the C-ABI (libb200mrc.so) against the CPU oracle.
Bar: bit-exact for every integer/byte result (masks, gray, fg/bg, thumbnails, sigma)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _plane(eng, arr, c=None):
    from archive_pdf_tools_b200 import Plane
    a = np.ascontiguousarray(arr)
    if a.dtype == np.bool_:
        a = a.view(np.uint8)
    if c is None:
        c = a.shape[3] if a.ndim == 4 else 1
    n, h, w = a.shape[:3]
    return Plane(n, h, w, c, eng.device).upload(a, non_blocking=False)


def _empty(eng, n, h, w, c=1):
    from archive_pdf_tools_b200 import Plane
    return Plane(n, h, w, c, eng.device)


def _imgs(rng, h, w):
    return [rng.integers(0, 256, (h, w), dtype=np.uint8),
            (rng.integers(0, 2, (h, w)) * 255).astype(np.uint8),
            np.clip(rng.normal(200, 30, (h, w)), 0, 255).astype(np.uint8),
            np.full((h, w), 255, np.uint8), np.zeros((h, w), np.uint8),
            ((np.add.outer(np.arange(h), np.arange(w)) % 2) * 255).astype(np.uint8)]


@pytest.mark.parametrize('shape', [(1, 1), (3, 7), (10, 10), (40, 200), (200, 40), (77, 133), (300, 1003), (131, 2000)])
def test_sauvola_matches_oracle(eng, orc, shape):
    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    imgs = np.stack(_imgs(rng, h, w))
    src = _plane(eng, imgs)
    dst = _empty(eng, len(imgs), h, w)
    for ww, wh, k in [(3, 3, 0.34), (33, 33, 0.34), (51, 51, 0.1), (101, 101, 0.34), (151, 151, 0.2), (255, 255, 0.34),
                      (33, 51, 0.34), (4, 6, 0.5), (75, 75, -0.2)]:
        eng.sauvola(src, dst, ww, wh, k=k)
        got = dst.numpy(np.bool_)
        for i in range(len(imgs)):
            exp = orc.sauvola(imgs[i], ww, wh, k=k)
            assert np.array_equal(got[i], exp), (shape, ww, wh, k, i, int((got[i] != exp).sum()))


def test_sauvola_flags_and_R(eng, orc):
    from archive_pdf_tools_b200 import _lib
    rng = np.random.default_rng(5)
    img = np.clip(rng.normal(170, 50, (1, 90, 150)), 0, 255).astype(np.uint8)
    src = _plane(eng, img)
    dst = _empty(eng, 1, 90, 150)
    exp = orc.sauvola(img[0], 33, k=0.34, R=100.0)
    eng.sauvola(src, dst, 33, k=0.34, R=100.0)
    assert np.array_equal(dst.numpy(np.bool_)[0], exp)
    eng.sauvola(src, dst, 33, k=0.34, R=100.0, flags=_lib.SAUVOLA_RAW_INVERTED)
    assert np.array_equal(dst.numpy(np.bool_)[0], ~exp)
    pre = rng.random((1, 90, 150)) < 0.2
    dst.upload(pre, non_blocking=False)
    eng.sauvola(src, dst, 33, k=0.34, R=100.0, flags=_lib.SAUVOLA_OR_INTO)
    assert np.array_equal(dst.numpy(np.bool_)[0], exp | pre[0])


def test_rgb2gray_all_colours(eng, orc):
    allc = np.stack(np.meshgrid(np.arange(256), np.arange(256), np.arange(256), indexing='ij'), -1)
    allc = allc.reshape(1, 4096, 4096, 3).astype(np.uint8)
    src = _plane(eng, allc)
    g = _empty(eng, 1, 4096, 4096)
    eng.gray_blur(src, g, None)
    assert np.array_equal(g.numpy()[0], orc.rgb2gray(allc[0]))


@pytest.mark.parametrize('shape,rgb', [((64, 64), False), ((2, 9), False), ((3, 3), True), ((1, 1), False), ((330, 255), True), ((513, 771), False)])
def test_noise_sigma_matches_oracle(eng, orc, synth, shape, rgb):
    h, w = shape
    pages = np.stack([synth.make_page(i, h, w, dpi=100, rgb=rgb, sigma_n=s) for i, s in enumerate((0.0, 0.7, 3.0, 12.0))])
    src = _plane(eng, pages)
    got = eng.estimate_noise(src).cpu().numpy()
    for i in range(len(pages)):
        gray = pages[i] if not rgb else orc.rgb2gray(pages[i])
        exp = orc.estimate_noise(gray)
        assert (np.isnan(exp) and np.isnan(got[i])) or got[i] == exp, (shape, i, got[i], exp)


def test_noise_sigma_flat_page_is_nan(eng, orc):
    src = _plane(eng, np.full((1, 40, 40), 128, np.uint8))
    got = eng.estimate_noise(src).cpu().numpy()[0]
    exp = orc.estimate_noise(np.full((40, 40), 128, np.uint8))
    assert (np.isnan(got) and np.isnan(exp)) or got == exp


@pytest.mark.parametrize('shape,rgb', [((1, 1), False), ((5, 9), True), ((60, 80), False), ((151, 260), True)])
def test_gray_blur_matches_oracle(eng, orc, shape, rgb):
    import torch
    h, w = shape
    rng = np.random.default_rng(h + w)
    sig_est = [0.5, 1.0, 1.1, 1.3, 3.7, 9.0, 21.0, 41.0, 44.0, 90.0, float('nan')]
    n = len(sig_est)
    pages = rng.integers(0, 256, (n, h, w, 3) if rgb else (n, h, w), dtype=np.uint8)
    src = _plane(eng, pages)
    dst = _empty(eng, n, h, w)
    eng.gray_blur(src, dst, torch.tensor(sig_est, dtype=torch.float64, device=eng.device))
    got = dst.numpy()
    for i, s in enumerate(sig_est):
        gray = pages[i] if not rgb else orc.rgb2gray(pages[i])
        exp = gray
        if s > 1.0:
            exp = orc.gauss_blur(gray.astype(np.float32), s * 0.1).astype(np.uint8)
        assert np.array_equal(got[i], exp), (shape, s, int((got[i] != exp).sum()))


def _masks(rng, h, w):
    ms = [rng.random((h, w)) < d for d in (0.02, 0.1, 0.3, 0.6, 0.95)]
    ms.append(np.eye(h, w, dtype=bool) | np.eye(h, w, 1, dtype=bool))
    ms.append(np.ones((h, w), bool)); ms.append(np.zeros((h, w), bool))
    d = np.zeros((h, w), bool); d[::3, :] = True; ms.append(d)
    return ms


@pytest.mark.parametrize('shape', [(1, 1), (4, 4), (5, 5), (6, 9), (50, 70), (70, 300), (200, 600), (131, 97), (64, 256), (65, 257)])
def test_denoise_matches_oracle(eng, orc, shape):
    h, w = shape
    rng = np.random.default_rng(h * 7 + w)
    ms = np.stack(_masks(rng, h, w))
    pl = _plane(eng, ms)
    eng.denoise(pl, 4, 2)
    got = pl.numpy(np.bool_)
    for i in range(len(ms)):
        exp = orc.denoise(ms[i])
        assert np.array_equal(got[i], exp), (shape, i, int((got[i] != exp).sum()))


@pytest.mark.parametrize('mincnt,n_size', [(1, 1), (5, 1), (8, 3), (12, 2), (0, 2), (1, 0), (40, 4)])
def test_denoise_general_parameters_match_oracle(eng, orc, mincnt, n_size):
    """fast_mask_denoise(mask, w, h, mincnt, n_size) for parameters the reference never uses: the general
    one-thread-per-pixel form of the same fixed point (several pages, a pitched plane, a cascade)."""
    rng = np.random.default_rng(mincnt * 10 + n_size)
    h, w = 83, 131
    ms = [rng.random((h, w)) < d for d in (0.05, 0.3, 0.6, 0.95)]
    diag = np.zeros((h, w), bool)
    idx = np.arange(min(h, w))
    diag[idx, idx] = True; diag[idx[:-1], idx[:-1] + 1] = True
    ms.append(diag)
    ms = np.stack(ms)
    pl = _plane(eng, ms)
    eng.denoise(pl, mincnt, n_size)
    got = pl.numpy(np.bool_)
    for i in range(len(ms)):
        exp = orc.denoise(ms[i], mincnt, n_size)
        assert np.array_equal(got[i], exp), (mincnt, n_size, i, int((got[i] != exp).sum()))


def test_denoise_long_cascade(eng, orc):
    # a 2-px diagonal band: every removal exposes the next pixel -> O(H) dependency chain across tiles
    h, w = 700, 700
    m = np.zeros((h, w), bool)
    idx = np.arange(h)
    m[idx, idx] = True
    m[idx[:-1], idx[:-1] + 1] = True
    pl = _plane(eng, m[None])
    eng.denoise(pl, 4, 2)
    assert np.array_equal(pl.numpy(np.bool_)[0], orc.denoise(m))


@pytest.mark.parametrize('shape,c', [((1, 1), 1), ((2, 3), 3), ((5, 5), 1), ((30, 41), 3), ((64, 200), 3), ((90, 700), 1), ((131, 97), 3)])
def test_optimise_matches_oracle(eng, orc, shape, c):
    h, w = shape
    rng = np.random.default_rng(h * 13 + w)
    dens = (0.0, 0.05, 0.5, 0.95, 1.0)
    masks = np.stack([rng.random((h, w)) < d for d in dens])
    imgs = rng.integers(0, 256, (len(dens), h, w, 3) if c == 3 else (len(dens), h, w), dtype=np.uint8)
    m = _plane(eng, masks)
    src = _plane(eng, imgs)
    fg = _empty(eng, len(dens), h, w, c)
    bg = _empty(eng, len(dens), h, w, c)
    for nfg, nbg in [(3, 10), (1, 16), (10, 3)]:
        eng.optimise(m, src, fg, nfg, bg, nbg)
        gf, gb = fg.numpy(), bg.numpy()
        for i in range(len(dens)):
            ef = orc.optimise(masks[i], imgs[i], nfg)
            eb = orc.optimise(~masks[i], imgs[i], nbg)
            assert np.array_equal(gf[i], ef), ('fg', shape, c, nfg, i, int((gf[i] != ef).sum()))
            assert np.array_equal(gb[i], eb), ('bg', shape, c, nbg, i, int((gb[i] != eb).sum()))
    eng.optimise(m, src, out_fg=fg, n_fg=5, out_bg=None)          # single reference call form
    assert np.array_equal(fg.numpy()[2], orc.optimise(masks[2], imgs[2], 5))


def test_optimise_text_page_many_strips(eng, orc, synth):
    page = synth.make_page(9, 500, 1300, dpi=200)
    gray = orc.rgb2gray(page)
    mask = orc.denoise(orc.sauvola(gray, 51))
    m = _plane(eng, mask[None]); src = _plane(eng, page[None])
    fg = _empty(eng, 1, 500, 1300, 3); bg = _empty(eng, 1, 500, 1300, 3)
    eng.optimise(m, src, fg, 3, bg, 10)
    assert np.array_equal(fg.numpy()[0], orc.optimise(mask, page, 3))
    assert np.array_equal(bg.numpy()[0], orc.optimise(~mask, page, 10))


@pytest.mark.parametrize('variant', [{'IIRW_MODE': 'trio'}, {'IIRW_MODE': 'trio', 'IIRW_TPC': '1'},
                                     {'IIRW_MODE': 'trio', 'IIRW_TPC': '3'},
                                     {'IIRW_MODE': 'single', 'IIRW_FEED': 'tma'},
                                     {'IIRW_MODE': 'single', 'IIRW_FEED': 'async'}],
                         ids=['trio', 'trio1', 'trio3', 'single-tma', 'single-async'])
def test_optimise_sweep_variants(eng, orc, synth, tuning, variant):
    """Every form of the row-sequential sweep (launcher picks by batch size; forced here) against the oracle:
    ragged strip widths, many strips, gray and RGB, random masks and a text page, several pages per launch."""
    for k, v in variant.items():
        tuning(k, v)
    rng = np.random.default_rng(77)
    for (h, w), c in [((37, 128), 3), ((64, 129), 3), ((90, 700), 1), ((131, 397), 3), ((70, 1027), 1), ((1, 300), 3), ((5, 4), 3)]:
        dens = (0.0, 0.05, 0.5, 1.0)
        masks = np.stack([rng.random((h, w)) < d for d in dens])
        imgs = rng.integers(0, 256, (len(dens), h, w, 3) if c == 3 else (len(dens), h, w), dtype=np.uint8)
        m = _plane(eng, masks); src = _plane(eng, imgs)
        fg = _empty(eng, len(dens), h, w, c); bg = _empty(eng, len(dens), h, w, c)
        for rep in range(2):                                  # second launch: mailbox reuse under a new epoch
            eng.optimise(m, src, fg, 3, bg, 10)
            gf, gb = fg.numpy(), bg.numpy()
            for i in range(len(dens)):
                assert np.array_equal(gf[i], orc.optimise(masks[i], imgs[i], 3)), ('fg', variant, h, w, c, i, rep)
                assert np.array_equal(gb[i], orc.optimise(~masks[i], imgs[i], 10)), ('bg', variant, h, w, c, i, rep)
    pages = np.stack([synth.make_page(20 + i, 420, 1300, dpi=200, halftone=(i == 1)) for i in range(3)])
    mks = np.stack([orc.denoise(orc.sauvola(orc.rgb2gray(pg), 51)) for pg in pages])
    m = _plane(eng, mks); src = _plane(eng, pages)
    fg = _empty(eng, 3, 420, 1300, 3); bg = _empty(eng, 3, 420, 1300, 3)
    eng.optimise(m, src, fg, 3, bg, 10)
    for i in range(3):
        assert np.array_equal(fg.numpy()[i], orc.optimise(mks[i], pages[i], 3))
        assert np.array_equal(bg.numpy()[i], orc.optimise(~mks[i], pages[i], 10))


@pytest.mark.parametrize('fill', [0xFF, 0xEE, 0x11])
def test_optimise_mailbox_survives_foreign_workspace_bytes(eng, orc, tuning, fill):
    """The sweep's strip hand-off rows live in a workspace that other stages (and callers) may overwrite between
    launches.  Every launch epoch 1..255 (and the wrap) is run on a workspace pre-filled with a byte pattern whose
    tag nibbles equal some epoch: the FIR pass clears the rows, so no stale or foreign word can be taken for a hand-off."""
    import torch
    rng = np.random.default_rng(fill)
    h, w = 24, 300                                                # 3 strips of 128 columns
    masks = np.stack([rng.random((h, w)) < d for d in (0.05, 0.5)])
    imgs = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
    m = _plane(eng, masks); src = _plane(eng, imgs)
    fg = _empty(eng, 2, h, w, 3); bg = _empty(eng, 2, h, w, 3)
    exp_f = [orc.optimise(masks[i], imgs[i], 3) for i in range(2)]
    exp_b = [orc.optimise(~masks[i], imgs[i], 10) for i in range(2)]
    eng.optimise(m, src, fg, 3, bg, 10)                           # allocates the workspace
    ws = eng._ws[('optimise', torch.cuda.current_stream(eng.device).cuda_stream)]
    for mode in ('trio', 'single'):
        tuning('IIRW_MODE', mode)
        for it in range(260):
            ws.fill_(fill)
            eng.optimise(m, src, fg, 3, bg, 10)
            if it % 13 == 0 or it > 250:
                gf, gb = fg.numpy(), bg.numpy()
                for i in range(2):
                    assert np.array_equal(gf[i], exp_f[i]) and np.array_equal(gb[i], exp_b[i]), (mode, it, i)


@pytest.mark.parametrize('path', ['fused', 'legacy'])
def test_sauvola_both_paths(eng, orc, tuning, path):
    """Gray planes with 16-byte aligned rows take the 8-columns-per-thread TMA-fed kernel (sauvola_fused.cu, direct form);
    everything else -- and THRESHOLD_PATH=legacy -- the general 4-byte kernel (sauvola.cu).  Window widths around the
    warp spans (127..129), widest window, several strips per row."""
    tuning('THRESHOLD_PATH', path)
    rng = np.random.default_rng(11)
    for h, w in [(150, 2600), (260, 1100), (40, 130)]:
        img = np.clip(rng.normal(190, 45, (2, h, w)), 0, 255).astype(np.uint8)
        src = _plane(eng, img); dst = _empty(eng, 2, h, w)
        for ww, k in [(33, 0.34), (101, 0.34), (127, 0.34), (128, 0.2), (129, 0.34), (151, 0.34), (255, 0.1)]:
            eng.sauvola(src, dst, ww, ww, k=k)
            got = dst.numpy(np.bool_)
            for i in range(2):
                assert np.array_equal(got[i], orc.sauvola(img[i], ww, ww, k=k)), (path, h, w, ww, i)


def test_sauvola_decision_table_over_k_and_R(eng, orc, tuning):
    """The fused kernel decides through a table vmin[mean][pixel] built per (k, R) (k_sauvola_vmin).  Images that reach the
    corners of the (pixel, mean, variance) space -- flat black / white, two-level patterns with every contrast, noise of
    every amplitude -- against the oracle's FP64 test, for k from 0 to beyond 1 and R other than 128."""
    tuning('THRESHOLD_PATH', 'fused')
    rng = np.random.default_rng(5)
    h, w = 96, 1040
    imgs = []
    lo = rng.integers(0, 256, (h, 1)).astype(np.int32); hi = rng.integers(0, 256, (1, w)).astype(np.int32)
    imgs.append(np.where(rng.random((h, w)) < 0.5, lo, hi))                        # two levels of every contrast
    imgs.append(np.clip(rng.normal(128, np.linspace(0, 90, w)[None, :], (h, w)), 0, 255))   # variance ramp
    imgs.append(np.tile(np.arange(w) % 256, (h, 1)))                               # every pixel value against slowly moving means
    imgs.append(np.zeros((h, w))); imgs.append(np.full((h, w), 255))
    imgs.append(np.where((np.add.outer(np.arange(h), np.arange(w)) // 3) % 2 == 0, 0, 255))
    img = np.stack(imgs).astype(np.uint8)
    src = _plane(eng, img); dst = _empty(eng, len(imgs), h, w)
    for k, R in [(0.34, 128.0), (0.1, 128.0), (0.0, 128.0), (0.05, 128.0), (0.5, 100.0), (0.9, 64.5), (1.0, 128.0), (1.7, 128.0), (0.34, 1.0)]:
        for ww in (5, 33, 75):
            eng.sauvola(src, dst, ww, ww, k=k, R=R)
            got = dst.numpy(np.bool_)
            for i in range(len(imgs)):
                exp = orc.sauvola(img[i], ww, ww, k=k, R=R)
                assert np.array_equal(got[i], exp), (k, R, ww, i, int((got[i] != exp).sum()))


def _threshold_expected(orc, page, sig_est, ww, wh, k=0.34):
    gray = page if page.ndim == 2 else orc.rgb2gray(page)
    if sig_est is not None and sig_est > 1.0:
        gray = orc.gauss_blur(gray.astype(np.float32), sig_est * 0.1).astype(np.uint8)
    return orc.sauvola(gray, ww, wh, k=k)


@pytest.mark.parametrize('shape,rgb', [((70, 64), True), ((90, 131), False), ((300, 1003), True), ((131, 2000), False),
                                       ((420, 2550), True), ((64, 3000), True), ((1, 1), True), ((9, 40), False)])
@pytest.mark.parametrize('path', ['fused', 'legacy'])
def test_threshold_mask_matches_oracle(eng, orc, tuning, shape, rgb, path):
    """create_threshold_mask as one fused kernel (b200mrc_threshold_mask): gray conversion, per-page blur decision
    (radius 0, 1, 2 in-kernel; radius > 2 through the tiled pre-blur; NaN / <= 1.0: none), Sauvola.  Small pages take
    the two-pass fallback behind the same entry point."""
    import torch
    tuning('THRESHOLD_PATH', path)
    h, w = shape
    rng = np.random.default_rng(h * 31 + w)
    sig_est = [0.5, 1.1, 1.3, 3.4, 4.0, 6.2, 9.0, 26.0, float('nan')]                    # radii 0 0 1 1 2 2 4 10 -
    n = len(sig_est)
    base = np.clip(rng.normal(180, 50, (n, h, w)), 0, 255)
    base[:, : h // 3, : w // 2] = rng.integers(0, 256, (n, h // 3, w // 2))
    base[:, -(h // 4 + 1):, :] = 255                                                     # a saturated flat band
    if rgb:
        pages = np.clip(base[..., None] + rng.integers(-20, 20, (n, h, w, 3)), 0, 255).astype(np.uint8)
    else:
        pages = base.astype(np.uint8)
    src = _plane(eng, pages)
    dst = _empty(eng, n, h, w)
    sig = torch.tensor(sig_est, dtype=torch.float64, device=eng.device)
    for ww, wh in [(101, 101), (33, 33), (51, 75), (151, 151), (255, 3), (3, 255)]:
        eng.threshold_mask(src, dst, ww, wh, k=0.34, sigma_dev=sig)
        got = dst.numpy(np.bool_)
        for i, s_ in enumerate(sig_est):
            exp = _threshold_expected(orc, pages[i], s_, ww, wh)
            assert np.array_equal(got[i], exp), (shape, rgb, ww, wh, s_, int((got[i] != exp).sum()))
    eng.threshold_mask(src, dst, 75, 75, k=0.2, sigma_dev=None)                           # no sigma array: never blurs
    got = dst.numpy(np.bool_)
    for i in range(n):
        assert np.array_equal(got[i], _threshold_expected(orc, pages[i], None, 75, 75, k=0.2)), (shape, rgb, i)


def test_fused_gray_all_colours(eng, orc, tuning):
    """The fused kernel's RGB -> L (two dp4a on the coefficient bytes) on all 2^24 colours: its gray delay-line plane
    (the threshold workspace) must equal PIL's convert('L')."""
    import torch
    tuning('THRESHOLD_PATH', 'fused')
    allc = np.stack(np.meshgrid(np.arange(256), np.arange(256), np.arange(256), indexing='ij'), -1)
    allc = allc.reshape(1, 4096, 4096, 3).astype(np.uint8)
    src = _plane(eng, allc)
    dst = _empty(eng, 1, 4096, 4096)
    eng.threshold_mask(src, dst, 33, 33)
    torch.cuda.synchronize()
    ws = eng._ws[('threshold', torch.cuda.current_stream(eng.device).cuda_stream)]
    off = (-ws.data_ptr()) % 16
    gray = ws[off:off + 4096 * 4096].reshape(4096, 4096).cpu().numpy()
    exp = orc.rgb2gray(allc[0])
    assert np.array_equal(gray, exp), int((gray != exp).sum())
    assert np.array_equal(dst.numpy(np.bool_)[0], orc.sauvola(exp, 33))


@pytest.mark.parametrize('knobs', [{'FUSED_NT': 128, 'FUSED_BANDS': 1}, {'FUSED_NT': 192, 'FUSED_BANDS': 3},
                                   {'FUSED_NT': 256, 'FUSED_BANDS': 7}, {'FUSED_NT': 128, 'FUSED_BANDS': 40},
                                   {'FUSED_NT': 128, 'FUSED_OCC': 5}, {'FUSED_NT': 128, 'FUSED_OCC': 3, 'FUSED_BANDS': 2}],
                         ids=['nt128-1band', 'nt192-3bands', 'nt256-7bands', 'nt128-40bands', 'occ5', 'occ3'])
def test_threshold_mask_geometries(eng, orc, tuning, knobs):
    """Every CTA width and several band heights (incl. bands shorter than the window) of the fused kernel, flags, and
    byte-identical results across launches (the gray delay line is rewritten by overlapping CTAs with equal bytes)."""
    import torch
    from archive_pdf_tools_b200 import _lib
    tuning('THRESHOLD_PATH', 'fused')
    for k_, v in knobs.items():
        tuning(k_, v)
    rng = np.random.default_rng(3)
    h, w = 500, 1900
    pages = np.clip(rng.normal(170, 60, (3, h, w, 3)), 0, 255).astype(np.uint8)
    sig_est = [0.0, 3.0, 5.5]
    src = _plane(eng, pages); dst = _empty(eng, 3, h, w)
    sig = torch.tensor(sig_est, dtype=torch.float64, device=eng.device)
    exp = [_threshold_expected(orc, pages[i], sig_est[i], 101, 101) for i in range(3)]
    for rep in range(2):
        eng.threshold_mask(src, dst, 101, 101, sigma_dev=sig)
        got = dst.numpy(np.bool_)
        for i in range(3):
            assert np.array_equal(got[i], exp[i]), (knobs, rep, i, int((got[i] != exp[i]).sum()))
    pre = rng.random((3, h, w)) < 0.1
    dst.upload(pre, non_blocking=False)
    eng.threshold_mask(src, dst, 101, 101, sigma_dev=sig, flags=_lib.SAUVOLA_OR_INTO)
    got = dst.numpy(np.bool_)
    for i in range(3):
        assert np.array_equal(got[i], exp[i] | pre[i])
    eng.threshold_mask(src, dst, 101, 101, sigma_dev=sig, flags=_lib.SAUVOLA_RAW_INVERTED)
    got = dst.numpy(np.bool_)
    for i in range(3):
        assert np.array_equal(got[i], ~exp[i])


@pytest.mark.parametrize('shape', [(33, 25), (100, 77), (330, 255), (64, 64), (7, 5), (600, 450)])
def test_thumbnail_matches_oracle_and_pillow(eng, orc, shape):
    from PIL import Image
    from archive_pdf_tools_b200 import ThumbnailPlan
    h, w = shape
    rng = np.random.default_rng(h + 3 * w)
    for f in [2, 3, 4, 5, 6, 8, 1.5, 2.5]:
        for ch in (1, 3):
            imgs = rng.integers(0, 256, (2, h, w, 3) if ch == 3 else (2, h, w), dtype=np.uint8)
            wd, hd = int(w / f), int(h / f)
            if wd <= 0 or hd <= 0:
                continue
            exp = orc.thumbnail(imgs[1], wd, hd)
            im = Image.fromarray(imgs[1]); im.thumbnail((wd, hd))
            assert np.array_equal(exp, np.array(im))
            plan = ThumbnailPlan(w, h, ch, wd, hd)
            if plan.noop:
                assert exp.shape[:2] == (h, w)
                continue
            assert (plan.out_h, plan.out_w) == exp.shape[:2], (shape, f)
            src = _plane(eng, imgs)
            dst = _empty(eng, 2, plan.out_h, plan.out_w, ch)
            eng.resample(plan, src, dst)
            assert np.array_equal(dst.numpy()[1], exp), (shape, f, ch)


def test_special_gray_matches_oracle(eng, orc, synth):
    import archive_pdf_tools_b200 as pkg
    rng = np.random.default_rng(4)
    for img in (synth.make_page(2, 200, 150, dpi=100), rng.integers(0, 256, (97, 131, 3), dtype=np.uint8),
                rng.integers(40, 200, (64, 64, 3), dtype=np.uint8)):
        assert np.array_equal(pkg.special_gray_convert(img), orc.special_gray_convert(img))
