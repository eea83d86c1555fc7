"""Synthetic scanned-page generator: the inputs for tests and bench.py.

Deterministic: numpy Generator(PCG64(20240000 + page_index)).  A page is paper U(225,245) with a
smooth +-8 illumination ramp + N(0, sigma_n) sensor noise + text lines at 0.17 in pitch made of
glyph-like dark strokes (level 30-70, stroke ~ dpi/100 px) covering ~5-8 % of the pixels;
"halftone" pages add 1-2 clustered-dot blocks (period ~ dpi/85 px) over ~25 % of the page.
RGB = gray with a fixed warm tint (R+0, G-5, B-15, clipped).
"""
import numpy as np

SEED_BASE = 20240000


def _glyph_atlas(rng, gh, gw, sw, count=48):
    atlas = np.zeros((count, gh, gw), bool)
    for g in range(count):
        a = atlas[g]
        for _ in range(rng.integers(2, 5)):
            kind = rng.integers(0, 3)
            if kind == 0:        # vertical stem
                x = rng.integers(0, max(1, gw - sw))
                y0 = rng.integers(0, gh // 2)
                a[y0:rng.integers(gh // 2, gh) + 1, x:x + sw] = True
            elif kind == 1:      # horizontal bar
                y = rng.integers(0, max(1, gh - sw))
                x0 = rng.integers(0, gw // 2)
                a[y:y + sw, x0:rng.integers(gw // 2, gw) + 1] = True
            else:                # diagonal
                n = min(gh, gw) - sw
                o = rng.integers(0, 2)
                for t in range(max(n, 1)):
                    xx = t if o else gw - sw - t
                    a[t:t + sw, max(xx, 0):max(xx, 0) + sw] = True
    return atlas


def _layout(H, W, dpi):
    sw = max(1, int(round(dpi / 100)))
    pitch = max(4, int(round(0.17 * dpi)))
    gh = max(3, int(round(0.085 * dpi)))
    gw = max(3, int(round(0.06 * dpi)))
    adv = gw + max(1, sw)
    margin_y, margin_x = min(H // 10, pitch * 2), min(W // 10, adv * 4)
    n_lines = max(0, (H - 2 * margin_y - gh) // pitch)
    n_glyphs = max(0, (W - 2 * margin_x - gw) // adv)
    return sw, pitch, gh, gw, adv, margin_y, margin_x, n_lines, n_glyphs


def page_hocr(H, W, dpi=400, pad=None, low_conf_every=7, scale=1.0):
    """hocr_word_data (the structure hocr.parse.hocr_page_to_word_data returns, as far as
    create_hocr_mask reads it: paragraphs -> lines -> bbox + words with text / confidence) for the text
    lines make_page draws.  Every `low_conf_every`-th line gets confidence 10 (skipped by the reference),
    one line has empty text, one box is degenerate and one leaves the page.  `scale` multiplies the
    coordinates (use it together with `downsample`)."""
    sw, pitch, gh, gw, adv, margin_y, margin_x, n_lines, n_glyphs = _layout(H, W, dpi)
    pad = max(1, sw) if pad is None else pad
    lines = []
    for i in range(n_lines):
        top = margin_y + i * pitch - pad
        box = [margin_x - pad, top, margin_x + n_glyphs * adv + pad, top + gh + 2 * pad]
        box = [max(0, box[0]), max(0, box[1]), min(W, box[2]), min(H, box[3])]
        conf = 10 if low_conf_every and i % low_conf_every == low_conf_every - 1 else 90
        words = [{'text': 'lorem', 'confidence': conf}, {'text': 'ipsum', 'confidence': conf}]
        lines.append({'bbox': [c * scale + 0.25 for c in box], 'words': words})
    extra = [{'bbox': [5.0, 5.0, 40.0, 20.0], 'words': [{'text': '  ', 'confidence': 95}]},
             {'bbox': [10.0, 10.0, 10.0, 30.0], 'words': [{'text': 'x', 'confidence': 95}]},
             {'bbox': [W - 10.0, 10.0, W + 30.0, 30.0], 'words': [{'text': 'x', 'confidence': 95}]},
             {'bbox': [30.0, 40.0, 20.0, 50.0], 'words': [{'text': 'x', 'confidence': 95}]}]
    half = len(lines) // 2
    return [{'lines': lines[:half]}, {'lines': extra}, {'lines': lines[half:]}]


def make_page(index, H, W, dpi=400, rgb=True, sigma_n=3.0, halftone=False, seed_base=SEED_BASE, invert_lines=(),
              noisy_dark_lines=()):
    """uint8 page, (H, W, 3) if rgb else (H, W).  invert_lines: indices of text lines drawn light-on-dark;
    noisy_dark_lines: lines replaced by light blocks on a dark, very noisy band (the case in which
    create_hocr_mask picks the inverted polarity through the sigma tie-break)."""
    rng = np.random.Generator(np.random.PCG64(seed_base + index))
    paper = rng.uniform(225, 245)
    ax, ay = rng.uniform(-4, 4, 2)
    yy = np.linspace(-1, 1, H, dtype=np.float32)[:, None]
    xx = np.linspace(-1, 1, W, dtype=np.float32)[None, :]
    page = (paper + ax * xx + ay * yy).astype(np.float32)

    sw, pitch, gh, gw, adv, margin_y, margin_x, n_lines, n_glyphs = _layout(H, W, dpi)
    atlas = _glyph_atlas(rng, gh, gw, sw)
    if n_lines and n_glyphs:
        idx = rng.integers(0, atlas.shape[0], (n_lines, n_glyphs))
        space = rng.random((n_lines, n_glyphs)) < 0.15           # word gaps
        ink = rng.uniform(30, 70, (n_lines, n_glyphs)).astype(np.float32)
        cells = atlas[idx] & ~space[:, :, None, None]            # L, G, gh, gw
        cell_full = np.zeros((n_lines, n_glyphs, pitch, adv), bool)
        cell_full[:, :, :gh, :gw] = cells
        text = cell_full.transpose(0, 2, 1, 3).reshape(n_lines * pitch, n_glyphs * adv)
        inkmap = np.broadcast_to(ink[:, None, :, None], (n_lines, pitch, n_glyphs, adv)).reshape(text.shape)
        region = page[margin_y:margin_y + text.shape[0], margin_x:margin_x + text.shape[1]]
        np.copyto(region, inkmap, where=text)

    for i in invert_lines:                                       # a light-on-dark band around text line i
        if 0 <= i < n_lines:
            y0 = max(0, margin_y + i * pitch - 2 * sw)
            band = page[y0:min(H, y0 + gh + 4 * sw), :]
            np.subtract(np.float32(255), band, out=band)

    for i in noisy_dark_lines:
        if 0 <= i < n_lines:
            r2 = np.random.Generator(np.random.PCG64(seed_base + index * 1000 + 500 + i))
            y0 = max(0, margin_y + i * pitch - 2 * sw)
            band = page[y0:min(H, y0 + gh + 4 * sw), :]
            band[...] = 60 + r2.standard_normal(band.shape, dtype=np.float32) * 30
            step = max(8, 4 * adv)
            for x in range(margin_x, W - margin_x - step // 3, step):
                band[sw:band.shape[0] - sw, x:x + max(2, int(step * 0.22))] = 230

    if halftone:
        period = max(2, int(round(dpi / 85)))
        for _ in range(rng.integers(1, 3)):
            bh, bw = int(H * rng.uniform(0.3, 0.5)), int(W * rng.uniform(0.3, 0.5))
            y0, x0 = rng.integers(0, H - bh + 1), rng.integers(0, W - bw + 1)
            ys = (np.arange(bh) % period)[:, None] - (period - 1) / 2
            xs = (np.arange(bw) % period)[None, :] - (period - 1) / 2
            tone = rng.uniform(0.25, 0.6) * period
            dots = (ys * ys + xs * xs) < tone * tone / 2
            blk = page[y0:y0 + bh, x0:x0 + bw]
            np.copyto(blk, np.float32(rng.uniform(40, 90)), where=dots)

    if sigma_n > 0:
        page += rng.standard_normal(page.shape, dtype=np.float32) * np.float32(sigma_n)
    gray = np.clip(np.rint(page), 0, 255).astype(np.uint8)
    if not rgb:
        return gray
    out = np.empty((H, W, 3), np.uint8)
    out[..., 0] = gray
    out[..., 1] = np.clip(gray.astype(np.int16) - 5, 0, 255).astype(np.uint8)
    out[..., 2] = np.clip(gray.astype(np.int16) - 15, 0, 255).astype(np.uint8)
    return out


def make_pages(start, count, H, W, **kw):
    return np.stack([make_page(start + i, H, W, **kw) for i in range(count)])
