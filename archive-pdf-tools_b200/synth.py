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


def make_page(index, H, W, dpi=400, rgb=True, sigma_n=3.0, halftone=False, seed_base=SEED_BASE):
    """uint8 page, (H, W, 3) if rgb else (H, W)."""
    rng = np.random.Generator(np.random.PCG64(seed_base + index))
    paper = rng.uniform(225, 245)
    ax, ay = rng.uniform(-4, 4, 2)
    yy = np.linspace(-1, 1, H, dtype=np.float32)[:, None]
    xx = np.linspace(-1, 1, W, dtype=np.float32)[None, :]
    page = (paper + ax * xx + ay * yy).astype(np.float32)

    sw = max(1, int(round(dpi / 100)))
    pitch = max(4, int(round(0.17 * dpi)))
    gh = max(3, int(round(0.085 * dpi)))
    gw = max(3, int(round(0.06 * dpi)))
    adv = gw + max(1, sw)
    atlas = _glyph_atlas(rng, gh, gw, sw)
    margin_y, margin_x = min(H // 10, pitch * 2), min(W // 10, adv * 4)
    n_lines = max(0, (H - 2 * margin_y - gh) // pitch)
    n_glyphs = max(0, (W - 2 * margin_x - gw) // adv)
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
