"""archive-pdf-tools_b200 -- B200-native MRC page-decomposition engine behind the
internetarchivepdf.mrc pixel-path surface (see DESIGN.md / INTEGRATION.md).

    from archive_pdf_tools_b200 import mrc
    gen = mrc.create_mrc_hocr_components(pil_image, [], dpi=400, bg_downsample=3, denoise_mask='fast')

`install()` makes an unmodified archive-pdf-tools checkout use this engine: it publishes the
drop-in `sauvola` / `optimiser` modules (mrc.py:36-37 import them by bare name) and, if
internetarchivepdf is already importable, rebinds its two pixel-path callables.
"""
import os
import sys

from . import _lib
from ._lib import B200MrcError, LIB_PATH
from .engine import (MrcEngine, Plane, ThumbnailPlan, DecomposeBatch, window_for_dpi,
                     DENOISE_NONE, DENOISE_FAST, DENOISE_BREGMAN)
from .mrc import (threshold_image, create_mrc_hocr_components, create_hocr_mask, decompose_pages, downsample_image,
                  packed_mask, get_engine)
from .grayconvert import special_gray_convert

__version__ = '0.1.0'

DROPIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'dropin')


def install(patch_reference=True):
    """Route the reference's pixel path through this engine (call before importing
    internetarchivepdf.recode, which binds the names at import time: recode.py:39-40)."""
    if DROPIN_DIR not in sys.path:
        sys.path.insert(0, DROPIN_DIR)
    for name in ('sauvola', 'optimiser'):
        sys.modules.pop(name, None)
    import sauvola, optimiser  # noqa: F401  (the drop-in modules)
    if patch_reference:
        ref = sys.modules.get('internetarchivepdf.mrc')
        if ref is not None:
            ref.threshold_image = threshold_image
            ref.create_mrc_hocr_components = create_mrc_hocr_components
            ref.binarise_sauvola = sauvola.binarise_sauvola
            for n in ('optimise_gray', 'optimise_rgb', 'optimise_gray2', 'optimise_rgb2', 'fast_mask_denoise'):
                setattr(ref, n, getattr(optimiser, n))
        gc = sys.modules.get('internetarchivepdf.grayconvert')
        if gc is not None:
            gc.special_gray_convert = special_gray_convert
