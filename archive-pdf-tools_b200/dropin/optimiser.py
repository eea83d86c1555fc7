"""Drop-in for the reference's top-level Cython module `optimiser` (cython/optimiser.pyx),
imported by bare name at internetarchivepdf/mrc.py:36.  Same six function names and calling
conventions; the arithmetic runs in libb200mrc.so on the GPU (no CPU fallback)."""
import numpy as np

import archive_pdf_tools_b200 as _pkg
from archive_pdf_tools_b200 import engine as _E


def _check_u8(a, ndim):
    if not isinstance(a, np.ndarray):
        raise TypeError('Argument has incorrect type (expected numpy.ndarray, got %s)' % type(a).__name__)
    if a.ndim != ndim:
        raise ValueError('Buffer has wrong number of dimensions (expected %d, got %d)' % (ndim, a.ndim))
    if a.dtype.itemsize != 1 or a.dtype.kind not in 'ub':
        raise ValueError("Buffer dtype mismatch, expected 'UINT8DTYPE_t' but got '%s'" % a.dtype)
    return a


def _optimise(mask, img, width, height, n_size, ndim):
    mask = _check_u8(mask, 2)
    img = _check_u8(img, ndim)
    eng = _pkg.get_engine()
    c = 1 if ndim == 2 else 3
    m = _E.Plane(1, height, width, 1, eng.device).upload(np.ascontiguousarray(mask[:height, :width]).view(np.uint8)[None], non_blocking=False)
    src = _E.Plane(1, height, width, c, eng.device).upload(np.ascontiguousarray(img[:height, :width]).view(np.uint8)[None], non_blocking=False)
    out = _E.Plane(1, height, width, c, eng.device)
    eng.optimise(m, src, out_fg=out, n_fg=n_size, out_bg=None)
    return out.numpy()[0]


def optimise_gray(mask, img, width, height, n_size):
    """cython/optimiser.pyx:22"""
    return _optimise(mask, img, width, height, n_size, 2)


def optimise_gray2(mask, img, width, height, n_size):
    """cython/optimiser.pyx:153"""
    return _optimise(mask, img, width, height, n_size, 2)


def optimise_rgb(mask, img, width, height, n_size):
    """cython/optimiser.pyx:83"""
    return _optimise(mask, img, width, height, n_size, 3)


def optimise_rgb2(mask, img, width, height, n_size):
    """cython/optimiser.pyx:280"""
    return _optimise(mask, img, width, height, n_size, 3)


def fast_mask_denoise(mask, width, height, mincnt, n_size):
    """cython/optimiser.pyx:436 -- in place; returns the same array object."""
    mask = _check_u8(mask, 2)
    if width <= 0 or height <= 0:
        return mask
    eng = _pkg.get_engine()
    m = _E.Plane(1, height, width, 1, eng.device).upload(np.ascontiguousarray(mask[:height, :width]).view(np.uint8)[None], non_blocking=False)
    eng.denoise(m, mincnt, n_size)
    mask.view(np.uint8)[:height, :width] = m.numpy()[0]
    return mask
