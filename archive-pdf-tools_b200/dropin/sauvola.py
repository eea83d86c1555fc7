"""Drop-in for the reference's top-level Cython module `sauvola` (cython/sauvola.pyx), imported
by bare name at internetarchivepdf/mrc.py:37.  Same function name, arguments and in-place
output convention; the arithmetic runs in libb200mrc.so on the GPU (no CPU fallback)."""
import numpy as np

import archive_pdf_tools_b200 as _pkg
from archive_pdf_tools_b200 import engine as _E, _lib as _L


def _check_u8(a, ndim, name):
    a = np.asarray(a) if not isinstance(a, np.ndarray) else a
    if a.ndim != ndim:
        raise ValueError('Buffer has wrong number of dimensions (expected %d, got %d)' % (ndim, a.ndim))
    if a.dtype.itemsize != 1 or a.dtype.kind not in 'ub':
        raise ValueError("Buffer dtype mismatch, expected 'UINT8DTYPE_t' but got '%s'" % a.dtype)
    return a


def binarise_sauvola(in_arr, out_arr, width, height, window_width, window_height, k, R):
    """cython/sauvola.pyx:29 -- writes 0 for foreground, 1 for background into out_arr; returns 0."""
    in_arr = _check_u8(in_arr, 1, 'in_arr')
    out_arr = _check_u8(out_arr, 1, 'out_arr')
    eng = _pkg.get_engine()
    if width <= 0 or height <= 0:
        return 0
    src = _E.Plane(1, height, width, 1, eng.device).upload(
        np.ascontiguousarray(in_arr[: width * height]).view(np.uint8).reshape(1, height, width), non_blocking=False)
    dst = _E.Plane(1, height, width, 1, eng.device)
    eng.sauvola(src, dst, window_width, window_height, k, R, _L.SAUVOLA_RAW_INVERTED)
    out_arr.view(np.uint8)[: width * height] = dst.numpy()[0].reshape(-1)
    return 0
