"""CPU suite, part 2: the C-ABI library loads and exports every symbol include/b200mrc.h declares
(no compute calls without a GPU), the ctypes mirror matches, and host logic behaves."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, 'include', 'b200mrc.h')).read()
    return sorted(set(re.findall(r'B200MRC_API[^;(]*?\b(b200mrc_\w+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from archive_pdf_tools_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), 'libb200mrc.so must be built (python -c "import __graft_entry__ as g; g.build()")'
    raw = C.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(raw, s), 'missing export: ' + s
    assert sorted(_lib.PROTOTYPES) == syms, 'ctypes prototypes and header disagree'
    assert _lib.lib().b200mrc_version() == 100
    assert _lib.lib().b200mrc_error_string(-3) == b'b200mrc: workspace too small'


def test_argument_validation_without_gpu():
    from archive_pdf_tools_b200 import _lib
    L = _lib.lib()
    assert L.b200mrc_sauvola(None, 0, 0, None, 0, 0, 10, 10, 1, 33, 33, 0.34, 128.0, 0, None) == _lib.ERR_INVALID
    buf = (C.c_uint8 * 64)()
    p = C.cast(buf, C.c_void_p)
    assert L.b200mrc_sauvola(p, 8, 64, p, 8, 64, 8, 8, 1, 300, 33, 0.34, 128.0, 0, None) == _lib.ERR_UNSUPPORTED
    assert L.b200mrc_sauvola(p, 7, 64, p, 8, 64, 7, 8, 1, 33, 33, 0.34, 128.0, 0, None) == _lib.ERR_ALIGNMENT
    assert L.b200mrc_denoise(p, 8, 64, 8, 8, 1, 3, 65, None, 0, None) == _lib.ERR_UNSUPPORTED    # n_size beyond the general form's range
    assert L.b200mrc_denoise(p, 8, 64, 8, 8, 1, 3, 4, None, 0, None) == 0                        # 8x8 page, n_size 4: no interior pixel, nothing to do
    assert L.b200mrc_optimise(p, 8, 64, p, 8, 64, 1, p, 8, 64, 17, None, 0, 0, 10, 8, 8, 1, None, 0, None) == _lib.ERR_UNSUPPORTED
    assert L.b200mrc_noise_workspace_bytes(2550, 3300, 1) >= 4 * 826 * 639
    assert L.b200mrc_denoise_workspace_bytes(2550, 3300, 64) > 0
    assert L.b200mrc_optimise_workspace_bytes(2550, 3300, 1) > 0


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    import archive_pdf_tools_b200 as pkg
    with pytest.raises(pkg.B200MrcError):
        pkg.threshold_image(np.zeros((8, 8), np.uint8), 400)


def test_window_rule():
    from archive_pdf_tools_b200 import window_for_dpi
    assert [window_for_dpi(d) for d in (None, 132, 200, 300, 400, 600)] == [51, 33, 51, 75, 101, 151]   # mrc.py:70-75


def test_product_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, 'archive-pdf-tools_b200')
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, re.M), f
                assert 'mrc_oracle' not in src.replace('oracle/mrc_oracle.c', ''), f


def test_synth_pages_are_deterministic(synth):
    a = synth.make_page(3, 120, 90, dpi=100)
    b = synth.make_page(3, 120, 90, dpi=100)
    assert a.shape == (120, 90, 3) and np.array_equal(a, b)
    g = synth.make_page(3, 120, 90, dpi=100, rgb=False, halftone=True)
    assert g.shape == (120, 90)


def test_host_unpack_mask_is_unpackbits():
    """b200mrc_host_unpack_mask runs on HOST memory (no CUDA call): np.unpackbits of mode-'1' rows, ragged widths,
    pitched destination left untouched outside the page."""
    import numpy as np
    from archive_pdf_tools_b200 import _lib as L
    rng = np.random.default_rng(7)
    for n, h, w in [(2, 37, 61), (1, 5, 8), (3, 20, 7), (1, 3, 1), (2, 64, 2550)]:
        m = rng.random((n, h, w)) < 0.3
        p = np.packbits(m, axis=2)
        out = np.full((n, h, w + 3), 7, np.uint8)
        rc = L.lib().b200mrc_host_unpack_mask(p.ctypes.data, p.shape[2], p.shape[1] * p.shape[2], out.ctypes.data, w + 3, h * (w + 3), w, h, n)
        assert rc == 0
        assert np.array_equal(out[:, :, :w], m.view(np.uint8)) and (out[:, :, w:] == 7).all()
    assert L.lib().b200mrc_host_unpack_mask(None, 1, 1, None, 1, 1, 1, 1, 1) != 0
