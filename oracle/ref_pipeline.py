"""The reference path itself, as far as it can run on this image.

TEST INFRASTRUCTURE ONLY (checker + CPU baseline).  Never imported by the product path.

* `ref_modules()`  -> the reference's OWN Cython (`sauvola`, `optimiser`), compiled unmodified from
  /root/reference/cython/*.pyx by oracle/build_ref.py into oracle/_ref (travels to the GPU box).
* `ref_decompose()` -> create_mrc_hocr_components (mrc.py:334-471) for hocr_word_data == [] with
  the reference's compiled kernels, real Pillow (`convert('L')`, `thumbnail`) and real scipy
  (`gaussian_filter`).  The ~40 lines of glue are restated here because /root/reference does not
  exist on the GPU box; tests/golden/make_golden.py checks this glue against the *imported*
  reference mrc.py inside the build container.  scikit-image's estimate_sigma is not installed:
  sigma comes from the oracle restatement (parity unpinned) or is injected by the caller.
* `load_reference_mrc()` -> imports /root/reference/internetarchivepdf/mrc.py unmodified with stub
  `fitz` / `skimage` modules (container only; used to make golden fixtures).
"""
import importlib.machinery
import importlib.util
import os
import sys
import types
import warnings

import numpy as np

from . import build_ref
from . import oracle as orc

_mods = None


def ref_modules():
    """(sauvola, optimiser) reference extension modules, or None if oracle/_ref is not built."""
    global _mods
    if _mods is None:
        if not build_ref.have_ref() and not build_ref.build():
            return None
        out = []
        for name in ('sauvola', 'optimiser'):
            path = build_ref.ref_so_path(name)
            loader = importlib.machinery.ExtensionFileLoader(name, path)
            spec = importlib.util.spec_from_loader(name, loader, origin=path)
            mod = importlib.util.module_from_spec(spec)
            loader.exec_module(mod)
            out.append(mod)
        _mods = tuple(out)
    return _mods


def ref_threshold_image(img, dpi, k=0.34, window=None):
    """mrc.py:58-87 on the reference's binarise_sauvola."""
    sauvola, _ = ref_modules()
    window_size = window if window is not None else orc.window_for_dpi(dpi)
    h, w = img.shape
    out_img = np.ndarray(img.shape, dtype=bool)
    out_img = np.reshape(out_img, w * h)
    in_img = np.reshape(np.ascontiguousarray(img), w * h)
    sauvola.binarise_sauvola(in_img, out_img.view(np.uint8), w, h, window_size, window_size, k, 128)
    return np.invert(np.reshape(out_img, (h, w)))


def ref_decompose(image, dpi=None, bg_downsample=None, fg_downsample=None, denoise_mask=None,
                  window=None, sigma_est=None, mask_only=False):
    """image: uint8 ndarray (H,W) or (H,W,3).  Same return dict as oracle.decompose()."""
    from PIL import Image
    from scipy import ndimage
    _, optimiser = ref_modules()
    pil = Image.fromarray(image)
    gray = pil if pil.mode == 'L' else pil.convert('L')                       # mrc.py:358-363
    width_, height_ = pil.size
    mask_arr = np.array(Image.new('1', pil.size))                             # mrc.py:367
    grayimgf = np.array(gray, dtype=np.float32)                               # mrc.py:372
    if sigma_est is None:
        sigma_est = orc.estimate_noise(np.array(gray))                        # mrc.py:305 (unpinned)
    imgf = grayimgf
    if sigma_est > 1.0:                                                       # mrc.py:309-311
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            imgf = ndimage.gaussian_filter(imgf, sigma=sigma_est * 0.1)
    mask_arr |= ref_threshold_image(imgf.astype(np.uint8), dpi, window=window)  # mrc.py:325-329
    if denoise_mask != orc.DENOISE_NONE:
        if denoise_mask == orc.DENOISE_FAST:
            optimiser.fast_mask_denoise(mask_arr.view(np.uint8), width_, height_, 4, 2)   # mrc.py:388
        else:
            raise ValueError('Invalid denoise option:', denoise_mask)
    res = dict(mask=mask_arr, sigma=sigma_est, errors=set())
    if mask_only:
        return res
    image_arr = np.array(pil)
    m8 = mask_arr.view(np.uint8)
    opt = optimiser.optimise_gray2 if pil.mode == 'L' else optimiser.optimise_rgb2
    fg = opt(m8, image_arr, width_, height_, 3)                               # mrc.py:412-415
    if fg_downsample is not None:
        im2 = Image.fromarray(fg)
        w, h = im2.size
        wd, hd = int(w / fg_downsample), int(h / fg_downsample)
        if wd > 0 and hd > 0:
            im2.thumbnail((wd, hd))
            fg = np.array(im2)
        else:
            res['errors'].add('too-small-to-downsample')
    mask_inv = mask_arr ^ np.ones(mask_arr.shape, dtype=bool)                 # mrc.py:439
    bg = opt(mask_inv.view(np.uint8), image_arr, width_, height_, 10)         # mrc.py:446-449
    if bg_downsample is not None:
        im2 = Image.fromarray(bg)
        w, h = im2.size
        wd, hd = int(w / bg_downsample), int(h / bg_downsample)
        if wd > 0 and hd > 0:
            im2.thumbnail((wd, hd))
            bg = np.array(im2)
        else:
            res['errors'].add('too-small-to-downsample')
    res['fg'], res['bg'] = fg, bg
    return res


def load_reference_mrc(reference_dir='/root/reference'):
    """Import the UNMODIFIED reference internetarchivepdf/mrc.py (container only).  Third-party
    modules that are not installed are stubbed: fitz (only fitz.TOOLS.set_icc is touched at
    import, mrc.py:39-41) and skimage, whose estimate_sigma is bound to the oracle restatement."""
    return _load_reference_mrc(reference_dir, dropin=False)


def load_reference_mrc_on_dropin(reference_dir='/root/reference'):
    """The UNMODIFIED reference mrc.py on top of whatever top-level `sauvola` / `optimiser` modules are importable --
    i.e. the GPU drop-ins after archive_pdf_tools_b200.install().  Falls back to the byte-compiled glue in
    oracle/_ref/internetarchivepdf (build_ref.build_glue) where the reference checkout does not exist (GPU box)."""
    return _load_reference_mrc(reference_dir, dropin=True)


def _load_reference_mrc(reference_dir, dropin):
    pkg_dir = os.path.join(reference_dir, 'internetarchivepdf')
    mrc_path = os.path.join(pkg_dir, 'mrc.py')
    if not os.path.exists(mrc_path):
        if not dropin or not build_ref.have_glue():
            return None
        pkg_dir = build_ref.GLUE_DIR
        mrc_path = os.path.join(pkg_dir, 'mrc.pyc')
    if dropin:
        import sauvola, optimiser                            # the drop-in modules (install() put them on sys.path)
    else:
        sauvola, optimiser = ref_modules()
    saved = {k: sys.modules.get(k) for k in ('sauvola', 'optimiser', 'fitz', 'skimage', 'skimage.filters',
                                             'skimage.restoration', 'internetarchivepdf',
                                             'internetarchivepdf.jpeg2000', 'internetarchivepdf.const')}
    sys.modules['sauvola'], sys.modules['optimiser'] = sauvola, optimiser
    fitz = types.ModuleType('fitz')
    fitz.TOOLS = types.SimpleNamespace(set_icc=lambda *_: None)
    sys.modules['fitz'] = fitz
    sk = types.ModuleType('skimage')
    skf = types.ModuleType('skimage.filters')
    skf.threshold_local = skf.threshold_otsu = None
    skr = types.ModuleType('skimage.restoration')
    skr.denoise_tv_bregman = None

    def estimate_sigma(arr):
        a = np.asarray(arr)
        if a.dtype == np.bool_:                              # create_hocr_mask (mrc.py:253-254)
            return orc.estimate_sigma_bool(a)
        if a.dtype != np.float32 or np.any(a != np.floor(a)) or a.min() < 0 or a.max() > 255:
            raise NotImplementedError('stub estimate_sigma: only uint8-valued float32 input')
        return orc.estimate_sigma_full(np.ascontiguousarray(a).astype(np.uint8))

    skr.estimate_sigma = estimate_sigma
    sys.modules.update({'skimage': sk, 'skimage.filters': skf, 'skimage.restoration': skr})
    pkg = types.ModuleType('internetarchivepdf')
    pkg.__path__ = [pkg_dir]
    sys.modules['internetarchivepdf'] = pkg
    try:
        if mrc_path.endswith('.pyc'):
            loader = importlib.machinery.SourcelessFileLoader('internetarchivepdf.mrc', mrc_path)
            spec = importlib.util.spec_from_loader('internetarchivepdf.mrc', loader, origin=mrc_path)
        else:
            spec = importlib.util.spec_from_file_location('internetarchivepdf.mrc', mrc_path)
        mod = importlib.util.module_from_spec(spec)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            spec.loader.exec_module(mod)
        # scipy.ndimage.filters alias (mrc.py:311) still resolves on scipy 1.18 with a warning
        return mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def load_reference_recode(mrc_module, reference_dir='/root/reference'):
    """Import the UNMODIFIED reference internetarchivepdf/recode.py (the checkout, or the byte-compiled glue on the GPU
    box) bound to `mrc_module` as internetarchivepdf.mrc.  Everything else recode.py imports and this image lacks --
    fitz, hocr.parse, pdfhacks, pdfrenderer, scandata -- is a stub: the tests drive insert_images_mrc (recode.py:266-530),
    whose page loop needs only the four hocr.parse names (given by the caller through `hocr_pages`) and page objects."""
    pkg_dir = os.path.join(reference_dir, 'internetarchivepdf')
    path = os.path.join(pkg_dir, 'recode.py')
    if not os.path.exists(path):
        if not build_ref.have_glue():
            return None
        pkg_dir = build_ref.GLUE_DIR
        path = os.path.join(pkg_dir, 'recode.pyc')
    names = ('fitz', 'hocr', 'hocr.parse', 'internetarchivepdf', 'internetarchivepdf.mrc', 'internetarchivepdf.grayconvert',
             'internetarchivepdf.pdfhacks', 'internetarchivepdf.pdfrenderer', 'internetarchivepdf.scandata',
             'internetarchivepdf.jpeg2000', 'internetarchivepdf.const')
    saved = {k: sys.modules.get(k) for k in names}

    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    def unused(*_a, **_k):
        raise NotImplementedError('stub: not on the page loop')

    stub('fitz', TOOLS=types.SimpleNamespace(set_icc=lambda *_: None))
    stub('hocr')
    # a "hocr page" is a dict here: {'dpi': (x, y) or None, 'words': hocr_word_data, 'dim': (w, h)}
    stub('hocr.parse', hocr_page_iterator=lambda pages: iter(pages), hocr_page_to_word_data=lambda p: p['words'],
         hocr_page_get_dimensions=lambda p: p['dim'], hocr_page_get_scan_res=lambda p: p.get('dpi') or (None, None))
    pkg = types.ModuleType('internetarchivepdf')
    pkg.__path__ = [pkg_dir]
    sys.modules['internetarchivepdf'] = pkg
    sys.modules['internetarchivepdf.mrc'] = mrc_module
    stub('internetarchivepdf.grayconvert', special_gray_convert=unused)
    stub('internetarchivepdf.pdfhacks', **{n: unused for n in ('fast_insert_image', 'write_pdfa', 'write_page_labels',
                                                               'write_basic_ua', 'write_metadata', 'write_pdf_toc')})
    stub('internetarchivepdf.pdfrenderer', TessPDFRenderer=unused)
    stub('internetarchivepdf.scandata', **{n: unused for n in ('scandata_xml_get_skip_pages', 'scandata_xml_get_page_numbers',
                                                               'scandata_xml_get_dpi_per_page', 'scandata_xml_get_document_dpi')})
    try:
        if path.endswith('.pyc'):
            loader = importlib.machinery.SourcelessFileLoader('internetarchivepdf.recode', path)
            spec = importlib.util.spec_from_loader('internetarchivepdf.recode', loader, origin=path)
        else:
            spec = importlib.util.spec_from_file_location('internetarchivepdf.recode', path)
        mod = importlib.util.module_from_spec(spec)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            spec.loader.exec_module(mod)                     # jpeg2000 / const come from pkg_dir (real reference code)
        return mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def run_reference_compress_pdf_images(mrc_module, fitz_module, argv, reference_dir='/root/reference'):
    """Execute the UNMODIFIED reference script bin/compress-pdf-images (the checkout, or its byte code on the GPU box) as
    __main__ with `argv`, bound to `mrc_module` as internetarchivepdf.mrc and to the caller's `fitz_module` (a fake
    PyMuPDF: the script opens sys.argv[1] with fitz.open and walks its pages, bin/compress-pdf-images:129-148)."""
    import marshal
    src = os.path.join(reference_dir, 'bin', 'compress-pdf-images')
    if os.path.exists(src):
        code = compile(open(src).read(), src, 'exec')
    elif os.path.exists(build_ref.SCRIPT_PYC):
        code = marshal.loads(open(build_ref.SCRIPT_PYC, 'rb').read()[16:])
    else:
        return False
    pkg_dir = os.path.join(reference_dir, 'internetarchivepdf')
    if not os.path.exists(os.path.join(pkg_dir, 'const.py')):
        pkg_dir = build_ref.GLUE_DIR
    names = ('fitz', 'hocr', 'hocr.parse', 'internetarchivepdf', 'internetarchivepdf.mrc', 'internetarchivepdf.const')
    saved = {k: sys.modules.get(k) for k in names}
    saved_argv = sys.argv
    hp = types.ModuleType('hocr.parse')
    hp.hocr_page_iterator = lambda pages: iter(pages)
    hp.hocr_page_to_word_data = lambda p: p['words']
    pkg = types.ModuleType('internetarchivepdf')
    pkg.__path__ = [pkg_dir]
    sys.modules.update({'fitz': fitz_module, 'hocr': types.ModuleType('hocr'), 'hocr.parse': hp,
                        'internetarchivepdf': pkg, 'internetarchivepdf.mrc': mrc_module})
    sys.modules.pop('internetarchivepdf.const', None)
    sys.argv = list(argv)
    try:
        exec(code, {'__name__': '__main__', '__file__': src})
        return True
    finally:
        sys.argv = saved_argv
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
