import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def orc():
    """The CPU oracle (test infrastructure; see oracle/mrc_oracle.c)."""
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope='session')
def refmods():
    """The reference's own compiled Cython (oracle/_ref), or skip when it was never built."""
    from oracle import ref_pipeline
    mods = ref_pipeline.ref_modules()
    if mods is None:
        pytest.skip('oracle/_ref not built (needs /root/reference once, see oracle/build_ref.py)')
    return mods


@pytest.fixture(scope='session')
def synth():
    import archive_pdf_tools_b200.synth as s
    return s


@pytest.fixture(scope='session')
def eng():
    import torch
    assert torch.cuda.is_available(), 'gpu tests need a CUDA device (no CPU fallback exists)'
    import archive_pdf_tools_b200 as pkg
    return pkg.get_engine()


GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def golden_cases():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith('.npz') and not f.startswith('hocr_'))


def hocr_cases():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith('.npz') and f.startswith('hocr_'))


def load_hocr_golden(name, synth):
    """Outputs of the imported reference create_mrc_hocr_components / create_hocr_mask with text-line boxes
    (tests/golden/make_golden.py HOCR_CASES); page and hocr_word_data are regenerated from the seeds."""
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    idx, H, W, dpi, rgb, ds, bgd = (int(v) for v in z['params'])
    dpi = None if dpi < 0 else dpi
    ds = None if ds < 0 else ds
    page = synth.make_page(idx, H, W, dpi=dpi or 100, rgb=bool(rgb), sigma_n=float(z['sigma_n']),
                           invert_lines=tuple(int(v) for v in z['invert_lines']),
                           noisy_dark_lines=tuple(int(v) for v in z['noisy_dark_lines']))
    hocr = synth.page_hocr(H, W, dpi=dpi or 100, scale=float(ds or 1))
    return dict(page=page, hocr=hocr, dpi=dpi, downsample=ds, bg_downsample=None if bgd < 0 else bgd, denoise=str(z['denoise']),
                fg=z['fg'], bg=z['bg'], mask=np.unpackbits(z['mask'])[: H * W].reshape(H, W).astype(bool),
                hocr_mask=np.unpackbits(z['hocr_mask'])[: H * W].reshape(H, W).astype(bool),
                timing_keys=[str(k) for k in z['timing_keys']])


def load_golden(name, synth):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    idx, H, W, dpi, rgb, ht, bgd, fgd = (int(v) for v in z['params'])
    dpi = None if dpi < 0 else dpi
    page = synth.make_page(idx, H, W, dpi=dpi or 200, rgb=bool(rgb), sigma_n=float(z['sigma_n']), halftone=bool(ht))
    g = dict(page=page, dpi=dpi, bg_downsample=None if bgd < 0 else bgd, fg_downsample=None if fgd < 0 else fgd,
             denoise=str(z['denoise']), sigma=float(z['sigma']), fg=z['fg'], bg=z['bg'],
             mask=np.unpackbits(z['mask'])[: H * W].reshape(H, W).astype(bool),
             t33=np.unpackbits(z['t33'])[: H * W].reshape(H, W).astype(bool),
             t01=np.unpackbits(z['t01'])[: H * W].reshape(H, W).astype(bool),
             timing_keys=[str(k) for k in z['timing_keys']])
    return g


@pytest.fixture
def tuning():
    """set(name, value) changes a b200mrc tuning knob (b200mrc_set_tuning) for the test and restores it afterwards."""
    from archive_pdf_tools_b200 import _lib
    saved = {}

    def set_(name, value):
        saved.setdefault(name, _lib.get_tuning(name))
        _lib.set_tuning(name, value)
    yield set_
    for k, v in saved.items():
        _lib.set_tuning(k, v)


def drive_reference_page_loop(recode, pages, hocr_words, tmp_path, dpi, force_1bit=False, **mrc_kw):
    """Run the UNMODIFIED reference page loop recode.insert_images_mrc (recode.py:266-530) over `pages` (numpy arrays,
    written as PNG files) with PDF objects and encoders mocked: pages are plain recorders of insert_image(), and
    encode_mrc_images / encode_mrc_mask (mrc.py:474-580; external jbig2 / JPEG2000 tools) hand the arrays they are given
    to the test instead of encoding them.  Returns (captured arrays per page, page recorders, timing keys seen)."""
    from PIL import Image
    captured, files = [], []
    for i, pg in enumerate(pages):
        f = str(tmp_path / ('page%03d.png' % i))
        Image.fromarray(pg).save(f)
        files.append(f)

    class Page:
        rect = (0, 0, 612, 792)

        def __init__(self):
            self.inserted = []

        def insert_image(self, rect, **kw):
            self.inserted.append(kw)

    def dummy(tag):
        f = str(tmp_path / ('%s_%d.bin' % (tag, len(captured))))
        open(f, 'wb').write(b'x')
        return f

    def fake_encode_mrc_images(mrc_gen, **kw):
        mask, fg, bg = list(mrc_gen)                          # the reference's own consumer pulls the three yields in order
        captured.append(dict(mask=mask, fg=fg, bg=bg))
        return dummy('mask'), dummy('bg'), (bg.shape[1], bg.shape[0]), dummy('fg'), (fg.shape[1], fg.shape[0])

    def fake_encode_mrc_mask(np_mask, **kw):
        captured.append(dict(mask_inverted=np_mask))
        return dummy('jb2'), dummy('png')

    saved = recode.encode_mrc_images, recode.encode_mrc_mask
    recode.encode_mrc_images, recode.encode_mrc_mask = fake_encode_mrc_images, fake_encode_mrc_mask
    try:
        to_pdf = [Page() for _ in pages]
        hocr_pages = [dict(dpi=None, words=w, dim=(p.shape[1], p.shape[0])) for p, w in zip(pages, hocr_words)]
        errors = set()
        recode.insert_images_mrc(to_pdf, hocr_pages, image_files=files, dpi=dpi, hq_pages=[False] * len(pages),
                                 tmp_dir=str(tmp_path), force_1bit_output=force_1bit, errors=errors, **mrc_kw)
    finally:
        recode.encode_mrc_images, recode.encode_mrc_mask = saved
    return captured, to_pdf, errors


def run_reference_compress_script(mrc, pages, hocr_words=None):
    """Run the UNMODIFIED reference bin/compress-pdf-images over a fake PyMuPDF document whose pages each hold one image
    (`pages`: numpy arrays), with mrc.encode_mrc_images replaced by a recorder.  Returns (captured arrays, fake document)
    or None when neither the reference checkout nor its byte code is available."""
    import io, types
    from PIL import Image
    from oracle import ref_pipeline
    captured = []

    class Page:
        def __init__(self, doc, xref):
            self.doc, self.xref, self.inserted, self.cleaned = doc, xref, [], 0

        def clean_contents(self):
            self.cleaned += 1

        def get_images(self, full=False):
            return [(self.xref, 0, 0, 0, 8, 'DeviceRGB', '', 'Im%d' % self.xref, 'FlateDecode')]

        def get_image_bbox(self, img_data):
            return (0, 0, 612, 792)

        def get_contents(self):
            return [1000 + self.xref]

        def read_contents(self):
            return b'q\n/Im%d Do\nQ' % self.xref

        def insert_image(self, bbox, **kw):
            self.inserted.append(kw)

    class Doc:
        def __init__(self, path):
            self.pages = [Page(self, i) for i in range(len(pages))]
            self.streams, self.saved = {}, None

        def __iter__(self):
            return iter(self.pages)

        def extract_image(self, xref):
            b = io.BytesIO()
            Image.fromarray(pages[xref]).save(b, format='PNG')
            return {'image': b.getvalue(), 'xres': 300, 'yres': 300}

        def update_stream(self, xref, data):
            self.streams[xref] = data

        def save(self, path, **kw):
            self.saved = path

    docs = []
    fitz = types.ModuleType('fitz')
    fitz.TOOLS = types.SimpleNamespace(set_icc=lambda *_: None)
    fitz.open = lambda path: docs.append(Doc(path)) or docs[-1]

    def fake_encode_mrc_images(mrc_gen, **kw):
        mask, fg, bg = list(mrc_gen)
        captured.append(dict(mask=mask, fg=fg, bg=bg))
        import tempfile
        fs = []
        for _ in range(3):
            fd, f = tempfile.mkstemp(suffix='.bin')
            os.write(fd, b'x'); os.close(fd)
            fs.append(f)
        return fs[0], fs[1], (bg.shape[1], bg.shape[0]), fs[2], (fg.shape[1], fg.shape[0])

    saved = mrc.encode_mrc_images
    mrc.encode_mrc_images = fake_encode_mrc_images
    try:
        argv = ['compress-pdf-images', 'in.pdf', 'out.pdf'] if hocr_words is None else \
               ['compress-pdf-images', 'in.pdf', [dict(words=w) for w in hocr_words], 'out.pdf']
        if not ref_pipeline.run_reference_compress_pdf_images(mrc, fitz, argv):
            return None
    finally:
        mrc.encode_mrc_images = saved
    return captured, docs[0]
