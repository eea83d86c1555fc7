#!/usr/bin/env python
"""Build oracle/_ref: the UNMODIFIED reference Cython kernels, compiled where they lie.

TEST INFRASTRUCTURE ONLY (checker / CPU baseline) -- never imported by the product path.

Compiles /root/reference/cython/{sauvola,optimiser}.pyx exactly the way the reference's
setup.py does (setup.py:7 CFLAGS='-Ofast -DNPY_NO_DEPRECATED_API', setup.py:24-27
language_level=3) into oracle/_ref/{sauvola,optimiser}.<abi>.so.  No reference source is
copied into this repository: Cython's generated C goes to a temp dir and is discarded.

oracle/_ref/ is git-ignored but NOT gpurun-ignored, so the built .so files travel to the GPU
box (same image => same ABI), where /root/reference does not exist.
"""
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')
REF = os.environ.get('MRC_REFERENCE_DIR', '/root/reference')


def ref_so_path(name):
    return os.path.join(OUT, name + sysconfig.get_config_var('EXT_SUFFIX'))


def have_ref():
    return all(os.path.exists(ref_so_path(n)) for n in ('sauvola', 'optimiser'))


# The reference's Python glue as byte code: internetarchivepdf/{mrc,const,jpeg2000,recode}.py compiled (py_compile, unmodified)
# into oracle/_ref/internetarchivepdf/*.pyc, so that the GPU box -- where /root/reference does not exist -- can run the
# UNMODIFIED reference create_mrc_hocr_components on top of the drop-in `sauvola` / `optimiser` modules
# (tests/test_gpu_pipeline.py::test_unmodified_reference_through_install) and the page loop of recode.py around it
# (::test_reference_recode_page_loop_on_the_dropin).  Built artefacts only; no source is copied.
GLUE = ('mrc', 'const', 'jpeg2000', 'recode')
GLUE_DIR = os.path.join(OUT, 'internetarchivepdf')


# ... and the reference's bin/compress-pdf-images script (a second caller of create_mrc_hocr_components, :66-70)
SCRIPT_SRC = os.path.join(REF, 'bin', 'compress-pdf-images')
SCRIPT_PYC = os.path.join(OUT, 'bin', 'compress_pdf_images.pyc')


def have_glue():
    return all(os.path.exists(os.path.join(GLUE_DIR, n + '.pyc')) for n in GLUE) and os.path.exists(SCRIPT_PYC)


def build_glue(force=False):
    if have_glue() and not force:
        return True
    srcs = [os.path.join(REF, 'internetarchivepdf', n + '.py') for n in GLUE]
    if not all(os.path.exists(p) for p in srcs):
        return have_glue()
    import py_compile
    os.makedirs(GLUE_DIR, exist_ok=True)
    for n, p in zip(GLUE, srcs):
        py_compile.compile(p, cfile=os.path.join(GLUE_DIR, n + '.pyc'), doraise=True)
    if os.path.exists(SCRIPT_SRC):
        os.makedirs(os.path.dirname(SCRIPT_PYC), exist_ok=True)
        py_compile.compile(SCRIPT_SRC, cfile=SCRIPT_PYC, doraise=True)
    return have_glue()


def build(force=False, verbose=False):
    """Returns True if oracle/_ref is usable after the call."""
    try:
        build_glue(force)
    except Exception:
        pass
    if have_ref() and not force:
        return True
    pyx = [os.path.join(REF, 'cython', n + '.pyx') for n in ('sauvola', 'optimiser')]
    if not all(os.path.exists(p) for p in pyx):
        return have_ref()
    import numpy
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix='mrc_ref_build_')
    try:
        for p in pyx:
            name = os.path.splitext(os.path.basename(p))[0]
            c_file = os.path.join(tmp, name + '.c')
            # cython reads the .pyx in place; only the generated C lands in tmp
            subprocess.check_call([sys.executable, '-m', 'cython', '-3', p, '-o', c_file],
                                  stdout=None if verbose else subprocess.DEVNULL,
                                  stderr=None if verbose else subprocess.DEVNULL)
            inc = [sysconfig.get_paths()['include'], numpy.get_include()]
            cmd = ['gcc', '-shared', '-fPIC', '-Ofast', '-DNPY_NO_DEPRECATED_API', '-w']
            for i in inc:
                cmd += ['-I', i]
            cmd += [c_file, '-o', ref_so_path(name)]
            subprocess.check_call(cmd)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return have_ref()


if __name__ == '__main__':
    ok = build(force='--force' in sys.argv, verbose=True)
    print('oracle/_ref:', 'ok' if ok else 'UNAVAILABLE', OUT)
    sys.exit(0 if ok else 1)
