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


def build(force=False, verbose=False):
    """Returns True if oracle/_ref is usable after the call."""
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
