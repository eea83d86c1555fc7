"""ctypes binding of libb200mrc.so (include/b200mrc.h).

There is NO fallback: if the CUDA library is missing or no CUDA device is present, every compute
call raises.  The oracle under oracle/ is test infrastructure and is never imported from here.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libb200mrc.so')

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_ALIGNMENT = 0, -1, -2, -3, -4
SAUVOLA_OR_INTO, SAUVOLA_RAW_INVERTED, SAUVOLA_INVERT_INPUT = 1, 2, 4
DECOMPOSE_DENOISE_FAST, DECOMPOSE_MASK_ONLY, DECOMPOSE_NO_NOISE_EST, DECOMPOSE_OR_INTO_MASK = 1, 2, 4, 8
MAX_WINDOW, MAX_OPT_N = 255, 16
COPY_H2D, COPY_D2H, COPY_D2D = 1, 2, 3

u8p = C.POINTER(C.c_uint8)
vp = C.c_void_p
i64 = C.c_int64


class Rect(C.Structure):
    _fields_ = [('ptr', vp), ('pitch', i64), ('width', C.c_int32), ('height', C.c_int32), ('keys', vp)]


class CopyRect(C.Structure):
    _fields_ = [('src', vp), ('src_pitch', i64), ('dst', vp), ('dst_pitch', i64), ('width', C.c_int32), ('height', C.c_int32)]


class SauvolaItem(C.Structure):
    _fields_ = [('in_', vp), ('in_pitch', i64), ('out', vp), ('out_pitch', i64), ('width', C.c_int32), ('height', C.c_int32),
                ('flags', C.c_int32), ('reserved', C.c_int32)]


class DecomposeArgs(C.Structure):
    _fields_ = [
        ('img', vp), ('img_pitch', i64), ('img_page_stride', i64), ('channels', C.c_int),
        ('width', C.c_int), ('height', C.c_int), ('n_pages', C.c_int),
        ('window', C.c_int),
        ('k', C.c_double), ('R', C.c_double),
        ('flags', C.c_int),
        ('sigma_in', vp), ('sigma_out', vp),
        ('mask', vp), ('mask_pitch', i64), ('mask_page_stride', i64),
        ('fg', vp), ('fg_pitch', i64), ('fg_page_stride', i64), ('fg_plan', vp),
        ('bg', vp), ('bg_pitch', i64), ('bg_page_stride', i64), ('bg_plan', vp),
        ('workspace', vp), ('workspace_bytes', C.c_size_t),
    ]


# name -> (restype, argtypes); every symbol include/b200mrc.h declares
PROTOTYPES = {
    'b200mrc_version': (C.c_int, []),
    'b200mrc_error_string': (C.c_char_p, [C.c_int]),
    'b200mrc_launch_count': (C.c_uint64, []),
    'b200mrc_set_tuning': (C.c_int, [C.c_char_p, C.c_int]),
    'b200mrc_get_tuning': (C.c_int, [C.c_char_p, C.POINTER(C.c_int)]),
    'b200mrc_profile_enable': (C.c_int, [C.c_int]),
    'b200mrc_profile_report': (C.c_int, [C.c_char_p, C.c_size_t]),
    'b200mrc_copy2d': (C.c_int, [vp, i64, vp, i64, i64, i64, C.c_int, vp]),
    'b200mrc_rgb2gray': (C.c_int, [vp, i64, i64, vp, i64, i64, C.c_int, C.c_int, C.c_int, vp]),
    'b200mrc_sauvola': (C.c_int, [vp, i64, i64, vp, i64, i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_double, C.c_double, C.c_int, vp]),
    'b200mrc_threshold_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    'b200mrc_threshold_mask': (C.c_int, [vp, i64, i64, C.c_int, vp, i64, i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_double, C.c_double, vp, C.c_int, vp, C.c_size_t, vp]),
    'b200mrc_noise_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    'b200mrc_estimate_noise': (C.c_int, [vp, i64, i64, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_size_t, vp]),
    'b200mrc_gray_blur': (C.c_int, [vp, i64, i64, C.c_int, vp, i64, i64, C.c_int, C.c_int, C.c_int, vp, vp]),
    'b200mrc_denoise_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    'b200mrc_denoise': (C.c_int, [vp, i64, i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_size_t, vp]),
    'b200mrc_optimise_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    'b200mrc_optimise': (C.c_int, [vp, i64, i64, vp, i64, i64, C.c_int, vp, i64, i64, C.c_int, vp, i64, i64, C.c_int,
                                   C.c_int, C.c_int, C.c_int, vp, C.c_size_t, vp]),
    'b200mrc_thumbnail_plan_create': (vp, [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                           C.POINTER(C.c_int)]),
    'b200mrc_resample_plan_destroy': (None, [vp]),
    'b200mrc_resample_plan_out_size': (None, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'b200mrc_resample_workspace_bytes': (C.c_size_t, [vp, C.c_int]),
    'b200mrc_resample': (C.c_int, [vp, vp, i64, i64, vp, i64, i64, C.c_int, vp, C.c_size_t, vp]),
    'b200mrc_rects_copy': (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp]),
    'b200mrc_sauvola_items': (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, vp]),
    'b200mrc_rects_count_nonzero': (C.c_int, [vp, C.c_int, vp, vp]),
    'b200mrc_rects_sigma_bool': (C.c_int, [vp, C.c_int, vp, vp]),
    'b200mrc_pack_mask': (C.c_int, [vp, i64, i64, vp, i64, i64, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    'b200mrc_host_unpack_mask': (C.c_int, [vp, i64, i64, vp, i64, i64, C.c_int, C.c_int, C.c_int]),
    'b200mrc_channel_stats': (C.c_int, [vp, i64, i64, C.c_int, C.c_int, C.c_int, vp, vp]),
    'b200mrc_special_gray': (C.c_int, [vp, i64, i64, vp, i64, i64, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
    'b200mrc_decompose_workspace_bytes': (C.c_size_t, [C.POINTER(DecomposeArgs)]),
    'b200mrc_decompose': (C.c_int, [C.POINTER(DecomposeArgs), vp]),
}

_lib = None


class B200MrcError(RuntimeError):
    pass


def lib():
    """The loaded library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200MrcError('libb200mrc.so not built (%s); run `python -c "import __graft_entry__ as g; g.build()"` '
                               'or `make -C archive-pdf-tools_b200/csrc`. There is no CPU fallback.' % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)          # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status, what=''):
    if status != 0:
        msg = lib().b200mrc_error_string(status)
        raise B200MrcError('%s failed (%d): %s' % (what or 'b200mrc call', status, msg.decode() if msg else '?'))


TUNING_WORDS = {'single': 1, 'trio': 2, 'tma': 1, 'async': 2, 'legacy': 1, 'generic': 1, 'fused': 2, 'split': 0, 'auto': 0}


def set_tuning(name, value):
    """b200mrc_set_tuning: name without the B200MRC_ prefix (e.g. 'IIRW_MODE'); value an int or one of the words the
    environment variable accepts ('single', 'trio', 'tma', 'async', 'legacy', 'generic', 'auto')."""
    if isinstance(value, str):
        value = TUNING_WORDS[value] if value in TUNING_WORDS else int(value)
    check(lib().b200mrc_set_tuning(name.encode(), int(value)), 'b200mrc_set_tuning(%s)' % name)


def get_tuning(name):
    v = C.c_int(0)
    check(lib().b200mrc_get_tuning(name.encode(), C.byref(v)), 'b200mrc_get_tuning(%s)' % name)
    return v.value


def profile_enable(on=True):
    check(lib().b200mrc_profile_enable(1 if on else 0), 'b200mrc_profile_enable')


def profile_report():
    """{kernel: (launches, total_ms)} for the launches recorded since profile_enable(True)."""
    buf = C.create_string_buffer(1 << 16)
    lib().b200mrc_profile_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms = line.rsplit(',', 2)
        out[name] = (int(n), float(ms))
    return out
