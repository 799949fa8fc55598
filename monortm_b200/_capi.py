"""ctypes declarations for libmonortm_b200.so (include/monortm_b200.h).

The library is the product: if it is missing or a symbol is absent this module
raises -- there is no Python/CPU fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MRTM_LIB", os.path.join(_HERE, "lib", "libmonortm_b200.so"))

MXMOL = 39
NSCOR = 42 * 9

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)
c_int32_p = C.POINTER(C.c_int32)
c_uint64_p = C.POINTER(C.c_uint64)


class MrtmOpts(C.Structure):
    _fields_ = [
        ("use_global_range", C.c_int32),
        ("line_mode", C.c_int32),
        ("v1_global", C.c_double),
        ("v2_global", C.c_double),
        ("iw0", C.c_int64),
        ("sel_count", c_int64_p),
        ("sel_hash", c_uint64_p),
        ("stream", C.c_void_p),
        ("xamnt", c_double_p),
        ("ld_xamnt", C.c_int64),
    ]


class MrtmXsRegion(C.Structure):
    _fields_ = [("ixmol", C.c_int32), ("ntemp", C.c_int32), ("npts", C.c_int64), ("v1fx", C.c_double), ("v2fx", C.c_double),
                ("v1x", C.c_double), ("v2x", C.c_double), ("xdoplr", C.c_double), ("tx", C.c_double * 6), ("pdx", C.c_double * 6),
                ("xsdat", C.c_void_p * 6)]


class MrtmStats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_int64),
        ("lines_staged", C.c_int64),
        ("last_lines_kernel_ms", C.c_double),
        ("last_rt_kernel_ms", C.c_double),
        ("last_derive_kernel_ms", C.c_double),
        ("nominal_evals", C.c_double),
        ("inwindow_evals", C.c_double),
        ("far_expansions", C.c_double),
        ("direct_evals", C.c_double),
        ("last_prep_ms", C.c_double),
    ]


class MrtmControl(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("ihirac", "icntnm", "iemit", "iplot", "iatm", "iod", "ixsect", "ispd", "ibrd")] + [
        ("cntnm", C.c_double * 7),
        ("v1", C.c_double), ("v2", C.c_double), ("dvset", C.c_double),
        ("nwn", C.c_int64), ("wn", c_double_p),
        ("tmpbnd", C.c_double), ("bndemi", C.c_double * 3), ("bndrfl", C.c_double * 3),
        ("nmol_scal", C.c_int64), ("hmol_scal", C.c_char * 64), ("xmol_scal", C.c_double * 64),
    ]


# every symbol include/monortm_b200.h declares: (restype, argtypes)
_D, _I, _P = C.c_double, C.c_int64, C.c_void_p
SIGNATURES = {
    "mrtm_init": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "mrtm_init_multi": (C.c_int, [C.c_uint64, C.POINTER(_P)]),
    "mrtm_num_devices": (C.c_int, [_P]),
    "mrtm_sync": (C.c_int, [_P]),
    "mrtm_free": (C.c_int, [_P]),
    "mrtm_strerror": (C.c_char_p, [C.c_int]),
    "mrtm_last_error": (C.c_char_p, [_P]),
    "mrtm_version": (C.c_char_p, []),
    "mrtm_stage_lines": (C.c_int, [_P, _P, _I] + [_P] * 15),
    "mrtm_num_lines": (C.c_int64, [_P]),
    "mrtm_modm": (C.c_int, [_P, _I, _P, _D, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P,
                            _D, _D, _D, _P, _I, _I, _P, _P]),
    "mrtm_stage_xsec": (C.c_int, [_P, _I, _P]),
    "mrtm_xsec": (C.c_int, [_P, _I, _P, _I, _P, _P, _I, _P, _P]),
    "mrtm_host_xsread": (C.c_int, [C.c_char_p, _I, C.c_char_p, _D, _D, c_int64_p, C.POINTER(C.POINTER(MrtmXsRegion))]),
    "mrtm_host_xs_free": (None, [C.POINTER(MrtmXsRegion), _I]),
    "mrtm_calctmr": (C.c_int, [_P, _I, _I, _P, _P, _P, _P, _P]),
    "mrtm_rtm": (C.c_int, [_P, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I]),
    "mrtm_profiles": (C.c_int, [_P, _I, _I, _P, _D, _I, _P, _P, _P, _P, _I, _P, _P, _P, _D, _D, _D,
                                _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mrtm_profiles_dev": (C.c_int, [_P, _I, _I, _P, _D, _I, _P, _P, _P, _P, _I, _P, _P, _P, _D, _D, _D,
                                    _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mrtm_get_stats": (C.c_int, [_P, C.POINTER(MrtmStats)]),
    "mrtm_reset_stats": (C.c_int, [_P]),
    "mrtm_fp64_peak": (C.c_int, [_P, c_double_p]),
    "mrtm_host_get_lnfl": (C.c_int, [C.c_char_p, _D, _D, _I] + [_P] * 16),
    "mrtm_host_tips_2003": (C.c_int, [_I, _D, _P]),
    "mrtm_host_last_error": (C.c_char_p, []),
    "mrtm_host_read_control": (C.c_int, [C.c_char_p, _I, C.POINTER(MrtmControl)]),
    "mrtm_host_free_control": (None, [C.POINTER(MrtmControl)]),
    "mrtm_host_count_profiles": (C.c_int, [C.c_char_p, _I, c_int64_p]),
    "mrtm_host_read_profile": (C.c_int, [C.c_char_p, _I, _I, c_int64_p, c_int64_p, c_int64_p, c_int64_p,
                                         c_double_p, c_double_p, c_double_p, c_double_p] + [_P] * 8),
    "mrtm_host_emiss_reflec": (C.c_int, [C.POINTER(MrtmControl), C.c_char_p, _I, _P, _P, _P]),
    "mrtm_host_storeout": (C.c_int, [C.c_char_p, C.c_int, _I, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P,
                                     _D, _D, _D, _P, _P, _I, _I, _D, _I]),
    "mrtm_host_run_monortm": (C.c_int, [C.c_char_p, C.c_int, _I, C.c_int]),
}

_lib = None


def load_library():
    """Load libmonortm_b200.so and bind every declared symbol; raise if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libmonortm_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
