"""ctypes binding of librnr_b200.so (the C ABI declared in include/rnr_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, an
exception is raised.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librnr_b200.so")

F16, BF16, F32 = 0, 1, 2
EPI_BIAS, EPI_TANH, EPI_STATS, EPI_GSTATS = 1, 2, 4, 8
MAX_VIEWS = 8


class View(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dim", C.c_int32 * 4), ("stride", C.c_int64 * 4)]


class KStep(C.Structure):
    _fields_ = [("view", C.c_int16), ("c0", C.c_int16), ("dx", C.c_int16), ("dy", C.c_int16)]


class ConvProblem(C.Structure):
    _fields_ = [
        ("views", View * MAX_VIEWS), ("n_views", C.c_int32), ("ab_dtype", C.c_int32), ("bk", C.c_int32),
        ("n_ksteps", C.c_int32), ("ksteps", C.POINTER(KStep)),
        ("wmat", C.c_void_p), ("n_rows_w", C.c_int32), ("cout", C.c_int32),
        ("mN", C.c_int32), ("mY", C.c_int32), ("mX", C.c_int32), ("th", C.c_int32), ("tw", C.c_int32),
        ("out", C.c_void_p), ("out_dtype", C.c_int32),
        ("out_sn", C.c_int64), ("out_sy", C.c_int64), ("out_sx", C.c_int64),
        ("out_my", C.c_int32), ("out_mx", C.c_int32), ("out_py", C.c_int32), ("out_px", C.c_int32),
        ("epi", C.c_int32), ("bias", C.c_void_p), ("stats", C.c_void_p), ("ldstats", C.c_int32),
    ]


class WTap(C.Structure):
    _fields_ = [("view", C.c_int16), ("dx", C.c_int16), ("dy", C.c_int16), ("gview", C.c_int16),
                ("c0", C.c_int32), ("ci0", C.c_int32), ("nci", C.c_int32), ("off", C.c_int64)]


class WgradProblem(C.Structure):
    _fields_ = [
        ("aviews", View * MAX_VIEWS), ("gviews", View * 4), ("n_aviews", C.c_int32), ("n_gviews", C.c_int32),
        ("a_dtype", C.c_int32), ("g_dtype", C.c_int32), ("n_taps", C.c_int32), ("taps", C.POINTER(WTap)),
        ("cout", C.c_int32), ("mN", C.c_int32), ("mY", C.c_int32), ("mX", C.c_int32),
        ("dw", C.c_void_p), ("s_co", C.c_int64), ("s_ci", C.c_int64),
    ]


class GSrc(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dtype", C.c_int32), ("fold", C.c_int32), ("ld", C.c_int32), ("c0", C.c_int32)]


class GStatSeg(C.Structure):
    _fields_ = [("raw", C.c_void_p), ("raw_dtype", C.c_int32), ("C", C.c_int32), ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("mean", C.c_void_p), ("drop", C.c_void_p), ("slope", C.c_float), ("c_lo", C.c_int32), ("c_hi", C.c_int32),
                ("totals", C.c_void_p)]


class WPrepJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("dst_dtype", C.c_int32), ("nr", C.c_int32), ("nr_pad", C.c_int32),
                ("nc", C.c_int32), ("cpad", C.c_int32), ("ntaps", C.c_int32), ("s_r", C.c_int64), ("s_c", C.c_int64),
                ("tapoff", C.c_int32 * 16), ("chunked", C.c_int32), ("ld", C.c_int64)]


class WUnpackJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("cout", C.c_int32), ("cin", C.c_int32), ("ntaps", C.c_int32),
                ("s_co", C.c_int64), ("s_ci", C.c_int64)]


class WMat(C.Structure):
    _fields_ = [("base", C.c_void_p), ("ld", C.c_int64), ("dtype", C.c_int32), ("ntaps", C.c_int32), ("sub_rows", C.c_int32),
                ("r0", C.c_int32), ("r1", C.c_int32), ("inv", C.c_int8 * 16)]


class AdamWJob(C.Structure):
    _fields_ = [("scratch", C.c_void_p), ("p", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("gdst", C.c_void_p),
                ("cout", C.c_int32), ("cin", C.c_int32), ("ntaps", C.c_int32), ("s_co", C.c_int64), ("s_ci", C.c_int64),
                ("fwd", WMat), ("dgrad", WMat)]


class AdamJob(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_int64)]


class RnrError(RuntimeError):
    pass


_lib = None


def lib():
    """Load (once) and return the ctypes handle.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RnrError(
            "librnr_b200.so not found at %s -- the CUDA library must be built "
            "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
    L.rnr_version.restype = C.c_char_p
    L.rnr_last_error.restype = C.c_char_p
    L.rnr_device_sm_count.argtypes = [i32]
    L.rnr_launch_count.restype = C.c_ulonglong
    L.rnr_launch_count.argtypes = []
    sigs = {
        "rnr_conv_plan_create": [C.POINTER(ConvProblem), i32, C.POINTER(vp)],
        "rnr_conv_plan_create_multi": [C.POINTER(ConvProblem), i32, i32, C.POINTER(vp)],
        "rnr_conv_run": [vp, vp],
        "rnr_conv_plan_tiles_m": [vp],
        "rnr_conv_plan_stat_rows": [vp],
        "rnr_conv_plan_set_bn": [vp, vp, vp, f64, f32, f32, vp, vp, vp, vp, vp, vp, vp, vp, i32],
        "rnr_conv_plan_set_gstats": [vp, C.POINTER(GStatSeg), i32, i32, i32, i32],
        "rnr_debug_set_trace": [vp],
        "rnr_wgrad_plan_create": [C.POINTER(WgradProblem), i32, C.POINTER(vp)],
        "rnr_wgrad_run": [vp, vp],
        "rnr_weight_prep": [vp, vp, i32, i32, i32, i32, i32, i32, i64, i64, vp, i32, vp],
        "rnr_wprep_plan_create": [C.POINTER(WPrepJob), i32, C.POINTER(vp)],
        "rnr_wprep_run": [vp, vp],
        "rnr_wgrad_unpack_plan_create": [C.POINTER(WUnpackJob), i32, C.POINTER(vp)],
        "rnr_wgrad_unpack_run": [vp, vp],
        "rnr_bn_finalize": [vp, i32, i32, i32, f64, vp, vp, f32, vp, vp, vp, vp, vp, vp, f32, vp],
        "rnr_bn_act_fwd": [vp, i32, vp, vp, vp, f32, vp, vp, i32, i32, i32, i32, vp],
        "rnr_conv_plan_set_stat_totals": [vp, vp],
        "rnr_bn_act_fwd_tot": [vp, i32, vp, vp, f64, vp, vp, f32, vp, vp, vp, vp, vp, vp, f32, vp, f32, vp, vp, i32, i32, i32, i32, vp],
        "rnr_bn_bwd_reduce": [C.POINTER(GSrc), i32, vp, vp, vp, vp, vp, vp, f32, vp, vp, C.POINTER(i32), i32, i32, i32, i32, vp],
        "rnr_bn_bwd_reduce_fin": [C.POINTER(GSrc), i32, vp, i32, vp, vp, vp, vp, vp, f32, vp, vp, vp, f64, vp, vp, vp, vp, i32, i32, i32, i32, vp],
        "rnr_bn_bwd_apply_src": [C.POINTER(GSrc), i32, vp, i32, vp, vp, vp, vp, vp, vp, f32, vp, vp, vp, f64, vp, vp, i32, i32, i32, i32, vp],
        "rnr_bn_bwd_finalize": [vp, i32, i32, f64, vp, vp, vp, vp, vp, vp, vp, vp, vp],
        "rnr_bn_bwd_apply": [vp, vp, i32, vp, i32, i32, i32, i32, vp],
        "rnr_pack_nchw_to_act": [vp, vp, vp, i32, i32, i32, i32, i32, vp],
        "rnr_unpack_nhwc_to_nchw": [vp, vp, i32, i32, i32, i32, i32, vp],
        "rnr_tanh_bwd_pack": [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
        "rnr_fold_to_nchw": [vp, i32, vp, i32, i32, i32, i32, i32, i32, vp],
        "rnr_fold_to_nchw_add": [vp, i32, vp, i32, i32, i32, i32, i32, i32, vp, i32, vp],
    }
    for name, argtypes in sigs.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = i32
    L.rnr_conv_plan_destroy.argtypes = [vp]
    L.rnr_conv_plan_destroy.restype = None
    L.rnr_wgrad_plan_destroy.argtypes = [vp]
    L.rnr_wgrad_plan_destroy.restype = None
    L.rnr_wprep_plan_destroy.argtypes = [vp]
    L.rnr_wprep_plan_destroy.restype = None
    L.rnr_wgrad_unpack_plan_destroy.argtypes = [vp]
    L.rnr_wgrad_unpack_plan_destroy.restype = None
    L.rnr_adam_plan_destroy.argtypes = [vp]
    L.rnr_adam_plan_destroy.restype = None
    _register_optional(L)
    _lib = L
    return L


_OPTIONAL_SIGS = {}


def register_sigs(sigs):
    """Other host modules (texture, rays, raster ...) declare their entry points here."""
    _OPTIONAL_SIGS.update(sigs)
    if _lib is not None:
        _register_optional(_lib)


def _register_optional(L):
    for name, argtypes in _OPTIONAL_SIGS.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int


def check(rc, what=""):
    if rc != 0:
        msg = lib().rnr_last_error().decode()
        raise RnrError("%s failed (cudaError %d): %s" % (what or "librnr_b200 call", rc, msg))


def exported_symbols():
    """Names declared in include/rnr_b200.h (parsed), used by the ABI test."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), "include", "rnr_b200.h")
    txt = open(hdr).read()
    return sorted(set(re.findall(r"\b(rnr_[a-z0-9_]+)\s*\(", txt)))
