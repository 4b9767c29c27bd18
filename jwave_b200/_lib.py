"""ctypes binding of libjwave_cuda.so (include/jwave_cuda.h).

There is NO CPU fallback: if the shared library is missing or cannot be loaded this module
raises, and every transform built on it fails loudly."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# JWAVE_CUDA_LIB points at another build of the same library (A/B runs of kernel variants)
SO_PATH = os.environ.get("JWAVE_CUDA_LIB") or os.path.join(_HERE, "libjwave_cuda.so")

OK, ERR_NOT_BINARY, ERR_LEVEL, ERR_ARG, ERR_CUDA, ERR_NCCL = range(6)
FORWARD, REVERSE = 0, 1
FWT, WPT = 0, 1
MAX_TAPS = 40

_dp = C.POINTER(C.c_double)
_vp = C.c_void_p
_i64 = C.c_int64
_int = C.c_int

# name -> (restype, argtypes); exactly the symbols include/jwave_cuda.h declares
SIGNATURES = {
    "jwc_version": (_int, []),
    "jwc_create": (_int, [C.POINTER(_vp), _int]),
    "jwc_create_multi": (_int, [C.POINTER(_vp), C.POINTER(_int), _int]),
    "jwc_device_count": (_int, [_vp]),
    "jwc_destroy": (_int, [_vp]),
    "jwc_last_error": (C.c_char_p, [_vp]),
    "jwc_set_stream": (_int, [_vp, _vp]),
    "jwc_reset_stream": (_int, [_vp]),
    "jwc_sync": (_int, [_vp]),
    "jwc_launch_count": (_i64, [_vp]),
    "jwc_profile_enable": (_int, [_vp, _int]),
    "jwc_profile_report": (_int, [_vp, C.c_char_p, C.c_size_t]),
    "jwc_set_wavelet": (_int, [_vp, _int, _dp, _dp, _dp, _dp, C.POINTER(_int)]),
    "jwc_fwt1d": (_int, [_vp, _int, _int, _vp, _vp, _i64, _int, _int]),
    "jwc_wpt1d": (_int, [_vp, _int, _int, _vp, _vp, _i64, _int, _int]),
    "jwc_fwt2d": (_int, [_vp, _int, _int, _vp, _vp, _i64, _int, _int, _int, _int]),
    "jwc_wpt2d": (_int, [_vp, _int, _int, _vp, _vp, _i64, _int, _int, _int, _int]),
    "jwc_fwt3d": (_int, [_vp, _int, _int, _vp, _vp, _int, _int, _int, _int, _int, _int]),
    "jwc_wpt3d": (_int, [_vp, _int, _int, _vp, _vp, _int, _int, _int, _int, _int, _int]),
    "jwc_aed1d": (_int, [_vp, _int, _int, _int, _vp, _vp, _i64, _int]),
    "jwc_aed1d_dev": (_int, [_vp, _int, _int, _int, _vp, _vp, _i64, _int]),
    "jwc_decompose1d": (_int, [_vp, _int, _int, _vp, _vp, _i64, _int]),
    "jwc_decompose1d_dev": (_int, [_vp, _int, _int, _vp, _vp, _i64, _int]),
    "jwc_compress_magnitude": (_int, [_vp, _vp, _vp, _i64, C.c_double, _dp]),
    "jwc_compress_magnitude_dev": (_int, [_vp, _vp, _vp, _i64, C.c_double, _vp]),
    "jwc_forward1d_compress_dev": (_int, [_vp, _int, _int, _vp, _vp, _i64, _int, _int, C.c_double, _vp]),
    "jwc_fwt1d_dev": (_int, [_vp, _int, _int, _vp, _vp, _i64, _int, _int]),
    "jwc_wpt1d_dev": (_int, [_vp, _int, _int, _vp, _vp, _i64, _int, _int]),
    "jwc_fwt2d_dev": (_int, [_vp, _int, _int, _vp, _vp, _i64, _int, _int, _int, _int]),
    "jwc_wpt2d_dev": (_int, [_vp, _int, _int, _vp, _vp, _i64, _int, _int, _int, _int]),
    "jwc_fwt3d_dev": (_int, [_vp, _int, _int, _vp, _vp, _int, _int, _int, _int, _int, _int]),
    "jwc_wpt3d_dev": (_int, [_vp, _int, _int, _vp, _vp, _int, _int, _int, _int, _int, _int]),
    "jwc_axis_dev": (_int, [_vp, _int, _int, _int, _vp, _vp, _i64, _int, _i64, _int]),
    "jwc_dev_alloc": (_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "jwc_dev_free": (_int, [_vp, _vp]),
    "jwc_h2d": (_int, [_vp, _vp, _vp, C.c_size_t]),
    "jwc_d2h": (_int, [_vp, _vp, _vp, C.c_size_t]),
    "jwc_host_alloc_pinned": (_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "jwc_host_free_pinned": (_int, [_vp, _vp]),
    "jwc_copy2d_dev": (_int, [_vp, _vp, C.c_size_t, _vp, C.c_size_t, C.c_size_t, C.c_size_t, _vp]),
    "jwc_set_staging_bytes": (_int, [_vp, C.c_size_t]),
}

class RemoteMap(C.Structure):
    """jwc_remote_map (include/jwave_cuda.h)"""
    _fields_ = [("mode", C.c_int), ("world", C.c_int), ("lg_seg", C.c_int), ("lg_hi", C.c_int),
                ("outer_stride", C.c_int64), ("row_stride", C.c_int64), ("base_off", C.c_int64),
                ("peer", C.c_void_p * 8)]


SIGNATURES["jwc_axis_dev_remote"] = (_int, [_vp, _int, _int, _int, _vp, _i64, _int, _i64, _int, C.POINTER(RemoteMap)])

_lib = None


def load():
    """Load libjwave_cuda.so and type its entry points.  Raises OSError when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise OSError(
                f"{SO_PATH} not found: build it with `make -C jwave_b200/csrc` "
                "(or __graft_entry__.build()); there is no CPU fallback")
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the ABI and the header drift apart
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
