"""ctypes binding of oracle/libjw_oracle.so (the C restatement of JWave's CPU path).

TEST INFRASTRUCTURE ONLY - see oracle/jw_oracle.h for scope and parity status.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libjw_oracle.so")

FWT, WPT = 0, 1
FORWARD, REVERSE = 0, 1
OK, ERR_NOT_BINARY, ERR_LEVEL, ERR_ARG = 0, 1, 2, 3
MAX_TAPS = 64


class Wavelet(C.Structure):
    _fields_ = [
        ("cls", C.c_char * 32),
        ("name", C.c_char * 32),
        ("motherWavelength", C.c_int),
        ("transformWavelength", C.c_int),
        ("scalingDeCom", C.c_double * MAX_TAPS),
        ("waveletDeCom", C.c_double * MAX_TAPS),
        ("scalingReCon", C.c_double * MAX_TAPS),
        ("waveletReCon", C.c_double * MAX_TAPS),
        ("reconFactor", C.c_double),
    ]

    def taps(self):
        L = self.motherWavelength
        return tuple(np.array(a[:L], dtype=np.float64) for a in
                     (self.scalingDeCom, self.waveletDeCom, self.scalingReCon, self.waveletReCon))


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc, -ffp-contract=off)."""
    src = [os.path.join(_HERE, f) for f in ("jw_oracle.c", "jw_oracle.h", "jw_taps_literal.inc", "jw_taps_bior.inc")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None
_dp = C.POINTER(C.c_double)
_wp = C.POINTER(Wavelet)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.jwo_wavelet_count.restype = C.c_int
        L.jwo_wavelet_at.restype = _wp
        L.jwo_wavelet_at.argtypes = [C.c_int]
        L.jwo_wavelet_find.restype = _wp
        L.jwo_wavelet_find.argtypes = [C.c_char_p]
        L.jwo_is_binary.argtypes = [C.c_int]
        L.jwo_get_exponent.argtypes = [C.c_double]
        L.jwo_wavelet_forward.restype = None
        L.jwo_wavelet_forward.argtypes = [_wp, _dp, C.c_int, _dp]
        L.jwo_wavelet_reverse.restype = None
        L.jwo_wavelet_reverse.argtypes = [_wp, _dp, C.c_int, _dp]
        L.jwo_1d.argtypes = [C.c_int, C.c_int, _wp, _dp, C.c_int, C.c_int, _dp]
        L.jwo_2d.argtypes = [C.c_int, C.c_int, _wp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
        L.jwo_3d.argtypes = [C.c_int, C.c_int, _wp, _dp] + [C.c_int] * 6 + [_dp]
        L.jwo_batch_1d.argtypes = [C.c_int, C.c_int, _wp, _dp, C.c_long, C.c_int, C.c_int, _dp, C.c_int]
        L.jwo_parallel_wpt.argtypes = [C.c_int, _wp, _dp, C.c_long, C.c_int, C.c_int, _dp, C.c_int]
        L.jwo_parallel_2d.argtypes = [C.c_int, C.c_int, _wp, _dp, C.c_long, C.c_int, C.c_int, C.c_int,
                                      C.c_int, _dp, C.c_int]
        L.jwo_parallel_3d.argtypes = [C.c_int, C.c_int, _wp, _dp] + [C.c_int] * 6 + [_dp, C.c_int]
        L.jwo_decompose.argtypes = [C.c_int, C.POINTER(C.c_int)]
        L.jwo_aed.argtypes = [C.c_int, C.c_int, _wp, _dp, C.c_int, _dp]
        L.jwo_compress_magnitude.restype = C.c_double
        L.jwo_compress_magnitude.argtypes = [_dp, C.c_long, C.c_double, _dp]
        L.jwo_max_threads.restype = C.c_int
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, status):
        super().__init__(f"oracle status {status}")
        self.status = status


def _p(a):
    return a.ctypes.data_as(_dp)


def _in(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _check(st):
    if st:
        raise OracleError(st)


def wavelet_names():
    L = lib()
    return [L.jwo_wavelet_at(i).contents.cls.decode() for i in range(L.jwo_wavelet_count())]


def wavelet(name):
    w = lib().jwo_wavelet_find(name.encode())
    if not w:
        raise KeyError(name)
    return w


def wavelet_forward(name, x, n=None):
    x = _in(x)
    n = len(x) if n is None else n
    out = np.empty(n)
    lib().jwo_wavelet_forward(wavelet(name), _p(x), n, _p(out))
    return out


def wavelet_reverse(name, x, n=None):
    x = _in(x)
    n = len(x) if n is None else n
    out = np.empty(n)
    lib().jwo_wavelet_reverse(wavelet(name), _p(x), n, _p(out))
    return out


def transform_1d(kind, direction, name, x, level=None):
    x = _in(x)
    n = x.shape[-1]
    if level is None:  # WaveletTransform.java:77-88 - default level is log2 N
        if n <= 0 or n & (n - 1):
            raise OracleError(ERR_NOT_BINARY)
        level = n.bit_length() - 1
    if x.ndim == 1:
        out = np.empty_like(x)
        _check(lib().jwo_1d(kind, direction, wavelet(name), _p(x), n, level, _p(out)))
        return out
    flat = x.reshape(-1, n)
    out = np.empty_like(flat)
    _check(lib().jwo_batch_1d(kind, direction, wavelet(name), _p(flat), flat.shape[0], n, level, _p(out), 0))
    return out.reshape(x.shape)


def transform_2d(kind, direction, name, m, lvlM=None, lvlN=None):
    m = _in(m)
    rows, cols = m.shape
    lvlM = lib().jwo_get_exponent(float(rows)) if lvlM is None else lvlM
    lvlN = lib().jwo_get_exponent(float(cols)) if lvlN is None else lvlN
    out = np.empty_like(m)
    _check(lib().jwo_2d(kind, direction, wavelet(name), _p(m), rows, cols, lvlM, lvlN, _p(out)))
    return out


def transform_3d(kind, direction, name, s, lvlP=None, lvlQ=None, lvlR=None):
    s = _in(s)
    P, Q, R = s.shape
    ge = lib().jwo_get_exponent
    lvlP = ge(float(P)) if lvlP is None else lvlP
    lvlQ = ge(float(Q)) if lvlQ is None else lvlQ
    lvlR = ge(float(R)) if lvlR is None else lvlR
    out = np.empty_like(s)
    _check(lib().jwo_3d(kind, direction, wavelet(name), _p(s), P, Q, R, lvlP, lvlQ, lvlR, _p(out)))
    return out


def batch_1d(kind, direction, name, x, level, threads=0):
    x = _in(x)
    out = np.empty_like(x)
    _check(lib().jwo_batch_1d(kind, direction, wavelet(name), _p(x), x.shape[0], x.shape[1], level, _p(out), threads))
    return out


def parallel_wpt(direction, name, x, level, threads=0):
    x = _in(x)
    out = np.empty_like(x)
    _check(lib().jwo_parallel_wpt(direction, wavelet(name), _p(x), x.shape[0], x.shape[1], level, _p(out), threads))
    return out


def parallel_2d(kind, direction, name, x, lvlM, lvlN, threads=0):
    x = _in(x)
    out = np.empty_like(x)
    b, rows, cols = x.shape
    _check(lib().jwo_parallel_2d(kind, direction, wavelet(name), _p(x), b, rows, cols, lvlM, lvlN, _p(out), threads))
    return out


def parallel_3d(kind, direction, name, s, lvlP, lvlQ, lvlR, threads=0):
    """ParallelTransform.forward/reverse(double[][][], lvlP, lvlQ, lvlR) on all host threads."""
    s = _in(s)
    P, Q, R = s.shape
    out = np.empty_like(s)
    _check(lib().jwo_parallel_3d(kind, direction, wavelet(name), _p(s), P, Q, R, lvlP, lvlQ, lvlR, _p(out), threads))
    return out


def decompose(number):
    buf = (C.c_int * 32)()
    cnt = lib().jwo_decompose(int(number), buf)
    return [buf[i] for i in range(cnt)]


def aed(kind, direction, name, x):
    """AncientEgyptianDecomposition over the given 1-D array of ANY length (rows of a 2-D array)."""
    x = _in(x)
    flat = x.reshape(-1, x.shape[-1])
    out = np.empty_like(flat)
    for i in range(flat.shape[0]):
        _check(lib().jwo_aed(kind, direction, wavelet(name), _p(flat[i]), flat.shape[1], _p(out[i])))
    return out.reshape(x.shape)


def compress_magnitude(x, threshold):
    """CompressorMagnitude.compress on an array of any rank -> (compressed, magnitude)."""
    x = _in(x)
    out = np.empty_like(x)
    mag = lib().jwo_compress_magnitude(_p(x.reshape(-1)), x.size, float(threshold), _p(out.reshape(-1)))
    return out, mag


def max_threads():
    return lib().jwo_max_threads()
