"""Host-side mirror of jwave/compressions/Compressor.java and CompressorMagnitude.java, backed by
libjwave_cuda.so (jwc_compress_magnitude).  Thresholding is the step JWave runs on the coefficients
right after the forward transform; there is no CPU fallback."""
import ctypes as C

import numpy as np

from . import _lib
from .exceptions import JWaveError, JWaveException, JWaveFailure
from .transforms import CudaContext, _as_f64


class Compressor:
    """compressions/Compressor.java:36-187 - holds the threshold and the last magnitude."""

    def __init__(self, threshold=1.0):
        self._magnitude = 0.0
        try:  # Compressor.java:52-66: a non-positive threshold is reported and replaced by 1
            if threshold <= 0.0:
                raise JWaveFailure("Compressor - given threshold should be larger than zero!")
        except JWaveException as e:
            e.showMessage()
            print("Compressor - setting threshold to default value: 1.0")
            threshold = 1.0
        self._threshold = float(threshold)

    def getThreshold(self):
        return self._threshold

    def getMagnitude(self):
        return self._magnitude

    @staticmethod
    def calcCompressionRate(arr):
        """Compressor.java:146-160: percentage of exact zeros"""
        arr = np.asarray(arr)
        zeros = int(np.count_nonzero(arr == 0.0))
        return zeros / arr.size * 100.0 if zeros else 0.0

    def compress(self, arrHilb):
        raise JWaveError("Compressor#compress - method is not implemented")


class CompressorMagnitude(Compressor):
    """compressions/CompressorMagnitude.java:36-118: magnitude = mean |c| of the whole array (any rank),
    coefficients below magnitude * threshold become zero."""

    def __init__(self, threshold=1.0, context=None, device=0):
        super().__init__(threshold)
        self._ctx = context if context is not None else CudaContext.default(device)

    def compress(self, arrHilb):
        src = _as_f64(arrHilb)
        if src.size == 0:
            raise JWaveFailure("CompressorMagnitude#compress - empty array")
        dst = np.empty_like(src)
        mag = C.c_double(0.0)
        with self._ctx.lock:
            st = self._ctx._lib.jwc_compress_magnitude(self._ctx.handle, src.ctypes.data, dst.ctypes.data, src.size,
                                                       self._threshold, C.byref(mag))
            self._ctx.check(st, "CompressorMagnitude#compress")
        self._magnitude = mag.value
        return dst
