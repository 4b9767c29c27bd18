"""Host-side mirror of JWave's transform plug-in API, backed by libjwave_cuda.so.

The reference's toolchain (Java 21) is absent from this image, so the host layer above the
C ABI is written in Python with the reference's names, argument meaning and error behaviour;
the Java 21 / java.lang.foreign classes a JWave maintainer would add are in java/ and
INTEGRATION.md.  Python has no overloading, so `forward` / `reverse` dispatch on the array
rank exactly as Java dispatches on double[], double[][] and double[][][].

    BasicTransform                 jwave/transforms/BasicTransform.java:42-699
    WaveletTransform               jwave/transforms/WaveletTransform.java:34-182
    CudaFastWaveletTransform       replaces jwave/transforms/FastWaveletTransform.java:71-153
    CudaWaveletPacketTransform     replaces jwave/transforms/WaveletPacketTransform.java:73-191
                                   (and the Pooled / Parallel variants, same arithmetic)
    Transform                      jwave/Transform.java:59-512 (the unchanged facade)

All arithmetic runs in the CUDA library; there is no CPU fallback.
"""
import ctypes as C
import threading
import traceback

import numpy as np

from . import _lib
from .exceptions import JWaveError, JWaveException, JWaveFailure
from .wavelets import Wavelet

_NOT_BINARY = ("given array length is not 2^p | p E N ... = 1, 2, 4, 8, 16, 32, .. "
               "please use the Ancient Egyptian Decomposition for any other array length!")


class CudaContext:
    """One jwc_ctx: one GPU, one stream.  Calls are serialised with a lock (the C context is
    not re-entrant); use one context per thread for concurrency, as include/jwave_cuda.h says."""

    _defaults = {}
    _defaults_lock = threading.Lock()

    def __init__(self, device=0):
        """device: a CUDA ordinal, or a list of ordinals for ONE context that drives several GPUs of the box
        (jwc_create_multi: batched host-buffer calls are sharded over them, a 3-D volume is slab-decomposed)."""
        self._lib = _lib.load()
        handle = C.c_void_p()
        if isinstance(device, (list, tuple)):
            devs = (C.c_int * len(device))(*[int(d) for d in device])
            st = self._lib.jwc_create_multi(C.byref(handle), devs, len(device))
            device = device[0] if device else 0
        else:
            st = self._lib.jwc_create(C.byref(handle), int(device))
        if st != _lib.OK:
            msg = self._lib.jwc_last_error(None)
            raise JWaveError("jwc_create failed: " + (msg.decode() if msg else f"status {st}"))
        self.handle = handle
        self.device = int(device)
        self.lock = threading.RLock()
        self._wids = {}

    @classmethod
    def default(cls, device=0):
        with cls._defaults_lock:
            ctx = cls._defaults.get(device)
            if ctx is None or ctx.handle is None:
                ctx = cls._defaults[device] = CudaContext(device)
            return ctx

    def register(self, wavelet):
        """jwc_set_wavelet with the four getter arrays (Wavelet.java:178-219)."""
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (
            wavelet.getScalingDeComposition(), wavelet.getWaveletDeComposition(),
            wavelet.getScalingReConstruction(), wavelet.getWaveletReConstruction())]
        # keyed by the tap CONTENTS: builders hand out a fresh Wavelet object per call, and the C side keeps
        # every registered filter set until the context is destroyed
        key = (wavelet.getMotherWavelength(),) + tuple(a.tobytes() for a in arrs)
        with self.lock:
            if key in self._wids:
                return self._wids[key]
            wid = C.c_int(-1)
            dp = C.POINTER(C.c_double)
            st = self._lib.jwc_set_wavelet(self.handle, wavelet.getMotherWavelength(),
                                           *[a.ctypes.data_as(dp) for a in arrs], C.byref(wid))
            self.check(st, "jwc_set_wavelet")
            self._wids[key] = wid.value
            return wid.value

    def last_error(self):
        msg = self._lib.jwc_last_error(self.handle)
        return msg.decode() if msg else ""

    def check(self, st, where):
        """Status codes of include/jwave_cuda.h -> the reference's exception classes."""
        if st == _lib.OK:
            return
        if st == _lib.ERR_NOT_BINARY:
            raise JWaveFailure(f"{where} - {_NOT_BINARY}")
        if st == _lib.ERR_LEVEL:
            raise JWaveFailure(f"{where} - given level is out of range for given array")
        if st == _lib.ERR_ARG:
            raise JWaveFailure(f"{where} - {self.last_error() or 'bad argument'}")
        raise JWaveError(f"{where} - {self.last_error() or 'device failure'} (status {st})")

    def launch_count(self):
        return int(self._lib.jwc_launch_count(self.handle))

    def device_count(self):
        return int(self._lib.jwc_device_count(self.handle))

    def profile(self, on):
        """Bracket every kernel launch with CUDA events (jwc_profile_enable)."""
        self.check(self._lib.jwc_profile_enable(self.handle, 1 if on else 0), "jwc_profile_enable")

    def profile_report(self):
        """[(label, launches, total_ms, samples_per_launch, levels)] since profiling was enabled / last read."""
        buf = C.create_string_buffer(1 << 16)
        self.check(self._lib.jwc_profile_report(self.handle, buf, len(buf)), "jwc_profile_report")
        rows = []
        for line in buf.value.decode().splitlines():
            name, n, ms, units, levels = line.split(",")
            rows.append((name, int(n), float(ms), float(units), int(levels)))
        return rows

    def set_stream(self, cuda_stream):
        """Use the given cudaStream_t handle (0 = CUDA's legacy default stream) for *_dev calls."""
        self.check(self._lib.jwc_set_stream(self.handle, C.c_void_p(cuda_stream or 0)), "jwc_set_stream")

    def reset_stream(self):
        self.check(self._lib.jwc_reset_stream(self.handle), "jwc_reset_stream")

    def sync(self):
        self.check(self._lib.jwc_sync(self.handle), "jwc_sync")

    def set_staging_bytes(self, nbytes):
        self.check(self._lib.jwc_set_staging_bytes(self.handle, int(nbytes)), "jwc_set_staging_bytes")

    def close(self):
        with self.lock:
            if self.handle is not None:
                self._lib.jwc_destroy(self.handle)
                self.handle = None


def _as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class BasicTransform:
    """BasicTransform.java:42 - the plug-in type.  Subclasses implement the 1-D methods; the
    2-D / 3-D drivers below compose them row by row exactly as BasicTransform.java:361-659
    does (the CUDA subclasses override them with whole-array passes)."""

    def __init__(self):
        self._name = None

    def getName(self):
        return self._name

    def getWavelet(self):
        """BasicTransform.java:82"""
        raise JWaveFailure("BasicTransform#getWavelet - not available")

    # ---- 1-D, to be implemented -------------------------------------------------------------
    def _forward1(self, arrTime, level=None):
        raise JWaveError("BasicTransform#forward - method is not implemented")

    def _reverse1(self, arrHilb, level=None):
        raise JWaveError("BasicTransform#reverse - method is not implemented")

    # ---- rank dispatch (Java overloads) ------------------------------------------------------
    def forward(self, data, *levels):
        return self._dispatch("forward", data, levels)

    def reverse(self, data, *levels):
        return self._dispatch("reverse", data, levels)

    def _dispatch(self, which, data, levels):
        a = np.asarray(data, dtype=np.float64)
        f1 = self._forward1 if which == "forward" else self._reverse1
        if a.ndim == 1:
            if len(levels) > 1:
                raise TypeError("1-D transform takes at most one level")
            return f1(a, *levels)
        if a.ndim == 2:
            if len(levels) == 0:  # BasicTransform.java:336-342
                levels = (MathToolKit.getExponent(a.shape[0]), MathToolKit.getExponent(a.shape[1]))
            if len(levels) != 2:
                raise TypeError("2-D transform takes (lvlM, lvlN)")
            return getattr(self, "_" + which + "2")(a, *levels)
        if a.ndim == 3:
            if len(levels) == 0:  # BasicTransform.java:487-495
                levels = tuple(MathToolKit.getExponent(s) for s in a.shape)
            if len(levels) != 3:
                raise TypeError("3-D transform takes (lvlP, lvlQ, lvlR)")
            return getattr(self, "_" + which + "3")(a, *levels)
        raise JWaveFailure("BasicTransform - only 1-D, 2-D and 3-D arrays are supported")

    # ---- 2-D / 3-D drivers of the reference --------------------------------------------------
    def _forward2(self, matTime, lvlM, lvlN):
        """BasicTransform.java:361-399: every row with lvlN, then every column with lvlM."""
        matHilb = np.empty_like(matTime)
        for i in range(matTime.shape[0]):
            matHilb[i, :] = self._forward1(matTime[i, :].copy(), lvlN)
        for j in range(matTime.shape[1]):
            matHilb[:, j] = self._forward1(matHilb[:, j].copy(), lvlM)
        return matHilb

    def _reverse2(self, matFreq, lvlM, lvlN):
        """BasicTransform.java:436-474: columns first, then rows."""
        matTime = np.empty_like(matFreq)
        for j in range(matFreq.shape[1]):
            matTime[:, j] = self._reverse1(matFreq[:, j].copy(), lvlM)
        for i in range(matFreq.shape[0]):
            matTime[i, :] = self._reverse1(matTime[i, :].copy(), lvlN)
        return matTime

    def _forward3(self, spcTime, lvlP, lvlQ, lvlR):
        """BasicTransform.java:509-566.  The slice transform receives (lvlP, lvlQ) - the
        reference's level shift - then the outer axis gets lvlR."""
        spcHilb = np.empty_like(spcTime)
        for i in range(spcTime.shape[0]):
            spcHilb[i] = self._forward2(spcTime[i], lvlP, lvlQ)
        for j in range(spcTime.shape[1]):
            for k in range(spcTime.shape[2]):
                spcHilb[:, j, k] = self._forward1(spcHilb[:, j, k].copy(), lvlR)
        return spcHilb

    def _reverse3(self, spcHilb, lvlP, lvlQ, lvlR):
        """BasicTransform.java:602-659"""
        spcTime = np.empty_like(spcHilb)
        for i in range(spcHilb.shape[0]):
            spcTime[i] = self._reverse2(spcHilb[i], lvlP, lvlQ)
        for j in range(spcHilb.shape[1]):
            for k in range(spcHilb.shape[2]):
                spcTime[:, j, k] = self._reverse1(spcTime[:, j, k].copy(), lvlR)
        return spcTime

    # ---- helpers -------------------------------------------------------------------------------
    def isBinary(self, number):
        """BasicTransform.java:671-675"""
        return MathToolKit.isBinary(number)

    def calcExponent(self, number):
        """BasicTransform.java:683-697"""
        if not self.isBinary(number):
            raise JWaveFailure("BasicTransform#calcExponent - given number is not binary: "
                               "2^p | pEN .. = 1, 2, 4, 8, 16, 32, .. ")
        return MathToolKit.getExponent(number)


class MathToolKit:
    """jwave/tools/MathToolKit.java:185-189 and :202-208 (the two functions on the path)."""

    @staticmethod
    def isBinary(number):
        number = int(number)
        return number > 0 and (number & (number - 1)) == 0

    @staticmethod
    def getExponent(f):
        """p with 2^p <= f < 2^(p+1); the reference's float log is exact on every 2^p (F14)."""
        f = int(f)
        return f.bit_length() - 1 if f > 0 else 0


class WaveletTransform(BasicTransform):
    """WaveletTransform.java:34-182: holds the wavelet, defaults the level to log2 N,
    decompose / recompose."""

    def __init__(self, wavelet):
        super().__init__()
        if not isinstance(wavelet, Wavelet):
            raise JWaveFailure("WaveletTransform - given object is not of type Wavelet")
        self._wavelet = wavelet

    def getWavelet(self):
        """WaveletTransform.java:62"""
        return self._wavelet

    def decompose(self, arrTime):
        """WaveletTransform.java:136-146: forward(x, p) for every p = 0 .. log2 N."""
        arrTime = _as_f64(arrTime)
        levels = self.calcExponent(arrTime.shape[0])
        return np.stack([self.forward(arrTime, p) for p in range(levels + 1)])

    def recompose(self, matDeComp, level):
        """WaveletTransform.java:166-176"""
        if level < 0 or level >= len(matDeComp):
            raise JWaveFailure("WaveletTransform#recompose - given level is out of range")
        return self.reverse(matDeComp[level], level)


class _CudaWaveletTransform(WaveletTransform):
    _KIND = None
    _CLS = None

    def __init__(self, wavelet, context=None, device=0):
        super().__init__(wavelet)
        self._ctx = context if context is not None else CudaContext.default(device)
        self._wid = self._ctx.register(wavelet)
        L = self._ctx._lib
        self._f1d, self._f2d, self._f3d = ((L.jwc_fwt1d, L.jwc_fwt2d, L.jwc_fwt3d) if self._KIND == _lib.FWT
                                           else (L.jwc_wpt1d, L.jwc_wpt2d, L.jwc_wpt3d))

    # -- validation in the reference's order and words ----------------------------------------
    def _check(self, length, level, where):
        if not self.isBinary(length):
            raise JWaveFailure(f"{self._CLS}#{where} - {_NOT_BINARY}")
        noOfLevels = self.calcExponent(length)
        if level is None:  # WaveletTransform.java:77-88, :101-112
            return noOfLevels
        if level < 0 or level > noOfLevels:
            raise JWaveFailure(f"{self._CLS}#{where} - given level is out of range for given array")
        return int(level)

    def _call(self, fn, where, direction, src, *dims):
        src = _as_f64(src)
        dst = np.empty_like(src)  # the reference never mutates its input
        with self._ctx.lock:
            if self._ctx.handle is None:
                raise JWaveError(f"{self._CLS}#{where} - the CUDA context is closed")
            st = fn(self._ctx.handle, self._wid, direction, src.ctypes.data, dst.ctypes.data, *dims)
            self._ctx.check(st, f"{self._CLS}#{where}")
        return dst

    # -- 1-D ----------------------------------------------------------------------------------------
    def _forward1(self, arrTime, level=None):
        level = self._check(arrTime.shape[0], level, "forward")
        return self._call(self._f1d, "forward", _lib.FORWARD, arrTime, 1, arrTime.shape[0], level)

    def _reverse1(self, arrHilb, level=None):
        level = self._check(arrHilb.shape[0], level, "reverse")
        return self._call(self._f1d, "reverse", _lib.REVERSE, arrHilb, 1, arrHilb.shape[0], level)

    # -- batched double[][] entry point (new; rows are independent signals) ---------------------
    def forwardBatch(self, signals, level=None):
        signals = _as_f64(signals)
        if signals.ndim != 2:
            raise JWaveFailure(f"{self._CLS}#forwardBatch - expected a [batch][n] array")
        level = self._check(signals.shape[1], level, "forwardBatch")
        return self._call(self._f1d, "forwardBatch", _lib.FORWARD, signals, signals.shape[0], signals.shape[1], level)

    def reverseBatch(self, coeffs, level=None):
        coeffs = _as_f64(coeffs)
        if coeffs.ndim != 2:
            raise JWaveFailure(f"{self._CLS}#reverseBatch - expected a [batch][n] array")
        level = self._check(coeffs.shape[1], level, "reverseBatch")
        return self._call(self._f1d, "reverseBatch", _lib.REVERSE, coeffs, coeffs.shape[0], coeffs.shape[1], level)

    # -- decompose: every level materialised, one upload / one launch per level / one download ----------
    def decompose(self, arrTime):
        """WaveletTransform.java:136-146 - row p is forward(arrTime, p)."""
        arrTime = _as_f64(arrTime)
        if arrTime.ndim != 1:
            raise JWaveFailure(f"{self._CLS}#decompose - expected a 1-D array")
        return self.decomposeBatch(arrTime[None, :])[0]

    def decomposeBatch(self, signals):
        """[batch][n] -> [batch][log2 n + 1][n] (new, like forwardBatch)."""
        signals = _as_f64(signals)
        if signals.ndim != 2:
            raise JWaveFailure(f"{self._CLS}#decomposeBatch - expected a [batch][n] array")
        batch, n = signals.shape
        self._check(n, None, "decompose")
        out = np.empty((batch, self.calcExponent(n) + 1, n))
        with self._ctx.lock:
            if self._ctx.handle is None:
                raise JWaveError(f"{self._CLS}#decompose - the CUDA context is closed")
            st = self._ctx._lib.jwc_decompose1d(self._ctx.handle, self._wid, self._KIND, signals.ctypes.data,
                                                out.ctypes.data, batch, n)
            self._ctx.check(st, f"{self._CLS}#decompose")
        return out

    # -- 2-D / 3-D: whole-array passes instead of 16 384 tiny launches ---------------------------
    def _forward2(self, matTime, lvlM, lvlN):
        rows, cols = matTime.shape
        self._check(cols, lvlN, "forward")  # rows are transformed first (BasicTransform.java:369)
        self._check(rows, lvlM, "forward")
        return self._call(self._f2d, "forward", _lib.FORWARD, matTime, 1, rows, cols, lvlM, lvlN)

    def _reverse2(self, matFreq, lvlM, lvlN):
        rows, cols = matFreq.shape
        self._check(rows, lvlM, "reverse")  # columns first (BasicTransform.java:444)
        self._check(cols, lvlN, "reverse")
        return self._call(self._f2d, "reverse", _lib.REVERSE, matFreq, 1, rows, cols, lvlM, lvlN)

    def forwardBatch2D(self, mats, lvlM=None, lvlN=None):
        mats = _as_f64(mats)
        if mats.ndim != 3:
            raise JWaveFailure("forwardBatch2D - expected a [batch][rows][cols] array")
        b, rows, cols = mats.shape
        lvlN = self._check(cols, lvlN, "forwardBatch2D")
        lvlM = self._check(rows, lvlM, "forwardBatch2D")
        return self._call(self._f2d, "forwardBatch2D", _lib.FORWARD, mats, b, rows, cols, lvlM, lvlN)

    def reverseBatch2D(self, mats, lvlM=None, lvlN=None):
        mats = _as_f64(mats)
        if mats.ndim != 3:
            raise JWaveFailure("reverseBatch2D - expected a [batch][rows][cols] array")
        b, rows, cols = mats.shape
        lvlM = self._check(rows, lvlM, "reverseBatch2D")
        lvlN = self._check(cols, lvlN, "reverseBatch2D")
        return self._call(self._f2d, "reverseBatch2D", _lib.REVERSE, mats, b, rows, cols, lvlM, lvlN)

    def _forward3(self, spcTime, lvlP, lvlQ, lvlR):
        P, Q, R = spcTime.shape
        self._check(R, lvlQ, "forward")  # F5: the innermost axis gets lvlQ ...
        self._check(Q, lvlP, "forward")  # ... the middle axis lvlP ...
        self._check(P, lvlR, "forward")  # ... the outer axis lvlR
        return self._call(self._f3d, "forward", _lib.FORWARD, spcTime, P, Q, R, lvlP, lvlQ, lvlR)

    def _reverse3(self, spcHilb, lvlP, lvlQ, lvlR):
        P, Q, R = spcHilb.shape
        self._check(Q, lvlP, "reverse")
        self._check(R, lvlQ, "reverse")
        self._check(P, lvlR, "reverse")
        return self._call(self._f3d, "reverse", _lib.REVERSE, spcHilb, P, Q, R, lvlP, lvlQ, lvlR)

    # -- lifecycle (ParallelWaveletPacketTransform.java:293-305 exposes shutdown()) ---------------
    def close(self):
        """Contexts obtained from CudaContext.default() are shared and stay open."""
        if self._ctx is not None and self._ctx not in CudaContext._defaults.values():
            self._ctx.close()

    shutdown = close


class CudaFastWaveletTransform(_CudaWaveletTransform):
    """Drop-in for FastWaveletTransform (jwave/transforms/FastWaveletTransform.java:38-154)."""
    _KIND = _lib.FWT
    _CLS = "FastWaveletTransform"

    def __init__(self, wavelet, context=None, device=0):
        super().__init__(wavelet, context, device)
        self._name = "Fast Wavelet Transform"  # FastWaveletTransform.java:51


class CudaWaveletPacketTransform(_CudaWaveletTransform):
    """Drop-in for WaveletPacketTransform and its Pooled / Parallel variants
    (jwave/transforms/WaveletPacketTransform.java:40-193)."""
    _KIND = _lib.WPT
    _CLS = "WaveletPacketTransform"

    def __init__(self, wavelet, context=None, device=0):
        super().__init__(wavelet, context, device)
        self._name = "Wavelet Packet Transform"  # WaveletPacketTransform.java:53


class AncientEgyptianDecomposition(BasicTransform):
    """Drop-in for jwave/transforms/AncientEgyptianDecomposition.java:38-185: accepts arrays of ANY
    length by splitting them into their binary expansion (13 = 8 + 4 + 1, MathToolKit.decompose,
    tools/MathToolKit.java:57-84) and running the wrapped transform at full depth on every block.
    Here the wrapped transform must be one of the CUDA transforms; all blocks of all signals run on
    the GPU in one native call (jwc_aed1d)."""

    def __init__(self, waveTransform):
        super().__init__()
        if not isinstance(waveTransform, _CudaWaveletTransform):
            raise JWaveFailure("AncientEgyptianDecomposition - expected a CudaFastWaveletTransform or "
                               "CudaWaveletPacketTransform")
        self._basicTransform = waveTransform
        self._name = "Ancient Egyptian Decomposition"

    def getWavelet(self):
        return self._basicTransform.getWavelet()

    def _run(self, direction, arr, where):
        arr = _as_f64(arr)
        if arr.ndim not in (1, 2) or arr.shape[-1] < 1:
            raise JWaveFailure("the supported number for decomposition is smaller than one")
        t = self._basicTransform
        batch = 1 if arr.ndim == 1 else arr.shape[0]
        return t._call(lambda h, wid, d, src, dst, b, n: t._ctx._lib.jwc_aed1d(h, wid, t._KIND, d, src, dst, b, n),
                       where, direction, arr, batch, arr.shape[-1])

    def _forward1(self, arrTime, level=None):
        if level is not None:  # the reference class has no levelled overload (BasicTransform.java:129 throws)
            raise JWaveError("BasicTransform#forward - method is not implemented")
        return self._run(_lib.FORWARD, arrTime, "AncientEgyptianDecomposition#forward")

    def _reverse1(self, arrHilb, level=None):
        if level is not None:
            raise JWaveError("BasicTransform#reverse - method is not implemented")
        return self._run(_lib.REVERSE, arrHilb, "AncientEgyptianDecomposition#reverse")

    def forwardBatch(self, signals):
        """rows are independent signals of one (arbitrary) length"""
        return self._run(_lib.FORWARD, signals, "AncientEgyptianDecomposition#forwardBatch")

    def reverseBatch(self, coeffs):
        return self._run(_lib.REVERSE, coeffs, "AncientEgyptianDecomposition#reverseBatch")

    @staticmethod
    def decompose(number):
        """MathToolKit.decompose (tools/MathToolKit.java:57-84): exponents, largest first."""
        if number < 1:
            raise JWaveFailure("the supported number for decomposition is smaller than one")
        return [p for p in range(int(number).bit_length() - 1, -1, -1) if (int(number) >> p) & 1]


class Transform:
    """jwave/Transform.java:59-512 - the facade user code holds.  It swallows JWaveException,
    prints it and returns None (Java: null), e.g. Transform.java:81-90."""

    def __init__(self, transform):
        self._basicTransform = transform
        try:
            if transform is None:
                raise JWaveFailure("given object is null!")
            if not isinstance(transform, BasicTransform):
                raise JWaveFailure("given object is not of type BasicTransform")
        except JWaveException as e:
            e.showMessage()
            traceback.print_exc()

    def _guard(self, fn, *args):
        try:
            return fn(*args)
        except JWaveException as e:
            e.showMessage()
            traceback.print_exc()
            return None

    def forward(self, data, *levels):
        return self._guard(self._basicTransform.forward, data, *levels)

    def reverse(self, data, *levels):
        return self._guard(self._basicTransform.reverse, data, *levels)

    def decompose(self, arrTime):
        return self._guard(self._basicTransform.decompose, arrTime)

    def recompose(self, matDeComp, level):
        return self._guard(self._basicTransform.recompose, matDeComp, level)

    def getBasicTransform(self):
        return self._basicTransform


class TransformBuilder:
    """TransformBuilder.java:40-95 with the GPU classes registered.  There are
    no CPU transforms in this package, so the reference's own names map to the CUDA subclasses too; the
    "Cuda ..." names are what a JWave maintainer would add to the switch."""
    _NAMES = {
        "Fast Wavelet Transform": "CudaFastWaveletTransform",
        "Wavelet Packet Transform": "CudaWaveletPacketTransform",
        "Cuda Fast Wavelet Transform": "CudaFastWaveletTransform",
        "Cuda Wavelet Packet Transform": "CudaWaveletPacketTransform",
    }

    @staticmethod
    def create(transformName, wavelet, context=None):
        """create(String, Wavelet) / create(String, String) (TransformBuilder.java:40-93)."""
        from .wavelets import WaveletBuilder
        if isinstance(wavelet, str):
            wavelet = WaveletBuilder.create(wavelet)
        cls = TransformBuilder._NAMES.get(transformName)
        if cls is None:  # incl. "Discrete Fourier Transform": not on this path
            raise JWaveFailure("TransformBuilder::create - unknown type of transform for given string!")
        return Transform(globals()[cls](wavelet, context=context))

    @staticmethod
    def identify(transform):
        """TransformBuilder.java:97-110: the name stored in the wrapped BasicTransform."""
        return transform.getBasicTransform().getName()
