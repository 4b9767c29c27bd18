"""Host-side mirror of JWave's Wavelet objects for the in-scope families.

Only the tap tables and the orthonormal-space construction live here - the arithmetic of
Wavelet.forward / Wavelet.reverse (jwave/transforms/wavelets/Wavelet.java:236-303) is the
CUDA library's job.  The literal tables come from _taps_literal.py (generated from the
reference's tap files by tools/gen_taps.py); the analytic families restate the reference
constructors' expression trees so the doubles are bit-identical to the JVM's.

Citations are relative to the reference's src/main/java/jwave/transforms/wavelets/."""
import math

import numpy as np

from ._taps_bior import BIOR_TAPS
from ._taps_literal import LITERAL_TAPS


class Wavelet:
    """Wavelet.java:38-219: name, wavelengths and the four filters (getters return copies)."""

    def __init__(self, name, scalingDeCom, waveletDeCom=None, scalingReCon=None, waveletReCon=None):
        self._name = name
        self._transformWavelength = 2
        self._scalingDeCom = np.array(scalingDeCom, dtype=np.float64)
        self._motherWavelength = len(self._scalingDeCom)
        if waveletDeCom is None:
            self._buildOrthonormalSpace()
        else:  # Haar1.java:62-72 writes its filters out by hand; the BiOrthogonal classes keep four
            self._waveletDeCom = np.array(waveletDeCom, dtype=np.float64)
            self._scalingReCon = (self._scalingDeCom.copy() if scalingReCon is None
                                  else np.array(scalingReCon, dtype=np.float64))
            self._waveletReCon = (self._waveletDeCom.copy() if waveletReCon is None
                                  else np.array(waveletReCon, dtype=np.float64))

    def _buildOrthonormalSpace(self):
        """Wavelet.java:104-122"""
        L = self._motherWavelength
        s = self._scalingDeCom
        self._waveletDeCom = np.array([s[(L - 1) - i] if i % 2 == 0 else -s[(L - 1) - i] for i in range(L)])
        self._scalingReCon = s.copy()
        self._waveletReCon = self._waveletDeCom.copy()

    def getName(self):
        return self._name

    def __str__(self):
        return self._name

    def getMotherWavelength(self):
        return self._motherWavelength

    def getTransformWavelength(self):
        return self._transformWavelength

    def getScalingDeComposition(self):
        return self._scalingDeCom.copy()

    def getWaveletDeComposition(self):
        return self._waveletDeCom.copy()

    def getScalingReConstruction(self):
        return self._scalingReCon.copy()

    def getWaveletReConstruction(self):
        return self._waveletReCon.copy()


def Haar1():
    """haar/Haar1.java:52-68"""
    sqrt2 = math.sqrt(2.0)
    s = [1.0 / sqrt2, 1.0 / sqrt2]
    return Wavelet("Haar", s, [s[1], -s[0]])


def Daubechies2():
    """daubechies/Daubechies2.java:53-63"""
    sqrt3 = math.sqrt(3.0)
    s = [(1.0 + sqrt3) / 4.0, (3.0 + sqrt3) / 4.0, (3.0 - sqrt3) / 4.0, (1.0 - sqrt3) / 4.0]
    sqrt02 = math.sqrt(2.0)
    return Wavelet("Daubechies 2", [v / sqrt02 for v in s])


def Daubechies3():
    """daubechies/Daubechies3.java:54-66"""
    sqrt10 = math.sqrt(10.0)
    constA = math.sqrt(5.0 + 2.0 * sqrt10)
    s = [(1.0 + 1.0 * sqrt10 + 1.0 * constA) / 16.0,
         (5.0 + 1.0 * sqrt10 + 3.0 * constA) / 16.0,
         (10.0 - 2.0 * sqrt10 + 2.0 * constA) / 16.0,
         (10.0 - 2.0 * sqrt10 - 2.0 * constA) / 16.0,
         (5.0 + 1.0 * sqrt10 - 3.0 * constA) / 16.0,
         (1.0 + 1.0 * sqrt10 - 1.0 * constA) / 16.0]
    sqrt02 = math.sqrt(2.0)
    return Wavelet("Daubechies 3", [v / sqrt02 for v in s])


def Coiflet1():
    """coiflet/Coiflet1.java:52-62 (the code, not its stale inline comments)"""
    sqrt02 = 1.4142135623730951
    sqrt15 = math.sqrt(15.0)
    return Wavelet("Coiflet 1", [
        sqrt02 * (sqrt15 - 3.0) / 32.0,
        sqrt02 * (1.0 - sqrt15) / 32.0,
        sqrt02 * (6.0 - 2 * sqrt15) / 32.0,
        sqrt02 * (2.0 * sqrt15 + 6.0) / 32.0,
        sqrt02 * (sqrt15 + 13.0) / 32.0,
        sqrt02 * (9.0 - sqrt15) / 32.0,
    ])


def _legendre(name, numerators, denominator):
    sqrt02 = math.sqrt(2.0)
    return Wavelet(name, [(v / denominator) / sqrt02 for v in numerators])


def Legendre1():
    """legendre/Legendre1.java:55-64"""
    return _legendre("Legendre 1", (-1.0, -1.0), 1.0)


def Legendre2():
    """legendre/Legendre2.java:52-63"""
    return _legendre("Legendre 2", (-5.0, -3.0, -3.0, -5.0), 8.0)


def Legendre3():
    """legendre/Legendre3.java:52-65"""
    return _legendre("Legendre 3", (-63.0, -35.0, -30.0, -30.0, -35.0, -63.0), 128.0)


def Haar1Orthogonal():
    """haar/Haar1Orthogonal.java:137-161 with the _energyCorrectionFactor of its reverse step
    (:39, :197-199) folded into the reconstruction filters: .5 * (a s + d w) == a (.5 s) + d (.5 w)
    exactly in binary floating point, so the device needs no special case."""
    s = np.array([1.0, 1.0])
    w = np.array([s[1], -s[0]])
    return Wavelet("Haar orthogonal", s, w, 0.5 * s, 0.5 * w)


def _bior_factory(cls):
    def make():
        name, built, arrays = BIOR_TAPS[cls]
        s_de, w_de = (np.array(a, dtype=np.float64) for a in arrays[:2])
        if built:  # biorthogonal/BiOrthogonal.java:43-66 (_buildBiOrthonormalSpace)
            sign = np.where(np.arange(len(s_de)) % 2 == 0, -1.0, 1.0)
            return Wavelet(name, s_de, w_de, sign * w_de, sign * s_de)
        return Wavelet(name, s_de, w_de, arrays[2], arrays[3])
    make.__name__ = cls
    make.__doc__ = f"the four filters of the reference's biorthogonal/{cls}.java"
    return make


def _literal_factory(cls):
    def make():
        name, taps = LITERAL_TAPS[cls]
        return Wavelet(name, taps)
    make.__name__ = cls
    make.__doc__ = f"literal taps of the reference's {cls}.java"
    return make


_FACTORIES = {f.__name__: f for f in (Haar1, Daubechies2, Daubechies3, Coiflet1, Legendre1, Legendre2, Legendre3)}
for _cls in LITERAL_TAPS:
    _FACTORIES[_cls] = _literal_factory(_cls)
    globals()[_cls] = _FACTORIES[_cls]

# section 8(f) row 3: four independent filters (generic one-level kernels, or the fused ones when the
# decomposition / reconstruction pairs happen to be mirrored)
_FACTORIES["Haar1Orthogonal"] = Haar1Orthogonal
for _cls in BIOR_TAPS:
    _FACTORIES[_cls] = _bior_factory(_cls)
    globals()[_cls] = _FACTORIES[_cls]

WAVELET_CLASSES = tuple(_FACTORIES)
# the reference's own test loops (WaveletBuilder.java:427-502) include these BiOrthogonal members only
_BIOR_IN_CREATE2ARR = ("BiOrthogonal11", "BiOrthogonal13", "BiOrthogonal15", "BiOrthogonal31", "BiOrthogonal33",
                       "BiOrthogonal35", "BiOrthogonal37", "BiOrthogonal39")


class WaveletBuilder:
    """WaveletBuilder.java:99-403 / :427-502 restricted to the in-scope families."""

    @staticmethod
    def create(waveletName):
        """Accepts the JWave display name ("Daubechies 4") or the class name ("Daubechies4")."""
        if waveletName in _FACTORIES:
            return _FACTORIES[waveletName]()
        for f in _FACTORIES.values():
            w = f()
            if w.getName() == waveletName:
                return w
        from .exceptions import JWaveFailure
        raise JWaveFailure("WaveletBuilder::create - unknown type of wavelet for given string!")

    @staticmethod
    def create2arr():
        """The wavelets the reference's own test loops run over (WaveletBuilder.java:427-502): no
        Legendre, no Haar1Orthogonal, and of the BiOrthogonal family only 1/x and 3/x."""
        return [f() for n, f in _FACTORIES.items()
                if not n.startswith("Legendre") and n != "Haar1Orthogonal"
                and (not n.startswith("BiOrthogonal") or n in _BIOR_IN_CREATE2ARR)]
