"""Test adapters: the same reference-derived checks run against

  * the CPU oracle  (OracleFWT / OracleWPT: the product's BasicTransform host drivers with the
    oracle's 1-D transform plugged in - CPU tests), and
  * the CUDA library (CudaFastWaveletTransform / CudaWaveletPacketTransform - GPU tests).
"""
import numpy as np

from jwave_b200 import WaveletBuilder
from jwave_b200.exceptions import JWaveFailure
from jwave_b200.transforms import WaveletTransform
from oracle import c_oracle as co


class _OracleTransform(WaveletTransform):
    KIND = None

    def __init__(self, wavelet_cls):
        super().__init__(WaveletBuilder.create(wavelet_cls))
        self.cls = wavelet_cls

    def _run(self, direction, arr, level):
        try:
            return co.transform_1d(self.KIND, direction, self.cls, arr, level)
        except co.OracleError as e:
            raise JWaveFailure(f"oracle status {e.status}")

    def _forward1(self, arrTime, level=None):
        return self._run(co.FORWARD, arrTime, level)

    def _reverse1(self, arrHilb, level=None):
        return self._run(co.REVERSE, arrHilb, level)


class OracleFWT(_OracleTransform):
    KIND = co.FWT


class OracleWPT(_OracleTransform):
    KIND = co.WPT


def rng_signal(seed, *shape):
    return np.random.default_rng(seed).standard_normal(shape)
