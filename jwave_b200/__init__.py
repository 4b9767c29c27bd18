"""jwave_b200 - B200 (sm_100a) implementation of JWave's discrete-wavelet hot path.

The product is libjwave_cuda.so (csrc/, C ABI in include/jwave_cuda.h).  This package is the
host-side mirror of the reference's plug-in interface on top of it; see transforms.py."""
from .compressions import Compressor, CompressorMagnitude
from .exceptions import JWaveError, JWaveException, JWaveFailure
from .transforms import (AncientEgyptianDecomposition, BasicTransform, CudaContext, CudaFastWaveletTransform,
                         CudaWaveletPacketTransform, MathToolKit, Transform, TransformBuilder, WaveletTransform)
from .wavelets import WAVELET_CLASSES, Wavelet, WaveletBuilder

__all__ = [
    "AncientEgyptianDecomposition", "BasicTransform", "Compressor", "CompressorMagnitude", "CudaContext", "CudaFastWaveletTransform", "CudaWaveletPacketTransform",
    "JWaveError", "JWaveException", "JWaveFailure", "MathToolKit", "Transform", "TransformBuilder", "WaveletTransform",
    "WAVELET_CLASSES", "Wavelet", "WaveletBuilder",
]
