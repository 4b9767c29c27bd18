/*
 * CudaWaveletPacketTransform - drop-in for WaveletPacketTransform (WaveletPacketTransform.java:40-193)
 * and its Pooled / Parallel variants, whose arithmetic runs in libjwave_cuda.so on a B200.
 */
package jwave.transforms;

import jwave.exceptions.JWaveException;
import jwave.transforms.cuda.JWaveCuda;
import jwave.transforms.wavelets.Wavelet;

public class CudaWaveletPacketTransform extends CudaWaveletTransform {

  public CudaWaveletPacketTransform( Wavelet wavelet ) throws JWaveException {
    this( wavelet, 0 );
  }

  public CudaWaveletPacketTransform( Wavelet wavelet, int device ) throws JWaveException {
    super( wavelet, JWaveCuda.WPT, "WaveletPacketTransform", device );
    _name = "Wavelet Packet Transform"; // WaveletPacketTransform.java:53
  }
}
