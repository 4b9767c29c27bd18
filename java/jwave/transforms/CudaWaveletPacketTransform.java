/*
 * CudaWaveletPacketTransform - drop-in for WaveletPacketTransform (WaveletPacketTransform.java:40-193; same arithmetic as the Pooled / Parallel variants)
 * whose arithmetic runs in libjwave_cuda.so on a B200.  Usage is the reference's:
 *
 *   Transform t = new Transform( new CudaWaveletPacketTransform( new Daubechies4( ) ) );
 *   double[ ] hilb = t.forward( arrTime );       // Transform.java:81
 *   double[ ] reco = t.reverse( hilb );
 *
 * new CudaWaveletPacketTransform( wavelet, 0, 1, 2, 3, 4, 5, 6, 7 ) puts all eight GPUs of a box behind the one object.
 */
package jwave.transforms;

import jwave.exceptions.JWaveException;
import jwave.transforms.cuda.JWaveCuda;
import jwave.transforms.wavelets.Wavelet;

public class CudaWaveletPacketTransform extends CudaWaveletTransform {

  public CudaWaveletPacketTransform( Wavelet wavelet ) throws JWaveException {
    this( wavelet, 0 );
  }

  public CudaWaveletPacketTransform( Wavelet wavelet, int... devices ) throws JWaveException {
    super( wavelet, JWaveCuda.WPT, "WaveletPacketTransform", devices );
    _name = "Wavelet Packet Transform";
  }
}
