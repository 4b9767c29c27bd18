/*
 * CudaFastWaveletTransform - drop-in for FastWaveletTransform (FastWaveletTransform.java:38-154)
 * whose arithmetic runs in libjwave_cuda.so on a B200.  Usage is the reference's:
 *
 *   Transform t = new Transform( new CudaFastWaveletTransform( new Daubechies4( ) ) );
 *   double[ ] hilb = t.forward( arrTime );       // Transform.java:81
 *   double[ ] reco = t.reverse( hilb );
 *
 * new CudaFastWaveletTransform( wavelet, 0, 1, 2, 3, 4, 5, 6, 7 ) puts all eight GPUs of a box behind the one object.
 */
package jwave.transforms;

import jwave.exceptions.JWaveException;
import jwave.transforms.cuda.JWaveCuda;
import jwave.transforms.wavelets.Wavelet;

public class CudaFastWaveletTransform extends CudaWaveletTransform {

  public CudaFastWaveletTransform( Wavelet wavelet ) throws JWaveException {
    this( wavelet, 0 );
  }

  public CudaFastWaveletTransform( Wavelet wavelet, int... devices ) throws JWaveException {
    super( wavelet, JWaveCuda.FWT, "FastWaveletTransform", devices );
    _name = "Fast Wavelet Transform";
  }
}
