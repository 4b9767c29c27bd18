/*
 * CudaWaveletTransform - shared base of the GPU drop-ins for FastWaveletTransform and
 * WaveletPacketTransform.  Extends the reference's WaveletTransform (WaveletTransform.java:34), so
 * the default-level overloads come for free and `new Transform( new CudaFastWaveletTransform( wavelet ) )`
 * works unchanged.
 *
 * Overrides, besides the two abstract 1-D methods (BasicTransform.java:99, :113, :129, :151), the
 * 2-D and 3-D drivers (BasicTransform.java:361, :436, :509, :602): the reference runs 16 384 tiny
 * 1-D transforms per 8192 x 8192 matrix through them; here each is ONE native call.  Adds the batched
 * entry points (forwardBatch / reverseBatch, forwardBatch2D / reverseBatch2D), the any-length form
 * (AncientEgyptianDecomposition over the whole batch in one call) and decompose in one call.
 *
 * NOT COMPILED in the build container (no JDK); see INTEGRATION.md.
 */
package jwave.transforms;

import jwave.exceptions.JWaveException;
import jwave.exceptions.JWaveFailure;
import jwave.transforms.cuda.JWaveCuda;
import jwave.transforms.wavelets.Wavelet;

public abstract class CudaWaveletTransform extends WaveletTransform implements AutoCloseable {

  private final JWaveCuda _cuda;
  private final int _wid;
  private final int _kind;
  private final String _cls;

  /** devices: one CUDA ordinal, or several for one context over a group of GPUs (jwc_create_multi). */
  protected CudaWaveletTransform( Wavelet wavelet, int kind, String cls, int... devices ) throws JWaveException {
    super( wavelet );
    _kind = kind;
    _cls = cls;
    _cuda = new JWaveCuda( devices.length == 0 ? new int[ ]{ 0 } : devices );
    _wid = _cuda.setWavelet( wavelet );
  }

  /** The native context, e.g. for a CudaCompressorMagnitude that shares it. */
  public JWaveCuda context( ) {
    return _cuda;
  }

  // ---- 1-D ------------------------------------------------------------------------------------

  @Override public double[ ] forward( double[ ] arrTime, int level ) throws JWaveException {
    return _cuda.transform1D( _kind, _wid, JWaveCuda.FORWARD, new double[ ][ ]{ arrTime }, arrTime.length, level,
        _cls + "#forward" )[ 0 ];
  }

  @Override public double[ ] reverse( double[ ] arrHilb, int level ) throws JWaveException {
    return _cuda.transform1D( _kind, _wid, JWaveCuda.REVERSE, new double[ ][ ]{ arrHilb }, arrHilb.length, level,
        _cls + "#reverse" )[ 0 ];
  }

  /** WaveletTransform.java:136-146 in one native call: row p of the result is forward( arrTime, p ). */
  @Override public double[ ][ ] decompose( double[ ] arrTime ) throws JWaveException {
    if( !isBinary( arrTime.length ) )
      throw new JWaveFailure( _cls + "#decompose - given array length is not 2^p | p E N ... = 1, 2, 4, 8, 16, 32, .. " );
    return _cuda.decompose1D( _kind, _wid, arrTime, calcExponent( arrTime.length ) + 1, _cls + "#decompose" );
  }

  // ---- batched entry points: rows are independent signals of one length -----------------------
  // (not called forward(double[][]): that overload means a 2-D transform, BasicTransform.java:336)

  public double[ ][ ] forwardBatch( double[ ][ ] signals, int level ) throws JWaveException {
    return signals.length == 0 ? new double[ 0 ][ ]
        : _cuda.transform1D( _kind, _wid, JWaveCuda.FORWARD, signals, signals[ 0 ].length, level, _cls + "#forwardBatch" );
  }

  public double[ ][ ] reverseBatch( double[ ][ ] coefficients, int level ) throws JWaveException {
    return coefficients.length == 0 ? new double[ 0 ][ ]
        : _cuda.transform1D( _kind, _wid, JWaveCuda.REVERSE, coefficients, coefficients[ 0 ].length, level, _cls + "#reverseBatch" );
  }

  /** AncientEgyptianDecomposition.forward(double[]) (AncientEgyptianDecomposition.java:97-129) for every row: any
   *  common length, each 2^p block of the binary expansion transformed at full depth, one native call. */
  public double[ ][ ] forwardAnyLength( double[ ][ ] signals ) throws JWaveException {
    return signals.length == 0 ? new double[ 0 ][ ]
        : _cuda.ancientEgyptian( _kind, _wid, JWaveCuda.FORWARD, signals, signals[ 0 ].length, _cls + "#forwardAnyLength" );
  }

  public double[ ][ ] reverseAnyLength( double[ ][ ] coefficients ) throws JWaveException {
    return coefficients.length == 0 ? new double[ 0 ][ ]
        : _cuda.ancientEgyptian( _kind, _wid, JWaveCuda.REVERSE, coefficients, coefficients[ 0 ].length, _cls + "#reverseAnyLength" );
  }

  // ---- 2-D (BasicTransform.java:361-399, :436-474) -------------------------------------------

  @Override public double[ ][ ] forward( double[ ][ ] matTime, int lvlM, int lvlN ) throws JWaveException {
    return _cuda.transform2D( _kind, _wid, JWaveCuda.FORWARD, new double[ ][ ][ ]{ matTime }, matTime.length,
        matTime[ 0 ].length, lvlM, lvlN, _cls + "#forward" )[ 0 ];
  }

  @Override public double[ ][ ] reverse( double[ ][ ] matFreq, int lvlM, int lvlN ) throws JWaveException {
    return _cuda.transform2D( _kind, _wid, JWaveCuda.REVERSE, new double[ ][ ][ ]{ matFreq }, matFreq.length,
        matFreq[ 0 ].length, lvlM, lvlN, _cls + "#reverse" )[ 0 ];
  }

  /** A batch of equally shaped matrices in one native call (sharded over the GPUs of a device group). */
  public double[ ][ ][ ] forwardBatch2D( double[ ][ ][ ] mats, int lvlM, int lvlN ) throws JWaveException {
    return mats.length == 0 ? new double[ 0 ][ ][ ] : _cuda.transform2D( _kind, _wid, JWaveCuda.FORWARD, mats,
        mats[ 0 ].length, mats[ 0 ][ 0 ].length, lvlM, lvlN, _cls + "#forwardBatch2D" );
  }

  public double[ ][ ][ ] reverseBatch2D( double[ ][ ][ ] mats, int lvlM, int lvlN ) throws JWaveException {
    return mats.length == 0 ? new double[ 0 ][ ][ ] : _cuda.transform2D( _kind, _wid, JWaveCuda.REVERSE, mats,
        mats[ 0 ].length, mats[ 0 ][ 0 ].length, lvlM, lvlN, _cls + "#reverseBatch2D" );
  }

  // ---- 3-D (BasicTransform.java:509-566, :602-659; the native side keeps the level shift) ------

  @Override public double[ ][ ][ ] forward( double[ ][ ][ ] spcTime, int lvlP, int lvlQ, int lvlR ) throws JWaveException {
    return _cuda.transform3D( _kind, _wid, JWaveCuda.FORWARD, spcTime, lvlP, lvlQ, lvlR, _cls + "#forward" );
  }

  @Override public double[ ][ ][ ] reverse( double[ ][ ][ ] spcHilb, int lvlP, int lvlQ, int lvlR ) throws JWaveException {
    return _cuda.transform3D( _kind, _wid, JWaveCuda.REVERSE, spcHilb, lvlP, lvlQ, lvlR, _cls + "#reverse" );
  }

  /** Releases the native context, like ParallelWaveletPacketTransform.shutdown() releases its pool. */
  public void shutdown( ) {
    _cuda.close( );
  }

  @Override public void close( ) {
    shutdown( );
  }
}
