/*
 * CudaWaveletTransform - shared base of the GPU drop-ins for FastWaveletTransform and
 * WaveletPacketTransform.  Extends the reference's WaveletTransform (WaveletTransform.java:34), so
 * decompose / recompose and the default-level overloads come for free and
 * `new Transform( new CudaFastWaveletTransform( wavelet ) )` works unchanged.
 *
 * Overrides, besides the two abstract 1-D methods (BasicTransform.java:99, :113, :129, :151), the
 * 2-D and 3-D drivers (BasicTransform.java:361, :436, :509, :602): the reference runs 16 384 tiny
 * 1-D transforms per 8192 x 8192 matrix through them; here each is ONE native call.
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

  protected CudaWaveletTransform( Wavelet wavelet, int kind, String cls, int device ) throws JWaveException {
    super( wavelet );
    _kind = kind;
    _cls = cls;
    _cuda = new JWaveCuda( device );
    _wid = _cuda.setWavelet( wavelet );
  }

  // ---- 1-D ------------------------------------------------------------------------------------

  @Override public double[ ] forward( double[ ] arrTime, int level ) throws JWaveException {
    return _cuda.transform1D( _kind, _wid, JWaveCuda.FORWARD, arrTime, 1, arrTime.length, level, _cls + "#forward" );
  }

  @Override public double[ ] reverse( double[ ] arrHilb, int level ) throws JWaveException {
    return _cuda.transform1D( _kind, _wid, JWaveCuda.REVERSE, arrHilb, 1, arrHilb.length, level, _cls + "#reverse" );
  }

  /** WaveletTransform.java:136-146 in one native call: row p of the result is forward( arrTime, p ). */
  @Override public double[ ][ ] decompose( double[ ] arrTime ) throws JWaveException {
    if( !isBinary( arrTime.length ) )
      throw new JWaveFailure( _cls + "#decompose - given array length is not 2^p | p E N ... = 1, 2, 4, 8, 16, 32, .. " );
    int rows = calcExponent( arrTime.length ) + 1;
    double[ ] flat = _cuda.decompose1D( _kind, _wid, arrTime, rows, _cls + "#decompose" );
    return unflatten( flat, rows, arrTime.length );
  }

  // ---- batched entry point: rows are independent signals of one length ------------------------
  // (not called forward(double[][]): that overload means a 2-D transform, BasicTransform.java:336)

  public double[ ][ ] forwardBatch( double[ ][ ] signals, int level ) throws JWaveException {
    return batch( signals, level, JWaveCuda.FORWARD, "#forwardBatch" );
  }

  public double[ ][ ] reverseBatch( double[ ][ ] coefficients, int level ) throws JWaveException {
    return batch( coefficients, level, JWaveCuda.REVERSE, "#reverseBatch" );
  }

  private double[ ][ ] batch( double[ ][ ] rows, int level, int dir, String where ) throws JWaveException {
    if( rows.length == 0 )
      return new double[ 0 ][ ];
    int n = rows[ 0 ].length;
    double[ ] flat = flatten( rows, n, _cls + where );
    double[ ] out = _cuda.transform1D( _kind, _wid, dir, flat, rows.length, n, level, _cls + where );
    return unflatten( out, rows.length, n );
  }

  // ---- 2-D (BasicTransform.java:361-399, :436-474) -------------------------------------------

  @Override public double[ ][ ] forward( double[ ][ ] matTime, int lvlM, int lvlN ) throws JWaveException {
    int rows = matTime.length, cols = matTime[ 0 ].length;
    double[ ] out = _cuda.transform2D( _kind, _wid, JWaveCuda.FORWARD, flatten( matTime, cols, _cls + "#forward" ), 1,
        rows, cols, lvlM, lvlN, _cls + "#forward" );
    return unflatten( out, rows, cols );
  }

  @Override public double[ ][ ] reverse( double[ ][ ] matFreq, int lvlM, int lvlN ) throws JWaveException {
    int rows = matFreq.length, cols = matFreq[ 0 ].length;
    double[ ] out = _cuda.transform2D( _kind, _wid, JWaveCuda.REVERSE, flatten( matFreq, cols, _cls + "#reverse" ), 1,
        rows, cols, lvlM, lvlN, _cls + "#reverse" );
    return unflatten( out, rows, cols );
  }

  // ---- 3-D (BasicTransform.java:509-566, :602-659; the native side keeps the level shift) ------

  @Override public double[ ][ ][ ] forward( double[ ][ ][ ] spcTime, int lvlP, int lvlQ, int lvlR ) throws JWaveException {
    return space( spcTime, lvlP, lvlQ, lvlR, JWaveCuda.FORWARD, "#forward" );
  }

  @Override public double[ ][ ][ ] reverse( double[ ][ ][ ] spcHilb, int lvlP, int lvlQ, int lvlR ) throws JWaveException {
    return space( spcHilb, lvlP, lvlQ, lvlR, JWaveCuda.REVERSE, "#reverse" );
  }

  private double[ ][ ][ ] space( double[ ][ ][ ] s, int lvlP, int lvlQ, int lvlR, int dir, String where )
      throws JWaveException {
    int p = s.length, q = s[ 0 ].length, r = s[ 0 ][ 0 ].length;
    double[ ] flat = new double[ p * q * r ];
    for( int i = 0; i < p; i++ )
      for( int j = 0; j < q; j++ )
        System.arraycopy( s[ i ][ j ], 0, flat, ( i * q + j ) * r, r );
    double[ ] out = _cuda.transform3D( _kind, _wid, dir, flat, p, q, r, lvlP, lvlQ, lvlR, _cls + where );
    double[ ][ ][ ] res = new double[ p ][ q ][ r ];
    for( int i = 0; i < p; i++ )
      for( int j = 0; j < q; j++ )
        System.arraycopy( out, ( i * q + j ) * r, res[ i ][ j ], 0, r );
    return res;
  }

  // ---- helpers ---------------------------------------------------------------------------------

  private static double[ ] flatten( double[ ][ ] rows, int n, String where ) throws JWaveException {
    double[ ] flat = new double[ rows.length * n ];
    for( int i = 0; i < rows.length; i++ ) {
      if( rows[ i ].length != n )
        throw new JWaveFailure( where + " - all rows must have the same length" );
      System.arraycopy( rows[ i ], 0, flat, i * n, n );
    }
    return flat;
  }

  private static double[ ][ ] unflatten( double[ ] flat, int rows, int n ) {
    double[ ][ ] out = new double[ rows ][ n ];
    for( int i = 0; i < rows; i++ )
      System.arraycopy( flat, i * n, out[ i ], 0, n );
    return out;
  }

  /** Releases the native context, like ParallelWaveletPacketTransform.shutdown() releases its pool. */
  public void shutdown( ) {
    _cuda.close( );
  }

  @Override public void close( ) {
    shutdown( );
  }
}
