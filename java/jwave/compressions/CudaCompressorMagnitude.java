/*
 * CudaCompressorMagnitude - drop-in for CompressorMagnitude (compressions/CompressorMagnitude.java:36-118):
 * magnitude = mean |c| over the whole array, coefficients below magnitude * threshold become zero
 * (Compressor.java:97-110), computed by libjwave_cuda.so (jwc_compress_magnitude).
 *
 * NOT COMPILED in the build container (no JDK); see INTEGRATION.md.
 */
package jwave.compressions;

import jwave.exceptions.JWaveException;
import jwave.transforms.cuda.JWaveCuda;

public class CudaCompressorMagnitude extends Compressor {

  private final JWaveCuda _cuda;

  /** Shares the native context of a transform: new CudaCompressorMagnitude( fwt.context( ), 1.5 ). */
  public CudaCompressorMagnitude( JWaveCuda cuda, double threshold ) {
    super( threshold );
    _cuda = cuda;
  }

  @Override public double[ ] compress( double[ ] arrHilb ) {
    return compress( new double[ ][ ]{ arrHilb } )[ 0 ];
  }

  @Override public double[ ][ ] compress( double[ ][ ] matHilb ) {
    try {
      double[ ] mag = new double[ 1 ];
      double[ ][ ] out = _cuda.compressMagnitude( matHilb, matHilb[ 0 ].length, _threshold, mag, "CompressorMagnitude#compress" );
      _magnitude = mag[ 0 ];
      return out;
    } catch( JWaveException e ) { // Compressor.java reports and carries on
      e.showMessage( );
      return null;
    }
  }

  @Override public double[ ][ ][ ] compress( double[ ][ ][ ] spcHilb ) {
    int p = spcHilb.length, q = spcHilb[ 0 ].length;
    double[ ][ ] flat = new double[ p * q ][ ]; // the mean runs over the whole volume: one call over all rows
    for( int i = 0; i < p; i++ )
      System.arraycopy( spcHilb[ i ], 0, flat, i * q, q );
    double[ ][ ] out = compress( flat );
    if( out == null )
      return null;
    double[ ][ ][ ] res = new double[ p ][ q ][ ];
    for( int i = 0; i < p; i++ )
      System.arraycopy( out, i * q, res[ i ], 0, q );
    return res;
  }
}
