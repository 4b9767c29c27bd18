/*
 * JWaveCuda - java.lang.foreign (FFM) binding of libjwave_cuda.so (include/jwave_cuda.h).
 *
 * JDK 21: java.lang.foreign is a preview API there (final in JDK 22), so compile and run with
 * --enable-preview (pom.xml targets release 21); on JDK 22+ rename allocateArray -> allocateFrom / allocate and
 * getUtf8String -> getString.  NOT COMPILED in the build container: no JDK is installed there or on the GPU box
 * (DESIGN.md); the C ABI it binds is exercised by the ctypes host layer (jwave_b200/) instead.
 *
 * One instance = one jwc_ctx: one GPU (JWaveCuda(int)) or a group of GPUs of the box behind one handle
 * (JWaveCuda(int[]), jwc_create_multi: batched calls are sharded over them, a 3-D volume is slab-decomposed).
 * The native context serialises the calls made on it; the methods here are synchronized as well because they
 * share the staging buffers.
 *
 * Staging: Java heap arrays cannot cross the FFM boundary, so every call copies its rows into a PINNED host
 * buffer owned by this object (jwc_host_alloc_pinned, grown on demand) - row by row, with long offsets, never
 * through one flat double[] (which overflows int past 2^31 elements) - and the library's H2D / D2H pipeline
 * runs at the bus rate the benchmark reports (pageable Arena memory would halve it).
 */
package jwave.transforms.cuda;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.foreign.ValueLayout;
import java.lang.invoke.MethodHandle;

import jwave.exceptions.JWaveError;
import jwave.exceptions.JWaveException;
import jwave.exceptions.JWaveFailure;
import jwave.transforms.wavelets.Wavelet;

public final class JWaveCuda implements AutoCloseable {

  public static final int FORWARD = 0, REVERSE = 1;
  public static final int FWT = 0, WPT = 1;

  private static final Linker LINKER = Linker.nativeLinker( );
  private static final SymbolLookup LIB =
      SymbolLookup.libraryLookup( System.getProperty( "jwave.cuda.library", "libjwave_cuda.so" ), Arena.global( ) );

  private static final ValueLayout.OfInt I = ValueLayout.JAVA_INT;
  private static final ValueLayout.OfLong J = ValueLayout.JAVA_LONG;
  private static final ValueLayout.OfDouble D = ValueLayout.JAVA_DOUBLE;
  private static final java.lang.foreign.AddressLayout P = ValueLayout.ADDRESS;

  private static MethodHandle fn( String name, FunctionDescriptor d ) {
    return LINKER.downcallHandle( LIB.find( name ).orElseThrow( ), d );
  }

  // int jwc_create(jwc_ctx**, int device); int jwc_create_multi(jwc_ctx**, const int* devices, int ndev)
  private static final MethodHandle CREATE = fn( "jwc_create", FunctionDescriptor.of( I, P, I ) );
  private static final MethodHandle CREATE_MULTI = fn( "jwc_create_multi", FunctionDescriptor.of( I, P, P, I ) );
  private static final MethodHandle DESTROY = fn( "jwc_destroy", FunctionDescriptor.of( I, P ) );
  private static final MethodHandle LAST_ERROR = fn( "jwc_last_error", FunctionDescriptor.of( P, P ) );
  // int jwc_set_wavelet(ctx, int L, const double* sDe, wDe, sRe, wRe, int* wid)
  private static final MethodHandle SET_WAVELET = fn( "jwc_set_wavelet", FunctionDescriptor.of( I, P, I, P, P, P, P, P ) );
  // int jwc_{fwt,wpt}1d(ctx, wid, dir, in, out, int64 batch, int n, int level)
  private static final FunctionDescriptor D1 = FunctionDescriptor.of( I, P, I, I, P, P, J, I, I );
  private static final MethodHandle[ ] T1D = { fn( "jwc_fwt1d", D1 ), fn( "jwc_wpt1d", D1 ) };
  // int jwc_{fwt,wpt}2d(ctx, wid, dir, in, out, int64 batch, int rows, int cols, int lvlM, int lvlN)
  private static final FunctionDescriptor D2 = FunctionDescriptor.of( I, P, I, I, P, P, J, I, I, I, I );
  private static final MethodHandle[ ] T2D = { fn( "jwc_fwt2d", D2 ), fn( "jwc_wpt2d", D2 ) };
  // int jwc_{fwt,wpt}3d(ctx, wid, dir, in, out, int P, int Q, int R, int lvlP, int lvlQ, int lvlR)
  private static final FunctionDescriptor D3 = FunctionDescriptor.of( I, P, I, I, P, P, I, I, I, I, I, I );
  private static final MethodHandle[ ] T3D = { fn( "jwc_fwt3d", D3 ), fn( "jwc_wpt3d", D3 ) };
  // int jwc_aed1d(ctx, wid, kind, dir, in, out, int64 batch, int n)
  private static final MethodHandle AED = fn( "jwc_aed1d", FunctionDescriptor.of( I, P, I, I, I, P, P, J, I ) );
  // int jwc_decompose1d(ctx, wid, kind, in, out, int64 batch, int n)
  private static final MethodHandle DECOMPOSE = fn( "jwc_decompose1d", FunctionDescriptor.of( I, P, I, I, P, P, J, I ) );
  // int jwc_compress_magnitude(ctx, in, out, int64 count, double threshold, double* magnitude)
  private static final MethodHandle COMPRESS = fn( "jwc_compress_magnitude", FunctionDescriptor.of( I, P, P, P, J, D, P ) );
  // int jwc_host_alloc_pinned(ctx, size_t bytes, void** hptr); int jwc_host_free_pinned(ctx, void* hptr)
  private static final MethodHandle PIN_ALLOC = fn( "jwc_host_alloc_pinned", FunctionDescriptor.of( I, P, J, P ) );
  private static final MethodHandle PIN_FREE = fn( "jwc_host_free_pinned", FunctionDescriptor.of( I, P, P ) );

  private final MemorySegment ctx;
  private MemorySegment pinIn = MemorySegment.NULL, pinOut = MemorySegment.NULL;
  private boolean closed;

  public JWaveCuda( int device ) throws JWaveException {
    this( new int[ ]{ device } );
  }

  /** One context over several GPUs of the box (jwc_create_multi); a single ordinal gives jwc_create. */
  public JWaveCuda( int[ ] devices ) throws JWaveException {
    try( Arena a = Arena.ofConfined( ) ) {
      MemorySegment out = a.allocate( P );
      int st = devices.length == 1 ? (int) CREATE.invokeExact( out, devices[ 0 ] )
          : (int) CREATE_MULTI.invokeExact( out, a.allocateArray( I, devices ), devices.length );
      if( st != 0 )
        throw new JWaveError( "jwc_create failed: " + text( (MemorySegment) LAST_ERROR.invokeExact( MemorySegment.NULL ) ) );
      ctx = out.get( P, 0 );
    } catch( JWaveException e ) {
      throw e;
    } catch( Throwable t ) {
      throw new JWaveError( "jwc_create: " + t );
    }
  }

  private static String text( MemorySegment cstr ) {
    return cstr.equals( MemorySegment.NULL ) ? "" : cstr.reinterpret( 4096 ).getUtf8String( 0 );
  }

  /** Status codes of include/jwave_cuda.h mapped onto the reference's exception classes. */
  private void check( int st, String where ) throws JWaveException {
    if( st == 0 )
      return;
    String msg;
    try {
      msg = text( (MemorySegment) LAST_ERROR.invokeExact( ctx ) );
    } catch( Throwable t ) {
      msg = "status " + st;
    }
    switch( st ){
      case 1: // FastWaveletTransform.java:74-78
        throw new JWaveFailure( where + " - given array length is not 2^p | p E N ... = 1, 2, 4, 8, 16, 32, .. "
            + "please use the Ancient Egyptian Decomposition for any other array length!" );
      case 2: // FastWaveletTransform.java:81-83
        throw new JWaveFailure( where + " - given level is out of range for given array" );
      case 3:
        throw new JWaveFailure( where + " - " + msg );
      default:
        throw new JWaveError( where + " - " + msg );
    }
  }

  /** Registers the four filters exactly as the Wavelet object hands them out (Wavelet.java:178-219). */
  public synchronized int setWavelet( Wavelet w ) throws JWaveException {
    try( Arena a = Arena.ofConfined( ) ) {
      MemorySegment sDe = a.allocateArray( D, w.getScalingDeComposition( ) );
      MemorySegment wDe = a.allocateArray( D, w.getWaveletDeComposition( ) );
      MemorySegment sRe = a.allocateArray( D, w.getScalingReConstruction( ) );
      MemorySegment wRe = a.allocateArray( D, w.getWaveletReConstruction( ) );
      MemorySegment wid = a.allocate( I );
      int st = (int) SET_WAVELET.invokeExact( ctx, w.getMotherWavelength( ), sDe, wDe, sRe, wRe, wid );
      check( st, "jwc_set_wavelet" );
      return wid.get( I, 0 );
    } catch( JWaveException e ) {
      throw e;
    } catch( Throwable t ) {
      throw new JWaveError( "jwc_set_wavelet: " + t );
    }
  }

  // ---- pinned staging --------------------------------------------------------------------------

  private MemorySegment pinned( MemorySegment cur, long bytes ) throws Throwable {
    if( cur.byteSize( ) >= bytes )
      return cur;
    if( !cur.equals( MemorySegment.NULL ) ) {
      int ignored = (int) PIN_FREE.invokeExact( ctx, cur );
    }
    try( Arena a = Arena.ofConfined( ) ) {
      MemorySegment slot = a.allocate( P );
      int st = (int) PIN_ALLOC.invokeExact( ctx, bytes, slot );
      if( st != 0 )
        throw new JWaveError( "jwc_host_alloc_pinned: " + text( (MemorySegment) LAST_ERROR.invokeExact( ctx ) ) );
      return slot.get( P, 0 ).reinterpret( bytes );
    }
  }

  /** rows x n doubles -> pinIn, row by row (offsets are longs: no 2^31-element limit) */
  private void stage( double[ ][ ] rows, int n, long rowOffset, String where ) throws JWaveException {
    for( int i = 0; i < rows.length; i++ ) {
      if( rows[ i ].length != n )
        throw new JWaveFailure( where + " - all rows must have the same length" );
      MemorySegment.copy( rows[ i ], 0, pinIn, D, ( rowOffset + i ) * (long) n * Double.BYTES, n );
    }
  }

  private static double[ ][ ] unstage( MemorySegment seg, long rowOffset, int rows, int n ) {
    double[ ][ ] out = new double[ rows ][ n ]; // fresh arrays: the reference never mutates its input
    for( int i = 0; i < rows; i++ )
      MemorySegment.copy( seg, D, ( rowOffset + i ) * (long) n * Double.BYTES, out[ i ], 0, n );
    return out;
  }

  @FunctionalInterface private interface Native {
    int call( MemorySegment in, MemorySegment out ) throws Throwable;
  }

  /** `elems` doubles in and `outElems` out through the pinned buffers; `fill` stages the input */
  private void run( long elems, long outElems, String where, Runnable2 fill, Native f ) throws JWaveException {
    try {
      pinIn = pinned( pinIn, elems * Double.BYTES );
      pinOut = pinned( pinOut, outElems * Double.BYTES );
      fill.run( );
      check( f.call( pinIn, pinOut ), where );
    } catch( JWaveException e ) {
      throw e;
    } catch( Throwable t ) {
      throw new JWaveError( where + ": " + t );
    }
  }

  @FunctionalInterface private interface Runnable2 {
    void run( ) throws JWaveException;
  }

  // ---- transforms ------------------------------------------------------------------------------

  /** FastWaveletTransform / WaveletPacketTransform.forward|reverse(double[], level) over rows.length signals. */
  public synchronized double[ ][ ] transform1D( int kind, int wid, int dir, double[ ][ ] rows, int n, int level, String where )
      throws JWaveException {
    long elems = (long) rows.length * n;
    run( elems, elems, where, ( ) -> stage( rows, n, 0, where ),
        ( in, out ) -> (int) T1D[ kind ].invokeExact( ctx, wid, dir, in, out, (long) rows.length, n, level ) );
    return unstage( pinOut, 0, rows.length, n );
  }

  /** BasicTransform.forward|reverse(double[][], lvlM, lvlN) over mats.length matrices in one native call. */
  public synchronized double[ ][ ][ ] transform2D( int kind, int wid, int dir, double[ ][ ][ ] mats, int rows, int cols,
      int lvlM, int lvlN, String where ) throws JWaveException {
    long elems = (long) mats.length * rows * cols;
    run( elems, elems, where, ( ) -> {
      for( int b = 0; b < mats.length; b++ ) {
        if( mats[ b ].length != rows )
          throw new JWaveFailure( where + " - all matrices must have the same shape" );
        stage( mats[ b ], cols, (long) b * rows, where );
      }
    }, ( in, out ) -> (int) T2D[ kind ].invokeExact( ctx, wid, dir, in, out, (long) mats.length, rows, cols, lvlM, lvlN ) );
    double[ ][ ][ ] res = new double[ mats.length ][ ][ ];
    for( int b = 0; b < mats.length; b++ )
      res[ b ] = unstage( pinOut, (long) b * rows, rows, cols );
    return res;
  }

  /** BasicTransform.forward|reverse(double[][][], lvlP, lvlQ, lvlR); the native side keeps the level shift. */
  public synchronized double[ ][ ][ ] transform3D( int kind, int wid, int dir, double[ ][ ][ ] s, int lvlP, int lvlQ,
      int lvlR, String where ) throws JWaveException {
    int p = s.length, q = s[ 0 ].length, r = s[ 0 ][ 0 ].length;
    long elems = (long) p * q * r;
    run( elems, elems, where, ( ) -> {
      for( int i = 0; i < p; i++ )
        stage( s[ i ], r, (long) i * q, where );
    }, ( in, out ) -> (int) T3D[ kind ].invokeExact( ctx, wid, dir, in, out, p, q, r, lvlP, lvlQ, lvlR ) );
    double[ ][ ][ ] res = new double[ p ][ ][ ];
    for( int i = 0; i < p; i++ )
      res[ i ] = unstage( pinOut, (long) i * q, q, r );
    return res;
  }

  /** AncientEgyptianDecomposition.forward|reverse(double[]) around an FWT / WPT: rows of ANY common length n. */
  public synchronized double[ ][ ] ancientEgyptian( int kind, int wid, int dir, double[ ][ ] rows, int n, String where )
      throws JWaveException {
    long elems = (long) rows.length * n;
    run( elems, elems, where, ( ) -> stage( rows, n, 0, where ),
        ( in, out ) -> (int) AED.invokeExact( ctx, wid, kind, dir, in, out, (long) rows.length, n ) );
    return unstage( pinOut, 0, rows.length, n );
  }

  /** WaveletTransform.decompose: one signal of length n -> (log2 n + 1) rows, row p = forward(x, p). */
  public synchronized double[ ][ ] decompose1D( int kind, int wid, double[ ] x, int levels, String where ) throws JWaveException {
    run( x.length, (long) levels * x.length, where, ( ) -> MemorySegment.copy( x, 0, pinIn, D, 0, x.length ),
        ( in, out ) -> (int) DECOMPOSE.invokeExact( ctx, wid, kind, in, out, 1L, x.length ) );
    return unstage( pinOut, 0, levels, x.length );
  }

  /** CompressorMagnitude.compress on a flat view of the coefficients; magnitude[0] receives the mean |c|. */
  public synchronized double[ ][ ] compressMagnitude( double[ ][ ] rows, int n, double threshold, double[ ] magnitude,
      String where ) throws JWaveException {
    long elems = (long) rows.length * n;
    try( Arena a = Arena.ofConfined( ) ) {
      MemorySegment mag = a.allocate( D );
      run( elems, elems, where, ( ) -> stage( rows, n, 0, where ),
          ( in, out ) -> (int) COMPRESS.invokeExact( ctx, in, out, elems, threshold, mag ) );
      magnitude[ 0 ] = mag.get( D, 0 );
    }
    return unstage( pinOut, 0, rows.length, n );
  }

  @Override public synchronized void close( ) {
    if( !closed ) {
      closed = true;
      try {
        if( !pinIn.equals( MemorySegment.NULL ) ) {
          int a = (int) PIN_FREE.invokeExact( ctx, pinIn );
        }
        if( !pinOut.equals( MemorySegment.NULL ) ) {
          int b = (int) PIN_FREE.invokeExact( ctx, pinOut );
        }
        int ignored = (int) DESTROY.invokeExact( ctx );
      } catch( Throwable t ) {
        // nothing sensible to do at shutdown
      }
    }
  }
}
