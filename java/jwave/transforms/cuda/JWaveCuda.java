/*
 * JWaveCuda - java.lang.foreign (FFM) binding of libjwave_cuda.so (include/jwave_cuda.h).
 *
 * JDK 21: java.lang.foreign is a preview API there (final in JDK 22), so compile and run with
 * --enable-preview (pom.xml targets release 21).  NOT COMPILED in the build container: no JDK is
 * installed there or on the GPU box (DESIGN.md); the C ABI it binds is exercised by the ctypes
 * host layer (jwave_b200/) instead.
 *
 * One instance = one jwc_ctx = one GPU + one stream.  Calls on an instance are synchronised.
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
  private static final java.lang.foreign.AddressLayout P = ValueLayout.ADDRESS;

  private static MethodHandle fn( String name, FunctionDescriptor d ) {
    return LINKER.downcallHandle( LIB.find( name ).orElseThrow( ), d );
  }

  // int jwc_create(jwc_ctx** out, int device); int jwc_destroy(jwc_ctx*); const char* jwc_last_error(const jwc_ctx*)
  private static final MethodHandle CREATE = fn( "jwc_create", FunctionDescriptor.of( I, P, I ) );
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

  private final MemorySegment ctx;
  private boolean closed;

  public JWaveCuda( int device ) throws JWaveException {
    try( Arena a = Arena.ofConfined( ) ) {
      MemorySegment out = a.allocate( P );
      int st = (int) CREATE.invokeExact( out, device );
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
      MemorySegment sDe = a.allocateArray( ValueLayout.JAVA_DOUBLE, w.getScalingDeComposition( ) );
      MemorySegment wDe = a.allocateArray( ValueLayout.JAVA_DOUBLE, w.getWaveletDeComposition( ) );
      MemorySegment sRe = a.allocateArray( ValueLayout.JAVA_DOUBLE, w.getScalingReConstruction( ) );
      MemorySegment wRe = a.allocateArray( ValueLayout.JAVA_DOUBLE, w.getWaveletReConstruction( ) );
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

  /** jwc_decompose1d: one signal of length n -> (log2 n + 1) x n, row p = forward(x, p). */
  private static final MethodHandle DECOMPOSE =
      fn( "jwc_decompose1d", FunctionDescriptor.of( I, P, I, I, P, P, ValueLayout.JAVA_LONG, I ) );

  public synchronized double[ ] decompose1D( int kind, int wid, double[ ] x, int rows, String where ) throws JWaveException {
    try( Arena a = Arena.ofConfined( ) ) {
      MemorySegment in = a.allocateArray( ValueLayout.JAVA_DOUBLE, x );
      MemorySegment out = a.allocateArray( ValueLayout.JAVA_DOUBLE, (long) rows * x.length );
      int st = (int) DECOMPOSE.invokeExact( ctx, wid, kind, in, out, 1L, x.length );
      check( st, where );
      return out.toArray( ValueLayout.JAVA_DOUBLE );
    } catch( JWaveException e ) {
      throw e;
    } catch( Throwable t ) {
      throw new JWaveError( where + ": " + t );
    }
  }

  /** batch x n signals, flattened row-major; returns a fresh array (the input is never modified). */
  public synchronized double[ ] transform1D( int kind, int wid, int dir, double[ ] flat, long batch, int n, int level,
      String where ) throws JWaveException {
    try( Arena a = Arena.ofConfined( ) ) {
      MemorySegment in = a.allocateArray( ValueLayout.JAVA_DOUBLE, flat ); // heap double[] cannot cross the FFM boundary
      MemorySegment out = a.allocateArray( ValueLayout.JAVA_DOUBLE, (long) flat.length );
      int st = (int) T1D[ kind ].invokeExact( ctx, wid, dir, in, out, batch, n, level );
      check( st, where );
      return out.toArray( ValueLayout.JAVA_DOUBLE );
    } catch( JWaveException e ) {
      throw e;
    } catch( Throwable t ) {
      throw new JWaveError( where + ": " + t );
    }
  }

  public synchronized double[ ] transform2D( int kind, int wid, int dir, double[ ] flat, long batch, int rows, int cols,
      int lvlM, int lvlN, String where ) throws JWaveException {
    try( Arena a = Arena.ofConfined( ) ) {
      MemorySegment in = a.allocateArray( ValueLayout.JAVA_DOUBLE, flat );
      MemorySegment out = a.allocateArray( ValueLayout.JAVA_DOUBLE, (long) flat.length );
      int st = (int) T2D[ kind ].invokeExact( ctx, wid, dir, in, out, batch, rows, cols, lvlM, lvlN );
      check( st, where );
      return out.toArray( ValueLayout.JAVA_DOUBLE );
    } catch( JWaveException e ) {
      throw e;
    } catch( Throwable t ) {
      throw new JWaveError( where + ": " + t );
    }
  }

  public synchronized double[ ] transform3D( int kind, int wid, int dir, double[ ] flat, int p, int q, int r, int lvlP,
      int lvlQ, int lvlR, String where ) throws JWaveException {
    try( Arena a = Arena.ofConfined( ) ) {
      MemorySegment in = a.allocateArray( ValueLayout.JAVA_DOUBLE, flat );
      MemorySegment out = a.allocateArray( ValueLayout.JAVA_DOUBLE, (long) flat.length );
      int st = (int) T3D[ kind ].invokeExact( ctx, wid, dir, in, out, p, q, r, lvlP, lvlQ, lvlR );
      check( st, where );
      return out.toArray( ValueLayout.JAVA_DOUBLE );
    } catch( JWaveException e ) {
      throw e;
    } catch( Throwable t ) {
      throw new JWaveError( where + ": " + t );
    }
  }

  @Override public synchronized void close( ) {
    if( !closed ) {
      closed = true;
      try {
        int ignored = (int) DESTROY.invokeExact( ctx );
      } catch( Throwable t ) {
        // nothing sensible to do at shutdown
      }
    }
  }
}
