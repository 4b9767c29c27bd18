/*
 * jwave_cuda.h - C ABI of libjwave_cuda.so, the B200 (sm_100a) implementation of JWave's
 * discrete-wavelet hot path.
 *
 * This is the drop-in boundary: the entry points below are exactly what a JWave-side FFI
 * binding (Java 21 java.lang.foreign, see INTEGRATION.md and java/) needs in order to put the
 * GPU behind `BasicTransform` subclasses while `Transform.forward/reverse` stays unchanged.
 * Plain pointers and sizes only; no C++ or torch types.  Every function that returns `int`
 * returns a jwc_status (0 = OK).
 *
 * Reference interfaces replaced (paths relative to /root/reference/src/main/java/jwave/):
 *   Wavelet.forward / Wavelet.reverse              transforms/wavelets/Wavelet.java:236-260, :277-303
 *   FastWaveletTransform.forward / reverse         transforms/FastWaveletTransform.java:71-101, :119-153
 *   WaveletPacketTransform.forward / reverse       transforms/WaveletPacketTransform.java:73-124, :141-191
 *   (Pooled|Parallel)WaveletPacketTransform        transforms/PooledWaveletPacketTransform.java:24-127,
 *                                                  transforms/ParallelWaveletPacketTransform.java:79-146
 *   BasicTransform 2-D forward / reverse           transforms/BasicTransform.java:361-399, :436-474
 *   BasicTransform 3-D forward / reverse           transforms/BasicTransform.java:509-566, :602-659
 *
 * Data layout: every array is dense, row-major, contiguous IEEE binary64.  A batch of 1-D
 * signals is [batch][n]; a batch of matrices is [batch][rows][cols]; a volume is [P][Q][R]
 * indexed [i][j][k] like Java's double[P][Q][R].
 *
 * Semantics are the reference's: periodic extension to the right, outputs laid out as
 * [a_l | d_l | ... | d_1] (FWT) or as 2^l packets in natural order (WPT); inputs are never
 * modified; `level` counts decomposition steps, 0 <= level <= log2(n).
 *
 * There is no CPU fallback: every entry point either runs the CUDA kernels or fails.
 */
#ifndef JWAVE_CUDA_H
#define JWAVE_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JWC_VERSION 100 /* 0.1.0 */
#define JWC_MAX_TAPS 40 /* Daubechies20 / Symlet20 */

typedef struct jwc_ctx jwc_ctx;

typedef enum jwc_status {
  JWC_OK = 0,
  JWC_ERR_NOT_BINARY = 1, /* a length is not 2^p      -> JWaveFailure (FastWaveletTransform.java:74-78) */
  JWC_ERR_LEVEL = 2,      /* level outside [0, log2 n] -> JWaveFailure (FastWaveletTransform.java:81-83) */
  JWC_ERR_ARG = 3,        /* null pointer, bad handle, odd filter length, aliasing, ... */
  JWC_ERR_CUDA = 4,       /* a CUDA runtime call failed; see jwc_last_error */
  JWC_ERR_NCCL = 5        /* reserved (the exchange step of the slab-decomposed 3-D path uses peer copies, not NCCL) */
} jwc_status;

enum { JWC_FORWARD = 0, JWC_REVERSE = 1 };
enum { JWC_FWT = 0, JWC_WPT = 1 };

int jwc_version(void);

/* One context = one GPU, one set of scratch buffers and one current stream.  `device` is a CUDA ordinal.
 * Every entry point that takes a context holds the context's lock for its duration, so concurrent callers are
 * serialised (the reference's transforms are stateless and thread-safe, BasicTransform.java:42; host threads
 * that want to overlap use one context each).  The device-resident entry points only ENQUEUE work: the rule for
 * them is one stream at a time per context - jwc_set_stream makes the new stream wait (on the device) for the
 * work enqueued under the previous one, because both use the context's scratch buffers. */
int jwc_create(jwc_ctx** out, int device);
/* One context that drives SEVERAL GPUs of the box (1 <= ndev <= 8 distinct CUDA ordinals, peer access between all
 * pairs): the form SURVEY.md section 8(b) specifies for a single JVM in front of 8 GPUs.  The handle behaves like
 * a context on devices[0] for every device-resident (*_dev) entry point; the host-buffer entry points use all
 * devices:
 *   jwc_fwt1d / jwc_wpt1d / jwc_fwt2d / jwc_wpt2d   the batch is cut into contiguous blocks, one per GPU, each
 *                                                   through its own staging pipeline (no exchange);
 *   jwc_fwt3d / jwc_wpt3d                           the volume is slab-decomposed along i when P and Q are
 *       multiples of ndev (BasicTransform.forward(double[][][]), BasicTransform.java:509-566): k and j passes on
 *       the owned slices, one re-cut over NVLink on the copy engines, the i pass, and the download scatters the
 *       result into the reference's layout.  The reverse rebuilds axis i first, as ParallelTransform.reverse does
 *       (ParallelTransform.java:193) - results differ from the single-device order at rounding level only.
 * jwc_set_wavelet registers the filters on every device; jwc_destroy releases all of them. */
int jwc_create_multi(jwc_ctx** out, const int* devices, int ndev);
int jwc_device_count(const jwc_ctx* ctx);
int jwc_destroy(jwc_ctx* ctx);
/* Text of the last failure on this context (never NULL). With ctx == NULL: creation failures. */
const char* jwc_last_error(const jwc_ctx* ctx);

/* Launch on a caller-owned cudaStream_t (e.g. torch's current stream) instead of the context's own.  The handle
 * is used as given: NULL is CUDA's legacy default stream.  jwc_reset_stream goes back to the context's own
 * stream.  The host-buffer entry points (jwc_fwt1d ... jwc_decompose1d, jwc_compress_magnitude) always run on the
 * context's own streams, after waiting for the caller's stream, and are synchronous on return. */
int jwc_set_stream(jwc_ctx* ctx, void* cuda_stream);
int jwc_reset_stream(jwc_ctx* ctx);
int jwc_sync(jwc_ctx* ctx);
/* Number of kernels this context has launched so far. */
int64_t jwc_launch_count(const jwc_ctx* ctx);

/* Per-launch timing for benchmarks: while enabled, every kernel launch is bracketed by a CUDA
 * event pair on the launching stream.  jwc_profile_report waits for the recorded launches and
 * writes one text line per kernel, "label,launches,total_ms,samples_per_launch,levels", then clears the
 * records.  Off by default (no events are created). */
int jwc_profile_enable(jwc_ctx* ctx, int on);
int jwc_profile_report(jwc_ctx* ctx, char* buf, size_t size);

/* Register a wavelet from the four arrays returned by the reference's getters
 * (Wavelet.getScalingDeComposition() ... getWaveletReConstruction(), Wavelet.java:178-219), so
 * the device filters are bit-identical to the JVM's.  L must be even, 2 <= L <= JWC_MAX_TAPS.
 * The taps travel to the kernels as a __grid_constant__ parameter, i.e. in constant memory.
 * On success *wid is a handle valid until the context is destroyed. */
int jwc_set_wavelet(jwc_ctx* ctx, int L, const double* scalingDeCom, const double* waveletDeCom,
                    const double* scalingReCon, const double* waveletReCon, int* wid);

/* ---- host-buffer entry points: H2D copy, kernels, D2H copy, synchronous on return ---------- */

/* FastWaveletTransform.forward/reverse(double[], int) over `batch` signals of length n. */
int jwc_fwt1d(jwc_ctx* ctx, int wid, int dir, const double* in, double* out, int64_t batch, int n,
              int level);
/* WaveletPacketTransform.forward/reverse(double[], int) over `batch` signals of length n. */
int jwc_wpt1d(jwc_ctx* ctx, int wid, int dir, const double* in, double* out, int64_t batch, int n,
              int level);
/* BasicTransform.forward/reverse(double[][], lvlM, lvlN) with an FWT / a WPT as the 1-D step:
 * forward = every row with lvlN then every column with lvlM; reverse = columns then rows. */
int jwc_fwt2d(jwc_ctx* ctx, int wid, int dir, const double* in, double* out, int64_t batch, int rows,
              int cols, int lvlM, int lvlN);
int jwc_wpt2d(jwc_ctx* ctx, int wid, int dir, const double* in, double* out, int64_t batch, int rows,
              int cols, int lvlM, int lvlN);
/* BasicTransform.forward/reverse(double[][][], lvlP, lvlQ, lvlR), including the reference's
 * level shift: axis k (length R) gets lvlQ, axis j (length Q) gets lvlP, axis i (length P)
 * gets lvlR (BasicTransform.java:532, :555). */
int jwc_fwt3d(jwc_ctx* ctx, int wid, int dir, const double* in, double* out, int P, int Q, int R,
              int lvlP, int lvlQ, int lvlR);
int jwc_wpt3d(jwc_ctx* ctx, int wid, int dir, const double* in, double* out, int P, int Q, int R,
              int lvlP, int lvlQ, int lvlR);

/* AncientEgyptianDecomposition.forward/reverse(double[]) around a FastWaveletTransform
 * (kind = JWC_FWT) or WaveletPacketTransform (kind = JWC_WPT): `batch` signals of ANY length n >= 1.
 * n is split into its binary expansion, largest block first (MathToolKit.decompose,
 * tools/MathToolKit.java:57-84), and every 2^p block is transformed at full depth
 * (transforms/AncientEgyptianDecomposition.java:97-129, :144-183). */
int jwc_aed1d(jwc_ctx* ctx, int wid, int kind, int dir, const double* in, double* out, int64_t batch, int n);

/* WaveletTransform.decompose(double[]) (transforms/WaveletTransform.java:136-146): row p of the result is
 * forward(x, p), p = 0 .. log2 n, for each of `batch` signals: out is [batch][log2 n + 1][n].  One upload,
 * one one-level launch per row (row p + 1 is row p with its packets / its approximation split once more),
 * one download - instead of log2 n + 1 separate transforms.  recompose(mat, level) is
 * jwc_fwt1d / jwc_wpt1d(JWC_REVERSE) of row `level`. */
int jwc_decompose1d(jwc_ctx* ctx, int wid, int kind, const double* in, double* out, int64_t batch, int n);

/* CompressorMagnitude.compress(double[] / double[][] / double[][][]) - the same operation on `count`
 * coefficients of any rank (compressions/CompressorMagnitude.java:52-118, compressions/Compressor.java:
 * 97-110): magnitude = mean |c|, then c -> (|c| >= magnitude * threshold ? c : 0).  *magnitude (may be
 * NULL) receives the mean.  threshold must be > 0. */
int jwc_compress_magnitude(jwc_ctx* ctx, const double* in, double* out, int64_t count, double threshold,
                           double* magnitude);

/* ---- device-resident entry points: `in`/`out` are device pointers on the context's GPU,
 *      must not overlap, and the work is enqueued on the context's stream (no sync) ---------- */

int jwc_fwt1d_dev(jwc_ctx* ctx, int wid, int dir, const double* in, double* out, int64_t batch, int n,
                  int level);
int jwc_wpt1d_dev(jwc_ctx* ctx, int wid, int dir, const double* in, double* out, int64_t batch, int n,
                  int level);
int jwc_fwt2d_dev(jwc_ctx* ctx, int wid, int dir, const double* in, double* out, int64_t batch,
                  int rows, int cols, int lvlM, int lvlN);
int jwc_wpt2d_dev(jwc_ctx* ctx, int wid, int dir, const double* in, double* out, int64_t batch,
                  int rows, int cols, int lvlM, int lvlN);
int jwc_fwt3d_dev(jwc_ctx* ctx, int wid, int dir, const double* in, double* out, int P, int Q, int R,
                  int lvlP, int lvlQ, int lvlR);
int jwc_wpt3d_dev(jwc_ctx* ctx, int wid, int dir, const double* in, double* out, int P, int Q, int R,
                  int lvlP, int lvlQ, int lvlR);
int jwc_aed1d_dev(jwc_ctx* ctx, int wid, int kind, int dir, const double* in, double* out, int64_t batch,
                  int n);
int jwc_decompose1d_dev(jwc_ctx* ctx, int wid, int kind, const double* in, double* out, int64_t batch, int n);
/* device variant: *magnitude_dev (device pointer, may be NULL) receives the mean, no host sync */
int jwc_compress_magnitude_dev(jwc_ctx* ctx, const double* in, double* out, int64_t count, double threshold,
                               double* magnitude_dev);
/* Forward transform (kind = JWC_FWT | JWC_WPT) of `batch` signals followed by CompressorMagnitude over ALL
 * coefficients, in one call: out = compress(forward(in)).  The |c| sum of every 48 MB chunk runs right behind its
 * transform, out of the L2, so the reduce pass costs no HBM read; the threshold pass is in place.  n % 4 == 0. */
int jwc_forward1d_compress_dev(jwc_ctx* ctx, int wid, int kind, const double* in, double* out, int64_t batch, int n,
                               int level, double threshold, double* magnitude_dev);
/* The building block the 2-D/3-D drivers and the slab-decomposed multi-GPU volume are made of:
 * a dense [outer][n][inner] array, 1-D transform (kind = JWC_FWT | JWC_WPT) along the middle
 * axis of every (outer, inner) line. */
int jwc_axis_dev(jwc_ctx* ctx, int wid, int kind, int dir, const double* in, double* out,
                 int64_t outer, int n, int64_t inner, int level);

/* jwc_axis_dev whose FINAL output is stored straight into buffers of peer GPUs (mapped with
 * jwc_ipc_open or any other peer mapping): the exchange that would follow the pass is folded into
 * the kernel's stores.  FWT only, and only where the fused kernels apply (returns JWC_ERR_ARG
 * otherwise - callers fall back to jwc_axis_dev + a collective).
 *   mode 1 - strided axis (inner a multiple of 8): row s of outer block o goes to
 *            peer[s >> lg_seg] + o * outer_stride + base_off + (s mod 2^lg_seg) * row_stride + column
 *   mode 2 - contiguous axis (inner == 1), reverse only: line o * 2^lg_hi + j goes to
 *            peer[j >> lg_seg] + o * outer_stride + base_off + (j mod 2^lg_seg) * row_stride + sample
 * Strides and offsets are in doubles.  The caller synchronises the GPUs around the call. */
typedef struct jwc_remote_map {
  int mode, world;
  int lg_seg, lg_hi;
  int64_t outer_stride, row_stride, base_off;
  void* peer[8];
} jwc_remote_map;
int jwc_axis_dev_remote(jwc_ctx* ctx, int wid, int kind, int dir, const double* in, int64_t outer, int n,
                        int64_t inner, int level, const jwc_remote_map* map);

/* ---- memory helpers for FFI callers ------------------------------------------------------- */
int jwc_dev_alloc(jwc_ctx* ctx, size_t bytes, void** dptr);
int jwc_dev_free(jwc_ctx* ctx, void* dptr);
int jwc_h2d(jwc_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes); /* async on the stream */
int jwc_d2h(jwc_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes); /* async on the stream */
int jwc_host_alloc_pinned(jwc_ctx* ctx, size_t bytes, void** hptr);
int jwc_host_free_pinned(jwc_ctx* ctx, void* hptr);
/* Strided device-to-device copy on the copy engines (cudaMemcpy2DAsync): `height` rows of `width` bytes, source /
 * destination pitches in bytes, enqueued on `cuda_stream` (a cudaStream_t; NULL = the context's current stream).
 * Either pointer may be a mapping of a PEER GPU's memory (same process with peer access, or a cross-process
 * mapping such as torch symmetric memory): this is what re-cuts a slab-decomposed volume over NVLink without
 * taking SMs away from the axis passes that run beside it. */
int jwc_copy2d_dev(jwc_ctx* ctx, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width,
                   size_t height, void* cuda_stream);
/* Upper bound (bytes) for the staging chunk the host-buffer entry points move per step. */
int jwc_set_staging_bytes(jwc_ctx* ctx, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* JWAVE_CUDA_H */
