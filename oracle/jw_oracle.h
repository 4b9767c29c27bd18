/*
 * jw_oracle.h - CPU restatement of JWave's discrete-wavelet hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under jwave_b200/ (the product) may include, link or
 * call this.  Allowed users: tests/, __graft_entry__.smoke(), and bench.py's cpu_baseline /
 * --impl reference legs, where it plays the role of "JWave on the host cores".
 *
 * Parity status: the JWave reference is Java and there is no JVM in the build container or
 * on the GPU box, so the reference itself cannot be run (DESIGN.md "Oracle").  This file is
 * a line-by-line restatement of the reference loops (citations on every function).  It is
 * PINNED against every known-answer vector the reference's own tests hold for this path
 * (Haar level-1 of [1..8], Haar/db filter fixtures, the all-ones ladder of SteppingTest and
 * DecomposeTest for every wavelet in WaveletBuilder.create2arr, the PropertyBasedTest
 * identities) - see tests/test_oracle_golden.py.  It is UNPINNED ("parity unpinned") for
 * non-constant coefficient values of wavelets other than Haar and for 2-D/3-D results,
 * because the reference has no such vectors; for those it is
 * cross-checked against an independent numpy restatement (oracle/np_oracle.py).
 *
 * All citations are relative to /root/reference/src/main/java/jwave/.
 */
#ifndef JW_ORACLE_H
#define JW_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define JWO_MAX_TAPS 64

enum { JWO_OK = 0, JWO_ERR_NOT_BINARY = 1, JWO_ERR_LEVEL = 2, JWO_ERR_ARG = 3 };
enum { JWO_FWT = 0, JWO_WPT = 1 };
enum { JWO_FORWARD = 0, JWO_REVERSE = 1 };

/* transforms/wavelets/Wavelet.java:52-75 - the four tap arrays and the two wavelengths */
typedef struct jwo_wavelet {
  char cls[32];   /* Java class name, e.g. "Daubechies4" */
  char name[32];  /* JWave display name, e.g. "Daubechies 4" */
  int motherWavelength;
  int transformWavelength;
  double scalingDeCom[JWO_MAX_TAPS];
  double waveletDeCom[JWO_MAX_TAPS];
  double scalingReCon[JWO_MAX_TAPS];
  double waveletReCon[JWO_MAX_TAPS];
  /* 1.0, except transforms/wavelets/haar/Haar1Orthogonal.java:39 (_energyCorrectionFactor = .5),
   * which scales every reconstruction term (Haar1Orthogonal.java:197-199) */
  double reconFactor;
} jwo_wavelet;

/* Registry: Haar1, Daubechies2-20, Symlet2-20, Coiflet1-5, Legendre1-3 (the north-star families), then
 * Haar1Orthogonal and BiOrthogonal 1/1 .. 6/8 (four independent filters). */
int jwo_wavelet_count(void);
const jwo_wavelet* jwo_wavelet_at(int idx);
/* Lookup by Java class name ("Symlet8") or display name ("Symlet 8"); NULL when unknown. */
const jwo_wavelet* jwo_wavelet_find(const char* name);

/* tools/MathToolKit.java:185-189, :202-208 */
int jwo_is_binary(int number);
int jwo_get_exponent(double f);

/* transforms/wavelets/Wavelet.java:236-260 and :277-303.  `out` has n entries. */
void jwo_wavelet_forward(const jwo_wavelet* w, const double* arrTime, int n, double* out);
void jwo_wavelet_reverse(const jwo_wavelet* w, const double* arrHilb, int n, double* out);

/* transforms/FastWaveletTransform.java:71-101, :119-153 */
int jwo_fwt_forward(const jwo_wavelet* w, const double* arrTime, int n, int level, double* out);
int jwo_fwt_reverse(const jwo_wavelet* w, const double* arrHilb, int n, int level, double* out);
/* transforms/WaveletPacketTransform.java:73-124, :141-191 */
int jwo_wpt_forward(const jwo_wavelet* w, const double* arrTime, int n, int level, double* out);
int jwo_wpt_reverse(const jwo_wavelet* w, const double* arrHilb, int n, int level, double* out);
/* kind = JWO_FWT|JWO_WPT, dir = JWO_FORWARD|JWO_REVERSE */
int jwo_1d(int kind, int dir, const jwo_wavelet* w, const double* in, int n, int level, double* out);

/* transforms/BasicTransform.java:361-399, :436-474 on a dense row-major rows x cols matrix */
int jwo_2d(int kind, int dir, const jwo_wavelet* w, const double* in, int rows, int cols,
           int lvlM, int lvlN, double* out);
/* transforms/BasicTransform.java:509-566, :602-659 on a dense P x Q x R volume (index [i][j][k]);
 * keeps the reference's level-argument shift. */
int jwo_3d(int kind, int dir, const jwo_wavelet* w, const double* in, int P, int Q, int R,
           int lvlP, int lvlQ, int lvlR, double* out);

/* tools/MathToolKit.java:57-84 (decompose): exponents of the binary expansion of `number`, largest
 * first, e.g. 13 -> {3, 2, 0}.  Returns the count (0 when number < 1); `out` needs 32 entries. */
int jwo_decompose(int number, int* out);
/* transforms/AncientEgyptianDecomposition.java:97-129, :144-183: arbitrary length n, split into
 * 2^p blocks by jwo_decompose, each block transformed at full depth by the wrapped FWT / WPT. */
int jwo_aed(int kind, int dir, const jwo_wavelet* w, const double* in, int n, double* out);

/* compressions/CompressorMagnitude.java:52-68 over compressions/Compressor.java:97-110: magnitude =
 * mean |x| (left-to-right sum), out[i] = |x[i]| >= magnitude * threshold ? x[i] : 0.  Returns the
 * magnitude. */
double jwo_compress_magnitude(const double* arr, long n, double threshold, double* out);

/* CPU-baseline drivers (OpenMP).  `threads` <= 0 means all available. Return the status of the
 * first failing signal or JWO_OK. */
/* independent signals on a fixed pool, pattern of test ParallelizationOpportunityTest.java:79-110 */
int jwo_batch_1d(int kind, int dir, const jwo_wavelet* w, const double* in, long batch, int n,
                 int level, double* out, int threads);
/* transforms/ParallelWaveletPacketTransform.java:79-146, :155-158: per level, packets run in
 * parallel only when packetSize >= 64 and packets >= 8; signals are looped. */
int jwo_parallel_wpt(int dir, const jwo_wavelet* w, const double* in, long batch, int n, int level,
                     double* out, int threads);
/* transforms/ParallelTransform.java:70-93, :137-173: rows / columns / slices in parallel */
int jwo_parallel_2d(int kind, int dir, const jwo_wavelet* w, const double* in, long batch, int rows,
                    int cols, int lvlM, int lvlN, double* out, int threads);
/* transforms/ParallelTransform.java:137-173, :175-213: slices as pool tasks + the i axis over blocks of j;
 * the reverse runs the i axis first */
int jwo_parallel_3d(int kind, int dir, const jwo_wavelet* w, const double* in, int P, int Q, int R,
                    int lvlP, int lvlQ, int lvlR, double* out, int threads);
int jwo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
