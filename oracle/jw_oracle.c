/*
 * jw_oracle.c - CPU restatement of JWave's FWT / WPT hot path (see jw_oracle.h for the
 * scope and parity status).  TEST INFRASTRUCTURE ONLY - never on the product path.
 *
 * Build with -ffp-contract=off: the JVM never contracts a*b+c into an FMA, so with
 * contraction off the arithmetic below is bit-faithful to the reference's operation order.
 *
 * Citations are relative to /root/reference/src/main/java/jwave/.
 */
#include "jw_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------
 * Tap tables
 * ---------------------------------------------------------------------------------------- */

#define JWO_MAX_WAVELETS 96

static jwo_wavelet g_wavelets[JWO_MAX_WAVELETS];
static int g_count = 0;

/* transforms/wavelets/Wavelet.java:104-122 (_buildOrthonormalSpace) */
static void build_orthonormal_space(jwo_wavelet* w) {
  int L = w->motherWavelength;
  for (int i = 0; i < L; i++) {
    if (i % 2 == 0)
      w->waveletDeCom[i] = w->scalingDeCom[(L - 1) - i];
    else
      w->waveletDeCom[i] = -w->scalingDeCom[(L - 1) - i];
  }
  for (int i = 0; i < L; i++) {
    w->scalingReCon[i] = w->scalingDeCom[i];
    w->waveletReCon[i] = w->waveletDeCom[i];
  }
}

static jwo_wavelet* new_wavelet(const char* cls, const char* name, int L) {
  jwo_wavelet* w = &g_wavelets[g_count++];
  memset(w, 0, sizeof(*w));
  strncpy(w->cls, cls, sizeof(w->cls) - 1);
  strncpy(w->name, name, sizeof(w->name) - 1);
  w->motherWavelength = L;
  w->transformWavelength = 2;
  w->reconFactor = 1.;
  return w;
}

static void add_literal(const char* cls, const char* name, int L, const double* taps) {
  jwo_wavelet* w = new_wavelet(cls, name, L);
  for (int i = 0; i < L; i++) w->scalingDeCom[i] = taps[i];
  build_orthonormal_space(w);
}

/* transforms/wavelets/biorthogonal/BiOrthogonal.java:43-66 (_buildBiOrthonormalSpace) */
static void build_biorthonormal_space(jwo_wavelet* w) {
  int L = w->motherWavelength;
  for (int i = 0; i < L; i++) {
    if (i % 2 == 0) {
      w->scalingReCon[i] = -w->waveletDeCom[i];
      w->waveletReCon[i] = -w->scalingDeCom[i];
    } else {
      w->scalingReCon[i] = w->waveletDeCom[i];
      w->waveletReCon[i] = w->scalingDeCom[i];
    }
  }
}

/* transforms/wavelets/biorthogonal/BiOrthogonal{11..68}.java: `lit` holds _scalingDeCom, _waveletDeCom
 * and, unless the constructor builds them, _scalingReCon and _waveletReCon (L literals each) */
static void add_bior(const char* cls, const char* name, int L, int built, const double* lit) {
  jwo_wavelet* w = new_wavelet(cls, name, L);
  for (int i = 0; i < L; i++) {
    w->scalingDeCom[i] = lit[i];
    w->waveletDeCom[i] = lit[L + i];
  }
  if (built) {
    build_biorthonormal_space(w);
  } else {
    for (int i = 0; i < L; i++) {
      w->scalingReCon[i] = lit[2 * L + i];
      w->waveletReCon[i] = lit[3 * L + i];
    }
  }
}

/* transforms/wavelets/haar/Haar1Orthogonal.java:137-161: un-normalised Haar, the reverse step carries
 * the factor .5 */
static void add_haar1_orthogonal(void) {
  jwo_wavelet* w = new_wavelet("Haar1Orthogonal", "Haar orthogonal", 2);
  w->scalingDeCom[0] = 1.;
  w->scalingDeCom[1] = 1.;
  w->waveletDeCom[0] = w->scalingDeCom[1];
  w->waveletDeCom[1] = -w->scalingDeCom[0];
  for (int i = 0; i < 2; i++) {
    w->scalingReCon[i] = w->scalingDeCom[i];
    w->waveletReCon[i] = w->waveletDeCom[i];
  }
  w->reconFactor = .5;
}

static void add_analytic(void) {
  jwo_wavelet* w;
  volatile double two = 2., three = 3., ten = 10., fifteen = 15.; /* keep sqrt() a run-time call */

  /* transforms/wavelets/haar/Haar1.java:52-68 - filters written out by hand, no
   * _buildOrthonormalSpace: wavelet = { s1, -s0 } */
  w = new_wavelet("Haar1", "Haar", 2);
  {
    double sqrt2 = sqrt(two);
    w->scalingDeCom[0] = 1. / sqrt2;
    w->scalingDeCom[1] = 1. / sqrt2;
    w->waveletDeCom[0] = w->scalingDeCom[1];
    w->waveletDeCom[1] = -w->scalingDeCom[0];
    for (int i = 0; i < 2; i++) {
      w->scalingReCon[i] = w->scalingDeCom[i];
      w->waveletReCon[i] = w->waveletDeCom[i];
    }
  }

  /* transforms/wavelets/daubechies/Daubechies2.java:53-63 */
  w = new_wavelet("Daubechies2", "Daubechies 2", 4);
  {
    double sqrt3 = sqrt(three);
    w->scalingDeCom[0] = ((1. + sqrt3) / 4.);
    w->scalingDeCom[1] = ((3. + sqrt3) / 4.);
    w->scalingDeCom[2] = ((3. - sqrt3) / 4.);
    w->scalingDeCom[3] = ((1. - sqrt3) / 4.);
    double sqrt02 = sqrt(two);
    for (int i = 0; i < 4; i++) w->scalingDeCom[i] /= sqrt02;
    build_orthonormal_space(w);
  }

  /* transforms/wavelets/daubechies/Daubechies3.java:54-66 */
  w = new_wavelet("Daubechies3", "Daubechies 3", 6);
  {
    double sqrt10 = sqrt(ten);
    double constA = sqrt(5. + 2. * sqrt10);
    w->scalingDeCom[0] = (1.0 + 1. * sqrt10 + 1. * constA) / 16.;
    w->scalingDeCom[1] = (5.0 + 1. * sqrt10 + 3. * constA) / 16.;
    w->scalingDeCom[2] = (10. - 2. * sqrt10 + 2. * constA) / 16.;
    w->scalingDeCom[3] = (10. - 2. * sqrt10 - 2. * constA) / 16.;
    w->scalingDeCom[4] = (5.0 + 1. * sqrt10 - 3. * constA) / 16.;
    w->scalingDeCom[5] = (1.0 + 1. * sqrt10 - 1. * constA) / 16.;
    double sqrt02 = sqrt(two);
    for (int i = 0; i < 6; i++) w->scalingDeCom[i] /= sqrt02;
    build_orthonormal_space(w);
  }

  /* transforms/wavelets/coiflet/Coiflet1.java:52-62 - follow the code, not its comments */
  w = new_wavelet("Coiflet1", "Coiflet 1", 6);
  {
    double sqrt02 = 1.4142135623730951;
    double sqrt15 = sqrt(fifteen);
    w->scalingDeCom[0] = sqrt02 * (sqrt15 - 3.) / 32.;
    w->scalingDeCom[1] = sqrt02 * (1. - sqrt15) / 32.;
    w->scalingDeCom[2] = sqrt02 * (6. - 2 * sqrt15) / 32.;
    w->scalingDeCom[3] = sqrt02 * (2. * sqrt15 + 6.) / 32.;
    w->scalingDeCom[4] = sqrt02 * (sqrt15 + 13.) / 32.;
    w->scalingDeCom[5] = sqrt02 * (9. - sqrt15) / 32.;
    build_orthonormal_space(w);
  }

  /* transforms/wavelets/legendre/Legendre1.java:55-64 */
  w = new_wavelet("Legendre1", "Legendre 1", 2);
  {
    w->scalingDeCom[0] = -1.;
    w->scalingDeCom[1] = -1.;
    double sqrt02 = sqrt(two);
    for (int i = 0; i < 2; i++) w->scalingDeCom[i] /= sqrt02;
    build_orthonormal_space(w);
  }

  /* transforms/wavelets/legendre/Legendre2.java:52-63 */
  w = new_wavelet("Legendre2", "Legendre 2", 4);
  {
    w->scalingDeCom[0] = -5. / 8.;
    w->scalingDeCom[1] = -3. / 8.;
    w->scalingDeCom[2] = -3. / 8.;
    w->scalingDeCom[3] = -5. / 8.;
    double sqrt02 = sqrt(two);
    for (int i = 0; i < 4; i++) w->scalingDeCom[i] /= sqrt02;
    build_orthonormal_space(w);
  }

  /* transforms/wavelets/legendre/Legendre3.java:52-65 */
  w = new_wavelet("Legendre3", "Legendre 3", 6);
  {
    w->scalingDeCom[0] = -63. / 128.;
    w->scalingDeCom[1] = -35. / 128.;
    w->scalingDeCom[2] = -30. / 128.;
    w->scalingDeCom[3] = -30. / 128.;
    w->scalingDeCom[4] = -35. / 128.;
    w->scalingDeCom[5] = -63. / 128.;
    double sqrt02 = sqrt(two);
    for (int i = 0; i < 6; i++) w->scalingDeCom[i] /= sqrt02;
    build_orthonormal_space(w);
  }
}

/* built once when the shared object is loaded, so lookups are read-only afterwards */
__attribute__((constructor)) static void init_registry(void) {
  if (g_count) return;
  add_analytic();
#define JW_LITERAL_WAVELET(cls, name, L, ...)    \
  {                                              \
    static const double taps_[] = {__VA_ARGS__}; \
    add_literal(cls, name, L, taps_);            \
  }
#include "jw_taps_literal.inc"
#undef JW_LITERAL_WAVELET
  add_haar1_orthogonal();
#define JW_BIOR_WAVELET(cls, name, L, built, ...) \
  {                                               \
    static const double lit_[] = {__VA_ARGS__};   \
    add_bior(cls, name, L, built, lit_);          \
  }
#include "jw_taps_bior.inc"
#undef JW_BIOR_WAVELET
}

int jwo_wavelet_count(void) {
  init_registry();
  return g_count;
}

const jwo_wavelet* jwo_wavelet_at(int idx) {
  init_registry();
  return (idx >= 0 && idx < g_count) ? &g_wavelets[idx] : NULL;
}

const jwo_wavelet* jwo_wavelet_find(const char* name) {
  init_registry();
  for (int i = 0; i < g_count; i++)
    if (!strcmp(name, g_wavelets[i].cls) || !strcmp(name, g_wavelets[i].name)) return &g_wavelets[i];
  return NULL;
}

/* ------------------------------------------------------------------------------------------
 * MathToolKit
 * ---------------------------------------------------------------------------------------- */

/* tools/MathToolKit.java:185-189 */
int jwo_is_binary(int number) { return number > 0 && ((number & (number - 1)) == 0); }

/* tools/MathToolKit.java:202-208 - float log, then truncate */
int jwo_get_exponent(double f) {
  int e = (int)(log(f) / log(2.));
  return e;
}

/* transforms/BasicTransform.java:683-697 (calcExponent); caller has checked isBinary */
static int calc_exponent(int number) { return jwo_get_exponent((double)number); }

/* ------------------------------------------------------------------------------------------
 * The two hot loops
 * ---------------------------------------------------------------------------------------- */

/* transforms/wavelets/Wavelet.java:236-260 */
void jwo_wavelet_forward(const jwo_wavelet* w, const double* arrTime, int n, double* arrHilb) {
  const int L = w->motherWavelength;
  int h = n >> 1;
  for (int i = 0; i < h; i++) {
    arrHilb[i] = arrHilb[i + h] = 0.;
    for (int j = 0; j < L; j++) {
      int k = (i << 1) + j;
      while (k >= n) k -= n;
      arrHilb[i] += arrTime[k] * w->scalingDeCom[j];
      arrHilb[i + h] += arrTime[k] * w->waveletDeCom[j];
    }
  }
}

/* transforms/wavelets/Wavelet.java:277-303; identical loops in biorthogonal/BiOrthogonal.java:108-133;
 * haar/Haar1Orthogonal.java:175-207 multiplies every term by its _energyCorrectionFactor */
void jwo_wavelet_reverse(const jwo_wavelet* w, const double* arrHilb, int n, double* arrTime) {
  const int L = w->motherWavelength;
  for (int i = 0; i < n; i++) arrTime[i] = 0.;
  int h = n >> 1;
  for (int i = 0; i < h; i++) {
    for (int j = 0; j < L; j++) {
      int k = (i << 1) + j;
      while (k >= n) k -= n;
      if (w->reconFactor == 1.)
        arrTime[k] += (arrHilb[i] * w->scalingReCon[j]) + (arrHilb[i + h] * w->waveletReCon[j]);
      else
        arrTime[k] += w->reconFactor * ((arrHilb[i] * w->scalingReCon[j]) + (arrHilb[i + h] * w->waveletReCon[j]));
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * 1-D level loops.  Like the reference, every Wavelet.forward/reverse call gets a fresh
 * buffer that is then copied back (this is what the JVM path pays for, too).
 * ---------------------------------------------------------------------------------------- */

static int check_1d(int n, int level) {
  if (!jwo_is_binary(n)) return JWO_ERR_NOT_BINARY;
  int noOfLevels = calc_exponent(n);
  if (level < 0 || level > noOfLevels) return JWO_ERR_LEVEL;
  return JWO_OK;
}

/* transforms/FastWaveletTransform.java:71-101 */
int jwo_fwt_forward(const jwo_wavelet* w, const double* arrTime, int n, int level, double* arrHilb) {
  int st = check_1d(n, level);
  if (st) return st;
  memcpy(arrHilb, arrTime, sizeof(double) * (size_t)n);
  int l = 0;
  int h = n;
  int transformWavelength = w->transformWavelength;
  while (h >= transformWavelength && l < level) {
    double* arrTempPart = (double*)malloc(sizeof(double) * (size_t)h);
    jwo_wavelet_forward(w, arrHilb, h, arrTempPart);
    memcpy(arrHilb, arrTempPart, sizeof(double) * (size_t)h);
    free(arrTempPart);
    h = h >> 1;
    l++;
  }
  return JWO_OK;
}

/* transforms/FastWaveletTransform.java:119-153 */
int jwo_fwt_reverse(const jwo_wavelet* w, const double* arrHilb, int n, int level, double* arrTime) {
  int st = check_1d(n, level);
  if (st) return st;
  memcpy(arrTime, arrHilb, sizeof(double) * (size_t)n);
  int transformWavelength = w->transformWavelength;
  long h = transformWavelength;
  int steps = calc_exponent(n);
  for (int l = level; l < steps; l++) h = h << 1;
  while (h <= n && h >= transformWavelength) {
    double* arrTempPart = (double*)malloc(sizeof(double) * (size_t)h);
    jwo_wavelet_reverse(w, arrTime, (int)h, arrTempPart);
    memcpy(arrTime, arrTempPart, sizeof(double) * (size_t)h);
    free(arrTempPart);
    h = h << 1;
  }
  return JWO_OK;
}

/* one packet of WaveletPacketTransform.java:102-113 / :170-181: copy out, transform, copy back */
static void wpt_packet(int dir, const jwo_wavelet* w, double* arr, int p, int h) {
  double* iBuf = (double*)malloc(sizeof(double) * (size_t)h);
  double* oBuf = (double*)malloc(sizeof(double) * (size_t)h);
  for (int i = 0; i < h; i++) iBuf[i] = arr[i + ((size_t)p * h)];
  if (dir == JWO_FORWARD)
    jwo_wavelet_forward(w, iBuf, h, oBuf);
  else
    jwo_wavelet_reverse(w, iBuf, h, oBuf);
  for (int i = 0; i < h; i++) arr[i + ((size_t)p * h)] = oBuf[i];
  free(iBuf);
  free(oBuf);
}

/* transforms/WaveletPacketTransform.java:73-124 */
int jwo_wpt_forward(const jwo_wavelet* w, const double* arrTime, int n, int level, double* arrHilb) {
  int st = check_1d(n, level);
  if (st) return st;
  for (int i = 0; i < n; i++) arrHilb[i] = arrTime[i];
  int k = n;
  int h = n;
  int transformWavelength = w->transformWavelength;
  int l = 0;
  while (h >= transformWavelength && l < level) {
    int g = k / h;
    for (int p = 0; p < g; p++) wpt_packet(JWO_FORWARD, w, arrHilb, p, h);
    h = h >> 1;
    l++;
  }
  return JWO_OK;
}

/* transforms/WaveletPacketTransform.java:141-191 */
int jwo_wpt_reverse(const jwo_wavelet* w, const double* arrHilb, int n, int level, double* arrTime) {
  int st = check_1d(n, level);
  if (st) return st;
  memcpy(arrTime, arrHilb, sizeof(double) * (size_t)n);
  int transformWavelength = w->transformWavelength;
  int k = n;
  long h = transformWavelength;
  int steps = calc_exponent(n);
  for (int l = level; l < steps; l++) h = h << 1;
  while (h <= n && h >= transformWavelength) {
    int g = (int)(k / h);
    for (int p = 0; p < g; p++) wpt_packet(JWO_REVERSE, w, arrTime, p, (int)h);
    h = h << 1;
  }
  return JWO_OK;
}

int jwo_1d(int kind, int dir, const jwo_wavelet* w, const double* in, int n, int level, double* out) {
  if (kind == JWO_FWT)
    return dir == JWO_FORWARD ? jwo_fwt_forward(w, in, n, level, out) : jwo_fwt_reverse(w, in, n, level, out);
  if (kind == JWO_WPT)
    return dir == JWO_FORWARD ? jwo_wpt_forward(w, in, n, level, out) : jwo_wpt_reverse(w, in, n, level, out);
  return JWO_ERR_ARG;
}

/* ------------------------------------------------------------------------------------------
 * 2-D and 3-D drivers (dense row-major storage instead of Java's array-of-arrays)
 * ---------------------------------------------------------------------------------------- */

/* one row of BasicTransform.java:369-381 (forward) / :458-470 (reverse): src row -> dst row */
static int row_1d(int kind, int dir, const jwo_wavelet* w, const double* src, double* dst, int row,
                  int cols, int lvl) {
  double* a = (double*)malloc(sizeof(double) * (size_t)cols);
  double* b = (double*)malloc(sizeof(double) * (size_t)cols);
  for (int j = 0; j < cols; j++) a[j] = src[(size_t)row * cols + j];
  int st = jwo_1d(kind, dir, w, a, cols, lvl, b);
  if (!st)
    for (int j = 0; j < cols; j++) dst[(size_t)row * cols + j] = b[j];
  free(a);
  free(b);
  return st;
}

/* one column of BasicTransform.java:383-395 (forward) / :444-456 (reverse): strided gather/scatter */
static int col_1d(int kind, int dir, const jwo_wavelet* w, const double* src, double* dst, int col,
                  int rows, int cols, int lvl) {
  double* a = (double*)malloc(sizeof(double) * (size_t)rows);
  double* b = (double*)malloc(sizeof(double) * (size_t)rows);
  for (int i = 0; i < rows; i++) a[i] = src[(size_t)i * cols + col];
  int st = jwo_1d(kind, dir, w, a, rows, lvl, b);
  if (!st)
    for (int i = 0; i < rows; i++) dst[(size_t)i * cols + col] = b[i];
  free(a);
  free(b);
  return st;
}

/* transforms/BasicTransform.java:361-399 (forward: rows with lvlN, then columns with lvlM) and
 * :436-474 (reverse: columns with lvlM, then rows with lvlN) */
int jwo_2d(int kind, int dir, const jwo_wavelet* w, const double* in, int rows, int cols, int lvlM,
           int lvlN, double* out) {
  int st;
  if (rows <= 0 || cols <= 0) return JWO_ERR_ARG;
  if (dir == JWO_FORWARD) {
    for (int i = 0; i < rows; i++)
      if ((st = row_1d(kind, dir, w, in, out, i, cols, lvlN))) return st;
    for (int j = 0; j < cols; j++)
      if ((st = col_1d(kind, dir, w, out, out, j, rows, cols, lvlM))) return st;
  } else {
    for (int j = 0; j < cols; j++)
      if ((st = col_1d(kind, dir, w, in, out, j, rows, cols, lvlM))) return st;
    for (int i = 0; i < rows; i++)
      if ((st = row_1d(kind, dir, w, out, out, i, cols, lvlN))) return st;
  }
  return JWO_OK;
}

/* transforms/BasicTransform.java:509-566 and :602-659.  Both directions run the 2-D transform on
 * every [j][k] slice with (lvlP, lvlQ) - the reference's level shift, SURVEY.md F5 - and then
 * the 1-D transform along i with lvlR. */
int jwo_3d(int kind, int dir, const jwo_wavelet* w, const double* in, int P, int Q, int R, int lvlP,
           int lvlQ, int lvlR, double* out) {
  int st;
  if (P <= 0 || Q <= 0 || R <= 0) return JWO_ERR_ARG;
  size_t slice = (size_t)Q * R;
  for (int i = 0; i < P; i++)
    if ((st = jwo_2d(kind, dir, w, in + i * slice, Q, R, lvlP, lvlQ, out + i * slice))) return st;
  double* a = (double*)malloc(sizeof(double) * (size_t)P);
  double* b = (double*)malloc(sizeof(double) * (size_t)P);
  st = JWO_OK;
  for (int j = 0; j < Q && !st; j++) {
    for (int k = 0; k < R && !st; k++) {
      for (int i = 0; i < P; i++) a[i] = out[i * slice + (size_t)j * R + k];
      st = jwo_1d(kind, dir, w, a, P, lvlR, b);
      if (!st)
        for (int i = 0; i < P; i++) out[i * slice + (size_t)j * R + k] = b[i];
    }
  }
  free(a);
  free(b);
  return st;
}

/* ------------------------------------------------------------------------------------------
 * Ancient Egyptian decomposition (arbitrary lengths)
 * ---------------------------------------------------------------------------------------- */

/* tools/MathToolKit.java:57-84 */
int jwo_decompose(int number, int* out) {
  if (number < 1) return 0;
  int pos = 0;
  double current = (double)number;
  while (current >= 1.) {
    int power = jwo_get_exponent(current);
    out[pos] = power;
    current = current - ldexp(1., power); /* MathToolKit.scalb( 1., power ) */
    pos++;
  }
  return pos;
}

/* transforms/AncientEgyptianDecomposition.java:97-129 (forward) and :144-183 (reverse) */
int jwo_aed(int kind, int dir, const jwo_wavelet* w, const double* in, int n, double* out) {
  int mult[32];
  int cnt = jwo_decompose(n, mult);
  if (cnt == 0) return JWO_ERR_ARG;
  int offSet = 0;
  for (int m = 0; m < cnt; m++) {
    int len = (int)ldexp(1., mult[m]);
    /* WaveletTransform.forward(double[]) / reverse(double[]): full depth = log2(len) */
    int st = jwo_1d(kind, dir, w, in + offSet, len, mult[m], out + offSet);
    if (st) return st;
    offSet += len;
  }
  return JWO_OK;
}

/* compressions/CompressorMagnitude.java:52-68 and compressions/Compressor.java:97-110 */
double jwo_compress_magnitude(const double* arr, long n, double threshold, double* out) {
  double magnitude = 0.;
  for (long i = 0; i < n; i++) magnitude += fabs(arr[i]);
  magnitude /= (double)n;
  for (long i = 0; i < n; i++) {
    if (fabs(arr[i]) >= magnitude * threshold)
      out[i] = arr[i];
    else
      out[i] = 0.;
  }
  return magnitude;
}

/* ------------------------------------------------------------------------------------------
 * CPU-baseline drivers.  A small persistent pthread pool stands in for the JVM's
 * ForkJoinPool / fixed executor (no OpenMP runtime in this image).
 * ---------------------------------------------------------------------------------------- */

typedef void (*jwo_body)(long idx, void* ctx);

static struct {
  pthread_mutex_t mu;
  pthread_cond_t go, done;
  pthread_t* workers;
  int nworkers;      /* threads besides the caller */
  long generation;   /* bumped for every parallel_for */
  int active;        /* workers admitted to the current loop */
  int running;       /* workers still inside the current loop */
  jwo_body body;
  void* ctx;
  long n, chunk;
  atomic_long next;
} g_pool = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER,
            NULL, 0, 0, 0, 0, NULL, NULL, 0, 1, 0};

static pthread_mutex_t g_pool_user = PTHREAD_MUTEX_INITIALIZER; /* one parallel_for at a time */

static void pool_drain(void) {
  for (;;) {
    long b = atomic_fetch_add(&g_pool.next, g_pool.chunk);
    if (b >= g_pool.n) break;
    long e = b + g_pool.chunk < g_pool.n ? b + g_pool.chunk : g_pool.n;
    for (long i = b; i < e; i++) g_pool.body(i, g_pool.ctx);
  }
}

static void* pool_worker(void* arg) {
  long id = (long)arg;
  long seen = 0;
  pthread_mutex_lock(&g_pool.mu);
  for (;;) {
    while (g_pool.generation == seen) pthread_cond_wait(&g_pool.go, &g_pool.mu);
    seen = g_pool.generation;
    if (id >= g_pool.active) continue;
    pthread_mutex_unlock(&g_pool.mu);
    pool_drain();
    pthread_mutex_lock(&g_pool.mu);
    if (--g_pool.running == 0) pthread_cond_signal(&g_pool.done);
  }
  return NULL;
}

int jwo_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n < 1 ? 1 : (int)n;
}

static int pick_threads(int threads) {
  int m = jwo_max_threads();
  return (threads <= 0 || threads > m) ? m : threads;
}

static void pool_grow(int want) {
  if (want <= g_pool.nworkers) return;
  g_pool.workers = (pthread_t*)realloc(g_pool.workers, sizeof(pthread_t) * (size_t)want);
  for (long i = g_pool.nworkers; i < want; i++)
    pthread_create(&g_pool.workers[i], NULL, pool_worker, (void*)i);
  g_pool.nworkers = want;
}

/* run body(0..n-1) on `threads` threads (the caller is one of them), dynamic chunks */
static void parallel_for(long n, long chunk, int threads, jwo_body body, void* ctx) {
  if (threads <= 1 || n <= 1) {
    for (long i = 0; i < n; i++) body(i, ctx);
    return;
  }
  pthread_mutex_lock(&g_pool_user);
  pthread_mutex_lock(&g_pool.mu);
  pool_grow(threads - 1);
  g_pool.body = body;
  g_pool.ctx = ctx;
  g_pool.n = n;
  g_pool.chunk = chunk < 1 ? 1 : chunk;
  atomic_store(&g_pool.next, 0);
  g_pool.active = threads - 1;
  g_pool.running = threads - 1;
  g_pool.generation++;
  pthread_cond_broadcast(&g_pool.go);
  pthread_mutex_unlock(&g_pool.mu);
  pool_drain();
  pthread_mutex_lock(&g_pool.mu);
  while (g_pool.running) pthread_cond_wait(&g_pool.done, &g_pool.mu);
  pthread_mutex_unlock(&g_pool.mu);
  pthread_mutex_unlock(&g_pool_user);
}

typedef struct {
  int kind, dir;
  const jwo_wavelet* w;
  const double* in;
  double* out;
  int n, level;
  int rows, cols, lvl, do_rows;
  int P, Q, R, lvlP, lvlQ, lvlR;
  double* arr;
  int h;
  atomic_int status;
} jwo_job;

static void note_status(jwo_job* j, int st) {
  int zero = 0;
  if (st) atomic_compare_exchange_strong(&j->status, &zero, st);
}

static void body_signal(long s, void* ctx) {
  jwo_job* j = (jwo_job*)ctx;
  note_status(j, jwo_1d(j->kind, j->dir, j->w, j->in + (size_t)s * j->n, j->n, j->level,
                        j->out + (size_t)s * j->n));
}

/* independent signals, one task per signal (test ParallelizationOpportunityTest.java:79-110) */
int jwo_batch_1d(int kind, int dir, const jwo_wavelet* w, const double* in, long batch, int n,
                 int level, double* out, int threads) {
  jwo_job j = {0};
  j.kind = kind; j.dir = dir; j.w = w; j.in = in; j.out = out; j.n = n; j.level = level;
  parallel_for(batch, 4, pick_threads(threads), body_signal, &j);
  return atomic_load(&j.status);
}

static void body_packet(long p, void* ctx) {
  jwo_job* j = (jwo_job*)ctx;
  wpt_packet(j->dir, j->w, j->arr, (int)p, j->h);
}

/* transforms/ParallelWaveletPacketTransform.java:79-146.  shouldUseParallel (:155-158):
 * packetSize >= 64 && packets >= 8.  Signals are looped by the caller, as a user of the
 * reference class would. */
int jwo_parallel_wpt(int dir, const jwo_wavelet* w, const double* in, long batch, int n, int level,
                     double* out, int threads) {
  int st = check_1d(n, level);
  if (st) return st;
  /* PooledWaveletPacketTransform.java:29 - the pooled/parallel forward rejects level <= 0 */
  if (dir == JWO_FORWARD && level <= 0) return JWO_ERR_LEVEL;
  int nt = pick_threads(threads);
  int steps = calc_exponent(n);
  jwo_job j = {0};
  j.dir = dir; j.w = w;
  for (long s = 0; s < batch; s++) {
    double* arr = out + (size_t)s * n;
    memcpy(arr, in + (size_t)s * n, sizeof(double) * (size_t)n);
    j.arr = arr;
    if (dir == JWO_FORWARD) {
      int h = n, l = 0;
      while (h >= w->transformWavelength && l < level) {
        int g = n / h;
        j.h = h;
        parallel_for(g, 1, (h >= 64 && g >= 8) ? nt : 1, body_packet, &j);
        h >>= 1;
        l++;
      }
    } else {
      long h = w->transformWavelength;
      for (int l = level; l < steps; l++) h <<= 1;
      while (h <= n && h >= w->transformWavelength) {
        int g = (int)(n / h);
        j.h = (int)h;
        parallel_for(g, 1, (h >= 64 && g >= 8) ? nt : 1, body_packet, &j);
        h <<= 1;
      }
    }
  }
  return JWO_OK;
}

static void body_line(long q, void* ctx) {
  jwo_job* j = (jwo_job*)ctx;
  note_status(j, j->do_rows ? row_1d(j->kind, j->dir, j->w, j->in, j->out, (int)q, j->cols, j->lvl)
                            : col_1d(j->kind, j->dir, j->w, j->in, j->out, (int)q, j->rows, j->cols, j->lvl));
}

/* transforms/ParallelTransform.java:70-93 (forward) and :111-134 (reverse): rows in parallel,
 * then columns in parallel (reverse: columns, then rows); images are looped. */
int jwo_parallel_2d(int kind, int dir, const jwo_wavelet* w, const double* in, long batch, int rows,
                    int cols, int lvlM, int lvlN, double* out, int threads) {
  int nt = pick_threads(threads);
  size_t img = (size_t)rows * cols;
  jwo_job j = {0};
  j.kind = kind; j.dir = dir; j.w = w; j.rows = rows; j.cols = cols;
  for (long b = 0; b < batch && !atomic_load(&j.status); b++) {
    for (int pass = 0; pass < 2; pass++) {
      j.do_rows = (dir == JWO_FORWARD) ? (pass == 0) : (pass == 1);
      j.in = (pass == 0) ? in + b * img : out + b * img;
      j.out = out + b * img;
      j.lvl = j.do_rows ? lvlN : lvlM;
      parallel_for(j.do_rows ? rows : cols, 8, nt, body_line, &j);
    }
  }
  return atomic_load(&j.status);
}

static void body_slice(long i, void* ctx) {
  jwo_job* j = (jwo_job*)ctx;
  size_t slice = (size_t)j->Q * j->R;
  note_status(j, jwo_2d(j->kind, j->dir, j->w, j->in + (size_t)i * slice, j->Q, j->R, j->lvlP, j->lvlQ,
                        j->out + (size_t)i * slice));
}

/* Space3DTransformTask.computeDirectly (ParallelTransform.java:376-404) for one j: every k, gather along i */
static void body_pencils(long jj, void* ctx) {
  jwo_job* j = (jwo_job*)ctx;
  size_t slice = (size_t)j->Q * j->R;
  double* a = (double*)malloc(sizeof(double) * (size_t)j->P);
  double* b = (double*)malloc(sizeof(double) * (size_t)j->P);
  for (int k = 0; k < j->R; k++) {
    size_t off = (size_t)jj * j->R + k;
    for (int i = 0; i < j->P; i++) a[i] = j->in[i * slice + off];
    int st = jwo_1d(j->kind, j->dir, j->w, a, j->P, j->lvlR, b);
    note_status(j, st);
    if (st) break;
    for (int i = 0; i < j->P; i++) j->out[i * slice + off] = b[i];
  }
  free(a);
  free(b);
}

/* transforms/ParallelTransform.java:137-173 (forward: every [j][k] slice as a pool task running the 2-D
 * transform with (lvlP, lvlQ) - the level shift of SURVEY.md F5 is inherited - then the i axis as
 * Space3DTransformTask over blocks of j, :338-406) and :175-213 (reverse: the i axis FIRST, then the slices -
 * the opposite order of BasicTransform.java:611-655). */
int jwo_parallel_3d(int kind, int dir, const jwo_wavelet* w, const double* in, int P, int Q, int R,
                    int lvlP, int lvlQ, int lvlR, double* out, int threads) {
  if (P <= 0 || Q <= 0 || R <= 0) return JWO_ERR_ARG;
  int nt = pick_threads(threads);
  jwo_job j = {0};
  j.kind = kind; j.dir = dir; j.w = w; j.P = P; j.Q = Q; j.R = R;
  j.lvlP = lvlP; j.lvlQ = lvlQ; j.lvlR = lvlR;
  j.out = out;
  if (dir == JWO_FORWARD) {
    j.in = in;
    parallel_for(P, 1, nt, body_slice, &j);
    if (atomic_load(&j.status)) return atomic_load(&j.status);
    j.in = out;
    parallel_for(Q, 1, nt, body_pencils, &j);
  } else {
    j.in = in;
    parallel_for(Q, 1, nt, body_pencils, &j);
    if (atomic_load(&j.status)) return atomic_load(&j.status);
    j.in = out;
    parallel_for(P, 1, nt, body_slice, &j);
  }
  return atomic_load(&j.status);
}
