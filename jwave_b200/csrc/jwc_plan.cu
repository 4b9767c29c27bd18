// jwc_plan.cu - the level planner.
//
// Buffers.  The reference copies the input and then overwrites a shrinking / growing prefix
// level by level (FastWaveletTransform.java:85-99).  On the GPU a level cannot run in place
// (every output reads L inputs that other threads overwrite), so:
//   FWT forward : details d_l go straight to their final place in `out`; approximations a_l
//                 ping-pong between two compact scratch buffers; the last a_l lands in `out`.
//   FWT reverse : d_l is read from `in`, the growing approximation ping-pongs in scratch and
//                 the last level writes `out`.
//   WPT         : whole-array ping-pong between `out` and one scratch buffer, phased so the
//                 last level writes `out`.
#include "jwc_plan.cuh"

#include "jwc_kernels.cuh"

namespace jwc {

cudaError_t ensure_scratch(jwc_ctx* ctx, int slot, size_t bytes, double** ptr) {
  Scratch& s = ctx->scratch[slot];
  if (s.bytes < bytes) {
    cudaError_t e;
    if (s.ptr) {
      // earlier launches on the stream may still read the old block
      if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return e;
      if ((e = cudaFree(s.ptr)) != cudaSuccess) return e;
      s.ptr = nullptr;
      s.bytes = 0;
    }
    if ((e = cudaMalloc(&s.ptr, bytes)) != cudaSuccess) return e;
    s.bytes = bytes;
  }
  *ptr = static_cast<double*>(s.ptr);
  return cudaSuccess;
}

static cudaError_t copy_through(jwc_ctx* ctx, const double* in, double* out, int64_t count) {
  return cudaMemcpyAsync(out, in, size_t(count) * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream);
}

#define JWC_TRY(call)                       \
  do {                                      \
    cudaError_t e__ = (call);               \
    if (e__ != cudaSuccess) return e__;     \
  } while (0)

// Lines per CTA in resident mode: as many as fit a shared-memory budget that keeps ~4 CTAs per SM,
// so short lines (the tail of every full-depth transform) still fill the CTA's threads.
static int resident_lines(const jwc_ctx* ctx, int h, int bytes_per_sample_x10) {
  const int64_t per_line = int64_t(h) * bytes_per_sample_x10 / 10 + 64;
  int64_t g = (int64_t(ctx->res_kb) * 1024) / per_line;
  if (g < 1) g = 1;
  if (g > 64) g = 64;
  return int(g);
}

// ---- FWT ---------------------------------------------------------------------------------------

static bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; }

// The fused kernels use 16-byte cp.async and 32-byte vector stores: contiguous lines whose
// bases are 32-byte aligned (true for every n >= 4 on a 32-byte aligned allocation).
static bool fused_ok(const jwc_ctx* ctx, const double* in, const double* out, int n, int64_t inner) {
  return !ctx->force_generic && inner == 1 && n >= 4 && aligned32(in) && aligned32(out);
}

static cudaError_t fwt_forward_generic(jwc_ctx* ctx, const WaveletRec& w, const double* in, double* out,
                                       int64_t outer, int n, int64_t inner, int level) {
  const int64_t line = int64_t(n) * inner;
  double* S[2] = {nullptr, nullptr};
  if (level >= 2) JWC_TRY(ensure_scratch(ctx, 1, size_t(outer) * (n / 2) * inner * sizeof(double), &S[1]));
  if (level >= 3) JWC_TRY(ensure_scratch(ctx, 0, size_t(outer) * (n / 4) * inner * sizeof(double), &S[0]));
  const double* src = in;
  int64_t src_os = line;
  int h = n;
  for (int l = 1; l <= level; ++l) {
    const int half = h >> 1;
    const bool last = (l == level);
    FwdLevelArgs a;
    a.src = src; a.src_os = src_os;
    a.dstA = last ? out : S[l & 1];
    a.dstA_os = last ? line : int64_t(half) * inner;
    a.dstD = out + int64_t(half) * inner;
    a.dstD_os = line;
    a.outer = outer; a.inner = inner; a.half = half;
    JWC_TRY(launch_fwd_level_generic(ctx, w.L, w.de, a));
    src = a.dstA; src_os = a.dstA_os; h = half;
  }
  return cudaSuccess;
}

// Strided axis (inner > 1): the same pass structure with the [sample][8 columns] kernels.
static bool strided_ok(const jwc_ctx* ctx, const double* in, const double* out, int n, int64_t inner) {
  return !ctx->force_generic && inner >= 8 && inner % 8 == 0 && n >= 4 && aligned32(in) && aligned32(out);
}

static cudaError_t fwt_forward_strided(jwc_ctx* ctx, const WaveletRec& w, const double* in, double* out,
                                       int64_t outer, int n, int64_t inner, int level) {
  // second generation (jwc_fwt_strided2.cu) where the geometry allows; it declines per launch with
  // cudaErrorNotSupported and the first generation, which takes the same pass, runs instead
  const bool v2 = ctx->str_v2 && inner % 16 == 0;
  const int cap = v2 ? ctx->str2_cap : ctx->str_cap, tileT = v2 ? ctx->str2_tile : ctx->str_tile;
  const int m_tile = v2 ? fwt_str2_tile_levels(w.L, tileT, ctx->str2_m) : fwt_str_tile_levels(w.L, tileT);
  double* S[2] = {nullptr, nullptr};
  if (n > cap && level > m_tile) {
    JWC_TRY(ensure_scratch(ctx, 1, size_t(outer) * (n >> m_tile) * inner * sizeof(double), &S[1]));
    if ((n >> m_tile) > cap && level > 2 * m_tile)
      JWC_TRY(ensure_scratch(ctx, 0, size_t(outer) * (n >> (2 * m_tile)) * inner * sizeof(double), &S[0]));
  }
  FwtFwdStrArgs a;
  a.src = in; a.src_os = int64_t(n) * inner;
  a.dstD = out; a.dstD_os = int64_t(n) * inner;
  a.outer = outer; a.inner = inner;
  if (ctx->remote) a.rmD = *ctx->remote;  // every d_k row is final output
  int h = n, left = level, pass = 0;
  while (left > 0) {
    const bool resident = (h <= cap) || (h < tileT);
    a.h = h;
    a.T = resident ? h : tileT;
    a.m = resident ? left : (left < m_tile ? left : m_tile);
    const bool last = (a.m == left);
    a.rmA = (last && ctx->remote) ? *ctx->remote : RemoteMap();  // a_m is final only in the last pass
    a.dstA = last ? out : S[(pass + 1) & 1];
    a.dstA_os = last ? int64_t(n) * inner : int64_t(h >> a.m) * inner;
    cudaError_t e = v2 ? launch_fwt_fwd_str2(ctx, w.L, w.de, a, resident) : cudaErrorNotSupported;
    if (e == cudaErrorNotSupported) e = launch_fwt_fwd_str(ctx, w.L, w.de, a, resident);
    JWC_TRY(e);
    a.src = a.dstA; a.src_os = a.dstA_os;
    h >>= a.m; left -= a.m; ++pass;
  }
  return cudaSuccess;
}

static cudaError_t fwt_reverse_strided(jwc_ctx* ctx, const WaveletRec& w, const double* in, double* out,
                                       int64_t outer, int n, int64_t inner, int level) {
  struct Pass { int h0, m; bool resident; };
  Pass passes[32];
  int npass = 0;
  size_t need[2] = {0, 0};
  const bool v2 = ctx->str_v2 && inner % 16 == 0;
  const int cap = v2 ? ctx->str2_cap : ctx->str_cap, tileT = v2 ? ctx->str2_rev_tile : ctx->str_rev_tile;
  // long filters are FP64-bound and every fused level adds a mostly idle step to each tile: fewer
  // levels per pass measured faster there (Daubechies20 columns: 2 levels 0.50, 5 levels 0.45)
  int rev_m = ctx->str_rev_m;
  const int cap_m = w.L >= 20 ? 2 : (w.L >= 12 ? 3 : rev_m);
  if (rev_m > cap_m) rev_m = cap_m;
  if (v2 && ctx->str2_rev_m > 0) rev_m = ctx->str2_rev_m;
  while ((tileT >> rev_m) < (v2 ? 8 : 4)) --rev_m;  // a tile keeps at least one group / one TMA box at its coarsest level
  int widths[32];
  int nw = 0;
  const int cur0 = n >> level;
  for (int wv = n; wv > cur0; wv >>= rev_m) {
    widths[nw++] = wv;
    if (wv <= cap || wv < tileT) break;  // produced by the resident pass
    if ((wv >> rev_m) <= cur0) break;
  }
  for (int i = nw - 1, cur = cur0; i >= 0; --i) {
    Pass p;
    p.h0 = widths[i];
    p.resident = (p.h0 <= cap) || (p.h0 < tileT);
    p.m = 0;
    while ((cur << p.m) < p.h0) ++p.m;
    if (p.h0 < n) {
      const size_t bytes = size_t(outer) * p.h0 * inner * sizeof(double);
      if (bytes > need[npass & 1]) need[npass & 1] = bytes;
    }
    passes[npass++] = p;
    cur = p.h0;
  }
  double* S[2] = {nullptr, nullptr};
  for (int i = 0; i < 2; ++i)
    if (need[i]) JWC_TRY(ensure_scratch(ctx, i, need[i], &S[i]));
  FwtRevStrArgs a;
  a.srcA = in; a.srcA_os = int64_t(n) * inner;
  a.srcD = in; a.srcD_os = int64_t(n) * inner;
  a.outer = outer; a.inner = inner;
  for (int i = 0; i < npass; ++i) {
    const Pass& p = passes[i];
    const bool last = (p.h0 == n);
    a.h0 = p.h0; a.m = p.m;
    a.T = p.resident ? p.h0 : tileT;
    a.dst = last ? out : S[i & 1];
    a.dst_os = last ? int64_t(n) * inner : int64_t(p.h0) * inner;
    a.rm = (last && ctx->remote) ? *ctx->remote : RemoteMap();
    cudaError_t e = v2 ? launch_fwt_rev_str2(ctx, w.L, w.re, a, p.resident) : cudaErrorNotSupported;
    if (e == cudaErrorNotSupported) e = launch_fwt_rev_str(ctx, w.L, w.re, a, p.resident);
    JWC_TRY(e);
    a.srcA = a.dst; a.srcA_os = a.dst_os;
  }
  return cudaSuccess;
}

// Fused plan: tile passes of m levels each while the width exceeds res_cap, then one resident
// launch for everything that is left.  a_m of a tile pass goes to a compact scratch buffer.
static cudaError_t fwt_forward(jwc_ctx* ctx, const WaveletRec& w, const double* in, double* out,
                               int64_t outer, int n, int64_t inner, int level) {
  if (inner > 1 && w.mirror_de && strided_ok(ctx, in, out, n, inner)) return fwt_forward_strided(ctx, w, in, out, outer, n, inner, level);
  if (ctx->remote) return cudaErrorNotSupported;  // peer stores exist in the strided forward kernels only
  if (!w.mirror_de || !fused_ok(ctx, in, out, n, inner)) return fwt_forward_generic(ctx, w, in, out, outer, n, inner, level);
  const int cap = ctx->res_cap;
  struct Pass { int h, T, m; bool resident, shfl; };
  Pass passes[32];
  int npass = 0;
  size_t need[2] = {0, 0};
  for (int h = n, left = level; left > 0;) {
    Pass p;
    p.h = h;
    p.resident = (h <= cap);
    p.T = p.resident ? h : (h < ctx->fwd_tile ? h : ctx->fwd_tile);
    // 2-tap filters: up to 8 levels per launch in registers and warp shuffles (jwc_shfl.cu) instead of a tile pass
    p.shfl = ctx->shfl && w.L == 2 && !p.resident && h % 256 == 0;
    if (p.shfl) {
      p.m = left < 8 ? left : 8;
      // a_m of a pass that is not the last one goes to compact lines of h >> m samples, which the next pass reads
      // with 32-byte loads: keep them at least 4 samples wide (h = 256, 9 levels to go: 6 + 3, not 8 + 1)
      while (p.m < left && (h >> p.m) < 4) --p.m;
    } else if (p.resident) {
      p.m = left;
      // deep tails: below ~32 samples per line a level keeps a fraction of the CTA's threads busy and still costs a
      // barrier; with res_split the resident work is two launches - down to res_split samples, then the rest on
      // (res_cap / res_split) times as many lines per CTA
      if (ctx->res_split >= 4 && h > ctx->res_split && (h >> left) < ctx->res_split) {
        p.m = 0;
        while ((h >> p.m) > ctx->res_split) ++p.m;
      }
    } else {
      int m_tile = fwt_tile_levels(w.L, p.T);
      if (ctx->fwd_m > 0 && ctx->fwd_m < m_tile) m_tile = ctx->fwd_m;
      p.m = left < m_tile ? left : m_tile;
    }
    if (p.m < left) {  // a_m of this pass goes to compact scratch lines
      const size_t bytes = size_t(outer) * (h >> p.m) * sizeof(double);
      if (bytes > need[(npass + 1) & 1]) need[(npass + 1) & 1] = bytes;
    }
    passes[npass++] = p;
    h >>= p.m; left -= p.m;
  }
  double* S[2] = {nullptr, nullptr};
  for (int i = 0; i < 2; ++i)
    if (need[i]) JWC_TRY(ensure_scratch(ctx, i, need[i], &S[i]));
  const int64_t pin = ctx->pitch_in ? ctx->pitch_in : n, pout = ctx->pitch_out ? ctx->pitch_out : n;
  FwtFwdArgs a;
  a.src = in; a.src_os = pin;
  a.dstD = out; a.dstD_os = pout;
  a.lines = outer;
  for (int i = 0; i < npass; ++i) {
    const Pass& p = passes[i];
    const bool last = (i == npass - 1);
    a.h = p.h; a.T = p.T; a.m = p.m;
    a.G = p.resident ? resident_lines(ctx, p.h, 150) : 1;   // fwd: (h/2 + h/4) double2, padded 1.25
    a.dstA = last ? out : S[(i + 1) & 1];
    a.dstA_os = last ? pout : (p.h >> p.m);
    cudaError_t e = p.shfl ? launch_fwt_fwd_shfl(ctx, w.L, w.de, a) : cudaErrorNotSupported;
    if (e == cudaErrorNotSupported) {
      if (p.shfl) {  // declined (alignment): the tile kernel takes the same pass if it can fuse that many levels
        int m_tile = fwt_tile_levels(w.L, p.T);
        if (p.m > m_tile) return cudaErrorInvalidValue;
      }
      e = launch_fwt_fwd(ctx, w.L, w.de, a, p.resident);
    }
    JWC_TRY(e);
    a.src = a.dstA; a.src_os = a.dstA_os;
  }
  return cudaSuccess;
}

static cudaError_t fwt_reverse_generic(jwc_ctx* ctx, const WaveletRec& w, const double* in, double* out,
                                       int64_t outer, int n, int64_t inner, int level) {
  const int64_t line = int64_t(n) * inner;
  double* S[2] = {nullptr, nullptr};
  if (level >= 2) JWC_TRY(ensure_scratch(ctx, 1, size_t(outer) * (n / 2) * inner * sizeof(double), &S[1]));
  if (level >= 3) JWC_TRY(ensure_scratch(ctx, 0, size_t(outer) * (n / 4) * inner * sizeof(double), &S[0]));
  // FastWaveletTransform.java:137-141: first width is 2 << (p - level)
  const double* srcA = in;
  int64_t srcA_os = line;
  for (int l = level; l >= 1; --l) {
    const int h = n >> (l - 1);  // width being rebuilt at this step
    const int half = h >> 1;
    const bool last = (l == 1);
    RevLevelArgs a;
    a.srcA = srcA; a.srcA_os = srcA_os;
    a.srcD = in + int64_t(half) * inner; a.srcD_os = line;
    // widths n/2, n/8, ... go to S[1] (n/2 per line), widths n/4, n/16, ... to S[0]
    a.dst = last ? out : S[(l + 1) & 1];
    a.dst_os = last ? line : int64_t(h) * inner;
    a.outer = outer; a.inner = inner; a.half = half;
    JWC_TRY(launch_rev_level_generic(ctx, w.L, w.re, a));
    srcA = a.dst; srcA_os = a.dst_os;
  }
  return cudaSuccess;
}

// Fused plan: one resident launch rebuilds everything up to width res_cap, then tile passes of up
// to rev_m levels each.  Intermediate approximations go to compact scratch lines.
static cudaError_t fwt_reverse(jwc_ctx* ctx, const WaveletRec& w, const double* in, double* out,
                               int64_t outer, int n, int64_t inner, int level) {
  if (inner > 1 && w.mirror_re && strided_ok(ctx, in, out, n, inner)) return fwt_reverse_strided(ctx, w, in, out, outer, n, inner, level);
  if (!w.mirror_re || !fused_ok(ctx, in, out, n, inner)) {
    if (ctx->remote) return cudaErrorNotSupported;  // the one-level kernels store locally only
    return fwt_reverse_generic(ctx, w, in, out, outer, n, inner, level);
  }
  struct Pass { int h0, m; bool resident; };
  Pass passes[32];
  int npass = 0;
  size_t need[2] = {0, 0};
  // Output widths of the passes, chosen backwards from n: every tile pass rebuilds
  // rev_m levels, so the resident pass ends at n >> (rev_m * tile passes) (a few hundred
  // samples per line, many lines per CTA) and the scratch round trip stays below 2 / 64 of the data.
  int widths[32];
  int nw = 0;
  const int cur0 = n >> level;
  const int cap = ctx->res_cap;
  for (int wv = n; wv > cur0;) {
    widths[nw++] = wv;
    if (wv <= cap) break;  // this one is produced by the resident pass
    const int T = wv < ctx->rev_tile ? wv : ctx->rev_tile;
    int rev_m = ctx->rev_m;
    while ((T >> rev_m) < 8) --rev_m;  // a tile keeps >= 8 slots at its coarsest level
    if ((wv >> rev_m) <= cur0) break;
    wv >>= rev_m;
  }
  for (int i = nw - 1, cur = cur0; i >= 0; --i) {
    Pass p;
    p.h0 = widths[i];
    p.resident = (p.h0 <= cap);
    p.m = 0;
    while ((cur << p.m) < p.h0) ++p.m;
    if (p.h0 < n) {
      const size_t bytes = size_t(outer) * p.h0 * sizeof(double);
      if (bytes > need[npass & 1]) need[npass & 1] = bytes;
    }
    passes[npass++] = p;
    cur = p.h0;
  }
  double* S[2] = {nullptr, nullptr};
  for (int i = 0; i < 2; ++i)
    if (need[i]) JWC_TRY(ensure_scratch(ctx, i, need[i], &S[i]));
  const int64_t pin = ctx->pitch_in ? ctx->pitch_in : n, pout = ctx->pitch_out ? ctx->pitch_out : n;
  FwtRevArgs a;
  a.srcA = in; a.srcA_os = pin;
  a.srcD = in; a.srcD_os = pin;
  a.lines = outer;
  for (int i = 0; i < npass; ++i) {
    const Pass& p = passes[i];
    const bool last = (p.h0 == n);
    a.h0 = p.h0; a.m = p.m;
    a.T = (p.resident || p.h0 < ctx->rev_tile) ? p.h0 : ctx->rev_tile;
    a.G = p.resident ? resident_lines(ctx, p.h0, 140) : 1;  // rev: (h + h/2 + h/4) samples, unpadded
    a.dst = last ? out : S[i & 1];
    a.dst_os = last ? pout : p.h0;
    a.rm = (last && ctx->remote) ? *ctx->remote : RemoteMap();
    JWC_TRY(launch_fwt_rev(ctx, w.L, w.re, a, p.resident));
    a.srcA = a.dst; a.srcA_os = a.dst_os;
  }
  return cudaSuccess;
}

// ---- WPT ---------------------------------------------------------------------------------------
// At width h a line holds n / h packets; packet (o, q) starts at o*n*inner + q*h*inner, i.e. the
// packets of all lines form `outer * n / h` lines of stride h * inner.

static cudaError_t wpt_forward_generic(jwc_ctx* ctx, const WaveletRec& w, const double* in, double* out,
                                       int64_t outer, int n, int64_t inner, int level) {
  double* S = nullptr;
  if (level >= 2) JWC_TRY(ensure_scratch(ctx, 0, size_t(outer) * n * inner * sizeof(double), &S));
  const double* src = in;
  int h = n;
  for (int l = 1; l <= level; ++l) {
    const int half = h >> 1;
    double* dst = ((level - l) & 1) ? S : out;
    FwdLevelArgs a;
    a.src = src; a.src_os = int64_t(h) * inner;
    a.dstA = dst; a.dstA_os = int64_t(h) * inner;
    a.dstD = dst + int64_t(half) * inner; a.dstD_os = int64_t(h) * inner;
    a.outer = outer * (n / h); a.inner = inner; a.half = half;
    JWC_TRY(launch_fwd_level_generic(ctx, w.L, w.de, a));
    src = dst; h = half;
  }
  return cudaSuccess;
}

static cudaError_t wpt_reverse_generic(jwc_ctx* ctx, const WaveletRec& w, const double* in, double* out,
                                       int64_t outer, int n, int64_t inner, int level) {
  double* S = nullptr;
  if (level >= 2) JWC_TRY(ensure_scratch(ctx, 0, size_t(outer) * n * inner * sizeof(double), &S));
  const double* src = in;
  for (int l = level; l >= 1; --l) {
    const int h = n >> (l - 1);
    const int half = h >> 1;
    double* dst = ((l - 1) & 1) ? S : out;
    RevLevelArgs a;
    a.srcA = src; a.srcA_os = int64_t(h) * inner;
    a.srcD = src + int64_t(half) * inner; a.srcD_os = int64_t(h) * inner;
    a.dst = dst; a.dst_os = int64_t(h) * inner;
    a.outer = outer * (n / h); a.inner = inner; a.half = half;
    JWC_TRY(launch_rev_level_generic(ctx, w.L, w.re, a));
    src = dst;
  }
  return cudaSuccess;
}

// Fused WPT plans.  A pass of m levels turns every packet of width h into 2^m packets of width
// h >> m, so the next pass sees 2^m times as many, shorter lines.  Whole-array ping-pong between
// `out` and one scratch buffer, phased so the last pass writes `out`.
static const size_t kWptSmemLimit = 112 * 1024;  // two CTAs per SM

// Packet transform along a STRIDED axis (inner > 1: matrix columns, the outer axes of a volume): transpose the
// [n][inner] planes so that the axis becomes contiguous, run the fused contiguous-line plan, transpose back
// (jwc_transpose.cu).  Buffers: `out` holds the transposed input, scratch[1] the transposed result; the fused plan
// ping-pongs between scratch[1] and scratch[0] as usual.
static bool wpt_transposed_ok(const jwc_ctx* ctx, bool mirror, const double* in, const double* out, int n, int64_t inner, int level) {
  return mirror && !ctx->force_generic && ctx->wpt_transpose && inner > 1 && inner <= 0x7fffffff && n >= 8 && level >= 2 &&
         aligned32(in) && aligned32(out);
}
static cudaError_t wpt_forward(jwc_ctx* ctx, const WaveletRec& w, const double* in, double* out,
                               int64_t outer, int n, int64_t inner, int level);
static cudaError_t wpt_reverse(jwc_ctx* ctx, const WaveletRec& w, const double* in, double* out,
                               int64_t outer, int n, int64_t inner, int level);
static cudaError_t wpt_transposed(jwc_ctx* ctx, const WaveletRec& w, int dir, const double* in, double* out,
                                  int64_t outer, int n, int64_t inner, int level) {
  double* X = nullptr;
  JWC_TRY(ensure_scratch(ctx, 1, size_t(outer) * n * inner * sizeof(double), &X));
  JWC_TRY(launch_transpose(ctx, in, out, outer, n, int(inner)));
  if (dir == JWC_FORWARD) JWC_TRY(wpt_forward(ctx, w, out, X, outer * inner, n, 1, level));
  else JWC_TRY(wpt_reverse(ctx, w, out, X, outer * inner, n, 1, level));
  return launch_transpose(ctx, X, out, outer, int(inner), n);
}

static cudaError_t wpt_forward(jwc_ctx* ctx, const WaveletRec& w, const double* in, double* out,
                               int64_t outer, int n, int64_t inner, int level) {
  if (wpt_transposed_ok(ctx, w.mirror_de, in, out, n, inner, level)) return wpt_transposed(ctx, w, JWC_FORWARD, in, out, outer, n, inner, level);
  if (!w.mirror_de || !fused_ok(ctx, in, out, n, inner) || n < 8) return wpt_forward_generic(ctx, w, in, out, outer, n, inner, level);
  struct Pass { int h, m; bool resident; };
  Pass passes[32];
  int npass = 0;
  for (int h = n, left = level; left > 0;) {
    Pass p;
    p.h = h;
    p.resident = (h <= ctx->res_cap);
    p.m = p.resident ? left : wpt_tile_levels(w.L, h < ctx->wpt_tile ? h : ctx->wpt_tile, left < ctx->wpt_m ? left : ctx->wpt_m, kWptSmemLimit, ctx->wpt_r);
    passes[npass++] = p;
    h >>= p.m; left -= p.m;
  }
  double* S = nullptr;
  if (npass >= 2) JWC_TRY(ensure_scratch(ctx, 0, size_t(outer) * n * sizeof(double), &S));
  const double* src = in;
  for (int i = 0; i < npass; ++i) {
    const Pass& p = passes[i];
    WptFwdArgs a;
    a.src = src; a.src_os = p.h;
    a.dst = ((npass - 1 - i) & 1) ? S : out; a.dst_os = p.h;
    a.lines = outer * (n / p.h);
    a.h = p.h; a.m = p.m;
    a.T = (p.resident || p.h < ctx->wpt_tile) ? p.h : ctx->wpt_tile;
    a.G = p.resident ? resident_lines(ctx, p.h, 200) : 1;   // wpt: two full line buffers, padded 1.25
    JWC_TRY(launch_wpt_fwd(ctx, w.L, w.de, a, p.resident));
    src = a.dst;
  }
  return cudaSuccess;
}

static cudaError_t wpt_reverse(jwc_ctx* ctx, const WaveletRec& w, const double* in, double* out,
                               int64_t outer, int n, int64_t inner, int level) {
  if (wpt_transposed_ok(ctx, w.mirror_re, in, out, n, inner, level)) return wpt_transposed(ctx, w, JWC_REVERSE, in, out, outer, n, inner, level);
  if (!w.mirror_re || !fused_ok(ctx, in, out, n, inner) || n < 8) return wpt_reverse_generic(ctx, w, in, out, outer, n, inner, level);
  // Output widths of the passes, chosen backwards from n: each tile pass rebuilds as many levels
  // as its shared memory allows; whatever is left below res_cap is one resident pass.
  struct Pass { int h0, m; bool resident; };
  Pass passes[32];
  int npass = 0;
  const int cur0 = n >> level;
  int widths[32];
  int nw = 0;
  for (int wv = n; wv > cur0;) {
    widths[nw++] = wv;
    if (wv <= ctx->res_cap) break;  // produced by the resident pass
    int want = 0;
    while ((cur0 << want) < wv) ++want;
    const int cap_m = ctx->wpt_rev_m > 0 ? ctx->wpt_rev_m : ctx->wpt_m;
    if (want > cap_m) want = cap_m;
    wv >>= wpt_rev_tile_levels(w.L, wv < ctx->wpt_tile ? wv : ctx->wpt_tile, want, kWptSmemLimit);
  }
  for (int i = nw - 1, cur = cur0; i >= 0; --i) {
    Pass p;
    p.h0 = widths[i];
    p.resident = (p.h0 <= ctx->res_cap);
    p.m = 0;
    while ((cur << p.m) < p.h0) ++p.m;
    passes[npass++] = p;
    cur = p.h0;
  }
  double* S = nullptr;
  if (npass >= 2) JWC_TRY(ensure_scratch(ctx, 0, size_t(outer) * n * sizeof(double), &S));
  const double* src = in;
  for (int i = 0; i < npass; ++i) {
    const Pass& p = passes[i];
    WptRevArgs a;
    a.src = src; a.src_os = p.h0;
    a.dst = ((npass - 1 - i) & 1) ? S : out; a.dst_os = p.h0;
    a.lines = outer * (n / p.h0);
    a.h0 = p.h0; a.m = p.m;
    a.T = (p.resident || p.h0 < ctx->wpt_tile) ? p.h0 : ctx->wpt_tile;
    a.G = p.resident ? resident_lines(ctx, p.h0, 200) : 1;
    JWC_TRY(launch_wpt_rev(ctx, w.L, w.re, a, p.resident));
    src = a.dst;
  }
  return cudaSuccess;
}

bool fwt_pitched_ok(const jwc_ctx* ctx, const WaveletRec& w, int dir, const double* in, const double* out, int n,
                    int64_t pitch_in, int64_t pitch_out) {
  return (dir == JWC_FORWARD ? w.mirror_de : w.mirror_re) && !ctx->remote && fused_ok(ctx, in, out, n, 1) &&
         pitch_in % 4 == 0 && pitch_out % 4 == 0;
}

cudaError_t run_axis(jwc_ctx* ctx, const WaveletRec& w, int kind, int dir, const double* in, double* out,
                     int64_t outer, int n, int64_t inner, int level) {
  if (ctx->remote && (level == 0 || n == 1 || kind != JWC_FWT)) return cudaErrorNotSupported;
  if (level == 0 || n == 1) return copy_through(ctx, in, out, outer * n * inner);
  if (kind == JWC_FWT)
    return dir == JWC_FORWARD ? fwt_forward(ctx, w, in, out, outer, n, inner, level)
                              : fwt_reverse(ctx, w, in, out, outer, n, inner, level);
  return dir == JWC_FORWARD ? wpt_forward(ctx, w, in, out, outer, n, inner, level)
                            : wpt_reverse(ctx, w, in, out, outer, n, inner, level);
}

}  // namespace jwc
