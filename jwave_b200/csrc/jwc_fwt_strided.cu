// jwc_fwt_strided.cu - fused multi-level FWT along a STRIDED axis (matrix columns, the two outer
// axes of a volume): the column loop of BasicTransform.forward/reverse(double[][], ...)
// (BasicTransform.java:383-395, :444-456) and the outer-axis loop of the 3-D driver (:546-562,
// :639-655), which in the reference gather every column into a temporary array, run
// FastWaveletTransform on it (FastWaveletTransform.java:88-97, :143-149) and scatter it back.
//
// Here a CTA owns kC = 8 adjacent lines (64 contiguous bytes per sample) and a run of samples
// along the axis, staged as [sample][8] in shared memory; no gather, no transpose, and `m` levels
// per launch.  Tile / resident modes and halo arithmetic are those of jwc_fwt_fwd.cu / jwc_fwt_rev.cu.
#include <cuda.h>

#include <cstring>

#include "jwc_kernels.cuh"
#include "jwc_strided.cuh"

namespace jwc {

// ================================ forward ======================================================

// Stage `rows` level-0 rows (first global row `y0` of the tensor, column x0) with TMA: whole boxes of
// kBoxRows rows; a tile that runs past the end of its line wraps to the line's first row (the
// periodic halo), and the wrap point is always a multiple of the box height.
__device__ __forceinline__ void tma_stage(double* buf, uint64_t* bar, const void* tmap, int x0, int64_t line_row0,
                                          int first, int rows, int h) {
  if (threadIdx.x == 0) {
    const int boxes = (rows + kBoxRows - 1) / kBoxRows;
    mbar_expect_tx(bar, unsigned(boxes) * kBoxRows * kC * sizeof(double));
    for (int b = 0; b < boxes; ++b) {
      const int s = (first + b * kBoxRows) & (h - 1);   // sample index inside the line
      tma_load_box(buf + b * kBoxRows * kC, tmap, x0, int(line_row0 + s), bar);
    }
  }
  mbar_wait(bar, 0);
}

template <int L, bool TMA0, class Store>
__device__ __forceinline__ void fwd_str_level_tile(const Taps& taps, const double* cur, int groups, int c, int g0,
                                                   int gpp, Store store) {
  for (int g = g0; g < groups; g += gpp) {
    double lo[kSR], hi[kSR];
    if constexpr (TMA0) fwd_run<L, kSR>(taps, tap_phase<L, false>(), [&](int s) { return tma_at(cur, 2 * kSR * g + s, c); }, lo, hi);
    else fwd_run<L, kSR>(taps, tap_phase<L, false>(), [&](int s) { return sat(cur, 2 * kSR * g + s, c); }, lo, hi);
    store(g, lo, hi);
  }
}

template <int L, bool RESIDENT, bool TMA>
__global__ void __launch_bounds__(kThreads)
k_fwt_fwd_str(const __grid_constant__ Taps taps, const __grid_constant__ FwtFwdStrArgs a,
              const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) double smem[];
  const int c = threadIdx.x % kC, g0 = threadIdx.x / kC;
  const int kGroupsPerPass = blockDim.x / kC;
  const int m = a.m, h = a.h;
  int64_t b = blockIdx.x;
  const int cb = int(b % a.cblocks); b /= a.cblocks;
  const int tile = RESIDENT ? 0 : int(b % a.tiles_per_line);
  const int64_t o = RESIDENT ? b : b / a.tiles_per_line;
  const int64_t inner = a.inner;
  const double* src = a.src + o * a.src_os + cb * kC;
  double* gD = a.dstD + o * a.dstD_os + cb * kC + c;
  double* gA = a.dstA + o * a.dstA_os + cb * kC + c;
  // output row -> address: local lines, or the peers' slabs (RemoteMap, jwc_internal.cuh)
  auto pD = [&](int64_t row) { return a.rmD.mode ? remote_row(a.rmD, o, row) + cb * kC + c : gD + row * inner; };
  auto pA = [&](int64_t row) { return a.rmA.mode ? remote_row(a.rmA, o, row) + cb * kC + c : gA + row * inner; };
  double* cur = smem;
  double* nxt = smem + a.rows0 * kC;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (a.rows0 + a.rows1) * kC);
  if constexpr (TMA) {
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
  }

  if constexpr (!RESIDENT) {
    const int T = a.T;
    const int n0 = T + ((1 << m) - 1) * (L - 2);
    if constexpr (TMA) {
      tma_stage(cur, bar, &tmap, cb * kC, o * a.rows_per_o, tile * T, n0, h);
    } else {
      stage_rows(cur, src, inner, tile * T, n0, h - 1);
      cp_async_wait_all();
      __syncthreads();
    }
    for (int k = 1; k <= m; ++k) {
      const int n_det = T >> k;
      const int n_out = n_det + ((1 << (m - k)) - 1) * (L - 2);
      const int groups = (n_out + kSR - 1) / kSR;
      const bool last = (k == m);
      const int64_t rowD = (h >> k) + tile * n_det, rowA = tile * n_det;
      auto store = [&](int g, const double (&lo)[kSR], const double (&hi)[kSR]) {
        const bool keep = kSR * g < n_det;
#pragma unroll
        for (int r = 0; r < kSR; ++r) {
          if (!last) sat(nxt, kSR * g + r, c) = lo[r];
          else if (keep) *pA(rowA + kSR * g + r) = lo[r];
          if (keep) *pD(rowD + kSR * g + r) = hi[r];
        }
      };
      if (TMA && k == 1) fwd_str_level_tile<L, TMA>(taps, cur, groups, c, g0, kGroupsPerPass, store);
      else fwd_str_level_tile<L, false>(taps, cur, groups, c, g0, kGroupsPerPass, store);
      __syncthreads();
      double* t = cur; cur = nxt; nxt = t;
    }
  } else {
    if constexpr (TMA) {
      tma_stage(cur, bar, &tmap, cb * kC, o * a.rows_per_o, 0, h, h);
    } else {
      stage_rows(cur, src, inner, 0, h, h - 1);
      cp_async_wait_all();
      __syncthreads();
    }
    for (int k = 1; k <= m; ++k) {
      const int h_in = h >> (k - 1), h_out = h_in >> 1;
      const int mask = h_in - 1;
      const bool last = (k == m);
      const bool tma0 = TMA && k == 1;
      if (h_out >= kSR) {
        for (int g = g0; g < h_out / kSR; g += kGroupsPerPass) {
          double lo[kSR], hi[kSR];
          if (tma0) fwd_run<L, kSR>(taps, tap_phase<L, false>(), [&](int s) { return tma_at(cur, (2 * kSR * g + s) & mask, c); }, lo, hi);
          else fwd_run<L, kSR>(taps, tap_phase<L, false>(), [&](int s) { return sat(cur, (2 * kSR * g + s) & mask, c); }, lo, hi);
#pragma unroll
          for (int r = 0; r < kSR; ++r) {
            if (!last) sat(nxt, kSR * g + r, c) = lo[r];
            else *pA(kSR * g + r) = lo[r];
            *pD(h_out + kSR * g + r) = hi[r];
          }
        }
      } else {
        // columns of 2 or 4 samples: one thread per output, true modular indexing (h < L wraps)
        for (int i = g0; i < h_out; i += kGroupsPerPass) {
          double lo = 0.0, hi = 0.0;
          const int z = tap_phase<L, false>();
#pragma unroll
          for (int j = 0; j < L; ++j) {
            const double v = sat(cur, (2 * i + j) & mask, c);   // never the TMA buffer: h >= kBoxRows there
            lo = fma(v, lo_tap<L>(taps, j, z), lo);
            hi = fma(v, hi_tap<L>(taps, j, z), hi);
          }
          if (!last) sat(nxt, i, c) = lo;
          else *pA(i) = lo;
          *pD(h_out + i) = hi;
        }
      }
      __syncthreads();
      double* t = cur; cur = nxt; nxt = t;
    }
  }
}

// ================================ reverse ======================================================
// Launch levels as in jwc_fwt_rev.cu: level 0 = output (width h0), level m = coarsest input a_m;
// d_k sits at sample (h0 >> k) of every coefficient line.

template <int L, bool RESIDENT>
__global__ void __launch_bounds__(kThreads)
k_fwt_rev_str(const __grid_constant__ Taps taps, const __grid_constant__ FwtRevStrArgs a) {
  extern __shared__ double smem[];
  const int c = threadIdx.x % kC, g0 = threadIdx.x / kC;
  const int kGroupsPerPass = blockDim.x / kC;
  const int m = a.m, h0 = a.h0;
  int64_t b = blockIdx.x;
  const int cb = int(b % a.cblocks); b /= a.cblocks;
  const int tile = RESIDENT ? 0 : int(b % a.tiles_per_line);
  const int64_t o = RESIDENT ? b : b / a.tiles_per_line;
  const int64_t inner = a.inner;
  const double* lineD = a.srcD + o * a.srcD_os + cb * kC;
  const double* lineA = a.srcA + o * a.srcA_os + cb * kC;
  double* gY = a.dst + o * a.dst_os + cb * kC + c;
  auto pY = [&](int64_t row) { return a.rm.mode ? remote_row(a.rm, o, row) + cb * kC + c : gY + row * inner; };

  if constexpr (!RESIDENT) {
    const int T = a.T, t0 = tile * T;
    // stage d_k (k = 1..m) and a_m; local row j of level k is slot O_k + j (periodic)
    for (int k = 1; k <= m; ++k) {
      const int wk = h0 >> k;
      const int O = (k == m) ? ((t0 >> k) - a.F[k] - a.ru) : 2 * ((t0 >> (k + 1)) - a.F[k + 1]);
      stage_rows(smem + a.offD[k], lineD + int64_t(wk) * inner, inner, O, a.len[k], wk - 1);
      if (k == m) stage_rows(smem + a.offA[m & 1], lineA, inner, O, a.len[k], wk - 1);
    }
    cp_async_wait_all();
    __syncthreads();
    for (int k = m; k >= 1; --k) {
      const double* A = smem + a.offA[k & 1];
      const double* D = smem + a.offD[k];
      double* Y = smem + a.offA[(k - 1) & 1];
      const int groups = ((T >> k) + a.F[k]) / kSR;
      const int s0 = a.s0[k];  // local row of the level's first slot
      for (int g = g0; g < groups; g += kGroupsPerPass) {
        const int top = s0 + kSR * g + kSR - 1;
        double t[2 * kSR];
        rev_run<L, kSR>(taps, tap_phase<L, true>(), [&](int s) { return sat(A, top - s, c); }, [&](int s) { return sat(D, top - s, c); }, t);
#pragma unroll
        for (int e = 0; e < 2 * kSR; ++e) {
          if (k > 1) sat(Y, 2 * kSR * g + e, c) = t[e];
          else *pY(t0 + 2 * kSR * g + e) = t[e];
        }
      }
      __syncthreads();
    }
  } else {
    // resident: C = coefficient prefix rows [0, h0) (a_m in rows [0, h0 >> m)); P[k & 1] = a_k
    double* C = smem;
    double* P[2] = {smem + a.rowsC * kC, smem + (a.rowsC + a.rowsP[0]) * kC};
    stage_rows(C, lineD, inner, 0, h0, h0 - 1);
    cp_async_wait_all();
    __syncthreads();
    for (int k = m; k >= 1; --k) {
      const int half = h0 >> k;
      const int mask = half - 1;
      const double* A = (k == m) ? C : P[k & 1];
      double* Y = P[(k - 1) & 1];
      const bool last = (k == 1);
      if (half >= kSR) {
        for (int g = g0; g < half / kSR; g += kGroupsPerPass) {
          const int top = kSR * g + kSR - 1;
          double t[2 * kSR];
          rev_run<L, kSR>(taps, tap_phase<L, true>(), [&](int s) { return sat(A, (top - s) & mask, c); },
                          [&](int s) { return sat(C, half + ((top - s) & mask), c); }, t);
#pragma unroll
          for (int e = 0; e < 2 * kSR; ++e) {
            if (!last) sat(Y, 2 * kSR * g + e, c) = t[e];
            else *pY(2 * kSR * g + e) = t[e];
          }
        }
      } else {
        for (int p = g0; p < half; p += kGroupsPerPass) {
          double t0v = 0.0, t1v = 0.0;
          const int z = tap_phase<L, true>();
#pragma unroll
          for (int q = 0; q < L / 2; ++q) {
            const int i = (p - q) & mask;
            const double av = sat(A, i, c), dv = sat(C, half + i, c);
            t0v = fma(av, lo_tap<L>(taps, 2 * q, z), t0v);
            t0v = fma(dv, hi_tap<L>(taps, 2 * q, z), t0v);
            t1v = fma(av, lo_tap<L>(taps, 2 * q + 1, z), t1v);
            t1v = fma(dv, hi_tap<L>(taps, 2 * q + 1, z), t1v);
          }
          if (!last) { sat(Y, 2 * p, c) = t0v; sat(Y, 2 * p + 1, c) = t1v; }
          else { *pY(2 * p) = t0v; *pY(2 * p + 1) = t1v; }
        }
      }
      __syncthreads();
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------

static int round_up(int v, int q) { return (v + q - 1) / q * q; }

int fwt_str_tile_levels(int L, int T) {
  // largest m with halo (2^m - 1)(L - 2) <= T / 4 and T / 2^m >= kSR
  int m = 1;
  while (((1 << (m + 1)) - 1) * (L - 2) <= T / 4 && (T >> (m + 1)) >= kSR) ++m;
  return m;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// 2-D map of the source as [rows][inner] doubles, box {8 columns, kBoxRows rows}, dense in shared memory
static bool make_tmap(CUtensorMap* map, const double* base, int64_t rows, int64_t inner) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc || (reinterpret_cast<uintptr_t>(base) & 15) || (inner * sizeof(double)) % 16) return false;
  const cuuint64_t dims[2] = {cuuint64_t(inner), cuuint64_t(rows)};
  const cuuint64_t strides[1] = {cuuint64_t(inner) * sizeof(double)};
  const cuuint32_t box[2] = {kC, kBoxRows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int L>
static cudaError_t launch_fwd_L(jwc_ctx* ctx, const Taps& taps, FwtFwdStrArgs a, bool resident) {
  if (a.inner % kC) return cudaErrorInvalidValue;
  a.cblocks = int(a.inner / kC);
  int64_t grid;
  // TMA needs whole boxes: tiles and resident lines of at least kBoxRows rows
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  a.rows_per_o = a.src_os / a.inner;
  bool tma = ctx->str_tma && (resident ? a.h : a.T) >= kBoxRows && a.src_os % a.inner == 0 &&
             make_tmap(&tmap, a.src, a.outer * a.rows_per_o, a.inner);
  if (!resident) {
    if ((a.T >> a.m) < kSR) return cudaErrorInvalidValue;
    const int n0 = a.T + ((1 << a.m) - 1) * (L - 2);
    a.rows0 = tma ? (n0 + kBoxRows - 1) / kBoxRows * kBoxRows : n0 + 8;   // whole boxes land in shared memory
    a.rows1 = (a.T >> 1) + ((1 << (a.m - 1)) - 1) * (L - 2) + 8;
    a.tiles_per_line = a.h / a.T;
    grid = a.outer * a.tiles_per_line * a.cblocks;
  } else {
    a.rows0 = tma ? (a.h + kBoxRows - 1) / kBoxRows * kBoxRows : a.h + 2;
    a.rows1 = max(2, a.h / 2) + 2;
    a.tiles_per_line = 1;
    grid = a.outer * a.cblocks;
  }
  const size_t smem = size_t(a.rows0 + a.rows1) * kC * sizeof(double) + 16;  // + the mbarrier
  if (grid > 0x7fffffff) return cudaErrorInvalidConfiguration;
  void (*kern)(const Taps, const FwtFwdStrArgs, const CUtensorMap) =
      resident ? (tma ? k_fwt_fwd_str<L, true, true> : k_fwt_fwd_str<L, true, false>)
               : (tma ? k_fwt_fwd_str<L, false, true> : k_fwt_fwd_str<L, false, false>);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
  }
  prof_begin(ctx, resident ? "k_fwt_fwd_str:resident" : "k_fwt_fwd_str:tile", double(a.outer) * a.h * a.inner, a.m);
  kern<<<int(grid), ctx->str_threads, smem, ctx->stream>>>(taps, a, tmap);
  prof_end(ctx);
  ctx->launches++;
  return cudaGetLastError();
}

cudaError_t launch_fwt_fwd_str(jwc_ctx* ctx, int L, const Taps& taps, const FwtFwdStrArgs& a, bool resident) {
  switch (L) {
#define JWC_CASE(LL) case LL: return launch_fwd_L<LL>(ctx, taps, a, resident);
    JWC_FOR_EACH_L(JWC_CASE)
#undef JWC_CASE
  }
  return cudaErrorInvalidValue;
}

template <int L>
static cudaError_t launch_rev_L(jwc_ctx* ctx, const Taps& taps, FwtRevStrArgs a, bool resident) {
  if (a.inner % kC) return cudaErrorInvalidValue;
  a.cblocks = int(a.inner / kC);
  int64_t grid;
  size_t smem;
  if (!resident) {
    if (a.m < 1 || a.m > kMaxFuse || (a.T >> a.m) < kSR) return cudaErrorInvalidValue;
    a.ru = round_up(L / 2 - 1, kSR);
    int N = 0;  // left extension of a_{k-1} that level k-1 needs
    for (int k = 1; k <= a.m; ++k) {
      a.F[k] = round_up((N + 1) / 2, kSR);
      N = a.F[k] + L / 2 - 1;
    }
    a.F[a.m + 1] = 0;
    int off = 0, capA[2] = {0, 0};
    for (int k = 1; k <= a.m; ++k) {
      a.len[k] = (k == a.m) ? (a.T >> k) + a.F[k] + a.ru : (a.T >> k) + 2 * a.F[k + 1];
      a.s0[k] = (k == a.m) ? a.ru : 2 * a.F[k + 1] - a.F[k];
      a.offD[k] = off * kC;
      off += a.len[k] + 2;
      if (a.len[k] + 2 > capA[k & 1]) capA[k & 1] = a.len[k] + 2;
    }
    a.offA[0] = off * kC;
    a.offA[1] = (off + capA[0]) * kC;
    smem = size_t(off + capA[0] + capA[1]) * kC * sizeof(double);
    a.tiles_per_line = a.h0 / a.T;
    grid = a.outer * a.tiles_per_line * a.cblocks;
  } else {
    a.rowsC = a.h0 + 2;
    a.rowsP[1] = max(2, a.h0 / 2) + 2;  // a_1
    a.rowsP[0] = max(2, a.h0 / 4) + 2;  // a_2
    smem = size_t(a.rowsC + a.rowsP[0] + a.rowsP[1]) * kC * sizeof(double);
    a.tiles_per_line = 1;
    grid = a.outer * a.cblocks;
  }
  if (grid > 0x7fffffff) return cudaErrorInvalidConfiguration;
  auto kern = resident ? k_fwt_rev_str<L, true> : k_fwt_rev_str<L, false>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
  }
  prof_begin(ctx, resident ? "k_fwt_rev_str:resident" : "k_fwt_rev_str:tile", double(a.outer) * a.h0 * a.inner, a.m);
  kern<<<int(grid), ctx->str_rev_threads, smem, ctx->stream>>>(taps, a);
  prof_end(ctx);
  ctx->launches++;
  return cudaGetLastError();
}

cudaError_t launch_fwt_rev_str(jwc_ctx* ctx, int L, const Taps& taps, const FwtRevStrArgs& a, bool resident) {
  switch (L) {
#define JWC_CASE(LL) case LL: return launch_rev_L<LL>(ctx, taps, a, resident);
    JWC_FOR_EACH_L(JWC_CASE)
#undef JWC_CASE
  }
  return cudaErrorInvalidValue;
}

}  // namespace jwc
