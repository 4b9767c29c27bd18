// jwc_fwt_strided2.cu - second-generation fused multi-level FWT along a STRIDED axis (matrix columns, the
// two outer axes of a volume): the column loop of BasicTransform.forward/reverse(double[][], ...)
// (BasicTransform.java:383-395, :444-456) and the outer-axis loop of the 3-D driver (:546-562, :639-655),
// which in the reference gather every column into a temporary array, run FastWaveletTransform on it
// (FastWaveletTransform.java:88-97, :143-149) and scatter it back.
//
// Layout, thread mapping and the reasons for them: jwc_strided2.cuh.  In short: CTA = 16 adjacent lines x
// a run of rows, dense [row][16] shared-memory tile staged by TMA boxes (forward AND reverse), two columns
// per thread (LDS.128 / STG.128, conflict-free), levels computed in place.  Arithmetic, level structure,
// halo rules and output layout are those of the first generation (jwc_fwt_strided.cu), which stays as the
// path for inner % 16 != 0 and for the shapes the launchers below decline (cudaErrorNotSupported).
#include <cuda.h>

#include <cstring>
#include <type_traits>

#include "jwc_kernels.cuh"
#include "jwc_strided2.cuh"

namespace jwc {

// ================================ forward ======================================================
// Tile mode: rows [tile T, tile T + n0) of the line (periodic), n0 = T + (2^m - 1)(L - 2).  Level k keeps
// n_det = T >> k outputs per filter (d_k rows are final, a_k rows feed level k + 1) and owes the levels
// below a halo of (2^(m-k) - 1)(L - 2) low-pass outputs.  Resident mode: the whole line (h rows) plus its
// periodic extension of L - 2 rows, and all remaining levels; after every level the extension of the new
// approximation is copied behind it, so both modes run the same wrap-free window code.
//
// A task = kR2 consecutive output rows of one column pair; halo tasks skip the high pass.  Levels are
// computed IN PLACE in rounds of blockDim / 8 tasks, lowest rows first: a round computes, waits at a
// barrier until every window of the round has been read, and then overwrites rows [R g, R g + R) - rows
// that later rounds never read (their windows start at row 2 R g' >= 2 R ngrp).  d_k rows and the last
// level's a_m rows go straight from registers to global memory (16-byte stores, 8 lanes = one 128-byte line).
// Columns of 2 or 4 samples (resident mode) take a one-row-per-thread path with true modular indexing
// (h < L wraps several times, Wavelet.java:248-249).
template <int L, bool RESIDENT>
__global__ void __launch_bounds__(RESIDENT ? kMaxThrRes2 : kMaxThr2, 2)
k_fwt_fwd_str2(const __grid_constant__ Taps taps, const __grid_constant__ FwtFwdStrArgs a,
               const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) double2 sm2[];
  constexpr int R = kR2;
  double2* X = sm2;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm2 + size_t(a.rows0) * kL8);
  const int tid = threadIdx.x, l8 = tid & (kL8 - 1), grp = tid >> 3, ngrp = blockDim.x >> 3;
  const int m = a.m, h = a.h;
  int64_t b = blockIdx.x;
  const int cb = int(b % a.cblocks); b /= a.cblocks;
  const int tile = RESIDENT ? 0 : int(b % a.tiles_per_line);
  const int64_t o = RESIDENT ? b : b / a.tiles_per_line;
  const int64_t inner = a.inner;
  double* gD = a.dstD + o * a.dstD_os + cb * kC2 + 2 * l8;
  double* gA = a.dstA + o * a.dstA_os + cb * kC2 + 2 * l8;
  // output row -> address: local lines, or the peers' slabs (RemoteMap, jwc_internal.cuh)
  auto pD = [&](int64_t row) { return a.rmD.mode ? remote_row(a.rmD, o, row) + cb * kC2 + 2 * l8 : gD + row * inner; };
  auto pA = [&](int64_t row) { return a.rmA.mode ? remote_row(a.rmA, o, row) + cb * kC2 + 2 * l8 : gA + row * inner; };
  const int T = RESIDENT ? h : a.T;
  const int n0 = RESIDENT ? h + (L - 2) : T + ((1 << m) - 1) * (L - 2);

  // ---- stage: TMA boxes {16 columns, kBoxF rows}; rows past the end of the line wrap to its first row,
  // always at a box boundary (T and h are multiples of the box height).  The boxes are counted on up to
  // kStagesF mbarriers in row order: round r of the first level waits only for the rows ITS windows read
  // (stage min(r, kStagesF - 1)), so a CTA computes from the first third of its tile while the rest is in flight.
  const int boxes = (n0 + kBoxF - 1) / kBoxF;
  const int rounds1 = a.rounds1;  // rounds of the first level
  auto stage_end = [&](int s) {   // first box NOT in stages 0 .. s
    if (s >= kStagesF - 1 || s >= rounds1 - 1) return boxes;
    return min(boxes, (2 * R * (s + 1) * ngrp + L - 2 + kBoxF - 1) / kBoxF);
  };
  if (tid == 0) {
    for (int s = 0; s < kStagesF; ++s) mbar_init(&bar[s], 1);
  }
  __syncthreads();
  if (tid == 0) {
    for (int s = 0, b0 = 0; s < kStagesF; ++s) {
      const int b1 = stage_end(s);
      mbar_expect_tx(&bar[s], unsigned(b1 - b0) * kBoxF * kC2 * sizeof(double));
      b0 = b1;
    }
  }
  if ((tid & 31) == 0) {  // one issuing lane per warp
    const int64_t row0 = o * a.rows_per_o;
    for (int bx = tid >> 5; bx < boxes; bx += blockDim.x >> 5) {
      int st = 0;
      while (bx >= stage_end(st)) ++st;
      const int s = (tile * T + bx * kBoxF) & (h - 1);
      tma_load_box(X + size_t(bx) * kBoxF * kL8, &tmap, cb * kC2, int(row0 + s), &bar[st]);
    }
  }

  const int z = tap_phase<L, true>();
  for (int k = 1; k <= m; ++k) {
    const bool last = (k == m);
    const int h_in = h >> (k - 1);  // resident: width of the level's input
    const int n_det = RESIDENT ? (h_in >> 1) : (T >> k);
    const int64_t rowD = RESIDENT ? n_det : (h >> k) + int64_t(tile) * n_det;
    const int64_t rowA = RESIDENT ? 0 : int64_t(tile) * n_det;
    if (!RESIDENT || n_det >= R) {
      // RR = output rows of one task.  Tile mode: always kR2.  Resident mode: the largest of 4, 2, 1 that still
      // gives every thread group a task - the short levels at the end of a full-depth transform would otherwise
      // keep a few threads busy for a whole 4-row task each while the rest of the CTA waits at the barrier
      // (9 task times for 4 of work on a 256-row line; the resident passes were 27 % of the 1024^3 step).
      auto level_rounds = [&](auto rc) {
        constexpr int RR = decltype(rc)::value;
        const int gkeep = n_det / RR;
        const int groups = RESIDENT ? gkeep : (n_det + ((1 << (m - k)) - 1) * (L - 2) + RR - 1) / RR;
        for (int g0 = 0, round = 0; g0 < groups; g0 += ngrp, ++round) {
          if (k == 1) mbar_wait(&bar[min(round, kStagesF - 1)], 0);
          const int g = g0 + grp;
          const bool has = g < groups;
          double2 lo[RR], hi[RR];
          if (has) {
            const double2* w = X + (2 * RR * g) * kL8 + l8;
            if (g < gkeep) {
              fwd_run2<L, RR, true>(taps, z, [&](int s) { return w[s * kL8]; }, lo, hi);
#pragma unroll
              for (int r = 0; r < RR; ++r) st2(pD(rowD + RR * g + r), hi[r]);
              if (last) {
#pragma unroll
                for (int r = 0; r < RR; ++r) st2(pA(rowA + RR * g + r), lo[r]);
              }
            } else {  // halo group (tile mode, never at the last level): low pass only
              fwd_run2<L, RR, false>(taps, z, [&](int s) { return w[s * kL8]; }, lo, hi);
            }
          }
          if (last) continue;
          __syncthreads();  // every window of this round (and of the rounds before it) has been read
          if (has) {
#pragma unroll
            for (int r = 0; r < RR; ++r) X[(RR * g + r) * kL8 + l8] = lo[r];
          }
        }
      };
      if constexpr (RESIDENT) {
        if (n_det >= 4 * ngrp) level_rounds(std::integral_constant<int, 4>{});
        else if (n_det >= 2 * ngrp) level_rounds(std::integral_constant<int, 2>{});
        else level_rounds(std::integral_constant<int, 1>{});
      } else {
        level_rounds(std::integral_constant<int, R>{});
      }
    } else {
      // resident, h_in = 2 or 4: one output row per thread group, true modular wrap
      if (k == 1) mbar_wait(&bar[kStagesF - 1], 0);
      const int mask = h_in - 1;
      double2 lo = make_double2(0.0, 0.0), hi = lo;
      const bool has = grp < n_det;
      if (has) {
#pragma unroll 1
        for (int j = 0; j < L; ++j) {
          const double2 v = X[((2 * grp + j) & mask) * kL8 + l8];
          fma2(lo, v, lo_tap<L>(taps, j, z));
          fma2(hi, v, hi_tap<L>(taps, j, z));
        }
        st2(pD(n_det + grp), hi);
        if (last) st2(pA(grp), lo);
      }
      if (!last) {
        __syncthreads();
        if (has) X[grp * kL8 + l8] = lo;
      }
    }
    if (last) break;
    __syncthreads();
    if constexpr (RESIDENT) {
      // periodic extension of a_k (width n_det) for the next level's windows
      if (n_det >= 2 * R) {
        for (int i = tid; i < (L - 2) * kL8; i += blockDim.x) {
          const int row = i >> 3, c = i & 7;
          X[(n_det + row) * kL8 + c] = X[(row & (n_det - 1)) * kL8 + c];
        }
        __syncthreads();
      }
    }
  }
}

// ================================ reverse ======================================================
// Launch levels as in jwc_fwt_rev.cu: level 0 = output (width h0), level m = coarsest input a_m; d_k sits
// at row (h0 >> k) of every coefficient line.  Tile mode (one CTA = T output rows): shared-memory rows
//   [ a_m | d_m | d_{m-1} | ... | d_1 ]      (a_m at row 0, d_k at row offD[k])
// each staged from slot O_k (a multiple of the box height, periodic) by its own mbarrier, coarsest level
// first, so level m starts while d_1 - half of all bytes - is still in flight.  Level k > 1 writes a_{k-1}
// IN PLACE over a_k and d_k (both dead by then; the launcher places d_{k-1} behind the rows a_{k-1} needs);
// level 1 stores to global memory.  Resident mode: the coefficient prefix [0, h0) of the line is already
// [a_m | d_m | ... | d_1]; level k overwrites rows [0, 2 (h0 >> k)).
template <int L, bool RESIDENT>
__global__ void __launch_bounds__(RESIDENT ? kMaxThrRes2 : kMaxThr2, 2)
k_fwt_rev_str2(const __grid_constant__ Taps taps, const __grid_constant__ FwtRevStrArgs a,
               const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapD) {
  extern __shared__ __align__(1024) double2 sm2[];
  constexpr int R = kR2;
  double2* X = sm2;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm2 + size_t(a.rowsC) * kL8);  // bar[k], k = 0 .. m
  const int tid = threadIdx.x, l8 = tid & (kL8 - 1), grp = tid >> 3, ngrp = blockDim.x >> 3;
  const int m = a.m, h0 = a.h0;
  int64_t b = blockIdx.x;
  const int cb = int(b % a.cblocks); b /= a.cblocks;
  const int tile = RESIDENT ? 0 : int(b % a.tiles_per_line);
  const int64_t o = RESIDENT ? b : b / a.tiles_per_line;
  const int64_t inner = a.inner;
  double* gY = a.dst + o * a.dst_os + cb * kC2 + 2 * l8;
  auto pY = [&](int64_t row) { return a.rm.mode ? remote_row(a.rm, o, row) + cb * kC2 + 2 * l8 : gY + row * inner; };
  const int z = tap_phase<L, true>();
  const int warp = tid >> 5, nwarp = blockDim.x >> 5;

  if constexpr (!RESIDENT) {
    const int T = a.T, t0 = tile * T;
    if (tid == 0) {
      for (int k = 1; k <= m; ++k) mbar_init(&bar[k], 1);
    }
    __syncthreads();
    if (tid == 0) {
      for (int k = 1; k <= m; ++k)
        mbar_expect_tx(&bar[k], unsigned(a.len[k]) * (k == m ? 2u : 1u) * kC2 * sizeof(double));
    }
    if ((tid & 31) == 0) {
      // every warp's first lane issues its share of every level's boxes (box j of a level: warps j mod nwarp),
      // coarsest level first
      const int64_t rowD0 = o * a.rowsD_per_o, rowA0 = o * a.rowsA_per_o;
      for (int k = m; k >= 1; --k) {
        const int wk = h0 >> k;
        const int O = (k == m) ? ((t0 >> k) - a.F[k] - a.ru) : 2 * ((t0 >> (k + 1)) - a.F[k + 1]);
        const int nb = a.len[k] / kBoxR;
        for (int j = warp; j < nb; j += nwarp) {
          const int slot = (O + j * kBoxR) & (wk - 1);
          if (k == m) tma_load_box(X + size_t(j) * kBoxR * kL8, &tmapA, cb * kC2, int(rowA0 + slot), &bar[k]);
          tma_load_box(X + size_t(a.offD[k] + j * kBoxR) * kL8, &tmapD, cb * kC2, int(rowD0 + wk + slot), &bar[k]);
        }
      }
    }
    for (int k = m; k >= 1; --k) {
      mbar_wait(&bar[k], 0);
      const double2* D = X + size_t(a.offD[k]) * kL8;
      const int groups = ((T >> k) + a.F[k]) / R;
      const int s0 = a.s0[k];  // local row of the level's first slot
      if (k > 1) {
        // one task per thread (the launcher sizes the CTA for it): a_{k-1} waits in registers
        double2 t[2 * R];
        const bool has = grp < groups;
        if (has) {
          const int top = s0 + R * grp + R - 1;
          const double2* wa = X + top * kL8 + l8;
          const double2* wd = D + top * kL8 + l8;
          rev_run2<L, R>(taps, z, [&](int s) { return wa[-s * kL8]; }, [&](int s) { return wd[-s * kL8]; }, t);
        }
        __syncthreads();
        if (has) {
#pragma unroll
          for (int e = 0; e < 2 * R; ++e) X[(2 * R * grp + e) * kL8 + l8] = t[e];
        }
        __syncthreads();
      } else if (tid < a.nmain) {
        // level 1 (no left extension): the main warps only - whole multiples of 4 warps, see the launcher
        for (int g = grp; g < groups; g += a.nmain >> 3) {
          double2 t[2 * R];
          const int top = s0 + R * g + R - 1;
          const double2* wa = X + top * kL8 + l8;
          const double2* wd = D + top * kL8 + l8;
          rev_run2<L, R>(taps, z, [&](int s) { return wa[-s * kL8]; }, [&](int s) { return wd[-s * kL8]; }, t);
#pragma unroll
          for (int e = 0; e < 2 * R; ++e) st2(pY(t0 + 2 * R * g + e), t[e]);
        }
      }
    }
  } else {
    if (tid == 0) mbar_init(&bar[0], 1);
    __syncthreads();
    if (tid == 0) mbar_expect_tx(&bar[0], unsigned(h0) * kC2 * sizeof(double));
    if ((tid & 31) == 0) {
      const int64_t rowD0 = o * a.rowsD_per_o;
      for (int j = warp; j < h0 / kBoxR; j += nwarp)
        tma_load_box(X + size_t(j) * kBoxR * kL8, &tmapD, cb * kC2, int(rowD0 + j * kBoxR), &bar[0]);
    }
    mbar_wait(&bar[0], 0);
    for (int k = m; k >= 1; --k) {
      const int half = h0 >> k, mask = half - 1;
      const bool last = (k == 1);
      if (half >= R) {
        // RR = slots of one task: kR2 at the last (largest) level, which loops; at the levels whose results wait in
        // registers the smallest of 1, 2, 4 that fits the level into one round, so that the short levels at the
        // start of a full-depth reverse spread over all thread groups (see the forward kernel)
        auto level = [&](auto rc) {
          constexpr int RR = decltype(rc)::value;
          const int groups = half / RR;
          auto task = [&](int g, double2 (&t)[2 * RR]) {
            const int top = RR * g + RR - 1;
            rev_run2<L, RR>(taps, z, [&](int s) { return X[((top - s) & mask) * kL8 + l8]; },
                            [&](int s) { return X[(half + ((top - s) & mask)) * kL8 + l8]; }, t);
          };
          if (!last) {
            double2 t[2 * RR];
            const bool has = grp < groups;
            if (has) task(grp, t);
            __syncthreads();
            if (has) {
#pragma unroll
              for (int e = 0; e < 2 * RR; ++e) X[(2 * RR * grp + e) * kL8 + l8] = t[e];
            }
            __syncthreads();
          } else {
            for (int g = grp; g < groups; g += ngrp) {
              double2 t[2 * RR];
              task(g, t);
#pragma unroll
              for (int e = 0; e < 2 * RR; ++e) st2(pY(2 * RR * g + e), t[e]);
            }
          }
        };
        if (last || half > 2 * ngrp) level(std::integral_constant<int, 4>{});
        else if (half > ngrp) level(std::integral_constant<int, 2>{});
        else level(std::integral_constant<int, 1>{});
      } else {
        // columns of 2 or 4 samples being rebuilt: one slot per thread group, true modular indexing
        double2 t0v = make_double2(0.0, 0.0), t1v = t0v;
        const bool has = grp < half;
        if (has) {
#pragma unroll 1
          for (int q = 0; q < L / 2; ++q) {
            const int i = (grp - q) & mask;
            const double2 av = X[i * kL8 + l8], dv = X[(half + i) * kL8 + l8];
            fma2(t0v, av, lo_tap<L>(taps, 2 * q, z));
            fma2(t1v, av, lo_tap<L>(taps, 2 * q + 1, z));
            fma2(t0v, dv, hi_tap<L>(taps, 2 * q, z));
            fma2(t1v, dv, hi_tap<L>(taps, 2 * q + 1, z));
          }
        }
        if (!last) {
          __syncthreads();
          if (has) {
            X[(2 * grp) * kL8 + l8] = t0v;
            X[(2 * grp + 1) * kL8 + l8] = t1v;
          }
          __syncthreads();
        } else if (has) {
          st2(pY(2 * grp), t0v);
          st2(pY(2 * grp + 1), t1v);
        }
      }
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------

static int round_up2(int v, int q) { return (v + q - 1) / q * q; }

int fwt_str2_tile_levels(int L, int T, int want) {
  // as v1: largest m with halo (2^m - 1)(L - 2) <= T / 4 and T / 2^m >= one TMA box; `want` > 0 overrides
  // the halo rule (experiments), never the box rule
  int m = 1;
  while ((want > 0 ? m < want : ((1 << (m + 1)) - 1) * (L - 2) <= T / 4) && (T >> (m + 1)) >= kBoxF) ++m;
  return m;
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn2 encode_tiled2() {
  static EncodeTiledFn2 fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn2>(p);
  }();
  return fn;
}

// 2-D map of a dense [rows][inner] array of doubles, box {16 columns, box_rows rows}, dense in shared memory
static bool make_tmap2(CUtensorMap* map, const double* base, int64_t rows, int64_t inner, int box_rows) {
  EncodeTiledFn2 enc = encode_tiled2();
  if (!enc || (reinterpret_cast<uintptr_t>(base) & 15) || inner % 2 || rows < 1 || rows >= (int64_t(1) << 31)) return false;
  const cuuint64_t dims[2] = {cuuint64_t(inner), cuuint64_t(rows)};
  const cuuint64_t strides[1] = {cuuint64_t(inner) * sizeof(double)};
  const cuuint32_t box[2] = {kC2, cuuint32_t(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int L>
static cudaError_t launch_fwd2_L(jwc_ctx* ctx, const Taps& taps, FwtFwdStrArgs a, bool resident) {
  constexpr int R = kR2;
  if (a.inner % kC2 || a.src_os % a.inner || !aligned16(a.dstD) || !aligned16(a.dstA) || a.m < 1) return cudaErrorNotSupported;
  if ((a.dstD_os | a.dstA_os) & 1) return cudaErrorNotSupported;
  a.cblocks = int(a.inner / kC2);
  a.rows_per_o = a.src_os / a.inner;
  int groups1;
  int64_t grid;
  if (!resident) {
    if (a.h % a.T || a.T % kBoxF || (a.T >> a.m) < R || (a.T >> a.m) % R) return cudaErrorNotSupported;
    const int n0 = a.T + ((1 << a.m) - 1) * (L - 2);
    groups1 = ((a.T >> 1) + ((1 << (a.m - 1)) - 1) * (L - 2) + R - 1) / R;
    a.rows0 = round_up2(max(n0, 2 * R * groups1 + L - 2), kBoxF);  // the last group's window may overshoot n0
    a.tiles_per_line = a.h / a.T;
    grid = a.outer * a.tiles_per_line * a.cblocks;
  } else {
    if (a.h < kBoxF || a.h % kBoxF) return cudaErrorNotSupported;
    groups1 = max(1, (a.h / 2) / R);
    a.rows0 = round_up2(a.h + L - 2, kBoxF);  // the line and its periodic extension
    a.tiles_per_line = 1;
    grid = a.outer * a.cblocks;
  }
  // Tile mode: T / 2 threads = whole multiples of 4 warps (warps map round-robin onto the 4 sub-partitions of
  // an SM; with 10 warps two sub-partitions carried 3 and two carried 2, and the per-level barriers turned
  // that into 17 % idle FP64 pipe, profiles/r02_ncu_c4_str2_first.md).  The kept groups of every level then
  // fill whole rounds (2 at level 1, 1 at level 2, ...) and the halo groups, low pass only, are one more short
  // round.  Resident mode: two rounds at the first level.
  int nthr;
  if (!resident) nthr = max(128, min(256, a.T / 2 / 128 * 128));
  else nthr = min(kMaxThrRes2, max(64, round_up2((groups1 + 1) / 2 * kL8, 32)));
  a.rounds1 = (groups1 + nthr / kL8 - 1) / (nthr / kL8);
  const size_t smem = size_t(a.rows0) * kC2 * sizeof(double) + kStagesF * sizeof(uint64_t);  // + the mbarriers
  if (nthr > (resident ? kMaxThrRes2 : kMaxThr2) || smem > ctx->smem_optin || grid > 0x7fffffff || grid < 1) return cudaErrorNotSupported;
  CUtensorMap tmap;
  if (!make_tmap2(&tmap, a.src, a.outer * a.rows_per_o, a.inner, kBoxF)) return cudaErrorNotSupported;
  void (*kern)(const Taps, const FwtFwdStrArgs, const CUtensorMap) =
      resident ? k_fwt_fwd_str2<L, true> : k_fwt_fwd_str2<L, false>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
  }
  prof_begin(ctx, resident ? "k_fwt_fwd_str2:resident" : "k_fwt_fwd_str2:tile", double(a.outer) * a.h * a.inner, a.m);
  kern<<<int(grid), nthr, smem, ctx->stream>>>(taps, a, tmap);
  prof_end(ctx);
  ctx->launches++;
  return cudaGetLastError();
}

cudaError_t launch_fwt_fwd_str2(jwc_ctx* ctx, int L, const Taps& taps, const FwtFwdStrArgs& a, bool resident) {
  switch (L) {
#define JWC_CASE(LL) case LL: return launch_fwd2_L<LL>(ctx, taps, a, resident);
    JWC_FOR_EACH_L(JWC_CASE)
#undef JWC_CASE
  }
  return cudaErrorInvalidValue;
}

template <int L>
static cudaError_t launch_rev2_L(jwc_ctx* ctx, const Taps& taps, FwtRevStrArgs a, bool resident) {
  constexpr int R = kR2;
  if (a.inner % kC2 || a.srcA_os % a.inner || a.srcD_os % a.inner || !aligned16(a.dst) || (a.dst_os & 1) || a.m < 1)
    return cudaErrorNotSupported;
  a.cblocks = int(a.inner / kC2);
  a.rowsA_per_o = a.srcA_os / a.inner;
  a.rowsD_per_o = a.srcD_os / a.inner;
  int64_t grid;
  int nthr;
  if (!resident) {
    if (a.m > kMaxFuse || a.h0 % a.T || (a.T >> a.m) < kBoxR || (a.T >> a.m) % kBoxR) return cudaErrorNotSupported;
    a.ru = round_up2(L / 2 - 1, R);
    int N = 0;  // left extension of a_{k-1} that level k-1 needs
    for (int k = 1; k <= a.m; ++k) {
      a.F[k] = round_up2((N + 1) / 2, R);
      N = a.F[k] + L / 2 - 1;
    }
    a.F[a.m + 1] = 0;
    if ((a.F[a.m] + a.ru) % kBoxR) a.ru += R;  // the staged range of level m starts on a box boundary
    int gmax = 0;
    for (int k = 1; k <= a.m; ++k) {
      a.len[k] = (k == a.m) ? (a.T >> k) + a.F[k] + a.ru : (a.T >> k) + 2 * a.F[k + 1];
      a.s0[k] = (k == a.m) ? a.ru : 2 * a.F[k + 1] - a.F[k];
      if (a.len[k] % kBoxR) return cudaErrorNotSupported;
      if (k >= 2) gmax = max(gmax, ((a.T >> k) + a.F[k]) / R);
    }
    // rows: a_m at 0, d_m behind it; d_{k-1} behind everything staged so far AND behind the rows of a_{k-1}
    a.offD[a.m] = a.len[a.m];
    int end = 2 * a.len[a.m];
    for (int k = a.m - 1; k >= 1; --k) {
      a.offD[k] = max(end, a.len[k]);
      end = a.offD[k] + a.len[k];
    }
    a.rowsC = end;
    // main warps: T / 2 threads (whole multiples of 4 warps: the SM's 4 sub-partitions stay balanced) take the
    // kept groups - one round at level 2, two at level 1; the left-extension groups of the levels that hold
    // their results in registers go to one extra warp
    a.nmain = max(128, min(256, a.T / 2 / 128 * 128));
    nthr = max(a.nmain, round_up2(gmax * kL8, 32));
    a.tiles_per_line = a.h0 / a.T;
    grid = a.outer * a.tiles_per_line * a.cblocks;
  } else {
    if (a.h0 < kBoxR || a.h0 % kBoxR) return cudaErrorNotSupported;
    a.rowsC = a.h0;
    nthr = max(64, round_up2(max(1, (a.h0 / 4) / R) * kL8, 32));
    a.tiles_per_line = 1;
    grid = a.outer * a.cblocks;
  }
  const size_t smem = size_t(a.rowsC) * kC2 * sizeof(double) + (kMaxFuse + 2) * sizeof(uint64_t);
  if (nthr > (resident ? kMaxThrRes2 : kMaxThr2) || smem > ctx->smem_optin || grid > 0x7fffffff || grid < 1) return cudaErrorNotSupported;
  CUtensorMap tmapA, tmapD;
  if (!make_tmap2(&tmapD, a.srcD, a.outer * a.rowsD_per_o, a.inner, kBoxR)) return cudaErrorNotSupported;
  if (!make_tmap2(&tmapA, a.srcA, a.outer * a.rowsA_per_o, a.inner, kBoxR)) return cudaErrorNotSupported;
  void (*kern)(const Taps, const FwtRevStrArgs, const CUtensorMap, const CUtensorMap) =
      resident ? k_fwt_rev_str2<L, true> : k_fwt_rev_str2<L, false>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
  }
  prof_begin(ctx, resident ? "k_fwt_rev_str2:resident" : "k_fwt_rev_str2:tile", double(a.outer) * a.h0 * a.inner, a.m);
  kern<<<int(grid), nthr, smem, ctx->stream>>>(taps, a, tmapA, tmapD);
  prof_end(ctx);
  ctx->launches++;
  return cudaGetLastError();
}

cudaError_t launch_fwt_rev_str2(jwc_ctx* ctx, int L, const Taps& taps, const FwtRevStrArgs& a, bool resident) {
  switch (L) {
#define JWC_CASE(LL) case LL: return launch_rev2_L<LL>(ctx, taps, a, resident);
    JWC_FOR_EACH_L(JWC_CASE)
#undef JWC_CASE
  }
  return cudaErrorInvalidValue;
}

}  // namespace jwc
