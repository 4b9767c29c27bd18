// jwc_fwt_rev.cu - fused multi-level reverse FWT along contiguous lines.
//
// Replaces the level loop of FastWaveletTransform.reverse (FastWaveletTransform.java:143-149)
// around Wavelet.reverse (Wavelet.java:277-303) for `m` consecutive levels per launch.  The
// reference's scatter `t[(2i+j) mod h] += a[i] s[j] + d[i] w[j]` is evaluated as a gather
// (rev_step below), so no atomics and no zero-fill are needed.
//
// Launch-level numbering: level 0 is the output of the launch (width h0), level m its coarsest
// input a_m (width h0 >> m); d_k (width h0 >> k) sits at `srcD + (h0 >> k)` in every line,
// which is where the FWT layout [a_l | d_l | ... | d_1] keeps it.
//
//   resident mode (h0 <= res_cap): G whole lines per CTA, wrap by index mask, every level down to
//                                h = 2 (a scalar path covers widths below 16).
//   tile mode     (h0  > res_cap): one CTA = T output samples.  Level k needs a_k / d_k only
//                                N_k = F_k + L/2 - 1 coefficients to the LEFT of the tile (F_k = 8-aligned
//                                extension computed at that level, < L) - the halo does not grow
//                                geometrically as it does in the forward direction.
#include "jwc_fused.cuh"
#include "jwc_kernels.cuh"

namespace jwc {

// Shared-memory layout of the reverse kernels.  A thread reads windows of consecutive double2 and
// neighbouring threads start kRS/2 = 2 double2 apart at ANY alignment, so the 8 lanes of an LDS.128
// phase cover one 16-slot window at stride 2: with double2 k stored at k ^ bit3(k) the slots of the
// upper half of every 16-block trade parity and the 8 lanes land on 8 distinct 16-byte bank groups -
// without pad slots (the stride-4 padding of the forward kernels, pad2, gives 2-way conflicts here).
__device__ __forceinline__ int lay(int k2) { return k2 ^ ((k2 >> 3) & 1); }
__host__ __device__ constexpr int lay_size(int n2) { return (n2 + 1) & ~1; }
__device__ __forceinline__ double lay_scalar(const double2* buf, int i) {
  return reinterpret_cast<const double*>(buf)[2 * lay(i >> 1) + (i & 1)];
}

// Store the kRS double2 a group produced at slots kRS g .. kRS g + kRS - 1.  For kRS = 4 the lanes of an
// STS.128 phase are 4 slots apart (32 slots in all): lanes 4-7 store their pairs in the order 2, 3, 0, 1,
// which together with the bit-3 swap makes every phase conflict-free.
template <int kRS>
__device__ __forceinline__ void store_group(double2* Y, int g, const double (&t)[2 * kRS]) {
  if constexpr (kRS == 4) {
    const bool rot = (g >> 2) & 1;
    const int base = 4 * g, x = (g >> 1) & 1, r2 = rot ? 2 : 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const double v0 = rot ? t[2 * (e ^ 2)] : t[2 * e];
      const double v1 = rot ? t[2 * (e ^ 2) + 1] : t[2 * e + 1];
      Y[(base + (e ^ r2)) ^ x] = make_double2(v0, v1);
    }
  } else {
#pragma unroll
    for (int e = 0; e < kRS; ++e) Y[lay(kRS * g + e)] = make_double2(t[2 * e], t[2 * e + 1]);
  }
}

// RS consecutive slots p = RS g' .. RS g' + RS - 1 (RS = 8 or 4) -> t[2 RS]:
//   t[2pp + r] = sum_q a[p - q] lo[2q + r] + d[p - q] hi[2q + r].
// `a2(w)` / `d2(w)` return double2 number (RS/2 g' + RS/2 - 1 - w) of the a / d arrays, w = 0 .. L/4 + RS/2 - 1.
template <int L, int RS, class A2, class D2>
__device__ __forceinline__ void rev_step(const Taps& taps, A2 a2, D2 d2, double (&t)[2 * RS]) {
  // long filters: keep the tap loads inside the step (tap_phase, jwc_internal.cuh) - with the main and the
  // tail step in one kernel ptxas otherwise hoists the taps into 100+ vector registers (L = 40: 168)
  const int z = tap_phase<L, true>();
#pragma unroll
  for (int r = 0; r < 2 * RS; ++r) t[r] = 0.0;
  constexpr int W = (L / 2) / 2 + RS / 2;
  // short filters: fetch the whole window first, so the LDS latency is paid once per group instead of
  // once per window step (the stall samples of the streaming form sat on short-scoreboard waits)
  constexpr bool kPrefetch = (W <= 6);
  double2 wa[kPrefetch ? W : 1], wd[kPrefetch ? W : 1];
  if constexpr (kPrefetch) {
#pragma unroll
    for (int w = 0; w < W; ++w) { wa[w] = a2(w); wd[w] = d2(w); }
  }
#pragma unroll
  for (int w = 0; w < W; ++w) {
    double2 av, dv;
    if constexpr (kPrefetch) { av = wa[w]; dv = wd[w]; }
    else { av = a2(w); dv = d2(w); }
#pragma unroll
    for (int pp = 0; pp < RS; ++pp) {
      const int qy = pp - (RS - 1) + 2 * w;  // .y is slot RS g' + RS - 1 - 2w
      const int qx = qy + 1;                 // .x is the slot before it
      if (qy >= 0 && qy < L / 2) {
        t[2 * pp] = fma(av.y, lo_tap<L>(taps, 2 * qy, z), t[2 * pp]);
        t[2 * pp] = fma(dv.y, hi_tap<L>(taps, 2 * qy, z), t[2 * pp]);
        t[2 * pp + 1] = fma(av.y, lo_tap<L>(taps, 2 * qy + 1, z), t[2 * pp + 1]);
        t[2 * pp + 1] = fma(dv.y, hi_tap<L>(taps, 2 * qy + 1, z), t[2 * pp + 1]);
      }
      if (qx >= 0 && qx < L / 2) {
        t[2 * pp] = fma(av.x, lo_tap<L>(taps, 2 * qx, z), t[2 * pp]);
        t[2 * pp] = fma(dv.x, hi_tap<L>(taps, 2 * qx, z), t[2 * pp]);
        t[2 * pp + 1] = fma(av.x, lo_tap<L>(taps, 2 * qx + 1, z), t[2 * pp + 1]);
        t[2 * pp + 1] = fma(dv.x, hi_tap<L>(taps, 2 * qx + 1, z), t[2 * pp + 1]);
      }
    }
  }
}

template <int L, bool RESIDENT, int kRS>
__global__ void __launch_bounds__(384)
k_fwt_rev(const __grid_constant__ Taps taps, const __grid_constant__ FwtRevArgs a) {
  extern __shared__ double2 smem2[];
  const int tid = rotated_tid(a.rot), nthr = blockDim.x;
  const int m = a.m, h0 = a.h0;

  if constexpr (!RESIDENT) {
    // ---------------- tile mode ----------------
    const int64_t line = blockIdx.x >> a.lg_tpl;  // tiles per line: a power of two
    const int tile = int(blockIdx.x) & (a.tiles_per_line - 1);
    const int T = a.T;
    const int t0 = tile * T;
    const double* lineD = a.srcD + line * a.srcD_os;
    // stage d_k (k = 1..m) and a_m, coarsest level first and one cp.async group per level: level k
    // starts as soon as ITS coefficients have landed, while d_1 (half of all bytes) is still in flight.
    // Local sample j of level k is absolute slot O_k + j (periodic).
    for (int k = m; k >= 1; --k) {
      const int wk = h0 >> k;  // width of a_k and d_k
      const int O = (k == m) ? ((t0 >> k) - a.F[k] - a.ru8) : 2 * ((t0 >> (k + 1)) - a.F[k + 1]);
      double2* D = smem2 + a.offD[k];
      const double* dk = lineD + wk;
#ifndef JWC_ABLATE_LD  // (timing experiments only: the kernel without its staging loads / FMA steps / output stores)
      for (int j2 = tid; j2 < a.len[k] / 2; j2 += nthr)
        cp_async16(&D[lay(j2)], dk + ((O + 2 * j2) & (wk - 1)));
      if (k == m) {
        double2* A = smem2 + a.offA[m & 1];
        const double* am = a.srcA + line * a.srcA_os;
        for (int j2 = tid; j2 < a.len[k] / 2; j2 += nthr)
          cp_async16(&A[lay(j2)], am + ((O + 2 * j2) & (wk - 1)));
      }
#endif
      cp_async_commit();
    }

    for (int k = m; k >= 1; --k) {
      cp_async_wait_pending(k - 1);  // groups of levels k-1 .. 1 may still be in flight
      __syncthreads();               // level k staged for every thread, a_k complete
      const double2* A = smem2 + a.offA[k & 1];
      const double2* D = smem2 + a.offD[k];
      double2* Y = smem2 + a.offA[(k - 1) & 1];
      const int groups = ((T >> k) + a.F[k]) / kRS;
      const int g0 = a.g0[k];
      // With a tail warp (a.tail) the F_k slots of left extension the levels below need - the first
      // F_k / kRS groups - are its job, two slots per lane; the main warps then run exactly (T >> k) / kRS
      // groups, a power of two, instead of one more partly filled kRS-wide step per level.
      const int nmain = nthr - 32 * a.tail;
      const int gl = a.tail ? a.F[k] / kRS : 0;
      if (tid < nmain) {
        for (int g = gl + tid; g < groups; g += nmain) {
          double t[2 * kRS];
          const int c = 4 * g0 + (kRS / 2) * g + kRS / 2 - 1;
#ifdef JWC_ABLATE_FMA
#pragma unroll
          for (int e = 0; e < 2 * kRS; ++e) t[e] = A[lay(c)].x;
#else
          rev_step<L, kRS>(taps, [&](int w) { return A[lay(c - w)]; }, [&](int w) { return D[lay(c - w)]; }, t);
#endif
          if (k > 1) {
            store_group<kRS>(Y, g, t);
          } else {
            double* y = (a.rm.mode ? remote_line(a.rm, line) : a.dst + line * a.dst_os) + t0 + 2 * kRS * g;
#ifdef JWC_ABLATE_STG  // timing experiment only (tools/build_variant.sh): the kernel without its output stores
            if (t[0] == 123.456)
#endif
#pragma unroll
            for (int e = 0; e < kRS / 2; ++e) st_global_v4(y + 4 * e, t[4 * e], t[4 * e + 1], t[4 * e + 2], t[4 * e + 3]);
            static_assert(kRS >= 2, "a group stores at least 4 samples");
          }
        }
      } else {
        // F_k < L <= 40, so the F_k / 2 steps fit the warp's lanes: no loop (a loop here makes ptxas hoist the
        // taps of long filters out of both loops into vector registers).  F_1 == 0: never at the output level.
        const int g2 = tid - nmain;
        if constexpr (L <= kTailMaxL) if (g2 < a.F[k] / 2) {
          double t4[4];
          const int c = 4 * g0 + g2;
          rev_step<L, 2>(taps, [&](int w) { return A[lay(c - w)]; }, [&](int w) { return D[lay(c - w)]; }, t4);
          store_group<2>(Y, g2, t4);
        }
      }
    }
  } else {
    // ---------------- resident mode ----------------
    const int G = a.G;
    const int64_t line0 = int64_t(blockIdx.x) * G;
    const int nlines = int(min(int64_t(G), a.lines - line0));
    const int capC = a.capC, capP0 = a.capP[0], capP1 = a.capP[1];
    double2* C = smem2;                                  // coefficient prefix [a_m | d_m | ... | d_1]
    double2* P[2] = {smem2 + size_t(G) * capC, smem2 + size_t(G) * (capC + capP0)};  // a_k: P[k & 1]
    const int capP[2] = {capP0, capP1};
    {
      const int per_line = h0 >> 1;
      for (int it = tid; it < nlines * per_line; it += nthr) {
        const int ln = it / per_line, k2 = it - ln * per_line;
        cp_async16(&C[ln * capC + lay(k2)], a.srcD + (line0 + ln) * a.srcD_os + 2 * k2);
      }
      cp_async_wait_all();
      __syncthreads();
    }
    for (int k = m; k >= 1; --k) {
      const int half = h0 >> k;  // length of a_k and d_k
      const bool from_c = (k == m);
      const bool last = (k == 1);
      if (half >= kRS) {
        const int gpl = half / kRS;
        const int mask2 = (half >> 1) - 1;
        const int doff = half >> 1;  // d_k starts at sample `half` of the prefix
        for (int it = tid; it < nlines * gpl; it += nthr) {
          const int ln = it / gpl, g = it - ln * gpl;
          const double2* cl = C + ln * capC;
          const double2* al = from_c ? cl : P[k & 1] + ln * capP[k & 1];
          const int c = (kRS / 2) * g + kRS / 2 - 1;
          double t[2 * kRS];
          rev_step<L, kRS>(taps, [&](int w) { return al[lay((c - w) & mask2)]; },
                       [&](int w) { return cl[lay(doff + ((c - w) & mask2))]; }, t);
          if (!last) {
            store_group<kRS>(P[(k - 1) & 1] + ln * capP[(k - 1) & 1], g, t);
          } else {
            double* y = (a.rm.mode ? remote_line(a.rm, line0 + ln) : a.dst + (line0 + ln) * a.dst_os) + 2 * kRS * g;
#pragma unroll
            for (int e = 0; e < kRS / 2; ++e) st_global_v4(y + 4 * e, t[4 * e], t[4 * e + 1], t[4 * e + 2], t[4 * e + 3]);
          static_assert(kRS >= 2, "a group stores at least 4 samples");
          }
        }
      } else {
        // widths 2, 4, 8: one thread per (line, slot), true modular indexing (h < L wraps)
        const int mask = half - 1;
        for (int it = tid; it < nlines * half; it += nthr) {
          const int ln = it / half, p = it - ln * half;
          const double2* cl = C + ln * capC;
          const double2* al = from_c ? cl : P[k & 1] + ln * capP[k & 1];
          double t0v = 0.0, t1v = 0.0;
#pragma unroll
          for (int q = 0; q < L / 2; ++q) {
            const int i = (p - q) & mask;
            const double av = lay_scalar(al, i), dv = lay_scalar(cl, half + i);
            t0v = fma(av, taps.lo[2 * q], t0v);
            t0v = fma(dv, hi_tap<L>(taps, 2 * q), t0v);
            t1v = fma(av, taps.lo[2 * q + 1], t1v);
            t1v = fma(dv, hi_tap<L>(taps, 2 * q + 1), t1v);
          }
          if (!last) {
            P[(k - 1) & 1][ln * capP[(k - 1) & 1] + lay(p)] = make_double2(t0v, t1v);
          } else {
            double* y = (a.rm.mode ? remote_line(a.rm, line0 + ln) : a.dst + (line0 + ln) * a.dst_os) + 2 * p;
            y[0] = t0v;
            y[1] = t1v;
          }
        }
      }
      __syncthreads();
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------

static int round_up8(int v) { return (v + 7) & ~7; }

template <int L>
static cudaError_t launch_L(jwc_ctx* ctx, const Taps& taps, FwtRevArgs a, bool resident) {
  size_t smem;
  int grid;
  if (!resident) {
    if (a.m < 1 || a.m > kMaxFuse || (a.T >> a.m) < 8) return cudaErrorInvalidValue;
    a.ru8 = round_up8(L / 2 - 1);
    int N = 0;  // N_{k-1}: left extension of a_{k-1} the level below needs
    for (int k = 1; k <= a.m; ++k) {
      a.F[k] = round_up8((N + 1) / 2);
      N = a.F[k] + L / 2 - 1;
    }
    a.F[a.m + 1] = 0;
    int off = 0;
    int capA[2] = {0, 0};
    for (int k = 1; k <= a.m; ++k) {
      a.len[k] = (k == a.m) ? (a.T >> k) + a.F[k] + a.ru8 : (a.T >> k) + 2 * a.F[k + 1];
      a.g0[k] = (k == a.m) ? a.ru8 / 8 : (2 * a.F[k + 1] - a.F[k]) / 8;
      a.offD[k] = off;
      off += lay_size(a.len[k] / 2);
      if (lay_size(a.len[k] / 2) > capA[k & 1]) capA[k & 1] = lay_size(a.len[k] / 2);
    }
    a.offA[0] = off;
    a.offA[1] = off + capA[0];
    smem = size_t(off + capA[0] + capA[1]) * sizeof(double2);
    a.tiles_per_line = a.h0 / a.T;
    a.lg_tpl = 0;
    while ((1 << a.lg_tpl) < a.tiles_per_line) ++a.lg_tpl;
    if ((1 << a.lg_tpl) != a.tiles_per_line) return cudaErrorInvalidValue;
    const int64_t ctas = a.lines * a.tiles_per_line;
    if (ctas > 0x7fffffff) return cudaErrorInvalidConfiguration;
    grid = int(ctas);
  } else {
    // + 1: successive lines start on the other slot parity (short lines share an LDS phase)
    a.capC = lay_size(a.h0 / 2) + 1;
    a.capP[1] = lay_size(max(1, a.h0 / 4)) + 1;  // a_1 (odd levels): h0 / 2 samples
    a.capP[0] = lay_size(max(1, a.h0 / 8)) + 1;  // a_2 (even levels): h0 / 4 samples
    smem = size_t(a.G) * (a.capC + a.capP[0] + a.capP[1]) * sizeof(double2);
    grid = int((a.lines + a.G - 1) / a.G);
  }
  auto kern = resident ? (ctx->rev_rs == 4 ? k_fwt_rev<L, true, 4> : ctx->rev_rs == 2 ? k_fwt_rev<L, true, 2> : k_fwt_rev<L, true, 8>)
                       : (ctx->rev_rs == 4 ? k_fwt_rev<L, false, 4> : ctx->rev_rs == 2 ? k_fwt_rev<L, false, 2> : k_fwt_rev<L, false, 8>);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
  }
  if (ctx->carve) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  prof_begin(ctx, resident ? "k_fwt_rev:resident" : "k_fwt_rev:tile", double(a.lines) * a.h0, a.m);
  a.tail = (!resident && ctx->rev_tail && L <= kTailMaxL) ? 1 : 0;
  a.rot = (a.tail && ctx->rot_warps) ? 1 : 0;
  kern<<<grid, resident ? ctx->res_threads : ctx->rev_threads + 32 * a.tail, smem, ctx->stream>>>(taps, a);
  prof_end(ctx);
  ctx->launches++;
  return cudaGetLastError();
}

cudaError_t launch_fwt_rev(jwc_ctx* ctx, int L, const Taps& taps, const FwtRevArgs& a, bool resident) {
  switch (L) {
#define JWC_CASE(LL) case LL: return launch_L<LL>(ctx, taps, a, resident);
    JWC_FOR_EACH_L(JWC_CASE)
#undef JWC_CASE
  }
  return cudaErrorInvalidValue;
}

}  // namespace jwc
