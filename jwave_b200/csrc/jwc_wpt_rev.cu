// jwc_wpt_rev.cu - fused multi-level reverse wavelet PACKET transform along contiguous lines.
//
// Replaces the level x packet loops of WaveletPacketTransform.reverse
// (WaveletPacketTransform.java:164-187; Pooled / Parallel variants:
// PooledWaveletPacketTransform.java:74-127, ParallelWaveletPacketTransform.java:113-146) around
// Wavelet.reverse (Wavelet.java:277-303), `m` levels per launch, gather form (jwc_fwt_rev.cu).
//
// Launch-level numbering: level 0 is the output (one packet of width h0 per line), level m the
// input: 2^m leaf packets of width h0 >> m, stored back to back in natural order.  Nodes 2j and
// 2j+1 of level k are the approximation / detail halves that rebuild node j of level k-1.
//
//   tile mode     (h0  > res_cap): one CTA = T output samples; every node of level k needs only
//                                  N_k = F_k + L/2 - 1 coefficients left of the tile (see jwc_fwt_rev.cu).
//   resident mode (h0 <= res_cap): G whole lines per CTA, wrap by index mask; nodes shorter than 8
//                                  take a scalar path with true modular indexing.
#include "jwc_fused.cuh"
#include "jwc_kernels.cuh"

#ifndef JWC_WPT_TAIL_WARP
#define JWC_WPT_TAIL_WARP 1
#endif

namespace jwc {

// RS consecutive slots -> t[2 RS]; a2(w)/d2(w) = double2 (RS/2 g' + RS/2 - 1 - w); see jwc_fwt_rev.cu
template <int L, int RS, class A2, class D2>
__device__ __forceinline__ void wrev_step(const Taps& taps, A2 a2, D2 d2, double (&t)[2 * RS]) {
#pragma unroll
  for (int r = 0; r < 2 * RS; ++r) t[r] = 0.0;
  constexpr int W = (L / 2) / 2 + RS / 2;
#pragma unroll
  for (int w = 0; w < W; ++w) {
    const double2 av = a2(w), dv = d2(w);
#pragma unroll
    for (int pp = 0; pp < RS; ++pp) {
      const int qy = pp - (RS - 1) + 2 * w;
      const int qx = qy + 1;
      if (qy >= 0 && qy < L / 2) {
        t[2 * pp] = fma(av.y, taps.lo[2 * qy], t[2 * pp]);
        t[2 * pp] = fma(dv.y, hi_tap<L>(taps, 2 * qy), t[2 * pp]);
        t[2 * pp + 1] = fma(av.y, taps.lo[2 * qy + 1], t[2 * pp + 1]);
        t[2 * pp + 1] = fma(dv.y, hi_tap<L>(taps, 2 * qy + 1), t[2 * pp + 1]);
      }
      if (qx >= 0 && qx < L / 2) {
        t[2 * pp] = fma(av.x, taps.lo[2 * qx], t[2 * pp]);
        t[2 * pp] = fma(dv.x, hi_tap<L>(taps, 2 * qx), t[2 * pp]);
        t[2 * pp + 1] = fma(av.x, taps.lo[2 * qx + 1], t[2 * pp + 1]);
        t[2 * pp + 1] = fma(dv.x, hi_tap<L>(taps, 2 * qx + 1), t[2 * pp + 1]);
      }
    }
  }
}

// ---- tile mode ------------------------------------------------------------------------------------
// At level k every parent owes T >> k slots the tile keeps - T / (2 kRS) groups over all parents, a
// power of two, so the warps are always full and (parent, group) is a shift and a mask - plus F_k
// slots of left extension for the levels below (none at level 1).  The extension is a separate short
// step, two slots per lane, run by a dedicated extra warp (JWC_WPT_TAIL_WARP = 1) or by one of the main
// warps, rotating with the CTA and the level (0); see jwc_wpt_fwd.cu.
template <int L, int kRS, bool INPLACE>
__global__ void __launch_bounds__(512)
k_wpt_rev_tile(const __grid_constant__ Taps taps, const __grid_constant__ WptRevArgs a) {
  extern __shared__ double2 smem2[];
  constexpr int lgRS = (kRS == 8) ? 3 : 2;
  static_assert(kRS == 8 || kRS == 4, "kRS is 4 or 8");
  const int tid = rotated_tid(a.rot), nthr = blockDim.x, nmain = nthr - 32 * JWC_WPT_TAIL_WARP;
  const int m = a.m, h0 = a.h0, T = a.T;
  const int64_t line = blockIdx.x >> a.lg_tpl;
  const int tile = int(blockIdx.x) & (a.tiles_per_line - 1);
  const int t0 = tile * T;
  double2* cur = smem2;
  double2* nxt = INPLACE ? smem2 : smem2 + a.buf_cap;
  {
    // stage all 2^m leaf packets: local sample i of node j is slot O + i (periodic) of packet j
    const int wm = h0 >> m;
    const int O = (t0 >> m) - a.F[m] - a.ru8;
    const int per_node = a.len[m] / 2;
    const double* src = a.src + line * a.src_os;
    const int total = per_node << m;
    int node = 0;
    for (int it = tid, j2 = tid; it < total; it += nthr, j2 += nthr) {
      while (j2 >= per_node) { j2 -= per_node; ++node; }
      cp_async16(&cur[node * a.cap[m] + pad2(j2)], src + node * wm + ((O + 2 * j2) & (wm - 1)));
    }
    cp_async_wait_all();
    __syncthreads();
  }
  // one group of kRS slots of parent `par`: t[2 kRS] from the children's windows in `cur`
  auto main_step = [&](int k, int par, int g, double (&t)[2 * kRS]) {
    const int cap_in = a.cap[k];
    if constexpr (kRS == 8) {
      const double2* A = cur + (2 * par) * cap_in + 5 * (a.g0[k] + g);  // pad2(4g'+3-w) = 5g' + (3-w) + floor((3-w)/4)
      const double2* D = A + cap_in;
      wrev_step<L, 8>(taps, [&](int w) { return A[(3 - w) + ((3 - w) >> 2)]; },
                      [&](int w) { return D[(3 - w) + ((3 - w) >> 2)]; }, t);
    } else {
      const double2* A = cur + (2 * par) * cap_in;
      const double2* D = A + cap_in;
      const int c = 4 * a.g0[k] + (kRS / 2) * g + kRS / 2 - 1;
      wrev_step<L, kRS>(taps, [&](int w) { return A[pad2(c - w)]; }, [&](int w) { return D[pad2(c - w)]; }, t);
    }
  };
  auto main_store = [&](int k, int par, int g, int gl, const double (&t)[2 * kRS]) {
    if (k > 1) {
      // pad2(kRS g + e) == kRS g + (kRS / 4) g + e + (e >> 2)
      double2* Y = nxt + par * a.cap[k - 1] + (kRS + kRS / 4) * g;
      if constexpr (kRS == 8) {
        // the lanes of an STS.128 phase are 10 slots apart - groups g and g + 4 share a bank group.
        // Lanes with bit 2 of g set store their upper four slots first: slot e ^ 4 sits 5 padded
        // slots from slot e, an odd distance, which separates the two halves of the phase.
        const bool rot = (g >> 2) & 1;
        double2* Ylo = Y + (rot ? 5 : 0);
        double2* Yhi = Y - (rot ? 5 : 0);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          Ylo[e] = make_double2(rot ? t[2 * e + 8] : t[2 * e], rot ? t[2 * e + 9] : t[2 * e + 1]);
#pragma unroll
        for (int e = 4; e < 8; ++e)
          Yhi[e + 1] = make_double2(rot ? t[2 * e - 8] : t[2 * e], rot ? t[2 * e - 7] : t[2 * e + 1]);
      } else {
#pragma unroll
        for (int e = 0; e < kRS; ++e) Y[e + (e >> 2)] = make_double2(t[2 * e], t[2 * e + 1]);
      }
    } else {
      double* y = a.dst + line * a.dst_os + t0 + 2 * kRS * (g - gl);
#pragma unroll
      for (int e = 0; e < kRS / 2; ++e) st_global_v4(y + 4 * e, t[4 * e], t[4 * e + 1], t[4 * e + 2], t[4 * e + 3]);
    }
  };
  // two slots of left extension (tail step number g of parent `par`)
  auto tail_step = [&](int k, int par, int g, double (&t)[4]) {
    const double2* A = cur + (2 * par) * a.cap[k];
    const double2* D = A + a.cap[k];
    const int c = 4 * a.g0[k] + g;
    wrev_step<L, 2>(taps, [&](int w) { return A[pad2(c - w)]; }, [&](int w) { return D[pad2(c - w)]; }, t);
  };
  auto tail_store = [&](int k, int par, int g, const double (&t)[4]) {
    double2* Y = nxt + par * a.cap[k - 1];
    Y[pad2(2 * g)] = make_double2(t[0], t[1]);
    Y[pad2(2 * g + 1)] = make_double2(t[2], t[3]);
  };

  if constexpr (INPLACE) {
    // One item per thread and level (the launcher checks it): results wait in registers until every
    // window of the level has been read, then overwrite the level's input - one buffer instead of two,
    // more CTAs per SM, one more barrier per level (see jwc_wpt_fwd.cu).
    for (int k = m; k >= 1; --k) {
      double t[2 * kRS];
      int par = 0, g = 0;
      bool has = false;
      const int gl = a.F[k] >> lgRS;
      if (tid < nmain) {
        const int lg_gpp = a.lg_T - k - lgRS;
        par = tid >> lg_gpp;
        g = gl + (tid & ((1 << lg_gpp) - 1));
        main_step(k, par, g, t);
      } else if (k > 1) {
        const int per_par = a.F[k] >> 1;
        g = tid - nmain;
        has = g < (per_par << (k - 1));
        if (has) {
          while (g >= per_par) { g -= per_par; ++par; }
          double t4[4];
          tail_step(k, par, g, t4);
          t[0] = t4[0]; t[1] = t4[1]; t[2] = t4[2]; t[3] = t4[3];
        }
      }
      if (k > 1) __syncthreads();
      if (tid < nmain) {
        main_store(k, par, g, gl, t);
      } else if (has) {
        const double t4[4] = {t[0], t[1], t[2], t[3]};
        tail_store(k, par, g, t4);
      }
      if (k > 1) __syncthreads();
    }
  } else {
    for (int k = m; k >= 1; --k) {
      if (tid < nmain) {
        const int lg_gpp = a.lg_T - k - lgRS;  // kept groups per parent = (T >> k) / kRS
        const int gl = a.F[k] >> lgRS;         // groups of left extension in front of them
        for (int it = tid; it < ((T >> 1) >> lgRS); it += nmain) {
          const int par = it >> lg_gpp, g = gl + (it & ((1 << lg_gpp) - 1));
          double t[2 * kRS];
          main_step(k, par, g, t);
          main_store(k, par, g, gl, t);
        }
      }
      const int tail_warp = JWC_WPT_TAIL_WARP ? (nmain >> 5) : int((blockIdx.x + k) % unsigned(nthr >> 5));
      if (k > 1 && (tid >> 5) == tail_warp) {
        const int per_par = a.F[k] >> 1;  // tail steps per parent (F_k is a multiple of 8)
        const int items = per_par << (k - 1);
        int par = 0;
        for (int it = tid & 31, g = it; it < items; it += 32, g += 32) {
          while (g >= per_par) { g -= per_par; ++par; }
          double t[4];
          tail_step(k, par, g, t);
          tail_store(k, par, g, t);
        }
      }
      __syncthreads();
      double2* tmp = cur; cur = nxt; nxt = tmp;
    }
  }
}

// ---- resident mode --------------------------------------------------------------------------------
template <int L, int kRS>
__global__ void __launch_bounds__(512)
k_wpt_rev_res(const __grid_constant__ Taps taps, const __grid_constant__ WptRevArgs a) {
  extern __shared__ double2 smem2[];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int m = a.m, h0 = a.h0;
  {

    const int G = a.G;
    const int64_t line0 = int64_t(blockIdx.x) * G;
    const int nlines = int(min(int64_t(G), a.lines - line0));
    const int cap = a.buf_cap;
    double2* cur = smem2;
    double2* nxt = smem2 + size_t(G) * cap;
    {
      const int per_line = h0 >> 1;
      for (int it = tid; it < nlines * per_line; it += nthr) {
        const int ln = it / per_line, k2 = it - ln * per_line;
        cp_async16(&cur[ln * cap + pad2(k2)], a.src + (line0 + ln) * a.src_os + 2 * k2);
      }
      cp_async_wait_all();
      __syncthreads();
    }
    for (int k = m; k >= 1; --k) {
      const int half = h0 >> k;      // width of every node at level k
      const bool last = (k == 1);
      if (half >= kRS) {
        const int gpp = half / kRS;                 // groups per parent
        const int per_line = gpp << (k - 1);        // == h0 / 16
        const int mask2 = (half >> 1) - 1;
        for (int it = tid; it < nlines * per_line; it += nthr) {
          const int ln = it / per_line, r = it - ln * per_line;
          const int par = r / gpp, g = r - par * gpp;
          const double2* cl = cur + ln * cap;
          const int offA = (2 * par) * (half >> 1), offD = offA + (half >> 1);
          const int c = (kRS / 2) * g + kRS / 2 - 1;
          double t[2 * kRS];
          wrev_step<L, kRS>(taps, [&](int w) { return cl[pad2(offA + ((c - w) & mask2))]; },
                       [&](int w) { return cl[pad2(offD + ((c - w) & mask2))]; }, t);
          if (!last) {
            double2* y = nxt + ln * cap;
            const int o = par * half + kRS * g;     // parent starts at sample par * 2 * half
#pragma unroll
            for (int e = 0; e < kRS; ++e) y[pad2(o + e)] = make_double2(t[2 * e], t[2 * e + 1]);
          } else {
            double* y = a.dst + (line0 + ln) * a.dst_os + 2 * kRS * g;
#pragma unroll
            for (int e = 0; e < kRS / 2; ++e) st_global_v4(y + 4 * e, t[4 * e], t[4 * e + 1], t[4 * e + 2], t[4 * e + 3]);
          }
        }
      } else {
        // nodes of 1, 2 or 4 coefficients: one thread per (line, parent, slot), true modular wrap
        const int per_line = half << (k - 1);       // == h0 / 2
        const int mask = half - 1;
        for (int it = tid; it < nlines * per_line; it += nthr) {
          const int ln = it / per_line, r = it - ln * per_line;
          const int par = r / half, p = r - par * half;
          const double2* cl = cur + ln * cap;
          const int offA = (2 * par) * half, offD = offA + half;
          double t0v = 0.0, t1v = 0.0;
#pragma unroll
          for (int q = 0; q < L / 2; ++q) {
            const int i = (p - q) & mask;
            const double av = sm_scalar(cl, offA + i), dv = sm_scalar(cl, offD + i);
            t0v = fma(av, taps.lo[2 * q], t0v);
            t0v = fma(dv, hi_tap<L>(taps, 2 * q), t0v);
            t1v = fma(av, taps.lo[2 * q + 1], t1v);
            t1v = fma(dv, hi_tap<L>(taps, 2 * q + 1), t1v);
          }
          if (!last) {
            nxt[ln * cap + pad2(par * half + p)] = make_double2(t0v, t1v);
          } else {
            double* y = a.dst + (line0 + ln) * a.dst_os + 2 * p;
            y[0] = t0v;
            y[1] = t1v;
          }
        }
      }
      __syncthreads();
      double2* tmp = cur; cur = nxt; nxt = tmp;
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------

static int round_up8(int v) { return (v + 7) & ~7; }

// Fills the per-level geometry of a tile-mode launch and returns its shared memory (bytes).
static size_t wpt_rev_tile_geometry(int L, WptRevArgs& a) {
  a.ru8 = round_up8(L / 2 - 1);
  int N = 0;
  for (int k = 1; k <= a.m; ++k) {
    a.F[k] = round_up8((N + 1) / 2);
    N = a.F[k] + L / 2 - 1;
  }
  a.F[a.m + 1] = 0;
  int cap = 0;
  for (int k = 1; k <= a.m; ++k) {
    a.len[k] = (k == a.m) ? (a.T >> k) + a.F[k] + a.ru8 : (a.T >> k) + 2 * a.F[k + 1];
    a.g0[k] = (k == a.m) ? a.ru8 / 8 : (2 * a.F[k + 1] - a.F[k]) / 8;
    a.cap[k] = pad2_size(a.len[k] / 2);
    const int c = (1 << k) * a.cap[k];
    if (c > cap) cap = c;
  }
  a.cap[0] = 0;
  a.buf_cap = cap;
  return size_t(2) * cap * sizeof(double2);
}

int wpt_rev_tile_levels(int L, int T, int want, size_t smem_limit) {
  WptRevArgs a;
  a.T = T;
  int m = 1;
  while (m < want && m < kMaxFuse && (T >> (m + 1)) >= 8) {
    a.m = m + 1;
    if (wpt_rev_tile_geometry(L, a) > smem_limit) break;
    ++m;
  }
  return m;
}

template <int L>
static cudaError_t launch_L(jwc_ctx* ctx, const Taps& taps, WptRevArgs a, bool resident) {
  size_t smem;
  int64_t grid;
  bool inplace = false;
  if (!resident) {
    if (a.m < 1 || a.m > kMaxFuse || (a.T >> a.m) < 8 || (a.T & (a.T - 1)) || ctx->wpt_threads < 32 + 32 * JWC_WPT_TAIL_WARP ||
        ctx->wpt_threads % 32)
      return cudaErrorInvalidValue;
    smem = wpt_rev_tile_geometry(L, a);
    inplace = JWC_WPT_TAIL_WARP && ctx->wpt_inplace && ctx->wpt_threads - 32 == (a.T / 2) / (ctx->wpt_rs == 4 ? 4 : 8);
    for (int k = 2; k <= a.m; ++k)  // tail steps of a level: one per lane of the tail warp
      if (((a.F[k] >> 1) << (k - 1)) > 32) inplace = false;
    if (inplace) smem /= 2;
    a.tiles_per_line = a.h0 / a.T;
    a.rot = (JWC_WPT_TAIL_WARP && ctx->rot_warps) ? 1 : 0;
    auto ilog2 = [](int v) { int l = 0; while ((1 << l) < v) ++l; return l; };
    a.lg_tpl = ilog2(a.tiles_per_line);
    a.lg_T = ilog2(a.T);
    grid = a.lines * a.tiles_per_line;
  } else {
    a.buf_cap = pad2_size(max(1, a.h0 / 2));
    smem = size_t(2) * a.G * a.buf_cap * sizeof(double2);
    grid = (a.lines + a.G - 1) / a.G;
  }
  if (grid > 0x7fffffff) return cudaErrorInvalidConfiguration;
  auto kern = resident ? (ctx->wpt_rs == 4 ? k_wpt_rev_res<L, 4> : k_wpt_rev_res<L, 8>)
                       : inplace ? (ctx->wpt_rs == 4 ? k_wpt_rev_tile<L, 4, true> : k_wpt_rev_tile<L, 8, true>)
                                 : (ctx->wpt_rs == 4 ? k_wpt_rev_tile<L, 4, false> : k_wpt_rev_tile<L, 8, false>);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
  }
  prof_begin(ctx, resident ? "k_wpt_rev:resident" : "k_wpt_rev:tile", double(a.lines) * a.h0, a.m);
  kern<<<int(grid), ctx->wpt_threads, smem, ctx->stream>>>(taps, a);
  prof_end(ctx);
  ctx->launches++;
  return cudaGetLastError();
}

cudaError_t launch_wpt_rev(jwc_ctx* ctx, int L, const Taps& taps, const WptRevArgs& a, bool resident) {
  switch (L) {
#define JWC_CASE(LL) case LL: return launch_L<LL>(ctx, taps, a, resident);
    JWC_FOR_EACH_L(JWC_CASE)
#undef JWC_CASE
  }
  return cudaErrorInvalidValue;
}

}  // namespace jwc
