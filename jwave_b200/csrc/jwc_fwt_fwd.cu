// jwc_fwt_fwd.cu - fused multi-level forward FWT along contiguous lines.
//
// Replaces the level loop of FastWaveletTransform.forward (FastWaveletTransform.java:88-97)
// around Wavelet.forward (Wavelet.java:236-260) for `m` consecutive levels per launch: the
// intermediate approximations a_1 .. a_{m-1} live only in shared memory.
//
//   tile mode      (h > res_cap) : one CTA = one tile of T level-0 samples of one line plus a
//                                 right-hand periodic halo of (2^m - 1)(L - 2) samples; details
//                                 d_1..d_m go to their final place, a_m to `dstA`.  For L <= 24 the
//                                 halo approximations of every level are a tail warp's job.
//   resident mode  (h <= res_cap): one CTA = G whole lines; the periodic wrap is an index mask, so
//                                 every remaining level (down to h = 2) runs in this launch.
//
// HBM traffic per launch: h samples read (+ halo re-reads that hit L2), h samples written.
#include "jwc_fused.cuh"
#include "jwc_kernels.cuh"

namespace jwc {

// Tile-mode layout.  Short and medium filters (L <= kXorMaxL): unpadded, double2 k at k ^ ((k >> 3) & 3) -
// conflict-free for the window loads (lanes 4 slots apart), the a_k stores (2 slots apart) and the
// cp.async staging (consecutive slots), where the padded layout is 2-way conflicted in the last two
// (tests/test_smem_layouts.py), and 20 % smaller.  The XOR costs two integer instructions per window load,
// which the FP64-bound long filters cannot spare: they keep pad2 (compile-time offsets from one base).
#ifndef JWC_FWD_XOR
#define JWC_FWD_XOR 1
#endif
constexpr int kXorMaxL = JWC_FWD_XOR ? 24 : 0;
template <int L> __device__ __forceinline__ int fl(int k2) {
  if constexpr (L <= kXorMaxL) return k2 ^ ((k2 >> 3) & 3);
  else return pad2(k2);
}
template <int L> static constexpr int fl_size(int n2) {
  return L <= kXorMaxL ? ((n2 + 3) & ~3) : pad2_size(n2);
}

template <int L, bool RESIDENT, int R = 4>
__global__ void __launch_bounds__(384)
k_fwt_fwd(const __grid_constant__ Taps taps, const FwtFwdArgs a) {
  extern __shared__ double2 smem2[];
  const int tid = rotated_tid(a.rot), nthr = blockDim.x;

  if constexpr (!RESIDENT) {
    // ---------------- tile mode ----------------
    const int64_t line = blockIdx.x >> a.lg_tpl;  // tiles per line: a power of two
    const int tile = int(blockIdx.x) & (a.tiles_per_line - 1);
    const int T = a.T, m = a.m, h = a.h;
    const int H0 = ((1 << m) - 1) * (L - 2);
    const int n0 = T + H0;                       // samples staged at level 0 (even)
    double2* cur = smem2;                        // level k-1 samples
    double2* nxt = smem2 + a.cap0;               // level k approximations
    const double* src = a.src + line * a.src_os;
    const int base = tile * T;
    if constexpr (L <= kXorMaxL) {
      for (int k2 = tid; k2 < n0 / 2; k2 += nthr)
        cp_async16(&cur[fl<L>(k2)], src + ((base + 2 * k2) & (h - 1)));   // periodic wrap of the halo
    } else {
      // Long filters are bound by instruction issue: running pointers instead of a mask, a 64-bit multiply-add and
      // a layout call per element (25 instructions each before).  Only the last tile of a line wraps, once
      // (halo <= T / 8 < h); nthr is a multiple of 4, so a step of nthr double2 is nthr + nthr / 4 padded slots.
      const int n2 = n0 >> 1, fit2 = min(n2, (h - base) >> 1), step = nthr + (nthr >> 2);
      const double* sp = src + base + 2 * tid;
      double2* dp = cur + pad2(tid);
      int k2 = tid;
      for (; k2 < fit2; k2 += nthr, sp += 2 * nthr, dp += step) cp_async16(dp, sp);
      sp -= h;
#pragma unroll 1
      for (; k2 < n2; k2 += nthr, sp += 2 * nthr, dp += step) cp_async16(dp, sp);
    }
    cp_async_wait_all();
    __syncthreads();

    double* outD = a.dstD + line * a.dstD_os;
    for (int k = 1; k <= m; ++k) {
      const int n_det = T >> k;                                   // details this tile owns
      const int n_out = n_det + ((1 << (m - k)) - 1) * (L - 2);   // approximations incl. halo
      // With a tail warp (a.tail, L <= kTailMaxL) the halo approximations are its job, two per lane and
      // low pass only; the main warps then run exactly n_det / R groups, a power of two.
      const int nmain = nthr - 32 * a.tail;
      const int groups = a.tail ? n_det / R : (n_out + R - 1) / R;
      const bool last = (k == m);
      double* dD = outD + (h >> k) + tile * n_det;
      double* dA = a.dstA + line * a.dstA_os + tile * n_det;      // used when last
      if constexpr (L <= kTailMaxL) {
        if (tid >= nmain) {  // only with a.tail; nothing to do at the last level (no halo)
          for (int j = tid - nmain; j < (n_out - n_det) / 2; j += 32) {
            const int o = n_det + 2 * j;  // first of the two outputs == first double2 of their window
            double lo2[2], hi2[2];
            fwd_stepR<L, 2>(taps, [&](int q) { return cur[fl<L>(o + q)]; }, lo2, hi2);
            nxt[fl<L>(o >> 1)] = make_double2(lo2[0], lo2[1]);
          }
        }
      }
      for (int g = tid; g < groups && tid < nmain; g += nmain) {
        double lo[R], hi[R];
        if constexpr (R == 4) {
          if constexpr (L <= kXorMaxL) {
            fwd_stepR<L, R>(taps, [&](int q) { return cur[fl<L>(R * g + q)]; }, lo, hi);
          } else {
            const double2* w = cur + pad2(R * g);   // pad2(4g + q) == 5g + q + (q >> 2)
            fwd_stepR<L, R>(taps, [&](int q) { return w[q + (q >> 2)]; }, lo, hi);
          }
          if (!last) {
            nxt[fl<L>(2 * g)] = make_double2(lo[0], lo[1]);
            nxt[fl<L>(2 * g + 1)] = make_double2(lo[2], lo[3]);
          } else {
            st_global_v4(dA + R * g, lo[0], lo[1], lo[2], lo[3]);
          }
          if (R * g < n_det) st_global_v4(dD + R * g, hi[0], hi[1], hi[2], hi[3]);
        } else {
          static_assert(R == 2 || R == 4, "R is 2 or 4");
          fwd_stepR<L, R>(taps, [&](int q) { return cur[fl<L>(R * g + q)]; }, lo, hi);
          if (!last) nxt[fl<L>(g)] = make_double2(lo[0], lo[1]);
          else *reinterpret_cast<double2*>(dA + R * g) = make_double2(lo[0], lo[1]);
          if (R * g < n_det) *reinterpret_cast<double2*>(dD + R * g) = make_double2(hi[0], hi[1]);
        }
      }
      __syncthreads();
      double2* t = cur; cur = nxt; nxt = t;
    }
  } else {
    // ---------------- resident mode ----------------
    const int h = a.h, m = a.m, G = a.G;
    const int64_t line0 = int64_t(blockIdx.x) * G;
    const int nlines = int(min(int64_t(G), a.lines - line0));
    const int capA = a.cap0, capB = a.cap1;   // per-line capacities (double2) of the two buffers
    double2* bufA = smem2;
    double2* bufB = smem2 + size_t(G) * capA;
    {
      const int per_line = h >> 1;            // double2 per line, >= 1
      const int total = nlines * per_line;
      for (int it = tid; it < total; it += nthr) {
        const int ln = it / per_line, k2 = it - ln * per_line;
        cp_async16(&bufA[ln * capA + pad2(k2)], a.src + (line0 + ln) * a.src_os + 2 * k2);
      }
      cp_async_wait_all();
      __syncthreads();
    }
    double2* cur = bufA; int cur_cap = capA;
    double2* nxt = bufB; int nxt_cap = capB;
    for (int k = 1; k <= m; ++k) {
      const int h_in = h >> (k - 1), h_out = h_in >> 1;
      const int gpl = max(1, h_out / kR);     // groups per line (power of two)
      const int mask2 = (h_in >> 1) - 1;      // wrap mask in double2 units
      const bool last = (k == m);
      const int items = nlines * gpl;
      for (int it = tid; it < items; it += nthr) {
        const int ln = it / gpl, g = it - ln * gpl;
        const double2* cl = cur + ln * cur_cap;
        double lo[kR], hi[kR];
        fwd_step4<L>(taps, [&](int q) { return cl[pad2((kR * g + q) & mask2)]; }, lo, hi);
        double* dD = a.dstD + (line0 + ln) * a.dstD_os + h_out + kR * g;
        double* dA = a.dstA + (line0 + ln) * a.dstA_os + kR * g;
        if (h_out >= kR) {
          if (!last) {
            double2* nl = nxt + ln * nxt_cap;
            nl[pad2(2 * g)] = make_double2(lo[0], lo[1]);
            nl[pad2(2 * g + 1)] = make_double2(lo[2], lo[3]);
          } else {
            st_global_v4(dA, lo[0], lo[1], lo[2], lo[3]);
          }
          st_global_v4(dD, hi[0], hi[1], hi[2], hi[3]);
        } else {
          // h_out is 1 or 2: outputs beyond h_out are periodic duplicates - drop them
          if (!last) nxt[ln * nxt_cap] = make_double2(lo[0], lo[1]);   // h_out == 2 (h_out == 1 is always last)
          else { dA[0] = lo[0]; if (h_out == 2) dA[1] = lo[1]; }
          dD[0] = hi[0];
          if (h_out == 2) dD[1] = hi[1];
        }
      }
      __syncthreads();
      double2* t = cur; cur = nxt; nxt = t;
      const int tc = cur_cap; cur_cap = nxt_cap; nxt_cap = tc;
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------

int fwt_tile_levels(int L, int T) {
  // largest m with halo (2^m - 1)(L - 2) <= T / 8 and T / 2^m >= 4 (256-bit stores stay aligned)
  int m = 1;
  while (((1 << (m + 1)) - 1) * (L - 2) <= T / 8 && (T >> (m + 1)) >= kR) ++m;
  return m;
}

template <int L>
static cudaError_t launch_L(jwc_ctx* ctx, const Taps& taps, FwtFwdArgs a, bool resident) {
  size_t smem;
  int grid;
  if (!resident) {
    const int H0 = ((1 << a.m) - 1) * (L - 2);
    const int H1 = ((1 << (a.m - 1)) - 1) * (L - 2);
    a.cap0 = fl_size<L>((a.T + H0) / 2 + 4);
    a.cap1 = fl_size<L>(((a.T >> 1) + H1) / 2 + 4);
    smem = size_t(a.cap0 + a.cap1) * sizeof(double2);
    a.tiles_per_line = a.h / a.T;
    a.lg_tpl = 0;
    while ((1 << a.lg_tpl) < a.tiles_per_line) ++a.lg_tpl;
    if ((1 << a.lg_tpl) != a.tiles_per_line || ctx->fwd_threads % 4) return cudaErrorInvalidValue;
    const int64_t ctas = a.lines * a.tiles_per_line;
    if (ctas > 0x7fffffff) return cudaErrorInvalidConfiguration;
    grid = int(ctas);
  } else {
    a.cap0 = pad2_size(max(1, a.h / 2));
    a.cap1 = pad2_size(max(1, a.h / 4));
    smem = size_t(a.G) * (a.cap0 + a.cap1) * sizeof(double2);
    grid = int((a.lines + a.G - 1) / a.G);
  }
  auto kern = resident ? k_fwt_fwd<L, true> : (ctx->fwd_r == 2 ? k_fwt_fwd<L, false, 2> : k_fwt_fwd<L, false, 4>);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
  }
  if (ctx->carve) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  prof_begin(ctx, resident ? "k_fwt_fwd:resident" : "k_fwt_fwd:tile", double(a.lines) * a.h, a.m);
  a.tail = (!resident && ctx->fwd_tail && L <= kTailMaxL) ? 1 : 0;
  a.rot = (a.tail && ctx->rot_warps) ? 1 : 0;
  kern<<<grid, resident ? ctx->res_threads : ctx->fwd_threads + 32 * a.tail, smem, ctx->stream>>>(taps, a);
  prof_end(ctx);
  ctx->launches++;
  return cudaGetLastError();
}

cudaError_t launch_fwt_fwd(jwc_ctx* ctx, int L, const Taps& taps, const FwtFwdArgs& a, bool resident) {
  switch (L) {
#define JWC_CASE(LL) case LL: return launch_L<LL>(ctx, taps, a, resident);
    JWC_FOR_EACH_L(JWC_CASE)
#undef JWC_CASE
  }
  return cudaErrorInvalidValue;
}

}  // namespace jwc
