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
#include <cuda.h>

#include <cstring>

#include "jwc_fused.cuh"
#include "jwc_kernels.cuh"

#ifndef JWC_WPT_TAIL_WARP
#define JWC_WPT_TAIL_WARP 1
#endif

namespace jwc {

// Launch bound of the tile kernels.  -DJWC_WPT_MINB=n builds them for 160-thread CTAs with n CTAs per SM (an A/B
// switch for the register budget; the default keeps run-time CTA sizes up to 512).
#ifdef JWC_WPT_MINB
#define JWC_WPT_TILE_BOUNDS __launch_bounds__(160, JWC_WPT_MINB)
#else
#define JWC_WPT_TILE_BOUNDS __launch_bounds__(512)
#endif

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
//
// Shared-memory layout of a node: double2 k at k + (k >> 3), one pad slot per 8 (rl).  A thread's window is two
// 4-slot blocks, G and G - 1; the lanes of an LDS.128 phase read blocks 4 slots apart, positions floor(9 j / 2): the 8
// bank groups 0,4,1,5,2,6,3,7 when the phase starts at an even block, one colliding lane pair (2-way) when it starts at
// an odd one - so one of the two blocks pays (ncu: 6 of the 16 window loads at 2 wavefront passes).  The lanes of an
// STS.128 phase store runs of 8 slots, 9 positions apart: 8 distinct groups, no rotated store order.  The alternative
// (JWC_WPT_REV_PAD4=1: one pad per 4, stride 5 - every load phase conflict-free, stores rotated with 32 FSEL per level)
// measures the same or 1-2 % slower (profiles/r02_lsu_bound.md); every 4-slot block is contiguous in both, so a
// window is two block pointers plus compile-time offsets.
#ifndef JWC_WPT_REV_PAD4
#define JWC_WPT_REV_PAD4 0
#endif
#if JWC_WPT_REV_PAD4  // A/B: one pad slot per 4 (window loads conflict-free at any phase, rotated stores)
__device__ __forceinline__ int rl(int k2) { return k2 + (k2 >> 2); }
__host__ __device__ constexpr int rl_size(int n2) { return n2 + (n2 >> 2) + 2; }
constexpr int kTwoBlocks = 10;
#else
__device__ __forceinline__ int rl(int k2) { return k2 + (k2 >> 3); }
__host__ __device__ constexpr int rl_size(int n2) { return n2 + (n2 >> 3) + 2; }
constexpr int kTwoBlocks = 9;
#endif

// TMA store of the finished tile (cp.async.bulk.tensor, SASS UTMASTG): box {16 doubles, T / 16 rows} of the output seen
// as a [rows][16] matrix, read from a dense shared-memory image whose 16-byte chunks are XOR-swizzled by the row number
// (SWIZZLE_128B) - the pattern that makes the threads' 128-byte row stores conflict-free.
__device__ __forceinline__ void tma_store_tile(const void* tmap, const void* smem_src, int x, int y) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_src));
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(x), "r"(y), "r"(s)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the CTA's shared memory may go once it has been read
}

template <int L, int kRS, bool INPLACE>
__global__ void JWC_WPT_TILE_BOUNDS
k_wpt_rev_tile(const __grid_constant__ Taps taps, const __grid_constant__ WptRevArgs a, const __grid_constant__ CUtensorMap tmapOut) {
  extern __shared__ __align__(1024) double2 smem2[];
  __shared__ unsigned s_geo;
  constexpr int lgRS = (kRS == 8) ? 3 : 2;
  static_assert(kRS == 8 || kRS == 4, "kRS is 4 or 8");
  const int tid = rotated_tid(a.rot), nthr = blockDim.x, nmain = nthr - 32 * JWC_WPT_TAIL_WARP;
  const int m = a.m, h0 = a.h0, T = a.T;
  const int64_t line = blockIdx.x >> a.lg_tpl;
  const int tile = int(blockIdx.x) & (a.tiles_per_line - 1);
  const int t0 = tile * T;
  double2* cur = smem2;
  double2* nxt = INPLACE ? smem2 : smem2 + a.buf_cap;
  first_wave_stagger(a.stagger_ns, a.stagger_ctas, a.stagger_div);
  {
    // stage all 2^m leaf packets: local sample i of node j is slot O + i (periodic) of packet j.
    // Thread j2 copies double2 number j2 of EVERY packet: the periodic source offset and the padded slot are the
    // same for all of them, so the inner loop is one LDGSTS and two pointer increments (the flat loop over
    // (packet, j2) items cost 40 instructions per item - half of this kernel's non-DFMA instructions).
    const int wm = h0 >> m;
    const int O = (t0 >> m) - a.stage_left;
    const int per_node = a.stage_len2, capm = a.cap_m, nn = 1 << m;
    const double* src = a.src + line * a.src_os;
    if (a.stage_lg_lpn < 0) {
      for (int j2 = tid; j2 < per_node; j2 += nthr) {
        const double* sp = src + ((O + 2 * j2) & (wm - 1));
        double2* dp = cur + rl(j2);
#pragma unroll 4
        for (int node = 0; node < nn; ++node, sp += wm, dp += capm) cp_async16(dp, sp);
      }
    } else {
      // many short packets (a deep pass: 64 packets of 24 double2 at m = 6): 2^lg_lpn threads per packet, several
      // packets per round
      const int lpn = 1 << a.stage_lg_lpn, npi = nthr >> a.stage_lg_lpn;
      const int j2 = tid & (lpn - 1), n0 = tid >> a.stage_lg_lpn;
      if (j2 < per_node && n0 < npi) {
        const double* sp = src + int64_t(n0) * wm + ((O + 2 * j2) & (wm - 1));
        double2* dp = cur + n0 * capm + rl(j2);
        const int64_t ss = int64_t(npi) * wm;
        const int ds = npi * capm;
#pragma unroll 4
        for (int node = n0; node < nn; node += npi, sp += ss, dp += ds) cp_async16(dp, sp);
      }
    }
    if (tid == 0) s_geo = a.geo;
    // L2 prefetch of the leaf-packet segments a CTA `pf_dist` launches later will stage (one instruction per packet)
    if (a.pf_dist > 0 && tid < nn) {
      const int64_t t2 = int64_t(blockIdx.x) + a.pf_dist;
      if (t2 < (a.lines << a.lg_tpl)) {
        const int O2 = (((int(t2) & (a.tiles_per_line - 1)) * T) >> m) & ~15;  // the kept slots, 128-byte aligned
        const double* p2 = a.src + (t2 >> a.lg_tpl) * a.src_os + int64_t(tid) * wm + O2;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p2), "r"((T >> m) * 8) : "memory");
      }
    }
    cp_async_wait_all();
    __syncthreads();
  }
  // one group of kRS slots of parent `par`: t[2 kRS] from the children's windows.  G = first 4-slot block of the
  // group's window in the children (g0_k + group number, in units of 4 double2).
  auto main_step = [&](const double2* A, int cap_in, int G, double (&t)[2 * kRS]) {
    const double2* D = A + cap_in;
    if constexpr (kRS == 8) {
      // slot 4G + 3 - w sits in 4-block G - (w >> 2).  Blocks G, G - 2, .. are 9 positions apart from o1 = rl(4G),
      // blocks G - 1, G - 3, .. from o0 = rl(4 (G - 1)): two run-time offsets, everything else folds (block G - 1 is
      // never read for L = 2, whose window is 4 slots).
      const int o1 = rl(4 * G), o0 = rl(4 * (G - 1));
      auto at = [&](const double2* X, int w) {
        const int blk = w >> 2, e = 3 - (w & 3) - kTwoBlocks * (blk >> 1);
        return (blk & 1) ? X[o0 + e] : X[o1 + e];
      };
      wrev_step<L, 8>(taps, [&](int w) { return at(A, w); }, [&](int w) { return at(D, w); }, t);
    } else {
      // slots 2G' + 1 - w with G' = 2 g0 + group number (units of 2 double2): blocks of 2 never straddle a pad
      wrev_step<L, kRS>(taps, [&](int w) { return A[rl(2 * G + 1 - w)]; }, [&](int w) { return D[rl(2 * G + 1 - w)]; }, t);
    }
  };
  auto main_store = [&](int k, double2* Yn, int g, int gl, const double (&t)[2 * kRS]) {
    if (k > 1) {
      double2* Y = Yn + rl(kRS * g);  // kRS = 8: 9 g; kRS = 4: runs of 4 inside one 8-block
#if JWC_WPT_REV_PAD4
      if constexpr (kRS == 8) {
        // lanes 10 slots apart: groups g and g + 4 share a bank group; lanes with bit 2 of g set store their upper
        // four slots first (slot e ^ 4 sits 5 padded slots from slot e)
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
        for (int e = 0; e < kRS; ++e) Y[e] = make_double2(t[2 * e], t[2 * e + 1]);
      }
#else
#pragma unroll
      for (int e = 0; e < kRS; ++e) Y[e] = make_double2(t[2 * e], t[2 * e + 1]);
#endif
    } else {
      double* y = a.dst + line * a.dst_os + t0 + 2 * kRS * (g - gl);
#pragma unroll
      for (int e = 0; e < kRS / 2; ++e) st_global_v4(y + 4 * e, t[4 * e], t[4 * e + 1], t[4 * e + 2], t[4 * e + 3]);
    }
  };
  // two slots of left extension (tail step number g of a parent): slots 4 g0 + g - w of the children
  auto tail_step = [&](const double2* A, int cap_in, int c, double (&t)[4]) {
    const double2* D = A + cap_in;
    wrev_step<L, 2>(taps, [&](int w) { return A[rl(c - w)]; }, [&](int w) { return D[rl(c - w)]; }, t);
  };
  auto tail_store = [&](double2* Yn, int g, const double* t) {
    Yn[rl(2 * g)] = make_double2(t[0], t[1]);
    Yn[rl(2 * g + 1)] = make_double2(t[2], t[3]);
  };

  if constexpr (INPLACE) {
    // One item per thread and level (the launcher checks it): results wait in registers until every
    // window of the level has been read, then overwrite the level's input - one buffer instead of two,
    // more CTAs per SM, one more barrier per level (see jwc_wpt_fwd.cu).
    // The per-level geometry is UNPACKED from one register (a.geo, read back from shared memory so that ptxas keeps
    // it in a register) instead of being read from arrays indexed by the loop counter: every such read is an LDC,
    // ptxas re-issues them in every level rather than hold registers, and constant loads into vector registers share
    // the LSU data pipe with the window loads (profiles/r02_lsu_bound.md).  Geometry (launcher): every level k >= 2
    // computes the same Fx slots of left extension, node capacities are capB >> k.
    const unsigned geo = s_geo;
    const int capB = geo & 0xffff, Fx = (geo >> 16) & 63, ru8 = (geo >> 22) & 31, lgT = geo >> 27;
    for (int k = m; k >= 1; --k) {
      double t[2 * kRS];
      int par = 0, g = 0;
      const int Fk = k == 1 ? 0 : Fx;                          // F_1 == 0: no extension at the output level
      const int cap_in = capB >> k, cap_out = capB >> (k - 1);
      const int g0k = k == m ? (ru8 >> 3) : ((2 * Fx - Fk) >> 3);
      const int gl = Fk >> lgRS;
      // tail warp: Fk / 2 two-slot steps per parent, up to kTailSteps per lane (a deep pass has 2^(k-1) parents:
      // 128 steps at k = 6); their results share the registers of a main step
      constexpr int kTailSteps = (2 * kRS) / 4;
      const int per_par = Fk >> 1, items = per_par << (k - 1);
      if (tid < nmain) {
        const int lg_gpp = lgT - k - lgRS;
        par = tid >> lg_gpp;
        g = gl + (tid & ((1 << lg_gpp) - 1));
        main_step(cur + (2 * par) * cap_in, cap_in, (kRS == 8 ? g0k : 2 * g0k) + g, t);
      } else if (k > 1) {
#pragma unroll
        for (int i = 0; i < kTailSteps; ++i) {
          const int it = tid - nmain + 32 * i;
          if (it < items) {
            const int tp = it / per_par, tg = it - tp * per_par;
            double t4[4];
            tail_step(cur + (2 * tp) * cap_in, cap_in, 4 * g0k + tg, t4);
            t[4 * i] = t4[0]; t[4 * i + 1] = t4[1]; t[4 * i + 2] = t4[2]; t[4 * i + 3] = t4[3];
          }
        }
      }
      // Output level with a TMA store (a.tma_out): once every window of level 1 has been read the buffer is free;
      // thread g writes its 16 samples as row g of a dense [T / 16][16] image, chunk e at e ^ (g & 7) (conflict-free,
      // registers in program order), and one thread hands the image to the copy engine - no 32 x 32-byte STG per
      // warp and instruction, no store drain before the CTA retires (profiles/r02_lsu_bound.md, ablation table).
      const bool tma_out = (k == 1) && a.tma_out && kRS == 8;
      if (k > 1 || tma_out) __syncthreads();
      if (tid < nmain) {
        if (tma_out) {
          double2* row = smem2 + 8 * g;  // k == 1: one parent, gl == 0, g == tid
          const int x = g & 7;
#pragma unroll
          for (int e = 0; e < 8; ++e) row[e ^ x] = make_double2(t[2 * e], t[2 * e + 1]);
        } else {
          main_store(k, nxt + par * cap_out, g, gl, t);
        }
      } else if (k > 1) {
#pragma unroll
        for (int i = 0; i < kTailSteps; ++i) {
          const int it = tid - nmain + 32 * i;
          if (it < items) {
            const int tp = it / per_par, tg = it - tp * per_par;
            tail_store(nxt + tp * cap_out, tg, &t[4 * i]);
          }
        }
      }
      if (k > 1) __syncthreads();
      else if (tma_out) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the TMA
        __syncthreads();
        if (tid == 0) tma_store_tile(&tmapOut, smem2, 0, int((line * h0 + t0) >> 4));
      }
    }
  } else {
    for (int k = m; k >= 1; --k) {
      const int cap_in = a.cap[k], cap_out = a.cap[k - 1], g0k = a.g0[k];
      if (tid < nmain) {
        const int lg_gpp = a.lg_T - k - lgRS;  // kept groups per parent = (T >> k) / kRS
        const int gl = a.F[k] >> lgRS;         // groups of left extension in front of them
        for (int it = tid; it < ((T >> 1) >> lgRS); it += nmain) {
          const int par = it >> lg_gpp, g = gl + (it & ((1 << lg_gpp) - 1));
          double t[2 * kRS];
          main_step(cur + (2 * par) * cap_in, cap_in, (kRS == 8 ? g0k : 2 * g0k) + g, t);
          main_store(k, nxt + par * cap_out, g, gl, t);
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
          tail_step(cur + (2 * par) * cap_in, cap_in, 4 * g0k + g, t);
          tail_store(nxt + par * cap_out, g, t);
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
// Level k computes, besides the T >> k slots the tile keeps, F_k slots of left extension for the levels below:
// F_1 = 0 and F_k = Fx = round_up8(L / 2) for every k >= 2 - the fixed point of F = round_up8((F + L/2) / 2), i.e.
// enough at every depth (level 2 alone would get by with round_up8(L / 4)) - and a node of level k owns capB >> k
// double2 of a buffer.  One F and one capacity base: the kernel derives the whole geometry from the packed word
// a.geo instead of per-level tables.
static size_t wpt_rev_tile_geometry(int L, WptRevArgs& a) {
  a.ru8 = round_up8(L / 2 - 1);
  const int Fx = round_up8(L / 2);
  a.F[1] = 0;
  for (int k = 2; k <= a.m; ++k) a.F[k] = Fx;
  a.F[a.m + 1] = 0;
  int capB = 0;
  for (int k = 1; k <= a.m; ++k) {
    a.len[k] = (k == a.m) ? (a.T >> k) + a.F[k] + a.ru8 : (a.T >> k) + 2 * a.F[k + 1];
    a.g0[k] = (k == a.m) ? a.ru8 / 8 : (2 * a.F[k + 1] - a.F[k]) / 8;
    const int c = rl_size(a.len[k] / 2) << k;
    if (c > capB) capB = c;
  }
  capB = (capB + (1 << a.m) - 1) & ~((1 << a.m) - 1);
  for (int k = 1; k <= a.m; ++k) a.cap[k] = capB >> k;
  a.cap[0] = 0;
  a.buf_cap = capB;
  int lgT = 0;
  while ((1 << lgT) < a.T) ++lgT;
  a.geo = unsigned(capB) | unsigned(Fx) << 16 | unsigned(a.ru8) << 22 | unsigned(lgT) << 27;
  a.geo_ok = capB < (1 << 16) && Fx < 64 && a.ru8 < 32;
  return size_t(2) * capB * sizeof(double2);
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

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFnW)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnW encode_tiled_w() {
  static EncodeTiledFnW fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFnW>(p);
  }();
  return fn;
}
// the output lines, dense, as a [rows][16 doubles] matrix; box {16, box_rows}; 128-byte swizzle on the shared side
static bool make_out_tmap(CUtensorMap* map, const double* base, int64_t rows, int box_rows) {
  EncodeTiledFnW enc = encode_tiled_w();
  if (!enc || (reinterpret_cast<uintptr_t>(base) & 127) || rows < 1 || rows >= (int64_t(1) << 31) || box_rows < 1 || box_rows > 256)
    return false;
  const cuuint64_t dims[2] = {16, cuuint64_t(rows)};
  const cuuint64_t strides[1] = {16 * sizeof(double)};
  const cuuint32_t box[2] = {16, cuuint32_t(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
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
    inplace = JWC_WPT_TAIL_WARP && ctx->wpt_inplace && a.geo_ok && ctx->wpt_threads - 32 == (a.T / 2) / (ctx->wpt_rs == 4 ? 4 : 8);
    a.stage_left = a.F[a.m] + a.ru8;
    a.stage_len2 = a.len[a.m] / 2;
    a.cap_m = a.cap[a.m];
    for (int k = 2; k <= a.m; ++k)  // tail steps of a level: up to 4 (kRS = 8) / 2 (kRS = 4) per lane of the tail warp
      if (((a.F[k] >> 1) << (k - 1)) > 32 * (ctx->wpt_rs == 4 ? 2 : 4)) inplace = false;
    a.stage_lg_lpn = -1;
    if (a.stage_len2 * 2 <= ctx->wpt_threads) {  // short packets: several per staging round
      a.stage_lg_lpn = 0;
      while ((1 << a.stage_lg_lpn) < a.stage_len2) ++a.stage_lg_lpn;
    }
    if (inplace) smem /= 2;
    a.tiles_per_line = a.h0 / a.T;
    a.rot = (JWC_WPT_TAIL_WARP && ctx->rot_warps) ? 1 : 0;
    a.stagger_ns = ctx->stagger;
    a.pf_dist = ctx->pf;
    a.stagger_div = ctx->sm_count;
    a.stagger_ctas = ctx->sm_count * 8;
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
  if (resident) {
    auto rk = ctx->wpt_rs == 4 ? k_wpt_rev_res<L, 4> : k_wpt_rev_res<L, 8>;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
      if (e != cudaSuccess) return e;
    }
    prof_begin(ctx, "k_wpt_rev:resident", double(a.lines) * a.h0, a.m);
    rk<<<int(grid), ctx->wpt_threads, smem, ctx->stream>>>(taps, a);
    prof_end(ctx);
    ctx->launches++;
    return cudaGetLastError();
  }
  auto kern = inplace ? (ctx->wpt_rs == 4 ? k_wpt_rev_tile<L, 4, true> : k_wpt_rev_tile<L, 8, true>)
                      : (ctx->wpt_rs == 4 ? k_wpt_rev_tile<L, 4, false> : k_wpt_rev_tile<L, 8, false>);
  CUtensorMap tmapOut;
  memset(&tmapOut, 0, sizeof tmapOut);
  a.tma_out = 0;
  if (inplace && ctx->wpt_tma_store && ctx->wpt_rs != 4 && a.dst_os == a.h0 && a.T >= 256 && a.T <= 4096 &&
      smem >= size_t(a.T) * sizeof(double) && make_out_tmap(&tmapOut, a.dst, a.lines * (a.h0 / 16), a.T / 16))
    a.tma_out = 1;
  smem += size_t(ctx->xsmem) << 10;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
  }
  if (ctx->carve) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  prof_begin(ctx, "k_wpt_rev:tile", double(a.lines) * a.h0, a.m);
  kern<<<int(grid), ctx->wpt_threads, smem, ctx->stream>>>(taps, a, tmapOut);
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
