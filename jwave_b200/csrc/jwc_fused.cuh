// jwc_fused.cuh - device helpers shared by the fused (multi-level, shared-memory) kernels.
#pragma once
#include <cuda_runtime.h>

#include "jwc_internal.cuh"

namespace jwc {

constexpr int kThreads = 256;  // upper bound of the CTA size of the strided kernels (launch bound)
constexpr int kR = 4;          // consecutive outputs (per filter) one thread produces per step

// Thread index with the CTA's warps rotated by the CTA number (JWC_TUNE rot_warps=1; OFF by default).  The
// tile kernels give their warps different roles - four main warps and a fifth, lighter tail warp.  If the
// hardware placed warp w of every CTA on SM sub-partition w mod 4, the tail warps of all resident CTAs would
// share sub-partition 0 with a main warp; rotating the roles with blockIdx was meant to spread them.  Measured:
// it is 5-7 % SLOWER on both the WPT and the FWT tile kernels (profiles/r02_ab_rot_warps.txt: Symlet8 WPT forward
// 0.725 -> 0.673 of the FP64 roofline, Daubechies4 FWT forward 0.976 -> 0.919 of the HBM peak), i.e. the
// unrotated placement is already the balanced one.  Kept as a switch because it is the experiment that says so.
__device__ __forceinline__ int rotated_tid(int rot) {
  if (!rot) return threadIdx.x;
  const unsigned nw = blockDim.x >> 5, w = threadIdx.x >> 5;
  return int((((w + blockIdx.x) % nw) << 5) | (threadIdx.x & 31));
}

// First-wave stagger (JWC_TUNE stagger=ns; OFF by default).  Every CTA of a tile kernel lives equally long and a new
// one starts when an old one retires, so the phase pattern of the first wave - all resident CTAs of an SM staging at
// once, then all in their FMA phase at once - can persist through the launch.  With the switch on, CTA b of the first
// wave (b < ctas) waits ((b / div) mod 8) * ns before it starts: div = SM count, i.e. the k-th CTA an SM receives.
__device__ __forceinline__ void first_wave_stagger(int ns, int ctas, int div) {
  if (ns > 0 && int(blockIdx.x) < ctas) {
    const unsigned wait = unsigned((blockIdx.x / unsigned(div)) & 7) * unsigned(ns);
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
      __nanosleep(200);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while (t1 - t0 < wait);
  }
}

// Padded shared-memory layout of a line segment: interleaved samples as double2 (x[2k], x[2k+1]), one
// pad slot after every 4 double2.  A thread that produces outputs 4g..4g+3 reads the window
// k = 4g .. 4g + L/2 + 2; consecutive threads are 5 slots (80 B) apart, so the 8 lanes of a
// quarter-warp LDS.128 phase hit 8 distinct 16-byte bank groups (20 t mod 32 words is a
// permutation) - no bank conflicts on the window loads; staging and 2-slot stores are 2-way.  Used by
// the resident kernels, the forward tile kernel for long filters and the WPT reverse kernel; the
// forward tile kernel for L <= 24 (fl, jwc_fwt_fwd.cu) and the FWT reverse kernels (lay, jwc_fwt_rev.cu)
// use XOR layouts instead (tests/test_smem_layouts.py enumerates all of them).
__device__ __forceinline__ int pad2(int k2) { return k2 + (k2 >> 2); }
__host__ __device__ constexpr int pad2_size(int n2) { return n2 + (n2 >> 2) + 1; }

// scalar view of a padded double2 line
__device__ __forceinline__ double sm_scalar(const double2* buf, int i) {
  return reinterpret_cast<const double*>(buf)[2 * pad2(i >> 1) + (i & 1)];
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most `pending` of the most recently committed groups are still in flight
__device__ __forceinline__ void cp_async_wait_pending(int pending) {
  switch (pending) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
  }
}

// 256-bit global store of 4 consecutive doubles (32-byte aligned): STG.E.ENL2.256 on sm_100a.
__device__ __forceinline__ void st_global_v4(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// Forward step for R consecutive output pairs i = R g' .. R g' + R - 1: `win(q)` returns double2
// number (R g' + q) of the input, q = 0 .. R + L/2 - 2 (same arithmetic as fwd_step4 below).
template <int L, int R, class Win>
__device__ __forceinline__ void fwd_stepR(const Taps& taps, Win win, double (&lo)[R], double (&hi)[R]) {
#pragma unroll
  for (int r = 0; r < R; ++r) lo[r] = hi[r] = 0.0;
  constexpr int W = L / 2 + R - 1;
  constexpr bool kPrefetch = (W <= 8);  // short filters: whole window in registers before the FMAs
  double2 wv[kPrefetch ? W : 1];
  if constexpr (kPrefetch) {
#pragma unroll
    for (int q = 0; q < W; ++q) wv[q] = win(q);
  }
#pragma unroll
  for (int q = 0; q < W; ++q) {
    double2 v;
    if constexpr (kPrefetch) v = wv[q];
    else v = win(q);
    // all .x updates first, then all .y updates: successive FMAs on one accumulator are 2R
    // instructions apart instead of 2 (the DFMA result latency showed up as `wait` stalls)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int jj = q - r;
      if (jj >= 0 && jj < L / 2) {
        lo[r] = fma(v.x, taps.lo[2 * jj], lo[r]);
        hi[r] = fma(v.x, hi_tap<L>(taps, 2 * jj), hi[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int jj = q - r;
      if (jj >= 0 && jj < L / 2) {
        lo[r] = fma(v.y, taps.lo[2 * jj + 1], lo[r]);
        hi[r] = fma(v.y, hi_tap<L>(taps, 2 * jj + 1), hi[r]);
      }
    }
  }
}

// Forward step for 4 consecutive output pairs from a window held in shared memory.
//   lo[r] = sum_j x[2(4g+r)+j] * taps.lo[j],  hi[r] likewise, j ascending (the reference's order,
//   Wavelet.java:244-254, contracted to FMAs).
// `win(q)` returns double2 number q of the window (q = 0 .. L/2 + 2).  Streaming over q keeps only
// one double2 live besides the 8 accumulators, so register use does not grow with L.
template <int L, class Win>
__device__ __forceinline__ void fwd_step4(const Taps& taps, Win win, double (&lo)[kR], double (&hi)[kR]) {
#pragma unroll
  for (int r = 0; r < kR; ++r) lo[r] = hi[r] = 0.0;
#pragma unroll
  for (int q = 0; q < L / 2 + kR - 1; ++q) {
    const double2 v = win(q);
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      const int jj = q - r;  // tap pair index for output r
      if (jj >= 0 && jj < L / 2) {
        lo[r] = fma(v.x, taps.lo[2 * jj], lo[r]);
        hi[r] = fma(v.x, hi_tap<L>(taps, 2 * jj), hi[r]);
        lo[r] = fma(v.y, taps.lo[2 * jj + 1], lo[r]);
        hi[r] = fma(v.y, hi_tap<L>(taps, 2 * jj + 1), hi[r]);
      }
    }
  }
}

}  // namespace jwc
