// jwc_strided2.cuh - device helpers of the second-generation strided-axis kernels (jwc_fwt_strided2.cu).
//
// Geometry as in jwc_strided.cuh: element s of line (o, c) lives at base + o * os + s * inner + c.  A CTA
// owns kC2 = 16 adjacent lines - ONE 128-byte line of HBM per sample row - and a run of rows along the
// axis, held in shared memory as dense [row][16] doubles, which is exactly what a TMA box {16 columns,
// B rows} writes (SWIZZLE_NONE).  A thread owns TWO adjacent columns: thread (l8 = tid % 8, grp = tid / 8)
// reads column pair l8 of a row with one LDS.128, so the 8 lanes of a quarter-warp phase cover one whole
// 128-byte row: every shared-memory access of these kernels is conflict-free by construction (no row
// permutation, no padding), loads are `base + s * 128` immediates, and every global store is a
// 16-byte STG of which 8 lanes fill one 128-byte line.
//
// Why: the first-generation kernels (8 columns, one per thread, LDS.64) spent 57 % of the shared-memory
// pipe's cycles per FP64-pipe cycle at their DFMA density and paid another 44 % of wavefronts in bank
// conflicts on the TMA-staged rows (profiles/r01_ncu_summary.md): they were shared-memory bound, not
// FP64 bound.  Levels are computed IN PLACE (results wait in registers across one barrier and then
// overwrite the level's input), which halves the footprint of a CTA: 2-3 CTAs of 9-10 warps per SM.
#pragma once
#include "jwc_strided.cuh"

namespace jwc {

constexpr int kC2 = 16;        // columns per CTA
constexpr int kL8 = 8;         // threads per row (two columns each) == double2 per row
constexpr int kR2 = 4;         // outputs (forward) / slots (reverse) of one task
constexpr int kBoxF = 16;      // TMA box rows: forward staging (2 KB boxes)
constexpr int kStagesF = 4;     // forward staging: mbarriers the boxes are counted on, in row order
constexpr int kBoxR = 8;       // reverse staging (1 KB boxes: the left extensions are 8-row aligned)
constexpr int kMaxThr2 = 320;   // tile-mode CTA size bound: 2 CTAs per SM at <= 96 registers
constexpr int kMaxThrRes2 = 256;  // resident-mode CTA size bound (lines of up to 512 rows)

__device__ __forceinline__ void fma2(double2& acc, const double2& v, double tap) {
  acc.x = fma(v.x, tap, acc.x);
  acc.y = fma(v.y, tap, acc.y);
}

// forward: outputs i = R g .. R g + R - 1 of a column pair; x(s) = input rows 2 R g + s, s = 0 .. 2R + L - 3
// (Wavelet.java:244-254, j ascending, FMA-contracted).  HI = false: a halo group, low pass only.
template <int L, int R, bool HI, class X>
__device__ __forceinline__ void fwd_run2(const Taps& taps, int z, X x, double2 (&lo)[R], double2 (&hi)[R]) {
#pragma unroll
  for (int r = 0; r < R; ++r) lo[r] = hi[r] = make_double2(0.0, 0.0);
#pragma unroll
  for (int s = 0; s < 2 * R + L - 2; ++s) {
    const double2 v = x(s);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int j = s - 2 * r;
      if (j >= 0 && j < L) {
        fma2(lo[r], v, lo_tap<L>(taps, j, z));
        if constexpr (HI) fma2(hi[r], v, hi_tap<L>(taps, j, z));
      }
    }
  }
}

// reverse (gather form of Wavelet.java:288-299): slots p = R g .. R g + R - 1 of a column pair give
// t[2 pp + r] = sum_q a[p - q] lo[2q + r] + d[p - q] hi[2q + r];  a(s) / d(s) = coefficient rows at slot
// R g + R - 1 - s, s = 0 .. R + L/2 - 2 (walking left)
template <int L, int R, class A, class D>
__device__ __forceinline__ void rev_run2(const Taps& taps, int z, A a, D d, double2 (&t)[2 * R]) {
#pragma unroll
  for (int r = 0; r < 2 * R; ++r) t[r] = make_double2(0.0, 0.0);
#pragma unroll
  for (int s = 0; s < R + L / 2 - 1; ++s) {
    const double2 av = a(s), dv = d(s);
#pragma unroll
    for (int pp = 0; pp < R; ++pp) {
      const int q = s - (R - 1 - pp);
      if (q >= 0 && q < L / 2) {
        fma2(t[2 * pp], av, lo_tap<L>(taps, 2 * q, z));
        fma2(t[2 * pp + 1], av, lo_tap<L>(taps, 2 * q + 1, z));
        fma2(t[2 * pp], dv, hi_tap<L>(taps, 2 * q, z));
        fma2(t[2 * pp + 1], dv, hi_tap<L>(taps, 2 * q + 1, z));
      }
    }
  }
}

// 16-byte global store of a column pair
__device__ __forceinline__ void st2(double* p, const double2& v) { *reinterpret_cast<double2*>(p) = v; }

}  // namespace jwc
