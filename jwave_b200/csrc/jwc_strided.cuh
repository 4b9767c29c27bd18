// jwc_strided.cuh - device helpers of the strided-axis kernels (matrix columns, volume axes).
//
// Geometry: element s of line (o, c) lives at base + o * os + s * inner + c (jwc_internal.cuh).
// A CTA owns kC adjacent lines (columns c0 .. c0 + kC - 1, 64 contiguous bytes per sample) and a
// run of samples along the axis.  Shared-memory layout: [sample row][kC] doubles.  Thread
// (c = tid % kC, g = tid / kC) produces a run of outputs of column c, so a warp reads 4 rows x 64
// B per LDS.64.  Rows are permuted r -> r ^ ((r >> 3) & 1): groups are 8 rows apart, and without
// the swap the 4 rows of a warp would all start on the same 16 banks (4 wavefronts instead of 2).
#pragma once
#include "jwc_fused.cuh"

namespace jwc {

constexpr int kC = 8;            // columns per CTA
constexpr int kSR = 4;           // outputs (forward) / slots (reverse) per thread and step

__device__ __forceinline__ int srow(int r) { return r ^ ((r >> 3) & 1); }
__device__ __forceinline__ double& sat(double* buf, int r, int c) { return buf[srow(r) * kC + c]; }
__device__ __forceinline__ const double& sat(const double* buf, int r, int c) { return buf[srow(r) * kC + c]; }

// stage `rows` sample rows of kC columns: row r <- global sample ((first + r) mod width) of the line
// block starting at `gsrc` (sample stride `inner` doubles); 16 bytes per cp.async
__device__ __forceinline__ void stage_rows(double* buf, const double* gsrc, int64_t inner, int first, int rows, int wmask) {
  for (int it = threadIdx.x; it < rows * (kC / 2); it += blockDim.x) {
    const int r = it / (kC / 2), c2 = it - r * (kC / 2);
    cp_async16(&buf[srow(r) * kC + 2 * c2], gsrc + int64_t((first + r) & wmask) * inner + 2 * c2);
  }
}

// forward: outputs i = R g .. R g + R - 1 of one column; x(s) = input sample 2 R g + s,
// s = 0 .. 2R + L - 3 (Wavelet.java:244-254, j ascending, FMA-contracted)
template <int L, int R, class X>
__device__ __forceinline__ void fwd_run(const Taps& taps, X x, double (&lo)[R], double (&hi)[R]) {
#pragma unroll
  for (int r = 0; r < R; ++r) lo[r] = hi[r] = 0.0;
#pragma unroll
  for (int s = 0; s < 2 * R + L - 2; ++s) {
    const double v = x(s);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int j = s - 2 * r;
      if (j >= 0 && j < L) {
        lo[r] = fma(v, taps.lo[j], lo[r]);
        hi[r] = fma(v, hi_tap<L>(taps, j), hi[r]);
      }
    }
  }
}

// reverse (gather form of Wavelet.java:288-299): slots p = RS g .. RS g + RS - 1 of one column give
// t[2 pp + r] = sum_q a[p - q] lo[2q + r] + d[p - q] hi[2q + r];  a(s) / d(s) = coefficient at slot
// RS g + RS - 1 - s, s = 0 .. RS + L/2 - 2 (walking left)
template <int L, int RS, class A, class D>
__device__ __forceinline__ void rev_run(const Taps& taps, A a, D d, double (&t)[2 * RS]) {
#pragma unroll
  for (int r = 0; r < 2 * RS; ++r) t[r] = 0.0;
#pragma unroll
  for (int s = 0; s < RS + L / 2 - 1; ++s) {
    const double av = a(s), dv = d(s);
#pragma unroll
    for (int pp = 0; pp < RS; ++pp) {
      const int q = s - (RS - 1 - pp);
      if (q >= 0 && q < L / 2) {
        t[2 * pp] = fma(av, taps.lo[2 * q], t[2 * pp]);
        t[2 * pp] = fma(dv, hi_tap<L>(taps, 2 * q), t[2 * pp]);
        t[2 * pp + 1] = fma(av, taps.lo[2 * q + 1], t[2 * pp + 1]);
        t[2 * pp + 1] = fma(dv, hi_tap<L>(taps, 2 * q + 1), t[2 * pp + 1]);
      }
    }
  }
}

}  // namespace jwc
