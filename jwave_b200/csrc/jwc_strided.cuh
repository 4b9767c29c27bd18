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

// ---- TMA staging (Blackwell / Hopper bulk tensor copy) -------------------------------------------
// The level-0 tile of the strided kernels can be staged as 2-D boxes {8 columns, kBoxRows rows} of
// the matrix [rows][inner]: one elected thread issues cp.async.bulk.tensor.2d per box, completion is
// counted on an mbarrier.  The box lands DENSE ([row][8], 64-byte rows, SWIZZLE_NONE).  Measured on
// B200 (tools/tma_decode.py): with SWIZZLE_128B a 64-byte box row occupies a whole 128-byte line
// (half the staging buffer wasted), and SWIZZLE_64B only permutes chunks inside a row - neither
// gives the row-pair swap of srow(), so the level-1 window reads of a TMA-staged tile pay 4
// instead of 2 wavefronts per LDS.64.  The copy engine still wins (no per-16-byte address
// arithmetic, one instruction per 8 KB): +7 % on the 8192^2 Daubechies20 column pass, so TMA is the
// default where whole boxes fit (tiles / lines of >= kBoxRows rows); cp.async + srow() otherwise.
constexpr int kBoxRows = 128;

__device__ __forceinline__ const double& tma_at(const double* buf, int r, int c) { return buf[r * kC + c]; }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned phase) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(a), "r"(phase) : "memory");
}
// one box: columns [x, x + 8), rows [y, y + kBoxRows) of the tensor -> smem_dst (128-byte aligned)
__device__ __forceinline__ void tma_load_box(void* smem_dst, const void* tmap, int x, int y, uint64_t* bar) {
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  const unsigned b = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(d), "l"(tmap), "r"(x), "r"(y), "r"(b) : "memory");
}

// forward: outputs i = R g .. R g + R - 1 of one column; x(s) = input sample 2 R g + s,
// s = 0 .. 2R + L - 3 (Wavelet.java:244-254, j ascending, FMA-contracted)
template <int L, int R, class X>
__device__ __forceinline__ void fwd_run(const Taps& taps, int z, X x, double (&lo)[R], double (&hi)[R]) {
#pragma unroll
  for (int r = 0; r < R; ++r) lo[r] = hi[r] = 0.0;
#pragma unroll
  for (int s = 0; s < 2 * R + L - 2; ++s) {
    const double v = x(s);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int j = s - 2 * r;
      if (j >= 0 && j < L) {
        lo[r] = fma(v, lo_tap<L>(taps, j, z), lo[r]);
        hi[r] = fma(v, hi_tap<L>(taps, j, z), hi[r]);
      }
    }
  }
}

// reverse (gather form of Wavelet.java:288-299): slots p = RS g .. RS g + RS - 1 of one column give
// t[2 pp + r] = sum_q a[p - q] lo[2q + r] + d[p - q] hi[2q + r];  a(s) / d(s) = coefficient at slot
// RS g + RS - 1 - s, s = 0 .. RS + L/2 - 2 (walking left)
template <int L, int RS, class A, class D>
__device__ __forceinline__ void rev_run(const Taps& taps, int z, A a, D d, double (&t)[2 * RS]) {
#pragma unroll
  for (int r = 0; r < 2 * RS; ++r) t[r] = 0.0;
#pragma unroll
  for (int s = 0; s < RS + L / 2 - 1; ++s) {
    const double av = a(s), dv = d(s);
#pragma unroll
    for (int pp = 0; pp < RS; ++pp) {
      const int q = s - (RS - 1 - pp);
      if (q >= 0 && q < L / 2) {
        t[2 * pp] = fma(av, lo_tap<L>(taps, 2 * q, z), t[2 * pp]);
        t[2 * pp] = fma(dv, hi_tap<L>(taps, 2 * q, z), t[2 * pp]);
        t[2 * pp + 1] = fma(av, lo_tap<L>(taps, 2 * q + 1, z), t[2 * pp + 1]);
        t[2 * pp + 1] = fma(dv, hi_tap<L>(taps, 2 * q + 1, z), t[2 * pp + 1]);
      }
    }
  }
}

}  // namespace jwc
