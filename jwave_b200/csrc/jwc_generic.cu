// jwc_generic.cu - one decomposition / reconstruction level, global -> global.
//
// Replaces Wavelet.forward / Wavelet.reverse (transforms/wavelets/Wavelet.java:236-260,
// :277-303) for ANY geometry and width: contiguous or strided lines, h >= 2, including the
// h < L case where the reference wraps several times (`while (k >= n) k -= n`).  h is a
// power of two, so the wrap is a mask.  Used for the shapes the fused kernels do not cover
// and as the on-device cross-check of those kernels (JWC_FORCE_GENERIC=1).
#include "jwc_internal.cuh"

namespace jwc {

// forward: out[i] = sum_j x[(2i+j) mod h] * lo[j],  out[i+h/2] = sum_j x[(2i+j) mod h] * hi[j]
// One thread per (line, i); threads run along `inner` first, so strided lines stay coalesced.
__global__ void __launch_bounds__(256)
k_fwd_level_generic(const __grid_constant__ Taps taps, int L, FwdLevelArgs a) {
  const int64_t total = a.outer * a.half * a.inner;
  const int mask = 2 * a.half - 1;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t c = idx % a.inner;
    const int64_t t = idx / a.inner;
    const int i = int(t % a.half);
    const int64_t o = t / a.half;
    const double* x = a.src + o * a.src_os + c;
    double lo = 0.0, hi = 0.0;
    for (int j = 0; j < L; ++j) {
      const double v = x[int64_t((2 * i + j) & mask) * a.inner];
      lo = fma(v, taps.lo[j], lo);
      hi = fma(v, taps.hi[j], hi);
    }
    a.dstA[o * a.dstA_os + int64_t(i) * a.inner + c] = lo;
    a.dstD[o * a.dstD_os + int64_t(i) * a.inner + c] = hi;
  }
}

// reverse, gather form of the reference's scatter: with k = 2p + r and j = 2q + r,
//   t[2p+r] = sum_q a[(p-q) mod h/2] * lo[2q+r] + d[(p-q) mod h/2] * hi[2q+r]
// One thread per (line, p) produces t[2p] and t[2p+1].
__global__ void __launch_bounds__(256)
k_rev_level_generic(const __grid_constant__ Taps taps, int L, RevLevelArgs a) {
  const int64_t total = a.outer * a.half * a.inner;
  const int mask = a.half - 1;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t c = idx % a.inner;
    const int64_t t = idx / a.inner;
    const int p = int(t % a.half);
    const int64_t o = t / a.half;
    const double* ca = a.srcA + o * a.srcA_os + c;
    const double* cd = a.srcD + o * a.srcD_os + c;
    double t0 = 0.0, t1 = 0.0;
    for (int q = 0; q < L / 2; ++q) {
      const int64_t i = int64_t((p - q) & mask) * a.inner;
      const double av = ca[i], dv = cd[i];
      t0 = fma(av, taps.lo[2 * q], t0);
      t0 = fma(dv, taps.hi[2 * q], t0);
      t1 = fma(av, taps.lo[2 * q + 1], t1);
      t1 = fma(dv, taps.hi[2 * q + 1], t1);
    }
    double* y = a.dst + o * a.dst_os + c;
    y[int64_t(2 * p) * a.inner] = t0;
    y[int64_t(2 * p + 1) * a.inner] = t1;
  }
}

static int grid_for(const jwc_ctx* ctx, int64_t total) {
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = int64_t(ctx->sm_count) * 32;  // grid-stride beyond 32 CTAs per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return int(blocks);
}

cudaError_t launch_fwd_level_generic(jwc_ctx* ctx, int L, const Taps& taps, const FwdLevelArgs& a) {
  const int64_t total = a.outer * a.half * a.inner;
  if (total <= 0) return cudaSuccess;
  prof_begin(ctx, "k_fwd_level_generic", 2.0 * double(total), 1);
  k_fwd_level_generic<<<grid_for(ctx, total), 256, 0, ctx->stream>>>(taps, L, a);
  prof_end(ctx);
  ctx->launches++;
  return cudaGetLastError();
}

cudaError_t launch_rev_level_generic(jwc_ctx* ctx, int L, const Taps& taps, const RevLevelArgs& a) {
  const int64_t total = a.outer * a.half * a.inner;
  if (total <= 0) return cudaSuccess;
  prof_begin(ctx, "k_rev_level_generic", 2.0 * double(total), 1);
  k_rev_level_generic<<<grid_for(ctx, total), 256, 0, ctx->stream>>>(taps, L, a);
  prof_end(ctx);
  ctx->launches++;
  return cudaGetLastError();
}

}  // namespace jwc
