// jwc_shfl.cu - forward FWT for 2-tap filters (Haar1, Legendre1, Haar1Orthogonal) entirely in registers and warp
// shuffles: the north star's "short filters use warp shuffles", built so that it can be measured against the
// shared-memory tile kernel (JWC_TUNE shfl=0|1, profiles/r02_ab_shuffle.txt).
//
// FastWaveletTransform.forward (FastWaveletTransform.java:88-97) around Wavelet.forward (Wavelet.java:236-260) with
// L = 2: a[i] = x[2i] h0 + x[2i+1] h1, d[i] = x[2i] g0 + x[2i+1] g1 - no window overlap, hence no halo and no
// periodic wrap inside a level.  A warp owns 256 consecutive samples of a line (8 per lane, two 32-byte loads):
// levels 1-3 are computed inside each lane's registers (8 -> 4 -> 2 -> 1), levels 4-8 by butterflies over
// __shfl_xor (lane distance 1, 2, 4, 8, 16).  No shared memory, no barrier, up to 8 levels per launch; every d_k
// goes straight to its final place and a_m to the next pass's input.
#include "jwc_fused.cuh"
#include "jwc_kernels.cuh"

namespace jwc {

constexpr int kShflSamples = 256;  // samples per warp
constexpr int kShflMax = 8;        // levels per launch

__global__ void __launch_bounds__(256)
k_fwt_fwd_shfl2(const __grid_constant__ Taps taps, const __grid_constant__ FwtFwdArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int wpl = a.h / kShflSamples;  // warps per line
  const int64_t line = warp / wpl;
  if (line >= a.lines) return;
  const int seg = int(warp - line * wpl);  // 256-sample segment of the line
  const double h0 = taps.lo[0], h1 = taps.lo[1], g0 = taps.hi[0], g1 = taps.hi[1];
  const double* src = a.src + line * a.src_os + seg * kShflSamples + 8 * lane;
  double* outD = a.dstD + line * a.dstD_os;
  double* outA = a.dstA + line * a.dstA_os;
  const int m = a.m, h = a.h;
  const double4 u = *reinterpret_cast<const double4*>(src), v = *reinterpret_cast<const double4*>(src + 4);
  const double x[8] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
  // levels 1-3 in registers; pos = index of this lane's first output at the level
  double a1[4], d1[4], a2[2], d2[2], a3, d3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a1[i] = fma(x[2 * i + 1], h1, x[2 * i] * h0);
    d1[i] = fma(x[2 * i + 1], g1, x[2 * i] * g0);
  }
  {
    double* p = outD + (h >> 1) + seg * (kShflSamples >> 1) + 4 * lane;
    st_global_v4(p, d1[0], d1[1], d1[2], d1[3]);
    if (m == 1) st_global_v4(outA + seg * (kShflSamples >> 1) + 4 * lane, a1[0], a1[1], a1[2], a1[3]);
  }
  if (m == 1) return;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    a2[i] = fma(a1[2 * i + 1], h1, a1[2 * i] * h0);
    d2[i] = fma(a1[2 * i + 1], g1, a1[2 * i] * g0);
  }
  *reinterpret_cast<double2*>(outD + (h >> 2) + seg * (kShflSamples >> 2) + 2 * lane) = make_double2(d2[0], d2[1]);
  if (m == 2) {
    *reinterpret_cast<double2*>(outA + seg * (kShflSamples >> 2) + 2 * lane) = make_double2(a2[0], a2[1]);
    return;
  }
  a3 = fma(a2[1], h1, a2[0] * h0);
  d3 = fma(a2[1], g1, a2[0] * g0);
  outD[(h >> 3) + seg * (kShflSamples >> 3) + lane] = d3;
  if (m == 3) {
    outA[seg * (kShflSamples >> 3) + lane] = a3;
    return;
  }
  // levels 4 .. m: butterflies; after step s the lanes with the low s + 1 bits clear hold a_{4+s}
  double cur = a3;
#pragma unroll
  for (int s = 0; s < kShflMax - 3; ++s) {
    const int k = 4 + s;
    if (k > m) break;
    const double other = __shfl_xor_sync(0xffffffffu, cur, 1 << s);
    const bool owner = (lane & ((2 << s) - 1)) == 0;  // holds the EARLIER sample of the pair
    const double an = fma(other, h1, cur * h0), dn = fma(other, g1, cur * g0);
    const int idx = seg * (kShflSamples >> k) + (lane >> (s + 1));
    if (owner) {
      outD[(h >> k) + idx] = dn;
      if (k == m) outA[idx] = an;
    }
    cur = an;  // only meaningful on owner lanes; the others are never read as `cur` of an owner again
  }
}

// Takes passes of a forward FWT with a mirrored 2-tap filter over contiguous lines whose width is a multiple of 256.
// cudaErrorNotSupported: not this kernel's shape - nothing was launched.
cudaError_t launch_fwt_fwd_shfl(jwc_ctx* ctx, int L, const Taps& taps, const FwtFwdArgs& a) {
  if (L != 2 || a.h % kShflSamples || a.m < 1 || a.m > kShflMax) return cudaErrorNotSupported;
  if ((reinterpret_cast<uintptr_t>(a.src) | reinterpret_cast<uintptr_t>(a.dstD) | reinterpret_cast<uintptr_t>(a.dstA)) & 31)
    return cudaErrorNotSupported;
  if ((a.src_os | a.dstD_os | a.dstA_os) & 3) return cudaErrorNotSupported;
  const int64_t warps = a.lines * (a.h / kShflSamples);
  const int64_t grid = (warps + 7) / 8;
  if (grid > 0x7fffffff || grid < 1) return cudaErrorNotSupported;
  prof_begin(ctx, "k_fwt_fwd_shfl", double(a.lines) * a.h, a.m);
  k_fwt_fwd_shfl2<<<int(grid), 256, 0, ctx->stream>>>(taps, a);
  prof_end(ctx);
  ctx->launches++;
  return cudaGetLastError();
}

}  // namespace jwc
