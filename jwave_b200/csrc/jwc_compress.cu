// jwc_compress.cu - CompressorMagnitude (compressions/CompressorMagnitude.java:52-118 over
// compressions/Compressor.java:97-110): magnitude = mean |c| over the whole array, then
// c -> (|c| >= magnitude * threshold ? c : 0).  The step JWave runs right after the forward
// transform in its compression use case; on the GPU it is two streaming passes over the
// coefficients (reduce, then threshold), HBM-bound at 8 + 16 bytes per coefficient - or 16 when the reduce
// runs chunk by chunk right behind the forward transform, while the chunk is still in the 126 MB L2
// (jwc_fwt1d_compress_dev, jwc_capi.cu).
//
// The reduction is deterministic (fixed grid, fixed order of partial sums) but not the reference's
// left-to-right sum, so the magnitude agrees to rounding (~1e-16 relative), not bit for bit.
#include "jwc_internal.cuh"

namespace jwc {

constexpr int kRedThreads = 256;

// |x| summed per CTA (fixed grid-stride order, 32-byte loads), partial sums to `partial`; the CTA that finishes
// LAST adds the partials in index order and writes the magnitude - deterministic, one launch instead of two.
// `accumulate`: add to the partials already there (the L2-resident chunks of jwc_fwt1d_compress_dev).
__global__ void __launch_bounds__(kRedThreads)
k_abs_partial(const double* __restrict__ x, int64_t n, double* __restrict__ partial, unsigned* __restrict__ done,
              int64_t n_total, double* __restrict__ magnitude, int accumulate, int finish) {
  double s = 0.0;
  const int64_t n4 = n >> 2;
  const double4* x4 = reinterpret_cast<const double4*>(x);
  for (int64_t i = blockIdx.x * int64_t(kRedThreads) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * kRedThreads) {
    const double4 v = x4[i];
    s += (fabs(v.x) + fabs(v.y)) + (fabs(v.z) + fabs(v.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) s += fabs(x[(n4 << 2) + threadIdx.x]);
  __shared__ double sh[kRedThreads / 32];
  __shared__ bool last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kRedThreads / 32; ++w) t += sh[w];
    partial[blockIdx.x] = accumulate ? partial[blockIdx.x] + t : t;
    __threadfence();
    last = finish && atomicAdd(done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {  // every other CTA's partial is visible (fence + atomic): fixed-order tree over the partials
    double t = 0.0;
    for (int b = threadIdx.x; b < int(gridDim.x); b += kRedThreads) t += __ldcg(&partial[b]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      double m = 0.0;
      for (int w = 0; w < kRedThreads / 32; ++w) m += sh[w];
      *magnitude = m / double(n_total);  // _magnitude /= (double)arrHilbLength
      *done = 0;                         // ready for the next call on this stream
    }
  }
}

__global__ void __launch_bounds__(256)
k_threshold(const double* x, double* y, int64_t n, const double* __restrict__ magnitude, double threshold) {  // y may be x
  const double cut = *magnitude * threshold;
  const int64_t n4 = n >> 2;
  const double4* x4 = reinterpret_cast<const double4*>(x);
  double4* y4 = reinterpret_cast<double4*>(y);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    double4 v = x4[i];
    v.x = (fabs(v.x) >= cut) ? v.x : 0.0;  // Compressor.java:103-107
    v.y = (fabs(v.y) >= cut) ? v.y : 0.0;
    v.z = (fabs(v.z) >= cut) ? v.z : 0.0;
    v.w = (fabs(v.w) >= cut) ? v.w : 0.0;
    y4[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const double v = x[(n4 << 2) + threadIdx.x];
    y[(n4 << 2) + threadIdx.x] = (fabs(v) >= cut) ? v : 0.0;
  }
}

// scratch layout: [0 .. blocks) partial sums, [blocks] the magnitude, [blocks + 1] the CTA counter (zero between
// calls).  Arrays must be 32-byte aligned (every cudaMalloc'd array is).
cudaError_t launch_abs_sum(jwc_ctx* ctx, const double* in, int64_t n, int64_t n_total, double* scratch, int blocks,
                           bool accumulate, bool finish) {
  prof_begin(ctx, "k_abs_partial", double(n), 0);
  k_abs_partial<<<blocks, kRedThreads, 0, ctx->stream>>>(in, n, scratch, reinterpret_cast<unsigned*>(scratch + blocks + 1),
                                                          n_total, scratch + blocks, accumulate ? 1 : 0, finish ? 1 : 0);
  prof_end(ctx);
  ctx->launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_threshold(jwc_ctx* ctx, const double* in, double* out, int64_t n, double threshold, double* scratch,
                             int blocks) {
  int64_t tb = (n / 4 + 255) / 256;
  const int64_t cap = int64_t(ctx->sm_count) * 16;
  if (tb > cap) tb = cap;
  if (tb < 1) tb = 1;
  prof_begin(ctx, "k_threshold", double(n), 0);
  k_threshold<<<int(tb), 256, 0, ctx->stream>>>(in, out, n, scratch + blocks, threshold);
  prof_end(ctx);
  ctx->launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_compress_magnitude(jwc_ctx* ctx, const double* in, double* out, int64_t n, double threshold,
                                      double* scratch, int blocks) {
  cudaError_t e = launch_abs_sum(ctx, in, n, n, scratch, blocks, false, true);
  if (e != cudaSuccess) return e;
  return launch_threshold(ctx, in, out, n, threshold, scratch, blocks);
}

}  // namespace jwc
