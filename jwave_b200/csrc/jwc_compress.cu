// jwc_compress.cu - CompressorMagnitude (compressions/CompressorMagnitude.java:52-118 over
// compressions/Compressor.java:97-110): magnitude = mean |c| over the whole array, then
// c -> (|c| >= magnitude * threshold ? c : 0).  The step JWave runs right after the forward
// transform in its compression use case; on the GPU it is two streaming passes over the
// coefficients (reduce, then threshold), HBM-bound at 8 + 16 bytes per coefficient.
//
// The reduction is deterministic (fixed grid, fixed order of partial sums) but not the reference's
// left-to-right sum, so the magnitude agrees to rounding (~1e-16 relative), not bit for bit.
#include "jwc_internal.cuh"

namespace jwc {

constexpr int kRedThreads = 256;

__global__ void __launch_bounds__(kRedThreads)
k_abs_partial(const double* __restrict__ x, int64_t n, double* __restrict__ partial) {
  double s = 0.0;
  for (int64_t i = blockIdx.x * int64_t(kRedThreads) + threadIdx.x; i < n; i += int64_t(gridDim.x) * kRedThreads)
    s += fabs(x[i]);
  __shared__ double sh[kRedThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kRedThreads / 32; ++w) t += sh[w];
    partial[blockIdx.x] = t;
  }
}

__global__ void k_abs_final(const double* __restrict__ partial, int blocks, int64_t n, double* __restrict__ magnitude) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int b = 0; b < blocks; ++b) t += partial[b];
    *magnitude = t / double(n);  // _magnitude /= (double)arrHilbLength
  }
}

__global__ void __launch_bounds__(256)
k_threshold(const double* __restrict__ x, double* __restrict__ y, int64_t n, const double* __restrict__ magnitude,
            double threshold) {
  const double cut = *magnitude * threshold;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const double v = x[i];
    y[i] = (fabs(v) >= cut) ? v : 0.0;  // Compressor.java:103-107
  }
}

// scratch layout: [0 .. blocks) partial sums, [blocks] the magnitude
cudaError_t launch_compress_magnitude(jwc_ctx* ctx, const double* in, double* out, int64_t n, double threshold,
                                      double* scratch, int blocks) {
  prof_begin(ctx, "k_abs_partial", double(n), 0);
  k_abs_partial<<<blocks, kRedThreads, 0, ctx->stream>>>(in, n, scratch);
  prof_end(ctx);
  k_abs_final<<<1, 32, 0, ctx->stream>>>(scratch, blocks, n, scratch + blocks);
  int64_t tb = (n + 255) / 256;
  const int64_t cap = int64_t(ctx->sm_count) * 16;
  if (tb > cap) tb = cap;
  prof_begin(ctx, "k_threshold", double(n), 0);
  k_threshold<<<int(tb), 256, 0, ctx->stream>>>(in, out, n, scratch + blocks, threshold);
  prof_end(ctx);
  ctx->launches += 3;
  return cudaGetLastError();
}

}  // namespace jwc
