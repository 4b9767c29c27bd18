// jwc_transpose.cu - batched matrix transpose [B][R][C] -> [B][C][R] of doubles.
//
// The packet transform along a STRIDED axis of a 2-D / 3-D array (BasicTransform.java:361-399, :509-566 calling
// WaveletPacketTransform.java:73-191 on gathered columns) is run as transpose -> fused contiguous-line WPT ->
// transpose (jwc_plan.cu): 2 x 16 B per sample for the two transposes plus 16 B per fused pass, where the one-level
// kernels it replaces moved 32 B per sample and LEVEL.  The reference gathers every column into a fresh array and
// scatters it back; this is the same idea on whole tiles.
//
// One CTA = one 32 x 32 tile through shared memory (33-column rows: conflict-free in both directions), 256 threads,
// coalesced 256-byte rows on both sides; the tiles of all matrices form a 1-D grid.
#include "jwc_kernels.cuh"

namespace jwc {

__global__ void __launch_bounds__(256)
k_transpose(const double* __restrict__ in, double* __restrict__ out, int R, int C, int tiles_r, int tiles_c) {
  __shared__ double tile[32][33];
  const int64_t t = blockIdx.x;
  const int tc = int(t % tiles_c);
  const int64_t q = t / tiles_c;
  const int tr = int(q % tiles_r);
  const int64_t b = q / tiles_r;
  const double* src = in + b * int64_t(R) * C;
  double* dst = out + b * int64_t(R) * C;
  const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
  const int r0 = tr * 32, c0 = tc * 32;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int r = r0 + y + j, c = c0 + x;
    if (r < R && c < C) tile[y + j][x] = src[int64_t(r) * C + c];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int c = c0 + y + j, r = r0 + x;
    if (r < R && c < C) dst[int64_t(c) * R + r] = tile[x][y + j];
  }
}

cudaError_t launch_transpose(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int R, int C) {
  if (batch < 1 || R < 1 || C < 1) return cudaSuccess;
  const int tiles_r = (R + 31) / 32, tiles_c = (C + 31) / 32;
  const int64_t grid = batch * tiles_r * tiles_c;
  if (grid > 0x7fffffff) return cudaErrorInvalidConfiguration;
  prof_begin(ctx, "k_transpose", double(batch) * R * C, 0);
  k_transpose<<<int(grid), 256, 0, ctx->stream>>>(in, out, R, C, tiles_r, tiles_c);
  prof_end(ctx);
  ctx->launches++;
  return cudaGetLastError();
}

}  // namespace jwc
