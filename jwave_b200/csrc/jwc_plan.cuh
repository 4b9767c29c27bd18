// jwc_plan.cuh - the level planner: turns one axis transform into kernel launches.
#pragma once
#include "jwc_internal.cuh"

namespace jwc {

// 1-D transform (kind = JWC_FWT | JWC_WPT) along the middle axis of a dense [outer][n][inner]
// array, `level` steps, in -> out (no overlap).  Arguments are already validated.
// Follows FastWaveletTransform.java:71-153 / WaveletPacketTransform.java:73-191.
cudaError_t run_axis(jwc_ctx* ctx, const WaveletRec& w, int kind, int dir, const double* in, double* out,
                     int64_t outer, int n, int64_t inner, int level);

cudaError_t ensure_scratch(jwc_ctx* ctx, int slot, size_t bytes, double** ptr);

}  // namespace jwc
