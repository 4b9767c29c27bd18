// jwc_plan.cuh - the level planner: turns one axis transform into kernel launches.
#pragma once
#include "jwc_internal.cuh"

namespace jwc {

// 1-D transform (kind = JWC_FWT | JWC_WPT) along the middle axis of a dense [outer][n][inner]
// array, `level` steps, in -> out (no overlap).  Arguments are already validated.
// Follows FastWaveletTransform.java:71-153 / WaveletPacketTransform.java:73-191.
cudaError_t run_axis(jwc_ctx* ctx, const WaveletRec& w, int kind, int dir, const double* in, double* out,
                     int64_t outer, int n, int64_t inner, int level);

// True if a contiguous FWT over lines of `n` samples that are pitch_in / pitch_out doubles apart (ctx->pitch_in,
// ctx->pitch_out) takes the fused kernels, which honour the pitches; the one-level and strided plans assume dense lines.
bool fwt_pitched_ok(const jwc_ctx* ctx, const WaveletRec& w, int dir, const double* in, const double* out, int n,
                    int64_t pitch_in, int64_t pitch_out);

cudaError_t ensure_scratch(jwc_ctx* ctx, int slot, size_t bytes, double** ptr);

}  // namespace jwc
