// jwc_kernels.cuh - argument blocks and host launchers of the fused kernels.
#pragma once
#include "jwc_internal.cuh"

namespace jwc {

// every even filter length in scope: Haar1 (2) .. Daubechies20 / Symlet20 (40)
#ifndef JWC_FOR_EACH_L
#define JWC_FOR_EACH_L(X) \
  X(2) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(18) X(20) X(22) X(24) X(26) X(28) X(30) X(32) X(34) X(36) X(38) X(40)
#endif


// ---- batched transpose [batch][R][C] -> [batch][C][R] (jwc_transpose.cu) ------------------------
cudaError_t launch_transpose(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int R, int C);

// ---- forward FWT, contiguous lines (jwc_fwt_fwd.cu) -----------------------------------------
struct FwtFwdArgs {
  const double* src; int64_t src_os;    // input lines of width h (stride between lines, in doubles)
  double* dstD; int64_t dstD_os;        // final output lines: d_k lands at line + (h >> k) + ...
  double* dstA; int64_t dstA_os;        // where a_m goes (compact scratch or the final output)
  int64_t lines;
  int h;                                // current width
  int m;                                // levels fused in this launch
  int T;                                // tile length (tile mode)
  int G;                                // lines per CTA (resident mode)
  int tiles_per_line, lg_tpl, cap0, cap1, tail; // filled in by the launcher
  int rot;                              // rotate the warps' roles with the CTA number (rotated_tid)
};
// Longest filter whose FWT tile kernels get a tail warp: with both steps in one kernel body ptxas hoists
// the taps of longer filters out of the level loop into vector registers (reverse L = 40: 66 -> 132).
#ifndef JWC_TAIL_MAXL
#define JWC_TAIL_MAXL 24
#endif
constexpr int kTailMaxL = JWC_TAIL_MAXL;
int fwt_tile_levels(int L, int T);
cudaError_t launch_fwt_fwd(jwc_ctx* ctx, int L, const Taps& taps, const FwtFwdArgs& a, bool resident);
// jwc_shfl.cu: 2-tap filters, registers + warp shuffles, up to 8 levels per launch (cudaErrorNotSupported = declined)
cudaError_t launch_fwt_fwd_shfl(jwc_ctx* ctx, int L, const Taps& taps, const FwtFwdArgs& a);

// ---- reverse FWT, contiguous lines (jwc_fwt_rev.cu) -----------------------------------------
constexpr int kMaxFuse = 12;
struct FwtRevArgs {
  const double* srcA; int64_t srcA_os;  // a_m lines (width h0 >> m); resident mode reads a_m from srcD
  const double* srcD; int64_t srcD_os;  // coefficient lines: d_k at line + (h0 >> k)
  double* dst; int64_t dst_os;          // a_0 lines (width h0)
  int64_t lines;
  int h0, m, T, G;
  RemoteMap rm;                         // mode 2: the output lines go to peer slabs
  // filled in by the launcher
  int tiles_per_line, lg_tpl, ru8, tail, rot;
  int F[kMaxFuse + 2], g0[kMaxFuse + 1], len[kMaxFuse + 1], offD[kMaxFuse + 1], offA[2];
  int capC, capP[2];
};
cudaError_t launch_fwt_rev(jwc_ctx* ctx, int L, const Taps& taps, const FwtRevArgs& a, bool resident);

// ---- FWT along a strided axis (jwc_fwt_strided.cu) ------------------------------------------------
struct FwtFwdStrArgs {
  const double* src; int64_t src_os;    // element (o, s, c) at src + o * src_os + s * inner + c
  double* dstD; int64_t dstD_os;        // final output: d_k at rows (h >> k) ...
  double* dstA; int64_t dstA_os;        // a_m destination
  int64_t outer, inner;
  int h, m, T;
  RemoteMap rmD, rmA;                         // mode 1: the d_k rows / the a_m rows go to peer slabs
  int tiles_per_line, cblocks, rows0, rows1;  // filled in by the launcher
  int64_t rows_per_o;                         // tensor rows between consecutive `outer` slices (TMA path)
  int rounds1;                                // second generation: rounds of the first level
};
int fwt_str_tile_levels(int L, int T);
cudaError_t launch_fwt_fwd_str(jwc_ctx* ctx, int L, const Taps& taps, const FwtFwdStrArgs& a, bool resident);

struct FwtRevStrArgs {
  const double* srcA; int64_t srcA_os;
  const double* srcD; int64_t srcD_os;
  double* dst; int64_t dst_os;
  int64_t outer, inner;
  int h0, m, T;
  RemoteMap rm;                               // mode 1: the output rows go to peer slabs
  // filled in by the launcher
  int tiles_per_line, cblocks, ru, rowsC, rowsP[2];
  int F[kMaxFuse + 2], s0[kMaxFuse + 1], len[kMaxFuse + 1], offD[kMaxFuse + 1], offA[2];
  int64_t rowsA_per_o, rowsD_per_o;           // tensor rows between consecutive `outer` slices (TMA path)
  int nmain;                                  // second generation: threads that take level 1
};
cudaError_t launch_fwt_rev_str(jwc_ctx* ctx, int L, const Taps& taps, const FwtRevStrArgs& a, bool resident);

// Second generation (jwc_fwt_strided2.cu): 16 columns per CTA, two per thread, levels in place, TMA-staged
// in both directions.  Same argument blocks (offD / rowsC count ROWS there).  cudaErrorNotSupported = the
// shape is not covered (inner % 16, box alignment, CTA size): nothing was launched, use the first generation.
int fwt_str2_tile_levels(int L, int T, int want);
cudaError_t launch_fwt_fwd_str2(jwc_ctx* ctx, int L, const Taps& taps, const FwtFwdStrArgs& a, bool resident);
cudaError_t launch_fwt_rev_str2(jwc_ctx* ctx, int L, const Taps& taps, const FwtRevStrArgs& a, bool resident);

// ---- forward WPT, contiguous lines (jwc_wpt_fwd.cu) -----------------------------------------
struct WptFwdArgs {
  const double* src; int64_t src_os;    // input lines (packets of width h)
  double* dst; int64_t dst_os;          // output lines: 2^m leaf packets of width h >> m, natural order
  int64_t lines;
  int h, m, T, G;
  // filled in by the launcher
  int tiles_per_line, lg_tpl, lg_T, buf_cap, rot;
  int stagger_ns, stagger_ctas, stagger_div;  // first-wave stagger (A/B switch, jwc_fused.cuh)
  int pf_dist;                          // L2 prefetch distance in CTAs (0 = off)
  int tma_out;                          // tile mode, in place: the leaf segments leave through TMA stores
  int cap[kMaxFuse + 1];                // tile mode: per-node capacity (double2) of level k
};
int wpt_tile_levels(int L, int T, int want, size_t smem_limit, int R);
cudaError_t launch_wpt_fwd(jwc_ctx* ctx, int L, const Taps& taps, const WptFwdArgs& a, bool resident);

// ---- reverse WPT, contiguous lines (jwc_wpt_rev.cu) -----------------------------------------
struct WptRevArgs {
  const double* src; int64_t src_os;    // input lines: 2^m leaf packets of width h0 >> m
  double* dst; int64_t dst_os;          // output lines (one packet of width h0)
  int64_t lines;
  int h0, m, T, G;
  // filled in by the launcher
  int tiles_per_line, lg_tpl, lg_T, ru8, buf_cap, rot;
  int stagger_ns, stagger_ctas, stagger_div;  // first-wave stagger (A/B switch, jwc_fused.cuh)
  int pf_dist;                          // L2 prefetch distance in CTAs (0 = off)
  int stage_left, stage_len2, cap_m;    // tile mode, staging: F[m] + ru8, len[m] / 2, cap[m]
  int stage_lg_lpn;                     // staging: log2 threads per packet when packets are short, else -1
  int tma_out;                          // tile mode, in place: the finished tile leaves through a TMA store
  unsigned geo; int geo_ok;             // tile mode: capB | Fx << 16 | ru8 << 22 | lg T << 27 (jwc_wpt_rev.cu)
  int F[kMaxFuse + 2], g0[kMaxFuse + 1], len[kMaxFuse + 1], cap[kMaxFuse + 1];
};
int wpt_rev_tile_levels(int L, int T, int want, size_t smem_limit);
cudaError_t launch_wpt_rev(jwc_ctx* ctx, int L, const Taps& taps, const WptRevArgs& a, bool resident);

}  // namespace jwc
