// jwc_internal.cuh - shared declarations of libjwave_cuda.so (not part of the public ABI).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include "jwave_cuda.h"

namespace jwc {

// One filter pair.  Passed BY VALUE as a __grid_constant__ kernel parameter, so the taps sit in
// the constant bank and an unrolled `taps.lo[j]` becomes a c[0x0][imm] operand of the DFMA.
struct Taps {
  double lo[JWC_MAX_TAPS];  // scaling (low pass)
  double hi[JWC_MAX_TAPS];  // wavelet (high pass)
};

// Every in-scope family builds its high pass from the low pass (Wavelet.java:104-122; Haar1.java:62
// writes the same thing out by hand): hi[j] = (j even ? + : -) lo[L - 1 - j].  The fused kernels use
// that: only the L low-pass taps have to stay in uniform registers (2L do not fit for L >= 16, and
// ptxas then shuffles taps between register files inside the inner loop).  jwc_set_wavelet checks the
// relation bit for bit; filter sets without it run on the one-level kernels, which use both arrays.
template <int L>
__device__ __forceinline__ double hi_tap(const Taps& t, int j) {
  return (j & 1) ? -t.lo[L - 1 - j] : t.lo[L - 1 - j];
}

// Long filters: 2 x L uniform registers do not exist for L >= 32 (63 per warp), and ptxas, which hoists
// the loop-invariant tap loads out of the group loop, then parks taps in 40-80 VECTOR registers (157-184
// registers per thread, 2-3 CTAs per SM).  Indexing the taps with a zero that comes out of an inline
// `mov` keeps the loads inside the step: ptxas folds the zero only after its loop-invariant code motion
// has run, and then streams the taps through the uniform registers with LDCU as the window slides
// (rev<40>: 178 -> 63 registers, rev<30>: 153 -> 40; C4 reverse +13 %, C5 +9..12 %, same-run A/B in
// profiles/r01_ab_tap_streaming.txt).  A zero derived from the loop counter is NOT equivalent: ptxas
// then issues per-thread LDC loads and prefetches them into 110-150 registers (no gain).  The
// forward kernel for L >= 38 is limited to 3 CTAs per SM by its shared memory either way and loses
// 6 % to the extra LDCU, so it keeps the hoisted taps.  The contiguous-line kernels (66-76 registers for
// L = 40) gain nothing from it either (same A/B file).
template <int L, bool REVERSE>
__device__ __forceinline__ int tap_phase() {
  int z = 0;
  if constexpr (L >= 30 && (REVERSE || L <= 36)) asm volatile("mov.u32 %0, 0;" : "=r"(z));
  return z;
}
template <int L>
__device__ __forceinline__ double lo_tap(const Taps& t, int j, int z) { return t.lo[j + z]; }
template <int L>
__device__ __forceinline__ double hi_tap(const Taps& t, int j, int z) {
  return (j & 1) ? -t.lo[L - 1 - j + z] : t.lo[L - 1 - j + z];
}


// Where the rows / lines of an axis pass's FINAL output go when they are stored straight into the slabs
// of peer GPUs (slab-decomposed volume, jwc_axis_dev_remote): the exchange that would follow the pass
// is folded into its stores.
//   mode 1 (strided axis, row = sample index s of outer block o):
//       peer[s >> lg_seg] + o * outer_stride + base_off + (s mod 2^lg_seg) * row_stride + column
//   mode 2 (contiguous axis, line = o * 2^lg_hi + j):
//       peer[j >> lg_seg] + o * outer_stride + base_off + (j mod 2^lg_seg) * row_stride + sample
struct RemoteMap {
  int mode = 0;
  int lg_seg = 0, lg_hi = 0;
  int64_t outer_stride = 0, row_stride = 0, base_off = 0;
  double* peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};
__device__ __forceinline__ double* remote_row(const RemoteMap& r, int64_t o, int64_t s) {
  return r.peer[s >> r.lg_seg] + o * r.outer_stride + r.base_off + (s & ((int64_t(1) << r.lg_seg) - 1)) * r.row_stride;
}
__device__ __forceinline__ double* remote_line(const RemoteMap& r, int64_t line) {
  const int64_t j = line & ((int64_t(1) << r.lg_hi) - 1), o = line >> r.lg_hi;
  return r.peer[j >> r.lg_seg] + o * r.outer_stride + r.base_off + (j & ((int64_t(1) << r.lg_seg) - 1)) * r.row_stride;
}

struct WaveletRec {
  bool mirror_de, mirror_re;  // the relation above holds for the decomposition / reconstruction pair
  int L;
  Taps de;  // decomposition: _scalingDeCom / _waveletDeCom
  Taps re;  // reconstruction: _scalingReCon / _waveletReCon
};

// Optional per-launch timing (jwc_profile_enable): an event pair around every kernel launch on the
// launching stream, summed per kernel label when the report is read.
struct ProfRec {
  const char* name;
  cudaEvent_t beg, end;
  double units;  // samples the launch transforms (its algorithmic work, for roofline arithmetic)
  int levels;    // decomposition levels fused in the launch
};

struct Scratch {
  void* ptr = nullptr;
  size_t bytes = 0;
};

}  // namespace jwc

struct jwc_group;  // the devices of a jwc_create_multi context (jwc_capi.cu)

struct jwc_ctx {
  jwc_group* group = nullptr;  // owner context of a device group: set by jwc_create_multi
  // Every entry point that takes a context holds this lock for its duration: a context is one GPU, one set of
  // scratch buffers and one "current stream", so concurrent callers are serialised here (include/jwave_cuda.h).
  std::recursive_mutex mu;
  // Recorded on ctx->stream after the last launch that touches the scratch buffers; jwc_set_stream makes the
  // new stream wait for it, so work enqueued under different streams never overlaps on the same scratch.
  cudaEvent_t scratch_ev = nullptr;
  bool scratch_ev_valid = false;
  int device = 0;
  int sm_count = 148;
  size_t smem_optin = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;  // where kernels go (own_stream unless jwc_set_stream)
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  std::vector<jwc::WaveletRec> wavelets;
  std::string err;
  int64_t launches = 0;
  const jwc::RemoteMap* remote = nullptr;  // set for the duration of one jwc_axis_dev_remote call
  // Line pitches (doubles) of `in` / `out` when the lines of a contiguous FWT are NOT dense (0 = dense): set by
  // jwc_aed1d for the duration of one block transform, honoured by the fused contiguous FWT plans only (jwc_plan.cu)
  int64_t pitch_in = 0, pitch_out = 0;
  bool prof_on = false;
  std::vector<jwc::ProfRec> prof;
  jwc::Scratch scratch[4];      // [0],[1]: level ping-pong; [2]: axis ping-pong; [3]: alias guard
  jwc::Scratch stage_in[2], stage_out[2];
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  size_t staging_bytes = size_t(64) << 20;  // pipeline ramp = 2 chunks: 64 MiB keeps it below 1 % of an 8 GiB batch
  bool force_generic = false;   // JWC_FORCE_GENERIC=1: only the one-level reference kernels
  // launch-shape tunables (JWC_TUNE="fwd_tile=2048,fwd_m=5,rev_tile=4096,rev_m=5,res_cap=4096")
  int fwd_tile = 2048, fwd_m = 4 /* cap on the halo rule */, fwd_r = 4, rev_tile = 4096, rev_m = 3, rev_rs = 4, fwd_threads = 128, rev_threads = 128, res_cap = 256, res_threads = 128, wpt_tile = 2048, wpt_m = 3, wpt_threads = 160, wpt_rs = 8, wpt_r = 8, wpt_inplace = 1, rev_tail = 1, fwd_tail = 1;
  int str_tile = 512, str_rev_tile = 512, str_rev_m = 5, str_cap = 512, str_threads = 128, str_rev_threads = 128, str_tma = 1;  // strided-axis kernels
  // second-generation strided kernels (inner % 16 == 0): on/off, tile rows, resident cap, forced levels per pass (0 = halo rule)
  int wpt_transpose = 1;  // WPT along strided axes: transpose -> fused contiguous plan -> transpose (0: one-level kernels)
  int wpt_rev_m = 0;  // WPT reverse: levels per tile pass (0 = wpt_m); the reverse's left extension does not grow with depth
  int wpt_tma_store = 1;  // WPT reverse tile kernel: finished tiles leave through cp.async.bulk.tensor stores
  int wpt_tma_store_fwd = 0;  // the same for the forward kernel's 2^m leaf segments (measured neutral: its 64-byte runs per lane were not the limit)
  int pf = 0;         // WPT tile kernels: L2 prefetch (cp.async.bulk.prefetch.L2) of the tile `pf` CTAs ahead; 0 = off
  int res_split = 0;  // resident FWT forward: split the resident work at this width (0 = one launch), jwc_plan.cu
  int res_kb = 48;    // resident kernels: shared-memory budget per CTA (KB) that sets the lines per CTA
  int stagger = 0;    // WPT tile kernels: first-wave stagger in ns per resident-CTA slot (jwc_fused.cuh)
  int xsmem = 0;      // extra dynamic shared memory (KB) per CTA of the WPT tile kernels: an A/B knob that LOWERS the CTAs per SM
  int carve = 0;      // tile kernels: ask for the largest shared-memory carve-out (A/B: does the driver's choice cap the CTAs per SM?)
  int shfl = 1;       // 2-tap filters, forward FWT: the register / warp-shuffle kernel (jwc_shfl.cu) instead of tile passes
  int rot_warps = 0;  // tile kernels with a tail warp: rotate the warps' roles with the CTA number (rotated_tid; measured slower)
  int str_v2 = 1, str2_tile = 512, str2_rev_tile = 256 /* A/B: profiles/r02_ab_strided_v2_tuning.txt */, str2_cap = 512, str2_m = 0, str2_rev_m = 0;
};

namespace jwc {

// bracket a kernel launch with profile events (no-ops unless profiling is on)
inline void prof_begin(jwc_ctx* ctx, const char* name, double units, int levels) {
  if (!ctx->prof_on) return;
  ProfRec r{name, nullptr, nullptr, units, levels};
  cudaEventCreate(&r.beg);
  cudaEventCreate(&r.end);
  cudaEventRecord(r.beg, ctx->stream);
  ctx->prof.push_back(r);
}
inline void prof_end(jwc_ctx* ctx) {
  if (!ctx->prof_on || ctx->prof.empty()) return;
  cudaEventRecord(ctx->prof.back().end, ctx->stream);
}

// ---- geometry of one level over a set of lines ---------------------------------------------
// A "line" is the 1-D sequence the transform runs along.  Element s of line (o, c) lives at
//   base + o * os + s * inner + c,      o in [0, outer), c in [0, inner).
// Contiguous signals have inner == 1; columns of a matrix have inner == cols.

struct FwdLevelArgs {
  const double* src; int64_t src_os;
  double* dstA; int64_t dstA_os;
  double* dstD; int64_t dstD_os;
  int64_t outer; int64_t inner;
  int half;  // outputs per line and filter: h / 2
};

struct RevLevelArgs {
  const double* srcA; int64_t srcA_os;
  const double* srcD; int64_t srcD_os;
  double* dst; int64_t dst_os;
  int64_t outer; int64_t inner;
  int half;
};

// jwc_generic.cu: one level, any geometry, any width (exact modular wrap); the slow-but-always
// -right kernels every other path is checked against and falls back to ON THE GPU.
cudaError_t launch_fwd_level_generic(jwc_ctx* ctx, int L, const Taps& taps, const FwdLevelArgs& a);
cudaError_t launch_rev_level_generic(jwc_ctx* ctx, int L, const Taps& taps, const RevLevelArgs& a);

// jwc_compress.cu: CompressorMagnitude; `scratch` holds blocks + 2 doubles (partials, magnitude, CTA counter = 0)
cudaError_t launch_compress_magnitude(jwc_ctx* ctx, const double* in, double* out, int64_t n, double threshold,
                                      double* scratch, int blocks);
// the two halves: |x| partial sums (optionally accumulated over several calls; `finish` = also write the mean over
// n_total) and the threshold pass; scratch holds blocks + 2 doubles, zero-initialised
cudaError_t launch_abs_sum(jwc_ctx* ctx, const double* in, int64_t n, int64_t n_total, double* scratch, int blocks,
                           bool accumulate, bool finish);
cudaError_t launch_threshold(jwc_ctx* ctx, const double* in, double* out, int64_t n, double threshold, double* scratch,
                             int blocks);

}  // namespace jwc
