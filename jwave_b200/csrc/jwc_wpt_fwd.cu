// jwc_wpt_fwd.cu - fused multi-level forward wavelet PACKET transform along contiguous lines.
//
// Replaces the level x packet loops of WaveletPacketTransform.forward
// (WaveletPacketTransform.java:96-120; same arithmetic as the Pooled / Parallel variants,
// PooledWaveletPacketTransform.java:24-71, ParallelWaveletPacketTransform.java:79-110) around
// Wavelet.forward (Wavelet.java:236-260) for `m` consecutive levels per launch.  Every tree node
// is expanded in shared memory: node j of level k-1 becomes nodes 2j (approximation) and 2j+1
// (detail) of level k - the reference's natural (Paley) packet order - and only the 2^m leaf
// packets return to HBM.
//
// A "line" here is one packet of width h of the current level (the planner folds signals x packets
// into lines), so the same kernel serves the first pass over whole signals and later passes over
// packets.  The wrap is per packet, as in the reference (WaveletPacketTransform.java:104-109).
//
//   tile mode      (h > res_cap) : one CTA = T samples of one line + periodic right halo
//                                  (2^m - 1)(L - 2); node j of the last level goes to
//                                  line + j * (h >> m) + tile * (T >> m).
//   resident mode  (h <= res_cap): one CTA = G whole lines, wrap by index mask; nodes shorter than
//                                  4 samples take a scalar path with true modular indexing.
#include "jwc_fused.cuh"
#include "jwc_kernels.cuh"

namespace jwc {

template <int L, bool RESIDENT>
__global__ void __launch_bounds__(512)
k_wpt_fwd(const __grid_constant__ Taps taps, const WptFwdArgs a) {
  extern __shared__ double2 smem2[];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int m = a.m, h = a.h;

  if constexpr (!RESIDENT) {
    // ---------------- tile mode ----------------
    const int64_t line = blockIdx.x / a.tiles_per_line;
    const int tile = int(blockIdx.x % a.tiles_per_line);
    const int T = a.T;
    const int n0 = T + ((1 << m) - 1) * (L - 2);
    double2* cur = smem2;
    double2* nxt = smem2 + a.buf_cap;
    const double* src = a.src + line * a.src_os;
    const int base = tile * T;
    for (int k2 = tid; k2 < n0 / 2; k2 += nthr)
      cp_async16(&cur[pad2(k2)], src + ((base + 2 * k2) & (h - 1)));
    cp_async_wait_all();
    __syncthreads();

    double* outl = a.dst + line * a.dst_os;
    int cap_in = 0;  // per-node capacity of the level being read (one node at level 0)
    for (int k = 1; k <= m; ++k) {
      const int n_keep = T >> k;                                    // outputs of each node this tile owns
      const int n_out = n_keep + ((1 << (m - k)) - 1) * (L - 2);    // incl. halo for the levels below
      const int groups = (n_out + kR - 1) / kR;
      const int cap_out = pad2_size(n_out / 2 + 4);
      const int items = groups << (k - 1);                          // nodes_in * groups
      const bool last = (k == m);
      for (int it = tid; it < items; it += nthr) {
        const int node = it / groups, g = it - node * groups;
        const double2* w = cur + node * cap_in + 5 * g;             // pad2(4g + q) == 5g + q + (q >> 2)
        double lo[kR], hi[kR];
        fwd_step4<L>(taps, [&](int q) { return w[q + (q >> 2)]; }, lo, hi);
        if (!last) {
          double2* na = nxt + (2 * node) * cap_out;
          double2* nd = na + cap_out;
          na[pad2(2 * g)] = make_double2(lo[0], lo[1]);
          na[pad2(2 * g + 1)] = make_double2(lo[2], lo[3]);
          nd[pad2(2 * g)] = make_double2(hi[0], hi[1]);
          nd[pad2(2 * g + 1)] = make_double2(hi[2], hi[3]);
        } else if (kR * g < n_keep) {
          double* pa = outl + int64_t(2 * node) * (h >> m) + tile * n_keep + kR * g;
          st_global_v4(pa, lo[0], lo[1], lo[2], lo[3]);
          st_global_v4(pa + (h >> m), hi[0], hi[1], hi[2], hi[3]);
        }
      }
      __syncthreads();
      double2* t = cur; cur = nxt; nxt = t;
      cap_in = cap_out;
    }
  } else {
    // ---------------- resident mode ----------------
    const int G = a.G;
    const int64_t line0 = int64_t(blockIdx.x) * G;
    const int nlines = int(min(int64_t(G), a.lines - line0));
    const int cap = a.buf_cap;               // per-line capacity (double2) of each buffer
    double2* cur = smem2;
    double2* nxt = smem2 + size_t(G) * cap;
    {
      const int per_line = h >> 1;
      for (int it = tid; it < nlines * per_line; it += nthr) {
        const int ln = it / per_line, k2 = it - ln * per_line;
        cp_async16(&cur[ln * cap + pad2(k2)], a.src + (line0 + ln) * a.src_os + 2 * k2);
      }
      cp_async_wait_all();
      __syncthreads();
    }
    for (int k = 1; k <= m; ++k) {
      const int h_in = h >> (k - 1), h_out = h_in >> 1;   // node length before / after this level
      const int lg_nodes = k - 1;                          // nodes per line at the input level = 2^(k-1)
      const bool last = (k == m);
      if (h_out >= kR) {
        const int gpn = h_out / kR;                        // groups per node (power of two)
        const int per_line = gpn << lg_nodes;              // == h / 8
        const int mask2 = (h_in >> 1) - 1;
        for (int it = tid; it < nlines * per_line; it += nthr) {
          const int ln = it / per_line, r = it - ln * per_line;
          const int node = r / gpn, g = r - node * gpn;
          const double2* cl = cur + ln * cap;
          const int off2 = node * (h_in >> 1);             // node start inside the line (double2)
          double lo[kR], hi[kR];
          fwd_step4<L>(taps, [&](int q) { return cl[pad2(off2 + ((kR * g + q) & mask2))]; }, lo, hi);
          if (!last) {
            double2* nl = nxt + ln * cap;
            const int oa = (2 * node) * (h_out >> 1) + 2 * g, od = oa + (h_out >> 1);
            nl[pad2(oa)] = make_double2(lo[0], lo[1]);
            nl[pad2(oa + 1)] = make_double2(lo[2], lo[3]);
            nl[pad2(od)] = make_double2(hi[0], hi[1]);
            nl[pad2(od + 1)] = make_double2(hi[2], hi[3]);
          } else {
            double* pa = a.dst + (line0 + ln) * a.dst_os + int64_t(2 * node) * h_out + kR * g;
            st_global_v4(pa, lo[0], lo[1], lo[2], lo[3]);
            st_global_v4(pa + h_out, hi[0], hi[1], hi[2], hi[3]);
          }
        }
      } else {
        // nodes of 2 or 4 samples in, 1 or 2 out: one thread per (line, node, i); true modular wrap
        const int per_line = h_out << lg_nodes;            // == h / 2
        for (int it = tid; it < nlines * per_line; it += nthr) {
          const int ln = it / per_line, r = it - ln * per_line;
          const int node = r / h_out, i = r - node * h_out;
          const double2* cl = cur + ln * cap;
          const int off = node * h_in;
          double lo = 0.0, hi = 0.0;
#pragma unroll
          for (int j = 0; j < L; ++j) {
            const double x = sm_scalar(cl, off + ((2 * i + j) & (h_in - 1)));
            lo = fma(x, taps.lo[j], lo);
            hi = fma(x, taps.hi[j], hi);
          }
          const int oa = (2 * node) * h_out + i, od = oa + h_out;
          if (!last) {
            double2* nl = nxt + ln * cap;
            sm_scalar_store(nl, oa, lo);
            sm_scalar_store(nl, od, hi);
          } else {
            double* pl = a.dst + (line0 + ln) * a.dst_os;
            pl[oa] = lo;
            pl[od] = hi;
          }
        }
      }
      __syncthreads();
      double2* t = cur; cur = nxt; nxt = t;
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------

// Shared memory (bytes) of a tile-mode launch; also fills the per-buffer capacity.
static size_t wpt_fwd_tile_smem(int L, int T, int m, int* buf_cap) {
  int cap = pad2_size((T + ((1 << m) - 1) * (L - 2)) / 2 + 4);  // level 0: one node
  for (int k = 1; k < m; ++k) {                                   // levels kept in shared memory
    const int n_out = (T >> k) + ((1 << (m - k)) - 1) * (L - 2);
    const int c = (1 << k) * pad2_size(n_out / 2 + 4);
    if (c > cap) cap = c;
  }
  *buf_cap = cap;
  return size_t(2) * cap * sizeof(double2);
}

int wpt_tile_levels(int L, int T, int want, size_t smem_limit) {
  // as many levels as asked for, while the halo stays below T / 4, the leaf runs keep >= 4 samples
  // (32-byte stores) and two level buffers fit in shared memory
  int m = 1, cap;
  while (m < want && ((1 << (m + 1)) - 1) * (L - 2) <= T / 4 && (T >> (m + 1)) >= kR &&
         wpt_fwd_tile_smem(L, T, m + 1, &cap) <= smem_limit)
    ++m;
  return m;
}

template <int L>
static cudaError_t launch_L(jwc_ctx* ctx, const Taps& taps, WptFwdArgs a, bool resident) {
  size_t smem;
  int64_t grid;
  if (!resident) {
    if ((a.T >> a.m) < kR) return cudaErrorInvalidValue;
    smem = wpt_fwd_tile_smem(L, a.T, a.m, &a.buf_cap);
    a.tiles_per_line = a.h / a.T;
    grid = a.lines * a.tiles_per_line;
  } else {
    a.buf_cap = pad2_size(max(1, a.h / 2));
    smem = size_t(2) * a.G * a.buf_cap * sizeof(double2);
    grid = (a.lines + a.G - 1) / a.G;
  }
  if (grid > 0x7fffffff) return cudaErrorInvalidConfiguration;
  auto kern = resident ? k_wpt_fwd<L, true> : k_wpt_fwd<L, false>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
  }
  kern<<<int(grid), ctx->wpt_threads, smem, ctx->stream>>>(taps, a);
  ctx->launches++;
  return cudaGetLastError();
}

cudaError_t launch_wpt_fwd(jwc_ctx* ctx, int L, const Taps& taps, const WptFwdArgs& a, bool resident) {
  switch (L) {
#define JWC_CASE(LL) case LL: return launch_L<LL>(ctx, taps, a, resident);
    JWC_FOR_EACH_L(JWC_CASE)
#undef JWC_CASE
  }
  return cudaErrorInvalidValue;
}

}  // namespace jwc
