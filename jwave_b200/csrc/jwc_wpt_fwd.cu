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
#include <cuda.h>

#include <cstring>

#include "jwc_fused.cuh"
#include "jwc_kernels.cuh"

#ifndef JWC_WPT_TAIL_WARP
#define JWC_WPT_TAIL_WARP 1
#endif

namespace jwc {

// Launch bound of the tile kernels.  -DJWC_WPT_MINB=n builds them for 160-thread CTAs with n CTAs per SM (an A/B
// switch for the register budget; the default keeps run-time CTA sizes up to 512).
#ifdef JWC_WPT_MINB
#define JWC_WPT_TILE_BOUNDS __launch_bounds__(160, JWC_WPT_MINB)
#else
#define JWC_WPT_TILE_BOUNDS __launch_bounds__(512)
#endif

// Layout of this kernel's shared-memory lines: one pad slot per R double2, so a thread that
// produces R outputs (window = R + L/2 - 1 consecutive double2 from a multiple of R) is R + 1 slots
// away from its neighbour - an odd stride, conflict-free LDS.128 for R = 4 and R = 8.  A longer run
// per thread cuts shared-memory wavefronts per DFMA (L = 16: 11 LDS per 128 DFMA at R = 4, 15 per
// 256 at R = 8), which is what bounds the long-filter WPT.
template <int R> __device__ __forceinline__ int padr(int k2) { return k2 + (k2 / R); }
template <int R> __host__ __device__ constexpr int padr_size(int n2) { return n2 + n2 / R + 2; }
template <int R> __device__ __forceinline__ double lscalar(const double2* buf, int i) {
  return reinterpret_cast<const double*>(buf)[2 * padr<R>(i >> 1) + (i & 1)];
}
template <int R> __device__ __forceinline__ void lscalar_store(double2* buf, int i, double v) {
  reinterpret_cast<double*>(buf)[2 * padr<R>(i >> 1) + (i & 1)] = v;
}

// ---- tile mode ------------------------------------------------------------------------------------
// At every level the T/2 outputs (per filter) the tile keeps are exactly T / (2R) groups of R, a power
// of two, so the warps are always full and the (node, group) split of an item is a shift and a mask.
// The halo outputs a node owes the levels below ((2^(m-k) - 1)(L - 2) per node, none at the last
// level) are a separate short step, two outputs per lane, instead of one more nearly empty R-wide step
// for every warp; JWC_WPT_TAIL_WARP selects who runs it: a dedicated extra warp (1) or one of the
// main warps, rotating with the CTA and the level so that no SM sub-partition collects all of it (0).
// TMA store of one finished leaf-packet segment (cp.async.bulk.tensor, SASS UTMASTG): box {16 doubles, (T >> m) / 16
// rows} of the output seen as a [rows][16] matrix, from a dense, 128-byte-swizzled shared-memory image (jwc_wpt_rev.cu).
__device__ __forceinline__ void tma_store_box_w(const void* tmap, const void* smem_src, int x, int y) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_src));
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(x), "r"(y), "r"(s)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <int L, int R, bool INPLACE>
__global__ void JWC_WPT_TILE_BOUNDS
k_wpt_fwd_tile(const __grid_constant__ Taps taps, const __grid_constant__ WptFwdArgs a, const __grid_constant__ CUtensorMap tmapOut) {
  extern __shared__ __align__(1024) double2 smem2[];
  constexpr int lgR = (R == 8) ? 3 : 2;
  static_assert(R == 8 || R == 4, "R is 4 or 8");
  const int tid = rotated_tid(a.rot), nthr = blockDim.x, nmain = nthr - 32 * JWC_WPT_TAIL_WARP;
  const int m = a.m, h = a.h, T = a.T;
  const int64_t line = blockIdx.x >> a.lg_tpl;
  const int tile = int(blockIdx.x) & (a.tiles_per_line - 1);
  const int base = tile * T;
  double2* cur = smem2;
  double2* nxt = INPLACE ? smem2 : smem2 + a.buf_cap;
  first_wave_stagger(a.stagger_ns, a.stagger_ctas, a.stagger_div);
  {
    // stage the tile and its right halo; only the last tile of a line wraps (once: halo <= T / 4 < h).
    // nthr is a multiple of R, so a step of nthr double2 is a constant step through the padded layout.
    const int n2 = (T + ((1 << m) - 1) * (L - 2)) >> 1;
    const int fit2 = min(n2, (h - base) >> 1);
    const int step = nthr + nthr / R;
    const double* sp = a.src + line * a.src_os + base + 2 * tid;
    double2* dp = cur + padr<R>(tid);
    int k2 = tid;
    for (; k2 < fit2; k2 += nthr, sp += 2 * nthr, dp += step) cp_async16(dp, sp);
    sp -= h;
#pragma unroll 1
    for (; k2 < n2; k2 += nthr, sp += 2 * nthr, dp += step) cp_async16(dp, sp);
    // L2 prefetch of the tile a CTA `pf_dist` launches later will stage (one instruction, no registers, no wait)
    if (a.pf_dist > 0 && tid == 0) {
      const int64_t t2 = int64_t(blockIdx.x) + a.pf_dist;
      if (t2 < (a.lines << a.lg_tpl)) {
        const double* p2 = a.src + (t2 >> a.lg_tpl) * a.src_os + (int(t2) & (a.tiles_per_line - 1)) * T;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p2), "r"(T * 8) : "memory");
      }
    }
    cp_async_wait_all();
    __syncthreads();
  }

  int cap_in = 0;  // per-node capacity of the level being read (one node at level 0)
  if constexpr (INPLACE) {
    // One item per thread and level (the launcher checks it): the results wait in registers until every
    // window of the level has been read, then overwrite the level's input.  One buffer instead of two
    // halves the shared memory of a CTA - more CTAs, i.e. more warps in their FMA phase, per SM - for
    // one more barrier per level.
    for (int k = 1; k <= m; ++k) {
      const int cap_out = a.cap[k];
      const bool last = (k == m);
      double lo[R], hi[R];
      int node = 0, g = 0, o = 0;
      bool has = false;
      if (tid < nmain) {
        const int lg_gpn = a.lg_T - k - lgR;
        node = tid >> lg_gpn;
        g = tid & ((1 << lg_gpn) - 1);
        const double2* w = cur + node * cap_in + (R + 1) * g;
        fwd_stepR<L, R>(taps, [&](int q) { return w[q + q / R]; }, lo, hi);
      } else if (!last) {
        const int per_node = (((1 << (m - k)) - 1) * (L - 2)) >> 1;
        int j = tid - nmain;
        has = j < (per_node << (k - 1));
        if (has) {
          while (j >= per_node) { j -= per_node; ++node; }
          const double2* w = cur + node * cap_in;
          o = (T >> k) + 2 * j;
          double l2[2], h2[2];
          fwd_stepR<L, 2>(taps, [&](int q) { return w[padr<R>(o + q)]; }, l2, h2);
          lo[0] = l2[0]; lo[1] = l2[1]; hi[0] = h2[0]; hi[1] = h2[1];
        }
      }
      if (last) {
        if (a.tma_out && R == 8) {
          // The 2^m leaf segments of the tile leave through TMA stores: the buffer is free once every window of the
          // last level has been read; the segment of leaf j is a dense [(T >> m) / 16][16] image at 16 (T >> m) bytes x j,
          // thread (node, g) owns half a row of leaves 2 node and 2 node + 1, chunks XOR-swizzled by the row number
          // (conflict-free); lanes 0 .. 2^m - 1 of warp 0 each hand one segment to the copy engine.
          __syncthreads();
          const int seg2 = (T >> m) >> 1;  // double2 per leaf segment
          if (tid < nmain) {
            const int row = g >> 1, x = row & 7, c0 = 4 * (g & 1);
            double2* ia = smem2 + (2 * node) * seg2 + 8 * row;
            double2* id = ia + seg2;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              ia[(c0 + e) ^ x] = make_double2(lo[2 * e], lo[2 * e + 1]);
              id[(c0 + e) ^ x] = make_double2(hi[2 * e], hi[2 * e + 1]);
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncthreads();
          if (tid < (1 << m)) {
            const int64_t leaf = h >> m;
            tma_store_box_w(&tmapOut, smem2 + tid * seg2, 0, int((line * a.dst_os + tid * leaf + (base >> m)) >> 4));
          }
          break;
        }
        if (tid < nmain) {
          const int leaf = h >> m;
          double* pa = a.dst + line * a.dst_os + int64_t(2 * node) * leaf + (base >> m) + R * g;
#pragma unroll
          for (int e = 0; e < R / 4; ++e) {
            st_global_v4(pa + 4 * e, lo[4 * e], lo[4 * e + 1], lo[4 * e + 2], lo[4 * e + 3]);
            st_global_v4(pa + leaf + 4 * e, hi[4 * e], hi[4 * e + 1], hi[4 * e + 2], hi[4 * e + 3]);
          }
        }
        break;
      }
      __syncthreads();
      if (tid < nmain) {
        double2* na = cur + (2 * node) * cap_out + (R / 2) * g + (g >> 1);
        double2* nd = na + cap_out;
#pragma unroll
        for (int e = 0; e < R / 2; ++e) {
          na[e] = make_double2(lo[2 * e], lo[2 * e + 1]);
          nd[e] = make_double2(hi[2 * e], hi[2 * e + 1]);
        }
      } else if (has) {
        double2* na = cur + (2 * node) * cap_out + padr<R>(o >> 1);
        na[0] = make_double2(lo[0], lo[1]);
        na[cap_out] = make_double2(hi[0], hi[1]);
      }
      __syncthreads();
      cap_in = cap_out;
    }
  } else {
  for (int k = 1; k <= m; ++k) {
    const int cap_out = a.cap[k];
    const bool last = (k == m);
    if (tid < nmain) {
      const int lg_gpn = a.lg_T - k - lgR;  // groups per node = (T >> k) / R
      for (int it = tid; it < ((T >> 1) >> lgR); it += nmain) {
        const int node = it >> lg_gpn, g = it & ((1 << lg_gpn) - 1);
        const double2* w = cur + node * cap_in + (R + 1) * g;  // padr(R g + q) == (R + 1) g + q + q / R
        double lo[R], hi[R];
        fwd_stepR<L, R>(taps, [&](int q) { return w[q + q / R]; }, lo, hi);
        if (!last) {
          // padr(R/2 g + e) == R/2 g + (g >> 1) + e for e < R/2
          double2* na = nxt + (2 * node) * cap_out + (R / 2) * g + (g >> 1);
          double2* nd = na + cap_out;
#pragma unroll
          for (int e = 0; e < R / 2; ++e) {
            na[e] = make_double2(lo[2 * e], lo[2 * e + 1]);
            nd[e] = make_double2(hi[2 * e], hi[2 * e + 1]);
          }
        } else {
          const int leaf = h >> m;
          double* pa = a.dst + line * a.dst_os + int64_t(2 * node) * leaf + (base >> m) + R * g;
#pragma unroll
          for (int e = 0; e < R / 4; ++e) {
            st_global_v4(pa + 4 * e, lo[4 * e], lo[4 * e + 1], lo[4 * e + 2], lo[4 * e + 3]);
            st_global_v4(pa + leaf + 4 * e, hi[4 * e], hi[4 * e + 1], hi[4 * e + 2], hi[4 * e + 3]);
          }
        }
      }
    }
    const int tail_warp = JWC_WPT_TAIL_WARP ? (nmain >> 5) : int((blockIdx.x + k) % unsigned(nthr >> 5));
    if (!last && (tid >> 5) == tail_warp) {
      const int n_keep = T >> k;
      const int per_node = (((1 << (m - k)) - 1) * (L - 2)) >> 1;  // halo steps per node (L - 2 is even)
      const int items = per_node << (k - 1);
      int node = 0;
      for (int it = tid & 31, j = it; it < items; it += 32, j += 32) {
        while (j >= per_node) { j -= per_node; ++node; }
        const double2* w = cur + node * cap_in;
        const int o = n_keep + 2 * j;  // first of the two outputs == first double2 of their window
        double lo[2], hi[2];
        fwd_stepR<L, 2>(taps, [&](int q) { return w[padr<R>(o + q)]; }, lo, hi);
        double2* na = nxt + (2 * node) * cap_out + padr<R>(o >> 1);
        na[0] = make_double2(lo[0], lo[1]);
        na[cap_out] = make_double2(hi[0], hi[1]);
      }
    }
    __syncthreads();
    double2* t = cur; cur = nxt; nxt = t;
    cap_in = cap_out;
  }
  }
}

// ---- resident mode --------------------------------------------------------------------------------
template <int L, int R>
__global__ void __launch_bounds__(512)
k_wpt_fwd_res(const __grid_constant__ Taps taps, const WptFwdArgs a) {
  extern __shared__ double2 smem2[];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int m = a.m, h = a.h;
  {

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
        cp_async16(&cur[ln * cap + padr<R>(k2)], a.src + (line0 + ln) * a.src_os + 2 * k2);
      }
      cp_async_wait_all();
      __syncthreads();
    }
    for (int k = 1; k <= m; ++k) {
      const int h_in = h >> (k - 1), h_out = h_in >> 1;   // node length before / after this level
      const int lg_nodes = k - 1;                          // nodes per line at the input level = 2^(k-1)
      const bool last = (k == m);
      if (h_out >= R) {
        const int gpn = h_out / R;                        // groups per node (power of two)
        const int per_line = gpn << lg_nodes;              // == h / 8
        const int mask2 = (h_in >> 1) - 1;
        for (int it = tid; it < nlines * per_line; it += nthr) {
          const int ln = it / per_line, r = it - ln * per_line;
          const int node = r / gpn, g = r - node * gpn;
          const double2* cl = cur + ln * cap;
          const int off2 = node * (h_in >> 1);             // node start inside the line (double2)
          double lo[R], hi[R];
          fwd_stepR<L, R>(taps, [&](int q) { return cl[padr<R>(off2 + ((R * g + q) & mask2))]; }, lo, hi);
          if (!last) {
            double2* nl = nxt + ln * cap;
            const int oa = (2 * node) * (h_out >> 1) + R / 2 * g, od = oa + (h_out >> 1);
#pragma unroll
            for (int e = 0; e < R / 2; ++e) {
              nl[padr<R>(oa + e)] = make_double2(lo[2 * e], lo[2 * e + 1]);
              nl[padr<R>(od + e)] = make_double2(hi[2 * e], hi[2 * e + 1]);
            }
          } else {
            double* pa = a.dst + (line0 + ln) * a.dst_os + int64_t(2 * node) * h_out + R * g;
#pragma unroll
            for (int e = 0; e < R / 4; ++e) {
              st_global_v4(pa + 4 * e, lo[4 * e], lo[4 * e + 1], lo[4 * e + 2], lo[4 * e + 3]);
              st_global_v4(pa + h_out + 4 * e, hi[4 * e], hi[4 * e + 1], hi[4 * e + 2], hi[4 * e + 3]);
            }
          }
        }
      } else {
        // nodes shorter than R outputs: one thread per (line, node, i); true modular wrap
        const int per_line = h_out << lg_nodes;            // == h / 2
        for (int it = tid; it < nlines * per_line; it += nthr) {
          const int ln = it / per_line, r = it - ln * per_line;
          const int node = r / h_out, i = r - node * h_out;
          const double2* cl = cur + ln * cap;
          const int off = node * h_in;
          double lo = 0.0, hi = 0.0;
#pragma unroll
          for (int j = 0; j < L; ++j) {
            const double x = lscalar<R>(cl, off + ((2 * i + j) & (h_in - 1)));
            lo = fma(x, taps.lo[j], lo);
            hi = fma(x, hi_tap<L>(taps, j), hi);
          }
          const int oa = (2 * node) * h_out + i, od = oa + h_out;
          if (!last) {
            double2* nl = nxt + ln * cap;
            lscalar_store<R>(nl, oa, lo);
            lscalar_store<R>(nl, od, hi);
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
static size_t wpt_fwd_tile_smem(int L, int T, int m, int R, int* buf_cap, int* caps = nullptr) {
  auto psize = [R](int n2) { return n2 + n2 / R + 2; };
  int cap = psize((T + ((1 << m) - 1) * (L - 2)) / 2 + R);  // level 0: one node
  for (int k = 1; k <= m; ++k) {                                  // levels kept in shared memory
    const int n_out = (T >> k) + ((1 << (m - k)) - 1) * (L - 2);
    const int per_node = psize(n_out / 2 + R);
    if (caps) caps[k] = per_node;
    if (k < m && (1 << k) * per_node > cap) cap = (1 << k) * per_node;
  }
  *buf_cap = cap;
  return size_t(2) * cap * sizeof(double2);
}

int wpt_tile_levels(int L, int T, int want, size_t smem_limit, int R) {
  // as many levels as asked for, while the halo stays below T / 4, the leaf runs keep >= 4 samples
  // (32-byte stores) and two level buffers fit in shared memory
  int m = 1, cap;
  while (m < want && ((1 << (m + 1)) - 1) * (L - 2) <= T / 4 && (T >> (m + 1)) >= R &&
         wpt_fwd_tile_smem(L, T, m + 1, R, &cap) <= smem_limit)
    ++m;
  return m;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFnF)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnF encode_tiled_f() {
  static EncodeTiledFnF fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFnF>(p);
  }();
  return fn;
}
// the output lines, dense, as a [rows][16 doubles] matrix; box {16, box_rows}; 128-byte swizzle on the shared side
static bool make_out_tmap_f(CUtensorMap* map, const double* base, int64_t rows, int box_rows) {
  EncodeTiledFnF enc = encode_tiled_f();
  if (!enc || (reinterpret_cast<uintptr_t>(base) & 127) || rows < 1 || rows >= (int64_t(1) << 31) || box_rows < 1 || box_rows > 256)
    return false;
  const cuuint64_t dims[2] = {16, cuuint64_t(rows)};
  const cuuint64_t strides[1] = {16 * sizeof(double)};
  const cuuint32_t box[2] = {16, cuuint32_t(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int L, int R>
static cudaError_t launch_LR(jwc_ctx* ctx, const Taps& taps, WptFwdArgs a, bool resident) {
  size_t smem;
  int64_t grid;
  auto ilog2 = [](int v) { int l = 0; while ((1 << l) < v) ++l; return l; };
  const int nthr = ctx->wpt_threads;
  bool inplace = false;
  if (!resident) {
    // T, h powers of two; at least one main warp beside the tail warp; steps of nthr keep the pad phase
    if ((a.T >> a.m) < R || a.m > kMaxFuse || (a.T & (a.T - 1)) || nthr < 32 + 32 * JWC_WPT_TAIL_WARP || nthr % 32)
      return cudaErrorInvalidValue;
    smem = wpt_fwd_tile_smem(L, a.T, a.m, R, &a.buf_cap, a.cap);
    inplace = JWC_WPT_TAIL_WARP && ctx->wpt_inplace && nthr - 32 == (a.T / 2) / R;
    for (int k = 1; k < a.m; ++k)  // tail steps of a level: one per lane of the tail warp
      if (((((1 << (a.m - k)) - 1) * (L - 2)) >> 1) << (k - 1) > 32) inplace = false;
    if (inplace) smem /= 2;
    a.tiles_per_line = a.h / a.T;
    a.rot = (JWC_WPT_TAIL_WARP && ctx->rot_warps) ? 1 : 0;
    a.stagger_ns = ctx->stagger;
    a.pf_dist = ctx->pf;
    a.stagger_div = ctx->sm_count;
    a.stagger_ctas = ctx->sm_count * 8;
    a.lg_tpl = ilog2(a.tiles_per_line);
    a.lg_T = ilog2(a.T);
    grid = a.lines * a.tiles_per_line;
  } else {
    a.buf_cap = padr_size<R>(max(1, a.h / 2));
    smem = size_t(2) * a.G * a.buf_cap * sizeof(double2);
    grid = (a.lines + a.G - 1) / a.G;
  }
  if (grid > 0x7fffffff) return cudaErrorInvalidConfiguration;
  if (!resident) smem += size_t(ctx->xsmem) << 10;
  {
    const void* kern = resident ? reinterpret_cast<const void*>(k_wpt_fwd_res<L, R>)
                                : inplace ? reinterpret_cast<const void*>(k_wpt_fwd_tile<L, R, true>)
                                          : reinterpret_cast<const void*>(k_wpt_fwd_tile<L, R, false>);
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
      if (e != cudaSuccess) return e;
    }
    if (ctx->carve) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  if (resident) {
    prof_begin(ctx, "k_wpt_fwd:resident", double(a.lines) * a.h, a.m);
    k_wpt_fwd_res<L, R><<<int(grid), nthr, smem, ctx->stream>>>(taps, a);
    prof_end(ctx);
    ctx->launches++;
    return cudaGetLastError();
  }
  CUtensorMap tmapOut;
  memset(&tmapOut, 0, sizeof tmapOut);
  a.tma_out = 0;
  const int seg = a.T >> a.m;  // samples per leaf segment of a tile
  if (inplace && R == 8 && ctx->wpt_tma_store_fwd && a.dst_os == a.h && seg % 128 == 0 && seg / 16 <= 256 && (a.h >> a.m) % 16 == 0 &&
      (1 << a.m) <= 32 && smem >= size_t(a.T) * sizeof(double) + (size_t(ctx->xsmem) << 10) &&
      make_out_tmap_f(&tmapOut, a.dst, a.lines * (a.h / 16), seg / 16))
    a.tma_out = 1;
  prof_begin(ctx, "k_wpt_fwd:tile", double(a.lines) * a.h, a.m);
  if (inplace) k_wpt_fwd_tile<L, R, true><<<int(grid), nthr, smem, ctx->stream>>>(taps, a, tmapOut);
  else k_wpt_fwd_tile<L, R, false><<<int(grid), nthr, smem, ctx->stream>>>(taps, a, tmapOut);
  prof_end(ctx);
  ctx->launches++;
  return cudaGetLastError();
}

template <int L>
static cudaError_t launch_L(jwc_ctx* ctx, const Taps& taps, const WptFwdArgs& a, bool resident) {
  return ctx->wpt_r == 8 ? launch_LR<L, 8>(ctx, taps, a, resident) : launch_LR<L, 4>(ctx, taps, a, resident);
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
