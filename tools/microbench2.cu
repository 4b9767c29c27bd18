// microbench2.cu - where is the ceiling of a DFMA stream that is shaped like the fused wavelet steps?
// The steps of the FP64-bound kernels are `acc[r] = fma(window value, tap (uniform register), acc[r])` with one
// LDS.128 per 16-32 DFMA; no kernel of the library gets the FP64 pipe above ~82 % busy, with 7 or with 10 warps per
// scheduler.  This program isolates the ingredients:
//   taps    : DFMA R, R, UR, R with 16 accumulators and a sliding window that lives in registers (no loads)
//   taps+lds: the same with the window read from shared memory, one conflict-free LDS.128 per `per` DFMA
//   regs    : the register-operand loop of tools/microbench.cu (the 36.7 TFLOP/s denominator) for reference
// each at 1, 2, 4, 7 and 10 warps per scheduler.  Output: TFLOP/s and the fraction of 148 x 4 x 16 lanes x 2 x clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench2 tools/microbench2.cu
#include <cstdio>
#include <cuda_runtime.h>

struct Taps { double lo[40]; };

__global__ void __launch_bounds__(128) k_regs(double* out, double a, double b, int iters) {
  double r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = fma(r[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// One "step" = 256 DFMA: 16 accumulators x 16 taps, window values v[q] (q = 0..15 + 7), as fwd_stepR<16, 8>.
template <int MODE, int XLD = 0, int XST = 0>  // 0: window in registers, 1: window from shared memory (LDS.128 per 32 DFMA), 2: LDS.128 per 16 DFMA
                                            // XLD / XST: extra LDS.128 / STS.128 per step (the in-place level stores, the staging)
__global__ void __launch_bounds__(128) k_step(double* out, const __grid_constant__ Taps taps, int iters) {
  extern __shared__ double2 sm[];
  for (int i = threadIdx.x; i < 2048 + 256; i += blockDim.x) sm[i] = make_double2(i * 1e-3, 1.0 - i * 1e-3);
  __syncthreads();
  double lo[8], hi[8];
  int sink = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) lo[i] = hi[i] = 0.0;
  double2 w0 = make_double2(threadIdx.x * 1e-9, 1.0), w1 = make_double2(0.5, threadIdx.x * 1e-7);
  const double2* base = sm + 9 * (threadIdx.x & 31) + 300 * (threadIdx.x >> 5);  // lanes 9 slots apart: conflict-free LDS.128
  for (int it = 0; it < iters; ++it) {
    const double2* w = base + (it & 7);
    if constexpr (MODE == 0) { w0 = w[0]; w1 = w[1]; }  // two loads per 256 DFMA: keeps the products out of loop-invariant code motion
#pragma unroll
    for (int q = 0; q < 15; ++q) {
      double2 v;
      if constexpr (MODE == 0) v = (q & 1) ? w1 : w0;
      else v = w[q + q / 8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int jj = q - r;
        if (jj >= 0 && jj < 8) {
          lo[r] = fma(v.x, taps.lo[2 * jj], lo[r]);
          hi[r] = fma(v.x, (jj & 1) ? -taps.lo[15 - 2 * jj] : taps.lo[15 - 2 * jj], hi[r]);
        }
      }
      if constexpr (MODE == 2) v = w[q + q / 8 + 20];
      if constexpr (XLD > 0) if (q < XLD) { const double2 e = w[q + 40]; sink ^= __double2loint(e.x) ^ __double2hiint(e.y); }
      if constexpr (XST > 0) if (q < XST) const_cast<double2*>(w)[q + 60 + 9 * 32] = make_double2(lo[q & 7], hi[q & 7]);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int jj = q - r;
        if (jj >= 0 && jj < 8) {
          lo[r] = fma(v.y, taps.lo[2 * jj + 1], lo[r]);
          hi[r] = fma(v.y, (jj & 1) ? taps.lo[14 - 2 * jj] : -taps.lo[14 - 2 * jj], hi[r]);
        }
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += lo[i] + hi[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + sink;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double peak = double(p.multiProcessorCount) * 4 * 16 * 2 * clk * 1e3;
  printf("device %s sms %d clock %d kHz  lane peak %.2f TFLOP/s\n", p.name, p.multiProcessorCount, clk, peak * 1e-12);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double* out;
  cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 16 * 128);
  Taps t;
  for (int i = 0; i < 40; ++i) t.lo[i] = 1e-3 * (i + 1);
  const int smem = 2048 * 16 + 4096;  // per CTA: 10 CTAs per SM fit
  for (int wps : {1, 2, 4, 7, 10}) {            // warps per scheduler = CTAs (4 warps) per SM
    const int blocks = p.multiProcessorCount * wps;
    for (int mode = -1; mode <= 5; ++mode) {
      const int iters = 2000;
      float best = 1e9f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (mode < 0) k_regs<<<blocks, 128>>>(out, 1.0000001, 1e-9, iters * 16);
        else if (mode == 0) k_step<0><<<blocks, 128, smem>>>(out, t, iters);
        else if (mode == 1) k_step<1><<<blocks, 128, smem>>>(out, t, iters);
        else if (mode == 2) k_step<2><<<blocks, 128, smem>>>(out, t, iters);
        else if (mode == 3) k_step<1, 0, 8><<<blocks, 128, smem>>>(out, t, iters);
        else if (mode == 4) k_step<1, 8, 8><<<blocks, 128, smem>>>(out, t, iters);
        else k_step<1, 15, 8><<<blocks, 128, smem>>>(out, t, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      const double flops = 2.0 * 256 * iters * double(blocks) * 128;
      const char* names[] = {"regs", "taps", "taps+lds/32", "taps+lds/16", "lds12+sts8", "lds20+sts8", "lds27+sts8"};
      printf("warps/sched %2d  %-12s %8.3f ms  %6.2f TFLOP/s  %.3f of lane peak\n", wps, names[mode + 1], best, flops / best * 1e-9,
             flops / best * 1e3 / peak);
    }
  }
  printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
