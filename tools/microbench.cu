// microbench.cu - measures the two roofline denominators MEASURED_PEAKS.json lacks for this
// path: the FP64 (DFMA) vector peak and a double2 streaming-copy bandwidth.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b, int iters) {
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

__global__ void __launch_bounds__(256) k_copy(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  size_t stride = size_t(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) out[i] = in[i];
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("device %s sms %d smem_optin %zu l2 %d MB\n", p.name, p.multiProcessorCount, p.sharedMemPerBlockOptin,
         p.l2CacheSize >> 20);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double* out;
  const int blocks = p.multiProcessorCount * 8;
  cudaMalloc(&out, sizeof(double) * blocks * 256);
  const int iters = 4096;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k_dfma<<<blocks, 256>>>(out, 1.0000001, 1e-9, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 16 * iters * double(blocks) * 256;
    printf("dfma rep %d: %.3f ms  %.2f TFLOP/s\n", rep, ms, flops / ms * 1e-9);
  }
  const size_t n = size_t(1) << 29;  // 2^29 double2 = 8 GiB per buffer
  double2 *a, *b;
  cudaMalloc(&a, n * sizeof(double2));
  cudaMalloc(&b, n * sizeof(double2));
  cudaMemset(a, 1, n * sizeof(double2));
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k_copy<<<p.multiProcessorCount * 16, 256>>>(a, b, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("copy rep %d: %.3f ms  %.1f GB/s (read+write)\n", rep, ms, 2.0 * n * sizeof(double2) / ms * 1e-6);
  }
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    cudaMemcpyAsync(b, a, n * sizeof(double2), cudaMemcpyDeviceToDevice);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemcpy D2D rep %d: %.3f ms  %.1f GB/s (read+write)\n", rep, ms, 2.0 * n * sizeof(double2) / ms * 1e-6);
  }
  printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
