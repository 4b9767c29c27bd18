// tma_probe.cu - where does a {8 x 128} fp64 box land in shared memory under SWIZZLE_128B?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap tmap, double* out, int x, int y) {
  extern __shared__ __align__(1024) double smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 128 * 8);   // no static shared: keeps the dynamic base at offset 0
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(128 * 8 * 8) : "memory");
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(d), "l"(&tmap), "r"(x), "r"(y), "r"(b) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(b) : "memory");
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) out[i] = smem[i];
}

int main() {
  const int rows = 256, cols = 64;
  std::vector<double> h(rows * cols);
  for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) h[r * cols + c] = r * 1000 + c;
  double *d, *o;
  cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, 1024 * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows}; cuuint64_t str[1] = {cols * 8}; cuuint32_t box[2] = {8, 128}; cuuint32_t es[2] = {1, 1};
  CUresult rc = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc %d\n", (int)rc);
  probe<<<1, 128, 128 * 8 * 8 + 16>>>(m, o, 8, 16);
  printf("launch %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  std::vector<double> s(1024);
  cudaMemcpy(s.data(), o, 1024 * 8, cudaMemcpyDeviceToHost);
  // print, for the first 20 rows, the 16-byte chunk each (row, column pair) landed in
  for (int i = 0; i < 1024; i += 2) {
    int v = (int)s[i]; int r = v / 1000 - 16, c = v % 1000 - 8;
    if (r < 20) printf("smem dbl %4d (line %3d chunk %d) <- row %3d col %d,%d\n", i, i / 16, (i % 16) / 2, r, c, (int)s[i + 1] % 1000 - 8);
  }
  return 0;
}
