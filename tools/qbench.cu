// qbench.cu - native A/B driver for the launch-shape switches of libjwave_cuda.so (no Python, no torch: a sweep of
// twenty JWC_TUNE variants costs seconds of GPU time instead of minutes).  Test infrastructure, not product.
//
//   tools/qbench <workload> <steps> [tune ...]          (tune = a JWC_TUNE string, "" = defaults)
//     workloads: c2 c3 c4 c5 (BASELINE.json's configs; c4 on QB_BATCH images, default 4), h1 (Haar1 FWT, c2's shape),
//                d10 / d20 / s20 (Daubechies10 / 20, Symlet20 FWT on c2's shape), w20 (Daubechies20 WPT, c3's shape),
//                w2d / w3d (Symlet8 WPT on 4 x 4096^2 images, 6 levels per axis / on a 512^3 volume, 5 levels)
//   QB_KERNELS=1 adds the per-kernel table of jwc_profile_report to every variant.
//
// Every variant runs on its own context (jwc_create reads JWC_TUNE): `steps` forward + reverse passes timed with CUDA
// events, the round-trip error, and the max abs difference of its forward output against the ONE-LEVEL GENERIC kernels
// (JWC_FORCE_GENERIC=1 - the on-GPU reference every fused kernel is tested against in tests/test_gpu_parity.py), so a
// faster variant that computes something else shows up in the same line.
//
// tools/qbench_taps.h: generated from jwave_b200/wavelets.py (hex float literals), see the snippet in its first line.
//   nvcc -O2 -std=c++17 -o tools/qbench tools/qbench.cu -Iinclude -Ljwave_b200 -ljwave_cuda -Xlinker -rpath='$ORIGIN/../jwave_b200'
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "jwave_cuda.h"
#include "qbench_taps.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__global__ void k_fill(double* x, size_t n, unsigned seed) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    unsigned long long z = (i + 1) * 0x9E3779B97F4A7C15ull + seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    x[i] = double(z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
  }
}
__global__ void k_maxdiff(const double* a, const double* b, size_t n, double* out) {
  double m = 0.0;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const double d = fabs(a[i] - b[i]);
    m = d > m || d != d ? (d != d ? 1e300 : d) : m;
  }
  for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(m));
}
static double maxdiff(const double* a, const double* b, size_t n, double* d_tmp) {
  CK(cudaMemset(d_tmp, 0, 8));
  k_maxdiff<<<148 * 8, 256>>>(a, b, n, d_tmp);
  double h;
  CK(cudaMemcpy(&h, d_tmp, 8, cudaMemcpyDeviceToHost));
  return h;
}

struct Work {
  const char* wavelet; int kind /*0 fwt1d 1 wpt1d 2 fwt2d 3 fwt3d*/; int n; int level; int64_t batch;
};

static int run(jwc_ctx* c, int wid, const Work& w, int dir, const double* in, double* out) {
  switch (w.kind) {
    case 0: return jwc_fwt1d_dev(c, wid, dir, in, out, w.batch, w.n, w.level);
    case 1: return jwc_wpt1d_dev(c, wid, dir, in, out, w.batch, w.n, w.level);
    case 2: return jwc_fwt2d_dev(c, wid, dir, in, out, w.batch, w.n, w.n, w.level, w.level);
    case 4: return jwc_wpt2d_dev(c, wid, dir, in, out, w.batch, w.n, w.n, w.level, w.level);
    case 5: return jwc_wpt3d_dev(c, wid, dir, in, out, w.n, w.n, w.n, w.level, w.level, w.level);
    default: return jwc_fwt3d_dev(c, wid, dir, in, out, w.n, w.n, w.n, w.level, w.level, w.level);
  }
}

int main(int argc, char** argv) {
  if (argc < 3) { printf("usage: qbench <workload> <steps> [tune ...]\n"); return 1; }
  const std::string wl = argv[1];
  const int steps = atoi(argv[2]);
  const int64_t qb = getenv("QB_BATCH") ? atoll(getenv("QB_BATCH")) : 0;
  Work w;
  if (wl == "c2") w = {"Daubechies4", 0, 1 << 14, 14, 65536};
  else if (wl == "h1") w = {"Haar1", 0, 1 << 14, 14, 65536};
  else if (wl == "d10") w = {"Daubechies10", 0, 1 << 14, 14, 65536};
  else if (wl == "d20") w = {"Daubechies20", 0, 1 << 14, 14, 65536};
  else if (wl == "s20") w = {"Symlet20", 0, 1 << 14, 14, 65536};
  else if (wl == "c3") w = {"Symlet8", 1, 1 << 16, 6, 4096};
  else if (wl == "w20") w = {"Daubechies20", 1, 1 << 16, 6, 4096};
  else if (wl == "c4") w = {"Daubechies20", 2, 8192, 13, 4};
  else if (wl == "c5") w = {"Coiflet5", 3, 1024, 10, 1};
  else if (wl == "w2d") w = {"Symlet8", 4, 4096, 6, 4};
  else if (wl == "w3d") w = {"Symlet8", 5, 512, 5, 1};
  else { printf("unknown workload %s\n", wl.c_str()); return 1; }
  if (qb > 0) w.batch = qb;
  const QTaps* tp = nullptr;
  for (const QTaps& t : kQTaps) if (!strcmp(t.name, w.wavelet)) tp = &t;
  if (!tp) { printf("no taps for %s\n", w.wavelet); return 1; }
  size_t count = size_t(w.batch) * w.n;
  if (w.kind == 2 || w.kind == 4) count *= w.n;
  if (w.kind == 3 || w.kind == 5) count = size_t(w.n) * w.n * w.n;
  double *x, *y, *z, *ref, *d_tmp;
  CK(cudaMalloc(&x, count * 8)); CK(cudaMalloc(&y, count * 8)); CK(cudaMalloc(&z, count * 8)); CK(cudaMalloc(&ref, count * 8));
  CK(cudaMalloc(&d_tmp, 8));
  k_fill<<<148 * 8, 256>>>(x, count, 12345u);
  CK(cudaDeviceSynchronize());
  // flops per sample (direct form), as bench.py counts them
  const int L = tp->L;
  double flops;
  const int axes = (w.kind == 0 || w.kind == 1) ? 1 : (w.kind == 2 || w.kind == 4) ? 2 : 3;
  if (w.kind == 1 || w.kind >= 4) flops = 2.0 * L * w.level * axes;
  else {
    const double per_axis = 4.0 * L * (1.0 - 1.0 / double(1 << w.level));
    flops = per_axis * axes;
  }
  const double t_hbm = 16.0 * axes / 6454.6e9, t_fp = flops / 36.7e12;
  const double t_roof = t_hbm > t_fp ? t_hbm : t_fp;
  printf("# %s: %s kind %d n %d level %d batch %lld  (%.3f G samples, roofline %s)\n", wl.c_str(), w.wavelet, w.kind, w.n, w.level,
         (long long)w.batch, count / 1e9, t_hbm > t_fp ? "hbm" : "fp64");

  // reference output: the one-level generic kernels
  {
    setenv("JWC_FORCE_GENERIC", "1", 1);
    unsetenv("JWC_TUNE");
    jwc_ctx* c; int wid;
    if (jwc_create(&c, 0)) { printf("jwc_create failed\n"); return 2; }
    if (jwc_set_wavelet(c, L, tp->f[0], tp->f[1], tp->f[2], tp->f[3], &wid)) { printf("set_wavelet: %s\n", jwc_last_error(c)); return 2; }
    if (run(c, wid, w, JWC_FORWARD, x, ref)) { printf("generic forward: %s\n", jwc_last_error(c)); return 2; }
    jwc_sync(c);
    jwc_destroy(c);
    unsetenv("JWC_FORCE_GENERIC");
  }
  std::vector<std::string> tunes;
  for (int i = 3; i < argc; ++i) tunes.push_back(argv[i]);
  if (tunes.empty()) tunes.push_back("");
  cudaEvent_t e0, e1, e2;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
  for (const std::string& t : tunes) {
    if (t.empty()) unsetenv("JWC_TUNE"); else setenv("JWC_TUNE", t.c_str(), 1);
    jwc_ctx* c; int wid;
    if (jwc_create(&c, 0)) { printf("%-44s jwc_create FAILED (bad tune?)\n", t.c_str()); continue; }
    if (jwc_set_wavelet(c, L, tp->f[0], tp->f[1], tp->f[2], tp->f[3], &wid)) { printf("set_wavelet: %s\n", jwc_last_error(c)); return 2; }
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    jwc_set_stream(c, s);
    bool ok = true;
    for (int i = 0; i < 2 && ok; ++i) {
      if (run(c, wid, w, JWC_FORWARD, x, y) || run(c, wid, w, JWC_REVERSE, y, z)) { printf("%-44s FAILED: %s\n", t.c_str(), jwc_last_error(c)); ok = false; }
    }
    if (!ok) { jwc_destroy(c); continue; }
    CK(cudaStreamSynchronize(s));
    const double err_fwd = maxdiff(y, ref, count, d_tmp), err_rt = maxdiff(z, x, count, d_tmp);
    if (getenv("QB_KERNELS")) jwc_profile_enable(c, 1);
    float ms_f = 0, ms_r = 0;
    for (int i = 0; i < steps; ++i) {
      CK(cudaEventRecord(e0, s));
      run(c, wid, w, JWC_FORWARD, x, y);
      CK(cudaEventRecord(e1, s));
      run(c, wid, w, JWC_REVERSE, y, z);
      CK(cudaEventRecord(e2, s));
      CK(cudaEventSynchronize(e2));
      float a, b;
      CK(cudaEventElapsedTime(&a, e0, e1)); CK(cudaEventElapsedTime(&b, e1, e2));
      ms_f += a; ms_r += b;
    }
    ms_f /= steps; ms_r /= steps;
    const double gf = count / (ms_f * 1e6), gr = count / (ms_r * 1e6);
    printf("%-44s fwd %8.3f ms %7.1f GS/s (%.3f)  rev %8.3f ms %7.1f GS/s (%.3f)  |fwd-generic| %.2e  rt %.2e\n",
           t.empty() ? "(default)" : t.c_str(), ms_f, gf, t_roof * count / (ms_f * 1e-3), ms_r, gr, t_roof * count / (ms_r * 1e-3), err_fwd, err_rt);
    if (getenv("QB_KERNELS")) {
      static char buf[1 << 16];
      jwc_profile_report(c, buf, sizeof buf);
      // label,launches,total_ms,samples_per_launch,levels
      char* save = nullptr;
      for (char* ln = strtok_r(buf, "\n", &save); ln; ln = strtok_r(nullptr, "\n", &save)) {
        char label[128]; long launches; double total, spl; int lev;
        if (sscanf(ln, "%127[^,],%ld,%lf,%lf,%d", label, &launches, &total, &spl, &lev) == 5) {
          const double per = total / launches, fl = (w.kind == 1 ? 2.0 * L * lev : 4.0 * L * (1.0 - 1.0 / double(1 << lev)));
          printf("      %-28s x%-4ld %8.4f ms  m=%-2d hbm %.3f  fp64 %.3f\n", label, launches / steps, per, lev, 16.0 * spl / (per * 1e-3) / 6454.6e9,
                 fl * spl / (per * 1e-3) / 36.7e12);
        }
      }
    }
    fflush(stdout);
    jwc_destroy(c);
    CK(cudaStreamDestroy(s));
  }
  return 0;
}
