// jwc_capi.cu - the C ABI of libjwave_cuda.so (include/jwave_cuda.h): contexts, wavelet
// registration, argument checks with the reference's failure classes, the host-buffer
// staging pipeline, and the 2-D / 3-D drivers expressed as axis passes.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

#include "jwc_internal.cuh"
#include "jwc_plan.cuh"

using namespace jwc;

static std::string g_create_err;
static std::mutex g_create_mu;

#define JWC_CUDA(ctx, call)                                                              \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                  \
      return JWC_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define JWC_LOCK(ctx) std::lock_guard<std::recursive_mutex> lk__((ctx)->mu)

// the scratch buffers are busy until everything enqueued so far on the current stream has run
static void mark_scratch(jwc_ctx* ctx) {
  if (ctx->scratch_ev && cudaEventRecord(ctx->scratch_ev, ctx->stream) == cudaSuccess) ctx->scratch_ev_valid = true;
}

static int fail(jwc_ctx* ctx, int status, const char* msg) {
  if (ctx) ctx->err = msg;
  return status;
}

static void prof_clear(jwc_ctx* ctx);
static void group_destroy(jwc_ctx* owner);
static int group_set_wavelet(jwc_ctx* owner, int L, const double* sDe, const double* wDe, const double* sRe,
                             const double* wRe, int wid);
static int group_size(const jwc_ctx* ctx);

extern "C" int jwc_version(void) { return JWC_VERSION; }

extern "C" int jwc_create(jwc_ctx** out, int device) {
  if (!out) return JWC_ERR_ARG;
  *out = nullptr;
  auto bail = [&](const char* what, cudaError_t e) {
    std::lock_guard<std::mutex> lk(g_create_mu);
    g_create_err = std::string(what) + ": " + cudaGetErrorString(e);
    return int(JWC_ERR_CUDA);
  };
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess) return bail("cudaGetDeviceCount", e);
  if (device < 0 || device >= count) {
    std::lock_guard<std::mutex> lk(g_create_mu);
    g_create_err = "jwc_create: no such CUDA device";
    return JWC_ERR_ARG;
  }
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  jwc_ctx* ctx = new jwc_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    delete ctx;
    return bail("cudaGetDeviceProperties", e);
  }
  if (prop.major < 10) {
    delete ctx;
    std::lock_guard<std::mutex> lk(g_create_mu);
    g_create_err = "jwc_create: libjwave_cuda.so is built for sm_100a (B200) only";
    return JWC_ERR_CUDA;
  }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking)) != cudaSuccess) {
    delete ctx;
    return bail("cudaStreamCreate", e);
  }
  for (int i = 0; i < 2; ++i) {
    cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming);
  }
  ctx->stream = ctx->own_stream;
  cudaEventCreateWithFlags(&ctx->scratch_ev, cudaEventDisableTiming);
  const char* fg = getenv("JWC_FORCE_GENERIC");
  ctx->force_generic = fg && fg[0] == '1';
  if (const char* tune = getenv("JWC_TUNE")) {
    // "key=value,key=value": whole-key comparison (rev_tile must not match str_rev_tile), values checked
    struct Key { const char* name; int* dst; bool pow2; int lo, hi; };
    const Key keys[] = {
        {"fwd_tile", &ctx->fwd_tile, true, 256, 1 << 16},     {"fwd_m", &ctx->fwd_m, false, 0, 12},
        {"fwd_r", &ctx->fwd_r, true, 2, 8},                   {"rev_tile", &ctx->rev_tile, true, 256, 1 << 16},
        {"rev_m", &ctx->rev_m, false, 1, 12},                 {"rev_rs", &ctx->rev_rs, true, 2, 8},
        {"fwd_threads", &ctx->fwd_threads, false, 32, 1024},  {"rev_threads", &ctx->rev_threads, false, 32, 384},
        {"res_cap", &ctx->res_cap, true, 16, 1 << 14},        {"res_threads", &ctx->res_threads, false, 32, 1024},
        {"wpt_tile", &ctx->wpt_tile, true, 256, 1 << 16},     {"wpt_m", &ctx->wpt_m, false, 1, 12},
        {"wpt_rs", &ctx->wpt_rs, true, 4, 8},                 {"wpt_r", &ctx->wpt_r, true, 4, 8},
        {"wpt_threads", &ctx->wpt_threads, false, 64, 512},   {"wpt_inplace", &ctx->wpt_inplace, false, 0, 1},
        {"rev_tail", &ctx->rev_tail, false, 0, 1},            {"fwd_tail", &ctx->fwd_tail, false, 0, 1},
        {"str_tile", &ctx->str_tile, true, 128, 1 << 14},     {"str_rev_tile", &ctx->str_rev_tile, true, 128, 1 << 14},
        {"str_rev_m", &ctx->str_rev_m, false, 1, 12},         {"str_cap", &ctx->str_cap, true, 16, 1 << 14},
        {"str_threads", &ctx->str_threads, false, 32, 256},   {"str_rev_threads", &ctx->str_rev_threads, false, 32, 256},
        {"str_tma", &ctx->str_tma, false, 0, 1},              {"str_v2", &ctx->str_v2, false, 0, 1},
        {"str2_tile", &ctx->str2_tile, true, 128, 1 << 12},   {"str2_rev_tile", &ctx->str2_rev_tile, true, 128, 1 << 12},
        {"str2_cap", &ctx->str2_cap, true, 16, 1 << 11},      {"str2_m", &ctx->str2_m, false, 0, 8},
        {"str2_rev_m", &ctx->str2_rev_m, false, 0, 8},
        {"rot_warps", &ctx->rot_warps, false, 0, 1},               {"shfl", &ctx->shfl, false, 0, 1},
        {"carve", &ctx->carve, false, 0, 1},                     {"xsmem", &ctx->xsmem, false, 0, 200},
        {"stagger", &ctx->stagger, false, 0, 100000},              {"res_kb", &ctx->res_kb, false, 4, 200},
        {"wpt_transpose", &ctx->wpt_transpose, false, 0, 1},  {"res_split", &ctx->res_split, true, 0, 4096},
        {"pf", &ctx->pf, false, 0, 1 << 20},
        {"wpt_tma_store", &ctx->wpt_tma_store, false, 0, 1},  {"wpt_tma_store_fwd", &ctx->wpt_tma_store_fwd, false, 0, 1},
        {"wpt_rev_m", &ctx->wpt_rev_m, false, 0, 12},
    };
    std::string bad;
    const char* p = tune;
    while (*p) {
      const char* end = strchr(p, ',');
      const std::string tok = end ? std::string(p, end) : std::string(p);
      p = end ? end + 1 : p + tok.size();
      if (tok.empty()) continue;
      const size_t eq = tok.find('=');
      bool ok = false;
      if (eq != std::string::npos && eq + 1 < tok.size()) {
        const std::string name = tok.substr(0, eq);
        char* rest = nullptr;
        const long v = strtol(tok.c_str() + eq + 1, &rest, 10);
        for (const Key& k : keys) {
          if (name != k.name) continue;
          ok = rest && *rest == 0 && v >= k.lo && v <= k.hi && (!k.pow2 || (v & (v - 1)) == 0) &&
               (strstr(k.name, "threads") == nullptr || v % 32 == 0);
          if (ok) *k.dst = int(v);
          break;
        }
      }
      if (!ok && bad.empty()) bad = tok;
    }
    if (!bad.empty()) {
      jwc_destroy(ctx);
      std::lock_guard<std::mutex> lk(g_create_mu);
      g_create_err = "jwc_create: bad JWC_TUNE entry '" + bad + "' (unknown key, or value out of range / not a power of two)";
      return JWC_ERR_ARG;
    }
  }
  *out = ctx;
  return JWC_OK;
}

static void free_scratch(Scratch& s) {
  if (s.ptr) cudaFree(s.ptr);
  s.ptr = nullptr;
  s.bytes = 0;
}

extern "C" int jwc_destroy(jwc_ctx* ctx) {
  if (!ctx) return JWC_ERR_ARG;
  { JWC_LOCK(ctx); }  // wait for a call still in flight on another thread; the caller must not start new ones
  if (ctx->group) group_destroy(ctx);
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  prof_clear(ctx);
  for (auto& s : ctx->scratch) free_scratch(s);
  for (int i = 0; i < 2; ++i) {
    free_scratch(ctx->stage_in[i]);
    free_scratch(ctx->stage_out[i]);
    if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
    if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
    if (ctx->ev_out[i]) cudaEventDestroy(ctx->ev_out[i]);
  }
  if (ctx->scratch_ev) cudaEventDestroy(ctx->scratch_ev);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
  if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
  delete ctx;
  return JWC_OK;
}

extern "C" const char* jwc_last_error(const jwc_ctx* ctx) {
  if (ctx) return ctx->err.c_str();
  std::lock_guard<std::mutex> lk(g_create_mu);
  static thread_local std::string copy;
  copy = g_create_err;
  return copy.c_str();
}

// Switching streams: work already enqueued under the old stream may still be using the scratch buffers, so the
// new stream first waits (on the device, no host sync) for the event recorded after the last such launch.
static int switch_stream(jwc_ctx* ctx, cudaStream_t next) {
  if (next != ctx->stream && ctx->scratch_ev_valid) {
    JWC_CUDA(ctx, cudaSetDevice(ctx->device));
    JWC_CUDA(ctx, cudaStreamWaitEvent(next, ctx->scratch_ev, 0));
  }
  ctx->stream = next;
  return JWC_OK;
}

extern "C" int jwc_set_stream(jwc_ctx* ctx, void* cuda_stream) {
  if (!ctx) return JWC_ERR_ARG;
  JWC_LOCK(ctx);
  return switch_stream(ctx, static_cast<cudaStream_t>(cuda_stream));
}

extern "C" int jwc_reset_stream(jwc_ctx* ctx) {
  if (!ctx) return JWC_ERR_ARG;
  JWC_LOCK(ctx);
  return switch_stream(ctx, ctx->own_stream);
}

extern "C" int jwc_sync(jwc_ctx* ctx) {
  if (!ctx) return JWC_ERR_ARG;
  JWC_LOCK(ctx);
  JWC_CUDA(ctx, cudaSetDevice(ctx->device));
  JWC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return JWC_OK;
}

extern "C" int64_t jwc_launch_count(const jwc_ctx* ctx) { return ctx ? ctx->launches : -1; }

static void prof_clear(jwc_ctx* ctx) {
  for (auto& r : ctx->prof) {
    cudaEventDestroy(r.beg);
    cudaEventDestroy(r.end);
  }
  ctx->prof.clear();
}

extern "C" int jwc_profile_enable(jwc_ctx* ctx, int on) {
  if (!ctx) return JWC_ERR_ARG;
  JWC_LOCK(ctx);
  JWC_CUDA(ctx, cudaSetDevice(ctx->device));
  prof_clear(ctx);
  ctx->prof_on = on != 0;
  return JWC_OK;
}

// Text report, one line per kernel label: "name,launches,total_ms,samples_per_launch,levels".  Waits for the
// recorded launches, then clears the records.
extern "C" int jwc_profile_report(jwc_ctx* ctx, char* buf, size_t size) {
  if (!ctx || !buf || size == 0) return JWC_ERR_ARG;
  JWC_LOCK(ctx);
  JWC_CUDA(ctx, cudaSetDevice(ctx->device));
  struct Acc { const char* name; int n; double ms, units; int levels; };
  std::vector<Acc> acc;
  for (auto& r : ctx->prof) {
    JWC_CUDA(ctx, cudaEventSynchronize(r.end));
    float ms = 0.f;
    JWC_CUDA(ctx, cudaEventElapsedTime(&ms, r.beg, r.end));
    Acc* hit = nullptr;
    for (auto& a : acc)
      if (!strcmp(a.name, r.name) && a.units == r.units && a.levels == r.levels) hit = &a;
    if (!hit) {
      acc.push_back({r.name, 0, 0.0, r.units, r.levels});
      hit = &acc.back();
    }
    hit->n++;
    hit->ms += ms;
  }
  std::string out;
  char line[256];
  for (auto& a : acc) {
    snprintf(line, sizeof(line), "%s,%d,%.6f,%.0f,%d\n", a.name, a.n, a.ms, a.units, a.levels);
    out += line;
  }
  prof_clear(ctx);
  if (out.size() + 1 > size) return fail(ctx, JWC_ERR_ARG, "jwc_profile_report: buffer too small");
  memcpy(buf, out.c_str(), out.size() + 1);
  return JWC_OK;
}

extern "C" int jwc_set_wavelet(jwc_ctx* ctx, int L, const double* sDe, const double* wDe,
                               const double* sRe, const double* wRe, int* wid) {
  if (!ctx) return JWC_ERR_ARG;
  JWC_LOCK(ctx);
  if (!sDe || !wDe || !sRe || !wRe || !wid) return fail(ctx, JWC_ERR_ARG, "jwc_set_wavelet: null argument");
  if (L < 2 || L > JWC_MAX_TAPS || (L & 1))
    return fail(ctx, JWC_ERR_ARG, "jwc_set_wavelet: filter length must be even and within 2..40");
  WaveletRec rec;
  memset(&rec, 0, sizeof(rec));
  rec.L = L;
  for (int j = 0; j < L; ++j) {
    rec.de.lo[j] = sDe[j];
    rec.de.hi[j] = wDe[j];
    rec.re.lo[j] = sRe[j];
    rec.re.hi[j] = wRe[j];
  }
  rec.mirror_de = rec.mirror_re = true;
  for (int j = 0; j < L; ++j) {
    const double de = (j & 1) ? -sDe[L - 1 - j] : sDe[L - 1 - j];
    const double re = (j & 1) ? -sRe[L - 1 - j] : sRe[L - 1 - j];
    if (memcmp(&de, &wDe[j], sizeof(double)) != 0) rec.mirror_de = false;
    if (memcmp(&re, &wRe[j], sizeof(double)) != 0) rec.mirror_re = false;
  }
  ctx->wavelets.push_back(rec);
  *wid = int(ctx->wavelets.size()) - 1;
  if (ctx->group) return group_set_wavelet(ctx, L, sDe, wDe, sRe, wRe, *wid);  // same handle on every device
  return JWC_OK;
}

// ---- argument checks with the reference's failure classes --------------------------------------

static bool is_binary(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }  // MathToolKit.java:185-189
static int exponent(int64_t v) {                                          // MathToolKit.java:202-208 (F14)
  int p = 0;
  while ((int64_t(1) << (p + 1)) <= v) ++p;
  return p;
}

static int check_axis(jwc_ctx* ctx, int n, int level) {
  if (!is_binary(n)) return fail(ctx, JWC_ERR_NOT_BINARY, "given array length is not 2^p | p E N");
  if (level < 0 || level > exponent(n)) return fail(ctx, JWC_ERR_LEVEL, "given level is out of range for given array");
  return JWC_OK;
}

static int check_common(jwc_ctx* ctx, int wid, int kind, int dir, const void* in, const void* out) {
  if (!ctx) return JWC_ERR_ARG;
  JWC_LOCK(ctx);
  if (wid < 0 || wid >= int(ctx->wavelets.size())) return fail(ctx, JWC_ERR_ARG, "unknown wavelet handle");
  if (kind != JWC_FWT && kind != JWC_WPT) return fail(ctx, JWC_ERR_ARG, "kind must be JWC_FWT or JWC_WPT");
  if (dir != JWC_FORWARD && dir != JWC_REVERSE) return fail(ctx, JWC_ERR_ARG, "dir must be JWC_FORWARD or JWC_REVERSE");
  if (!in || !out) return fail(ctx, JWC_ERR_ARG, "null data pointer");
  return JWC_OK;
}

static bool overlaps(const double* a, const double* b, int64_t count) {
  return a < b + count && b < a + count;
}

// ---- device-resident drivers --------------------------------------------------------------------

static int axis_dev(jwc_ctx* ctx, int wid, int kind, int dir, const double* in, double* out,
                    int64_t outer, int n, int64_t inner, int level) {
  int st = check_axis(ctx, n, level);
  if (st) return st;
  if (outer < 0 || inner < 1) return fail(ctx, JWC_ERR_ARG, "negative batch");
  if (outer == 0) return JWC_OK;
  if (overlaps(in, out, outer * n * inner)) return fail(ctx, JWC_ERR_ARG, "in and out overlap");
  JWC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaError_t e = run_axis(ctx, ctx->wavelets[wid], kind, dir, in, out, outer, n, inner, level);
  mark_scratch(ctx);
  if (e != cudaSuccess) {
    ctx->err = std::string("axis transform: ") + cudaGetErrorString(e);
    return JWC_ERR_CUDA;
  }
  return JWC_OK;
}

extern "C" int jwc_axis_dev(jwc_ctx* ctx, int wid, int kind, int dir, const double* in, double* out,
                            int64_t outer, int n, int64_t inner, int level) {
  int st = check_common(ctx, wid, kind, dir, in, out);
  if (st) return st;
  JWC_LOCK(ctx);
  return axis_dev(ctx, wid, kind, dir, in, out, outer, n, inner, level);
}

extern "C" int jwc_axis_dev_remote(jwc_ctx* ctx, int wid, int kind, int dir, const double* in, int64_t outer, int n,
                                   int64_t inner, int level, const jwc_remote_map* map) {
  if (!ctx) return JWC_ERR_ARG;
  JWC_LOCK(ctx);
  if (!map) return fail(ctx, JWC_ERR_ARG, "null remote map");
  int st = check_common(ctx, wid, kind, dir, in, map->peer[0]);
  if (st) return st;
  if ((st = check_axis(ctx, n, level))) return st;
  if (kind != JWC_FWT || level < 1) return fail(ctx, JWC_ERR_ARG, "remote stores: FWT with level >= 1 only");
  if (map->world < 1 || map->world > 8 || map->lg_seg < 0 || map->lg_hi < 0)
    return fail(ctx, JWC_ERR_ARG, "remote map: bad geometry");
  const int64_t rows = map->mode == 1 ? n : (map->mode == 2 ? (int64_t(1) << map->lg_hi) : -1);
  if (rows < 0 || ((rows - 1) >> map->lg_seg) >= map->world) return fail(ctx, JWC_ERR_ARG, "remote map: bad geometry");
  if (map->mode == 1 && (inner < 8 || inner % 8)) return fail(ctx, JWC_ERR_ARG, "remote map mode 1 needs inner % 8 == 0");
  if (map->mode == 2 && (inner != 1 || dir != JWC_REVERSE || (outer & (rows - 1))))
    return fail(ctx, JWC_ERR_ARG, "remote map mode 2: contiguous reverse passes over whole outer blocks only");
  if (outer < 1) return fail(ctx, JWC_ERR_ARG, "negative batch");
  RemoteMap rm;
  rm.mode = map->mode; rm.lg_seg = map->lg_seg; rm.lg_hi = map->lg_hi;
  rm.outer_stride = map->outer_stride; rm.row_stride = map->row_stride; rm.base_off = map->base_off;
  for (int i = 0; i < map->world; ++i) {
    if (!map->peer[i] || (reinterpret_cast<uintptr_t>(map->peer[i]) & 31)) return fail(ctx, JWC_ERR_ARG, "remote map: peer pointer");
    rm.peer[i] = static_cast<double*>(map->peer[i]);
  }
  if ((rm.outer_stride | rm.row_stride | rm.base_off) & 3) return fail(ctx, JWC_ERR_ARG, "remote map: strides must keep 32-byte alignment");
  JWC_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->remote = &rm;
  cudaError_t e = run_axis(ctx, ctx->wavelets[wid], kind, dir, in, rm.peer[0], outer, n, inner, level);
  ctx->remote = nullptr;
  mark_scratch(ctx);
  if (e == cudaErrorNotSupported) return fail(ctx, JWC_ERR_ARG, "remote stores: shape not covered by the fused kernels");
  if (e != cudaSuccess) {
    ctx->err = std::string("axis transform: ") + cudaGetErrorString(e);
    return JWC_ERR_CUDA;
  }
  return JWC_OK;
}

static int t1d_dev(jwc_ctx* ctx, int wid, int kind, int dir, const double* in, double* out,
                   int64_t batch, int n, int level) {
  int st = check_common(ctx, wid, kind, dir, in, out);
  if (st) return st;
  JWC_LOCK(ctx);
  return axis_dev(ctx, wid, kind, dir, in, out, batch, n, 1, level);
}

extern "C" int jwc_fwt1d_dev(jwc_ctx* ctx, int wid, int dir, const double* in, double* out,
                             int64_t batch, int n, int level) {
  return t1d_dev(ctx, wid, JWC_FWT, dir, in, out, batch, n, level);
}
extern "C" int jwc_wpt1d_dev(jwc_ctx* ctx, int wid, int dir, const double* in, double* out,
                             int64_t batch, int n, int level) {
  return t1d_dev(ctx, wid, JWC_WPT, dir, in, out, batch, n, level);
}

static int ensure(jwc_ctx* ctx, Scratch& s, size_t bytes) {
  if (s.bytes >= bytes) return JWC_OK;
  if (s.ptr) {
    JWC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    JWC_CUDA(ctx, cudaFree(s.ptr));
    s.ptr = nullptr;
    s.bytes = 0;
  }
  JWC_CUDA(ctx, cudaMalloc(&s.ptr, bytes));
  s.bytes = bytes;
  return JWC_OK;
}

// BasicTransform.java:361-399 / :436-474.  forward: rows (axis 1, lvlN) then columns (axis 0,
// lvlM); reverse: columns then rows.  Both passes run over the whole batch.
static int t2d_dev(jwc_ctx* ctx, int wid, int kind, int dir, const double* in, double* out,
                   int64_t batch, int rows, int cols, int lvlM, int lvlN) {
  int st = check_common(ctx, wid, kind, dir, in, out);
  if (st) return st;
  JWC_LOCK(ctx);
  // the reference transforms rows first (forward) / columns first (reverse): report in that order
  if (dir == JWC_FORWARD) {
    if ((st = check_axis(ctx, cols, lvlN)) || (st = check_axis(ctx, rows, lvlM))) return st;
  } else {
    if ((st = check_axis(ctx, rows, lvlM)) || (st = check_axis(ctx, cols, lvlN))) return st;
  }
  if (batch < 0) return fail(ctx, JWC_ERR_ARG, "negative batch");
  if (batch == 0) return JWC_OK;
  const int64_t total = batch * rows * cols;
  if (overlaps(in, out, total)) return fail(ctx, JWC_ERR_ARG, "in and out overlap");
  JWC_CUDA(ctx, cudaSetDevice(ctx->device));
  if ((st = ensure(ctx, ctx->scratch[2], size_t(total) * sizeof(double)))) return st;
  double* tmp = static_cast<double*>(ctx->scratch[2].ptr);
  if (dir == JWC_FORWARD) {
    if ((st = axis_dev(ctx, wid, kind, dir, in, tmp, batch * rows, cols, 1, lvlN))) return st;
    return axis_dev(ctx, wid, kind, dir, tmp, out, batch, rows, cols, lvlM);
  }
  if ((st = axis_dev(ctx, wid, kind, dir, in, tmp, batch, rows, cols, lvlM))) return st;
  return axis_dev(ctx, wid, kind, dir, tmp, out, batch * rows, cols, 1, lvlN);
}

extern "C" int jwc_fwt2d_dev(jwc_ctx* ctx, int wid, int dir, const double* in, double* out,
                             int64_t batch, int rows, int cols, int lvlM, int lvlN) {
  return t2d_dev(ctx, wid, JWC_FWT, dir, in, out, batch, rows, cols, lvlM, lvlN);
}
extern "C" int jwc_wpt2d_dev(jwc_ctx* ctx, int wid, int dir, const double* in, double* out,
                             int64_t batch, int rows, int cols, int lvlM, int lvlN) {
  return t2d_dev(ctx, wid, JWC_WPT, dir, in, out, batch, rows, cols, lvlM, lvlN);
}

// BasicTransform.java:509-566 / :602-659.  The 2-D call on each [j][k] slice receives
// (lvlP, lvlQ) as (lvlM, lvlN): axis k (length R) gets lvlQ, axis j (length Q) gets lvlP; the
// outer axis i (length P) gets lvlR.  forward: k, j, i.  reverse: j, k, i.
static int t3d_dev(jwc_ctx* ctx, int wid, int kind, int dir, const double* in, double* out, int P,
                   int Q, int R, int lvlP, int lvlQ, int lvlR) {
  int st = check_common(ctx, wid, kind, dir, in, out);
  if (st) return st;
  JWC_LOCK(ctx);
  if (dir == JWC_FORWARD) {
    if ((st = check_axis(ctx, R, lvlQ)) || (st = check_axis(ctx, Q, lvlP))) return st;
  } else {
    if ((st = check_axis(ctx, Q, lvlP)) || (st = check_axis(ctx, R, lvlQ))) return st;
  }
  if ((st = check_axis(ctx, P, lvlR))) return st;
  const int64_t total = int64_t(P) * Q * R;
  if (overlaps(in, out, total)) return fail(ctx, JWC_ERR_ARG, "in and out overlap");
  JWC_CUDA(ctx, cudaSetDevice(ctx->device));
  if ((st = ensure(ctx, ctx->scratch[2], size_t(total) * sizeof(double)))) return st;
  double* tmp = static_cast<double*>(ctx->scratch[2].ptr);
  // in -> out -> tmp -> out
  if (dir == JWC_FORWARD) {
    if ((st = axis_dev(ctx, wid, kind, dir, in, out, int64_t(P) * Q, R, 1, lvlQ))) return st;
    if ((st = axis_dev(ctx, wid, kind, dir, out, tmp, P, Q, R, lvlP))) return st;
  } else {
    if ((st = axis_dev(ctx, wid, kind, dir, in, out, P, Q, R, lvlP))) return st;
    if ((st = axis_dev(ctx, wid, kind, dir, out, tmp, int64_t(P) * Q, R, 1, lvlQ))) return st;
  }
  return axis_dev(ctx, wid, kind, dir, tmp, out, 1, P, int64_t(Q) * R, lvlR);
}

extern "C" int jwc_fwt3d_dev(jwc_ctx* ctx, int wid, int dir, const double* in, double* out, int P,
                             int Q, int R, int lvlP, int lvlQ, int lvlR) {
  return t3d_dev(ctx, wid, JWC_FWT, dir, in, out, P, Q, R, lvlP, lvlQ, lvlR);
}
extern "C" int jwc_wpt3d_dev(jwc_ctx* ctx, int wid, int dir, const double* in, double* out, int P,
                             int Q, int R, int lvlP, int lvlQ, int lvlR) {
  return t3d_dev(ctx, wid, JWC_WPT, dir, in, out, P, Q, R, lvlP, lvlQ, lvlR);
}

// AncientEgyptianDecomposition.java:97-129 / :144-183.  Every 2^p block of the binary expansion of n
// (MathToolKit.java:57-84, largest first) is gathered from all signals into a dense [batch][2^p]
// scratch array, transformed at full depth and scattered back (2-D device copies; the blocks of a
// signal are neither aligned nor equally strided, which the fused kernels want).
static int aed_dev(jwc_ctx* ctx, int wid, int kind, int dir, const double* in, double* out, int64_t batch, int n) {
  if (n < 1) return fail(ctx, JWC_ERR_ARG, "the supported number for decomposition is smaller than one");
  if (batch < 0) return fail(ctx, JWC_ERR_ARG, "negative batch");
  if (batch == 0) return JWC_OK;
  if (overlaps(in, out, batch * n)) return fail(ctx, JWC_ERR_ARG, "in and out overlap");
  JWC_CUDA(ctx, cudaSetDevice(ctx->device));
  const int top = exponent(n);  // largest block
  int st = ensure(ctx, ctx->scratch[2], size_t(2) * batch * (size_t(1) << top) * sizeof(double));
  if (st) return st;
  double* dense_in = static_cast<double*>(ctx->scratch[2].ptr);
  double* dense_out = dense_in + batch * (int64_t(1) << top);
  const size_t pitch = size_t(n) * sizeof(double);
  int64_t off = 0;
  for (int p = top; p >= 0; --p) {
    if (!((n >> p) & 1)) continue;
    const int len = 1 << p;
    const size_t row = size_t(len) * sizeof(double);
    // FWT blocks of at least 4 samples in signals whose length is a multiple of 4: the fused kernels read and write
    // the block in place of the signal (line pitch n), no gather / scatter passes (2 x 16 B per sample less)
    if (kind == JWC_FWT && p >= 2 && jwc::fwt_pitched_ok(ctx, ctx->wavelets[wid], dir, in + off, out + off, len, n, n)) {
      ctx->pitch_in = ctx->pitch_out = n;
      st = axis_dev(ctx, wid, kind, dir, in + off, out + off, batch, len, 1, p);
      ctx->pitch_in = ctx->pitch_out = 0;
      if (st) return st;
      off += len;
      continue;
    }
    JWC_CUDA(ctx, cudaMemcpy2DAsync(dense_in, row, in + off, pitch, row, size_t(batch), cudaMemcpyDeviceToDevice, ctx->stream));
    if ((st = axis_dev(ctx, wid, kind, dir, dense_in, dense_out, batch, len, 1, p))) return st;
    JWC_CUDA(ctx, cudaMemcpy2DAsync(out + off, pitch, dense_out, row, row, size_t(batch), cudaMemcpyDeviceToDevice, ctx->stream));
    off += len;
  }
  return JWC_OK;
}

extern "C" int jwc_aed1d_dev(jwc_ctx* ctx, int wid, int kind, int dir, const double* in, double* out,
                             int64_t batch, int n) {
  int st = check_common(ctx, wid, kind, dir, in, out);
  if (st) return st;
  JWC_LOCK(ctx);
  return aed_dev(ctx, wid, kind, dir, in, out, batch, n);
}

// WaveletTransform.decompose (WaveletTransform.java:136-146): out[b][p][.] = forward(in[b], p).  Row p + 1
// is row p with one more level applied - to its approximation prefix (FWT; the detail tail is copied) or
// to every one of its 2^p packets (WPT) - so each row costs one one-level launch.
static int decompose_dev(jwc_ctx* ctx, int wid, int kind, const double* in, double* out, int64_t batch, int n) {
  int st = check_axis(ctx, n, 0);
  if (st) return st;
  if (batch < 0) return fail(ctx, JWC_ERR_ARG, "negative batch");
  if (batch == 0) return JWC_OK;
  const int P = exponent(n);
  const int64_t sig = int64_t(P + 1) * n;  // doubles per signal in `out`
  if (in < out + batch * sig && out < in + batch * n) return fail(ctx, JWC_ERR_ARG, "in and out overlap");
  JWC_CUDA(ctx, cudaSetDevice(ctx->device));
  const WaveletRec& w = ctx->wavelets[wid];
  double* D[2] = {nullptr, nullptr};
  if (kind == JWC_WPT && P > 0) {
    cudaError_t e0 = ensure_scratch(ctx, 0, size_t(batch) * n * sizeof(double), &D[0]);
    if (e0 == cudaSuccess) e0 = ensure_scratch(ctx, 1, size_t(batch) * n * sizeof(double), &D[1]);
    if (e0 != cudaSuccess) {
      ctx->err = std::string("decompose: ") + cudaGetErrorString(e0);
      return JWC_ERR_CUDA;
    }
  }
  JWC_CUDA(ctx, cudaMemcpy2DAsync(out, size_t(sig) * sizeof(double), in, size_t(n) * sizeof(double),
                                  size_t(n) * sizeof(double), size_t(batch), cudaMemcpyDeviceToDevice, ctx->stream));
  for (int p = 0; p < P; ++p) {
    const int h = n >> p;
    FwdLevelArgs a;
    a.inner = 1;
    a.half = h / 2;
    cudaError_t e = cudaSuccess;
    if (kind == JWC_FWT) {
      a.src = out + int64_t(p) * n;            a.src_os = sig;
      a.dstA = out + int64_t(p + 1) * n;       a.dstA_os = sig;
      a.dstD = a.dstA + h / 2;                 a.dstD_os = sig;
      a.outer = batch;
      e = launch_fwd_level_generic(ctx, w.L, w.de, a);
      if (e == cudaSuccess && h < n)
        e = cudaMemcpy2DAsync(out + int64_t(p + 1) * n + h, size_t(sig) * sizeof(double), out + int64_t(p) * n + h,
                              size_t(sig) * sizeof(double), size_t(n - h) * sizeof(double), size_t(batch),
                              cudaMemcpyDeviceToDevice, ctx->stream);
    } else {
      // every signal's 2^p packets of width h in ONE launch: level p runs dense [batch][n] -> dense [batch][n]
      // (batch * 2^p lines of stride h) between two scratch arrays, and the new row is copied into out[b][p + 1]
      const double* src = (p == 0) ? in : D[p & 1];
      a.src = src;                    a.src_os = h;
      a.dstA = D[(p + 1) & 1];        a.dstA_os = h;
      a.dstD = a.dstA + h / 2;        a.dstD_os = h;
      a.outer = batch << p;
      e = launch_fwd_level_generic(ctx, w.L, w.de, a);
      if (e == cudaSuccess)
        e = cudaMemcpy2DAsync(out + int64_t(p + 1) * n, size_t(sig) * sizeof(double), D[(p + 1) & 1], size_t(n) * sizeof(double),
                              size_t(n) * sizeof(double), size_t(batch), cudaMemcpyDeviceToDevice, ctx->stream);
    }
    if (e != cudaSuccess) {
      ctx->err = std::string("decompose: ") + cudaGetErrorString(e);
      return JWC_ERR_CUDA;
    }
  }
  mark_scratch(ctx);
  return JWC_OK;
}

extern "C" int jwc_decompose1d_dev(jwc_ctx* ctx, int wid, int kind, const double* in, double* out, int64_t batch,
                                   int n) {
  int st = check_common(ctx, wid, kind, JWC_FORWARD, in, out);
  if (st) return st;
  JWC_LOCK(ctx);
  return decompose_dev(ctx, wid, kind, in, out, batch, n);
}

// scratch of the compressor kernels: blocks partial sums, the magnitude, the CTA counter (zero between calls)
static int compress_scratch(jwc_ctx* ctx, int blocks, double** scratch) {
  const size_t need = size_t(blocks + 2) * sizeof(double);
  if (ctx->scratch[3].bytes < need) {
    int st = ensure(ctx, ctx->scratch[3], need);
    if (st) return st;
    JWC_CUDA(ctx, cudaMemsetAsync(ctx->scratch[3].ptr, 0, need, ctx->stream));
  }
  *scratch = static_cast<double*>(ctx->scratch[3].ptr);
  return JWC_OK;
}

// CompressorMagnitude on device-resident coefficients; the magnitude stays on the device
static int compress_dev(jwc_ctx* ctx, const double* in, double* out, int64_t count, double threshold,
                        double* magnitude_dev) {
  if (!ctx) return JWC_ERR_ARG;
  JWC_LOCK(ctx);
  if (!in || !out) return fail(ctx, JWC_ERR_ARG, "null data pointer");
  if (!(threshold > 0.)) return fail(ctx, JWC_ERR_ARG, "Compressor - given threshold should be larger than zero!");
  if (count < 1) return fail(ctx, JWC_ERR_ARG, "Compressor - empty array");
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 31)
    return fail(ctx, JWC_ERR_ARG, "Compressor - device arrays must be 32-byte aligned");
  JWC_CUDA(ctx, cudaSetDevice(ctx->device));
  const int blocks = ctx->sm_count * 8;
  double* scratch = nullptr;
  int st = compress_scratch(ctx, blocks, &scratch);
  if (st) return st;
  cudaError_t e = launch_compress_magnitude(ctx, in, out, count, threshold, scratch, blocks);
  mark_scratch(ctx);
  if (e != cudaSuccess) {
    ctx->err = std::string("compress: ") + cudaGetErrorString(e);
    return JWC_ERR_CUDA;
  }
  if (magnitude_dev)
    JWC_CUDA(ctx, cudaMemcpyAsync(magnitude_dev, scratch + blocks, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  return JWC_OK;
}

// Forward transform + CompressorMagnitude in one call (the compression use case: Compressor.java:97-110 applied to
// the output of FastWaveletTransform / WaveletPacketTransform.forward).  The magnitude is a mean over ALL
// coefficients, so the threshold pass cannot start before the last coefficient exists; what CAN be saved is the
// read of the reduce pass: the batch is transformed in chunks of at most 48 MB and the |c| sum of a chunk runs
// right behind its forward transform, while the chunk still sits in the 126 MB L2.  HBM traffic per coefficient:
// 16 B (transform) + 16 B (threshold, in place) instead of + 8 B for a separate reduce.
extern "C" int jwc_forward1d_compress_dev(jwc_ctx* ctx, int wid, int kind, const double* in, double* out, int64_t batch,
                                          int n, int level, double threshold, double* magnitude_dev) {
  int st = check_common(ctx, wid, kind, JWC_FORWARD, in, out);
  if (st) return st;
  JWC_LOCK(ctx);
  if ((st = check_axis(ctx, n, level))) return st;
  if (!(threshold > 0.)) return fail(ctx, JWC_ERR_ARG, "Compressor - given threshold should be larger than zero!");
  if (batch < 1) return fail(ctx, JWC_ERR_ARG, "Compressor - empty array");
  if ((reinterpret_cast<uintptr_t>(out) & 31) || (n & 3)) return fail(ctx, JWC_ERR_ARG, "Compressor - n % 4 == 0 and 32-byte aligned arrays");
  if (overlaps(in, out, batch * n)) return fail(ctx, JWC_ERR_ARG, "in and out overlap");
  JWC_CUDA(ctx, cudaSetDevice(ctx->device));
  const int blocks = ctx->sm_count * 8;
  double* scratch = nullptr;
  if ((st = compress_scratch(ctx, blocks, &scratch))) return st;
  int64_t chunk = (int64_t(48) << 20) / (int64_t(n) * int64_t(sizeof(double)));
  if (chunk < 1) chunk = 1;
  const int64_t total = batch * n;
  for (int64_t first = 0; first < batch; first += chunk) {
    const int64_t cnt = batch - first < chunk ? batch - first : chunk;
    if ((st = axis_dev(ctx, wid, kind, JWC_FORWARD, in + first * n, out + first * n, cnt, n, 1, level))) return st;
    cudaError_t e = launch_abs_sum(ctx, out + first * n, cnt * n, total, scratch, blocks, first > 0, first + cnt >= batch);
    if (e != cudaSuccess) {
      ctx->err = std::string("compress: ") + cudaGetErrorString(e);
      return JWC_ERR_CUDA;
    }
  }
  cudaError_t e = launch_threshold(ctx, out, out, total, threshold, scratch, blocks);
  mark_scratch(ctx);
  if (e != cudaSuccess) {
    ctx->err = std::string("compress: ") + cudaGetErrorString(e);
    return JWC_ERR_CUDA;
  }
  if (magnitude_dev)
    JWC_CUDA(ctx, cudaMemcpyAsync(magnitude_dev, scratch + blocks, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  return JWC_OK;
}

extern "C" int jwc_compress_magnitude_dev(jwc_ctx* ctx, const double* in, double* out, int64_t count,
                                          double threshold, double* magnitude_dev) {
  return compress_dev(ctx, in, out, count, threshold, magnitude_dev);
}

// Host buffers: the magnitude is a global mean, so the whole array has to be resident at once.
extern "C" int jwc_compress_magnitude(jwc_ctx* ctx, const double* in, double* out, int64_t count, double threshold,
                                      double* magnitude) {
  if (!ctx) return JWC_ERR_ARG;
  JWC_LOCK(ctx);
  if (!in || !out) return fail(ctx, JWC_ERR_ARG, "null data pointer");
  if (!(threshold > 0.)) return fail(ctx, JWC_ERR_ARG, "Compressor - given threshold should be larger than zero!");
  if (count < 1) return fail(ctx, JWC_ERR_ARG, "Compressor - empty array");
  JWC_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t bytes = size_t(count) * sizeof(double);
  int st;
  if ((st = ensure(ctx, ctx->stage_in[0], bytes)) || (st = ensure(ctx, ctx->stage_out[0], bytes))) return st;
  double* d_in = static_cast<double*>(ctx->stage_in[0].ptr);
  double* d_out = static_cast<double*>(ctx->stage_out[0].ptr);
  cudaStream_t user_stream = ctx->stream;
  if (user_stream != ctx->own_stream) JWC_CUDA(ctx, cudaStreamSynchronize(user_stream));
  ctx->stream = ctx->own_stream;
  cudaError_t e = cudaMemcpyAsync(d_in, in, bytes, cudaMemcpyHostToDevice, ctx->stream);
  st = (e == cudaSuccess) ? compress_dev(ctx, d_in, d_out, count, threshold, nullptr) : JWC_ERR_CUDA;
  if (!st) {
    e = cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && magnitude)
      e = cudaMemcpyAsync(magnitude, static_cast<double*>(ctx->scratch[3].ptr) + ctx->sm_count * 8 /* = blocks */, sizeof(double),
                          cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) st = JWC_ERR_CUDA;
  }
  ctx->stream = user_stream;
  if (st == JWC_ERR_CUDA && e != cudaSuccess) ctx->err = std::string("compress: ") + cudaGetErrorString(e);
  return st;
}

// ---- host-buffer entry points ---------------------------------------------------------------------
//
// The batch is cut into chunks of whole items (signals / matrices); chunk c uses staging slot
// c & 1, so the H2D copy of chunk c+1 and the D2H copy of chunk c-1 overlap the kernels of
// chunk c (three streams, events between them).  Pinned host memory (jwc_host_alloc_pinned)
// makes the copies truly asynchronous; pageable memory works, only slower.

template <class Run>
static int staged(jwc_ctx* ctx, const double* in, double* out, int64_t items, int64_t item_elems, Run run) {
  if (items == 0) return JWC_OK;
  JWC_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t item_bytes = size_t(item_elems) * sizeof(double);
  int64_t per_chunk = int64_t(ctx->staging_bytes / item_bytes);
  if (per_chunk < 1) per_chunk = 1;
  if (per_chunk > items) per_chunk = items;
  const int slots = per_chunk < items ? 2 : 1;
  int st;
  for (int s = 0; s < slots; ++s) {
    if ((st = ensure(ctx, ctx->stage_in[s], size_t(per_chunk) * item_bytes))) return st;
    if ((st = ensure(ctx, ctx->stage_out[s], size_t(per_chunk) * item_bytes))) return st;
  }
  cudaStream_t user_stream = ctx->stream;
  if (user_stream != ctx->own_stream)  // device-resident work may still be using the scratch buffers
    JWC_CUDA(ctx, cudaStreamSynchronize(user_stream));
  ctx->stream = ctx->own_stream;  // the pipeline owns its three streams
  int status = JWC_OK;
  int64_t c = 0;
  for (int64_t first = 0; first < items && !status; first += per_chunk, ++c) {
    const int s = int(c % slots);
    const int64_t cnt = (items - first < per_chunk) ? items - first : per_chunk;
    double* d_in = static_cast<double*>(ctx->stage_in[s].ptr);
    double* d_out = static_cast<double*>(ctx->stage_out[s].ptr);
    cudaError_t e;
    // slot reuse: the kernels that read d_in[s] and the D2H that read d_out[s] two chunks ago
    if (c >= slots) {
      if ((e = cudaStreamWaitEvent(ctx->h2d_stream, ctx->ev_done[s], 0)) != cudaSuccess) goto cuda_fail;
      if ((e = cudaStreamWaitEvent(ctx->own_stream, ctx->ev_out[s], 0)) != cudaSuccess) goto cuda_fail;
    }
    if ((e = cudaMemcpyAsync(d_in, in + first * item_elems, size_t(cnt) * item_bytes, cudaMemcpyHostToDevice,
                             ctx->h2d_stream)) != cudaSuccess) goto cuda_fail;
    if ((e = cudaEventRecord(ctx->ev_in[s], ctx->h2d_stream)) != cudaSuccess) goto cuda_fail;
    if ((e = cudaStreamWaitEvent(ctx->own_stream, ctx->ev_in[s], 0)) != cudaSuccess) goto cuda_fail;
    status = run(d_in, d_out, cnt);
    if (status) break;
    if ((e = cudaEventRecord(ctx->ev_done[s], ctx->own_stream)) != cudaSuccess) goto cuda_fail;
    if ((e = cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_done[s], 0)) != cudaSuccess) goto cuda_fail;
    if ((e = cudaMemcpyAsync(out + first * item_elems, d_out, size_t(cnt) * item_bytes, cudaMemcpyDeviceToHost,
                             ctx->d2h_stream)) != cudaSuccess) goto cuda_fail;
    if ((e = cudaEventRecord(ctx->ev_out[s], ctx->d2h_stream)) != cudaSuccess) goto cuda_fail;
    continue;
  cuda_fail:
    ctx->err = std::string("staging pipeline: ") + cudaGetErrorString(e);
    status = JWC_ERR_CUDA;
  }
  cudaError_t e1 = cudaStreamSynchronize(ctx->h2d_stream);
  cudaError_t e2 = cudaStreamSynchronize(ctx->own_stream);
  cudaError_t e3 = cudaStreamSynchronize(ctx->d2h_stream);
  ctx->stream = user_stream;
  if (!status && (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)) {
    cudaError_t e = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
    ctx->err = std::string("staging pipeline: ") + cudaGetErrorString(e);
    status = JWC_ERR_CUDA;
  }
  return status;
}


// ---- device groups: one context that drives several GPUs of the box (jwc_create_multi) ---------------------
//
// `jwc_create(out, devices, ndev)` and a `jwc_fwt3d` that slab-shards internally.  One
// process, one host thread issuing asynchronous work to every GPU, peer access between all pairs
// (cudaDeviceEnablePeerAccess), events for the cross-device dependencies; no torch, no NCCL.
//   * batched 1-D / 2-D entry points: the batch is cut into contiguous blocks, one per GPU, each through that
//     GPU's own staging pipeline on its own host thread - independent signals / images, no exchange;
//   * jwc_fwt3d / jwc_wpt3d: the volume is slab-decomposed along i.  Forward: every GPU uploads its i-slab in
//     chunks of slices, runs the k and j passes on a chunk while the next one uploads, and its copy engines
//     write the chunk's rows into the j-slabs of all GPUs (cudaMemcpy2DAsync, 1 MiB rows); after the re-cut the
//     i pass runs on the j-slab and the result goes to the host with one strided copy per GPU - the download
//     IS the second re-cut.  Reverse: mirror image, axis i first (the order of ParallelTransform.reverse,
//     ParallelTransform.java:193; rounding-level different from BasicTransform.java:602-659).
struct jwc_group {
  std::vector<jwc_ctx*> dev;              // dev[0] is the owner context itself
  std::vector<cudaStream_t> copy;         // one peer-copy stream per device
  std::vector<Scratch> in, a, b, J, Y;    // per device: i-slab in, two i-slab temporaries, j-slab in / out
};

static int group_size(const jwc_ctx* ctx) { return (ctx && ctx->group) ? int(ctx->group->dev.size()) : 1; }

extern "C" int jwc_device_count(const jwc_ctx* ctx) { return ctx ? group_size(ctx) : 0; }

static void group_destroy(jwc_ctx* owner) {
  jwc_group* g = owner->group;
  owner->group = nullptr;
  for (size_t i = 0; i < g->dev.size(); ++i) {
    cudaSetDevice(g->dev[i]->device);
    cudaDeviceSynchronize();
    for (auto* v : {&g->in, &g->a, &g->b, &g->J, &g->Y}) free_scratch((*v)[i]);
    if (g->copy[i]) cudaStreamDestroy(g->copy[i]);
    if (i > 0) jwc_destroy(g->dev[i]);
  }
  delete g;
}

static int group_set_wavelet(jwc_ctx* owner, int L, const double* sDe, const double* wDe, const double* sRe,
                             const double* wRe, int wid) {
  for (size_t i = 1; i < owner->group->dev.size(); ++i) {
    int w = -1;
    int st = jwc_set_wavelet(owner->group->dev[i], L, sDe, wDe, sRe, wRe, &w);
    if (st) return fail(owner, st, jwc_last_error(owner->group->dev[i]));
    if (w != wid) return fail(owner, JWC_ERR_ARG, "device group: wavelet handles out of step");
  }
  return JWC_OK;
}

extern "C" int jwc_create_multi(jwc_ctx** out, const int* devices, int ndev) {
  if (!out) return JWC_ERR_ARG;
  *out = nullptr;
  auto bad = [&](const std::string& msg, int status) {
    std::lock_guard<std::mutex> lk(g_create_mu);
    g_create_err = msg;
    return status;
  };
  if (!devices || ndev < 1 || ndev > 8) return bad("jwc_create_multi: 1 to 8 devices", JWC_ERR_ARG);
  for (int i = 0; i < ndev; ++i)
    for (int j = 0; j < i; ++j)
      if (devices[i] == devices[j]) return bad("jwc_create_multi: duplicate device", JWC_ERR_ARG);
  jwc_group* g = new jwc_group();
  int st = JWC_OK;
  for (int i = 0; i < ndev && !st; ++i) {
    jwc_ctx* c = nullptr;
    st = jwc_create(&c, devices[i]);
    if (!st) g->dev.push_back(c);
  }
  for (int i = 0; i < ndev && !st; ++i) {
    cudaSetDevice(devices[i]);
    for (int j = 0; j < ndev && !st; ++j) {
      if (i == j) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
      if (!can) st = bad("jwc_create_multi: no peer access between the devices", JWC_ERR_CUDA);
      else {
        cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) st = bad(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e), JWC_ERR_CUDA);
      }
    }
  }
  if (st) {
    for (auto* c : g->dev) jwc_destroy(c);
    delete g;
    return st;
  }
  g->copy.assign(ndev, nullptr);
  for (int i = 0; i < ndev; ++i) {
    cudaSetDevice(devices[i]);
    cudaStreamCreateWithFlags(&g->copy[i], cudaStreamNonBlocking);
  }
  for (auto* v : {&g->in, &g->a, &g->b, &g->J, &g->Y}) v->resize(ndev);
  g->dev[0]->group = g;
  *out = g->dev[0];
  return JWC_OK;
}

// batched 1-D / 2-D on a group: contiguous blocks of items, one per GPU, each on its own host thread through
// that GPU's single-device entry point (independent items: no exchange)
template <class Call>
static int group_batch(jwc_ctx* owner, int64_t items, int64_t item_elems, const double* in, double* out, Call call) {
  jwc_group* g = owner->group;
  const int nd = int(g->dev.size());
  std::vector<int> status(nd, JWC_OK);
  std::vector<std::thread> th;
  const int64_t base = items / nd, extra = items % nd;
  int64_t first = 0;
  for (int i = 0; i < nd; ++i) {
    const int64_t cnt = base + (i < extra ? 1 : 0);
    if (cnt > 0 && i > 0) {
      jwc_ctx* c = g->dev[i];
      const double* pi = in + first * item_elems;
      double* po = out + first * item_elems;
      th.emplace_back([&status, i, c, pi, po, cnt, &call] { status[i] = call(c, pi, po, cnt); });
    }
    first += cnt;
  }
  // the owner's own block runs on the calling thread, which already holds the owner's (per-thread recursive) lock
  if (base + (extra > 0 ? 1 : 0) > 0) status[0] = call(g->dev[0], in, out, base + (extra > 0 ? 1 : 0));
  for (auto& t : th) t.join();
  for (int i = 0; i < nd; ++i)
    if (status[i]) return fail(owner, status[i], jwc_last_error(g->dev[i]));
  return JWC_OK;
}

#define JWC_G(call)                                                                  \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      owner->err = std::string("device group: " #call ": ") + cudaGetErrorString(e__); \
      return JWC_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

static int group_axis(jwc_ctx* owner, jwc_ctx* c, int wid, int kind, int dir, const double* in, double* out,
                      int64_t outer, int n, int64_t inner, int level) {
  cudaError_t e = cudaSetDevice(c->device);
  if (e == cudaSuccess) e = run_axis(c, c->wavelets[wid], kind, dir, in, out, outer, n, inner, level);
  if (e != cudaSuccess) {
    owner->err = std::string("device group: axis transform: ") + cudaGetErrorString(e);
    return JWC_ERR_CUDA;
  }
  return JWC_OK;
}

// the slab-decomposed volume; arguments are validated, P and Q are multiples of the group size
static int group_3d(jwc_ctx* owner, int wid, int kind, int dir, const double* in, double* out, int P, int Q, int R,
                    int lvlP, int lvlQ, int lvlR) {
  jwc_group* g = owner->group;
  const int nd = int(g->dev.size());
  const int64_t p = P / nd, q = Q / nd;
  int C = 4;
  while (C > 1 && p % C) C >>= 1;
  const int64_t S = p / C;
  const size_t islab = size_t(p) * Q * R * sizeof(double), jslab = size_t(P) * q * R * sizeof(double);
  const size_t rowJ = size_t(q) * R * sizeof(double), rowI = size_t(Q) * R * sizeof(double);
  std::vector<cudaEvent_t> evs;
  auto new_event = [&](cudaEvent_t* e) {
    cudaError_t r = cudaEventCreateWithFlags(e, cudaEventDisableTiming);
    if (r == cudaSuccess) evs.push_back(*e);
    return r;
  };
  struct Cleanup {
    std::vector<cudaEvent_t>& v;
    ~Cleanup() { for (auto e : v) cudaEventDestroy(e); }
  } cleanup{evs};
  auto ptr = [](Scratch& s) { return static_cast<double*>(s.ptr); };
  for (int i = 0; i < nd; ++i) {
    jwc_ctx* c = g->dev[i];
    JWC_G(cudaSetDevice(c->device));
    int st;
    if ((st = ensure(c, g->in[i], islab)) || (st = ensure(c, g->a[i], islab)) || (st = ensure(c, g->b[i], islab)) ||
        (st = ensure(c, g->J[i], jslab)) || (st = ensure(c, g->Y[i], jslab)))
      return fail(owner, st, jwc_last_error(c));
  }
  std::vector<std::vector<cudaEvent_t>> copied(nd, std::vector<cudaEvent_t>(C));
  if (dir == JWC_FORWARD) {
    // BasicTransform.java:509-566 (F5): axis k gets lvlQ, axis j gets lvlP on every slice, then axis i gets lvlR
    for (int c = 0; c < C; ++c) {
      for (int i = 0; i < nd; ++i) {
        jwc_ctx* x = g->dev[i];
        JWC_G(cudaSetDevice(x->device));
        const int64_t off = c * S * Q * R;  // chunk c of my i-slab
        cudaEvent_t up, done;
        JWC_G(new_event(&up));
        JWC_G(new_event(&done));
        JWC_G(new_event(&copied[i][c]));
        JWC_G(cudaMemcpyAsync(ptr(g->in[i]) + off, in + (int64_t(i) * p + c * S) * Q * R, size_t(S) * Q * R * sizeof(double),
                              cudaMemcpyHostToDevice, x->h2d_stream));
        JWC_G(cudaEventRecord(up, x->h2d_stream));
        JWC_G(cudaStreamWaitEvent(x->own_stream, up, 0));
        int st;
        if ((st = group_axis(owner, x, wid, kind, dir, ptr(g->in[i]) + off, ptr(g->a[i]) + off, S * Q, R, 1, lvlQ))) return st;
        if ((st = group_axis(owner, x, wid, kind, dir, ptr(g->a[i]) + off, ptr(g->b[i]) + off, S, Q, R, lvlP))) return st;
        JWC_G(cudaEventRecord(done, x->own_stream));
        JWC_G(cudaStreamWaitEvent(g->copy[i], done, 0));
        for (int k = 0; k < nd; ++k) {  // my own block first, then round the ring
          const int d = (i + k) % nd;
          JWC_G(cudaMemcpy2DAsync(ptr(g->J[d]) + (int64_t(i) * p + c * S) * q * R, rowJ, ptr(g->b[i]) + off + int64_t(d) * q * R,
                                  rowI, rowJ, size_t(S), cudaMemcpyDefault, g->copy[i]));
        }
        JWC_G(cudaEventRecord(copied[i][c], g->copy[i]));
      }
    }
    for (int d = 0; d < nd; ++d) {
      jwc_ctx* x = g->dev[d];
      JWC_G(cudaSetDevice(x->device));
      for (int i = 0; i < nd; ++i)
        for (int c = 0; c < C; ++c) JWC_G(cudaStreamWaitEvent(x->own_stream, copied[i][c], 0));
      int st;
      if ((st = group_axis(owner, x, wid, kind, dir, ptr(g->J[d]), ptr(g->Y[d]), 1, P, q * R, lvlR))) return st;
      cudaEvent_t done;
      JWC_G(new_event(&done));
      JWC_G(cudaEventRecord(done, x->own_stream));
      JWC_G(cudaStreamWaitEvent(x->d2h_stream, done, 0));
      // the download is the second re-cut: my j range of every (i, .) row of the host volume
      JWC_G(cudaMemcpy2DAsync(out + int64_t(d) * q * R, rowI, ptr(g->Y[d]), rowJ, rowJ, size_t(P), cudaMemcpyDeviceToHost,
                              x->d2h_stream));
    }
  } else {
    // axis i first (ParallelTransform.java:193), then every slice: columns, then rows (BasicTransform.java:611-655)
    for (int d = 0; d < nd; ++d) {
      jwc_ctx* x = g->dev[d];
      JWC_G(cudaSetDevice(x->device));
      cudaEvent_t up, done;
      JWC_G(new_event(&up));
      JWC_G(new_event(&done));
      JWC_G(cudaMemcpy2DAsync(ptr(g->J[d]), rowJ, in + int64_t(d) * q * R, rowI, rowJ, size_t(P), cudaMemcpyHostToDevice,
                              x->h2d_stream));
      JWC_G(cudaEventRecord(up, x->h2d_stream));
      JWC_G(cudaStreamWaitEvent(x->own_stream, up, 0));
      int st;
      if ((st = group_axis(owner, x, wid, kind, dir, ptr(g->J[d]), ptr(g->Y[d]), 1, P, q * R, lvlR))) return st;
      JWC_G(cudaEventRecord(done, x->own_stream));
      JWC_G(cudaStreamWaitEvent(g->copy[d], done, 0));
      for (int c = 0; c < C; ++c) {
        for (int k = 0; k < nd; ++k) {
          const int i = (d + k) % nd;
          JWC_G(cudaMemcpy2DAsync(ptr(g->in[i]) + c * S * Q * R + int64_t(d) * q * R, rowI,
                                  ptr(g->Y[d]) + (int64_t(i) * p + c * S) * q * R, rowJ, rowJ, size_t(S), cudaMemcpyDefault,
                                  g->copy[d]));
        }
        JWC_G(new_event(&copied[d][c]));
        JWC_G(cudaEventRecord(copied[d][c], g->copy[d]));
      }
    }
    for (int c = 0; c < C; ++c) {
      for (int i = 0; i < nd; ++i) {
        jwc_ctx* x = g->dev[i];
        JWC_G(cudaSetDevice(x->device));
        for (int d = 0; d < nd; ++d) JWC_G(cudaStreamWaitEvent(x->own_stream, copied[d][c], 0));
        const int64_t off = c * S * Q * R;
        int st;
        if ((st = group_axis(owner, x, wid, kind, dir, ptr(g->in[i]) + off, ptr(g->a[i]) + off, S, Q, R, lvlP))) return st;
        if ((st = group_axis(owner, x, wid, kind, dir, ptr(g->a[i]) + off, ptr(g->b[i]) + off, S * Q, R, 1, lvlQ))) return st;
        cudaEvent_t done;
        JWC_G(new_event(&done));
        JWC_G(cudaEventRecord(done, x->own_stream));
        JWC_G(cudaStreamWaitEvent(x->d2h_stream, done, 0));
        JWC_G(cudaMemcpyAsync(out + (int64_t(i) * p + c * S) * Q * R, ptr(g->b[i]) + off, size_t(S) * Q * R * sizeof(double),
                              cudaMemcpyDeviceToHost, x->d2h_stream));
      }
    }
  }
  for (int i = 0; i < nd; ++i) {
    jwc_ctx* x = g->dev[i];
    JWC_G(cudaSetDevice(x->device));
    JWC_G(cudaStreamSynchronize(x->d2h_stream));
    JWC_G(cudaStreamSynchronize(g->copy[i]));
    JWC_G(cudaStreamSynchronize(x->own_stream));
  }
  JWC_G(cudaSetDevice(owner->device));
  return JWC_OK;
}

static int t1d_host(jwc_ctx* ctx, int wid, int kind, int dir, const double* in, double* out,
                    int64_t batch, int n, int level, bool sub = false) {
  int st = check_common(ctx, wid, kind, dir, in, out);
  if (st) return st;
  JWC_LOCK(ctx);
  if ((st = check_axis(ctx, n, level))) return st;
  if (batch < 0) return fail(ctx, JWC_ERR_ARG, "negative batch");
  if (group_size(ctx) > 1 && batch >= group_size(ctx) && !sub)
    return group_batch(ctx, batch, n, in, out, [=](jwc_ctx* c, const double* pi, double* po, int64_t cnt) {
      return t1d_host(c, wid, kind, dir, pi, po, cnt, n, level, true);
    });
  return staged(ctx, in, out, batch, n, [&](const double* di, double* dout, int64_t cnt) {
    return axis_dev(ctx, wid, kind, dir, di, dout, cnt, n, 1, level);
  });
}

extern "C" int jwc_fwt1d(jwc_ctx* ctx, int wid, int dir, const double* in, double* out, int64_t batch,
                         int n, int level) {
  return t1d_host(ctx, wid, JWC_FWT, dir, in, out, batch, n, level);
}
extern "C" int jwc_wpt1d(jwc_ctx* ctx, int wid, int dir, const double* in, double* out, int64_t batch,
                         int n, int level) {
  return t1d_host(ctx, wid, JWC_WPT, dir, in, out, batch, n, level);
}

extern "C" int jwc_aed1d(jwc_ctx* ctx, int wid, int kind, int dir, const double* in, double* out, int64_t batch,
                         int n) {
  int st = check_common(ctx, wid, kind, dir, in, out);
  if (st) return st;
  JWC_LOCK(ctx);
  if (n < 1) return fail(ctx, JWC_ERR_ARG, "the supported number for decomposition is smaller than one");
  if (batch < 0) return fail(ctx, JWC_ERR_ARG, "negative batch");
  return staged(ctx, in, out, batch, n, [&](const double* di, double* dout, int64_t cnt) {
    return aed_dev(ctx, wid, kind, dir, di, dout, cnt, n);
  });
}

// host buffers: chunks of whole signals (the output is log2 n + 1 times the input, so the equal-size
// staging pipeline above does not apply); upload, decompose, download per chunk on the context's stream
extern "C" int jwc_decompose1d(jwc_ctx* ctx, int wid, int kind, const double* in, double* out, int64_t batch, int n) {
  int st = check_common(ctx, wid, kind, JWC_FORWARD, in, out);
  if (st) return st;
  JWC_LOCK(ctx);
  if ((st = check_axis(ctx, n, 0))) return st;
  if (batch < 0) return fail(ctx, JWC_ERR_ARG, "negative batch");
  if (batch == 0) return JWC_OK;
  JWC_CUDA(ctx, cudaSetDevice(ctx->device));
  // host-buffer entry points run on the context's own stream (include/jwave_cuda.h), whatever jwc_set_stream chose
  struct OwnStream {
    jwc_ctx* c; cudaStream_t user;
    explicit OwnStream(jwc_ctx* c_) : c(c_), user(c_->stream) { switch_stream(c, c->own_stream); }
    ~OwnStream() { cudaStreamSynchronize(c->own_stream); switch_stream(c, user); }
  } own(ctx);
  const int64_t sig = int64_t(exponent(n) + 1) * n;
  int64_t per_chunk = int64_t(ctx->staging_bytes / (size_t(sig) * sizeof(double)));
  if (per_chunk < 1) per_chunk = 1;
  if (per_chunk > batch) per_chunk = batch;
  if ((st = ensure(ctx, ctx->stage_in[0], size_t(per_chunk) * n * sizeof(double)))) return st;
  if ((st = ensure(ctx, ctx->stage_out[0], size_t(per_chunk) * sig * sizeof(double)))) return st;
  double* d_in = static_cast<double*>(ctx->stage_in[0].ptr);
  double* d_out = static_cast<double*>(ctx->stage_out[0].ptr);
  for (int64_t first = 0; first < batch; first += per_chunk) {
    const int64_t cnt = batch - first < per_chunk ? batch - first : per_chunk;
    JWC_CUDA(ctx, cudaMemcpyAsync(d_in, in + first * n, size_t(cnt) * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if ((st = decompose_dev(ctx, wid, kind, d_in, d_out, cnt, n))) return st;
    JWC_CUDA(ctx, cudaMemcpyAsync(out + first * sig, d_out, size_t(cnt) * sig * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    JWC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return JWC_OK;
}

static int t2d_host(jwc_ctx* ctx, int wid, int kind, int dir, const double* in, double* out,
                    int64_t batch, int rows, int cols, int lvlM, int lvlN, bool sub = false) {
  int st = check_common(ctx, wid, kind, dir, in, out);
  if (st) return st;
  JWC_LOCK(ctx);
  if (dir == JWC_FORWARD) {
    if ((st = check_axis(ctx, cols, lvlN)) || (st = check_axis(ctx, rows, lvlM))) return st;
  } else {
    if ((st = check_axis(ctx, rows, lvlM)) || (st = check_axis(ctx, cols, lvlN))) return st;
  }
  if (batch < 0) return fail(ctx, JWC_ERR_ARG, "negative batch");
  if (group_size(ctx) > 1 && batch >= group_size(ctx) && !sub)
    return group_batch(ctx, batch, int64_t(rows) * cols, in, out, [=](jwc_ctx* c, const double* pi, double* po, int64_t cnt) {
      return t2d_host(c, wid, kind, dir, pi, po, cnt, rows, cols, lvlM, lvlN, true);
    });
  return staged(ctx, in, out, batch, int64_t(rows) * cols, [&](const double* di, double* dout, int64_t cnt) {
    return t2d_dev(ctx, wid, kind, dir, di, dout, cnt, rows, cols, lvlM, lvlN);
  });
}

extern "C" int jwc_fwt2d(jwc_ctx* ctx, int wid, int dir, const double* in, double* out, int64_t batch,
                         int rows, int cols, int lvlM, int lvlN) {
  return t2d_host(ctx, wid, JWC_FWT, dir, in, out, batch, rows, cols, lvlM, lvlN);
}
extern "C" int jwc_wpt2d(jwc_ctx* ctx, int wid, int dir, const double* in, double* out, int64_t batch,
                         int rows, int cols, int lvlM, int lvlN) {
  return t2d_host(ctx, wid, JWC_WPT, dir, in, out, batch, rows, cols, lvlM, lvlN);
}

static int t3d_host(jwc_ctx* ctx, int wid, int kind, int dir, const double* in, double* out, int P,
                    int Q, int R, int lvlP, int lvlQ, int lvlR) {
  int st = check_common(ctx, wid, kind, dir, in, out);
  if (st) return st;
  JWC_LOCK(ctx);
  if (P <= 0 || Q <= 0 || R <= 0) return fail(ctx, JWC_ERR_NOT_BINARY, "given array length is not 2^p | p E N");
  const int nd = group_size(ctx);
  if (nd > 1 && P % nd == 0 && Q % nd == 0) {
    // the same checks, in the same order, as the single-device driver (t3d_dev)
    if (dir == JWC_FORWARD) {
      if ((st = check_axis(ctx, R, lvlQ)) || (st = check_axis(ctx, Q, lvlP))) return st;
    } else {
      if ((st = check_axis(ctx, Q, lvlP)) || (st = check_axis(ctx, R, lvlQ))) return st;
    }
    if ((st = check_axis(ctx, P, lvlR))) return st;
    if (overlaps(in, out, int64_t(P) * Q * R)) return fail(ctx, JWC_ERR_ARG, "in and out overlap");
    cudaStream_t user_stream = ctx->stream;
    if (user_stream != ctx->own_stream) JWC_CUDA(ctx, cudaStreamSynchronize(user_stream));
    ctx->stream = ctx->own_stream;
    st = group_3d(ctx, wid, kind, dir, in, out, P, Q, R, lvlP, lvlQ, lvlR);
    ctx->stream = user_stream;
    return st;
  }
  const size_t keep = ctx->staging_bytes;
  ctx->staging_bytes = size_t(-1) / 2;  // one volume is one item
  st = staged(ctx, in, out, 1, int64_t(P) * Q * R, [&](const double* di, double* dout, int64_t) {
    return t3d_dev(ctx, wid, kind, dir, di, dout, P, Q, R, lvlP, lvlQ, lvlR);
  });
  ctx->staging_bytes = keep;
  return st;
}

extern "C" int jwc_fwt3d(jwc_ctx* ctx, int wid, int dir, const double* in, double* out, int P, int Q,
                         int R, int lvlP, int lvlQ, int lvlR) {
  return t3d_host(ctx, wid, JWC_FWT, dir, in, out, P, Q, R, lvlP, lvlQ, lvlR);
}
extern "C" int jwc_wpt3d(jwc_ctx* ctx, int wid, int dir, const double* in, double* out, int P, int Q,
                         int R, int lvlP, int lvlQ, int lvlR) {
  return t3d_host(ctx, wid, JWC_WPT, dir, in, out, P, Q, R, lvlP, lvlQ, lvlR);
}

// ---- memory helpers ----------------------------------------------------------------------------------

extern "C" int jwc_dev_alloc(jwc_ctx* ctx, size_t bytes, void** dptr) {
  if (!ctx || !dptr) return JWC_ERR_ARG;
  JWC_LOCK(ctx);
  JWC_CUDA(ctx, cudaSetDevice(ctx->device));
  JWC_CUDA(ctx, cudaMalloc(dptr, bytes ? bytes : 1));
  return JWC_OK;
}
extern "C" int jwc_dev_free(jwc_ctx* ctx, void* dptr) {
  if (!ctx) return JWC_ERR_ARG;
  JWC_LOCK(ctx);
  JWC_CUDA(ctx, cudaSetDevice(ctx->device));
  JWC_CUDA(ctx, cudaFree(dptr));
  return JWC_OK;
}
extern "C" int jwc_h2d(jwc_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
  if (!ctx || !dst_dev || !src_host) return JWC_ERR_ARG;
  JWC_LOCK(ctx);
  JWC_CUDA(ctx, cudaSetDevice(ctx->device));
  JWC_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return JWC_OK;
}
extern "C" int jwc_d2h(jwc_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
  if (!ctx || !dst_host || !src_dev) return JWC_ERR_ARG;
  JWC_LOCK(ctx);
  JWC_CUDA(ctx, cudaSetDevice(ctx->device));
  JWC_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return JWC_OK;
}
extern "C" int jwc_host_alloc_pinned(jwc_ctx* ctx, size_t bytes, void** hptr) {
  if (!ctx || !hptr) return JWC_ERR_ARG;
  JWC_LOCK(ctx);
  JWC_CUDA(ctx, cudaSetDevice(ctx->device));
  JWC_CUDA(ctx, cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return JWC_OK;
}
extern "C" int jwc_host_free_pinned(jwc_ctx* ctx, void* hptr) {
  if (!ctx) return JWC_ERR_ARG;
  JWC_LOCK(ctx);
  JWC_CUDA(ctx, cudaFreeHost(hptr));
  return JWC_OK;
}
extern "C" int jwc_copy2d_dev(jwc_ctx* ctx, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width,
                              size_t height, void* cuda_stream) {
  if (!ctx || !dst || !src) return JWC_ERR_ARG;
  JWC_LOCK(ctx);
  if (width == 0 || height == 0) return JWC_OK;
  if (dpitch < width || spitch < width) return fail(ctx, JWC_ERR_ARG, "jwc_copy2d_dev: pitch smaller than the row");
  JWC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
  JWC_CUDA(ctx, cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDefault, st));
  return JWC_OK;
}

extern "C" int jwc_set_staging_bytes(jwc_ctx* ctx, size_t bytes) {
  if (!ctx || bytes < sizeof(double)) return JWC_ERR_ARG;
  JWC_LOCK(ctx);
  ctx->staging_bytes = bytes;
  return JWC_OK;
}
