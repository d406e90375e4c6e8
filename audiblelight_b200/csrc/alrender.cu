// alrender.cu — host planner + C-ABI (include/alrender.h) of the B200-native AudibleLight synthesis renderer.
//
// The planner turns a batch of (event, microphone) renders and (scene, microphone) mixdowns into flat device
// descriptors: uniform partitions of P samples, per-IR active source-block ranges derived from the reference's
// interpolation matrix (generate_interpolation_matrix, synthesize.py:148-181), spectrum slots in a bounded
// workspace, and per-kernel work prefixes.  Kernels are in alr_kernels.cuh.  There is no CPU compute path.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/alrender.h"
#include "alr_kernels.cuh"

using namespace alr;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t e__ = (expr);                                                                           \
    if (e__ != cudaSuccess)                                                                             \
      return fail(ALR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return ALR_OK;
    if (p) {
      CUDA_TRY(cudaFree(p));
      p = nullptr;
      cap = 0;
    }
    size_t want = bytes + bytes / 8 + 4096;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      p = nullptr;
      return fail(ALR_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    cap = want;
    return ALR_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct HostBuf {  // pinned
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return ALR_OK;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 4096;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) return fail(ALR_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
    cap = want;
    return ALR_OK;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

enum ProfCat { kCatIrFft = 0, kCatXFft, kCatCmac, kCatIfft, kCatMix, kCatOther, kNumCat };

}  // namespace

struct alr_context {
  int device = 0;
  float2* d_tw = nullptr;    // exp(-2 pi i m / P), m < P
  float2* d_zeta = nullptr;  // exp(+i pi t / 2P), t < 64 (twist seed of thread t)
  float* d_win = nullptr;  // sin^2(pi p / 256), p < 128
  DevBuf spec, desc, misc, arena;
  HostBuf stage, stage_out;
  int64_t ws_limit = (int64_t)2 << 30;
  int profiling = 0;
  alr_profile prof{};
  std::vector<cudaEvent_t> ev_pool;
  std::vector<std::pair<int, int>> ev_marks;  // (category, index of the event recorded AFTER the launch)
  size_t ev_used = 0;
};

namespace {

// ---- planning ------------------------------------------------------------------------------------------------
struct Chunk {
  int ev_begin = 0, ev_end = 0;  // internal event range
  int ir_begin = 0, ir_end = 0;
  long long hslots = 0, xslots = 0, yslots = 0;
  int n_irfft = 0, n_xfft = 0, n_cmac = 0, n_ifft = 0;
  size_t off_irfft = 0, off_ir = 0, off_xfft = 0, off_cmac = 0, off_ifft = 0;  // byte offsets of the prefix arrays
  size_t off_tile = 0, off_dry = 0;
  int n_tile = 0, n_dry = 0;
  int part_base = 0;
  bool is_dry = false;
};

struct Blob {
  std::vector<unsigned char> bytes;
  size_t add(const void* src, size_t n) {
    size_t off = (bytes.size() + 15) & ~size_t(15);
    bytes.resize(off + n);
    if (n) memcpy(bytes.data() + off, src, n);
    return off;
  }
};

constexpr int kTileSlices = 8;
constexpr int kGainSlices = 64;
constexpr int kAmbSlices = 64;

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

struct Plan {
  std::vector<EvDev> evs;      // main events [0, n_main) then dry events
  std::vector<IrDev> irs;
  std::vector<float> wband;
  std::vector<int2> lrange;
  std::vector<Chunk> chunks;
  std::vector<SceneDev> scenes;
  std::vector<AmbDev> ambs;
  std::vector<MixEv> mevs;
  int n_main = 0;
  int n_parts = 0;
  int n_amb_parts = 0;
  long long max_h = 0, max_x = 0, max_y = 0;
};

int plan_event(const alr_event& u, int idx, EvDev& d, Plan& pl) {
  if (u.n_irs != -1 && (!u.audio || u.n_audio < 1)) return fail(ALR_ERR_INVALID, "event %d: empty audio", idx);
  if (u.n_channels < 1) return fail(ALR_ERR_INVALID, "event %d: n_channels must be >= 1", idx);
  if (u.n_out < 1 || !u.spatial) return fail(ALR_ERR_INVALID, "event %d: no output buffer", idx);
  if (u.n_irs < -1) return fail(ALR_ERR_INVALID, "event %d: n_irs < -1", idx);
  if (u.n_irs == -1) {  // pre-rendered: `spatial` is an input that is only mixed
    memset(&d, 0, sizeof(d));
    d.y = u.spatial;
    d.C = u.n_channels;
    d.n_out = (int)u.n_out;
    d.gain_mode = kGainPass;
    d.parent = -1;
    d.stat = idx;
    d.ir0 = (int)pl.irs.size();
    d.blk0 = (int)pl.lrange.size();
    if (u.n_channels < 1 || u.n_out < 1 || !u.spatial) return fail(ALR_ERR_INVALID, "event %d: no spatial buffer", idx);
    return ALR_OK;
  }
  if (u.n_irs > 0 && (!u.irs || u.n_ir_samples < 1)) return fail(ALR_ERR_INVALID, "event %d: empty IRs", idx);
  if (u.n_audio > 0x3fffffff || u.n_out > 0x3fffffff || u.n_ir_samples > 0x3fffffff)
    return fail(ALR_ERR_INVALID, "event %d: signal too long", idx);
  if (u.n_irs > 1 && (!u.ir_frames || u.n_frames < 0))
    return fail(ALR_ERR_INVALID, "event %d: moving event without ir_frames / n_frames", idx);
  memset(&d, 0, sizeof(d));
  d.x = u.audio;
  d.irs = u.irs;
  d.y = u.spatial;
  d.ir_stride_c = u.ir_stride_c;
  d.ir_stride_n = u.ir_stride_n;
  d.Lx = (int)u.n_audio;
  d.Lh = (int)u.n_ir_samples;
  d.C = u.n_channels;
  d.N = u.n_irs;
  d.n_out = (int)u.n_out;
  d.moving = u.n_irs > 1;
  d.normalize = u.normalize_irs != 0;
  d.gain_mode = u.gain_mode == ALR_GAIN_NONE ? kGainNone : kGainEvent;
  d.parent = -1;
  d.stat = idx;
  d.snr = u.snr;
  d.ref_db = u.ref_db;
  d.dry_channel = u.dry_channel;
  d.dry_low = u.dry_low;
  d.dry_high = u.dry_high;
  d.mask_lo = 0;
  d.mask_hi = d.Lh;
  d.ir0 = (int)pl.irs.size();
  d.blk0 = (int)pl.lrange.size();
  if (d.N == 0) {
    d.K = 0;
    d.n_valid = std::min(d.n_out, d.Lx);
    d.B_valid = 0;
    d.B_out = 0;
    d.xlimit = 0;
    return ALR_OK;
  }
  d.K = ceil_div(d.Lh, kP);
  long long natural = d.moving ? std::max<long long>(0, (long long)u.n_frames * 128 - 256)
                               : (long long)d.Lx + d.Lh - 1;
  d.n_valid = (int)std::min<long long>(d.n_out, natural);
  d.B_valid = ceil_div(d.n_valid, kP);
  d.B_out = ceil_div(d.n_out, kP);
  d.xlimit = std::min(d.Lx, d.n_valid);
  // ---- per-IR activity
  int xslot = 0;
  if (!d.moving) {
    IrDev ir{};
    ir.xb0 = 0;
    ir.xnb = ceil_div(d.xlimit, kP);
    ir.xslot = 0;
    xslot = ir.xnb;
    pl.irs.push_back(ir);
  } else {
    const int N = d.N;
    const int32_t* fr = u.ir_frames;
    if (fr[0] < 1) return fail(ALR_ERR_INVALID, "event %d: ir_frames[0] must be >= 1", idx);
    for (int l = 1; l < N; ++l)
      if (fr[l] < fr[l - 1]) return fail(ALR_ERR_INVALID, "event %d: ir_frames must be non-decreasing", idx);
    // banded columns of the interpolation matrix, filled with the reference's assignment order
    std::vector<int> jmin(N), jlen(N), woff(N);
    for (int l = 0; l < N; ++l) {
      int lo = (l > 0 ? fr[l - 1] : fr[0]) - 1;
      int hi = (l < N - 1 ? fr[l + 1] : fr[N - 1]) - 1;
      jmin[l] = lo;
      jlen[l] = hi - lo + 1;
      woff[l] = (int)pl.wband.size();
      pl.wband.resize(pl.wband.size() + jlen[l], 0.f);
    }
    for (int ni = 0; ni + 1 < N; ++ni) {
      const int r0 = fr[ni] - 1, len = fr[ni + 1] - fr[ni] + 1;
      const double step = len > 1 ? 1.0 / (double)(len - 1) : 0.0;
      for (int i = 0; i < len; ++i) {
        double ratio = (len > 1 && i == len - 1) ? 1.0 : (double)i * step;  // np.linspace(0, 1, len)
        int r = r0 + i;
        pl.wband[woff[ni] + (r - jmin[ni])] = (float)(1.0 - ratio);
        pl.wband[woff[ni + 1] + (r - jmin[ni + 1])] = (float)ratio;
      }
    }
    for (int l = 0; l < N; ++l) {
      IrDev ir{};
      ir.woff = woff[l];
      ir.jmin = jmin[l];
      ir.nrows = jlen[l];
      // tighten to the non-zero rows; frame j covers samples [128 j - 128, 128 j + 128)
      int a = 0, b = jlen[l] - 1;
      const float* w = pl.wband.data() + woff[l];
      while (a <= b && w[a] == 0.f) ++a;
      while (b >= a && w[b] == 0.f) --b;
      ir.xslot = xslot;
      if (a <= b) {
        long long t_lo = std::max<long long>(0, 128LL * (jmin[l] + a) - 128);
        long long t_hi = std::min<long long>(d.xlimit, 128LL * (jmin[l] + b) + 128);
        if (t_hi > t_lo) {
          ir.xb0 = (int)(t_lo / kP);
          ir.xnb = (int)((t_hi - 1) / kP) - ir.xb0 + 1;
        }
      }
      xslot += ir.xnb;
      pl.irs.push_back(ir);
    }
  }
  // ---- IR range per output block (two-pointer sweep; xb0 and xb0+xnb are non-decreasing in l)
  {
    const IrDev* ir = pl.irs.data() + d.ir0;
    int lo = 0;
    for (int b = 0; b < d.B_valid; ++b) {
      while (lo < d.N && (ir[lo].xnb == 0 || ir[lo].xb0 + ir[lo].xnb - 1 + d.K - 1 < b)) ++lo;
      int hi = lo - 1;
      for (int l = lo; l < d.N && (ir[l].xnb == 0 || ir[l].xb0 <= b); ++l)
        if (ir[l].xnb > 0) hi = l;
      int2 r;
      r.x = lo;
      r.y = hi;
      pl.lrange.push_back(r);
    }
  }
  d.xslot0 = xslot;  // temporarily: number of X slots (replaced by the chunk-local base later)
  return ALR_OK;
}

inline long long ev_hslots(const EvDev& d) { return (long long)d.N * d.K * d.C; }
inline long long ev_yslots(const EvDev& d) { return (long long)d.B_valid * d.C; }

int build_plan(const alr_event* events, int64_t n_events, const alr_scene* scenes, int64_t n_scenes, int64_t ws_limit,
               Plan& pl) {
  pl.n_main = (int)n_events;
  pl.evs.resize(n_events);
  for (int64_t i = 0; i < n_events; ++i) {
    int rc = plan_event(events[i], (int)i, pl.evs[i], pl);
    if (rc) return rc;
    const alr_event& u = events[i];
    if (u.scene >= n_scenes) return fail(ALR_ERR_INVALID, "event %d: scene index %d out of range", (int)i, u.scene);
    if (u.scene >= 0 && scenes[u.scene].n_channels != u.n_channels)
      return fail(ALR_ERR_INVALID, "event %d: %d channels but scene %d has %d", (int)i, u.n_channels, u.scene,
                  scenes[u.scene].n_channels);
  }
  // dry / direct-path sub-events: a static mono render of IR (ref channel, 0) windowed around its peak
  for (int64_t i = 0; i < n_events; ++i) {
    const alr_event& u = events[i];
    if (!u.dry || u.n_irs == -1) continue;
    if (u.n_irs < 1) return fail(ALR_ERR_INVALID, "event %d: dry audio needs at least one IR", (int)i);
    if (u.dry_channel < 0 || u.dry_channel >= u.n_channels)
      return fail(ALR_ERR_INVALID, "Reference channel index out of range for IRs with %d channels", u.n_channels);
    alr_event s = u;
    s.irs = u.irs + (long long)u.dry_channel * u.ir_stride_c;
    s.n_channels = 1;
    s.n_irs = 1;
    s.ir_frames = nullptr;
    s.spatial = u.dry;
    s.n_out = u.n_audio + u.n_ir_samples - 1;
    s.dry = nullptr;
    EvDev d;
    int rc = plan_event(s, (int)i, d, pl);
    if (rc) return rc;
    d.gain_mode = kGainDry;
    d.parent = (int)i;
    d.stat = (int)i;
    d.normalize = u.normalize_irs != 0;
    pl.evs.push_back(d);
  }
  // ---- chunks: consecutive events whose spectra fit the workspace limit
  const long long slot_bytes = (long long)kP * sizeof(float2);
  const int n_all = (int)pl.evs.size();
  int e = 0;
  while (e < n_all) {
    Chunk ch;
    ch.ev_begin = e;
    ch.is_dry = e >= pl.n_main;
    ch.ir_begin = pl.evs[e].ir0;
    long long bytes = 0;
    while (e < n_all) {
      if (!ch.is_dry && e >= pl.n_main) break;  // dry events start their own chunk (they depend on main results)
      EvDev& d = pl.evs[e];
      long long h = ev_hslots(d), x = d.xslot0, y = ev_yslots(d);
      long long add = (h + x + y) * slot_bytes;
      if (e > ch.ev_begin && bytes + add > ws_limit) break;
      bytes += add;
      ch.hslots += h;
      ch.xslots += x;
      ch.yslots += y;
      ++e;
    }
    ch.ev_end = e;
    ch.ir_end = (e < n_all) ? pl.evs[e].ir0 : (int)pl.irs.size();
    pl.chunks.push_back(ch);
  }
  // ---- scenes
  pl.scenes.resize(n_scenes);
  for (int64_t s = 0; s < n_scenes; ++s) {
    const alr_scene& u = scenes[s];
    if (u.n_channels < 1 || u.n_samples < 1 || !u.mix) return fail(ALR_ERR_INVALID, "scene %d: bad shape", (int)s);
    if (u.n_ambience < 0 || (u.n_ambience > 0 && (!u.ambience || !u.ambience_ref_db)))
      return fail(ALR_ERR_INVALID, "scene %d: bad ambience list", (int)s);
    SceneDev& d = pl.scenes[s];
    memset(&d, 0, sizeof(d));
    d.mix = u.mix;
    d.C = u.n_channels;
    d.T = u.n_samples;
    d.n_amb = u.n_ambience;
    d.amb0 = (int)pl.ambs.size();
    for (int a = 0; a < u.n_ambience; ++a) {
      if (!u.ambience[a]) return fail(ALR_ERR_INVALID, "scene %d: null ambience %d", (int)s, a);
      AmbDev ad;
      memset(&ad, 0, sizeof(ad));
      ad.data = u.ambience[a];
      ad.n = (long long)u.n_channels * u.n_samples;
      ad.ref_db = u.ambience_ref_db[a];
      ad.part0 = pl.n_amb_parts;
      ad.nparts = kAmbSlices;
      pl.n_amb_parts += kAmbSlices;
      pl.ambs.push_back(ad);
    }
  }
  // events of each scene, in call order (== the reference's dict order)
  {
    std::vector<int> count(n_scenes, 0);
    for (int64_t i = 0; i < n_events; ++i)
      if (events[i].scene >= 0 && events[i].scene_end > events[i].scene_start) count[events[i].scene]++;
    int off = 0;
    for (int64_t s = 0; s < n_scenes; ++s) {
      pl.scenes[s].ev0 = off;
      pl.scenes[s].nev = 0;
      off += count[s];
    }
    pl.mevs.resize(off);
    for (int64_t i = 0; i < n_events; ++i) {
      const alr_event& u = events[i];
      if (u.scene < 0 || u.scene_end <= u.scene_start) continue;
      if (u.scene_start < 0 || u.scene_end > scenes[u.scene].n_samples)
        return fail(ALR_ERR_INVALID, "event %d: scene slice [%lld, %lld) outside the scene", (int)i,
                    (long long)u.scene_start, (long long)u.scene_end);
      SceneDev& sd = pl.scenes[u.scene];
      MixEv& m = pl.mevs[sd.ev0 + sd.nev++];
      m.y = u.spatial;
      m.start = u.scene_start;
      m.end = u.scene_end;
      m.n_out = (int)u.n_out;
      m.pad = 0;
    }
  }
  return ALR_OK;
}

// ---- profiling helpers ---------------------------------------------------------------------------------------
int prof_mark(alr_context* ctx, cudaStream_t st, int cat) {
  ctx->prof.kernel_launches += (cat >= 0 && cat != kNumCat) ? 1 : 0;
  if (!ctx->profiling) return ALR_OK;
  if (ctx->ev_used == ctx->ev_pool.size()) {
    cudaEvent_t ev;
    CUDA_TRY(cudaEventCreate(&ev));
    ctx->ev_pool.push_back(ev);
  }
  CUDA_TRY(cudaEventRecord(ctx->ev_pool[ctx->ev_used], st));
  ctx->ev_marks.push_back({cat, (int)ctx->ev_used});
  ctx->ev_used++;
  return ALR_OK;
}

#define LAUNCH_CHECK(cat)                                   \
  do {                                                      \
    CUDA_TRY(cudaGetLastError());                           \
    int rc__ = prof_mark(ctx, st, cat);                     \
    if (rc__) return rc__;                                  \
  } while (0)

int init_tables(alr_context* ctx) {
  std::vector<float2> tw(kP);
  for (int m = 0; m < kP; ++m) {
    double a = -2.0 * M_PI * (double)m / (double)kP;
    tw[m] = make_float2((float)cos(a), (float)sin(a));
  }
  std::vector<float2> zeta(kGroup);
  for (int t = 0; t < kGroup; ++t) {
    double a = M_PI * (double)t / (double)(2 * kP);
    zeta[t] = make_float2((float)cos(a), (float)sin(a));
  }
  CUDA_TRY(cudaMalloc(&ctx->d_zeta, zeta.size() * sizeof(float2)));
  CUDA_TRY(cudaMemcpy(ctx->d_zeta, zeta.data(), zeta.size() * sizeof(float2), cudaMemcpyHostToDevice));
  std::vector<float> win(128);
  for (int p = 0; p < 128; ++p) {
    double s = sin(M_PI * (double)p / 256.0);
    win[p] = (float)(s * s);
  }
  CUDA_TRY(cudaMalloc(&ctx->d_tw, tw.size() * sizeof(float2)));
  CUDA_TRY(cudaMalloc(&ctx->d_win, win.size() * sizeof(float)));
  CUDA_TRY(cudaMemcpy(ctx->d_tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(ctx->d_win, win.data(), win.size() * sizeof(float), cudaMemcpyHostToDevice));
  return ALR_OK;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

// ================================================================================================================
extern "C" {

int alr_version(void) { return ALR_VERSION; }
const char* alr_last_error(void) { return g_err.c_str(); }
int alr_partition_size(void) { return kP; }
int alr_struct_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(alr_event);
    case 1: return (int)sizeof(alr_scene);
    case 2: return (int)sizeof(alr_event_stats);
    case 3: return (int)sizeof(alr_profile);
    default: return -1;
  }
}

int alr_create(int device, alr_context** out) {
  if (!out) return fail(ALR_ERR_INVALID, "alr_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(ALR_ERR_NO_DEVICE, "no CUDA device available (%s); alrender has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  if (device < 0) CUDA_TRY(cudaGetDevice(&device));
  if (device >= n) return fail(ALR_ERR_INVALID, "alr_create: device %d out of range (%d devices)", device, n);
  CUDA_TRY(cudaSetDevice(device));
  alr_context* ctx = new alr_context();
  ctx->device = device;
  int rc = init_tables(ctx);
  if (rc) {
    delete ctx;
    return rc;
  }
  *out = ctx;
  return ALR_OK;
}

void alr_destroy(alr_context* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  if (ctx->d_tw) cudaFree(ctx->d_tw);
  if (ctx->d_win) cudaFree(ctx->d_win);
  if (ctx->d_zeta) cudaFree(ctx->d_zeta);
  ctx->spec.release();
  ctx->desc.release();
  ctx->misc.release();
  ctx->arena.release();
  ctx->stage.release();
  ctx->stage_out.release();
  for (auto ev : ctx->ev_pool) cudaEventDestroy(ev);
  delete ctx;
}

int alr_set_workspace_limit(alr_context* ctx, int64_t bytes) {
  if (!ctx || bytes < (1 << 16)) return fail(ALR_ERR_INVALID, "alr_set_workspace_limit: bad argument");
  ctx->ws_limit = bytes;
  return ALR_OK;
}

int alr_set_profiling(alr_context* ctx, int enable) {
  if (!ctx) return fail(ALR_ERR_INVALID, "alr_set_profiling: ctx is NULL");
  ctx->profiling = enable != 0;
  return ALR_OK;
}

int alr_get_profile(alr_context* ctx, alr_profile* out) {
  if (!ctx || !out) return fail(ALR_ERR_INVALID, "alr_get_profile: bad argument");
  *out = ctx->prof;
  return ALR_OK;
}

int alr_render(alr_context* ctx, const alr_event* events_in, int64_t n_events, const alr_scene* scenes_in,
               int64_t n_scenes, int mem_space, alr_event_stats* stats_out, void* stream) {
  if (!ctx) return fail(ALR_ERR_INVALID, "alr_render: ctx is NULL");
  if (n_events < 0 || n_scenes < 0 || (n_events > 0 && !events_in) || (n_scenes > 0 && !scenes_in))
    return fail(ALR_ERR_INVALID, "alr_render: bad event / scene arrays");
  if (mem_space != ALR_MEM_HOST && mem_space != ALR_MEM_DEVICE)
    return fail(ALR_ERR_INVALID, "alr_render: bad mem_space %d", mem_space);
  if (n_events > 0x3fffffff) return fail(ALR_ERR_INVALID, "alr_render: too many events");
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  memset(&ctx->prof, 0, sizeof(ctx->prof));
  ctx->ev_used = 0;
  ctx->ev_marks.clear();
  if (n_events == 0 && n_scenes == 0) return ALR_OK;

  std::vector<alr_event> events(events_in, events_in + n_events);
  std::vector<alr_scene> scenes(scenes_in, scenes_in + n_scenes);
  std::vector<std::vector<const float*>> amb_ptrs(n_scenes);
  for (int64_t s = 0; s < n_scenes; ++s) {
    if (scenes[s].n_ambience > 0 && scenes[s].ambience) {
      amb_ptrs[s].assign(scenes[s].ambience, scenes[s].ambience + scenes[s].n_ambience);
      scenes[s].ambience = amb_ptrs[s].data();
    }
  }

  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
  CUDA_TRY(cudaEventCreate(&ev_t0));
  CUDA_TRY(cudaEventCreate(&ev_t1));
  struct EvGuard {
    cudaEvent_t a, b;
    ~EvGuard() {
      cudaEventDestroy(a);
      cudaEventDestroy(b);
    }
  } ev_guard{ev_t0, ev_t1};
  CUDA_TRY(cudaEventRecord(ev_t0, st));

  // ---- host mode: stage every buffer on the device -------------------------------------------------------------
  struct OutCopy {
    void* host;
    const void* dev;
    size_t bytes;
  };
  std::vector<OutCopy> out_copies;
  if (mem_space == ALR_MEM_HOST) {
    struct InCopy {
      const float* host;
      size_t off;
      size_t bytes;
      int kind;  // 0 linear, 1 IR block
      int ev;
    };
    std::vector<InCopy> in_copies;
    std::unordered_map<const void*, size_t> seen;  // host pointer -> arena offset
    size_t total = 0;
    auto reserve = [&](size_t bytes) {
      size_t off = total;
      total = align_up(total + bytes, 256);
      return off;
    };
    std::vector<size_t> off_audio(n_events), off_irs(n_events), off_sp(n_events), off_dry(n_events);
    for (int64_t i = 0; i < n_events; ++i) {
      alr_event& u = events[i];
      if (!u.spatial || u.n_out < 1 || u.n_channels < 1)
        return fail(ALR_ERR_INVALID, "event %d: missing buffers", (int)i);
      if (u.n_irs == -1) {
        size_t bytes = (size_t)u.n_channels * u.n_out * sizeof(float);
        off_sp[i] = reserve(bytes);
        in_copies.push_back({u.spatial, off_sp[i], bytes, 0, (int)i});
        continue;
      }
      if (!u.audio || u.n_audio < 1) return fail(ALR_ERR_INVALID, "event %d: missing buffers", (int)i);
      auto it = seen.find(u.audio);
      if (it == seen.end()) {
        size_t off = reserve(u.n_audio * sizeof(float));
        seen[u.audio] = off;
        in_copies.push_back({u.audio, off, (size_t)u.n_audio * sizeof(float), 0, (int)i});
        off_audio[i] = off;
      } else {
        off_audio[i] = it->second;
      }
      if (u.n_irs > 0) {
        if (!u.irs || u.n_ir_samples < 1) return fail(ALR_ERR_INVALID, "event %d: empty IRs", (int)i);
        auto it2 = seen.find(u.irs);
        if (it2 == seen.end()) {
          size_t bytes = (size_t)u.n_channels * u.n_irs * u.n_ir_samples * sizeof(float);
          size_t off = reserve(bytes);
          seen[u.irs] = off;
          in_copies.push_back({u.irs, off, bytes, 1, (int)i});
          off_irs[i] = off;
        } else {
          off_irs[i] = it2->second;
        }
      }
      off_sp[i] = reserve((size_t)u.n_channels * u.n_out * sizeof(float));
      if (u.dry) off_dry[i] = reserve((size_t)(u.n_audio + u.n_ir_samples - 1) * sizeof(float));
    }
    std::vector<size_t> off_mix(n_scenes);
    std::vector<std::vector<size_t>> off_amb(n_scenes);
    for (int64_t s = 0; s < n_scenes; ++s) {
      alr_scene& u = scenes[s];
      if (u.n_channels < 1 || u.n_samples < 1 || !u.mix) return fail(ALR_ERR_INVALID, "scene %d: bad shape", (int)s);
      size_t bytes = (size_t)u.n_channels * u.n_samples * sizeof(float);
      for (int a = 0; a < u.n_ambience; ++a) {
        const float* p = amb_ptrs[s][a];
        if (!p) return fail(ALR_ERR_INVALID, "scene %d: null ambience", (int)s);
        auto it = seen.find(p);
        size_t off;
        if (it == seen.end()) {
          off = reserve(bytes);
          seen[p] = off;
          in_copies.push_back({p, off, bytes, 0, -1});
        } else {
          off = it->second;
        }
        off_amb[s].push_back(off);
      }
      off_mix[s] = reserve(bytes);
    }
    int rc = ctx->arena.ensure(total);
    if (rc) return rc;
    char* base = (char*)ctx->arena.p;
    for (const InCopy& c : in_copies) {
      if (c.kind == 0) {
        CUDA_TRY(cudaMemcpyAsync(base + c.off, c.host, c.bytes, cudaMemcpyHostToDevice, st));
      } else {
        const alr_event& u = events_in[c.ev];
        const size_t row = (size_t)u.n_ir_samples * sizeof(float);
        if (u.ir_stride_n == u.n_ir_samples) {
          CUDA_TRY(cudaMemcpy2DAsync(base + c.off, row * u.n_irs, u.irs, (size_t)u.ir_stride_c * sizeof(float),
                                     row * u.n_irs, u.n_channels, cudaMemcpyHostToDevice, st));
        } else {
          for (int ch = 0; ch < u.n_channels; ++ch)
            CUDA_TRY(cudaMemcpy2DAsync(base + c.off + (size_t)ch * u.n_irs * row, row,
                                       u.irs + (long long)ch * u.ir_stride_c, (size_t)u.ir_stride_n * sizeof(float),
                                       row, u.n_irs, cudaMemcpyHostToDevice, st));
        }
      }
      ctx->prof.h2d_bytes += (int64_t)c.bytes;
    }
    for (int64_t i = 0; i < n_events; ++i) {
      alr_event& u = events[i];
      if (u.n_irs == -1) {
        u.spatial = (float*)(base + off_sp[i]);
        continue;
      }
      u.audio = (const float*)(base + off_audio[i]);
      if (u.n_irs > 0) {
        u.irs = (const float*)(base + off_irs[i]);
        u.ir_stride_n = u.n_ir_samples;
        u.ir_stride_c = (int64_t)u.n_irs * u.n_ir_samples;
      }
      size_t sp_bytes = (size_t)u.n_channels * u.n_out * sizeof(float);
      out_copies.push_back({u.spatial, base + off_sp[i], sp_bytes});
      u.spatial = (float*)(base + off_sp[i]);
      if (u.dry) {
        out_copies.push_back({u.dry, base + off_dry[i], (size_t)(u.n_audio + u.n_ir_samples - 1) * sizeof(float)});
        u.dry = (float*)(base + off_dry[i]);
      }
    }
    for (int64_t s = 0; s < n_scenes; ++s) {
      alr_scene& u = scenes[s];
      for (int a = 0; a < u.n_ambience; ++a) amb_ptrs[s][a] = (const float*)(base + off_amb[s][a]);
      out_copies.push_back({u.mix, base + off_mix[s], (size_t)u.n_channels * u.n_samples * sizeof(float)});
      u.mix = (float*)(base + off_mix[s]);
    }
  }

  // ---- plan -------------------------------------------------------------------------------------------------------
  Plan pl;
  {
    int rc = build_plan(events.data(), n_events, scenes.data(), n_scenes, ctx->ws_limit, pl);
    if (rc) return rc;
  }
  const int n_all = (int)pl.evs.size();
  Blob blob;
  // per chunk: slot bases, prefixes, part ranges
  int part_total = 0;
  for (Chunk& ch : pl.chunks) {
    const int ne = ch.ev_end - ch.ev_begin;
    std::vector<int> p_irfft(ne + 1, 0), p_ir(ne + 1, 0), p_xfft(ne + 1, 0), p_cmac(ne + 1, 0), p_ifft(ne + 1, 0);
    std::vector<int> tiles, drys;
    long long h = 0, x = 0, y = 0;
    ch.part_base = part_total;
    for (int i = 0; i < ne; ++i) {
      EvDev& d = pl.evs[ch.ev_begin + i];
      const long long nx = d.xslot0;  // X slot count stored by plan_event
      d.hslot0 = h;
      d.xslot0 = x;
      d.yslot0 = y;
      h += ev_hslots(d);
      x += nx;
      y += ev_yslots(d);
      const int ncg = (d.C + kChanGroup - 1) / kChanGroup;
      const long long n_irfft = ev_hslots(d);
      const long long n_cmac = (long long)ceil_div(d.B_valid, kG) * ncg * kBinCtas;
      const long long n_ifft = d.N > 0 ? (long long)ncg * ceil_div(d.B_out, kRun) : 0;
      if (p_irfft[i] + n_irfft > 0x7ffffff0LL || p_cmac[i] + n_cmac > 0x7ffffff0LL || p_xfft[i] + nx > 0x7ffffff0LL)
        return fail(ALR_ERR_INVALID, "chunk too large for 32-bit task indices; lower the workspace limit");
      p_irfft[i + 1] = p_irfft[i] + (int)n_irfft;
      p_ir[i + 1] = p_ir[i] + d.N;
      p_xfft[i + 1] = p_xfft[i] + (int)nx;
      p_cmac[i + 1] = p_cmac[i] + (int)n_cmac;
      p_ifft[i + 1] = p_ifft[i] + (int)n_ifft;
      if (d.N > 0) {
        d.part0 = ch.part_base + p_ifft[i];
        d.nparts = (int)n_ifft;
      }
      if (d.N == 0 && d.gain_mode != kGainPass) tiles.push_back(ch.ev_begin + i);
      if (d.gain_mode == kGainDry) drys.push_back(ch.ev_begin + i);
    }
    part_total += p_ifft[ne];
    for (int ti : tiles) {
      pl.evs[ti].part0 = part_total;
      pl.evs[ti].nparts = kTileSlices;
      part_total += kTileSlices;
    }
    ch.n_irfft = p_irfft[ne];
    ch.n_xfft = p_xfft[ne];
    ch.n_cmac = p_cmac[ne];
    ch.n_ifft = p_ifft[ne];
    ch.n_tile = (int)tiles.size();
    ch.n_dry = (int)drys.size();
    ch.off_irfft = blob.add(p_irfft.data(), p_irfft.size() * sizeof(int));
    ch.off_ir = blob.add(p_ir.data(), p_ir.size() * sizeof(int));
    ch.off_xfft = blob.add(p_xfft.data(), p_xfft.size() * sizeof(int));
    ch.off_cmac = blob.add(p_cmac.data(), p_cmac.size() * sizeof(int));
    ch.off_ifft = blob.add(p_ifft.data(), p_ifft.size() * sizeof(int));
    ch.off_tile = blob.add(tiles.data(), tiles.size() * sizeof(int));
    ch.off_dry = blob.add(drys.data(), drys.size() * sizeof(int));
    pl.max_h = std::max(pl.max_h, h);
    pl.max_x = std::max(pl.max_x, x);
    pl.max_y = std::max(pl.max_y, y);
  }
  pl.n_parts = part_total;
  // ambience pointers were patched in `scenes` (host mode) — refresh AmbDev.data
  {
    size_t k = 0;
    for (int64_t s = 0; s < n_scenes; ++s)
      for (int a = 0; a < scenes[s].n_ambience; ++a) pl.ambs[k++].data = amb_ptrs[s][a];
  }
  const size_t off_evs = blob.add(pl.evs.data(), pl.evs.size() * sizeof(EvDev));
  const size_t off_irs = blob.add(pl.irs.data(), pl.irs.size() * sizeof(IrDev));
  const size_t off_wband = blob.add(pl.wband.data(), pl.wband.size() * sizeof(float));
  const size_t off_lrange = blob.add(pl.lrange.data(), pl.lrange.size() * sizeof(int2));
  const size_t off_scenes = blob.add(pl.scenes.data(), pl.scenes.size() * sizeof(SceneDev));
  const size_t off_ambs = blob.add(pl.ambs.data(), pl.ambs.size() * sizeof(AmbDev));
  const size_t off_mevs = blob.add(pl.mevs.data(), pl.mevs.size() * sizeof(MixEv));

  {
    int rc = ctx->stage.ensure(blob.bytes.size());
    if (rc) return rc;
    rc = ctx->desc.ensure(blob.bytes.size());
    if (rc) return rc;
  }
  memcpy(ctx->stage.p, blob.bytes.data(), blob.bytes.size());
  CUDA_TRY(cudaMemcpyAsync(ctx->desc.p, ctx->stage.p, blob.bytes.size(), cudaMemcpyHostToDevice, st));
  ctx->prof.h2d_bytes += (int64_t)blob.bytes.size();
  char* dbase = (char*)ctx->desc.p;
  EvDev* d_evs = (EvDev*)(dbase + off_evs);
  IrDev* d_irs = (IrDev*)(dbase + off_irs);
  float* d_wband = (float*)(dbase + off_wband);
  int2* d_lrange = (int2*)(dbase + off_lrange);
  SceneDev* d_scenes = (SceneDev*)(dbase + off_scenes);
  AmbDev* d_ambs = (AmbDev*)(dbase + off_ambs);
  MixEv* d_mevs = (MixEv*)(dbase + off_mevs);

  // ---- workspaces ---------------------------------------------------------------------------------------------------
  const size_t slot_bytes = (size_t)kP * sizeof(float2);
  const size_t spec_bytes = (size_t)(pl.max_h + pl.max_x + pl.max_y) * slot_bytes;
  {
    int rc = ctx->spec.ensure(std::max<size_t>(spec_bytes, 16));
    if (rc) return rc;
  }
  float2* d_hspec = (float2*)ctx->spec.p;
  float2* d_xspec = d_hspec + (size_t)pl.max_h * kP;
  float2* d_yspec = d_xspec + (size_t)pl.max_x * kP;
  size_t m_off = 0;
  auto m_take = [&](size_t bytes) {
    size_t o = m_off;
    m_off = align_up(m_off + bytes, 256);
    return o;
  };
  const size_t mo_irscale = m_take(std::max<size_t>(pl.irs.size(), 1) * sizeof(float));
  const size_t mo_hen = m_take(std::max<size_t>((size_t)pl.max_h, 1) * sizeof(float));
  const size_t mo_parts = m_take(std::max<size_t>(pl.n_parts, 1) * sizeof(float2));
  const size_t mo_gain = m_take(std::max<size_t>(n_all, 1) * sizeof(float));
  const size_t mo_stats = m_take(std::max<size_t>(n_events, 1) * sizeof(EvStat));
  const size_t mo_amb = m_take(std::max<size_t>(pl.n_amb_parts, 1) * sizeof(float));
  {
    int rc = ctx->misc.ensure(m_off);
    if (rc) return rc;
  }
  char* mbase = (char*)ctx->misc.p;
  float* d_irscale = (float*)(mbase + mo_irscale);
  float* d_hen = (float*)(mbase + mo_hen);
  float2* d_parts = (float2*)(mbase + mo_parts);
  float* d_gain = (float*)(mbase + mo_gain);
  EvStat* d_stats = (EvStat*)(mbase + mo_stats);
  float* d_ambparts = (float*)(mbase + mo_amb);
  CUDA_TRY(cudaMemsetAsync(d_stats, 0, std::max<size_t>(n_events, 1) * sizeof(EvStat), st));
  ctx->prof.workspace_bytes = (int64_t)(ctx->spec.cap + ctx->misc.cap + ctx->desc.cap + ctx->arena.cap);
  ctx->prof.n_chunks = (int64_t)pl.chunks.size();
  {
    int rc = prof_mark(ctx, st, kNumCat);  // start marker
    if (rc) return rc;
  }

  // ---- launches -----------------------------------------------------------------------------------------------------
  for (const Chunk& ch : pl.chunks) {
    const int ne = ch.ev_end - ch.ev_begin;
    const EvDev* c_evs = d_evs + ch.ev_begin;
    const int* p_irfft = (const int*)(dbase + ch.off_irfft);
    const int* p_ir = (const int*)(dbase + ch.off_ir);
    const int* p_xfft = (const int*)(dbase + ch.off_xfft);
    const int* p_cmac = (const int*)(dbase + ch.off_cmac);
    const int* p_ifft = (const int*)(dbase + ch.off_ifft);
    if (ch.n_dry > 0) {
      k_dry_window<<<ch.n_dry, 256, 0, st>>>(d_evs, (const int*)(dbase + ch.off_dry), d_stats);
      LAUNCH_CHECK(kCatOther);
    }
    if (ch.n_irfft > 0) {
      k_ir_fft<<<ceil_div(ch.n_irfft, kGroupsPerCta), kCtaThreads, 0, st>>>(c_evs, ne, p_irfft, ch.n_irfft, ctx->d_tw,
                                                                         ctx->d_zeta, d_hspec, d_hen);
      LAUNCH_CHECK(kCatIrFft);
      const int n_irs = ch.ir_end - ch.ir_begin;
      k_ir_scale<<<ceil_div((long long)n_irs * 32, 128), 128, 0, st>>>(c_evs, ne, p_ir, n_irs, d_hen, d_irscale, d_stats);
      LAUNCH_CHECK(kCatOther);
    }
    if (ch.n_xfft > 0) {
      k_x_fft<<<ceil_div(ch.n_xfft, kGroupsPerCta), kCtaThreads, 0, st>>>(c_evs, ne, p_xfft, ch.n_xfft, d_irs, d_wband,
                                                                        d_irscale, ctx->d_tw, ctx->d_zeta, ctx->d_win, d_xspec);
      LAUNCH_CHECK(kCatXFft);
    }
    if (ch.n_cmac > 0) {
      k_cmac<<<ch.n_cmac, kCtaThreads, 0, st>>>(c_evs, ne, p_cmac, d_irs, d_lrange, d_xspec, d_hspec, d_yspec);
      LAUNCH_CHECK(kCatCmac);
    }
    if (ch.n_ifft > 0) {
      k_ifft_ola<<<ch.n_ifft, kCtaThreads, 0, st>>>(c_evs, ne, p_ifft, ctx->d_tw, ctx->d_zeta, d_yspec, d_parts, ch.part_base);
      LAUNCH_CHECK(kCatIfft);
    }
    if (ch.n_tile > 0) {
      k_tile<<<dim3(kTileSlices, ch.n_tile), 256, 0, st>>>(d_evs, (const int*)(dbase + ch.off_tile), kTileSlices, d_parts);
      LAUNCH_CHECK(kCatOther);
    }
    k_event_gain<<<ceil_div((long long)ne * 32, 128), 128, 0, st>>>(d_evs, ch.ev_begin, ch.ev_end, d_parts, d_stats, d_gain);
    LAUNCH_CHECK(kCatMix);
    for (int e0 = 0; e0 < ne; e0 += 32768) {
      const int cnt = std::min(32768, ne - e0);
      k_apply_gain<<<dim3(kGainSlices, cnt), 256, 0, st>>>(d_evs, ch.ev_begin + e0, d_gain);
      LAUNCH_CHECK(kCatMix);
    }
  }
  if (n_scenes > 0) {
    if (!pl.ambs.empty()) {
      for (size_t a0 = 0; a0 < pl.ambs.size(); a0 += 32768) {
        const int cnt = (int)std::min<size_t>(32768, pl.ambs.size() - a0);
        k_amb_partial<<<dim3(kAmbSlices, cnt), 256, 0, st>>>(d_ambs + a0, d_ambparts);
        LAUNCH_CHECK(kCatMix);
      }
      k_amb_final<<<ceil_div((long long)pl.ambs.size() * 32, 128), 128, 0, st>>>(d_ambs, (int)pl.ambs.size(), d_ambparts);
      LAUNCH_CHECK(kCatMix);
    }
    long long max_t = 0;
    for (const SceneDev& s : pl.scenes) max_t = std::max(max_t, s.T);
    for (int64_t s0 = 0; s0 < n_scenes; s0 += 32768) {
      const int cnt = (int)std::min<int64_t>(32768, n_scenes - s0);
      k_mix<<<dim3(ceil_div(max_t, 1024), cnt), 256, 0, st>>>(d_scenes + s0, d_ambs, d_mevs);
      LAUNCH_CHECK(kCatMix);
    }
  }

  // ---- results back ---------------------------------------------------------------------------------------------------
  {
    int rc = ctx->stage_out.ensure(std::max<size_t>(n_events, 1) * sizeof(EvStat));
    if (rc) return rc;
  }
  if (n_events > 0) {
    CUDA_TRY(cudaMemcpyAsync(ctx->stage_out.p, d_stats, n_events * sizeof(EvStat), cudaMemcpyDeviceToHost, st));
    ctx->prof.d2h_bytes += (int64_t)(n_events * sizeof(EvStat));
  }
  for (const OutCopy& c : out_copies) {
    CUDA_TRY(cudaMemcpyAsync(c.host, c.dev, c.bytes, cudaMemcpyDeviceToHost, st));
    ctx->prof.d2h_bytes += (int64_t)c.bytes;
  }
  CUDA_TRY(cudaEventRecord(ev_t1, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, ev_t0, ev_t1));
  ctx->prof.ms_total = ms;
  if (ctx->profiling && ctx->ev_marks.size() > 1) {
    double acc[kNumCat] = {0};
    for (size_t i = 1; i < ctx->ev_marks.size(); ++i) {
      float dt = 0.f;
      CUDA_TRY(cudaEventElapsedTime(&dt, ctx->ev_pool[ctx->ev_marks[i - 1].second], ctx->ev_pool[ctx->ev_marks[i].second]));
      int cat = ctx->ev_marks[i].first;
      if (cat >= 0 && cat < kNumCat) acc[cat] += dt;
    }
    ctx->prof.ms_ir_fft = acc[kCatIrFft];
    ctx->prof.ms_x_fft = acc[kCatXFft];
    ctx->prof.ms_cmac = acc[kCatCmac];
    ctx->prof.ms_ifft = acc[kCatIfft];
    ctx->prof.ms_mix = acc[kCatMix];
    ctx->prof.ms_other = acc[kCatOther];
  }
  if (stats_out) {
    const EvStat* hs = (const EvStat*)ctx->stage_out.p;
    for (int64_t i = 0; i < n_events; ++i) {
      stats_out[i].peak = hs[i].peak;
      stats_out[i].mean_abs = hs[i].mean_abs;
      stats_out[i].gain = hs[i].gain;
      stats_out[i].event_scale = hs[i].event_scale;
      stats_out[i].nonfinite = hs[i].nonfinite;
      stats_out[i].dry_peak = events_in[i].dry ? hs[i].dry_peak : -1;
    }
  }
  return ALR_OK;
}

int alr_debug_rfft(alr_context* ctx, const float* in, int64_t n_blocks, int64_t in_stride, int32_t n_valid,
                   float* spec_out, void* stream) {
  if (!ctx || !in || !spec_out || n_blocks < 1 || n_valid < 0 || n_valid > kP)
    return fail(ALR_ERR_INVALID, "alr_debug_rfft: bad argument");
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  k_debug_rfft<<<ceil_div(n_blocks, kGroupsPerCta), kCtaThreads, 0, st>>>(in, n_blocks, in_stride, n_valid, ctx->d_tw,
                                                                        ctx->d_zeta, (float2*)spec_out);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(st));
  return ALR_OK;
}

int alr_debug_irfft(alr_context* ctx, const float* spec_in, int64_t n_blocks, float* out, void* stream) {
  if (!ctx || !spec_in || !out || n_blocks < 1) return fail(ALR_ERR_INVALID, "alr_debug_irfft: bad argument");
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  k_debug_irfft<<<ceil_div(n_blocks, kGroupsPerCta), kCtaThreads, 0, st>>>((const float2*)spec_in, n_blocks, ctx->d_tw, ctx->d_zeta, out);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(st));
  return ALR_OK;
}

int alr_debug_plan(const alr_event* ev, int32_t* header, int32_t* irs, int64_t irs_cap, float* wband,
                   int64_t wband_cap, int32_t* lrange, int64_t lrange_cap) {
  if (!ev || !header) return fail(ALR_ERR_INVALID, "alr_debug_plan: bad argument");
  Plan pl;
  EvDev d;
  int rc = plan_event(*ev, 0, d, pl);
  if (rc) return rc;
  header[0] = d.K;
  header[1] = d.B_valid;
  header[2] = d.B_out;
  header[3] = d.n_valid;
  header[4] = d.xlimit;
  header[5] = (int32_t)pl.irs.size();
  header[6] = (int32_t)pl.wband.size();
  header[7] = (int32_t)pl.lrange.size();
  if ((int64_t)pl.irs.size() * 6 > irs_cap || (int64_t)pl.wband.size() > wband_cap ||
      (int64_t)pl.lrange.size() * 2 > lrange_cap)
    return fail(ALR_ERR_INVALID, "alr_debug_plan: output buffers too small");
  for (size_t i = 0; i < pl.irs.size(); ++i) {
    const IrDev& r = pl.irs[i];
    int32_t* o = irs + 6 * i;
    o[0] = r.xb0; o[1] = r.xnb; o[2] = r.xslot; o[3] = r.woff; o[4] = r.jmin; o[5] = r.nrows;
  }
  if (!pl.wband.empty()) memcpy(wband, pl.wband.data(), pl.wband.size() * sizeof(float));
  for (size_t i = 0; i < pl.lrange.size(); ++i) {
    lrange[2 * i] = pl.lrange[i].x;
    lrange[2 * i + 1] = pl.lrange[i].y;
  }
  return ALR_OK;
}

}  // extern "C"
