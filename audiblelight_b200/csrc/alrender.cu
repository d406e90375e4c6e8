// alrender.cu — host planner + C-ABI (include/alrender.h) of the B200-native AudibleLight synthesis renderer.
//
// The planner turns a batch of (event, microphone) renders and (scene, microphone) mixdowns into flat device
// descriptors: uniform partitions of P samples, per-IR active source-block ranges derived from the reference's
// interpolation matrix (generate_interpolation_matrix, synthesize.py:148-181), spectrum slots in a bounded
// workspace, and per-kernel work prefixes.  Kernels are in alr_kernels.cuh.  There is no CPU compute path.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/alrender.h"
#include "alr_kernels.cuh"
#include "alr_fused.cuh"
#include "alr_sweep.cuh"

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

enum ProfCat { kCatIrFft = 0, kCatXFft, kCatCmac, kCatCmacStatic, kCatIfft, kCatMix, kCatOther, kCatFused, kNumCat };

}  // namespace

struct alr_context {
  int device = 0;
  float2* d_tw = nullptr;    // exp(-2 pi i m / P), m < P
  float2* d_zeta = nullptr;  // exp(+i pi t / 2P), t < 64 (twist seed of thread t)
  float* d_win = nullptr;  // sin^2(pi p / 256), p < 128
  DevBuf spec, desc, misc, arena, augbuf, augdesc, ring, ambgen, visbuf;
  // persistent producer/consumer launch for moving events (alr_fused.cuh)
  int fused = 0;                          // moving events: 0 = k_ir_fft + k_cmac, 1 = k_mov_fused (alr_fused.cuh, experiment,
                                          // profiles/r02_fused_ring.txt), 2 = k_mov_sweep (alr_sweep.cuh)
  int sweep_grid = 0;                     // CTAs of k_mov_sweep (one per SM), 0: not available
  int64_t ring_bytes = (int64_t)64 << 20; // H-spectra ring (must stay L2 resident: 126 MB on B200)
  int lookahead = 2;                      // runs whose RIRs are produced ahead of the consumer tasks
  int fused_grid = 0;                     // resident CTAs of k_mov_fused (SMs x occupancy)
  int sm_clock_khz = 0;
  int mix_group = 0;                      // scenes per ambience-reduction + mixdown group (0: all at once)
  int cmac_merge = 0;                     // k_cmac and k_cmac_static CTAs interleaved in one grid (k_cmac_both)
  long long watchdog_ms = 2000;           // ALR_WATCHDOG_MS: how long a persistent kernel may wait for one dependency
  int small_rir = 1;                      // k_small_rir for RIRs of at most one partition (ALR_SMALL=0: general pipeline)
  int64_t l2_persist_bytes = 0;           // L2 set aside for persisting lines (the ring of k_mov_sweep); 0: off
  int64_t l2_window_max = 0;
  HostBuf stage, stage_out, stage_aug;
  int64_t ws_limit = (int64_t)16 << 30;  // spectra workspace bound (device inputs). Benchmark step at 4 / 8 / 16 / 32 GiB: 21.2 / 20.7 /
                                         // 20.3 / 20.5 ms (fewer, fuller launches; B200 has 180 GB). Only what a call needs is allocated.
  int profiling = 0;
  alr_profile prof{};
  std::vector<cudaEvent_t> ev_pool;
  std::vector<std::pair<int, int>> ev_marks;  // (category, index of the event recorded AFTER the launch)
  size_t ev_used = 0;
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
  // host-mode pipeline: uploads / downloads run on their own streams and overlap the kernels of other chunks
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  std::vector<cudaEvent_t> sync_pool;
  size_t sync_used = 0;
};

namespace {

// ---- planning ------------------------------------------------------------------------------------------------
// Two levels.  size_event() is a cheap pre-pass (no allocation) that validates an event and derives every size the
// launch geometry depends on; with it the whole call can be cut into workspace-bounded chunks and every buffer can
// be sized before any detailed planning.  plan_event_into() then writes the device descriptors of one event
// straight into the pinned staging blob of its chunk, so a chunk can be uploaded and launched while the host is
// already planning the next one (planning overlaps GPU execution).
constexpr int kTileSlices = 8;
constexpr int kGainSlices = 64;
constexpr int kAmbSlices = 64;

constexpr int kMaxSweepSlots = 64;  // sweeper slots of k_mov_sweep (SMs / 8 bin slices)
inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Blob {  // host-side byte blob with 16-byte aligned sections
  std::vector<unsigned char> bytes;
  size_t add(const void* src, size_t n) {
    size_t off = (bytes.size() + 15) & ~size_t(15);
    bytes.resize(off + n);
    if (n) memcpy(bytes.data() + off, src, n);
    return off;
  }
};

struct EvSize {
  int K = 0, n_valid = 0, B_valid = 0, B_out = 0, xlimit = 0;
  long long h = 0, xb = 0, y = 0;  // spectrum slots: H exact, X upper bound, Y exact
  int n_ir = 0;                    // IrDev entries
  int wband = 0;                   // upper bound of cross-fade weight floats
  int n_blk = 0;                   // lrange entries
  int n_irfft = 0, n_cmac = 0, n_cmac_static = 0, n_ifft = 0, n_parts = 0;
  bool pass = false;
  // moving events taken by k_mov_fused: H lives in the ring (z.h_ws = 0), tasks instead of k_ir_fft / k_cmac CTAs
  int fused = 0;       // 0 / 1 (k_mov_fused) / 2 (k_mov_sweep)
  long long h_ws = 0;  // H slots in the chunk workspace
  int n_ptask = 0, n_ctask = 0;
  // static renders whose effective RIR fits one partition: k_small_rir, nothing in the workspace
  bool small = false;
  int n_small = 0, small_runs = 0;
};

int size_event(const alr_event& u, int idx, EvSize& z, long long ring_slots = 0, int mover = 0, int small_mode = 0) {
  z = EvSize();
  if (u.n_channels < 1) return fail(ALR_ERR_INVALID, "event %d: n_channels must be >= 1", idx);
  if (u.n_out < 1 || !u.spatial) return fail(ALR_ERR_INVALID, "event %d: no output buffer", idx);
  if (u.n_irs < -1) return fail(ALR_ERR_INVALID, "event %d: n_irs < -1", idx);
  if (u.n_out > 0x3fffffff) return fail(ALR_ERR_INVALID, "event %d: signal too long", idx);
  if (u.n_irs == -1) {  // pre-rendered: `spatial` is an input that is only mixed
    z.pass = true;
    return ALR_OK;
  }
  if (!u.audio || u.n_audio < 1) return fail(ALR_ERR_INVALID, "event %d: empty audio", idx);
  if (u.n_irs > 0 && (!u.irs || u.n_ir_samples < 1)) return fail(ALR_ERR_INVALID, "event %d: empty IRs", idx);
  if (u.n_audio > 0x3fffffff || u.n_ir_samples > 0x3fffffff)
    return fail(ALR_ERR_INVALID, "event %d: signal too long", idx);
  if (u.n_irs > 1 && (!u.ir_frames || u.n_frames < 0))
    return fail(ALR_ERR_INVALID, "event %d: moving event without ir_frames / n_frames", idx);
  const int Lx = (int)u.n_audio, Lh = (int)u.n_ir_samples, C = u.n_channels, N = u.n_irs, n_out = (int)u.n_out;
  if (N == 0) {
    z.n_valid = std::min(n_out, Lx);
    z.n_parts = kTileSlices;
    return ALR_OK;
  }
  const bool moving = N > 1;
  z.K = ceil_div(Lh, kP);
  const long long natural = moving ? std::max<long long>(0, (long long)u.n_frames * 128 - 256) : (long long)Lx + Lh - 1;
  z.n_valid = (int)std::min<long long>(n_out, natural);
  z.B_valid = ceil_div(z.n_valid, kP);
  z.B_out = ceil_div(n_out, kP);
  z.xlimit = std::min(Lx, z.n_valid);
  z.h = (long long)N * z.K * C;
  z.y = (long long)z.B_valid * C;
  z.n_ir = N;
  z.n_blk = z.B_valid;
  if (!moving) {
    z.xb = ceil_div(z.xlimit, kP);
  } else {
    const int32_t* fr = u.ir_frames;
    if (fr[0] < 1) return fail(ALR_ERR_INVALID, "event %d: ir_frames[0] must be >= 1", idx);
    long long xb = 0, wb = 0;
    for (int l = 0; l < N; ++l) {
      if (l > 0 && fr[l] < fr[l - 1]) return fail(ALR_ERR_INVALID, "event %d: ir_frames must be non-decreasing", idx);
      const int lo = (l > 0 ? fr[l - 1] : fr[0]) - 1, hi = (l < N - 1 ? fr[l + 1] : fr[N - 1]) - 1;
      wb += hi - lo + 1;
      const long long t_lo = std::max<long long>(0, 128LL * lo - 128), t_hi = std::min<long long>(z.xlimit, 128LL * hi + 128);
      if (t_hi > t_lo) xb += (t_hi - 1) / kP - t_lo / kP + 1;
    }
    if (wb > 0x3fffffff) return fail(ALR_ERR_INVALID, "event %d: too many STFT frames", idx);
    z.xb = xb;
    z.wband = (int)wb;
    // Eligibility for the fused launch: the RIRs any single run of kGm output blocks reads (its dependency window)
    // plus the one being produced must fit the ring. Upper bound from the un-tightened bands (same expressions as
    // above); the planner's own ordering adapts to anything smaller than that.
    if (ring_slots > 0 && C <= 0x7fff && z.B_valid > 0) {
      static thread_local std::vector<int> fb, le;  // first / last active source block per RIR (-1: inactive)
      fb.assign(N, -1);
      le.assign(N, -1);
      for (int l = 0; l < N; ++l) {
        const int lo = (l > 0 ? fr[l - 1] : fr[0]) - 1, hi = (l < N - 1 ? fr[l + 1] : fr[N - 1]) - 1;
        const long long t_lo = std::max<long long>(0, 128LL * lo - 128), t_hi = std::min<long long>(z.xlimit, 128LL * hi + 128);
        if (t_hi > t_lo) {
          fb[l] = (int)(t_lo / kP);
          le[l] = (int)((t_hi - 1) / kP);
        }
      }
      int wmax = 0, lmin = 0, xnb_max = 0;
      for (int l = 0; l < N; ++l)
        if (fb[l] >= 0) xnb_max = std::max(xnb_max, le[l] - fb[l] + 1);
      for (int b0 = 0; b0 < z.B_valid; b0 += kGm) {
        const int b1 = std::min(b0 + kGm, z.B_valid) - 1;
        while (lmin < N && (fb[lmin] < 0 || le[lmin] + z.K - 1 < b0)) ++lmin;
        int lmax = lmin - 1;
        for (int l = lmin; l < N && (fb[l] < 0 || fb[l] <= b1); ++l)
          if (fb[l] >= 0) lmax = l;
        wmax = std::max(wmax, lmax - lmin + 1);
      }
      const long long ir_slots = (long long)z.K * C;
      if (mover == 1) {
        z.fused = (long long)(wmax + 2) * ir_slots <= ring_slots ? 1 : 0;
      } else if (mover == 2) {
        // k_mov_sweep: one capsule group, the RIR's output span fits the accumulator window, its source blocks fit the
        // register FIR, and the ring holds a few RIRs for every sweeper slot
        z.fused = (C <= kChanGroup && xnb_max >= 1 && xnb_max <= kSwMaxXnb && z.K + xnb_max - 1 <= kSwW &&
                   64 * ir_slots <= ring_slots) ? 2 : 0;
      }
    }
  }
  const int ncg = (C + kChanGroup - 1) / kChanGroup;
  const long long n_cmac = (long long)ceil_div(ceil_div(z.B_valid, kGm), kCmacRuns) * ncg * kBinCtas;
  const long long n_ifft = (long long)((C + kIfftCh - 1) / kIfftCh) * ceil_div(z.B_out, kRun);
  if (z.h > 0x3ffffff0LL || n_cmac * kCmacRuns > 0x3ffffff0LL || z.xb > 0x3ffffff0LL)
    return fail(ALR_ERR_INVALID, "event %d: too large for 32-bit task indices", idx);
  // small_mode: 0 off, 1 caller's event, 2 dry / direct-path sub-event (its RIR is the window of at most
  // dry_low + dry_high taps that k_dry_window selects)
  if (small_mode && !moving) {
    const long long wmax = small_mode == 2 ? std::min<long long>(Lh, (long long)std::max(u.dry_low, 0) + std::max(u.dry_high, 0)) : Lh;
    if (wmax >= 1 && wmax <= kP && (small_mode == 2 || C <= 2)) {  // short RIRs with more capsules: the source transform would be repeated per capsule (tie at C = 4, profiles/r02_small_rir.txt)
      const int bc = ceil_div(Lx + wmax - 1, kP);
      z.small = true;
      z.small_runs = ceil_div(bc, kRun);
      z.n_small = ((C + kGroupsPerCta - 1) / kGroupsPerCta) * z.small_runs;
    }
  }
  if (z.small) {
    z.h = z.xb = z.y = 0;
    z.n_parts = z.n_small;
    z.h_ws = 0;
    return ALR_OK;
  }
  z.h_ws = z.h;
  z.n_irfft = (int)z.h;
  z.n_cmac = moving ? (int)n_cmac : 0;         // generic kernel: moving events
  if (z.fused) {
    if ((long long)N * C * kSwKSplit > 0x3ffffff0LL) z.fused = 0;
  }
  if (z.fused) {
    z.n_ptask = N * C * (z.fused == 2 ? kSwKSplit : 1);
    z.n_ctask = z.fused == 1 ? ceil_div(z.B_valid, kGm) * ncg * kBinCtas : 0;  // k_mov_fused: one C-task per run
    z.h_ws = 0;
    z.n_irfft = 0;
    z.n_cmac = 0;
  }
  const long long n_cmac_s = (long long)ceil_div(ceil_div(z.B_valid, kG), kStaticRuns) * ((C + kStaticCh - 1) / kStaticCh) * kBinCtas;
  z.n_cmac_static = moving ? 0 : (int)n_cmac_s;  // regular block-FIR kernel: static events (through k_cmac's item list they
                                                 // take the same time: 5.78 vs 4.26 + 1.58 ms, profiles/r02_micro_variants.txt)
  z.n_ifft = (int)n_ifft;
  z.n_parts = (int)n_ifft;
  return ALR_OK;
}

// Detailed plan of one event. irs_out / wband_out / lr_out point at the event's slices of the chunk arrays; ir0, w0
// and blk0 are the offsets of those slices inside the chunk arrays (what the kernels index with).
// Returns the number of X slots actually used in *x_used.
void plan_event_into(const alr_event& u, int idx, const EvSize& z, EvDev& d, IrDev* irs_out, int ir0, float* wband_out,
                     int w0, int2* lr_out, int blk0, int* x_used) {
  memset(&d, 0, sizeof(d));
  d.y = u.spatial;
  d.C = u.n_channels;
  d.n_out = (int)u.n_out;
  d.parent = -1;
  d.stat = idx;
  d.ir0 = ir0;
  d.blk0 = blk0;
  *x_used = 0;
  if (z.pass) {
    d.gain_mode = kGainPass;
    return;
  }
  d.x = u.audio;
  d.irs = u.irs;
  d.ir_stride_c = u.ir_stride_c;
  d.ir_stride_n = u.ir_stride_n;
  d.Lx = (int)u.n_audio;
  d.Lh = (int)u.n_ir_samples;
  d.N = u.n_irs;
  d.moving = u.n_irs > 1;
  d.normalize = u.normalize_irs != 0;
  d.gain_mode = u.gain_mode == ALR_GAIN_NONE ? kGainNone : kGainEvent;
  d.snr = u.snr;
  d.ref_db = u.ref_db;
  d.dry_channel = u.dry_channel;
  d.dry_low = u.dry_low;
  d.dry_high = u.dry_high;
  d.mask_lo = 0;
  d.mask_hi = d.Lh;
  d.K = z.K;
  d.n_valid = z.n_valid;
  d.B_valid = z.B_valid;
  d.B_out = z.B_out;
  d.xlimit = z.xlimit;
  if (d.N == 0) return;
  int xslot = 0;
  if (!d.moving) {
    IrDev ir{};
    ir.xnb = ceil_div(d.xlimit, kP);
    xslot = ir.xnb;
    irs_out[0] = ir;
  } else {
    const int N = d.N;
    const int32_t* fr = u.ir_frames;
    // banded columns of the interpolation matrix (generate_interpolation_matrix, synthesize.py:172-179), filled
    // in the reference's assignment order; column l spans rows [fr[l-1]-1, fr[l+1]-1]
    int off = 0;
    for (int l = 0; l < N; ++l) {
      const int lo = (l > 0 ? fr[l - 1] : fr[0]) - 1, hi = (l < N - 1 ? fr[l + 1] : fr[N - 1]) - 1;
      IrDev ir{};
      ir.woff = w0 + off;
      ir.jmin = lo;
      ir.nrows = hi - lo + 1;
      irs_out[l] = ir;
      off += ir.nrows;
    }
    memset(wband_out, 0, (size_t)off * sizeof(float));
    for (int ni = 0; ni + 1 < N; ++ni) {
      const int r0 = fr[ni] - 1, len = fr[ni + 1] - fr[ni] + 1;
      const double step = len > 1 ? 1.0 / (double)(len - 1) : 0.0;
      float* wa = wband_out + (irs_out[ni].woff - w0) + (r0 - irs_out[ni].jmin);
      float* wb = wband_out + (irs_out[ni + 1].woff - w0) + (r0 - irs_out[ni + 1].jmin);
      for (int i = 0; i < len; ++i) {
        const double ratio = (len > 1 && i == len - 1) ? 1.0 : (double)i * step;  // np.linspace(0, 1, len)
        wa[i] = (float)(1.0 - ratio);
        wb[i] = (float)ratio;
      }
    }
    for (int l = 0; l < N; ++l) {
      IrDev& ir = irs_out[l];
      // tighten to the non-zero rows; frame j covers samples [128 j - 128, 128 j + 128)
      int a = 0, b = ir.nrows - 1;
      const float* w = wband_out + (ir.woff - w0);
      while (a <= b && w[a] == 0.f) ++a;
      while (b >= a && w[b] == 0.f) --b;
      ir.xslot = xslot;
      if (a <= b) {
        const long long t_lo = std::max<long long>(0, 128LL * (ir.jmin + a) - 128);
        const long long t_hi = std::min<long long>(d.xlimit, 128LL * (ir.jmin + b) + 128);
        if (t_hi > t_lo) {
          ir.xb0 = (int)(t_lo / kP);
          ir.xnb = (int)((t_hi - 1) / kP) - ir.xb0 + 1;
        }
      }
      xslot += ir.xnb;
    }
  }
  // IR range per output block (two-pointer sweep; xb0 and xb0+xnb are non-decreasing over the active IRs)
  int lo = 0;
  for (int b = 0; b < d.B_valid; ++b) {
    while (lo < d.N && (irs_out[lo].xnb == 0 || irs_out[lo].xb0 + irs_out[lo].xnb - 1 + d.K - 1 < b)) ++lo;
    int hi = lo - 1;
    for (int l = lo; l < d.N && (irs_out[l].xnb == 0 || irs_out[l].xb0 <= b); ++l)
      if (irs_out[l].xnb > 0) hi = l;
    lr_out[b].x = lo;
    lr_out[b].y = hi;
  }
  *x_used = xslot;
}

// byte layout of one chunk's descriptor blob (identical in the pinned staging buffer and on the device)
struct Chunk {
  int ev_begin = 0, ev_end = 0;  // range in the phase's event list
  long long hslots = 0, xslots = 0, yslots = 0;
  int n_ir = 0, n_wband = 0, n_blk = 0;
  size_t off_evs = 0, off_irs = 0, off_wband = 0, off_lrange = 0;
  size_t off_irfft = 0, off_ir = 0, off_xfft = 0, off_cmac = 0, off_cmacs = 0, off_ifft = 0, off_tile = 0, off_dry = 0, off_small = 0;
  size_t bytes = 0;  // blob size
  size_t base = 0;   // offset of the blob in the staging / descriptor buffers
  int part_base = 0, ir_base = 0, gain_base = 0;
  // fused launch: RIRs, tasks and (RIR, capsule) energy entries of the chunk's fused events
  int n_fo = 0, n_tasks = 0, n_sweep_ev = 0;
  long long n_ecap = 0;
  size_t off_tasks = 0, off_pop = 0, off_ncons = 0, off_prod = 0, off_slotoff = 0, off_slotjobs = 0;
};

void layout_chunk(Chunk& ch) {
  const int ne = ch.ev_end - ch.ev_begin;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o = align_up(o + bytes, 16);
    return r;
  };
  ch.off_evs = take((size_t)ne * sizeof(EvDev));
  ch.off_irs = take((size_t)ch.n_ir * sizeof(IrDev));
  ch.off_wband = take((size_t)ch.n_wband * sizeof(float));
  ch.off_lrange = take((size_t)ch.n_blk * sizeof(int2));
  ch.off_irfft = take((size_t)(ne + 1) * sizeof(int));
  ch.off_ir = take((size_t)(ne + 1) * sizeof(int));
  ch.off_xfft = take((size_t)(ne + 1) * sizeof(int));
  ch.off_cmac = take((size_t)(ne + 1) * sizeof(int));
  ch.off_cmacs = take((size_t)(ne + 1) * sizeof(int));
  ch.off_ifft = take((size_t)(ne + 1) * sizeof(int));
  ch.off_tile = take((size_t)ne * sizeof(int));
  ch.off_dry = take((size_t)ne * sizeof(int));
  ch.off_small = take((size_t)(ne + 1) * sizeof(int));
  ch.off_tasks = take((size_t)ch.n_tasks * sizeof(FusedTask));
  ch.off_pop = take((size_t)ch.n_fo * sizeof(int2));
  ch.off_ncons = take((size_t)ch.n_fo * sizeof(int2));
  ch.off_prod = take((size_t)ch.n_fo * sizeof(int));
  ch.off_slotoff = take((size_t)(kMaxSweepSlots + 1) * sizeof(int));
  ch.off_slotjobs = take((size_t)ch.n_sweep_ev * sizeof(int));
  ch.bytes = align_up(o, 256);
}

// cuts events [0, n) into chunks whose spectra fit `ws_limit`
void make_chunks(const std::vector<EvSize>& sz, int64_t ws_limit, std::vector<Chunk>& out) {
  const long long slot_bytes = (long long)kP * sizeof(float2);
  const int n = (int)sz.size();
  int e = 0;
  while (e < n) {
    Chunk ch;
    ch.ev_begin = e;
    long long bytes = 0;
    while (e < n) {
      const EvSize& z = sz[e];
      const long long add = (z.h_ws + z.xb + z.y) * slot_bytes;
      if (e > ch.ev_begin && bytes + add > ws_limit) break;
      if (e > ch.ev_begin && z.fused && (long long)ch.n_tasks + z.n_ptask + z.n_ctask > 0x3ffffff0LL) break;
      bytes += add;
      ch.hslots += z.h_ws;
      if (z.fused == 2) ch.n_sweep_ev += 1;
      if (z.fused) {
        ch.n_fo += z.n_ir;
        ch.n_tasks += z.n_ptask + z.n_ctask;
        ch.n_ecap += (long long)z.n_ptask;
      }
      ch.xslots += z.xb;
      ch.yslots += z.y;
      ch.n_ir += z.n_ir;
      ch.n_wband += z.wband;
      ch.n_blk += z.n_blk;
      ++e;
    }
    ch.ev_end = e;
    layout_chunk(ch);
    out.push_back(ch);
  }
}

// ---- task queue of the fused launch ---------------------------------------------------------------------------------
// Orders the P- and C-tasks of a chunk's fused events (see alr_fused.cuh) and assigns ring regions.
//   * events are processed one after the other, runs in ascending order; before the C-tasks of run i are queued, the
//     P-tasks of every RIR that run i + lookahead reads are queued (across event boundaries), so consumers find their
//     inputs ready and the queue never makes a CTA wait for a task that comes later;
//   * a RIR gets the next K*C slots of the ring (wrapping to 0 when it does not fit behind the head); the RIRs whose
//     region it overwrites are recorded in pop[] — its P-tasks wait until consumed[] shows all their readers done. If
//     one of those readers has not even been queued yet (ring tighter than lookahead), its run is queued first, i.e.
//     the lookahead shrinks on the spot. size_event() admitted the event only if a single run's window fits the ring,
//     so that always works.
// Returns false on an internal inconsistency (never expected).
struct FusedRun {
  int ev, run, lmin, lmax;  // lmax < lmin: the run reads no RIR
};
bool plan_fused(EvDev* evs, int ne, IrDev* irs, const int2* lr, long long ring_slots, int lookahead,
                FusedTask* tasks, int n_tasks_expected, int2* pop, int2* need, int n_fo_expected) {
  static thread_local std::vector<FusedRun> runs;
  static thread_local std::vector<int> last_run;      // per fused RIR ordinal: global index of the last run reading it
  static thread_local std::vector<int> fo_ev;         // per fused ordinal: event
  runs.clear();
  int n_fo = 0;
  for (int e = 0; e < ne; ++e) {
    EvDev& d = evs[e];
    if (!d.fused) continue;
    d.fo0 = n_fo;
    n_fo += d.N;
    const int nruns = ceil_div(d.B_valid, kGm);
    for (int r = 0; r < nruns; ++r) {
      const int b0 = r * kGm, nb = std::min(kGm, d.B_valid - b0);
      runs.push_back({e, r, lr[d.blk0 + b0].x, lr[d.blk0 + b0 + nb - 1].y});
    }
  }
  if (n_fo != n_fo_expected) return false;
  last_run.assign(n_fo, -1);
  fo_ev.resize(n_fo);
  for (int i = 0; i < n_fo; ++i) need[i] = make_int2(0, 0);
  for (int e = 0; e < ne; ++e)
    if (evs[e].fused)
      for (int l = 0; l < evs[e].N; ++l) {
        fo_ev[evs[e].fo0 + l] = e;
        need[evs[e].fo0 + l].y = evs[e].C + 1;  // ready[] value of a published RIR
      }
  for (int i = 0; i < (int)runs.size(); ++i) {
    const FusedRun& r = runs[i];
    const EvDev& d = evs[r.ev];
    const int per_run = ((d.C + kChanGroup - 1) / kChanGroup) * kBinCtas;
    for (int l = r.lmin; l <= r.lmax; ++l) {
      need[d.fo0 + l].x += per_run;
      last_run[d.fo0 + l] = i;
    }
  }
  int nt = 0;
  int c_emitted = 0;          // runs [0, c_emitted) have their C-tasks queued
  int p_emitted = 0;          // fused ordinals [0, p_emitted) have their P-tasks queued
  long long head = 0;         // next free ring slot
  int live_lo = 0;            // ordinals [live_lo, p_emitted) occupy ring regions
  static thread_local std::vector<long long> reg_lo;  // ring region start per ordinal
  reg_lo.resize(n_fo);
  bool ok = true;
  auto emit_c = [&](int i) {
    const FusedRun& r = runs[i];
    const EvDev& d = evs[r.ev];
    const int per_run = ((d.C + kChanGroup - 1) / kChanGroup) * kBinCtas;
    for (int s = 0; s < per_run; ++s) tasks[nt++] = FusedTask{kTaskC, r.ev, r.run, s};
  };
  auto emit_p = [&](int fo) {  // queue the P-tasks of ordinal fo (== p_emitted)
    const int e = fo_ev[fo];
    const EvDev& d = evs[e];
    const int l = fo - d.fo0;
    const long long size = (long long)d.K * d.C;
    const long long old_head = head;
    const bool wrapped = head + size > ring_slots;
    if (wrapped) head = 0;
    const long long lo = head, hi = head + size;
    // Previous occupants overlapping [lo, hi): a prefix of the live list (the ring is a FIFO). On a wrap the regions left
    // behind in the abandoned tail [old_head, ring_slots) are the OLDEST live entries; they are released first although
    // nothing overwrites them yet, otherwise they would sit in front of the entries at the start of the ring that the
    // new region does overwrite (RIR sizes differ between events).
    int pop_lo = live_lo;
    if (live_lo > 0 && !wrapped) {
      // the last region released by the PREVIOUS allocation may straddle `lo` (RIR sizes differ between events): this
      // producer overwrites its upper part and has to wait for its readers as well — the previous allocation's P-tasks
      // do wait for them, but nothing orders those P-tasks before ours
      const EvDev& o = evs[fo_ev[live_lo - 1]];
      const long long olo = reg_lo[live_lo - 1], ohi = olo + (long long)o.K * o.C;
      if (olo < hi && ohi > lo) pop_lo = live_lo - 1;
    }
    while (live_lo < fo) {
      const EvDev& o = evs[fo_ev[live_lo]];
      const long long olo = reg_lo[live_lo], ohi = olo + (long long)o.K * o.C;
      const bool in_tail = wrapped && olo >= old_head;
      if (!in_tail && (ohi <= lo || olo >= hi)) break;
      // its readers must be in the queue before this producer
      while (c_emitted <= last_run[live_lo]) {
        if (runs[c_emitted].lmax >= 0 && evs[runs[c_emitted].ev].fo0 + runs[c_emitted].lmax >= fo) ok = false;
        emit_c(c_emitted++);
      }
      ++live_lo;
    }
    reg_lo[fo] = lo;
    head = hi;
    irs[d.ir0 + l].hring = (int)lo;
    pop[fo] = make_int2(pop_lo, live_lo);
    for (int c = 0; c < d.C; ++c) tasks[nt++] = FusedTask{kTaskP, e, l, c};
    p_emitted = fo + 1;
  };
  for (int i = 0; i < (int)runs.size(); ++i) {
    if (i < c_emitted) continue;  // pulled forward by a tight ring
    const FusedRun& tgt = runs[std::min<int>(i + lookahead, (int)runs.size() - 1)];
    // everything of earlier events, and RIRs <= lmax of the target run
    int upto = evs[tgt.ev].fo0 + (tgt.lmax >= tgt.lmin ? tgt.lmax + 1 : 0);
    // the run itself must have its inputs whatever the lookahead target says
    if (runs[i].lmax >= runs[i].lmin) upto = std::max(upto, evs[runs[i].ev].fo0 + runs[i].lmax + 1);
    while (p_emitted < upto && i >= c_emitted) emit_p(p_emitted);
    if (i >= c_emitted) emit_c(c_emitted++);
  }
  while (p_emitted < n_fo) emit_p(p_emitted);  // RIRs no run reads (their a_0 may still be needed by a dry render)
  return ok && nt == n_tasks_expected;
}

// ---- production order of the sweep launch ------------------------------------------------------------------------------
// k_mov_sweep (alr_sweep.cuh): the chunk's sweep events are dealt to `n_slots` sweeper slots (8 CTAs each, one per
// 256-bin slice; greedy by RIR count), every slot walks its events' RIRs in order, and all slots advance at the same
// rate. The producers therefore make "RIR t of every slot" for t = 0, 1, 2, ...: that production order is the ticket
// order of the P-tasks and the allocation order of the ring, so whatever a producer overwrites was produced about one
// ring-full earlier and has long been consumed.
//   prod[p]  ordinal (event-major: ev.fo0 + l) of the p-th RIR produced
//   pop[fo]  production indices [x, y) whose ring regions RIR fo overwrites (wait for ready + consumed of prod[x..y))
bool plan_sweep(EvDev* evs, int ne, IrDev* irs, long long ring_slots, int n_slots_max, FusedTask* tasks,
                int n_tasks_expected, int2* pop, int2* need, int* prod, int n_fo_expected, int* slot_off, int* slot_jobs,
                int n_ev_expected, int* n_slots_out) {
  static thread_local std::vector<int> sweep_ev, load, cur_job, cur_l;
  static thread_local std::vector<std::vector<int>> jobs;
  sweep_ev.clear();
  int n_fo = 0;
  for (int e = 0; e < ne; ++e)
    if (evs[e].fused == 2) {
      evs[e].fo0 = n_fo;
      n_fo += evs[e].N;
      sweep_ev.push_back(e);
    }
  if (n_fo != n_fo_expected || (int)sweep_ev.size() != n_ev_expected) return false;
  const int n_slots = std::max(1, std::min<int>(n_slots_max, (int)sweep_ev.size()));
  jobs.assign(n_slots, {});
  load.assign(n_slots, 0);
  for (int e : sweep_ev) {  // in event order, to the least loaded slot
    int best = 0;
    for (int s2 = 1; s2 < n_slots; ++s2)
      if (load[s2] < load[best]) best = s2;
    jobs[best].push_back(e);
    load[best] += evs[e].N;
  }
  int nj = 0;
  for (int s2 = 0; s2 < n_slots; ++s2) {
    slot_off[s2] = nj;
    for (int e : jobs[s2]) slot_jobs[nj++] = e;
  }
  for (int s2 = n_slots; s2 <= kMaxSweepSlots; ++s2) slot_off[s2] = nj;
  *n_slots_out = n_slots;
  for (int e : sweep_ev)
    for (int l = 0; l < evs[e].N; ++l)
      need[evs[e].fo0 + l] = make_int2(kSwBinCtas * (kSwSweepThreads / 32), evs[e].C * kSwKSplit + 1);
  // production order + ring allocation (FIFO; see plan_fused for the wrap / straddle rules)
  cur_job.assign(n_slots, 0);
  cur_l.assign(n_slots, 0);
  static thread_local std::vector<long long> reg_lo, reg_hi;  // ring region per production index
  reg_lo.resize(n_fo);
  reg_hi.resize(n_fo);
  long long head = 0;
  int live_lo = 0, p = 0, nt = 0;
  bool any = true;
  while (any) {
    any = false;
    for (int s2 = 0; s2 < n_slots; ++s2) {
      while (cur_job[s2] < (int)jobs[s2].size() && cur_l[s2] >= evs[jobs[s2][cur_job[s2]]].N) {
        ++cur_job[s2];
        cur_l[s2] = 0;
      }
      if (cur_job[s2] >= (int)jobs[s2].size()) continue;
      any = true;
      const int e = jobs[s2][cur_job[s2]], l = cur_l[s2]++;
      const EvDev& d = evs[e];
      const long long size = (long long)d.K * d.C;
      if (size > ring_slots) return false;
      const long long old_head = head;
      const bool wrapped = head + size > ring_slots;
      if (wrapped) head = 0;
      const long long lo = head, hi = head + size;
      int pop_lo = live_lo;
      if (live_lo > 0 && !wrapped && reg_lo[live_lo - 1] < hi && reg_hi[live_lo - 1] > lo) pop_lo = live_lo - 1;
      while (live_lo < p) {
        const bool in_tail = wrapped && reg_lo[live_lo] >= old_head;
        if (!in_tail && (reg_hi[live_lo] <= lo || reg_lo[live_lo] >= hi)) break;
        ++live_lo;
      }
      reg_lo[p] = lo;
      reg_hi[p] = hi;
      head = hi;
      const int fo = d.fo0 + l;
      irs[d.ir0 + l].hring = (int)lo;
      pop[fo] = make_int2(pop_lo, live_lo);
      prod[p] = fo;
      for (int c = 0; c < d.C * kSwKSplit; ++c) tasks[nt++] = FusedTask{kTaskP, e, l, c};  // sub = capsule * kSwKSplit + part
      ++p;
    }
  }
  return p == n_fo && nt == n_tasks_expected;
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
  std::vector<float2> tw(kP + 64);
  for (int m = 0; m < kP; ++m) {
    double a = -2.0 * M_PI * (double)m / (double)kP;
    tw[m] = make_float2((float)cos(a), (float)sin(a));
  }
  // compact copy of the pass-B seeds w^(1,2,4,8) of fft_core: entry kP + 16 j + tq = tw[(kP/256 << j) * tq]. The strided
  // originals sit in 16 different 128-byte lines per warp request; the copy is one line per request.
  for (int j = 0; j < 4; ++j)
    for (int tq = 0; tq < 16; ++tq) tw[kP + 16 * j + tq] = tw[((kP / 256) << j) * tq];
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

}  // namespace

// ================================================================================================================
extern "C" {

int alr_version(void) { return ALR_VERSION; }
const char* alr_last_error(void) { return g_err.c_str(); }
int alr_partition_size(void) { return kP; }
int alr_pinned_alloc(alr_context* ctx, size_t bytes, void** out) {
  if (!ctx || !out || bytes == 0) return fail(ALR_ERR_INVALID, "alr_pinned_alloc: bad argument");
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(ALR_ERR_NOMEM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  }
  return ALR_OK;
}
void alr_pinned_free(alr_context* ctx, void* p) {
  if (!ctx || !p) return;
  cudaSetDevice(ctx->device);
  cudaFreeHost(p);
}
int alr_struct_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(alr_event);
    case 1: return (int)sizeof(alr_scene);
    case 2: return (int)sizeof(alr_event_stats);
    case 3: return (int)sizeof(alr_profile);
    case 4: return (int)sizeof(alr_aug_op);
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
  if (cudaFuncSetAttribute(k_cmac, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmacSmem) != cudaSuccess ||
      cudaFuncSetAttribute(k_cmac_both, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmacSmem) != cudaSuccess) {
    cudaGetLastError();
    delete ctx;
    return fail(ALR_ERR_CUDA, "k_cmac: cannot reserve %zu bytes of shared memory", kCmacSmem);
  }
  if (kIrFftSmem > 0 &&
      cudaFuncSetAttribute(k_ir_fft, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kIrFftSmem) != cudaSuccess) {
    cudaGetLastError();
    delete ctx;
    return fail(ALR_ERR_CUDA, "k_ir_fft: cannot reserve %zu bytes of shared memory", kIrFftSmem);
  }
  {
    // k_mov_fused is a persistent launch: exactly as many CTAs as can be resident at once
    int sms = 0, per_sm = 0, khz = 0;
    cudaError_t e1 = cudaFuncSetAttribute(k_mov_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem);
    cudaError_t e2 = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaError_t e3 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mov_fused, kCtaThreads, kFusedSmem);
    cudaError_t e4 = cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess || sms < 1 || per_sm < 1) {
      cudaGetLastError();
      ctx->fused = 0;  // unfused kernels only
      ctx->fused_grid = 0;
    } else {
      ctx->fused_grid = sms * per_sm;
      ctx->sm_clock_khz = std::max(khz, 500000);
    }
    {
      int per_sm2 = 0;
      cudaError_t s1 = cudaFuncSetAttribute(k_mov_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSwSmem);
      cudaError_t s2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, k_mov_sweep, kSwThreads, kSwSmem);
      if (kSweepOk && s1 == cudaSuccess && s2 == cudaSuccess && per_sm2 >= 1 && sms >= kSwBinCtas) ctx->sweep_grid = sms;
      else cudaGetLastError();
    }
    // Experiment (ALR_L2_PERSIST=1, off): pin the ring with a persisting access-policy window. Measured: k_mov_sweep
    // unchanged (13.06 vs 13.10 ms per benchmark step) while every OTHER kernel of the step got slower with 72 MB of L2
    // set aside (step 23.3 -> 26.7 ms), profiles/r02_sweep_tuning.txt.
    if (ctx->sweep_grid > 0 && getenv("ALR_L2_PERSIST") && atoi(getenv("ALR_L2_PERSIST")) != 0) {
      int max_persist = 0, max_window = 0;
      cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
      cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, device);
      if (max_persist > 0 && max_window > 0) {
        const size_t want = std::min<size_t>((size_t)max_persist, (size_t)ctx->ring_bytes + ((size_t)8 << 20));
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
          ctx->l2_persist_bytes = (int64_t)want;
          ctx->l2_window_max = max_window;
        } else {
          cudaGetLastError();
        }
      }
    }
    if (getenv("ALR_TRACE")) fprintf(stderr, "[alr] L2 persisting %lld bytes, window max %lld\n", (long long)ctx->l2_persist_bytes, (long long)ctx->l2_window_max);
    if (const char* v = getenv("ALR_FUSED")) {
      const int m = atoi(v);
      ctx->fused = (m == 1 && ctx->fused_grid > 0) ? 1 : (m == 2 && ctx->sweep_grid > 0) ? 2 : 0;
    }
    if (const char* v = getenv("ALR_RING_MB")) ctx->ring_bytes = std::max<int64_t>(1, atoll(v)) << 20;
    if (const char* v = getenv("ALR_LOOKAHEAD")) ctx->lookahead = std::max(0, atoi(v));
    if (const char* v = getenv("ALR_MIX_GROUP")) ctx->mix_group = std::max(0, atoi(v));
    if (const char* v = getenv("ALR_CMAC_MERGE")) ctx->cmac_merge = atoi(v) != 0;
    if (const char* v = getenv("ALR_SMALL")) ctx->small_rir = atoi(v) != 0;
    if (const char* v = getenv("ALR_WATCHDOG_MS")) ctx->watchdog_ms = std::max(1LL, atoll(v));
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
  ctx->augbuf.release();
  ctx->augdesc.release();
  ctx->ring.release();
  ctx->ambgen.release();
  ctx->visbuf.release();
  ctx->stage.release();
  ctx->stage_out.release();
  ctx->stage_aug.release();
  for (auto ev : ctx->ev_pool) cudaEventDestroy(ev);
  for (auto ev : ctx->sync_pool) cudaEventDestroy(ev);
  if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
  if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
  if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
  if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
  delete ctx;
}

int alr_set_workspace_limit(alr_context* ctx, int64_t bytes) {
  if (!ctx || bytes < (1 << 16)) return fail(ALR_ERR_INVALID, "alr_set_workspace_limit: bad argument");
  ctx->ws_limit = bytes;
  return ALR_OK;
}

int alr_set_option(alr_context* ctx, const char* name, int64_t value) {
  if (!ctx || !name) return fail(ALR_ERR_INVALID, "alr_set_option: bad argument");
  const std::string n(name);
  if (n == "fused") {
    if (value < 0 || value > 2) return fail(ALR_ERR_INVALID, "alr_set_option: fused must be 0, 1 or 2");
    if ((value == 1 && ctx->fused_grid <= 0) || (value == 2 && ctx->sweep_grid <= 0))
      return fail(ALR_ERR_INVALID, "alr_set_option: that launch is not available on this device");
    ctx->fused = (int)value;
  } else if (n == "ring_bytes") {
    if (value < (1 << 20)) return fail(ALR_ERR_INVALID, "alr_set_option: ring_bytes must be >= 1 MiB");
    ctx->ring_bytes = value;
  } else if (n == "lookahead") {
    if (value < 0 || value > 1024) return fail(ALR_ERR_INVALID, "alr_set_option: lookahead out of range");
    ctx->lookahead = (int)value;
  } else if (n == "small_rir") {
    ctx->small_rir = value != 0;
  } else if (n == "cmac_merge") {
    ctx->cmac_merge = value != 0;
  } else if (n == "mix_group") {
    if (value < 0) return fail(ALR_ERR_INVALID, "alr_set_option: mix_group must be >= 0");
    ctx->mix_group = (int)value;
  } else {
    return fail(ALR_ERR_INVALID, "alr_set_option: unknown option '%s'", name);
  }
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
  // f3: ambience layers without a buffer are Gaussian noise generated on the device (alr_scene.ambience_seed)
  struct GenLayer {
    int scene, layer;
    size_t off;
  };
  std::vector<GenLayer> gen_layers;
  std::vector<std::vector<char>> amb_gen(n_scenes);
  size_t gen_bytes = 0;
  int gen_channels = 0;
  for (int64_t s = 0; s < n_scenes; ++s) {
    if (scenes[s].n_ambience > 0 && scenes[s].ambience) {
      amb_ptrs[s].assign(scenes[s].ambience, scenes[s].ambience + scenes[s].n_ambience);
      amb_gen[s].assign(scenes[s].n_ambience, 0);
      for (int a = 0; a < scenes[s].n_ambience; ++a) {
        if (amb_ptrs[s][a]) continue;
        if (!scenes[s].ambience_seed) return fail(ALR_ERR_INVALID, "scene %d: null ambience %d (and no ambience_seed)", (int)s, a);
        if (scenes[s].n_channels < 1 || scenes[s].n_samples < 1) return fail(ALR_ERR_INVALID, "scene %d: bad shape", (int)s);
        amb_gen[s][a] = 1;
        gen_layers.push_back({(int)s, a, gen_bytes});
        gen_bytes += align_up((size_t)scenes[s].n_channels * scenes[s].n_samples * sizeof(float), 256);
        gen_channels += scenes[s].n_channels;
      }
      scenes[s].ambience = amb_ptrs[s].data();
    }
  }
  std::vector<GenDev> h_gens;
  size_t gen_off_peaks = 0, gen_off_data = 0;
  if (!gen_layers.empty()) {
    gen_off_peaks = align_up(gen_layers.size() * sizeof(GenDev), 256);
    gen_off_data = align_up(gen_off_peaks + (size_t)gen_channels * kGenSlices * sizeof(float), 256);
    int rc = ctx->ambgen.ensure(gen_off_data + gen_bytes);
    if (rc) return rc;
    int part = 0;
    for (const GenLayer& gl : gen_layers) {
      GenDev gd;
      gd.out = (float*)((char*)ctx->ambgen.p + gen_off_data + gl.off);
      gd.T = scenes[gl.scene].n_samples;
      gd.C = scenes[gl.scene].n_channels;
      gd.part0 = part;
      gd.seed = scenes[gl.scene].ambience_seed[gl.layer];
      part += gd.C;
      h_gens.push_back(gd);
      amb_ptrs[gl.scene][gl.layer] = gd.out;  // a DEVICE pointer from here on, also in host mode
    }
  }

  if (!ctx->ev_t0) {
    CUDA_TRY(cudaEventCreate(&ctx->ev_t0));
    CUDA_TRY(cudaEventCreate(&ctx->ev_t1));
  }
  CUDA_TRY(cudaEventRecord(ctx->ev_t0, st));

  // ---- host mode: stage every buffer on the device -------------------------------------------------------------
  struct OutCopy {
    void* host;
    const void* dev;
    size_t bytes;
  };
  struct InCopy {
    const float* host;
    size_t off;
    size_t bytes;
    int kind;  // 0 linear, 1 IR block
    int ev;    // event that first needs it (copies are listed in event order); -1: ambience
    bool is_audio = false;  // dry audio (small): uploaded up front when any event is augmented on the device
    bool done = false;
  };
  std::vector<OutCopy> out_spatial(mem_space == ALR_MEM_HOST ? n_events : 0), out_dry(mem_space == ALR_MEM_HOST ? n_events : 0);
  std::vector<OutCopy> out_mix, out_pcm;
  std::vector<InCopy> in_copies;   // event inputs, in event order
  std::vector<InCopy> amb_copies;  // ambience layers, `ev` holds the scene index; in scene order
  const bool host_mode = mem_space == ALR_MEM_HOST;
  char* arena_base = nullptr;
  if (host_mode) {
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
      if (u.n_out < 1 || u.n_channels < 1 || (!u.spatial && (u.scene < 0 || u.n_irs == -1)))
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
        in_copies.push_back({u.audio, off, (size_t)u.n_audio * sizeof(float), 0, (int)i, true, false});
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
    std::vector<size_t> off_mix(n_scenes), off_pcm(n_scenes);
    std::vector<std::vector<size_t>> off_amb(n_scenes);
    for (int64_t s = 0; s < n_scenes; ++s) {
      alr_scene& u = scenes[s];
      if (u.n_channels < 1 || u.n_samples < 1 || (!u.mix && !u.pcm16)) return fail(ALR_ERR_INVALID, "scene %d: bad shape", (int)s);
      size_t bytes = (size_t)u.n_channels * u.n_samples * sizeof(float);
      for (int a = 0; a < u.n_ambience; ++a) {
        const float* p = amb_ptrs[s][a];
        if (amb_gen[s][a]) {  // generated on the device: nothing to upload
          off_amb[s].push_back(0);
          continue;
        }
        if (!p) return fail(ALR_ERR_INVALID, "scene %d: null ambience", (int)s);
        auto it = seen.find(p);
        size_t off;
        if (it == seen.end()) {
          off = reserve(bytes);
          seen[p] = off;
          amb_copies.push_back({p, off, bytes, 0, (int)s});
        } else {
          off = it->second;
        }
        off_amb[s].push_back(off);
      }
      off_mix[s] = reserve(bytes);
      if (u.pcm16) off_pcm[s] = reserve(bytes / 2);
    }
    int rc = ctx->arena.ensure(total);
    if (rc) return rc;
    char* base = (char*)ctx->arena.p;
    arena_base = base;
    if (!ctx->s_h2d) {
      CUDA_TRY(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
      CUDA_TRY(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
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
      out_spatial[i] = {u.spatial, base + off_sp[i], sp_bytes};
      u.spatial = (float*)(base + off_sp[i]);
      if (u.dry) {
        out_dry[i] = {u.dry, base + off_dry[i], (size_t)(u.n_audio + u.n_ir_samples - 1) * sizeof(float)};
        u.dry = (float*)(base + off_dry[i]);
      }
    }
    for (int64_t s = 0; s < n_scenes; ++s) {
      alr_scene& u = scenes[s];
      for (int a = 0; a < u.n_ambience; ++a)
        if (!amb_gen[s][a]) amb_ptrs[s][a] = (const float*)(base + off_amb[s][a]);
      out_mix.push_back({u.mix, base + off_mix[s], (size_t)u.n_channels * u.n_samples * sizeof(float)});
      u.mix = (float*)(base + off_mix[s]);
      if (u.pcm16) {
        out_pcm.push_back({u.pcm16, base + off_pcm[s], (size_t)u.n_channels * u.n_samples * sizeof(int16_t)});
        u.pcm16 = (int16_t*)(base + off_pcm[s]);
      } else {
        out_pcm.push_back({nullptr, nullptr, 0});
      }
    }
  }

  // ---- f1: device-side augmentation of the dry audio (ping-pong buffers; the event then reads the final buffer) -------
  constexpr int kAugLevels = 8, kAugSlices = 16;
  std::vector<AugDev> aug_point[kAugLevels], aug_iir[kAugLevels];
  std::vector<int> aug_iir_prefix[kAugLevels];
  std::vector<NormDev> aug_norm;
  struct AudioOut {
    void* dst;
    const void* src;
    size_t bytes;
  };
  std::vector<AudioOut> aug_audio_out;
  int aug_chunks_total = 0, aug_any = 0;
  std::vector<int> ev_norm_idx(n_events, -1);  // index of the event's peak-normalisation scalar, or -1
  {
    size_t aug_floats = 0;
    for (int64_t i = 0; i < n_events; ++i) {
      const alr_event& u = events_in[i];
      if (u.n_irs == -1 || (u.n_aug_ops <= 0 && !u.normalize_audio)) continue;
      if (u.n_aug_ops > kAugLevels) return fail(ALR_ERR_INVALID, "event %d: more than %d augmentations", (int)i, kAugLevels);
      if (u.n_aug_ops > 0 && !u.aug_ops) return fail(ALR_ERR_INVALID, "event %d: aug_ops is NULL", (int)i);
      aug_floats += 2 * align_up((size_t)u.n_audio, 64);
      ++aug_any;
    }
    if (aug_any) {
      int rc = ctx->augbuf.ensure(aug_floats * sizeof(float));
      if (rc) return rc;
      float* base = (float*)ctx->augbuf.p;
      size_t off = 0;
      for (int l = 0; l < kAugLevels; ++l) aug_iir_prefix[l].push_back(0);
      for (int64_t i = 0; i < n_events; ++i) {
        const alr_event& uin = events_in[i];
        alr_event& u = events[i];
        if (uin.n_irs == -1 || (uin.n_aug_ops <= 0 && !uin.normalize_audio)) continue;
        const int L = (int)u.n_audio;
        float* bufA = base + off;
        float* bufB = bufA + align_up((size_t)L, 64);
        off += 2 * align_up((size_t)L, 64);
        const float* src = u.audio;  // device pointer (already staged in host mode)
        const int n_ops = std::max(1, uin.n_aug_ops);  // normalise-only events get an identity gain (copy)
        for (int l = 0; l < n_ops; ++l) {
          AugDev a;
          memset(&a, 0, sizeof(a));
          a.src = src;
          a.dst = (l & 1) ? bufB : bufA;
          a.L = L;
          if (uin.n_aug_ops <= 0) {
            a.type = kAugGain;
            a.p[0] = 1.0;
          } else {
            const alr_aug_op& op = uin.aug_ops[l];
            if (op.type < 0 || op.type > kAugDelay) return fail(ALR_ERR_INVALID, "event %d: bad augmentation type %d", (int)i, op.type);
            a.type = op.type;
            a.fin_shape = op.fade_in_shape;
            a.fout_shape = op.fade_out_shape;
            a.fin = std::min(std::max(op.fade_in_samples, 0), L);
            a.fout = std::min(std::max(op.fade_out_samples, 0), L);
            for (int q = 0; q < 6; ++q) a.p[q] = op.p[q];
          }
          if (a.type == kAugBiquad || a.type == kAugDeemph) {
            a.nchunks = ceil_div(L, kIirChunk);
            a.chunk0 = aug_chunks_total;
            aug_chunks_total += a.nchunks;
            aug_iir[l].push_back(a);
            aug_iir_prefix[l].push_back(aug_iir_prefix[l].back() + a.nchunks);
          } else {
            aug_point[l].push_back(a);
          }
          src = a.dst;
        }
        float* fin = const_cast<float*>(src);
        if (uin.normalize_audio) {
          NormDev nd;
          nd.x = fin;
          nd.L = L;
          nd.part0 = (int)aug_norm.size() * kAugSlices;
          nd.scale_in_place = uin.audio_out ? 1 : 0;
          ev_norm_idx[i] = (int)aug_norm.size();
          aug_norm.push_back(nd);
        }
        if (uin.audio_out) aug_audio_out.push_back({uin.audio_out, fin, (size_t)L * sizeof(float)});
        u.audio = fin;
      }
    }
  }

  // the augmentation descriptor buffer is sized now (upper bound): the normalisation scalars live at its start and the
  // event descriptors built below point at them
  const size_t xnorm_bytes = align_up(std::max<size_t>(aug_norm.size(), 1) * sizeof(float), 256);
  size_t aug_desc_reserved = 0;
  if (aug_any) {
    size_t n_ops_total = 0;
    for (int l = 0; l < kAugLevels; ++l) n_ops_total += aug_point[l].size() + aug_iir[l].size();
    aug_desc_reserved = xnorm_bytes + n_ops_total * (sizeof(AugDev) + sizeof(int)) + kAugLevels * 3 * 64 +
                        aug_norm.size() * sizeof(NormDev) + (size_t)std::max(aug_chunks_total, 1) * sizeof(double2) +
                        std::max<size_t>(aug_norm.size(), 1) * kAugSlices * sizeof(float) + 4096;
    int rc = ctx->augdesc.ensure(aug_desc_reserved);
    if (rc) return rc;
  }
  const float* d_xnorm = aug_any ? (const float*)ctx->augdesc.p : nullptr;

  // ---- sizes, chunks and buffers (cheap pre-pass; no detailed planning yet) -----------------------------------------
  const auto host_t0 = std::chrono::steady_clock::now();
  // phase 0: the caller's events; phase 1: dry / direct-path sub-events (static mono renders of IR (ref channel, 0)
  // windowed around its peak, compute_dry_audio synthesize.py:432-504), which depend on phase-0 results
  std::vector<alr_event> dry_events;
  std::vector<int> dry_parent;
  for (int64_t i = 0; i < n_events; ++i) {
    const alr_event& u = events[i];
    if (u.scene >= n_scenes) return fail(ALR_ERR_INVALID, "event %d: scene index %d out of range", (int)i, u.scene);
    if (u.scene >= 0 && scenes[u.scene].n_channels != u.n_channels)
      return fail(ALR_ERR_INVALID, "event %d: %d channels but scene %d has %d", (int)i, u.n_channels, u.scene,
                  scenes[u.scene].n_channels);
    if (!u.dry || u.n_irs == -1) continue;
    if (u.n_irs < 1) return fail(ALR_ERR_INVALID, "event %d: dry audio needs at least one IR", (int)i);
    if (u.dry_channel < 0 || u.dry_channel >= u.n_channels)
      return fail(ALR_ERR_INVALID, "Reference channel index out of range for IRs with %d channels", u.n_channels);
    alr_event sdry = u;
    sdry.irs = u.irs + (long long)u.dry_channel * u.ir_stride_c;
    sdry.n_channels = 1;
    sdry.n_irs = 1;
    sdry.ir_frames = nullptr;
    sdry.spatial = u.dry;
    sdry.n_out = u.n_audio + u.n_ir_samples - 1;
    sdry.dry = nullptr;
    dry_events.push_back(sdry);
    dry_parent.push_back((int)i);
  }
  const long long ring_slots = ctx->fused ? (long long)(ctx->ring_bytes / ((int64_t)kP * sizeof(float2))) : 0;
  const std::vector<alr_event>* phase_events[2] = {&events, &dry_events};
  std::vector<EvSize> sizes[2];
  std::vector<Chunk> chunks[2];
  for (int ph = 0; ph < 2; ++ph) {
    const auto& evl = *phase_events[ph];
    sizes[ph].resize(evl.size());
    for (size_t i = 0; i < evl.size(); ++i) {
      int rc = size_event(evl[i], ph == 0 ? (int)i : dry_parent[i], sizes[ph][i], ph == 0 ? ring_slots : 0, ctx->fused,
                          ctx->small_rir ? (ph == 0 ? 1 : 2) : 0);
      if (rc) return rc;
    }
    // host mode: smaller chunks give the upload / compute / download pipeline something to overlap
    make_chunks(sizes[ph], host_mode ? std::min<int64_t>(ctx->ws_limit, (int64_t)512 << 20) : ctx->ws_limit, chunks[ph]);
  }
  long long max_h = 0, max_x = 0, max_y = 0, max_ecap = 0;
  int max_fo = 0;
  size_t blob_total = 0;
  int part_total = 0, ir_total = 0, gain_total = 0;
  for (int ph = 0; ph < 2; ++ph)
    for (Chunk& ch : chunks[ph]) {
      max_h = std::max(max_h, ch.hslots);
      max_x = std::max(max_x, ch.xslots);
      max_y = std::max(max_y, ch.yslots);
      max_fo = std::max(max_fo, ch.n_fo);
      max_ecap = std::max(max_ecap, ch.n_ecap);
      ch.base = blob_total;
      blob_total += ch.bytes;
      ch.part_base = part_total;
      ch.ir_base = ir_total;
      ch.gain_base = gain_total;
      for (int e = ch.ev_begin; e < ch.ev_end; ++e) part_total += sizes[ph][e].n_parts;
      ir_total += ch.n_ir;
      gain_total += ch.ev_end - ch.ev_begin;
    }
  // ---- scene / ambience / mix descriptors (independent of the event plans) ------------------------------------------
  std::vector<SceneDev> h_scenes(n_scenes);
  std::vector<int> mev_event;                 // per mix entry: the event whose gain k_mix applies, or -1
  std::vector<int> gain_from(n_events, 0);    // per event: first sample k_apply_gain still has to scale
  bool any_pcm = false;
  std::vector<AmbDev> h_ambs;
  std::vector<MixEv> h_mevs;
  int n_amb_parts = 0;
  for (int64_t sidx = 0; sidx < n_scenes; ++sidx) {
    const alr_scene& u = scenes[sidx];
    if (u.n_channels < 1 || u.n_samples < 1 || !u.mix) return fail(ALR_ERR_INVALID, "scene %d: bad shape", (int)sidx);
    if (u.n_samples > 0x7ffff000LL) return fail(ALR_ERR_INVALID, "scene %d: too many samples", (int)sidx);
    if (u.n_ambience < 0 || (u.n_ambience > 0 && (!u.ambience || !u.ambience_ref_db)))
      return fail(ALR_ERR_INVALID, "scene %d: bad ambience list", (int)sidx);
    SceneDev& d = h_scenes[sidx];
    memset(&d, 0, sizeof(d));
    d.mix = u.mix;
    d.pcm = u.pcm16;
    any_pcm |= u.pcm16 != nullptr;
    d.C = u.n_channels;
    d.T = u.n_samples;
    d.n_amb = u.n_ambience;
    d.amb0 = (int)h_ambs.size();
    for (int k = 0; k < u.n_ambience; ++k) {
      if (!amb_ptrs[sidx][k]) return fail(ALR_ERR_INVALID, "scene %d: null ambience %d", (int)sidx, k);
      AmbDev ad;
      memset(&ad, 0, sizeof(ad));
      ad.data = amb_ptrs[sidx][k];
      ad.n = (long long)u.n_channels * u.n_samples;
      ad.ref_db = u.ambience_ref_db[k];
      ad.part0 = n_amb_parts;
      ad.nparts = kAmbSlices;
      n_amb_parts += kAmbSlices;
      h_ambs.push_back(ad);
    }
  }
  {
    std::vector<int> count(n_scenes, 0);
    for (int64_t i = 0; i < n_events; ++i)
      if (events[i].scene >= 0 && events[i].scene_end > events[i].scene_start) count[events[i].scene]++;
    int off = 0;
    for (int64_t sidx = 0; sidx < n_scenes; ++sidx) {
      h_scenes[sidx].ev0 = off;
      h_scenes[sidx].nev = 0;
      off += count[sidx];
    }
    h_mevs.resize(off);
    mev_event.assign(off, -1);
    for (int64_t i = 0; i < n_events; ++i) {  // call order == the reference's dict order
      const alr_event& u = events[i];
      if (u.scene < 0 || u.scene_end <= u.scene_start) continue;
      if (u.scene_start < 0 || u.scene_end > scenes[u.scene].n_samples)
        return fail(ALR_ERR_INVALID, "event %d: scene slice [%lld, %lld) outside the scene", (int)i,
                    (long long)u.scene_start, (long long)u.scene_end);
      SceneDev& sd = h_scenes[u.scene];
      MixEv& m = h_mevs[sd.ev0 + sd.nev++];
      m.y = u.spatial;
      m.gain = nullptr;
      // Device buffers: the event gain is applied by k_mix while it reads the event anyway (k_apply_gain then only
      // scales what the scene does not cover). With host buffers the event audio is downloaded right after its chunk,
      // before the mix, so it is scaled per chunk as usual.
      if (!host_mode && u.n_irs != -1 && u.gain_mode != ALR_GAIN_NONE) {
        mev_event[sd.ev0 + sd.nev - 1] = (int)i;
        gain_from[i] = (int)std::min<int64_t>(u.scene_end - u.scene_start, u.n_out);
      }
      m.start = u.scene_start;
      m.end = u.scene_end;
      m.n_out = (int)u.n_out;
      m.pad = 0;
    }
  }
  const size_t mix_off_scenes = align_up(blob_total, 256);
  const size_t mix_off_ambs = align_up(mix_off_scenes + h_scenes.size() * sizeof(SceneDev), 16);
  const size_t mix_off_mevs = align_up(mix_off_ambs + h_ambs.size() * sizeof(AmbDev), 16);
  const size_t desc_total = align_up(mix_off_mevs + h_mevs.size() * sizeof(MixEv), 256) + 256;
  {
    int rc = ctx->stage.ensure(desc_total);
    if (rc) return rc;
    rc = ctx->desc.ensure(desc_total);
    if (rc) return rc;
  }
  char* hbase = (char*)ctx->stage.p;
  char* dbase = (char*)ctx->desc.p;
  // ---- workspaces -----------------------------------------------------------------------------------------------------
  const size_t slot_bytes = (size_t)kP * sizeof(float2);
  {
    int rc = ctx->spec.ensure(std::max<size_t>((size_t)(max_h + max_x + max_y) * slot_bytes, 16));
    if (rc) return rc;
  }
  float2* d_hspec = (float2*)ctx->spec.p;
  float2* d_xspec = d_hspec + (size_t)max_h * kP;
  float2* d_yspec = d_xspec + (size_t)max_x * kP;
  size_t m_off = 0;
  auto m_take = [&](size_t bytes) {
    size_t o = m_off;
    m_off = align_up(m_off + bytes, 256);
    return o;
  };
  const size_t mo_irscale = m_take(std::max<size_t>(ir_total, 1) * sizeof(float));
  const size_t mo_hen = m_take(std::max<size_t>((size_t)max_h, 1) * kEnWarps * sizeof(float));
  const size_t mo_ctl = m_take(sizeof(FusedCtl));
  const size_t mo_flags = m_take(std::max<size_t>((size_t)max_fo, 1) * 2 * sizeof(int));  // ready[], consumed[]
  const size_t mo_ecap = m_take(std::max<size_t>((size_t)max_ecap, 1) * sizeof(float));
  const size_t mo_parts = m_take(std::max<size_t>(part_total, 1) * sizeof(float2));
  const size_t mo_gain = m_take(std::max<size_t>(gain_total, 1) * sizeof(float));
  const size_t mo_stats = m_take(std::max<size_t>(n_events, 1) * sizeof(EvStat));
  const size_t mo_amb = m_take(std::max<size_t>(n_amb_parts, 1) * sizeof(float));
  {
    int rc = ctx->misc.ensure(m_off);
    if (rc) return rc;
    rc = ctx->stage_out.ensure(std::max<size_t>(n_events, 1) * sizeof(EvStat) + sizeof(FusedCtl));
    if (rc) return rc;
    if (max_fo > 0) {
      rc = ctx->ring.ensure((size_t)ctx->ring_bytes);
      if (rc) return rc;
    }
  }
  char* mbase = (char*)ctx->misc.p;
  float* d_irscale = (float*)(mbase + mo_irscale);
  float* d_hen = (float*)(mbase + mo_hen);
  float2* d_parts = (float2*)(mbase + mo_parts);
  float* d_gain = (float*)(mbase + mo_gain);
  for (size_t k = 0; k < h_mevs.size(); ++k)  // phase-0 events own gain slots [0, n_events) in call order
    if (mev_event[k] >= 0) h_mevs[k].gain = d_gain + mev_event[k];
  EvStat* d_stats = (EvStat*)(mbase + mo_stats);
  float* d_ambparts = (float*)(mbase + mo_amb);
  FusedCtl* d_ctl = (FusedCtl*)(mbase + mo_ctl);
  int* d_flags = (int*)(mbase + mo_flags);
  float* d_ecap = (float*)(mbase + mo_ecap);
  CUDA_TRY(cudaMemsetAsync(d_stats, 0, std::max<size_t>(n_events, 1) * sizeof(EvStat), st));
  CUDA_TRY(cudaMemsetAsync(d_ctl, 0, sizeof(FusedCtl), st));
  if (!h_gens.empty()) {
    CUDA_TRY(cudaMemcpyAsync(ctx->ambgen.p, h_gens.data(), h_gens.size() * sizeof(GenDev), cudaMemcpyHostToDevice, st));
    int max_c = 0;
    for (const GenDev& gd : h_gens) max_c = std::max(max_c, gd.C);
    float* d_peaks = (float*)((char*)ctx->ambgen.p + gen_off_peaks);
    for (size_t g0 = 0; g0 < h_gens.size(); g0 += 32768) {
      const unsigned cnt = (unsigned)std::min<size_t>(32768, h_gens.size() - g0);
      const dim3 grid(kGenSlices, (unsigned)max_c, cnt);
      k_amb_gauss<false><<<grid, 256, 0, st>>>((const GenDev*)ctx->ambgen.p + g0, d_peaks);
      CUDA_TRY(cudaGetLastError());
      k_amb_gauss<true><<<grid, 256, 0, st>>>((const GenDev*)ctx->ambgen.p + g0, d_peaks);
      CUDA_TRY(cudaGetLastError());
      ctx->prof.kernel_launches += 2;
    }
  }
  ctx->prof.workspace_bytes = (int64_t)(ctx->spec.cap + ctx->misc.cap + ctx->desc.cap + ctx->arena.cap);
  ctx->prof.n_chunks = (int64_t)(chunks[0].size() + chunks[1].size());
  {
    int rc = prof_mark(ctx, st, kNumCat);  // start marker
    if (rc) return rc;
  }
  double host_plan_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();

  // ---- host-mode copy pipeline ------------------------------------------------------------------------------------------
  // ALR_TRACE=1: timeline of the three streams (timed events, printed relative to the start of the call)
  const bool trace = getenv("ALR_TRACE") != nullptr;
  struct TraceMark {
    cudaEvent_t ev;
    const char* what;
    int chunk;
  };
  std::vector<TraceMark> trace_marks;
  auto trace_mark = [&](cudaStream_t s_, const char* what, int chunk) {
    if (!trace) return;
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, s_);
    trace_marks.push_back({ev, what, chunk});
  };
  ctx->sync_used = 0;
  auto next_sync_event = [&](cudaEvent_t* out) -> int {
    if (ctx->sync_used == ctx->sync_pool.size()) {
      cudaEvent_t ev;
      CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      ctx->sync_pool.push_back(ev);
    }
    *out = ctx->sync_pool[ctx->sync_used++];
    return ALR_OK;
  };
  size_t in_cursor = 0;
  // uploads every input first needed by an event < ev_end (copies are in event order) on the upload stream and
  // makes the compute stream wait for them
  size_t amb_cursor = 0;
  auto upload_ambience_until = [&](int scene_end) -> int {
    while (amb_cursor < amb_copies.size() && amb_copies[amb_cursor].ev < scene_end) {
      const InCopy& c = amb_copies[amb_cursor++];
      CUDA_TRY(cudaMemcpyAsync(arena_base + c.off, c.host, c.bytes, cudaMemcpyHostToDevice, ctx->s_h2d));
      ctx->prof.h2d_bytes += (int64_t)c.bytes;
    }
    return ALR_OK;
  };
  auto upload_until = [&](int ev_end) -> int {
    while (in_cursor < in_copies.size()) {
      const InCopy& c = in_copies[in_cursor];
      if (c.ev >= ev_end) break;
      if (c.done) {
        ++in_cursor;
        continue;
      }
      if (c.kind == 0) {
        CUDA_TRY(cudaMemcpyAsync(arena_base + c.off, c.host, c.bytes, cudaMemcpyHostToDevice, ctx->s_h2d));
      } else {
        const alr_event& u = events_in[c.ev];
        const size_t row = (size_t)u.n_ir_samples * sizeof(float);
        if (u.ir_stride_n == u.n_ir_samples && u.ir_stride_c == (int64_t)u.n_irs * u.n_ir_samples) {
          // fully contiguous (C, N, Lh): one linear copy (2-D copies do not overlap with downloads on this platform)
          CUDA_TRY(cudaMemcpyAsync(arena_base + c.off, u.irs, c.bytes, cudaMemcpyHostToDevice, ctx->s_h2d));
        } else if (u.ir_stride_n == u.n_ir_samples) {
          CUDA_TRY(cudaMemcpy2DAsync(arena_base + c.off, row * u.n_irs, u.irs, (size_t)u.ir_stride_c * sizeof(float),
                                     row * u.n_irs, u.n_channels, cudaMemcpyHostToDevice, ctx->s_h2d));
        } else {
          for (int chn = 0; chn < u.n_channels; ++chn)
            CUDA_TRY(cudaMemcpy2DAsync(arena_base + c.off + (size_t)chn * u.n_irs * row, row,
                                       u.irs + (long long)chn * u.ir_stride_c, (size_t)u.ir_stride_n * sizeof(float),
                                       row, u.n_irs, cudaMemcpyHostToDevice, ctx->s_h2d));
        }
      }
      ctx->prof.h2d_bytes += (int64_t)c.bytes;
      ++in_cursor;
    }
    return ALR_OK;
  };
  auto compute_waits_for_uploads = [&]() -> int {
    cudaEvent_t ev;
    int rc = next_sync_event(&ev);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(ev, ctx->s_h2d));
    CUDA_TRY(cudaStreamWaitEvent(st, ev, 0));
    return ALR_OK;
  };
  auto download_after_compute = [&](const OutCopy* list, size_t n) -> int {
    cudaEvent_t ev;
    int rc = next_sync_event(&ev);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(ev, st));
    CUDA_TRY(cudaStreamWaitEvent(ctx->s_d2h, ev, 0));
    for (size_t i = 0; i < n; ++i) {
      if (!list[i].host) continue;
      CUDA_TRY(cudaMemcpyAsync(list[i].host, list[i].dev, list[i].bytes, cudaMemcpyDeviceToHost, ctx->s_d2h));
      ctx->prof.d2h_bytes += (int64_t)list[i].bytes;
    }
    return ALR_OK;
  };
  // ---- mixdown of scenes [s0, s1): ambience scale, then k_mix; in host mode followed by the download of the mixes
  int scenes_mixed = 0;
  std::vector<int> scene_last_ev(n_scenes, -1);  // a scene can be mixed once this event's chunk has been rendered
  for (int64_t i = 0; i < n_events; ++i)
    if (events[i].scene >= 0) scene_last_ev[events[i].scene] = (int)i;
  // number of leading scenes whose events are all < ev_end (scenes are mixed in index order)
  auto scenes_done_after = [&](int ev_end) -> int {
    int sidx = scenes_mixed;
    while (sidx < (int)n_scenes && scene_last_ev[sidx] < ev_end) ++sidx;
    return sidx;
  };
  bool mix_desc_uploaded = false;
  SceneDev* d_scenes = (SceneDev*)(dbase + mix_off_scenes);
  AmbDev* d_ambs = (AmbDev*)(dbase + mix_off_ambs);
  MixEv* d_mevs = (MixEv*)(dbase + mix_off_mevs);
  auto upload_mix_desc = [&](cudaStream_t s_) -> int {
    memcpy(hbase + mix_off_scenes, h_scenes.data(), h_scenes.size() * sizeof(SceneDev));
    if (!h_ambs.empty()) memcpy(hbase + mix_off_ambs, h_ambs.data(), h_ambs.size() * sizeof(AmbDev));
    if (!h_mevs.empty()) memcpy(hbase + mix_off_mevs, h_mevs.data(), h_mevs.size() * sizeof(MixEv));
    CUDA_TRY(cudaMemcpyAsync(dbase + mix_off_scenes, hbase + mix_off_scenes, desc_total - 256 - mix_off_scenes,
                             cudaMemcpyHostToDevice, s_));
    ctx->prof.h2d_bytes += (int64_t)(desc_total - 256 - mix_off_scenes);
    mix_desc_uploaded = true;
    return ALR_OK;
  };
  auto launch_mix = [&](int s0, int s1) -> int {
    if (s1 <= s0) return ALR_OK;
    if (!mix_desc_uploaded) {
      int rc = upload_mix_desc(st);
      if (rc) return rc;
    }
    // Groups of `mix_group` scenes: the ambience of a group (23 MB per one-minute 4-channel layer) is reduced and then
    // mixed right away, so that the mixdown's second read of it can hit the 126 MB L2 (profiles/r02_mix_group.txt).
    const int G = ctx->mix_group > 0 ? ctx->mix_group : (s1 - s0);
    for (int g0 = s0; g0 < s1; g0 += G) {
      const int g1 = std::min(g0 + G, s1);
      const int a0 = h_scenes[g0].amb0;
      const int a1 = h_scenes[g1 - 1].amb0 + h_scenes[g1 - 1].n_amb;
      for (int a = a0; a < a1; a += 32768) {
        const int cnt = std::min(32768, a1 - a);
        k_amb_partial<<<dim3(kAmbSlices, cnt), 256, 0, st>>>(d_ambs + a, d_ambparts);
        LAUNCH_CHECK(kCatMix);
      }
      if (a1 > a0) {
        k_amb_final<<<ceil_div((long long)(a1 - a0) * 32, 128), 128, 0, st>>>(d_ambs + a0, a1 - a0, d_ambparts);
        LAUNCH_CHECK(kCatMix);
      }
      long long max_t = 0;
      for (int sidx = g0; sidx < g1; ++sidx) max_t = std::max(max_t, h_scenes[sidx].T);
      for (int sidx = g0; sidx < g1; sidx += 32768) {
        const int cnt = std::min(32768, g1 - sidx);
        k_mix<<<dim3(ceil_div(max_t, 1024), cnt), 256, 0, st>>>(d_scenes + sidx, d_ambs, d_mevs);
        LAUNCH_CHECK(kCatMix);
        if (any_pcm) {
          k_pcm16<<<dim3(ceil_div(max_t, 256), cnt), 256, 0, st>>>(d_scenes + sidx);
          LAUNCH_CHECK(kCatMix);
        }
      }
    }
    if (host_mode) {
      int rc = download_after_compute(out_mix.data() + s0, (size_t)(s1 - s0));
      if (rc) return rc;
      rc = download_after_compute(out_pcm.data() + s0, (size_t)(s1 - s0));
      if (rc) return rc;
    }
    scenes_mixed = s1;
    return ALR_OK;
  };

  if (host_mode) {
    // the upload stream must not start before work already queued on the caller's stream (e.g. a previous call)
    cudaEvent_t ev;
    int rc = next_sync_event(&ev);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(ev, st));
    CUDA_TRY(cudaStreamWaitEvent(ctx->s_h2d, ev, 0));
    CUDA_TRY(cudaStreamWaitEvent(ctx->s_d2h, ev, 0));
    if (n_scenes > 0) {
      rc = upload_mix_desc(ctx->s_h2d);  // every later compute_waits_for_uploads() covers it
      if (rc) return rc;
    }
    if (!chunks[0].empty()) {
      rc = upload_until(chunks[0][0].ev_end);
      if (rc) return rc;
    }
  }

  // ---- f1: run the augmentation chains (all events, before the chunk pipeline; the dry audio is ~1 % of the bytes)
  if (aug_any) {
    if (host_mode) {  // dry audio first
      for (InCopy& c : in_copies) {
        if (!c.is_audio || c.done) continue;
        CUDA_TRY(cudaMemcpyAsync(arena_base + c.off, c.host, c.bytes, cudaMemcpyHostToDevice, ctx->s_h2d));
        ctx->prof.h2d_bytes += (int64_t)c.bytes;
        c.done = true;
      }
    }
    // descriptor blob: per level [pointwise ops][iir ops][iir chunk prefix], then the normalisation list
    Blob ab;
    size_t off_point[kAugLevels], off_iir[kAugLevels], off_pref[kAugLevels];
    ab.bytes.resize(xnorm_bytes, 0);  // [0, xnorm_bytes): the per-event normalisation scalars (written by k_peak_final)
    for (int l = 0; l < kAugLevels; ++l) {
      off_point[l] = ab.add(aug_point[l].data(), aug_point[l].size() * sizeof(AugDev));
      off_iir[l] = ab.add(aug_iir[l].data(), aug_iir[l].size() * sizeof(AugDev));
      off_pref[l] = ab.add(aug_iir_prefix[l].data(), aug_iir_prefix[l].size() * sizeof(int));
    }
    const size_t off_norm = ab.add(aug_norm.data(), aug_norm.size() * sizeof(NormDev));
    const size_t off_zs = align_up(ab.bytes.size(), 256);
    const size_t off_peaks = align_up(off_zs + (size_t)std::max(aug_chunks_total, 1) * sizeof(double2), 256);
    const size_t aug_total = off_peaks + std::max<size_t>(aug_norm.size(), 1) * kAugSlices * sizeof(float);
    if (aug_total > aug_desc_reserved) return fail(ALR_ERR_INVALID, "internal: augmentation descriptor bound exceeded");
    {
      int rc = ctx->stage_aug.ensure(ab.bytes.size() + 16);
      if (rc) return rc;
    }
    memcpy(ctx->stage_aug.p, ab.bytes.data(), ab.bytes.size());
    char* ad = (char*)ctx->augdesc.p;
    CUDA_TRY(cudaMemcpyAsync(ad, ctx->stage_aug.p, ab.bytes.size(), cudaMemcpyHostToDevice, host_mode ? ctx->s_h2d : st));
    ctx->prof.h2d_bytes += (int64_t)ab.bytes.size();
    if (host_mode) {
      int rc = compute_waits_for_uploads();
      if (rc) return rc;
    }
    double2* d_zs = (double2*)(ad + off_zs);
    float* d_peaks = (float*)(ad + off_peaks);
    for (int l = 0; l < kAugLevels; ++l) {
      if (!aug_point[l].empty()) {
        k_aug_pointwise<<<dim3(kAugSlices, (unsigned)aug_point[l].size()), 256, 0, st>>>((const AugDev*)(ad + off_point[l]));
        LAUNCH_CHECK(kCatOther);
      }
      if (!aug_iir[l].empty()) {
        const int n_ops = (int)aug_iir[l].size(), n_ch = aug_iir_prefix[l].back();
        const AugDev* dops = (const AugDev*)(ad + off_iir[l]);
        const int* dpref = (const int*)(ad + off_pref[l]);
        // the zero-state scratch is indexed by the level-local chunk number
        k_iir_pass<false><<<ceil_div(n_ch, kIirCta), kIirCta, 0, st>>>(dops, dpref, n_ops, n_ch, d_zs);
        LAUNCH_CHECK(kCatOther);
        k_iir_combine<<<ceil_div(n_ops, 4), 128, 0, st>>>(dops, dpref, n_ops, d_zs);
        LAUNCH_CHECK(kCatOther);
        k_iir_pass<true><<<ceil_div(n_ch, kIirCta), kIirCta, 0, st>>>(dops, dpref, n_ops, n_ch, d_zs);
        LAUNCH_CHECK(kCatOther);
      }
    }
    if (!aug_norm.empty()) {
      const NormDev* dn = (const NormDev*)(ad + off_norm);
      for (size_t n0 = 0; n0 < aug_norm.size(); n0 += 32768) {
        const unsigned cnt = (unsigned)std::min<size_t>(32768, aug_norm.size() - n0);
        k_peak_partial<<<dim3(kAugSlices, cnt), 256, 0, st>>>(dn + n0, d_peaks);
        LAUNCH_CHECK(kCatOther);
        k_peak_final<<<dim3(kAugSlices, cnt), 256, 0, st>>>(dn + n0, d_peaks, kAugSlices, (float*)ad + n0);
        LAUNCH_CHECK(kCatOther);
      }
    }
    for (const AudioOut& ao : aug_audio_out) {
      CUDA_TRY(cudaMemcpyAsync(ao.dst, ao.src, ao.bytes, host_mode ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
      if (host_mode) ctx->prof.d2h_bytes += (int64_t)ao.bytes;
    }
  }

  // ---- stream the chunks: plan into pinned memory, upload, launch; the host plans chunk i+1 while the GPU runs chunk i
  for (int ph = 0; ph < 2; ++ph) {
    const auto& evl = *phase_events[ph];
    for (size_t ci = 0; ci < chunks[ph].size(); ++ci) {
      const Chunk& ch = chunks[ph][ci];

      const auto t_plan0 = std::chrono::steady_clock::now();
      const int ne = ch.ev_end - ch.ev_begin;
      char* hb = hbase + ch.base;
      EvDev* h_evs = (EvDev*)(hb + ch.off_evs);
      IrDev* h_irs = (IrDev*)(hb + ch.off_irs);
      float* h_wband = (float*)(hb + ch.off_wband);
      int2* h_lr = (int2*)(hb + ch.off_lrange);
      int* p_irfft = (int*)(hb + ch.off_irfft);
      int* p_ir = (int*)(hb + ch.off_ir);
      int* p_xfft = (int*)(hb + ch.off_xfft);
      int* p_cmac = (int*)(hb + ch.off_cmac);
      int* p_cmacs = (int*)(hb + ch.off_cmacs);
      int* p_ifft = (int*)(hb + ch.off_ifft);
      int* l_tile = (int*)(hb + ch.off_tile);
      int* l_dry = (int*)(hb + ch.off_dry);
      int* p_small = (int*)(hb + ch.off_small);
      int n_tile = 0, n_dry = 0, ir_off = 0, w_off = 0, blk_off = 0, parts = ch.part_base;
      long long hs = 0, xs = 0, ys = 0, ecap_off = 0;
      p_irfft[0] = p_ir[0] = p_xfft[0] = p_cmac[0] = p_cmacs[0] = p_ifft[0] = p_small[0] = 0;
      for (int i = 0; i < ne; ++i) {
        const int ei = ch.ev_begin + i;
        const EvSize& z = sizes[ph][ei];
        EvDev& d = h_evs[i];
        int x_used = 0;
        plan_event_into(evl[ei], ph == 0 ? ei : dry_parent[ei], z, d, h_irs + ir_off, ir_off, h_wband + w_off, w_off,
                        h_lr + blk_off, blk_off, &x_used);
        {
          const int parent_ev = ph == 0 ? ei : dry_parent[ei];
          d.xnorm = (d_xnorm && ev_norm_idx[parent_ev] >= 0) ? d_xnorm + ev_norm_idx[parent_ev] : nullptr;
        }
        if (ph == 0) d.gain_from = gain_from[ei];
        if (ph == 1) {
          d.gain_mode = kGainDry;
          d.parent = dry_parent[ei];
          d.normalize = events[dry_parent[ei]].normalize_irs != 0;
          l_dry[n_dry++] = i;
        }
        d.hslot0 = hs;
        d.xslot0 = xs;
        d.yslot0 = ys;
        d.part0 = parts;
        d.nparts = z.n_parts;
        if (z.small) {
          d.small = 1;
          d.B_out = z.small_runs;  // k_small_rir: runs of kRun convolution blocks
          x_used = 0;
        }
        if (z.fused) {
          d.fused = z.fused;
          d.ecap0 = ecap_off;
          ecap_off += z.n_ptask;
        }
        hs += z.h_ws;
        xs += z.xb;  // the bound, so that slot bases do not depend on the detailed plan
        ys += z.y;
        parts += z.n_parts;
        ir_off += z.n_ir;
        w_off += z.wband;
        blk_off += z.n_blk;
        p_irfft[i + 1] = p_irfft[i] + z.n_irfft;
        p_ir[i + 1] = p_ir[i] + z.n_ir;
        p_xfft[i + 1] = p_xfft[i] + x_used;
        p_cmac[i + 1] = p_cmac[i] + z.n_cmac;
        p_cmacs[i + 1] = p_cmacs[i] + z.n_cmac_static;
        p_ifft[i + 1] = p_ifft[i] + z.n_ifft;
        p_small[i + 1] = p_small[i] + z.n_small;
        if (!z.pass && d.N == 0) l_tile[n_tile++] = i;
      }
      int n_sweep_slots = 0;
      if (ch.n_tasks > 0 && ch.n_sweep_ev == 0 &&
          !plan_fused(h_evs, ne, h_irs, h_lr, ring_slots, ctx->lookahead, (FusedTask*)(hb + ch.off_tasks), ch.n_tasks,
                      (int2*)(hb + ch.off_pop), (int2*)(hb + ch.off_ncons), ch.n_fo))
        return fail(ALR_ERR_INVALID, "internal: inconsistent task plan for the fused launch");
      if (ch.n_sweep_ev > 0 &&
          !plan_sweep(h_evs, ne, h_irs, ring_slots, std::min(kMaxSweepSlots, ctx->sweep_grid / kSwBinCtas),
                      (FusedTask*)(hb + ch.off_tasks), ch.n_tasks, (int2*)(hb + ch.off_pop), (int2*)(hb + ch.off_ncons),
                      (int*)(hb + ch.off_prod), ch.n_fo, (int*)(hb + ch.off_slotoff), (int*)(hb + ch.off_slotjobs),
                      ch.n_sweep_ev, &n_sweep_slots))
        return fail(ALR_ERR_INVALID, "internal: inconsistent production plan for the sweep launch");
      host_plan_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_plan0).count();
      // Descriptors. In host mode they travel on the UPLOAD stream, right behind this chunk's inputs and ahead of the
      // next chunk's: a copy on the compute stream would sit behind everything already queued on the H2D copy
      // engine, which starved the kernels until all uploads had finished (ALR_TRACE timeline, round 1).
      CUDA_TRY(cudaMemcpyAsync(dbase + ch.base, hb, ch.bytes, cudaMemcpyHostToDevice, host_mode ? ctx->s_h2d : st));
      ctx->prof.h2d_bytes += (int64_t)ch.bytes;
      if (host_mode) {
        trace_mark(ctx->s_h2d, "upload done", (int)ci);
        int rc = compute_waits_for_uploads();  // inputs (uploaded one chunk ahead) + descriptors of this chunk
        if (rc) return rc;
        if (ph == 0) {
          // keep the upload stream one chunk ahead of the kernels: first the ambience of the scenes that become
          // complete with this chunk (needed right after its kernels), then the inputs of the next chunk
          rc = upload_ambience_until(scenes_done_after(ch.ev_end));
          if (rc) return rc;
          rc = upload_until(ci + 1 < chunks[0].size() ? chunks[0][ci + 1].ev_end : (int)n_events);
          if (rc) return rc;
        }
      }
      char* db = dbase + ch.base;
      EvDev* c_evs = (EvDev*)(db + ch.off_evs);
      const IrDev* c_irs = (const IrDev*)(db + ch.off_irs);
      const float* c_wband = (const float*)(db + ch.off_wband);
      const int2* c_lr = (const int2*)(db + ch.off_lrange);
      float* c_irscale = d_irscale + ch.ir_base;
      float* c_gain = d_gain + ch.gain_base;
      const int n_irfft = p_irfft[ne], n_xfft = p_xfft[ne], n_cmac = p_cmac[ne], n_cmacs = p_cmacs[ne], n_ifft = p_ifft[ne],
                n_irs = p_ir[ne];
      if (n_dry > 0) {
        k_dry_window<<<n_dry, 256, 0, st>>>(c_evs, (const int*)(db + ch.off_dry), d_stats);
        LAUNCH_CHECK(kCatOther);
      }
      if (p_small[ne] > 0) {
        k_small_rir<<<p_small[ne], kCtaThreads, 0, st>>>(c_evs, ne, (const int*)(db + ch.off_small), ctx->d_tw, ctx->d_zeta,
                                                        d_stats, d_parts);
        LAUNCH_CHECK(kCatIfft);
      }
      if (n_irfft > 0) {
        k_ir_fft<<<ceil_div(n_irfft, kGroupsPerCta * kIrTasks), kCtaThreads, kIrFftSmem, st>>>(c_evs, ne, (const int*)(db + ch.off_irfft),
                                                                        n_irfft, ctx->d_tw, ctx->d_zeta, d_hspec, d_hen);
        LAUNCH_CHECK(kCatIrFft);
        k_ir_scale<<<ceil_div((long long)n_irs * 32, 128), 128, 0, st>>>(c_evs, ne, (const int*)(db + ch.off_ir), n_irs,
                                                                       d_hen, c_irscale, d_stats);
        LAUNCH_CHECK(kCatOther);
      }
      if (n_xfft > 0) {
        k_x_fft<<<ceil_div(n_xfft, kGroupsPerCta * kXTasks), kCtaThreads, 0, st>>>(c_evs, ne, (const int*)(db + ch.off_xfft), n_xfft,
                                                                       c_irs, c_wband, c_irscale, ctx->d_tw, ctx->d_zeta,
                                                                       ctx->d_win, d_xspec);
        LAUNCH_CHECK(kCatXFft);
      }
      if (ch.n_sweep_ev > 0) {
        CUDA_TRY(cudaMemsetAsync(d_flags, 0, (size_t)ch.n_fo * 2 * sizeof(int), st));
        CUDA_TRY(cudaMemsetAsync(&d_ctl->ticket, 0, sizeof(int), st));
        SweepArgs sa;
        sa.evs = c_evs;
        sa.irs = c_irs;
        sa.tasks = (const FusedTask*)(db + ch.off_tasks);
        sa.n_tasks = ch.n_tasks;
        sa.pop = (const int2*)(db + ch.off_pop);
        sa.need = (const int2*)(db + ch.off_ncons);
        sa.prod = (const int*)(db + ch.off_prod);
        sa.slot_off = (const int*)(db + ch.off_slotoff);
        sa.slot_jobs = (const int*)(db + ch.off_slotjobs);
        sa.n_slots = n_sweep_slots;
        sa.ctl = d_ctl;
        sa.ready = d_flags;
        sa.consumed = d_flags + ch.n_fo;
        sa.ecap = d_ecap;
        sa.irscale = c_irscale;
        sa.stats = d_stats;
        sa.tw = ctx->d_tw;
        sa.zeta = ctx->d_zeta;
        sa.xspec = d_xspec;
        sa.hring = (float2*)ctx->ring.p;
        sa.yspec = d_yspec;
        sa.spin_limit = (long long)ctx->sm_clock_khz * ctx->watchdog_ms;
        // Pin the ring in L2 for this launch: the taps (9 GB per benchmark step), X and Y stream through the same cache and
        // would otherwise push ring lines out between a producer's store and the sweepers' reads (56 % hit rate and 5 GB of
        // ring write-backs per launch without this, profiles/r02_sweep_v1.txt).
        if (ctx->l2_persist_bytes > 0) {
          cudaStreamAttrValue av;
          memset(&av, 0, sizeof(av));
          av.accessPolicyWindow.base_ptr = ctx->ring.p;
          av.accessPolicyWindow.num_bytes = std::min<size_t>((size_t)ctx->ring_bytes, (size_t)ctx->l2_window_max);
          av.accessPolicyWindow.hitRatio = std::min(1.0f, (float)ctx->l2_persist_bytes / (float)av.accessPolicyWindow.num_bytes);
          av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
          av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
          CUDA_TRY(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
        }
        k_mov_sweep<<<ctx->sweep_grid, kSwThreads, kSwSmem, st>>>(sa);
        LAUNCH_CHECK(kCatFused);
        if (ctx->l2_persist_bytes > 0) {
          cudaStreamAttrValue av;
          memset(&av, 0, sizeof(av));
          av.accessPolicyWindow.num_bytes = 0;
          CUDA_TRY(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
        }
      } else if (ch.n_tasks > 0) {
        CUDA_TRY(cudaMemsetAsync(d_flags, 0, (size_t)ch.n_fo * 2 * sizeof(int), st));
        CUDA_TRY(cudaMemsetAsync(&d_ctl->ticket, 0, sizeof(int), st));
        FusedArgs fa;
        fa.evs = c_evs;
        fa.irs = c_irs;
        fa.lrange = c_lr;
        fa.tasks = (const FusedTask*)(db + ch.off_tasks);
        fa.n_tasks = ch.n_tasks;
        fa.pop = (const int2*)(db + ch.off_pop);
        fa.need = (const int2*)(db + ch.off_ncons);
        fa.ctl = d_ctl;
        fa.ready = d_flags;
        fa.consumed = d_flags + ch.n_fo;
        fa.ecap = d_ecap;
        fa.irscale = c_irscale;
        fa.stats = d_stats;
        fa.tw = ctx->d_tw;
        fa.zeta = ctx->d_zeta;
        fa.xspec = d_xspec;
        fa.hring = (float2*)ctx->ring.p;
        fa.yspec = d_yspec;
        fa.spin_limit = (long long)ctx->sm_clock_khz * ctx->watchdog_ms;  // ~2 s of SM clocks
        k_mov_fused<<<std::min(ctx->fused_grid, ch.n_tasks), kCtaThreads, kFusedSmem, st>>>(fa);
        LAUNCH_CHECK(kCatFused);
      }
      if (ctx->cmac_merge && n_cmac > 0 && n_cmacs > 0 && (long long)n_cmac + n_cmacs < 0x7fffffffLL) {
        const int period = std::max(1, (int)(((long long)n_cmac + n_cmacs) / n_cmacs));
        k_cmac_both<<<n_cmac + n_cmacs, kCtaThreads, kCmacSmem, st>>>(c_evs, ne, (const int*)(db + ch.off_cmac),
                                                             (const int*)(db + ch.off_cmacs), n_cmacs, period, c_irs, c_lr,
                                                             d_xspec, d_hspec, d_yspec);
        LAUNCH_CHECK(kCatCmac);
      } else {
      if (n_cmac > 0) {
        k_cmac<<<n_cmac, kCtaThreads, kCmacSmem, st>>>(c_evs, ne, (const int*)(db + ch.off_cmac), c_irs, c_lr, d_xspec, d_hspec,
                                              d_yspec);
        LAUNCH_CHECK(kCatCmac);
      }
      if (n_cmacs > 0) {
        k_cmac_static<<<n_cmacs, kCtaThreads, 0, st>>>(c_evs, ne, (const int*)(db + ch.off_cmacs), c_irs, d_xspec, d_hspec,
                                                      d_yspec);
        LAUNCH_CHECK(kCatCmacStatic);
      }
      }
      if (n_ifft > 0) {
        k_ifft_ola<<<n_ifft, kCtaThreads, 0, st>>>(c_evs, ne, (const int*)(db + ch.off_ifft), ctx->d_tw, ctx->d_zeta,
                                                  d_yspec, d_parts);
        LAUNCH_CHECK(kCatIfft);
      }
      if (n_tile > 0) {
        k_tile<<<dim3(kTileSlices, n_tile), 256, 0, st>>>(c_evs, (const int*)(db + ch.off_tile), kTileSlices, d_parts);
        LAUNCH_CHECK(kCatOther);
      }
      k_event_gain<<<ceil_div((long long)ne * 32, 128), 128, 0, st>>>(c_evs, ne, d_parts, d_stats, c_gain);
      LAUNCH_CHECK(kCatMix);
      for (int e0 = 0; e0 < ne; e0 += 32768) {
        const int cnt = std::min(32768, ne - e0);
        k_apply_gain<<<dim3(kGainSlices, cnt), 256, 0, st>>>(c_evs + e0, c_gain + e0);
        LAUNCH_CHECK(kCatMix);
      }
      trace_mark(st, "kernels done", (int)ci);
      if (host_mode) {  // results of this chunk go home while the next chunk computes
        int rc;
        if (ph == 0) {
          std::vector<OutCopy> list;
          for (int e = ch.ev_begin; e < ch.ev_end; ++e)
            if (events_in[e].n_irs != -1) list.push_back(out_spatial[e]);
          rc = download_after_compute(list.data(), list.size());
          if (rc) return rc;
          trace_mark(ctx->s_d2h, "event download done", (int)ci);
          // scenes whose last event is in this chunk: mix them now so that their download overlaps later uploads
          const int s1 = scenes_done_after(ch.ev_end);
          if (s1 > scenes_mixed) {
            rc = compute_waits_for_uploads();  // their ambience
            if (rc) return rc;
            rc = launch_mix(scenes_mixed, s1);
            trace_mark(st, "mix done", (int)ci);
            trace_mark(ctx->s_d2h, "mix download done", (int)ci);
          }
        } else {
          std::vector<OutCopy> list;
          for (int e = ch.ev_begin; e < ch.ev_end; ++e) list.push_back(out_dry[dry_parent[e]]);
          rc = download_after_compute(list.data(), list.size());
        }
        if (rc) return rc;
      }
    }
  }

  if (host_mode) {  // whatever is left (scenes without events, ...) before the final mixdown
    int rc = upload_until((int)n_events);
    if (rc) return rc;
    rc = upload_ambience_until((int)n_scenes);
    if (rc) return rc;
    rc = compute_waits_for_uploads();
    if (rc) return rc;
  }
  {
    int rc = launch_mix(scenes_mixed, (int)n_scenes);
    if (rc) return rc;
  }
  ctx->prof.ms_host_plan = host_plan_ms;

  // ---- results back ---------------------------------------------------------------------------------------------------
  FusedCtl* h_ctl = (FusedCtl*)((char*)ctx->stage_out.p + std::max<size_t>(n_events, 1) * sizeof(EvStat));
  h_ctl->abort = 0;
  if (n_events > 0) {
    CUDA_TRY(cudaMemcpyAsync(ctx->stage_out.p, d_stats, n_events * sizeof(EvStat), cudaMemcpyDeviceToHost, st));
    ctx->prof.d2h_bytes += (int64_t)(n_events * sizeof(EvStat));
    if (max_fo > 0) CUDA_TRY(cudaMemcpyAsync(h_ctl, d_ctl, sizeof(FusedCtl), cudaMemcpyDeviceToHost, st));
  }
  if (host_mode) {
    cudaEvent_t ev;  // the caller's stream finishes only when every download has finished
    int rc = next_sync_event(&ev);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(ev, ctx->s_d2h));
    CUDA_TRY(cudaStreamWaitEvent(st, ev, 0));
  }
  CUDA_TRY(cudaEventRecord(ctx->ev_t1, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (trace) {
    for (auto& m : trace_marks) {
      float t = 0.f;
      cudaEventSynchronize(m.ev);
      cudaEventElapsedTime(&t, ctx->ev_t0, m.ev);
      fprintf(stderr, "[alr trace] %8.3f ms  chunk %2d  %s\n", t, m.chunk, m.what);
      cudaEventDestroy(m.ev);
    }
  }
  if (h_ctl->abort)
    return fail(ALR_ERR_CUDA, "internal: the producer/consumer launch timed out waiting for a dependency (watchdog); "
                              "results of this call are invalid");
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1));
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
    ctx->prof.ms_cmac_static = acc[kCatCmacStatic];
    ctx->prof.ms_ifft = acc[kCatIfft];
    ctx->prof.ms_mix = acc[kCatMix];
    ctx->prof.ms_other = acc[kCatOther];
    ctx->prof.ms_fused = acc[kCatFused];
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

int alr_visibilities(alr_context* ctx, const float* mix, int32_t n_channels, int64_t n_samples, double rate, double t_sti,
                     const double* fc, int32_t n_bands, double bw, int32_t n_sti_per_block, double tukey_alpha,
                     double* out, int mem_space, void* stream) {
  if (!ctx || !mix || !fc || !out) return fail(ALR_ERR_INVALID, "alr_visibilities: NULL argument");
  if (mem_space != ALR_MEM_HOST && mem_space != ALR_MEM_DEVICE) return fail(ALR_ERR_INVALID, "alr_visibilities: bad mem_space");
  if (n_channels < 1 || n_channels > 1024 || n_samples < 1 || n_bands < 1 || n_sti_per_block < 1 || !(rate > 0) || !(t_sti > 0))
    return fail(ALR_ERR_INVALID, "alr_visibilities: bad shape");
  const long long N = (long long)(rate * t_sti);  // int(rate_ * t), imaging.py:467
  if (N == 0) return fail(ALR_ERR_INVALID, "Not enough samples per time frame.");  // imaging.py:469
  if (N > 0x3fffffff) return fail(ALR_ERR_INVALID, "alr_visibilities: frame too long");
  const long long n_stf = n_samples / N, n_blocks = n_stf / n_sti_per_block;
  if (n_blocks < 1) return ALR_OK;  // nothing to write
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int C = n_channels;
  // modulated windows, float64: g[b][n] = tukey(n) * sum_{k in band b} exp(-2 pi i k n / N)
  std::vector<double> win((size_t)N);
  {
    // scipy.signal.windows.tukey(M, alpha, sym=True)
    const long long M = N;
    if (tukey_alpha <= 0.0 || M == 1) {
      for (long long n = 0; n < M; ++n) win[n] = 1.0;
    } else if (tukey_alpha >= 1.0) {
      for (long long n = 0; n < M; ++n) win[n] = 0.5 - 0.5 * cos(2.0 * M_PI * (double)n / (double)(M - 1));  // hann(M, sym=True)
    } else {
      const long long width = (long long)floor(tukey_alpha * (double)(M - 1) / 2.0);
      for (long long n = 0; n < M; ++n) {
        if (n <= width) win[n] = 0.5 * (1.0 + cos(M_PI * (-1.0 + 2.0 * (double)n / tukey_alpha / (double)(M - 1))));
        else if (n < M - width - 1) win[n] = 1.0;
        else win[n] = 0.5 * (1.0 + cos(M_PI * (-2.0 / tukey_alpha + 1.0 + 2.0 * (double)n / tukey_alpha / (double)(M - 1))));
      }
    }
  }
  std::vector<double2> g((size_t)n_bands * N);
  for (int b = 0; b < n_bands; ++b) {
    // bins stft_data[:, idx_start : idx_end + 1] with Python slice semantics (imaging.py:485-487)
    long long i0 = (long long)((fc[b] - 0.5 * bw) * (double)N / rate), i1 = (long long)((fc[b] + 0.5 * bw) * (double)N / rate) + 1;
    if (i0 < 0) i0 = std::max<long long>(i0 + N, 0);
    if (i1 < 0) i1 = std::max<long long>(i1 + N, 0);
    i0 = std::min(i0, N);
    i1 = std::min(i1, N);
    for (long long n = 0; n < N; ++n) {
      double re = 0.0, im = 0.0;
      for (long long k = i0; k < i1; ++k) {
        const double ph = -2.0 * M_PI * (double)((k * n) % N) / (double)N;
        re += cos(ph);
        im += sin(ph);
      }
      g[(size_t)b * N + n] = make_double2(win[n] * re, win[n] * im);
    }
  }
  const size_t off_g = 0, bytes_g = g.size() * sizeof(double2);
  const size_t off_s = align_up(off_g + bytes_g, 256), bytes_s = (size_t)n_stf * n_bands * C * sizeof(double2);
  const size_t off_v = align_up(off_s + bytes_s, 256), bytes_v = (size_t)n_blocks * n_bands * C * C * sizeof(double2);
  const size_t off_x = align_up(off_v + bytes_v, 256), bytes_x = mem_space == ALR_MEM_HOST ? (size_t)C * n_samples * sizeof(float) : 0;
  int rc = ctx->visbuf.ensure(off_x + bytes_x + 256);
  if (rc) return rc;
  char* base = (char*)ctx->visbuf.p;
  CUDA_TRY(cudaMemcpyAsync(base + off_g, g.data(), bytes_g, cudaMemcpyHostToDevice, st));
  const float* d_mix = mix;
  if (mem_space == ALR_MEM_HOST) {
    CUDA_TRY(cudaMemcpyAsync(base + off_x, mix, bytes_x, cudaMemcpyHostToDevice, st));
    d_mix = (const float*)(base + off_x);
  }
  if (n_stf > 0x7fffffffLL || (long long)n_bands * C > 65535)
    return fail(ALR_ERR_INVALID, "alr_visibilities: too many frames or band x channel pairs");
  k_vis_spectrum<<<dim3((unsigned)n_stf, (unsigned)(n_bands * C)), 64, 0, st>>>(d_mix, n_samples, C, (int)N, n_bands,
                                                                                (const double2*)(base + off_g),
                                                                                (double2*)(base + off_s));
  CUDA_TRY(cudaGetLastError());
  k_vis_outer<<<dim3((unsigned)n_blocks, (unsigned)n_bands), std::min(C * C, 256), 0, st>>>(
      (const double2*)(base + off_s), C, n_bands, n_sti_per_block, (double2*)(base + off_v));
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(out, base + off_v, bytes_v, mem_space == ALR_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return ALR_OK;
}

int alr_debug_plan_movers(const alr_event* events, int32_t n_events, int32_t mode, int64_t ring_bytes, int32_t lookahead,
                          int32_t n_slots, int32_t* header, int32_t* tasks, int64_t tasks_cap, int32_t* per_ir,
                          int64_t per_ir_cap) {
  if (!events || n_events < 1 || !header || !tasks || !per_ir || (mode != 1 && mode != 2))
    return fail(ALR_ERR_INVALID, "alr_debug_plan_movers: bad argument");
  const long long ring_slots = ring_bytes / ((int64_t)kP * sizeof(float2));
  std::vector<EvSize> sz(n_events);
  int n_ir = 0, n_w = 0, n_blk = 0, n_fo = 0, n_tasks = 0, n_sweep_ev = 0;
  for (int i = 0; i < n_events; ++i) {
    int rc = size_event(events[i], i, sz[i], ring_slots, mode, 0);
    if (rc) return rc;
    n_ir += sz[i].n_ir;
    n_w += sz[i].wband;
    n_blk += sz[i].n_blk;
    if (sz[i].fused) {
      n_fo += sz[i].n_ir;
      n_tasks += sz[i].n_ptask + sz[i].n_ctask;
      n_sweep_ev += sz[i].fused == 2;
    }
  }
  std::vector<EvDev> evs(n_events);
  std::vector<IrDev> irs(std::max(n_ir, 1));
  std::vector<float> wband(std::max(n_w, 1));
  std::vector<int2> lr(std::max(n_blk, 1));
  int ir_off = 0, w_off = 0, blk_off = 0;
  for (int i = 0; i < n_events; ++i) {
    int x_used = 0;
    plan_event_into(events[i], i, sz[i], evs[i], irs.data() + ir_off, ir_off, wband.data() + w_off, w_off, lr.data() + blk_off,
                    blk_off, &x_used);
    evs[i].fused = sz[i].fused;
    ir_off += sz[i].n_ir;
    w_off += sz[i].wband;
    blk_off += sz[i].n_blk;
  }
  header[0] = n_fo;
  header[1] = n_tasks;
  header[2] = (int)std::min<long long>(ring_slots, 0x7fffffff);
  header[3] = 0;
  for (int i = 0; i < n_events; ++i) header[3] += sz[i].fused != 0;
  if (n_fo == 0) return ALR_OK;
  if ((int64_t)n_tasks * 4 > tasks_cap || (int64_t)n_fo * 10 > per_ir_cap)
    return fail(ALR_ERR_INVALID, "alr_debug_plan_movers: output buffers too small (%d tasks, %d RIRs)", n_tasks, n_fo);
  std::vector<FusedTask> tk(n_tasks);
  std::vector<int2> pop(n_fo), need(n_fo);
  std::vector<int> prod(n_fo, -1), slot_off(kMaxSweepSlots + 1), slot_jobs(std::max(n_sweep_ev, 1));
  int slots_used = 0;
  bool ok;
  if (mode == 1) {
    ok = plan_fused(evs.data(), n_events, irs.data(), lr.data(), ring_slots, lookahead, tk.data(), n_tasks, pop.data(), need.data(), n_fo);
    for (int i = 0; i < n_fo; ++i) prod[i] = i;  // ordinal == production order
  } else {
    ok = plan_sweep(evs.data(), n_events, irs.data(), ring_slots, std::max(1, std::min(n_slots, kMaxSweepSlots)), tk.data(), n_tasks,
                    pop.data(), need.data(), prod.data(), n_fo, slot_off.data(), slot_jobs.data(), n_sweep_ev, &slots_used);
  }
  if (!ok) return fail(ALR_ERR_INVALID, "internal: inconsistent plan");
  header[4] = slots_used;
  for (int i = 0; i < n_tasks; ++i) {
    tasks[4 * i] = tk[i].type;
    tasks[4 * i + 1] = tk[i].ev;
    tasks[4 * i + 2] = tk[i].idx;
    tasks[4 * i + 3] = tk[i].sub;
  }
  std::vector<int> prod_index(n_fo, -1);
  for (int p2 = 0; p2 < n_fo; ++p2)
    if (prod[p2] >= 0) prod_index[prod[p2]] = p2;
  for (int e = 0; e < n_events; ++e) {
    if (!evs[e].fused) continue;
    for (int l = 0; l < evs[e].N; ++l) {
      const int fo = evs[e].fo0 + l;
      int32_t* o = per_ir + 10 * fo;
      o[0] = e; o[1] = l; o[2] = irs[evs[e].ir0 + l].hring; o[3] = evs[e].K * evs[e].C;
      o[4] = pop[fo].x; o[5] = pop[fo].y; o[6] = need[fo].x; o[7] = need[fo].y; o[8] = prod_index[fo];
      o[9] = (mode == 1) ? 0 : irs[evs[e].ir0 + l].xnb;
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
  EvSize z;
  int rc = size_event(*ev, 0, z);
  if (rc) return rc;
  std::vector<IrDev> v_irs(std::max(z.n_ir, 1));
  std::vector<float> v_w(std::max(z.wband, 1));
  std::vector<int2> v_lr(std::max(z.n_blk, 1));
  EvDev d;
  int x_used = 0;
  plan_event_into(*ev, 0, z, d, v_irs.data(), 0, v_w.data(), 0, v_lr.data(), 0, &x_used);
  header[0] = d.K;
  header[1] = d.B_valid;
  header[2] = d.B_out;
  header[3] = d.n_valid;
  header[4] = d.xlimit;
  header[5] = z.n_ir;
  header[6] = z.wband;
  header[7] = z.n_blk;
  if ((int64_t)z.n_ir * 6 > irs_cap || (int64_t)z.wband > wband_cap || (int64_t)z.n_blk * 2 > lrange_cap)
    return fail(ALR_ERR_INVALID, "alr_debug_plan: output buffers too small");
  if ((long long)x_used > z.xb) return fail(ALR_ERR_INVALID, "alr_debug_plan: X slot bound violated (%d > %lld)", x_used, z.xb);
  for (int i = 0; i < z.n_ir; ++i) {
    const IrDev& r = v_irs[i];
    int32_t* o = irs + 6 * i;
    o[0] = r.xb0; o[1] = r.xnb; o[2] = r.xslot; o[3] = r.woff; o[4] = r.jmin; o[5] = r.nrows;
  }
  if (z.wband > 0) memcpy(wband, v_w.data(), (size_t)z.wband * sizeof(float));
  for (int i = 0; i < z.n_blk; ++i) {
    lrange[2 * i] = v_lr[i].x;
    lrange[2 * i + 1] = v_lr[i].y;
  }
  return ALR_OK;
}

}  // extern "C"
