// alr_kernels.cuh — device descriptors and kernels of the renderer (sm_100a).
//
// Pipeline per chunk of events (each kernel cites the reference code it replaces):
//   k_ir_fft      RIR partitions -> block spectra (+ per-partition energy)            synthesize.py:103,298,425
//   k_ir_scale    per-IR normalisation scalar a_l                                      synthesize.py:404-428,560
//   k_x_fft       source blocks x cross-fade weight g_l -> spectra                     synthesize.py:148-181,299
//   k_cmac        Y[b,c] = sum_{l,j,k: xb0_l+j+k=b} X_l[j] * H_l[k,c]  (moving events)  synthesize.py:184-252
//   k_cmac_static the same contraction for one-IR events (regular block FIR)           synthesize.py:103
//   k_ifft_ola    inverse FFT + overlap-add + truncation + max|y|, sum|y| partials     synthesize.py:255-274,590
//   k_event_gain  apply_snr / db_to_multiplier scalars                                 synthesize.py:40-68,594-599
//   k_apply_gain  y *= gain (device buffers: only what k_mix does not cover)           synthesize.py:594,599
//   k_amb_*       mean|ambience| -> scale                                              synthesize.py:350-352
//   k_mix         scene = sum ambience*scale + sum events in [start,end)               synthesize.py:328-378
//   k_pcm16       (C,T) float mix -> (T,C) PCM_16 as Scene.generate writes it         core.py:1840-1847
// Before the chunk loop (f1): k_aug_pointwise, k_iir_pass / k_iir_combine, k_peak_partial / k_peak_final apply the
// linear event augmentations and the peak normalisation of Event.load_audio           event.py:530-536
#pragma once
#include "alr_fft.cuh"

namespace alr {

constexpr int kCtaThreads = kGroup * kGroupsPerCta;  // 256
#ifndef ALR_CMAC_CH
#define ALR_CMAC_CH 4
#endif
#ifndef ALR_IR_TASKS
#define ALR_IR_TASKS 8
#endif
#ifndef ALR_X_TASKS
#define ALR_X_TASKS 4
#endif
constexpr int kXTasks = ALR_X_TASKS;                 // source blocks transformed per FFT group of k_x_fft
constexpr int kIrTasks = ALR_IR_TASKS;               // RIR partitions transformed per FFT group of k_ir_fft
constexpr int kChanGroup = ALR_CMAC_CH;              // capsules per k_cmac thread / CTA
constexpr int kIfftCh = kGroupsPerCta;               // capsules per IFFT CTA (one FFT group each)
#ifndef ALR_IFFT_RUN
#define ALR_IFFT_RUN 16  // at P = 4096: 8 / 16 / 32 blocks per run -> k_ifft_ola 1.96 / 1.87 / 1.89 ms per benchmark step (a run re-does the
#endif                   // inverse transform of the block before it for the overlap tail: 1/16 extra work instead of 1/8)
constexpr int kRun = ALR_IFFT_RUN;                              // consecutive output blocks per IFFT CTA (tail kept in registers)
constexpr int kBinCtas = kP / kCtaThreads;           // CMAC CTAs per spectrum (each thread owns one bin)

enum { kGainEvent = 0, kGainNone = 1, kGainDry = 2, kGainPass = 3 };  // Pass: already rendered, only mixed

struct EvDev {
  const float* x;    // dry audio (Lx)
  const float* irs;  // RIR taps
  float* y;          // out (C, n_out)
  long long ir_stride_c, ir_stride_n;
  long long hslot0, xslot0, yslot0;  // first spectrum slot of this event in the chunk workspace
  int Lx, Lh, C, N, K;
  int n_out;    // samples per channel in y
  int n_valid;  // samples that carry signal; [n_valid, n_out) is zero filled
  int B_valid;  // ceil(n_valid / P): blocks with spectra
  int B_out;    // ceil(n_out / P)
  int xlimit;   // source samples >= xlimit are not needed / zero
  int ir0;      // first IrDev of this event
  int blk0;     // first entry of the per-output-block IR range table
  int moving, normalize, gain_mode;
  int mask_lo, mask_hi;  // taps outside [mask_lo, mask_hi) are treated as zero (dry / direct-path window)
  int parent;            // dry events: stats index of the parent event, else -1
  int stat;              // index into the stats array
  int part0, nparts;     // range of reduction partials written by k_ifft_ola
  int gain_from;         // k_apply_gain scales samples [gain_from, n_out) of every channel; [0, gain_from) is left to k_mix
  int dry_channel, dry_low, dry_high;
  const float* xnorm;    // optional scalar the dry audio is multiplied with (peak normalisation), or NULL
  double snr, ref_db;
  // moving events rendered by the persistent producer/consumer launch (alr_fused.cuh); 0 for everything else
  int small;             // 1: rendered by k_small_rir (effective RIR length <= P): no spectra in the workspace
  int fused;             // 1 / 2: H spectra live in the L2-resident ring (k_mov_fused / k_mov_sweep), a_l applied by the consumer
  int fo0;               // ordinal of the event's first RIR among the chunk's fused RIRs (ready / consumed counters)
  long long ecap0;       // first entry of the event's (RIR, capsule) tap-energy table
};

struct IrDev {
  int xb0;    // first source block in which IR l is active
  int xnb;    // number of active source blocks
  int xslot;  // event-local index of the first X spectrum of this IR (prefix sum of xnb)
  int woff;   // offset of this IR's cross-fade weight band
  int jmin;   // first STFT frame (row of the interpolation matrix) of the band
  int nrows;  // band length
  int hring;  // fused events: first spectrum slot of this RIR in the ring ([k][c] order), else 0
  int pad;
};

struct EvStat {
  double peak, mean_abs, gain, event_scale, a0;
  int nonfinite, dry_peak;
};

struct SceneDev {
  float* mix;
  short* pcm;  // optional (T, C) interleaved 16-bit copy of the mix
  int C;
  int n_amb;
  long long T;
  int amb0;  // first AmbDev
  int ev0;   // first entry in the scene's event list
  int nev;
};

struct AmbDev {
  const float* data;
  long long n;  // C*T
  double ref_db;
  float scale;
  int part0, nparts;
};

struct MixEv {
  float* y;
  const float* gain;  // non-NULL: y still holds the un-scaled render; k_mix applies (and stores) the event gain
  long long start, end;
  int n_out;
  int pad;
};

// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_segment(const int* __restrict__ prefix, int n, int idx) {
  // largest e in [0, n) with prefix[e] <= idx   (prefix has n + 1 entries, prefix[0] = 0)
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(prefix + mid) <= idx) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------------------
// k_ir_fft: one group of kGroup threads per RIR partition (event e, IR l, partition k, capsule c), kIrTasks
// consecutive partitions per group.
// Spectrum slot = hslot0 + (l*K + k)*C + c, i.e. layout [l][k][c][P] so that k_cmac reads the C capsules of one
// (l,k) contiguously.  Also writes the partition's energy sum(h^2) for normalize_irs: every WARP stores its own
// partial (kEnWarps floats per slot, summed in a fixed order by k_ir_scale). Round 1 reduced the warps through a
// shared-memory buffer that thread 0 read behind the transform's group barrier while the next task's writes were
// already under way in other warps; a single-buffered version of that mixed the energies of neighbouring tasks
// about every second run of a 40-event batch (racecheck blind), and the double-buffered fix was never explained.
// There is no shared reduction state any more: nothing is read that another warp may be rewriting.
// ALR_IRFFT_NT=2 (experiment, off): two partitions per pass through the transform share every derived twiddle (a third
// of the transform's floating-point instructions), but need 128 registers and 70 KB of shared memory, i.e. 2 CTAs per SM
// instead of 4: k_ir_fft 6.35 -> 8.40 ms per benchmark step (profiles/r02_micro_variants.txt). Latency hiding by
// occupancy is worth more here than the saved arithmetic.
// ALR_H_TILED: layout of the RIR spectra of one (RIR, partition): 0 = [capsule][P bins]; 1 = [256-bin tile][capsule][256 bins],
// so that the 4 capsules x 256 bins a multiply-accumulate CTA reads per item are one contiguous 8 KB piece instead of four
// 2 KB pieces 32 KB apart.
#ifndef ALR_H_TILED
#define ALR_H_TILED 0
#endif
#ifndef ALR_IRFFT_NT
#define ALR_IRFFT_NT 1
#endif
#ifndef ALR_IRFFT_MINB  // occupancy experiment, profiles/r01_irfft_occupancy.txt: 4 CTAs per SM (64 registers) is fastest
#define ALR_IRFFT_MINB 4
#endif
constexpr int kEnWarps = kGroup / 32;  // energy partials per spectrum slot
constexpr size_t kIrFftSmem = ALR_IRFFT_NT == 2 ? sizeof(FftSmem) * 2 * kGroupsPerCta : 0;  // dynamic part
__global__ void __launch_bounds__(kCtaThreads, ALR_IRFFT_MINB)
k_ir_fft(const EvDev* __restrict__ evs, int n_ev, const int* __restrict__ prefix, int n_tasks,
         const float2* __restrict__ tw, const float2* __restrict__ zeta, float2* __restrict__ hspec,
         float* __restrict__ hen) {
#if ALR_IRFFT_NT == 2
  extern __shared__ __align__(16) unsigned char ir_dyn[];  // 2 x FftSmem per group: beyond the 48 KB static limit at P = 4096
  FftSmem (*sm)[2] = reinterpret_cast<FftSmem (*)[2]>(ir_dyn);
#else
  __shared__ FftSmem sm[kGroupsPerCta];
#endif
  const int g = threadIdx.x / kGroup, t = threadIdx.x % kGroup, bar = 1 + g;
  // Each group transforms kIrTasks consecutive partitions: the event lookup (a chain of ~8 dependent loads) and the
  // descriptor reads are paid once per group instead of once per transform (consecutive tasks share the event).
  const int task0 = (blockIdx.x * kGroupsPerCta + g) * kIrTasks;
  if (task0 >= n_tasks) return;  // whole group leaves together
  int e = find_segment(prefix, n_ev, task0);
  int seg_lo = __ldg(prefix + e), seg_hi = __ldg(prefix + e + 1);
  const float2 zt = __ldg(zeta + t);
  int c = 0, k = 0, l = 0;
  bool fresh = true;  // (c, k, l) must be derived from the task index: first task of the group or a new event
  // one task: event / partition bookkeeping, the 16 taps of this thread, the warp's energy partial; returns the slot
  int stride256 = 256;
  auto load_task = [&](int task, float (&a)[16]) -> float2* {
    while (task >= seg_hi) {  // next event (empty segments are skipped)
      ++e;
      seg_lo = seg_hi;
      seg_hi = __ldg(prefix + e + 1);
      fresh = true;
    }
    const EvDev& ev = evs[e];
    if (fresh) {  // two integer divisions; the following tasks of the same event just count on
      int local = task - seg_lo;
      c = local % ev.C;
      local /= ev.C;
      k = local % ev.K;
      l = local / ev.K;
      fresh = false;
    } else if (++c == ev.C) {
      c = 0;
      if (++k == ev.K) {
        k = 0;
        ++l;
      }
    }
    const float* __restrict__ src = ev.irs + (long long)c * ev.ir_stride_c + (long long)l * ev.ir_stride_n;
    const int t0 = k * kP;
    const int lo = max(ev.mask_lo, t0), hi = min(min(ev.mask_hi, ev.Lh), t0 + kP);
    float en = 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int n = t0 + t + kGroup * r;
      a[r] = (n >= lo && n < hi) ? __ldg(src + n) : 0.f;
      en = fmaf(a[r], a[r], en);
    }
    const long long slot = ev.hslot0 + (long long)(l * ev.K + k) * ev.C + c;
    en = warp_sum(en);
    if ((t & 31) == 0) hen[slot * kEnWarps + (t >> 5)] = en;
#if ALR_H_TILED
    stride256 = ev.C * 256;
    return hspec + (ev.hslot0 + (long long)(l * ev.K + k) * ev.C) * kP + c * 256;
#else
    return hspec + slot * kP;
#endif
  };
#if ALR_IRFFT_NT == 2
  // two partitions per pass through the transform: they share every derived twiddle (a third of the transform's
  // floating-point instructions) and the group barriers
  for (int i = 0; i < kIrTasks; i += 2) {
    const int task = task0 + i;
    if (task >= n_tasks) break;
    float a[2][16];
    float2* spec[2];
    spec[0] = load_task(task, a[0]);
    if (i + 1 < kIrTasks && task + 1 < n_tasks) {
      spec[1] = load_task(task + 1, a[1]);  // (stride256 of the second task: same event assumed, see below)
    } else {
      spec[1] = nullptr;
#pragma unroll
      for (int r = 0; r < 16; ++r) a[1][r] = 0.f;
    }
    fwd_blocks_to_global<2>(a, zt, sm[g], tw, t, bar, spec, stride256);  // contains group barriers, ends with one
  }
#else
  for (int i = 0; i < kIrTasks; ++i) {
    const int task = task0 + i;
    if (task >= n_tasks) break;
    float a[16];
    float2* spec = load_task(task, a);
    fwd_block_to_global(a, zt, sm[g], tw, t, bar, spec, stride256);  // contains group barriers, ends with one
  }
#endif
}

// k_ir_scale: a_l = 1 / mean_c( sqrt(sum_t h_{l,c}^2) + tiny )  (normalize_irs on the (N, C, Lh) view), times 512
// for moving events (the un-normalised irfft of istft_overlap_synthesis, synthesize.py:267). One warp per IR.
__global__ void k_ir_scale(const EvDev* __restrict__ evs, int n_ev, const int* __restrict__ ir_prefix, int n_irs,
                           const float* __restrict__ hen, float* __restrict__ irscale, EvStat* __restrict__ stats) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_irs) return;
  const int e = find_segment(ir_prefix, n_ev, w);
  const EvDev& ev = evs[e];
  const int l = w - __ldg(ir_prefix + e);
  if (ev.fused || ev.small) return;  // a_l of fused events comes from the P-tasks of k_mov_fused / k_mov_sweep, k_small_rir has its own
  double a = 1.0;
  if (ev.gain_mode == kGainDry) {
    a = stats[ev.parent].a0;  // the parent's a_0: compute_dry_audio gets the normalised IRs (synthesize.py:608)
  } else if (ev.normalize) {
    double mean_e = 0.0;
    for (int c = 0; c < ev.C; ++c) {
      float s = 0.f;
      for (int k = lane; k < ev.K; k += 32) {
        const float* p = hen + (ev.hslot0 + (long long)(l * ev.K + k) * ev.C + c) * kEnWarps;
        float sk = 0.f;
#pragma unroll
        for (int w4 = 0; w4 < kEnWarps; ++w4) sk += p[w4];
        s += sk;
      }
      s = warp_sum(s);
      mean_e += sqrt((double)s) + 2.2250738585072014e-308;
    }
    mean_e /= ev.C;
    a = mean_e > 0.0 ? 1.0 / mean_e : 0.0;
    if (!(a < 3.0e38)) a = 0.0;  // all-zero IR: the reference yields exact zeros (0 / tiny)
  }
  if (lane == 0) {
    if (l == 0 && ev.gain_mode != kGainDry) stats[ev.stat].a0 = a;
    irscale[ev.ir0 + l] = (float)(ev.moving ? 512.0 * a : a);
  }
}

// k_x_fft: one group per (event, IR l, active source block j). Input sample t = (xb0+j)*P + n is
// x[t] * irscale_l * g_l(t), with g_l(t) = w[q,l] cos^2(pi p/256) + w[q+1,l] sin^2(pi p/256), q = t / 128,
// p = t % 128: the Hann-smoothed source-side cross-fade that is equivalent to the reference's STFT-domain
// interpolation (generate_interpolation_matrix + stft window, synthesize.py:120,148-181).
#ifndef ALR_XFFT_MINB
#define ALR_XFFT_MINB 4
#endif
__global__ void __launch_bounds__(kCtaThreads, ALR_XFFT_MINB)
k_x_fft(const EvDev* __restrict__ evs, int n_ev, const int* __restrict__ prefix, int n_tasks,
        const IrDev* __restrict__ irs, const float* __restrict__ wband, const float* __restrict__ irscale,
        const float2* __restrict__ tw, const float2* __restrict__ zeta, const float* __restrict__ win,
        float2* __restrict__ xspec) {
  __shared__ FftSmem sm[kGroupsPerCta];
  // cross-fade weights of the kP / 128 + 1 STFT frames a block touches, staged once per block: every thread needs 32 of them,
  // and as predicated global loads they were the kernel's main stall (long scoreboard 7 per issue, profiles/r01_xfft_v4.txt)
  constexpr int kFrames = kP / 128 + 1;
  __shared__ float s_w[kGroupsPerCta][kFrames + 1];
  const int g = threadIdx.x / kGroup, t = threadIdx.x % kGroup, bar = 1 + g;
  // kXTasks consecutive source blocks per group: the two lookups (event, then IR inside the event: ~15 dependent
  // loads) are done once and then advanced incrementally
  const int task0 = (blockIdx.x * kGroupsPerCta + g) * kXTasks;
  if (task0 >= n_tasks) return;
  int e = find_segment(prefix, n_ev, task0);
  int seg_lo = __ldg(prefix + e), seg_hi = __ldg(prefix + e + 1);
  int l = -1;  // IR index inside the event, -1: search
  const float2 zt = __ldg(zeta + t);
  for (int i = 0; i < kXTasks; ++i) {
    const int task = task0 + i;
    if (task >= n_tasks) break;
    while (task >= seg_hi) {
      ++e;
      seg_lo = seg_hi;
      seg_hi = __ldg(prefix + e + 1);
      l = -1;
    }
    const EvDev& ev = evs[e];
    const int local = task - seg_lo;
    // IR owning this X slot: largest l with xslot[l] <= local
    if (l < 0) {
      int lo = 0, hi = ev.N;
      while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (irs[ev.ir0 + mid].xslot <= local) lo = mid; else hi = mid;
      }
      l = lo;
    } else {
      while (l + 1 < ev.N && irs[ev.ir0 + l + 1].xslot <= local) ++l;
    }
    const IrDev ir = irs[ev.ir0 + l];
    const int j = local - ir.xslot;
    const int t0 = (ir.xb0 + j) * kP;
    // fused events: the C-tasks of k_mov_fused apply 512 a_l themselves (a_l is not known yet when this kernel runs)
    const float sc = (ev.fused ? 1.f : irscale[ev.ir0 + l]) * (ev.xnorm ? __ldg(ev.xnorm) : 1.f);
    const float* __restrict__ x = ev.x;
    // sin^2(pi p / 256) for the in-frame positions of this thread's samples: offset t + kGroup r inside the block, i.e.
    // p = t (+ 64 for odd r when kGroup == 64)
    float s_even = 0.f, s_odd = 0.f;
    if (ev.moving) {
      s_even = __ldg(win + (t & 127));
      s_odd = __ldg(win + ((t + 64) & 127));
    }
    if (ev.moving) {  // (uniform over the group)
      if (t <= kFrames) {
        const int q = (t0 >> 7) - ir.jmin + t;  // row of the weight band for frame (t0 >> 7) + t
        s_w[g][t] = (q >= 0 && q < ir.nrows && t < kFrames) ? __ldg(wband + ir.woff + q) : 0.f;
      }
      group_sync(bar);  // the previous block's reads of s_w finished before its transform's first barrier
    }
    float a[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int n = t0 + t + kGroup * r;
      float v = (n < ev.xlimit) ? __ldg(x + n) : 0.f;
      float gw = sc;
      if (ev.moving) {
        const int f = (t + kGroup * r) >> 7;  // STFT frame of sample n, relative to the block's first
        const float w0 = s_w[g][f], w1 = s_w[g][f + 1];
        gw = sc * fmaf(w1 - w0, (kGroup == 64 && (r & 1)) ? s_odd : s_even, w0);  // w0 (1 - s) + w1 s
      }
      a[r] = v * gw;
    }
    fwd_block_to_global(a, zt, sm[g], tw, t, bar, xspec + (ev.xslot0 + local) * kP);  // ends with a group barrier
  }
}

// k_cmac: Y[b,c] = sum over IRs l, source blocks j of l and partitions k with xb0_l + j + k = b of X_l[j] * H_l[k,c].
// One CTA per (event, kCmacRuns consecutive runs of kGm output blocks, group of 4 capsules, 256 bins); a thread owns ONE bin
// of 4 capsules for a whole run: 8 x 4 complex accumulators in registers.
//  * The work of a run is a flat sequence of items (IR l, partition k), LISTED in shared memory before the pipeline starts
//    (CmacItem: 16 bytes, one thread per RIR builds its items, warp-shuffle scan for the offsets; windows of kMaxHeads
//    RIRs, passes of kMaxItems items for dense trajectories). The steady state of an item is one broadcast LDS.128 of its
//    record, the H copy of the item kStages-1 ahead and the multiply-accumulates: ~87 instructions for 44 FFMAs; round 1's
//    producer / consumer ITERATORS over per-RIR headers cost 172 (profiles/r02_micro_variants.txt).
//  * H_l[k, c0..c0+3][256 bins] is copied with 16-byte cp.async.cg (LDGSTS.128, L1 bypassed) into a kStages-deep ring; a
//    warp owns its 4 x 32-bin slice of every stage, so cp.async.wait_group + __syncwarp is the only synchronisation: no
//    CTA barriers in the pipeline, and it does not drain at IR boundaries. (Register prefetching could not cover the DRAM
//    latency with the 16 warps/SM that 64 accumulator registers allow: profiles/r01c_cmac_*.)
//  * Each H value is used for every output block of the run it contributes to (<= 8 x 4 FFMA per 8 bytes).
//  * The source spectra X_l[j] an item touches for the first time are pulled into L1 with prefetch.global.L1 when its H
//    copy is issued (kStages-1 items early) and read through L1 at row xrow + s.
constexpr int kG = 8;       // output blocks per CTA (k_cmac_static)
#ifndef ALR_CMAC_G
#define ALR_CMAC_G 8
#endif
#ifndef ALR_CMAC_STAGES
#define ALR_CMAC_STAGES 5
#endif
constexpr int kGm = ALR_CMAC_G;            // output blocks per CTA (k_cmac); even
#ifndef ALR_CMAC_ORDER
#define ALR_CMAC_ORDER 0
#endif
#ifndef ALR_CMAC_RUNS
#define ALR_CMAC_RUNS 4
#endif
constexpr int kCmacRuns = ALR_CMAC_RUNS;  // consecutive runs of kGm output blocks one k_cmac CTA works through
constexpr int kStages = ALR_CMAC_STAGES;  // cp.async ring depth: 5 x 4 capsules x 256 threads x 8 B = 40 KB

// ---- packed fp32 pairs (Blackwell FFMA2, experiment): `fma.rn.f32x2` does two FMAs per instruction on a 64-bit
// register pair. A complex multiply-accumulate
//   acc += x * h   is   acc = fma2({x.re, x.re}, {h.re, h.im}, acc);  acc = fma2({-x.im, x.im}, {h.im, h.re}, acc)
// i.e. 2 instructions instead of 4, with the same per-component operation order (bit-identical results).
// Measured (tools/micro/ffma2_bench.cu, profiles/r01_ffma2.txt): plain 3-register FFMA already reaches 120 of the 128
// FMA/clk/SM on B200, FFMA2 123 — packing saves issue slots, not pipe time. In k_cmac_static the operand packing
// (MOVs) and the extra live registers (spills at the 128-register cap) made it SLOWER (2.33 -> 2.86 ms), so it is off.
#ifndef ALR_FFMA2
#define ALR_FFMA2 0
#endif
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 unpack2(f32x2 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One (RIR l, partition k) item of a k_cmac CTA, built into shared memory before the pipeline runs (16 bytes, read with
// one broadcast LDS.128): rows are in units of kP float2 elements.
struct __align__(16) CmacItem {
  int hrow;  // H_l[k][c0] relative to the event's first spectrum row of capsule c0: (l * K + k) * C
  int xrow;  // X_l[j] of output 0 relative to the event's first source row: xslot_l + d0 - k (output s reads row xrow + s)
  int mask;  // bit s: output s of the run takes this item (0 <= s + d0 - k < xnb, s < nb); bits 8.. see pfmask
  int pfrow; // the one source row no earlier item of the RIR has touched (prefetched with the item's H copy)
};
constexpr int kMaxHeads = 64;   // RIRs per list-building window (one thread each)
constexpr int kMaxItems = 448;  // list capacity (7 KB); longer windows are worked through in passes
#ifndef ALR_CMAC_PF
#define ALR_CMAC_PF 1
#endif

struct CmacSmem {
  float2 ring[kStages][kChanGroup][kCtaThreads];
  CmacItem items[kMaxItems];
  int wtot[2][2];  // [window parity][warp]: totals of the two scanning warps
};
constexpr size_t kCmacSmem = sizeof(CmacSmem);  // dynamic shared memory of k_cmac / k_cmac_both (beyond 48 KB with deeper rings)

// `vb` is the CTA's index among the mover CTAs (blockIdx.x for k_cmac, the de-interleaved index for k_cmac_both)
__device__ __forceinline__ void cmac_mover_cta(CmacSmem& sh, int vb, const EvDev* __restrict__ evs, int n_ev,
                                               const int* __restrict__ prefix, const IrDev* __restrict__ irs,
                                               const int2* __restrict__ lrange, const float2* __restrict__ xspec,
                                               const float2* __restrict__ hspec, float2* __restrict__ yspec) {
  float2 (&ring)[kStages][kChanGroup][kCtaThreads] = sh.ring;
  CmacItem (&items)[kMaxItems] = sh.items;
  int (&s_wtot)[2][2] = sh.wtot;
  const int e = find_segment(prefix, n_ev, vb);
  const EvDev& ev = evs[e];
  int local = vb - __ldg(prefix + e);
  const int ncg = (ev.C + kChanGroup - 1) / kChanGroup;
#if ALR_CMAC_ORDER == 1
  // Experiment (off): CTA order [256-bin tile][run group][capsule group], so that CTAs starting within microseconds of each
  // other share what they read (capsule groups of a run: the source spectra; neighbouring runs: the RIRs straddling them).
  // No gain (4.26-4.29 ms at 2-4 runs per CTA, 4.53 at 1): with the default order k_cmac already reads only 1.11x its
  // unique bytes from DRAM (20.8 GB per step for 16.7 GB of RIR spectra + 2 GB of source spectra) at 5.3 TB/s.
  const int nrg = ((ev.B_valid + kGm - 1) / kGm + kCmacRuns - 1) / kCmacRuns;
  const int cg = local % ncg;
  local /= ncg;
  const int run0 = (local % nrg) * kCmacRuns;
  const int br = local / nrg;
#else
  const int br = local % kBinCtas;
  local /= kBinCtas;
  const int cg = local % ncg;
  const int run0 = (local / ncg) * kCmacRuns;
#endif
  const int c0 = cg * kChanGroup;
  const int nc = min(kChanGroup, ev.C - c0);
  const int tid = threadIdx.x;
  const int bin = br * kCtaThreads + tid;
  const int K = ev.K, C = ev.C;
  const IrDev* __restrict__ irp = irs + ev.ir0;
  const float2* __restrict__ xbase = xspec + ev.xslot0 * kP + bin;
  const float2* const ring_t = &ring[0][0][tid];  // this thread's column: stage stride 4*256, capsule stride 256 elements
  // H staging: 16-byte cp.async.cg copies (the north star's "coalesced float4 loads" of the spectra; L1 bypassed for data
  // that is read once per CTA). The warp's 4 x 32-bin slice of an item is 64 pieces of 16 bytes, two per lane (capsule =
  // piece / 16, bin pair = piece % 16), so a ring stage is warp-private: cp.async.wait_group + __syncwarp is the only
  // synchronisation of the pipeline, no CTA barriers.
  const int piece = (tid & ~31) + (tid & 15) * 2, pc = (tid >> 4) & 1;  // this lane's bin pair and its first capsule
#if ALR_H_TILED
  const float2* const hsrc = hspec + ev.hslot0 * kP + (long long)br * (ev.C * 256) + (c0 + pc) * 256 + piece;
  constexpr int kCapStride = 256;  // capsules of one 256-bin tile are adjacent: an item is ONE contiguous 8 KB read
#else
  const float2* const hsrc = hspec + (ev.hslot0 + c0 + pc) * kP + br * kCtaThreads + piece;
  constexpr int kCapStride = kP;
#endif
  float2* const rdst = &ring[0][pc][piece];
  const bool cp0 = pc < nc, cp1 = pc + 2 < nc;

  // kCmacRuns consecutive runs per CTA, one after the other: an RIR whose outputs straddle a run boundary (40 % of the
  // (RIR, partition) items with runs of 8 blocks) is read again by the next run. As separate CTAs the two reads are a CTA
  // lifetime apart (~140 MB of other traffic: both miss L2, k_cmac read 1.4x its spectra from DRAM); in one CTA the second
  // read follows the first within a few items and hits L2 (4.93 -> 4.73 ms per benchmark step).
  int wpar = 0;
  for (int run = run0; run < run0 + kCmacRuns && run * kGm < ev.B_valid; ++run) {
    const int b0 = run * kGm;
    const int nb = min(kGm, ev.B_valid - b0);
    const int lmin = lrange[ev.blk0 + b0].x, lmax = lrange[ev.blk0 + b0 + nb - 1].y;
    float2 acc[kGm][kChanGroup];
#pragma unroll
    for (int s = 0; s < kGm; ++s)
#pragma unroll
      for (int c = 0; c < kChanGroup; ++c) acc[s][c] = make_float2(0.f, 0.f);

    for (int l0 = lmin; l0 <= lmax; l0 += kMaxHeads) {
      // ---- the window's RIRs, one per thread: partitions k_lo..k_hi contribute to the run
      const int n_heads = min(kMaxHeads, lmax - l0 + 1);
      int k_lo = 0, cnt = 0, d0 = 0, xnb = 0, xslot = 0;
      if (tid < n_heads) {
        const IrDev ir = irp[l0 + tid];
        d0 = b0 - ir.xb0;  // source block of output s and partition k: j = s + d0 - k, valid for 0 <= j < xnb
        xnb = ir.xnb;
        xslot = ir.xslot;
        k_lo = max(0, d0 - xnb + 1);
        cnt = xnb > 0 ? max(0, min(K - 1, d0 + nb - 1) - k_lo + 1) : 0;
      }
      int off = 0;  // exclusive prefix of cnt over the window's threads (warps 0 and 1)
      if (tid < kMaxHeads) {
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, d);
          if ((tid & 31) >= d) incl += v;
        }
        off = incl - cnt;
        if ((tid & 31) == 31) s_wtot[wpar][tid >> 5] = incl;
      }
      __syncthreads();  // s_wtot written; the previous window's list is consumed
      const int total = s_wtot[wpar][0] + s_wtot[wpar][1];
      if (tid >= 32) off += s_wtot[wpar][0];
      wpar ^= 1;  // the next window's totals go to the other pair: a thread may still be reading these
      for (int pass = 0; pass < total; pass += kMaxItems) {
        if (pass > 0) __syncthreads();  // previous pass consumed
        if (cnt > 0) {
          const int i1 = min(off + cnt, pass + kMaxItems);
          for (int i = max(off, pass); i < i1; ++i) {
            const int k = k_lo + (i - off), jb = d0 - k;
            const int s_lo = max(0, -jb), s_hi = min(nb, xnb - jb);
            CmacItem it;
            it.hrow = ((l0 + tid) * K + k) * C;
            it.xrow = xslot + jb;
            it.mask = ((1 << s_hi) - 1) & ~((1 << s_lo) - 1);
            it.pfrow = it.xrow + s_lo;
            if (k == k_lo) it.mask |= (it.mask & (it.mask - 1)) << 8;  // entering the RIR: every other row is new as well
            items[i - pass] = it;
          }
        }
        __syncthreads();
        const int n_items = min(kMaxItems, total - pass);

        // ---- pipeline: the producer side runs kStages-1 items ahead of the consumer side in the same thread
        auto produce = [&](int i, int stage) {
          if (i < n_items) {
#if ALR_CMAC_PF
            const CmacItem it = items[i];
            const float2* src = hsrc + (long long)it.hrow * kP;
#else
            const float2* src = hsrc + (long long)items[i].hrow * kP;
#endif
            float2* dst = rdst + stage * (kChanGroup * kCtaThreads);
            if (cp0) cp_async16(dst, src);
            if (cp1) cp_async16(dst + 2 * kCtaThreads, src + 2 * kCapStride);
#if ALR_CMAC_PF
            // the source spectra the item reads for the first time go to L1 now, kStages-1 items before they are used
            prefetch_l1(xbase + (long long)it.pfrow * kP);
            for (int m = it.mask >> 8; m; m &= m - 1) prefetch_l1(xbase + (long long)(it.xrow + __ffs(m) - 1) * kP);
#endif
          }
          cp_async_commit();  // (possibly empty) group: keeps the group count in step with the item count
        };
#pragma unroll
        for (int i = 0; i < kStages - 1; ++i) produce(i, i);
        int stage = 0;
        for (int i = 0; i < n_items; ++i) {
          cp_async_wait<kStages - 2>();  // this lane's pieces of item i have landed ...
          __syncwarp();                  // ... and the other lanes'; every lane is done with the previous stage
          const float2* src = ring_t + stage * (kChanGroup * kCtaThreads);
          float2 h[kChanGroup];
#pragma unroll
          for (int c = 0; c < kChanGroup; ++c) h[c] = (c < nc) ? src[c * kCtaThreads] : make_float2(0.f, 0.f);
          produce(i + kStages - 1, stage == 0 ? kStages - 1 : stage - 1);  // refill the slot consumed one iteration ago
          const CmacItem it = items[i];
          const float2* cxq = xbase + (long long)it.xrow * kP;  // X of output s is cxq[s * kP]
          // The validity test is warp-uniform. ptxas if-converts a single 16-FFMA body into predicated code, which
          // still costs an issue slot per predicated-off FFMA (45 % of them for moving events, profiles/r01d_cmac.txt),
          // so the bodies are PAIRS of output blocks (32 FFMA): large enough to stay real basic blocks behind a
          // uniform branch. A block of a pair that is itself out of range gets a zero source value.
#pragma unroll
          for (int s = 0; s < kGm; s += 2) {
            if (it.mask & (3 << s)) {
              float2 x0 = make_float2(0.f, 0.f), x1 = make_float2(0.f, 0.f);
              if (it.mask & (1 << s)) x0 = __ldg(cxq + s * kP);
              if (it.mask & (2 << s)) x1 = __ldg(cxq + (s + 1) * kP);
#pragma unroll
              for (int c = 0; c < kChanGroup; ++c) {
                acc[s][c].x = fmaf(x0.x, h[c].x, acc[s][c].x);
                acc[s][c].x = fmaf(-x0.y, h[c].y, acc[s][c].x);
                acc[s][c].y = fmaf(x0.x, h[c].y, acc[s][c].y);
                acc[s][c].y = fmaf(x0.y, h[c].x, acc[s][c].y);
                acc[s + 1][c].x = fmaf(x1.x, h[c].x, acc[s + 1][c].x);
                acc[s + 1][c].x = fmaf(-x1.y, h[c].y, acc[s + 1][c].x);
                acc[s + 1][c].y = fmaf(x1.x, h[c].y, acc[s + 1][c].y);
                acc[s + 1][c].y = fmaf(x1.y, h[c].x, acc[s + 1][c].y);
              }
            }
          }
          stage = (stage + 1 == kStages) ? 0 : stage + 1;
        }
        cp_async_wait<0>();
      }
    }
#pragma unroll
    for (int s = 0; s < kGm; ++s)
      if (s < nb)
#pragma unroll
        for (int c = 0; c < kChanGroup; ++c)
          if (c < nc) ALR_SPEC_STORE(yspec + (ev.yslot0 + (long long)(b0 + s) * C + c0 + c) * kP + bin, acc[s][c]);
  }  // run
}

__global__ void __launch_bounds__(kCtaThreads, 2)
k_cmac(const EvDev* __restrict__ evs, int n_ev, const int* __restrict__ prefix, const IrDev* __restrict__ irs,
       const int2* __restrict__ lrange, const float2* __restrict__ xspec, const float2* __restrict__ hspec,
       float2* __restrict__ yspec) {
  extern __shared__ __align__(16) unsigned char cmac_dyn[];
  CmacSmem& sh = *reinterpret_cast<CmacSmem*>(cmac_dyn);
  cmac_mover_cta(sh, blockIdx.x, evs, n_ev, prefix, irs, lrange, xspec, hspec, yspec);
}

// k_cmac_static: the same contraction for STATIC events (one IR, every source block active), where it is a plain
// block-FIR  Y[b,c] = sum_k X[b-k] * H[k,c]  and completely regular: all kG outputs of the run are live in every
// partition step, so there are no per-IR headers, iterators or validity branches. Static events are 40 % of the
// multiply-accumulates and 72 % of the CTAs of the benchmark workload (and the common case in DCASE-style scenes).
// One CTA per (event, run of 8 output blocks, 4 capsules, 256 bins), a thread owns one bin:
//  * source window W[s] = X[b0 + s - k] in registers; the partition loop is unrolled 8x so that the window slides
//    by register RENAMING (W[(s - r) & 7]), one new row per step, requested one step ahead;
//  * H[k, c0..c0+3] double-buffered by name (hA / hB), requested one step ahead (static-event H is L2 resident:
//    96 rows per event shared by all of its CTAs);
//  * rows outside [0, xnb) are zeros (start of the signal / beyond its end), so the 128 FFMA of a step are
//    unconditional: ~145 instructions per step instead of ~290 in the generic kernel.
// Tile per thread: 1 bin x 4 capsules x 8 blocks at 2 CTAs/SM measured best (3.78 ms per benchmark step); 2 capsules
// (77 registers, 3 CTAs/SM) 4.19 ms, 1 capsule (4 CTAs/SM) 4.49 ms: the extra loads per FFMA outweigh the occupancy;
// 1 CTA/SM 6.14 ms, i.e. the kernel is latency-bound (profiles/r01_cmac_variants.txt).
#ifndef ALR_STATIC_CH
#define ALR_STATIC_CH 4
#endif
#ifndef ALR_STATIC_OCC
#define ALR_STATIC_OCC 2
#endif
constexpr int kStaticCh = ALR_STATIC_CH;  // capsules per thread in k_cmac_static
// A 16-block x 2-capsule tile for long RIRs (K >= 12 partitions; fewer re-reads of the RIR spectra, which every run of output
// blocks reads in full: 80 GB of L2 traffic per launch on the em64 shape) was SLOWER: C4 k_cmac_static 15.4 -> 19.1 ms.
#ifndef ALR_STATIC_RUNS
#define ALR_STATIC_RUNS 8
#endif
constexpr int kStaticRuns = ALR_STATIC_RUNS;  // consecutive runs of kG output blocks per k_cmac_static CTA
__device__ __forceinline__ void cmac_static_cta(int vb, const EvDev* __restrict__ evs, int n_ev,
                                                const int* __restrict__ prefix, const IrDev* __restrict__ irs,
                                                const float2* __restrict__ xspec, const float2* __restrict__ hspec,
                                                float2* __restrict__ yspec) {
  const int e = find_segment(prefix, n_ev, vb);
  const EvDev& ev = evs[e];
  int local = vb - __ldg(prefix + e);
  const int br = local % kBinCtas;
  local /= kBinCtas;
  const int ncg = (ev.C + kStaticCh - 1) / kStaticCh;
  const int cg = local % ncg;
  const int run0 = (local / ncg) * kStaticRuns;
  const int c0 = cg * kStaticCh;
  const int nc = min(kStaticCh, ev.C - c0);
  const int bin = br * kCtaThreads + threadIdx.x;
  const int K = ev.K, C = ev.C;
  const int xnb = irs[ev.ir0].xnb;
  const long long kstride = (long long)C * kP;
  const float2* __restrict__ xbase = xspec + ev.xslot0 * kP + bin;
#if ALR_H_TILED
  const float2* __restrict__ hbase = hspec + ev.hslot0 * kP + (long long)br * (ev.C * 256) + c0 * 256 + threadIdx.x;
  constexpr int kCapStride = 256;
#else
  const float2* __restrict__ hbase = hspec + (ev.hslot0 + c0) * kP + bin;
  constexpr int kCapStride = kP;
#endif
  const float2 zero = make_float2(0.f, 0.f);
  auto load_x = [&](int j) -> float2 { return (j >= 0 && j < xnb) ? __ldg(xbase + (long long)j * kP) : zero; };
  auto load_h = [&](int k, float2 (&h)[kStaticCh]) {
    if (k < K) {
#pragma unroll
      for (int c = 0; c < kStaticCh; ++c) h[c] = (c < nc) ? __ldg(hbase + k * kstride + (long long)c * kCapStride) : zero;
    }
  };
  // kStaticRuns consecutive runs per CTA: the event lookup and descriptor reads (a chain of ~10 dependent loads in front of
  // only K = 6 steps of arithmetic at P = 4096) are paid once. 1 / 2 / 4 / 8 / 16 runs: k_cmac_static 1.59 / 1.35 / 1.24 /
  // 1.19 / 1.18 ms per benchmark step, C2 8.01 -> 7.34 ms, C1 9.14 -> 8.36 ms (profiles/r02_micro_variants.txt).
  for (int run = run0; run < run0 + kStaticRuns && run * kG < ev.B_valid; ++run) {
  const int b0 = run * kG;
  const int nb = min(kG, ev.B_valid - b0);
#if ALR_FFMA2
  f32x2 acc[kG][kStaticCh];
#pragma unroll
  for (int s = 0; s < kG; ++s)
#pragma unroll
    for (int c = 0; c < kStaticCh; ++c) acc[s][c] = 0ull;
#else
  float2 acc[kG][kStaticCh];
#pragma unroll
  for (int s = 0; s < kG; ++s)
#pragma unroll
    for (int c = 0; c < kStaticCh; ++c) acc[s][c] = zero;
#endif
  float2 W[kG], hA[kStaticCh], hB[kStaticCh];
#pragma unroll
  for (int s = 0; s < kG; ++s) W[s] = load_x(b0 + s);  // window of partition 0
#pragma unroll
  for (int c = 0; c < kStaticCh; ++c) hA[c] = hB[c] = zero;
  load_h(0, hA);
  // New source rows are requested TWO steps ahead into two registers that alternate by step parity: the copy into
  // the window then always reads a value loaded a full step earlier. (Copying a register in the step that loads it
  // waits for the load: 38 % of all stall samples sat on that single MOV, profiles/r01_cmac_static_v1.txt.)
  float2 xe = load_x(b0 - 2);  // row of step 2 (even steps)
  float2 xo = load_x(b0 - 1);  // row of step 1 (odd steps)
  for (int k0 = 0; k0 < K; k0 += 8) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int k = k0 + r;
      if (k < K) {
        // W[(s - r) & 7] == X[b0 + s - k]; the H request for step k + 1 goes out before this step's arithmetic
        if (r & 1) load_h(k + 1, hA); else load_h(k + 1, hB);
        const float2 (&h)[kStaticCh] = (r & 1) ? hB : hA;
#if ALR_FFMA2
        f32x2 hd[kStaticCh], hs[kStaticCh];  // {re, im} and {im, re}
#pragma unroll
        for (int c = 0; c < kStaticCh; ++c) {
          hd[c] = pack2(h[c].x, h[c].y);
          hs[c] = pack2(h[c].y, h[c].x);
        }
#pragma unroll
        for (int s = 0; s < kG; ++s) {
          const float2 xv = W[(s - r) & 7];
          const f32x2 xx = pack2(xv.x, xv.x), xy = pack2(-xv.y, xv.y);
#pragma unroll
          for (int c = 0; c < kStaticCh; ++c) {
            acc[s][c] = fma2(xx, hd[c], acc[s][c]);
            acc[s][c] = fma2(xy, hs[c], acc[s][c]);
          }
        }
#else
#pragma unroll
        for (int s = 0; s < kG; ++s) {
          const float2 xv = W[(s - r) & 7];
#pragma unroll
          for (int c = 0; c < kStaticCh; ++c) {
            acc[s][c].x = fmaf(xv.x, h[c].x, acc[s][c].x);
            acc[s][c].x = fmaf(-xv.y, h[c].y, acc[s][c].x);
            acc[s][c].y = fmaf(xv.x, h[c].y, acc[s][c].y);
            acc[s][c].y = fmaf(xv.y, h[c].x, acc[s][c].y);
          }
        }
#endif
        // the slot of output 7 of this step becomes output 0 of step k + 1: row b0 - (k + 1), requested during
        // step k - 1; its register is then free for the row of step k + 3
        if (r & 1) {  // k + 1 is even
          W[(7 - r) & 7] = xe;
          xe = load_x(b0 - k - 3);
        } else {
          W[(7 - r) & 7] = xo;
          xo = load_x(b0 - k - 3);
        }
      }
    }
  }
#pragma unroll
  for (int s = 0; s < kG; ++s)
    if (s < nb)
#pragma unroll
      for (int c = 0; c < kStaticCh; ++c)
#if ALR_FFMA2
        if (c < nc) yspec[(ev.yslot0 + (long long)(b0 + s) * C + c0 + c) * kP + bin] = unpack2(acc[s][c]);
#else
        if (c < nc) ALR_SPEC_STORE(yspec + (ev.yslot0 + (long long)(b0 + s) * C + c0 + c) * kP + bin, acc[s][c]);
#endif
  }  // run
}

__global__ void __launch_bounds__(kCtaThreads, ALR_STATIC_OCC)
k_cmac_static(const EvDev* __restrict__ evs, int n_ev, const int* __restrict__ prefix, const IrDev* __restrict__ irs,
              const float2* __restrict__ xspec, const float2* __restrict__ hspec, float2* __restrict__ yspec) {
  cmac_static_cta(blockIdx.x, evs, n_ev, prefix, irs, xspec, hspec, yspec);
}

// k_cmac_both: the CTAs of k_cmac and k_cmac_static in ONE grid, interleaved at their ratio. Launched one after the other
// the two kernels leave complementary halves of the SM idle: k_cmac streams spectra (DRAM- and L2-bound pipeline), a
// k_cmac_static CTA is 6 short steps behind a long chain of dependent loads (48 % long-scoreboard stalls, 38 % of the DRAM
// bandwidth). Both are 2 CTAs per SM at 128 registers, so a mixed grid makes the typical SM hold one of each.
// Every `period`-th CTA is a static one until the n_static are used up: CTA b is static iff (b + 1) % period == 0 and
// (b + 1) / period <= n_static.
__global__ void __launch_bounds__(kCtaThreads, 2)
k_cmac_both(const EvDev* __restrict__ evs, int n_ev, const int* __restrict__ prefix_m, const int* __restrict__ prefix_s,
            int n_static, int period, const IrDev* __restrict__ irs, const int2* __restrict__ lrange,
            const float2* __restrict__ xspec, const float2* __restrict__ hspec, float2* __restrict__ yspec) {
  extern __shared__ __align__(16) unsigned char cmac_dyn[];
  CmacSmem& sh = *reinterpret_cast<CmacSmem*>(cmac_dyn);
  const int q = (blockIdx.x + 1) / period;
  if (q * period == blockIdx.x + 1 && q <= n_static)
    cmac_static_cta(q - 1, evs, n_ev, prefix_s, irs, xspec, hspec, yspec);
  else
    cmac_mover_cta(sh, blockIdx.x - min(q, n_static), evs, n_ev, prefix_m, irs, lrange, xspec, hspec, yspec);
}

// k_ifft_ola: one CTA per (event, group of 4 capsules, run of kRun output blocks); group g handles capsule c0+g.
// Inverse transform of Y[b]; the real part of element e is sample e of the block, the imaginary part its overlap
// tail (sample P + e), which the SAME thread adds to block b+1 — the tail never leaves registers. Scale 1/P,
// truncate to n_valid, zero fill up to n_out (pad_or_truncate_audio, utils.py:667), reduce max|y| and sum|y|.
#ifndef ALR_IFFT_MINB
#define ALR_IFFT_MINB 3  // 80 registers, 3 CTAs per SM: 2.11 ms per benchmark step (2 CTAs at 127 registers 2.20, 4 CTAs at 64 with spills 2.30)
#endif
__global__ void __launch_bounds__(kCtaThreads, ALR_IFFT_MINB)
k_ifft_ola(const EvDev* __restrict__ evs, int n_ev, const int* __restrict__ prefix, const float2* __restrict__ tw,
           const float2* __restrict__ zeta, const float2* __restrict__ yspec, float2* __restrict__ partials) {
  __shared__ FftSmem sm[kGroupsPerCta];
  __shared__ float s_max[kCtaThreads / 32], s_sum[kCtaThreads / 32];
  const int g = threadIdx.x / kGroup, t = threadIdx.x % kGroup, bar = 1 + g;
  const int e = find_segment(prefix, n_ev, blockIdx.x);
  const EvDev& ev = evs[e];
  int local = blockIdx.x - __ldg(prefix + e);
  const int nruns = (ev.B_out + kRun - 1) / kRun;
  const int run = local % nruns;
  const int cg = local / nruns;
  const int c = cg * kIfftCh + g;
  float vmax = 0.f, vsum = 0.f;
  if (c < ev.C) {
    const float inv = 1.0f / kP;
    const float2 zt = __ldg(zeta + t);
    float* __restrict__ y = ev.y + (long long)c * ev.n_out;
    float tail[kM3][kR3];
#pragma unroll
    for (int m = 0; m < kM3; ++m)
#pragma unroll
      for (int k = 0; k < kR3; ++k) tail[m][k] = 0.f;
    const int b0 = run * kRun;
    const int b1 = min(b0 + kRun, ev.B_out);
    for (int b = max(b0 - 1, 0); b < b1; ++b) {
      float2 o[kM3][kR3];
      if (b < ev.B_valid) {
        inv_block_from_global(yspec + (ev.yslot0 + (long long)b * ev.C + c) * kP, zt, sm[g], tw, t, bar, o);
      } else {
#pragma unroll
        for (int m = 0; m < kM3; ++m)
#pragma unroll
          for (int k = 0; k < kR3; ++k) o[m][k] = make_float2(0.f, 0.f);
      }
      if (b >= b0) {
#pragma unroll
        for (int k = 0; k < kR3; ++k)
#pragma unroll
          for (int m = 0; m < kM3; ++m) {
            const int n = b * kP + t + kGroup * m + 256 * k;
            float a = fmaf(o[m][k].x, inv, tail[m][k]);
            if (n >= ev.n_valid) a = 0.f;
            if (n < ev.n_out) y[n] = a;
            vmax = fmaxf(vmax, fabsf(a));
            vsum += fabsf(a);
          }
      }
#pragma unroll
      for (int m = 0; m < kM3; ++m)
#pragma unroll
        for (int k = 0; k < kR3; ++k) tail[m][k] = o[m][k].y * inv;
    }
  }
  vmax = warp_max(vmax);
  vsum = warp_sum(vsum);
  if ((threadIdx.x & 31) == 0) {
    s_max[threadIdx.x >> 5] = vmax;
    s_sum[threadIdx.x >> 5] = vsum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.f, s = 0.f;
#pragma unroll
    for (int w = 0; w < kCtaThreads / 32; ++w) {
      m = fmaxf(m, s_max[w]);
      s += s_sum[w];
    }
    // slot relative to the EVENT's range: events without IRs (k_tile) own partial slots too, so the chunk-wide CTA
    // index is not the slot index (found by tests/test_gpu_fuzz.py: every event after a no-IR event got a wrong gain)
    partials[ev.part0 + local] = make_float2(m, s);
  }
}

// k_small_rir: "batched small-RIR kernel" for static renders whose effective RIR fits ONE partition (taps <= P): short
// RIRs (SOFA / anechoic HRIRs) and, above all, the dry / direct-path sub-events of compute_dry_audio, whose RIR is the
// <= 65 ms window around the direct-path peak (synthesize.py:432-504; 1 560 taps at 24 kHz) — the general pipeline
// spends 12 partition transforms, a spectra round trip through the workspace and four launches on them.
// One FFT group per (event, capsule, run of kRun blocks): the RIR spectrum is computed once and STAYS IN REGISTERS; per
// block one forward transform of the source, the pointwise product, one inverse transform, overlap-add with the tail in
// registers, gain statistics. No spectrum ever leaves the SM. The tap window [mask_lo, mask_hi) (set by k_dry_window
// for dry events) is shifted to tap 0 and the output written at offset mask_lo, which is the same convolution.
// a_0 (normalize_irs) comes from the parent for dry events and is computed here otherwise (every group reduces all C
// energies in the same fixed order, so all CTAs of an event agree bit for bit).
__global__ void __launch_bounds__(kCtaThreads, 2)
k_small_rir(const EvDev* __restrict__ evs, int n_ev, const int* __restrict__ prefix, const float2* __restrict__ tw,
            const float2* __restrict__ zeta, EvStat* __restrict__ stats, float2* __restrict__ partials) {
  __shared__ FftSmem sm[kGroupsPerCta];
  __shared__ float s_red[kGroupsPerCta][kGroup / 32];
  __shared__ float s_max[kCtaThreads / 32], s_sum[kCtaThreads / 32];
  const int g = threadIdx.x / kGroup, t = threadIdx.x % kGroup, bar = 1 + g;
  const int e = find_segment(prefix, n_ev, blockIdx.x);
  const EvDev& ev = evs[e];
  const int local = blockIdx.x - __ldg(prefix + e);
  const int lo = ev.mask_lo, hi = min(ev.mask_hi, ev.Lh);
  const int wlen = max(hi - lo, 0);
  // convolution blocks (relative to tap lo): the host sized the grid for the largest window it could be
  const int nruns = ev.B_out;  // runs of kRun blocks (host: ceil(ceil((Lx + wmax - 1) / P) / kRun))
  const int run = local % nruns, cg = local / nruns;
  const int c = cg * kGroupsPerCta + g;
  float vmax = 0.f, vsum = 0.f;
  if (c < ev.C) {
    const float2 zt = __ldg(zeta + t);
    // ---- scale: a_0 * peak-normalisation scalar
    double a0 = 1.0;
    if (ev.gain_mode == kGainDry) {
      a0 = stats[ev.parent].a0;
    } else if (ev.normalize) {
      double mean_e = 0.0;
      for (int cc = 0; cc < ev.C; ++cc) {
        const float* __restrict__ hc = ev.irs + (long long)cc * ev.ir_stride_c;
        float en = 0.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const int n = t + kGroup * r;
          const float v = (n < ev.Lh) ? __ldg(hc + n) : 0.f;
          en = fmaf(v, v, en);
        }
        en = warp_sum(en);
        if ((t & 31) == 0) s_red[g][t >> 5] = en;
        group_sync(bar);
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < kGroup / 32; ++w) tot += s_red[g][w];
        group_sync(bar);
        mean_e += sqrt((double)tot) + 2.2250738585072014e-308;
      }
      mean_e /= ev.C;
      a0 = mean_e > 0.0 ? 1.0 / mean_e : 0.0;
      if (!(a0 < 3.0e38)) a0 = 0.0;
      if (run == 0 && c == 0 && t == 0) stats[ev.stat].a0 = a0;
    }
    const float sc = (float)a0 * (ev.xnorm ? __ldg(ev.xnorm) : 1.f);
    // ---- RIR spectrum (window shifted to tap 0), kept in registers in the inverse transform's input order
    float2 H[16];
    {
      const float* __restrict__ src = ev.irs + (long long)c * ev.ir_stride_c + lo;
      float a[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int n = t + kGroup * r;
        a[r] = (n < wlen) ? __ldg(src + n) : 0.f;
      }
      float2 o[kM3][kR3];
      fwd_block_to_regs(a, zt, sm[g], tw, t, bar, o);
#pragma unroll
      for (int m = 0; m < kM3; ++m)
#pragma unroll
        for (int k = 0; k < kR3; ++k) H[m + (256 / kGroup) * k] = o[m][k];  // element t + kGroup (m + 2k)
    }
    const float inv = 1.0f / kP;
    const int n_conv = ev.Lx + wlen - 1;  // samples of the (shifted) convolution that carry signal
    float* __restrict__ y = ev.y + (long long)c * ev.n_out;
    const float* __restrict__ x = ev.x;
    const int b0 = run * kRun, b1 = b0 + kRun;
    // the outputs before the window offset and behind the last convolution block are exact zeros
    if (run == 0)
      for (int n = t; n < min(lo, ev.n_out); n += kGroup) y[n] = 0.f;
    if (run == nruns - 1)
      for (int n = lo + b1 * kP + t; n < ev.n_out; n += kGroup) y[n] = 0.f;
    float tail[kM3][kR3];
#pragma unroll
    for (int m = 0; m < kM3; ++m)
#pragma unroll
      for (int k = 0; k < kR3; ++k) tail[m][k] = 0.f;
    for (int b = max(b0 - 1, 0); b < b1; ++b) {
      float2 o[kM3][kR3];
      if (b * kP < ev.xlimit) {
        float a[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const int n = b * kP + t + kGroup * r;
          a[r] = (n < ev.xlimit) ? __ldg(x + n) * sc : 0.f;
        }
        float2 xs[kM3][kR3];
        fwd_block_to_regs(a, zt, sm[g], tw, t, bar, xs);
        float2 v[16];
#pragma unroll
        for (int m = 0; m < kM3; ++m)
#pragma unroll
          for (int k = 0; k < kR3; ++k) v[m + (256 / kGroup) * k] = cmul(xs[m][k], H[m + (256 / kGroup) * k]);
        inv_block_from_regs(v, zt, sm[g], tw, t, bar, o);
      } else {
#pragma unroll
        for (int m = 0; m < kM3; ++m)
#pragma unroll
          for (int k = 0; k < kR3; ++k) o[m][k] = make_float2(0.f, 0.f);
      }
      if (b >= b0) {
#pragma unroll
        for (int k = 0; k < kR3; ++k)
#pragma unroll
          for (int m = 0; m < kM3; ++m) {
            const int ci = b * kP + t + kGroup * m + 256 * k;  // index into the shifted convolution
            const int n = lo + ci;
            float a = fmaf(o[m][k].x, inv, tail[m][k]);
            if (ci >= n_conv || n >= ev.n_valid) a = 0.f;
            if (n < ev.n_out) {
              y[n] = a;
              vmax = fmaxf(vmax, fabsf(a));
              vsum += fabsf(a);
            }
          }
      }
#pragma unroll
      for (int m = 0; m < kM3; ++m)
#pragma unroll
        for (int k = 0; k < kR3; ++k) tail[m][k] = o[m][k].y * inv;
    }
  }
  vmax = warp_max(vmax);
  vsum = warp_sum(vsum);
  if ((threadIdx.x & 31) == 0) {
    s_max[threadIdx.x >> 5] = vmax;
    s_sum[threadIdx.x >> 5] = vsum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.f, s = 0.f;
#pragma unroll
    for (int w = 0; w < kCtaThreads / 32; ++w) {
      m = fmaxf(m, s_max[w]);
      s += s_sum[w];
    }
    partials[ev.part0 + local] = make_float2(m, s);
  }
}

// k_tile: events without IRs — dry audio repeated on every capsule (synthesize.py:572-577). One CTA per
// (event, slice); partial reductions like k_ifft_ola.
__global__ void k_tile(const EvDev* __restrict__ evs, const int* __restrict__ list, int slices,
                       float2* __restrict__ partials) {
  __shared__ float s_max[32], s_sum[32];
  const EvDev& ev = evs[list[blockIdx.y]];
  float vmax = 0.f, vsum = 0.f;
  const long long total = (long long)ev.C * ev.n_out;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % ev.n_out);
    const float a = n < ev.Lx ? ev.x[n] * (ev.xnorm ? __ldg(ev.xnorm) : 1.f) : 0.f;
    ev.y[i] = a;
    vmax = fmaxf(vmax, fabsf(a));
    vsum += fabsf(a);
  }
  vmax = warp_max(vmax);
  vsum = warp_sum(vsum);
  if ((threadIdx.x & 31) == 0) {
    s_max[threadIdx.x >> 5] = vmax;
    s_sum[threadIdx.x >> 5] = vsum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.f, s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      m = fmaxf(m, s_max[w]);
      s += s_sum[w];
    }
    partials[ev.part0 + blockIdx.x] = make_float2(m, s);
  }
  (void)slices;
}

// k_event_gain: one warp per event. Follows apply_snr + db_to_multiplier literally, in double:
//   M = max(1e-15, max|y|); y1 = y * snr / M; m = mean|y1|; S = 10^((ref_db+snr)/20) / (m + tiny); out = S * y1.
__global__ void k_event_gain(const EvDev* __restrict__ evs, int n_ev, const float2* __restrict__ partials,
                             EvStat* __restrict__ stats, float* __restrict__ gains) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_ev) return;
  const EvDev& ev = evs[w];
  if (ev.gain_mode == kGainPass) return;
  float m = 0.f;
  double s = 0.0;
  for (int p = lane; p < ev.nparts; p += 32) {
    float2 v = partials[ev.part0 + p];
    m = fmaxf(m, v.x);
    s += (double)v.y;
  }
  m = warp_max(m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane != 0) return;
  EvStat& st = stats[ev.stat];
  const double count = (double)ev.C * (double)ev.n_out;
  const double mean_abs = s / count;
  const int bad = !(isfinite(s) && isfinite((double)m));
  double gain = 1.0, scale = 1.0;
  const double peak = fmax(1e-15, (double)m);
  if (ev.gain_mode == kGainEvent) {
    const double m1 = fabs(ev.snr) * mean_abs / peak;  // mean|y * snr / M|
    scale = pow(10.0, (ev.ref_db + ev.snr) / 20.0) / (m1 + 2.2250738585072014e-308);
    gain = (m > 0.f) ? scale * ev.snr / peak : 0.0;    // y == 0 stays 0 (the reference multiplies 0 first)
  } else if (ev.gain_mode == kGainDry) {
    gain = stats[ev.parent].event_scale;                // dry * event_scale (synthesize.py:493)
    scale = gain;
  }
  if (ev.gain_mode == kGainDry) {
    st.nonfinite |= bad;
  } else {
    st.peak = peak;
    st.mean_abs = mean_abs;
    st.gain = gain;
    st.event_scale = scale;
    st.nonfinite = bad;
  }
  gains[w] = (float)gain;
}

// k_apply_gain: y *= gain for every event of the chunk. grid = (slices, events)
__global__ void k_apply_gain(const EvDev* __restrict__ evs, const float* __restrict__ gains) {
  const EvDev& ev = evs[blockIdx.y];
  if (ev.gain_mode == kGainNone || ev.gain_mode == kGainPass) return;
  const float gain = gains[blockIdx.y];
  const long long total = (long long)ev.C * ev.n_out;
  float* __restrict__ y = ev.y;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ev.gain_from > 0) {  // the head of every channel is scaled by k_mix while it is mixed (one pass over y less)
    const int tail = ev.n_out - ev.gain_from;
    if (tail <= 0) return;
    for (long long q = i; q < (long long)ev.C * tail; q += stride) {
      const long long c = q / tail, n = ev.gain_from + q % tail;
      y[c * ev.n_out + n] *= gain;
    }
    return;
  }
  if ((reinterpret_cast<uintptr_t>(y) & 15u) == 0) {
    float4* y4 = reinterpret_cast<float4*>(y);
    const long long n4 = total >> 2;
    for (long long q = i; q < n4; q += stride) {
      float4 v = y4[q];
      v.x *= gain; v.y *= gain; v.z *= gain; v.w *= gain;
      y4[q] = v;
    }
    for (long long q = (n4 << 2) + i; q < total; q += stride) y[q] *= gain;
  } else {
    for (; i < total; i += stride) y[i] *= gain;
  }
}

// k_dry_window: peak = argmax(irs[ref, 0, :]) (signed, first occurrence; synthesize.py:482) and the tap window
// [peak - low, peak + high) that compute_dry_audio keeps (:484-487). One CTA per dry event; writes the mask into
// the event descriptor that k_ir_fft reads afterwards.
__global__ void k_dry_window(EvDev* __restrict__ evs, const int* __restrict__ list, EvStat* __restrict__ stats) {
  __shared__ float s_v[32];
  __shared__ int s_i[32];
  EvDev& ev = evs[list[blockIdx.x]];
  const float* h = ev.irs;  // already offset to (ref channel, IR 0)
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int n = threadIdx.x; n < ev.Lh; n += blockDim.x) {
    float v = h[n];
    if (v > best) { best = v; bi = n; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = best; s_i[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (s_v[w] > best || (s_v[w] == best && s_i[w] < bi)) { best = s_v[w]; bi = s_i[w]; }
    if (bi == 0x7fffffff) bi = 0;  // all-NaN row: numpy's argmax would return the first NaN; flagged via nonfinite
    // ir[peak+high:] = 0 if peak+high < Lh ; ir[:peak-low] = 0 if peak-low > 0
    ev.mask_hi = (bi + ev.dry_high < ev.Lh) ? bi + ev.dry_high : ev.Lh;
    ev.mask_lo = (bi - ev.dry_low > 0) ? bi - ev.dry_low : 0;
    stats[ev.parent].dry_peak = bi;
  }
}

// ------------------------------------------------------------------------------------------------------------
// ambience: sum|a| partials -> scale = 10^(ref_db/20) / (mean|a| + tiny)      (synthesize.py:350-352)
__global__ void k_amb_partial(const AmbDev* __restrict__ ambs, float* __restrict__ partials) {
  __shared__ float s_sum[32];
  const AmbDev& a = ambs[blockIdx.y];
  float s = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(a.data) & 15u) == 0) {
    const float4* d4 = reinterpret_cast<const float4*>(a.data);
    const long long n4 = a.n >> 2;
    for (long long q = i; q < n4; q += stride) {
      float4 v = __ldg(d4 + q);
      s += (fabsf(v.x) + fabsf(v.y)) + (fabsf(v.z) + fabsf(v.w));
    }
    for (long long q = (n4 << 2) + i; q < a.n; q += stride) s += fabsf(a.data[q]);
  } else {
    for (; i < a.n; i += stride) s += fabsf(a.data[i]);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_sum[w];
    partials[a.part0 + blockIdx.x] = tot;
  }
}

__global__ void k_amb_final(AmbDev* __restrict__ ambs, int n_amb, const float* __restrict__ partials) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_amb) return;
  AmbDev& a = ambs[w];
  double s = 0.0;
  for (int p = lane; p < a.nparts; p += 32) s += (double)partials[a.part0 + p];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    const double mean = s / (double)a.n;
    const double sc = pow(10.0, a.ref_db / 20.0) / (mean + 2.2250738585072014e-308);
    a.scale = mean > 0.0 ? (float)sc : 0.f;  // all-zero ambience contributes zeros in the reference too
  }
}

// k_mix: scene[c, t] = sum_a scale_a * amb_a[c, t] + sum_e y_e[c, t - start_e] for start_e <= t < end_e, added in
// the reference's order (ambience first, then events in dict order) with one float32 rounding per term like
// the reference's float32 scene buffer (synthesize.py:332,356,378). grid = (time tiles, scenes); every thread
// owns 4 samples strided by the CTA width so all accesses are coalesced whatever the event offsets are.
__global__ void __launch_bounds__(256)
k_mix(const SceneDev* __restrict__ scenes, const AmbDev* __restrict__ ambs, const MixEv* __restrict__ mevs) {
  const SceneDev& sc = scenes[blockIdx.y];
  const long long T = sc.T, tile0 = (long long)blockIdx.x * 1024;
  if (tile0 >= T) return;
  const int tlen = (int)min(1024LL, T - tile0), tid = threadIdx.x;
  // The kernel was issue-bound (73 % issue, 54 % DRAM) on 64-bit index arithmetic repeated per channel and sample:
  // events are now filtered once per group of 4 channels, and one unsigned compare per sample replaces the bounds
  // checks (rel = sample index inside the event; valid iff 0 <= rel < lim).
  for (int c0 = 0; c0 < sc.C; c0 += 4) {
    const int nc = min(4, sc.C - c0);
    float acc[4][4];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[cc][q] = 0.f;
    for (int a = 0; a < sc.n_amb; ++a) {
      const AmbDev& am = ambs[sc.amb0 + a];
      const float scale = am.scale;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        if (cc < nc) {
          const float* __restrict__ d = am.data + (long long)(c0 + cc) * T + tile0 + tid;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (tid + 256 * q < tlen) acc[cc][q] = fmaf(scale, __ldg(d + 256 * q), acc[cc][q]);
        }
      }
    }
    for (int k = 0; k < sc.nev; ++k) {
      const MixEv& me = mevs[sc.ev0 + k];
      if (me.end <= tile0 || me.start >= tile0 + tlen) continue;  // uniform per CTA
      const int base = (int)(tile0 - me.start);                   // in (-1024, end - start)
      const int lim = (int)min(min(me.end - me.start, (long long)me.n_out), (long long)base + tlen);
      if (lim <= 0) continue;
      if (me.gain) {
        // fused k_apply_gain: the stored event audio becomes gain * y (rounded to float32 once, as before), and that
        // rounded value is what the mix adds
        const float g = __ldg(me.gain);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          if (cc < nc) {
            float* __restrict__ y = me.y + (long long)(c0 + cc) * me.n_out + base + tid;
            float v[4];  // all loads first: the stores below must not serialise them
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = ((unsigned)(base + tid + 256 * q) < (unsigned)lim) ? y[256 * q] : 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if ((unsigned)(base + tid + 256 * q) < (unsigned)lim) {
                v[q] *= g;
                y[256 * q] = v[q];
                acc[cc][q] += v[q];
              }
          }
        }
        continue;
      }
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        if (cc < nc) {
          const float* __restrict__ y = me.y + (long long)(c0 + cc) * me.n_out + base + tid;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if ((unsigned)(base + tid + 256 * q) < (unsigned)lim) acc[cc][q] += __ldg(y + 256 * q);
        }
      }
    }
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      if (cc < nc) {
        float* __restrict__ out = sc.mix + (long long)(c0 + cc) * T + tile0 + tid;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (tid + 256 * q < tlen) out[256 * q] = acc[cc][q];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// f1: linear event augmentations on the dry audio (audiblelight/augmentation.py, applied by Event.load_audio,
// event.py:530-536). Every op reads `src` and writes `dst` (ping-pong), so Reverse needs no in-place swap.
enum { kAugGain = 0, kAugInvert = 1, kAugReverse = 2, kAugFade = 3, kAugBiquad = 4, kAugPreemph = 5, kAugDeemph = 6,
       kAugDelay = 7 };
constexpr int kIirChunk = 512;  // samples per thread of the chunked IIR scan

struct AugDev {       // one (event, op) application
  const float* src;
  float* dst;
  int L;              // samples
  int type;
  int fin_shape, fout_shape, fin, fout;
  int chunk0;         // first entry of the chunk-state scratch (IIR ops)
  int nchunks;
  double p[6];        // coefficients in double: rounding a pole near the unit circle to float32 alone costs ~1e-5 of output
};

__device__ __forceinline__ float fade_in_curve(int shape, float f) {
  // augmentation.py:1494-1508 (f = linspace(0, 1, fade_len)[n])
  const float pi = 3.14159265358979323846f;
  switch (shape) {
    case 1: f = exp2f(f - 1.f) * f; break;                        // exponential
    case 2: f = log10f(0.1f + f) + 1.f; break;                    // logarithmic
    case 3: f = sinf(f * pi / 2.f); break;                        // quarter sine
    case 4: f = sinf(f * pi - pi / 2.f) / 2.f + 0.5f; break;      // half sine
    default: break;                                               // linear
  }
  return fminf(fmaxf(f, 0.f), 1.f);
}
__device__ __forceinline__ float fade_out_curve(int shape, float f) {
  // augmentation.py:1516-1530
  const float pi = 3.14159265358979323846f;
  switch (shape) {
    case 0: f = 1.f - f; break;
    case 1: f = exp2f(-f) * (1.f - f); break;
    case 2: f = log10f(1.1f - f) + 1.f; break;
    case 3: f = sinf(f * pi / 2.f + pi / 2.f); break;
    case 4: f = sinf(f * pi + pi / 2.f) / 2.f + 0.5f; break;
    default: break;
  }
  return fminf(fmaxf(f, 0.f), 1.f);
}

// Gain / Invert / Reverse / Fade / Preemphasis (all without recursion). grid = (slices, ops)
__global__ void k_aug_pointwise(const AugDev* __restrict__ ops) {
  const AugDev& o = ops[blockIdx.y];
  const int L = o.L;
  if (o.type == kAugDelay) {
    // pedalboard.Delay: every sample pops the delay line (d[n] = w[n-D]), pushes w[n] = x[n] + feedback d[n] and
    // outputs (1 - mix) x[n] + mix d[n]. The D residue classes n = r (mod D) are independent first-order recurrences
    // d[n] = x[n-D] + feedback d[n-D]: one thread per residue, coalesced across r.
    const int D = (int)o.p[0];
    const float fb = (float)o.p[1], wet = (float)o.p[2], dry = 1.f - (float)o.p[2];
    if (D <= 0) {
      for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < L; n += gridDim.x * blockDim.x) o.dst[n] = o.src[n];
      return;
    }
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < min(D, L); r += gridDim.x * blockDim.x) {
      float xp = 0.f, dp = 0.f;  // x[n-D], d[n-D]
      for (int n = r; n < L; n += D) {
        const float x = o.src[n];
        const float d = (n >= D) ? fmaf(fb, dp, xp) : 0.f;
        o.dst[n] = fmaf(wet, d, dry * x);
        xp = x;
        dp = d;
      }
    }
    return;
  }
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < L; n += gridDim.x * blockDim.x) {
    float v;
    switch (o.type) {
      case kAugGain: v = (float)o.p[0] * o.src[n]; break;
      case kAugInvert: v = -o.src[n]; break;
      case kAugReverse: v = o.src[L - 1 - n]; break;
      case kAugFade: {
        float f = 1.f;
        if (o.fin > 0 && o.fin_shape != 5 && n < o.fin)
          f *= fade_in_curve(o.fin_shape, o.fin > 1 ? (float)n / (float)(o.fin - 1) : 0.f);
        if (o.fout > 0 && o.fout_shape != 5 && n >= L - o.fout) {
          const int i = n - (L - o.fout);
          f *= fade_out_curve(o.fout_shape, o.fout > 1 ? (float)i / (float)(o.fout - 1) : 0.f);
        }
        v = o.src[n] * f;
        break;
      }
      case kAugPreemph: {
        // librosa.effects.preemphasis: lfilter([1, -coef], [1], x, zi = 2 x[0] - x[1])  ->  y[0] = x[0] + zi
        const float c = (float)o.p[0];
        if (n == 0) v = o.src[0] + (L > 1 ? 2.f * o.src[0] - o.src[1] : o.src[0]);
        else v = o.src[n] - c * o.src[n - 1];
        break;
      }
      default: v = o.src[n]; break;
    }
    o.dst[n] = v;
  }
}

// Recursive filters (biquad, de-emphasis) as a chunked scan: with the transposed direct form II state s = (s1, s2)
//   y = b0 x + s1;  s1' = b1 x - a1 y + s2;  s2' = b2 x - a2 y
// a chunk maps s_in -> A^Lc s_in + s_zs (A = [[-a1, 1], [-a2, 0]]). Pass 1: zero-state run of every chunk (s_zs);
// combine: sequential over the chunks of one op in double; pass 2: re-run every chunk from its true s_in.
__device__ __forceinline__ void iir_coeffs(const AugDev& o, double& b0, double& b1, double& b2, double& a1, double& a2) {
  if (o.type == kAugDeemph) {  // y[n] = x[n] + coef y[n-1]
    b0 = 1.0; b1 = 0.0; b2 = 0.0; a1 = -o.p[0]; a2 = 0.0;
  } else {
    b0 = o.p[0]; b1 = o.p[1]; b2 = o.p[2]; a1 = o.p[3]; a2 = o.p[4];
  }
}
// One thread per 512-sample chunk runs the recurrence; the samples travel through shared memory so that every global
// access is a coalesced 128-byte row: a warp owns 32 chunks and, per 32-sample step, issues 32 independent row loads
// (one per chunk, all in flight together), transposes them through a padded 32 x 33 tile, filters its own row and
// writes the rows back the same way. History of this kernel on the benchmark step (both passes): 3.9 ms walking memory
// sample by sample at a 2 KB lane stride; 1.2 ms with a transposing tile but 4 loads in flight; 0.54 ms with
// per-thread 16-byte loads (each warp request still touched 32 lines: L1TEX wavefront bound); this version is bound
// by the recurrence itself.
// The state (s1, s2), the coefficients and the chunk hand-over run in float64: a float32 recurrence loses ~eps / (1 - |pole|)
// of the signal (2e-5 of full scale for a 32 Hz high-pass at 24 kHz), more than the 1e-5 the whole path is allowed. Samples
// stay float32 in memory (what Event.load_audio returns), each output is rounded once.
constexpr int kIirCta = 128, kIirSub = 32;

template <bool WRITE>
__global__ void __launch_bounds__(kIirCta)
k_iir_pass(const AugDev* __restrict__ ops, const int* __restrict__ chunk_prefix, int n_ops, int n_chunks,
           double2* __restrict__ zs) {
  __shared__ float tile[kIirCta / 32][32][33];
  __shared__ const float* s_src[kIirCta];
  __shared__ float* s_dst[kIirCta];
  __shared__ int s_n0[kIirCta], s_n1[kIirCta];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row0 = warp * 32;
  const int c = blockIdx.x * kIirCta + threadIdx.x;
  const bool live = c < n_chunks;
  double b0 = 0.0, b1 = 0.0, b2 = 0.0, a1 = 0.0, a2 = 0.0;
  int n0 = 0, n1 = 0;
  const float* src = nullptr;
  float* dst = nullptr;
  double s1 = 0.0, s2 = 0.0, corr = 0.0, cpow = 1.0, coef = 0.0;
  bool deemph = false;
  if (live) {
    const int oi = find_segment(chunk_prefix, n_ops, c);
    const AugDev& o = ops[oi];
    iir_coeffs(o, b0, b1, b2, a1, a2);
    n0 = (c - chunk_prefix[oi]) * kIirChunk;
    n1 = min(n0 + kIirChunk, o.L);
    src = o.src;
    dst = o.dst;
    if (WRITE) {
      s1 = zs[c].x;
      s2 = zs[c].y;
      // librosa.effects.deemphasis subtracts ((2 - coef) x0 - x1) / (3 - coef) * coef^n afterwards
      if (o.type == kAugDeemph && o.L > 1) {
        deemph = true;
        coef = o.p[0];
        corr = ((2.0 - coef) * (double)src[0] - (double)src[1]) / (3.0 - coef);
        cpow = pow(coef, (double)n0);
      }
    }
  }
  s_src[threadIdx.x] = src;
  s_dst[threadIdx.x] = dst;
  s_n0[threadIdx.x] = n0;
  s_n1[threadIdx.x] = n1;
  __syncwarp();
  float (*T)[33] = tile[warp];
  for (int sub = 0; sub < kIirChunk / kIirSub; ++sub) {
    const int off = sub * kIirSub;
    if (__ballot_sync(0xffffffffu, n0 + off < n1) == 0u) break;  // every chunk of this warp is finished
    float v[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      const int n = s_n0[row0 + r] + off + lane;
      v[r] = (n < s_n1[row0 + r]) ? __ldg(s_src[row0 + r] + n) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 32; ++r) T[r][lane] = v[r];
    __syncwarp();
    const int cnt = min(kIirSub, n1 - n0 - off);  // samples of this thread's chunk in the tile (may be <= 0)
#pragma unroll
    for (int j = 0; j < kIirSub; ++j) {
      if (j < cnt) {
        const double x = (double)T[lane][j];
        const double y = fma(b0, x, s1);
        s1 = fma(b1, x, fma(-a1, y, s2));
        s2 = fma(b2, x, -a2 * y);
        if (WRITE) {
          if (deemph) {
            T[lane][j] = (float)(y - corr * cpow);
            cpow *= coef;
          } else {
            T[lane][j] = (float)y;
          }
        }
      }
    }
    __syncwarp();
    if (WRITE) {
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const int n = s_n0[row0 + r] + off + lane;
        if (n < s_n1[row0 + r]) s_dst[row0 + r][n] = T[r][lane];
      }
      __syncwarp();
    }
  }
  if (!WRITE && live) zs[c] = make_double2(s1, s2);
}

// One WARP per op: lanes fetch 32 zero-state results at once (one coalesced request instead of 32 dependent
// round trips), the carry is then chained through the batch from registers with shuffles. blockDim = 128.
__global__ void k_iir_combine(const AugDev* __restrict__ ops, const int* __restrict__ chunk_prefix, int n_ops,
                              double2* __restrict__ zs) {
  const int oi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (oi >= n_ops) return;
  const AugDev& o = ops[oi];
  double b0, b1, b2, a1f, a2f;
  iir_coeffs(o, b0, b1, b2, a1f, a2f);
  // M = A^kIirChunk by repeated squaring (kIirChunk is a power of two)
  double m00 = -a1f, m01 = 1.0, m10 = -a2f, m11 = 0.0;
  for (int q = 1; q < kIirChunk; q <<= 1) {
    const double n00 = m00 * m00 + m01 * m10, n01 = m00 * m01 + m01 * m11;
    const double n10 = m10 * m00 + m11 * m10, n11 = m10 * m01 + m11 * m11;
    m00 = n00; m01 = n01; m10 = n10; m11 = n11;
  }
  double s1 = 0.0, s2 = 0.0;  // state entering chunk 0 (pedalboard runs with reset=True; de-emphasis starts from zeros)
  const int c0 = chunk_prefix[oi], c1 = chunk_prefix[oi + 1];
  for (int cb = c0; cb < c1; cb += 32) {
    const int c = cb + lane;
    const double2 z = c < c1 ? zs[c] : make_double2(0.0, 0.0);
    double in1 = 0.0, in2 = 0.0;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
      const double zx = __shfl_sync(0xffffffffu, z.x, j), zy = __shfl_sync(0xffffffffu, z.y, j);
      if (lane == j) {  // the true entry state of chunk cb + j replaces its zero-state result
        in1 = s1;
        in2 = s2;
      }
      const double t1 = m00 * s1 + m01 * s2 + zx, t2 = m10 * s1 + m11 * s2 + zy;
      s1 = t1;
      s2 = t2;
    }
    if (c < c1) zs[c] = make_double2(in1, in2);
  }
}

// peak normalisation of Event.load_audio: x / max(|x| + tiny)   (event.py:535-536). The reciprocal peak is a per-event
// SCALAR: it is folded into the source-spectrum scale of k_x_fft (and k_tile) instead of rewriting the audio; only
// events whose normalised audio is requested back (alr_event.audio_out) are scaled in place.
struct NormDev {
  float* x;
  int L;
  int part0;
  int scale_in_place;  // 1: audio_out requested
};
__global__ void k_peak_partial(const NormDev* __restrict__ nd, float* __restrict__ partials) {
  __shared__ float s_max[32];
  const NormDev& d = nd[blockIdx.y];
  const float* __restrict__ x = d.x;
  const int L = d.L;
  float m = 0.f;
  const int stride = gridDim.x * blockDim.x, i0 = blockIdx.x * blockDim.x + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const int n4 = L >> 2;
    for (int q = i0; q < n4; q += stride) {
      const float4 v = __ldg(x4 + q);
      m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (int n = (n4 << 2) + i0; n < L; n += stride) m = fmaxf(m, fabsf(x[n]));
  } else {
    for (int n = i0; n < L; n += stride) m = fmaxf(m, fabsf(x[n]));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, s_max[w]);
    partials[d.part0 + blockIdx.x] = m;
  }
}
// xnorm[i] = 1 / (max + tiny(float32)); events with scale_in_place get their audio scaled and xnorm = 1. grid = (slices, events)
__global__ void k_peak_final(const NormDev* __restrict__ nd, const float* __restrict__ partials, int slices,
                             float* __restrict__ xnorm) {
  const NormDev& d = nd[blockIdx.y];
  float m = 0.f;
  for (int i = 0; i < slices; ++i) m = fmaxf(m, partials[d.part0 + i]);
  const float inv = 1.0f / (m + 1.17549435e-38f);
  if (!d.scale_in_place) {
    if (blockIdx.x == 0 && threadIdx.x == 0) xnorm[blockIdx.y] = inv;
    return;
  }
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < d.L; n += gridDim.x * blockDim.x) d.x[n] *= inv;
  if (blockIdx.x == 0 && threadIdx.x == 0) xnorm[blockIdx.y] = 1.0f;
}

// (C, T) float mix -> (T, C) interleaved PCM_16 as libsndfile writes it for sf.write(path, mix.T, sr): psf_lrintf(x * 0x7FFF)
// stored in a short (round to nearest even, 16-bit truncation, no clipping). grid = (ceil(T / 256), scenes)
__global__ void k_pcm16(const SceneDev* __restrict__ scenes) {
  const SceneDev& s = scenes[blockIdx.y];
  if (!s.pcm) return;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= s.T) return;
  const int C = s.C;
  short* __restrict__ out = s.pcm + t * C;
  if (C == 4 && (reinterpret_cast<uintptr_t>(s.pcm) & 7u) == 0) {
    short4 v;
    v.x = (short)__float2int_rn(s.mix[t] * 32767.f);
    v.y = (short)__float2int_rn(s.mix[s.T + t] * 32767.f);
    v.z = (short)__float2int_rn(s.mix[2 * s.T + t] * 32767.f);
    v.w = (short)__float2int_rn(s.mix[3 * s.T + t] * 32767.f);
    *reinterpret_cast<short4*>(out) = v;
    return;
  }
  for (int c = 0; c < C; ++c) out[c] = (short)__float2int_rn(s.mix[(long long)c * s.T + t] * 32767.f);
}

// ------------------------------------------------------------------------------------------------------------
// f3: Gaussian ambience generated on the device (Ambience.load_ambience with noise="gaussian", ambience.py:155-163 +
// per-channel peak normalisation :210-214). The reference draws it from numpy's UNSEEDED global generator, so there is
// no sample-level parity to keep: what is kept is the distribution (i.i.d. N(0, 1) per sample, channels independent)
// and the normalisation. Counter-based Philox4x32-10 keyed by (seed, layer), Box-Muller; the stream of a layer depends
// only on its seed, channel and sample index, never on the launch geometry. Two passes that both REGENERATE the
// numbers: the first only reduces max|x| per channel (no memory traffic), the second writes x / (max + tiny).
struct GenDev {
  float* out;            // (C, T)
  long long T;
  int C;
  int part0;             // first per-channel peak slot
  unsigned long long seed;
};
__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1,
                                              unsigned (&r)[4]) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  r[0] = c0; r[1] = c1; r[2] = c2; r[3] = c3;
}
// four N(0, 1) samples for (seed, channel c, samples 4 q .. 4 q + 3)
__device__ __forceinline__ void gauss4(unsigned long long seed, int c, unsigned long long q, float (&g)[4]) {
  unsigned r[4];
  philox4x32_10((unsigned)q, (unsigned)(q >> 32), (unsigned)c, 0x414c5221u, (unsigned)seed, (unsigned)(seed >> 32), r);
  // Box-Muller on uniforms in (0, 1]
  const float u0 = ((float)(r[0] >> 8) + 1.0f) * (1.0f / 16777216.0f), u1 = (float)(r[1] >> 8) * (1.0f / 16777216.0f);
  const float u2 = ((float)(r[2] >> 8) + 1.0f) * (1.0f / 16777216.0f), u3 = (float)(r[3] >> 8) * (1.0f / 16777216.0f);
  const float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
  float sa, ca, sb, cb;
  sincospif(2.0f * u1, &sa, &ca);
  sincospif(2.0f * u3, &sb, &cb);
  g[0] = ra * ca; g[1] = ra * sa; g[2] = rb * cb; g[3] = rb * sb;
}
constexpr int kGenSlices = 32;
// grid = (kGenSlices, C, layers). WRITE = false: per-slice max|x| -> peaks[(part0 + c) * kGenSlices + slice]
template <bool WRITE>
__global__ void __launch_bounds__(256)
k_amb_gauss(const GenDev* __restrict__ gens, float* __restrict__ peaks) {
  __shared__ float s_max[8];
  const GenDev& gd = gens[blockIdx.z];
  const int c = blockIdx.y;
  if (c >= gd.C) return;
  const unsigned long long nq = (unsigned long long)((gd.T + 3) >> 2);
  float inv = 1.f;
  if (WRITE) {
    float m = 0.f;
    for (int i = 0; i < kGenSlices; ++i) m = fmaxf(m, peaks[(gd.part0 + c) * kGenSlices + i]);
    inv = 1.0f / (m + 1.17549435e-38f);  // channel / max(|channel| + tiny(float32 result))
  }
  float vmax = 0.f;
  float* __restrict__ out = gd.out + (long long)c * gd.T;
  for (unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq;
       q += (unsigned long long)gridDim.x * blockDim.x) {
    float g[4];
    gauss4(gd.seed, c, q, g);
    const long long n = (long long)(q << 2);
    if (WRITE) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (n + i < gd.T) out[n + i] = g[i] * inv;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (n + i < gd.T) vmax = fmaxf(vmax, fabsf(g[i]));
    }
  }
  if (!WRITE) {
    vmax = warp_max(vmax);
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = vmax;
    __syncthreads();
    if (threadIdx.x == 0) {
      float m = 0.f;
      for (int w = 0; w < 8; ++w) m = fmaxf(m, s_max[w]);
      peaks[(gd.part0 + c) * kGenSlices + blockIdx.x] = m;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// f4: STFT / visibility front-end of acoustic imaging (imaging.py:455-719). A band's "collapsed spectrum" is the sum of a
// few DFT bins of a windowed frame, i.e. ONE complex dot product of the frame with the modulated window
// g_b[n] = w[n] sum_k exp(-2 pi i k n / N): no FFT of the (non power-of-two) frame length is needed.
// grid = (frames, bands * channels), 64 threads; S[(f * n_bands + b) * C + c]
__global__ void __launch_bounds__(64)
k_vis_spectrum(const float* __restrict__ mix, long long T, int C, int N, int n_bands, const double2* __restrict__ g,
               double2* __restrict__ S) {
  __shared__ double s_re[2], s_im[2];
  const int f = blockIdx.x, b = blockIdx.y / C, c = blockIdx.y % C;
  const float* __restrict__ x = mix + (long long)c * T + (long long)f * N;
  const double2* __restrict__ gb = g + (long long)b * N;
  double re = 0.0, im = 0.0;
  for (int n = threadIdx.x; n < N; n += 64) {
    const double v = (double)__ldg(x + n);
    const double2 w = gb[n];
    re = fma(v, w.x, re);
    im = fma(v, w.y, im);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    re += __shfl_xor_sync(0xffffffffu, re, o);
    im += __shfl_xor_sync(0xffffffffu, im, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_re[threadIdx.x >> 5] = re;
    s_im[threadIdx.x >> 5] = im;
  }
  __syncthreads();
  if (threadIdx.x == 0) S[((long long)f * n_bands + b) * C + c] = make_double2(s_re[0] + s_re[1], s_im[0] + s_im[1]);
}
// grid = (blocks, bands), C*C threads (strided); V[((blk * n_bands + b) * C + i) * C + j] = sum_f conj(S_i) S_j
__global__ void k_vis_outer(const double2* __restrict__ S, int C, int n_bands, int per_block, double2* __restrict__ V) {
  const int blk = blockIdx.x, b = blockIdx.y;
  for (int ij = threadIdx.x; ij < C * C; ij += blockDim.x) {
    const int i = ij / C, j = ij % C;
    double re = 0.0, im = 0.0;
    for (int q = 0; q < per_block; ++q) {
      const long long f = (long long)blk * per_block + q;
      const double2 a = S[(f * n_bands + b) * C + i], bb = S[(f * n_bands + b) * C + j];
      re += a.x * bb.x + a.y * bb.y;   // conj(a) * bb
      im += a.x * bb.y - a.y * bb.x;
    }
    V[(((long long)blk * n_bands + b) * C + i) * C + j] = make_double2(re, im);
  }
}

// ---- unit-test kernels for the FFT core ---------------------------------------------------------------------
__global__ void __launch_bounds__(kCtaThreads)
k_debug_rfft(const float* __restrict__ in, long long n_blocks, long long in_stride, int n_valid,
             const float2* __restrict__ tw, const float2* __restrict__ zeta, float2* __restrict__ spec) {
  __shared__ FftSmem sm[kGroupsPerCta];
  const int g = threadIdx.x / kGroup, t = threadIdx.x % kGroup, bar = 1 + g;
  const long long blk = (long long)blockIdx.x * kGroupsPerCta + g;
  if (blk >= n_blocks) return;
  const float* src = in + blk * in_stride;
  float a[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int n = t + kGroup * r;
    a[r] = n < n_valid ? src[n] : 0.f;
  }
  fwd_block_to_global(a, __ldg(zeta + t), sm[g], tw, t, bar, spec + blk * kP);
}

__global__ void __launch_bounds__(kCtaThreads)
k_debug_irfft(const float2* __restrict__ spec, long long n_blocks, const float2* __restrict__ tw,
              const float2* __restrict__ zeta, float* __restrict__ out) {
  __shared__ FftSmem sm[kGroupsPerCta];
  const int g = threadIdx.x / kGroup, t = threadIdx.x % kGroup, bar = 1 + g;
  const long long blk = (long long)blockIdx.x * kGroupsPerCta + g;
  if (blk >= n_blocks) return;
  float2 o[kM3][kR3];
  inv_block_from_global(spec + blk * kP, __ldg(zeta + t), sm[g], tw, t, bar, o);
  const float inv = 1.0f / kP;
  float* dst = out + blk * 2 * kP;
#pragma unroll
  for (int m = 0; m < kM3; ++m)
#pragma unroll
    for (int k = 0; k < kR3; ++k) {
      const int n = t + kGroup * m + 256 * k;
      dst[n] = o[m][k].x * inv;
      dst[kP + n] = o[m][k].y * inv;
    }
}

}  // namespace alr
