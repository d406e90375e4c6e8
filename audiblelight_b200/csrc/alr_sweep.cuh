// alr_sweep.cuh — k_mov_sweep: warp-specialised persistent kernel for the moving events (sm_100a).
//
// Replaces k_ir_fft -> k_ir_scale -> k_cmac for moving events (the RIR STFTs, normalize_irs and the contraction of
// perform_time_variant_convolution, synthesize.py:184-252,298,404-428). The unfused kernels write the RIR partition
// spectra H (twice the size of the taps) to HBM and read them back 1.4x: 37 of the 46 GB they move per benchmark step.
// k_mov_fused (alr_fused.cuh) showed that an output-stationary consumer cannot keep H in L2: a run of output blocks
// reads a window of ~20 RIRs and the machine needs ~37 such windows in flight (profiles/r02_fused_ring.txt). This kernel
// turns the contraction around: the consumer is INPUT-driven. One CTA per SM, three kinds of warps:
//
//   warps 0-15  SWEEPER  owns 256 bins of ONE event at a time (thread = one bin x two capsules) and walks its RIRs in
//                        trajectory order. The outputs a RIR can touch form a window of W = K + xnb - 1 <= 15 blocks, kept
//                        as accumulators in shared memory (15 x 4 capsules x 256 bins x 8 B = 120 KB). Per RIR partition k
//                        the thread reads H_l[k, c][bin] once, does the xnb <= 4 multiply-accumulates against the source
//                        spectra X_l[j] it holds in registers (a 4-tap FIR along k in registers) and retires ONE finished
//                        partial sum per capsule into the window; blocks no later RIR can reach are stored to Y and the
//                        slot is recycled.
//   warp  28    COPY     one thread streams, per RIR, its source spectra and scale (one stage) and the H tiles of its
//                        partitions (4 capsules x 2 KB per stage) from global memory / the ring into a 6-stage shared-
//                        memory pipeline with TMA bulk copies (cp.async.bulk -> UBLKCP) completing on mbarriers; it runs
//                        ahead of the sweeper across RIR boundaries, so the sweeper never waits for a global load.
//   warps 16-27 PRODUCER three independent FFT groups claim P-tasks (event, RIR, capsule) from a global ticket queue and
//                        write the K partition spectra of h_{l,c} into the ring (same arithmetic as k_ir_fft), the tap
//                        energy, a_l by the last capsule to finish, then publish: ready[RIR] = C + 1.
//
// Every H value is consumed by exactly 8 sweepers (one per 256-bin slice) within microseconds of being produced, in the
// order the host planner fixed (RIR t of every sweeper slot, then RIR t + 1, ...), so a ring of a few dozen MB holds
// everything in flight and stays in the 126 MB L2; a producer reuses a ring region once consumed[] shows that all 128
// sweeper warps that read the previous occupant are done. All waits point backwards in the production order and the grid
// is one CTA per SM (all resident), so the pipeline cannot deadlock; a clock64 watchdog turns a violation into an error
// code. Accumulation order is fixed by the RIR order: results are bit-reproducible.
#pragma once
#include "alr_fused.cuh"

#ifndef ALR_SWEEP_LEADER_WAIT
#define ALR_SWEEP_LEADER_WAIT 0
#endif
#ifndef ALR_SWEEP_SLEEP_NS
#define ALR_SWEEP_SLEEP_NS 128
#endif

namespace alr {

// Two layouts. P = 2048 (128-thread FFT groups): 16 sweeper warps on capsule pairs, 3 FFT groups, 15-block window.
// P = 4096 (256-thread FFT groups): 8 sweeper warps with 4 capsules per thread, 2 FFT groups, 8-block window (a 1 s RIR
// at 24 kHz is 6 partitions there).
constexpr bool kSwWide = kGroup == 256;
constexpr int kSwW = kSwWide ? 8 : 15;  // accumulator window in output blocks (K + max xnb - 1 must fit)
constexpr int kSwMaxXnb = 4;         // source blocks per RIR held in registers
constexpr int kSwStages = 6;         // pipeline depth (stages of 4 rows x 256 bins x 8 B)
constexpr int kSwBins = 256;         // bins per sweeper CTA
constexpr int kSwCapPerThread = kSwWide ? 4 : 2;   // capsules per sweeper thread
constexpr bool kSwPipelined = kSwCapPerThread <= 2;  // software-pipelined loads need registers the 4-capsule layout lacks
constexpr int kSwSweepThreads = kSwBins * (kChanGroup / kSwCapPerThread);  // 512
constexpr int kSwSweepWarps = kSwSweepThreads / 32;
constexpr int kSwGroups = kSwWide ? 2 : 3;  // independent FFT groups of kGroup threads
constexpr int kSwKSplit = 3;         // a (RIR, capsule) is produced as kSwKSplit P-tasks (partitions k = part, part + 3, ...):
                                     // shorter tasks, and the RIRs in production at any time fit the ring
constexpr int kSwProdThreads = kSwGroups * kGroup;
constexpr int kSwCopyWarp = (kSwSweepThreads + kSwProdThreads) / 32;
constexpr int kSwThreads = kSwSweepThreads + kSwProdThreads + 32;  // 928 (29 warps, allocated as 32 x 64 registers)
constexpr int kSwBinCtas = kP / kSwBins;                            // sweeper CTAs per event (8)

constexpr size_t kSwAccBytes = (size_t)kSwW * kChanGroup * kSwBins * sizeof(float2);                  // 122880
constexpr size_t kSwStageBytes = (size_t)kChanGroup * kSwBins * sizeof(float2);                      // 8192
constexpr size_t kSwFftBytes = sizeof(FftSmem) * kSwGroups;                                            // 52224
constexpr size_t kSwOffStage = kSwAccBytes;
constexpr size_t kSwOffFft = kSwOffStage + kSwStages * kSwStageBytes;
constexpr size_t kSwOffBar = kSwOffFft + kSwFftBytes;
constexpr size_t kSwOffMisc = kSwOffBar + 2 * kSwStages * sizeof(unsigned long long);
constexpr size_t kSwSmem = kSwOffMisc + 512;  // misc: 64 B per FFT group, scale slots at +256, fail flag at +320
// laid out for 4 capsules per group and P = 2048 (128-thread FFT groups); other partition sizes build without it
constexpr bool kSweepOk = kChanGroup == 4 && (kGroup == 128 || kGroup == 256) && kSwSmem <= 232448 &&
                          kSwSweepThreads + kSwProdThreads + 32 <= 1024;

struct SweepArgs {
  const EvDev* evs;
  const IrDev* irs;
  const FusedTask* tasks;  // P-tasks in production order
  int n_tasks;
  const int2* pop;         // per fused RIR ordinal: ordinals [x, y) whose ring region this RIR overwrites
  const int2* need;        // per fused RIR ordinal: (sweeper warps that read it, ready[] value once published)
  const int* prod;         // production index -> fused RIR ordinal (pop[] ranges are production indices)
  const int* slot_off;     // n_slots + 1 prefix into slot_jobs
  const int* slot_jobs;    // events (chunk-local index) each sweeper slot renders, in order
  int n_slots;
  FusedCtl* ctl;
  int* ready;
  int* consumed;
  float* ecap;
  float* irscale;
  EvStat* stats;
  const float2* tw;
  const float2* zeta;
  const float2* xspec;
  float2* hring;
  float2* yspec;
  long long spin_limit;
};

// ---- mbarrier + TMA bulk-copy primitives ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// waits for the phase with the given parity; false when the watchdog fired. try_wait suspends the thread in hardware
// for up to the hinted time, so a waiting warp costs a handful of issue slots per microsecond; the abort flag and the
// clock are looked at every 64th round only.
__device__ __forceinline__ bool mbar_try_wait_hint(unsigned long long* bar, unsigned parity, unsigned ns) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(unsigned long long* bar, unsigned parity, FusedCtl* ctl, long long limit) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  unsigned ns = ALR_SWEEP_SLEEP_NS;
  for (unsigned it = 1;; ++it) {
    if (mbar_try_wait_hint(bar, parity, 1000u)) return true;
    __nanosleep(ns);  // a waiting warp must not spin: it shares its scheduler with the FFT warps (measured: 44 % of all
    if (ns < 4 * ALR_SWEEP_SLEEP_NS) ns <<= 1;  // executed instructions were this loop before the back-off)
    if ((it & 63u) == 0u) {
      if (ld_relaxed(&ctl->abort) != 0) return false;
      if (clock64() - t0 > limit) {
        atomicExch(&ctl->abort, 1);
        return false;
      }
    }
  }
}
// TMA 1-D bulk copy global -> shared, completion (bytes) signalled on an mbarrier. SASS: UBLKCP.
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// range of partitions of a RIR that reach valid output blocks
__device__ __forceinline__ int sweep_k_count(const EvDev& ev, const IrDev& ir) {
  return ir.xnb > 0 ? max(0, min(ev.K, ev.B_valid - ir.xb0)) : 0;
}

#ifdef ALR_SWEEP_DEBUG
#define SW_T0() const long long t0__ = clock64()
#define SW_ACC(var) var += clock64() - t0__
#else
#define SW_T0()
#define SW_ACC(var)
#endif

// ---- COPY warp (one thread) ----------------------------------------------------------------------------------------------
// Stage sequence of an active RIR: one X stage (its xnb <= 4 source spectra, 2 KB each, plus the scale 512 a_l in
// sc_slot[stage]) followed by one H stage per partition (4 capsules x 2 KB). Everything the sweeper needs arrives
// through shared memory: it never waits for a global load.
__device__ __forceinline__ void sweep_copy_role(const SweepArgs& A, unsigned char* smem, int slot, int br) {
  unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + kSwOffBar);
  unsigned long long* empty = full + kSwStages;
  float* sc_slot = reinterpret_cast<float*>(smem + kSwOffMisc + 256);
  unsigned stage = 0, par = 1;  // waiting on an empty barrier's "previous" phase succeeds at once on the first lap
  [[maybe_unused]] long long w_empty = 0, w_ready = 0, n_stages = 0;
  const long long t_start = clock64();
  auto acquire_stage = [&]() -> bool {  // the stage `stage` is free again
    SW_T0();
    const bool ok = mbar_wait(empty + stage, par, A.ctl, A.spin_limit);
    SW_ACC(w_empty);
    ++n_stages;
    return ok;
  };
  auto next_stage = [&]() {
    if (++stage == kSwStages) {
      stage = 0;
      par ^= 1u;
    }
  };
  for (int ji = A.slot_off[slot]; ji < A.slot_off[slot + 1]; ++ji) {
    const EvDev& ev = A.evs[A.slot_jobs[ji]];
    const int C = ev.C, nc = min(kChanGroup, C);
    const float2* __restrict__ xbase = A.xspec + ev.xslot0 * kP + br * kSwBins;
    for (int l = 0; l < ev.N; ++l) {
      const IrDev ir = A.irs[ev.ir0 + l];
      const int kn = sweep_k_count(ev, ir);
      if (kn == 0) continue;
      // X stage: does not depend on the producers except for the scale
      if (!acquire_stage()) return;
      {
        SW_T0();
        if (!spin_ge(A.ready + ev.fo0 + l, C * kSwKSplit + 1, A.ctl, A.spin_limit)) return;
        SW_ACC(w_ready);
      }
      sc_slot[stage] = __ldcg(A.irscale + ev.ir0 + l);
      asm volatile("fence.proxy.async;" ::: "memory");  // producers' generic-proxy stores -> this thread's async-proxy reads
      mbar_arrive_expect_tx(full + stage, (unsigned)(ir.xnb * kSwBins * sizeof(float2)));
      {
        float2* dst = reinterpret_cast<float2*>(smem + kSwOffStage + stage * kSwStageBytes);
        for (int j = 0; j < ir.xnb; ++j)
          tma_bulk_g2s(dst + j * kSwBins, xbase + (long long)(ir.xslot + j) * kP, (unsigned)(kSwBins * sizeof(float2)),
                       full + stage);
      }
      next_stage();
      const float2* src = A.hring + (long long)ir.hring * kP + br * kSwBins;
      for (int k = 0; k < kn; ++k) {
        if (!acquire_stage()) return;
        mbar_arrive_expect_tx(full + stage, (unsigned)(nc * kSwBins * sizeof(float2)));
        float2* dst = reinterpret_cast<float2*>(smem + kSwOffStage + stage * kSwStageBytes);
        for (int c = 0; c < nc; ++c)
          tma_bulk_g2s(dst + c * kSwBins, src + ((long long)k * C + c) * kP, (unsigned)(kSwBins * sizeof(float2)),
                       full + stage);
        next_stage();
      }
    }
  }
#ifdef ALR_SWEEP_DEBUG
  if (blockIdx.x % 37 == 0)
    printf("[sweep dbg] cta %d copy: total %lld clk, wait empty %lld, wait ready %lld, stages %lld\n", blockIdx.x,
           clock64() - t_start, w_empty, w_ready, n_stages);
#else
  (void)t_start; (void)w_empty; (void)w_ready; (void)n_stages;
#endif
}

// ---- SWEEPER warps ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sweep_consume_role(const SweepArgs& A, unsigned char* smem, int slot, int br) {
  float2* acc = reinterpret_cast<float2*>(smem);  // [w][c][bin]
  unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + kSwOffBar);
  unsigned long long* empty = full + kSwStages;
  const float* sc_slot = reinterpret_cast<const float*>(smem + kSwOffMisc + 256);
  const int tid = threadIdx.x, lane = tid & 31;
  const int lb = tid & (kSwBins - 1);                 // bin inside the CTA's slice
  const int c0 = (tid / kSwBins) * kSwCapPerThread;   // first of this thread's two capsules
  const int bin = br * kSwBins + lb;
  unsigned stage = 0, par = 0;  // stage / phase parity of the next stage in sequence
  [[maybe_unused]] long long w_full = 0, n_irs_done = 0;
  const long long t_start = clock64();
  auto next_stage = [&]() {
    if (++stage == kSwStages) {
      stage = 0;
      par ^= 1u;
    }
  };
  // Only warp 0 polls the mbarrier; the other 15 warps join through a named barrier (a warp blocked in bar.sync costs no
  // issue slots, a warp blocked in mbarrier.try_wait is replayed by the hardware: with all 16 warps polling, a third of
  // the kernel's executed instructions were try_wait replays, profiles/r02_sweep.txt).
  [[maybe_unused]] int* fail_flag = reinterpret_cast<int*>(smem + kSwOffMisc + 320);
  auto wait_stage = [&]() -> bool {
#if ALR_SWEEP_LEADER_WAIT
    if (tid < 32) {
      if (!mbar_wait(full + stage, par, A.ctl, A.spin_limit)) *fail_flag = 1;
    }
    named_sync(12, kSwSweepThreads);
    return *reinterpret_cast<volatile int*>(fail_flag) == 0;
#else
    return mbar_wait(full + stage, par, A.ctl, A.spin_limit);
#endif
  };
  auto release = [&]() {  // every lane of the warp has its values of the current stage in registers
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + stage);
    next_stage();
  };
  for (int ji = A.slot_off[slot]; ji < A.slot_off[slot + 1]; ++ji) {
    const EvDev& ev = A.evs[A.slot_jobs[ji]];
    const int C = ev.C, B = ev.B_valid, N = ev.N;
    const int nc = max(0, min(kSwCapPerThread, C - c0));  // capsules of this thread that exist
    // window slot of output block b: b % kSwW (tracked incrementally); this thread's columns only
    for (int i = 0; i < kSwW; ++i)
#pragma unroll
      for (int c = 0; c < kSwCapPerThread; ++c) acc[(i * kChanGroup + c0 + c) * kSwBins + lb] = make_float2(0.f, 0.f);
    float2* __restrict__ ybase = A.yspec + (ev.yslot0 + c0) * kP + bin;
    int emitted = 0, wslot = 0;  // blocks [0, emitted) are stored; wslot = emitted % kSwW
    auto emit_until = [&](int b_end) {
      for (; emitted < b_end; ++emitted) {
        float2* a = acc + (wslot * kChanGroup + c0) * kSwBins + lb;
#pragma unroll
        for (int c = 0; c < kSwCapPerThread; ++c)
          if (c < nc) {
            __stcs(ybase + ((long long)emitted * C + c) * kP, a[c * kSwBins]);
            a[c * kSwBins] = make_float2(0.f, 0.f);
          }
        wslot = (wslot + 1 == kSwW) ? 0 : wslot + 1;
      }
    };
    const IrDev* __restrict__ irp = A.irs + ev.ir0;
    IrDev ir_next = irp[0];
    for (int l = 0; l < N; ++l) {
      const IrDev ir = ir_next;
      if (l + 1 < N) ir_next = irp[l + 1];  // descriptor of the next RIR: its latency hides behind this RIR's work
      const int kn = sweep_k_count(ev, ir);
      if (kn > 0) {
        emit_until(min(ir.xb0, B));  // no RIR from here on reaches blocks below xb0 (xb0 is non-decreasing)
        // X stage: source spectra of this RIR times 512 a_l
        float2 X[kSwMaxXnb];
        {
          {
            SW_T0();
            if (!wait_stage()) return;
            SW_ACC(w_full);
          }
          const float2* st = reinterpret_cast<const float2*>(smem + kSwOffStage + stage * kSwStageBytes) + lb;
          const float sc = sc_slot[stage];
#pragma unroll
          for (int j = 0; j < kSwMaxXnb; ++j) {
            X[j] = (j < ir.xnb) ? st[j * kSwBins] : make_float2(0.f, 0.f);
            X[j].x *= sc;
            X[j].y *= sc;
          }
          release();
        }
        float2 z[kSwCapPerThread][4];
#pragma unroll
        for (int c = 0; c < kSwCapPerThread; ++c)
#pragma unroll
          for (int s2 = 0; s2 < 4; ++s2) z[c][s2] = make_float2(0.f, 0.f);
        int ws = wslot + (ir.xb0 - emitted);  // window slot of block xb0 + k (emitted <= xb0 < emitted + kSwW)
        if (ws >= kSwW) ws -= kSwW;
        // Software pipeline: the H values of step k + 1 and the window entries step k retires into are requested BEFORE
        // the FMAs of step k. The shared-memory pipe is shared with the FFT exchanges of the producer warps, so an LDS
        // can take hundreds of cycles here.
        float2 h[kSwCapPerThread], hn[kSwCapPerThread];
        auto fetch = [&](float2 (&dst)[kSwCapPerThread]) -> bool {
          {
            SW_T0();
            if (!wait_stage()) return false;
            SW_ACC(w_full);
          }
          const float2* st = reinterpret_cast<const float2*>(smem + kSwOffStage + stage * kSwStageBytes) + c0 * kSwBins + lb;
#pragma unroll
          for (int c = 0; c < kSwCapPerThread; ++c) dst[c] = (c < nc) ? st[c * kSwBins] : make_float2(0.f, 0.f);
          return true;
        };
        if (kSwPipelined && !fetch(h)) return;
        for (int k4 = 0; k4 < kn; k4 += 4) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (k4 + kk < kn) {
              if (!kSwPipelined && !fetch(h)) return;
              release();  // stage of step k (pipelined: its values were requested a full step ago)
              if (kSwPipelined && k4 + kk + 1 < kn && !fetch(hn)) return;
              float2* a = acc + (ws * kChanGroup + c0) * kSwBins + lb;
              float2 av[kSwCapPerThread];
              if (kSwPipelined) {
#pragma unroll
                for (int c = 0; c < kSwCapPerThread; ++c) av[c] = (c < nc) ? a[c * kSwBins] : make_float2(0.f, 0.f);
              }
              // 4-tap FIR along k: X[j] * H[k] belongs to output block xb0 + k + j, held in z[.][(k + j) & 3]
#pragma unroll
              for (int j = 0; j < kSwMaxXnb; ++j) {
                const float2 x = X[j];
#pragma unroll
                for (int c = 0; c < kSwCapPerThread; ++c) {
                  float2& zz = z[c][(kk + j) & 3];
                  zz.x = fmaf(x.x, h[c].x, zz.x);
                  zz.x = fmaf(-x.y, h[c].y, zz.x);
                  zz.y = fmaf(x.x, h[c].y, zz.y);
                  zz.y = fmaf(x.y, h[c].x, zz.y);
                }
              }
              // block xb0 + k is complete for this RIR: retire it into the window (it is < B by the choice of kn)
#pragma unroll
              for (int c = 0; c < kSwCapPerThread; ++c) {
                if (c < nc) {
                  const float2 old = kSwPipelined ? av[c] : a[c * kSwBins];
                  a[c * kSwBins] = make_float2(old.x + z[c][kk].x, old.y + z[c][kk].y);
                }
                z[c][kk] = make_float2(0.f, 0.f);
              }
              ws = (ws + 1 == kSwW) ? 0 : ws + 1;
              if (kSwPipelined) {
#pragma unroll
                for (int c = 0; c < kSwCapPerThread; ++c) h[c] = hn[c];
              }
            }
          }
        }
        // tails: blocks xb0 + kn - 1 + j, j = 1 .. xnb - 1, sit in z[.][(kn - 1 + j) & 3]
#pragma unroll
        for (int s2 = 0; s2 < 4; ++s2) {
          const int j = (s2 - (kn - 1)) & 3;  // j in 0..3 with (kn - 1 + j) & 3 == s2
          const int b = ir.xb0 + kn - 1 + j;
          if (j >= 1 && j < ir.xnb && b < B) {
            int w2 = ws + (j - 1);
            if (w2 >= kSwW) w2 -= kSwW;
            float2* a = acc + (w2 * kChanGroup + c0) * kSwBins + lb;
#pragma unroll
            for (int c = 0; c < kSwCapPerThread; ++c)
              if (c < nc) {
                float2 v = a[c * kSwBins];
                v.x += z[c][s2].x;
                v.y += z[c][s2].y;
                a[c * kSwBins] = v;
              }
          }
        }
      }
      // this warp is done with the RIR's ring region (also for RIRs it never read: the producers' count is uniform)
      __syncwarp();
      if (lane == 0) atomicAdd(A.consumed + ev.fo0 + l, 1);
      ++n_irs_done;
    }
    emit_until(B);
  }
#ifdef ALR_SWEEP_DEBUG
  if (blockIdx.x % 37 == 0 && (tid == 0 || tid == 256))
    printf("[sweep dbg] cta %d sweeper %d: total %lld clk, wait full %lld, RIRs %lld\n", blockIdx.x, tid / 256,
           clock64() - t_start, w_full, n_irs_done);
#else
  (void)t_start; (void)w_full; (void)n_irs_done;
#endif
}

// ---- PRODUCER groups ---------------------------------------------------------------------------------------------------------
// The P-task of alr_fused.cuh, run by ONE FFT group (128 threads, one transform in flight): the 64-register budget next
// to the sweeper rules out two transforms per thread, and independent groups need no barrier wider than the FFT's own.
__device__ __forceinline__ void sweep_produce_role(const SweepArgs& A, unsigned char* smem, int grp, int t) {
  int* misc = reinterpret_cast<int*>(smem + kSwOffMisc) + grp * 16;  // [0] ticket, [1] fail flag, [2..10) energy partials
  float* red = reinterpret_cast<float*>(misc + 2);
  const int bar = 1 + grp;  // the group's named barrier (also used inside the FFT)
  FftSmem* fs = reinterpret_cast<FftSmem*>(smem + kSwOffFft) + grp;
  const float2 zt = __ldg(A.zeta + t);
  [[maybe_unused]] long long w_pop = 0, n_tasks_done = 0;
  const long long t_start = clock64();
  for (;;) {
    if (t == 0) {
      misc[0] = atomicAdd(&A.ctl->ticket, 1);
      misc[1] = 0;
    }
    group_sync(bar);
    const int ti = misc[0];
    if (ti >= A.n_tasks) {
#ifdef ALR_SWEEP_DEBUG
      if (blockIdx.x % 37 == 0 && t == 0)
        printf("[sweep dbg] cta %d producer %d: total %lld clk, wait ring %lld, tasks %lld\n", blockIdx.x, grp,
               clock64() - t_start, w_pop, n_tasks_done);
#else
      (void)t_start; (void)w_pop; (void)n_tasks_done;
#endif
      return;
    }
    ++n_tasks_done;
    const FusedTask tk = A.tasks[ti];
    const EvDev& ev = A.evs[tk.ev];
    const int l = tk.idx, c = tk.sub / kSwKSplit, part = tk.sub % kSwKSplit;
    const int g = ev.ir0 + l, fo = ev.fo0 + l;
    {
      SW_T0();
      const int2 pr = A.pop[fo];
      bool ok = true;
      for (int i = pr.x + t; i < pr.y; i += kGroup) {
        const int o = __ldg(A.prod + i);
        const int2 nd = __ldg(A.need + o);
        ok = ok && spin_ge(A.ready + o, nd.y, A.ctl, A.spin_limit) && spin_ge(A.consumed + o, nd.x, A.ctl, A.spin_limit);
      }
      if (!ok) misc[1] = 1;
      group_sync(bar);
      SW_ACC(w_pop);
      if (misc[1]) return;
    }
    const int K = ev.K, Cn = ev.C, Lh = ev.Lh;
    const float* __restrict__ src = ev.irs + (long long)c * ev.ir_stride_c + (long long)l * ev.ir_stride_n;
    const long long hslot = A.irs[g].hring;
    float en = 0.f;
    for (int k = part; k < K; k += kSwKSplit) {
      float a[1][16];
      const int t0 = k * kP, hi = min(Lh, t0 + kP);
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int n = t0 + t + kGroup * r;
        a[0][r] = (n < hi) ? __ldcs(src + n) : 0.f;
        en = fmaf(a[0][r], a[0][r], en);
      }
      float2* dst[1] = {A.hring + (hslot + (long long)k * Cn + c) * kP};
      fwd_blocks_to_global<1>(a, zt, fs, A.tw, t, bar, dst);
    }
    en = warp_sum(en);
    if ((t & 31) == 0) red[t >> 5] = en;
    group_sync(bar);  // every spectrum store of the group has been issued; red[] complete
    if (t == 0) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < kGroup / 32; ++w) tot += red[w];
      __stcg(A.ecap + ev.ecap0 + ((long long)l * Cn + c) * kSwKSplit + part, tot);
      __threadfence();
      const int old = atomicAdd(A.ready + fo, 1);
      if (old == Cn * kSwKSplit - 1) {  // last piece of this RIR: a_l = 1 / mean_c(||h_{l,c}|| + tiny)   (synthesize.py:425-428)
        __threadfence();
        double a = 1.0;
        if (ev.normalize) {
          double mean_e = 0.0;
          for (int cc = 0; cc < Cn; ++cc) {
            const float* e3 = A.ecap + ev.ecap0 + ((long long)l * Cn + cc) * kSwKSplit;
            float ec = 0.f;
#pragma unroll
            for (int q = 0; q < kSwKSplit; ++q) ec += __ldcg(e3 + q);  // fixed order: deterministic
            mean_e += sqrt((double)ec) + 2.2250738585072014e-308;
          }
          mean_e /= Cn;
          a = mean_e > 0.0 ? 1.0 / mean_e : 0.0;
          if (!(a < 3.0e38)) a = 0.0;
        }
        if (l == 0) A.stats[ev.stat].a0 = a;
        __stcg(A.irscale + g, (float)(512.0 * a));
        __threadfence();
        atomicAdd(A.ready + fo, 1);
      }
    }
    // the next ticket's group_sync orders red[] / misc[] reuse
  }
}

// P = 2048: 29 warps are allocated as 32 -> 64 registers per thread; P = 4096: 25 warps as 28 -> 72
__global__ void __launch_bounds__(kSweepOk ? kSwThreads : 32, 1)
k_mov_sweep(const SweepArgs A) {
  extern __shared__ __align__(16) unsigned char sweep_smem[];
  unsigned long long* full = reinterpret_cast<unsigned long long*>(sweep_smem + kSwOffBar);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSwStages; ++s) {
      mbar_init(full + s, 1);                          // the copy thread's arrive.expect_tx
      mbar_init(full + kSwStages + s, kSwSweepWarps);  // one arrival per sweeper warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    *reinterpret_cast<int*>(sweep_smem + kSwOffMisc + 320) = 0;
  }
  __syncthreads();
  const int tid = threadIdx.x;
  const int slot = blockIdx.x / kSwBinCtas, br = blockIdx.x % kSwBinCtas;
  if (tid < kSwSweepThreads) {
    if (slot < A.n_slots) sweep_consume_role(A, sweep_smem, slot, br);
  } else if (tid < kSwSweepThreads + kSwProdThreads) {
    sweep_produce_role(A, sweep_smem, (tid - kSwSweepThreads) / kGroup, (tid - kSwSweepThreads) % kGroup);
  } else if (tid == kSwCopyWarp * 32) {
    if (slot < A.n_slots) sweep_copy_role(A, sweep_smem, slot, br);
  }
}

}  // namespace alr
