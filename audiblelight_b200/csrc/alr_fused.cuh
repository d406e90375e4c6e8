// alr_fused.cuh — k_mov_fused: ONE persistent launch per chunk for the moving events (sm_100a).
//
// Replaces k_ir_fft -> k_ir_scale -> k_cmac for moving events (perform_time_variant_convolution and the RIR STFTs,
// synthesize.py:184-252,298; normalize_irs :404-428). In the unfused pipeline the RIR partition spectra H (twice the
// size of the taps) are written to HBM by one kernel and read back by the next: 37 of the 46 GB those two kernels
// move per benchmark step. Here producer and consumer tasks of the SAME launch hand the spectra over through a RING
// of a few dozen MB that stays resident in the 126 MB L2:
//
//   P-task (event, RIR l, capsule c)   FFT of the K partitions of h_{l,c} into the ring (two transforms in flight per
//                                      FFT group, fft_core_n<.., 2>), tap energy; the last capsule to finish derives
//                                      a_l (normalize_irs) and publishes the RIR:           ready[l] = C + 1
//   C-task (event, run of 8 output     waits for ready[lmin..lmax], streams H out of the ring (cp.async.cg, 16-byte,
//           blocks, 4 capsules, 256    warp-private stages) and accumulates Y[b,c] += a_l X_l[j] H_l[k,c] in
//           bins)                      registers exactly like k_cmac; then             consumed[l] += 1
//
// Tasks are claimed from one queue (atomic ticket) in an order the host planner fixes: the P-tasks of the RIRs that
// run i + lookahead needs come before the C-tasks of run i, so a consumer rarely waits, and a producer may overwrite
// a ring region only when consumed[] says every reader of the previous occupant has finished. Every wait targets
// tasks EARLIER in the queue and the grid is sized to be fully resident, so the dependency graph cannot deadlock; a
// clock64 watchdog turns any violation into an error code instead of a hang.
#pragma once
#include "alr_kernels.cuh"

namespace alr {

constexpr int kFusedNT = 2;  // transforms in flight per FFT group of a P-task
enum { kTaskP = 0, kTaskC = 1 };

struct FusedTask {
  int type;  // kTaskP / kTaskC
  int ev;    // event (chunk-local index)
  int idx;   // P: RIR l            C: run
  int sub;   // P: capsule c        C: cg * kBinCtas + br
};

struct FusedCtl {
  int ticket;
  int abort;  // set by the watchdog
  int pad[2];
};

struct FusedArgs {
  const EvDev* evs;
  const IrDev* irs;
  const int2* lrange;
  const FusedTask* tasks;
  int n_tasks;
  const int2* pop;      // per fused RIR ordinal: ordinals [x, y) whose ring region this RIR overwrites
  const int2* need;     // per fused RIR ordinal: (number of C-tasks that read it, ready[] value once published)
  FusedCtl* ctl;
  int* ready;           // per fused RIR ordinal
  int* consumed;        // per fused RIR ordinal
  float* ecap;          // tap energy per (RIR, capsule)
  float* irscale;       // per RIR of the chunk (512 a_l)
  EvStat* stats;
  const float2* tw;
  const float2* zeta;
  const float2* xspec;
  float2* hring;
  float2* yspec;
  long long spin_limit;  // watchdog, in clock64 ticks
};

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// spins until *p >= target; false when the watchdog fired (here or in another CTA)
__device__ __forceinline__ bool spin_ge(const int* p, int target, FusedCtl* ctl, long long limit) {
  if (ld_acquire(p) >= target) return true;
  const long long t0 = clock64();
  unsigned ns = 64;
  for (;;) {
    __nanosleep(ns);
    if (ns < 1024) ns <<= 1;
    if (ld_acquire(p) >= target) return true;
    if (ld_relaxed(&ctl->abort) != 0) return false;
    if (clock64() - t0 > limit) {
      atomicExch(&ctl->abort, 1);
      return false;
    }
  }
}

__device__ __forceinline__ void cp_async16_cg(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

constexpr size_t kFusedSmemP = sizeof(FftSmem) * kGroupsPerCta * kFusedNT + sizeof(float) * (kCtaThreads / 32);
struct FusedHead {  // one RIR as seen by one C-task
  int k_lo, k_hi;
  int d0, xnb;
  long long xoff;   // float2 offset of X_l[0] from the event's X base
  long long hoff;   // float2 offset of H_l[k_lo][c0] from the RING base
  float sc;         // 512 a_l
  int pad;
};
constexpr size_t kFusedRingBytes = sizeof(float2) * kStages * kChanGroup * kCtaThreads;
constexpr size_t kFusedSmemC = kFusedRingBytes + sizeof(FusedHead) * kMaxHeads;
constexpr size_t kFusedSmem = kFusedSmemP > kFusedSmemC ? kFusedSmemP : kFusedSmemC;

// ---- P-task -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool fused_p_task(const FusedArgs& A, const FusedTask& tk, unsigned char* smem) {
  const EvDev& ev = A.evs[tk.ev];
  const int l = tk.idx, c = tk.sub, tid = threadIdx.x;
  const int g = ev.ir0 + l, fo = ev.fo0 + l;
  // ring region release: every reader of the previous occupants has finished
  {
    const int2 pr = A.pop[fo];
    bool ok = true;
    for (int i = pr.x + tid; i < pr.y; i += kCtaThreads) {
      const int2 nd = __ldg(A.need + i);
      // the previous occupant has been written completely (matters when nobody reads it) and read by everyone
      ok = ok && spin_ge(A.ready + i, nd.y, A.ctl, A.spin_limit) && spin_ge(A.consumed + i, nd.x, A.ctl, A.spin_limit);
    }
    if (__syncthreads_or(!ok)) return false;
  }
  const int grp = tid / kGroup, t = tid % kGroup, bar = 1 + grp;
  FftSmem* fs = reinterpret_cast<FftSmem*>(smem) + grp * kFusedNT;
  float* red = reinterpret_cast<float*>(smem + sizeof(FftSmem) * kGroupsPerCta * kFusedNT);
  const int K = ev.K, Cn = ev.C, Lh = ev.Lh;
  const float* __restrict__ src = ev.irs + (long long)c * ev.ir_stride_c + (long long)l * ev.ir_stride_n;
  const long long hslot = A.irs[g].hring;
  const float2 zt = __ldg(A.zeta + t);
  float en = 0.f;
  for (int k0 = grp * kFusedNT; k0 < K; k0 += kGroupsPerCta * kFusedNT) {
    float a[kFusedNT][16];
    float2* dst[kFusedNT];
#pragma unroll
    for (int q = 0; q < kFusedNT; ++q) {
      const int k = k0 + q;
      const bool valid = k < K;
      const int t0 = k * kP, hi = min(Lh, t0 + kP);
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int n = t0 + t + kGroup * r;
        a[q][r] = (valid && n < hi) ? __ldcs(src + n) : 0.f;  // taps are read once: streaming (evict-first) loads
        en = fmaf(a[q][r], a[q][r], en);
      }
      dst[q] = valid ? A.hring + (hslot + (long long)k * Cn + c) * kP : nullptr;
    }
    fwd_blocks_to_global<kFusedNT>(a, zt, fs, A.tw, t, bar, dst);
  }
  en = warp_sum(en);
  if ((tid & 31) == 0) red[tid >> 5] = en;
  __syncthreads();  // every spectrum store of the CTA has been issued; red[] complete
  if (tid == 0) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kCtaThreads / 32; ++w) tot += red[w];
    __stcg(A.ecap + ev.ecap0 + (long long)l * Cn + c, tot);
    __threadfence();
    const int old = atomicAdd(A.ready + fo, 1);
    if (old == Cn - 1) {  // last capsule of this RIR: a_l = 1 / mean_c(||h_{l,c}|| + tiny)   (synthesize.py:425-428)
      __threadfence();
      double a = 1.0;
      if (ev.normalize) {
        double mean_e = 0.0;
        for (int cc = 0; cc < Cn; ++cc)
          mean_e += sqrt((double)__ldcg(A.ecap + ev.ecap0 + (long long)l * Cn + cc)) + 2.2250738585072014e-308;
        mean_e /= Cn;
        a = mean_e > 0.0 ? 1.0 / mean_e : 0.0;
        if (!(a < 3.0e38)) a = 0.0;
      }
      if (l == 0) A.stats[ev.stat].a0 = a;
      __stcg(A.irscale + g, (float)(512.0 * a));  // x 512: un-normalised irfft of istft_overlap_synthesis (:267)
      __threadfence();
      atomicAdd(A.ready + fo, 1);
    }
  }
  return true;
}

// ---- C-task -----------------------------------------------------------------------------------------------------
// Same contraction, tile and register blocking as k_cmac (alr_kernels.cuh). Differences: the H items come out of the
// L2-resident ring with 16-byte cp.async.cg copies (L1 is bypassed: the ring is rewritten during the launch), a WARP
// owns its 4 x 32-bin slice of every stage (lane i copies 2 of the warp's 64 16-byte pieces, __syncwarp hands them
// over), and the per-RIR scale a_l is applied to the source value instead of being folded into X by k_x_fft.
__device__ __forceinline__ bool fused_c_task(const FusedArgs& A, const FusedTask& tk, unsigned char* smem) {
  float2 (*ring)[kChanGroup][kCtaThreads] = reinterpret_cast<float2 (*)[kChanGroup][kCtaThreads]>(smem);
  FusedHead* heads = reinterpret_cast<FusedHead*>(smem + kFusedRingBytes);
  const EvDev& ev = A.evs[tk.ev];
  const int run = tk.idx;
  const int br = tk.sub % kBinCtas, cg = tk.sub / kBinCtas;
  const int c0 = cg * kChanGroup;
  const int nc = min(kChanGroup, ev.C - c0);
  const int b0 = run * kGm;
  const int nb = min(kGm, ev.B_valid - b0);
  const int tid = threadIdx.x, lane = tid & 31, wbase = tid & ~31;
  const int bin = br * kCtaThreads + tid;
  const int K = ev.K, C = ev.C;
  const long long kstride = (long long)C * kP;
  const int lmin = A.lrange[ev.blk0 + b0].x, lmax = A.lrange[ev.blk0 + b0 + nb - 1].y;
  const IrDev* __restrict__ irp = A.irs + ev.ir0;
  const float2* __restrict__ xbase = A.xspec + ev.xslot0 * kP + bin;
  // copy geometry of this lane: pieces p = lane, lane + 32 of the warp's 64 (capsule = p / 16, bin pair = p % 16)
  const int pc0 = lane >> 4, pb = (lane & 15) * 2;
  const float2* __restrict__ hcopy = A.hring + (long long)c0 * kP + br * kCtaThreads + wbase + pb;

  float2 acc[kGm][kChanGroup];
#pragma unroll
  for (int s = 0; s < kGm; ++s)
#pragma unroll
    for (int c = 0; c < kChanGroup; ++c) acc[s][c] = make_float2(0.f, 0.f);

  const int n_win = (lmax - lmin + kMaxHeads) / kMaxHeads;
  for (int wi = 0; wi < n_win; ++wi) {
    const int l0 = lmin + kMaxHeads * wi;
    const int n_heads = min(kMaxHeads, lmax - l0 + 1);
    __syncthreads();  // previous window fully consumed
    bool ok = true;
    if (tid < n_heads) {
      const int l = l0 + tid;
      ok = spin_ge(A.ready + ev.fo0 + l, C + 1, A.ctl, A.spin_limit);
      const IrDev ir = irp[l];
      FusedHead h;
      h.d0 = b0 - ir.xb0;
      h.xnb = ir.xnb;
      h.k_lo = max(0, h.d0 - ir.xnb + 1);
      h.k_hi = ir.xnb > 0 ? min(K - 1, h.d0 + nb - 1) : -1;
      h.xoff = (long long)ir.xslot * kP;
      h.hoff = ((long long)ir.hring + (long long)h.k_lo * C) * kP;
      h.sc = ok ? __ldcg(A.irscale + ev.ir0 + l) : 0.f;
      h.pad = 0;
      heads[tid] = h;
    }
    if (__syncthreads_or(!ok)) return false;

    // ---- producer state: next (RIR, partition) item whose H values get copied into the ring
    int ph = -1, pk_left = 0;
    const float2* php = hcopy;
    bool prod_ok = true;
    auto prod_advance = [&]() {
      if (--pk_left > 0) {
        php += kstride;
        return;
      }
      while (++ph < n_heads) {
        const FusedHead h = heads[ph];
        if (h.k_lo <= h.k_hi) {
          pk_left = h.k_hi - h.k_lo + 1;
          php = hcopy + h.hoff;
          const int j_lo = max(0, h.d0 - h.k_hi), j_hi = min(h.xnb - 1, h.d0 + nb - 1 - h.k_lo);
          const float2* xp = xbase + h.xoff;
          for (int j = j_lo; j <= j_hi; ++j) prefetch_l1(xp + (long long)j * kP);
          return;
        }
      }
      prod_ok = false;
    };
    auto produce = [&](int stage) {
      if (prod_ok) {
        float2* dst = &ring[stage][0][wbase + pb];
#pragma unroll
        for (int q = 0; q < kChanGroup / 2; ++q) {
          const int c = pc0 + 2 * q;
          if (c < nc) cp_async16_cg(dst + c * kCtaThreads, php + (long long)c * kP);
        }
        prod_advance();
      }
      cp_async_commit();
    };
    // ---- consumer state
    int ch = -1, ck_left = 0, cjb = 0, cxnb = 0;
    float csc = 0.f;
    const float2* cxq = xbase;
    bool cons_ok = true;
    auto cons_advance = [&]() {
      if (--ck_left > 0) {
        cjb -= 1;
        cxq -= kP;
        return;
      }
      while (++ch < n_heads) {
        const FusedHead h = heads[ch];
        if (h.k_lo <= h.k_hi) {
          ck_left = h.k_hi - h.k_lo + 1;
          cjb = h.d0 - h.k_lo;
          cxnb = h.xnb;
          csc = h.sc;
          cxq = xbase + h.xoff + (long long)cjb * kP;
          return;
        }
      }
      cons_ok = false;
    };
    prod_advance();
    cons_advance();
#pragma unroll
    for (int i = 0; i < kStages - 1; ++i) produce(i);
    int stage = 0;
    while (cons_ok) {
      cp_async_wait<kStages - 2>();  // this lane's pieces of the consumer's item have landed
      __syncwarp();                  // ... and every other lane's; all lanes are done with the previous stage
      const float2* src = &ring[stage][0][tid];
      float2 h[kChanGroup];
#pragma unroll
      for (int c = 0; c < kChanGroup; ++c) h[c] = (c < nc) ? src[c * kCtaThreads] : make_float2(0.f, 0.f);
      produce(stage == 0 ? kStages - 1 : stage - 1);  // refill the slot consumed in the previous iteration
      const int s_lo = max(0, -cjb);
      const unsigned s_cnt = (unsigned)max(0, min(nb, cxnb - cjb) - s_lo);
#pragma unroll
      for (int s = 0; s < kGm; s += 2) {
        const bool v0 = (unsigned)(s - s_lo) < s_cnt, v1 = (unsigned)(s + 1 - s_lo) < s_cnt;
        if (v0 || v1) {
          float2 x0 = make_float2(0.f, 0.f), x1 = make_float2(0.f, 0.f);
          if (v0) x0 = __ldg(cxq + s * kP);
          if (v1) x1 = __ldg(cxq + (s + 1) * kP);
          x0.x *= csc; x0.y *= csc; x1.x *= csc; x1.y *= csc;
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
      cons_advance();
      stage = (stage + 1 == kStages) ? 0 : stage + 1;
    }
    cp_async_wait<0>();
    __syncwarp();
  }
#pragma unroll
  for (int s = 0; s < kGm; ++s)
    if (s < nb)
#pragma unroll
      for (int c = 0; c < kChanGroup; ++c)
        if (c < nc) __stcg(A.yspec + (ev.yslot0 + (long long)(b0 + s) * C + c0 + c) * kP + bin, acc[s][c]);
  // release the RIRs of this run's range (all ring reads of the CTA are complete)
  __syncthreads();
  if (tid <= lmax - lmin) __threadfence();
  for (int l = lmin + tid; l <= lmax; l += kCtaThreads) atomicAdd(A.consumed + ev.fo0 + l, 1);
  return true;
}

__global__ void __launch_bounds__(kCtaThreads, 2)
k_mov_fused(const FusedArgs A) {
  extern __shared__ __align__(16) unsigned char fused_smem[];
  __shared__ int s_ticket;
  for (;;) {
    if (threadIdx.x == 0) s_ticket = atomicAdd(&A.ctl->ticket, 1);
    __syncthreads();
    const int ti = s_ticket;
    __syncthreads();
    if (ti >= A.n_tasks) return;
    const FusedTask tk = A.tasks[ti];
    const bool ok = tk.type == kTaskP ? fused_p_task(A, tk, fused_smem) : fused_c_task(A, tk, fused_smem);
    if (!ok) return;
  }
}

}  // namespace alr
