// alr_fft.cuh — shared-memory Stockham FFT core (sm_100a), used by every spectral kernel of the renderer.
//
// A real block of 2P samples (P signal + P zero padding) is transformed with the negacyclic fold+twist map
// (see below) and ONE P-point complex FFT done by a group of P/16 threads (16 elements each): Stockham autosort
// passes of radix 16, 16 and P/256, butterflies in registers, two exchanges through shared memory.  This replaces
// scipy's pocketfft calls of the reference (scipy.fft.rfft/irfft, synthesize.py:138,267; scipy.signal.fftconvolve,
// :103,490).  P is a compile-time choice (-DALR_P=1024|2048|4096). Benchmark step with the round-2 kernels: 24.5 / 21.8 /
// 21.1 ms — larger partitions trade spectral multiply-accumulates (and spectra re-reads) for FFT work; round 1 measured
// 26.5 / 24.5 / 24.9 and shipped 2048, the multiply-accumulate kernels have since gained more than the FFTs
// (profiles/r01_partition_sweep.txt, profiles/r02_partition_sweep.txt). Default: 4096.
//
// Shared-memory layout: float2 elements, one pad element per 16 (index i -> i + i/16).  With 64-bit accesses the
// hardware serves a warp as two half-warps of 16 lanes x 8 B; with this padding the stride-16 scatter of pass A
// (17 t + k), the scatter of pass B and both gathers hit 16 distinct 8-byte bank pairs per half-warp, i.e. every
// exchange is bank-conflict free (ncu on the first version, which used split re/im planes and 32-bit accesses,
// showed the kernel L1TEX-bound at 96 % with 108 conflict wavefronts per transform: profiles/r01_*).
//
// Twiddles: only 5 table loads per thread and transform (w^1, w^2, w^4, w^8 of pass B and w_t of pass C); the rest
// are products of at most three exact table values, so the rounding error stays ~3 ulp.
//
// A block spectrum is P ordinary complex values (8 bytes each).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Cache hints for spectra that are written by one kernel and read once by the next (far larger than L2):
// ALR_STREAM_SPECTRA=1 uses streaming (evict-first) stores / loads.
#ifndef ALR_STREAM_SPECTRA
#define ALR_STREAM_SPECTRA 0
#endif
#if ALR_STREAM_SPECTRA
#define ALR_SPEC_STORE(p, v) __stcs((p), (v))
#define ALR_SPEC_LOAD(p) __ldcs((p))
#else
#define ALR_SPEC_STORE(p, v) (*(p) = (v))
#define ALR_SPEC_LOAD(p) (*(p))
#endif

namespace alr {

#ifndef ALR_P
#define ALR_P 4096
#endif
constexpr int kP = ALR_P;                    // partition length in samples == complex FFT size (1024, 2048 or 4096)
static_assert(kP == 1024 || kP == 2048 || kP == 4096, "partition must be 1024, 2048 or 4096");
constexpr int kGroup = kP / 16;              // threads per FFT (each holds 16 elements): 64, 128 or 256
constexpr int kGroupsPerCta = 256 / kGroup;  // FFTs in flight per CTA: 4, 2 or 1
constexpr int kR3 = kP / 256;                // radix of the last pass (16 * 16 * kR3 == kP): 4, 8 or 16
constexpr int kM3 = 16 / kR3;                // last-pass butterflies per thread: 4, 2 or 1
constexpr int kPad = kP + kP / 16;

struct FftSmem {
  float2 d[kPad];
};

__device__ __forceinline__ int padi(int i) { return i + (i >> 4); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// (bit-identical results; 12 % fewer instructions in the FFT kernels, 1 % of the benchmark step: profiles/r01_ffma2.txt)
#ifndef ALR_FADD2
#define ALR_FADD2 1
#endif
#if ALR_FADD2
// complex add / subtract as ONE packed instruction (Blackwell FADD2, add.rn.f32x2 on a 64-bit register pair)
__device__ __forceinline__ unsigned long long f2_bits(float2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ float2 bits_f2(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(r);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(r);
}
#else
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// barrier over one FFT group of kGroup threads (named barriers 1..4; 0 stays __syncthreads)
__device__ __forceinline__ void group_sync(int bar) {
  asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(kGroup) : "memory");
}

// 4-point DFT, natural order in and out. Forward: exp(-i..); INV: exp(+i..)
template <bool INV>
__device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d) {
  float2 s02 = cadd(a, c), d02 = csub(a, c), s13 = cadd(b, d), d13 = csub(b, d);
  float2 r = INV ? make_float2(-d13.y, d13.x) : make_float2(d13.y, -d13.x);  // (+-i) * d13
  a = cadd(s02, s13);
  c = csub(s02, s13);
  b = cadd(d02, r);
  d = csub(d02, r);
}

template <bool INV>
__device__ __forceinline__ float2 mul_w16(float2 v, int m) {
  // multiply by W16^m (forward) or its conjugate (INV); m is a compile-time constant after unrolling
  const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
  float wr, wi;
  switch (m) {
    case 0: return v;
    case 1: wr = c1; wi = -s1; break;
    case 2: wr = h; wi = -h; break;
    case 3: wr = s1; wi = -c1; break;
    case 4: wr = 0.f; wi = -1.f; break;
    case 6: wr = -h; wi = -h; break;
    default: wr = -c1; wi = s1; break;  // m == 9
  }
  if (INV) wi = -wi;
  return make_float2(v.x * wr - v.y * wi, v.x * wi + v.y * wr);
}

// 16-point DFT as 4x4 (n = 4*n1 + n2, k = k1 + 4*k2). Output X[k] ends up in v[4*(k&3) + (k>>2)].
template <bool INV>
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4<INV>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
#pragma unroll
  for (int k1 = 1; k1 < 4; ++k1)
#pragma unroll
    for (int n2 = 1; n2 < 4; ++n2) v[4 * k1 + n2] = mul_w16<INV>(v[4 * k1 + n2], n2 * k1);
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4<INV>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
}
__device__ __forceinline__ constexpr int perm16(int k) { return 4 * (k & 3) + (k >> 2); }

// 8-point DFT as 4x2 (n = 2*n1 + n2, k = k1 + 4*k2). Output X[k] ends up in v[2*(k&3) + (k>>2)].
template <bool INV>
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
  dft4<INV>(v[0], v[2], v[4], v[6]);
  dft4<INV>(v[1], v[3], v[5], v[7]);
  // v[2*k1 + 1] *= W8^k1
  const float h = 0.70710678118654752f;
  {
    float2 a = v[3];  // W8^1 = (h, -h) forward, (h, +h) inverse
    v[3] = INV ? make_float2(h * (a.x - a.y), h * (a.x + a.y)) : make_float2(h * (a.x + a.y), h * (a.y - a.x));
    a = v[5];         // W8^2 = -i forward, +i inverse
    v[5] = INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
    a = v[7];         // W8^3 = (-h, -h) forward, (-h, +h) inverse
    v[7] = INV ? make_float2(-h * (a.x + a.y), h * (a.x - a.y)) : make_float2(h * (a.y - a.x), -h * (a.x + a.y));
  }
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) {
    const float2 a = v[2 * k1], b = v[2 * k1 + 1];
    v[2 * k1] = cadd(a, b);
    v[2 * k1 + 1] = csub(a, b);
  }
}
__device__ __forceinline__ constexpr int perm8(int k) { return 2 * (k & 3) + (k >> 2); }

// P-point complex FFT by one group of kGroup = P/16 threads; Stockham autosort, radix 16 x 16 x kR3.
//   in : v[q][r] = element (t + kGroup r) of transform q, r = 0..15                 (t = thread index in the group)
//   out: o[q][m][k] = spectrum element (t + kGroup m) + 256 k, m < kM3, k < kR3      (natural order, un-normalised)
// tw[m] = exp(-2*pi*i*m/P), m < P.  NT independent transforms (exchange buffers s[0..NT)) are carried through the
// passes TOGETHER: they share the barriers and every twiddle product, and give the scheduler NT independent
// instruction streams per thread — the persistent kernel (k_mov_fused) runs at 2 CTAs per SM and uses NT = 2 to get
// the latency hiding the stand-alone FFT kernels get from 4 CTAs per SM.
// The caller must have a group_sync between any earlier use of `s` by other threads and this call; on return the
// group may still be reading `s` (last-pass gather), so the caller needs a group_sync before the next transform
// scatters into `s`.
template <bool INV, int NT>
__device__ __forceinline__ void fft_core_n(float2 (&v)[NT][16], FftSmem* __restrict__ s, const float2* __restrict__ tw,
                                           int t, int bar, float2 (&o)[NT][kM3][kR3]) {
  // twiddle seeds: issue the table loads first so that their latency hides behind pass A
  const int tq = t & 15;
  constexpr int kB = kP / 256;  // pass-B twiddle exp(-2 pi i tq r / 256) = tw[tq * r * kB]
  // (the seeds come from the compact copy behind the table, same values as tw[kB * tq << j]: the strided originals cost
  // 16 L1 wavefronts per warp request, 40 % of the kernel's global-memory wavefronts)
  static_assert(kB >= 1, "");
  float2 w1 = __ldg(tw + kP + tq), w2 = __ldg(tw + kP + 16 + tq), w4 = __ldg(tw + kP + 32 + tq), w8 = __ldg(tw + kP + 48 + tq);
  float2 wt = __ldg(tw + t);
  if (INV) {
    w1.y = -w1.y; w2.y = -w2.y; w4.y = -w4.y; w8.y = -w8.y; wt.y = -wt.y;
  }
  // ---- pass A: radix 16, Ns = 1, no twiddles; scatter to 16*t + k
#pragma unroll
  for (int q = 0; q < NT; ++q) {
    dft16<INV>(v[q]);
#pragma unroll
    for (int k = 0; k < 16; ++k) s[q].d[padi(16 * t + k)] = v[q][perm16(k)];
  }
  group_sync(bar);
#pragma unroll
  for (int q = 0; q < NT; ++q)
#pragma unroll
    for (int r = 0; r < 16; ++r) v[q][r] = s[q].d[padi(t + kGroup * r)];
  group_sync(bar);
  // ---- pass B: radix 16, Ns = 16; twiddle exp(-2 pi i (t%16) r / 256) = w1^r, built from w1, w2, w4, w8
  {
    const float2 w3 = cmul(w1, w2), w5 = cmul(w1, w4), w6 = cmul(w2, w4), w9 = cmul(w1, w8), w10 = cmul(w2, w8),
                 w12 = cmul(w4, w8);
    const float2 w7 = cmul(w3, w4), w11 = cmul(w3, w8), w13 = cmul(w5, w8), w14 = cmul(w6, w8);
    const float2 w15 = cmul(w7, w8);
#pragma unroll
    for (int q = 0; q < NT; ++q) {
      float2 (&u)[16] = v[q];
      u[1] = cmul(u[1], w1);   u[2] = cmul(u[2], w2);   u[3] = cmul(u[3], w3);   u[4] = cmul(u[4], w4);
      u[5] = cmul(u[5], w5);   u[6] = cmul(u[6], w6);   u[7] = cmul(u[7], w7);   u[8] = cmul(u[8], w8);
      u[9] = cmul(u[9], w9);   u[10] = cmul(u[10], w10); u[11] = cmul(u[11], w11); u[12] = cmul(u[12], w12);
      u[13] = cmul(u[13], w13); u[14] = cmul(u[14], w14); u[15] = cmul(u[15], w15);
    }
  }
  const int base = (t >> 4) * 256 + tq;
#pragma unroll
  for (int q = 0; q < NT; ++q) {
    dft16<INV>(v[q]);
#pragma unroll
    for (int k = 0; k < 16; ++k) s[q].d[padi(base + 16 * k)] = v[q][perm16(k)];
  }
  group_sync(bar);
  // ---- pass C: radix kR3, Ns = 256; butterfly j = t + kGroup m; twiddle exp(-2 pi i j r / P) = (wt * c_m)^r,
  //      c_m = exp(-2 pi i kGroup m / P) = exp(-i pi m / 8)
#pragma unroll
  for (int q = 0; q < NT; ++q)
#pragma unroll
    for (int m = 0; m < kM3; ++m) {
      const int j = t + kGroup * m;
#pragma unroll
      for (int r = 0; r < kR3; ++r) o[q][m][r] = s[q].d[padi(j + 256 * r)];
    }
#pragma unroll
  for (int m = 0; m < kM3; ++m) {
    const float cr[4] = {1.f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f};
    const float ci[4] = {0.f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f};
    const float2 cm = make_float2(cr[m], INV ? ci[m] : -ci[m]);
    const float2 u1 = (m == 0) ? wt : cmul(wt, cm);
    const float2 u2 = cmul(u1, u1);
    const float2 u3 = cmul(u2, u1);
    if constexpr (kR3 == 16) {
      const float2 u4 = cmul(u2, u2), u8 = cmul(u4, u4);
      const float2 u5 = cmul(u4, u1), u6 = cmul(u4, u2), u7 = cmul(u4, u3);
      const float2 u9 = cmul(u8, u1), u10 = cmul(u8, u2), u11 = cmul(u8, u3), u12 = cmul(u8, u4), u13 = cmul(u8, u5),
                   u14 = cmul(u8, u6), u15 = cmul(u8, u7);
#pragma unroll
      for (int q = 0; q < NT; ++q) {
        float2 (&x)[kR3] = o[q][m];
        x[1] = cmul(x[1], u1);   x[2] = cmul(x[2], u2);   x[3] = cmul(x[3], u3);   x[4] = cmul(x[4], u4);
        x[5] = cmul(x[5], u5);   x[6] = cmul(x[6], u6);   x[7] = cmul(x[7], u7);   x[8] = cmul(x[8], u8);
        x[9] = cmul(x[9], u9);   x[10] = cmul(x[10], u10); x[11] = cmul(x[11], u11); x[12] = cmul(x[12], u12);
        x[13] = cmul(x[13], u13); x[14] = cmul(x[14], u14); x[15] = cmul(x[15], u15);
        dft16<INV>(x);
        float2 tmp[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) tmp[k] = x[perm16(k)];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = tmp[k];
      }
    } else if constexpr (kR3 == 8) {
      const float2 u4 = cmul(u2, u2);
      const float2 u5 = cmul(u4, u1), u6 = cmul(u4, u2), u7 = cmul(u4, u3);
#pragma unroll
      for (int q = 0; q < NT; ++q) {
        float2 (&x)[kR3] = o[q][m];
        x[1] = cmul(x[1], u1);   x[2] = cmul(x[2], u2);   x[3] = cmul(x[3], u3);   x[4] = cmul(x[4], u4);
        x[5] = cmul(x[5], u5);   x[6] = cmul(x[6], u6);   x[7] = cmul(x[7], u7);
        dft8<INV>(x);
        // natural order: X[k] sits in slot perm8(k)
        float2 tmp[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) tmp[k] = x[perm8(k)];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = tmp[k];
      }
    } else {
#pragma unroll
      for (int q = 0; q < NT; ++q) {
        float2 (&x)[kR3] = o[q][m];
        x[1] = cmul(x[1], u1);   x[2] = cmul(x[2], u2);   x[3] = cmul(x[3], u3);
        dft4<INV>(x[0], x[1], x[2], x[3]);
      }
    }
  }
}

template <bool INV>
__device__ __forceinline__ void fft_core(float2 (&v)[16], FftSmem& s, const float2* __restrict__ tw, int t, int bar,
                                         float2 (&o)[kM3][kR3]) {
  fft_core_n<INV, 1>(reinterpret_cast<float2 (&)[1][16]>(v), &s, tw, t, bar,
                     reinterpret_cast<float2 (&)[1][kM3][kR3]>(o));
}

// exp(+i*pi*kGroup*r/(2P)) = exp(i*pi*r/32): the r-dependent factor of the twist zeta^(t + kGroup r) (kGroup = P/16),
// compile-time after unrolling
__device__ __forceinline__ float2 zeta_step(int r) {
  constexpr float c[16] = {1.0f,          0.99518472667f, 0.98078528040f, 0.95694033573f, 0.92387953251f, 0.88192126435f,
                           0.83146961230f, 0.77301045336f, 0.70710678119f, 0.63439328416f, 0.55557023302f, 0.47139673683f,
                           0.38268343237f, 0.29028467725f, 0.19509032202f, 0.09801714033f};
  constexpr float sn[16] = {0.0f,          0.09801714033f, 0.19509032202f, 0.29028467725f, 0.38268343237f, 0.47139673683f,
                            0.55557023302f, 0.63439328416f, 0.70710678119f, 0.77301045336f, 0.83146961230f, 0.88192126435f,
                            0.92387953251f, 0.95694033573f, 0.98078528040f, 0.99518472667f};
  return make_float2(c[r], sn[r]);
}

// ---- negacyclic ("fold + twist") block transform ---------------------------------------------------------------------
// A real block a[0..2P) is mapped to z[n] = (a[n] + i a[n+P]) * zeta^n, zeta = exp(i pi / 2P), n < P, followed by a
// plain P-point complex FFT.  Pointwise products of such spectra give the NEGACYCLIC convolution of the 2P-blocks,
// which equals the linear convolution when both blocks are zero in their second half (P + P - 1 < 2P): exactly the
// uniform-partitioned overlap-add setting.  Every one of the P bins is an ordinary complex number (no packed
// DC/Nyquist bin, no real-FFT untangle pass), and after the inverse the first half of the block is the real part and
// the overlap tail the imaginary part of the SAME element.
//
// Forward: the caller passes the real samples a[t + kGroup r] in a[r] (second half implicitly zero); zt = zeta^t.
__device__ __forceinline__ void fwd_block_to_global(const float (&a)[16], float2 zt, FftSmem& s,
                                                    const float2* __restrict__ tw, int t, int bar,
                                                    float2* __restrict__ spec, int stride256 = 256) {
  float2 v[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const float2 z = (r == 0) ? zt : cmul(zt, zeta_step(r));
    v[r] = make_float2(a[r] * z.x, a[r] * z.y);
  }
  float2 o[kM3][kR3];
  fft_core<false>(v, s, tw, t, bar, o);
#pragma unroll
  for (int m = 0; m < kM3; ++m)
#pragma unroll
    for (int k = 0; k < kR3; ++k) ALR_SPEC_STORE(spec + t + kGroup * m + stride256 * k, o[m][k]);  // stride256: distance of the 256-bin tiles
  group_sync(bar);  // pass-C gathers done before the next transform's scatter
}

// NT blocks at once (fft_core_n): a[q] = real samples of block q, spec[q] = destination or nullptr (block not stored).
template <int NT>
__device__ __forceinline__ void fwd_blocks_to_global(const float (&a)[NT][16], float2 zt, FftSmem* __restrict__ s,
                                                     const float2* __restrict__ tw, int t, int bar,
                                                     float2* (&spec)[NT], int stride256 = 256) {
  float2 v[NT][16];
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const float2 z = (r == 0) ? zt : cmul(zt, zeta_step(r));
#pragma unroll
    for (int q = 0; q < NT; ++q) v[q][r] = make_float2(a[q][r] * z.x, a[q][r] * z.y);
  }
  float2 o[NT][kM3][kR3];
  fft_core_n<false, NT>(v, s, tw, t, bar, o);
#pragma unroll
  for (int q = 0; q < NT; ++q)
    if (spec[q]) {
#pragma unroll
      for (int m = 0; m < kM3; ++m)
#pragma unroll
        for (int k = 0; k < kR3; ++k) __stcg(spec[q] + t + kGroup * m + stride256 * k, o[q][m][k]);
    }
  group_sync(bar);  // pass-C gathers done before the next transform's scatter
}

// Inverse: spectrum (global) -> o[m][k] = un-normalised z[e] * conj(zeta^e), e = t + kGroup m + 256 k:
// real part = block sample e (first half), imaginary part = block sample P + e (overlap tail). Scale by 1/P.
__device__ __forceinline__ void inv_block_from_global(const float2* __restrict__ spec, float2 zt, FftSmem& s,
                                                      const float2* __restrict__ tw, int t, int bar,
                                                      float2 (&o)[kM3][kR3]) {
  float2 v[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = ALR_SPEC_LOAD(spec + t + kGroup * r);
  fft_core<true>(v, s, tw, t, bar, o);
  const float2 ztc = cconj(zt);
#pragma unroll
  for (int m = 0; m < kM3; ++m)
#pragma unroll
    for (int k = 0; k < kR3; ++k) {
      const int r = m + (256 / kGroup) * k;  // e = t + kGroup r
      const float2 z = (r == 0) ? ztc : cmul(ztc, cconj(zeta_step(r)));
      o[m][k] = cmul(o[m][k], z);
    }
  group_sync(bar);
}

// Register-to-register variants (k_small_rir keeps the RIR spectrum and every intermediate spectrum in registers):
// forward: real samples a[r] -> spectrum o[m][k] (element t + kGroup m + 256 k); ends with a group barrier.
__device__ __forceinline__ void fwd_block_to_regs(const float (&a)[16], float2 zt, FftSmem& s, const float2* __restrict__ tw,
                                                  int t, int bar, float2 (&o)[kM3][kR3]) {
  float2 v[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const float2 z = (r == 0) ? zt : cmul(zt, zeta_step(r));
    v[r] = make_float2(a[r] * z.x, a[r] * z.y);
  }
  fft_core<false>(v, s, tw, t, bar, o);
  group_sync(bar);
}
// inverse: spectrum v[r] (element t + kGroup r) -> o[m][k] = un-normalised (block sample e, overlap-tail sample P + e),
// e = t + kGroup m + 256 k; ends with a group barrier.
__device__ __forceinline__ void inv_block_from_regs(float2 (&v)[16], float2 zt, FftSmem& s, const float2* __restrict__ tw,
                                                    int t, int bar, float2 (&o)[kM3][kR3]) {
  fft_core<true>(v, s, tw, t, bar, o);
  const float2 ztc = cconj(zt);
#pragma unroll
  for (int m = 0; m < kM3; ++m)
#pragma unroll
    for (int k = 0; k < kR3; ++k) {
      const int r = m + (256 / kGroup) * k;  // e = t + kGroup r
      const float2 z = (r == 0) ? ztc : cmul(ztc, cconj(zeta_step(r)));
      o[m][k] = cmul(o[m][k], z);
    }
  group_sync(bar);
}

}  // namespace alr
