// alr_fft.cuh — shared-memory Stockham FFT core (sm_100a), used by every spectral kernel of the renderer.
//
// A real block of 2P samples is transformed as one P-point complex FFT of the even/odd-packed signal plus an
// "untangle" step (standard real-FFT trick), so the FFT that replaces scipy's pocketfft calls of the reference
// (scipy.fft.rfft/irfft, synthesize.py:138,267 and scipy.signal.fftconvolve, :103,490) is a P = 1024 point
// complex transform done by a group of 64 threads: Stockham autosort passes of radix 16, 16 and 4, butterflies
// in registers, two exchanges through padded shared memory (split re/im planes, 1 pad word per 32 -> the
// stride-16 scatter of the first pass is bank-conflict free).
//
// Half spectra are stored PACKED: P complex values per block, bin 0 = (Re X[0], Re X[P]) (both are real).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace alr {

constexpr int kP = 1024;          // partition length in samples == complex FFT size
constexpr int kGroup = 64;        // threads per FFT
constexpr int kGroupsPerCta = 4;  // FFTs in flight per CTA
constexpr int kPad = kP + kP / 32;

struct FftSmem {
  float re[kPad];
  float im[kPad];
};

__device__ __forceinline__ int padi(int i) { return i + (i >> 5); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// barrier over one 64-thread FFT group (named barriers 1..4; 0 stays __syncthreads)
__device__ __forceinline__ void group_sync(int bar) {
  asm volatile("bar.sync %0, 64;" ::"r"(bar) : "memory");
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

// P-point complex FFT by one 64-thread group.
//   in : v[r] = element (t + 64 r), r = 0..15                         (t = thread index in the group)
//   out: o[m][k] = spectrum element (t + 64 m) + 256 k, m,k = 0..3     (natural order, un-normalised)
// tw[m] = exp(-2*pi*i*m/(2P)), m < 2P.  The caller must have a group_sync between any earlier use of `s` by
// other threads and this call; on return the group may still be reading `s` (pass C), but only at indices that
// the reading thread owns, so the caller may overwrite its own o-indices without a further barrier.
template <bool INV>
__device__ __forceinline__ void fft_core(float2 (&v)[16], FftSmem& s, const float2* __restrict__ tw, int t, int bar,
                                         float2 (&o)[4][4]) {
  // ---- pass A: radix 16, Ns = 1, no twiddles; scatter to 16*t + k
  dft16<INV>(v);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    int i = padi(16 * t + k);
    s.re[i] = v[perm16(k)].x;
    s.im[i] = v[perm16(k)].y;
  }
  group_sync(bar);
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    int i = padi(t + 64 * r);
    v[r] = make_float2(s.re[i], s.im[i]);
  }
  group_sync(bar);
  // ---- pass B: radix 16, Ns = 16; twiddle exp(-2 pi i (t%16) r / 256) = tw[(t%16)*r*8]
  const int tq = t & 15;
#pragma unroll
  for (int r = 1; r < 16; ++r) {
    float2 w = __ldg(tw + tq * r * 8);
    if (INV) w.y = -w.y;
    v[r] = cmul(v[r], w);
  }
  dft16<INV>(v);
  const int base = (t >> 4) * 256 + tq;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    int i = padi(base + 16 * k);
    s.re[i] = v[perm16(k)].x;
    s.im[i] = v[perm16(k)].y;
  }
  group_sync(bar);
  // ---- pass C: radix 4, Ns = 256; butterfly j = t + 64 m; twiddle exp(-2 pi i j r / 1024) = tw[2 j r]
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int j = t + 64 * m;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int i = padi(j + 256 * r);
      o[m][r] = make_float2(s.re[i], s.im[i]);
    }
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int j = t + 64 * m;
#pragma unroll
    for (int r = 1; r < 4; ++r) {
      float2 w = __ldg(tw + 2 * j * r);
      if (INV) w.y = -w.y;
      o[m][r] = cmul(o[m][r], w);
    }
    dft4<INV>(o[m][0], o[m][1], o[m][2], o[m][3]);
  }
}

// Forward real FFT of a zero-padded block: the caller provides the P/2 packed complex inputs
// z[i] = (x[2i], x[2i+1]) for i = t + 64 r, r = 0..7 in v[0..7] (v[8..15] are the zero padding), and receives
// the packed half spectrum in global memory at `spec` (P float2).
__device__ __forceinline__ void rfft_block_to_global(float2 (&v)[16], FftSmem& s, const float2* __restrict__ tw,
                                                     int t, int bar, float2* __restrict__ spec) {
  float2 o[4][4];
  fft_core<false>(v, s, tw, t, bar, o);
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int i = padi(t + 64 * m + 256 * k);  // indices owned by this thread in pass C
      s.re[i] = o[m][k].x;
      s.im[i] = o[m][k].y;
    }
  group_sync(bar);
  // untangle: X[k] = E + W^k O, X[P-k] = conj(E - W^k O), E = (Z[k] + conj Z[P-k])/2, O = -i (Z[k] - conj Z[P-k])/2
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int k = t + 64 * m;  // 0 .. 511
    if (k == 0) {
      float zr = s.re[0], zi = s.im[0];
      spec[0] = make_float2(zr + zi, zr - zi);
      // k = P/2 pairs with itself: X[P/2] = conj(Z[P/2])
      int ih = padi(kP / 2);
      spec[kP / 2] = make_float2(s.re[ih], -s.im[ih]);
    } else {
      int i1 = padi(k), i2 = padi(kP - k);
      float2 zk = make_float2(s.re[i1], s.im[i1]);
      float2 zc = make_float2(s.re[i2], -s.im[i2]);
      float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
      float2 d = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y - zc.y));
      float2 od = make_float2(d.y, -d.x);  // -i * d
      float2 wo = cmul(__ldg(tw + k), od);
      spec[k] = cadd(e, wo);
      spec[kP - k] = cconj(csub(e, wo));
    }
  }
  group_sync(bar);  // smem free for the next transform
}

// Inverse of the above: packed half spectrum (global) -> o[m][k] holding complex z[(t+64m) + 256k] with
// real block sample 2i = Re z[i], 2i+1 = Im z[i]; un-normalised (scale by 1/(2P) for a true inverse).
__device__ __forceinline__ void irfft_block_from_global(const float2* __restrict__ spec, FftSmem& s,
                                                        const float2* __restrict__ tw, int t, int bar,
                                                        float2 (&o)[4][4]) {
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int k = t + 64 * m;
    if (k == 0) {
      float2 x0 = spec[0];
      s.re[0] = x0.x + x0.y;
      s.im[0] = x0.x - x0.y;
      float2 xh = spec[kP / 2];
      int ih = padi(kP / 2);
      s.re[ih] = 2.f * xh.x;  // same 2x scale as the E/O sums below
      s.im[ih] = -2.f * xh.y;
    } else {
      float2 a = spec[k];
      float2 b = cconj(spec[kP - k]);
      float2 e = cadd(a, b);
      float2 wo = csub(a, b);
      float2 od = cmul(wo, cconj(__ldg(tw + k)));
      // Z[k] = E + i O ; Z[P-k] = conj(E - i O)
      int i1 = padi(k), i2 = padi(kP - k);
      s.re[i1] = e.x - od.y;
      s.im[i1] = e.y + od.x;
      s.re[i2] = e.x + od.y;
      s.im[i2] = -(e.y - od.x);
    }
  }
  group_sync(bar);
  float2 v[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    int i = padi(t + 64 * r);
    v[r] = make_float2(s.re[i], s.im[i]);
  }
  group_sync(bar);
  fft_core<true>(v, s, tw, t, bar, o);
  group_sync(bar);  // pass-C reads done before the next transform overwrites smem
}

}  // namespace alr
