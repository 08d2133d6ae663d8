// Packed warp-level FFT engine (sm_100a): every lane register holds the SAME quantity of TWO frames in
// the two halves of a 64-bit register pair, and all arithmetic is FFMA2 / FADD2 / FMUL2 (fma.rn.f32x2 ...).
//
// Measured on B200 (tools/ubench): a packed instruction occupies the FMA pipe for 2 cycles but only ONE
// issue slot, and it takes 32-bit immediates / scalar registers broadcast to both halves for free.  The
// scalar engine (fftcore.cuh) was issue-bound; the packed one leaves half of the issue slots to LDS / STS /
// MUFU / integer work, so the kernel becomes FMA-pipe bound.
//
// Transform (same mathematics as fftcore.cuh): a frame is win = N/2 window taps a[m] centred in N points,
//   z[n] = a[2n] + i a[2n+1]   (n < Nz/2, Nz = N/2; upper half of the Nz-point input is zero)
//   Z = DFT_Nz(z) = pass A (in-lane radix-2R DIT, first stage pruned) -> twiddle -> 32x32 transpose through
//       shared memory (two planes of 8-byte packed reals) -> pass B (in-lane radix-32 DIT)
//   split: A[k] = Zk + conj Zr + g_k (Zk - conj Zr), A[Nz-k] = conj(Zk + conj Zr - g_k (Zk - conj Zr)),
//          Zr = Z[Nz-k], g_k = -i w_N^k, with the 1/2 folded into the window table;  X[k] = (-i)^k A[k].
// Butterflies are in Linzer-Feig form: w b = c (b.re + t b.im, b.im - t b.re), t = tan, and the scale c is
// absorbed into the following a +- c t' FMAs: 6 FMAs per non-trivial radix-2 butterfly.
//
// One warp "item" is 32 lanes x 32 packed complex registers = P = 1024/N... pairs of frames:
//   n_fft 2048: 1 pair (2 frames), 1024: 2 pairs (4 frames), 512: 4 pairs (8 frames).
#pragma once
#include <cuda_runtime.h>

#include "fftcore.cuh"

namespace sb200 {

typedef unsigned long long pf;   // packed float pair: lo = first frame of the pair, hi = second

__device__ __forceinline__ pf pk(float lo, float hi) {
  pf r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float plo(pf a) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
  return lo;
}
__device__ __forceinline__ float phi(pf a) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
  return hi;
}
__device__ __forceinline__ pf add2(pf a, pf b) {
  pf d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ pf sub2(pf a, pf b) {
  pf d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ pf mul2(pf a, pf b) {
  pf d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ pf fma2(pf a, pf b, pf c) {
  pf d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// scalar (immediate or register) broadcast to both halves: ptxas folds the mov.b64 into the operand
__device__ __forceinline__ pf mul2s(pf a, float s) { return mul2(a, pk(s, s)); }
__device__ __forceinline__ pf fma2s(pf a, float s, pf c) { return fma2(a, pk(s, s), c); }

struct PC {   // packed complex: (re, im) of two frames
  pf re, im;
};

__device__ __forceinline__ PC pc_shfl(const PC& a, int src) {
  PC r;
  r.re = __shfl_sync(kFullMask, a.re, src);
  r.im = __shfl_sync(kFullMask, a.im, src);
  return r;
}
__device__ __forceinline__ PC pc_sel(bool c, const PC& a, const PC& b) {
  PC r;
  r.re = c ? a.re : b.re;
  r.im = c ? a.im : b.im;
  return r;
}

// 16-byte shared-memory access of one packed complex (re pair, im pair); OFF = compile-time byte offset
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
template <int OFF>
__device__ __forceinline__ void sts_pc(unsigned addr, pf re, pf im) {
  asm volatile("st.shared.v2.b64 [%0+%3], {%1, %2};" ::"r"(addr), "l"(re), "l"(im), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ PC lds_pc(unsigned addr) {
  PC r;
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2+%3];" : "=l"(r.re), "=l"(r.im) : "r"(addr), "n"(OFF) : "memory");
  return r;
}

// 8-byte accesses of one packed real (both frames of a pair).  The forward transpose keeps the real and imaginary parts
// in two planes of 8-byte elements: a 16-byte st.shared.v2.b64 needs its four source registers consecutive, which costs
// four MOVs per store after packed arithmetic (ptxas does not allocate the producers into quads).
template <int OFF>
__device__ __forceinline__ void sts_pf(unsigned addr, pf x) {
  asm volatile("st.shared.b64 [%0+%2], %1;" ::"r"(addr), "l"(x), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ pf lds_pf(unsigned addr) {
  pf r;
  asm volatile("ld.shared.b64 %0, [%1+%2];" : "=l"(r) : "r"(addr), "n"(OFF) : "memory");
  return r;
}

// cos / sin of 2 pi i / 32 in double (compile-time), i in [0, 16]
__host__ __device__ constexpr double dcos32_q(int i) {   // i in [0, 8]
  return i == 0 ? 1.0
       : i == 1 ? 0.98078528040323044913
       : i == 2 ? 0.92387953251128675613
       : i == 3 ? 0.83146961230254523708
       : i == 4 ? 0.70710678118654752440
       : i == 5 ? 0.55557023301960222474
       : i == 6 ? 0.38268343236508977173
       : i == 7 ? 0.19509032201612826785
                : 0.0;
}
__host__ __device__ constexpr double dcos32(int i) { return i <= 8 ? dcos32_q(i) : -dcos32_q(16 - i); }
__host__ __device__ constexpr double dsin32(int i) { return i <= 8 ? dcos32_q(8 - i) : dcos32_q(i - 8); }

// Radix-2 DIT butterfly with twiddle w_32^I (forward: e^{-2 pi i I/32}; INV: conjugate), I in [0, 16):
//   a' = a + w b,  b' = a - w b
template <int I, bool INV>
__device__ __forceinline__ void bfly(PC& a, PC& b) {
  static_assert(I >= 0 && I < 16, "twiddle index");
  if constexpr (I == 0) {
    const PC t = b;
    b.re = sub2(a.re, t.re); b.im = sub2(a.im, t.im);
    a.re = add2(a.re, t.re); a.im = add2(a.im, t.im);
  } else if constexpr (I == 8) {
    // forward: w b = -i b = (b.im, -b.re); inverse: +i b = (-b.im, b.re)
    const PC t = b;
    if constexpr (!INV) {
      b.re = sub2(a.re, t.im); b.im = add2(a.im, t.re);
      a.re = add2(a.re, t.im); a.im = sub2(a.im, t.re);
    } else {
      b.re = add2(a.re, t.im); b.im = sub2(a.im, t.re);
      a.re = sub2(a.re, t.im); a.im = add2(a.im, t.re);
    }
  } else {
    constexpr double c = dcos32(I), s = INV ? -dsin32(I) : dsin32(I);   // w = c - i s
    // w b = (c b.re + s b.im, c b.im - s b.re)
    if constexpr ((c < 0 ? -c : c) >= (s < 0 ? -s : s)) {
      constexpr float t = static_cast<float>(s / c), cf = static_cast<float>(c);
      const pf tr = fma2s(b.im, t, b.re);     // (w b).re / c
      const pf ti = fma2s(b.re, -t, b.im);    // (w b).im / c
      b.re = fma2s(tr, -cf, a.re); b.im = fma2s(ti, -cf, a.im);
      a.re = fma2s(tr, cf, a.re);  a.im = fma2s(ti, cf, a.im);
    } else {
      constexpr float k = static_cast<float>(c / s), sf = static_cast<float>(s);
      const pf tr = fma2s(b.re, k, b.im);     // (w b).re / s
      const pf tn = fma2s(b.im, -k, b.re);    // -(w b).im / s
      b.re = fma2s(tr, -sf, a.re); b.im = fma2s(tn, sf, a.im);
      a.re = fma2s(tr, sf, a.re);  a.im = fma2s(tn, -sf, a.im);
    }
  }
}

// In-register radix-2 DIT DFT of LEN points v[BASE .. BASE+LEN).  Input in bit-reversed order
// (v[BASE + brev(n)] = x[n]), output in natural order.  Stages with block size < MIN_M are skipped:
// MIN_M = 4 with v[BASE+2i+1] == v[BASE+2i] is the transform of an input whose upper half is zero.
template <int LEN, int BASE, bool INV, int MIN_M>
__device__ __forceinline__ void dit(PC (&v)[32]) {
  if constexpr (LEN > 1) {
    constexpr int H = LEN / 2;
    dit<H, BASE, INV, MIN_M>(v);
    dit<H, BASE + H, INV, MIN_M>(v);
    if constexpr (LEN >= MIN_M) {
      static_for<0, H>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        bfly<j * (32 / LEN), INV>(v[BASE + j], v[BASE + j + H]);
      });
    }
  }
}

template <int N>
struct Fft2Cfg {
  static_assert(N == 2048 || N == 1024 || N == 512, "supported n_fft: 512, 1024, 2048 (win = n_fft/2)");
  static constexpr int kN = N;
  static constexpr int kNz = N / 2;          // complex FFT length
  static constexpr int kWin = N / 2;         // window support
  static constexpr int kR = N / 128;         // non-zero pass-A inputs per lane per frame
  static constexpr int kR2 = 2 * kR;         // pass-A radix (32 / 16 / 8)
  static constexpr int kLogR2 = ilog2(kR2);
  static constexpr int kP = 2048 / N;        // frame PAIRS per item (1 / 2 / 4)
  static constexpr int kFrames = 2 * kP;     // frames per item
  static constexpr int kF = N / 2 + 1;       // one-sided bins
  static constexpr int kTwCount = (kR2 - 1) * 32;   // float2 w_Nz^{k1*lane}, row k1-1
  static constexpr int kSplitSlots = 17;     // pair slots per lane: k2 = 0..15 and the self pair k2 = 16 of column 0
  // transpose buffer: element (row, col) of 16 bytes at 33*row + col: the 8 rows of a quarter-warp start in 8
  // different 16-byte bank groups, so both the row-wise stores and the column-wise loads are conflict free
  __host__ __device__ static constexpr int xoff(int row) { return row * 33; }
  static constexpr int kXElems = 32 * 33;
  static constexpr int kXBytes = kXElems * 16;   // per-warp exchange buffer
  static constexpr int kPlane = kXElems * 8;     // forward transpose: imaginary plane offset (8-byte elements, same 33 padding)
};

// ---- pass A + twiddle + transpose + pass B: v (pass-A inputs, see below) -> v[k2] = Z_p[k1 + R2 k2] --------
// On entry v[p*R2 + brev(r)] = v[p*R2 + brev(r) + 1] = z_p[lane + 32 r], r < R (brev over log2(R2) bits).
// On exit lane j = p*R2 + k1 holds Z_p[k1 + R2*k2] in v[k2], k2 = 0..31.  xbuf: this warp's exchange buffer.
// tw: smem table tw[(k1-1)*32 + lane] = w_Nz^{k1*lane}.
// xbuf_free(): called once every lane has read its transposed column, i.e. as soon as the exchange buffer may be reused
// (feat3.cuh issues the bulk-tensor copy of the NEXT item's samples into it there, under pass B and the split).
// EARLY = true: the warp synchronises right after the transposed read and calls xbuf_free() before pass B; false: pass B first
// (its leading butterflies overlap the tail of the loads), synchronisation at the end, no hook.
template <int N, bool EARLY, class Hook>
__device__ __forceinline__ void fft2_forward_h(PC (&v)[32], uint4* __restrict__ xbuf, const float2* __restrict__ tw,
                                               int lane, Hook&& xbuf_free) {
  using C = Fft2Cfg<N>;
  static_for<0, C::kP>([&](auto pc_) {
    constexpr int p = decltype(pc_)::value;
    dit<C::kR2, p * C::kR2, false, 4>(v);
  });
  // twiddle (shared by all pairs) + transposed store: element (row = lane, col = p*R2 + k1).  The twiddles are read in
  // batches of 8, one batch ahead of their use: the stores are volatile asm with a memory clobber, so a table load
  // written after a store is never hoisted above it and would cost one shared-memory round trip per column.
  const unsigned wrow = smem_u32(xbuf) + 8 * C::xoff(lane);
  constexpr int kTwBatch = 8;
  float2 wb[2][kTwBatch];   // batch b lives in wb[b & 1] (compile-time index: no copies between batches)
  static_for<0, kTwBatch>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    if constexpr (j >= 1 && j < C::kR2) wb[0][j] = tw[(j - 1) * 32 + lane];
  });
  static_for<0, (C::kR2 + kTwBatch - 1) / kTwBatch>([&](auto bc) {
    constexpr int b = decltype(bc)::value;
    static_for<0, kTwBatch>([&](auto jc) {
      constexpr int j = decltype(jc)::value;
      constexpr int k1n = (b + 1) * kTwBatch + j;
      if constexpr (k1n < C::kR2) wb[(b + 1) & 1][j] = tw[(k1n - 1) * 32 + lane];
    });
    static_for<0, kTwBatch>([&](auto jc) {
      constexpr int j = decltype(jc)::value;
      constexpr int k1 = b * kTwBatch + j;
      if constexpr (k1 == 0) {
        static_for<0, C::kP>([&](auto pc_) {
          constexpr int p = decltype(pc_)::value;
          sts_pf<8 * (p * C::kR2)>(wrow, v[p * C::kR2].re);
          sts_pf<C::kPlane + 8 * (p * C::kR2)>(wrow, v[p * C::kR2].im);
        });
      } else if constexpr (k1 < C::kR2) {
        const float2 w = wb[b & 1][j];
        static_for<0, C::kP>([&](auto pc_) {
          constexpr int p = decltype(pc_)::value;
          const PC& y = v[p * C::kR2 + k1];
          const pf re = fma2s(y.im, -w.y, mul2s(y.re, w.x));
          const pf im = fma2s(y.im, w.x, mul2s(y.re, w.y));
          sts_pf<8 * (p * C::kR2 + k1)>(wrow, re);
          sts_pf<C::kPlane + 8 * (p * C::kR2 + k1)>(wrow, im);
        });
      }
    });
  });
  __syncwarp();
  // transposed read: lane j takes column j of every row n1, placed bit-reversed for the DIT
  const unsigned rcol = smem_u32(xbuf) + 8 * lane;
  static_for<0, 32>([&](auto nc) {
    constexpr int n1 = decltype(nc)::value;
    v[brev(n1, 5)].re = lds_pf<8 * C::xoff(n1)>(rcol);
    v[brev(n1, 5)].im = lds_pf<C::kPlane + 8 * C::xoff(n1)>(rcol);
  });
  if constexpr (EARLY) {
    __syncwarp();   // exchange buffer free again
    xbuf_free();
    dit<32, 0, false, 2>(v);
  } else {
    dit<32, 0, false, 2>(v);
    __syncwarp();   // exchange buffer free again
  }
}
template <int N>
__device__ __forceinline__ void fft2_forward(PC (&v)[32], uint4* __restrict__ xbuf, const float2* __restrict__ tw,
                                             int lane) {
  fft2_forward_h<N, false>(v, xbuf, tw, lane, [] {});
}

// ---- Hermitian split on packed data -------------------------------------------------------------------------
// Zk, Zr = Z[k], Z[Nz-k];  (t, c) = L-F twiddle of g_k = -i w_N^k = (-s, -c') with phi = 2 pi k / N:
//   SINFORM == false (phi <= pi/4): tw = (tan phi, cos phi);  true (phi >= pi/4): tw = (cot phi, sin phi).
// Returns A[k] in ak and conj(A[Nz-k]) in am (imaginary part of am has the sign of the conjugate).
template <bool SINFORM>
__device__ __forceinline__ void split2(const PC& Zk, const PC& Zr, float2 tw, PC& ak, PC& am) {
  const pf fer = add2(Zk.re, Zr.re), fei = sub2(Zk.im, Zr.im);   // Zk + conj Zr
  const pf fr = sub2(Zk.re, Zr.re), fi = add2(Zk.im, Zr.im);     // Zk - conj Zr
  // t = g fo = (c fi - s fr, -s fi - c fr)
  if constexpr (!SINFORM) {
    const pf tr = fma2s(fr, -tw.x, fi);    // t.re / c
    const pf tn = fma2s(fi, tw.x, fr);     // -t.im / c
    ak.re = fma2s(tr, tw.y, fer);  ak.im = fma2s(tn, -tw.y, fei);
    am.re = fma2s(tr, -tw.y, fer); am.im = fma2s(tn, tw.y, fei);
  } else {
    const pf nr = fma2s(fi, -tw.x, fr);    // -t.re / s = fr - cot fi
    const pf tn = fma2s(fr, tw.x, fi);     // -t.im / s = fi + cot fr
    ak.re = fma2s(nr, -tw.y, fer); ak.im = fma2s(tn, -tw.y, fei);
    am.re = fma2s(nr, tw.y, fer);  am.im = fma2s(tn, tw.y, fei);
  }
}

__device__ __forceinline__ pf norm2(const PC& a) { return fma2(a.re, a.re, mul2(a.im, a.im)); }

// ======================================= inverse direction ==========================================================
// Half butterfly a' = a + w b (the b' = a - w b output is pruned), same twiddle convention as bfly<I, INV>.
template <int I, bool INV>
__device__ __forceinline__ void bfly_half(PC& a, const PC& b) {
  static_assert(I >= 0 && I < 16, "twiddle index");
  if constexpr (I == 0) {
    a.re = add2(a.re, b.re); a.im = add2(a.im, b.im);
  } else if constexpr (I == 8) {
    if constexpr (!INV) { a.re = add2(a.re, b.im); a.im = sub2(a.im, b.re); }
    else { a.re = sub2(a.re, b.im); a.im = add2(a.im, b.re); }
  } else {
    constexpr double c = dcos32(I), s = INV ? -dsin32(I) : dsin32(I);   // w = c - i s
    if constexpr ((c < 0 ? -c : c) >= (s < 0 ? -s : s)) {
      constexpr float t = static_cast<float>(s / c), cf = static_cast<float>(c);
      const pf tr = fma2s(b.im, t, b.re);
      const pf ti = fma2s(b.re, -t, b.im);
      a.re = fma2s(tr, cf, a.re); a.im = fma2s(ti, cf, a.im);
    } else {
      constexpr float k = static_cast<float>(c / s), sf = static_cast<float>(s);
      const pf tr = fma2s(b.re, k, b.im);
      const pf tn = fma2s(b.im, -k, b.re);
      a.re = fma2s(tr, sf, a.re); a.im = fma2s(tn, -sf, a.im);
    }
  }
}

// Radix-2 DIT DFT of LEN points whose upper half of OUTPUTS is not needed: v[BASE + j], j < LEN/2, on exit.
template <int LEN, int BASE, bool INV>
__device__ __forceinline__ void dit_pruned_out(PC (&v)[32]) {
  constexpr int H = LEN / 2;
  dit<H, BASE, INV, 2>(v);
  dit<H, BASE + H, INV, 2>(v);
  static_for<0, H>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    bfly_half<j * (32 / LEN), INV>(v[BASE + j], v[BASE + j + H]);
  });
}

// ---- inverse of fft2_forward: v[k2] = Z'_p[k1 + R2 k2] in lane j = p*R2 + k1  ->  v[p*R2 + r] = sum_k Z'_p[k] e^{+2 pi i k n / Nz},
// n = lane + 32 r, r < R (outputs n >= Nz/2, outside the window support, are pruned).  No 1/Nz scaling.
template <int N>
__device__ __forceinline__ void fft2_inverse(PC (&v)[32], uint4* __restrict__ xbuf, const float2* __restrict__ tw, int lane) {
  using C = Fft2Cfg<N>;
  {
    PC u[32];
    static_for<0, 32>([&](auto kc) {
      constexpr int k2 = decltype(kc)::value;
      u[brev(k2, 5)] = v[k2];
    });
    dit<32, 0, true, 2>(u);   // u[n1] = sum_k2 Z'[k1 + R2 k2] e^{+2 pi i n1 k2 / 32}
    // transposed store: element (row = n1, col = lane)
    const unsigned wcol = smem_u32(xbuf + lane);
    static_for<0, 32>([&](auto nc) {
      constexpr int n1 = decltype(nc)::value;
      sts_pc<16 * C::xoff(n1)>(wcol, u[n1].re, u[n1].im);
    });
  }
  __syncwarp();
  // row read: lane n1 takes (p, k1) for all columns, conjugate twiddle, placed bit-reversed for the DIT over k1
  const unsigned rrow = smem_u32(xbuf + C::xoff(lane));
  static_for<0, C::kR2>([&](auto kc) {
    constexpr int k1 = decltype(kc)::value;
    if constexpr (k1 == 0) {
      static_for<0, C::kP>([&](auto pc_) {
        constexpr int p = decltype(pc_)::value;
        v[p * C::kR2] = lds_pc<16 * (p * C::kR2)>(rrow);
      });
    } else {
      const float2 w = tw[(k1 - 1) * 32 + lane];
      static_for<0, C::kP>([&](auto pc_) {
        constexpr int p = decltype(pc_)::value;
        const PC y = lds_pc<16 * (p * C::kR2 + k1)>(rrow);
        PC& o = v[p * C::kR2 + brev(k1, C::kLogR2)];
        o.re = fma2s(y.im, w.y, mul2s(y.re, w.x));     // y * conj(w)
        o.im = fma2s(y.re, -w.y, mul2s(y.im, w.x));
      });
    }
  });
  __syncwarp();   // exchange buffer free again
  static_for<0, C::kP>([&](auto pc_) {
    constexpr int p = decltype(pc_)::value;
    dit_pruned_out<C::kR2, p * C::kR2, true>(v);
  });
}

// ---- inverse Hermitian split on packed data ----------------------------------------------------------------------
// P = A'[k], Q = conj(A'[Nz-k]) (the form split2 returns); tw as in split2.  Returns Zk = 2 Z'[k] and Zr = 2 Z'[Nz-k]
// with Z' the spectrum of z'[n] = a'[2n] + i a'[2n+1], a' = N * irfft(A') restricted to the window support:
//   Zk = Fe + conj(g) G,  Zr = conj(Fe - conj(g) G),  Fe = P + Q, G = P - Q, g = -i w_N^k.
// The factor 2 and the 1/Nz of the inverse FFT are folded into the synthesis window table (1/N overall).
template <bool SINFORM>
__device__ __forceinline__ void split2_inv(const PC& P, const PC& Q, float2 tw, PC& Zk, PC& Zr) {
  const pf fer = add2(P.re, Q.re), fei = add2(P.im, Q.im);
  const pf gr = sub2(P.re, Q.re), gi = sub2(P.im, Q.im);
  if constexpr (!SINFORM) {   // conj(g) G = c [(-t gr - gi) + i (gr - t gi)]
    const pf nr = fma2s(gr, tw.x, gi);
    const pf ui = fma2s(gi, -tw.x, gr);
    const pf cu = mul2s(ui, tw.y);
    Zk.re = fma2s(nr, -tw.y, fer); Zk.im = add2(cu, fei);
    Zr.re = fma2s(nr, tw.y, fer);  Zr.im = sub2(cu, fei);
  } else {                    // conj(g) G = s [(-gr - k gi) + i (k gr - gi)]
    const pf nr = fma2s(gi, tw.x, gr);
    const pf nui = fma2s(gr, -tw.x, gi);
    const pf cu = mul2s(nui, -tw.y);
    Zk.re = fma2s(nr, -tw.y, fer); Zk.im = add2(cu, fei);
    Zr.re = fma2s(nr, tw.y, fer);  Zr.im = sub2(cu, fei);
  }
}

}  // namespace sb200
