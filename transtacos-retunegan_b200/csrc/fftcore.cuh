// Warp-level real FFT engine for centred, zero-padded STFT frames (sm_100a).
//
// The reference's STFT (librosa.stft / torch.stft as called at transtacos/audio.py:143-144,
// retunegan/audio.py:117,161-163) uses win_length = n_fft/2: the analysis window is n_fft/2
// non-zero taps centred in an n_fft frame.  Writing a[m] = w[m]*x[t*hop - N/4 + m], m < N/2,
//   X[k] = (-i)^k * A[k],   A = DFT_N(a zero-padded to N)
// and A comes from ONE complex FFT of length Nz = N/2 whose upper input half is zero:
//   z[n] = a[2n] + i a[2n+1] (n < Nz/2),  Z = DFT_Nz(z),
//   A[k]    = (Z[k] + conj Z[Nz-k])/2 - (i/2) w_N^k (Z[k] - conj Z[Nz-k]),
//   A[Nz-k] = conj( (Z[k] + conj Z[Nz-k])/2 + (i/2) w_N^k (Z[k] - conj Z[Nz-k]) ).
// The Nz-point FFT is Cooley-Tukey with n = lane + 32 r (r < 2R, R = N/128; inputs r >= R are
// zero and pruned) and k = k1 + 2R k2:
//   pass A (per lane, registers): radix-2R over r, then twiddle w_Nz^{lane*k1}
//   transpose through shared memory (32 x 33 float2 per warp, conflict free)
//   pass B (per lane, registers): one radix-32 over the 32 lanes' values of one (frame, k1)
// One warp "pass" always carries 1024 complex points = Q = 2048/N frames, 32 points per lane,
// so n_fft = 2048 / 1024 / 512 (retunegan/hparam.py:72-81) share one code shape.
// The inverse (irfft restricted to the window support) is the exact transpose of this flow.
#pragma once
#include <cuda_runtime.h>
#include <type_traits>

namespace sb200 {

constexpr unsigned kFullMask = 0xffffffffu;

template <int I, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < E) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, E>(static_cast<F&&>(f));
  }
}

__host__ __device__ constexpr int ilog2(int x) { return x <= 1 ? 0 : 1 + ilog2(x >> 1); }
__host__ __device__ constexpr int brev(int x, int bits) {
  int r = 0;
  for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
  return r;
}

// cos(2 pi i / 32), sin(2 pi i / 32) for i in [0, 16]
__host__ __device__ constexpr float cos32_q(int i) {  // i in [0, 8]
  return i == 0 ? 1.0f
       : i == 1 ? 0.98078528040323043f
       : i == 2 ? 0.92387953251128674f
       : i == 3 ? 0.83146961230254524f
       : i == 4 ? 0.70710678118654752f
       : i == 5 ? 0.55557023301960218f
       : i == 6 ? 0.38268343236508978f
       : i == 7 ? 0.19509032201612825f
                : 0.0f;
}
__host__ __device__ constexpr float cos32(int i) { return i <= 8 ? cos32_q(i) : -cos32_q(16 - i); }
__host__ __device__ constexpr float sin32(int i) { return i <= 8 ? cos32_q(8 - i) : cos32_q(i - 8); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(-a.y, b.y, a.x * b.x), fmaf(a.y, b.x, a.x * b.y));
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {  // a * conj(b)
  return make_float2(fmaf(a.y, b.y, a.x * b.x), fmaf(a.y, b.x, -a.x * b.y));
}

// d * w_32^I (forward, w = e^{-2 pi i/32}) or d * conj(w_32^I) (INV); I in [0, 16)
template <int I, bool INV>
__device__ __forceinline__ float2 twmul(float2 d) {
  static_assert(I >= 0 && I < 16, "twiddle index");
  if constexpr (I == 0) {
    return d;
  } else if constexpr (I == 8) {
    return INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
  } else if constexpr (I == 4) {
    constexpr float h = 0.70710678118654752f;
    return INV ? make_float2(h * (d.x - d.y), h * (d.x + d.y)) : make_float2(h * (d.x + d.y), h * (d.y - d.x));
  } else if constexpr (I == 12) {
    constexpr float h = 0.70710678118654752f;
    return INV ? make_float2(-h * (d.x + d.y), h * (d.x - d.y)) : make_float2(h * (d.y - d.x), -h * (d.x + d.y));
  } else {
    constexpr float c = cos32(I), s = sin32(I);
    return INV ? make_float2(fmaf(-d.y, s, d.x * c), fmaf(d.x, s, d.y * c))
               : make_float2(fmaf(d.y, s, d.x * c), fmaf(-d.x, s, d.y * c));
  }
}

// In-register radix-2 DIF DFT of LEN points v[BASE + STRIDE*i].  X[k] lands at
// v[BASE + STRIDE*brev(k, log2 LEN)].  INV uses conjugate twiddles (no scaling).
template <int LEN, int BASE, int STRIDE, bool INV, int NV>
__device__ __forceinline__ void dif(float2 (&v)[NV]) {
  if constexpr (LEN > 1) {
    constexpr int H = LEN / 2;
    static_for<0, H>([&](auto ic) {
      constexpr int i = decltype(ic)::value;
      const float2 a = v[BASE + STRIDE * i], b = v[BASE + STRIDE * (i + H)];
      v[BASE + STRIDE * i] = make_float2(a.x + b.x, a.y + b.y);
      v[BASE + STRIDE * (i + H)] = twmul<i * (32 / LEN), INV>(make_float2(a.x - b.x, a.y - b.y));
    });
    dif<H, BASE, STRIDE, INV>(v);
    dif<H, BASE + STRIDE * H, STRIDE, INV>(v);
  }
}

template <int N>
struct FftCfg {
  static_assert(N == 2048 || N == 1024 || N == 512, "supported n_fft: 512, 1024, 2048 (win = n_fft/2)");
  static constexpr int kN = N;
  static constexpr int kNz = N / 2;         // complex FFT length
  static constexpr int kWin = N / 2;        // window support
  static constexpr int kR = N / 128;        // non-zero pass-A inputs per lane per frame
  static constexpr int kR2 = 2 * kR;        // pass-A radix
  static constexpr int kQ = 2048 / N;       // frames per warp pass
  static constexpr int kF = N / 2 + 1;      // one-sided bins
  static constexpr int kZS = kNz + (kR2 < 16 ? kR2 : 0);   // smem row stride (float2) of the natural-order Z
  static constexpr int kPairIters = kNz / 64;              // pair iterations per frame: k = lane + 32 i < Nz/2
  static constexpr int kBufF2 = 1056;       // per-warp smem buffer, float2 units: max(32*33, Q*ZS)
  static_assert(kQ * kZS <= kBufF2, "buffer");
  // register position of pass-A output Y[k1] of frame q
  __host__ __device__ static constexpr int posA(int q, int k1) {
    return q * kR2 + (k1 & 1) * kR + brev(k1 >> 1, ilog2(kR));
  }
};

// ---- forward: v (pass-A inputs) -> natural-order Z_q[k] at buf[q*ZS + k] --------------------
// On entry v[q*R2 + r] = z_q[lane + 32 r] (r < R) and v[q*R2 + R + r] = z_q[lane + 32 r] * w_{2R}^r
// (use fwd_put() to fill both).  tw: smem table, tw[(k1-1)*32 + lane] = w_Nz^{k1*lane}.
template <int N, int Q, int RR>
__device__ __forceinline__ void fwd_put(float2 (&v)[32], float2 z) {
  using C = FftCfg<N>;
  v[Q * C::kR2 + RR] = z;
  v[Q * C::kR2 + C::kR + RR] = twmul<RR * (32 / C::kR2), false>(z);
}

template <int N>
__device__ __forceinline__ void fft_forward(float2 (&v)[32], float2* __restrict__ buf,
                                            const float2* __restrict__ tw, int lane) {
  using C = FftCfg<N>;
  // pass A: pruned first stage already applied by fwd_put; finish the two radix-R halves
  static_for<0, C::kQ>([&](auto qc) {
    constexpr int q = decltype(qc)::value;
    dif<C::kR, q * C::kR2, 1, false>(v);
    dif<C::kR, q * C::kR2 + C::kR, 1, false>(v);
  });
  // twiddle + transpose write: T[lane*33 + q*R2 + k1]
  float2* trow = buf + lane * 33;
  static_for<0, C::kR2>([&](auto kc) {
    constexpr int k1 = decltype(kc)::value;
    if constexpr (k1 == 0) {
      static_for<0, C::kQ>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        trow[q * C::kR2] = v[C::posA(q, 0)];
      });
    } else {
      const float2 w = tw[(k1 - 1) * 32 + lane];
      static_for<0, C::kQ>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        trow[q * C::kR2 + k1] = cmul(v[C::posA(q, k1)], w);
      });
    }
  });
  __syncwarp();
  // transpose read: lane j = (q, k1) takes the 32 lanes' values
  static_for<0, 32>([&](auto nc) {
    constexpr int n2 = decltype(nc)::value;
    v[n2] = buf[n2 * 33 + lane];
  });
  // pass B
  dif<32, 0, 1, false>(v);
  __syncwarp();   // all transpose reads done before the aliasing natural-order store
  float2* zrow = buf + (lane / C::kR2) * C::kZS + (lane % C::kR2);
  static_for<0, 32>([&](auto kc) {
    constexpr int k2 = decltype(kc)::value;
    zrow[C::kR2 * k2] = v[brev(k2, 5)];
  });
  __syncwarp();
}

// ---- inverse: natural-order Z'_q[k] in buf -> v[q*R + r] = sum_k Z'_q[k] e^{+2 pi i k n / Nz},
// n = lane + 32 r, r < R (outputs n >= Nz/2 are pruned).  No 1/Nz scaling.
template <int N>
__device__ __forceinline__ void fft_inverse(float2 (&v)[32], float2* __restrict__ buf,
                                            const float2* __restrict__ tw, int lane) {
  using C = FftCfg<N>;
  const float2* zrow = buf + (lane / C::kR2) * C::kZS + (lane % C::kR2);
  static_for<0, 32>([&](auto kc) {
    constexpr int k2 = decltype(kc)::value;
    v[k2] = zrow[C::kR2 * k2];
  });
  dif<32, 0, 1, true>(v);
  __syncwarp();
  static_for<0, 32>([&](auto nc) {
    constexpr int n2 = decltype(nc)::value;
    buf[n2 * 33 + lane] = v[brev(n2, 5)];
  });
  __syncwarp();
  const float2* trow = buf + lane * 33;
  static_for<0, 32>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    v[j] = trow[j];
  });
  __syncwarp();   // buffer free for the next pass
  static_for<1, C::kR2>([&](auto kc) {
    constexpr int k1 = decltype(kc)::value;
    const float2 w = tw[(k1 - 1) * 32 + lane];
    static_for<0, C::kQ>([&](auto qc) {
      constexpr int q = decltype(qc)::value;
      v[q * C::kR2 + k1] = cmul_conj(v[q * C::kR2 + k1], w);
    });
  });
  // pruned DIT radix-2R: E over even k1, O over odd k1, keep outputs r < R
  float2 out[32];
  static_for<0, C::kQ>([&](auto qc) {
    constexpr int q = decltype(qc)::value;
    dif<C::kR, q * C::kR2, 2, true>(v);
    dif<C::kR, q * C::kR2 + 1, 2, true>(v);
    static_for<0, C::kR>([&](auto rc) {
      constexpr int r = decltype(rc)::value;
      constexpr int p = q * C::kR2 + 2 * brev(r, ilog2(C::kR));
      const float2 e = v[p];
      const float2 o = twmul<r * (32 / C::kR2), true>(v[p + 1]);
      out[q * C::kR + r] = make_float2(e.x + o.x, e.y + o.y);
    });
  });
  static_for<0, C::kQ * C::kR>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    v[i] = out[i];
  });
}

// ---- Hermitian split (forward): pair (k, Nz-k) of Z -> A[k], A[Nz-k] ------------------------
// wsk = -0.5i * w_N^k.  X[k] = (-i)^k A[k] (apply rot_fwd when the complex value / phase matters).
__device__ __forceinline__ void split_fwd(float2 Zk, float2 Zr, float2 wsk, float2& Ak, float2& Am) {
  const float2 fe = make_float2(Zk.x + Zr.x, Zk.y - Zr.y);   // Zk + conj Zr
  const float2 fo = make_float2(Zk.x - Zr.x, Zk.y + Zr.y);   // Zk - conj Zr
  const float2 t = cmul(wsk, fo);
  Ak = make_float2(fmaf(0.5f, fe.x, t.x), fmaf(0.5f, fe.y, t.y));
  Am = make_float2(fmaf(0.5f, fe.x, -t.x), fmaf(-0.5f, fe.y, t.y));
}

// ---- Hermitian split (inverse): A'[k], A'[Nz-k] -> Z'[k], Z'[Nz-k]  (scaled by 2; caller folds 1/N)
// A' = i^k X (apply rot_inv first).  Z'[k] = Fe + H, Z'[Nz-k] = conj(Fe - H),
// Fe = A'k + conj A'm, H = 2 (A'k - conj A'm) conj(wsk).
__device__ __forceinline__ void split_inv(float2 Ak, float2 Am, float2 wsk, float2& Zk, float2& Zr) {
  const float2 fe = make_float2(Ak.x + Am.x, Ak.y - Am.y);
  const float2 g = make_float2(Ak.x - Am.x, Ak.y + Am.y);
  const float2 h = cmul_conj(g, wsk);
  Zk = make_float2(fmaf(2.0f, h.x, fe.x), fmaf(2.0f, h.y, fe.y));
  Zr = make_float2(fmaf(-2.0f, h.x, fe.x), fmaf(2.0f, h.y, -fe.y));
}

// (-i)^j * a  and  i^j * a
__device__ __forceinline__ float2 rot_fwd(float2 a, int j) {
  j &= 3;
  const float x = (j & 1) ? a.y : a.x, y = (j & 1) ? a.x : a.y;
  return make_float2((j >= 2) ? -x : x, (j == 1 || j == 2) ? -y : y);
}
__device__ __forceinline__ float2 rot_inv(float2 a, int j) { return rot_fwd(a, (4 - (j & 3)) & 3); }

}  // namespace sb200
