// Fused STFT-magnitude + mel feature kernel on the packed FFT engine (fft2.cuh):
//   [pre-emphasis ->] reflect pad -> frame gather -> window -> real FFT -> |.|^2 -> banded mel -> dB-normalise / ln / raw
//   transtacos/audio.py:73-77 get_specs;  retunegan/audio.py:116-128 get_mag / get_mel.
// One warp owns one item = 4096/n_fft consecutive frames of one utterance, processed as frame PAIRS packed
// in the halves of 64-bit registers (2 frames for n_fft 2048).  A CTA is 8 warps sharing the plan tables in
// shared memory, each warp with a private 16.5 KB exchange buffer; the grid is persistent (1 CTA per SM).
// Log-scaled magnitudes are taken from |A|^2 directly (a/2 log2 max(floor^2, p) + b): no sqrt on that path.
#pragma once
#include "feat.cuh"
#include "fft2.cuh"

namespace sb200 {

#ifdef kFeat2WarpsOverride
constexpr int kFeat2Warps = kFeat2WarpsOverride;
#else
constexpr int kFeat2Warps = 8;
#endif
// true: the next item's samples are requested before the mel phase of the current one and held in registers
// (needs ~60 more registers: 8 warps per SM); false: fetched at the start of the item (fits 12 warps per SM)
#ifdef kFeat2PrefetchOverride
constexpr bool kFeat2Prefetch = kFeat2PrefetchOverride;
#else
constexpr bool kFeat2Prefetch = true;
#endif

template <int N>
struct Smem2 {
  using C = Fft2Cfg<N>;
  uint4* xbufs;   // [warps][kXElems] exchange buffers (first: 16-byte aligned)
  float* win;     // [win] 0.5 * analysis window
  float2* tw;     // [kTwCount]
  float2* sp2;    // [17*32]
  float* melw;    // [melw_count]
  int* mel_lo;    // [32*rounds]
  __host__ __device__ static size_t bytes(int melw_count, int mel_rounds, int warps) {
    return static_cast<size_t>(warps) * C::kXBytes + sizeof(float) * C::kWin + sizeof(float2) * (C::kTwCount + 17 * 32) +
           sizeof(float) * melw_count + sizeof(int) * 32 * mel_rounds;
  }
  __device__ __forceinline__ void carve(unsigned char* raw, const PlanDev& p, int warps) {
    xbufs = reinterpret_cast<uint4*>(raw);
    win = reinterpret_cast<float*>(raw + static_cast<size_t>(warps) * C::kXBytes);
    tw = reinterpret_cast<float2*>(win + C::kWin);
    sp2 = tw + C::kTwCount;
    melw = reinterpret_cast<float*>(sp2 + 17 * 32);
    mel_lo = reinterpret_cast<int*>(melw + p.melw_count);
  }
  // 16-byte copies, several in flight per thread (all tables are multiples of 16 bytes and 16-byte aligned)
  template <class T>
  static __device__ __forceinline__ void copy16(T* dst, const T* src, int count, float scale = 1.f) {
    const int n16 = count * static_cast<int>(sizeof(T)) / 16;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll 4
    for (int i = threadIdx.x; i < n16; i += kFeat2Warps * 32) {
      float4 t = __ldg(s4 + i);
      if (scale != 1.f) t = make_float4(t.x * scale, t.y * scale, t.z * scale, t.w * scale);
      d4[i] = t;
    }
  }
  __device__ __forceinline__ void fill(const PlanDev& p, bool with_mel) {
    copy16(win, p.window, C::kWin, 0.5f);
    copy16(tw, p.tw, C::kTwCount);
    copy16(sp2, p.sp2, 17 * 32);
    if (with_mel) {
      copy16(melw, p.melw, p.melw_count);
      copy16(mel_lo, p.mel_lo, 32 * p.mel_rounds);
    }
  }
};

// One frame's windowed sample pairs z[lane + 32 r] = (re[r], im[r]), r < R: reflect padding (np.pad mode='reflect') and
// the optional pre-emphasis FIR are applied on the fly.  stage: >= win floats of this warp's shared memory (edge frames).  s_win holds 0.5 * window (the 1/2 of the Hermitian split).
template <int N, bool PRE>
__device__ __forceinline__ void load_frame2(float (&re)[Fft2Cfg<N>::kR], float (&im)[Fft2Cfg<N>::kR],
                                            const float* __restrict__ x, long long L, int t, int T, int hop, float pre,
                                            const float* __restrict__ s_win, float* stage, int lane) {
  using C = Fft2Cfg<N>;
  const long long p0 = static_cast<long long>(t) * hop - N / 4;
  if (t < T && p0 >= 1 && p0 + C::kWin <= L) {
    const float* xp = x + p0 + 2 * lane;
    static_for<0, C::kR>([&](auto rc) {
      constexpr int r = decltype(rc)::value;
      const float lo = __ldg(xp + 64 * r), hi = __ldg(xp + 64 * r + 1);
      float a0 = lo, a1 = hi;
      if constexpr (PRE) {
        a0 = fmaf(-pre, __ldg(xp + 64 * r - 1), lo);
        a1 = fmaf(-pre, lo, hi);
      }
      const float2 w = *reinterpret_cast<const float2*>(s_win + 2 * lane + 64 * r);
      re[r] = a0 * w.x;
      im[r] = a1 * w.y;
    });
  } else if (t < T) {
    // edge frame: a rolled loop stages the reflected (pre-emphasised) samples in shared memory (keeps the code small)
#pragma unroll 1
    for (int m = lane; m < C::kWin; m += 32) {
      long long i = p0 + m;
      if (i < 0) i = -i;
      if (i >= L) i = 2 * (L - 1) - i;
      float s = __ldg(x + i);
      if constexpr (PRE) s = fmaf(-pre, i > 0 ? __ldg(x + i - 1) : 0.f, s);
      stage[m] = s;
    }
    __syncwarp();
    static_for<0, C::kR>([&](auto rc) {
      constexpr int r = decltype(rc)::value;
      const int m = 2 * lane + 64 * r;
      const float2 w = *reinterpret_cast<const float2*>(s_win + m);
      const float2 sv = *reinterpret_cast<const float2*>(stage + m);
      re[r] = sv.x * w.x;
      im[r] = sv.y * w.y;
    });
    __syncwarp();
  } else {
    static_for<0, C::kR>([&](auto rc) {
      constexpr int r = decltype(rc)::value;
      re[r] = 0.f;
      im[r] = 0.f;
    });
  }
}

// Raw samples of one frame pair, loaded ahead of use (software prefetch across the mel phase of the previous item).
// HS > 0 (requires hop == 64*HS, n_fft 2048): the two frames of a pair overlap by R - HS lane slots, so an interior
// pair needs R + HS slots of (previous, even, odd) samples once instead of 2R.
template <int N, bool PRE, int HS>
struct Fetch2 {
  using C = Fft2Cfg<N>;
  static constexpr int kSlots = HS > 0 ? C::kR + HS : 1;
  Item it;
  bool shared;
  float lo[kSlots], hi[kSlots], pv[PRE ? kSlots : 1];
  __device__ __forceinline__ void issue(const BatchDev& bd, long long item, const float* __restrict__ x, int hop, int lane) {
    it = decode_item(bd, item, C::kFrames);
    shared = false;
    if constexpr (HS > 0) {
      const long long p0 = static_cast<long long>(it.t0) * hop - N / 4;
      shared = (it.t0 + 1 < it.T) && p0 >= 1 && p0 + hop + C::kWin <= it.L;
      if (shared) {
        const float* xp = x + it.sig_base + p0 + 2 * lane;
        static_for<0, kSlots>([&](auto rc) {
          constexpr int r = decltype(rc)::value;
          lo[r] = __ldg(xp + 64 * r);
          hi[r] = __ldg(xp + 64 * r + 1);
          if constexpr (PRE) pv[r] = __ldg(xp + 64 * r - 1);
        });
      }
    }
  }
};

// Build the pass-A registers of one item (see fft2_forward).  Frame t = t0 + 2p + h sits in half h of pair p.
template <int N, bool PRE, int HS>
__device__ __forceinline__ void load_item2(PC (&v)[32], const Fetch2<N, PRE, HS>& f, const float* __restrict__ x, int hop,
                                           float pre, const float* __restrict__ s_win, float* stage, int lane) {
  using C = Fft2Cfg<N>;
  static_for<0, C::kP>([&](auto pc_) {
    constexpr int p = decltype(pc_)::value;
    float re[2][C::kR], im[2][C::kR];
    if (HS > 0 && f.shared) {
      if constexpr (HS > 0) {
        static_for<0, C::kR>([&](auto rc) {
          constexpr int r = decltype(rc)::value;
          const float2 w = *reinterpret_cast<const float2*>(s_win + 2 * lane + 64 * r);
          if constexpr (PRE) {
            re[0][r] = fmaf(-pre, f.pv[r], f.lo[r]) * w.x;
            im[0][r] = fmaf(-pre, f.lo[r], f.hi[r]) * w.y;
            re[1][r] = fmaf(-pre, f.pv[r + HS], f.lo[r + HS]) * w.x;
            im[1][r] = fmaf(-pre, f.lo[r + HS], f.hi[r + HS]) * w.y;
          } else {
            re[0][r] = f.lo[r] * w.x;
            im[0][r] = f.hi[r] * w.y;
            re[1][r] = f.lo[r + HS] * w.x;
            im[1][r] = f.hi[r + HS] * w.y;
          }
        });
      }
    } else {
      const int t = f.it.t0 + 2 * p;
      load_frame2<N, PRE>(re[0], im[0], x + f.it.sig_base, f.it.L, t, f.it.T, hop, pre, s_win, stage, lane);
      load_frame2<N, PRE>(re[1], im[1], x + f.it.sig_base, f.it.L, t + 1, f.it.T, hop, pre, s_win, stage, lane);
    }
    static_for<0, C::kR>([&](auto rc) {
      constexpr int r = decltype(rc)::value;
      constexpr int idx = p * C::kR2 + brev(r, C::kLogR2);
      v[idx].re = pk(re[0][r], re[1][r]);
      v[idx].im = pk(im[0][r], im[1][r]);
      v[idx + 1] = v[idx];
    });
  });
}

__device__ __forceinline__ pf sqrt2(pf p) { return pk(fast_sqrt(plo(p)), fast_sqrt(phi(p))); }

template <int N, bool PRE, bool LOGMAG, int HS>
__global__ void __launch_bounds__(kFeat2Warps * 32, 1) stft_feature2_kernel(const PlanDev p, const FeatArgs a) {
  using C = Fft2Cfg<N>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem2<N> sm;
  sm.carve(smem_raw, p, kFeat2Warps);
  sm.fill(p, a.mel != nullptr);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint4* xbuf = sm.xbufs + warp * C::kXElems;
  pf* sbuf = reinterpret_cast<pf*>(xbuf);   // [P][Nz] magnitudes of both frames of a pair (aliases the exchange buffer)
  const int k1 = lane & (C::kR2 - 1), pl = lane / C::kR2;   // pass-B role of this lane: column k1 of pair pl
  const bool col0 = (k1 == 0);
  const int partner = (lane & ~(C::kR2 - 1)) | ((C::kR2 - k1) & (C::kR2 - 1));
  const bool want_mag = a.mag != nullptr, want_mel = a.mel != nullptr;
  // squared-magnitude form of the log scale: a log2 max(floor, sqrt p) + b = a/2 log2 max(floor^2, p) + b
  const float mag_a = 0.5f * a.mag_scale.a, mag_b = a.mag_scale.b, mag_fl = a.mag_scale.floor * a.mag_scale.floor;
  pf* const sa = sbuf + pl * C::kNz + k1;              // bins k1 + R2 s
  pf* const sb = sbuf + pl * C::kNz + C::kNz - k1;     // bins Nz - k1 - R2 s
  const float2* const sp = sm.sp2 + lane;
  // slots whose bins the mel filterbank reads: bit s of need_a for bins [R2 s, R2 (s+1)), of need_b for (Nz - R2 (s+1), Nz - R2 s]
  unsigned need_a = 0, need_b = 0;
  if (want_mel) {
    for (int s = 0; s < 16; ++s) {
      if (C::kR2 * s <= p.mel_kmax && C::kR2 * (s + 1) > p.mel_kmin) need_a |= 1u << s;
      if (C::kNz - C::kR2 * (s + 1) < p.mel_kmax && C::kNz - C::kR2 * s >= p.mel_kmin) need_b |= 1u << s;
    }
  }
  const long long warps_total = static_cast<long long>(gridDim.x) * kFeat2Warps;
  long long item = static_cast<long long>(blockIdx.x) * kFeat2Warps + warp;
  Fetch2<N, PRE, HS> nx;
  bool have = item < a.bd.total_items;
  if (kFeat2Prefetch && have) nx.issue(a.bd, item, a.x, p.hop, lane);
  while (have) {
    if (!kFeat2Prefetch) nx.issue(a.bd, item, a.x, p.hop, lane);
    const Item it = nx.it;
    PC v[32];
    load_item2<N, PRE, HS>(v, nx, a.x, p.hop, a.pre, sm.win, reinterpret_cast<float*>(xbuf), lane);
    fft2_forward<N>(v, xbuf, sm.tw, lane);
    // lane (pl, k1) now holds Z[k1 + R2*k2] of frames fA = t0 + 2 pl, fB = fA + 1
    const int fA = it.t0 + 2 * pl;
    const bool stA = want_mag && fA < it.T, stB = want_mag && fA + 1 < it.T;
    float* const pa = a.mag + (it.frame_base + fA) * C::kF + k1;             // bins k1 + R2 s       (frame B: + F)
    float* const pb = a.mag + (it.frame_base + fA) * C::kF + C::kNz - k1;    // bins Nz - k1 - R2 s
    auto scaled = [&](pf pw) -> pf {   // |A|^2 of both frames -> output values
      if constexpr (LOGMAG) {
        return fma2s(pk(fast_lg2(fmaxf(mag_fl, plo(pw))), fast_lg2(fmaxf(mag_fl, phi(pw)))), mag_a, pk(mag_b, mag_b));
      } else {
        return sqrt2(pw);
      }
    };
    {
      // self pair of column 0 (bin Nz/2) first: the exchange below overwrites v[16]
      PC ak, am;
      split2<true>(v[16], v[16], sp[16 * 32], ak, am);
      const pf pa2 = norm2(ak);
      const pf oa = scaled(pa2);
      if (stA && col0) pa[C::kR2 * 16] = plo(oa);
      if (stB && col0) pa[C::kF + C::kR2 * 16] = phi(oa);
      if (want_mel && col0) sa[C::kR2 * 16] = LOGMAG ? sqrt2(pa2) : oa;
    }
    // exchange with the partner lane, all slots back to back and in place: slot s receives Z[Nz - k] into v[31 - s]
    // (descending s: column-0 lanes send v[32 - s], which slot s - 1 overwrites afterwards)
    static_for<0, 16>([&](auto sc) {
      constexpr int s = 15 - decltype(sc)::value;
      const PC send = pc_sel(col0, v[(32 - s) & 31], v[31 - s]);
      v[31 - s] = pc_shfl(send, partner);
    });
    static_for<0, 16>([&](auto sc) {
      constexpr int s = decltype(sc)::value;
      PC ak, am;
      split2<(s >= 8)>(v[s], v[31 - s], sp[s * 32], ak, am);
      const pf pa2 = norm2(ak), pm2 = norm2(am);
      const pf oa = scaled(pa2), om = scaled(pm2);
      if (stA) {
        pa[C::kR2 * s] = plo(oa);
        pb[-C::kR2 * s] = plo(om);
      }
      if (stB) {
        pa[C::kF + C::kR2 * s] = phi(oa);
        pb[C::kF - C::kR2 * s] = phi(om);
      }
      // magnitudes for the mel filterbank (bin Nz is never part of a filter)
      if (need_a & (1u << s)) sa[C::kR2 * s] = LOGMAG ? sqrt2(pa2) : oa;
      if ((need_b & (1u << s)) && (s > 0 || !col0)) sb[-C::kR2 * s] = LOGMAG ? sqrt2(pm2) : om;
    });
    // next item's samples are requested now and land during the mel phase
    item += warps_total;
    have = item < a.bd.total_items;
    if (kFeat2Prefetch && have) nx.issue(a.bd, item, a.x, p.hop, lane);
    __syncwarp();
    if (want_mel) {
#pragma unroll
      for (int rd = 0; rd < kMaxMelRounds; ++rd) {
        if (rd < p.mel_rounds) {
          const int slot = sm.mel_lo[rd * 32 + lane];
          const int m = slot >> 16, lo = slot & 0xffff;
          const float* wr = sm.melw + p.mel_round_off[rd] + lane;
          const pf* sr = sbuf + lo;   // reads may run past the row end (zero weights) into finite stale data
          const int n = p.mel_round_len[rd];   // multiple of 8
          pf acc[C::kP], acc2[C::kP];
#pragma unroll
          for (int q = 0; q < C::kP; ++q) acc[q] = acc2[q] = 0ull;
#pragma unroll 1
          for (int i0 = 0; i0 < n; i0 += 8) {
            float w[8];
            pf sv[C::kP][8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = wr[(i0 + j) * 32];
#pragma unroll
            for (int q = 0; q < C::kP; ++q)
#pragma unroll
              for (int j = 0; j < 8; ++j) sv[q][j] = sr[q * C::kNz + i0 + j];
#pragma unroll
            for (int q = 0; q < C::kP; ++q)
#pragma unroll
              for (int j = 0; j < 8; j += 2) {
                acc[q] = fma2s(sv[q][j], w[j], acc[q]);
                acc2[q] = fma2s(sv[q][j + 1], w[j + 1], acc2[q]);
              }
          }
#pragma unroll
          for (int q = 0; q < C::kP; ++q) acc[q] = add2(acc[q], acc2[q]);
          if (m < p.n_mel) {
#pragma unroll
            for (int q = 0; q < C::kP; ++q) {
              const int f = it.t0 + 2 * q;
              float* dst = a.mel + (it.frame_base + f) * p.n_mel + m;
              if (f < it.T) dst[0] = apply_scale(a.mel_scale, plo(acc[q]));
              if (f + 1 < it.T) dst[p.n_mel] = apply_scale(a.mel_scale, phi(acc[q]));
            }
          }
        }
      }
    }
    __syncwarp();
  }
}

template <int N>
inline size_t feat2_smem_bytes(const PlanDev& p) {
  return Smem2<N>::bytes(p.melw_count, p.mel_rounds, kFeat2Warps);
}

}  // namespace sb200
