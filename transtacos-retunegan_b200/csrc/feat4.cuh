// STFT-magnitude + mel feature kernel, 12 + 4 role split (n_fft 2048; the hot kernel of bench.py's config 3):
//   12 "transform" warps per CTA (3 per scheduler):  [pre-emphasis ->] reflect pad -> frame gather -> window -> packed real FFT ->
//       Hermitian split -> |.|^2 -> dB-normalise / ln / raw magnitudes stored STRAIGHT FROM REGISTERS (128-byte row segments),
//       the magnitudes of the mel band handed over in shared memory;
//    4 "mel" warps per CTA (1 per scheduler), each serving three transform warps: banded mel filterbank -> scale -> stores.
//   transtacos/audio.py:73-77 get_specs;  retunegan/audio.py:116-128 get_mag / get_mel.
//
// EXPERIMENT, NOT THE DEFAULT (SB200_FEAT_KERNEL=4 selects it): measured 96.9 us on config 3 against 86.8 us for the 8 + 8 kernel
// (feat3.cuh) -- results are bit-identical, the parity tests pass on it.
// The idea (profiles/r02_feat_summary.md): the 8 + 8 kernel is bound by the dependent instruction chain of its analysis warps, 2
// per scheduler: with the epilogue warps doing nothing it still needs 62.7 us (87.5 us in all), issue slots, FMA pipe and shared
// memory all about half idle.  More FFT warps per scheduler looked like the lever, and the register file (64 K) is what caps them.
// Moving the log / store of the magnitudes back into the FFT warp costs it ~330 instructions per item but removes the |A|^2 round
// trip through shared memory and leaves the second role so little work (the mel: ~400 instructions per item) that one warp of 56
// registers serves three FFT warps: 12 x 152 + 4 x 56 registers per thread position = the whole file.  The forward transpose goes
// through the exchange buffer one 8-byte plane at a time, which halves the buffer (8.25 KB per warp) so that twelve fit.
// What happened: at 152 registers the transform role spills (216 B), its item grows to ~2300 instructions with MUFU / store phases
// that again run one after the other inside the warp (the lock-step problem of feat2.cuh), and a warp's item takes 24.5 k cycles
// instead of 14.6 k: three slower warps per scheduler lose against two faster ones plus two helpers.
#pragma once
#include "feat3.cuh"

namespace sb200 {

constexpr int kF4Fft = 12;                       // transform warps per CTA
constexpr int kF4Mel = 4;                        // mel warps per CTA
constexpr int kF4PerMel = kF4Fft / kF4Mel;       // transform warps served by one mel warp
constexpr int kF4Threads = (kF4Fft + kF4Mel) * 32;
constexpr int kF4BandElems = 1024 + 16;          // packed magnitudes of a frame pair by bin (only the mel band is written) + ELL over-read pad
#ifndef kF4FftRegs
#define kF4FftRegs 152
#endif
#ifndef kF4MelRegs
#define kF4MelRegs 56
#endif

template <int N>
struct Smem4 {
  using C = Fft2Cfg<N>;
  static constexpr int kXHalfBytes = C::kPlane;   // one 8-byte plane of the transpose
  unsigned char* xbufs;     // [kF4Fft][kXHalfBytes]
  pf* bands;                // [kF4Fft][kF4BandElems]
  float* win;               // [win] 0.5 * analysis window
  float2* tw;               // [kTwCount]
  float2* sp2;              // [17*32]
  float* melw;              // [melw_count]
  int* mel_lo;              // [32*rounds]
  unsigned long long* bar;  // [2*kF4Fft]: full[w], empty[w]
  __host__ __device__ static size_t bytes(int melw_count, int mel_rounds) {
    return static_cast<size_t>(kF4Fft) * (kXHalfBytes + kF4BandElems * sizeof(pf)) + sizeof(float) * C::kWin +
           sizeof(float2) * (C::kTwCount + 17 * 32) + sizeof(float) * melw_count + sizeof(int) * 32 * mel_rounds +
           sizeof(unsigned long long) * 2 * kF4Fft;
  }
  __device__ __forceinline__ void carve(unsigned char* raw, const PlanDev& p) {
    xbufs = raw;
    bands = reinterpret_cast<pf*>(raw + static_cast<size_t>(kF4Fft) * kXHalfBytes);
    win = reinterpret_cast<float*>(bands + kF4Fft * kF4BandElems);
    tw = reinterpret_cast<float2*>(win + C::kWin);
    sp2 = tw + C::kTwCount;
    melw = reinterpret_cast<float*>(sp2 + 17 * 32);
    mel_lo = reinterpret_cast<int*>(melw + p.melw_count);
    bar = reinterpret_cast<unsigned long long*>(mel_lo + 32 * p.mel_rounds);
  }
  template <class T>
  static __device__ __forceinline__ void copy16(T* dst, const T* src, int count, float scale = 1.f) {
    const int n16 = count * static_cast<int>(sizeof(T)) / 16;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll 2
    for (int i = threadIdx.x; i < n16; i += kF4Threads) {
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
    for (int i = threadIdx.x; i < kF4Fft * kF4BandElems; i += kF4Threads) bands[i] = 0ull;
  }
};

// fft2_forward with the transpose done one 8-byte plane at a time through a buffer of HALF the size (n_fft 2048: one frame pair
// per warp).  The twiddled imaginary parts wait in their registers while the real parts cross; results are bit-identical.
template <int N>
__device__ __forceinline__ void fft2_forward_planes(PC (&v)[32], unsigned char* __restrict__ xbuf, const float2* __restrict__ tw,
                                                    int lane) {
  using C = Fft2Cfg<N>;
  static_assert(C::kP == 1, "one frame pair per warp");
  dit<C::kR2, 0, false, 4>(v);
  const unsigned wrow = smem_u32(xbuf) + 8 * C::xoff(lane);
  constexpr int kTwBatch = 8;
  float2 wb[2][kTwBatch];
  static_for<0, kTwBatch>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    if constexpr (j >= 1) wb[0][j] = tw[(j - 1) * 32 + lane];
  });
  static_for<0, C::kR2 / kTwBatch>([&](auto bc) {
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
        sts_pf<0>(wrow, v[0].re);
      } else {
        const float2 w = wb[b & 1][j];
        const PC y = v[k1];
        v[k1].im = fma2s(y.im, w.x, mul2s(y.re, w.y));
        sts_pf<8 * k1>(wrow, fma2s(y.im, -w.y, mul2s(y.re, w.x)));
      }
    });
  });
  __syncwarp();
  const unsigned rcol = smem_u32(xbuf) + 8 * lane;
  static_for<0, 32>([&](auto nc) {
    constexpr int n1 = decltype(nc)::value;
    v[brev(n1, 5)].re = lds_pf<8 * C::xoff(n1)>(rcol);
  });
  __syncwarp();   // every lane has its real parts: the plane can take the imaginary parts
  static_for<0, C::kR2>([&](auto kc) {
    constexpr int k1 = decltype(kc)::value;
    sts_pf<8 * k1>(wrow, v[k1].im);
  });
  __syncwarp();
  static_for<0, 32>([&](auto nc) {
    constexpr int n1 = decltype(nc)::value;
    v[brev(n1, 5)].im = lds_pf<8 * C::xoff(n1)>(rcol);
  });
  dit<32, 0, false, 2>(v);
  __syncwarp();   // exchange buffer free again
}

// ---- transform warp --------------------------------------------------------------------------------------------------------
template <int N, bool PRE, bool LOGMAG, int HS>
__device__ __forceinline__ void feat4_transform(const PlanDev& p, const FeatArgs& a, Smem4<N>& sm, int w, int lane) {
  using C = Fft2Cfg<N>;
  unsigned char* xbuf = sm.xbufs + static_cast<size_t>(w) * Smem4<N>::kXHalfBytes;
  pf* band = sm.bands + w * kF4BandElems;
  const unsigned full = smem_u32(sm.bar + w), empty = smem_u32(sm.bar + kF4Fft + w);
  const int k1 = lane;                                       // pass-B role of this lane: column k1 (R2 = 32)
  const bool col0 = (k1 == 0);
  const int partner = (C::kR2 - k1) & (C::kR2 - 1);
  const bool want_mag = a.mag != nullptr, want_mel = a.mel != nullptr;
  const float mag_a = 0.5f * a.mag_scale.a, mag_b = a.mag_scale.b, mag_fl = a.mag_scale.floor * a.mag_scale.floor;
  // slots whose bins the mel filterbank reads: bit s of need_a for bins [32 s, 32 (s+1)), of need_b for (Nz - 32 (s+1), Nz - 32 s]
  unsigned need_a = 0, need_b = 0;
  if (want_mel) {
    for (int s = 0; s < 17; ++s) {
      if (C::kR2 * s <= p.mel_kmax && C::kR2 * (s + 1) > p.mel_kmin) need_a |= 1u << s;
      if (s < 16 && C::kNz - C::kR2 * (s + 1) < p.mel_kmax && C::kNz - C::kR2 * s >= p.mel_kmin) need_b |= 1u << s;
    }
  }
  const float2* const sp = sm.sp2 + lane;
  auto scaled = [&](pf pw) -> pf {   // |A|^2 of both frames -> output values
    if constexpr (LOGMAG) {
      return fma2s(pk(fast_lg2(fmaxf(mag_fl, plo(pw))), fast_lg2(fmaxf(mag_fl, phi(pw)))), mag_a, pk(mag_b, mag_b));
    } else {
      return sqrt2(pw);
    }
  };
  const long long warps_total = static_cast<long long>(gridDim.x) * kF4Fft;
  long long item = static_cast<long long>(blockIdx.x) * kF4Fft + w;
  unsigned round = 0;
  for (; item < a.bd.total_items; item += warps_total, ++round) {
    const Item it = decode_item(a.bd, item, C::kFrames);
    PC v[32];
    gather_item3<N, PRE, HS>(v, it, a.x, p.hop, a.pre, sm.win, reinterpret_cast<float*>(xbuf), lane);
    fft2_forward_planes<N>(v, xbuf, sm.tw, lane);
    // lane k1 now holds Z[k1 + 32 k2] of frames fA = t0, fB = t0 + 1
    const bool stA = want_mag && it.t0 < it.T, stB = want_mag && it.t0 + 1 < it.T;
    float* const pa = a.mag + (it.frame_base + it.t0) * C::kF + k1;             // bins k1 + 32 s       (frame B: + F)
    float* const pb = a.mag + (it.frame_base + it.t0) * C::kF + C::kNz - k1;    // bins Nz - k1 - 32 s
    if (want_mel) {
      SB200_HANDOFF_MBAR(mbar_wait(empty, (round & 1) ^ 1));   // the mel warp is done with the previous item's band
      pair_bar_sync(w);
    }
    {
      // self pair of column 0 (bin Nz/2) first: the exchange below reuses v[16]
      PC ak, am;
      split2<true>(v[16], v[16], sp[16 * 32], ak, am);
      const pf p2 = norm2(ak);
      const pf o = scaled(p2);
      if (stA && col0) pa[C::kR2 * 16] = plo(o);
      if (stB && col0) pa[C::kF + C::kR2 * 16] = phi(o);
      if ((need_a >> 16 & 1u) && col0) band[C::kR2 * 16] = LOGMAG ? sqrt2(p2) : o;
    }
    // split twiddles in two batches (a table load written after a store cannot be hoisted above it; sixteen at once cost registers)
    static_for<0, 2>([&](auto hc) {
      constexpr int h = decltype(hc)::value;
      float2 spv[8];
      static_for<0, 8>([&](auto sc) {
        constexpr int s = 8 * h + decltype(sc)::value;
        spv[s - 8 * h] = sp[s * 32];
      });
      static_for<0, 8>([&](auto sc) {
        constexpr int s = 8 * h + decltype(sc)::value;
        const PC zr = pc_shfl(pc_sel(col0, v[(32 - s) & 31], v[31 - s]), partner);
        PC ak, am;
        split2<(s >= 8)>(v[s], zr, spv[s - 8 * h], ak, am);
        const pf pa2 = norm2(ak), pm2 = norm2(am);
        const pf oa = scaled(pa2), om = scaled(pm2);
        // bins k1 + 32 s (ak) and Nz - k1 - 32 s (am); for s = 0 column 0 holds bins 0 and Nz
        if (stA) {
          pa[C::kR2 * s] = plo(oa);
          pb[-C::kR2 * s] = plo(om);
        }
        if (stB) {
          pa[C::kF + C::kR2 * s] = phi(oa);
          pb[C::kF - C::kR2 * s] = phi(om);
        }
        // magnitudes for the mel filterbank (bin Nz is never part of a filter: fmax < sr/2)
        if (need_a >> s & 1u) band[k1 + C::kR2 * s] = LOGMAG ? sqrt2(pa2) : oa;
        if ((need_b >> s & 1u) && (s > 0 || !col0)) band[C::kNz - k1 - C::kR2 * s] = LOGMAG ? sqrt2(pm2) : om;
      });
    });
    if (want_mel) {
      SB200_HANDOFF_MBAR(mbar_arrive(full));
      pair_bar_sync(w);
    }
  }
}

// ---- mel warp: serves kF4PerMel transform warps in turn ------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void feat4_mel(const PlanDev& p, const FeatArgs& a, Smem4<N>& sm, int mw, int lane) {
  using C = Fft2Cfg<N>;
  if (a.mel == nullptr) return;
  const long long warps_total = static_cast<long long>(gridDim.x) * kF4Fft;
  const long long base = static_cast<long long>(blockIdx.x) * kF4Fft + mw * kF4PerMel;
  for (unsigned round = 0; base + static_cast<long long>(round) * warps_total < a.bd.total_items; ++round) {
#pragma unroll 1
    for (int f = 0; f < kF4PerMel; ++f) {
      const long long item = base + f + static_cast<long long>(round) * warps_total;
      if (item >= a.bd.total_items) break;
      const int w = mw * kF4PerMel + f;
      const pf* band = sm.bands + w * kF4BandElems;
      const Item it = decode_item(a.bd, item, C::kFrames);
      pair_bar_sync(w);
      SB200_HANDOFF_MBAR(mbar_wait(smem_u32(sm.bar + w), round & 1));
      pair_bar_sync(w);
#pragma unroll
      for (int rd = 0; rd < kMaxMelRounds; ++rd) {
        if (rd < p.mel_rounds) {
          const int slot = sm.mel_lo[rd * 32 + lane];
          const int m = slot >> 16, lo = slot & 0xffff;
          const float* wr = sm.melw + p.mel_round_off[rd] + lane;
          const pf* sr = band + lo;   // reads may run past the row end (zero weights) into finite stale data
          const int n = p.mel_round_len[rd];   // multiple of 8
          pf acc[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j] = 0ull;
#pragma unroll 1
          for (int i0 = 0; i0 < n; i0 += 8) {
            float wv[8];
            pf sv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) wv[j] = wr[(i0 + j) * 32];
#pragma unroll
            for (int j = 0; j < 8; ++j) sv[j] = sr[i0 + j];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j & 3] = fma2s(sv[j], wv[j], acc[j & 3]);
          }
          if (m < p.n_mel) {
            const pf s = add2(add2(acc[0], acc[1]), add2(acc[2], acc[3]));
            float* dst = a.mel + (it.frame_base + it.t0) * p.n_mel + m;
            if (it.t0 < it.T) dst[0] = apply_scale(a.mel_scale, plo(s));
            if (it.t0 + 1 < it.T) dst[p.n_mel] = apply_scale(a.mel_scale, phi(s));
          }
        }
      }
      SB200_HANDOFF_MBAR(mbar_arrive(smem_u32(sm.bar + kF4Fft + w)));
    }
  }
}

template <int N, bool PRE, bool LOGMAG, int HS>
__global__ void __launch_bounds__(kF4Threads, 1) stft_feature4_kernel(const PlanDev p, const FeatArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem4<N> sm;
  sm.carve(smem_raw, p);
  sm.fill(p, a.mel != nullptr);
  if (threadIdx.x < 2 * kF4Fft) {   // full / empty: 32 lane arrivals each
    mbar_init(smem_u32(sm.bar + threadIdx.x), 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp < kF4Fft) {
#ifndef SB200_NO_SETMAXNREG
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kF4FftRegs));
#endif
    feat4_transform<N, PRE, LOGMAG, HS>(p, a, sm, warp, lane);
  } else {
#ifndef SB200_NO_SETMAXNREG
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kF4MelRegs));
#endif
    feat4_mel<N>(p, a, sm, warp - kF4Fft, lane);
  }
}

template <int N>
inline size_t feat4_smem_bytes(const PlanDev& p) {
  return Smem4<N>::bytes(p.melw_count, p.mel_rounds);
}

}  // namespace sb200
