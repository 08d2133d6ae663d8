// Fused STFT-magnitude + mel feature kernel:  [pre-emphasis ->] reflect pad -> frame gather ->
// window -> real FFT -> |.| -> banded mel -> dB-normalise / ln / raw.
//   transtacos/audio.py:73-77 get_specs;  retunegan/audio.py:116-128 get_mag / get_mel;
//   retunegan/audio.py:161-166 torch.stft + abs.
// One warp owns one "pass" (Q = 2048/n_fft consecutive frames of one utterance); a CTA is 8 such warps
// sharing the plan tables in shared memory; the grid is persistent (2 CTAs per SM).
#pragma once
#include "fftcore.cuh"
#include "plan.cuh"

namespace sb200 {

struct BatchDev {
  int B;
  long long len, stride;
  const long long* sig_off;
  const long long* sig_len;
  const long long* frame_off;
  const long long* item_off;
  long long total_items;
  long long total_samples;   // ragged: samples in the flat buffer (0 = unknown)
  long long items_per_row;   // uniform
  long long frames_per_row;  // uniform
};

struct Item {
  int b;
  long long sig_base;    // first sample of the utterance
  long long L;           // samples
  long long frame_base;  // first output frame of the utterance
  int T;                 // frames of the utterance
  int t0;                // first frame of this pass
};

__device__ __forceinline__ Item decode_item(const BatchDev& bd, long long item, int Q) {
  Item it;
  if (bd.item_off == nullptr) {
    const long long b = item / bd.items_per_row;
    it.b = static_cast<int>(b);
    it.sig_base = b * bd.stride;
    it.L = bd.len;
    it.frame_base = b * bd.frames_per_row;
    it.T = static_cast<int>(bd.frames_per_row);
    it.t0 = static_cast<int>(item - b * bd.items_per_row) * Q;
  } else {
    int lo = 0, hi = bd.B;   // largest b with item_off[b] <= item
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(bd.item_off + mid) <= item) lo = mid; else hi = mid;
    }
    it.b = lo;
    it.sig_base = __ldg(bd.sig_off + lo);
    it.L = __ldg(bd.sig_len + lo);
    it.frame_base = __ldg(bd.frame_off + lo);
    it.T = static_cast<int>(__ldg(bd.frame_off + lo + 1) - it.frame_base);
    it.t0 = static_cast<int>(item - __ldg(bd.item_off + lo)) * Q;
  }
  return it;
}

struct ScaleDev {
  int log;
  float a, b, floor;
};

__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_lg2(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float apply_scale(const ScaleDev& s, float v) {
  return s.log ? fmaf(s.a, fast_lg2(fmaxf(s.floor, v)), s.b) : v;
}

// Table fill: up to NS segments of 16-byte chunks copied global -> shared as ONE index space, four chunks per thread in flight
// (all loads of a batch before its stores).  A CTA of the mstft kernels lives for one pass per warp, so the latency of its table
// fill is on every warp's critical path (profiles: 10-13 % of the warp time with one scalar loop per table).
struct CopySeg {
  const uint4* src;
  uint4* dst;
  int n16;
};
template <class T>
__device__ __forceinline__ CopySeg copy_seg(T* dst, const T* src, int count) {
  return CopySeg{reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), static_cast<int>(count * sizeof(T) / 16)};
}
template <int NS>
__device__ __forceinline__ void copy_segments(const CopySeg (&seg)[NS], int ns) {
  int total = 0;
#pragma unroll
  for (int s = 0; s < NS; ++s) total += s < ns ? seg[s].n16 : 0;
  constexpr int kBatch = 4;
  for (int i0 = threadIdx.x; i0 < total; i0 += kBatch * blockDim.x) {
    uint4 r[kBatch];
    uint4* d[kBatch];
#pragma unroll
    for (int j = 0; j < kBatch; ++j) {
      int i = i0 + j * blockDim.x;
      d[j] = nullptr;
      if (i < total) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          if (s < ns && d[j] == nullptr) {
            if (i < seg[s].n16) {
              r[j] = __ldg(seg[s].src + i);
              d[j] = seg[s].dst + i;
            } else {
              i -= seg[s].n16;
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kBatch; ++j)
      if (d[j] != nullptr) *d[j] = r[j];
  }
}

// The same fill without blocking: cp.async (LDGSTS) copies the chunks of segments [s0, s1) straight into shared memory and the
// thread's copies arrive on an mbarrier (initialised with one arrival per thread of the CTA) when they have landed.  A warp waits on
// the barrier of a table group where it first needs one of its tables (tbl_wait), so the fill runs under whatever comes before:
// the window / twiddle group under the item decode, everything else under the frame gather and the first FFT.
__device__ __forceinline__ unsigned tbl_smem_addr(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void tbl_bar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tbl_smem_addr(bar)), "r"(count) : "memory");
}
template <int NS>
__device__ __forceinline__ void copy_segments_async(const CopySeg (&seg)[NS], int s0, int s1, unsigned long long* bar) {
  int total = 0;
#pragma unroll
  for (int s = 0; s < NS; ++s) total += (s >= s0 && s < s1) ? seg[s].n16 : 0;
  for (int i0 = threadIdx.x; i0 < total; i0 += blockDim.x) {
    int i = i0;
    bool done = false;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (s >= s0 && s < s1 && !done) {
        if (i < seg[s].n16) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tbl_smem_addr(seg[s].dst + i)), "l"(seg[s].src + i) : "memory");
          done = true;
        } else {
          i -= seg[s].n16;
        }
      }
    }
  }
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tbl_smem_addr(bar)) : "memory");
}
// wait for phase 0 of the barrier (returns at once when it has completed: the barrier is used for one fill per CTA)
__device__ __forceinline__ void tbl_wait(unsigned long long* bar) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "TBL_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
      "@p bra TBL_DONE;\n\t"
      "bra TBL_WAIT;\n\t"
      "TBL_DONE:\n\t}" ::"r"(tbl_smem_addr(bar))
      : "memory");
}

// Shared-memory table block common to all FFT kernels.
template <int N>
struct SmemTables {
  using C = FftCfg<N>;
  static constexpr int kTwCount = (C::kR2 - 1) * 32;
  static constexpr int kWsCount = C::kNz / 2 + 2;
  float* win;     // [win]
  float2* tw;     // [kTwCount]
  float2* ws;     // [kWsCount]
  float* melw;    // [melw_count]
  int* mel_lo;    // [32*rounds]
  float2* bufs;   // per-warp buffers
  __host__ __device__ static size_t table_bytes(int melw_count, int mel_rounds) {
    return sizeof(float) * C::kWin + sizeof(float2) * (kTwCount + kWsCount) + sizeof(float) * melw_count +
           sizeof(int) * 32 * mel_rounds;
  }
  __device__ __forceinline__ void carve(unsigned char* raw, const PlanDev& p) {
    win = reinterpret_cast<float*>(raw);
    tw = reinterpret_cast<float2*>(win + C::kWin);
    ws = tw + kTwCount;
    melw = reinterpret_cast<float*>(ws + kWsCount);
    mel_lo = reinterpret_cast<int*>(melw + p.melw_count);
    bufs = reinterpret_cast<float2*>(mel_lo + 32 * p.mel_rounds);
  }
  // the same fill as copy segments (every table is a whole number of 16-byte chunks); returns the number of segments
  __device__ __forceinline__ int segments(const PlanDev& p, const float* win_src, bool with_mel, CopySeg* out) {
    out[0] = copy_seg(win, win_src, C::kWin);
    out[1] = copy_seg(tw, p.tw, kTwCount);
    out[2] = copy_seg(ws, p.ws, kWsCount);
    if (!with_mel) return 3;
    out[3] = copy_seg(melw, p.melw, p.melw_count);
    out[4] = copy_seg(mel_lo, p.mel_lo, 32 * p.mel_rounds);
    return 5;
  }
  // which: analysis window (p.window) or any other [win] table
  __device__ __forceinline__ void fill(const PlanDev& p, const float* win_src, bool with_mel) {
    for (int i = threadIdx.x; i < C::kWin; i += blockDim.x) win[i] = win_src[i];
    for (int i = threadIdx.x; i < kTwCount; i += blockDim.x) tw[i] = p.tw[i];
    for (int i = threadIdx.x; i < kWsCount; i += blockDim.x) ws[i] = p.ws[i];
    if (with_mel) {
      for (int i = threadIdx.x; i < p.melw_count; i += blockDim.x) melw[i] = p.melw[i];
      for (int i = threadIdx.x; i < 32 * p.mel_rounds; i += blockDim.x) mel_lo[i] = p.mel_lo[i];
    }
  }
};

// Gather + window Q frames into pass-A registers.  x points at the utterance; reflect padding
// (np.pad mode='reflect') and the optional pre-emphasis FIR are applied on the fly.
template <int N, bool PRE>
__device__ __forceinline__ void load_frames(float2 (&v)[32], const float* __restrict__ x, long long L, int t0, int T,
                                            int hop, float pre, const float* __restrict__ s_win, int lane) {
  using C = FftCfg<N>;
  static_for<0, C::kQ>([&](auto qc) {
    constexpr int q = decltype(qc)::value;
    const int t = t0 + q;
    const long long p0 = static_cast<long long>(t) * hop - N / 4;
    if (t < T && p0 >= 1 && p0 + C::kWin <= L) {
      const float* xp = x + p0 + 2 * lane;
      static_for<0, C::kR>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        const float lo = __ldg(xp + 64 * r), hi = __ldg(xp + 64 * r + 1);
        float a0 = lo, a1 = hi;
        if constexpr (PRE) {
          float prev = __shfl_up_sync(kFullMask, hi, 1);
          if (lane == 0) prev = __ldg(xp + 64 * r - 1);
          a0 = fmaf(-pre, prev, lo);
          a1 = fmaf(-pre, lo, hi);
        }
        const float2 w = *reinterpret_cast<const float2*>(s_win + 2 * lane + 64 * r);
        fwd_put<N, q, r>(v, make_float2(a0 * w.x, a1 * w.y));
      });
    } else if (t < T) {
      auto sample = [&](long long i) -> float {
        if (i < 0) i = -i;
        if (i >= L) i = 2 * (L - 1) - i;
        float s = __ldg(x + i);
        if constexpr (PRE) s = fmaf(-pre, i > 0 ? __ldg(x + i - 1) : 0.f, s);
        return s;
      };
      static_for<0, C::kR>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        const int m = 2 * lane + 64 * r;
        const float2 w = *reinterpret_cast<const float2*>(s_win + m);
        fwd_put<N, q, r>(v, make_float2(sample(p0 + m) * w.x, sample(p0 + m + 1) * w.y));
      });
    } else {
      static_for<0, C::kR>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        fwd_put<N, q, r>(v, make_float2(0.f, 0.f));
      });
    }
  });
}

// Banded mel projection of the raw magnitudes held in buf[q*ZS + k].x; calls emit(q, rd, m, value) with
// m = the mel row of slot (rd, lane), >= n_mel for an empty slot (q and rd come from fully unrolled loops, so register arrays indexed by them stay in registers).
template <int N, class Emit>
__device__ __forceinline__ void mel_project_smem(const PlanDev& p, const float* __restrict__ s_melw,
                                                 const int* __restrict__ s_lo, const float2* __restrict__ buf, int lane,
                                                 Emit&& emit) {
  using C = FftCfg<N>;
#pragma unroll
  for (int rd = 0; rd < kMaxMelRounds; ++rd) {
    if (rd < p.mel_rounds) {
      const int slot = s_lo[rd * 32 + lane];
      const int m = slot >> 16, lo = slot & 0xffff;   // mel row of this lane in this round (>= n_mel: none)
      const float* wr = s_melw + p.mel_round_off[rd] + lane;
      const int n = p.mel_round_len[rd];
      float acc[C::kQ];
#pragma unroll
      for (int q = 0; q < C::kQ; ++q) acc[q] = 0.f;
      for (int it = 0; it < n; ++it) {
        const float w = wr[it * 32];
        const int idx = min(lo + it, C::kNz - 1);
#pragma unroll
        for (int q = 0; q < C::kQ; ++q) acc[q] = fmaf(w, buf[q * C::kZS + idx].x, acc[q]);
      }
#pragma unroll
      for (int q = 0; q < C::kQ; ++q) emit(q, rd, m, acc[q]);
    }
  }
}

// Sum of the frames covering offset coordinate pp (= padded position - n_fft/4) of one utterance whose
// frames are stored as fb[frame][win] (overlap-add written as a gather: deterministic, no atomics).
// L2 = true: the frames were written earlier in the SAME launch by other CTAs (behind a grid barrier): read them through L2
// (ld.global.cg), never through the non-coherent L1 / read-only path.
template <bool L2 = false>
__device__ __forceinline__ float ola_gather(const float* __restrict__ fb, int n_frames, int hop, int win, long long pp) {
  // frames tp_lo .. tp_hi cover the position (offset pp - tp hop in [0, win)).  Their loads are issued together, six at a time, and
  // summed from the last frame backwards (as a rolled loop with an early exit the loads went out one by one: the border
  // samples of the gradient overlap-add, which take this path, were the tail of that launch)
  int tp_hi, tp_lo;
  if (pp < (1LL << 30)) {   // 32-bit divisions (the 64-bit one is a subroutine)
    const int q = static_cast<int>(pp);
    tp_hi = min(n_frames - 1, q / hop);
    tp_lo = q < win ? 0 : (q - win) / hop + 1;
  } else {
    tp_hi = static_cast<int>(min(static_cast<long long>(n_frames - 1), pp / hop));
    tp_lo = static_cast<int>((pp - win) / hop) + 1;
  }
  float acc = 0.f;
  constexpr int kFrames = 6;
  for (int t1 = tp_hi; t1 >= tp_lo; t1 -= kFrames) {
    float v[kFrames];
#pragma unroll
    for (int j = 0; j < kFrames; ++j) {
      const int tp = t1 - j;
      v[j] = 0.f;
      if (tp >= tp_lo) {
        const float* q = fb + static_cast<long long>(tp) * win + (pp - static_cast<long long>(tp) * hop);
        v[j] = L2 ? __ldcg(q) : *q;
      }
    }
#pragma unroll
    for (int j = 0; j < kFrames; ++j)
      if (t1 - j >= tp_lo) acc += v[j];
  }
  return acc;
}

struct FeatArgs {
  const float* x;
  BatchDev bd;
  float pre;
  ScaleDev mag_scale, mel_scale;
  float* mag;     // [frames, F] or null
  float* mel;     // [frames, n_mel] or null
  float2* spec;   // [frames, F] or null
};

constexpr int kFeatWarps = 8;

template <int N, bool PRE>
__global__ void __launch_bounds__(kFeatWarps * 32, 2) stft_feature_kernel(const PlanDev p, const FeatArgs a) {
  using C = FftCfg<N>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemTables<N> sm;
  sm.carve(smem_raw, p);
  sm.fill(p, p.window, a.mel != nullptr);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2* buf = sm.bufs + warp * C::kBufF2;
  const int rk = lane & 3, rm = (4 - rk) & 3;   // (-i)^k rotation index for bins k = lane + 32 i and Nz - k
  const long long warps_total = static_cast<long long>(gridDim.x) * kFeatWarps;
  // an item is 2Q frames (packed-engine granularity); this engine takes it as two independent Q-frame passes
  for (long long sub = static_cast<long long>(blockIdx.x) * kFeatWarps + warp; sub < 2 * a.bd.total_items;
       sub += warps_total) {
    Item it = decode_item(a.bd, sub >> 1, 2 * C::kQ);
    it.t0 += static_cast<int>(sub & 1) * C::kQ;
    if (it.t0 < it.T) {
    float2 v[32];
    load_frames<N, PRE>(v, a.x + it.sig_base, it.L, it.t0, it.T, p.hop, a.pre, sm.win, lane);
    fft_forward<N>(v, buf, sm.tw, lane);
    static_for<0, C::kQ>([&](auto qc) {
      constexpr int q = decltype(qc)::value;
      if (it.t0 + q < it.T) {
        float2* zq = buf + q * C::kZS;
        const long long row = (it.frame_base + it.t0 + q) * C::kF;
#pragma unroll 4
        for (int i = 0; i < C::kPairIters; ++i) {
          const int k = lane + 32 * i;
          const int km = (C::kNz - k) & (C::kNz - 1);
          float2 Ak, Am;
          split_fwd(zq[k], zq[km], sm.ws[k], Ak, Am);
          const float sk = fast_sqrt(fmaf(Ak.x, Ak.x, Ak.y * Ak.y));
          const float smg = fast_sqrt(fmaf(Am.x, Am.x, Am.y * Am.y));
          zq[k].x = sk;
          if (k != 0) zq[km].x = smg;
          if (a.mag) {
            a.mag[row + k] = apply_scale(a.mag_scale, sk);
            a.mag[row + C::kNz - k] = apply_scale(a.mag_scale, smg);
          }
          if (a.spec) {
            a.spec[row + k] = rot_fwd(Ak, rk);
            a.spec[row + C::kNz - k] = rot_fwd(Am, rm);
          }
        }
        if (lane == 0) {
          constexpr int k = C::kNz / 2;
          float2 Ak, Am;
          split_fwd(zq[k], zq[k], sm.ws[k], Ak, Am);
          const float sk = fast_sqrt(fmaf(Ak.x, Ak.x, Ak.y * Ak.y));
          zq[k].x = sk;
          if (a.mag) a.mag[row + k] = apply_scale(a.mag_scale, sk);
          if (a.spec) a.spec[row + k] = rot_fwd(Ak, k);
        }
      }
    });
    __syncwarp();
    if (a.mel) {
      mel_project_smem<N>(p, sm.melw, sm.mel_lo, buf, lane, [&](int q, int, int m, float val) {
        if (m < p.n_mel && it.t0 + q < it.T)
          a.mel[(it.frame_base + it.t0 + q) * p.n_mel + m] = apply_scale(a.mel_scale, val);
      });
    }
    __syncwarp();
    }
  }
}

template <int N>
inline size_t feat_smem_bytes(const PlanDev& p) {
  return SmemTables<N>::table_bytes(p.melw_count, p.mel_rounds) + sizeof(float2) * FftCfg<N>::kBufF2 * kFeatWarps;
}

}  // namespace sb200
