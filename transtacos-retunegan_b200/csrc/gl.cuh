// ISTFT and fused Griffin-Lim iteration kernels.
//   librosa.istft   (transtacos/audio.py:147-148)               -> istft_frames_kernel + ola_finish_kernel
//   _griffin_lim    (transtacos/audio.py:130-140, angle form)   -> gl_iter_kernel<N, 0>
//   librosa.griffinlim (retunegan/audio.py:131-136, fast/momentum form) -> gl_iter_kernel<N, 1>
//
// State between iterations is a "frames buffer" fb[frame][win]: each synthesised frame, already
// multiplied by the synthesis window and by 1/(n_fft * window-sum-square) at its position, so that the
// time signal is the plain sum of the (<= ceil(win/hop)) frames covering a sample.  One iteration is ONE
// launch: every warp gathers its Q frames of the current signal from the previous buffer (overlap-add as a
// gather -> deterministic, no atomics), re-analyses them (forward FFT), applies the phase update to the
// Hermitian pairs in registers, and synthesises the next buffer (inverse FFT) -- the time signal itself
// is only materialised once, by ola_finish_kernel.  The two buffers ping-pong; for fp32 / 5 s utterances the
// whole state (2 x 1.8 MB + S 1.8 MB [+ tprev 3.5 MB]) stays L2-resident between launches.
#pragma once
#include "feat.cuh"

namespace sb200 {

struct GlBatch {
  BatchDev bd;         // frame-described: frames_per_row / frame_off give T_b; sig_off / stride place the output rows
  long long length;    // uniform: output samples per row (<= 0: hop*(T-1)); ragged: sig_len[b] when length > 0
};

struct GlRow {
  int T;               // spectrogram frames
  int n_frames;        // frames synthesised: min(T, ceil((Ly + N)/hop)) (librosa.istft)
  long long Ly;        // output samples
  long long frame_base;
  long long out_base;
};

__device__ __forceinline__ GlRow gl_row(const GlBatch& g, int b, int N, int hop) {
  GlRow r;
  if (g.bd.item_off == nullptr) {
    r.T = static_cast<int>(g.bd.frames_per_row);
    r.frame_base = static_cast<long long>(b) * g.bd.frames_per_row;
    r.out_base = static_cast<long long>(b) * g.bd.stride;
    r.Ly = g.length > 0 ? g.length : static_cast<long long>(hop) * (r.T - 1);
  } else {
    r.frame_base = __ldg(g.bd.frame_off + b);
    r.T = static_cast<int>(__ldg(g.bd.frame_off + b + 1) - r.frame_base);
    r.out_base = __ldg(g.bd.sig_off + b);
    r.Ly = g.length > 0 ? __ldg(g.bd.sig_len + b) : static_cast<long long>(hop) * (r.T - 1);
  }
  const long long nf = (r.Ly + N + hop - 1) / hop;
  r.n_frames = g.length > 0 ? static_cast<int>(min(static_cast<long long>(r.T), nf)) : r.T;
  return r;
}

// Item decode for frame-described batches (no signal lengths involved).
__device__ __forceinline__ void gl_decode(const GlBatch& g, long long item, int Q, int* b, int* t0) {
  if (g.bd.item_off == nullptr) {
    const long long bb = item / g.bd.items_per_row;
    *b = static_cast<int>(bb);
    *t0 = static_cast<int>(item - bb * g.bd.items_per_row) * Q;
  } else {
    int lo = 0, hi = g.bd.B;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(g.bd.item_off + mid) <= item) lo = mid; else hi = mid;
    }
    *b = lo;
    *t0 = static_cast<int>(item - __ldg(g.bd.item_off + lo)) * Q;
  }
}

// Synthesis scale for sample m of frame t: window / (n_fft * wss) (librosa.istft: y[wss > tiny] /= wss).
__device__ __forceinline__ float synth_scale_edge(const PlanDev& p, int t, int n_frames, int m) {
  const int nov = (p.win - 1) / p.hop;
  float wss = 0.f;
  for (int d = -nov; d <= nov; ++d) {
    const int off = m - d * p.hop;
    if (t + d >= 0 && t + d < n_frames && off >= 0 && off < p.win) wss += __ldg(p.wsq + off);
  }
  const float w = __ldg(p.window + m) / p.n_fft;
  return wss > 1.1754943508222875e-38f ? w / wss : w;
}

// Inverse FFT of the natural-order Z' in buf, window / normalise, store Q frames to fb_out.
template <int N>
__device__ __forceinline__ void synth_store(const PlanDev& p, float2 (&v)[32], float2* buf, const float2* s_tw,
                                            const float* __restrict__ s_wnorm, float* __restrict__ fb_out /*utterance*/,
                                            int t0, const GlRow& row, int lane) {
  using C = FftCfg<N>;
  fft_inverse<N>(v, buf, s_tw, lane);
  const int nov = (p.win - 1) / p.hop;
  static_for<0, C::kQ>([&](auto qc) {
    constexpr int q = decltype(qc)::value;
    const int t = t0 + q;
    if (t < row.n_frames) {
      float* dst = fb_out + static_cast<long long>(t) * C::kWin;
      const bool interior = (t - nov >= 0) && (t + nov <= row.n_frames - 1);
      static_for<0, C::kR>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        const int m = 2 * lane + 64 * r;
        float2 w;
        if (interior) {
          w = *reinterpret_cast<const float2*>(s_wnorm + m);
        } else {
          w = make_float2(synth_scale_edge(p, t, row.n_frames, m), synth_scale_edge(p, t, row.n_frames, m + 1));
        }
        const float2 z = v[q * C::kR + r];
        *reinterpret_cast<float2*>(dst + m) = make_float2(z.x * w.x, z.y * w.y);
      });
    }
  });
}

struct GlArgs {
  GlBatch g;
  const float* S;           // [frames, F] magnitudes, or null when spec is given
  const float* init_phase;  // [frames, F] u in [0,1)
  const float2* spec;       // [frames, F] complex (istft API)
  const float* fb_in;       // [frames, win]
  float* fb_out;            // [frames, win]
  float2* tprev;            // [frames, F] (form 1)
  float alpha;              // momentum / (1 + momentum)
  int first;                // form 1: tprev not yet written (rebuilt = 0)
};

constexpr int kGlWarps = 8;

template <int N>
struct GlSmem {
  using C = FftCfg<N>;
  float* win;     // analysis window
  float* wnorm;   // synthesis window / (N * wss_interior)
  float2* tw;
  float2* ws;
  float2* bufs;
  static constexpr int kTwCount = (C::kR2 - 1) * 32;
  static constexpr int kWsCount = C::kNz / 2 + 2;
  static constexpr size_t kBytes =
      sizeof(float) * 2 * C::kWin + sizeof(float2) * (kTwCount + kWsCount) + sizeof(float2) * C::kBufF2 * kGlWarps;
  __device__ __forceinline__ void init(unsigned char* raw, const PlanDev& p) {
    win = reinterpret_cast<float*>(raw);
    wnorm = win + C::kWin;
    tw = reinterpret_cast<float2*>(wnorm + C::kWin);
    ws = tw + kTwCount;
    bufs = ws + kWsCount;
    for (int i = threadIdx.x; i < C::kWin; i += blockDim.x) {
      win[i] = p.window[i];
      wnorm[i] = p.wnorm[i];
    }
    for (int i = threadIdx.x; i < kTwCount; i += blockDim.x) tw[i] = p.tw[i];
    for (int i = threadIdx.x; i < kWsCount; i += blockDim.x) ws[i] = p.ws[i];
    __syncthreads();
  }
};

// spec (or S * exp(2 pi i u)) -> first frames buffer
template <int N>
__global__ void __launch_bounds__(kGlWarps * 32, 2) istft_frames_kernel(const PlanDev p, const GlArgs a) {
  using C = FftCfg<N>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GlSmem<N> sm;
  sm.init(smem_raw, p);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2* buf = sm.bufs + warp * C::kBufF2;
  const int rk = lane & 3, rm = (4 - rk) & 3;
  // an item is 2Q frames (packed-engine granularity); this engine takes it as two independent Q-frame passes
  for (long long sub = static_cast<long long>(blockIdx.x) * kGlWarps + warp; sub < 2 * a.g.bd.total_items;
       sub += static_cast<long long>(gridDim.x) * kGlWarps) {
    int b, t0;
    gl_decode(a.g, sub >> 1, 2 * C::kQ, &b, &t0);
    t0 += static_cast<int>(sub & 1) * C::kQ;
    const GlRow row = gl_row(a.g, b, N, p.hop);
    if (t0 < row.T) {
    auto fetch = [&](long long idx) -> float2 {
      if (a.spec) return __ldg(a.spec + idx);
      float s, c;
      sincospif(2.f * __ldg(a.init_phase + idx), &s, &c);
      const float mag = __ldg(a.S + idx);
      return make_float2(mag * c, mag * s);
    };
    static_for<0, C::kQ>([&](auto qc) {
      constexpr int q = decltype(qc)::value;
      float2* zq = buf + q * C::kZS;
      if (t0 + q < row.T) {
        const long long base = (row.frame_base + t0 + q) * C::kF;
#pragma unroll 2
        for (int i = 0; i < C::kPairIters; ++i) {
          const int k = lane + 32 * i;
          float2 Ak = rot_inv(fetch(base + k), rk), Am = rot_inv(fetch(base + C::kNz - k), rm);
          if (k == 0) { Ak.y = 0.f; Am.y = 0.f; }   // irfft ignores imag of DC / Nyquist
          float2 Zk, Zr;
          split_inv(Ak, Am, sm.ws[k], Zk, Zr);
          if (k != 0) zq[C::kNz - k] = Zr;
          zq[k] = Zk;
        }
        if (lane == 0) {
          constexpr int k = C::kNz / 2;
          const float2 A = rot_inv(fetch(base + k), k);
          float2 Zk, Zr;
          split_inv(A, A, sm.ws[k], Zk, Zr);
          zq[k] = Zk;
        }
      } else {
        for (int k = lane; k < C::kNz; k += 32) zq[k] = make_float2(0.f, 0.f);
      }
    });
    __syncwarp();
    float2 v[32];
    synth_store<N>(p, v, buf, sm.tw, sm.wnorm, a.fb_out + row.frame_base * C::kWin, t0, row, lane);
    }
  }
}

// One Griffin-Lim iteration: fb_in -> (signal gather, STFT, phase update, ISTFT) -> fb_out.
template <int N, int FORM>
__global__ void __launch_bounds__(kGlWarps * 32, 2) gl_iter_kernel(const PlanDev p, const GlArgs a) {
  using C = FftCfg<N>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GlSmem<N> sm;
  sm.init(smem_raw, p);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2* buf = sm.bufs + warp * C::kBufF2;
  const int rk = lane & 3, rm = (4 - rk) & 3;
  for (long long sub = static_cast<long long>(blockIdx.x) * kGlWarps + warp; sub < 2 * a.g.bd.total_items;
       sub += static_cast<long long>(gridDim.x) * kGlWarps) {
    int b, t0;
    gl_decode(a.g, sub >> 1, 2 * C::kQ, &b, &t0);
    t0 += static_cast<int>(sub & 1) * C::kQ;
    const GlRow row = gl_row(a.g, b, N, p.hop);
    const float* fb = a.fb_in + row.frame_base * C::kWin;
    if (t0 < row.T) {
    float2 v[32];
    // 1. gather the current signal under each analysis frame (reflect padded, np.pad mode='reflect')
    static_for<0, C::kQ>([&](auto qc) {
      constexpr int q = decltype(qc)::value;
      const int t = t0 + q;
      const long long p0 = static_cast<long long>(t) * p.hop - N / 4;
      static_for<0, C::kR>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        const int m = 2 * lane + 64 * r;
        float2 z = make_float2(0.f, 0.f);
        if (t < row.T) {
          auto sample = [&](long long i) -> float {
            if (i < 0) i = -i;
            if (i >= row.Ly) i = 2 * (row.Ly - 1) - i;
            return ola_gather(fb, row.n_frames, p.hop, C::kWin, i + N / 4);
          };
          const float2 w = *reinterpret_cast<const float2*>(sm.win + m);
          z = make_float2(sample(p0 + m) * w.x, sample(p0 + m + 1) * w.y);
        }
        fwd_put<N, q, r>(v, z);
      });
    });
    // 2. analysis
    fft_forward<N>(v, buf, sm.tw, lane);
    // 3. phase update on Hermitian pairs, in place Z -> Z'
    static_for<0, C::kQ>([&](auto qc) {
      constexpr int q = decltype(qc)::value;
      float2* zq = buf + q * C::kZS;
      if (t0 + q < row.T) {
        const long long base = (row.frame_base + t0 + q) * C::kF;
        auto update = [&](float2 X, long long idx) -> float2 {
          float2 ang;
          if constexpr (FORM == 0) {
            const float n2 = fmaf(X.x, X.x, X.y * X.y);
            const float inv = rsqrtf(n2);
            ang = n2 > 0.f ? make_float2(X.x * inv, X.y * inv) : make_float2(1.f, 0.f);
          } else {
            float2 c = X;
            if (!a.first) {
              const float2 tp = a.tprev[idx];
              c = make_float2(fmaf(-a.alpha, tp.x, X.x), fmaf(-a.alpha, tp.y, X.y));
            }
            a.tprev[idx] = X;
            const float inv = 1.f / (sqrtf(fmaf(c.x, c.x, c.y * c.y)) + 1e-16f);
            ang = make_float2(c.x * inv, c.y * inv);
          }
          const float mag = __ldg(a.S + idx);
          return make_float2(mag * ang.x, mag * ang.y);
        };
#pragma unroll 2
        for (int i = 0; i < C::kPairIters; ++i) {
          const int k = lane + 32 * i;
          const int km = (C::kNz - k) & (C::kNz - 1);
          float2 Ak, Am;
          split_fwd(zq[k], zq[km], sm.ws[k], Ak, Am);
          const float2 Yk = update(rot_fwd(Ak, rk), base + k);
          const float2 Ym = update(rot_fwd(Am, rm), base + C::kNz - k);
          float2 Bk = rot_inv(Yk, rk), Bm = rot_inv(Ym, rm);
          if (k == 0) { Bk.y = 0.f; Bm.y = 0.f; }
          float2 Zk, Zr;
          split_inv(Bk, Bm, sm.ws[k], Zk, Zr);
          if (k != 0) zq[km] = Zr;
          zq[k] = Zk;
        }
        if (lane == 0) {
          constexpr int k = C::kNz / 2;
          float2 Ak, Am;
          split_fwd(zq[k], zq[k], sm.ws[k], Ak, Am);
          const float2 Y = update(rot_fwd(Ak, k), base + k);
          const float2 B = rot_inv(Y, k);
          float2 Zk, Zr;
          split_inv(B, B, sm.ws[k], Zk, Zr);
          zq[k] = Zk;
        }
      } else {
        for (int k = lane; k < C::kNz; k += 32) zq[k] = make_float2(0.f, 0.f);
      }
    });
    __syncwarp();
    // 4. synthesis into the other buffer
    synth_store<N>(p, v, buf, sm.tw, sm.wnorm, a.fb_out + row.frame_base * C::kWin, t0, row, lane);
    }
  }
}

// y[j] = sum of the frames covering j  (librosa.istft tail: trim n_fft/2, fix_length)
struct OlaArgs {
  GlBatch g;
  const float* fb;
  float* y;
};
template <int N>
__global__ void ola_finish_kernel(const PlanDev p, const OlaArgs a) {
  const int b = blockIdx.y;
  const GlRow row = gl_row(a.g, b, N, p.hop);
  const float* fb = a.fb + row.frame_base * (N / 2);
  for (long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; j < row.Ly;
       j += static_cast<long long>(gridDim.x) * blockDim.x)
    a.y[row.out_base + j] = ola_gather(fb, row.n_frames, p.hop, N / 2, j + N / 4);
}

}  // namespace sb200
