// Small kernels around the FFT path: banded mel projection of an arbitrary spectrogram, the
// element-wise inverse scalings, and the pre-emphasis FIR / de-emphasis IIR (parallel scan).
#pragma once
#include "feat.cuh"

namespace sb200 {

// out[t, m] = sum_k basis[m, k] in[t, k]   (retunegan/audio.py:21, transtacos/audio.py:154-155)
template <int N>
__global__ void __launch_bounds__(kFeatWarps * 32) mel_project_kernel(const PlanDev p, const float* __restrict__ in,
                                                                      long long frames, const ScaleDev sc,
                                                                      float* __restrict__ out) {
  using C = FftCfg<N>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemTables<N> sm;
  sm.carve(smem_raw, p);
  sm.fill(p, p.window, true);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2* buf = sm.bufs + warp * C::kBufF2;
  const long long items = (frames + C::kQ - 1) / C::kQ;
  for (long long item = static_cast<long long>(blockIdx.x) * kFeatWarps + warp; item < items;
       item += static_cast<long long>(gridDim.x) * kFeatWarps) {
    const long long t0 = item * C::kQ;
#pragma unroll
    for (int q = 0; q < C::kQ; ++q) {
      const long long t = t0 + q;
      for (int k = lane; k < C::kNz; k += 32) buf[q * C::kZS + k].x = (t < frames) ? __ldg(in + t * C::kF + k) : 0.f;
    }
    __syncwarp();
    mel_project_smem<N>(p, sm.melw, sm.mel_lo, buf, lane, [&](int q, int, int m, float val) {
      if (m < p.n_mel && t0 + q < frames) out[(t0 + q) * p.n_mel + m] = apply_scale(sc, val);
    });
    __syncwarp();
  }
}

// out[t, k] = sum_m lin[k, m] mel[t, m] with the <= 2 non-zeros of row k (transtacos/audio.py:164-165 _mel_to_linear).
// HBM-bound: 4 (n_mel + F) bytes per frame; one thread per output element, consecutive bins per warp.
__global__ void mel_to_linear_kernel(const PlanDev p, const float* __restrict__ mel, long long frames, float* __restrict__ out) {
  const long long n = frames * p.F;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long t = i / p.F;
    const int k = static_cast<int>(i - t * p.F);
    const int r0 = __ldg(p.col_r0 + k);
    const float* row = mel + t * p.n_mel;
    float v = __ldg(p.lin_c0 + k) * __ldg(row + r0);
    if (r0 + 1 < p.n_mel) v = fmaf(__ldg(p.lin_c1 + k), __ldg(row + r0 + 1), v);
    out[i] = v;
  }
}

// mode 0: 10^(((in + p0) * (-p1) / (2 p0) + p1 + p2) / 20) ^ power ; mode 1: exp(in) ^ power ; mode 2: in ^ power (in >= 0)
__global__ void spec_to_amplitude_kernel(const float* __restrict__ in, long long n, int mode, float p0, float p1, float p2,
                                         float power, float* __restrict__ out) {
  const float kLog2_10 = 3.3219280948873623f, kLog2_e = 1.4426950408889634f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = __ldg(in + i);
    float l2;   // log2 of the amplitude
    if (mode == 0) {
      const float db = (v + p0) * (-p1) / (2.f * p0) + p1 + p2;
      l2 = db * 0.05f * kLog2_10;
    } else if (mode == 1) {
      l2 = v * kLog2_e;
    } else {
      l2 = log2f(v);
    }
    out[i] = exp2f(l2 * power);
  }
}

__device__ __forceinline__ void row_of(const BatchDev& bd, int b, long long* base, long long* L) {
  if (bd.sig_off) {
    *base = __ldg(bd.sig_off + b);
    *L = __ldg(bd.sig_len + b);
  } else {
    *base = b * bd.stride;
    *L = bd.len;
  }
}

// y[n] = x[n] - k x[n-1], zero initial state (transtacos/audio.py:64-66)
__global__ void preemphasis_kernel(const float* __restrict__ x, const BatchDev bd, float k, float* __restrict__ y) {
  long long base, L;
  row_of(bd, blockIdx.y, &base, &L);
  for (long long n = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; n < L;
       n += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float prev = n > 0 ? __ldg(x + base + n - 1) : 0.f;
    y[base + n] = fmaf(-k, prev, __ldg(x + base + n));
  }
}

// y[n] = x[n] + k y[n-1] (transtacos/audio.py:69-70): one CTA per row, tiles of 256 x 8 samples.
// Within a tile every thread runs the recurrence over its 8 samples from a zero state, the
// per-thread carries (a = k^8, b = local tail) are combined with a warp-shuffle + smem scan of the
// affine maps s -> a s + b, and the incoming state is folded back in.
constexpr int kScanThreads = 256, kScanPer = 8;
__global__ void __launch_bounds__(kScanThreads) inv_preemphasis_kernel(const float* x, const BatchDev bd, float k,
                                                                       float* y) {   // x == y allowed (in place)
  __shared__ float tile[kScanThreads * kScanPer];
  __shared__ float warp_b[kScanThreads / 32];
  long long base, L;
  row_of(bd, blockIdx.x, &base, &L);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float kp[kScanPer + 1];   // k^1 .. k^8
  kp[0] = 1.f;
#pragma unroll
  for (int i = 1; i <= kScanPer; ++i) kp[i] = kp[i - 1] * k;
  const float a_thread = kp[kScanPer];
  float a_pow_lane = 1.f;   // a_thread^lane
  for (int i = 0; i < lane; ++i) a_pow_lane *= a_thread;
  float a_warp = 1.f;       // a_thread^32
  for (int i = 0; i < 32; ++i) a_warp *= a_thread;
  float state = 0.f;        // y[tile_start - 1]
  for (long long t0 = 0; t0 < L; t0 += kScanThreads * kScanPer) {
    for (int i = tid; i < kScanThreads * kScanPer; i += kScanThreads)
      tile[i] = (t0 + i < L) ? x[base + t0 + i] : 0.f;
    __syncthreads();
    float loc[kScanPer];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kScanPer; ++i) {
      s = fmaf(k, s, tile[tid * kScanPer + i]);
      loc[i] = s;
    }
    // inclusive scan over lanes of affine maps (a, b): combine(prev, cur) = (a_p a_c, a_c b_p + b_c); a is a power of a_thread
    float b = s, a = a_thread;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float bp = __shfl_up_sync(kFullMask, b, d);
      const float ap = __shfl_up_sync(kFullMask, a, d);
      if (lane >= d) {
        b = fmaf(a, bp, b);
        a *= ap;
      }
    }
    if (lane == 31) warp_b[warp] = b;
    __syncthreads();
    // state entering this warp
    float win_state = state;
    for (int w = 0; w < warp; ++w) win_state = fmaf(a_warp, win_state, warp_b[w]);
    // state entering this thread: exclusive prefix within the warp applied to win_state
    float b_excl = __shfl_up_sync(kFullMask, b, 1);
    if (lane == 0) b_excl = 0.f;
    const float tin = fmaf(a_pow_lane, win_state, b_excl);
#pragma unroll
    for (int i = 0; i < kScanPer; ++i) tile[tid * kScanPer + i] = fmaf(kp[i + 1], tin, loc[i]);
    // next tile's incoming state = value of the last sample of this tile
    float next_state = state;
    for (int w = 0; w < kScanThreads / 32; ++w) next_state = fmaf(a_warp, next_state, warp_b[w]);
    state = next_state;
    __syncthreads();
    for (int i = tid; i < kScanThreads * kScanPer; i += kScanThreads)
      if (t0 + i < L) y[base + t0 + i] = tile[i];
    __syncthreads();
  }
}

// Frame statistics that share the STFT framing (frame_length samples, hop, centred):
//   rms[t] = sqrt(mean(x_reflect[t*hop + m]^2))                 librosa.feature.rms, pad_mode='reflect'
//            (transtacos/audio.py:112-114, retunegan/audio.py:103-105 get_c0; librosa.effects.trim at transtacos/audio.py:59-61)
//   zcr[t] = #{m in [1, FL): signbit(x_edge[m]) != signbit(x_edge[m-1])} / FL   librosa.feature.zero_crossing_rate
//            (retunegan/audio.py:98-100 get_zcr): edge padding, |x| <= 1e-10 counts as +0, the first sample of a frame never counts.
// One warp per frame, lanes stride the frame; the two sums are reduced by shuffles (deterministic).  frames_per_row /
// frame_off describe the output rows; HBM-bound: 4 B per sample (neighbouring frames hit L1 / L2), 4-8 B per frame.
struct FrameStatsArgs {
  const float* x;
  BatchDev bd;          // sig_off / sig_len / frame_off (ragged) or len / stride / frames_per_row (uniform)
  int frame_length, hop;
  float* rms;           // [frames] or null
  float* zcr;           // [frames] or null
  long long total_frames;
};

__global__ void __launch_bounds__(256) frame_stats_kernel(const FrameStatsArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * static_cast<long long>(blockDim.x)) >> 5;
  const int FL = a.frame_length, half = FL / 2;
  for (long long f = warp; f < a.total_frames; f += nwarps) {
    long long base, L;
    int t;
    if (a.bd.frame_off == nullptr) {
      const long long b = f / a.bd.frames_per_row;
      base = b * a.bd.stride;
      L = a.bd.len;
      t = static_cast<int>(f - b * a.bd.frames_per_row);
    } else {
      int lo = 0, hi = a.bd.B;   // largest b with frame_off[b] <= f
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a.bd.frame_off + mid) <= f) lo = mid; else hi = mid;
      }
      base = __ldg(a.bd.sig_off + lo);
      L = __ldg(a.bd.sig_len + lo);
      t = static_cast<int>(f - __ldg(a.bd.frame_off + lo));
    }
    const long long i0 = static_cast<long long>(t) * a.hop - half;   // signal index of the frame's first sample
    const float* x = a.x + base;
    float ss = 0.f;
    int cross = 0;
    const bool interior = i0 >= 1 && i0 + FL <= L;
    for (int m = lane; m < FL; m += 32) {
      const long long i = i0 + m;
      float xr, xe, xp;
      if (interior) {
        xr = xe = __ldg(x + i);
        xp = __ldg(x + i - 1);
      } else {
        long long ir = i < 0 ? -i : i;
        ir = ir >= L ? 2 * (L - 1) - ir : ir;
        ir = min(max(ir, 0LL), L - 1);     // signals shorter than frame_length/2: librosa raises; stay in bounds
        xr = __ldg(x + ir);
        xe = __ldg(x + min(max(i, 0LL), L - 1));
        xp = __ldg(x + min(max(i - 1, 0LL), L - 1));
      }
      ss = fmaf(xr, xr, ss);
      const bool neg = (fabsf(xe) <= 1e-10f) ? false : signbit(xe);
      const bool negp = (fabsf(xp) <= 1e-10f) ? false : signbit(xp);
      cross += (m > 0 && neg != negp) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ss += __shfl_xor_sync(kFullMask, ss, o);
      cross += __shfl_xor_sync(kFullMask, cross, o);
    }
    if (lane == 0) {
      if (a.rms) a.rms[f] = sqrtf(ss / static_cast<float>(FL));
      if (a.zcr) a.zcr[f] = static_cast<float>(cross) / static_cast<float>(FL);
    }
  }
}

// YIN fundamental-frequency track (librosa 0.8.1 yin; transtacos/audio.py:107-109 get_f0), one CTA per frame:
//   frame of frame_length samples (centred, reflect padded), W = frame_length / 2
//   acf[tau] = sum_{j=1..W} y[j] y[j+tau],  e[tau] = sum_{j=tau+1..tau+W} y[j]^2   (both zeroed where |.| < 1e-6, as librosa does)
//   d[tau]   = e[0] + e[tau] - 2 acf[tau]                                             (difference function)
//   d'[i]    = d[pmin+i] / (mean_{1..pmin+i} d + tiny),  i = 0..pmax-pmin             (cumulative mean normalised difference)
//   trough   = local minimum of d' (librosa.util.localmax of -d'; first element: d'[0] < d'[1]) with d' < trough_threshold
//   period   = pmin + first such i (else argmin d') + parabolic shift;  f0 = sr / period
// FP32-bound: W (pmax+1) multiply-adds per frame (155 k at the reference settings) from shared memory.  A thread owns FOUR
// consecutive lags and walks j four samples at a time: the 7 samples y[j+tau0 .. j+tau0+6] slide through registers, so a trip costs
// one 16-byte broadcast load of y[j..j+3] plus four 4-byte loads for 16 multiply-adds.  The frame is kept twice: in natural
// order (broadcast side) and de-interleaved by 4 (plane c holds y[4i+c]: the lane-strided side is then consecutive across the
// threads, conflict free).  The energy terms come from a prefix sum of y^2 (as librosa's cumsum does).
constexpr int kYinThreads = 96;       // 76 lag groups at the reference settings: 3 warps (128 threads: 0.53 ms, 96: 0.44 ms per 64 x 5 s)
constexpr int kYinMaxFrame = 4096;   // samples per frame the shared-memory layout supports
struct YinArgs {
  const float* x;
  BatchDev bd;
  int frame_length, hop, pmin, pmax;
  float sr, threshold;
  float* f0;            // [frames]
  long long total_frames;
};

__host__ __device__ inline int yin_plane_len(int frame_length) { return frame_length / 4 + 4; }
__host__ __device__ inline size_t yin_smem_floats(int frame_length, int pmax, int pmin) {
  // ynat [FL + 4] | planes [4][FL/4 + 4] | cs [FL + 1 (+3 pad)] | d [pmax + 2 (+ pad)] | dn [n]
  return static_cast<size_t>(frame_length + 4) + 4 * yin_plane_len(frame_length) + (frame_length + 4) + (pmax + 8) + (pmax - pmin + 1);
}

__global__ void __launch_bounds__(kYinThreads) yin_kernel(const YinArgs a) {
  extern __shared__ __align__(16) float ysm[];
  const int FL = a.frame_length, W = FL / 2, n = a.pmax - a.pmin + 1, PL = yin_plane_len(FL);
  float* ynat = ysm;                     // ynat[i] = y[i + 1]: y[j0 .. j0+3] with j0 = 1 (mod 4) is one aligned 16-byte load
  float* yp = ynat + FL + 4;             // yp[c * PL + i] = y[4 i + c]
  float* cs = yp + 4 * PL;               // cs[i] = sum_{j=1..i} y[j]^2
  float* d = cs + FL + 4;
  float* dn = d + a.pmax + 8;
  __shared__ float s_warp[kYinThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (long long f = blockIdx.x; f < a.total_frames; f += gridDim.x) {
    long long base, L;
    int t;
    if (a.bd.frame_off == nullptr) {
      const long long b = f / a.bd.frames_per_row;
      base = b * a.bd.stride;
      L = a.bd.len;
      t = static_cast<int>(f - b * a.bd.frames_per_row);
    } else {
      int lo = 0, hi = a.bd.B;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a.bd.frame_off + mid) <= f) lo = mid; else hi = mid;
      }
      base = __ldg(a.bd.sig_off + lo);
      L = __ldg(a.bd.sig_len + lo);
      t = static_cast<int>(f - __ldg(a.bd.frame_off + lo));
    }
    const long long i0 = static_cast<long long>(t) * a.hop - FL / 2;
    for (int m = tid; m < FL + 4; m += kYinThreads) {   // 4 zero samples of padding behind the frame
      float v = 0.f;
      if (m < FL) {
        long long i = i0 + m;
        i = i < 0 ? -i : i;
        i = i >= L ? 2 * (L - 1) - i : i;
        i = min(max(i, 0LL), L - 1);
        v = __ldg(a.x + base + i);
      }
      if (m >= 1) ynat[m - 1] = v;
      yp[(m & 3) * PL + (m >> 2)] = v;
    }
    if (tid == 0) ynat[FL + 3] = 0.f;
    __syncthreads();
    // prefix sums of y[j]^2, j = 1..FL-1: per-thread runs, block scan of the run totals
    {
      const int per = (FL + kYinThreads - 1) / kYinThreads;
      const int lo = 1 + tid * per, hi = min(FL, lo + per);
      float run = 0.f;
      for (int j = lo; j < hi; ++j) run = fmaf(ynat[j - 1], ynat[j - 1], run);
      float incl = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) s_warp[warp] = incl;
      __syncthreads();
      float off = 0.f;
      for (int w = 0; w < warp; ++w) off += s_warp[w];
      float c = off + incl - run;
      if (tid == 0) cs[0] = 0.f;
      for (int j = lo; j < hi; ++j) {
        c = fmaf(ynat[j - 1], ynat[j - 1], c);
        cs[j] = c;
      }
    }
    __syncthreads();
    // difference function, four lags per thread
    {
      float e0 = cs[W] - cs[0];
      if (fabsf(e0) < 1e-6f) e0 = 0.f;
      const float4* ynat4 = reinterpret_cast<const float4*>(ynat);
      for (int g = tid; 4 * g <= a.pmax; g += kYinThreads) {
        float acf[4] = {0.f, 0.f, 0.f, 0.f};
        float w0 = yp[1 * PL + g], w1 = yp[2 * PL + g], w2 = yp[3 * PL + g];   // y[tau0+1], y[tau0+2], y[tau0+3], tau0 = 4 g
        const float* q0 = yp + g + 1;
#pragma unroll 2
        for (int m = 0; m < W / 4; ++m) {                 // j0 = 1 + 4 m
          const float4 yj = ynat4[m];                     // y[j0 .. j0+3]
          const float n0 = q0[m], n1 = q0[PL + m], n2 = q0[2 * PL + m], n3 = q0[3 * PL + m];   // y[j0+tau0+3 .. +6]
          acf[0] = fmaf(yj.x, w0, acf[0]); acf[1] = fmaf(yj.x, w1, acf[1]); acf[2] = fmaf(yj.x, w2, acf[2]); acf[3] = fmaf(yj.x, n0, acf[3]);
          acf[0] = fmaf(yj.y, w1, acf[0]); acf[1] = fmaf(yj.y, w2, acf[1]); acf[2] = fmaf(yj.y, n0, acf[2]); acf[3] = fmaf(yj.y, n1, acf[3]);
          acf[0] = fmaf(yj.z, w2, acf[0]); acf[1] = fmaf(yj.z, n0, acf[1]); acf[2] = fmaf(yj.z, n1, acf[2]); acf[3] = fmaf(yj.z, n2, acf[3]);
          acf[0] = fmaf(yj.w, n0, acf[0]); acf[1] = fmaf(yj.w, n1, acf[1]); acf[2] = fmaf(yj.w, n2, acf[2]); acf[3] = fmaf(yj.w, n3, acf[3]);
          w0 = n1; w1 = n2; w2 = n3;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int tau = 4 * g + q;
          if (tau <= a.pmax) {
            float ac = acf[q], e = cs[tau + W] - cs[tau];
            if (fabsf(ac) < 1e-6f) ac = 0.f;
            if (fabsf(e) < 1e-6f) e = 0.f;
            d[tau] = e0 + e - 2.f * ac;
          }
        }
      }
    }
    __syncthreads();
    // cumulative sums of d[1..pmax] (warp 0: per-lane runs + shuffle scan), then d'
    if (warp == 0) {
      const int per = (a.pmax + 31) / 32;
      const int lo = 1 + lane * per, hi = min(a.pmax + 1, lo + per);
      float run = 0.f;
      for (int k = lo; k < hi; ++k) run += d[k];
      float incl = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl += v;
      }
      float cs = incl - run;       // sum of d[1 .. lo-1]
      for (int k = lo; k < hi; ++k) {
        cs += d[k];
        const int i = k - a.pmin;
        if (i >= 0) dn[i] = d[k] / (cs / static_cast<float>(k) + 1.17549435e-38f);
      }
    }
    __syncthreads();
    // first trough below the threshold, else the global minimum (warp 0)
    if (warp == 0) {
      int first = 1 << 30, amin = 0;
      float vmin = 3.4e38f;
      for (int i = lane; i < n; i += 32) {
        const float v = dn[i];
        bool trough;
        if (i == 0) trough = n > 1 && v < dn[1];
        else if (i == n - 1) trough = v < dn[i - 1];
        else trough = v < dn[i - 1] && v <= dn[i + 1];
        if (trough && v < a.threshold) first = min(first, i);
        if (v < vmin) { vmin = v; amin = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        first = min(first, __shfl_xor_sync(kFullMask, first, o));
        const float ov = __shfl_xor_sync(kFullMask, vmin, o);
        const int oi = __shfl_xor_sync(kFullMask, amin, o);
        if (ov < vmin || (ov == vmin && oi < amin)) { vmin = ov; amin = oi; }
      }
      if (lane == 0) {
        const int i = first < (1 << 30) ? first : amin;
        float shift = 0.f;
        if (i >= 1 && i <= n - 2) {
          const float pa = (dn[i - 1] + dn[i + 1] - 2.f * dn[i]) * 0.5f, pb = (dn[i + 1] - dn[i - 1]) * 0.5f;
          shift = -pb / (2.f * pa + 1.17549435e-38f);
          if (fabsf(shift) > 1.f) shift = 0.f;
        }
        a.f0[f] = a.sr / (static_cast<float>(a.pmin + i) + shift);
      }
    }
    __syncthreads();
  }
}

// Waveform max-pool losses of RetuneGAN (retunegan/models/loss.py:66-82, MaxPool1d(envelope_pool_k = 160), stride = kernel):
//   hi = max over the window, lo = max(-y) = -min;
//   envelope = mean|hi - hi_g| + mean|lo - lo_g|;   dynamic = mean| |hi + lo| - |hi_g + lo_g| |
// One warp per window; value AND the gradient w.r.t. y_g for a unit upstream gradient in the same pass (the gradient of a max
// pool goes to the first arg-max, as in ATen).  Partial sums per warp, summed in a fixed order by pool_loss_finalize_kernel.
// HBM-bound: 8 B per sample read, 4 B per sample of gradient written.
struct PoolLossArgs {
  const float* y;
  const float* yg;
  int B, k, mode;           // mode 0: envelope, 1: dynamic
  long long T, n_win;       // samples per row, windows per row (T / k)
  float* grad;              // [B, T], zero-filled by the kernel, or null
  float* partials;          // [gridDim.x * warps per block]
  float inv_n;              // 1 / (B * n_win)
};

__device__ __forceinline__ void warp_argmax(float& v, int& i) {   // larger value wins, ties -> lower index
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(kFullMask, v, o);
    const int oi = __shfl_xor_sync(kFullMask, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
}

__global__ void __launch_bounds__(256) pool_loss_kernel(const PoolLossArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * static_cast<long long>(blockDim.x)) >> 5;
  const long long total = static_cast<long long>(a.B) * a.n_win;
  float acc = 0.f;
  for (long long w = warp; w < total; w += nwarps) {
    const long long b = w / a.n_win, base = b * a.T + (w - b * a.n_win) * a.k;
    float hi = -3.4e38f, nlo = -3.4e38f, hig = -3.4e38f, nlog = -3.4e38f;   // nlo = max(-y)
    int ihig = 0, ilog = 0, d0 = 0, d1 = 0;
    for (int m = lane; m < a.k; m += 32) {
      const float v = __ldg(a.y + base + m), g = __ldg(a.yg + base + m);
      if (v > hi) hi = v;
      if (-v > nlo) nlo = -v;
      if (g > hig) { hig = g; ihig = m; }
      if (-g > nlog) { nlog = -g; ilog = m; }
      if (a.grad) a.grad[base + m] = 0.f;
    }
    warp_argmax(hi, d0);
    warp_argmax(nlo, d1);
    warp_argmax(hig, ihig);
    warp_argmax(nlog, ilog);
    __syncwarp();
    auto sgn = [](float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); };
    if (a.mode == 0) {
      const float e0 = hi - hig, e1 = nlo - nlog;
      acc += fabsf(e0) + fabsf(e1);
      if (a.grad && lane == 0) {
        a.grad[base + ihig] += -sgn(e0) * a.inv_n;      // d|hi - hi_g| / d y_g[argmax]
        a.grad[base + ilog] += sgn(e1) * a.inv_n;       // lo_g = -y_g[argmin]
      }
    } else {
      const float dyn = fabsf(hi + nlo), dg = hig + nlog, e = dyn - fabsf(dg);
      acc += fabsf(e);
      if (a.grad && lane == 0) {
        const float c = -sgn(e) * sgn(dg) * a.inv_n;    // d|dyn - |dg|| / d dg
        a.grad[base + ihig] += c;
        a.grad[base + ilog] += -c;
      }
    }
  }
  // the tail of every row (T % k samples) is outside every window: zero gradient
  if (a.grad) {
    const long long tail = a.T - a.n_win * a.k;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < tail * a.B;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const long long b = i / tail;
      a.grad[b * a.T + a.n_win * a.k + (i - b * tail)] = 0.f;
    }
  }
  if (lane == 0) a.partials[warp] = acc;    // every lane holds the same acc (the reductions broadcast)
}

__global__ void pool_loss_finalize_kernel(const float* __restrict__ partials, int n, float inv_n, float* __restrict__ loss) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partials[i];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(kFullMask, s, d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += red[w];
    *loss = t * inv_n;
  }
}

// ---- librosa.effects.trim bounds from a precomputed RMS track (transtacos/audio.py:59-61 trim_silence) ---------------------
// Row b owns frames [frame_off[b], frame_off[b+1]) of rms (uniform: frames_per_row each).  db = 10 log10(max(1e-10, rms^2)) -
// 10 log10(max(1e-10, max_t rms^2)) (power_to_db with ref = max, amin = 1e-10, top_db = None); non-silent frames are those with
// db > -top_db; out[2b] = first non-silent frame, out[2b+1] = last non-silent frame + 1, (0, 0) if there is none.  The
// comparison is carried out in double like numpy does (the track itself is float32).  One warp per row.
struct TrimArgs {
  const float* rms;
  const long long* frame_off;   // [B+1] or null (uniform)
  long long frames_per_row;
  int B;
  double top_db;
  long long* out;               // [B, 2]
};
__global__ void __launch_bounds__(128) trim_bounds_kernel(const TrimArgs a) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int b = warp; b < a.B; b += nwarps) {
    const long long f0 = a.frame_off ? __ldg(a.frame_off + b) : b * a.frames_per_row;
    const long long T = a.frame_off ? __ldg(a.frame_off + b + 1) - f0 : a.frames_per_row;
    double mx = 0.0;
    for (long long t = lane; t < T; t += 32) {
      const double r = static_cast<double>(__ldg(a.rms + f0 + t));
      mx = fmax(mx, r * r);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    const double ref_db = 10.0 * log10(fmax(1e-10, mx));
    long long first = T, last = -1;
    for (long long t = lane; t < T; t += 32) {
      const double r = static_cast<double>(__ldg(a.rms + f0 + t));
      const double db = 10.0 * log10(fmax(1e-10, r * r)) - ref_db;
      if (db > -a.top_db) {
        first = min(first, t);
        last = max(last, t);
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      first = min(first, __shfl_xor_sync(0xffffffffu, first, d));
      last = max(last, __shfl_xor_sync(0xffffffffu, last, d));
    }
    if (lane == 0) {
      a.out[2 * b] = last >= 0 ? first : 0;
      a.out[2 * b + 1] = last >= 0 ? last + 1 : 0;
    }
  }
}

}  // namespace sb200
