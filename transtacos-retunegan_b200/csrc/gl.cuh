// Batch / row descriptors and edge-frame synthesis weights shared by the ISTFT / Griffin-Lim kernels (gl2.cuh).
// (The first-generation kernels that lived here -- one warp per item with a frames buffer between launches -- were replaced
// by the tiled kernels of gl2.cuh.)
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

}  // namespace sb200
