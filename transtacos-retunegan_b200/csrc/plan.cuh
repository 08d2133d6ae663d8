// Immutable per-configuration tables (window, twiddles, banded mel filterbank) -- the B200-side
// replacement of the reference's cached mel_basis / window_fn_torch (retunegan/audio.py:20,25-26,
// 153-159; transtacos/audio.py:151-162).  Built once on the host in double precision, rounded to
// float32 the way the reference's libraries do, and uploaded to the current device.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <functional>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/spectral_b200.h"

namespace sb200 {

constexpr int kMaxMelRounds = 4;   // n_mel <= 128
constexpr int kMaxSegRounds = 5;   // n_mel + 1 <= 129 segments between neighbouring filter centres

// Device view of a plan, passed by value to kernels.
struct PlanDev {
  int n_fft, win, hop, n_mel, F;
  const float* window;    // [win]            analysis window
  const float* wout;      // [win]            window / n_fft                      (synthesis, un-normalised OLA)
  const float* wnorm;     // [win]            window / (n_fft * wss_interior[m])  (synthesis, interior frames)
  const float* wsq;       // [win]            window^2
  const float* wedge;     // [2][nov][win]    synthesis window / (n_fft * wss) for the first nov frames ([0][t]) and the last
                          //                  nov frames ([1][t'], t' = n_frames-1-t) of an utterance with >= 2 nov + 1 frames
  int nov;                // (win - 1) / hop: frames overlapping a given frame on each side
  const float2* tw;       // [(2R-1)*32]      w_Nz^{k1*lane}, row k1-1
  const float2* ws;       // [Nz/2+1]         -0.5i * w_N^k
  const float2* sp2;      // [17*32]          packed-engine split twiddles, slot s, lane l: k = l%R2 + R2*s, phi = 2 pi k/N:
                          //                  s < 8: (tan phi, cos phi); s >= 8: (cot phi, sin phi)   (fft2.cuh split2)
  const float2* spn;      // [Nz/2+2]         split twiddles in natural bin order, phi = 2 pi k/N: k < Nz/4: (tan phi, cos phi),
                          //                  k >= Nz/4: (cot phi, sin phi)   (gl2.cuh spectrum pass)
  // banded mel, ELL per round of 32 rows: melw[round_off[r] + it*32 + lane], it < round_len[r]
  const float* melw;
  const int* mel_lo;      // [32*rounds]      slot (round, lane): first bin read | mel row << 16 (row 0x7fff: empty slot)
  int mel_rounds;
  int mel_round_off[kMaxMelRounds];
  int mel_round_len[kMaxMelRounds];
  int melw_count;         // floats in melw
  int mel_kmin, mel_kmax; // first / last bin with a non-zero filter weight
  // Segment view of the triangular filterbank (feat3.cuh): segment j = the bins between filter centres c_j and c_{j+1}
  // (j = 0 .. n_mel), where filter j rises with r_k = r0_j + a_j (k - k0_j) and filter j-1 falls with 1 - r_k.  With
  // A_j = sum S_k and B_j = sum (k - k0_j) S_k over the segment, R_j = r0_j A_j + a_j B_j, E_j = A_j - R_j:
  //   mel[m] = enorm_m (R_m + E_{m+1}).
  // Every band bin is read ONCE and no weight is loaded.  Segments sit in their natural order (round j / 32, lane j % 32);
  // lane l of round r starts its reads seg_slot[].y bins early (masked) so that the 16 lanes of a half-warp hit 16 different
  // 8-byte bank pairs in every trip (exact matching, minimal round length).
  const int4* seg_slot;   // [32*seg_rounds]  {first bin read = k0 - d, d, len, 1 if the slot holds a segment}
  const float2* seg_coef; // [32*seg_rounds]  {r0, a}
  const float* mel_enorm; // [n_mel]          2 / (c_{m+2} - c_m)  (Slaney area normalisation)
  int seg_rounds;
  int seg_round_len[kMaxSegRounds];
  // column view (<= 2 non-zeros per column, consecutive rows): basis[r0[k], k] = c0[k], basis[r0[k]+1, k] = c1[k]
  const int* col_r0;      // [F]
  const float* col_c0;    // [F]
  const float* col_c1;    // [F]
  // the same view packed for shared memory (mstft backward): {c0, c1} per bin, padded to an even count, and r0 as one byte per bin,
  // padded to a multiple of 16 (n_mel <= 128)
  const float2* col_c01;  // [(F + 1) & ~1]
  const unsigned char* col_r8;   // [(F + 15) & ~15]
  // pseudo-inverse ("linear") basis of transtacos/audio.py:167-175: lin[k, m] = basis[m, k] * dinv[m] with
  // dinv = 1 / column sums of basis basis^T; in the column view: lin[k, r0[k]] = lin_c0[k], lin[k, r0[k]+1] = lin_c1[k]
  const float* lin_c0;    // [F]
  const float* lin_c1;    // [F]
};

}  // namespace sb200

struct sb200_plan {
  sb200_config cfg;
  sb200::PlanDev dev;
  int device;
  std::vector<float> mel_dense;   // [n_mel * F] float32, == librosa.filters.mel
  std::vector<float> window_f32;  // [win]
  std::vector<void*> allocs;
};

namespace sb200 {

inline double hz_to_mel(double f, bool htk) {
  if (htk) return 2595.0 * std::log10(1.0 + f / 700.0);
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
inline double mel_to_hz(double m, bool htk) {
  if (htk) return 700.0 * (std::pow(10.0, m / 2595.0) - 1.0);
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

// librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk, norm='slaney', dtype=float32)  
inline std::vector<float> mel_filterbank(int sr, int n_fft, int n_mels, double fmin, double fmax, bool htk) {
  const int F = 1 + n_fft / 2;
  std::vector<double> mel_f(n_mels + 2);
  const double m0 = hz_to_mel(fmin, htk), m1 = hz_to_mel(fmax, htk);
  for (int i = 0; i < n_mels + 2; ++i) {
    // np.linspace(m0, m1, n): start + i*step with the last point pinned to m1
    const double step = (m1 - m0) / (n_mels + 1);
    mel_f[i] = mel_to_hz(i == n_mels + 1 ? m1 : m0 + i * step, htk);
  }
  std::vector<float> w(static_cast<size_t>(n_mels) * F, 0.f);
  for (int i = 0; i < n_mels; ++i) {
    const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
    const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
    for (int k = 0; k < F; ++k) {
      const double f = (k == F - 1) ? sr / 2.0 : k * ((sr / 2.0) / (F - 1));   // np.linspace(0, sr/2, F)
      const double lower = -(mel_f[i] - f) / fd0, upper = (mel_f[i + 2] - f) / fd1;
      const float tri = static_cast<float>(std::fmax(0.0, std::fmin(lower, upper)));   // stored f32 ...
      w[static_cast<size_t>(i) * F + k] = static_cast<float>(static_cast<double>(tri) * enorm);   // ... then *= enorm
    }
  }
  return w;
}

// scipy.signal.get_window(name, M, fftbins=True) in double
inline std::vector<double> make_window(int kind, int M) {
  std::vector<double> w(M);
  const double pi = 3.14159265358979323846;
  for (int n = 0; n < M; ++n) {
    const double t = 2.0 * pi * n / M;
    switch (kind) {
      case SB200_WIN_HANN: w[n] = 0.5 - 0.5 * std::cos(t); break;
      case SB200_WIN_HAMMING: w[n] = 0.54 - 0.46 * std::cos(t); break;
      case SB200_WIN_BLACKMAN: w[n] = 0.42 - 0.5 * std::cos(t) + 0.08 * std::cos(2 * t); break;
      default: w[n] = (n <= M / 2) ? 2.0 * n / M : 2.0 - 2.0 * n / M; break;   // bartlett(M+1)[:-1]
    }
  }
  return w;
}

template <class T>
inline cudaError_t upload(sb200_plan* p, const std::vector<T>& h, const T** out) {
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, std::max<size_t>(h.size(), 1) * sizeof(T));
  if (e != cudaSuccess) return e;
  p->allocs.push_back(d);
  e = cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  *out = static_cast<const T*>(d);
  return e;
}

// Returns "" on success, otherwise an error message (status in *st).
inline std::string build_plan(const sb200_config& c, sb200_plan* p, int* st) {
  *st = SB200_ERR_INVALID;
  if (c.n_fft != 2048 && c.n_fft != 1024 && c.n_fft != 512) {
    *st = SB200_ERR_UNSUPPORTED;
    return "n_fft must be 512, 1024 or 2048";
  }
  if (c.win_length != c.n_fft / 2) {
    *st = SB200_ERR_UNSUPPORTED;
    return "win_length must equal n_fft/2 (all reference configurations: hparam.py n_fft/win = 2048/1024, 1024/512, 512/256)";
  }
  if (c.hop_length < 2 || c.hop_length > c.win_length || (c.hop_length & 1)) return "hop_length must be even and in [2, win_length]";
  if (c.n_mel < 1 || c.n_mel > 32 * kMaxMelRounds) return "n_mel must be in [1, 128]";
  if (!(c.fmax < c.sample_rate / 2)) return "fmax must be < sample_rate // 2 (transtacos/audio.py:160)";
  if (!(c.fmin >= 0 && c.fmin < c.fmax)) return "need 0 <= fmin < fmax";
  if (c.window < 0 || c.window > 3) return "unknown window";
  const double pi = 3.14159265358979323846;
  const int N = c.n_fft, Nz = N / 2, win = c.win_length, hop = c.hop_length, F = N / 2 + 1, R2 = N / 64;
  p->cfg = c;
  cudaGetDevice(&p->device);
  PlanDev& d = p->dev;
  d.n_fft = N; d.win = win; d.hop = hop; d.n_mel = c.n_mel; d.F = F;

  const std::vector<double> w = make_window(c.window, win);
  std::vector<float> wf(win), wout(win), wnorm(win), wsq(win);
  // interior window-sum-square at offset m (periodic in hop): sum over all shifts of w^2
  std::vector<double> wss(win, 0.0);
  for (int m = 0; m < win; ++m)
    for (int j = m % hop; j < win; j += hop) wss[m] += w[j] * w[j];
  for (int m = 0; m < win; ++m) {
    wf[m] = static_cast<float>(w[m]);
    wout[m] = static_cast<float>(w[m] / N);
    wsq[m] = static_cast<float>(w[m] * w[m]);
    wnorm[m] = static_cast<float>(wss[m] > 1.1754943508222875e-38 ? w[m] / (N * wss[m]) : w[m] / N);
  }
  p->window_f32 = wf;
  // edge-frame synthesis scales (librosa.istft divides by the window-sum-square of the frames that exist)
  const int nov = (win - 1) / hop;
  d.nov = nov;
  std::vector<float> wedge(static_cast<size_t>(2) * std::max(nov, 1) * win, 0.f);
  for (int side = 0; side < 2; ++side)
    for (int t = 0; t < nov; ++t)
      for (int m = 0; m < win; ++m) {
        float acc = 0.f;   // same float accumulation order as synth_scale_edge (gl.cuh)
        for (int dd = -nov; dd <= nov; ++dd) {
          const int off = m - dd * hop;
          const bool exists = side == 0 ? (t + dd >= 0) : (dd <= t);
          if (exists && off >= 0 && off < win) acc += wsq[off];
        }
        const float wv = wf[m] / N;
        wedge[(static_cast<size_t>(side) * nov + t) * win + m] = acc > 1.1754943508222875e-38f ? wv / acc : wv;
      }
  std::vector<float2> tw(static_cast<size_t>(R2 - 1) * 32), ws(Nz / 2 + 2);
  for (int k1 = 1; k1 < R2; ++k1)
    for (int l = 0; l < 32; ++l) {
      const double th = 2.0 * pi * ((k1 * l) % Nz) / Nz;
      tw[(k1 - 1) * 32 + l] = make_float2(static_cast<float>(std::cos(th)), static_cast<float>(-std::sin(th)));
    }
  for (int k = 0; k <= Nz / 2; ++k) {
    const double th = 2.0 * pi * k / N;   // -0.5i (cos - i sin) = -0.5 sin - 0.5i cos
    ws[k] = make_float2(static_cast<float>(-0.5 * std::sin(th)), static_cast<float>(-0.5 * std::cos(th)));
  }
  ws[Nz / 2 + 1] = make_float2(0.f, 0.f);
  std::vector<float2> sp2(17 * 32);
  for (int s = 0; s < 17; ++s)
    for (int l = 0; l < 32; ++l) {
      const int k = (l % R2) + R2 * s;
      const double ph = 2.0 * pi * k / N;
      sp2[s * 32 + l] = s < 8 ? make_float2(static_cast<float>(std::tan(ph)), static_cast<float>(std::cos(ph)))
                              : make_float2(static_cast<float>(std::cos(ph) / std::sin(ph)), static_cast<float>(std::sin(ph)));
    }

  std::vector<float2> spn(Nz / 2 + 2, make_float2(0.f, 0.f));
  for (int k = 0; k <= Nz / 2; ++k) {
    const double ph = 2.0 * pi * k / N;
    spn[k] = k < Nz / 4 ? make_float2(static_cast<float>(std::tan(ph)), static_cast<float>(std::cos(ph)))
                        : make_float2(static_cast<float>(std::cos(ph) / std::sin(ph)), static_cast<float>(std::sin(ph)));
  }
  p->mel_dense = mel_filterbank(c.sample_rate, N, c.n_mel, c.fmin, c.fmax, c.mel_htk != 0);
  const std::vector<float>& mb = p->mel_dense;
  const int rounds = (c.n_mel + 31) / 32;
  d.mel_rounds = rounds;
  std::vector<int> lo(c.n_mel, 0), len(c.n_mel, 0);
  for (int m = 0; m < c.n_mel; ++m) {
    int first = -1, last = -1;
    for (int k = 0; k < F; ++k)
      if (mb[static_cast<size_t>(m) * F + k] != 0.f) { if (first < 0) first = k; last = k; }
    if (first >= 0) { lo[m] = first; len[m] = last - first + 1; }
    if (last >= F - 1) return "mel filter reaches the Nyquist bin (requires fmax < sample_rate/2)";
  }
  // Rows are dealt to (round, lane) slots longest first, so a round's 32 lanes have similar lengths; inside a round the
  // lanes of each half-warp get distinct start residues mod 16 (the 64-bit magnitude loads of a half-warp then hit 16
  // different bank pairs) by starting a row up to (round length - row length) bins early with zero weights: bipartite
  // matching rows -> (half, residue) slots.
  std::vector<int> order(c.n_mel);
  for (int m = 0; m < c.n_mel; ++m) order[m] = m;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return len[a] > len[b]; });
  std::vector<int> slot_row(32 * rounds, -1), slot_lo(32 * rounds, 0);
  std::vector<float> melw;
  for (int r = 0; r < kMaxMelRounds; ++r) { d.mel_round_off[r] = 0; d.mel_round_len[r] = 0; }
  for (int r = 0; r < rounds; ++r) {
    const int nrows = std::min(32, c.n_mel - 32 * r);
    const int* rows = order.data() + 32 * r;
    int mx = 0;
    for (int i = 0; i < nrows; ++i) mx = std::max(mx, len[rows[i]]);
    std::vector<int> owner(32, -1), shift(nrows, -1);   // slot = half * 16 + residue
    std::function<bool(int, std::vector<char>&)> place = [&](int i, std::vector<char>& seen) -> bool {
      const int m = rows[i];
      for (int dsh = 0; dsh <= std::min(mx - len[m], lo[m]); ++dsh)
        for (int h = 0; h < 2; ++h) {
          const int sl = h * 16 + ((lo[m] - dsh) & 15);
          if (seen[sl]) continue;
          seen[sl] = 1;
          if (owner[sl] < 0 || place(owner[sl], seen)) {
            owner[sl] = i;
            shift[i] = dsh;
            return true;
          }
        }
      return false;
    };
    std::vector<int> lane_of(nrows, -1);
    for (int i = 0; i < nrows; ++i) {
      std::vector<char> seen(32, 0);
      place(i, seen);
    }
    std::vector<char> lane_used(32, 0);
    std::vector<int> half_fill(2, 0);
    for (int sl = 0; sl < 32; ++sl)
      if (owner[sl] >= 0) {
        const int h = sl / 16, lane = h * 16 + half_fill[h]++;
        lane_of[owner[sl]] = lane;
        lane_used[lane] = 1;
      }
    for (int i = 0; i < nrows; ++i)
      if (lane_of[i] < 0) {   // unmatched (bank conflicts accepted): any free lane, no shift
        int lane = 0;
        while (lane_used[lane]) ++lane;
        lane_used[lane] = 1;
        lane_of[i] = lane;
        shift[i] = 0;
      }
    d.mel_round_off[r] = static_cast<int>(melw.size());
    const int mx8 = (mx + 7) / 8 * 8;   // kernels walk a round in fully unrolled chunks of 8 taps (zero weights pad)
    d.mel_round_len[r] = mx8;
    melw.resize(melw.size() + static_cast<size_t>(mx8) * 32, 0.f);
    for (int i = 0; i < nrows; ++i) {
      const int m = rows[i], lane = lane_of[i], start = lo[m] - shift[i];
      slot_row[32 * r + lane] = m;
      slot_lo[32 * r + lane] = start;
      for (int it = 0; it < len[m]; ++it)
        melw[d.mel_round_off[r] + static_cast<size_t>(it + shift[i]) * 32 + lane] = mb[static_cast<size_t>(m) * F + lo[m] + it];
    }
  }
  // packed slot table: low 16 bits = first bin read, high bits = mel row (0x7fff: empty slot)
  std::vector<int> slots(32 * rounds);
  for (int i = 0; i < 32 * rounds; ++i) slots[i] = slot_lo[i] | ((slot_row[i] < 0 ? 0x7fff : slot_row[i]) << 16);
  d.melw_count = static_cast<int>(melw.size());
  d.mel_kmin = F;
  d.mel_kmax = -1;
  for (int m = 0; m < c.n_mel; ++m)
    if (len[m] > 0) {
      d.mel_kmin = std::min(d.mel_kmin, lo[m]);
      d.mel_kmax = std::max(d.mel_kmax, lo[m] + len[m] - 1);
    }
  // ---- segment view (see PlanDev) ----
  const int n_seg = c.n_mel + 1;
  const int seg_rounds = (n_seg + 31) / 32;
  d.seg_rounds = seg_rounds;
  std::vector<int4> seg_slot(static_cast<size_t>(32) * seg_rounds, make_int4(0, 0, 0, 0));
  std::vector<float2> seg_coef(static_cast<size_t>(32) * seg_rounds, make_float2(0.f, 0.f));
  std::vector<float> enorm(c.n_mel, 0.f);
  for (int r = 0; r < kMaxSegRounds; ++r) d.seg_round_len[r] = 0;
  {
    const bool htk = c.mel_htk != 0;
    std::vector<double> cf(c.n_mel + 2);
    const double m0 = hz_to_mel(c.fmin, htk), m1 = hz_to_mel(c.fmax, htk), step = (m1 - m0) / (c.n_mel + 1);
    for (int i = 0; i < c.n_mel + 2; ++i) cf[i] = mel_to_hz(i == c.n_mel + 1 ? m1 : m0 + i * step, htk);
    const double df = (c.sample_rate / 2.0) / (F - 1);
    auto first_bin_at = [&](double hz) {   // first k with f_k >= hz
      int k = static_cast<int>(std::floor(hz / df)) - 1;
      if (k < 0) k = 0;
      while (k < F && k * df < hz) ++k;
      return k;
    };
    std::vector<int> sk0(n_seg), slen(n_seg);
    for (int j = 0; j < n_seg; ++j) {
      sk0[j] = first_bin_at(cf[j]);
      slen[j] = std::max(0, first_bin_at(cf[j + 1]) - sk0[j]);
    }
    for (int m = 0; m < c.n_mel; ++m) enorm[m] = static_cast<float>(2.0 / (cf[m + 2] - cf[m]));
    for (int r = 0; r < seg_rounds; ++r) {
      int k0r[32], lenr[32], dsh[32];
      int mx = 0;
      for (int l = 0; l < 32; ++l) {
        const int j = 32 * r + l;
        k0r[l] = j < n_seg ? sk0[j] : 64 + l;   // empty slots: masked entirely, any bank will do
        lenr[l] = j < n_seg ? slen[j] : 0;
        mx = std::max(mx, lenr[l]);
      }
      // smallest round length L for which both half-warps have a perfect matching lane -> residue (k0 - d) mod 16, d <= L - len
      int L = std::max(mx, 1);
      for (;; ++L) {
        bool ok = true;
        for (int h = 0; h < 2 && ok; ++h) {
          int owner[16];
          for (int q = 0; q < 16; ++q) owner[q] = -1;
          std::function<bool(int, std::vector<char>&)> place = [&](int l, std::vector<char>& seen) -> bool {
            for (int dd = 0; dd <= std::min(L - lenr[l], k0r[l]); ++dd) {
              const int q = (k0r[l] - dd) & 15;
              if (seen[q]) continue;
              seen[q] = 1;
              if (owner[q] < 0 || place(owner[q], seen)) {
                owner[q] = l;
                dsh[l] = dd;
                return true;
              }
            }
            return false;
          };
          for (int l = 16 * h; l < 16 * h + 16 && ok; ++l) {
            std::vector<char> seen(16, 0);
            ok = place(l, seen);
          }
        }
        if (ok || L > mx + 32) break;   // (L > mx + 32 cannot happen: 16 shifts reach every residue)
      }
      d.seg_round_len[r] = (L + 3) / 4 * 4;   // the kernel walks a round in chunks of 4 trips
      for (int l = 0; l < 32; ++l) {
        const int j = 32 * r + l;
        seg_slot[32 * r + l] = make_int4(k0r[l] - dsh[l], dsh[l], lenr[l], j < n_seg ? 1 : 0);
        if (j < n_seg) {
          const double delta = cf[j + 1] - cf[j];
          seg_coef[32 * r + l] = make_float2(static_cast<float>((sk0[j] * df - cf[j]) / delta), static_cast<float>(df / delta));
        }
      }
    }
  }
  std::vector<int> r0(F, 0);
  std::vector<float> c0(F, 0.f), c1(F, 0.f);
  for (int k = 0; k < F; ++k) {
    int first = -1, cnt = 0;
    for (int m = 0; m < c.n_mel; ++m)
      if (mb[static_cast<size_t>(m) * F + k] != 0.f) { if (first < 0) first = m; ++cnt; }
    if (cnt > 2) return "mel filterbank has more than two non-zeros in a column";
    if (first >= 0) {
      r0[k] = std::min(first, c.n_mel - 2 < 0 ? 0 : c.n_mel - 2);
      c0[k] = mb[static_cast<size_t>(r0[k]) * F + k];
      c1[k] = (r0[k] + 1 < c.n_mel) ? mb[static_cast<size_t>(r0[k] + 1) * F + k] : 0.f;
    }
  }
  std::vector<float2> c01((F + 1) & ~1, make_float2(0.f, 0.f));
  std::vector<unsigned char> r8((F + 15) & ~15, 0);
  for (int k = 0; k < F; ++k) {
    c01[k] = make_float2(c0[k], c1[k]);
    r8[k] = static_cast<unsigned char>(r0[k]);
  }
  // _get_linear_basis: p = m m^T (float32 like np.matmul on the float32 basis), d = 1 / column sums where |x| > 1e-8
  std::vector<float> lc0(F, 0.f), lc1(F, 0.f);
  {
    std::vector<double> dinv(c.n_mel, 0.0);
    for (int j = 0; j < c.n_mel; ++j) {
      float colsum = 0.f;
      for (int i = 0; i < c.n_mel; ++i) {
        float pij = 0.f;
        for (int k = 0; k < F; ++k) pij += mb[static_cast<size_t>(i) * F + k] * mb[static_cast<size_t>(j) * F + k];
        colsum += pij;
      }
      dinv[j] = std::fabs(colsum) > 1.0e-8f ? 1.0 / static_cast<double>(colsum) : static_cast<double>(colsum);
    }
    for (int k = 0; k < F; ++k) {
      lc0[k] = static_cast<float>(static_cast<double>(c0[k]) * dinv[r0[k]]);
      lc1[k] = (r0[k] + 1 < c.n_mel) ? static_cast<float>(static_cast<double>(c1[k]) * dinv[r0[k] + 1]) : 0.f;
    }
  }
  *st = SB200_ERR_CUDA;
  cudaError_t e;
#define SB200_UP(vec, field) \
  if ((e = upload(p, vec, &d.field)) != cudaSuccess) return std::string("cuda upload: ") + cudaGetErrorString(e);
  SB200_UP(wf, window) SB200_UP(wout, wout) SB200_UP(wnorm, wnorm) SB200_UP(wsq, wsq) SB200_UP(wedge, wedge) SB200_UP(tw, tw) SB200_UP(ws, ws) SB200_UP(sp2, sp2) SB200_UP(spn, spn)
  SB200_UP(seg_slot, seg_slot) SB200_UP(seg_coef, seg_coef) SB200_UP(enorm, mel_enorm)
  SB200_UP(melw, melw) SB200_UP(slots, mel_lo) SB200_UP(r0, col_r0) SB200_UP(c0, col_c0) SB200_UP(c1, col_c1) SB200_UP(lc0, lin_c0) SB200_UP(lc1, lin_c1) SB200_UP(c01, col_c01) SB200_UP(r8, col_r8)
#undef SB200_UP
  *st = SB200_OK;
  return "";
}

}  // namespace sb200
