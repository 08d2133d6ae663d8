// ISTFT and fused Griffin-Lim iteration on the packed FFT engine (fft2.cuh).
//   librosa.istft      (transtacos/audio.py:147-148)                                  -> gl2_kernel<N, 0> + gl2_finish_kernel
//   _griffin_lim       (transtacos/audio.py:130-140, angle form)                      -> gl2_kernel<N, 1> (init), <N, 2> (iteration)
//   librosa.griffinlim (retunegan/audio.py:131-136, fast form)         -> gl2_kernel<N, 1> (init), <N, 3> (iteration)
//
// A CTA of 8 warps owns a TILE of 8 items = 8 * 4096/n_fft consecutive frames of one utterance.  One iteration is one
// launch and, per tile:
//   1. the CTA stages the time signal under the tile (reflect padded, np.pad mode='reflect') in shared memory,
//   2. every warp analyses its frame pairs (packed forward FFT), applies the phase update to the Hermitian pairs in
//      registers (X/|X|, or librosa's c/(|c|+1e-16) with c = X - alpha*tprev), and synthesises them (packed inverse
//      FFT, synthesis window / (n_fft * window-sum-square)) into its shared-memory buffer,
//   3. the CTA overlap-adds the tile's frames from shared memory (a gather: deterministic, no atomics) and writes the
//      tile's span of the new signal.
// Consecutive tiles overlap by win - hop samples.  The signal therefore lives in TWO buffers by tile parity: a tile
// writes its whole span into the buffer of its parity and zeros into the interior of the other one, so every element
// of both buffers has exactly one writer per launch and the signal is ya[u] + yb[u].  Per iteration the kernel
// touches S (4 B/bin), tprev (16 B/bin, fast form only) and 16 B per signal sample: the streaming minimum of SURVEY.md 8d
// plus one extra signal write; the spectrum itself never leaves the registers.
#pragma once
#include "fft2.cuh"
#include "gl.cuh"

namespace sb200 {

constexpr int kGl2Warps = 8;        // warps per CTA
constexpr int kGl2GroupWarps = 4;   // warps per tile group: a CTA runs kGl2Warps / kGl2GroupWarps tiles concurrently, each
                                    // group synchronising on its own named barrier so their phases interleave
constexpr int kGl2Groups = kGl2Warps / kGl2GroupWarps;
#ifdef kGl2TprevAheadOverride
constexpr bool kGl2TprevAhead = kGl2TprevAheadOverride;
#else
constexpr bool kGl2TprevAhead = true;   // fast form: fetch the previous spectrum one slot ahead (more registers) or at its use
#endif

__device__ __forceinline__ void gl2_group_sync(int group) {
  asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kGl2GroupWarps * 32) : "memory");
}

struct Gl2Args {
  GlBatch g;
  int tiles_per_row;         // ceil(max frames of a row / frames per tile)
  const float* S;            // [frames, F] magnitudes (modes 1-3)
  const float* init_phase;   // [frames, F] u in [0,1) (mode 1)
  const float2* spec;        // [frames, F] complex (mode 0)
  const float* ya_in;        // signal written by the previous launch (modes 2, 3)
  const float* yb_in;
  float* ya_out;
  float* yb_out;
  float2* tprev;             // [frames, F] (mode 3): previous rebuilt spectrum in the engine's internal (rotated) form
  float alpha;               // momentum / (1 + momentum)
  int first;                 // mode 3: tprev not yet written (rebuilt = 0)
};

// first element of utterance b in the signal buffers; an utterance owns (T - 1) * hop + win elements
__device__ __forceinline__ long long gl2_sig_base(const GlRow& row, int b, int hop, int win) {
  return row.frame_base * hop + static_cast<long long>(b) * win;
}

template <int N>
struct Gl2Smem {
  using C = Fft2Cfg<N>;
  uint4* xbufs;    // [warps][kXElems]
  float* win;      // [win] 0.5 * analysis window
  float* wnorm;    // [win] synthesis window / (N * interior window-sum-square)
  float2* tw;      // [kTwCount]
  float2* sp2;     // [17*32]
  float* ytile;    // [groups][span]
  __host__ __device__ static size_t bytes(int span) {
    return static_cast<size_t>(kGl2Warps) * C::kXBytes + sizeof(float) * 2 * C::kWin + sizeof(float2) * (C::kTwCount + 17 * 32) +
           sizeof(float) * span * kGl2Groups;
  }
  __device__ __forceinline__ void init(unsigned char* raw, const PlanDev& p) {
    xbufs = reinterpret_cast<uint4*>(raw);
    win = reinterpret_cast<float*>(raw + static_cast<size_t>(kGl2Warps) * C::kXBytes);
    wnorm = win + C::kWin;
    tw = reinterpret_cast<float2*>(wnorm + C::kWin);
    sp2 = tw + C::kTwCount;
    ytile = reinterpret_cast<float*>(sp2 + 17 * 32);
    for (int i = threadIdx.x; i < C::kWin; i += blockDim.x) {
      win[i] = 0.5f * p.window[i];
      wnorm[i] = p.wnorm[i];
    }
    for (int i = threadIdx.x; i < C::kTwCount; i += blockDim.x) tw[i] = p.tw[i];
    for (int i = threadIdx.x; i < 17 * 32; i += blockDim.x) sp2[i] = p.sp2[i];
    __syncthreads();
  }
};

// i^j * (x, y)
__device__ __forceinline__ float2 rot_i(float2 a, int j) {
  j &= 3;
  const float x = (j & 1) ? a.y : a.x, y = (j & 1) ? a.x : a.y;
  return make_float2((j == 1 || j == 2) ? -x : x, (j >= 2) ? -y : y);
}

// Edge-frame synthesis weights for an utterance too short for the plan's head / tail tables (rolled, out of line).
__device__ __noinline__ void gl2_edge_weights(const PlanDev& p, int t, int n_frames, float* out, int lane) {
#pragma unroll 1
  for (int m = lane; m < p.win; m += 32) out[m] = synth_scale_edge(p, t, n_frames, m);
}

template <int N, int MODE>
__global__ void __launch_bounds__(kGl2Warps * 32, 1) gl2_kernel(const PlanDev p, const Gl2Args a) {
  using C = Fft2Cfg<N>;
  constexpr int FT = kGl2GroupWarps * C::kFrames;   // frames per tile
  constexpr int GT = kGl2GroupWarps * 32;           // threads per group
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Gl2Smem<N> sm;
  sm.init(smem_raw, p);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int group = warp / kGl2GroupWarps, wg = warp % kGl2GroupWarps, gt = threadIdx.x % GT;   // tile group, warp / thread in it
  const int hop = p.hop;
  float* const ytile = sm.ytile + group * ((FT - 1) * hop + C::kWin);
  uint4* xbuf = sm.xbufs + warp * C::kXElems;
  float* stage = reinterpret_cast<float*>(xbuf);   // [kFrames][win] synthesised frames of this warp (aliases the exchange buffer)
  const int k1 = lane & (C::kR2 - 1), pl = lane / C::kR2;   // spectral role of this lane: column k1 of pair pl
  const bool col0 = (k1 == 0);
  const int partner = (lane & ~(C::kR2 - 1)) | ((C::kR2 - k1) & (C::kR2 - 1));
  const int rk = k1 & 3, rm = (4 - rk) & 3;                 // bin index mod 4 of the a side (k1 + R2 s) and b side (Nz - k)
  const float2* const sp = sm.sp2 + lane;
  const long long n_tiles = static_cast<long long>(a.g.bd.B) * a.tiles_per_row;
  for (long long vt = static_cast<long long>(blockIdx.x) * kGl2Groups + group; vt < n_tiles;
       vt += static_cast<long long>(gridDim.x) * kGl2Groups) {
    const int b = static_cast<int>(vt / a.tiles_per_row), tk = static_cast<int>(vt - static_cast<long long>(b) * a.tiles_per_row);
    const GlRow row = gl_row(a.g, b, N, hop);
    const int tile_t0 = tk * FT;
    if (tile_t0 >= row.T) continue;   // uniform over the group
    const long long sbase = gl2_sig_base(row, b, hop, C::kWin);
    const long long cover = static_cast<long long>(row.T - 1) * hop + C::kWin;   // elements of this utterance in the signal buffers
    const long long u0 = static_cast<long long>(tile_t0) * hop;
    // ---- 1. signal under the tile ------------------------------------------------------------------------------
    if constexpr (MODE >= 2) {
      const int nf = min(FT, row.T - tile_t0);
      const int span = (nf - 1) * hop + C::kWin;
      const float* ya = a.ya_in + sbase;
      const float* yb = a.yb_in + sbase;
      const long long i_lo = u0 - N / 4, i_hi = u0 + span - 1 - N / 4;   // signal indices under the tile
      if (i_lo >= 0 && i_hi < row.Ly && u0 + span <= cover && (hop & 3) == 0) {
        // no reflection: offset coordinate u0 + j maps to itself; 16-byte loads (sbase, u0 and span are multiples of 4)
        const float4* ya4 = reinterpret_cast<const float4*>(ya + u0);
        const float4* yb4 = reinterpret_cast<const float4*>(yb + u0);
        float4* yt4 = reinterpret_cast<float4*>(ytile);
#pragma unroll 5
        for (int j = gt; j < span / 4; j += GT) {
          const float4 p4 = __ldg(ya4 + j), q4 = __ldg(yb4 + j);
          yt4[j] = make_float4(p4.x + q4.x, p4.y + q4.y, p4.z + q4.z, p4.w + q4.w);
        }
      } else {
        const int Ly = static_cast<int>(row.Ly), cov = static_cast<int>(cover), base = static_cast<int>(u0) - N / 4;
#pragma unroll 4
        for (int j = gt; j < span; j += GT) {
          int i = base + j;                            // signal index of offset coordinate u0 + j
          i = i < 0 ? -i : i;
          i = i >= Ly ? 2 * (Ly - 1) - i : i;
          const int uu = min(i + N / 4, cov - 1);
          const float val = __ldg(ya + uu) + __ldg(yb + uu);
          ytile[j] = (i + N / 4 < cov) ? val : 0.f;
        }
      }
      gl2_group_sync(group);
    }
    // ---- 2. per warp: analysis, phase update, synthesis ---------------------------------------------------------
    {
      const int item_t0 = tile_t0 + wg * C::kFrames;
      if constexpr (MODE >= 2) {
        // the per-bin state of this item's frames is consecutive in memory: pull it into L2 while the analysis FFT runs
        const long long r0 = (row.frame_base + min(item_t0, row.T - 1)) * C::kF;
        const int nfr = max(0, min(C::kFrames, row.T - item_t0));
        const char* sp0 = reinterpret_cast<const char*>(a.S + r0);
        for (int o = lane * 128; o < nfr * C::kF * 4; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp0 + o));
        if constexpr (MODE == 3) {
          if (!a.first) {
            const char* tp0 = reinterpret_cast<const char*>(a.tprev + r0);
            for (int o = lane * 128; o < nfr * C::kF * 8; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(tp0 + o));
          }
        }
      }
      PC v[32];
      if constexpr (MODE >= 2) {
        static_for<0, C::kP>([&](auto pc_) {
          constexpr int pp = decltype(pc_)::value;
          const int tA = item_t0 + 2 * pp;
          const bool okA = tA < row.T, okB = tA + 1 < row.T;
          const float* yA = ytile + (tA - tile_t0) * hop + 2 * lane;
          static_for<0, C::kR>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            const float2 w = *reinterpret_cast<const float2*>(sm.win + 2 * lane + 64 * r);
            float2 sa = make_float2(0.f, 0.f), sb = make_float2(0.f, 0.f);
            if (okA) sa = *reinterpret_cast<const float2*>(yA + 64 * r);
            if (okB) sb = *reinterpret_cast<const float2*>(yA + hop + 64 * r);
            constexpr int idx = pp * C::kR2 + brev(r, C::kLogR2);
            v[idx].re = pk(sa.x * w.x, sb.x * w.x);
            v[idx].im = pk(sa.y * w.y, sb.y * w.y);
            v[idx + 1] = v[idx];
          });
        });
        fft2_forward<N>(v, xbuf, sm.tw, lane);
      }
      // lane (pl, k1) handles frames fA = item_t0 + 2 pl (low halves) and fB = fA + 1 (high halves)
      const int fA = item_t0 + 2 * pl;
      const bool okA = fA < row.T, okB = fA + 1 < row.T;
      const long long rowA = (row.frame_base + fA) * C::kF;   // frame B: + F
      // spectral value of bin kb of both frames in the engine's internal form (modes 0 / 1): i^kb X, conjugated on the b side
      auto fetch = [&](int kb, int rot, bool conj_it) -> PC {
        float2 xa = make_float2(0.f, 0.f), xb = make_float2(0.f, 0.f);
        if constexpr (MODE == 0) {
          if (okA) xa = __ldg(a.spec + rowA + kb);
          if (okB) xb = __ldg(a.spec + rowA + C::kF + kb);
        } else {
          if (okA) {
            float s, c;
            sincospif(2.f * __ldg(a.init_phase + rowA + kb), &s, &c);
            const float mg = __ldg(a.S + rowA + kb);
            xa = make_float2(mg * c, mg * s);
          }
          if (okB) {
            float s, c;
            sincospif(2.f * __ldg(a.init_phase + rowA + C::kF + kb), &s, &c);
            const float mg = __ldg(a.S + rowA + C::kF + kb);
            xb = make_float2(mg * c, mg * s);
          }
        }
        xa = rot_i(xa, rot);
        xb = rot_i(xb, rot);
        PC r;
        r.re = pk(xa.x, xb.x);
        r.im = conj_it ? pk(-xa.y, -xb.y) : pk(xa.y, xb.y);
        return r;
      };
      // ---- modes 2 / 3: per-bin state of the phase update, fetched one slot ahead of its use -------------------------
      // Loads are unconditional (frames past the end are clamped to the last one and masked), addresses are one base
      // register per (side, frame) plus a compile-time offset.
      const long long rowAc = (row.frame_base + min(fA, row.T - 1)) * C::kF, rowBc = (row.frame_base + min(fA + 1, row.T - 1)) * C::kF;
      const float mA = okA ? 1.f : 0.f, mB = okB ? 1.f : 0.f;
      // element offsets of (side, frame); S and tprev share them (the bases are kernel parameters: no registers)
      const long long oaA = rowAc + k1, oaB = rowBc + k1;                       // a side, bin k1 + R2 s: + R2 s
      const long long obA = rowAc + C::kNz - k1, obB = rowBc + C::kNz - k1;     // b side, bin Nz - k1 - R2 s: - R2 s
      struct SlotLd {
        float saA, saB, sbA, sbB;
        float2 taA, taB, tbA, tbB;
      };
      auto load_slot = [&](auto sc) -> SlotLd {
        constexpr int off = C::kR2 * decltype(sc)::value;
        SlotLd L;
        L.saA = __ldg(a.S + oaA + off) * mA;
        L.saB = __ldg(a.S + oaB + off) * mB;
        L.sbA = __ldg(a.S + obA - off) * mA;
        L.sbB = __ldg(a.S + obB - off) * mB;
        if constexpr (MODE == 3 && kGl2TprevAhead) {
          if (!a.first) {
            L.taA = a.tprev[oaA + off];
            L.taB = a.tprev[oaB + off];
            L.tbA = a.tprev[obA - off];
            L.tbB = a.tprev[obB - off];
          }
        }
        return L;
      };
      // phase update of one held value X of both frames: magnitudes sA / sB, previous spectrum tA / tB, stored to dA / dB
      auto update = [&](const PC& X, float sA, float sB, float2 tA, float2 tB, float2* dA, float2* dB, bool on, int rot,
                        bool conj_held) -> PC {
        PC o;
        if constexpr (MODE == 2) {
          const pf n2 = norm2(X);
          const float nA = plo(n2), nB = phi(n2);
          const float iA = nA > 0.f ? rsqrtf(nA) : 0.f, iB = nB > 0.f ? rsqrtf(nB) : 0.f;
          const pf sc = pk(sA * iA, sB * iB);
          o.re = mul2(X.re, sc);
          o.im = mul2(X.im, sc);
          if (nA == 0.f || nB == 0.f) {   // exp(1j * angle(0)) = 1 (transtacos/audio.py:138), in the internal form: i^kb (conjugated on the b side)
            float2 un = rot_i(make_float2(1.f, 0.f), rot);
            if (conj_held) un.y = -un.y;
            o.re = pk(nA > 0.f ? plo(o.re) : sA * un.x, nB > 0.f ? phi(o.re) : sB * un.x);
            o.im = pk(nA > 0.f ? plo(o.im) : sA * un.y, nB > 0.f ? phi(o.im) : sB * un.y);
          }
        } else {
          PC c = X;
          if (!a.first) {
            c.re = fma2s(pk(tA.x, tB.x), -a.alpha, X.re);
            c.im = fma2s(pk(tA.y, tB.y), -a.alpha, X.im);
          }
          if (on && okA) *dA = make_float2(plo(X.re), plo(X.im));
          if (on && okB) *dB = make_float2(phi(X.re), phi(X.im));
          const pf n2 = norm2(c);
          const pf sc = pk(sA / (sqrtf(plo(n2)) + 1e-16f), sB / (sqrtf(phi(n2)) + 1e-16f));
          o.re = mul2(c.re, sc);
          o.im = mul2(c.im, sc);
        }
        return o;
      };
      PC zself;
      if constexpr (MODE >= 2) {
        // self pair of column 0 (bin Nz/2) first: the exchange below overwrites v[16]
        PC ak, am;
        split2<true>(v[16], v[16], sp[16 * 32], ak, am);
        {
          const float sA = __ldg(a.S + rowAc + C::kNz / 2) * mA, sB = __ldg(a.S + rowBc + C::kNz / 2) * mB;
          float2 tA = make_float2(0.f, 0.f), tB = tA;
          if constexpr (MODE == 3) {
            if (!a.first) {
              tA = a.tprev[rowAc + C::kNz / 2];
              tB = a.tprev[rowBc + C::kNz / 2];
            }
          }
          ak = update(ak, sA, sB, tA, tB, a.tprev + rowAc + C::kNz / 2, a.tprev + rowBc + C::kNz / 2, col0, 0, false);
        }
        PC q;
        q.re = ak.re;
        q.im = sub2(0ull, ak.im);
        PC zr;
        split2_inv<true>(ak, q, sp[16 * 32], zself, zr);
        // partner exchange: slot s receives Z[Nz - k] into v[31 - s]
        static_for<0, 16>([&](auto sc) {
          constexpr int s = 15 - decltype(sc)::value;
          const PC send = pc_sel(col0, v[(32 - s) & 31], v[31 - s]);
          v[31 - s] = pc_shfl(send, partner);
        });
      } else {
        const PC pS = fetch(C::kNz / 2, 0, false);
        PC q;
        q.re = pS.re;
        q.im = sub2(0ull, pS.im);
        PC zr;
        split2_inv<true>(pS, q, sp[16 * 32], zself, zr);
      }
      SlotLd nxt{};
      if constexpr (MODE >= 2) nxt = load_slot(std::integral_constant<int, 0>{});
      static_for<0, 16>([&](auto sc) {
        constexpr int s = decltype(sc)::value;
        PC P, Q;
        if constexpr (MODE >= 2) {
          SlotLd cur = nxt;
          if constexpr (MODE == 3 && !kGl2TprevAhead) {   // previous spectrum fetched at its use (it was pulled into L2 at item start)
            if (!a.first) {
              cur.taA = a.tprev[oaA + C::kR2 * s];
              cur.taB = a.tprev[oaB + C::kR2 * s];
              cur.tbA = a.tprev[obA - C::kR2 * s];
              cur.tbB = a.tprev[obB - C::kR2 * s];
            }
          }
          if constexpr (s < 15) nxt = load_slot(std::integral_constant<int, s + 1>{});
          split2<(s >= 8)>(v[s], v[31 - s], sp[s * 32], P, Q);
          P = update(P, cur.saA, cur.saB, cur.taA, cur.taB, a.tprev + oaA + C::kR2 * s, a.tprev + oaB + C::kR2 * s, true, rk, false);
          Q = update(Q, cur.sbA, cur.sbB, cur.tbA, cur.tbB, a.tprev + obA - C::kR2 * s, a.tprev + obB - C::kR2 * s, true, rm, true);
        } else {
          const int ka = k1 + C::kR2 * s, kb = C::kNz - ka;
          P = fetch(ka, rk, false);
          Q = fetch(kb, rm, true);
        }
        if constexpr (s == 0) {
          if (col0) {   // irfft ignores the imaginary parts of DC and Nyquist
            P.im = 0ull;
            Q.im = 0ull;
          }
        }
        split2_inv<(s >= 8)>(P, Q, sp[s * 32], v[s], v[31 - s]);
      });
      // reverse exchange: v[31 - s] <- Z'[Nz - k] computed by the partner (column 0: by this lane's slot s + 1, bin Nz/2: self pair)
      static_for<0, 16>([&](auto sc) {
        constexpr int s = decltype(sc)::value;
        PC send;
        if constexpr (s < 15) send = pc_sel(col0, v[30 - s], v[31 - s]);
        else send = pc_sel(col0, zself, v[16]);
        v[31 - s] = pc_shfl(send, partner);
      });
      fft2_inverse<N>(v, xbuf, sm.tw, lane);
      // synthesis window and window-sum-square normaliser (librosa.istft), frames into this warp's staging buffer
      static_for<0, C::kP>([&](auto pc_) {
        constexpr int pp = decltype(pc_)::value;
        const int tA = item_t0 + 2 * pp;
        static_for<0, 2>([&](auto hc) {
          constexpr int h = decltype(hc)::value;
          const int t = tA + h;
          float* dst = stage + (2 * pp + h) * C::kWin + 2 * lane;
          const bool live = t < row.n_frames;
          const float* wt = sm.wnorm;   // interior frames
          if (live) {
            const bool head = t < p.nov, tail = t > row.n_frames - 1 - p.nov;
            if (head && tail) {
              // utterance shorter than 2 nov + 1 frames (rare): weights computed by a rolled loop into the free upper part of the buffer
              float* wtmp = stage + C::kFrames * C::kWin;
              gl2_edge_weights(p, t, row.n_frames, wtmp, lane);
              __syncwarp();
              wt = wtmp;
            } else if (head) {
              wt = p.wedge + static_cast<long long>(t) * C::kWin;
            } else if (tail) {
              wt = p.wedge + static_cast<long long>(p.nov + row.n_frames - 1 - t) * C::kWin;
            }
          }
          const bool edge = wt != sm.wnorm;
          wt += 2 * lane;
          const float lv = live ? 1.f : 0.f;
          float2 w[C::kR];
          if (!edge) {   // interior frame: shared-memory table
            static_for<0, C::kR>([&](auto rc) {
              constexpr int r = decltype(rc)::value;
              w[r] = *reinterpret_cast<const float2*>(sm.wnorm + 2 * lane + 64 * r);
            });
          } else {       // head / tail frame: plan table in global memory, or the buffer filled above (generic loads)
            static_for<0, C::kR>([&](auto rc) {
              constexpr int r = decltype(rc)::value;
              w[r] = *reinterpret_cast<const float2*>(wt + 64 * r);
            });
          }
          static_for<0, C::kR>([&](auto rc) {
            constexpr int r = decltype(rc)::value;
            const PC z = v[pp * C::kR2 + r];
            const float zr = h ? phi(z.re) : plo(z.re), zi = h ? phi(z.im) : plo(z.im);
            *reinterpret_cast<float2*>(dst + 64 * r) = make_float2(zr * w[r].x * lv, zi * w[r].y * lv);
          });
          __syncwarp();   // wtmp may be rewritten by the next frame
        });
      });
    }
    gl2_group_sync(group);
    // ---- 3. overlap-add of the tile's frames, written as this tile's span of the signal ----------------------------
    {
      const bool last_tile = tile_t0 + FT >= row.T;
      const long long left = cover - u0;
      const int span_full = (FT - 1) * hop + C::kWin;
      const int span = static_cast<int>(min(static_cast<long long>(span_full), left));
      float* mine = ((tk & 1) ? a.yb_out : a.ya_out) + sbase + u0;
      float* other = ((tk & 1) ? a.ya_out : a.yb_out) + sbase + u0;
      const float* frames0 = reinterpret_cast<const float*>(sm.xbufs + group * kGl2GroupWarps * C::kXElems);
      constexpr int kWarpStride = C::kXBytes / 4;   // floats between the staging buffers of consecutive warps
      const int nov = p.nov;
      // thread owns offsets o within a hop; sample j = q * hop + o is covered by frames q - d at offset o + d * hop, d <= nov
      for (int o = gt; o < hop; o += GT) {
        for (int q = 0, j = o; j < span; ++q, j += hop) {
          float acc = 0.f;
#pragma unroll 4
          for (int d = 0; d <= nov; ++d) {
            const int f = q - d, off = o + d * hop;
            if (f >= 0 && f < FT && off < C::kWin) acc += frames0[(f / C::kFrames) * kWarpStride + (f % C::kFrames) * C::kWin + off];
          }
          mine[j] = acc;
          const bool interior = j >= C::kWin - hop && j < FT * hop;
          if (interior || (tk == 0 && j < C::kWin - hop) || (last_tile && j >= FT * hop)) other[j] = 0.f;
        }
      }
    }
    gl2_group_sync(group);
  }
}

// y[j] = ya[u] + yb[u], u = j + n_fft/4  (librosa.istft tail: trim n_fft/2, fix_length)
struct Gl2FinishArgs {
  GlBatch g;
  const float* ya;
  const float* yb;
  float* y;
};
template <int N>
__global__ void gl2_finish_kernel(const PlanDev p, const Gl2FinishArgs a) {
  const int b = blockIdx.y;
  const GlRow row = gl_row(a.g, b, N, p.hop);
  const long long sbase = gl2_sig_base(row, b, p.hop, N / 2);
  const long long cover = static_cast<long long>(row.T - 1) * p.hop + N / 2;
  for (long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; j < row.Ly;
       j += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long u = j + N / 4;
    a.y[row.out_base + j] = u < cover ? a.ya[sbase + u] + a.yb[sbase + u] : 0.f;
  }
}

}  // namespace sb200
