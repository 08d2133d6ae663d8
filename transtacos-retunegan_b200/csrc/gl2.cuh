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
// Unroll factor of the Hermitian-pair loop of the spectrum pass: the loop carries the next trip's operands (24 registers per pair)
// in a register rotation, which a rolled loop pays as ~50 MOVs per trip.  Batches (one launch per iteration, code warm after the
// first tile): 4 (1.105 -> 1.035 ms per 64 x 5 s x 4 iterations).  The persistent single-utterance kernel takes the same factor (0.410 -> 0.402 ms per 1 x 5 s x 30 iterations; 2 measured 0.399): with different factors ptxas contracts the scalar parts differently and a ragged batch is no longer bit-equal to single calls (tests).
#ifndef SB200_GL2_PAIR_UNROLL
#define SB200_GL2_PAIR_UNROLL 4
#endif
#ifndef SB200_GL2_PAIR_UNROLL_COH
#define SB200_GL2_PAIR_UNROLL_COH 4
#endif
constexpr int kGl2PairUnroll = SB200_GL2_PAIR_UNROLL, kGl2PairUnrollCoh = SB200_GL2_PAIR_UNROLL_COH;
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
  ulonglong2* tprev;         // (mode 3) previous rebuilt spectrum in the engine's internal (rotated) form, PACKED per frame pair the
                             // way the registers hold it: element (pair slot, bin) = {re of frames (2m, 2m+1), im of (2m, 2m+1)}, 16 bytes;
                             // pair slot of frames (2m, 2m+1) of utterance b = gl2_pair_base(frame_base, b) + m
  float alpha;               // momentum / (1 + momentum)
  int first;                 // mode 3: tprev not yet written (rebuilt = 0)
  int groups_active;         // tile groups of a CTA that take tiles (0 = all): 1 gives a small batch one tile per SM
};

// First pair slot of utterance b in the packed previous-spectrum array: utterances with an odd number of frames leave half a slot
// unused, so slot bases are rounded up from frame_base + 2 b + 1 (consecutive utterances never share a slot; at most
// (total_frames + 2 B + 1) / 2 + 1 slots in all).
__host__ __device__ __forceinline__ long long gl2_pair_base(long long frame_base, int b) { return (frame_base + 2LL * b + 1) >> 1; }
__host__ __device__ __forceinline__ long long gl2_pair_slots(long long total_frames, long long B) { return (total_frames + 2 * B + 1) / 2 + 1; }

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
  float2* spn;     // [Nz/2 + 1] split twiddles in natural bin order (padded to a multiple of 2)
  float* ytile;    // [groups][span]
  __host__ __device__ static size_t bytes(int span) {
    return static_cast<size_t>(kGl2Warps) * C::kXBytes + sizeof(float) * 2 * C::kWin + sizeof(float2) * (C::kTwCount + C::kNz / 2 + 2) +
           sizeof(float) * span * kGl2Groups;
  }
  __device__ __forceinline__ void init(unsigned char* raw, const PlanDev& p) {
    xbufs = reinterpret_cast<uint4*>(raw);
    win = reinterpret_cast<float*>(raw + static_cast<size_t>(kGl2Warps) * C::kXBytes);
    wnorm = win + C::kWin;
    tw = reinterpret_cast<float2*>(wnorm + C::kWin);
    spn = tw + C::kTwCount;
    ytile = reinterpret_cast<float*>(spn + C::kNz / 2 + 2);
    for (int i = threadIdx.x; i < C::kWin; i += blockDim.x) {
      win[i] = 0.5f * p.window[i];
      wnorm[i] = p.wnorm[i];
    }
    for (int i = threadIdx.x; i < C::kTwCount; i += blockDim.x) tw[i] = p.tw[i];
    for (int i = threadIdx.x; i < C::kNz / 2 + 1; i += blockDim.x) spn[i] = p.spn[i];
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

__device__ __forceinline__ float gl2_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Phase update of one held spectral value X of both frames of a pair (modes 2 / 3): magnitudes sA / sB, previous
// spectrum T (packed like X).  rot = bin index mod 4, conj_held = b side (value held conjugated): only used for the exact-zero case.
template <int MODE>
__device__ __forceinline__ PC gl2_update(const PC& X, float sA, float sB, const PC& T, float alpha, int first, int rot,
                                         bool conj_held) {
  PC o;
  if constexpr (MODE == 2) {   // transtacos/audio.py:137-138: angles = exp(1j * angle(X))
    const pf n2 = norm2(X);
    const float nA = plo(n2), nB = phi(n2);
    const float iA = nA > 0.f ? rsqrtf(nA) : 0.f, iB = nB > 0.f ? rsqrtf(nB) : 0.f;
    const pf sc = pk(sA * iA, sB * iB);
    o.re = mul2(X.re, sc);
    o.im = mul2(X.im, sc);
    if (nA == 0.f || nB == 0.f) {   // exp(1j * angle(0)) = 1, in the engine's internal form: i^k (conjugated on the b side)
      float2 un = rot_i(make_float2(1.f, 0.f), rot);
      if (conj_held) un.y = -un.y;
      o.re = pk(nA > 0.f ? plo(o.re) : sA * un.x, nB > 0.f ? phi(o.re) : sB * un.x);
      o.im = pk(nA > 0.f ? plo(o.im) : sA * un.y, nB > 0.f ? phi(o.im) : sB * un.y);
    }
  } else {                     // librosa.griffinlim: c = rebuilt - alpha * tprev; angles = c / (|c| + 1e-16)
    // tprev is loaded as zero on the first iteration (rebuilt = 0): the update is then exactly X, without a select
    PC c;
    c.re = fma2s(T.re, -alpha, X.re);
    c.im = fma2s(T.im, -alpha, X.im);
    const pf n2 = norm2(c);
    // MUFU sqrt / reciprocal (~1e-7 relative): far inside the Griffin-Lim tolerance, and branch-free
    const pf sc = pk(sA * gl2_rcp(fast_sqrt(plo(n2)) + 1e-16f), sB * gl2_rcp(fast_sqrt(phi(n2)) + 1e-16f));
    o.re = mul2(c.re, sc);
    o.im = mul2(c.im, sc);
  }
  return o;
}

// exp(2 pi i u), u in [0, 1): MUFU sine / cosine on the reduced argument 2 pi (u - 1/2) in [-pi, pi) (absolute error 2^-21.4,
// CUDA programming guide), negated.  The initial phase is a random draw; 5e-7 on it is far inside the Griffin-Lim tolerance,
// and sincospif cost as many instructions per item as a transform.
__device__ __forceinline__ void gl2_unit_phase(float u, float* s, float* c) {
  const float th = 6.283185307179586f * (u - 0.5f);
  *s = -__sinf(th);
  *c = -__cosf(th);
}

// Input spectrum of bin kb of both frames in the engine's internal form (modes 0 / 1): i^kb X, conjugated on the b side.
template <int MODE>
__device__ __forceinline__ PC gl2_fetch(const Gl2Args& a, long long rowA, long long rowB, bool okA, bool okB, int kb, int rot,
                                        bool conj_it) {
  float2 xa = make_float2(0.f, 0.f), xb = make_float2(0.f, 0.f);
  if constexpr (MODE == 0) {
    if (okA) xa = __ldg(a.spec + rowA + kb);
    if (okB) xb = __ldg(a.spec + rowB + kb);
  } else {
    if (okA) {
      float s, c;
      gl2_unit_phase(__ldg(a.init_phase + rowA + kb), &s, &c);
      const float mg = __ldg(a.S + rowA + kb);
      xa = make_float2(mg * c, mg * s);
    }
    if (okB) {
      float s, c;
      gl2_unit_phase(__ldg(a.init_phase + rowB + kb), &s, &c);
      const float mg = __ldg(a.S + rowB + kb);
      xb = make_float2(mg * c, mg * s);
    }
  }
  xa = rot_i(xa, rot);
  xb = rot_i(xb, rot);
  PC r;
  r.re = pk(xa.x, xb.x);
  r.im = conj_it ? pk(-xa.y, -xb.y) : pk(xa.y, xb.y);
  return r;
}

// Same from values fetched ahead of their use (modes 0 / 1): mode 0: tA / tB = spectrum values (already masked); mode 1:
// sA / sB = magnitudes (masked), tA.x / tB.x = initial phase u in [0, 1).
template <int MODE>
__device__ __forceinline__ PC gl2_fetch_v(float sA, float sB, float2 tA, float2 tB, int rot, bool conj_it) {
  float2 xa = tA, xb = tB;
  if constexpr (MODE == 1) {
    float s, c;
    gl2_unit_phase(tA.x, &s, &c);
    xa = make_float2(sA * c, sA * s);
    gl2_unit_phase(tB.x, &s, &c);
    xb = make_float2(sB * c, sB * s);
  }
  xa = rot_i(xa, rot);
  xb = rot_i(xb, rot);
  PC r;
  r.re = pk(xa.x, xb.x);
  r.im = conj_it ? pk(-xa.y, -xb.y) : pk(xa.y, xb.y);
  return r;
}

// Signal loads: read-only path when the buffers were written by an earlier LAUNCH; L2 (coherent across SMs after a grid
// barrier) when they were written earlier in the SAME launch (persistent kernel below).
template <bool COH, class T>
__device__ __forceinline__ T gl2_ldsig(const T* ptr) {
  if constexpr (COH) return __ldcg(ptr);
  else return __ldg(ptr);
}

// One pass over all tiles (see the file header).  COH: the signal buffers may have been written earlier in this launch.
template <int N, int MODE, bool COH>
__device__ __forceinline__ void gl2_pass(const PlanDev& p, const Gl2Args& a, Gl2Smem<N>& sm) {
  using C = Fft2Cfg<N>;
  constexpr int FT = kGl2GroupWarps * C::kFrames;   // frames per tile
  constexpr int GT = kGl2GroupWarps * 32;           // threads per group
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
  const long long n_tiles = static_cast<long long>(a.g.bd.B) * a.tiles_per_row;
  const int ga = a.groups_active > 0 ? a.groups_active : kGl2Groups;
  for (long long vt = group < ga ? static_cast<long long>(blockIdx.x) * ga + group : n_tiles; vt < n_tiles;
       vt += static_cast<long long>(gridDim.x) * ga) {
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
          const float4 p4 = gl2_ldsig<COH>(ya4 + j), q4 = gl2_ldsig<COH>(yb4 + j);
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
          const float val = gl2_ldsig<COH>(ya + uu) + gl2_ldsig<COH>(yb + uu);
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
            const char* tp0 = reinterpret_cast<const char*>(
                a.tprev + (gl2_pair_base(row.frame_base, b) + (min(item_t0, row.T - 1) >> 1)) * C::kF);
            for (int o = lane * 128; o < ((nfr + 1) >> 1) * C::kF * 16; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(tp0 + o));
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
      // ---- spectrum pass.  The analysed spectrum Z goes to the warp's buffer in natural bin order (16 bytes per bin: both
      // frames of a pair), the registers of the FFT are released, and the Hermitian pairs (k, Nz - k) are updated in place
      // by a short loop: forward split -> phase update against S (and the previous spectrum) -> inverse split.  Lanes take
      // consecutive bins, so S / tprev accesses are coalesced and several pairs are in flight per lane.
      ulonglong2* const xz = reinterpret_cast<ulonglong2*>(xbuf);   // [P][Nz] (re pair, im pair)
      if constexpr (MODE >= 2) {
        const unsigned wz = smem_u32(xz + pl * C::kNz + k1);
        static_for<0, 32>([&](auto kc) {
          constexpr int k2 = decltype(kc)::value;
          sts_pc<16 * C::kR2 * k2>(wz, v[k2].re, v[k2].im);
        });
        __syncwarp();
      }
      static_for<0, C::kP>([&](auto pc_) {
        constexpr int pp = decltype(pc_)::value;
        const int fA = item_t0 + 2 * pp;
        const bool okA = fA < row.T, okB = fA + 1 < row.T;
        // frames past the end are clamped to the last one (loads stay in bounds) and masked
        const long long rowA = (row.frame_base + min(fA, row.T - 1)) * C::kF, rowB = (row.frame_base + min(fA + 1, row.T - 1)) * C::kF;
        const float mA = okA ? 1.f : 0.f, mB = okB ? 1.f : 0.f;
        // packed previous spectrum of this pair (a pair past the end is clamped to the last one: loads stay in bounds and finite)
        ulonglong2* const tpp = a.tprev + (gl2_pair_base(row.frame_base, b) + (min(fA, row.T - 1) >> 1)) * C::kF;
        ulonglong2* const zp = xz + pp * C::kNz;
        // everything one Hermitian pair (k, Nz - k) reads, fetched ahead of its use
        struct PairLd {
          ulonglong2 zk, zr;
          float sAa, sBa, sAb, sBb;
          float2 tAa, tBa, tAb, tBb;   // modes 0 / 1: input spectrum / initial phase
          ulonglong2 ta, tb;           // mode 3: previous rebuilt spectrum of bins k / Nz - k, packed (re pair, im pair)
        };
        // ld_pair_g: what comes from global memory (S, previous spectrum / input spectrum / initial phase), fetched TWO trips ahead
        // (L2 hits at ~700 cycles: one trip of lead left a third of the spectrum pass's stalls on the long scoreboard);
        // ld_pair_s: the pair's analysed spectrum from the warp's shared-memory buffer, one trip ahead.
        auto ld_pair_s = [&](int k, PairLd& L) {
          if constexpr (MODE >= 2) {
            L.zk = zp[k];
            L.zr = zp[(C::kNz - k) & (C::kNz - 1)];
          }
        };
        auto ld_pair_g = [&](int k) -> PairLd {
          PairLd L;
          if constexpr (MODE == 0) {   // the input spectrum, masked (rows past the end are clamped)
            const int kb = C::kNz - k;
            const float2 a0 = __ldg(a.spec + rowA + k), b0 = __ldg(a.spec + rowB + k);
            const float2 a1 = __ldg(a.spec + rowA + kb), b1 = __ldg(a.spec + rowB + kb);
            L.tAa = a0;   // raw: the masks are applied where the values are used -- a multiply here would wait for the
            L.tBa = b0;   // load at once and undo the one-trip-ahead fetch
            L.tAb = a1;
            L.tBb = b1;
          } else if constexpr (MODE == 1) {   // magnitudes and initial phases: without this the loads sit behind each sincospif
            const int kb = C::kNz - k;
            L.sAa = __ldg(a.S + rowA + k);    // raw: masked at the use (see above)
            L.sBa = __ldg(a.S + rowB + k);
            L.sAb = __ldg(a.S + rowA + kb);
            L.sBb = __ldg(a.S + rowB + kb);
            L.tAa.x = __ldg(a.init_phase + rowA + k);
            L.tBa.x = __ldg(a.init_phase + rowB + k);
            L.tAb.x = __ldg(a.init_phase + rowA + kb);
            L.tBb.x = __ldg(a.init_phase + rowB + kb);
          }
          if constexpr (MODE >= 2) {
            const int kb = C::kNz - k;
            L.sAa = __ldg(a.S + rowA + k);    // raw: masked at the use (see above)
            L.sBa = __ldg(a.S + rowB + k);
            L.sAb = __ldg(a.S + rowA + kb);
            L.sBb = __ldg(a.S + rowB + kb);
            L.ta = L.tb = make_ulonglong2(0ull, 0ull);
            if constexpr (MODE == 3) {
              if (!a.first) {
                L.ta = tpp[k];
                L.tb = tpp[kb];
              }
            }
          }
          return L;
        };
        // one Hermitian pair: bins k (a side) and Nz - k (b side, held conjugated)
        auto do_pair = [&](int k, const PairLd& L, auto sinform_c) {
          constexpr bool SINFORM = decltype(sinform_c)::value;
          const int km = (C::kNz - k) & (C::kNz - 1), kb = C::kNz - k;
          const float2 tws = sm.spn[k];
          PC P, Q;
          if constexpr (MODE >= 2) {
            PC Zk, Zr;
            Zk.re = L.zk.x; Zk.im = L.zk.y;
            Zr.re = L.zr.x; Zr.im = L.zr.y;
            split2<SINFORM>(Zk, Zr, tws, P, Q);
            if constexpr (MODE == 3) {   // the rebuilt spectrum becomes the next iteration's tprev (both frames of the pair: the
              if (okA) {                 // second half of an odd utterance's last pair is the spectrum of a zero frame, finite)
                tpp[k] = make_ulonglong2(P.re, P.im);
                tpp[kb] = make_ulonglong2(Q.re, Q.im);
              }
            }
            PC Ta, Tb;
            Ta.re = L.ta.x; Ta.im = L.ta.y;
            Tb.re = L.tb.x; Tb.im = L.tb.y;
            P = gl2_update<MODE>(P, L.sAa * mA, L.sBa * mB, Ta, a.alpha, a.first, k & 3, false);
            Q = gl2_update<MODE>(Q, L.sAb * mA, L.sBb * mB, Tb, a.alpha, a.first, (4 - (k & 3)) & 3, true);
          } else {
            if constexpr (MODE == 0) {
              P = gl2_fetch_v<MODE>(0.f, 0.f, make_float2(L.tAa.x * mA, L.tAa.y * mA), make_float2(L.tBa.x * mB, L.tBa.y * mB), k & 3, false);
              Q = gl2_fetch_v<MODE>(0.f, 0.f, make_float2(L.tAb.x * mA, L.tAb.y * mA), make_float2(L.tBb.x * mB, L.tBb.y * mB),
                                    (4 - (k & 3)) & 3, true);
            } else {
              P = gl2_fetch_v<MODE>(L.sAa * mA, L.sBa * mB, L.tAa, L.tBa, k & 3, false);
              Q = gl2_fetch_v<MODE>(L.sAb * mA, L.sBb * mB, L.tAb, L.tBb, (4 - (k & 3)) & 3, true);
            }
          }
          if (k == 0) {   // irfft ignores the imaginary parts of DC and Nyquist
            P.im = 0ull;
            Q.im = 0ull;
          }
          PC Zk2, Zr2;
          split2_inv<SINFORM>(P, Q, tws, Zk2, Zr2);
          zp[k] = make_ulonglong2(Zk2.re, Zk2.im);
          if (k != 0) zp[km] = make_ulonglong2(Zr2.re, Zr2.im);
        };
        // bins [0, Nz/4) use the tan / cos form of the split twiddle, [Nz/4, Nz/2) the cot / sin form: one of each per trip;
        // the loads of the next trip are issued before the current one is computed
        {
          constexpr int kTrips = C::kNz / 128;
          static_assert(kTrips >= 2, "the pair loop runs two trips ahead");
          PairLd c0 = ld_pair_g(lane), c1 = ld_pair_g(lane + C::kNz / 4);
          PairLd d0 = ld_pair_g(lane + 32), d1 = ld_pair_g(lane + 32 + C::kNz / 4);
          ld_pair_s(lane, c0);
          ld_pair_s(lane + C::kNz / 4, c1);
#pragma unroll (COH ? kGl2PairUnrollCoh : kGl2PairUnroll)
          for (int ii = 0; ii < kTrips; ++ii) {   // rolled: a single-tile launch runs this code once, from a cold instruction cache
            const int k = lane + 32 * ii;
            PairLd e0 = d0, e1 = d1;
            if (ii + 2 < kTrips) {
              e0 = ld_pair_g(k + 64);
              e1 = ld_pair_g(k + 64 + C::kNz / 4);
            }
            if (ii + 1 < kTrips) {
              ld_pair_s(k + 32, d0);
              ld_pair_s(k + 32 + C::kNz / 4, d1);
            }
            do_pair(k, c0, std::false_type{});
            do_pair(k + C::kNz / 4, c1, std::true_type{});
            c0 = d0;
            c1 = d1;
            d0 = e0;
            d1 = e1;
          }
        }
        // bin Nz/2 pairs with itself: Q = conj(P) (computed by every lane, stored by lane 0)
        {
          constexpr int k = C::kNz / 2;
          const float2 tws = sm.spn[k];
          PC P, Q;
          if constexpr (MODE >= 2) {
            const ulonglong2 zk = zp[k];
            const float sAa = __ldg(a.S + rowA + k) * mA, sBa = __ldg(a.S + rowB + k) * mB;
            PC Ta;
            Ta.re = Ta.im = 0ull;
            if constexpr (MODE == 3) {
              if (!a.first) {
                const ulonglong2 t = tpp[k];
                Ta.re = t.x; Ta.im = t.y;
              }
            }
            PC Zk;
            Zk.re = zk.x; Zk.im = zk.y;
            split2<true>(Zk, Zk, tws, P, Q);
            if constexpr (MODE == 3) {
              if (lane == 0 && okA) tpp[k] = make_ulonglong2(P.re, P.im);
            }
            P = gl2_update<MODE>(P, sAa, sBa, Ta, a.alpha, a.first, 0, false);
          } else {
            P = gl2_fetch<MODE>(a, rowA, rowB, okA, okB, k, 0, false);
          }
          Q.re = P.re;
          Q.im = sub2(0ull, P.im);
          PC Zk2, Zr2;
          split2_inv<true>(P, Q, tws, Zk2, Zr2);
          __syncwarp();   // (MODE 3: all lanes have read tprev[k] before lane 0's store above is visible -- same value anyway)
          if (lane == 0) zp[k] = make_ulonglong2(Zk2.re, Zk2.im);
        }
      });
      __syncwarp();
      {
        const ulonglong2* rz = xz + pl * C::kNz + k1;
        static_for<0, 32>([&](auto kc) {
          constexpr int k2 = decltype(kc)::value;
          const ulonglong2 z = rz[C::kR2 * k2];
          v[k2].re = z.x;
          v[k2].im = z.y;
        });
      }
      __syncwarp();   // spectrum read before the inverse transform reuses the buffer
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
    // the next tile of this group starts by staging its signal: pull it towards L1 now (multi-launch path: the buffers were
    // written by the previous launch and sit in L2)
    if constexpr (MODE >= 2 && !COH) {
      const long long nvt = vt + static_cast<long long>(gridDim.x) * ga;
      if (nvt < n_tiles) {
        const int nb = static_cast<int>(nvt / a.tiles_per_row), ntk = static_cast<int>(nvt - static_cast<long long>(nb) * a.tiles_per_row);
        const GlRow nrow = gl_row(a.g, nb, N, hop);
        if (ntk * FT < nrow.T) {
          const long long nbase = gl2_sig_base(nrow, nb, hop, C::kWin) + static_cast<long long>(ntk) * FT * hop;
          const int nbytes = ((FT - 1) * hop + C::kWin) * 4;
          for (int o = gt * 128; o < nbytes; o += GT * 128) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(a.ya_in + nbase) + o));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(a.yb_in + nbase) + o));
          }
        }
      }
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
      bool done = false;
      if constexpr (N == 2048) {
        if (C::kWin == 4 * hop && (hop & 1) == 0) {
          // The reference's framing (n_fft 2048, win 1024, hop 256), fully unrolled: frame f of the tile adds its four hop-sized
          // chunks to samples q = f .. f + 3 of the thread's offset o.  Every shared-memory address is the thread's base plus a
          // compile-time constant and every accumulator a fixed register: 32 loads, 64 adds and the stores, where the rolled loops
          // below spend ~1000 instructions per thread on index arithmetic and branches (the overlap-add was 15 % of the kernel's
          // time for 64 useful additions).  Frames run from the last to the first, so each sample adds its terms in the order of the
          // rolled loop (d ascending): same bits.
          constexpr int ND = 4, NQ = FT + ND - 1, kHop = C::kWin / ND;
          for (int o = 2 * gt; o < hop; o += 2 * GT) {
            float2 acc[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) acc[q] = make_float2(0.f, 0.f);
            const float* base = frames0 + o;
            static_for<0, FT>([&](auto fc) {
              constexpr int f = FT - 1 - decltype(fc)::value;
              const float* fp = base + (f / C::kFrames) * kWarpStride + (f % C::kFrames) * C::kWin;
              static_for<0, ND>([&](auto dc) {
                constexpr int d = decltype(dc)::value;
                const float2 t2 = *reinterpret_cast<const float2*>(fp + d * kHop);
                acc[f + d].x += t2.x;
                acc[f + d].y += t2.y;
              });
            });
            static_for<0, NQ>([&](auto qc) {
              constexpr int q = decltype(qc)::value;
              const int j = q * kHop + o;
              if (j < span) {
                *reinterpret_cast<float2*>(mine + j) = acc[q];
                constexpr bool interior = q >= ND - 1 && q < FT;   // j >= win - hop && j < FT * hop for every o < hop
                if (interior || (tk == 0 && q < ND - 1) || (last_tile && q >= FT))
                  *reinterpret_cast<float2*>(other + j) = make_float2(0.f, 0.f);
              }
            });
          }
          done = true;
        }
      }
      if (done) {
      } else if (C::kWin % hop == 0 && (hop & 1) == 0) {
        // win a multiple of hop (the reference framing): every d in [max(0, q - FT + 1), min(nov, q)] is a valid term, no
        // per-term tests; two neighbouring offsets per thread (8-byte shared loads and global stores; same summation order)
        for (int o = 2 * gt; o < hop; o += 2 * GT) {
          for (int q = 0, j = o; j < span; ++q, j += hop) {
            const int dlo = max(0, q - (FT - 1)), dhi = min(nov, q);
            float2 acc = make_float2(0.f, 0.f);
            for (int d = dlo; d <= dhi; ++d) {
              const int f = q - d;
              const float2 t2 = *reinterpret_cast<const float2*>(frames0 + (f / C::kFrames) * kWarpStride + (f % C::kFrames) * C::kWin + o + d * hop);
              acc.x += t2.x;
              acc.y += t2.y;
            }
            *reinterpret_cast<float2*>(mine + j) = acc;
            const bool interior = j >= C::kWin - hop && j < FT * hop;
            if (interior || (tk == 0 && j < C::kWin - hop) || (last_tile && j >= FT * hop))
              *reinterpret_cast<float2*>(other + j) = make_float2(0.f, 0.f);
          }
        }
      } else {
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
    }
    gl2_group_sync(group);
  }
}

template <int N, int MODE>
__global__ void __launch_bounds__(kGl2Warps * 32, 1) gl2_kernel(const PlanDev p, const Gl2Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Gl2Smem<N> sm;
  sm.init(smem_raw, p);
  gl2_pass<N, MODE, false>(p, a, sm);
}

// ---- persistent Griffin-Lim: init + all iterations in ONE cooperative launch --------------------------------------------
// For small batches (every tile group resident at once: tiles <= 2 x SMs, e.g. one 5 s utterance = 54 tiles) an iteration
// is the latency of ONE tile, and with one launch per iteration most of that latency is fetching the kernel's ~5000
// straight-line instructions again.  Here the CTAs stay resident, the code stays in the instruction cache, and the
// iterations are separated by a grid barrier (arrive / spin on a global counter; the launch is cooperative, so all CTAs
// are co-resident).  The signal written before a barrier is read after it through L2 (gl2_ldsig<true>); the per-bin state
// (tprev) of a tile is only ever touched by the CTA that owns the tile.
__device__ __forceinline__ void gl2_grid_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
    __threadfence();
  }
  __syncthreads();
}

template <int N, int FORM>   // FORM 0: angle form (mode 2), 1: fast form (mode 3)
__global__ void __launch_bounds__(kGl2Warps * 32, 1) gl2_persistent_kernel(const PlanDev p, Gl2Args a, int n_iter, float* sig,
                                                                             long long sig_elems, unsigned* counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Gl2Smem<N> sm;
  sm.init(smem_raw, p);
  // signal buffers: [parity pair cur][a / b]
  float* const b00 = sig;
  float* const b01 = sig + sig_elems;
  float* const b10 = sig + 2 * sig_elems;
  float* const b11 = sig + 3 * sig_elems;
  a.ya_out = b00;
  a.yb_out = b01;
  gl2_pass<N, 1, true>(p, a, sm);
  unsigned target = gridDim.x;
  gl2_grid_barrier(counter, target);
  int cur = 0;
#pragma unroll 1
  for (int it = 0; it < n_iter; ++it) {
    a.ya_in = cur ? b10 : b00;
    a.yb_in = cur ? b11 : b01;
    a.ya_out = cur ? b00 : b10;
    a.yb_out = cur ? b01 : b11;
    a.first = (it == 0);
    gl2_pass<N, 2 + FORM, true>(p, a, sm);
    target += gridDim.x;
    gl2_grid_barrier(counter, target);
    cur ^= 1;
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
