// Warp-specialised STFT-magnitude + mel feature kernel (packed FFT engine, fft2.cuh):
//   [pre-emphasis ->] reflect pad -> frame gather -> window -> real FFT -> |.|^2   (8 "analysis" warps, FMA pipe)
//   |.|^2 -> dB-normalise / ln / raw magnitude stores, sqrt -> banded mel -> scale   (8 "epilogue" warps, MUFU / LSU)
//   transtacos/audio.py:73-77 get_specs;  retunegan/audio.py:116-128 get_mag / get_mel.
// Why two roles: in the single-role kernel (feat2.cuh) the 8 warps of an SM drift into lock-step, so the FMA-bound
// FFT, the shared-memory-bound exchange / mel phases and the MUFU-bound log phase run one after the other and each
// pipe idles most of the time.  Here analysis warp w hands the squared magnitudes of one item (a frame pair set,
// see feat2.cuh) to epilogue warp w through a private shared-memory buffer guarded by a named barrier per warp pair (pair_bar_sync),
// and moves on to the next item's FFT while the epilogue warp takes logs, stores and runs the mel filterbank.
// Registers are re-partitioned with setmaxnreg (analysis 200, epilogue 56 per thread; 128 at launch).
#pragma once
#include <cuda.h>   // CUtensorMap (type only: the encoder is reached through cudaGetDriverEntryPoint, no libcuda link)

#include "feat2.cuh"

namespace sb200 {

// Mel formulation of the epilogue: the banded ELL filterbank (default) or, with -DSB200_MEL_SEGMENTS (tools/variants.py,
// tools/ab_mel.sh), the segment formulation (PlanDev::seg_slot: every band bin read once, no weight loads).  Measured on
// config 3: segments 91.6 us, ELL 87.0 us -- fewer shared-memory wavefronts, but the masked trips and the carry chain between
// rounds cost more issue slots than the weights did.
#ifndef SB200_MEL_SEGMENTS
#define SB200_MEL_ELL
#endif

constexpr int kFeat3Pairs = 8;                 // analysis / epilogue warp pairs per CTA
constexpr int kFeat3Threads = 2 * kFeat3Pairs * 32;
constexpr int kFeat3PbufElems = 1024 + 16;     // [P][Nz] powers of both frames of a pair + P Nyquist bins + zero pad
#ifndef kFeat3AnalysisRegs
#define kFeat3AnalysisRegs 200
#endif
#ifndef kFeat3EpilogueRegs
#define kFeat3EpilogueRegs 56
#endif

__device__ __forceinline__ void mbar_init(unsigned addr, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned addr) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(addr) : "memory");
}
// try_wait suspends the thread until the phase completes or a time limit expires; the explicit (large) suspend-time hint keeps
// a waiting warp from waking up every few hundred cycles to re-issue the instruction (ncu: ~190 re-issues per item and epilogue
// warp without it -- issue slots taken from the analysis warps).
__device__ __forceinline__ void mbar_wait(unsigned addr, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
#ifndef SB200_TRYWAIT_NO_HINT
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
#else
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
#endif
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(addr), "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned addr, unsigned bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(addr), "r"(bytes) : "memory");
}
// One bulk-tensor (TMA) copy of a 2-D box global -> shared, completion counted in bytes on an mbarrier (SASS: UTMALDG).
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// Staging of an interior item's raw samples (hop 256, n_fft 2048).  A bulk-tensor copy must START on a 16-byte boundary of global
// memory (measured: tools/sanitizer/tma_align.cu traps with "illegal instruction" at element offset 1), but the rows of a [B, L]
// batch with odd L start anywhere.  The flat signal buffer is therefore described to the TMA unit as the 2-D tensor
// t[j][c] = x[64 j + c] (overlapping rows: row stride 256 B, row length = the whole buffer), and the box {64 columns, 21 rows}
// at (c0, 0) is the 1344 CONTIGUOUS samples x[c0 .. c0 + 1343]; c0 = (first sample of the pair's first frame - 4) rounded down
// to a multiple of 4.  Sample i of the pair then sits at float index lead + i of the staging area, lead = 4 .. 7, and
// x[p0 - 1], which the pre-emphasis FIR needs, at lead - 1.
constexpr int kStageBoxCols = 64, kStageBoxRows = 21;
constexpr int kStageFloats = kStageBoxCols * kStageBoxRows;     // 1344 >= 7 + 1280
constexpr unsigned kStageBytes = kStageFloats * 4;

// Hand-off between analysis warp w and epilogue warp w.  DEFAULT: one named barrier per warp pair (bar.sync id, 64), hit twice
// per item by both warps: "empty" (the epilogue warp is done reading the previous item's powers; the analysis warp may
// overwrite them) and "full" (the powers of this item are written; the epilogue warp may read them).  The hardware parks the
// waiting warp: no polling loop, and compute-sanitizer's racecheck / synccheck model bar.sync exactly (0 hazards / 0 errors on
// this build, profiles/r02_sanitizer/).  -DSB200_HANDOFF_MBARRIER (tools/variants.py) selects round 1's formulation instead, two
// mbarriers per pair (full / empty, 32 arrivals each, try_wait loops): same speed within 0.3 % (87.3 vs 87.0 us on config 3), but
// the two tools report the accesses on either side of it as hazards / "missing init" although a minimal kernel with the same
// PTX (tools/sanitizer/mbar_minimal.cu) is clean under both -- an unexplained tool report the product build should not carry.
__device__ __forceinline__ void pair_bar_sync(int w) {
#ifndef SB200_HANDOFF_MBARRIER
  asm volatile("bar.sync %0, 64;" ::"r"(w + 1) : "memory");
#endif
}
#ifdef SB200_HANDOFF_MBARRIER
#define SB200_HANDOFF_MBAR(x) x
#else
#define SB200_HANDOFF_MBAR(x)
#endif

template <int N>
struct Smem3 {
  using C = Fft2Cfg<N>;
  uint4* xbufs;             // [pairs][kXElems] exchange buffers of the analysis warps
  pf* pbufs;                // [pairs][kFeat3PbufElems]
  float* win;               // [win] 0.5 * analysis window
  float2* tw;               // [kTwCount]
  float2* sp2;              // [17*32]
#ifdef SB200_MEL_ELL
  float* melw;              // [melw_count]
  int* mel_lo;              // [32*rounds]
#else
  int4* seg_slot;           // [32*seg_rounds] segment view of the filterbank (plan.cuh)
  float2* seg_coef;         // [32*seg_rounds]
  float* enorm;             // [n_mel rounded up to 4]
#endif
  unsigned long long* bar;  // [3*pairs]: full[w], empty[w], staged[w] (bulk-tensor copy of the next item's samples landed)
  __host__ __device__ static size_t mel_table_bytes(const PlanDev& p) {
#ifdef SB200_MEL_ELL
    return sizeof(float) * p.melw_count + sizeof(int) * 32 * p.mel_rounds;
#else
    return (sizeof(int4) + sizeof(float2)) * 32 * p.seg_rounds + sizeof(float) * ((p.n_mel + 3) / 4 * 4);
#endif
  }
  __host__ __device__ static size_t bytes(const PlanDev& p) {
    return static_cast<size_t>(kFeat3Pairs) * (C::kXBytes + kFeat3PbufElems * sizeof(pf)) + sizeof(float) * C::kWin +
           sizeof(float2) * (C::kTwCount + 17 * 32) + mel_table_bytes(p) + sizeof(unsigned long long) * 3 * kFeat3Pairs;
  }
  __device__ __forceinline__ void carve(unsigned char* raw, const PlanDev& p) {
    xbufs = reinterpret_cast<uint4*>(raw);
    pbufs = reinterpret_cast<pf*>(raw + static_cast<size_t>(kFeat3Pairs) * C::kXBytes);
    win = reinterpret_cast<float*>(pbufs + kFeat3Pairs * kFeat3PbufElems);
    tw = reinterpret_cast<float2*>(win + C::kWin);
    sp2 = tw + C::kTwCount;
#ifdef SB200_MEL_ELL
    melw = reinterpret_cast<float*>(sp2 + 17 * 32);
    mel_lo = reinterpret_cast<int*>(melw + p.melw_count);
    bar = reinterpret_cast<unsigned long long*>(mel_lo + 32 * p.mel_rounds);
#else
    seg_slot = reinterpret_cast<int4*>(sp2 + 17 * 32);
    seg_coef = reinterpret_cast<float2*>(seg_slot + 32 * p.seg_rounds);
    enorm = reinterpret_cast<float*>(seg_coef + 32 * p.seg_rounds);
    bar = reinterpret_cast<unsigned long long*>(enorm + (p.n_mel + 3) / 4 * 4);
#endif
  }
  template <class T>
  static __device__ __forceinline__ void copy16(T* dst, const T* src, int count, float scale = 1.f) {
    const int n16 = count * static_cast<int>(sizeof(T)) / 16;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll 2
    for (int i = threadIdx.x; i < n16; i += kFeat3Threads) {
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
#ifdef SB200_MEL_ELL
      copy16(melw, p.melw, p.melw_count);
      copy16(mel_lo, p.mel_lo, 32 * p.mel_rounds);
#else
      copy16(seg_slot, p.seg_slot, 32 * p.seg_rounds);
      copy16(seg_coef, p.seg_coef, 32 * p.seg_rounds);
      for (int i = threadIdx.x; i < p.n_mel; i += kFeat3Threads) enorm[i] = p.mel_enorm[i];
#endif
    }
    for (int i = threadIdx.x; i < kFeat3Pairs * kFeat3PbufElems; i += kFeat3Threads) pbufs[i] = 0ull;
  }
};

// Pass-A registers of one item (layout: load_item2 in feat2.cuh).  HS > 0 (hop == 64 HS, n_fft 2048): the two frames
// of an interior pair overlap by R - HS lane slots, so the pair needs R + HS slots of (even, odd) samples once.  The
// "previous sample" of the pre-emphasis FIR, x[p0 + 2 lane + 64 r - 1], is the odd sample of lane - 1 (slot r), or of
// lane 31 (slot r - 1) for lane 0: it comes from a shuffle, only x[p0 - 1] is loaded on top.
template <int N, bool PRE, int HS>
__device__ __forceinline__ void gather_item3(PC (&v)[32], const Item& it, const float* __restrict__ x, int hop, float pre,
                                             const float* __restrict__ s_win, float* stage, int lane) {
  using C = Fft2Cfg<N>;
  bool shared = false;
  if constexpr (HS > 0) {
    const long long p0 = static_cast<long long>(it.t0) * hop - N / 4;
    shared = (it.t0 + 1 < it.T) && p0 >= 1 && p0 + hop + C::kWin <= it.L;
    if (shared) {   // loads and their use stay inside one block: conditionally defined registers end up in local memory
      constexpr int kSlots = C::kR + HS;
      float lo[kSlots], hi[kSlots];
      const float* xp = x + it.sig_base + p0 + 2 * lane;
      static_for<0, kSlots>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        lo[r] = __ldg(xp + 64 * r);
        hi[r] = __ldg(xp + 64 * r + 1);
      });
      if constexpr (PRE) {
        float carry = __ldg(x + it.sig_base + p0 - 1);
        static_for<0, kSlots>([&](auto rc) {
          constexpr int r = decltype(rc)::value;
          const float t = __shfl_sync(kFullMask, hi[r], (lane + 31) & 31);
          const float pv = lane == 0 ? carry : t;
          carry = t;
          const float l = lo[r];
          lo[r] = fmaf(-pre, pv, l);
          hi[r] = fmaf(-pre, l, hi[r]);
        });
      }
      static_for<0, C::kR>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        const float2 w = *reinterpret_cast<const float2*>(s_win + 2 * lane + 64 * r);
        constexpr int idx = brev(r, C::kLogR2);
        v[idx].re = pk(lo[r] * w.x, lo[r + HS] * w.x);   // scalar products land in the two halves directly: no packing MOVs
        v[idx].im = pk(hi[r] * w.y, hi[r + HS] * w.y);
        v[idx + 1] = v[idx];
      });
    }
  }
  if (!shared) {
    static_for<0, C::kP>([&](auto pc_) {
      constexpr int p = decltype(pc_)::value;
      float re[2][C::kR], im[2][C::kR];
      const int t = it.t0 + 2 * p;
      load_frame2<N, PRE>(re[0], im[0], x + it.sig_base, it.L, t, it.T, hop, pre, s_win, stage, lane);
      load_frame2<N, PRE>(re[1], im[1], x + it.sig_base, it.L, t + 1, it.T, hop, pre, s_win, stage, lane);
      static_for<0, C::kR>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        constexpr int idx = p * C::kR2 + brev(r, C::kLogR2);
        v[idx].re = pk(re[0][r], re[1][r]);
        v[idx].im = pk(im[0][r], im[1][r]);
        v[idx + 1] = v[idx];
      });
    });
  }
}

// Pass-A registers of an interior item whose samples a bulk-tensor copy has staged in shared memory: sample i of the pair at
// stage[lead + i].  Same arithmetic as the global-load path of gather_item3 (bit-identical results); the FIR's previous sample
// is read from the staging area too (no shuffle, no carry).  ODD = lead is odd: the 8-byte aligned pair then holds
// (previous, even) and the odd sample comes from the 4-byte load; else (even, odd) and the previous sample does.
template <int N, bool PRE, int HS, bool ODD>
__device__ __forceinline__ void gather_staged3(PC (&v)[32], const float* __restrict__ q /* stage + (lead & ~1) + 2 lane */,
                                               float pre, const float* __restrict__ s_win, int lane) {
  using C = Fft2Cfg<N>;
  constexpr int kSlots = C::kR + HS;
  float lo[kSlots], hi[kSlots];
  static_for<0, kSlots>([&](auto rc) {
    constexpr int r = decltype(rc)::value;
    const float2 t = *reinterpret_cast<const float2*>(q + 64 * r);
    if constexpr (!ODD) {
      lo[r] = t.x;
      hi[r] = t.y;
      if constexpr (PRE) {
        const float pv = q[64 * r - 1];
        hi[r] = fmaf(-pre, t.x, t.y);
        lo[r] = fmaf(-pre, pv, t.x);
      }
    } else {
      const float h = q[64 * r + 2];
      lo[r] = t.y;
      hi[r] = h;
      if constexpr (PRE) {
        hi[r] = fmaf(-pre, t.y, h);
        lo[r] = fmaf(-pre, t.x, t.y);
      }
    }
  });
  static_for<0, C::kR>([&](auto rc) {
    constexpr int r = decltype(rc)::value;
    const float2 w = *reinterpret_cast<const float2*>(s_win + 2 * lane + 64 * r);
    constexpr int idx = brev(r, C::kLogR2);
    v[idx].re = pk(lo[r] * w.x, lo[r + HS] * w.x);
    v[idx].im = pk(hi[r] * w.y, hi[r + HS] * w.y);
    v[idx + 1] = v[idx];
  });
}

// What the analysis warps need to know about the bulk-tensor staging (TMA == true only).
struct StageArgs {
  long long total;   // samples the flat buffer holds from x: a staged box must end inside it (else the item takes the edge path)
};

// ---- analysis warp: gather + window + FFT + Hermitian split; |A|^2 of every bin goes to the pair's power buffer ----
template <int N, bool PRE, int HS, bool TMA>
__device__ __forceinline__ void feat3_analysis(const PlanDev& p, const FeatArgs& a, Smem3<N>& sm, int w, int lane,
                                               const CUtensorMap* tmap, const StageArgs sa_) {
  using C = Fft2Cfg<N>;
  uint4* xbuf = sm.xbufs + w * C::kXElems;
  pf* pbuf = sm.pbufs + w * kFeat3PbufElems;
  const unsigned full = smem_u32(sm.bar + w), empty = smem_u32(sm.bar + kFeat3Pairs + w);
  const int k1 = lane & (C::kR2 - 1), pl = lane / C::kR2;   // pass-B role of this lane: column k1 of pair pl
  const bool col0 = (k1 == 0);
  const int partner = (lane & ~(C::kR2 - 1)) | ((C::kR2 - k1) & (C::kR2 - 1));
  pf* const sa = pbuf + pl * C::kNz + k1;                                  // bins k1 + R2 s
  pf* const sb = pbuf + pl * C::kNz + C::kNz - k1;     // bins Nz - k1 - R2 s
  pf* const sb0 = col0 ? pbuf + C::kP * C::kNz + pl : sb;   // s = 0: column 0 holds the Nyquist bin, kept after the P rows
  const float2* const sp = sm.sp2 + lane;
  const long long warps_total = static_cast<long long>(gridDim.x) * kFeat3Pairs;
  long long item = static_cast<long long>(blockIdx.x) * kFeat3Pairs + w;
  unsigned round = 0;
  // TMA: while pass B and the split of item n run, ONE elected lane has the TMA unit copy the samples of item n + 1 (an
  // interior frame pair: 1280 contiguous samples and the few before them) into this warp's exchange buffer, which is idle from
  // the transposed read of item n to the transposed write of item n + 1; the copy completes on an mbarrier (bytes).  The
  // gather then costs 56 shared-memory loads and no global-load latency on the critical warp.
  const unsigned staged_bar = smem_u32(sm.bar + 2 * kFeat3Pairs + w);
  unsigned staged_phase = 0;
  bool staged = false;
  auto stage_start = [&](const Item& ni) -> long long {   // first sample of the box, or -1 if the item is not staged
    const long long p0 = static_cast<long long>(ni.t0) * p.hop - N / 4;
    const long long g0 = (ni.sig_base + p0 - 4) & ~3LL;
    const bool interior = (ni.t0 + 1 < ni.T) && p0 >= 1 && p0 + p.hop + C::kWin <= ni.L && ni.sig_base + p0 >= 4 &&
                          g0 + kStageFloats <= sa_.total;
    return interior ? g0 : -1;
  };
  auto stage_next = [&](long long nxt) -> bool {
    if (nxt >= a.bd.total_items) return false;
    const long long g0 = stage_start(decode_item(a.bd, nxt, C::kFrames));
    if (g0 >= 0 && lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the warp's generic-proxy reads of the buffer come first
      mbar_expect_tx(staged_bar, kStageBytes);
      tma_load_2d(smem_u32(xbuf), tmap, static_cast<int>(g0), 0, staged_bar);
    }
    return g0 >= 0;
  };
  if constexpr (TMA) staged = stage_next(item);
  for (; item < a.bd.total_items; item += warps_total, ++round) {
    // No register prefetch across items: ptxas spills whatever stays live over the loop edge when the register budget comes
    // from setmaxnreg.  (Measured and dropped: an L2 prefetch of the next item from the epilogue warp, +1 us; a second copy
    // of the loop body for the edge path, +3.7 us; cp.async staging, +7 us.)
    const Item it = decode_item(a.bd, item, C::kFrames);
    PC v[32];
    if constexpr (TMA) {
      if (staged) {
        const long long p0 = static_cast<long long>(it.t0) * p.hop - N / 4;
        const int lead = static_cast<int>(it.sig_base + p0 - stage_start(it));   // 4 .. 7
        const float* q = reinterpret_cast<const float*>(xbuf) + (lead & ~1) + 2 * lane;
        mbar_wait(staged_bar, staged_phase);
        staged_phase ^= 1;
        if (lead & 1) gather_staged3<N, PRE, HS, true>(v, q, a.pre, sm.win, lane);
        else gather_staged3<N, PRE, HS, false>(v, q, a.pre, sm.win, lane);
        __syncwarp();   // every lane has its samples before the transpose overwrites the staging area
      } else {          // edge items (reflect padding, odd tail): per-frame path
        gather_item3<N, PRE, 0>(v, it, a.x, p.hop, a.pre, sm.win, reinterpret_cast<float*>(xbuf), lane);
      }
      fft2_forward_h<N, true>(v, xbuf, sm.tw, lane, [&] { staged = stage_next(item + warps_total); });
    } else {
      gather_item3<N, PRE, HS>(v, it, a.x, p.hop, a.pre, sm.win, reinterpret_cast<float*>(xbuf), lane);
      fft2_forward<N>(v, xbuf, sm.tw, lane);
    }
    // lane (pl, k1) now holds Z[k1 + R2*k2] of frames t0 + 2 pl, t0 + 2 pl + 1
    // all split twiddles up front: a table load written after a store to the power buffer cannot be hoisted above it
    float2 spv[17];
    static_for<0, 17>([&](auto sc) {
      constexpr int s = decltype(sc)::value;
      spv[s] = sp[s * 32];
    });
    SB200_HANDOFF_MBAR(mbar_wait(empty, (round & 1) ^ 1));   // the epilogue warp is done with the previous item's powers
    pair_bar_sync(w);
    {
      // self pair of column 0 (bin Nz/2) first: the exchange below overwrites v[16]
      PC ak, am;
      split2<true>(v[16], v[16], spv[16], ak, am);
      if (col0) sa[C::kR2 * 16] = norm2(ak);
    }
    // Z[Nz - k] comes from the partner lane slot by slot, consumed at once (nothing is overwritten: no ordering constraint)
    static_for<0, 16>([&](auto sc) {
      constexpr int s = decltype(sc)::value;
      const PC zr = pc_shfl(pc_sel(col0, v[(32 - s) & 31], v[31 - s]), partner);
      PC ak, am;
      split2<(s >= 8)>(v[s], zr, spv[s], ak, am);
      sa[C::kR2 * s] = norm2(ak);
      if constexpr (s == 0) sb0[0] = norm2(am);
      else sb[-C::kR2 * s] = norm2(am);
    });
    SB200_HANDOFF_MBAR(mbar_arrive(full));
    pair_bar_sync(w);
  }
}

// ---- epilogue warp: powers -> scaled magnitudes (coalesced row stores), sqrt -> banded mel -> scale -> stores ----
template <int N>
__device__ __forceinline__ void feat3_epilogue(const PlanDev& p, const FeatArgs& a, Smem3<N>& sm, int w, int lane) {
  using C = Fft2Cfg<N>;
  pf* pbuf = sm.pbufs + w * kFeat3PbufElems;
  const unsigned full = smem_u32(sm.bar + w), empty = smem_u32(sm.bar + kFeat3Pairs + w);
  const bool want_mag = a.mag != nullptr, want_mel = a.mel != nullptr;
  const bool logmag = a.mag_scale.log != 0;
  // squared-magnitude form of the log scale: a log2 max(floor, sqrt p) + b = a/2 log2 max(floor^2, p) + b
  const float mag_a = 0.5f * a.mag_scale.a, mag_b = a.mag_scale.b, mag_fl = a.mag_scale.floor * a.mag_scale.floor;
#if defined(SB200_MEL_ELL) || !defined(SB200_MEL_SEG_SQRT)
  const int c_lo = want_mel ? p.mel_kmin / 32 : 1 << 30, c_hi = want_mel ? p.mel_kmax / 32 : -1;   // chunks the mel filters read
#else
  constexpr int c_lo = 1 << 30, c_hi = -1;   // segment mel taking the square roots itself: nothing is written back
#endif
  const long long warps_total = static_cast<long long>(gridDim.x) * kFeat3Pairs;
  unsigned round = 0;
  for (long long item = static_cast<long long>(blockIdx.x) * kFeat3Pairs + w; item < a.bd.total_items; item += warps_total, ++round) {
    const Item it = decode_item(a.bd, item, C::kFrames);
    pair_bar_sync(w);
    SB200_HANDOFF_MBAR(mbar_wait(full, round & 1));
    pair_bar_sync(w);
#ifdef SB200_ABLATE_EPILOGUE   // ablation builds (tools/variants.py): what the analysis role costs on its own
    if (a.mag == reinterpret_cast<float*>(1)) a.mag[0] = plo(pbuf[lane]);
    SB200_HANDOFF_MBAR(mbar_arrive(empty));
    continue;
#endif
#pragma unroll
    for (int q = 0; q < C::kP; ++q) {
      const int fA = it.t0 + 2 * q;
      const bool stA = want_mag && fA < it.T, stB = want_mag && fA + 1 < it.T;
      float* const rowA = a.mag + (it.frame_base + fA) * C::kF + lane;
      pf* const src = pbuf + q * C::kNz + lane;
      // 4 chunks of 32 bins per trip, all loads first: the write-back of the square roots may alias later loads as far
      // as the compiler can tell, so loads written after a store are not hoisted
#pragma unroll 1
      for (int j0 = 0; j0 < C::kNz / 32; j0 += 4) {
        pf pw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) pw[u] = src[32 * (j0 + u)];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + u;
          const bool band = (j >= c_lo && j <= c_hi);
          if (logmag) {
            const pf o = fma2s(pk(fast_lg2(fmaxf(mag_fl, plo(pw[u]))), fast_lg2(fmaxf(mag_fl, phi(pw[u])))), mag_a, pk(mag_b, mag_b));
            if (stA) rowA[32 * j] = plo(o);
            if (stB) rowA[C::kF + 32 * j] = phi(o);
            if (band) src[32 * j] = sqrt2(pw[u]);
          } else {
            const pf o = sqrt2(pw[u]);
            if (stA) rowA[32 * j] = plo(o);
            if (stB) rowA[C::kF + 32 * j] = phi(o);
            if (band) src[32 * j] = o;
          }
        }
      }
    }
    if (want_mag && lane < C::kP) {   // Nyquist bins
      const int fA = it.t0 + 2 * lane;
      const pf pw = pbuf[C::kP * C::kNz + lane];
      const pf o = logmag ? fma2s(pk(fast_lg2(fmaxf(mag_fl, plo(pw))), fast_lg2(fmaxf(mag_fl, phi(pw)))), mag_a, pk(mag_b, mag_b))
                          : sqrt2(pw);
      float* const row = a.mag + (it.frame_base + fA) * C::kF + C::kNz;
      if (fA < it.T) row[0] = plo(o);
      if (fA + 1 < it.T) row[C::kF] = phi(o);
    }
    __syncwarp();
#ifdef SB200_ABLATE_MEL
    const bool mel_on = false;
#else
    const bool mel_on = want_mel;
#endif
#ifdef SB200_MEL_ELL
    if (mel_on) {
#pragma unroll
      for (int rd = 0; rd < kMaxMelRounds; ++rd) {
        if (rd < p.mel_rounds) {
          const int slot = sm.mel_lo[rd * 32 + lane];
          const int m = slot >> 16, lo = slot & 0xffff;
          const float* wr = sm.melw + p.mel_round_off[rd] + lane;
          const pf* sr = pbuf + lo;   // reads may run past the row end (zero weights) into finite stale data
          const int n = p.mel_round_len[rd];   // multiple of 8
          pf acc[C::kP][4];
#pragma unroll
          for (int q = 0; q < C::kP; ++q)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[q][j] = 0ull;
          constexpr int kChunk = C::kP == 1 ? 8 : 4;   // loads in flight per trip (round lengths are multiples of 8)
#pragma unroll 1
          for (int i0 = 0; i0 < n; i0 += kChunk) {
            float wv[kChunk];
            pf sv[C::kP][kChunk];
#pragma unroll
            for (int j = 0; j < kChunk; ++j) wv[j] = wr[(i0 + j) * 32];
#pragma unroll
            for (int q = 0; q < C::kP; ++q)
#pragma unroll
              for (int j = 0; j < kChunk; ++j) sv[q][j] = sr[q * C::kNz + i0 + j];
#pragma unroll
            for (int q = 0; q < C::kP; ++q)
#pragma unroll
              for (int j = 0; j < kChunk; ++j) acc[q][j & 3] = fma2s(sv[q][j], wv[j], acc[q][j & 3]);
          }
          if (m < p.n_mel) {
#pragma unroll
            for (int q = 0; q < C::kP; ++q) {
              const pf s = add2(add2(acc[q][0], acc[q][1]), add2(acc[q][2], acc[q][3]));
              const int f = it.t0 + 2 * q;
              float* dst = a.mel + (it.frame_base + f) * p.n_mel + m;
              if (f < it.T) dst[0] = apply_scale(a.mel_scale, plo(s));
              if (f + 1 < it.T) dst[p.n_mel] = apply_scale(a.mel_scale, phi(s));
            }
          }
        }
      }
    }
#else
    // Mel by segments (PlanDev::seg_slot): lane l of round r owns segment j = 32 r + l, the bins between filter centres c_j and
    // c_{j+1}.  It reads every bin of its segment ONCE (the power itself: the square root is taken here, nothing is written
    // back to the buffer and no weight is loaded), accumulating A = sum S_k and B = sum (k - k0) S_k; then R = r0 A + a B is
    // what the segment gives the rising filter j and E = A - R what it gives the falling filter j - 1, and
    // mel[m] = enorm_m (R_m + E_{m+1}) with E_{m+1} from the next lane (rounds run from the last to the first, the first lane
    // of the later round hands its E to lane 31 of the earlier one).  Trips before the segment starts (the d bins of bank
    // alignment) and after it ends are masked by predication, so whatever lies there never enters the sums.
    if (mel_on) {
      pf carryE[C::kP];   // E of segment 32 (r + 1), from round r + 1
#pragma unroll
      for (int q = 0; q < C::kP; ++q) carryE[q] = 0ull;
#pragma unroll
      for (int rdi = 0; rdi < kMaxSegRounds; ++rdi) {
        const int rd = p.seg_rounds - 1 - rdi;
        if (rd >= 0) {
          const int4 slot = sm.seg_slot[rd * 32 + lane];
          const float2 cf = sm.seg_coef[rd * 32 + lane];
          const pf* sr = pbuf + slot.x;
          const int n = p.seg_round_len[rd];   // multiple of 4
          const unsigned d = static_cast<unsigned>(slot.y), len = static_cast<unsigned>(slot.z);
          pf accA[C::kP], accB[C::kP];
#pragma unroll
          for (int q = 0; q < C::kP; ++q) accA[q] = accB[q] = 0ull;
          float kk = -static_cast<float>(slot.y);
#pragma unroll 1
          for (int i0 = 0; i0 < n; i0 += 4) {
            pf pv[C::kP][4];
#pragma unroll
            for (int q = 0; q < C::kP; ++q)
#pragma unroll
              for (int j = 0; j < 4; ++j) pv[q][j] = sr[q * C::kNz + i0 + j];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const bool in = static_cast<unsigned>(i0 + j) - d < len;
#pragma unroll
              for (int q = 0; q < C::kP; ++q) {
#ifdef SB200_MEL_SEG_SQRT
                const pf sq = sqrt2(pv[q][j]);
#else
                const pf sq = pv[q][j];   // the log loop above has written the magnitudes of the band back in place
#endif
                if (in) {
                  accA[q] = add2(accA[q], sq);
                  accB[q] = fma2s(sq, kk, accB[q]);
                }
              }
              kk += 1.f;
            }
          }
          const int m = rd * 32 + lane;   // filter m rises in segment m and falls in segment m + 1
#pragma unroll
          for (int q = 0; q < C::kP; ++q) {
            const pf R = fma2s(accB[q], cf.y, mul2s(accA[q], cf.x));
            const pf E = sub2(accA[q], R);
            pf En = __shfl_down_sync(kFullMask, E, 1);
            if (lane == 31) En = carryE[q];
            carryE[q] = __shfl_sync(kFullMask, E, 0);
            if (m < p.n_mel) {
              const pf s = mul2s(add2(R, En), sm.enorm[m]);
              const int f = it.t0 + 2 * q;
              float* dst = a.mel + (it.frame_base + f) * p.n_mel + m;
              if (f < it.T) dst[0] = apply_scale(a.mel_scale, plo(s));
              if (f + 1 < it.T) dst[p.n_mel] = apply_scale(a.mel_scale, phi(s));
            }
          }
        }
      }
    }
#endif
    SB200_HANDOFF_MBAR(mbar_arrive(empty));
  }
}

template <int N, bool PRE, int HS, bool TMA>
__global__ void __launch_bounds__(kFeat3Threads, 1) stft_feature3_kernel(const PlanDev p, const FeatArgs a,
                                                                          const __grid_constant__ CUtensorMap tmap,
                                                                          const StageArgs sa) {
  extern __shared__ __align__(128) unsigned char smem_raw[];   // bulk-tensor copies land in 128-byte aligned shared memory
  Smem3<N> sm;
  sm.carve(smem_raw, p);
#ifndef SB200_BISECT_NO_FILL   // (sanitizer bisection variants: results are garbage, only the tools' reports matter)
  sm.fill(p, a.mel != nullptr);
#endif
#ifdef SB200_BISECT_SYNC_BEFORE_INIT
  __syncthreads();
#endif
  if (threadIdx.x < 3 * kFeat3Pairs) {   // full / empty: 32 lane arrivals; staged: the one expect_tx arrival of the issuing lane
    mbar_init(smem_u32(sm.bar + threadIdx.x), threadIdx.x < 2 * kFeat3Pairs ? 32 : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // SB200_NO_SETMAXNREG (sanitizer variant, tools/variants.py): no register re-partitioning, the whole kernel is compiled for
  // the 128 registers per thread of the launch (the analysis role spills).  compute-sanitizer patches the kernel with code of its
  // own that needs registers too: with setmaxnreg in the kernel its synccheck / racecheck lose track of the mbarriers.
  if (warp < kFeat3Pairs) {
#ifndef SB200_NO_SETMAXNREG
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kFeat3AnalysisRegs));
#endif
    feat3_analysis<N, PRE, HS, TMA>(p, a, sm, warp, lane, &tmap, sa);
  } else {
#ifndef SB200_NO_SETMAXNREG
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kFeat3EpilogueRegs));
#endif
    feat3_epilogue<N>(p, a, sm, warp - kFeat3Pairs, lane);
  }
}

template <int N>
inline size_t feat3_smem_bytes(const PlanDev& p) {
  return Smem3<N>::bytes(p);
}

}  // namespace sb200
