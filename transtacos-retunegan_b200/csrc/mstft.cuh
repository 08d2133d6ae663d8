// Multi-resolution STFT loss, forward and backward (retunegan/models/loss.py:22-62,
// retunegan/audio.py:150-170; backward = the autograd graph of retunegan/train.py:192, closed form in
// SURVEY.md 8a row L2).
//
// All resolutions of a phase run as ONE launch (mstft_multi_*_kernel below: the per-resolution grids concatenated, a CTA works
// for the resolution its index falls into, one pass per warp).
// Forward, per resolution: a warp analyses the same Q frames of y and of y_g back to back,
// keeps the real mel rows in registers, and accumulates |M - M_g| + |ln M - ln M_g| into a per-warp
// partial (deterministic two-stage reduction, no atomics).  Optionally writes the [B,2,T',F]
// (ln|D+1e-9|, angle(D)/PI) stacks for the STFT discriminator.
// Backward, per resolution: recomputes the y_g analysis (cheaper than saving 8 B/bin), forms
// gD = gS (D+1e-9)/S + (gP/PI) i D/|D|^2 with gS = basis^T gM + g_lnS/S on the Hermitian pairs in
// registers, runs the adjoint of the one-sided rFFT through the inverse engine, windows, and stores
// gradient frames; grad_ola_kernel then overlap-adds all resolutions as a gather and folds the
// reflect padding back.  The loss-only training step is one fused launch (value + gradient, no recomputation) + the overlap-add,
// whose extra block reduces the loss partial sums and, under DDP, averages the loss over the ranks of the box in place
// (peer_allreduce_mean: NVLink peer memory, no collective call).
#pragma once
#include "feat.cuh"

namespace sb200 {

constexpr float kRefPI = 3.14159265358979f;   // retunegan/utils.py:12
constexpr int kMstftWarps = 8;
constexpr int kMaxRes = 4;
// unroll factor of the rolled Hermitian-pair / mel-tap loops of the backward kernels (independent iterations in flight per warp
// against instruction-cache footprint: every warp runs the body once)
#ifndef SB200_MSTFT_UNROLL
#define SB200_MSTFT_UNROLL 2
#endif
constexpr int kMstftUnroll = SB200_MSTFT_UNROLL;

// atan2f to ~3e-7 rad: octant reduction with one approximate division, degree-7 minimax polynomial in a^2 (fitted in double,
// evaluated in float: max error 1.4e-7 on [0, 1]), same results as atan2f at the axes; atan2(+-0, -0) is taken as +-0.  The library atan2f
// (IEEE division with a slow path) and logf were a seventh of the forward kernel's instructions at two calls per bin.
__device__ __forceinline__ float fast_atan2(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float a = __fdividef(mn, fmaxf(mx, 1e-37f));   // 0 / 1e-37 = 0 at the origin
  const float s = a * a;
  float r = -0.0040545654483139515f;
  r = fmaf(r, s, 0.021862952038645744f);
  r = fmaf(r, s, -0.0559123195707798f);
  r = fmaf(r, s, 0.0964219719171524f);
  r = fmaf(r, s, -0.1390862911939621f);
  r = fmaf(r, s, 0.19946566224098206f);
  r = fmaf(r, s, -0.33329859375953674f);
  r = fmaf(r, s, 0.9999993443489075f);
  r *= a;
  if (ay > ax) r = 1.57079632679489662f - r;
  if (x < 0.f) r = 3.14159265358979324f - r;   // not the sign bit: a zero spectrum (digital silence) has phase 0 whatever the sign
                                               // of its zeros, as torch.angle of the reference's all-(+0) torch.stft output
  return copysignf(r, y);
}

struct MstftFwdArgs {
  const float* y;
  const float* yg;
  BatchDev bd;          // uniform [B, T]
  int Tf;               // frames per row
  float* spec_r;        // [B, 2, Tf, F] or null
  float* spec_g;        // [B, 2, Tf, F] or null
  int phd_phase;
  float* mel_r;         // [B*Tf, n_mel] saved for backward (never null)
  float* partials;      // [gridDim.x * kMstftWarps]
  int want_loss;
};

// RAW (get_stft_torch, retunegan/audio.py:166-168): ch0 receives S = |D + 1e-9| itself and ch1 angle(D), instead of ln S and angle / PI
template <int N, bool RAW = false>
__device__ __forceinline__ void mstft_analyse(const PlanDev& p, const SmemTables<N>& sm, float2* buf, float2 (&v)[32],
                                              const float* x, long long L, int t0, int T, int lane, float* ch0,
                                              float* ch0_alias, float* ch1, long long row0 /* (b*2*Tf + t0) * F */,
                                              long long ch_stride /* Tf*F */, unsigned long long* tbl_bar = nullptr) {
  // analysis of Q frames; leaves S = |D + 1e-9| in buf[q*ZS + k].x; optionally writes ln S (ch0, ch0_alias) and angle/PI (ch1)
  // tbl_bar: barriers of an asynchronous table fill ([0] window + FFT twiddles, [1] the rest), waited on at first use
  using C = FftCfg<N>;
  if (tbl_bar) tbl_wait(tbl_bar);
  load_frames<N, false>(v, x, L, t0, T, p.hop, 0.f, sm.win, lane);
  fft_forward<N>(v, buf, sm.tw, lane);
  if (tbl_bar) tbl_wait(tbl_bar + 1);
  const int rk = lane & 3, rm = (4 - rk) & 3;
  static_for<0, C::kQ>([&](auto qc) {
    constexpr int q = decltype(qc)::value;
    if (t0 + q < T) {
      float2* zq = buf + q * C::kZS;
      const long long row = row0 + static_cast<long long>(q) * C::kF;
      auto emit = [&](float2 X, int k, float S) {
        if (ch0) ch0[row + k] = RAW ? S : __logf(S);
        if (ch0_alias) ch0_alias[row + k] = __logf(S);
        if (ch1) ch1[row + ch_stride + k] = RAW ? fast_atan2(X.y, X.x) : fast_atan2(X.y, X.x) * (1.f / kRefPI);
      };
      if (!RAW && ch0 != nullptr && ch1 != nullptr && ch0_alias == nullptr) {
        // the training step's case (both channels of one stack): no per-bin pointer tests, the quarter-turn (-i)^k as a
        // multiplication by the lane's constant (c, s) in {(1,0), (0,1), (-1,0), (0,-1)} instead of selects, two row pointers.
        // Same values as the general loop below (products with 0 and +-1 are exact; only the sign of a zero can differ).
        float* const p0 = ch0 + row;
        float* const p1 = ch1 + row + ch_stride;
        const float ck = (rk == 0) ? 1.f : (rk == 2 ? -1.f : 0.f), sk_ = (rk == 1) ? 1.f : (rk == 3 ? -1.f : 0.f);
        const float cm = (rm == 0) ? 1.f : (rm == 2 ? -1.f : 0.f), sm_ = (rm == 1) ? 1.f : (rm == 3 ? -1.f : 0.f);
#pragma unroll 2
        for (int i = 0; i < C::kPairIters; ++i) {
          const int k = lane + 32 * i;
          const int km = (C::kNz - k) & (C::kNz - 1);
          float2 Ak, Am;
          split_fwd(zq[k], zq[km], sm.ws[k], Ak, Am);
          const float2 Xk = make_float2(fmaf(sk_, Ak.y, ck * Ak.x), fmaf(-sk_, Ak.x, ck * Ak.y));
          const float2 Xm = make_float2(fmaf(sm_, Am.y, cm * Am.x), fmaf(-sm_, Am.x, cm * Am.y));
          const float rek = Xk.x + 1e-9f, rem = Xm.x + 1e-9f;
          const float sk = fast_sqrt(fmaf(rek, rek, Xk.y * Xk.y)), smg = fast_sqrt(fmaf(rem, rem, Xm.y * Xm.y));
          zq[k].x = sk;
          if (k != 0) zq[km].x = smg;
          p0[k] = __logf(sk);
          p0[C::kNz - k] = __logf(smg);
          p1[k] = fast_atan2(Xk.y, Xk.x) * (1.f / kRefPI);
          p1[C::kNz - k] = fast_atan2(Xm.y, Xm.x) * (1.f / kRefPI);
        }
      } else {
#pragma unroll 2
      for (int i = 0; i < C::kPairIters; ++i) {
        const int k = lane + 32 * i;
        const int km = (C::kNz - k) & (C::kNz - 1);
        float2 Ak, Am;
        split_fwd(zq[k], zq[km], sm.ws[k], Ak, Am);
        const float2 Xk = rot_fwd(Ak, rk), Xm = rot_fwd(Am, rm);
        const float rek = Xk.x + 1e-9f, rem = Xm.x + 1e-9f;
        const float sk = fast_sqrt(fmaf(rek, rek, Xk.y * Xk.y)), smg = fast_sqrt(fmaf(rem, rem, Xm.y * Xm.y));
        zq[k].x = sk;
        if (k != 0) zq[km].x = smg;
        emit(Xk, k, sk);
        emit(Xm, C::kNz - k, smg);
      }
      }
      if (lane == 0) {
        constexpr int k = C::kNz / 2;
        float2 Ak, Am;
        split_fwd(zq[k], zq[k], sm.ws[k], Ak, Am);
        const float2 Xk = rot_fwd(Ak, k);
        const float rek = Xk.x + 1e-9f;
        const float sk = fast_sqrt(fmaf(rek, rek, Xk.y * Xk.y));
        zq[k].x = sk;
        emit(Xk, k, sk);
      }
    }
  });
  __syncwarp();
}

// vb / nvb: index of this CTA among the CTAs working on this resolution and their number (the per-resolution kernel passes
// blockIdx.x / gridDim.x; the all-resolution kernel deals its CTAs to the resolutions round-robin).
template <int N>
__device__ __forceinline__ void mstft_fwd_body(const PlanDev& p, const MstftFwdArgs& a, unsigned char* smem_raw, int vb, int nvb) {
  using C = FftCfg<N>;
  SmemTables<N> sm;
  sm.carve(smem_raw, p);
  __shared__ unsigned long long tbl_bar[2];   // [0] window + FFT twiddles, [1] split twiddles + mel tables
  if (threadIdx.x == 0) {
    tbl_bar_init(tbl_bar, blockDim.x);
    tbl_bar_init(tbl_bar + 1, blockDim.x);
  }
  __syncthreads();
  {
    CopySeg seg[5];
    const int ns = sm.segments(p, p.window, true, seg);
    copy_segments_async(seg, 0, 2, tbl_bar);
    copy_segments_async(seg, 2, ns, tbl_bar + 1);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2* buf = sm.bufs + warp * C::kBufF2;
  float acc = 0.f;
  const long long chs = static_cast<long long>(a.Tf) * C::kF;
  // an item is 2Q frames (packed-engine granularity); this engine takes it as two independent Q-frame passes
  for (long long sub = static_cast<long long>(vb) * kMstftWarps + warp; sub < 2 * a.bd.total_items;
       sub += static_cast<long long>(nvb) * kMstftWarps) {
    Item it = decode_item(a.bd, sub >> 1, 2 * C::kQ);
    it.t0 += static_cast<int>(sub & 1) * C::kQ;
    if (it.t0 < it.T) {
    const long long row0 = (static_cast<long long>(it.b) * 2 * a.Tf + it.t0) * C::kF;
    float2 v[32];
    float mr[kMaxMelRounds][C::kQ];
    // real, then generated signal through ONE copy of the analysis code (rolled on purpose: every warp runs this body once, so
    // the kernel's time is the instruction fetch of its straight-line code; profiles/r01_mstft_gl2_ncu_summary.md)
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
      const float* x = side ? a.yg : a.y;
      float* ch0 = side ? (a.phd_phase ? nullptr : a.spec_g) : a.spec_r;
      float* ch0_alias = side ? nullptr : (a.phd_phase ? a.spec_g : nullptr);
      float* ch1 = side ? a.spec_g : a.spec_r;
      mstft_analyse<N>(p, sm, buf, v, x + it.sig_base, it.L, it.t0, it.T, lane, ch0, ch0_alias, ch1, row0, chs, tbl_bar);
      mel_project_smem<N>(p, sm.melw, sm.mel_lo, buf, lane, [&](int q, int rd, int m, float val) {
        if (side == 0) {
          mr[rd][q] = val;
          if (m < p.n_mel && it.t0 + q < it.T) a.mel_r[(it.frame_base + it.t0 + q) * p.n_mel + m] = val;
        } else if (a.want_loss && m < p.n_mel && it.t0 + q < it.T) {
          const float r = mr[rd][q];
          acc += fabsf(r - val) + fabsf(logf(r) - logf(val));
        }
      });
      __syncwarp();
    }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(kFullMask, acc, d);
  if (lane == 0) a.partials[vb * kMstftWarps + warp] = acc;
  asm volatile("cp.async.wait_all;" ::: "memory");   // a warp without an item never waited: its copies land before it exits
}

template <int N>
__global__ void __launch_bounds__(kMstftWarps * 32, 2) mstft_fwd_kernel(const PlanDev p, const MstftFwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  mstft_fwd_body<N>(p, a, smem_raw, blockIdx.x, gridDim.x);
}

// get_stft_torch (retunegan/audio.py:150-170) as one launch: S = |D + 1e-9| [B, Tf, F], P = angle(D) [B, Tf, F] and
// M = mel_basis S [B, Tf, n_mel], frame-major; any output may be null.
struct StftSmpArgs {
  const float* y;
  BatchDev bd;     // uniform [B, T]
  int Tf;
  float* S;
  float* P;
  float* M;
};
template <int N>
__global__ void __launch_bounds__(kMstftWarps * 32, 2) stft_smp_kernel(const PlanDev p, const StftSmpArgs a) {
  using C = FftCfg<N>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemTables<N> sm;
  sm.carve(smem_raw, p);
  {
    CopySeg seg[5];
    copy_segments(seg, sm.segments(p, p.window, a.M != nullptr, seg));
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2* buf = sm.bufs + warp * C::kBufF2;
  for (long long sub = static_cast<long long>(blockIdx.x) * kMstftWarps + warp; sub < 2 * a.bd.total_items;
       sub += static_cast<long long>(gridDim.x) * kMstftWarps) {
    Item it = decode_item(a.bd, sub >> 1, 2 * C::kQ);
    it.t0 += static_cast<int>(sub & 1) * C::kQ;
    if (it.t0 >= it.T) continue;
    float2 v[32];
    const long long row0 = (it.frame_base + it.t0) * C::kF;
    mstft_analyse<N, true>(p, sm, buf, v, a.y + it.sig_base, it.L, it.t0, it.T, lane, a.S, nullptr, a.P, row0, 0);
    if (a.M) {
      mel_project_smem<N>(p, sm.melw, sm.mel_lo, buf, lane, [&](int q, int, int m, float val) {
        if (m < p.n_mel && it.t0 + q < it.T) a.M[(it.frame_base + it.t0 + q) * p.n_mel + m] = val;
      });
    }
    __syncwarp();
  }
}

// ---- loss reduction over the ranks of one box, inside the reducing block (DDP) -----------------------------------------------
// Every rank owns one exchange buffer (PeerBuf, cudaMalloc + CUDA IPC: all of them are mapped in every process, reached over
// NVLink / NVSwitch peer access).  The thread that has just reduced the rank's loss stores it into slot [rank] of EVERY rank's
// buffer, publishes it with a release store of the call's epoch into the matching flag, waits until all flags of its own
// buffer carry the epoch and adds the slots in rank order: every rank gets the same mean, bit for bit, with no extra launch and
// no host call.  The epoch is a counter in the rank's own buffer incremented by the kernel itself (CUDA-graph replay safe);
// slots are double-buffered by epoch parity: a rank can only be one call ahead of the slowest one, because its next call waits
// for that rank's flag.  A wait that exceeds kPeerTimeoutNs gives up and yields NaN instead of hanging the GPU.
constexpr int kMaxPeers = 8;
constexpr unsigned long long kPeerTimeoutNs = 2000000000ull;
struct PeerBuf {
  unsigned epoch;
  unsigned pad[31];
  float vals[2][kMaxPeers];
  unsigned flags[2][kMaxPeers];
};
struct PeerDev {
  void* buf[kMaxPeers];
  int rank, world;   // world <= 1: no reduction
  float* out;        // device scalar: mean over the ranks
};
__device__ __forceinline__ unsigned long long peer_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void peer_allreduce_mean(const PeerDev& pd, float local) {
  PeerBuf* me = static_cast<PeerBuf*>(pd.buf[pd.rank]);
  const unsigned e = me->epoch + 1;   // written by this thread of this rank only
  me->epoch = e;
  const int par = e & 1;
  for (int q = 0; q < pd.world; ++q) {
    PeerBuf* pb = static_cast<PeerBuf*>(pd.buf[q]);
    asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(&pb->vals[par][pd.rank]), "f"(local) : "memory");
  }
  __threadfence_system();
  for (int q = 0; q < pd.world; ++q) {
    PeerBuf* pb = static_cast<PeerBuf*>(pd.buf[q]);
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&pb->flags[par][pd.rank]), "r"(e) : "memory");
  }
  const unsigned long long t0 = peer_now_ns();
  float sum = 0.f;
  bool ok = true;
  for (int q = 0; q < pd.world && ok; ++q) {
    unsigned seen;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(&me->flags[par][q]) : "memory");
      if (seen == e) break;
      if (peer_now_ns() - t0 > kPeerTimeoutNs) { ok = false; break; }
    }
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(&me->vals[par][q]) : "memory");
    sum += v;
  }
  *pd.out = ok ? sum / static_cast<float>(pd.world) : __int_as_float(0x7fc00000);
}

struct MstftFinArgs {
  int n_res;
  const float* partials[kMaxRes];
  int n_partials[kMaxRes];
  float inv_count[kMaxRes];   // 1 / (B * n_mel * Tf)
  float* loss;
  PeerDev peer;               // DDP: the rank's loss is also averaged over the box (peer.world > 1)
};
// loss = 1/n_res * sum_res (sum of partials) / count   (F.l1_loss reduction='mean', loss.py:51-54); one CTA, fixed order:
// thread i sums partials i, i + blockDim, ... of every resolution (all loads independent), warps reduce by shuffle, thread 0 adds
// the warp sums in warp order.  One barrier.
__device__ __forceinline__ void mstft_finalize_body(const MstftFinArgs& a) {
  __shared__ float red[kMaxRes][32];
  float s[kMaxRes];
#pragma unroll
  for (int r = 0; r < kMaxRes; ++r) {
    s[r] = 0.f;
    if (r < a.n_res)
      for (int i = threadIdx.x; i < a.n_partials[r]; i += blockDim.x) s[r] += __ldcg(a.partials[r] + i);
  }
#pragma unroll
  for (int r = 0; r < kMaxRes; ++r) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s[r] += __shfl_xor_sync(kFullMask, s[r], d);
    if ((threadIdx.x & 31) == 0) red[r][threadIdx.x >> 5] = s[r];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float total = 0.f;
    for (int r = 0; r < a.n_res; ++r) {
      float t = 0.f;
      for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += red[r][w];
      total += t * a.inv_count[r];
    }
    const float local = total / a.n_res;
    *a.loss = local;
    if (a.peer.world > 1) peer_allreduce_mean(a.peer, local);
  }
}
__global__ void mstft_finalize_kernel(const MstftFinArgs a) { mstft_finalize_body(a); }

struct MstftBwdArgs {
  const float* yg;
  BatchDev bd;
  int Tf;
  const float* mel_r;      // saved [B*Tf, n_mel]
  const float* g_loss;     // device scalar or null
  float loss_scale;        // 1 / (n_res * B * n_mel * Tf)
  const float* g_spec;     // [B, 2, Tf, F] upstream or null
  int phd_phase;
  float* gfb;              // [B*Tf, win] gradient frames
  const float* y;          // FUSED: real audio (its mel rows are computed here instead of read from mel_r)
  float* partials;         // FUSED: [gridDim.x * kMstftWarps] loss partial sums
  // backward of get_stft_torch (raw != 0): upstream gradients of S, P [B, Tf, F] and M [B, Tf, n_mel], any may be null
  int raw;
  const float* g_s_raw;
  const float* g_p_raw;
  const float* g_m_raw;
};

// Shared-memory bytes of mstft_bwd_kernel: the tables and per-warp FFT buffers of the forward kernel plus, per warp, the
// mel-row gradients [Q][128], and the column view of the mel basis ({c0, c1} and the first row per bin).
template <int N>
inline size_t mstft_bwd_smem_bytes(const PlanDev& p) {
  return feat_smem_bytes<N>(p) + sizeof(float) * kMstftWarps * FftCfg<N>::kQ * 128 + sizeof(float2) * ((FftCfg<N>::kF + 1) & ~1) +
         ((FftCfg<N>::kF + 15) & ~15);
}

// mel projection of |X + 1e-9| with the complex spectrum X in buf (natural bin order), magnitudes taken on the fly
template <int N, class Emit>
__device__ __forceinline__ void mel_project_cplx(const PlanDev& p, const float* __restrict__ s_melw, const int* __restrict__ s_lo,
                                                 const float2* __restrict__ buf, int lane, Emit&& emit) {
  using C = FftCfg<N>;
#pragma unroll 1
  for (int rd = 0; rd < p.mel_rounds; ++rd) {
    const int slot = s_lo[rd * 32 + lane];
    const int m = slot >> 16, lo = slot & 0xffff;
    const float* wr = s_melw + p.mel_round_off[rd] + lane;
    const int n = p.mel_round_len[rd];
    float acc[C::kQ];
#pragma unroll
    for (int q = 0; q < C::kQ; ++q) acc[q] = 0.f;
#pragma unroll kMstftUnroll
    for (int it = 0; it < n; ++it) {
      const float w = wr[it * 32];
      const int idx = min(lo + it, C::kNz - 1);
#pragma unroll
      for (int q = 0; q < C::kQ; ++q) {
        const float2 X = buf[q * C::kZS + idx];
        const float re = X.x + 1e-9f;
        acc[q] = fmaf(w, fast_sqrt(fmaf(re, re, X.y * X.y)), acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < C::kQ; ++q) emit(q, rd, m, acc[q]);
  }
}

// FUSED = loss value and its gradient in ONE pass (loss-only training steps): analyses y, then does the backward work on
// y_g with a unit upstream gradient while accumulating the loss -- no recomputation, no second launch per resolution.
//
// Every warp runs this body once per item and an item is all a warp gets at training sizes, so the kernel's time is the
// instruction fetch of its own straight-line code (profiles/r01_mstft_gl2_ncu_summary.md: 44 % no_instruction stalls at
// 14.6 k instructions).  The Hermitian-pair work (forward split, gradient of every bin, inverse split) therefore runs as
// ROLLED loops over the spectrum kept in the warp's shared-memory buffer instead of unrolled over register arrays.
template <int N, bool FUSED>
__device__ __forceinline__ void mstft_bwd_body(const PlanDev& p, const MstftBwdArgs& a, unsigned char* smem_raw, int vb, int nvb) {
  using C = FftCfg<N>;
  SmemTables<N> sm;
  sm.carve(smem_raw, p);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2* buf = sm.bufs + warp * C::kBufF2;
  float* gmbuf = reinterpret_cast<float*>(sm.bufs + kMstftWarps * C::kBufF2) + warp * C::kQ * 128;   // [Q][128] mel-row gradients
  // column view of the mel basis: gS[k] = c0[k] gM[r0[k]] + c1[k] gM[r0[k] + 1].  In shared memory: read from global inside the
  // pair loop it was the largest stall of the kernel (long scoreboard, 20 % of the warp time).
  float2* colc = reinterpret_cast<float2*>(reinterpret_cast<float*>(sm.bufs + kMstftWarps * C::kBufF2) + kMstftWarps * C::kQ * 128);
  unsigned char* colr = reinterpret_cast<unsigned char*>(colc + ((C::kF + 1) & ~1));
  __shared__ unsigned long long tbl_bar[2];   // [0] window + FFT twiddles, [1] split twiddles + mel tables (asynchronous fill)
  if (threadIdx.x == 0) {
    tbl_bar_init(tbl_bar, blockDim.x);
    tbl_bar_init(tbl_bar + 1, blockDim.x);
  }
  __syncthreads();
  {
    CopySeg seg[7];
    const int ns = sm.segments(p, p.window, true, seg);
    seg[ns] = copy_seg(colc, p.col_c01, (C::kF + 1) & ~1);
    seg[ns + 1] = copy_seg(colr, p.col_r8, (C::kF + 15) & ~15);
    copy_segments_async(seg, 0, 2, tbl_bar);
    copy_segments_async(seg, 2, ns + 2, tbl_bar + 1);
  }
  // bin Nz (Nyquist) of frame q: the pad slot behind the row where rows are padded, else behind the last row
  auto nyq = [](int q) { return C::kZS > C::kNz ? q * C::kZS + C::kNz : C::kQ * C::kZS + q; };
  static_assert(C::kZS > C::kNz || C::kQ * C::kZS + C::kQ <= C::kBufF2, "no room for the Nyquist bins");
  const int rk = lane & 3, rm = (4 - rk) & 3;
  // quarter turns as constants: (-i)^j (x + i y) = (c x + s y) + i (c y - s x), i^j (x + i y) = (c x - s y) + i (c y + s x)
  const float rck = (rk == 0) ? 1.f : (rk == 2 ? -1.f : 0.f), rsk = (rk == 1) ? 1.f : (rk == 3 ? -1.f : 0.f);
  const float rcm = (rm == 0) ? 1.f : (rm == 2 ? -1.f : 0.f), rsm = (rm == 1) ? 1.f : (rm == 3 ? -1.f : 0.f);
  const float gl = FUSED ? a.loss_scale : (a.g_loss ? __ldg(a.g_loss) * a.loss_scale : 0.f);
  float loss_acc = 0.f;
  const long long chs = static_cast<long long>(a.Tf) * C::kF;
  for (long long sub = static_cast<long long>(vb) * kMstftWarps + warp; sub < 2 * a.bd.total_items;
       sub += static_cast<long long>(nvb) * kMstftWarps) {
    Item it = decode_item(a.bd, sub >> 1, 2 * C::kQ);
    it.t0 += static_cast<int>(sub & 1) * C::kQ;
    if (it.t0 < it.T) {
    float2 v[32];
    float mr[kMaxMelRounds][C::kQ];
    float gm[kMaxMelRounds][C::kQ];
#pragma unroll
    for (int rd = 0; rd < kMaxMelRounds; ++rd)
#pragma unroll
      for (int q = 0; q < C::kQ; ++q) { gm[rd][q] = 0.f; mr[rd][q] = 1.f; }
    // FUSED: the real signal (side 0: only its mel rows are kept), then the generated one (side 1) through ONE copy of the
    // analysis code.  Not fused: side 1 only.
#pragma unroll 1
    for (int side = FUSED ? 0 : 1; side < 2; ++side) {
      tbl_wait(tbl_bar);
      load_frames<N, false>(v, (side ? a.yg : a.y) + it.sig_base, it.L, it.t0, it.T, p.hop, 0.f, sm.win, lane);
      fft_forward<N>(v, buf, sm.tw, lane);
      tbl_wait(tbl_bar + 1);
      // forward split in place: Z[k], Z[Nz-k] -> X[k], X[Nz-k] (bin Nz to its own slot)
#pragma unroll 1
      for (int q = 0; q < C::kQ; ++q) {
        float2* zq = buf + q * C::kZS;
#pragma unroll kMstftUnroll
        for (int i = 0; i < C::kPairIters; ++i) {
          const int k = lane + 32 * i;
          const int km = (C::kNz - k) & (C::kNz - 1);
          float2 Ak, Am;
          split_fwd(zq[k], zq[km], sm.ws[k], Ak, Am);
          const float2 xk = make_float2(fmaf(rsk, Ak.y, rck * Ak.x), fmaf(-rsk, Ak.x, rck * Ak.y));
          const float2 xm = make_float2(fmaf(rsm, Am.y, rcm * Am.x), fmaf(-rsm, Am.x, rcm * Am.y));
          if (FUSED && side == 0) {   // the real signal: only |X + 1e-9| is needed (the Nyquist bin carries no mel weight)
            const float rek = xk.x + 1e-9f, rem = xm.x + 1e-9f;
            zq[k].x = fast_sqrt(fmaf(rek, rek, xk.y * xk.y));
            if (k != 0) zq[km].x = fast_sqrt(fmaf(rem, rem, xm.y * xm.y));
          } else {
            zq[k] = xk;
            if (k != 0) zq[km] = xm;
            else buf[nyq(q)] = xm;
          }
        }
        if (lane == 0) {
          constexpr int k = C::kNz / 2;
          float2 Ak, Am;
          split_fwd(zq[k], zq[k], sm.ws[k], Ak, Am);
          const float2 xk = rot_fwd(Ak, k);
          if (FUSED && side == 0) {
            const float rek = xk.x + 1e-9f;
            zq[k].x = fast_sqrt(fmaf(rek, rek, xk.y * xk.y));
          } else {
            zq[k] = xk;
          }
        }
      }
      __syncwarp();
      // side 0: mel of the real signal from the stored magnitudes (one square root per bin instead of one per filter tap)
      if (FUSED && side == 0) {
        mel_project_smem<N>(p, sm.melw, sm.mel_lo, buf, lane, [&](int q, int rd, int, float val) { mr[rd][q] = val; });
        __syncwarp();
        continue;
      }
      // side 1: mel of the generated signal -> gradient of the loss w.r.t. each mel row
      mel_project_cplx<N>(p, sm.melw, sm.mel_lo, buf, lane, [&](int q, int rd, int m, float mg) {
#pragma unroll
        for (int r2 = 0; r2 < kMaxMelRounds; ++r2) {
          if (r2 != rd) continue;
          if (side == 0) {
            mr[r2][q] = mg;
          } else {
            float g = 0.f;
            if (a.raw) {
              if (a.g_m_raw && m < p.n_mel && it.t0 + q < it.T) g = __ldg(a.g_m_raw + (it.frame_base + it.t0 + q) * p.n_mel + m);
            } else if (gl != 0.f && m < p.n_mel && it.t0 + q < it.T) {
              float r;
              if constexpr (FUSED) {
                r = mr[r2][q];
                loss_acc += fabsf(r - mg) + fabsf(logf(r) - logf(mg));
              } else {
                r = __ldg(a.mel_r + (it.frame_base + it.t0 + q) * p.n_mel + m);
              }
              const float sgn = (mg > r) ? 1.f : ((mg < r) ? -1.f : 0.f);
              g = gl * (sgn + __fdividef(sgn, mg));
            }
            gm[r2][q] = g;
          }
        }
      });
      if (side == 0) __syncwarp();
    }
    // mel-row gradients by row (their own buffer: the spectrum stays in buf)
#pragma unroll
    for (int rd = 0; rd < kMaxMelRounds; ++rd)
#pragma unroll
      for (int q = 0; q < C::kQ; ++q) {
        const int row = (rd < p.mel_rounds) ? (sm.mel_lo[rd * 32 + lane] >> 16) : 0x7fff;
        if (row < 128) gmbuf[q * 128 + row] = gm[rd][q];
      }
    __syncwarp();
    // gD of one bin, scaled for the adjoint: interior bins / 2, DC and Nyquist real.  (r0, c0, c1): column view of the mel basis at
    // bin k; (us, up): upstream gradients of the bin's magnitude / phase outputs (not FUSED: spec stacks or get_stft_torch outputs).
    const bool s_div = !a.raw;                              // spec stacks hold ln S: d ln S / dS = 1 / S
    const float p_scale = a.raw ? 1.f : 1.f / kRefPI;       // and angle / PI
    auto grad_bin = [&](float2 X, int q, float half, int r0, float c0, float c1, float us, float up) -> float2 {
      float2 G;
      {   // (the frame is valid: frames past the end of the utterance are zero-filled by the caller)
        const float re = X.x + 1e-9f;
        const float S = fast_sqrt(fmaf(re, re, X.y * X.y));
        float gS = fmaf(c0, gmbuf[q * 128 + r0], c1 * gmbuf[q * 128 + r0 + 1]);
        float gP = 0.f;
        if constexpr (!FUSED) {
          gS += s_div ? __fdividef(us, S) : us;
          gP = up * p_scale;
        }
        const float d2 = fmaf(X.x, X.x, X.y * X.y);
        // abs / angle backward are 0 at 0.  div.approx (2 ulp) instead of the IEEE division: its slow-path subroutine was 10 % of
        // the warp time, and the gradient tolerance is 1e-4
        const float gs = S > 0.f ? __fdividef(gS, S) : 0.f, gp = (!FUSED && d2 > 0.f) ? __fdividef(gP, d2) : 0.f;
        // gS (X + 1e-9)/S + gP i X / |X|^2
        G = make_float2(half * fmaf(gs, re, -gp * X.y), half * fmaf(gs, X.y, gp * X.x));
      }
      return G;
    };
    // upstream gradients of the pair (k, Nz - k), k = lane + 32 i, of frame q: loaded TWO iterations ahead of their use (they come
    // from L2 / HBM, and inside the iteration their latency sat on the dependent chain of the pair)
    struct Up { float sk, pk, sm, pm; };
    const float* ups = nullptr;   // rows of the current frame in the upstream magnitude-type / phase-type gradients (null: none)
    const float* upp = nullptr;
    auto set_up_rows = [&](int q) {
      ups = upp = nullptr;
      if constexpr (!FUSED) {
        const long long t = it.t0 + q;
        if (a.raw) {
          const long long idx = (it.frame_base + t) * C::kF;
          if (a.g_s_raw) ups = a.g_s_raw + idx;
          if (a.g_p_raw) upp = a.g_p_raw + idx;
        } else if (a.g_spec) {
          const long long idx = (static_cast<long long>(it.b) * 2 * a.Tf + t) * C::kF;
          if (!a.phd_phase) ups = a.g_spec + idx;
          upp = a.g_spec + idx + chs;
        }
      }
    };
    auto load_up = [&](int k) -> Up {
      Up u{0.f, 0.f, 0.f, 0.f};
      if constexpr (!FUSED) {
        if (k < C::kNz) {
          if (ups) { u.sk = __ldg(ups + k); u.sm = __ldg(ups + C::kNz - k); }
          if (upp) { u.pk = __ldg(upp + k); u.pm = __ldg(upp + C::kNz - k); }
        }
      }
      return u;
    };
    // gradient of one Hermitian pair and inverse split, in place: X[k], X[Nz-k] -> Z'[k], Z'[Nz-k]
    auto pair_iter = [&](float2* zq, int q, int i, const Up& u) {
      const int k = lane + 32 * i;
      const int km = (C::kNz - k) & (C::kNz - 1);
      const float half = k == 0 ? 1.f : 0.5f;
      const int r0k = colr[k], r0m = colr[C::kNz - k];
      const float2 ck = colc[k], cm = colc[C::kNz - k];
      const float2 Gk = grad_bin(zq[k], q, half, r0k, ck.x, ck.y, u.sk, u.pk);
      const float2 Gm = grad_bin(k != 0 ? zq[km] : buf[nyq(q)], q, half, r0m, cm.x, cm.y, u.sm, u.pm);
      // i^k G as a product with the lane's constant quarter turn (c, s): no selects
      float2 Bk = make_float2(fmaf(-rsk, Gk.y, rck * Gk.x), fmaf(rsk, Gk.x, rck * Gk.y));
      float2 Bm = make_float2(fmaf(-rsm, Gm.y, rcm * Gm.x), fmaf(rsm, Gm.x, rcm * Gm.y));
      if (k == 0) { Bk.y = 0.f; Bm.y = 0.f; }
      float2 Zk, Zr;
      split_inv(Bk, Bm, sm.ws[k], Zk, Zr);
      if (k != 0) zq[km] = Zr;
      zq[k] = Zk;
    };
#pragma unroll 1
    for (int q = 0; q < C::kQ; ++q) {
      float2* zq = buf + q * C::kZS;
      if (it.t0 + q >= it.T) {   // frame past the end of the utterance: no gradient (the inverse transform of zeros)
#pragma unroll 1
        for (int i = 0; i < C::kPairIters; ++i) {
          const int k = lane + 32 * i;
          zq[k] = make_float2(0.f, 0.f);
          zq[(C::kNz - k) & (C::kNz - 1)] = make_float2(0.f, 0.f);
        }
        if (lane == 0) zq[C::kNz / 2] = make_float2(0.f, 0.f);
        continue;
      }
      set_up_rows(q);
      if constexpr (FUSED) {
#pragma unroll kMstftUnroll
        for (int i = 0; i < C::kPairIters; ++i) pair_iter(zq, q, i, Up{0.f, 0.f, 0.f, 0.f});
      } else {
        static_assert(C::kPairIters % 2 == 0, "pair loop is unrolled by two");
        Up u0 = load_up(lane), u1 = load_up(lane + 32);
#pragma unroll 1
        for (int i = 0; i < C::kPairIters; i += 2) {
          const Up n0 = load_up(lane + 32 * (i + 2) + (i + 2 < C::kPairIters ? 0 : C::kNz));
          const Up n1 = load_up(lane + 32 * (i + 3) + (i + 3 < C::kPairIters ? 0 : C::kNz));
          pair_iter(zq, q, i, u0);
          pair_iter(zq, q, i + 1, u1);
          u0 = n0;
          u1 = n1;
        }
      }
      if (lane == 0) {
        constexpr int k = C::kNz / 2;
        const Up u = load_up(k);   // the self pair: sm / pm address the same bin
        const float2 B = rot_inv(grad_bin(zq[k], q, 0.5f, colr[k], colc[k].x, colc[k].y, u.sk, u.pk), k);
        float2 Zk, Zr;
        split_inv(B, B, sm.ws[k], Zk, Zr);
        zq[k] = Zk;
      }
    }
    __syncwarp();
    fft_inverse<N>(v, buf, sm.tw, lane);
    static_for<0, C::kQ>([&](auto qc) {
      constexpr int q = decltype(qc)::value;
      if (it.t0 + q < it.T) {
        float* dst = a.gfb + (it.frame_base + it.t0 + q) * C::kWin;
        static_for<0, C::kR>([&](auto rc) {
          constexpr int r = decltype(rc)::value;
          const int m = 2 * lane + 64 * r;
          const float2 w = *reinterpret_cast<const float2*>(sm.win + m);
          const float2 z = v[q * C::kR + r];
          *reinterpret_cast<float2*>(dst + m) = make_float2(z.x * w.x, z.y * w.y);
        });
      }
    });
    __syncwarp();
    }
  }
  if constexpr (FUSED) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) loss_acc += __shfl_xor_sync(kFullMask, loss_acc, d);
    if (lane == 0) a.partials[vb * kMstftWarps + warp] = loss_acc;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");   // a warp without an item never waited: its copies land before it exits
}

template <int N, bool FUSED>
__global__ void __launch_bounds__(kMstftWarps * 32, 2) mstft_bwd_kernel(const PlanDev p, const MstftBwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  mstft_bwd_body<N, FUSED>(p, a, smem_raw, blockIdx.x, gridDim.x);
}

// g_yg[b, j] = sum_res ( gp[h + j] + [1 <= j <= h] gp[h - j] + [T-2-h < j <= T-2] gp[h + 2T - 2 - j] ),
// gp[P] = sum of the gradient frames covering padded position P  (adjoint of reflect pad + framing)
struct GradOlaArgs {
  int n_res;
  const float* gfb[kMaxRes];
  int n_fft[kMaxRes], hop[kMaxRes], Tf[kMaxRes];
  int B;
  long long T;
  float* g;
};
// the mirror images of sample j in the reflected borders of resolution r (zero for an interior sample)
template <bool L2>
__device__ __forceinline__ float grad_ola_mirrors(const GradOlaArgs& a, int r, const float* fb, long long j) {
  const int N = a.n_fft[r], h = N / 2, win = N / 2, hop = a.hop[r], Tf = a.Tf[r];
  // padded position P -> offset coordinate P - N/4; positions outside every window contribute 0
  auto gp = [&](long long P) -> float {
    const long long pp = P - N / 4;
    return pp >= 0 ? ola_gather<L2>(fb, Tf, hop, win, pp) : 0.f;
  };
  float m = 0.f;
  if (j >= 1 && j <= h) m += gp(h - j);
  if (j > a.T - 2 - h && j <= a.T - 2) m += gp(h + 2 * a.T - 2 - j);
  return m;
}
template <bool L2>
__device__ __forceinline__ float grad_ola_sample(const GradOlaArgs& a, int b, long long j) {
  float acc = 0.f;
  for (int r = 0; r < a.n_res; ++r) {
    const int N = a.n_fft[r], h = N / 2, win = N / 2, hop = a.hop[r], Tf = a.Tf[r];
    const float* fb = a.gfb[r] + static_cast<long long>(b) * Tf * win;
    const long long pp = h + j - N / 4;
    acc += ola_gather<L2>(fb, Tf, hop, win, pp);
    acc += grad_ola_mirrors<L2>(a, r, fb, j);
  }
  return acc;
}
// Four consecutive samples j0 .. j0+3 (j0 % 4 == 0): with hop % 4 == 0 their own positions are covered by the same frames at offsets
// that are multiples of 4, so each frame contributes one 16-byte load; samples inside a reflected border of a resolution add that
// resolution's mirror terms one by one.  Same summation order per sample as grad_ola_sample: resolutions in order; the position
// itself (frames from the last covering one backwards), then the mirrors.
__device__ __forceinline__ float4 grad_ola_quad(const GradOlaArgs& a, int b, long long j0) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = 0; r < a.n_res; ++r) {
    const int N = a.n_fft[r], h = N / 2, win = N / 2, hop = a.hop[r], Tf = a.Tf[r];
    const float* fb = a.gfb[r] + static_cast<long long>(b) * Tf * win;
    const long long pp = N / 4 + j0;                         // (h + j0) - N/4
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int tp_hi, tp_lo;
    if (pp < (1LL << 30)) {   // 32-bit divisions (the 64-bit one is a subroutine)
      const int q = static_cast<int>(pp);
      tp_hi = min(Tf - 1, q / hop);
      tp_lo = q < win ? 0 : (q - win) / hop + 1;
    } else {
      tp_hi = static_cast<int>(min(static_cast<long long>(Tf - 1), pp / hop));
      tp_lo = static_cast<int>((pp - win) / hop) + 1;
    }
    constexpr int kFrames = 6;   // win / hop <= 5 at the reference's resolutions: one round of loads, all in flight together
    for (int t1 = tp_hi; t1 >= tp_lo; t1 -= kFrames) {
      float4 v[kFrames];
#pragma unroll
      for (int j = 0; j < kFrames; ++j) {
        const int tp = t1 - j;
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tp >= tp_lo)
          v[j] = *reinterpret_cast<const float4*>(fb + static_cast<long long>(tp) * win + (pp - static_cast<long long>(tp) * hop));
      }
#pragma unroll
      for (int j = 0; j < kFrames; ++j) {
        if (t1 - j >= tp_lo) { s.x += v[j].x; s.y += v[j].y; s.z += v[j].z; s.w += v[j].w; }
      }
    }
    acc.x += s.x; acc.y += s.y; acc.z += s.z; acc.w += s.w;
    if (j0 <= h || j0 + 3 > a.T - 2 - h) {   // some sample of the quad lies in a reflected border of this resolution
      acc.x += grad_ola_mirrors<false>(a, r, fb, j0);
      acc.y += grad_ola_mirrors<false>(a, r, fb, j0 + 1);
      acc.z += grad_ola_mirrors<false>(a, r, fb, j0 + 2);
      acc.w += grad_ola_mirrors<false>(a, r, fb, j0 + 3);
    }
  }
  return acc;
}

// One thread per 4 samples; block (0, 0) first reduces the loss partial sums when fin.loss is set (the loss-only step: saves the
// separate one-block launch, which sat between the analysis kernels and this one).
__global__ void __launch_bounds__(256) grad_ola_kernel(const GradOlaArgs a, const MstftFinArgs fin) {
  // fin.loss set: the grid carries one extra column of blocks; the first of them reduces the loss partial sums and none of them
  // takes overlap-add work (the reduction is a chain of dependent round trips: inside a working block it was the kernel's tail)
  const int gx = static_cast<int>(gridDim.x) - (fin.loss != nullptr ? 1 : 0);
  if (static_cast<int>(blockIdx.x) == gx) {
    if (blockIdx.y == 0) mstft_finalize_body(fin);
    return;
  }
  const int b = blockIdx.y;
  bool vec = (a.T >= 8);
  for (int r = 0; r < a.n_res; ++r) vec = vec && (a.hop[r] % 4 == 0) && (a.n_fft[r] % 16 == 0);
  float* g = a.g + static_cast<long long>(b) * a.T;
  for (long long j0 = 4 * (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x); j0 < a.T;
       j0 += 4 * static_cast<long long>(gx) * blockDim.x) {
    if (vec) {
      const float4 v = grad_ola_quad(a, b, j0);
      g[j0] = v.x;
      if (j0 + 1 < a.T) g[j0 + 1] = v.y;
      if (j0 + 2 < a.T) g[j0 + 2] = v.z;
      if (j0 + 3 < a.T) g[j0 + 3] = v.w;
    } else {
      for (long long j = j0; j < min(j0 + 4, a.T); ++j) g[j] = grad_ola_sample<false>(a, b, j);
    }
  }
}

// ---- all resolutions as ONE grid (default) ----------------------------------------------------------------------------------
// The per-resolution kernels are each well under one wave of the GPU.  Running them on three streams costs a fork event, two
// stream waits, two join events and two more waits per phase, and once the kernels themselves were tuned the eager training step
// was bound by the host issuing those calls.  Here the grids are simply concatenated: CTA c belongs to the resolution whose
// [cta_end[r - 1], cta_end[r]) range holds it and works exactly like CTA c - cta_end[r - 1] of that resolution's own kernel (one
// pass per warp, same numbers bit for bit).  No cooperative launch, no grid barrier: the reduction / overlap-add stay a second launch.
struct MstftMultiFwdArgs {
  int n_res;
  int cta_end[kMaxRes];
  PlanDev plan[kMaxRes];
  MstftFwdArgs f[kMaxRes];
};
struct MstftMultiBwdArgs {
  int n_res;
  int cta_end[kMaxRes];
  PlanDev plan[kMaxRes];
  MstftBwdArgs b[kMaxRes];
};
__device__ __forceinline__ void mstft_multi_locate(int n_res, const int* cta_end, int* r, int* vb, int* nvb) {
  int i = 0, base = 0;
  while (i + 1 < n_res && static_cast<int>(blockIdx.x) >= cta_end[i]) base = cta_end[i++];
  *r = i;
  *vb = static_cast<int>(blockIdx.x) - base;
  *nvb = cta_end[i] - base;
}
__global__ void __launch_bounds__(kMstftWarps * 32, 2) mstft_multi_fwd_kernel(const __grid_constant__ MstftMultiFwdArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int r, vb, nvb;
  mstft_multi_locate(A.n_res, A.cta_end, &r, &vb, &nvb);
  switch (A.plan[r].n_fft) {
    case 2048: mstft_fwd_body<2048>(A.plan[r], A.f[r], smem_raw, vb, nvb); break;
    case 1024: mstft_fwd_body<1024>(A.plan[r], A.f[r], smem_raw, vb, nvb); break;
    default: mstft_fwd_body<512>(A.plan[r], A.f[r], smem_raw, vb, nvb); break;
  }
}
template <bool FUSED>
__global__ void __launch_bounds__(kMstftWarps * 32, 2) mstft_multi_bwd_kernel(const __grid_constant__ MstftMultiBwdArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int r, vb, nvb;
  mstft_multi_locate(A.n_res, A.cta_end, &r, &vb, &nvb);
  switch (A.plan[r].n_fft) {
    case 2048: mstft_bwd_body<2048, FUSED>(A.plan[r], A.b[r], smem_raw, vb, nvb); break;
    case 1024: mstft_bwd_body<1024, FUSED>(A.plan[r], A.b[r], smem_raw, vb, nvb); break;
    default: mstft_bwd_body<512, FUSED>(A.plan[r], A.b[r], smem_raw, vb, nvb); break;
  }
}

// ---- all resolutions in ONE launch ------------------------------------------------------------------------------------------
// The per-resolution kernels above are each well under one wave of the GPU (184 CTAs at the training size) and were run on three
// streams, followed by the reduction and the overlap-add as two more launches: 552 CTAs over 296 resident slots plus fork / join
// events and two small kernels per step.  Here ONE grid of resident CTAs (<= 2 per SM) deals its CTAs to the resolutions round-robin
// (CTA c works on resolution c % n_res: one set of tables, one code body per CTA), every CTA loops over the items of its
// resolution, and the tail runs in the same launch: the loss reduction by the last CTA to finish (threadfence reduction, no
// spinning), the gradient overlap-add by ALL CTAs after a grid barrier (cooperative launch: the CTAs are co-resident).
struct MstftAllFwdArgs {
  int n_res;
  PlanDev plan[kMaxRes];
  MstftFwdArgs f[kMaxRes];
  MstftFinArgs fin;        // fin.loss == nullptr: no reduction
  unsigned* counter;       // zeroed before the launch
};
struct MstftAllBwdArgs {
  int n_res;
  PlanDev plan[kMaxRes];
  MstftBwdArgs b[kMaxRes];
  MstftFinArgs fin;        // FUSED only
  GradOlaArgs ola;
  unsigned* counter;       // [2]: finished-CTA count (loss reduction), grid barrier; zeroed before the launch
};

__device__ __forceinline__ void mstft_split_grid(int n_res, int* r, int* vb, int* nvb) {
  *r = static_cast<int>(blockIdx.x) % n_res;
  *vb = static_cast<int>(blockIdx.x) / n_res;
  *nvb = (static_cast<int>(gridDim.x) - *r + n_res - 1) / n_res;
}
// true in exactly one CTA: the last one to get here (everything the others wrote before is visible to it)
__device__ __forceinline__ bool mstft_last_cta(unsigned* counter) {
  __shared__ unsigned last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
    __threadfence();
  }
  __syncthreads();
  return last != 0;
}

__global__ void __launch_bounds__(kMstftWarps * 32, 2) mstft_all_fwd_kernel(const __grid_constant__ MstftAllFwdArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int r, vb, nvb;
  mstft_split_grid(A.n_res, &r, &vb, &nvb);
  switch (A.plan[r].n_fft) {
    case 2048: mstft_fwd_body<2048>(A.plan[r], A.f[r], smem_raw, vb, nvb); break;
    case 1024: mstft_fwd_body<1024>(A.plan[r], A.f[r], smem_raw, vb, nvb); break;
    default: mstft_fwd_body<512>(A.plan[r], A.f[r], smem_raw, vb, nvb); break;
  }
  if (A.fin.loss != nullptr && mstft_last_cta(A.counter)) mstft_finalize_body(A.fin);
}

template <bool FUSED>
__global__ void __launch_bounds__(kMstftWarps * 32, 2) mstft_all_bwd_kernel(const __grid_constant__ MstftAllBwdArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int r, vb, nvb;
  mstft_split_grid(A.n_res, &r, &vb, &nvb);
  switch (A.plan[r].n_fft) {
    case 2048: mstft_bwd_body<2048, FUSED>(A.plan[r], A.b[r], smem_raw, vb, nvb); break;
    case 1024: mstft_bwd_body<1024, FUSED>(A.plan[r], A.b[r], smem_raw, vb, nvb); break;
    default: mstft_bwd_body<512, FUSED>(A.plan[r], A.b[r], smem_raw, vb, nvb); break;
  }
  // grid barrier: every gradient frame and partial sum is written and visible
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(A.counter, 1u);
    unsigned seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(A.counter) : "memory");
    } while (seen < gridDim.x);
    __threadfence();
  }
  __syncthreads();
  if constexpr (FUSED) {
    if (blockIdx.x == 0) mstft_finalize_body(A.fin);
  }
  const long long total = static_cast<long long>(A.ola.B) * A.ola.T;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / A.ola.T);
    A.ola.g[i] = grad_ola_sample<true>(A.ola, b, i - static_cast<long long>(b) * A.ola.T);
  }
}

}  // namespace sb200
