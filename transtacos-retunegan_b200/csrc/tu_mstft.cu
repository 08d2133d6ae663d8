// C-ABI entry points: multi-resolution STFT loss forward / backward (include/spectral_b200.h).
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "capi_common.cuh"
#include "mstft.cuh"

using namespace sb200;
using namespace sb200::host;

// ---- multi-resolution STFT loss ---------------------------------------------------------------------

static int mstft_check(const sb200_plan* const* plans, int32_t n_res, int32_t B, int64_t T) {
  if (!plans || n_res < 1 || n_res > kMaxRes) return fail(SB200_ERR_INVALID, "mstft: need 1..4 resolutions");
  if (B < 1) return fail(SB200_ERR_INVALID, "mstft: B must be >= 1");
  for (int r = 0; r < n_res; ++r) {
    if (!plans[r]) return fail(SB200_ERR_INVALID, "mstft: null plan");
    if (T <= plans[r]->cfg.n_fft / 2)
      return fail(SB200_ERR_INVALID, "mstft: reflect padding needs T > n_fft/2 (torch.stft raises)");
  }
  return SB200_OK;
}

// The resolutions of one mstft call are independent until the final reduction / overlap-add and each is well under one
// wave of the GPU, so they run concurrently: resolution 0 on the caller's stream, the others on library-owned side
// streams forked from and joined back into it with events (capturable in a CUDA graph).
struct MstftSide {
  cudaStream_t s[kMaxRes - 1];
  cudaEvent_t fork, join[kMaxRes - 1];
  bool ok = false;
};
// The side streams and the fork / join events are one set per device: the whole enqueue of a call holds this mutex, so
// two host threads (or two caller streams) never interleave their event records / waits (an event is re-recorded only
// after every wait on its previous record has been enqueued).
static std::mutex g_mstft_enqueue_mu;
static MstftSide* mstft_side() {
  static std::mutex mu;
  static MstftSide side[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  MstftSide& m = side[dev];
  if (!m.ok) {
    bool good = cudaEventCreateWithFlags(&m.fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < kMaxRes - 1 && good; ++i)
      good = cudaStreamCreateWithFlags(&m.s[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&m.join[i], cudaEventDisableTiming) == cudaSuccess;
    if (!good) { cudaGetLastError(); return nullptr; }
    m.ok = true;
  }
  return &m;
}
// stream of resolution r (forks the side stream off `st` on first use in this call)
static cudaStream_t mstft_stream(MstftSide* side, int r, cudaStream_t st, bool* forked) {
  if (r == 0 || side == nullptr) return st;
  if (!*forked) {
    cudaEventRecord(side->fork, st);
    *forked = true;
  }
  cudaStreamWaitEvent(side->s[r - 1], side->fork, 0);
  return side->s[r - 1];
}
static void mstft_join(MstftSide* side, int n_res, cudaStream_t st) {
  if (side == nullptr) return;
  for (int r = 1; r < n_res; ++r) {
    cudaEventRecord(side->join[r - 1], side->s[r - 1]);
    cudaStreamWaitEvent(st, side->join[r - 1], 0);
  }
}

static int mstft_grid(long long items) {
  return static_cast<int>(std::min<long long>((items + kMstftWarps - 1) / kMstftWarps, 2LL * sm_count()));
}

// saved layout: per resolution mel_r [B*Tf*n_mel] floats (256 B aligned)
static int64_t mstft_saved_off(const sb200_plan* const* plans, int r, int32_t B, int64_t T) {
  int64_t off = 0;
  for (int i = 0; i < r; ++i) {
    const int64_t Tf = 1 + T / plans[i]->cfg.hop_length;
    off += ((B * Tf * plans[i]->cfg.n_mel * 4 + 255) / 256) * 256;
  }
  return off;
}
// workspace layout: per resolution gradient frames [B*Tf*win] floats, then partial sums [2*SMs*warps] per resolution
static int64_t mstft_ws_gfb_off(const sb200_plan* const* plans, int r, int32_t B, int64_t T) {
  int64_t off = 0;
  for (int i = 0; i < r; ++i) {
    const int64_t Tf = 1 + T / plans[i]->cfg.hop_length;
    off += ((B * Tf * plans[i]->cfg.win_length * 4 + 255) / 256) * 256;
  }
  return off;
}
static int64_t mstft_partials_bytes() { return ((2LL * sm_count() * kMstftWarps * 4 + 255) / 256) * 256; }
// two 32-bit counters of the single-launch kernels (finished CTAs, grid barrier), behind the partial sums
static int64_t mstft_counter_off(const sb200_plan* const* plans, int32_t n_res, int32_t B, int64_t T) {
  return mstft_ws_gfb_off(plans, n_res, B, T) + n_res * mstft_partials_bytes();
}

// SB200_MSTFT_SINGLE=1 selects ONE cooperative launch for all resolutions with the loss reduction and the overlap-add folded in
// (mstft_all_*_kernel).  OFF BY DEFAULT: measured on B200 at 16 x 22 050 it is slower than the per-resolution launches on side
// streams, 190 us against 180 us per loss-only step and 410 against 393 us with spec stacks (tools/ab_mstft.sh,
// profiles/r02_mstft_summary.md): a warp's pass is a long dependent chain (~6000 instructions, issue-active 15-17 %), the
// resident grid gives every warp two passes back to back where the three concurrent kernels give the hardware scheduler
// 552 one-pass CTAs to pack, and the folded tail does not make up for it.
static bool mstft_single_enabled() {
  static const int v = [] {
    const char* e = std::getenv("SB200_MSTFT_SINGLE");
    return e ? std::atoi(e) : 0;
  }();
  return v != 0;
}

// SB200_MSTFT_STREAMS=1 selects round 1's formulation, one launch per resolution on library-owned side streams (fork / join
// events): the kernels' own time is the same, the step pays nine more driver calls per phase (tools/ab_mstft.sh).
static bool mstft_streams_enabled() {
  static const int v = [] {
    const char* e = std::getenv("SB200_MSTFT_STREAMS");
    return e ? std::atoi(e) : 0;
  }();
  return v != 0;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per device and sticky: set it when a device first needs more than it has had
// (a driver call per launch otherwise).  The race between two host threads is benign (both set the same value).
static bool mstft_smem_grow(size_t (&have)[64], size_t smem) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (smem <= have[dev]) return false;
  have[dev] = smem;
  return true;
}

// Grids of the concatenated launch.  Every resolution gets ceil(sub-items / warps) CTAs: one pass per warp, 552 CTAs against 296
// resident slots at the training size, packed by the hardware scheduler.  SB200_MSTFT_RESIDENT=1 scales the grids down to the
// resident count instead and lets the CTAs loop (one table fill per CTA instead of two, second pass from a warm instruction
// cache): measured SLOWER, 105 against 100 us of GPU time per loss-only step and 216 against 208 us with spec stacks
// (tools/probe_mstft_graph.py, CUDA-graph replay) -- a warp's second pass starts only when its own first one ends, while a
// fresh CTA starts as soon as any slot frees up.  Off by default.
static void mstft_multi_grids(const long long* subs, int n_res, int resident, int* grid) {
  static const int on = [] {
    const char* e = std::getenv("SB200_MSTFT_RESIDENT");
    return e ? std::atoi(e) : 0;
  }();
  long long need[kMaxRes], total = 0;
  for (int r = 0; r < n_res; ++r) {
    need[r] = (subs[r] + kMstftWarps - 1) / kMstftWarps;
    total += need[r];
  }
  for (int r = 0; r < n_res; ++r) {
    grid[r] = static_cast<int>(std::min<long long>(need[r], 2LL * sm_count()));
    if (on && resident > 0 && total > resident)
      grid[r] = static_cast<int>(std::max<long long>(1, std::min<long long>(need[r], need[r] * resident / total)));
  }
}
template <class Kern>
static int mstft_resident(Kern kern, size_t smem, int (&cache)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (cache[dev] == 0) {
    int per_sm = 0;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kMstftWarps * 32, smem) != cudaSuccess) {
      cudaGetLastError();
      per_sm = 0;
    }
    cache[dev] = per_sm > 0 ? per_sm * sm_count() : -1;
  }
  return cache[dev] > 0 ? cache[dev] : 0;
}

static size_t mstft_all_smem(const sb200_plan* const* plans, int32_t n_res, bool bwd) {
  size_t smem = 0;
  for (int r = 0; r < n_res; ++r) {
    size_t s = 0;
    SB200_DISPATCH_N(plans[r], s = bwd ? mstft_bwd_smem_bytes<kN>(plans[r]->dev) : feat_smem_bytes<kN>(plans[r]->dev));
    smem = std::max(smem, s);
  }
  return smem;
}

// Grid of the all-resolution kernels: the CTAs the work needs, capped at what is resident at once (cooperative launch).
template <class Kern>
static int mstft_all_grid(Kern kern, size_t smem, const long long* subs, int n_res) {
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kMstftWarps * 32, smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    return 0;
  }
  long long need = 0;   // round-robin dealing: every resolution gets ceil(grid / n_res) or floor(grid / n_res) CTAs
  for (int r = 0; r < n_res; ++r) need = std::max(need, (subs[r] + kMstftWarps - 1) / kMstftWarps);
  need *= n_res;
  return static_cast<int>(std::min<long long>(need, static_cast<long long>(per_sm) * sm_count()));
}

int64_t sb200_mstft_saved_bytes(const sb200_plan* const* plans, int32_t n_res, int32_t B, int64_t T) {
  if (mstft_check(plans, n_res, B, T)) return -1;
  return mstft_saved_off(plans, n_res, B, T) + 256;
}

int64_t sb200_mstft_workspace_bytes(const sb200_plan* const* plans, int32_t n_res, int32_t B, int64_t T) {
  if (mstft_check(plans, n_res, B, T)) return -1;
  return mstft_ws_gfb_off(plans, n_res, B, T) + n_res * mstft_partials_bytes() + 256;
}

template <int N>
static void launch_mstft_fwd(const sb200_plan* plan, const MstftFwdArgs& a, int grid, cudaStream_t st) {
  const size_t smem = feat_smem_bytes<N>(plan->dev);
  cudaFuncSetAttribute(mstft_fwd_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  mstft_fwd_kernel<N><<<grid, kMstftWarps * 32, smem, st>>>(plan->dev, a);
}
template <int N>
static void launch_mstft_bwd(const sb200_plan* plan, const MstftBwdArgs& a, int grid, cudaStream_t st, bool fused = false) {
  const size_t smem = mstft_bwd_smem_bytes<N>(plan->dev);
  if (fused) {
    cudaFuncSetAttribute(mstft_bwd_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    mstft_bwd_kernel<N, true><<<grid, kMstftWarps * 32, smem, st>>>(plan->dev, a);
  } else {
    cudaFuncSetAttribute(mstft_bwd_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    mstft_bwd_kernel<N, false><<<grid, kMstftWarps * 32, smem, st>>>(plan->dev, a);
  }
}

static BatchDev mstft_batch(const sb200_plan* plan, int32_t B, int64_t T) {
  const int Q = 4096 / plan->cfg.n_fft;   // frames per item
  BatchDev d{};
  d.B = B;
  d.len = T;
  d.stride = T;
  d.frames_per_row = 1 + T / plan->cfg.hop_length;
  d.items_per_row = (d.frames_per_row + Q - 1) / Q;
  d.total_items = d.items_per_row * B;
  return d;
}

// Backward (g_specs / g_loss) or fused loss + gradient, all resolutions + overlap-add (+ loss reduction) in one cooperative launch.
// Returns SB200_OK, an error, or 1 when the single-launch path is not available (caller falls back).
static int mstft_all_bwd(const sb200_plan* const* plans, int32_t n_res, const float* y, const float* y_g, int32_t B, int64_t T,
                         int32_t phd_phase, const float* g_loss, const float* const* g_specs_g, const void* saved, float* loss,
                         float* g_yg, char* ws, bool fused, cudaStream_t st) {
  if (!mstft_single_enabled()) return 1;
  MstftAllBwdArgs A{};
  A.n_res = n_res;
  A.fin.n_res = n_res;
  A.fin.loss = loss;
  A.ola.n_res = n_res;
  A.ola.B = B;
  A.ola.T = T;
  A.ola.g = g_yg;
  A.counter = reinterpret_cast<unsigned*>(ws + mstft_counter_off(plans, n_res, B, T));
  const int64_t part0 = mstft_ws_gfb_off(plans, n_res, B, T);
  long long subs[kMaxRes];
  for (int r = 0; r < n_res; ++r) {
    const sb200_plan* plan = plans[r];
    MstftBwdArgs& a = A.b[r];
    A.plan[r] = plan->dev;
    a.y = y;
    a.yg = y_g;
    a.bd = mstft_batch(plan, B, T);
    a.Tf = static_cast<int>(a.bd.frames_per_row);
    a.mel_r = saved ? reinterpret_cast<const float*>(static_cast<const char*>(saved) + mstft_saved_off(plans, r, B, T)) : nullptr;
    a.g_loss = g_loss;
    a.loss_scale = static_cast<float>(1.0 / (static_cast<double>(n_res) * B * plan->cfg.n_mel * a.Tf));
    a.g_spec = g_specs_g ? g_specs_g[r] : nullptr;
    a.phd_phase = phd_phase;
    a.gfb = reinterpret_cast<float*>(ws + mstft_ws_gfb_off(plans, r, B, T));
    a.partials = reinterpret_cast<float*>(ws + part0 + r * mstft_partials_bytes());
    subs[r] = 2 * a.bd.total_items;
    A.fin.partials[r] = a.partials;
    A.fin.inv_count[r] = static_cast<float>(1.0 / (static_cast<double>(B) * plan->cfg.n_mel * a.Tf));
    A.ola.gfb[r] = a.gfb;
    A.ola.n_fft[r] = plan->cfg.n_fft;
    A.ola.hop[r] = plan->cfg.hop_length;
    A.ola.Tf[r] = a.Tf;
  }
  const size_t smem = mstft_all_smem(plans, n_res, true);
  const void* kern = fused ? reinterpret_cast<const void*>(mstft_all_bwd_kernel<true>) : reinterpret_cast<const void*>(mstft_all_bwd_kernel<false>);
  const int grid = fused ? mstft_all_grid(mstft_all_bwd_kernel<true>, smem, subs, n_res) : mstft_all_grid(mstft_all_bwd_kernel<false>, smem, subs, n_res);
  if (grid <= 0) return 1;
  for (int r = 0; r < n_res; ++r) A.fin.n_partials[r] = ((grid - r + n_res - 1) / n_res) * kMstftWarps;
  cudaMemsetAsync(A.counter, 0, 2 * sizeof(unsigned), st);
  void* args[] = {&A};
  if (cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kMstftWarps * 32), args, smem, st) != cudaSuccess) {
    cudaGetLastError();
    return 1;
  }
  return check_launch(fused ? "mstft_all_bwd_kernel<fused>" : "mstft_all_bwd_kernel");
}

// sb200_peer_reduce -> device view, or world = 0 (no reduction)
static int peer_dev(const sb200_peer_reduce* pr, PeerDev* out) {
  *out = PeerDev{};
  if (pr == nullptr) return SB200_OK;
  if (pr->world < 1 || pr->world > kMaxPeers || pr->rank < 0 || pr->rank >= pr->world || !pr->loss_global)
    return fail(SB200_ERR_INVALID, "peer reduce: need 1 <= world <= 8, 0 <= rank < world and a loss_global pointer");
  for (int q = 0; q < pr->world; ++q) {
    if (!pr->peer[q]) return fail(SB200_ERR_INVALID, "peer reduce: null exchange buffer");
    out->buf[q] = pr->peer[q];
  }
  out->rank = pr->rank;
  out->world = pr->world;
  out->out = pr->loss_global;
  return SB200_OK;
}

static int mstft_forward_impl(const sb200_plan* const* plans, int32_t n_res, const float* y, const float* y_g, int32_t B,
                              int64_t T, int32_t phd_phase, float* loss, float* const* specs_r, float* const* specs_g,
                              void* saved, void* workspace, const sb200_peer_reduce* pr, sb200_stream stream) {
  if (int rc = mstft_check(plans, n_res, B, T)) return rc;
  PeerDev peer;
  if (int rc = peer_dev(pr, &peer)) return rc;
  if (pr && !loss) return fail(SB200_ERR_INVALID, "mstft_forward_ddp: the loss reduction needs the loss");
  if (!y || !y_g || !saved || !workspace) return fail(SB200_ERR_INVALID, "mstft_forward: null argument");
  if (!loss && !specs_r && !specs_g) return fail(SB200_ERR_INVALID, "mstft_forward: neither loss nor specs requested (loss.py:62 raises)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  const int64_t part0 = mstft_ws_gfb_off(plans, n_res, B, T);
  if (mstft_single_enabled() && !pr) {
    MstftAllFwdArgs A{};
    A.n_res = n_res;
    A.fin.n_res = n_res;
    A.fin.loss = loss;
    A.counter = reinterpret_cast<unsigned*>(ws + mstft_counter_off(plans, n_res, B, T));
    long long subs[kMaxRes];
    for (int r = 0; r < n_res; ++r) {
      const sb200_plan* plan = plans[r];
      MstftFwdArgs& a = A.f[r];
      A.plan[r] = plan->dev;
      a.y = y;
      a.yg = y_g;
      a.bd = mstft_batch(plan, B, T);
      a.Tf = static_cast<int>(a.bd.frames_per_row);
      a.spec_r = specs_r ? specs_r[r] : nullptr;
      a.spec_g = specs_g ? specs_g[r] : nullptr;
      a.phd_phase = phd_phase;
      a.mel_r = reinterpret_cast<float*>(static_cast<char*>(saved) + mstft_saved_off(plans, r, B, T));
      a.partials = reinterpret_cast<float*>(ws + part0 + r * mstft_partials_bytes());
      a.want_loss = loss != nullptr;
      subs[r] = 2 * a.bd.total_items;
      A.fin.partials[r] = a.partials;
      A.fin.inv_count[r] = static_cast<float>(1.0 / (static_cast<double>(B) * plan->cfg.n_mel * a.Tf));
    }
    const size_t smem = mstft_all_smem(plans, n_res, false);
    const int grid = mstft_all_grid(mstft_all_fwd_kernel, smem, subs, n_res);
    if (grid > 0) {
      for (int r = 0; r < n_res; ++r) A.fin.n_partials[r] = ((grid - r + n_res - 1) / n_res) * kMstftWarps;
      if (loss) cudaMemsetAsync(A.counter, 0, 2 * sizeof(unsigned), st);
      mstft_all_fwd_kernel<<<grid, kMstftWarps * 32, smem, st>>>(A);
      return check_launch("mstft_all_fwd_kernel");
    }
  }
  if (!mstft_streams_enabled() || pr) {   // default: the resolutions' grids concatenated into one launch
    MstftMultiFwdArgs A{};
    MstftFinArgs fin{};
    A.n_res = fin.n_res = n_res;
    fin.loss = loss;
    fin.peer = peer;
    int total = 0;
    long long subs[kMaxRes];
    for (int r = 0; r < n_res; ++r) {
      const sb200_plan* plan = plans[r];
      MstftFwdArgs& a = A.f[r];
      A.plan[r] = plan->dev;
      a.y = y;
      a.yg = y_g;
      a.bd = mstft_batch(plan, B, T);
      a.Tf = static_cast<int>(a.bd.frames_per_row);
      a.spec_r = specs_r ? specs_r[r] : nullptr;
      a.spec_g = specs_g ? specs_g[r] : nullptr;
      a.phd_phase = phd_phase;
      a.mel_r = reinterpret_cast<float*>(static_cast<char*>(saved) + mstft_saved_off(plans, r, B, T));
      a.partials = reinterpret_cast<float*>(ws + part0 + r * mstft_partials_bytes());
      a.want_loss = loss != nullptr;
      subs[r] = 2 * a.bd.total_items;
      fin.partials[r] = a.partials;
      fin.inv_count[r] = static_cast<float>(1.0 / (static_cast<double>(B) * plan->cfg.n_mel * a.Tf));
    }
    const size_t smem = mstft_all_smem(plans, n_res, false);
    static int resident[64] = {};
    int grid[kMaxRes];
    mstft_multi_grids(subs, n_res, mstft_resident(mstft_multi_fwd_kernel, smem, resident), grid);
    for (int r = 0; r < n_res; ++r) {
      total += grid[r];
      A.cta_end[r] = total;
      fin.n_partials[r] = grid[r] * kMstftWarps;
    }
    static size_t smem_set[64] = {};
    if (mstft_smem_grow(smem_set, smem))
      cudaFuncSetAttribute(mstft_multi_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    mstft_multi_fwd_kernel<<<total, kMstftWarps * 32, smem, st>>>(A);
    if (int rc = check_launch("mstft_multi_fwd_kernel")) return rc;
    if (loss) {
      mstft_finalize_kernel<<<1, 256, 0, st>>>(fin);
      if (int rc = check_launch("mstft_finalize_kernel")) return rc;
    }
    return SB200_OK;
  }
  std::lock_guard<std::mutex> enqueue_lock(g_mstft_enqueue_mu);
  MstftFinArgs fin{};
  fin.n_res = n_res;
  fin.loss = loss;
  MstftSide* side = mstft_side();
  bool forked = false;
  for (int r = 0; r < n_res; ++r) {
    const sb200_plan* plan = plans[r];
    cudaStream_t sr = mstft_stream(side, r, st, &forked);
    MstftFwdArgs a{};
    a.y = y;
    a.yg = y_g;
    a.bd = mstft_batch(plan, B, T);
    a.Tf = static_cast<int>(a.bd.frames_per_row);
    a.spec_r = specs_r ? specs_r[r] : nullptr;
    a.spec_g = specs_g ? specs_g[r] : nullptr;
    a.phd_phase = phd_phase;
    a.mel_r = reinterpret_cast<float*>(static_cast<char*>(saved) + mstft_saved_off(plans, r, B, T));
    a.partials = reinterpret_cast<float*>(ws + part0 + r * mstft_partials_bytes());
    a.want_loss = loss != nullptr;
    const int grid = mstft_grid(2 * a.bd.total_items);
    SB200_DISPATCH_N(plan, launch_mstft_fwd<kN>(plan, a, grid, sr));
    if (int rc = check_launch("mstft_fwd_kernel")) { mstft_join(side, r + 1, st); return rc; }
    fin.partials[r] = a.partials;
    fin.n_partials[r] = grid * kMstftWarps;
    fin.inv_count[r] = static_cast<float>(1.0 / (static_cast<double>(B) * plan->cfg.n_mel * a.Tf));
  }
  mstft_join(side, n_res, st);
  if (loss) {
    mstft_finalize_kernel<<<1, 256, 0, st>>>(fin);
    if (int rc = check_launch("mstft_finalize_kernel")) return rc;
  }
  return SB200_OK;
}

// Backward (g_specs / g_loss) or fused loss + gradient of all resolutions as one concatenated grid, then the overlap-add (which
// also reduces the loss partial sums when `loss` is set).
static int mstft_multi_bwd(const sb200_plan* const* plans, int32_t n_res, const float* y, const float* y_g, int32_t B, int64_t T,
                           int32_t phd_phase, const float* g_loss, const float* const* g_specs_g, const void* saved, float* loss,
                           float* g_yg, char* ws, bool fused, cudaStream_t st, const PeerDev& peer = PeerDev{}) {
  MstftMultiBwdArgs A{};
  MstftFinArgs fin{};
  GradOlaArgs o{};
  A.n_res = fin.n_res = o.n_res = n_res;
  fin.loss = loss;
  fin.peer = peer;
  o.B = B;
  o.T = T;
  o.g = g_yg;
  const int64_t part0 = mstft_ws_gfb_off(plans, n_res, B, T);
  int total = 0;
  long long subs[kMaxRes];
  for (int r = 0; r < n_res; ++r) {
    const sb200_plan* plan = plans[r];
    MstftBwdArgs& a = A.b[r];
    A.plan[r] = plan->dev;
    a.y = y;
    a.yg = y_g;
    a.bd = mstft_batch(plan, B, T);
    a.Tf = static_cast<int>(a.bd.frames_per_row);
    a.mel_r = saved ? reinterpret_cast<const float*>(static_cast<const char*>(saved) + mstft_saved_off(plans, r, B, T)) : nullptr;
    a.g_loss = g_loss;
    a.loss_scale = static_cast<float>(1.0 / (static_cast<double>(n_res) * B * plan->cfg.n_mel * a.Tf));
    a.g_spec = g_specs_g ? g_specs_g[r] : nullptr;
    a.phd_phase = phd_phase;
    a.gfb = reinterpret_cast<float*>(ws + mstft_ws_gfb_off(plans, r, B, T));
    a.partials = reinterpret_cast<float*>(ws + part0 + r * mstft_partials_bytes());
    subs[r] = 2 * a.bd.total_items;
    fin.partials[r] = a.partials;
    fin.inv_count[r] = static_cast<float>(1.0 / (static_cast<double>(B) * plan->cfg.n_mel * a.Tf));
    o.gfb[r] = a.gfb;
    o.n_fft[r] = plan->cfg.n_fft;
    o.hop[r] = plan->cfg.hop_length;
    o.Tf[r] = a.Tf;
  }
  const size_t smem = mstft_all_smem(plans, n_res, true);
  static int resident[2][64] = {};
  int grid_r[kMaxRes];
  mstft_multi_grids(subs, n_res,
                    fused ? mstft_resident(mstft_multi_bwd_kernel<true>, smem, resident[1])
                          : mstft_resident(mstft_multi_bwd_kernel<false>, smem, resident[0]), grid_r);
  for (int r = 0; r < n_res; ++r) {
    total += grid_r[r];
    A.cta_end[r] = total;
    fin.n_partials[r] = grid_r[r] * kMstftWarps;
  }
  static size_t smem_set[2][64] = {};
  if (mstft_smem_grow(smem_set[fused], smem)) {
    if (fused) cudaFuncSetAttribute(mstft_multi_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    else cudaFuncSetAttribute(mstft_multi_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  }
  if (fused) mstft_multi_bwd_kernel<true><<<total, kMstftWarps * 32, smem, st>>>(A);
  else mstft_multi_bwd_kernel<false><<<total, kMstftWarps * 32, smem, st>>>(A);
  if (int rc = check_launch(fused ? "mstft_multi_bwd_kernel<fused>" : "mstft_multi_bwd_kernel")) return rc;
  dim3 grid(grid_for((T + 3) / 4, 256, 2) + (loss ? 1 : 0), B);
  grad_ola_kernel<<<grid, 256, 0, st>>>(o, loss ? fin : MstftFinArgs{});
  return check_launch("grad_ola_kernel");
}

int sb200_mstft_backward(const sb200_plan* const* plans, int32_t n_res, const float* y_g, int32_t B, int64_t T,
                         int32_t phd_phase, const float* g_loss, const float* const* g_specs_g, const void* saved,
                         float* g_yg, void* workspace, sb200_stream stream) {
  if (int rc = mstft_check(plans, n_res, B, T)) return rc;
  if (!y_g || !saved || !g_yg || !workspace) return fail(SB200_ERR_INVALID, "mstft_backward: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  {
    const int rc = mstft_all_bwd(plans, n_res, nullptr, y_g, B, T, phd_phase, g_loss, g_specs_g, saved, nullptr, g_yg, ws, false, st);
    if (rc <= 0) return rc;
  }
  if (!mstft_streams_enabled())
    return mstft_multi_bwd(plans, n_res, nullptr, y_g, B, T, phd_phase, g_loss, g_specs_g, saved, nullptr, g_yg, ws, false, st);
  std::lock_guard<std::mutex> enqueue_lock(g_mstft_enqueue_mu);
  GradOlaArgs o{};
  o.n_res = n_res;
  o.B = B;
  o.T = T;
  o.g = g_yg;
  MstftSide* side = mstft_side();
  bool forked = false;
  for (int r = 0; r < n_res; ++r) {
    const sb200_plan* plan = plans[r];
    cudaStream_t sr = mstft_stream(side, r, st, &forked);
    MstftBwdArgs a{};
    a.yg = y_g;
    a.bd = mstft_batch(plan, B, T);
    a.Tf = static_cast<int>(a.bd.frames_per_row);
    a.mel_r = reinterpret_cast<const float*>(static_cast<const char*>(saved) + mstft_saved_off(plans, r, B, T));
    a.g_loss = g_loss;
    a.loss_scale = static_cast<float>(1.0 / (static_cast<double>(n_res) * B * plan->cfg.n_mel * a.Tf));
    a.g_spec = g_specs_g ? g_specs_g[r] : nullptr;
    a.phd_phase = phd_phase;
    a.gfb = reinterpret_cast<float*>(ws + mstft_ws_gfb_off(plans, r, B, T));
    const int grid = mstft_grid(2 * a.bd.total_items);
    SB200_DISPATCH_N(plan, launch_mstft_bwd<kN>(plan, a, grid, sr));
    if (int rc = check_launch("mstft_bwd_kernel")) { mstft_join(side, r + 1, st); return rc; }
    o.gfb[r] = a.gfb;
    o.n_fft[r] = plan->cfg.n_fft;
    o.hop[r] = plan->cfg.hop_length;
    o.Tf[r] = a.Tf;
  }
  mstft_join(side, n_res, st);
  dim3 grid(grid_for((T + 3) / 4, 256, 2), B);
  grad_ola_kernel<<<grid, 256, 0, st>>>(o, MstftFinArgs{});
  return check_launch("grad_ola_kernel");
}

// Loss value AND d loss / d y_g in one pass (one launch for all resolutions + the overlap-add, which carries the reduction): the loss-only training
// step of retunegan/train.py:165,192 without a second analysis in backward.  grad_yg [B, T] is the gradient for a unit
// upstream gradient; the autograd wrapper scales it by the incoming gradient.
static int mstft_loss_and_grad_impl(const sb200_plan* const* plans, int32_t n_res, const float* y, const float* y_g, int32_t B,
                                    int64_t T, float* loss, float* grad_yg, void* workspace, const sb200_peer_reduce* pr,
                                    sb200_stream stream) {
  if (int rc = mstft_check(plans, n_res, B, T)) return rc;
  if (!y || !y_g || !loss || !grad_yg || !workspace) return fail(SB200_ERR_INVALID, "mstft_loss_and_grad: null argument");
  PeerDev peer;
  if (int rc = peer_dev(pr, &peer)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  if (!pr) {
    const int rc = mstft_all_bwd(plans, n_res, y, y_g, B, T, 0, nullptr, nullptr, nullptr, loss, grad_yg, ws, true, st);
    if (rc <= 0) return rc;
  }
  if (!mstft_streams_enabled() || pr)
    return mstft_multi_bwd(plans, n_res, y, y_g, B, T, 0, nullptr, nullptr, nullptr, loss, grad_yg, ws, true, st, peer);
  std::lock_guard<std::mutex> enqueue_lock(g_mstft_enqueue_mu);
  const int64_t part0 = mstft_ws_gfb_off(plans, n_res, B, T);
  MstftFinArgs fin{};
  fin.n_res = n_res;
  fin.loss = loss;
  GradOlaArgs o{};
  o.n_res = n_res;
  o.B = B;
  o.T = T;
  o.g = grad_yg;
  MstftSide* side = mstft_side();
  bool forked = false;
  for (int r = 0; r < n_res; ++r) {
    const sb200_plan* plan = plans[r];
    cudaStream_t sr = mstft_stream(side, r, st, &forked);
    MstftBwdArgs a{};
    a.y = y;
    a.yg = y_g;
    a.bd = mstft_batch(plan, B, T);
    a.Tf = static_cast<int>(a.bd.frames_per_row);
    a.loss_scale = static_cast<float>(1.0 / (static_cast<double>(n_res) * B * plan->cfg.n_mel * a.Tf));
    a.gfb = reinterpret_cast<float*>(ws + mstft_ws_gfb_off(plans, r, B, T));
    a.partials = reinterpret_cast<float*>(ws + part0 + r * mstft_partials_bytes());
    const int grid = mstft_grid(2 * a.bd.total_items);
    SB200_DISPATCH_N(plan, launch_mstft_bwd<kN>(plan, a, grid, sr, true));
    if (int rc = check_launch("mstft_bwd_kernel<fused>")) { mstft_join(side, r + 1, st); return rc; }
    fin.partials[r] = a.partials;
    fin.n_partials[r] = grid * kMstftWarps;
    fin.inv_count[r] = static_cast<float>(1.0 / (static_cast<double>(B) * plan->cfg.n_mel * a.Tf));
    o.gfb[r] = a.gfb;
    o.n_fft[r] = plan->cfg.n_fft;
    o.hop[r] = plan->cfg.hop_length;
    o.Tf[r] = a.Tf;
  }
  mstft_join(side, n_res, st);
  dim3 grid(grid_for((T + 3) / 4, 256, 2) + 1, B);
  grad_ola_kernel<<<grid, 256, 0, st>>>(o, fin);   // the extra block column reduces the loss partial sums
  return check_launch("grad_ola_kernel");
}

int sb200_mstft_forward(const sb200_plan* const* plans, int32_t n_res, const float* y, const float* y_g, int32_t B,
                        int64_t T, int32_t phd_phase, float* loss, float* const* specs_r, float* const* specs_g,
                        void* saved, void* workspace, sb200_stream stream) {
  return mstft_forward_impl(plans, n_res, y, y_g, B, T, phd_phase, loss, specs_r, specs_g, saved, workspace, nullptr, stream);
}
int sb200_mstft_forward_ddp(const sb200_plan* const* plans, int32_t n_res, const float* y, const float* y_g, int32_t B,
                            int64_t T, int32_t phd_phase, float* loss, float* const* specs_r, float* const* specs_g,
                            void* saved, void* workspace, const sb200_peer_reduce* peers, sb200_stream stream) {
  if (!peers) return fail(SB200_ERR_INVALID, "mstft_forward_ddp: null peer descriptor");
  return mstft_forward_impl(plans, n_res, y, y_g, B, T, phd_phase, loss, specs_r, specs_g, saved, workspace, peers, stream);
}
int sb200_mstft_loss_and_grad(const sb200_plan* const* plans, int32_t n_res, const float* y, const float* y_g, int32_t B,
                              int64_t T, float* loss, float* grad_yg, void* workspace, sb200_stream stream) {
  return mstft_loss_and_grad_impl(plans, n_res, y, y_g, B, T, loss, grad_yg, workspace, nullptr, stream);
}
int sb200_mstft_loss_and_grad_ddp(const sb200_plan* const* plans, int32_t n_res, const float* y, const float* y_g, int32_t B,
                                  int64_t T, float* loss, float* grad_yg, void* workspace, const sb200_peer_reduce* peers,
                                  sb200_stream stream) {
  if (!peers) return fail(SB200_ERR_INVALID, "mstft_loss_and_grad_ddp: null peer descriptor");
  return mstft_loss_and_grad_impl(plans, n_res, y, y_g, B, T, loss, grad_yg, workspace, peers, stream);
}

// ---- exchange buffers of the in-kernel loss reduction (mstft.cuh: PeerBuf) ---------------------------------------------------
int64_t sb200_peer_buffer_bytes(void) { return static_cast<int64_t>(sizeof(PeerBuf)); }
int sb200_peer_buffer_create(void** buf, void* ipc_handle) {
  if (!buf) return fail(SB200_ERR_INVALID, "peer_buffer_create: null argument");
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, sizeof(PeerBuf));   // cudaMalloc, not a pool: the allocation is exported through CUDA IPC
  if (e == cudaSuccess) e = cudaMemset(d, 0, sizeof(PeerBuf));
  if (e == cudaSuccess && ipc_handle) {
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == 64, "CUDA IPC handle size");
    e = cudaIpcGetMemHandle(&h, d);
    if (e == cudaSuccess) std::memcpy(ipc_handle, &h, sizeof(h));
  }
  if (e != cudaSuccess) {
    if (d) cudaFree(d);
    cudaGetLastError();
    return fail(SB200_ERR_CUDA, std::string("peer_buffer_create: ") + cudaGetErrorString(e));
  }
  *buf = d;
  return SB200_OK;
}
int sb200_peer_buffer_open(const void* ipc_handle, void** buf) {
  if (!ipc_handle || !buf) return fail(SB200_ERR_INVALID, "peer_buffer_open: null argument");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, ipc_handle, sizeof(h));
  void* d = nullptr;
  const cudaError_t e = cudaIpcOpenMemHandle(&d, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(SB200_ERR_CUDA, std::string("peer_buffer_open: ") + cudaGetErrorString(e));
  }
  *buf = d;
  return SB200_OK;
}
int sb200_peer_buffer_close(void* buf) {
  if (buf && cudaIpcCloseMemHandle(buf) != cudaSuccess) {
    cudaGetLastError();
    return fail(SB200_ERR_CUDA, "peer_buffer_close failed");
  }
  return SB200_OK;
}
int sb200_peer_buffer_destroy(void* buf) {
  if (buf && cudaFree(buf) != cudaSuccess) {
    cudaGetLastError();
    return fail(SB200_ERR_CUDA, "peer_buffer_destroy failed");
  }
  return SB200_OK;
}

// ---- get_stft_torch, differentiable (retunegan/audio.py:150-170) ------------------------------------------------------------

static int stft_smp_check(const sb200_plan* plan, int32_t B, int64_t T) {
  if (!plan) return fail(SB200_ERR_INVALID, "stft_smp: null plan");
  if (B < 1) return fail(SB200_ERR_INVALID, "stft_smp: B must be >= 1");
  if (T <= plan->cfg.n_fft / 2) return fail(SB200_ERR_INVALID, "stft_smp: reflect padding needs T > n_fft/2 (torch.stft raises)");
  return SB200_OK;
}

int64_t sb200_stft_smp_workspace_bytes(const sb200_plan* plan, int32_t B, int64_t T) {
  if (stft_smp_check(plan, B, T)) return -1;
  const int64_t Tf = 1 + T / plan->cfg.hop_length;
  return B * Tf * plan->cfg.win_length * 4 + 256;
}

template <int N>
static void launch_stft_smp(const sb200_plan* plan, const StftSmpArgs& a, int grid, cudaStream_t st) {
  const size_t smem = feat_smem_bytes<N>(plan->dev);
  cudaFuncSetAttribute(stft_smp_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  stft_smp_kernel<N><<<grid, kMstftWarps * 32, smem, st>>>(plan->dev, a);
}

int sb200_stft_smp_forward(const sb200_plan* plan, const float* y, int32_t B, int64_t T, float* S, float* M, float* P,
                           sb200_stream stream) {
  if (int rc = stft_smp_check(plan, B, T)) return rc;
  if (!y || (!S && !M && !P)) return fail(SB200_ERR_INVALID, "stft_smp_forward: null argument");
  StftSmpArgs a{};
  a.y = y;
  a.bd = mstft_batch(plan, B, T);
  a.Tf = static_cast<int>(a.bd.frames_per_row);
  a.S = S;
  a.P = P;
  a.M = M;
  const int grid = mstft_grid(2 * a.bd.total_items);
  SB200_DISPATCH_N(plan, launch_stft_smp<kN>(plan, a, grid, static_cast<cudaStream_t>(stream)));
  return check_launch("stft_smp_kernel");
}

int sb200_stft_smp_backward(const sb200_plan* plan, const float* y, int32_t B, int64_t T, const float* g_S, const float* g_M,
                            const float* g_P, float* g_y, void* workspace, sb200_stream stream) {
  if (int rc = stft_smp_check(plan, B, T)) return rc;
  if (!y || !g_y || !workspace) return fail(SB200_ERR_INVALID, "stft_smp_backward: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MstftBwdArgs a{};
  a.yg = y;
  a.bd = mstft_batch(plan, B, T);
  a.Tf = static_cast<int>(a.bd.frames_per_row);
  a.raw = 1;
  a.g_s_raw = g_S;
  a.g_p_raw = g_P;
  a.g_m_raw = g_M;
  a.gfb = static_cast<float*>(workspace);
  const int grid = mstft_grid(2 * a.bd.total_items);
  SB200_DISPATCH_N(plan, launch_mstft_bwd<kN>(plan, a, grid, st));
  if (int rc = check_launch("mstft_bwd_kernel<raw>")) return rc;
  GradOlaArgs o{};
  o.n_res = 1;
  o.B = B;
  o.T = T;
  o.g = g_y;
  o.gfb[0] = a.gfb;
  o.n_fft[0] = plan->cfg.n_fft;
  o.hop[0] = plan->cfg.hop_length;
  o.Tf[0] = a.Tf;
  dim3 og(grid_for((T + 3) / 4, 256, 2), B);
  grad_ola_kernel<<<og, 256, 0, st>>>(o, MstftFinArgs{});
  return check_launch("grad_ola_kernel");
}
