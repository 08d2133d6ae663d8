// C-ABI entry points: ISTFT / Griffin-Lim (include/spectral_b200.h).  Host glue + the gl2 kernels.
#include <cstdlib>
#include <mutex>

#include "capi_common.cuh"
#include "gl2.cuh"

using namespace sb200;
using namespace sb200::host;

// ---- ISTFT / Griffin-Lim --------------------------------------------------------------------------

static int make_batch_frames(const sb200_plan* plan, const sb200_batch* b, int64_t length, GlBatch* out,
                             long long* total_frames, long long* max_rows_len) {
  if (!b || b->B < 1) return fail(SB200_ERR_INVALID, "batch: B must be >= 1");
  const int Q = 4096 / plan->cfg.n_fft, hop = plan->cfg.hop_length, N = plan->cfg.n_fft;   // frames per item
  GlBatch g{};
  g.bd.B = b->B;
  g.length = length;
  if (b->item_off == nullptr) {
    if (b->len < 1) return fail(SB200_ERR_INVALID, "frames batch: need at least one frame");
    g.bd.frames_per_row = b->len;
    g.bd.items_per_row = (b->len + Q - 1) / Q;
    g.bd.total_items = g.bd.items_per_row * b->B;
    const long long Ly = length > 0 ? length : static_cast<long long>(hop) * (b->len - 1);
    if (Ly < N / 4 + 1) return fail(SB200_ERR_INVALID, "output signal shorter than n_fft/4 + 1 samples");
    if (length > 0 && 1 + length / hop != b->len)
      return fail(SB200_ERR_INVALID, "length inconsistent with the number of frames: need 1 + length // hop == n_frames");
    g.bd.stride = b->stride > 0 ? b->stride : Ly;
    *total_frames = b->len * b->B;
    *max_rows_len = Ly;
  } else {
    if (!b->sig_off || !b->frame_off || (length > 0 && !b->sig_len)) return fail(SB200_ERR_INVALID, "ragged frames batch: missing table");
    g.bd.sig_off = reinterpret_cast<const long long*>(b->sig_off);
    g.bd.sig_len = reinterpret_cast<const long long*>(b->sig_len);
    g.bd.frame_off = reinterpret_cast<const long long*>(b->frame_off);
    g.bd.item_off = reinterpret_cast<const long long*>(b->item_off);
    g.bd.total_items = b->total_items;
    *total_frames = b->total_frames;
    *max_rows_len = b->len > 0 ? b->len : (1 << 20);   // ragged: `len` may carry the longest output row
  }
  *out = g;
  return SB200_OK;
}

// workspace layout: 4 signal buffers (2 parities x ping-pong) of gl2_sig_elems() floats, then tprev (form 1)
static int64_t gl2_sig_elems(const sb200_plan* plan, int64_t total_frames, int64_t B) {
  const int64_t n = total_frames * plan->cfg.hop_length + B * plan->cfg.win_length;
  return (n + 63) / 64 * 64;
}

int64_t sb200_griffinlim_workspace_bytes(const sb200_plan* plan, int64_t total_frames, int32_t n_rows, int32_t form) {
  if (!plan || total_frames < 1 || n_rows < 1) return -1;
  const int64_t F = plan->cfg.n_fft / 2 + 1;
  int64_t bytes = 4 * gl2_sig_elems(plan, total_frames, n_rows) * 4;
  if (form == 1) bytes += gl2_pair_slots(total_frames, n_rows) * F * 16;   // tprev: complex64 of both frames of a pair per bin
  return bytes + 512;
}

template <int N, int MODE>
static void launch_gl2_mode(const sb200_plan* plan, const Gl2Args& a, int grid, size_t smem, cudaStream_t st) {
  cudaFuncSetAttribute(gl2_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  gl2_kernel<N, MODE><<<grid, kGl2Warps * 32, smem, st>>>(plan->dev, a);
}

// SB200_GL_PERSISTENT=0 keeps one launch per iteration for every batch size (A/B measurements)
static bool gl2_persistent_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("SB200_GL_PERSISTENT");
    return !(e && e[0] == '0');
  }();
  return on;
}

static bool gl2_one_group_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("SB200_GL_ONE_GROUP");
    return !(e && e[0] == '0');
  }();
  return on;
}

// init (mode 0: complex spectrogram, mode 1: S * exp(2 pi i u)) -> n_iter iterations -> finish [-> inv_preemphasis]
// counter: 4 zeroable bytes of workspace for the grid barrier of the persistent kernel (null: never persistent)
template <int N>
static int launch_gl2(const sb200_plan* plan, Gl2Args a, int n_iter, int form, float* y, float inv_pre, float* sig,
                      long long sig_elems, long long max_row_len, cudaStream_t st, unsigned* counter = nullptr) {
  using C = Fft2Cfg<N>;
  constexpr int FT = kGl2GroupWarps * C::kFrames;
  const int hop = plan->cfg.hop_length;
  if (static_cast<long long>(FT) * hop < C::kWin - hop)   // a sample must be covered by at most two consecutive tiles
    return fail(SB200_ERR_UNSUPPORTED, "istft / griffinlim: hop_length too small for the tiled overlap-add (need tile_frames * hop >= win - hop)");
  const long long max_frames = a.g.bd.item_off == nullptr ? a.g.bd.frames_per_row : max_row_len / hop + 1;
  a.tiles_per_row = static_cast<int>((max_frames + FT - 1) / FT);
  const size_t smem = Gl2Smem<N>::bytes((FT - 1) * hop + C::kWin);
  const long long tiles = static_cast<long long>(a.g.bd.B) * a.tiles_per_row;
  const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((tiles + kGl2Groups - 1) / kGl2Groups, sm_count())));
  float* buf[2][2] = {{sig, sig + sig_elems}, {sig + 2 * sig_elems, sig + 3 * sig_elems}};
  int cur = 0;
  // small batch, Griffin-Lim proper: one cooperative launch for init + all iterations (gl2_persistent_kernel)
  if (!a.spec && n_iter > 0 && counter && tiles <= static_cast<long long>(kGl2Groups) * sm_count() && gl2_persistent_enabled()) {
    void (*kern)(const PlanDev, Gl2Args, int, float*, long long, unsigned*) =
        form == 0 ? gl2_persistent_kernel<N, 0> : gl2_persistent_kernel<N, 1>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cudaMemsetAsync(counter, 0, sizeof(unsigned), st);
    int pgrid = grid;
    if (tiles <= sm_count() && gl2_one_group_enabled()) {   // one tile per SM: the tile's latency is the iteration time
      a.groups_active = 1;
      pgrid = static_cast<int>(tiles);
    }
    PlanDev pd = plan->dev;
    long long se = sig_elems;
    void* args[] = {&pd, &a, &n_iter, &sig, &se, &counter};
    const cudaError_t ce = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(pgrid), dim3(kGl2Warps * 32), args, smem, st);
    if (ce == cudaSuccess) {
      if (int rc = check_launch("gl2_persistent_kernel")) return rc;
      cur = n_iter & 1;
      goto finish;
    }
    cudaGetLastError();   // not co-resident / not supported: fall back to one launch per iteration
    a.groups_active = 0;
  }
  a.ya_out = buf[0][0];
  a.yb_out = buf[0][1];
  if (a.spec) launch_gl2_mode<N, 0>(plan, a, grid, smem, st);
  else launch_gl2_mode<N, 1>(plan, a, grid, smem, st);
  if (int rc = check_launch("gl2_kernel (init)")) return rc;
  for (int it = 0; it < n_iter; ++it) {
    a.ya_in = buf[cur][0];
    a.yb_in = buf[cur][1];
    a.ya_out = buf[cur ^ 1][0];
    a.yb_out = buf[cur ^ 1][1];
    a.first = (it == 0);
    if (form == 0) launch_gl2_mode<N, 2>(plan, a, grid, smem, st);
    else launch_gl2_mode<N, 3>(plan, a, grid, smem, st);
    if (int rc = check_launch("gl2_kernel (iteration)")) return rc;
    cur ^= 1;
  }
finish:
  Gl2FinishArgs f{a.g, buf[cur][0], buf[cur][1], y};
  dim3 ogrid(grid_for(max_row_len, 256, 4), a.g.bd.B);
  gl2_finish_kernel<N><<<ogrid, 256, 0, st>>>(plan->dev, f);
  if (int rc = check_launch("gl2_finish_kernel")) return rc;
  if (inv_pre != 0.f) {
    BatchDev rows{};
    rows.B = a.g.bd.B;
    if (a.g.bd.item_off == nullptr) {
      rows.len = a.g.length > 0 ? a.g.length : static_cast<long long>(hop) * (a.g.bd.frames_per_row - 1);
      rows.stride = a.g.bd.stride;
      if (int rc = launch_inv_preemphasis(y, rows, inv_pre, y, st)) return rc;
    } else {
      return fail(SB200_ERR_UNSUPPORTED, "inv_preemph inside griffinlim is only supported for uniform batches; call sb200_inv_preemphasis");
    }
  }
  return SB200_OK;
}

int sb200_istft(const sb200_plan* plan, const float* spec, const sb200_batch* frames_batch, int64_t length, float* y,
                void* workspace, sb200_stream stream) {
  if (!plan || !spec || !y || !workspace) return fail(SB200_ERR_INVALID, "istft: null argument");
  Gl2Args a{};
  long long total_frames = 0, max_len = 0;
  if (int rc = make_batch_frames(plan, frames_batch, length, &a.g, &total_frames, &max_len)) return rc;
  a.spec = reinterpret_cast<const float2*>(spec);
  int rc = 0;
  SB200_DISPATCH_N(plan, rc = launch_gl2<kN>(plan, a, 0, 0, y, 0.f, static_cast<float*>(workspace),
                                             gl2_sig_elems(plan, total_frames, a.g.bd.B), max_len, static_cast<cudaStream_t>(stream)));
  return rc;
}

int sb200_griffinlim(const sb200_plan* plan, const float* S, const float* init_phase, const sb200_batch* frames_batch,
                     int64_t length, int32_t n_iter, float momentum, int32_t form, float inv_preemph, float* y,
                     void* workspace, sb200_stream stream) {
  if (!plan || !S || !init_phase || !y || !workspace) return fail(SB200_ERR_INVALID, "griffinlim: null argument");
  if (n_iter < 0 || (form != 0 && form != 1)) return fail(SB200_ERR_INVALID, "griffinlim: bad n_iter / form");
  if (form == 1 && !(momentum >= 0.f)) return fail(SB200_ERR_INVALID, "griffinlim: momentum must be >= 0");   // librosa ParameterError
  Gl2Args a{};
  long long total_frames = 0, max_len = 0;
  if (int rc = make_batch_frames(plan, frames_batch, length, &a.g, &total_frames, &max_len)) return rc;
  a.S = S;
  a.init_phase = init_phase;
  a.alpha = momentum / (1.f + momentum);
  const long long sig_elems = gl2_sig_elems(plan, total_frames, a.g.bd.B);
  float* sig = static_cast<float*>(workspace);
  a.tprev = reinterpret_cast<ulonglong2*>(sig + 4 * sig_elems);
  int rc = 0;
  // grid-barrier counter: first 256-byte boundary behind the signal buffers and tprev (inside the workspace's 512 bytes of slack)
  const size_t used = static_cast<size_t>(4) * sig_elems * sizeof(float) +
                      (form == 1 ? static_cast<size_t>(gl2_pair_slots(total_frames, a.g.bd.B)) * (plan->cfg.n_fft / 2 + 1) * 16 : 0);
  unsigned* counter = reinterpret_cast<unsigned*>(static_cast<char*>(workspace) + (used + 255) / 256 * 256);
  SB200_DISPATCH_N(plan, rc = launch_gl2<kN>(plan, a, n_iter, form, y, inv_preemph, sig, sig_elems, max_len,
                                             static_cast<cudaStream_t>(stream), counter));
  return rc;
}

