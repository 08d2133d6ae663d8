// C-ABI entry points of libspectral_b200.so (see include/spectral_b200.h).
// Host-side glue only: argument validation, launch configuration, kernel launches on the caller's stream.
#include <atomic>
#include <cstdio>
#include <mutex>
#include <string>

#include "capi_common.cuh"
#include "feat.cuh"
#include "misc.cuh"

using namespace sb200;
using namespace sb200::host;

namespace {
thread_local std::string g_err;
std::atomic<long long> g_launches{0};
}  // namespace

int sb200::host::fail(int st, const std::string& msg) {
  g_err = msg;
  return st;
}

int sb200::host::check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SB200_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return SB200_OK;
}

int sb200::host::sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

namespace {

// Signal-described batch (STFT direction): frames = 1 + len/hop.
int make_batch_signal(const sb200_plan* plan, const sb200_batch* b, BatchDev* out, long long* total_frames) {
  if (!b || b->B < 1) return fail(SB200_ERR_INVALID, "batch: B must be >= 1");
  const int Q = 4096 / plan->cfg.n_fft, hop = plan->cfg.hop_length;   // frames per item
  BatchDev d{};
  d.B = b->B;
  if (b->item_off == nullptr) {
    if (b->len < plan->cfg.n_fft / 4 + 1)
      return fail(SB200_ERR_INVALID, "signal shorter than n_fft/4 + 1 samples: reflect padding undefined");
    d.len = b->len;
    d.stride = b->stride > 0 ? b->stride : b->len;
    d.frames_per_row = 1 + b->len / hop;
    d.items_per_row = (d.frames_per_row + Q - 1) / Q;
    d.total_items = d.items_per_row * b->B;
    *total_frames = d.frames_per_row * b->B;
  } else {
    if (!b->sig_off || !b->sig_len || !b->frame_off) return fail(SB200_ERR_INVALID, "ragged batch: missing offset table");
    d.sig_off = reinterpret_cast<const long long*>(b->sig_off);
    d.sig_len = reinterpret_cast<const long long*>(b->sig_len);
    d.frame_off = reinterpret_cast<const long long*>(b->frame_off);
    d.item_off = reinterpret_cast<const long long*>(b->item_off);
    d.total_items = b->total_items;
    d.total_samples = b->total_samples;
    *total_frames = b->total_frames;
  }
  *out = d;
  return SB200_OK;
}

// Scalar engine: only used when the complex STFT itself is requested (get_stft_torch).
template <int N>
int launch_features_spec(const sb200_plan* plan, const FeatArgs& a, cudaStream_t st) {
  const size_t smem = feat_smem_bytes<N>(plan->dev);
  const long long ctas_needed = (2 * a.bd.total_items + kFeatWarps - 1) / kFeatWarps;
  const int grid = static_cast<int>(std::min<long long>(ctas_needed, 2LL * sm_count()));
  if (a.pre != 0.f) {
    cudaFuncSetAttribute(stft_feature_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    stft_feature_kernel<N, true><<<grid, kFeatWarps * 32, smem, st>>>(plan->dev, a);
  } else {
    cudaFuncSetAttribute(stft_feature_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    stft_feature_kernel<N, false><<<grid, kFeatWarps * 32, smem, st>>>(plan->dev, a);
  }
  return check_launch("stft_feature_kernel");
}

}  // namespace

// All sb200_* definitions below get C linkage from their declarations in include/spectral_b200.h.

const char* sb200_version(void) { return "spectral_b200 0.1 (sm_100a)"; }
const char* sb200_last_error_string(void) { return g_err.c_str(); }
int64_t sb200_launch_count(void) { return g_launches.load(); }

int sb200_plan_create(const sb200_config* cfg, sb200_plan** out) {
  if (!cfg || !out) return fail(SB200_ERR_INVALID, "plan_create: null argument");
  int devcount = 0;
  if (cudaGetDeviceCount(&devcount) != cudaSuccess || devcount == 0) {
    cudaGetLastError();
    return fail(SB200_ERR_CUDA, "plan_create: no CUDA device (this library has no CPU path)");
  }
  sb200_plan* p = new sb200_plan();
  int st = 0;
  const std::string msg = build_plan(*cfg, p, &st);
  if (!msg.empty()) {
    for (void* d : p->allocs) cudaFree(d);
    delete p;
    return fail(st, "plan_create: " + msg);
  }
  *out = p;
  return SB200_OK;
}

int sb200_plan_destroy(sb200_plan* plan) {
  if (!plan) return SB200_OK;
  for (void* d : plan->allocs) cudaFree(d);
  delete plan;
  return SB200_OK;
}

int sb200_plan_frames_per_pass(const sb200_plan* plan) { return plan ? 4096 / plan->cfg.n_fft : 0; }

int sb200_plan_mel_basis_host(const sb200_plan* plan, float* out_host) {
  if (!plan || !out_host) return fail(SB200_ERR_INVALID, "mel_basis_host: null argument");
  std::memcpy(out_host, plan->mel_dense.data(), plan->mel_dense.size() * sizeof(float));
  return SB200_OK;
}

int sb200_plan_window_host(const sb200_plan* plan, float* out_host) {
  if (!plan || !out_host) return fail(SB200_ERR_INVALID, "window_host: null argument");
  std::memcpy(out_host, plan->window_f32.data(), plan->window_f32.size() * sizeof(float));
  return SB200_OK;
}

int sb200_stft_features(const sb200_plan* plan, const float* x, const sb200_batch* batch, float preemph,
                        sb200_scale mag_scale, sb200_scale mel_scale, float* mag, float* mel, float* spec,
                        sb200_stream stream) {
  if (!plan || !x) return fail(SB200_ERR_INVALID, "stft_features: null argument");
  if (!mag && !mel && !spec) return fail(SB200_ERR_INVALID, "stft_features: no output requested");
  FeatArgs a{};
  long long total_frames = 0;
  if (int rc = make_batch_signal(plan, batch, &a.bd, &total_frames)) return rc;
  a.x = x;
  a.pre = preemph;
  a.mag_scale = to_dev(mag_scale);
  a.mel_scale = to_dev(mel_scale);
  a.mag = mag;
  a.mel = mel;
  a.spec = reinterpret_cast<float2*>(spec);
  int rc = 0;
  if (spec) {
    SB200_DISPATCH_N(plan, rc = launch_features_spec<kN>(plan, a, static_cast<cudaStream_t>(stream)));
  } else {
    rc = launch_features2_any(plan, a, static_cast<cudaStream_t>(stream));
  }
  return rc;
}

int sb200_mel_project(const sb200_plan* plan, const float* in, int64_t frames, sb200_scale scale, float* out,
                      sb200_stream stream) {
  if (!plan || !in || !out || frames < 1) return fail(SB200_ERR_INVALID, "mel_project: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SB200_DISPATCH_N(plan, {
    const size_t smem = feat_smem_bytes<kN>(plan->dev);
    const long long items = (frames + FftCfg<kN>::kQ - 1) / FftCfg<kN>::kQ;
    const int grid = static_cast<int>(std::min<long long>((items + kFeatWarps - 1) / kFeatWarps, 2LL * sm_count()));
    cudaFuncSetAttribute(mel_project_kernel<kN>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    mel_project_kernel<kN><<<grid, kFeatWarps * 32, smem, st>>>(plan->dev, in, frames, to_dev(scale), out);
  });
  return check_launch("mel_project_kernel");
}

int sb200_mel_to_linear(const sb200_plan* plan, const float* mel, int64_t frames, float* out, sb200_stream stream) {
  if (!plan || !mel || !out || frames < 1) return fail(SB200_ERR_INVALID, "mel_to_linear: bad argument");
  mel_to_linear_kernel<<<grid_for(frames * plan->dev.F, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(plan->dev, mel, frames, out);
  return check_launch("mel_to_linear_kernel");
}

int sb200_spec_to_amplitude(const float* in, int64_t n, int32_t mode, float p0, float p1, float p2, float power,
                            float* out, sb200_stream stream) {
  if (!in || !out || n < 1 || mode < 0 || mode > 2) return fail(SB200_ERR_INVALID, "spec_to_amplitude: bad argument");
  spec_to_amplitude_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, n, mode, p0, p1, p2, power, out);
  return check_launch("spec_to_amplitude_kernel");
}

static int make_batch_rows(const sb200_batch* b, BatchDev* out) {
  if (!b || b->B < 1) return fail(SB200_ERR_INVALID, "batch: B must be >= 1");
  BatchDev d{};
  d.B = b->B;
  if (b->sig_off == nullptr) {
    if (b->len < 1) return fail(SB200_ERR_INVALID, "batch: len must be >= 1");
    d.len = b->len;
    d.stride = b->stride > 0 ? b->stride : b->len;
  } else {
    if (!b->sig_len) return fail(SB200_ERR_INVALID, "ragged batch: missing sig_len");
    d.sig_off = reinterpret_cast<const long long*>(b->sig_off);
    d.sig_len = reinterpret_cast<const long long*>(b->sig_len);
  }
  *out = d;
  return SB200_OK;
}

int sb200_frame_stats(const float* x, const sb200_batch* batch, int32_t frame_length, int32_t hop_length, float* rms,
                      float* zcr, sb200_stream stream) {
  if (!x || (!rms && !zcr)) return fail(SB200_ERR_INVALID, "frame_stats: null argument");
  if (frame_length < 2 || hop_length < 1) return fail(SB200_ERR_INVALID, "frame_stats: need frame_length >= 2 and hop_length >= 1");
  FrameStatsArgs a{};
  if (int rc = make_batch_rows(batch, &a.bd)) return rc;
  if (a.bd.sig_off) {
    if (!batch->frame_off) return fail(SB200_ERR_INVALID, "frame_stats: ragged batch needs frame_off for this hop_length");
    a.bd.frame_off = reinterpret_cast<const long long*>(batch->frame_off);
    a.total_frames = batch->total_frames;
  } else {
    a.bd.frames_per_row = 1 + a.bd.len / hop_length;
    a.total_frames = a.bd.frames_per_row * a.bd.B;
  }
  if (a.total_frames < 1) return fail(SB200_ERR_INVALID, "frame_stats: no frames");
  a.x = x;
  a.frame_length = frame_length;
  a.hop = hop_length;
  a.rms = rms;
  a.zcr = zcr;
  frame_stats_kernel<<<grid_for(a.total_frames * 32, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return check_launch("frame_stats_kernel");
}

int sb200_trim_bounds(const float* rms, const int64_t* frame_off, int64_t frames_per_row, int32_t B, float top_db,
                      int64_t* bounds, sb200_stream stream) {
  if (!rms || !bounds || B < 1 || (!frame_off && frames_per_row < 1)) return fail(SB200_ERR_INVALID, "trim_bounds: bad argument");
  TrimArgs a{};
  a.rms = rms;
  a.frame_off = reinterpret_cast<const long long*>(frame_off);
  a.frames_per_row = frames_per_row;
  a.B = B;
  a.top_db = static_cast<double>(top_db);
  a.out = reinterpret_cast<long long*>(bounds);
  trim_bounds_kernel<<<grid_for(static_cast<long long>(B) * 32, 128, 8), 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return check_launch("trim_bounds_kernel");
}

int sb200_yin(const float* x, const sb200_batch* batch, int32_t sample_rate, float fmin, float fmax, int32_t frame_length,
              int32_t hop_length, float trough_threshold, float* f0, sb200_stream stream) {
  if (!x || !f0) return fail(SB200_ERR_INVALID, "yin: null argument");
  if (frame_length < 4 || frame_length > kYinMaxFrame || hop_length < 1 || sample_rate < 1)
    return fail(SB200_ERR_INVALID, "yin: need 4 <= frame_length <= 4096, hop_length >= 1");
  if (!(fmin > 0.f) || !(fmax > fmin)) return fail(SB200_ERR_INVALID, "yin: need 0 < fmin < fmax");
  YinArgs a{};
  if (int rc = make_batch_rows(batch, &a.bd)) return rc;
  if (a.bd.sig_off) {
    if (!batch->frame_off) return fail(SB200_ERR_INVALID, "yin: ragged batch needs frame_off for this hop_length");
    a.bd.frame_off = reinterpret_cast<const long long*>(batch->frame_off);
    a.total_frames = batch->total_frames;
  } else {
    a.bd.frames_per_row = 1 + a.bd.len / hop_length;
    a.total_frames = a.bd.frames_per_row * a.bd.B;
  }
  if (a.total_frames < 1) return fail(SB200_ERR_INVALID, "yin: no frames");
  const int win = frame_length / 2;
  // librosa: min_period = max(floor(sr / fmax), 1); max_period = min(ceil(sr / fmin), frame_length - win_length - 1)
  a.pmin = std::max(static_cast<int>(std::floor(static_cast<double>(sample_rate) / fmax)), 1);
  a.pmax = std::min(static_cast<int>(std::ceil(static_cast<double>(sample_rate) / fmin)), frame_length - win - 1);
  if (a.pmax < a.pmin + 2) return fail(SB200_ERR_INVALID, "yin: period range too small for this frame_length");
  a.x = x;
  a.frame_length = frame_length;
  a.hop = hop_length;
  a.sr = static_cast<float>(sample_rate);
  a.threshold = trough_threshold;
  a.f0 = f0;
  const size_t smem = sizeof(float) * yin_smem_floats(frame_length, a.pmax, a.pmin);
  if (frame_length % 8) return fail(SB200_ERR_INVALID, "yin: frame_length must be a multiple of 8");
  const int grid = static_cast<int>(std::min<long long>(a.total_frames, 16LL * sm_count()));
  yin_kernel<<<grid, kYinThreads, smem, static_cast<cudaStream_t>(stream)>>>(a);
  return check_launch("yin_kernel");
}

static int pool_loss_grid() { return 4 * sm_count(); }
int64_t sb200_pool_loss_workspace_bytes(void) { return static_cast<int64_t>(pool_loss_grid()) * 8 * sizeof(float) + 256; }

int sb200_pool_loss(const float* y, const float* y_g, int32_t B, int64_t T, int32_t pool_k, int32_t mode, float* loss,
                    float* grad_yg, void* workspace, sb200_stream stream) {
  if (!y || !y_g || !loss || !workspace) return fail(SB200_ERR_INVALID, "pool_loss: null argument");
  if (B < 1 || pool_k < 1 || T < pool_k || (mode != 0 && mode != 1))
    return fail(SB200_ERR_INVALID, "pool_loss: need B >= 1, 1 <= pool_k <= T, mode 0 or 1");
  PoolLossArgs a{};
  a.y = y;
  a.yg = y_g;
  a.B = B;
  a.k = pool_k;
  a.mode = mode;
  a.T = T;
  a.n_win = T / pool_k;
  a.grad = grad_yg;
  a.partials = static_cast<float*>(workspace);
  a.inv_n = static_cast<float>(1.0 / (static_cast<double>(B) * a.n_win));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = static_cast<int>(std::min<long long>(pool_loss_grid(), (static_cast<long long>(B) * a.n_win + 7) / 8));
  pool_loss_kernel<<<grid, 256, 0, st>>>(a);
  if (int rc = check_launch("pool_loss_kernel")) return rc;
  pool_loss_finalize_kernel<<<1, 256, 0, st>>>(a.partials, grid * 8, a.inv_n, loss);
  return check_launch("pool_loss_finalize_kernel");
}

int sb200_preemphasis(const float* x, const sb200_batch* batch, float k, float* y, sb200_stream stream) {
  if (!x || !y) return fail(SB200_ERR_INVALID, "preemphasis: null argument");
  BatchDev bd;
  if (int rc = make_batch_rows(batch, &bd)) return rc;
  const long long per_row = bd.sig_off ? (1 << 20) : bd.len;
  dim3 grid(grid_for(per_row, 256, 4), bd.B);
  preemphasis_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, bd, k, y);
  return check_launch("preemphasis_kernel");
}

int sb200_inv_preemphasis(const float* x, const sb200_batch* batch, float k, float* y, sb200_stream stream) {
  if (!x || !y) return fail(SB200_ERR_INVALID, "inv_preemphasis: null argument");
  BatchDev bd;
  if (int rc = make_batch_rows(batch, &bd)) return rc;
  return host::launch_inv_preemphasis(x, bd, k, y, static_cast<cudaStream_t>(stream));
}

int sb200::host::launch_inv_preemphasis(const float* x, const BatchDev& rows, float k, float* y, cudaStream_t st) {
  inv_preemphasis_kernel<<<rows.B, kScanThreads, 0, st>>>(x, rows, k, y);
  return check_launch("inv_preemphasis_kernel");
}

