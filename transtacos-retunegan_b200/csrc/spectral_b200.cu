// C-ABI entry points of libspectral_b200.so (see include/spectral_b200.h).
// Host-side glue only: argument validation, launch configuration, kernel launches on the caller's stream.
#include <atomic>
#include <cstdio>
#include <mutex>
#include <string>

#include "feat.cuh"
#include "feat2.cuh"
#include "gl.cuh"
#include "gl2.cuh"
#include "mstft.cuh"
#include "misc.cuh"

using namespace sb200;

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int st, const std::string& msg) {
  g_err = msg;
  return st;
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SB200_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return SB200_OK;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

ScaleDev to_dev(const sb200_scale& s) { return ScaleDev{s.log, s.a, s.b, s.floor}; }

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

template <int N, bool PRE, bool LOGMAG, int HS>
void launch_features2_t(const sb200_plan* plan, const FeatArgs& a, int grid, size_t smem, cudaStream_t st) {
  cudaFuncSetAttribute(stft_feature2_kernel<N, PRE, LOGMAG, HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  stft_feature2_kernel<N, PRE, LOGMAG, HS><<<grid, kFeat2Warps * 32, smem, st>>>(plan->dev, a);
}

// Packed engine: magnitude / mel features (the hot path).
template <int N>
int launch_features2(const sb200_plan* plan, const FeatArgs& a, cudaStream_t st) {
  const size_t smem = feat2_smem_bytes<N>(plan->dev);
  const long long ctas_needed = (a.bd.total_items + kFeat2Warps - 1) / kFeat2Warps;
  const int grid = static_cast<int>(std::min<long long>(ctas_needed, sm_count()));
  const bool pre = a.pre != 0.f, lg = a.mag_scale.log != 0;
  if constexpr (N == 2048) {
    if (plan->cfg.hop_length == 256) {   // the reference hop (hparam.py): frames of a pair share 3/4 of their samples
      if (pre && lg) launch_features2_t<N, true, true, 4>(plan, a, grid, smem, st);
      else if (pre) launch_features2_t<N, true, false, 4>(plan, a, grid, smem, st);
      else if (lg) launch_features2_t<N, false, true, 4>(plan, a, grid, smem, st);
      else launch_features2_t<N, false, false, 4>(plan, a, grid, smem, st);
      return check_launch("stft_feature2_kernel");
    }
  }
  if (pre && lg) launch_features2_t<N, true, true, 0>(plan, a, grid, smem, st);
  else if (pre) launch_features2_t<N, true, false, 0>(plan, a, grid, smem, st);
  else if (lg) launch_features2_t<N, false, true, 0>(plan, a, grid, smem, st);
  else launch_features2_t<N, false, false, 0>(plan, a, grid, smem, st);
  return check_launch("stft_feature2_kernel");
}

}  // namespace

#define SB200_DISPATCH_N(plan, ...)                           \
  switch ((plan)->cfg.n_fft) {                                \
    case 2048: { constexpr int kN = 2048; __VA_ARGS__; } break; \
    case 1024: { constexpr int kN = 1024; __VA_ARGS__; } break; \
    default:   { constexpr int kN = 512;  __VA_ARGS__; } break; \
  }

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
    SB200_DISPATCH_N(plan, rc = launch_features2<kN>(plan, a, static_cast<cudaStream_t>(stream)));
  }
  return rc;
}

#include "capi_rest.inc"
