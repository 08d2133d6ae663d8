// Launcher of the packed-engine STFT + mel feature kernel (the hot path of sb200_stft_features).
#include "capi_common.cuh"
#include <cstdlib>

#include "feat3.cuh"

using namespace sb200;
using namespace sb200::host;

namespace {

template <int N, bool PRE, bool LOGMAG, int HS>
void launch_features2_t(const sb200_plan* plan, const FeatArgs& a, int grid, size_t smem, cudaStream_t st) {
  cudaFuncSetAttribute(stft_feature2_kernel<N, PRE, LOGMAG, HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  stft_feature2_kernel<N, PRE, LOGMAG, HS><<<grid, kFeat2Warps * 32, smem, st>>>(plan->dev, a);
}

template <int N, bool PRE, int HS>
void launch_features3_t(const sb200_plan* plan, const FeatArgs& a, int grid, size_t smem, cudaStream_t st) {
  cudaFuncSetAttribute(stft_feature3_kernel<N, PRE, HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  stft_feature3_kernel<N, PRE, HS><<<grid, kFeat3Threads, smem, st>>>(plan->dev, a);
}

// SB200_FEAT_KERNEL=2 selects the single-role kernel (feat2.cuh) for A/B measurements; default: warp-specialised (feat3.cuh)
bool use_feat3(size_t smem3) {
  static const int forced = [] {
    const char* e = std::getenv("SB200_FEAT_KERNEL");
    return e ? std::atoi(e) : 0;
  }();
  return forced != 2 && smem3 <= 227u * 1024u;
}

// Warp-specialised packed engine (the hot path).  Returns false if the configuration does not fit (tables too large).
template <int N>
bool launch_features3(const sb200_plan* plan, const FeatArgs& a, cudaStream_t st) {
  if constexpr (N != 2048) {
    return false;   // n_fft 1024 / 512 (2 / 4 frame pairs per item) stay on the single-role kernel
  } else {
  const size_t smem = feat3_smem_bytes<N>(plan->dev);
  if (!use_feat3(smem)) return false;
  const long long ctas_needed = (a.bd.total_items + kFeat3Pairs - 1) / kFeat3Pairs;
  const int grid = static_cast<int>(std::min<long long>(ctas_needed, sm_count()));
  const bool pre = a.pre != 0.f;
  if (plan->cfg.hop_length == 256) {   // the reference hop (hparam.py): frames of a pair share 3/4 of their samples
    if (pre) launch_features3_t<N, true, 4>(plan, a, grid, smem, st);
    else launch_features3_t<N, false, 4>(plan, a, grid, smem, st);
  } else {
    if (pre) launch_features3_t<N, true, 0>(plan, a, grid, smem, st);
    else launch_features3_t<N, false, 0>(plan, a, grid, smem, st);
  }
  return true;
  }
}

// Packed engine: magnitude / mel features.
template <int N>
int launch_features2(const sb200_plan* plan, const FeatArgs& a, cudaStream_t st) {
  if (launch_features3<N>(plan, a, st)) return check_launch("stft_feature3_kernel");
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

int sb200::host::launch_features2_any(const sb200_plan* plan, const FeatArgs& a, cudaStream_t st) {
  int rc = 0;
  SB200_DISPATCH_N(plan, rc = launch_features2<kN>(plan, a, st));
  return rc;
}
