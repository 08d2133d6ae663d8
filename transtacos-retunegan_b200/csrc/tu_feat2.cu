// Launcher of the packed-engine STFT + mel feature kernel (the hot path of sb200_stft_features).
#include "capi_common.cuh"
#include "feat2.cuh"

using namespace sb200;
using namespace sb200::host;

namespace {

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

int sb200::host::launch_features2_any(const sb200_plan* plan, const FeatArgs& a, cudaStream_t st) {
  int rc = 0;
  SB200_DISPATCH_N(plan, rc = launch_features2<kN>(plan, a, st));
  return rc;
}
