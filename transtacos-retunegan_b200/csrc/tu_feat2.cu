// Launcher of the packed-engine STFT + mel feature kernel (the hot path of sb200_stft_features).
#include "capi_common.cuh"
#include <cstdio>
#include <cstdlib>

#include "feat4.cuh"

using namespace sb200;
using namespace sb200::host;

namespace {

template <int N, bool PRE, bool LOGMAG, int HS>
void launch_features2_t(const sb200_plan* plan, const FeatArgs& a, int grid, size_t smem, cudaStream_t st) {
  cudaFuncSetAttribute(stft_feature2_kernel<N, PRE, LOGMAG, HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  stft_feature2_kernel<N, PRE, LOGMAG, HS><<<grid, kFeat2Warps * 32, smem, st>>>(plan->dev, a);
}

template <int N, bool PRE, int HS, bool TMA>
void launch_features3_t(const sb200_plan* plan, const FeatArgs& a, int grid, size_t smem, cudaStream_t st, const CUtensorMap& map,
                        StageArgs sa) {
  cudaFuncSetAttribute(stft_feature3_kernel<N, PRE, HS, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  stft_feature3_kernel<N, PRE, HS, TMA><<<grid, kFeat3Threads, smem, st>>>(plan->dev, a, map, sa);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static const EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      f = nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// Tensor map of the flat sample buffer for the staging copies of feat3.cuh (2-D view with overlapping rows, see kStageBoxCols
// there).  Returns the number of samples the buffer spans, or 0 when staging is off or the buffer cannot be described
// (misaligned base, too long, no encoder): the caller then launches the global-load kernel.
// OFF BY DEFAULT (SB200_FEAT_TMA=1 turns it on): measured on B200, config 3, the staged kernel is SLOWER, 90.9 us against 87.3 us
// (profiles/r02_feat3_ncu_summary.md).  The kernel is bound by shared-memory wavefronts and issue slots, not by the latency of the
// gather: loads straight into registers cost no shared-memory traffic, the staged samples cost 42 wavefronts to land and 80 to read
// per item (+9 %), and removing the exposed global-load latency (long-scoreboard stalls 1.75 -> 1.51 per issue) buys nothing
// because the scheduler's other warps were covering it.
long long make_stage_map(const FeatArgs& a, CUtensorMap* map) {
  static const int mode = [] {
    const char* e = std::getenv("SB200_FEAT_TMA");
    return e ? std::atoi(e) : 0;
  }();
  static const bool debug = std::getenv("SB200_FEAT_DEBUG") != nullptr;
  if (mode == 0 || encode_tiled() == nullptr) return 0;
  const long long total = a.bd.item_off == nullptr ? (a.bd.B - 1) * a.bd.stride + a.bd.len : a.bd.total_samples;
  if ((reinterpret_cast<uintptr_t>(a.x) & 15) != 0 || total < 4096 || total >= (1LL << 31)) return 0;
  const cuuint32_t ones[2] = {1, 1};
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(total - (kStageBoxRows - 1) * kStageBoxCols), kStageBoxRows};
  const cuuint64_t strides[1] = {kStageBoxCols * sizeof(float)};
  const cuuint32_t box[2] = {kStageBoxCols, kStageBoxRows};
  const CUresult r = encode_tiled()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(a.x), dims, strides, box, ones,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (debug) std::fprintf(stderr, "[sb200] stage map: x %p, %lld samples, encode -> %d\n", static_cast<const void*>(a.x), total, static_cast<int>(r));
  return r == CUDA_SUCCESS ? total : 0;
}

// SB200_FEAT_KERNEL selects the n_fft 2048 feature kernel for A/B measurements: 2 = single role (feat2.cuh), 3 = 8 analysis +
// 8 epilogue warps (feat3.cuh, the default), 4 = 12 transform + 4 mel warps (feat4.cuh).  Measured on config 3 (tools/ab_feat.sh):
// 97.6 / 86.8 / 96.9 us.
int feat_kernel_choice() {
  static const int forced = [] {
    const char* e = std::getenv("SB200_FEAT_KERNEL");
    return e ? std::atoi(e) : 0;
  }();
  return forced;
}
bool use_feat3(size_t smem3) { return feat_kernel_choice() != 2 && smem3 <= 227u * 1024u; }

template <int N, bool PRE, bool LOGMAG, int HS>
void launch_features4_t(const sb200_plan* plan, const FeatArgs& a, int grid, size_t smem, cudaStream_t st) {
  cudaFuncSetAttribute(stft_feature4_kernel<N, PRE, LOGMAG, HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  stft_feature4_kernel<N, PRE, LOGMAG, HS><<<grid, kF4Threads, smem, st>>>(plan->dev, a);
}

// 12 + 4 kernel (n_fft 2048).  Returns false if not selected or the tables do not fit.
template <int N>
bool launch_features4(const sb200_plan* plan, const FeatArgs& a, cudaStream_t st) {
  if constexpr (N != 2048) {
    return false;
  } else {
    const int choice = feat_kernel_choice();
    if (choice != 4) return false;
    const size_t smem = feat4_smem_bytes<N>(plan->dev);
    if (smem > 227u * 1024u) return false;
    const long long ctas_needed = (a.bd.total_items + kF4Fft - 1) / kF4Fft;
    const int grid = static_cast<int>(std::min<long long>(ctas_needed, sm_count()));
    const bool pre = a.pre != 0.f, lg = a.mag_scale.log != 0;
    if (plan->cfg.hop_length == 256) {
      if (pre && lg) launch_features4_t<N, true, true, 4>(plan, a, grid, smem, st);
      else if (pre) launch_features4_t<N, true, false, 4>(plan, a, grid, smem, st);
      else if (lg) launch_features4_t<N, false, true, 4>(plan, a, grid, smem, st);
      else launch_features4_t<N, false, false, 4>(plan, a, grid, smem, st);
    } else {
      if (pre && lg) launch_features4_t<N, true, true, 0>(plan, a, grid, smem, st);
      else if (pre) launch_features4_t<N, true, false, 0>(plan, a, grid, smem, st);
      else if (lg) launch_features4_t<N, false, true, 0>(plan, a, grid, smem, st);
      else launch_features4_t<N, false, false, 0>(plan, a, grid, smem, st);
    }
    return true;
  }
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
  CUtensorMap map{};
  StageArgs sa{0};
  if (plan->cfg.hop_length == 256) {   // the reference hop (hparam.py): frames of a pair share 3/4 of their samples
    sa.total = make_stage_map(a, &map);
    if (sa.total > 0) {                // interior items staged by bulk-tensor copies
      if (pre) launch_features3_t<N, true, 4, true>(plan, a, grid, smem, st, map, sa);
      else launch_features3_t<N, false, 4, true>(plan, a, grid, smem, st, map, sa);
    } else {
      if (pre) launch_features3_t<N, true, 4, false>(plan, a, grid, smem, st, map, sa);
      else launch_features3_t<N, false, 4, false>(plan, a, grid, smem, st, map, sa);
    }
  } else {
    if (pre) launch_features3_t<N, true, 0, false>(plan, a, grid, smem, st, map, sa);
    else launch_features3_t<N, false, 0, false>(plan, a, grid, smem, st, map, sa);
  }
  return true;
  }
}

// Packed engine: magnitude / mel features.
template <int N>
int launch_features2(const sb200_plan* plan, const FeatArgs& a, cudaStream_t st) {
  if (launch_features4<N>(plan, a, st)) return check_launch("stft_feature4_kernel");
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
