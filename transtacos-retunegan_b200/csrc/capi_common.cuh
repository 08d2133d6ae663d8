// Host-side helpers shared by the translation units of libspectral_b200.so (defined in spectral_b200.cu).
// The library is built as one .o per kernel family (feat2 / gl2 / mstft / rest), compiled in parallel without
// --split-compile: ptxas then produces the same code for the same source on every build.
#pragma once
#include <algorithm>
#include <string>

#include "feat.cuh"

namespace sb200 {
namespace host {

int fail(int st, const std::string& msg);      // records the thread's last error string, returns st
int check_launch(const char* what);            // counts the launch, maps cudaGetLastError to a status
int sm_count();
inline ScaleDev to_dev(const sb200_scale& s) { return ScaleDev{s.log, s.a, s.b, s.floor}; }
inline int grid_for(long long n, int block, int cap_mult = 8) {
  const long long need = (n + block - 1) / block;
  return static_cast<int>(std::max<long long>(1, std::min<long long>(need, static_cast<long long>(cap_mult) * sm_count())));
}

// tu_feat2.cu: the packed-engine feature kernel (hot path)
int launch_features2_any(const sb200_plan* plan, const FeatArgs& a, cudaStream_t st);
// spectral_b200.cu: y[n] = x[n] + k y[n-1] over the rows of a batch (x == y allowed)
int launch_inv_preemphasis(const float* x, const BatchDev& rows, float k, float* y, cudaStream_t st);

}  // namespace host
}  // namespace sb200

#define SB200_DISPATCH_N(plan, ...)                           \
  switch ((plan)->cfg.n_fft) {                                \
    case 2048: { constexpr int kN = 2048; __VA_ARGS__; } break; \
    case 1024: { constexpr int kN = 1024; __VA_ARGS__; } break; \
    default:   { constexpr int kN = 512;  __VA_ARGS__; } break; \
  }
