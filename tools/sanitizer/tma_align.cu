// Which element offsets may a bulk-tensor (TMA) box start at?  Loads 256 floats from x[off ..] through (a) a plain 1-D tensor map,
// (b) the 2-D overlapping-row view t[j][c] = x[256 j + c] used by feat3.cuh, for off = 0..8, one launch per case, and reports
// the launch status and whether the data arrived.   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_align tma_align.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, float* out, int n) {
  __shared__ __align__(128) float buf[2048];
  __shared__ __align__(8) unsigned long long bar;
  const unsigned b = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(b), "r"(n * 4) : "memory");
    if (RANK == 1)
      asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];"
                   ::"r"(smem_u32(buf)), "l"(reinterpret_cast<unsigned long long>(&map)), "r"(c0), "r"(b) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(buf)), "l"(reinterpret_cast<unsigned long long>(&map)), "r"(c0), "r"(0), "r"(b) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(b) : "memory");
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = buf[i];
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(fn);
  const int total = 1 << 16;
  std::vector<float> h(total);
  for (int i = 0; i < total; ++i) h[i] = static_cast<float>(i);
  float *x, *out;
  cudaMalloc(&x, total * 4);
  cudaMalloc(&out, 2048 * 4);
  cudaMemcpy(x, h.data(), total * 4, cudaMemcpyHostToDevice);
  const cuuint32_t ones[2] = {1, 1};
  for (int rank = 1; rank <= 2; ++rank) {
    CUtensorMap map;
    CUresult r;
    int n;
    if (rank == 1) {
      const cuuint64_t dims[1] = {total};
      const cuuint64_t strides[1] = {0};
      const cuuint32_t box[1] = {256};
      n = 256;
      r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 1, x, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      const cuuint64_t dims[2] = {total - 1280, 6};
      const cuuint64_t strides[1] = {1024};
      const cuuint32_t box[2] = {256, 6};
      n = 1536;
      r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    printf("rank %d: encode -> %d\n", rank, static_cast<int>(r));
    if (r != CUDA_SUCCESS) continue;
    for (int off = 0; off <= 8; ++off) {
      cudaMemset(out, 0, 2048 * 4);
      if (rank == 1) probe<1><<<1, 128>>>(map, off, out, n);
      else probe<2><<<1, 128>>>(map, off, out, n);
      cudaError_t e = cudaDeviceSynchronize();
      float g[2048];
      bool ok = false;
      if (e == cudaSuccess) {
        cudaMemcpy(g, out, n * 4, cudaMemcpyDeviceToHost);
        ok = true;
        for (int i = 0; i < n; ++i) ok = ok && g[i] == static_cast<float>(off + i);
      }
      printf("  rank %d offset %d: %s, data %s\n", rank, off, cudaGetErrorString(e), ok ? "ok" : "WRONG");
      if (e != cudaSuccess) { printf("  (context lost; stopping)\n"); return 0; }
    }
  }
  return 0;
}
