// Microbenchmarks: issue throughput of scalar vs packed FP32, MUFU, LDS/STS widths, SHFL on sm_100a.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define NACC 16
#define UNROLL 4
enum { K_FFMA, K_FFMA2, K_FADD, K_FADD2, K_FMUL2, K_MIX2, K_MUFU_LG2, K_MUFU_SQRT, K_MUFU_RSQ, K_LDS32, K_LDS64, K_LDS128, K_STS64, K_STS128, K_SHFL, K_FFMA_SHFL, K_FFMA2_LDS128, K_IADD, K_FFMA2_IADD, K_COUNT };
const char* names[] = {"FFMA","FFMA2","FADD","FADD2","FMUL2","FADD2+FFMA2","MUFU.LG2","MUFU.SQRT","MUFU.RSQ","LDS.32","LDS.64","LDS.128","STS.64","STS.128","SHFL","FFMA+SHFL(1:1)","FFMA2+LDS128(4:1)","IADD3","FFMA2+IADD(1:1)"};

template <int KIND>
__global__ void __launch_bounds__(1024) bench(float* out, long long* cyc, int iters) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  for (int i = tid; i < 8192; i += blockDim.x) sm[i] = i * 0.001f;
  __syncthreads();
  float a[NACC]; unsigned long long p[NACC]; int ia[NACC];
  for (int i = 0; i < NACC; ++i) { a[i] = tid * 0.01f + i; p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 0.5f); ia[i] = tid + i; }
  const float c1 = 1.0001f, c2 = 0.0001f;
  unsigned long long pc1 = ((unsigned long long)__float_as_uint(c1) << 32) | __float_as_uint(c1);
  unsigned long long pc2 = ((unsigned long long)__float_as_uint(c2) << 32) | __float_as_uint(c2);
  float4 v4 = make_float4(0, 0, 0, 0); float2 v2 = make_float2(0, 0); float v1 = 0;
  const float4* s4 = reinterpret_cast<const float4*>(sm) + (tid & 31);
  const float2* s2 = reinterpret_cast<const float2*>(sm) + (tid & 31);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        if (KIND == K_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(c2));
        if (KIND == K_FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pc1), "l"(pc2));
        if (KIND == K_FADD) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c2));
        if (KIND == K_FADD2) asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc2));
        if (KIND == K_FMUL2) asm volatile("mul.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc1));
        if (KIND == K_MIX2) { if (i & 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pc1), "l"(pc2)); else asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc2)); }
        if (KIND == K_MUFU_LG2) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        if (KIND == K_MUFU_SQRT) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        if (KIND == K_MUFU_RSQ) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        if (KIND == K_LDS32) { float t; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"((unsigned)__cvta_generic_to_shared(sm + ((tid & 31) + 32 * ((i + u * NACC) & 63))))); v1 += t; }
        if (KIND == K_LDS64) { float2 t; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(t.x), "=f"(t.y) : "r"((unsigned)__cvta_generic_to_shared(s2 + 32 * ((i + u * NACC) & 63)))); v2.x += t.x; v2.y += t.y; }
        if (KIND == K_LDS128) { float4 t; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"((unsigned)__cvta_generic_to_shared(s4 + 32 * ((i + u * NACC) & 31)))); v4.x += t.x; v4.w += t.w; }
        if (KIND == K_STS64) asm volatile("st.shared.v2.f32 [%0], {%1,%2};" :: "r"((unsigned)__cvta_generic_to_shared(s2 + 32 * ((i + u * NACC) & 63))), "f"(a[i]), "f"(a[(i + 1) % NACC]));
        if (KIND == K_STS128) asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"((unsigned)__cvta_generic_to_shared(s4 + 32 * ((i + u * NACC) & 31))), "f"(a[i]), "f"(a[(i + 1) % NACC]), "f"(a[(i + 2) % NACC]), "f"(a[(i + 3) % NACC]));
        if (KIND == K_SHFL) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 15));
        if (KIND == K_FFMA_SHFL) { if (i & 1) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 15)); else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(c2)); }
        if (KIND == K_FFMA2_LDS128) { if ((i & 3) == 3) { float4 t; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"((unsigned)__cvta_generic_to_shared(s4 + 32 * ((i + u * NACC) & 31)))); v4.x += t.x; } else asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pc1), "l"(pc2)); }
        if (KIND == K_IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(ia[i]) : "r"(tid));
        if (KIND == K_FFMA2_IADD) { if (i & 1) asm volatile("add.s32 %0, %0, %1;" : "+r"(ia[i]) : "r"(tid)); else asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pc1), "l"(pc2)); }
      }
    }
  }
  long long t1 = clock64();
  float s = v1 + v2.x + v2.y + v4.x + v4.w;
  for (int i = 0; i < NACC; ++i) s += a[i] + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]) + ia[i];
  out[blockIdx.x * blockDim.x + tid] = s;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(float* out, long long* cyc, int threads) {
  const int iters = 2000, grid = 148;
  cudaFuncSetAttribute(bench<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  bench<KIND><<<grid, threads, 32768>>>(out, cyc, 10);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<KIND><<<grid, threads, 32768>>>(out, cyc, iters);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < grid; ++i) c += h[i]; c /= grid;
  const double ops = (double)iters * UNROLL * NACC;             // warp-instructions per warp
  const int warps_per_smsp = threads / 32 / 4;
  printf("%-20s threads=%4d  cyc/instr/warp=%6.2f  cyc/instr/SMSP=%6.3f  ms=%.3f  eff_clk=%.0f MHz %s\n", names[KIND], threads,
         c / ops, c / ops / (warps_per_smsp > 0 ? warps_per_smsp : 1), ms, c / (ms * 1e3), err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int threads : {128, 256, 512, 1024}) {
    run<K_FFMA>(out, cyc, threads); run<K_FFMA2>(out, cyc, threads); run<K_FADD>(out, cyc, threads); run<K_FADD2>(out, cyc, threads);
    run<K_FMUL2>(out, cyc, threads); run<K_MIX2>(out, cyc, threads); run<K_MUFU_LG2>(out, cyc, threads); run<K_MUFU_SQRT>(out, cyc, threads);
    run<K_MUFU_RSQ>(out, cyc, threads); run<K_LDS32>(out, cyc, threads); run<K_LDS64>(out, cyc, threads); run<K_LDS128>(out, cyc, threads);
    run<K_STS64>(out, cyc, threads); run<K_STS128>(out, cyc, threads); run<K_SHFL>(out, cyc, threads); run<K_FFMA_SHFL>(out, cyc, threads);
    run<K_FFMA2_LDS128>(out, cyc, threads); run<K_IADD>(out, cyc, threads); run<K_FFMA2_IADD>(out, cyc, threads);
    printf("\n");
  }
  return 0;
}
