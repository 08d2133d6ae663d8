// Minimal producer / consumer hand-off through two mbarriers (full / empty), written the same way as feat3.cuh
// (inline PTX: mbarrier.init by one thread + fence.mbarrier_init + __syncthreads, 32-lane arrive, try_wait.parity loop).
// Used to find out what compute-sanitizer's synccheck / racecheck report for the bare pattern:
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o mbar_minimal mbar_minimal.cu
//   compute-sanitizer --tool synccheck ./mbar_minimal ; compute-sanitizer --tool racecheck ./mbar_minimal
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned addr, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned addr) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned addr, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(addr), "r"(parity)
      : "memory");
}

template <bool DYNAMIC>
__global__ void handoff(float* out, int rounds) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ __align__(16) unsigned char stat[32 * 4 + 16];
  unsigned char* raw = DYNAMIC ? dyn : stat;
  float* buf = reinterpret_cast<float*>(raw);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(raw + 32 * 4);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(bar), 32);
    mbar_init(smem_u32(bar + 1), 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned full = smem_u32(bar), empty = smem_u32(bar + 1);
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int r = 0; r < rounds; ++r) {
    if (threadIdx.x < 32) {
      mbar_wait(empty, (r & 1) ^ 1);
      buf[lane] = static_cast<float>(r * 32 + lane);
      mbar_arrive(full);
    } else {
      mbar_wait(full, r & 1);
      acc += buf[31 - lane];
      mbar_arrive(empty);
    }
  }
  if (threadIdx.x >= 32) out[lane] = acc;
}

// The same hand-off for PAIRS producer / consumer warp pairs in one CTA, laid out like feat3.cuh: producers are warps 0 .. PAIRS-1,
// consumers warps PAIRS .. 2 PAIRS-1, barriers full[PAIRS] then empty[PAIRS] back to back (8 bytes apart), one lane per barrier
// initialises it.
// delay: cycles the producer spends on an item before it arrives (feat3's analysis warp needs ~15 k cycles per item, so its
// consumer sits in try_wait across many suspend time-outs; with delay = 0 the waits are over at once).
// hi: byte offset of the buffers and barriers inside the dynamic shared memory (feat3 keeps them at 132 .. 224 KB).
template <int PAIRS>
__global__ void handoff_pairs(float* out, int rounds, long long delay, int hi) {
  extern __shared__ __align__(16) unsigned char dyn_base[];
  unsigned char* dyn = dyn_base + hi;
  float* bufs = reinterpret_cast<float*>(dyn);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(dyn + PAIRS * 32 * 4);
  if (threadIdx.x < 2 * PAIRS) {
    mbar_init(smem_u32(bar + threadIdx.x), 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, w = warp % PAIRS;
  const unsigned full = smem_u32(bar + w), empty = smem_u32(bar + PAIRS + w);
  float* buf = bufs + 32 * w;
  float acc = 0.f;
  for (int r = 0; r < rounds; ++r) {
    if (warp < PAIRS) {
      const long long t0 = clock64();
      while (clock64() - t0 < delay) {}
      mbar_wait(empty, (r & 1) ^ 1);
      buf[lane] = static_cast<float>(r * 32 + lane);
      mbar_arrive(full);
    } else {
      mbar_wait(full, r & 1);
      acc += buf[31 - lane];
      mbar_arrive(empty);
    }
  }
  if (warp >= PAIRS) out[32 * w + lane] = acc;
}

// The same hand-off with ONE named barrier per warp pair (bar.sync id, 64), hit twice per round by both warps from their own
// branches -- the producer / consumer use of named barriers (as CUTLASS's NamedBarrier), and feat3.cuh's default hand-off.
template <int PAIRS>
__global__ void handoff_named(float* out, int rounds) {
  extern __shared__ __align__(16) unsigned char dyn[];
  float* bufs = reinterpret_cast<float*>(dyn);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, w = warp % PAIRS;
  float* buf = bufs + 32 * w;
  float acc = 0.f;
  for (int r = 0; r < rounds; ++r) {
    if (warp < PAIRS) {
      asm volatile("bar.sync %0, 64;" ::"r"(w + 1) : "memory");   // empty
      buf[lane] = static_cast<float>(r * 32 + lane);
      asm volatile("bar.sync %0, 64;" ::"r"(w + 1) : "memory");   // full
    } else {
      asm volatile("bar.sync %0, 64;" ::"r"(w + 1) : "memory");   // empty
      asm volatile("bar.sync %0, 64;" ::"r"(w + 1) : "memory");   // full
      acc += buf[31 - lane];
    }
  }
  if (warp >= PAIRS) out[32 * w + lane] = acc;
}

int main() {
  float* out;
  cudaMalloc(&out, 32 * sizeof(float));
  handoff<true><<<1, 64, 32 * 4 + 16>>>(out, 8);
  cudaError_t e1 = cudaDeviceSynchronize();
  handoff<false><<<1, 64>>>(out, 8);
  cudaError_t e2 = cudaDeviceSynchronize();
  float h[32];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("dynamic smem: %s, static smem: %s, out[0] = %g (expect %g)\n", cudaGetErrorString(e1), cudaGetErrorString(e2), h[0],
         8 * 31.f + 32.f * 28);
  float* out8;
  cudaMalloc(&out8, 8 * 32 * sizeof(float));
  cudaFuncSetAttribute(handoff_pairs<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const int his[] = {0, 40 * 1024, 100 * 1024, 220 * 1024};
  for (int hi : his)
    for (long long delay : {0LL, 20000LL}) {
      handoff_pairs<8><<<1, 512, hi + 8 * 32 * 4 + 16 * 8>>>(out8, 8, delay, hi);
      cudaError_t e3 = cudaDeviceSynchronize();
      float h8[256];
      cudaMemcpy(h8, out8, sizeof(h8), cudaMemcpyDeviceToHost);
      printf("8 warp pairs, 16 barriers at +%d KB, producer delay %lld cycles: %s, out[0] = %g, out[255] = %g (expect %g, %g)\n",
             hi / 1024, delay, cudaGetErrorString(e3), h8[0], h8[255], 8 * 31.f + 32.f * 28, 8 * 0.f + 32.f * 28);
    }
  handoff_named<8><<<1, 512, 8 * 32 * 4>>>(out8, 8);
  cudaError_t e4 = cudaDeviceSynchronize();
  float h8[256];
  cudaMemcpy(h8, out8, sizeof(h8), cudaMemcpyDeviceToHost);
  printf("8 warp pairs, named barriers (bar.sync id, 64 from both branches): %s, out[0] = %g, out[255] = %g (expect %g, %g)\n",
         cudaGetErrorString(e4), h8[0], h8[255], 8 * 31.f + 32.f * 28, 8 * 0.f + 32.f * 28);
  return 0;
}
