// tcgen05.mma issue cost, two ways of electing the issuing thread:
//   (a) `if (threadIdx.x == 0)`: the compiler cannot prove the descriptor registers warp-uniform inside a thread-dependent
//       branch, so every UTCHMMA is wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall plus four R2UR
//   (b) warp-uniform branch (`__shfl_sync`-derived warp index) + elect.sync: descriptors live in uniform registers, the
//       UTCHMMAs are issued back to back
#include <cstdio>
#include <cuda_bf16.h>
#include "ptx.cuh"
using namespace mae;
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
template <int MODE>
__global__ void __launch_bounds__(128) probe(int N, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tbase, 256);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t a = smem_u32(smem), b = smem_u32(smem + 64 * 1024);
  const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
  const uint64_t ad = umma_sdesc_sw128(a, 16, 1024), bd = umma_sdesc_sw128(b, 16, 1024);
  const uint32_t d = tbase;
  if (MODE == 0) {
    if (threadIdx.x == 0) {
      for (int w = 0; w < 2; ++w) {
        const long long t0 = clock64();
        for (int i = 0; i < reps; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) umma_bf16(d, ad + 2 * (j & 3), bd + 2 * (j & 3), idesc, 1);
        const long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, w);
        out[2 * w] = t1 - t0; out[2 * w + 1] = clock64() - t0;
      }
    }
  } else {
    const int warp_u = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    if (warp_u == 0) {
      for (int w = 0; w < 2; ++w) {
        const long long t0 = clock64();
        long long t1 = 0;
        if (elect_one()) {
          for (int i = 0; i < reps; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) umma_bf16(d, ad + 2 * (j & 3), bd + 2 * (j & 3), idesc, 1);
          t1 = clock64();
          umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, w);
        if (t1) { out[2 * w] = t1 - t0; out[2 * w + 1] = clock64() - t0; }
      }
    }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tbase, 256);
}
int main() {
  long long* d; cudaMalloc(&d, 4 * sizeof(long long));
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 64;
  for (int N : {32, 64, 128, 144, 192, 256})
    for (int mode = 0; mode < 2; ++mode) {
      if (mode == 0) probe<0><<<1, 128, 100 * 1024>>>(N, reps, d); else probe<1><<<1, 128, 100 * 1024>>>(N, reps, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[4]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      printf("N=%3d %-26s issue %.1f cycles/MMA, issue+complete %.1f cycles/MMA (%s)\n", N, mode ? "uniform branch + elect.sync" : "if (threadIdx.x == 0)",
             double(h[2]) / (reps * 8), double(h[3]) / (reps * 8), cudaGetErrorString(e));
    }
  return 0;
}
