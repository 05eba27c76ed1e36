// Issue cost of tcgen05.mma (cta_group::1, kind::f16, bf16) by operand layout and shape: one lane issues `reps` MMAs
// back to back into one accumulator, commits, waits; cycles per MMA = (t1 - t0) / reps.
#include <cstdio>
#include <cuda_bf16.h>
#include "ptx.cuh"
using namespace mae;

__device__ __forceinline__ uint64_t sdesc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return uint64_t((addr & 0x3ffffu) >> 4) | (uint64_t((lbo >> 4) & 0x3fffu) << 16) | (uint64_t((sbo >> 4) & 0x3fffu) << 32) |
         (uint64_t(1) << 46) | (uint64_t(layout) << 61);
}

struct Cfg { int M, N, a_mn, b_mn, a_sw64, b_sw64; };

__global__ void __launch_bounds__(128) probe(Cfg c, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tbase, 256);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 64 * 1024);
    const uint32_t idesc = umma_idesc_bf16(c.M, c.N, c.a_mn, c.b_mn);
    // K-major: SBO = 8 rows; MN-major: LBO = next 64/32-element group (16 KB box / unused), SBO = 8 K-rows
    const uint64_t ad = c.a_mn ? sdesc(a, 16384, c.a_sw64 ? 512 : 1024, c.a_sw64 ? 4 : 2) : sdesc(a, 16, c.a_sw64 ? 512 : 1024, c.a_sw64 ? 4 : 2);
    const uint64_t bd = c.b_mn ? sdesc(b, 16384, c.b_sw64 ? 512 : 1024, c.b_sw64 ? 4 : 2) : sdesc(b, 16, c.b_sw64 ? 512 : 1024, c.b_sw64 ? 4 : 2);
    for (int w = 0; w < 2; ++w) {  // second round is the measurement
      const long long t0 = clock64();
      for (int i = 0; i < reps; ++i) umma_bf16(tbase, ad, bd, idesc, 1);
      umma_commit(&bar);
      mbar_wait(&bar, w);
      out[w] = clock64() - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tbase, 256);
}

int main() {
  long long* d;
  cudaMalloc(&d, 2 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const Cfg cfgs[] = {{128, 256, 0, 0, 0, 0}, {128, 192, 0, 0, 0, 0}, {128, 160, 0, 0, 0, 0}, {128, 160, 0, 0, 1, 1}, {128, 32, 0, 0, 0, 0},
                      {128, 32, 0, 1, 0, 0}, {128, 32, 0, 1, 0, 1}, {128, 32, 1, 1, 0, 1}, {128, 32, 1, 1, 0, 0}, {128, 64, 1, 1, 0, 0},
                      {64, 32, 0, 1, 0, 1}, {64, 32, 1, 1, 0, 1}, {64, 160, 0, 0, 1, 1}, {64, 160, 1, 0, 0, 0}, {128, 16, 0, 0, 0, 0}, {128, 8, 0, 0, 0, 0}};
  const int reps = 512;
  for (const Cfg& c : cfgs) {
    probe<<<1, 128, 100 * 1024>>>(c, reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("M=%3d N=%3d A %s/%s  B %s/%s : %6.1f cycles per MMA  (%s)\n", c.M, c.N, c.a_mn ? "MN" : "K ", c.a_sw64 ? "sw64 " : "sw128",
           c.b_mn ? "MN" : "K ", c.b_sw64 ? "sw64 " : "sw128", double(h[1]) / reps, cudaGetErrorString(e));
  }
  return 0;
}
