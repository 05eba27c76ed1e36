// Where the ~47 cycles per small tcgen05.mma go: (a) the issuing thread's own time per MMA (clock after the issue loop, before
// the commit), (b) commit + wait cost, (c) two warps issuing concurrently into different accumulators (tensor-pipe floor or
// per-thread issue cost?), (d) the cost of tcgen05.commit and of an mbarrier wait that is already complete.
#include <cstdio>
#include <cuda_bf16.h>
#include "ptx.cuh"
using namespace mae;

__device__ __forceinline__ uint64_t sdesc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return uint64_t((addr & 0x3ffffu) >> 4) | (uint64_t((lbo >> 4) & 0x3fffu) << 16) | (uint64_t((sbo >> 4) & 0x3fffu) << 32) |
         (uint64_t(1) << 46) | (uint64_t(layout) << 61);
}

__global__ void __launch_bounds__(128) probe(int N, int reps, int nissuers, int same_acc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[4], bar2[64];
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); for (int i = 0; i < 64; ++i) mbar_init(&bar2[i], 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tbase, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < nissuers) {
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 64 * 1024);
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const uint64_t ad = sdesc(a, 16, 1024, 2), bd = sdesc(b, 16, 1024, 2);
    const uint32_t d = tbase + (same_acc ? 0 : w * 128);
    for (int r = 0; r < 2; ++r) {
      const long long t0 = clock64();
      for (int i = 0; i < reps; ++i) umma_bf16(d, ad, bd, idesc, 1);
      const long long t1 = clock64();
      umma_commit(&bar[w]);
      const long long t2 = clock64();
      mbar_wait(&bar[w], r);
      const long long t3 = clock64();
      if (r == 1) { out[w * 8 + 0] = t1 - t0; out[w * 8 + 1] = t2 - t1; out[w * 8 + 2] = t3 - t2; }
    }
    if (w == 0) {
      // 32 commits back to back (nothing outstanding), then 32 waits on completed barriers
      long long t0 = clock64();
      for (int i = 0; i < 32; ++i) umma_commit(&bar2[i]);
      long long t1 = clock64();
      for (int i = 0; i < 32; ++i) mbar_wait(&bar2[i], 0);
      long long t2 = clock64();
      for (int i = 0; i < 32; ++i) mbar_wait(&bar2[i], 0);   // certainly complete now
      long long t3 = clock64();
      out[3] = t1 - t0; out[4] = t2 - t1; out[5] = t3 - t2;
      // groups of 8 MMAs + commit + wait (the attention pattern): latency of a short dependent burst
      t0 = clock64();
      for (int g = 0; g < 16; ++g) {
        for (int i = 0; i < 8; ++i) umma_bf16(d, ad, bd, idesc, 1);
        umma_commit(&bar2[32 + g]);
        mbar_wait(&bar2[32 + g], 0);
      }
      out[6] = (clock64() - t0) / 16;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tbase, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 256;
  for (int N : {32, 128})
    for (int nis : {1, 2, 4})
      for (int same : {0, 1}) {
        if (nis == 1 && same) continue;
        cudaMemset(d, 0, 64 * sizeof(long long));
        probe<<<1, 128, 100 * 1024>>>(N, reps, nis, same, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[64];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("N=%3d issuers=%d same_acc=%d (%s):", N, nis, same, cudaGetErrorString(e));
        for (int w = 0; w < nis; ++w) printf("  w%d issue %.1f/MMA commit %lld wait %lld total %.1f/MMA;", w, double(h[w * 8]) / reps, h[w * 8 + 1], h[w * 8 + 2], double(h[w * 8] + h[w * 8 + 1] + h[w * 8 + 2]) / reps);
        printf("\n   32 commits: %.1f each; 32 waits (first) %.1f each; 32 waits (complete) %.1f each; burst of 8 MMAs+commit+wait: %lld cycles\n",
               h[3] / 32.0, h[4] / 32.0, h[5] / 32.0, h[6]);
      }
  return 0;
}
