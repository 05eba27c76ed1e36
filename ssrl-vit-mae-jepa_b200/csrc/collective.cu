// Gradient exchange of the data-parallel step over NVLink peer memory (sm_100a, one process per GPU):
//
//   out[i] = (1 / W) * sum_{r < W} grads_r[i]        and        partials[block] = sum over the block's i of out[i]^2
//
// One-shot all-reduce fused with the first pass of the global-norm clip.  Every rank's flat gradient buffer lives in
// symmetric memory (torch.distributed._symmetric_memory: the same allocation mapped into every process of the node), so a
// kernel on rank k simply LOADS the other ranks' gradients through NVLink / NVSwitch (`peers[r]` are peer-mapped device
// pointers) -- no staging copies, no ring steps, no separate library kernel competing for SMs with the persistent GEMMs:
// the CTAs are 256 threads x <= 40 registers and co-reside with them.  All ranks add the W buffers in the same order
// (r = 0 .. W-1), so every rank computes bit-identical averages and the replicas never drift.  The 7.9 MB gradient is cut
// into one bucket per transformer block; bucket i's launch follows a cross-rank barrier that every rank enters once its
// own backward has produced bucket i, so all but the last bucket overlap the rest of backward.
//
// Replaces ncclAllReduce(avg) + mae_grad_sumsq in FusedTrainer (reference: Lightning's DDP strategy would do the
// all-reduce; the reference itself is devices=1, scripts/training/pretrain_mae.py:116-118).
#include "mae_internal.h"
#include "ptx.cuh"

namespace mae {

static constexpr int kArMaxWorld = 16;

__global__ void __launch_bounds__(256, 6) allreduce_avg_sumsq_kernel(const float* const* __restrict__ peers, int world,
                                                                    long long off, long long n4, float inv_world,
                                                                    float* __restrict__ out, float* __restrict__ partials) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s[8];
  __shared__ const float4* src[kArMaxWorld];
  if (threadIdx.x < kArMaxWorld)
    src[threadIdx.x] = threadIdx.x < world ? reinterpret_cast<const float4*>(peers[threadIdx.x] + off) : nullptr;
  __syncthreads();
  float4* dst = reinterpret_cast<float4*>(out + off);
  float t = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r0 = 0; r0 < world; r0 += 4) {  // four peer loads in flight at a time (<= 40 registers: the CTAs co-reside
      float4 v[4];                           // with the persistent GEMM CTAs); fixed order r = 0..W-1 on every rank
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = (r0 + k < world) ? src[r0 + k][i] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) { a.x += v[k].x; a.y += v[k].y; a.z += v[k].z; a.w += v[k].w; }
    }
    a.x *= inv_world; a.y *= inv_world; a.z *= inv_world; a.w *= inv_world;
    dst[i] = a;
    t += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
  }
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int i = 0; i < 8; ++i) a += s[i];
    partials[blockIdx.x] = a;
  }
}

}  // namespace mae

using namespace mae;

extern "C" int mae_allreduce_avg_sumsq(const void* peer_ptrs_dev, int world, long long offset, long long n, float* out,
                                       float* partials, int nblocks, cudaStream_t stream) {
  if (n <= 0) return MAE_OK;
  if (!peer_ptrs_dev || !out || !partials) return MAE_ERR_NULL;
  if (world < 1 || world > kArMaxWorld || (n & 3) || (offset & 3) || nblocks < 1) return MAE_ERR_SHAPE;
  mae_launch(allreduce_avg_sumsq_kernel, dim3(nblocks), dim3(256), 0, stream,
             reinterpret_cast<const float* const*>(peer_ptrs_dev), world, offset, n >> 2, 1.0f / float(world), out, partials);
  return MAE_CHECK_LAUNCH();
}
