// Internal declarations shared by the .cu translation units of libmae_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mae_b200.h"

namespace mae {
int make_tmap_bf16(CUtensorMap* map, const void* ptr, int rows, int cols, int ld, int box_rows);
// attention_tc.cu: tcgen05 attention for 128 < L <= 160, head_dim 32, even H
// q_rows: number of leading query rows whose outputs (gradients) are wanted; <= 128 skips the second query tile
int attn_fwd_tc(const void* qkv, int B, int L, int H, void* out, float* lse, int q_rows, cudaStream_t st);
// delta_ws: optional (B, L, H) f32 scratch: rowsum(dO o O) is then computed by a coalesced pre-pass instead of per-thread gathers
int attn_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, int B, int L, int H, void* dqkv,
                float* delta_ws, int q_rows, cudaStream_t st);
// attention_enc.cu: tcgen05 attention for short rows (L <= 64, head_dim 8..32): several samples per 128-row tile
bool attn_enc_covers(int L, int hd);
int attn_fwd_enc(const void* qkv, int B, int L, int H, int hd, void* out, float* lse, cudaStream_t st);
int attn_bwd_enc(const void* qkv, const void* out, const void* dout, const float* lse, int B, int L, int H, int hd,
                 void* dqkv, cudaStream_t st);
}

// Launch with the programmatic-dependent-launch attribute (see ptx.cuh: pdl_trigger / pdl_wait).
template <typename... KArgs, typename... Args>
inline cudaError_t mae_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Per-device state.  A process may drive several GPUs (one engine per device): the SM count and the opt-in for more than
// 48 KB of dynamic shared memory (cudaFuncSetAttribute is per device) are cached per CUDA device, not per process.
namespace mae {
constexpr int kMaxDevices = 64;
inline int cur_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
  return d;
}
inline int num_sms() {
  static int n[kMaxDevices] = {};
  const int d = cur_device();
  if (!n[d]) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || v <= 0) v = 148;
    n[d] = v;
  }
  return n[d];
}
// `done` is a function-local `static bool[kMaxDevices]` of the caller: one flag per (kernel instantiation, device).
template <typename K>
inline bool ensure_dyn_smem(K kernel, size_t bytes, bool (&done)[kMaxDevices]) {
  const int d = cur_device();
  if (done[d]) return true;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)) != cudaSuccess) return false;
  done[d] = true;
  return true;
}
}  // namespace mae

#define MAE_CHECK_LAUNCH() (cudaGetLastError() == cudaSuccess ? MAE_OK : MAE_ERR_LAUNCH)
