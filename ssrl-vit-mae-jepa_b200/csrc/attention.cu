// Fused small-sequence attention (timm Attention core = F.scaled_dot_product_attention, no mask, no dropout).
// One CTA per (sample, head); the whole Q/K/V tile of that head lives in shared memory; one warp owns a
// 16-row slab.  L <= 160 (encoder 36, decoder 145), head_dim 24 or 32 (24 is zero-padded to 32 for Q.K^T).
// Tensor work uses warp-level mma.sync m16n8k16 (bf16 in, fp32 accumulate): each (sample, head) problem is
// 36..145 rows, far below a 128-row tcgen05 tile, and S/P never leave registers.
//
// Backward recomputes S twice, once per orientation, so that every gradient is produced by the warp that
// owns its rows (no cross-warp reduction, no atomics, no P/dS round trip through shared memory):
//   phase A (warp = 16 queries): P, dP = dO.V^T, dS = P o (dP - delta), dQ = dS.K
//   phase B (warp = 16 keys):    P^T, dP^T = V.dO^T, dS^T,  dV = P^T.dO, dK = dS^T.Q
#include "mae_internal.h"
#include "ptx.cuh"

namespace mae {

static constexpr int ST = 40;  // smem row stride (bf16) of [row][d] tiles: 32 data + 8 pad -> conflict-free frags

__device__ __forceinline__ uint32_t ld32(const __nv_bfloat16* p) { return *reinterpret_cast<const uint32_t*>(p); }

// A fragment (16 rows x 16 k) of a row-major [row][k] smem tile
__device__ __forceinline__ void load_a_frag(uint32_t (&a)[4], const __nv_bfloat16* tile, int row0, int k0, int g, int t) {
  a[0] = ld32(tile + (row0 + g) * ST + k0 + t * 2);
  a[1] = ld32(tile + (row0 + g + 8) * ST + k0 + t * 2);
  a[2] = ld32(tile + (row0 + g) * ST + k0 + 8 + t * 2);
  a[3] = ld32(tile + (row0 + g + 8) * ST + k0 + 8 + t * 2);
}

template <int HD, int NT>
__global__ void __launch_bounds__(NT * 16) attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                           __nv_bfloat16* __restrict__ out, float* __restrict__ lse,
                                                           int L, int H, float scale) {
  constexpr int Lp = NT * 8, LpP = Lp + 8, CH = HD / 8, DN = HD / 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* sK = sQ + Lp * ST;
  __nv_bfloat16* sVt = sK + Lp * ST;  // [32][LpP]
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int D = H * HD;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;

  for (int i = tid; i < Lp * 4; i += NT * 16) {
    const int r = i >> 2, c = i & 3;
    uint4 q = make_uint4(0, 0, 0, 0), k = q, v = q;
    if (r < L && c < CH) {
      const __nv_bfloat16* base = qkv + (size_t(b) * L + r) * 3 * D + h * HD + c * 8;
      q = *reinterpret_cast<const uint4*>(base);
      k = *reinterpret_cast<const uint4*>(base + D);
      v = *reinterpret_cast<const uint4*>(base + 2 * D);
    }
    *reinterpret_cast<uint4*>(sQ + r * ST + c * 8) = q;
    *reinterpret_cast<uint4*>(sK + r * ST + c * 8) = k;
    const __nv_bfloat16* vv = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
    for (int e = 0; e < 8; ++e) sVt[(c * 8 + e) * LpP + r] = vv[e];
  }
  __syncthreads();

  const int q0 = w * 16;
  uint32_t aq[2][4];
  load_a_frag(aq[0], sQ, q0, 0, g, t);
  load_a_frag(aq[1], sQ, q0, 16, g, t);
  float s[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const __nv_bfloat16* kr = sK + (nt * 8 + g) * ST + ks * 16 + t * 2;
      mma_16816(s[nt], aq[ks], ld32(kr), ld32(kr + 8));
    }
  }
  const float sl2 = scale * 1.4426950408889634f;
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int col = nt * 8 + t * 2 + (i & 1);
      if (col >= L) s[nt][i] = -INFINITY;
    }
    m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
    m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    s[nt][0] = exp2f((s[nt][0] - m0) * sl2); s[nt][1] = exp2f((s[nt][1] - m0) * sl2);
    s[nt][2] = exp2f((s[nt][2] - m1) * sl2); s[nt][3] = exp2f((s[nt][3] - m1) * sl2);
    sum0 += s[nt][0] + s[nt][1];
    sum1 += s[nt][2] + s[nt][3];
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

  float o[DN][4];
#pragma unroll
  for (int dn = 0; dn < DN; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    uint32_t a[4];
    a[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
    a[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
    a[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    a[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
    for (int dn = 0; dn < DN; ++dn) {
      const __nv_bfloat16* vr = sVt + (dn * 8 + g) * LpP + kk * 16 + t * 2;
      mma_16816(o[dn], a, ld32(vr), ld32(vr + 8));
    }
  }
  const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
  const int r0 = q0 + g, r1 = q0 + g + 8;
#pragma unroll
  for (int dn = 0; dn < DN; ++dn) {
    const int d = h * HD + dn * 8 + t * 2;
    if (r0 < L) *reinterpret_cast<uint32_t*>(out + (size_t(b) * L + r0) * D + d) = pack_bf16x2(o[dn][0] * inv0, o[dn][1] * inv0);
    if (r1 < L) *reinterpret_cast<uint32_t*>(out + (size_t(b) * L + r1) * D + d) = pack_bf16x2(o[dn][2] * inv1, o[dn][3] * inv1);
  }
  if (t == 0) {
    if (r0 < L) lse[(size_t(b) * H + h) * L + r0] = m0 * scale + logf(sum0);
    if (r1 < L) lse[(size_t(b) * H + h) * L + r1] = m1 * scale + logf(sum1);
  }
}

template <int HD, int NT>
__global__ void __launch_bounds__(NT * 16) attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                           const __nv_bfloat16* __restrict__ out,
                                                           const __nv_bfloat16* __restrict__ dout,
                                                           const float* __restrict__ lse,
                                                           __nv_bfloat16* __restrict__ dqkv, int L, int H, float scale) {
  constexpr int Lp = NT * 8, LpP = Lp + 8, CH = HD / 8, DN = HD / 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* sK = sQ + Lp * ST;
  __nv_bfloat16* sV = sK + Lp * ST;
  __nv_bfloat16* sdO = sV + Lp * ST;
  __nv_bfloat16* sQt = sdO + Lp * ST;  // [32][LpP]
  __nv_bfloat16* sKt = sQt + 32 * LpP;
  __nv_bfloat16* sdOt = sKt + 32 * LpP;
  float* sLse = reinterpret_cast<float*>(sdOt + 32 * LpP);
  float* sDelta = sLse + Lp;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int D = H * HD;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;

  for (int i = tid; i < Lp * 4; i += NT * 16) {  // Lp*4 == 2 * blockDim: every lane runs both iterations
    const int r = i >> 2, c = i & 3;
    uint4 q = make_uint4(0, 0, 0, 0), k = q, v = q, d_o = q, oo = q;
    if (r < L && c < CH) {
      const __nv_bfloat16* base = qkv + (size_t(b) * L + r) * 3 * D + h * HD + c * 8;
      q = *reinterpret_cast<const uint4*>(base);
      k = *reinterpret_cast<const uint4*>(base + D);
      v = *reinterpret_cast<const uint4*>(base + 2 * D);
      const size_t ob = (size_t(b) * L + r) * D + h * HD + c * 8;
      d_o = *reinterpret_cast<const uint4*>(dout + ob);
      oo = *reinterpret_cast<const uint4*>(out + ob);
    }
    *reinterpret_cast<uint4*>(sQ + r * ST + c * 8) = q;
    *reinterpret_cast<uint4*>(sK + r * ST + c * 8) = k;
    *reinterpret_cast<uint4*>(sV + r * ST + c * 8) = v;
    *reinterpret_cast<uint4*>(sdO + r * ST + c * 8) = d_o;
    const __nv_bfloat16* qq = reinterpret_cast<const __nv_bfloat16*>(&q);
    const __nv_bfloat16* kk2 = reinterpret_cast<const __nv_bfloat16*>(&k);
    const __nv_bfloat16* dd = reinterpret_cast<const __nv_bfloat16*>(&d_o);
    const __nv_bfloat16* op = reinterpret_cast<const __nv_bfloat16*>(&oo);
    float part = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sQt[(c * 8 + e) * LpP + r] = qq[e];
      sKt[(c * 8 + e) * LpP + r] = kk2[e];
      sdOt[(c * 8 + e) * LpP + r] = dd[e];
      part += __bfloat162float(dd[e]) * __bfloat162float(op[e]);
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    if (c == 0) {
      sDelta[r] = part;
      sLse[r] = (r < L) ? lse[(size_t(b) * H + h) * L + r] : 0.f;
    }
  }
  __syncthreads();

  const float sl2 = scale * 1.4426950408889634f;
  const float l2e = 1.4426950408889634f;
  const int x0 = w * 16;  // first row of this warp's slab (queries in phase A, keys in phase B)
  const int r0 = x0 + g, r1 = x0 + g + 8;

  // ------------------------------ phase A: dQ ------------------------------
  {
    uint32_t aq[2][4], ado[2][4];
    load_a_frag(aq[0], sQ, x0, 0, g, t);  load_a_frag(aq[1], sQ, x0, 16, g, t);
    load_a_frag(ado[0], sdO, x0, 0, g, t); load_a_frag(ado[1], sdO, x0, 16, g, t);
    float s[NT][4];
    const float l0 = sLse[r0] * l2e, l1 = sLse[r1] * l2e;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const __nv_bfloat16* kr = sK + (nt * 8 + g) * ST + ks * 16 + t * 2;
        mma_16816(s[nt], aq[ks], ld32(kr), ld32(kr + 8));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int col = nt * 8 + t * 2 + (i & 1);
        s[nt][i] = (col < L) ? exp2f(s[nt][i] * sl2 - (i < 2 ? l0 : l1)) : 0.f;
      }
    }
    const float d0 = sDelta[r0], d1 = sDelta[r1];
    float dq[DN][4];
#pragma unroll
    for (int dn = 0; dn < DN; ++dn) dq[dn][0] = dq[dn][1] = dq[dn][2] = dq[dn][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
      float dp[2][4];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        dp[hf][0] = dp[hf][1] = dp[hf][2] = dp[hf][3] = 0.f;
        const int nt = 2 * kk + hf;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const __nv_bfloat16* vr = sV + (nt * 8 + g) * ST + ks * 16 + t * 2;
          mma_16816(dp[hf], ado[ks], ld32(vr), ld32(vr + 8));
        }
        dp[hf][0] = s[nt][0] * (dp[hf][0] - d0); dp[hf][1] = s[nt][1] * (dp[hf][1] - d0);
        dp[hf][2] = s[nt][2] * (dp[hf][2] - d1); dp[hf][3] = s[nt][3] * (dp[hf][3] - d1);
      }
      uint32_t a[4];
      a[0] = pack_bf16x2(dp[0][0], dp[0][1]); a[1] = pack_bf16x2(dp[0][2], dp[0][3]);
      a[2] = pack_bf16x2(dp[1][0], dp[1][1]); a[3] = pack_bf16x2(dp[1][2], dp[1][3]);
#pragma unroll
      for (int dn = 0; dn < DN; ++dn) {
        const __nv_bfloat16* kr = sKt + (dn * 8 + g) * LpP + kk * 16 + t * 2;
        mma_16816(dq[dn], a, ld32(kr), ld32(kr + 8));
      }
    }
#pragma unroll
    for (int dn = 0; dn < DN; ++dn) {
      const int d = h * HD + dn * 8 + t * 2;
      if (r0 < L) *reinterpret_cast<uint32_t*>(dqkv + (size_t(b) * L + r0) * 3 * D + d) = pack_bf16x2(dq[dn][0] * scale, dq[dn][1] * scale);
      if (r1 < L) *reinterpret_cast<uint32_t*>(dqkv + (size_t(b) * L + r1) * 3 * D + d) = pack_bf16x2(dq[dn][2] * scale, dq[dn][3] * scale);
    }
  }

  // ------------------------------ phase B: dK, dV ------------------------------
  {
    uint32_t ak[2][4], av[2][4];
    load_a_frag(ak[0], sK, x0, 0, g, t); load_a_frag(ak[1], sK, x0, 16, g, t);
    load_a_frag(av[0], sV, x0, 0, g, t); load_a_frag(av[1], sV, x0, 16, g, t);
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const __nv_bfloat16* qr = sQ + (nt * 8 + g) * ST + ks * 16 + t * 2;
        mma_16816(s[nt], ak[ks], ld32(qr), ld32(qr + 8));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int qc = nt * 8 + t * 2 + (i & 1);
        const int kr = (i < 2) ? r0 : r1;
        s[nt][i] = (qc < L && kr < L) ? exp2f(s[nt][i] * sl2 - sLse[qc] * l2e) : 0.f;
      }
    }
    float dk[DN][4], dv[DN][4];
#pragma unroll
    for (int dn = 0; dn < DN; ++dn) {
      dk[dn][0] = dk[dn][1] = dk[dn][2] = dk[dn][3] = 0.f;
      dv[dn][0] = dv[dn][1] = dv[dn][2] = dv[dn][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
      float ds[2][4];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        ds[hf][0] = ds[hf][1] = ds[hf][2] = ds[hf][3] = 0.f;
        const int nt = 2 * kk + hf;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const __nv_bfloat16* dr = sdO + (nt * 8 + g) * ST + ks * 16 + t * 2;
          mma_16816(ds[hf], av[ks], ld32(dr), ld32(dr + 8));
        }
        const float dl0 = sDelta[nt * 8 + t * 2], dl1 = sDelta[nt * 8 + t * 2 + 1];
        ds[hf][0] = s[nt][0] * (ds[hf][0] - dl0); ds[hf][1] = s[nt][1] * (ds[hf][1] - dl1);
        ds[hf][2] = s[nt][2] * (ds[hf][2] - dl0); ds[hf][3] = s[nt][3] * (ds[hf][3] - dl1);
      }
      uint32_t ap[4], as_[4];
      ap[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);         ap[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      ap[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]); ap[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      as_[0] = pack_bf16x2(ds[0][0], ds[0][1]); as_[1] = pack_bf16x2(ds[0][2], ds[0][3]);
      as_[2] = pack_bf16x2(ds[1][0], ds[1][1]); as_[3] = pack_bf16x2(ds[1][2], ds[1][3]);
#pragma unroll
      for (int dn = 0; dn < DN; ++dn) {
        const __nv_bfloat16* dor = sdOt + (dn * 8 + g) * LpP + kk * 16 + t * 2;
        mma_16816(dv[dn], ap, ld32(dor), ld32(dor + 8));
        const __nv_bfloat16* qr = sQt + (dn * 8 + g) * LpP + kk * 16 + t * 2;
        mma_16816(dk[dn], as_, ld32(qr), ld32(qr + 8));
      }
    }
#pragma unroll
    for (int dn = 0; dn < DN; ++dn) {
      const int d = h * HD + dn * 8 + t * 2;
      if (r0 < L) {
        __nv_bfloat16* base = dqkv + (size_t(b) * L + r0) * 3 * D + d;
        *reinterpret_cast<uint32_t*>(base + D) = pack_bf16x2(dk[dn][0] * scale, dk[dn][1] * scale);
        *reinterpret_cast<uint32_t*>(base + 2 * D) = pack_bf16x2(dv[dn][0], dv[dn][1]);
      }
      if (r1 < L) {
        __nv_bfloat16* base = dqkv + (size_t(b) * L + r1) * 3 * D + d;
        *reinterpret_cast<uint32_t*>(base + D) = pack_bf16x2(dk[dn][2] * scale, dk[dn][3] * scale);
        *reinterpret_cast<uint32_t*>(base + 2 * D) = pack_bf16x2(dv[dn][2], dv[dn][3]);
      }
    }
  }
}

template <int HD, int NT>
static int launch_fwd(const void* qkv, int B, int L, int H, void* out, float* lse, cudaStream_t st) {
  constexpr int Lp = NT * 8, LpP = Lp + 8;
  const size_t smem = size_t(2 * Lp * ST + 32 * LpP) * 2;
  static bool done = false;
  if (!done) {
    if (cudaFuncSetAttribute(attn_fwd_kernel<HD, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
      return MAE_ERR_LAUNCH;
    done = true;
  }
  attn_fwd_kernel<HD, NT><<<B * H, NT * 16, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                        reinterpret_cast<__nv_bfloat16*>(out), lse, L, H,
                                                        1.0f / sqrtf(float(HD)));
  return MAE_CHECK_LAUNCH();
}

template <int HD, int NT>
static int launch_bwd(const void* qkv, const void* out, const void* dout, const float* lse, int B, int L, int H,
                      void* dqkv, cudaStream_t st) {
  constexpr int Lp = NT * 8, LpP = Lp + 8;
  const size_t smem = size_t(4 * Lp * ST + 3 * 32 * LpP) * 2 + size_t(2 * Lp) * 4;
  static bool done = false;
  if (!done) {
    if (cudaFuncSetAttribute(attn_bwd_kernel<HD, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
      return MAE_ERR_LAUNCH;
    done = true;
  }
  attn_bwd_kernel<HD, NT><<<B * H, NT * 16, smem, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(out),
      reinterpret_cast<const __nv_bfloat16*>(dout), lse, reinterpret_cast<__nv_bfloat16*>(dqkv), L, H,
      1.0f / sqrtf(float(HD)));
  return MAE_CHECK_LAUNCH();
}

static int pick_nt(int L) {
  const int need = 2 * ((L + 15) / 16);
  const int opts[4] = {4, 6, 10, 20};
  for (int i = 0; i < 4; ++i)
    if (opts[i] >= need) return opts[i];
  return 0;
}

}  // namespace mae

using namespace mae;

#define MAE_ATTN_DISPATCH(FN, ...)                                           \
  switch (hd * 100 + nt) {                                                   \
    case 2404: return FN<24, 4>(__VA_ARGS__);                                \
    case 2406: return FN<24, 6>(__VA_ARGS__);                                \
    case 2410: return FN<24, 10>(__VA_ARGS__);                               \
    case 2420: return FN<24, 20>(__VA_ARGS__);                               \
    case 3204: return FN<32, 4>(__VA_ARGS__);                                \
    case 3206: return FN<32, 6>(__VA_ARGS__);                                \
    case 3210: return FN<32, 10>(__VA_ARGS__);                               \
    case 3220: return FN<32, 20>(__VA_ARGS__);                               \
    default: return MAE_ERR_SHAPE;                                           \
  }

extern "C" int mae_attention_fwd(const void* qkv, int B, int L, int H, int hd, void* out, float* lse,
                                 cudaStream_t stream) {
  if (B <= 0) return MAE_OK;
  if (!qkv || !out || !lse) return MAE_ERR_NULL;
  const int nt = pick_nt(L);
  if (!nt || L < 1 || H < 1) return MAE_ERR_SHAPE;
  MAE_ATTN_DISPATCH(launch_fwd, qkv, B, L, H, out, lse, stream)
}

extern "C" int mae_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, int B, int L,
                                 int H, int hd, void* dqkv, cudaStream_t stream) {
  if (B <= 0) return MAE_OK;
  if (!qkv || !out || !dout || !lse || !dqkv) return MAE_ERR_NULL;
  const int nt = pick_nt(L);
  if (!nt || L < 1 || H < 1) return MAE_ERR_SHAPE;
  MAE_ATTN_DISPATCH(launch_bwd, qkv, out, dout, lse, B, L, H, dqkv, stream)
}
