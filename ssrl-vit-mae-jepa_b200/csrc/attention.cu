// Fused small-sequence attention (timm Attention core = F.scaled_dot_product_attention, no mask, no dropout).
// One CTA per (sample, head); the whole Q/K/V tile of that head lives in shared memory (row-major, padded
// rows); one warp owns a 16-row slab.  L <= 160 (encoder 36, decoder 145), head_dim 24 or 32 (24 is zero-padded
// to 32).  Tensor work uses warp-level mma.sync m16n8k16 (bf16 in, fp32 accumulate) fed by ldmatrix: each
// (sample, head) problem is 36..145 rows, far below a 128-row tcgen05 tile, and S/P never leave registers.
//
// Register pressure is kept low so that several CTAs are resident per SM (the load phase of one overlaps the
// math of another): the forward keeps the scores of short rows (<= 64 keys) in registers and otherwise makes two passes
// over the keys (pass 1: row max, pass 2: recompute
// S per 16-key block, P = exp(S - max), O += P.V); the backward uses the saved log-sum-exp, so every 16-wide
// block of P is recomputed independently.  Backward recomputes S once per orientation so that every gradient
// is produced by the warp that owns its rows (no cross-warp reduction, no atomics, no smem round trip for P/dS):
//   phase A (warp = 16 queries): P, dP = dO.V^T, dS = P o (dP - delta), dQ = dS.K
//   phase B (warp = 16 keys):    P^T, dP^T = V.dO^T, dS^T,  dV = P^T.dO, dK = dS^T.Q
#include <stdlib.h>

#include "mae_internal.h"
#include "ptx.cuh"

namespace mae {

// smem row stride (bf16) = padded head dim + 8: 80-byte (hd 24/32) or 144-byte (hd 64) rows, conflict-free ldmatrix

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}

// All fragment loads take a 32-bit shared address that already contains the lane's own offset, so the loops only
// add small compile-time-scaled offsets (no 64-bit pointer arithmetic in the hot loops):
//   "a-type" lane offset  ((lane&7) + ((lane>>3)&1)*8) rows, (lane>>4)*8 columns : A fragments and X[k][n] B fragments
//   "r-type" lane offset  (lane&7) rows, (lane>>3)*8 columns                      : X[n][k] B fragments
template <int ST>
__device__ __forceinline__ uint32_t lane_off_a(int lane) {
  return uint32_t((((lane & 7) + ((lane >> 3) & 1) * 8) * ST + (lane >> 4) * 8) * 2);
}
template <int ST>
__device__ __forceinline__ uint32_t lane_off_r(int lane) {
  return uint32_t(((lane & 7) * ST + (lane >> 3) * 8) * 2);
}
// A fragment: 16 rows [row0, +16) x 16 k [k0, +16) of a row-major tile (tile_a = tile base + lane_off_a).
template <int ST>
__device__ __forceinline__ void frag_a(uint32_t (&a)[4], uint32_t tile_a, int row0, int k0) {
  ldsm_x4(a, tile_a + uint32_t((row0 * ST + k0) * 2));
}
// B fragments for  D[m][n] += A[m][k] * X[n][k]  (X row-major [n][k]) for the 8 rows n0..n0+7 of X and the 32
// k-columns starting at kc: {b[0], b[1]} serve k-step kc/16, {b[2], b[3]} k-step kc/16 + 1 (tile_r = base + lane_off_r).
template <int ST>
__device__ __forceinline__ void frag_b_rows(uint32_t (&b)[4], uint32_t tile_r, int n0, int kc) {
  ldsm_x4(b, tile_r + uint32_t((n0 * ST + kc) * 2));
}
// acc(16x8) += A(16 x 16*KS) . X[n0..n0+7][0..16*KS)^T  with A fragments a[KS]
template <int ST, int KS>
__device__ __forceinline__ void mma_rows(float (&acc)[4], const uint32_t (&a)[KS][4], uint32_t tile_r, int n0) {
#pragma unroll
  for (int kp = 0; kp < KS / 2; ++kp) {
    uint32_t b[4];
    frag_b_rows<ST>(b, tile_r, n0, kp * 32);
    mma_16816(acc, a[2 * kp], b[0], b[1]);
    mma_16816(acc, a[2 * kp + 1], b[2], b[3]);
  }
}
// B fragments for  D[m][n] += A[m][k] * X[k][n]  (X row-major [k][n]): 16 k-rows k0.., 16 n-columns d0..:
// {b[0], b[1]} serve n-tile d0, {b[2], b[3]} n-tile d0 + 8 (tile_a = base + lane_off_a).
template <int ST>
__device__ __forceinline__ void frag_b_cols(uint32_t (&b)[4], uint32_t tile_a, int k0, int d0) {
  ldsm_x4_t(b, tile_a + uint32_t((k0 * ST + d0) * 2));
}

template <int NARR, int ST, int CPR>  // CPR = 16-byte chunks per padded row (4 or 8)
__device__ __forceinline__ void load_tiles(const __nv_bfloat16* const (&src)[NARR], const int (&ld)[NARR],
                                           __nv_bfloat16* const (&dst)[NARR], int L, int Lp, int chunks, int tid,
                                           int nthreads) {
  for (int i = tid; i < Lp * CPR; i += nthreads) {
    const int r = i / CPR, c = i % CPR;
    const bool ok = r < L && c < chunks;
#pragma unroll
    for (int a = 0; a < NARR; ++a) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (ok) v = *reinterpret_cast<const uint4*>(src[a] + size_t(r) * ld[a] + c * 8);
      *reinterpret_cast<uint4*>(dst[a] + r * ST + c * 8) = v;
    }
  }
}

template <int HD, int NT>
__global__ void __launch_bounds__(NT * 16) attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                           __nv_bfloat16* __restrict__ out, float* __restrict__ lse,
                                                           int L, int H, float scale) {
  pdl_trigger();
  constexpr int Lp = NT * 8, DN = HD / 8, KD = HD <= 32 ? 32 : 64, KS = KD / 16, ST = KD + 8, DNP = KD / 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* sK = sQ + Lp * ST;
  __nv_bfloat16* sV = sK + Lp * ST;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int D = H * HD;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
  pdl_wait();
  {
    const __nv_bfloat16* base = qkv + size_t(b) * L * 3 * D + h * HD;
    const __nv_bfloat16* const src[3] = {base, base + D, base + 2 * D};
    const int ld[3] = {3 * D, 3 * D, 3 * D};
    __nv_bfloat16* const dst[3] = {sQ, sK, sV};
    load_tiles<3, ST, KD / 8>(src, ld, dst, L, Lp, HD / 8, tid, NT * 16);
  }
  __syncthreads();

  const int q0 = w * 16;
  uint32_t aq[KS][4];
#pragma unroll
  const uint32_t qA = smem_u32(sQ) + lane_off_a<ST>(lane), kR = smem_u32(sK) + lane_off_r<ST>(lane),
                 vA = smem_u32(sV) + lane_off_a<ST>(lane);
  for (int ks = 0; ks < KS; ++ks) frag_a<ST>(aq[ks], qA, q0, ks * 16);
  const float sl2 = scale * 1.4426950408889634f;

  float o[DNP][4];
#pragma unroll
  for (int dn = 0; dn < DNP; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
  float M0, M1, l0 = 0.f, l1 = 0.f;
  if constexpr (NT <= 8) {
    // short sequences (encoder: 36 tokens, NT = 6): the whole 16 x (8 NT) score slab stays in registers (4 NT floats), so
    // Q.K^T is computed once: scores -> row max -> exp -> P.V.  Masked key columns are set to -inf (exp2 gives 0).
    float sc[NT][4];
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
      mma_rows<ST, KS>(sc[nt], aq, kR, nt * 8);
      if (nt * 8 + 8 > L) {  // warp-uniform
        const int c = nt * 8 + t * 2;
        if (c >= L) { sc[nt][0] = -INFINITY; sc[nt][2] = -INFINITY; }
        if (c + 1 >= L) { sc[nt][1] = -INFINITY; sc[nt][3] = -INFINITY; }
      }
      m0 = fmaxf(m0, fmaxf(sc[nt][0], sc[nt][1]));
      m1 = fmaxf(m1, fmaxf(sc[nt][2], sc[nt][3]));
    }
    M0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); M0 = fmaxf(M0, __shfl_xor_sync(0xffffffffu, M0, 2));
    M1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); M1 = fmaxf(M1, __shfl_xor_sync(0xffffffffu, M1, 2));
    const float ms0 = M0 * sl2, ms1 = M1 * sl2;
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
      float p[2][4];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int nt = 2 * kk + hf;
        p[hf][0] = ex2_approx(sc[nt][0] * sl2 - ms0); p[hf][1] = ex2_approx(sc[nt][1] * sl2 - ms0);
        p[hf][2] = ex2_approx(sc[nt][2] * sl2 - ms1); p[hf][3] = ex2_approx(sc[nt][3] * sl2 - ms1);
        l0 += p[hf][0] + p[hf][1];
        l1 += p[hf][2] + p[hf][3];
      }
      uint32_t a[4];
      a[0] = pack_bf16x2(p[0][0], p[0][1]); a[1] = pack_bf16x2(p[0][2], p[0][3]);
      a[2] = pack_bf16x2(p[1][0], p[1][1]); a[3] = pack_bf16x2(p[1][2], p[1][3]);
#pragma unroll
      for (int dp = 0; dp < DNP / 2; ++dp) {
        uint32_t bv[4];
        frag_b_cols<ST>(bv, vA, kk * 16, dp * 16);
        mma_16816(o[2 * dp], a, bv[0], bv[1]);
        if (2 * dp + 1 < DN) mma_16816(o[2 * dp + 1], a, bv[2], bv[3]);
      }
    }
  } else {
    // pass 1: row max only (no exponentials); key columns >= L exist only in the last tiles, so only those mask
    float m0 = -INFINITY, m1 = -INFINITY;
  #pragma unroll 2
    for (int nt = 0; nt < NT; ++nt) {
      float s[4] = {0.f, 0.f, 0.f, 0.f};
      mma_rows<ST, KS>(s, aq, kR, nt * 8);
      if (nt * 8 + 8 > L) {  // warp-uniform
        const int c = nt * 8 + t * 2;
        if (c >= L) { s[0] = -INFINITY; s[2] = -INFINITY; }
        if (c + 1 >= L) { s[1] = -INFINITY; s[3] = -INFINITY; }
      }
      m0 = fmaxf(m0, fmaxf(s[0], s[1]));
      m1 = fmaxf(m1, fmaxf(s[2], s[3]));
    }
    M0 = m0; M1 = m1;
    M0 = fmaxf(M0, __shfl_xor_sync(0xffffffffu, M0, 1)); M0 = fmaxf(M0, __shfl_xor_sync(0xffffffffu, M0, 2));
    M1 = fmaxf(M1, __shfl_xor_sync(0xffffffffu, M1, 1)); M1 = fmaxf(M1, __shfl_xor_sync(0xffffffffu, M1, 2));

    // pass 2: recompute S per 16-key block, P = exp(S - max), row sums and O += P.V; normalised at the end
    const float ms0 = M0 * sl2, ms1 = M1 * sl2;
  #pragma unroll 2
    for (int kk = 0; kk < NT / 2; ++kk) {
      float p[2][4];
  #pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int nt = 2 * kk + hf;
        p[hf][0] = p[hf][1] = p[hf][2] = p[hf][3] = 0.f;
        mma_rows<ST, KS>(p[hf], aq, kR, nt * 8);
        p[hf][0] = ex2_approx(p[hf][0] * sl2 - ms0); p[hf][1] = ex2_approx(p[hf][1] * sl2 - ms0);
        p[hf][2] = ex2_approx(p[hf][2] * sl2 - ms1); p[hf][3] = ex2_approx(p[hf][3] * sl2 - ms1);
        if (nt * 8 + 8 > L) {  // warp-uniform
          const int c = nt * 8 + t * 2;
          if (c >= L) { p[hf][0] = 0.f; p[hf][2] = 0.f; }
          if (c + 1 >= L) { p[hf][1] = 0.f; p[hf][3] = 0.f; }
        }
        l0 += p[hf][0] + p[hf][1];
        l1 += p[hf][2] + p[hf][3];
      }
      uint32_t a[4];
      a[0] = pack_bf16x2(p[0][0], p[0][1]); a[1] = pack_bf16x2(p[0][2], p[0][3]);
      a[2] = pack_bf16x2(p[1][0], p[1][1]); a[3] = pack_bf16x2(p[1][2], p[1][3]);
  #pragma unroll
      for (int dp = 0; dp < DNP / 2; ++dp) {
        uint32_t bv[4];
        frag_b_cols<ST>(bv, vA, kk * 16, dp * 16);
        mma_16816(o[2 * dp], a, bv[0], bv[1]);
        if (2 * dp + 1 < DN) mma_16816(o[2 * dp + 1], a, bv[2], bv[3]);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
  const int r0 = q0 + g, r1 = q0 + g + 8;
#pragma unroll
  for (int dn = 0; dn < DN; ++dn) {
    const int d = h * HD + dn * 8 + t * 2;
    if (r0 < L) *reinterpret_cast<uint32_t*>(out + (size_t(b) * L + r0) * D + d) = pack_bf16x2(o[dn][0] * inv0, o[dn][1] * inv0);
    if (r1 < L) *reinterpret_cast<uint32_t*>(out + (size_t(b) * L + r1) * D + d) = pack_bf16x2(o[dn][2] * inv1, o[dn][3] * inv1);
  }
  if (t == 0) {
    if (r0 < L) lse[(size_t(b) * H + h) * L + r0] = M0 * scale + logf(l0);
    if (r1 < L) lse[(size_t(b) * H + h) * L + r1] = M1 * scale + logf(l1);
  }
}

template <int HD, int NT>
__global__ void __launch_bounds__(NT * 16) attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                           const __nv_bfloat16* __restrict__ out,
                                                           const __nv_bfloat16* __restrict__ dout,
                                                           const float* __restrict__ lse,
                                                           __nv_bfloat16* __restrict__ dqkv, int L, int H, float scale) {
  pdl_trigger();
  constexpr int Lp = NT * 8, DN = HD / 8, KD = HD <= 32 ? 32 : 64, KS = KD / 16, ST = KD + 8, DNP = KD / 8, CPR = KD / 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* sK = sQ + Lp * ST;
  __nv_bfloat16* sV = sK + Lp * ST;
  __nv_bfloat16* sdO = sV + Lp * ST;
  float* sLse = reinterpret_cast<float*>(sdO + Lp * ST);  // pre-multiplied by log2(e)
  float* sDelta = sLse + Lp;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int D = H * HD;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
  pdl_wait();
  const float l2e = 1.4426950408889634f;
  {
    const __nv_bfloat16* base = qkv + size_t(b) * L * 3 * D + h * HD;
    const __nv_bfloat16* const src[4] = {base, base + D, base + 2 * D, dout + size_t(b) * L * D + h * HD};
    const int ld[4] = {3 * D, 3 * D, 3 * D, D};
    __nv_bfloat16* const dst[4] = {sQ, sK, sV, sdO};
    load_tiles<4, ST, CPR>(src, ld, dst, L, Lp, HD / 8, tid, NT * 16);
  }
  // delta[q] = sum_d dO[q,d] * O[q,d];  Lp*CPR == (CPR/2) * blockDim, so every lane runs every iteration
  // (full-warp shuffles); the CPR lanes of a row are adjacent
  for (int i = tid; i < Lp * CPR; i += NT * 16) {
    const int r = i / CPR, c = i % CPR;
    float part = 0.f;
    if (r < L && c < HD / 8) {
      const size_t ob = (size_t(b) * L + r) * D + h * HD + c * 8;
      const uint4 d_o = *reinterpret_cast<const uint4*>(dout + ob);
      const uint4 oo = *reinterpret_cast<const uint4*>(out + ob);
      const uint32_t dw[4] = {d_o.x, d_o.y, d_o.z, d_o.w}, ow[4] = {oo.x, oo.y, oo.z, oo.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) part += bf16_lo(dw[e]) * bf16_lo(ow[e]) + bf16_hi(dw[e]) * bf16_hi(ow[e]);
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    if (CPR == 8) part += __shfl_xor_sync(0xffffffffu, part, 4);
    if (c == 0) {
      sDelta[r] = part;
      sLse[r] = (r < L) ? lse[(size_t(b) * H + h) * L + r] * l2e : 0.f;
    }
  }
  __syncthreads();

  const float sl2 = scale * l2e;
  const uint32_t qA = smem_u32(sQ) + lane_off_a<ST>(lane), kA = smem_u32(sK) + lane_off_a<ST>(lane),
                 vA = smem_u32(sV) + lane_off_a<ST>(lane), oA = smem_u32(sdO) + lane_off_a<ST>(lane);
  const uint32_t qR = smem_u32(sQ) + lane_off_r<ST>(lane), kR = smem_u32(sK) + lane_off_r<ST>(lane),
                 vR = smem_u32(sV) + lane_off_r<ST>(lane), oR = smem_u32(sdO) + lane_off_r<ST>(lane);
  const int x0 = w * 16;  // first row of this warp's slab (queries in phase A, keys in phase B)
  const int r0 = x0 + g, r1 = x0 + g + 8;

  // ------------------------------ phase A: dQ ------------------------------
  {
    uint32_t aq[KS][4], ado[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) { frag_a<ST>(aq[ks], qA, x0, ks * 16); frag_a<ST>(ado[ks], oA, x0, ks * 16); }
    const float l0 = sLse[r0], l1 = sLse[r1], d0 = sDelta[r0], d1 = sDelta[r1];
    float dq[DNP][4];
#pragma unroll
    for (int dn = 0; dn < DNP; ++dn) dq[dn][0] = dq[dn][1] = dq[dn][2] = dq[dn][3] = 0.f;
#pragma unroll 2
    for (int kk = 0; kk < NT / 2; ++kk) {
      float ds[2][4];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int nt = 2 * kk + hf;
        float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
        mma_rows<ST, KS>(s, aq, kR, nt * 8);
        mma_rows<ST, KS>(dp, ado, vR, nt * 8);
        float p0 = ex2_approx(s[0] * sl2 - l0), p1 = ex2_approx(s[1] * sl2 - l0);
        float p2 = ex2_approx(s[2] * sl2 - l1), p3 = ex2_approx(s[3] * sl2 - l1);
        if (nt * 8 + 8 > L) {  // warp-uniform: only the tail tiles hold padded keys (s = 0 there, exp(-lse) may overflow)
          const int c = nt * 8 + t * 2;
          if (c >= L) { p0 = 0.f; p2 = 0.f; }
          if (c + 1 >= L) { p1 = 0.f; p3 = 0.f; }
        }
        ds[hf][0] = p0 * (dp[0] - d0); ds[hf][1] = p1 * (dp[1] - d0);
        ds[hf][2] = p2 * (dp[2] - d1); ds[hf][3] = p3 * (dp[3] - d1);
      }
      uint32_t a[4];
      a[0] = pack_bf16x2(ds[0][0], ds[0][1]); a[1] = pack_bf16x2(ds[0][2], ds[0][3]);
      a[2] = pack_bf16x2(ds[1][0], ds[1][1]); a[3] = pack_bf16x2(ds[1][2], ds[1][3]);
#pragma unroll
      for (int dp2 = 0; dp2 < DNP / 2; ++dp2) {
        uint32_t kt[4];
        frag_b_cols<ST>(kt, kA, kk * 16, dp2 * 16);
        mma_16816(dq[2 * dp2], a, kt[0], kt[1]);
        if (2 * dp2 + 1 < DN) mma_16816(dq[2 * dp2 + 1], a, kt[2], kt[3]);
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
    uint32_t ak[KS][4], av[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) { frag_a<ST>(ak[ks], kA, x0, ks * 16); frag_a<ST>(av[ks], vA, x0, ks * 16); }
    float dk[DNP][4], dv[DNP][4];
#pragma unroll
    for (int dn = 0; dn < DNP; ++dn) {
      dk[dn][0] = dk[dn][1] = dk[dn][2] = dk[dn][3] = 0.f;
      dv[dn][0] = dv[dn][1] = dv[dn][2] = dv[dn][3] = 0.f;
    }
    const bool k0ok = r0 < L, k1ok = r1 < L;
#pragma unroll 2
    for (int kk = 0; kk < NT / 2; ++kk) {
      float pt[2][4], ds[2][4];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int nt = 2 * kk + hf;
        float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
        mma_rows<ST, KS>(s, ak, qR, nt * 8);
        mma_rows<ST, KS>(dp, av, oR, nt * 8);
        // no masks needed: for query >= L the Q and dO rows are exactly zero (P^T there meets zero dO / Q rows,
        // sLse/sDelta are 0), and key rows >= L only feed dK/dV rows that are never stored
        const int c = nt * 8 + t * 2;  // query column
        const float2 lq = *reinterpret_cast<const float2*>(sLse + c), dl = *reinterpret_cast<const float2*>(sDelta + c);
        const float lq0 = lq.x, lq1 = lq.y, dl0 = dl.x, dl1 = dl.y;
        pt[hf][0] = ex2_approx(s[0] * sl2 - lq0);
        pt[hf][1] = ex2_approx(s[1] * sl2 - lq1);
        pt[hf][2] = ex2_approx(s[2] * sl2 - lq0);
        pt[hf][3] = ex2_approx(s[3] * sl2 - lq1);
        ds[hf][0] = pt[hf][0] * (dp[0] - dl0); ds[hf][1] = pt[hf][1] * (dp[1] - dl1);
        ds[hf][2] = pt[hf][2] * (dp[2] - dl0); ds[hf][3] = pt[hf][3] * (dp[3] - dl1);
      }
      uint32_t ap[4], as_[4];
      ap[0] = pack_bf16x2(pt[0][0], pt[0][1]); ap[1] = pack_bf16x2(pt[0][2], pt[0][3]);
      ap[2] = pack_bf16x2(pt[1][0], pt[1][1]); ap[3] = pack_bf16x2(pt[1][2], pt[1][3]);
      as_[0] = pack_bf16x2(ds[0][0], ds[0][1]); as_[1] = pack_bf16x2(ds[0][2], ds[0][3]);
      as_[2] = pack_bf16x2(ds[1][0], ds[1][1]); as_[3] = pack_bf16x2(ds[1][2], ds[1][3]);
#pragma unroll
      for (int dp2 = 0; dp2 < DNP / 2; ++dp2) {
        uint32_t dot[4], qt[4];
        frag_b_cols<ST>(dot, oA, kk * 16, dp2 * 16);
        frag_b_cols<ST>(qt, qA, kk * 16, dp2 * 16);
        mma_16816(dv[2 * dp2], ap, dot[0], dot[1]);
        mma_16816(dk[2 * dp2], as_, qt[0], qt[1]);
        if (2 * dp2 + 1 < DN) {
          mma_16816(dv[2 * dp2 + 1], ap, dot[2], dot[3]);
          mma_16816(dk[2 * dp2 + 1], as_, qt[2], qt[3]);
        }
      }
    }
#pragma unroll
    for (int dn = 0; dn < DN; ++dn) {
      const int d = h * HD + dn * 8 + t * 2;
      if (k0ok) {
        __nv_bfloat16* base = dqkv + (size_t(b) * L + r0) * 3 * D + d;
        *reinterpret_cast<uint32_t*>(base + D) = pack_bf16x2(dk[dn][0] * scale, dk[dn][1] * scale);
        *reinterpret_cast<uint32_t*>(base + 2 * D) = pack_bf16x2(dv[dn][0], dv[dn][1]);
      }
      if (k1ok) {
        __nv_bfloat16* base = dqkv + (size_t(b) * L + r1) * 3 * D + d;
        *reinterpret_cast<uint32_t*>(base + D) = pack_bf16x2(dk[dn][2] * scale, dk[dn][3] * scale);
        *reinterpret_cast<uint32_t*>(base + 2 * D) = pack_bf16x2(dv[dn][2], dv[dn][3]);
      }
    }
  }
}

template <int HD, int NT>
static int launch_fwd(const void* qkv, int B, int L, int H, void* out, float* lse, cudaStream_t st) {
  constexpr int Lp = NT * 8, ST = (HD <= 32 ? 32 : 64) + 8;
  const size_t smem = size_t(3 * Lp * ST) * 2;
  static bool done[kMaxDevices] = {};
  if (!ensure_dyn_smem(attn_fwd_kernel<HD, NT>, smem, done)) return MAE_ERR_LAUNCH;
  mae_launch(attn_fwd_kernel<HD, NT>, dim3(B * H), dim3(NT * 16), smem, st, reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                        reinterpret_cast<__nv_bfloat16*>(out), lse, L, H,
                                                        1.0f / sqrtf(float(HD)));
  return MAE_CHECK_LAUNCH();
}

template <int HD, int NT>
static int launch_bwd(const void* qkv, const void* out, const void* dout, const float* lse, int B, int L, int H,
                      void* dqkv, cudaStream_t st) {
  constexpr int Lp = NT * 8, ST = (HD <= 32 ? 32 : 64) + 8;
  const size_t smem = size_t(4 * Lp * ST) * 2 + size_t(2 * Lp) * 4;
  static bool done[kMaxDevices] = {};
  if (!ensure_dyn_smem(attn_bwd_kernel<HD, NT>, smem, done)) return MAE_ERR_LAUNCH;
  mae_launch(attn_bwd_kernel<HD, NT>, dim3(B * H), dim3(NT * 16), smem, st, reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<const __nv_bfloat16*>(out),
      reinterpret_cast<const __nv_bfloat16*>(dout), lse, reinterpret_cast<__nv_bfloat16*>(dqkv), L, H,
      1.0f / sqrtf(float(HD)));
  return MAE_CHECK_LAUNCH();
}

static int pick_nt(int L) {
  const int need = 2 * ((L + 15) / 16);
  const int opts[5] = {4, 6, 10, 20, 26};
  for (int i = 0; i < 5; ++i)
    if (opts[i] >= need) return opts[i];
  return 0;
}

}  // namespace mae

using namespace mae;

#define MAE_ATTN_CASE(FN, HDV, NTV, ...) case HDV * 100 + NTV: return FN<HDV, NTV>(__VA_ARGS__);
#define MAE_ATTN_DISPATCH(FN, ...)                                                                            \
  switch (hd * 100 + nt) {                                                                                    \
    MAE_ATTN_CASE(FN, 24, 4, __VA_ARGS__) MAE_ATTN_CASE(FN, 24, 6, __VA_ARGS__) MAE_ATTN_CASE(FN, 24, 10, __VA_ARGS__) \
    MAE_ATTN_CASE(FN, 24, 20, __VA_ARGS__) MAE_ATTN_CASE(FN, 24, 26, __VA_ARGS__)                             \
    MAE_ATTN_CASE(FN, 32, 4, __VA_ARGS__) MAE_ATTN_CASE(FN, 32, 6, __VA_ARGS__) MAE_ATTN_CASE(FN, 32, 10, __VA_ARGS__) \
    MAE_ATTN_CASE(FN, 32, 20, __VA_ARGS__) MAE_ATTN_CASE(FN, 32, 26, __VA_ARGS__)                             \
    MAE_ATTN_CASE(FN, 64, 4, __VA_ARGS__) MAE_ATTN_CASE(FN, 64, 6, __VA_ARGS__) MAE_ATTN_CASE(FN, 64, 10, __VA_ARGS__) \
    MAE_ATTN_CASE(FN, 64, 20, __VA_ARGS__) MAE_ATTN_CASE(FN, 64, 26, __VA_ARGS__)                             \
    default: return MAE_ERR_SHAPE;                                                                            \
  }

static int attention_fwd_impl(const void* qkv, int B, int L, int H, int hd, void* out, float* lse, int q_rows, cudaStream_t stream) {
  if (B <= 0) return MAE_OK;
  if (!qkv || !out || !lse) return MAE_ERR_NULL;
  const int nt = pick_nt(L);
  if (!nt || L < 1 || H < 1) return MAE_ERR_SHAPE;
  {
    static const bool no_tc = getenv("MAE_ATTN_NO_TC") != nullptr;  // A/B switch: keep everything on the mma.sync kernels
    if (!no_tc && hd == 32 && L > 128 && L <= 160 && (H % 2) == 0) return attn_fwd_tc(qkv, B, L, H, out, lse, q_rows, stream);
    if (!no_tc && attn_enc_covers(L, hd)) return attn_fwd_enc(qkv, B, L, H, hd, out, lse, stream);
  }
  MAE_ATTN_DISPATCH(launch_fwd, qkv, B, L, H, out, lse, stream)
}

extern "C" int mae_attention_fwd(const void* qkv, int B, int L, int H, int hd, void* out, float* lse,
                                 cudaStream_t stream) {
  return attention_fwd_impl(qkv, B, L, H, hd, out, lse, L, stream);
}

extern "C" int mae_attention_fwd_q(const void* qkv, int B, int L, int H, int hd, int q_rows, void* out, float* lse,
                                   cudaStream_t stream) {
  if (q_rows < 1 || q_rows > L) return MAE_ERR_SHAPE;
  return attention_fwd_impl(qkv, B, L, H, hd, out, lse, q_rows, stream);
}

static int attention_bwd_impl(const void* qkv, const void* out, const void* dout, const float* lse, int B, int L, int H, int hd,
                              void* dqkv, float* delta_ws, int q_rows, cudaStream_t stream) {
  if (B <= 0) return MAE_OK;
  if (!qkv || !out || !dout || !lse || !dqkv) return MAE_ERR_NULL;
  const int nt = pick_nt(L);
  if (!nt || L < 1 || H < 1) return MAE_ERR_SHAPE;
  {
    static const bool no_tc = getenv("MAE_ATTN_NO_TC") != nullptr || getenv("MAE_ATTN_NO_TC_BWD") != nullptr;
    static const bool no_delta = getenv("MAE_ATTN_NO_DELTA_PREPASS") != nullptr;  // A/B switch
    if (!no_tc && hd == 32 && L > 128 && L <= 160 && (H % 2) == 0)
      return attn_bwd_tc(qkv, out, dout, lse, B, L, H, dqkv, no_delta ? nullptr : delta_ws, q_rows, stream);
    if (!no_tc && attn_enc_covers(L, hd)) return attn_bwd_enc(qkv, out, dout, lse, B, L, H, hd, dqkv, stream);
  }
  MAE_ATTN_DISPATCH(launch_bwd, qkv, out, dout, lse, B, L, H, dqkv, stream)
}

extern "C" int mae_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, int B, int L,
                                 int H, int hd, void* dqkv, cudaStream_t stream) {
  return attention_bwd_impl(qkv, out, dout, lse, B, L, H, hd, dqkv, nullptr, L, stream);
}

extern "C" int mae_attention_bwd_ws(const void* qkv, const void* out, const void* dout, const float* lse, int B, int L,
                                    int H, int hd, void* dqkv, void* ws, long long ws_bytes, cudaStream_t stream) {
  float* delta = (ws && ws_bytes >= (long long)B * L * H * 4 && (reinterpret_cast<uintptr_t>(ws) & 15) == 0) ? reinterpret_cast<float*>(ws) : nullptr;
  return attention_bwd_impl(qkv, out, dout, lse, B, L, H, hd, dqkv, delta, L, stream);
}

extern "C" int mae_attention_bwd_q(const void* qkv, const void* out, const void* dout, const float* lse, int B, int L,
                                   int H, int hd, int q_rows, void* dqkv, cudaStream_t stream) {
  if (q_rows < 1 || q_rows > L) return MAE_ERR_SHAPE;
  return attention_bwd_impl(qkv, out, dout, lse, B, L, H, hd, dqkv, nullptr, q_rows, stream);
}
