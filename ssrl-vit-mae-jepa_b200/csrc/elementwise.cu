// HBM-bound kernels of the MAE step: token masking (warp bitonic sort), patch gather, token assembly,
// LayerNorm fwd/bwd, decoder scatter, masked MSE, column sums, casts, clip+AdamW.
#include "mae_internal.h"
#include <stdlib.h>
#include "ptx.cuh"

namespace mae {

// ---------------------------------------------------------------------------------------------
// random_token_mask: one warp sorts one row of <= 256 (key,index) pairs with a register/shuffle
// bitonic network.  Composite 64-bit keys (monotone float bits << 32 | index) make the order total,
// which is exactly a stable ascending sort.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_to_ordered(float x) {
  x = x + 0.0f;  // -0.0 -> +0.0 so that they compare equal like torch
  uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(128) token_mask_kernel(const float* __restrict__ noise, int B, int L, int num_keep,
                                                        int64_t* __restrict__ idx_keep,
                                                        int64_t* __restrict__ idx_mask) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const int lane = threadIdx.x & 31;
  unsigned long long v[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = lane * 8 + r;
    if (i < L) {
      const float x = (i == 0) ? -1.0f : noise[size_t(row) * L + i];
      v[r] = (static_cast<unsigned long long>(float_to_ordered(x)) << 32) | static_cast<unsigned>(i);
    } else {
      v[r] = ~0ull;
    }
  }
#pragma unroll
  for (int k = 2; k <= 256; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 8) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, v[r], j >> 3);
          const int i = lane * 8 + r;
          const bool up = (i & k) == 0;
          const bool lower = (i & j) == 0;
          const unsigned long long mn = v[r] < other ? v[r] : other;
          const unsigned long long mx = v[r] < other ? other : v[r];
          v[r] = (up == lower) ? mn : mx;
        }
      } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          if ((r & j) == 0) {
            const int i = lane * 8 + r;
            const bool up = (i & k) == 0;
            const unsigned long long a = v[r], b = v[r | j];
            const bool sw = (a > b) == up;
            v[r] = sw ? b : a;
            v[r | j] = sw ? a : b;
          }
        }
      }
    }
  }
  const int n_mask = L - num_keep;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = lane * 8 + r;
    const int64_t idx = static_cast<int64_t>(v[r] & 0xffffffffull);
    if (i < num_keep) idx_keep[size_t(row) * num_keep + i] = idx;
    else if (i < L) idx_mask[size_t(row) * n_mask + (i - num_keep)] = idx;
  }
}

// ---------------------------------------------------------------------------------------------
// patch gather (im2col of the kept patches only), one warp per kept patch
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_patches_kernel(const float* __restrict__ images,
                                                            const int64_t* __restrict__ idx_keep, int B, int C,
                                                            int img, int patch, int num_keep,
                                                            __nv_bfloat16* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int nk = num_keep - 1;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (long long)B * nk) return;
  const int lane = threadIdx.x & 31;
  const int b = int(w / nk), j = int(w % nk) + 1;
  const int pidx = int(idx_keep[size_t(b) * num_keep + j]) - 1;
  const int grid = img / patch;
  const int pr = pidx / grid, pcn = pidx % grid;
  const int pp = patch * patch, pd = C * pp;
  const float* src = images + size_t(b) * C * img * img;
  __nv_bfloat16* dst = out + size_t(w) * pd;
  if ((patch & 3) == 0) {  // 4 consecutive px of one (c, py) row per lane: 16-byte loads, 8-byte stores
    const int p4 = patch >> 2;
    for (int q = lane; q < (pd >> 2); q += 32) {
      const int c = q / (patch * p4), py = (q / p4) % patch, px = (q % p4) * 4;
      const float4 v = *reinterpret_cast<const float4*>(src + (size_t(c) * img + pr * patch + py) * img + pcn * patch + px);
      *reinterpret_cast<uint2*>(dst + q * 4) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
  } else {
    for (int q = lane; q < pd; q += 32) {
      const int c = q / pp, py = (q % pp) / patch, px = q % patch;
      dst[q] = __float2bfloat16_rn(src[(size_t(c) * img + pr * patch + py) * img + pcn * patch + px]);
    }
  }
}

// Grid-stride walk over a (batch b, row j < L, float4 column d4 < D4) index space without a division per element: the
// 64-bit divisions happen once per thread, every further step is three adds with carries.
struct RowWalk {
  int d4, j, b;      // current position
  int dd, dj, db;    // the grid stride decomposed the same way
  int D4, L;
  __device__ __forceinline__ RowWalk(long long i0, long long stride, int D4_, int L_) : D4(D4_), L(L_) {
    d4 = int(i0 % D4); const long long r0 = i0 / D4; j = int(r0 % L); b = int(r0 / L);
    dd = int(stride % D4); const long long rs = stride / D4; dj = int(rs % L); db = int(rs / L);
  }
  __device__ __forceinline__ void next() {
    d4 += dd;
    const int c = d4 >= D4;
    d4 -= c ? D4 : 0;
    j += dj + c;
    const int c2 = j >= L;
    j -= c2 ? L : 0;
    b += db + c2;
  }
};

__global__ void __launch_bounds__(256) assemble_tokens_kernel(const __nv_bfloat16* __restrict__ tok,
                                                             const float* __restrict__ cls,
                                                             const float* __restrict__ pos,
                                                             const int64_t* __restrict__ idx_keep, int B,
                                                             int num_keep, int D, float* __restrict__ x) {
  pdl_trigger();
  pdl_wait();
  const int D4 = D >> 2;
  const long long n = (long long)B * num_keep * D4, stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  RowWalk w(i0, stride, D4, num_keep);
  for (long long i = i0; i < n; i += stride, w.next()) {
    const int d = w.d4 * 4, j = w.j, b = w.b;
    const long long r = (long long)b * num_keep + j;
    const int tokidx = int(idx_keep[r]);
    float4 v;
    if (j == 0) {
      v = *reinterpret_cast<const float4*>(cls + d);
    } else {
      const uint2 t = *reinterpret_cast<const uint2*>(tok + (size_t(b) * (num_keep - 1) + (j - 1)) * D + d);
      v = make_float4(bf16_lo(t.x), bf16_hi(t.x), bf16_lo(t.y), bf16_hi(t.y));
    }
    const float4 p = *reinterpret_cast<const float4*>(pos + size_t(tokidx) * D + d);
    v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    *reinterpret_cast<float4*>(x + r * D + d) = v;
  }
}

// sum over the LPR-lane group of the calling lane (LPR = 32: the warp)
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm forward / backward: LPR lanes per row (32: one warp per row; 16: two rows per warp), D <= 4*LPR*NV,
// D % 4 == 0; a lane owns the float4 column groups {4*LPR*j + 4*sub}, j < NV, sub = lane % LPR.  With one warp per row
// D = 144 / 192 (configs/mae.yaml) fill 36 / 48 of the 64 float4 slots of two passes (the second pass is a 64- / 256-byte
// request per row); 16 lanes per row and NV = 3 fill 36 / 48 of 48 (D = 192: every lane busy in every pass).
// NV = 4 / 8 with LPR = 32 cover D <= 512 / 1024.
// ROWS rows are processed per warp iteration with all their loads issued before the first reduction (these
// kernels are pure HBM streams, so bytes in flight per warp is what matters).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ long long map_row(long long r, const int64_t* gather_idx, int n_sel, int L_src) {
  if (gather_idx == nullptr) return r;
  const long long b = (long long)(uint32_t(r) / uint32_t(n_sel));  // rows is an int: 32-bit division
  return b * L_src + __ldg(gather_idx + r);
}

// ADD: the normalised row is x[map(r)] + delta[r] (delta bf16, compact rows: a Linear output that has not been added to
// the residual stream yet); the sum is also written to x_out[r] (fp32, compact) for the backward.
template <int NV, int ROWS, bool GATHER, bool ADD = false, int LPR = 32>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta,
                                                           const int64_t* __restrict__ gather_idx, int n_sel,
                                                           int L_src, int rows, int D, float eps,
                                                           __nv_bfloat16* __restrict__ xn, float* __restrict__ mean,
                                                           float* __restrict__ rstd,
                                                           const __nv_bfloat16* __restrict__ delta,
                                                           float* __restrict__ x_out) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & (LPR - 1);  // position inside the row's lane group
  const int wpb = blockDim.x / LPR;          // rows in flight per block and pass
  float4 gm[NV], bt[NV];
  bool has[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = 4 * LPR * j + 4 * lane;
    has[j] = c < D;
    gm[j] = bt[j] = make_float4(0, 0, 0, 0);
    if (has[j]) { gm[j] = *reinterpret_cast<const float4*>(gamma + c); bt[j] = *reinterpret_cast<const float4*>(beta + c); }
  }
  const float invD = 1.0f / D;
  const long long stride = (long long)gridDim.x * wpb;
  // source rows of the NEXT iteration (the index load of a gathered norm is fetched one iteration ahead, so that it is
  // not a dependent round trip in front of every row load)
  long long xr_next[ROWS];
  const long long ra0 = (long long)blockIdx.x * wpb + (threadIdx.x / LPR);
  // two rows per warp (LPR = 16): the row loop must stay warp-uniform (full-mask shuffles inside), so it runs while the
  // warp's FIRST row exists; a missing second row is processed as a duplicate of the first and not stored
  const int gsel = int(threadIdx.x / LPR) % (32 / LPR);
#pragma unroll
  for (int k = 0; k < ROWS; ++k) {
    const long long r = ra0 + k * stride;
    xr_next[k] = (GATHER && ra0 - gsel < rows) ? map_row(r < rows ? r : ra0 - gsel, gather_idx, n_sel, L_src) : 0;
  }
  for (long long ra = ra0; ra - gsel < rows; ra += ROWS * stride) {
    float4 v[ROWS][NV];
    uint2 dl[ROWS][NV];
    long long rr[ROWS];
    bool ok[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      rr[k] = ra + k * stride;
      ok[k] = rr[k] < rows;
      if (!ok[k]) rr[k] = ra - gsel;
      const float* src = x + (GATHER ? xr_next[k] : rr[k]) * D;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        v[k][j] = make_float4(0, 0, 0, 0);
        dl[k][j] = make_uint2(0u, 0u);
        if (has[j]) {
          v[k][j] = *reinterpret_cast<const float4*>(src + 4 * LPR * j + 4 * lane);
          if (ADD) dl[k][j] = *reinterpret_cast<const uint2*>(delta + rr[k] * D + 4 * LPR * j + 4 * lane);
        }
      }
    }
    if (ADD) {
#pragma unroll
      for (int k = 0; k < ROWS; ++k)
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          v[k][j].x += bf16_lo(dl[k][j].x); v[k][j].y += bf16_hi(dl[k][j].x);
          v[k][j].z += bf16_lo(dl[k][j].y); v[k][j].w += bf16_hi(dl[k][j].y);
          if (has[j] && ok[k]) *reinterpret_cast<float4*>(x_out + rr[k] * D + 4 * LPR * j + 4 * lane) = v[k][j];
        }
    }
    if (GATHER) {
      const long long rn = ra + ROWS * stride;
      if (rn - gsel < rows) {
#pragma unroll
        for (int k = 0; k < ROWS; ++k) {
          const long long r = rn + k * stride;
          xr_next[k] = map_row(r < rows ? r : rn - gsel, gather_idx, n_sel, L_src);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) sum += v[k][j].x + v[k][j].y + v[k][j].z + v[k][j].w;
      const float mu = group_sum<LPR>(sum) * invD;
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (has[j]) {
          const float a = v[k][j].x - mu, b = v[k][j].y - mu, c = v[k][j].z - mu, d = v[k][j].w - mu;
          sq += a * a + b * b + c * c + d * d;
        }
      const float rs = rsqrtf(group_sum<LPR>(sq) * invD + eps);
      if (!ok[k]) continue;
      const long long r = rr[k];
      __nv_bfloat16* dst = xn + r * D;
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (has[j]) {
          uint2 o;
          o.x = pack_bf16x2((v[k][j].x - mu) * rs * gm[j].x + bt[j].x, (v[k][j].y - mu) * rs * gm[j].y + bt[j].y);
          o.y = pack_bf16x2((v[k][j].z - mu) * rs * gm[j].z + bt[j].z, (v[k][j].w - mu) * rs * gm[j].w + bt[j].w);
          *reinterpret_cast<uint2*>(dst + 4 * LPR * j + 4 * lane) = o;
        }
      if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
    }
  }
}

// DINC: dx_in is indexed by the compact row r (like g), not by the mapped row (the residual gradient of a gathered
// sub-stream: the masked-rows-only tail of the last decoder block).
template <int NV, int ROWS, int MINB, bool GATHER, bool DINC = false, int LPR = 32>
__global__ void __launch_bounds__(256, MINB) layernorm_bwd_kernel(
    const __nv_bfloat16* __restrict__ g, const float* __restrict__ x, const float* __restrict__ mean,
    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ dx_in,
    const int64_t* __restrict__ gather_idx, int n_sel, int L_src, int rows, int D, float* __restrict__ dx_out,
    __nv_bfloat16* __restrict__ dx_out_bf16, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_dg[128 * NV], s_db[128 * NV];
  const int lane = threadIdx.x & (LPR - 1);  // position inside the row's lane group
  const int wpb = blockDim.x / LPR;          // rows in flight per block and pass
  for (int i = threadIdx.x; i < 128 * NV; i += blockDim.x) { s_dg[i] = 0.f; s_db[i] = 0.f; }
  __syncthreads();
  float4 wv[NV];
  bool has[NV];
  float dg[NV][4], db[NV][4];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = 4 * LPR * j + 4 * lane;
    has[j] = c < D;
    wv[j] = make_float4(0, 0, 0, 0);
    if (has[j]) wv[j] = *reinterpret_cast<const float4*>(gamma + c);
#pragma unroll
    for (int e = 0; e < 4; ++e) { dg[j][e] = 0.f; db[j][e] = 0.f; }
  }
  const float invD = 1.0f / D;
  const long long stride = (long long)gridDim.x * wpb;
  long long xr_next[ROWS];  // see layernorm_fwd_kernel
  const long long ra0 = (long long)blockIdx.x * wpb + (threadIdx.x / LPR);
  // two rows per warp (LPR = 16): the row loop must stay warp-uniform (full-mask shuffles inside), so it runs while the
  // warp's FIRST row exists; a missing second row is processed as a duplicate of the first and not stored
  const int gsel = int(threadIdx.x / LPR) % (32 / LPR);
#pragma unroll
  for (int k = 0; k < ROWS; ++k) {
    const long long r = ra0 + k * stride;
    xr_next[k] = (GATHER && ra0 - gsel < rows) ? map_row(r < rows ? r : ra0 - gsel, gather_idx, n_sel, L_src) : 0;
  }
  for (long long ra = ra0; ra - gsel < rows; ra += ROWS * stride) {
    long long rrow[ROWS], xr[ROWS];
    bool ok[ROWS];
    uint2 gg[ROWS][NV];
    float4 xv[ROWS][NV], di[ROWS][NV];
    float mu[ROWS], rs[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      rrow[k] = ra + k * stride;
      ok[k] = rrow[k] < rows;
      if (!ok[k]) rrow[k] = ra - gsel;
      xr[k] = GATHER ? xr_next[k] : rrow[k];
      mu[k] = mean[rrow[k]]; rs[k] = rstd[rrow[k]];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = 4 * LPR * j + 4 * lane;
        gg[k][j] = make_uint2(0, 0);
        xv[k][j] = di[k][j] = make_float4(0, 0, 0, 0);
        if (has[j]) {
          gg[k][j] = *reinterpret_cast<const uint2*>(g + rrow[k] * D + c);
          xv[k][j] = *reinterpret_cast<const float4*>(x + xr[k] * D + c);
          if (dx_in) di[k][j] = *reinterpret_cast<const float4*>(dx_in + (DINC ? rrow[k] : xr[k]) * D + c);
        }
      }
    }
    if (GATHER) {
      const long long rn = ra + ROWS * stride;
      if (rn - gsel < rows) {
#pragma unroll
        for (int k = 0; k < ROWS; ++k) {
          const long long r = rn + k * stride;
          xr_next[k] = map_row(r < rows ? r : rn - gsel, gather_idx, n_sel, L_src);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      const float live = ok[k] ? 1.f : 0.f;  // the duplicate row of a ragged tail must not count twice
      float s1 = 0.f, s2 = 0.f;
      float xh[NV][4], dxh[NV][4];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float gv[4] = {bf16_lo(gg[k][j].x), bf16_hi(gg[k][j].x), bf16_lo(gg[k][j].y), bf16_hi(gg[k][j].y)};
        const float xs[4] = {xv[k][j].x, xv[k][j].y, xv[k][j].z, xv[k][j].w};
        const float ws[4] = {wv[j].x, wv[j].y, wv[j].z, wv[j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          xh[j][e] = has[j] ? (xs[e] - mu[k]) * rs[k] : 0.f;
          dg[j][e] += live * gv[e] * xh[j][e];
          db[j][e] += live * gv[e];
          dxh[j][e] = gv[e] * ws[e];
          s1 += dxh[j][e];
          s2 += dxh[j][e] * xh[j][e];
        }
      }
      s1 = group_sum<LPR>(s1) * invD;
      s2 = group_sum<LPR>(s2) * invD;
      if (!ok[k]) continue;
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (has[j]) {
          const int c = 4 * LPR * j + 4 * lane;
          float4 o;
          o.x = rs[k] * (dxh[j][0] - s1 - xh[j][0] * s2) + di[k][j].x;
          o.y = rs[k] * (dxh[j][1] - s1 - xh[j][1] * s2) + di[k][j].y;
          o.z = rs[k] * (dxh[j][2] - s1 - xh[j][2] * s2) + di[k][j].z;
          o.w = rs[k] * (dxh[j][3] - s1 - xh[j][3] * s2) + di[k][j].w;
          *reinterpret_cast<float4*>(dx_out + xr[k] * D + c) = o;
          if (dx_out_bf16)
            *reinterpret_cast<uint2*>(dx_out_bf16 + xr[k] * D + c) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        }
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j)
    if (has[j]) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        atomicAdd(&s_dg[4 * LPR * j + 4 * lane + e], dg[j][e]);
        atomicAdd(&s_db[4 * LPR * j + 4 * lane + e], db[j][e]);
      }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(dgamma + i, s_dg[i]);
    atomicAdd(dbeta + i, s_db[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// decoder input: mask tokens everywhere, embedded kept tokens at idx_keep, + decoder pos embed
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scatter_decoder_kernel(const __nv_bfloat16* __restrict__ emb,
                                                             const float* __restrict__ mask_token,
                                                             const float* __restrict__ dec_pos,
                                                             const int64_t* __restrict__ idx_keep,
                                                             const int64_t* __restrict__ idx_mask, int B, int L,
                                                             int num_keep, int D, float* __restrict__ x_dec, int grouped) {
  // grouped: the stream is stored as [mask tokens in idx_mask order | kept tokens in idx_keep order] per sample instead
  // of in token order (row-wise operators and attention do not care; the positional embedding still follows the token)
  pdl_trigger();
  pdl_wait();
  const int D4 = D >> 2;
  const long long n = (long long)B * L * D4, stride = (long long)gridDim.x * blockDim.x;
  const int n_mask = L - num_keep;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  RowWalk w(i0, stride, D4, L);
  for (long long i = i0; i < n; i += stride, w.next()) {
    const int d = w.d4 * 4, jj = w.j, b = w.b;
    int l;
    float4 v;
    if (jj < num_keep) {
      l = int(idx_keep[size_t(b) * num_keep + jj]);
      const uint2 e = *reinterpret_cast<const uint2*>(emb + (size_t(b) * num_keep + jj) * D + d);
      v = make_float4(bf16_lo(e.x), bf16_hi(e.x), bf16_lo(e.y), bf16_hi(e.y));
    } else {
      l = int(idx_mask[size_t(b) * n_mask + (jj - num_keep)]);
      v = *reinterpret_cast<const float4*>(mask_token + d);
    }
    const float4 p = *reinterpret_cast<const float4*>(dec_pos + size_t(l) * D + d);
    v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    const int dst_row = grouped ? (jj < num_keep ? n_mask + jj : jj - num_keep) : l;
    *reinterpret_cast<float4*>(x_dec + (size_t(b) * L + dst_row) * D + d) = v;
  }
}

// scatter_decoder_kernel fused with the first decoder block's norm1: one warp per token (two tokens per iteration); the
// row is assembled in registers (embedded kept token or mask token, + pos), written to the fp32 stream, normalised and
// written as bf16 with its statistics: the stream is not read back by a separate LayerNorm launch.
template <int NV>
__global__ void __launch_bounds__(256, 3) scatter_decoder_ln_kernel(
    const __nv_bfloat16* __restrict__ emb, const float* __restrict__ mask_token, const float* __restrict__ dec_pos,
    const int64_t* __restrict__ idx_keep, const int64_t* __restrict__ idx_mask, int B, int L, int num_keep, int D,
    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float* __restrict__ x_dec,
    __nv_bfloat16* __restrict__ xn, float* __restrict__ mean, float* __restrict__ rstd) {
  pdl_trigger();
  pdl_wait();
  constexpr int ROWS = 2;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int n_mask = L - num_keep;
  float4 gm[NV], bt[NV], mt[NV];
  bool has[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = 128 * j + 4 * lane;
    has[j] = c < D;
    gm[j] = bt[j] = mt[j] = make_float4(0, 0, 0, 0);
    if (has[j]) {
      gm[j] = *reinterpret_cast<const float4*>(gamma + c);
      bt[j] = *reinterpret_cast<const float4*>(beta + c);
      mt[j] = *reinterpret_cast<const float4*>(mask_token + c);
    }
  }
  const float invD = 1.0f / D;
  const long long rows = (long long)B * L, stride = (long long)gridDim.x * wpb;
  for (long long ra = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); ra < rows; ra += ROWS * stride) {
    float4 v[ROWS][NV], pp[ROWS][NV];
    long long orow[ROWS];  // destination row b * L + l
    bool ok[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      long long r = ra + k * stride;
      ok[k] = r < rows;
      if (!ok[k]) r = ra;
      const uint32_t b = uint32_t(r) / uint32_t(L), jj = uint32_t(r) - b * uint32_t(L);  // B * L is an int
      const bool kept = jj < uint32_t(num_keep);
      const long long src = kept ? (long long)b * num_keep + jj : (long long)b * n_mask + (jj - num_keep);
      const int l = int(__ldg((kept ? idx_keep : idx_mask) + src));
      orow[k] = (long long)b * L + l;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = 128 * j + 4 * lane;
        v[k][j] = mt[j];
        pp[k][j] = make_float4(0, 0, 0, 0);
        if (has[j]) {
          pp[k][j] = *reinterpret_cast<const float4*>(dec_pos + (long long)l * D + c);
          if (kept) {
            const uint2 e = *reinterpret_cast<const uint2*>(emb + src * D + c);
            v[k][j] = make_float4(bf16_lo(e.x), bf16_hi(e.x), bf16_lo(e.y), bf16_hi(e.y));
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        v[k][j].x += pp[k][j].x; v[k][j].y += pp[k][j].y; v[k][j].z += pp[k][j].z; v[k][j].w += pp[k][j].w;
        sum += v[k][j].x + v[k][j].y + v[k][j].z + v[k][j].w;  // lanes past D hold zeros
      }
      const float mu = warp_sum(sum) * invD;
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (has[j]) {
          const float a = v[k][j].x - mu, b2 = v[k][j].y - mu, c2 = v[k][j].z - mu, d2 = v[k][j].w - mu;
          sq += a * a + b2 * b2 + c2 * c2 + d2 * d2;
        }
      const float rs = rsqrtf(warp_sum(sq) * invD + eps);
      if (!ok[k]) continue;
      float* xd = x_dec + orow[k] * D;
      __nv_bfloat16* dst = xn + orow[k] * D;
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (has[j]) {
          const int c = 128 * j + 4 * lane;
          *reinterpret_cast<float4*>(xd + c) = v[k][j];
          uint2 o;
          o.x = pack_bf16x2((v[k][j].x - mu) * rs * gm[j].x + bt[j].x, (v[k][j].y - mu) * rs * gm[j].y + bt[j].y);
          o.y = pack_bf16x2((v[k][j].z - mu) * rs * gm[j].z + bt[j].z, (v[k][j].w - mu) * rs * gm[j].w + bt[j].w);
          *reinterpret_cast<uint2*>(dst + c) = o;
        }
      if (lane == 0) { mean[orow[k]] = mu; rstd[orow[k]] = rs; }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// masked MSE: one warp per masked patch; the target is read straight from the image
// ---------------------------------------------------------------------------------------------
static constexpr int kMsePartials = 4096;
static constexpr int kMaxPatchDim = 768;  // 16*16*3

__global__ void __launch_bounds__(256) masked_mse_kernel(const __nv_bfloat16* __restrict__ pred,
                                                        const float* __restrict__ images,
                                                        const int64_t* __restrict__ idx_mask, int B, int C, int img,
                                                        int patch, int n_mask, int norm_pix, float inv_numel,
                                                        float* __restrict__ partials,
                                                        __nv_bfloat16* __restrict__ dpred,
                                                        float* __restrict__ target) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_t[];  // per warp: patch_dim floats (target in (py,px,c) order)
  __shared__ float s_red[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int pp = patch * patch, pd = C * pp, grid = img / patch;
  float* st = s_t + wid * pd;
  float acc = 0.f;
  const long long nrows = (long long)B * n_mask;
  for (long long r = (long long)blockIdx.x * wpb + wid; r < nrows; r += (long long)gridDim.x * wpb) {
    const int b = int(r / n_mask);
    int pidx = int(idx_mask[r]) - 1;
    if (pidx < 0) pidx = 0;  // torch.clamp(idx_mask - 1, min=0)
    const int pr = pidx / grid, pcn = pidx % grid;
    const float* src = images + size_t(b) * C * img * img;
    float s = 0.f, ss = 0.f;
    if ((patch & 3) == 0) {
      // vector path: one float4 = 4 consecutive px of one (c, py) image row
      const int p4 = patch >> 2, nq = pd >> 2;
      for (int q = lane; q < nq; q += 32) {
        const int c = q / (patch * p4), py = (q / p4) % patch, px = (q % p4) * 4;
        const float4 v = *reinterpret_cast<const float4*>(src + (size_t(c) * img + pr * patch + py) * img + pcn * patch + px);
        float* d = st + (py * patch + px) * C + c;
        d[0] = v.x; d[C] = v.y; d[2 * C] = v.z; d[3 * C] = v.w;
        s += v.x + v.y + v.z + v.w; ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
    } else {
      for (int q = lane; q < pd; q += 32) {  // q in image order (c,py,px): 32-byte runs
        const int c = q / pp, py = (q % pp) / patch, px = q % patch;
        const float v = src[(size_t(c) * img + pr * patch + py) * img + pcn * patch + px];
        st[(py * patch + px) * C + c] = v;
        s += v; ss += v * v;
      }
    }
    __syncwarp();
    float mu = 0.f, rsd = 1.f;
    if (norm_pix) {  // opt-in MAE-paper variant: (t - mean) / sqrt(var_unbiased + 1e-6)
      s = warp_sum(s); ss = warp_sum(ss);
      mu = s / pd;
      const float var = (ss - pd * mu * mu) / (pd - 1);
      rsd = rsqrtf(var + 1e-6f);
    }
    const __nv_bfloat16* pr_row = pred + r * pd;
    for (int e = lane * 2; e < pd; e += 64) {  // pd is even: bf16x2 per lane, coalesced
      const uint32_t pv = *reinterpret_cast<const uint32_t*>(pr_row + e);
      const float t0 = (st[e] - mu) * rsd, t1 = (st[e + 1] - mu) * rsd;
      const float d0 = bf16_lo(pv) - t0, d1 = bf16_hi(pv) - t1;
      acc += d0 * d0 + d1 * d1;
      if (dpred) *reinterpret_cast<uint32_t*>(dpred + r * pd + e) = pack_bf16x2(2.0f * d0 * inv_numel, 2.0f * d1 * inv_numel);
      if (target) *reinterpret_cast<float2*>(target + r * pd + e) = make_float2(t0, t1);
    }
    __syncwarp();
  }
  acc = warp_sum(acc);
  if (lane == 0) s_red[wid] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < wpb; ++i) t += s_red[i];
    partials[blockIdx.x] = t;
  }
}

// Fast path for 8x8 RGB patches (configs/mae.yaml): lane l owns pixels (py = l/4, px = 2*(l%4) + {0,1}) of the patch in
// all three channels, i.e. six consecutive elements of the (py,px,c)-ordered target vector: three 8-byte image loads
// (each 32-byte patch row is covered by 4 adjacent lanes = one sector), three 4-byte pred loads / dpred stores
// (384 contiguous bytes per warp) and no shared-memory transpose.  Two patches per warp iteration, loads first.
__global__ void __launch_bounds__(256) masked_mse_p8c3_kernel(const __nv_bfloat16* __restrict__ pred,
                                                             const float* __restrict__ images,
                                                             const int64_t* __restrict__ idx_mask, int B, int img,
                                                             int n_mask, int norm_pix, float inv_numel,
                                                             float* __restrict__ partials,
                                                             __nv_bfloat16* __restrict__ dpred,
                                                             float* __restrict__ target) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_red[8];
  constexpr int ROWS = 2, PD = 192;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int grid = img >> 3;
  const size_t plane = size_t(img) * img;
  const int py = lane >> 2, px = (lane & 3) * 2;
  float acc = 0.f;
  const long long nrows = (long long)B * n_mask;
  const long long stride = (long long)gridDim.x * wpb;
  for (long long ra = (long long)blockIdx.x * wpb + wid; ra < nrows; ra += ROWS * stride) {
    float2 ch[ROWS][3];
    uint32_t pv[ROWS][3];
    long long rr[ROWS];
    bool ok[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      rr[k] = ra + k * stride;
      ok[k] = rr[k] < nrows;
      if (!ok[k]) rr[k] = ra;
      const int b = int(rr[k] / n_mask);
      int pidx = int(idx_mask[rr[k]]) - 1;
      if (pidx < 0) pidx = 0;  // torch.clamp(idx_mask - 1, min=0)
      const int pr = pidx / grid, pcn = pidx % grid;
      const float* src = images + size_t(b) * 3 * plane + size_t(pr * 8 + py) * img + pcn * 8 + px;
#pragma unroll
      for (int c = 0; c < 3; ++c) ch[k][c] = *reinterpret_cast<const float2*>(src + c * plane);
      const uint32_t* pp = reinterpret_cast<const uint32_t*>(pred + rr[k] * PD) + lane * 3;
#pragma unroll
      for (int j = 0; j < 3; ++j) pv[k][j] = pp[j];
    }
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      float t[6] = {ch[k][0].x, ch[k][1].x, ch[k][2].x, ch[k][0].y, ch[k][1].y, ch[k][2].y};
      if (norm_pix) {  // opt-in MAE-paper variant: (t - mean) / sqrt(var_unbiased + 1e-6)
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int e = 0; e < 6; ++e) { s += t[e]; ss += t[e] * t[e]; }
        s = warp_sum(s); ss = warp_sum(ss);
        const float mu = s / PD;
        const float rsd = rsqrtf((ss - PD * mu * mu) / (PD - 1) + 1e-6f);
#pragma unroll
        for (int e = 0; e < 6; ++e) t[e] = (t[e] - mu) * rsd;
      }
      if (!ok[k]) continue;
      float d[6];
#pragma unroll
      for (int j = 0; j < 3; ++j) { d[2 * j] = bf16_lo(pv[k][j]) - t[2 * j]; d[2 * j + 1] = bf16_hi(pv[k][j]) - t[2 * j + 1]; }
#pragma unroll
      for (int e = 0; e < 6; ++e) acc += d[e] * d[e];
      if (dpred) {
        uint32_t* dp = reinterpret_cast<uint32_t*>(dpred + rr[k] * PD) + lane * 3;
#pragma unroll
        for (int j = 0; j < 3; ++j) dp[j] = pack_bf16x2(2.0f * d[2 * j] * inv_numel, 2.0f * d[2 * j + 1] * inv_numel);
      }
      if (target) {
        float2* tp = reinterpret_cast<float2*>(target + rr[k] * PD) + lane * 3;
#pragma unroll
        for (int j = 0; j < 3; ++j) tp[j] = make_float2(t[2 * j], t[2 * j + 1]);
      }
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) s_red[wid] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < wpb; ++i) t += s_red[i];
    partials[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) finalize_loss_kernel(const float* __restrict__ partials, int n, float scale,
                                                           float* __restrict__ loss) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s[256];
  float t = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) t += partials[i];
  s[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = s[0] * scale;
}

// ---------------------------------------------------------------------------------------------
// column sums and row gathers
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, int M, int N,
                                                         int rows_per_block, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int col = (blockIdx.y * 128 + threadIdx.x) * 2;
  if (col >= N) return;
  const int m0 = blockIdx.x * rows_per_block;
  int m1 = m0 + rows_per_block;
  if (m1 > M) m1 = M;
  float a = 0.f, b = 0.f;
  for (int m = m0; m < m1; ++m) {
    const uint32_t v = *reinterpret_cast<const uint32_t*>(x + size_t(m) * N + col);
    a += bf16_lo(v); b += bf16_hi(v);
  }
  atomicAdd(out + col, a);
  atomicAdd(out + col + 1, b);
}

__device__ __forceinline__ long long sel_row(long long r, const int64_t* idx, int n_sel, int L_src, int row_offset) {
  const uint32_t b = uint32_t(r) / uint32_t(n_sel), j = uint32_t(r) - b * uint32_t(n_sel);  // B * n_sel is an int
  return (long long)b * L_src + (idx ? __ldg(idx + r) : (long long)(row_offset + j));
}

__global__ void __launch_bounds__(256) gather_rows_bf16_kernel(const float* __restrict__ src,
                                                              const int64_t* __restrict__ idx, int B, int n_sel,
                                                              int L_src, int row_offset, int D,
                                                              __nv_bfloat16* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int D4 = D >> 2;
  const long long n = (long long)B * n_sel * D4, stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  RowWalk w(i0, stride, D4, n_sel);
  for (long long i = i0; i < n; i += stride, w.next()) {
    const int d = w.d4 * 4;
    const long long r = (long long)w.b * n_sel + w.j;
    const long long srow = (long long)w.b * L_src + (idx ? idx[r] : (long long)(row_offset + w.j));
    const float4 v = *reinterpret_cast<const float4*>(src + srow * D + d);
    *reinterpret_cast<uint2*>(out + r * D + d) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

// dst rows b*L_src + idx[b*n_sel+j] <- 0 (fp32 and optional bf16 copy): the kept rows of the decoder stream receive no
// gradient from decoder_norm / decoder_pred (reference src/models/mae.py:72: only idx_mask rows are predicted).
__global__ void __launch_bounds__(256) zero_rows_kernel(float* __restrict__ dst, __nv_bfloat16* __restrict__ dst_bf16,
                                                       const int64_t* __restrict__ idx, int B, int n_sel, int L_src, int D) {
  pdl_trigger();
  pdl_wait();
  const int D4 = D >> 2;
  const long long n = (long long)B * n_sel * D4, stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  RowWalk w(i0, stride, D4, n_sel);
  for (long long i = i0; i < n; i += stride, w.next()) {
    const int d = w.d4 * 4;
    const long long r = (long long)w.b * n_sel + w.j;
    const long long row = (long long)w.b * L_src + idx[r];
    *reinterpret_cast<float4*>(dst + row * D + d) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dst_bf16) *reinterpret_cast<uint2*>(dst_bf16 + row * D + d) = make_uint2(0u, 0u);
  }
}

template <int NV>
__global__ void __launch_bounds__(256) colsum_rows_f32_kernel(const float* __restrict__ src,
                                                             const int64_t* __restrict__ idx, int B, int n_sel,
                                                             int L_src, int row_offset, int D,
                                                             float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  // one warp per row (grid-stride), lane owns the float4 column groups {128*j + 4*lane}; block-level smem reduce
  __shared__ float s_acc[128 * NV];
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 128 * NV; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  float4 a[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) a[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long nrows = (long long)B * n_sel, stride = (long long)gridDim.x * wpb;
  constexpr int RB = 4;  // rows in flight per warp: the index loads, then the row loads, are issued back to back
  for (long long r = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); r < nrows; r += RB * stride) {
    // rows past the end are clamped to a valid row and weighted 0, so that every load is unconditional and the
    // compiler keeps all RB index loads, then all RB row loads, in flight together
    long long rows_[RB];
    float wt[RB];
#pragma unroll
    for (int k = 0; k < RB; ++k) {
      const long long rk = r + k * stride;
      wt[k] = rk < nrows ? 1.f : 0.f;
      rows_[k] = sel_row(rk < nrows ? rk : r, idx, n_sel, L_src, row_offset);
    }
    __syncwarp();  // scheduling fence: without it nvcc re-serialises index load -> row load -> add per row
    float4 v[RB][NV];
#pragma unroll
    for (int k = 0; k < RB; ++k)
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = 128 * j + 4 * lane;
        v[k][j] = *reinterpret_cast<const float4*>(src + rows_[k] * D + (c < D ? c : 0));
      }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < RB; ++k)
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        a[j].x += wt[k] * v[k][j].x; a[j].y += wt[k] * v[k][j].y; a[j].z += wt[k] * v[k][j].z; a[j].w += wt[k] * v[k][j].w;
      }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = 128 * j + 4 * lane;
    if (c < D) { atomicAdd(&s_acc[c], a[j].x); atomicAdd(&s_acc[c + 1], a[j].y); atomicAdd(&s_acc[c + 2], a[j].z); atomicAdd(&s_acc[c + 3], a[j].w); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) atomicAdd(out + i, s_acc[i]);
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                       long long n) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i]);
}

__global__ void __launch_bounds__(256) transpose_cast_kernel(const float* __restrict__ src,
                                                            __nv_bfloat16* __restrict__ dst,
                                                            const long long* __restrict__ table) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[32][33];
  const long long* t = table + size_t(blockIdx.z) * 4;
  const long long src_off = t[0], dst_off = t[2 + 1];
  const int rows = int(t[1]), cols = int(t[2]);
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  if (r0 >= rows || c0 >= cols) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? src[src_off + (long long)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < rows && c < cols) dst[dst_off + (long long)c * rows + r] = __float2bfloat16_rn(tile[tx][i]);
  }
}

// ---------------------------------------------------------------------------------------------
// clip-by-global-norm + AdamW over one flat buffer
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, long long n,
                                                        float* __restrict__ partials) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s[8];
  float t = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = g[i];
    t += v * v;
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

__global__ void __launch_bounds__(256) clip_adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v,
                                                        __nv_bfloat16* __restrict__ p_bf16, long long n, float lr,
                                                        float beta1, float beta2, float eps, float wd, float max_norm,
                                                        float inv_bc1, float inv_sqrt_bc2,
                                                        const float* __restrict__ partials, int n_partials,
                                                        float* __restrict__ grad_norm_out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s[256];
  // every block reduces the same partials in the same order -> identical, deterministic norm
  float t = 0.f;
  for (int i = threadIdx.x; i < n_partials; i += 256) t += partials[i];
  s[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  const float norm = sqrtf(s[0]);
  const float coef = fminf(1.0f, max_norm / (norm + 1e-6f));
  if (blockIdx.x == 0 && threadIdx.x == 0 && grad_norm_out) grad_norm_out[0] = norm;
  const float decay = 1.0f - lr * wd;
  const float step_size = lr * inv_bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    float pi = p[i] * decay;
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    pi -= step_size * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (p_bf16) p_bf16[i] = __float2bfloat16_rn(pi);
  }
}

static int grid_for(long long n, int threads, int cap = 148 * 16) {
  long long g = (n + threads - 1) / threads;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return int(g);
}

}  // namespace mae

using namespace mae;

extern "C" int mae_version(void) { return 100; }

extern "C" const char* mae_error_string(int code) {
  switch (code) {
    case MAE_OK: return "ok";
    case MAE_ERR_SHAPE: return "unsupported or inconsistent shape";
    case MAE_ERR_NULL: return "required pointer is null";
    case MAE_ERR_ALIGN: return "pointer or leading dimension not 16-byte aligned";
    case MAE_ERR_DRIVER: return "CUDA driver entry point / tensor map encode failed";
    case MAE_ERR_LAUNCH: return "kernel launch failed";
    default: return "unknown error";
  }
}

extern "C" int mae_random_token_mask(const float* noise, int B, int L, int num_keep, int64_t* idx_keep,
                                     int64_t* idx_mask, cudaStream_t stream) {
  if (B <= 0) return MAE_OK;
  if (!noise || !idx_keep || (!idx_mask && num_keep < L)) return MAE_ERR_NULL;
  if (L < 1 || L > 256 || num_keep < 1 || num_keep > L) return MAE_ERR_SHAPE;
  mae_launch(token_mask_kernel, dim3((B + 3) / 4), dim3(128), 0, stream, noise, B, L, num_keep, idx_keep, idx_mask);
  return MAE_CHECK_LAUNCH();
}

extern "C" int mae_gather_patches(const float* images, const int64_t* idx_keep, int B, int C, int img, int patch,
                                  int num_keep, void* patches, cudaStream_t stream) {
  if (B <= 0 || num_keep <= 1) return MAE_OK;
  if (!images || !idx_keep || !patches) return MAE_ERR_NULL;
  if (patch <= 0 || img % patch) return MAE_ERR_SHAPE;
  const long long warps = (long long)B * (num_keep - 1);
  mae_launch(gather_patches_kernel, dim3(unsigned((warps + 7) / 8)), dim3(256), 0, stream, images, idx_keep, B, C, img, patch, num_keep,
                                                                       reinterpret_cast<__nv_bfloat16*>(patches));
  return MAE_CHECK_LAUNCH();
}

extern "C" int mae_assemble_tokens(const void* tok, const float* cls, const float* pos, const int64_t* idx_keep, int B,
                                   int num_keep, int D, float* x, cudaStream_t stream) {
  if (B <= 0) return MAE_OK;
  if (!tok || !cls || !pos || !idx_keep || !x) return MAE_ERR_NULL;
  if (D % 4) return MAE_ERR_SHAPE;
  mae_launch(assemble_tokens_kernel, dim3(grid_for((long long)B * num_keep * (D / 4), 256)), dim3(256), 0, stream, reinterpret_cast<const __nv_bfloat16*>(tok), cls, pos, idx_keep, B, num_keep, D, x);
  return MAE_CHECK_LAUNCH();
}

// A/B switch: MAE_LN_LPR=32 keeps one warp per row for D <= 192
static bool ln_half_warp_rows() {
  static const bool v = !(getenv("MAE_LN_LPR") && atoi(getenv("MAE_LN_LPR")) == 32);
  return v;
}

extern "C" int mae_layernorm_fwd(const float* x, const float* gamma, const float* beta, const int64_t* gather_idx,
                                 int n_sel, int L_src, int rows, int D, float eps, void* xn, float* mean, float* rstd,
                                 cudaStream_t stream) {
  if (rows <= 0) return MAE_OK;
  if (!x || !gamma || !beta || !xn || !mean || !rstd) return MAE_ERR_NULL;
  if (D > 1024 || (D % 4) || (gather_idx && (n_sel <= 0 || L_src <= 0))) return MAE_ERR_SHAPE;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(xn);
  static const int bps = getenv("MAE_LN_FWD_BPS") ? atoi(getenv("MAE_LN_FWD_BPS")) : 16;
  const dim3 grid(grid_for(rows, 8, 148 * bps)), block(256);
  if (D <= 192 && ln_half_warp_rows()) {  // 16 lanes per row, two rows per warp (see the kernel comment)
    const dim3 grid2(grid_for(rows, 16, 148 * bps));
    if (gather_idx) mae_launch(layernorm_fwd_kernel<3, 2, true, false, 16>, grid2, block, 0, stream, x, gamma, beta, gather_idx, n_sel, L_src, rows, D, eps, o, mean, rstd, (const __nv_bfloat16*)nullptr, (float*)nullptr);
    else mae_launch(layernorm_fwd_kernel<3, 2, false, false, 16>, grid2, block, 0, stream, x, gamma, beta, gather_idx, n_sel, L_src, rows, D, eps, o, mean, rstd, (const __nv_bfloat16*)nullptr, (float*)nullptr);
  } else if (D <= 256) {
    if (gather_idx) mae_launch(layernorm_fwd_kernel<2, 2, true>, grid, block, 0, stream, x, gamma, beta, gather_idx, n_sel, L_src, rows, D, eps, o, mean, rstd, (const __nv_bfloat16*)nullptr, (float*)nullptr);
    else mae_launch(layernorm_fwd_kernel<2, 2, false>, grid, block, 0, stream, x, gamma, beta, gather_idx, n_sel, L_src, rows, D, eps, o, mean, rstd, (const __nv_bfloat16*)nullptr, (float*)nullptr);
  } else if (D <= 512) {
    if (gather_idx) mae_launch(layernorm_fwd_kernel<4, 2, true>, grid, block, 0, stream, x, gamma, beta, gather_idx, n_sel, L_src, rows, D, eps, o, mean, rstd, (const __nv_bfloat16*)nullptr, (float*)nullptr);
    else mae_launch(layernorm_fwd_kernel<4, 2, false>, grid, block, 0, stream, x, gamma, beta, gather_idx, n_sel, L_src, rows, D, eps, o, mean, rstd, (const __nv_bfloat16*)nullptr, (float*)nullptr);
  } else {
    if (gather_idx) mae_launch(layernorm_fwd_kernel<8, 1, true>, grid, block, 0, stream, x, gamma, beta, gather_idx, n_sel, L_src, rows, D, eps, o, mean, rstd, (const __nv_bfloat16*)nullptr, (float*)nullptr);
    else mae_launch(layernorm_fwd_kernel<8, 1, false>, grid, block, 0, stream, x, gamma, beta, gather_idx, n_sel, L_src, rows, D, eps, o, mean, rstd, (const __nv_bfloat16*)nullptr, (float*)nullptr);
  }
  return MAE_CHECK_LAUNCH();
}

extern "C" int mae_layernorm_fwd_add(const float* x, const void* delta, const float* gamma, const float* beta,
                                     const int64_t* gather_idx, int n_sel, int L_src, int rows, int D, float eps,
                                     float* x_out, void* xn, float* mean, float* rstd, cudaStream_t stream) {
  if (rows <= 0) return MAE_OK;
  if (!x || !delta || !gamma || !beta || !x_out || !xn || !mean || !rstd) return MAE_ERR_NULL;
  if (D > 256 || (D % 4) || (gather_idx && (n_sel <= 0 || L_src <= 0))) return MAE_ERR_SHAPE;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(xn);
  const __nv_bfloat16* dl = reinterpret_cast<const __nv_bfloat16*>(delta);
  const dim3 grid(grid_for(rows, 8, 148 * 16)), block(256);
  if (D <= 192 && ln_half_warp_rows()) {
    const dim3 grid2(grid_for(rows, 16, 148 * 16));
    if (gather_idx) mae_launch(layernorm_fwd_kernel<3, 2, true, true, 16>, grid2, block, 0, stream, x, gamma, beta, gather_idx, n_sel, L_src, rows, D, eps, o, mean, rstd, dl, x_out);
    else mae_launch(layernorm_fwd_kernel<3, 2, false, true, 16>, grid2, block, 0, stream, x, gamma, beta, gather_idx, n_sel, L_src, rows, D, eps, o, mean, rstd, dl, x_out);
  } else if (gather_idx) mae_launch(layernorm_fwd_kernel<2, 2, true, true>, grid, block, 0, stream, x, gamma, beta, gather_idx, n_sel, L_src, rows, D, eps, o, mean, rstd, dl, x_out);
  else mae_launch(layernorm_fwd_kernel<2, 2, false, true>, grid, block, 0, stream, x, gamma, beta, gather_idx, n_sel, L_src, rows, D, eps, o, mean, rstd, dl, x_out);
  return MAE_CHECK_LAUNCH();
}

extern "C" int mae_layernorm_bwd_scatter(const void* g, const float* x, const float* mean, const float* rstd,
                                         const float* gamma, const float* dx_in_compact, const int64_t* gather_idx,
                                         int n_sel, int L_src, int rows, int D, float* dx_out, void* dx_out_bf16,
                                         float* dgamma, float* dbeta, cudaStream_t stream) {
  if (rows <= 0) return MAE_OK;
  if (!g || !x || !mean || !rstd || !gamma || !dx_in_compact || !gather_idx || !dx_out || !dgamma || !dbeta) return MAE_ERR_NULL;
  if (D > 256 || (D % 4) || n_sel <= 0 || L_src <= 0) return MAE_ERR_SHAPE;
  const __nv_bfloat16* gp = reinterpret_cast<const __nv_bfloat16*>(g);
  __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(dx_out_bf16);
  const dim3 grid(grid_for(rows, 8, 148 * 6)), block(256);
  if (D <= 192 && ln_half_warp_rows())
    mae_launch(layernorm_bwd_kernel<3, 1, 2, true, true, 16>, dim3(grid_for(rows, 16, 148 * 2)), block, 0, stream, gp, x, mean, rstd, gamma, dx_in_compact, gather_idx, n_sel, L_src, rows, D, dx_out, ob, dgamma, dbeta);
  else
    mae_launch(layernorm_bwd_kernel<2, 2, 3, true, true>, grid, block, 0, stream, gp, x, mean, rstd, gamma, dx_in_compact, gather_idx, n_sel, L_src, rows, D, dx_out, ob, dgamma, dbeta);
  return MAE_CHECK_LAUNCH();
}

extern "C" int mae_layernorm_bwd(const void* g, const float* x, const float* mean, const float* rstd,
                                 const float* gamma, const float* dx_in, const int64_t* gather_idx, int n_sel,
                                 int L_src, int rows, int D, float* dx_out, void* dx_out_bf16, float* dgamma,
                                 float* dbeta, cudaStream_t stream) {
  if (rows <= 0) return MAE_OK;
  if (!g || !x || !mean || !rstd || !gamma || !dx_out || !dgamma || !dbeta) return MAE_ERR_NULL;
  if (D > 1024 || (D % 4) || (gather_idx && (n_sel <= 0 || L_src <= 0))) return MAE_ERR_SHAPE;
  const __nv_bfloat16* gp = reinterpret_cast<const __nv_bfloat16*>(g);
  __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(dx_out_bf16);
  static const int bps_env = getenv("MAE_LN_BWD_BPS") ? atoi(getenv("MAE_LN_BWD_BPS")) : 0;
  const int bps = bps_env ? bps_env : 6;  // 3 resident CTAs per SM: two full waves
  const dim3 grid(grid_for(rows, 8, 148 * bps)), block(256);
  if (D <= 192 && ln_half_warp_rows()) {
    // 2 resident CTAs per SM, one wave: measured best (D = 192: 137 us = 0.99 of the copy bandwidth; more CTAs only add
    // d-gamma / d-beta atomics on the same 2 D addresses)
    const dim3 grid2(grid_for(rows, 16, 148 * (bps_env ? bps_env : 2)));
    if (gather_idx) mae_launch(layernorm_bwd_kernel<3, 1, 2, true, false, 16>, grid2, block, 0, stream, gp, x, mean, rstd, gamma, dx_in, gather_idx, n_sel, L_src, rows, D, dx_out, ob, dgamma, dbeta);
    else mae_launch(layernorm_bwd_kernel<3, 1, 2, false, false, 16>, grid2, block, 0, stream, gp, x, mean, rstd, gamma, dx_in, gather_idx, n_sel, L_src, rows, D, dx_out, ob, dgamma, dbeta);
  } else if (D <= 256) {
    if (gather_idx) mae_launch(layernorm_bwd_kernel<2, 2, 3, true>, grid, block, 0, stream, gp, x, mean, rstd, gamma, dx_in, gather_idx, n_sel, L_src, rows, D, dx_out, ob, dgamma, dbeta);
    else mae_launch(layernorm_bwd_kernel<2, 2, 3, false>, grid, block, 0, stream, gp, x, mean, rstd, gamma, dx_in, gather_idx, n_sel, L_src, rows, D, dx_out, ob, dgamma, dbeta);
  } else if (D <= 512) {
    if (gather_idx) mae_launch(layernorm_bwd_kernel<4, 1, 2, true>, grid, block, 0, stream, gp, x, mean, rstd, gamma, dx_in, gather_idx, n_sel, L_src, rows, D, dx_out, ob, dgamma, dbeta);
    else mae_launch(layernorm_bwd_kernel<4, 1, 2, false>, grid, block, 0, stream, gp, x, mean, rstd, gamma, dx_in, gather_idx, n_sel, L_src, rows, D, dx_out, ob, dgamma, dbeta);
  } else {
    if (gather_idx) mae_launch(layernorm_bwd_kernel<8, 1, 1, true>, grid, block, 0, stream, gp, x, mean, rstd, gamma, dx_in, gather_idx, n_sel, L_src, rows, D, dx_out, ob, dgamma, dbeta);
    else mae_launch(layernorm_bwd_kernel<8, 1, 1, false>, grid, block, 0, stream, gp, x, mean, rstd, gamma, dx_in, gather_idx, n_sel, L_src, rows, D, dx_out, ob, dgamma, dbeta);
  }
  return MAE_CHECK_LAUNCH();
}

extern "C" int mae_scatter_decoder_tokens(const void* emb, const float* mask_token, const float* dec_pos,
                                          const int64_t* idx_keep, const int64_t* idx_mask, int B, int L, int num_keep,
                                          int D, float* x_dec, cudaStream_t stream) {
  if (B <= 0) return MAE_OK;
  if (!emb || !mask_token || !dec_pos || !idx_keep || !idx_mask || !x_dec) return MAE_ERR_NULL;
  if (D % 4 || num_keep < 1 || num_keep > L) return MAE_ERR_SHAPE;
  mae_launch(scatter_decoder_kernel, dim3(grid_for((long long)B * L * (D / 4), 256)), dim3(256), 0, stream, reinterpret_cast<const __nv_bfloat16*>(emb), mask_token, dec_pos, idx_keep, idx_mask, B, L, num_keep, D, x_dec, 0);
  return MAE_CHECK_LAUNCH();
}

extern "C" int mae_scatter_decoder_tokens_grouped(const void* emb, const float* mask_token, const float* dec_pos,
                                                  const int64_t* idx_keep, const int64_t* idx_mask, int B, int L,
                                                  int num_keep, int D, float* x_dec, cudaStream_t stream) {
  if (B <= 0) return MAE_OK;
  if (!emb || !mask_token || !dec_pos || !idx_keep || !idx_mask || !x_dec) return MAE_ERR_NULL;
  if (D % 4 || num_keep < 1 || num_keep > L) return MAE_ERR_SHAPE;
  mae_launch(scatter_decoder_kernel, dim3(grid_for((long long)B * L * (D / 4), 256)), dim3(256), 0, stream, reinterpret_cast<const __nv_bfloat16*>(emb), mask_token, dec_pos, idx_keep, idx_mask, B, L, num_keep, D, x_dec, 1);
  return MAE_CHECK_LAUNCH();
}

extern "C" int mae_scatter_decoder_tokens_ln(const void* emb, const float* mask_token, const float* dec_pos,
                                             const int64_t* idx_keep, const int64_t* idx_mask, int B, int L, int num_keep,
                                             int D, const float* gamma, const float* beta, float eps, float* x_dec,
                                             void* xn, float* mean, float* rstd, cudaStream_t stream) {
  if (B <= 0) return MAE_OK;
  if (!emb || !mask_token || !dec_pos || !idx_keep || !idx_mask || !gamma || !beta || !x_dec || !xn || !mean || !rstd)
    return MAE_ERR_NULL;
  if (D % 4 || D > 256 || num_keep < 1 || num_keep > L) return MAE_ERR_SHAPE;
  mae_launch(scatter_decoder_ln_kernel<2>, dim3(grid_for((long long)B * L, 8, 148 * 16)), dim3(256), 0, stream,
             reinterpret_cast<const __nv_bfloat16*>(emb), mask_token, dec_pos, idx_keep, idx_mask, B, L, num_keep, D, gamma,
             beta, eps, x_dec, reinterpret_cast<__nv_bfloat16*>(xn), mean, rstd);
  return MAE_CHECK_LAUNCH();
}

extern "C" int mae_masked_mse(const void* pred, const float* images, const int64_t* idx_mask, int B, int C, int img,
                              int patch, int n_mask, int norm_pix, float* loss, float* partials, void* dpred,
                              float* target, cudaStream_t stream) {
  if (!pred || !images || !idx_mask || !loss || !partials) return MAE_ERR_NULL;
  if (B <= 0 || n_mask <= 0) return MAE_ERR_SHAPE;
  const int pd = C * patch * patch;
  if (patch <= 0 || img % patch || pd > kMaxPatchDim) return MAE_ERR_SHAPE;
  const long long nrows = (long long)B * n_mask;
  int grid = int((nrows + 7) / 8);
  if (grid > kMsePartials) grid = kMsePartials;
  const float inv_numel = 1.0f / (float(nrows) * float(pd));
  if (patch == 8 && C == 3 && (img % 8) == 0) {
    const int g2 = int((nrows + 15) / 16) < grid ? int((nrows + 15) / 16) : grid;  // two patches per warp iteration
    grid = g2 < 1 ? 1 : g2;
    mae_launch(masked_mse_p8c3_kernel, dim3(grid), dim3(256), 0, stream, reinterpret_cast<const __nv_bfloat16*>(pred), images,
               idx_mask, B, img, n_mask, norm_pix, inv_numel, partials, reinterpret_cast<__nv_bfloat16*>(dpred), target);
  } else {
    mae_launch(masked_mse_kernel, dim3(grid), dim3(256), 8 * pd * sizeof(float), stream, reinterpret_cast<const __nv_bfloat16*>(pred), images, idx_mask, B, C, img, patch, n_mask, norm_pix, inv_numel,
        partials, reinterpret_cast<__nv_bfloat16*>(dpred), target);
  }
  if (cudaGetLastError() != cudaSuccess) return MAE_ERR_LAUNCH;
  mae_launch(finalize_loss_kernel, dim3(1), dim3(256), 0, stream, partials, grid, inv_numel, loss);
  return MAE_CHECK_LAUNCH();
}

extern "C" int mae_colsum_bf16(const void* x, int M, int N, float* out, cudaStream_t stream) {
  if (M <= 0) return MAE_OK;
  if (!x || !out) return MAE_ERR_NULL;
  if (N % 2) return MAE_ERR_SHAPE;
  const int ny = (N + 255) / 256;
  int nx = (148 * 8) / ny;
  int rpb = (M + nx - 1) / nx;
  if (rpb < 16) rpb = 16;
  nx = (M + rpb - 1) / rpb;
  mae_launch(colsum_bf16_kernel, dim3(dim3(nx, ny)), dim3(128), 0, stream, reinterpret_cast<const __nv_bfloat16*>(x), M, N, rpb, out);
  return MAE_CHECK_LAUNCH();
}

extern "C" int mae_gather_rows_bf16(const float* src, const int64_t* idx, int B, int n_sel, int L_src, int row_offset,
                                    int D, void* out, cudaStream_t stream) {
  if (B <= 0 || n_sel <= 0) return MAE_OK;
  if (!src || !out) return MAE_ERR_NULL;
  if (D % 4) return MAE_ERR_SHAPE;
  mae_launch(gather_rows_bf16_kernel, dim3(grid_for((long long)B * n_sel * (D / 4), 256)), dim3(256), 0, stream, src, idx, B, n_sel, L_src, row_offset, D, reinterpret_cast<__nv_bfloat16*>(out));
  return MAE_CHECK_LAUNCH();
}

extern "C" int mae_zero_rows(float* dst, void* dst_bf16, const int64_t* idx, int B, int n_sel, int L_src, int D,
                             cudaStream_t stream) {
  if (B <= 0 || n_sel <= 0) return MAE_OK;
  if (!dst || !idx) return MAE_ERR_NULL;
  if (D % 4 || L_src <= 0) return MAE_ERR_SHAPE;
  mae_launch(zero_rows_kernel, dim3(grid_for((long long)B * n_sel * (D / 4), 256)), dim3(256), 0, stream, dst,
             reinterpret_cast<__nv_bfloat16*>(dst_bf16), idx, B, n_sel, L_src, D);
  return MAE_CHECK_LAUNCH();
}

extern "C" int mae_colsum_rows_f32(const float* src, const int64_t* idx, int B, int n_sel, int L_src, int row_offset,
                                   int D, float* out, cudaStream_t stream) {
  if (B <= 0 || n_sel <= 0) return MAE_OK;
  if (!src || !out) return MAE_ERR_NULL;
  if (D > 1024 || (D % 4)) return MAE_ERR_SHAPE;
  const long long nrows = (long long)B * n_sel;
  const dim3 grid(grid_for(nrows, 8 * 4, 148 * 8)), block(256);
  if (D <= 256) mae_launch(colsum_rows_f32_kernel<2>, grid, block, 0, stream, src, idx, B, n_sel, L_src, row_offset, D, out);
  else if (D <= 512) mae_launch(colsum_rows_f32_kernel<4>, grid, block, 0, stream, src, idx, B, n_sel, L_src, row_offset, D, out);
  else mae_launch(colsum_rows_f32_kernel<8>, grid, block, 0, stream, src, idx, B, n_sel, L_src, row_offset, D, out);
  return MAE_CHECK_LAUNCH();
}

extern "C" int mae_cast_bf16(const float* src, void* dst, long long n, cudaStream_t stream) {
  if (n <= 0) return MAE_OK;
  if (!src || !dst) return MAE_ERR_NULL;
  mae_launch(cast_bf16_kernel, dim3(grid_for(n, 256)), dim3(256), 0, stream, src, reinterpret_cast<__nv_bfloat16*>(dst), n);
  return MAE_CHECK_LAUNCH();
}

extern "C" int mae_transpose_cast_bf16(const float* src, void* dst, const long long* table, int n_entries,
                                       int max_rows, int max_cols, cudaStream_t stream) {
  if (n_entries <= 0) return MAE_OK;
  if (!src || !dst || !table) return MAE_ERR_NULL;
  if (max_rows <= 0 || max_cols <= 0) return MAE_ERR_SHAPE;
  // grid covers the largest weight of the table; blocks outside a smaller entry exit early
  mae_launch(transpose_cast_kernel, dim3((max_cols + 31) / 32, (max_rows + 31) / 32, n_entries), dim3(256), 0, stream, src,
             reinterpret_cast<__nv_bfloat16*>(dst), table);
  return MAE_CHECK_LAUNCH();
}

extern "C" int mae_grad_sumsq(const float* g, long long n, float* partials, int max_partials, cudaStream_t stream) {
  if (!g || !partials) return MAE_ERR_NULL;
  if (n <= 0 || max_partials <= 0) return MAE_ERR_SHAPE;
  mae_launch(grad_sumsq_kernel, dim3(max_partials), dim3(256), 0, stream, g, n, partials);
  return MAE_CHECK_LAUNCH();
}

extern "C" int mae_clip_adamw(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, float max_norm, int step,
                              const float* partials, int n_partials, float* grad_norm_out, cudaStream_t stream) {
  if (!p || !g || !m || !v || !partials) return MAE_ERR_NULL;
  if (n <= 0 || step < 1 || n_partials < 1) return MAE_ERR_SHAPE;
  const double bc1 = 1.0 - pow(double(beta1), double(step));
  const double bc2 = 1.0 - pow(double(beta2), double(step));
  mae_launch(clip_adamw_kernel, dim3(grid_for(n, 256, 148 * 8)), dim3(256), 0, stream, p, g, m, v, reinterpret_cast<__nv_bfloat16*>(p_bf16), n, lr, beta1, beta2, eps, weight_decay, max_norm,
      float(1.0 / bc1), float(1.0 / sqrt(bc2)), partials, n_partials, grad_norm_out);
  return MAE_CHECK_LAUNCH();
}
