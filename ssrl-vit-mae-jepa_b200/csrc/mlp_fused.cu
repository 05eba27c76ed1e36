// Fused MLP half of a pre-norm transformer block (timm Block: x = x + mlp(norm2(x))) for sm_100a, forward, ONE launch:
//
//   xn = LayerNorm(x)  ->  pre = xn W1^T + b1  ->  h = GELU(pre), g = GELU'(pre)  ->  y = x + h W2^T + b2
//
// The hidden activations never leave the SM on their way from fc1 to fc2: a 128-row tile of the stream is normalised
// straight into a swizzled bf16 A tile, fc1 runs in chunks of 64 hidden columns (tcgen05.mma into one of two 64-column
// TMEM buffers), the epilogue warps turn a chunk into GELU / GELU' (bf16 table look-up, exactly rounded) and write both
// into swizzled shared-memory tiles: the GELU tile is the A operand of the chunk's fc2 MMAs (accumulating the 128 x D
// output tile in TMEM over all chunks) and, together with the GELU' tile, leaves through TMA bulk stores because the
// backward needs both.  Compared with LayerNorm + fc1 + fc2 as three launches this removes the xn re-read, the h
// re-read and the residual re-read from HBM (7.7 -> 5.0 KB per row at D = 192) and two launches.
//
// Warp roles (19 warps, one CTA per SM, persistent over row tiles):
//   warp 0      TMA producer of the weight chunks (W1 rows [64c, 64c+64) as ceil(D/64) K-major boxes, W2 columns
//               [64c, 64c+64) as one [D x 64] box) through a 3-stage ring
//   warp 1      TMEM owner + tcgen05.mma issuer: fc1(c+1) is issued before fc2(c), so the tensor pipe works on the next
//               chunk while the epilogue warps are busy with the current one
//   warps 2-17  LayerNorm prologue (8 rows per warp, 8 lanes per row, coalesced fp32 loads, two-pass statistics, bf16 A tile +
//               xn / mean / rstd for the backward), GELU epilogue (even chunks: warps with grp 0-1, odd chunks: grp 2-3; 32
//               columns per warp; four TMEM chunk buffers keep the tensor pipe two chunks ahead of each group), output
//               epilogue (bias + residual, fp32)
//   warp 18     TMA bulk stores of the GELU / GELU' tiles
// Shared memory: [GELU table +][A 48 KB][h 2x16 KB][g 2x16 KB][barriers] ... 128 KB above the + table: [GELU table -][ring].
#include <stdlib.h>

#include "gemm_common.cuh"

namespace mae {

static constexpr int kMlpEpiWarps = 16;
static constexpr int kMlpThreads = (3 + kMlpEpiWarps) * 32;  // 608
static constexpr int kMlpHC = 64;                             // hidden columns per chunk
static constexpr int kMlpStages = 3;
static constexpr uint32_t kMlpStageBytes = 24 * 1024;         // W1 chunk: 3 boxes of [64 x 64]; W2 chunk: [D <= 192 x 64]
static constexpr uint32_t kMlpOffTabP = 0;
static constexpr uint32_t kMlpOffA = 13 * 1024;               // 3 K panels of [128 rows x 128 B]
static constexpr uint32_t kMlpOffH = kMlpOffA + 3 * 16384;    // 2 x [128 x 64] bf16, 128-byte swizzle
static constexpr uint32_t kMlpOffG = kMlpOffH + 2 * 16384;
static constexpr uint32_t kMlpOffHdr = kMlpOffG + 2 * 16384;  // 128000
static constexpr uint32_t kMlpOffTabN = kGeluNegOff;          // 131072: pinned 0x8000 table entries above the + table
static constexpr uint32_t kMlpOffRing = kMlpOffTabN + 13 * 1024;
static constexpr uint32_t kMlpSmem = kMlpOffRing + kMlpStages * kMlpStageBytes;  // 218112
static constexpr int kMlpStgRow = 80;                         // output staging: 64 data bytes + 16 pad per row
static constexpr uint32_t kMlpTmemPre = 256;                  // TMEM columns: [0, D) output tile, 256 + 64 i: four fc1 chunk buffers

struct MlpHdr {
  uint64_t w_full[kMlpStages], w_empty[kMlpStages];
  uint64_t a_ready;
  uint64_t pre_full[4], pre_free[4];
  uint64_t h_ready[2], h_free[2];
  uint64_t out_full, out_free;
  uint32_t tmem_base;
  int rowmap[BM];  // source row of every tile row (-1: past the end)
};
static_assert(sizeof(MlpHdr) <= kMlpOffTabN - kMlpOffHdr, "header must fit below the negative GELU table");
static_assert(kGeluTabBytes <= 13 * 1024, "GELU table region");
static_assert(kMlpEpiWarps * 32 * kMlpStgRow <= 3 * 16384, "output staging lives in the A region");

struct MlpFwdParams {
  int rows, D, HID, NC, nkb, ntiles;
  const float* x;
  const int64_t* gather_idx;
  int n_sel, L_src;
  const float *gamma, *beta;
  float eps;
  const float *b1, *b2;
  __nv_bfloat16* xn;
  float *mean, *rstd;
  float* x_out;
  const uint32_t* gelu_tab;
  long long* dbg;  // optional clock64 timeline of CTA 0 (MAE_MLP_TIMELINE, tools/mlp_timeline.py); nullptr in production
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// timeline slot layout: [role][event] with 64 events per role; role 0 = epilogue warp 2 (even chunks), 1 = epilogue warp 10
// (odd chunks), 2 = MMA issuer
#define MLP_TS(role, ev) do { if (p.dbg != nullptr && blockIdx.x == 0 && lane == 0 && (ev) < 64) p.dbg[(role) * 64 + (ev)] = clock64(); } while (0)

__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_fwd_kernel(const MlpFwdParams p, const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2,
               const __grid_constant__ CUtensorMap mapH, const __grid_constant__ CUtensorMap mapG) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  MlpHdr* hdr = reinterpret_cast<MlpHdr*>(smem + kMlpOffHdr);
  const int warp = warp_id_uniform();  // provably warp-uniform: role branches are uniform control flow
  const int lane = threadIdx.x & 31;
  const int D = p.D, NC = p.NC;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapW1);
    tma_prefetch_desc(&mapW2);
    tma_prefetch_desc(&mapH);
    tma_prefetch_desc(&mapG);
    for (int s = 0; s < kMlpStages; ++s) { mbar_init(&hdr->w_full[s], 1); mbar_init(&hdr->w_empty[s], 1); }
    mbar_init(&hdr->a_ready, kMlpEpiWarps);
    for (int b = 0; b < 4; ++b) {
      mbar_init(&hdr->pre_full[b], 1);
      mbar_init(&hdr->pre_free[b], kMlpEpiWarps / 2);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&hdr->h_ready[b], kMlpEpiWarps / 2);
      mbar_init(&hdr->h_free[b], 2);  // fc2 MMAs of the chunk have completed + its TMA stores have read the tiles
    }
    mbar_init(&hdr->out_full, 1);
    mbar_init(&hdr->out_free, kMlpEpiWarps);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&hdr->tmem_base, 512);
  {  // constant table (not produced by the previous kernel): staged before pdl_wait
    uint32_t* tpos = reinterpret_cast<uint32_t*>(smem + kMlpOffTabP);
    uint32_t* tneg = reinterpret_cast<uint32_t*>(smem + kMlpOffTabN);
    for (uint32_t i = threadIdx.x; i < kGeluNT; i += kMlpThreads) { tpos[i] = p.gelu_tab[i]; tneg[i] = p.gelu_tab[kGeluNT + i]; }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;
  pdl_wait();

  const uint32_t a_s = smem_u32(smem + kMlpOffA), h_s = smem_u32(smem + kMlpOffH), g_s = smem_u32(smem + kMlpOffG);
  const uint32_t ring_s = smem_u32(smem + kMlpOffRing);

  // (TMA / MMA warps: all lanes run the loops, one elected lane issues -- no ELECT / R2UR waterfall per UTCHMMA, see ptx.cuh)
  if (warp == 0) {
    {
      int stage = 0; uint32_t phase = 0;
      const uint32_t w1_bytes = uint32_t(p.nkb) * 8192u, w2_bytes = uint32_t(D) * 128u;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        // same order as the MMA issuer consumes them: W1_0..W1_3, then per chunk c: W2_c, W1_{c+4}
        for (int c = -4; c < NC; ++c) {
          if (c >= 0) {  // W2 columns [64c, +64): one box of [D rows x 64 k]
            mbar_wait(&hdr->w_empty[stage], phase ^ 1);
            uint8_t* dst = smem + kMlpOffRing + stage * kMlpStageBytes;
            if (elect_one_sync()) {
              mbar_expect_tx(&hdr->w_full[stage], w2_bytes);
              tma_load_2d(dst, &mapW2, &hdr->w_full[stage], c * kMlpHC, 0);
            }
            __syncwarp();
            if (++stage == kMlpStages) { stage = 0; phase ^= 1; }
          }
          if (c + 4 < NC) {  // W1 rows [64(c+4), +64): nkb boxes of [64 rows x 64 k]; k >= D is zero-filled
            mbar_wait(&hdr->w_empty[stage], phase ^ 1);
            uint8_t* dst = smem + kMlpOffRing + stage * kMlpStageBytes;
            if (elect_one_sync()) {
              mbar_expect_tx(&hdr->w_full[stage], w1_bytes);
              for (int kb = 0; kb < p.nkb; ++kb) tma_load_2d(dst + kb * 8192, &mapW1, &hdr->w_full[stage], kb * BK, (c + 4) * kMlpHC);
            }
            __syncwarp();
            if (++stage == kMlpStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc1 = umma_idesc_bf16(BM, kMlpHC, 0, 0), idesc2 = umma_idesc_bf16(BM, D, 0, 0);
      const int ks1 = D >> 4;
      int stage = 0; uint32_t phase = 0;
      uint32_t g0 = 0, t = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++t) {
        if (t == 1) MLP_TS(2, 0);
        mbar_wait(&hdr->a_ready, t & 1);
        tc_fence_after();
        if (t == 1) MLP_TS(2, 1);
        // fc1(0..3) up front (four 64-column TMEM buffers), then per chunk c: fc2(c) as soon as the epilogue warps have
        // written its GELU tile, and fc1(c+4) into the buffer chunk c came from: the tensor pipe stays two chunks ahead of
        // each of the two epilogue groups, which therefore never wait for it
        for (int c = -4; c < NC; ++c) {
          if (c >= 0) {  // fc2 chunk c: out += h[hb][128 x 64] * W2[D x 64]^T
            const uint32_t g = g0 + uint32_t(c), hb = g & 1u;
            if (c == 0) mbar_wait(&hdr->out_free, (t & 1u) ^ 1u);  // previous tile's output has been read out of TMEM
            mbar_wait(&hdr->h_ready[hb], (g >> 1) & 1u);
            if (t == 1) MLP_TS(2, 2 + 2 * c);
            mbar_wait(&hdr->w_full[stage], phase);
            tc_fence_after();
            const uint32_t sb = ring_s + stage * kMlpStageBytes;
            if (elect_one_sync()) {
#pragma unroll
              for (int k = 0; k < kMlpHC / 16; ++k) {
                const uint64_t adesc = umma_sdesc_sw128(h_s + hb * 16384 + k * 32, 16, 1024);
                const uint64_t bdesc = umma_sdesc_sw128(sb + k * 32, 16, 1024);
                umma_bf16(tmem_base, adesc, bdesc, idesc2, (c > 0 || k != 0) ? 1u : 0u);
              }
              umma_commit(&hdr->w_empty[stage]);
              umma_commit(&hdr->h_free[hb]);
              if (c == NC - 1) umma_commit(&hdr->out_full);
            }
            __syncwarp();
            if (++stage == kMlpStages) { stage = 0; phase ^= 1; }
          }
          if (c + 4 < NC) {  // fc1 chunk c+4: pre[pb] = A[128 x D] * W1[64 x D]^T
            const uint32_t g = g0 + uint32_t(c + 4), pb = g & 3u;
            mbar_wait(&hdr->pre_free[pb], ((g >> 2) & 1u) ^ 1u);
            mbar_wait(&hdr->w_full[stage], phase);
            tc_fence_after();
            const uint32_t sb = ring_s + stage * kMlpStageBytes;
            const uint32_t d_tmem = tmem_base + kMlpTmemPre + pb * kMlpHC;
            if (elect_one_sync()) {
              for (int ks = 0; ks < ks1; ++ks) {
                const int kb = ks >> 2, k = ks & 3;
                const uint64_t adesc = umma_sdesc_sw128(a_s + kb * 16384 + k * 32, 16, 1024);
                const uint64_t bdesc = umma_sdesc_sw128(sb + kb * 8192 + k * 32, 16, 1024);
                umma_bf16(d_tmem, adesc, bdesc, idesc1, ks != 0);
              }
              umma_commit(&hdr->w_empty[stage]);
              umma_commit(&hdr->pre_full[pb]);
            }
            __syncwarp();
            if (++stage == kMlpStages) { stage = 0; phase ^= 1; }
            if (t == 1 && c >= 0) MLP_TS(2, 3 + 2 * c);
          }
        }
        if (t == 1) MLP_TS(2, 40);
        g0 += uint32_t(NC);
      }
    }
  } else if (warp == 2 + kMlpEpiWarps) {
    if (lane == 0) {
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        for (int c = 0; c < NC; ++c, ++g) {
          const uint32_t b = g & 1u;
          mbar_wait(&hdr->h_ready[b], (g >> 1) & 1u);
          tma_store_2d(&mapH, h_s + b * 16384, c * kMlpHC, tile * BM);  // rows past the end are clipped by the TMA
          tma_store_2d(&mapG, g_s + b * 16384, c * kMlpHC, tile * BM);
          tma_store_wait_read();
          mbar_arrive(&hdr->h_free[b]);
        }
      }
      tma_store_wait_all();
    }
  } else {
    const int ew = warp - 2;        // 0..15
    const int q = warp & 3;         // TMEM lane quarter this warp may access
    const int grp = ew >> 2;        // which of the quarter's four warps
    const uint32_t half = uint32_t(grp >> 1), sub = uint32_t(grp & 1);  // GELU epilogue: chunk parity / 32-column half
    const uint32_t tab_s = smem_u32(smem + kMlpOffTabP) - kGeluLo * 4;
    // LayerNorm prologue mapping: 8 lanes per row (4 rows per warp instruction); lane `sub` owns the float4 column groups
    // 4 sub + 32 j, j = 0..5 (D <= 192), i.e. 16 B of every 128-byte line of the row: coalesced loads, 3 shuffle steps
    const int lsub = lane & 7, lrow = lane >> 3;
    bool has[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) has[j] = 4 * lsub + 32 * j < D;
    const float invD = 1.0f / float(D);
    uint8_t* stg = smem + kMlpOffA + ew * (32 * kMlpStgRow);
    uint32_t g0 = 0, t = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++t) {
      const int row0 = tile * BM;
      const int role = ew == 0 ? 0 : (ew == 8 ? 1 : 3);  // timeline only
      int kth = 0;
      if (t == 1 && role < 3) MLP_TS(role, 0);
      // ---------------- LayerNorm prologue: rows [8 ew, 8 ew + 8) of the tile ----------------
      {
        float4 v[2][6];
        long long srow[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int lr = ew * 8 + i * 4 + lrow;
          const int gr = row0 + lr;
          long long sr = -1;
          if (gr < p.rows) {
            sr = gr;
            if (p.gather_idx != nullptr)
              sr = (long long)(uint32_t(gr) / uint32_t(p.n_sel)) * p.L_src + __ldg(p.gather_idx + gr);
          }
          srow[i] = sr;
          if (lsub == 0) hdr->rowmap[lr] = int(sr);
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            v[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has[j] && sr >= 0) v[i][j] = *reinterpret_cast<const float4*>(p.x + sr * D + 4 * lsub + 32 * j);
          }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int lr = ew * 8 + i * 4 + lrow;
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 6; ++j) s += (v[i][j].x + v[i][j].y) + (v[i][j].z + v[i][j].w);
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          const float mu = s * invD;
          float sq = 0.f;
#pragma unroll
          for (int j = 0; j < 6; ++j)
            if (has[j]) {
              const float a = v[i][j].x - mu, b = v[i][j].y - mu, c = v[i][j].z - mu, d = v[i][j].w - mu;
              sq += (a * a + b * b) + (c * c + d * d);
            }
          sq += __shfl_xor_sync(0xffffffffu, sq, 1);
          sq += __shfl_xor_sync(0xffffffffu, sq, 2);
          sq += __shfl_xor_sync(0xffffffffu, sq, 4);
          const float rs = rsqrtf(sq * invD + p.eps);
          const bool ok = srow[i] >= 0;
          const long long gr = row0 + lr;
          // row lr of K panel j/2: 16-byte chunk 4 (j & 1) + sub / 2 (swizzled by the row), 8 bytes at (sub & 1) * 8
          const uint32_t srow_s = a_s + uint32_t(lr) * 128u + uint32_t(lsub & 1) * 8u;
#pragma unroll
          for (int j = 0; j < 6; ++j)
            if (has[j]) {
              uint32_t lo = 0u, hi = 0u;
              if (ok) {  // (the affine parameters come from L1: keeping 12 float4 of them in registers spills)
                const float4 gm = __ldg(reinterpret_cast<const float4*>(p.gamma + 4 * lsub + 32 * j));
                const float4 bt = __ldg(reinterpret_cast<const float4*>(p.beta + 4 * lsub + 32 * j));
                lo = pack_bf16x2((v[i][j].x - mu) * rs * gm.x + bt.x, (v[i][j].y - mu) * rs * gm.y + bt.y);
                hi = pack_bf16x2((v[i][j].z - mu) * rs * gm.z + bt.z, (v[i][j].w - mu) * rs * gm.w + bt.w);
              }
              const uint32_t chunk = uint32_t(4 * (j & 1) + (lsub >> 1));
              const uint32_t sa = srow_s + uint32_t(j >> 1) * 16384u + ((chunk ^ uint32_t(lr & 7)) << 4);
              asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(sa), "r"(lo), "r"(hi) : "memory");
              if (ok) *reinterpret_cast<uint2*>(p.xn + gr * D + 4 * lsub + 32 * j) = make_uint2(lo, hi);
            }
          if (ok && lsub == 0) { p.mean[gr] = mu; p.rstd[gr] = rs; }
        }
        fence_proxy_async_smem();  // generic-proxy writes of the A tile -> visible to tcgen05.mma (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&hdr->a_ready);
        if (t == 1 && role < 3) MLP_TS(role, 1);
      }
      {  // pull this warp's 8 rows of the CTA's next tile into L2 while the current tile is computed
        const int ntile = tile + gridDim.x;
        if (ntile < p.ntiles) {
          const int gr = ntile * BM + ew * 8 + (lane >> 2);
          if (gr < p.rows) {
            long long sr = gr;
            if (p.gather_idx != nullptr)
              sr = (long long)(uint32_t(gr) / uint32_t(p.n_sel)) * p.L_src + __ldg(p.gather_idx + gr);
            const char* rowp = reinterpret_cast<const char*>(p.x + sr * D);
            const int row_bytes = D * 4;
#pragma unroll
            for (int k = 0; k < 2; ++k) {  // 4 lanes per row, 2 lines each: covers rows of up to 1 KB
              const int off = ((lane & 3) * 2 + k) * 128;
              if (off < row_bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(rowp + off));
            }
          }
        }
      }
      // ---------------- GELU epilogue of this warp's chunks ----------------
      for (int c = 0; c < NC; ++c) {
        const uint32_t g = g0 + uint32_t(c);
        if ((g & 1u) != half) continue;
        const uint32_t b = half, pb = g & 3u;  // GELU tile buffer / fc1 TMEM buffer of this chunk
        if (t == 1 && role < 3) MLP_TS(role, 2 + 4 * kth);
        mbar_wait(&hdr->pre_full[pb], (g >> 2) & 1u);
        tc_fence_after();
        if (t == 1 && role < 3) MLP_TS(role, 3 + 4 * kth);
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + kMlpTmemPre + pb * kMlpHC + sub * 32;
        uint32_t raw[2][16];
        tmem_ld16_issue(taddr, raw[0]);
        tmem_ld16_issue(taddr + 16, raw[1]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&hdr->pre_free[pb]);  // the accumulator is in registers: fc1 of chunk c+4 may overwrite it
        const int row = q * 32 + lane;
        const uint32_t hb = h_s + b * 16384, gb = g_s + b * 16384;
#pragma unroll
        for (int u = 0; u < 2; ++u) {  // two 16-column groups
          float bia[16];
          const float4* bp = reinterpret_cast<const float4*>(p.b1 + c * kMlpHC + sub * 32 + u * 16);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 t4 = __ldg(bp + i);
            bia[4 * i] = t4.x; bia[4 * i + 1] = t4.y; bia[4 * i + 2] = t4.z; bia[4 * i + 3] = t4.w;
          }
          uint32_t pre[8], hh[8], dd[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)  // autocast semantics: the linear's output is rounded to bf16 and GELU reads that value
            pre[i] = pack_bf16x2(__uint_as_float(raw[u][2 * i]) + bia[2 * i], __uint_as_float(raw[u][2 * i + 1]) + bia[2 * i + 1]);
          gelu_table8(pre, tab_s, hh, dd);
          // the first group's look-ups overlap chunk c-2's fc2 MMAs and TMA stores, which must be done with the tiles now
          // (measured: waiting before the loop instead, so that both groups' look-ups can interleave, is 10 % slower)
          if (u == 0) { mbar_wait(&hdr->h_free[b], ((g >> 1) & 1u) ^ 1u); if (t == 1 && role < 3) MLP_TS(role, 4 + 4 * kth); }
          const int ch = int(sub) * 4 + u * 2;  // 16-byte chunks of the 128-byte row
          sts128(swz128(hb, row, ch), hh[0], hh[1], hh[2], hh[3]);
          sts128(swz128(hb, row, ch + 1), hh[4], hh[5], hh[6], hh[7]);
          sts128(swz128(gb, row, ch), dd[0], dd[1], dd[2], dd[3]);
          sts128(swz128(gb, row, ch + 1), dd[4], dd[5], dd[6], dd[7]);
        }
        fence_proxy_async_smem();  // -> visible to tcgen05.mma (A operand) and to the TMA stores
        __syncwarp();
        if (lane == 0) mbar_arrive(&hdr->h_ready[b]);
        if (t == 1 && role < 3) MLP_TS(role, 5 + 4 * kth);
        ++kth;
      }
      // ---------------- output epilogue: y = acc + b2 + x, fp32 ----------------
      {
        const int nsc = D >> 4;  // 16-column units, dealt in contiguous ranges to the four warps of the lane quarter
        const int sc_begin = (nsc * grp) / 4, sc_end = (nsc * (grp + 1)) / 4;
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16);
        const int rr = lane >> 2, chk = lane & 3;  // flush: 4 lanes per 64-byte row segment, 8 rows per instruction
        // residual rows of all (<= 3) units of this warp, coalesced, issued before the accumulator is touched
        float4 xr[3][4];
        int orow[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) orow[i] = hdr->rowmap[q * 32 + i * 8 + rr];
#pragma unroll
        for (int u = 0; u < 3; ++u)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            xr[u][i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (sc_begin + u < sc_end && orow[i] >= 0)
              xr[u][i] = *reinterpret_cast<const float4*>(p.x + (long long)orow[i] * D + (sc_begin + u) * 16 + chk * 4);
          }
        if (t == 1 && role < 3) MLP_TS(role, 30);
        mbar_wait(&hdr->out_full, t & 1u);  // all fc2 MMAs of the tile have completed
        tc_fence_after();
        if (t == 1 && role < 3) MLP_TS(role, 31);
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int sc = sc_begin + u;
          if (sc >= sc_end) break;
          uint32_t raw[16];
          tmem_ld16_issue(taddr + sc * 16, raw);
          float bia[16];
          {
            const float4* bp = reinterpret_cast<const float4*>(p.b2 + sc * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 t4 = __ldg(bp + i);
              bia[4 * i] = t4.x; bia[4 * i + 1] = t4.y; bia[4 * i + 2] = t4.z; bia[4 * i + 3] = t4.w;
            }
          }
          tmem_ld_wait();
          float4* my = reinterpret_cast<float4*>(stg + lane * kMlpStgRow);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            my[i] = make_float4(__uint_as_float(raw[4 * i]) + bia[4 * i], __uint_as_float(raw[4 * i + 1]) + bia[4 * i + 1],
                                __uint_as_float(raw[4 * i + 2]) + bia[4 * i + 2], __uint_as_float(raw[4 * i + 3]) + bia[4 * i + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = i * 8 + rr;
            if (orow[i] >= 0) {
              const float4 a = *reinterpret_cast<const float4*>(stg + r * kMlpStgRow + chk * 16);
              *reinterpret_cast<float4*>(p.x_out + (long long)(row0 + q * 32 + r) * D + sc * 16 + chk * 4) =
                  make_float4(a.x + xr[u][i].x, a.y + xr[u][i].y, a.z + xr[u][i].z, a.w + xr[u][i].w);
            }
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&hdr->out_free);
      // the staging tiles alias the A panels and rowmap is rewritten by the next prologue: all epilogue warps first
      if (t == 1 && role < 3) MLP_TS(role, 32);
      named_bar_sync(1, kMlpEpiWarps * 32);
      if (t == 1 && role < 3) MLP_TS(role, 33);
      g0 += uint32_t(NC);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace mae

using namespace mae;

extern "C" int mae_mlp_fwd(const float* x, const int64_t* gather_idx, int n_sel, int L_src, int rows, int D,
                           const float* gamma, const float* beta, float eps, const void* w1_bf16, const float* b1,
                           const void* w2_bf16, const float* b2, void* xn_bf16, float* mean, float* rstd, void* h_bf16,
                           void* dgelu_bf16, float* x_out, cudaStream_t stream) {
  if (rows <= 0) return MAE_OK;
  if (!x || !gamma || !beta || !w1_bf16 || !b1 || !w2_bf16 || !b2 || !xn_bf16 || !mean || !rstd || !h_bf16 || !dgelu_bf16 ||
      !x_out)
    return MAE_ERR_NULL;
  if (D < 64 || D > 192 || (D % 16) || (gather_idx && (n_sel <= 0 || L_src <= 0))) return MAE_ERR_SHAPE;
  const int HID = 4 * D;
  if (HID % kMlpHC) return MAE_ERR_SHAPE;
  MlpFwdParams p;
  p.rows = rows; p.D = D; p.HID = HID; p.NC = HID / kMlpHC; p.nkb = (D + BK - 1) / BK; p.ntiles = (rows + BM - 1) / BM;
  p.x = x; p.gather_idx = gather_idx; p.n_sel = n_sel; p.L_src = L_src;
  p.gamma = gamma; p.beta = beta; p.eps = eps; p.b1 = b1; p.b2 = b2;
  p.xn = reinterpret_cast<__nv_bfloat16*>(xn_bf16); p.mean = mean; p.rstd = rstd; p.x_out = x_out;
  {
    const char* e = getenv("MAE_MLP_TIMELINE");  // device address of >= 3 x 64 int64 slots (tools/mlp_timeline.py)
    p.dbg = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr;
  }
  p.gelu_tab = gelu_table_device_ptr();
  if (!p.gelu_tab) return MAE_ERR_DRIVER;
  CUtensorMap m1, m2, mh, mg;
  int rc = make_tmap_bf16(&m1, w1_bf16, HID, D, D, kMlpHC);  // W1 (HID, D): boxes of [64 rows x 64 k]
  if (rc) return rc;
  rc = make_tmap_bf16(&m2, w2_bf16, D, HID, HID, D);          // W2 (D, HID): boxes of [D rows x 64 k]
  if (rc) return rc;
  rc = make_tmap_bf16(&mh, h_bf16, rows, HID, HID, BM);       // stores: [128 rows x 64 cols]
  if (rc) return rc;
  rc = make_tmap_bf16(&mg, dgelu_bf16, rows, HID, HID, BM);
  if (rc) return rc;
  static bool attr_done[kMaxDevices] = {};
  if (!ensure_dyn_smem(mlp_fwd_kernel, kMlpSmem + 1024, attr_done)) return MAE_ERR_LAUNCH;
  const int grid = p.ntiles < num_sms() ? p.ntiles : num_sms();
  mae_launch(mlp_fwd_kernel, dim3(grid), dim3(kMlpThreads), kMlpSmem + 1024, stream, p, m1, m2, mh, mg);
  return MAE_CHECK_LAUNCH();
}
