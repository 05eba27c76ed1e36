// tcgen05 attention forward for the decoder-sized problems (128 < L <= 160 tokens, head_dim 32, even head count):
// timm Attention core = softmax(q k^T / sqrt(hd)) v per (sample, head)  (F.scaled_dot_product_attention, no mask).
//
// Persistent kernel, one CTA per SM, work item = (sample, pair of heads).  Warp roles:
//   warp 2      TMA producer: Q / K / V boxes [160 rows x 64 cols] (two heads side by side, 128-byte swizzle, rows >= L
//               zero-filled by a 3-D tensor map over (col, token, sample)), double buffered across items
//   warps 1, 3  tcgen05.mma issuers (one lane each; warp 1: query tile 0 = rows 0..127, warp 3: query tile 1 = rows
//               128..159): S = Q K^T (M=128, N=160, K=32) and O = P V (M=128, N=32, K=160; V is the MN-major B operand
//               straight from its tile).  Separate issuers keep the two softmax pipelines from waiting on each other.
//   warps 4..7  softmax of query tile 0, one TMEM lane quarter each (row per thread): the whole S row is pulled into
//               registers at once (tcgen05.ld x128 + x32), which frees the TMEM buffer for the next head's MMAs; row
//               max, exp2 / sum, P written as bf16 into a 128-byte-swizzled K-major smem tile (the A operand of P V)
//   warp 0      the same for query tile 1 (17 live rows)
// TMEM: S of each query tile (2 x 160 columns, single buffered) and O of each tile double buffered across heads
// (4 x 32): the O epilogue of head n is deferred until after the softmax of head n+1, so P V runs underneath.
// O goes through a swizzled staging tile and one TMA store per item; the log-sum-exp rows (for the backward) are stored
// directly.  One copy of the softmax code serves both tiles: with a single warp per scheduler an instruction-cache miss
// is not hidden by anything (60 KB of SASS cost 27 % "no instruction" stalls; 37 KB: 3 %).
#include "mae_internal.h"
#include "attn_tc_common.cuh"

namespace mae {


static constexpr int kTcLp = 160;                       // padded tokens (keys = MMA N, query tiles 128 + 32)
static constexpr int kTcTile = kTcLp * 128;             // one [160 x 64] bf16 box = 20480 B
static constexpr int kTcThreads = 8 * 32;  // 8 warps: up to 255 registers per thread (a softmax thread holds a whole S row)
// TMEM columns: S of query tile 0 / 1 (single buffered: a softmax thread pulls its whole row into registers at once, which
// frees the buffer for the next head's MMAs), O of both tiles double buffered across heads (their epilogue is deferred)
static constexpr uint32_t kColS0 = 0, kColS1 = 160, kColO0 = 320, kColO1 = 384;

struct TcBars {
  uint64_t kv_full[2], kv_empty[2];   // Q/K/V of an item landed / the item no longer needs its buffer
  uint64_t s_full[2], s_free[2];      // [tile]: S complete in TMEM / S pulled into registers
  uint64_t p_ready[2], p_free[2];     // [tile]: P tile written / P V finished reading it
  uint64_t o_full[2][2], o_free[2][2];  // [tile][head parity]
  uint64_t ostg_free;                 // the O staging tile may be overwritten (previous item's TMA store has read it)
  uint32_t tmem_base;
};

struct TcSmem {  // offsets from the 1024-aligned base
  static constexpr uint32_t qkv = 0;                               // [2 buffers][Q,K,V] x 20480
  static constexpr uint32_t p0 = 2 * 3 * kTcTile;                  // 3 boxes [128 x 64]
  static constexpr uint32_t p1 = p0 + 3 * 16384;                   // 3 boxes [32 x 64] (the MMA over-reads 12 KB beyond)
  static constexpr uint32_t ostg = p1 + 3 * 4096;                  // [160 x 64] O staging
  static constexpr uint32_t bars = ostg + kTcTile;
  static constexpr uint32_t total = bars + 1024;
};
static_assert(TcSmem::p1 + 2 * 4096 + 16384 <= TcSmem::total, "P1 over-read stays inside the allocation");

// One softmax warp (tile = query tile; row = TMEM lane = thread): persistent loop over the CTA's heads.  Iteration n does
// the softmax of head n and then the O epilogue of head n - 1 (whose P V ran underneath).  One copy of the code serves
// both tiles and every iteration: the unrolled row loops are large and an instruction-cache miss is not hidden by
// anything with one warp per scheduler.
__device__ __forceinline__ void tc_softmax_warp(TcBars* bars, uint8_t* smem, uint32_t tmem_base, const CUtensorMap* mapO,
                                                const int tile, int quarter, int lane, bool storer, int nitems, int L, int H,
                                                float scale, float* __restrict__ lse, const uint32_t nbar) {
  const int HP = H >> 1;
  const float sl2 = scale * 1.4426950408889634f;
  const int row = tile ? 128 + lane : quarter * 32 + lane;   // query token of this thread
  const int prow = tile ? lane : row;                         // its row inside the P tile
  const uint32_t lane_off = uint32_t(tile ? 0 : quarter * 32) << 16;
  const uint32_t s_addr = tmem_base + lane_off + (tile ? kColS1 : kColS0);
  const uint32_t o_addr = tmem_base + lane_off + (tile ? kColO1 : kColO0);
  const uint32_t p_base = smem_u32(smem + (tile ? TcSmem::p1 : TcSmem::p0));
  const uint32_t box_bytes = tile ? 4096u : 16384u;
  const uint32_t ost = smem_u32(smem + TcSmem::ostg);
  const int my_items = nitems > int(blockIdx.x) ? (nitems - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x) : 0;
  const int nheads = 2 * my_items;
  uint32_t ph_sfull = 0, ph_pfree = 0, ph_ofull0 = 0, ph_ofull1 = 0, ph_ostg = 0;
  float inv_prev = 0.f;  // 1 / row sum of the head whose O epilogue is pending

#pragma unroll 1
  for (int n = 0; n <= nheads; ++n) {
    if (n < nheads) {
      const int item = int(blockIdx.x) + (n >> 1) * int(gridDim.x);
      const int b = item / HP, h = (item % HP) * 2 + (n & 1);
      // the whole S row into registers, then hand the TMEM buffer back
      mbar_wait(&bars->s_full[tile], ph_sfull); ph_sfull ^= 1;
      tc_fence_after();
      uint32_t sr[160];
      tmem_ld128_nowait(s_addr, sr);
      tmem_ld32_nowait(s_addr + 128, sr + 128);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->s_free[tile]);
      float m4[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) m4[j] = __uint_as_float(sr[j]);
#pragma unroll
      for (int i = 8; i < 160; ++i)  // 8 independent chains (a single warp per scheduler has no other latency hiding)
        if (i < 128 || i < L) m4[i & 7] = fmaxf(m4[i & 7], __uint_as_float(sr[i]));
      const float mx = fmaxf(fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])), fmaxf(fmaxf(m4[4], m4[5]), fmaxf(m4[6], m4[7])));
      const float msl2 = mx * sl2;
      // P = exp2(S * sl2 - max * sl2) -> bf16 -> swizzled K-major smem tile; row sum
      if (n > 0) { mbar_wait(&bars->p_free[tile], ph_pfree); ph_pfree ^= 1; }  // previous head's P V has read the tile
      float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = ex2_approx(__uint_as_float(sr[c * 32 + 2 * i]) * sl2 - msl2);
          float p1 = ex2_approx(__uint_as_float(sr[c * 32 + 2 * i + 1]) * sl2 - msl2);
          if (c == 4) {
            if (128 + 2 * i >= L) p0 = 0.f;
            if (128 + 2 * i + 1 >= L) p1 = 0.f;
          }
          l4[i & 3] += p0 + p1;
          pk[i] = pack_bf16x2(p0, p1);
        }
        const uint32_t box = p_base + uint32_t(c >> 1) * box_bytes;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          sts16(swz(box, prow, (c & 1) * 4 + k), pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
      }
      const float l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_ready[tile]);
      if (row < L) lse[(size_t(b) * H + h) * L + row] = mx * scale + logf(l);
      const float inv_cur = 1.0f / l;
      // swap: the epilogue below needs the PREVIOUS head's normaliser
      const float t = inv_prev; inv_prev = inv_cur;
      if (n == 0) continue;
      // fall through to the epilogue of head n - 1 with its normaliser
      {
        const int m = n - 1, par = m & 1;
        mbar_wait(&bars->o_full[tile][par], par ? ph_ofull1 : ph_ofull0);
        if (par) ph_ofull1 ^= 1; else ph_ofull0 ^= 1;
        tc_fence_after();
        uint32_t o[32];
        tmem_ld32_nowait(o_addr + uint32_t(par) * 32u, o);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->o_free[tile][par]);
        if (par == 0 && m >= 2) { mbar_wait(&bars->ostg_free, ph_ostg); ph_ostg ^= 1; }  // previous item's store has read the tile
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            w[j] = pack_bf16x2(__uint_as_float(o[8 * k + 2 * j]) * t, __uint_as_float(o[8 * k + 2 * j + 1]) * t);
          sts16(swz(ost, row, par * 4 + k), w[0], w[1], w[2], w[3]);
        }
        if (par == 1) {  // both heads of the item are staged: one TMA store [160 x 64] for the pair
          const int item_m = int(blockIdx.x) + (m >> 1) * int(gridDim.x);
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, %0;" ::"r"(nbar) : "memory");  // the softmax warps of the item (160 threads; 128 without tile 1)
          if (storer && lane == 0) {
            tma_store_3d(mapO, ost, (item_m % HP) * 64, 0, item_m / HP);
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            mbar_arrive(&bars->ostg_free);
          }
        }
      }
    } else if (n > 0) {
      // epilogue of the last head
      const int m = n - 1, par = m & 1;
      const float t = inv_prev;
      mbar_wait(&bars->o_full[tile][par], par ? ph_ofull1 : ph_ofull0);
      tc_fence_after();
      uint32_t o[32];
      tmem_ld32_nowait(o_addr + uint32_t(par) * 32u, o);
      tmem_wait_ld();
      if (par == 0 && m >= 2) { mbar_wait(&bars->ostg_free, ph_ostg); ph_ostg ^= 1; }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          w[j] = pack_bf16x2(__uint_as_float(o[8 * k + 2 * j]) * t, __uint_as_float(o[8 * k + 2 * j + 1]) * t);
        sts16(swz(ost, row, par * 4 + k), w[0], w[1], w[2], w[3]);
      }
      const int item_m = int(blockIdx.x) + (m >> 1) * int(gridDim.x);
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, %0;" ::"r"(nbar) : "memory");  // the softmax warps of the item (160 threads; 128 without tile 1)
      if (storer && lane == 0) {
        tma_store_3d(mapO, ost, (item_m % HP) * 64, 0, item_m / HP);
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      }
    }
  }
}

__global__ void __launch_bounds__(kTcThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap mapQKV, const __grid_constant__ CUtensorMap mapO,
                   float* __restrict__ lse, int B, int L, int H, float scale, int skip1) {
  // skip1: only query rows 0..127 are wanted (the caller keeps the rows it needs first: last decoder block, whose rows past
  // the masked tokens are keys / values only): query tile 1, its issuer and its softmax warp do nothing; O rows >= 128 are
  // written as zeros, lse rows >= 128 are not written.
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  TcBars* bars = reinterpret_cast<TcBars*>(smem + TcSmem::bars);
  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;  // uniform: role branches are warp-uniform control flow
  const int HP = H >> 1, D = H * 32;
  const int nitems = B * HP;

  if (warp == 2 && lane == 0) {
    tma_prefetch_desc(&mapQKV);
    tma_prefetch_desc(&mapO);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->kv_full[i], 1); mbar_init(&bars->kv_empty[i], skip1 ? 1 : 2);
      mbar_init(&bars->s_full[i], 1); mbar_init(&bars->s_free[i], i == 0 ? 4 : 1);
      mbar_init(&bars->p_ready[i], i == 0 ? 4 : 1); mbar_init(&bars->p_free[i], 1);
      for (int j = 0; j < 2; ++j) { mbar_init(&bars->o_full[i][j], 1); mbar_init(&bars->o_free[i][j], i == 0 ? 4 : 1); }
    }
    mbar_init(&bars->ostg_free, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  // rows 128..159 of the O staging tile: zero unless query tile 1 writes them
  for (uint32_t i = threadIdx.x; i < 32u * 128u / 16u; i += kTcThreads) reinterpret_cast<uint4*>(smem + TcSmem::ostg + 128 * 128)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_wait();

  // The TMA and tcgen05.mma issuers run their loops with the whole warp (uniform control flow: descriptors and
  // coordinates stay in uniform registers) and elect one lane per burst; under `if (lane == 0)` the compiler wraps every
  // UTCHMMA / UTMALDG in an ELECT / R2UR.BROADCAST waterfall (~70 cycles of the issuing thread per MMA, ptx.cuh).
  if (warp == 2) {
    {
      int k = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++k) {
        const int s = k & 1;
        mbar_wait(&bars->kv_empty[s], ((k >> 1) & 1) ^ 1);
        const int b = item / HP, hp = item % HP;
        uint8_t* dst = smem + TcSmem::qkv + s * 3 * kTcTile;
        if (elect_one_sync()) {
          mbar_expect_tx(&bars->kv_full[s], 3 * kTcTile);
          tma_load_3d(dst, &mapQKV, &bars->kv_full[s], hp * 64, 0, b);
          tma_load_3d(dst + kTcTile, &mapQKV, &bars->kv_full[s], D + hp * 64, 0, b);
          tma_load_3d(dst + 2 * kTcTile, &mapQKV, &bars->kv_full[s], 2 * D + hp * 64, 0, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1 || (warp == 3 && !skip1)) {
    // one tcgen05.mma issuer per query tile (warp 1: tile 0, warp 3: tile 1): their waits do not couple the two softmax
    // pipelines; the TMEM regions are disjoint and the smem operands read-only
    {
      const int tile = warp == 1 ? 0 : 1;
      const uint32_t idesc_s = umma_idesc_bf16(128, kTcLp, 0, 0);   // S: both operands K-major
      const uint32_t idesc_o = umma_idesc_bf16(128, 32, 0, 1);      // O: A = P K-major, B = V MN-major
      const int my_items = nitems > int(blockIdx.x) ? (nitems - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x) : 0;
      const int nheads = 2 * my_items;
      const uint32_t p_s = smem_u32(smem + (tile ? TcSmem::p1 : TcSmem::p0));
      const uint32_t box_bytes = tile ? 4096u : 16384u;
      auto tiles = [&](int n) { return smem_u32(smem + TcSmem::qkv + ((n >> 1) & 1) * 3 * kTcTile); };
      auto issue_s = [&](int n) {  // S tile of head n
        const uint32_t q = tiles(n) + uint32_t(tile) * 128u * 128u + uint32_t(n & 1) * 64u;
        const uint32_t kk = tiles(n) + kTcTile + uint32_t(n & 1) * 64u;
        const uint32_t d = tmem_base + (tile ? kColS1 : kColS0);
        if (elect_one_sync()) {
#pragma unroll
          for (int k16 = 0; k16 < 2; ++k16)
            umma_bf16(d, umma_sdesc_sw128(q + k16 * 32, 16, 1024), umma_sdesc_sw128(kk + k16 * 32, 16, 1024), idesc_s, k16 != 0);
          umma_commit(&bars->s_full[tile]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int n, bool last_of_item) {
        const uint32_t v = tiles(n) + 2 * kTcTile + uint32_t(n & 1) * 64u;
        const uint32_t d = tmem_base + (tile ? kColO1 : kColO0) + uint32_t(n & 1) * 32u;
        if (elect_one_sync()) {
#pragma unroll
          for (int j = 0; j < kTcLp / 16; ++j)
            umma_bf16(d, umma_sdesc_sw128(p_s + uint32_t(j >> 2) * box_bytes + uint32_t(j & 3) * 32u, 16, 1024),
                      umma_sdesc_sw128(v + uint32_t(j) * 2048u, kTcTile, 1024), idesc_o, j != 0);
          umma_commit(&bars->o_full[tile][n & 1]);
          umma_commit(&bars->p_free[tile]);
          if (last_of_item) umma_commit(&bars->kv_empty[(n >> 1) & 1]);  // this issuer's MMAs on the item's Q / K / V are all in flight
        }
        __syncwarp();
      };
      uint32_t ph_sfree = 0, ph_pready = 0, ph_ofree[2] = {0, 0};
      if (nheads > 0) {
        mbar_wait(&bars->kv_full[0], 0);
        tc_fence_after();
        issue_s(0);
      }
      for (int n = 0; n < nheads; ++n) {
        if (n + 1 < nheads) {  // the next head's S as soon as the softmax warps hold this head's S in registers
          if (((n + 1) & 1) == 0) {  // first head of the next item
            const int k = (n + 1) >> 1;
            mbar_wait(&bars->kv_full[k & 1], (k >> 1) & 1);
          }
          mbar_wait(&bars->s_free[tile], ph_sfree); ph_sfree ^= 1;
          tc_fence_after();
          issue_s(n + 1);
        }
        mbar_wait(&bars->p_ready[tile], ph_pready); ph_pready ^= 1;
        if (n >= 2) { mbar_wait(&bars->o_free[tile][n & 1], ph_ofree[n & 1]); ph_ofree[n & 1] ^= 1; }
        tc_fence_after();
        issue_pv(n, (n & 1) != 0);
      }
    }
  } else if ((warp == 0 && !skip1) || warp >= 4) {
    // warp 0: query tile 1 (TMEM lane quarter 0 = rows 128..159 of the tile-1 accumulators); warps 4..7: tile 0
    tc_softmax_warp(bars, smem, tmem_base, &mapO, warp == 0 ? 1 : 0, warp & 3, lane, warp == 4, nitems, L, H, scale, lse,
                    skip1 ? 128u : 160u);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int attn_fwd_tc(const void* qkv, int B, int L, int H, void* out, float* lse, int q_rows, cudaStream_t st) {
  CUtensorMap mq, mo;
  const int D = H * 32;
  int rc = make_tmap_bf16_3d(&mq, qkv, 3 * D, L, B, 3 * D, kTcLp);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&mo, out, D, L, B, D, kTcLp);
  if (rc) return rc;
  static bool done[kMaxDevices] = {};
  if (!ensure_dyn_smem(attn_fwd_tc_kernel, TcSmem::total + 1024, done)) return MAE_ERR_LAUNCH;
  const int sms = num_sms();
  const int nitems = B * (H / 2);
  const int grid = nitems < sms ? nitems : sms;
  mae_launch(attn_fwd_tc_kernel, dim3(grid), dim3(kTcThreads), size_t(TcSmem::total) + 1024, st, mq, mo, lse, B, L, H,
             1.0f / sqrtf(32.0f), q_rows <= 128 ? 1 : 0);
  return MAE_CHECK_LAUNCH();
}


// =============================================================================================
// tcgen05 attention backward, same problem sizes.  Per (sample, head), with the forward's log-sum-exp:
//   S = Q K^T,  P = exp(S / sqrt(hd) - lse),  dP = dO V^T,  delta = rowsum(dO o O),  dS = P o (dP - delta)
//   dQ = dS K / sqrt(hd),  dK = dS^T Q / sqrt(hd),  dV = P^T dO
// Work item = (sample, pair of heads): Q, K, V, dO boxes [160 x 64] by TMA (single buffered: P / dS tiles take the rest
// of shared memory).  Per head the two query tiles (rows 0..127, 128..159) go through S and dP in TMEM one after the
// other; the softmax warps (4 for tile 0, 1 for tile 1, row per thread) turn them into bf16 P and dS tiles in shared
// memory (128-byte-swizzled, rows = queries), which feed three more MMA groups without any transpose:
//   dQ_t  = dS_t . K       A = dS tile as K-major, B = K tile as MN-major
//   dV   += P_t^T . dO_t   A = the SAME P tile read as MN-major (M = keys), B = dO tile as MN-major
//   dK   += dS_t^T . Q_t   likewise
// dK / dV (keys on the TMEM lanes, two 128-lane tiles) accumulate over both query tiles.  TMEM: S 160 + dP 160 + dQ 2x32 +
// dK 2x32 + dV 2x32 = 512 columns.  Gradients leave as 64-byte row segments straight from registers.
// =============================================================================================
struct TcbBars {
  uint64_t ld_full[2], ld_empty[2];   // [buffer]: Q / dO / K / V of a head landed / every MMA reading them has completed
  uint64_t sd_full[2][2], sd_free[2][2];  // [query tile][key half]: S and dP of 80 keys complete / pulled out of TMEM
  uint64_t pds_ready[2], pds_free[2]; // [query tile]: P and dS tiles written / the MMAs reading them are done
  uint64_t dq_full[2], dq_free[2];
  uint64_t dkv_full, dkv_free;
  uint32_t tmem_base;
};
struct TcbSmem {  // P0/dS0 first: the MN-major reads of the second key tile run 16 KB past a tile's end, into valid memory
  static constexpr uint32_t p0 = 0, ds0 = 49152;
  static constexpr uint32_t p1 = 98304, ds1 = 98304 + 12288;
  // per head and buffer: Q, dO, K, V boxes [160 rows x 32 cols] (64-byte rows, 64-byte swizzle), double buffered across
  // heads.  Q and dO first: as A operands of tile 1 (rows 128..255) they are read 6 KB past their 160 rows, into the next tile
  static constexpr uint32_t tile = kTcLp * 64;            // 10240 B
  static constexpr uint32_t tiles = 122880;               // [2 buffers][Q, dO, K, V]
  static constexpr uint32_t bars = tiles + 2 * 4 * tile;  // 204800
  // gradient staging of the tile-0 warps: per warp [dQ | dK | dV] x [32 rows x 64 bytes] (64-byte swizzle) for TMA stores
  static constexpr uint32_t stg = bars + 1024;
  static constexpr uint32_t total = stg + 4 * 3 * 2048;
};
// TMEM: two S / dP buffers of 80 key columns each (buffer = key half; S at +0, dP at +80): the two query tiles of a head
// take turns in them half by half, so tile 1's softmax overlaps tile 0's second half and the next head's first half
static constexpr uint32_t kbSD = 0, kbDQ = 320, kbDK = 384, kbDV = 448;

// 32 values of one TMEM row segment -> 64 contiguous bytes of a global bf16 row
__device__ __forceinline__ void store_row32(__nv_bfloat16* dst, const uint32_t (&o)[32], float mul) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint4 w;
    w.x = pack_bf16x2(__uint_as_float(o[8 * k]) * mul, __uint_as_float(o[8 * k + 1]) * mul);
    w.y = pack_bf16x2(__uint_as_float(o[8 * k + 2]) * mul, __uint_as_float(o[8 * k + 3]) * mul);
    w.z = pack_bf16x2(__uint_as_float(o[8 * k + 4]) * mul, __uint_as_float(o[8 * k + 5]) * mul);
    w.w = pack_bf16x2(__uint_as_float(o[8 * k + 6]) * mul, __uint_as_float(o[8 * k + 7]) * mul);
    *reinterpret_cast<uint4*>(dst + 8 * k) = w;
  }
}

// delta = rowsum(dO o O): DELTA = true reads it from `delta` (B, L, H) f32, written by attn_delta_kernel right before
// this launch (coalesced pre-pass); DELTA = false gathers the O and dO row segments per thread (64-byte pieces, 32 live
// registers, on the softmax warps' critical path: ~70 us of the launch).
template <bool DELTA>
__global__ void __launch_bounds__(kTcThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap mapQKV, const __grid_constant__ CUtensorMap mapDO,
                   const __grid_constant__ CUtensorMap mapG, const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                   const float* __restrict__ lse, const float* __restrict__ delta_in, __nv_bfloat16* __restrict__ dqkv,
                   int B, int L, int H, float scale, int q_rows) {
  // q_rows <= 128 ("skip1"): only the first q_rows query rows carry a gradient (dout is treated as zero on the others: the
  // last decoder block, whose rows past the masked tokens are keys / values only).  Query tile 1 then contributes nothing:
  // no S1 / dP1, no tile-1 softmax, no tile-1 gradient MMAs; the S / dP buffers are drained by tile 0 alone, so the next
  // head's S0 / dP0 are issued while tile 0 still works on this head.  dQ rows >= q_rows are written as zeros.
  const bool skip1 = q_rows <= 128;
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  TcbBars* bars = reinterpret_cast<TcbBars*>(smem + TcbSmem::bars);
  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;  // uniform: role branches are warp-uniform control flow
  const int D = H * 32;
  const int nitems = B * H;  // work item = (sample, head)
  const int nheads = nitems > int(blockIdx.x) ? (nitems - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x) : 0;

  if (warp == 2 && lane == 0) {
    tma_prefetch_desc(&mapQKV);
    tma_prefetch_desc(&mapDO);
    tma_prefetch_desc(&mapG);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&bars->ld_full[t], 1); mbar_init(&bars->ld_empty[t], 1);
      for (int hf = 0; hf < 2; ++hf) { mbar_init(&bars->sd_full[t][hf], 1); mbar_init(&bars->sd_free[t][hf], t == 0 ? 4 : 1); }
      mbar_init(&bars->pds_ready[t], t == 0 ? 4 : 1); mbar_init(&bars->pds_free[t], 1);
      mbar_init(&bars->dq_full[t], 1); mbar_init(&bars->dq_free[t], t == 0 ? 4 : 1);
    }
    mbar_init(&bars->dkv_full, 1); mbar_init(&bars->dkv_free, 5);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_wait();

  if (warp == 2) {
    {  // (whole warp + elect.sync: see the forward kernel)
      for (int k = 0; k < nheads; ++k) {
        const int item = int(blockIdx.x) + k * int(gridDim.x);
        const int b = item / H, h = item % H, buf = k & 1;
        uint8_t* dst = smem + TcbSmem::tiles + buf * 4 * TcbSmem::tile;
        mbar_wait(&bars->ld_empty[buf], ((k >> 1) & 1) ^ 1);
        if (elect_one_sync()) {
          mbar_expect_tx(&bars->ld_full[buf], 4 * TcbSmem::tile);
          tma_load_3d(dst, &mapQKV, &bars->ld_full[buf], h * 32, 0, b);
          tma_load_3d(dst + TcbSmem::tile, &mapDO, &bars->ld_full[buf], h * 32, 0, b);
          tma_load_3d(dst + 2 * TcbSmem::tile, &mapQKV, &bars->ld_full[buf], D + h * 32, 0, b);
          tma_load_3d(dst + 3 * TcbSmem::tile, &mapQKV, &bars->ld_full[buf], 2 * D + h * 32, 0, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t id_s = umma_idesc_bf16(128, kTcLp, 0, 0);   // S, dP: both operands K-major
      const uint32_t id_q = umma_idesc_bf16(128, 32, 0, 1);      // dQ: A K-major, B MN-major
      const uint32_t id_kv = umma_idesc_bf16(128, 32, 1, 1);     // dK, dV: both MN-major
      const uint32_t tiles0 = smem_u32(smem + TcbSmem::tiles);
      const uint32_t sp[2] = {smem_u32(smem + TcbSmem::p0), smem_u32(smem + TcbSmem::p1)};
      const uint32_t sds[2] = {smem_u32(smem + TcbSmem::ds0), smem_u32(smem + TcbSmem::ds1)};
      // Descriptor templates: the start-address field is the low 14 bits (address >> 4) and every operand lives below
      // 256 KB, so stepping through a tile is one 64-bit add of (byte offset >> 4): the issuing thread is on the critical
      // path (60+ MMAs per head from one lane), building each descriptor from scratch cost ~65 cycles per MMA.
      const uint64_t t64k = umma_sdesc_sw64(0, 16, 512);            // 64-byte rows: tiles loaded by TMA (K- or MN-major)
      const uint64_t t128k = umma_sdesc_sw128(0, 16, 1024);         // P / dS tiles as K-major A
      auto d_at = [](uint64_t templ, uint32_t addr) { return templ + uint64_t((addr & 0x3ffffu) >> 4); };
      const uint32_t id_s80 = umma_idesc_bf16(128, 80, 0, 0);
      auto issue_sdp = [&](int n, int t, int hf) {  // S_t = Q_t K^T and dP_t = dO_t V^T of head n, keys [80 hf, 80 hf + 80)
        const uint32_t tb = tiles0 + uint32_t(n & 1) * 4u * TcbSmem::tile, ro = uint32_t(t) * 8192u, ko = uint32_t(hf) * 80u * 64u;
        const uint64_t dq_ = d_at(t64k, tb + ro), ddo = d_at(t64k, tb + TcbSmem::tile + ro);
        const uint64_t dk_ = d_at(t64k, tb + 2 * TcbSmem::tile + ko), dv_ = d_at(t64k, tb + 3 * TcbSmem::tile + ko);
        const uint32_t d = tmem_base + kbSD + uint32_t(hf) * 160u;
        if (elect_one_sync()) {
          umma_bf16(d, dq_, dk_, id_s80, 0);
          umma_bf16(d, dq_ + 2, dk_ + 2, id_s80, 1);       // +32 B per K = 16
          umma_bf16(d + 80, ddo, dv_, id_s80, 0);
          umma_bf16(d + 80, ddo + 2, dv_ + 2, id_s80, 1);
          umma_commit(&bars->sd_full[t][hf]);
        }
        __syncwarp();
      };
      auto issue_grads = [&](int n, int t, bool last_of_head) {
        const uint32_t tb = tiles0 + uint32_t(n & 1) * 4u * TcbSmem::tile, ro = uint32_t(t) * 8192u;
        const uint32_t box = t ? 4096u : 16384u;
        const int ksteps = t ? 2 : 8;
        const uint64_t tmn = umma_sdesc_sw128(0, box, 1024);          // P / dS tiles as MN-major A (LBO = box stride)
        if (elect_one_sync()) {
          // dQ_t = dS_t . K   (B = K tile as MN-major: 32 head dims per 64-byte row, +1024 B per 16 keys)
          {
            const uint64_t a0 = d_at(t128k, sds[t]);
            uint64_t b = d_at(t64k, tb + 2 * TcbSmem::tile);
#pragma unroll
            for (int j = 0; j < kTcLp / 16; ++j, b += 64)
              umma_bf16(tmem_base + kbDQ + uint32_t(t) * 32u, a0 + uint64_t((uint32_t(j >> 2) * box + uint32_t(j & 3) * 32u) >> 4), b, id_q, j != 0);
          }
          umma_commit(&bars->dq_full[t]);
          // dV += P_t^T dO_t,  dK += dS_t^T Q_t   (M = keys: tile m covers keys 128 m ..; K = the query rows of tile t)
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            uint64_t a = d_at(tmn, sp[t] + uint32_t(2 * m) * box), b = d_at(t64k, tb + TcbSmem::tile + ro);
            for (int s = 0; s < ksteps; ++s, a += 128, b += 64)          // +2048 B / +1024 B per 16 query rows
              umma_bf16(tmem_base + kbDV + uint32_t(m) * 32u, a, b, id_kv, (t != 0 || s != 0) ? 1u : 0u);
            a = d_at(tmn, sds[t] + uint32_t(2 * m) * box); b = d_at(t64k, tb + ro);
            for (int s = 0; s < ksteps; ++s, a += 128, b += 64)
              umma_bf16(tmem_base + kbDK + uint32_t(m) * 32u, a, b, id_kv, (t != 0 || s != 0) ? 1u : 0u);
          }
          umma_commit(&bars->pds_free[t]);
          if (last_of_head) {
            umma_commit(&bars->dkv_full);
            umma_commit(&bars->ld_empty[n & 1]);
          }
        }
        __syncwarp();
      };
      uint32_t ph_sdfree[2][2] = {{0, 0}, {0, 0}}, ph_ready[2] = {0, 0}, ph_dqfree[2] = {0, 0}, ph_dkvfree = 0;
      auto wait_sdfree = [&](int t, int hf) { mbar_wait(&bars->sd_free[t][hf], ph_sdfree[t][hf]); ph_sdfree[t][hf] ^= 1; };
      if (skip1) {
        if (nheads > 0) {
          mbar_wait(&bars->ld_full[0], 0);
          tc_fence_after();
          issue_sdp(0, 0, 0);
          issue_sdp(0, 0, 1);
        }
        for (int n = 0; n < nheads; ++n) {
          if (n + 1 < nheads) {  // next head's S0 / dP0 as soon as tile 0 has pulled each buffer of this head
            mbar_wait(&bars->ld_full[(n + 1) & 1], ((n + 1) >> 1) & 1);
            for (int hf = 0; hf < 2; ++hf) {
              wait_sdfree(0, hf);
              tc_fence_after();
              issue_sdp(n + 1, 0, hf);
            }
          }
          mbar_wait(&bars->pds_ready[0], ph_ready[0]); ph_ready[0] ^= 1;
          if (n > 0) {
            mbar_wait(&bars->dq_free[0], ph_dqfree[0]); ph_dqfree[0] ^= 1;
            mbar_wait(&bars->dkv_free, ph_dkvfree); ph_dkvfree ^= 1;
          }
          tc_fence_after();
          issue_grads(n, 0, true);
        }
      }
      bool sdp0_issued = false;  // tile 0's S / dP of the current head were issued ahead (inside the previous head)
      for (int n = 0; n < (skip1 ? 0 : nheads); ++n) {
        if (!sdp0_issued) {
          mbar_wait(&bars->ld_full[n & 1], (n >> 1) & 1);
          for (int hf = 0; hf < 2; ++hf) {  // each buffer was last used by the previous head's tile 1
            if (n > 0) wait_sdfree(1, hf);
            tc_fence_after();
            issue_sdp(n, 0, hf);
          }
        }
        sdp0_issued = false;
        for (int hf = 0; hf < 2; ++hf) {  // tile 1 follows tile 0 through the two buffers
          wait_sdfree(0, hf);
          tc_fence_after();
          issue_sdp(n, 1, hf);
        }
        mbar_wait(&bars->pds_ready[0], ph_ready[0]); ph_ready[0] ^= 1;
        if (n > 0) {
          mbar_wait(&bars->dq_free[0], ph_dqfree[0]); ph_dqfree[0] ^= 1;
          mbar_wait(&bars->dkv_free, ph_dkvfree); ph_dkvfree ^= 1;
        }
        tc_fence_after();
        issue_grads(n, 0, false);
        if (n + 1 < nheads) {  // the next head's tile 0 as soon as tile 1 has drained each buffer
          mbar_wait(&bars->ld_full[(n + 1) & 1], ((n + 1) >> 1) & 1);
          for (int hf = 0; hf < 2; ++hf) {
            wait_sdfree(1, hf);
            tc_fence_after();
            issue_sdp(n + 1, 0, hf);
          }
          sdp0_issued = true;
        }
        mbar_wait(&bars->pds_ready[1], ph_ready[1]); ph_ready[1] ^= 1;
        if (n > 0) { mbar_wait(&bars->dq_free[1], ph_dqfree[1]); ph_dqfree[1] ^= 1; }
        tc_fence_after();
        issue_grads(n, 1, true);
      }
    }
  } else if (warp == 0 || warp >= 4) {
    const int tile = warp == 0 ? 1 : 0;
    const int quarter = warp & 3;
    const float l2e = 1.4426950408889634f, sl2 = scale * l2e;
    const int row = tile ? 128 + lane : quarter * 32 + lane;   // query token (softmax) and key token (dK / dV epilogue)
    const int prow = tile ? lane : row;
    const uint32_t lane_off = uint32_t(tile ? 0 : quarter * 32) << 16;
    const uint32_t p_base = smem_u32(smem + (tile ? TcbSmem::p1 : TcbSmem::p0));
    const uint32_t ds_base = smem_u32(smem + (tile ? TcbSmem::ds1 : TcbSmem::ds0));
    const uint32_t box_bytes = tile ? 4096u : 16384u;
    uint32_t ph_sdfull = 0, ph_pdsfree = 0, ph_dqfull = 0, ph_dkvfull = 0;
    // per-row constants of head m: log-sum-exp (in log2 units) and delta = sum_d dO o O, straight from global memory, one
    // head ahead (rows >= L: lse = +inf makes P and dS vanish).  issue (loads only) and finish (arithmetic) are separate
    // so that the loads of head n + 1 stay in flight across the epilogue waits of head n.
    uint4 rc_o[DELTA ? 1 : 4], rc_d[DELTA ? 1 : 4];
    float rc_l = INFINITY, rc_delta = 0.f;
    auto row_consts_issue = [&](int m) {
      rc_l = INFINITY; rc_delta = 0.f;
      if constexpr (!DELTA) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { rc_o[k] = make_uint4(0, 0, 0, 0); rc_d[k] = make_uint4(0, 0, 0, 0); }
      }
      if (m >= nheads || row >= L || row >= q_rows) return;  // (rows without a gradient: lse = +inf makes P and dS vanish)
      const int item = int(blockIdx.x) + m * int(gridDim.x);
      const int b = item / H, h = item % H;
      rc_l = __ldg(lse + (size_t(b) * H + h) * L + row);
      if constexpr (DELTA) {
        rc_delta = __ldg(delta_in + (size_t(b) * L + row) * H + h);
      } else {
        const uint4* op = reinterpret_cast<const uint4*>(out + (size_t(b) * L + row) * D + h * 32);
        const uint4* dp = reinterpret_cast<const uint4*>(dout + (size_t(b) * L + row) * D + h * 32);
#pragma unroll
        for (int k = 0; k < 4; ++k) { rc_o[k] = __ldg(op + k); rc_d[k] = __ldg(dp + k); }
      }
    };
    auto row_consts_finish = [&](float& lse2, float& delta) {
      lse2 = rc_l * l2e;
      delta = rc_delta;
      if constexpr (!DELTA) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint4 o4 = rc_o[k], d4 = rc_d[k];
        delta += bf16_lo(o4.x) * bf16_lo(d4.x) + bf16_hi(o4.x) * bf16_hi(d4.x) + bf16_lo(o4.y) * bf16_lo(d4.y) + bf16_hi(o4.y) * bf16_hi(d4.y) +
                 bf16_lo(o4.z) * bf16_lo(d4.z) + bf16_hi(o4.z) * bf16_hi(d4.z) + bf16_lo(o4.w) * bf16_lo(d4.w) + bf16_hi(o4.w) * bf16_hi(d4.w);
      }
      }
    };
    float lse2_next, delta_next;
    row_consts_issue(0);
    row_consts_finish(lse2_next, delta_next);
    // Gradient rows leave through shared memory and TMA stores for tile 0 (a row per thread written straight to global
    // memory is a 16-byte piece of 32 different 1152-byte-strided rows per instruction: ~80 us of LSU time per launch,
    // measured by skipping the stores); tile 1's 17 rows keep the direct stores.
    const uint32_t stg_s = smem_u32(smem + TcbSmem::stg) + uint32_t(quarter) * 3u * 2048u;
    auto stage_row32 = [&](uint32_t tile_s, const uint32_t (&o)[32], float mul) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        sts16(tile_s + uint32_t(lane) * 64u + (uint32_t(k ^ ((lane >> 1) & 3)) << 4),
              pack_bf16x2(__uint_as_float(o[8 * k]) * mul, __uint_as_float(o[8 * k + 1]) * mul),
              pack_bf16x2(__uint_as_float(o[8 * k + 2]) * mul, __uint_as_float(o[8 * k + 3]) * mul),
              pack_bf16x2(__uint_as_float(o[8 * k + 4]) * mul, __uint_as_float(o[8 * k + 5]) * mul),
              pack_bf16x2(__uint_as_float(o[8 * k + 6]) * mul, __uint_as_float(o[8 * k + 7]) * mul));
    };
    int prev_item = -1;  // tile-0 warps: the head whose dK / dV epilogue is pending
    auto dkv_epilogue = [&](int item_m) {
      const int bm = item_m / H, hm = item_m % H;
      mbar_wait(&bars->dkv_full, ph_dkvfull); ph_dkvfull ^= 1;
      tc_fence_after();
      uint32_t o[32], o2[32];
      tmem_ld32_nowait(tmem_base + lane_off + kbDK + uint32_t(tile) * 32u, o);
      tmem_ld32_nowait(tmem_base + lane_off + kbDV + uint32_t(tile) * 32u, o2);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->dkv_free);
      if (tile) {
        if (row < L) {
          __nv_bfloat16* base = dqkv + (size_t(bm) * L + row) * 3 * D + hm * 32;
          store_row32(base + D, o, scale);
          store_row32(base + 2 * D, o2, 1.0f);
        }
      } else {
        // the staging tiles were last read by the stores of two heads ago; only the dQ store of the previous head may be pending
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        stage_row32(stg_s + 2048u, o, scale);
        stage_row32(stg_s + 4096u, o2, 1.0f);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&mapG, stg_s + 2048u, D + hm * 32, quarter * 32, bm);
          tma_store_3d(&mapG, stg_s + 4096u, 2 * D + hm * 32, quarter * 32, bm);
        }
      }
    };
    if (tile && skip1) {
      // no tile-1 queries: this warp only drains the dK / dV rows of keys 128.. (accumulated from tile 0's queries) and
      // writes the zero dQ rows of the tokens >= 128
#pragma unroll 1
      for (int n = 0; n < nheads; ++n) {
        const int item = int(blockIdx.x) + n * int(gridDim.x);
        dkv_epilogue(item);
        if (row < L) {
          uint4* dst = reinterpret_cast<uint4*>(dqkv + (size_t(item / H) * L + row) * 3 * D + (item % H) * 32);
#pragma unroll
          for (int k = 0; k < 4; ++k) dst[k] = make_uint4(0, 0, 0, 0);
        }
      }
    }
#pragma unroll 1
    for (int n = 0; n < ((tile && skip1) ? 0 : nheads); ++n) {
      const int item = int(blockIdx.x) + n * int(gridDim.x);
      const int b = item / H, h = item % H;
      const float lse2 = lse2_next, delta = delta_next;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {  // keys [80 hf, 80 hf + 80): chunks of 32, 32 and 16 columns
        mbar_wait(&bars->sd_full[tile][hf], ph_sdfull); 
        tc_fence_after();
        if (hf == 0 && n > 0) { mbar_wait(&bars->pds_free[tile], ph_pdsfree); ph_pdsfree ^= 1; }  // previous head's MMAs have read the tiles
        const uint32_t s_addr = tmem_base + lane_off + kbSD + uint32_t(hf) * 160u, dp_addr = s_addr + 80;
        uint32_t sA[32], dA[32], sB[32], dB[32];
        tmem_ld32_nowait(s_addr, sA);
        tmem_ld32_nowait(dp_addr, dA);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          uint32_t (&sc)[32] = (c & 1) ? sB : sA;
          uint32_t (&dc)[32] = (c & 1) ? dB : dA;
          tmem_wait_ld();
          if (c == 0) {
            tmem_ld32_nowait(s_addr + 32, sB);
            tmem_ld32_nowait(dp_addr + 32, dB);
          } else if (c == 1) {
            tmem_ld16_nowait(s_addr + 64, sA);
            tmem_ld16_nowait(dp_addr + 64, dA);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->sd_free[tile][hf]);
          }
          const int npair = c < 2 ? 16 : 8;
          const int col0 = hf * 80 + c * 32;
          uint32_t pk[16], dk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (i >= npair) break;
            float p0 = ex2_approx(__uint_as_float(sc[2 * i]) * sl2 - lse2);
            float p1 = ex2_approx(__uint_as_float(sc[2 * i + 1]) * sl2 - lse2);
            if (hf == 1 && c >= 1) {  // only columns >= 112 can reach L (> 128)
              if (col0 + 2 * i >= L) p0 = 0.f;
              if (col0 + 2 * i + 1 >= L) p1 = 0.f;
            }
            pk[i] = pack_bf16x2(p0, p1);
            dk[i] = pack_bf16x2(p0 * (__uint_as_float(dc[2 * i]) - delta), p1 * (__uint_as_float(dc[2 * i + 1]) - delta));
          }
          const int g0 = col0 >> 3;  // first 16-byte chunk of this step inside the tile row: box g / 8, chunk g % 8
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k >= npair / 4) break;
            const int g = g0 + k;
            sts16(swz(p_base + uint32_t(g >> 3) * box_bytes, prow, g & 7), pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
            sts16(swz(ds_base + uint32_t(g >> 3) * box_bytes, prow, g & 7), dk[4 * k], dk[4 * k + 1], dk[4 * k + 2], dk[4 * k + 3]);
          }
        }
      }
      ph_sdfull ^= 1;
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->pds_ready[tile]);
      row_consts_issue(n + 1);  // in flight during the epilogue waits below
      if (!tile && n > 0) dkv_epilogue(prev_item);  // the previous head's dK / dV rows
      // dQ rows of this tile
      {
        mbar_wait(&bars->dq_full[tile], ph_dqfull); ph_dqfull ^= 1;
        tc_fence_after();
        uint32_t o[32];
        tmem_ld32_nowait(tmem_base + lane_off + kbDQ + uint32_t(tile) * 32u, o);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->dq_free[tile]);
        if (tile) {
          if (row < L) store_row32(dqkv + (size_t(b) * L + row) * 3 * D + h * 32, o, scale);
        } else {
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");  // (the dK / dV stores just issued may be pending)
          __syncwarp();
          stage_row32(stg_s, o, scale);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) tma_store_3d(&mapG, stg_s, h * 32, quarter * 32, b);
        }
      }
      // dK / dV rows (keys = TMEM lanes) once both query tiles have been accumulated.  The tile-0 warps defer this to
      // after the next head's softmax (dkv_epilogue below): the wait for tile 1 and its gradient MMAs is the longest of the
      // head, and the next head's S0 / dP0 are already in TMEM.
      if (tile) dkv_epilogue(item);
      else prev_item = item;
      row_consts_finish(lse2_next, delta_next);
    }
    if (!tile && nheads > 0) dkv_epilogue(prev_item);
    if (!tile && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores must complete before the smem goes away
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// delta[(b L + l) H + h] = sum_d out[b, l, h, d] * dout[b, l, h, d]: one thread per (token, head), hd contiguous bf16 each
// from both tensors (consecutive threads read consecutive 2 hd-byte pieces: fully coalesced), one float out.
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                                                        long long n, int hd, float* __restrict__ delta) {
  pdl_trigger();
  pdl_wait();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint4* op = reinterpret_cast<const uint4*>(out + t * hd);
  const uint4* dp = reinterpret_cast<const uint4*>(dout + t * hd);
  float acc = 0.f;
  for (int k = 0; k < hd / 8; ++k) {
    const uint4 o4 = __ldg(op + k), d4 = __ldg(dp + k);
    acc += bf16_lo(o4.x) * bf16_lo(d4.x) + bf16_hi(o4.x) * bf16_hi(d4.x) + bf16_lo(o4.y) * bf16_lo(d4.y) + bf16_hi(o4.y) * bf16_hi(d4.y) +
           bf16_lo(o4.z) * bf16_lo(d4.z) + bf16_hi(o4.z) * bf16_hi(d4.z) + bf16_lo(o4.w) * bf16_lo(d4.w) + bf16_hi(o4.w) * bf16_hi(d4.w);
  }
  delta[t] = acc;
}

int attn_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, int B, int L, int H, void* dqkv,
                float* delta_ws, int q_rows, cudaStream_t st) {
  CUtensorMap mq, md;
  const int D = H * 32;
  int rc = make_tmap_bf16_3d(&mq, qkv, 3 * D, L, B, 3 * D, kTcLp, 32);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&md, dout, D, L, B, D, kTcLp, 32);
  if (rc) return rc;
  CUtensorMap mg;  // gradient stores: [32 cols x 32 rows] boxes of dqkv (B, L, 3 D), 64-byte swizzle
  rc = make_tmap_bf16_3d(&mg, dqkv, 3 * D, L, B, 3 * D, 32, 32);
  if (rc) return rc;
  const int sms = num_sms();
  const int nitems = B * H;
  const int grid = nitems < sms ? nitems : sms;
  const __nv_bfloat16* o = reinterpret_cast<const __nv_bfloat16*>(out);
  const __nv_bfloat16* d_o = reinterpret_cast<const __nv_bfloat16*>(dout);
  if (delta_ws) {
    const long long n = (long long)B * L * H;
    mae_launch(attn_delta_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, o, d_o, n, 32, delta_ws);
    static bool done[kMaxDevices] = {};
    if (!ensure_dyn_smem(attn_bwd_tc_kernel<true>, TcbSmem::total + 1024, done)) return MAE_ERR_LAUNCH;
    mae_launch(attn_bwd_tc_kernel<true>, dim3(grid), dim3(kTcThreads), size_t(TcbSmem::total) + 1024, st, mq, md, mg, o, d_o, lse,
               (const float*)delta_ws, reinterpret_cast<__nv_bfloat16*>(dqkv), B, L, H, 1.0f / sqrtf(32.0f), q_rows);
  } else {
    static bool done[kMaxDevices] = {};
    if (!ensure_dyn_smem(attn_bwd_tc_kernel<false>, TcbSmem::total + 1024, done)) return MAE_ERR_LAUNCH;
    mae_launch(attn_bwd_tc_kernel<false>, dim3(grid), dim3(kTcThreads), size_t(TcbSmem::total) + 1024, st, mq, md, mg, o, d_o, lse,
               (const float*)nullptr, reinterpret_cast<__nv_bfloat16*>(dqkv), B, L, H, 1.0f / sqrtf(32.0f), q_rows);
  }
  return MAE_CHECK_LAUNCH();
}

}  // namespace mae
