// tcgen05 attention for SHORT rows (the encoder of configs/mae.yaml: L = 36 kept tokens, 6 heads of 24): timm Attention
// core = softmax(q k^T / sqrt(hd)) v per (sample, head), forward and backward.
//
// A 128-row UMMA tile would be 72 % padding for one 36-token sample, so a work item is (GROUP of G samples, head): the
// G samples sit one after the other at a pitch of LP = L rounded up to 8 rows (L = 36: LP = 40, G = 3, 120 of 128 rows
// live), S = Q K^T is computed for the whole tile and only its G diagonal [LP x LP] blocks are used: a softmax thread
// (row per thread = TMEM lane) reads the LP score columns of its own sample, and the P (dS) tile in shared memory is
// block diagonal -- the off-diagonal part is zeroed once at kernel start and never written again, because row r always
// belongs to the same sample slot.  Padding costs tensor FLOPs (which this shape has to spare: 3 % of peak before), not
// HBM bytes.
//
// Loads: ONE 4-D TMA box per operand and item over (head dim, 3H slices, token, sample) = [32 x 1 x LP x G]: the box is
// larger than the tensor along the head dim (24 -> 32) and the token dim (36 -> 40), so TMA zero-fills the K padding of
// the MMAs and the rows between samples; samples past the batch end are zero-filled the same way.  Tiles are
// [128 rows x 64 bytes], 64-byte swizzle; rows G*LP..127 are zeroed once.  The same tile serves as K-major operand
// (S = Q K^T, dP = dO V^T) and as MN-major operand (O = P V, dQ = dS K, dK = dS^T Q, dV = P^T dO): no transposes.
//
// Persistent, one CTA per SM, 10 warps: two softmax groups of 4 warps (one TMEM lane quarter each) take the CTA's items
// alternately, so one group's softmax runs under the other group's MMAs; warp 8 = TMA producer, warp 9 = tcgen05.mma
// issuer.  MMAs complete in issue order, which the schedule uses: the accumulators of a group's previous item are
// complete whenever the S of its next item is, so each group drains them right after pulling S (no extra waits).
#include "mae_internal.h"
#include "attn_tc_common.cuh"

namespace mae {

typedef CUresult (*PFN_encodeTiledN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// bf16 tensor (samples, tokens, slices, hd) viewed as (hd, slices, tokens, samples); box = [32 x 1 x LP x G], 64-byte
// swizzle, out-of-bounds elements (head dims >= hd, tokens >= L, samples >= B) are zero-filled.
static int make_tmap_enc(CUtensorMap* map, const void* ptr, int hd, int slices, int L, int B, int LP, int G) {
  static PFN_encodeTiledN enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return MAE_ERR_DRIVER;
    enc = reinterpret_cast<PFN_encodeTiledN>(p);
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (hd & 7)) return MAE_ERR_ALIGN;
  const cuuint64_t row = cuuint64_t(slices) * cuuint64_t(hd) * 2;
  cuuint64_t dims[4] = {cuuint64_t(hd), cuuint64_t(slices), cuuint64_t(L), cuuint64_t(B)};
  cuuint64_t strides[3] = {cuuint64_t(hd) * 2, row, row * cuuint64_t(L)};
  cuuint32_t box[4] = {32, 1, cuuint32_t(LP), cuuint32_t(G)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MAE_OK : MAE_ERR_DRIVER;
}

__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
// LP consecutive TMEM columns (LP in {16, 32, 40, 48, 64}) of this warp's lanes
template <int LP>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[LP]) {
  static_assert(LP == 16 || LP == 32 || LP == 40 || LP == 48 || LP == 64, "row pitch");
  if constexpr (LP == 16) tmem_ld16_nowait(taddr, r);
  if constexpr (LP >= 32) tmem_ld32_nowait(taddr, r);
  if constexpr (LP == 40) tmem_ld8_nowait(taddr + 32, r + 32);
  if constexpr (LP == 48) tmem_ld16_nowait(taddr + 32, r + 32);
  if constexpr (LP == 64) tmem_ld32_nowait(taddr + 32, r + 32);
  tmem_wait_ld();
}
// the score columns of this thread's own sample: a warp's 32 rows touch at most two samples (LP >= 16), the warp pulls
// both column segments and each thread keeps its own
template <int LP>
__device__ __forceinline__ void tmem_ld_own(uint32_t taddr, int g_lo, bool two, bool second, uint32_t (&r)[LP]) {
  tmem_ld_cols<LP>(taddr + uint32_t(g_lo * LP), r);
  if (two) {  // warp uniform
    uint32_t t[LP];
    tmem_ld_cols<LP>(taddr + uint32_t((g_lo + 1) * LP), t);
#pragma unroll
    for (int i = 0; i < LP; ++i) r[i] = second ? t[i] : r[i];
  }
}
// 16-byte chunk c of row r in a [rows x 64 bytes] tile with the 64-byte swizzle (tile base 512-byte aligned)
__device__ __forceinline__ uint32_t swz64(uint32_t base, int row, int chunk) {
  return base + uint32_t(row) * 64u + (uint32_t(chunk ^ ((row >> 1) & 3)) << 4);
}
__device__ __forceinline__ uint4 lds16(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}

static constexpr int kEncThreads = 10 * 32;
static constexpr uint32_t kEncTile = 128 * 64;   // one [128 x 32] bf16 operand tile
static constexpr uint32_t kEncPTile = 2 * 16384;  // [128 queries x 128 keys] bf16, two 128-byte-swizzled boxes of 64 keys

// geometry of a pitch
template <int LP>
struct EncGeo {
  static constexpr int G = 128 / LP;                       // samples per tile
  static constexpr int ROWS = G * LP;                      // live rows / keys (<= 128)
  static constexpr int KS = (ROWS + 15) / 16;              // K = 16 steps over keys / query rows
  static constexpr int NS = KS * 16;                       // MMA N of S / dP
  static constexpr uint32_t kBox = uint32_t(ROWS) * 64u;   // bytes one TMA box writes
};

// =============================================================================================
// forward
// =============================================================================================
struct EfBars {
  uint64_t kv_full[4], kv_empty[4];  // Q / K / V of an item landed / its MMAs have completed
  uint64_t s_full[2], s_free[2];     // [group]: S complete in TMEM / pulled into registers
  uint64_t p_ready[2], p_free[2];    // [group]: P tile written / P V has read it
  uint64_t o_full[2];                // [group]: O complete in TMEM
  uint32_t tmem_base;
};
struct EfSmem {
  static constexpr int kStages = 4;
  static constexpr uint32_t qkv = 0;                              // [4 stages][Q, K, V]
  static constexpr uint32_t p = kStages * 3 * kEncTile;           // [2 groups] P tiles
  static constexpr uint32_t bars = p + 2 * kEncPTile;
  static constexpr uint32_t total = bars + 1024;
};
static constexpr uint32_t kEfColS = 0, kEfColO = 256;  // S: 2 x 128 columns, O: 2 x 32

template <int LP>
__global__ void __launch_bounds__(kEncThreads, 1)
attn_fwd_enc_kernel(const __grid_constant__ CUtensorMap mapQKV, __nv_bfloat16* __restrict__ out, float* __restrict__ lse,
                    int B, int L, int H, int hd, float scale) {
  using Geo = EncGeo<LP>;
  constexpr int G = Geo::G;
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  EfBars* bars = reinterpret_cast<EfBars*>(smem + EfSmem::bars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ngroups = (B + G - 1) / G;
  const int nitems = ngroups * H;
  const int my_items = nitems > int(blockIdx.x) ? (nitems - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x) : 0;

  // zero the operand and P tiles once: rows past G * LP, and the off-diagonal blocks of P, are never written afterwards
  for (uint32_t i = threadIdx.x; i < EfSmem::bars / 16; i += kEncThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&mapQKV);
    for (int i = 0; i < EfSmem::kStages; ++i) { mbar_init(&bars->kv_full[i], 1); mbar_init(&bars->kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->s_full[i], 1); mbar_init(&bars->s_free[i], 4);
      mbar_init(&bars->p_ready[i], 4); mbar_init(&bars->p_free[i], 1);
      mbar_init(&bars->o_full[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&bars->tmem_base, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_wait();

  if (warp == 8) {
    if (lane == 0) {
      for (int n = 0; n < my_items; ++n) {
        const int item = int(blockIdx.x) + n * int(gridDim.x);
        const int grp = item / H, h = item % H, s = n % EfSmem::kStages;
        mbar_wait(&bars->kv_empty[s], ((n / EfSmem::kStages) & 1) ^ 1);
        const uint32_t dst = smem_u32(smem + EfSmem::qkv + s * 3 * kEncTile);
        mbar_expect_tx(&bars->kv_full[s], 3 * Geo::kBox);
        tma_load_4d(dst, &mapQKV, &bars->kv_full[s], 0, h, 0, grp * G);
        tma_load_4d(dst + kEncTile, &mapQKV, &bars->kv_full[s], 0, H + h, 0, grp * G);
        tma_load_4d(dst + 2 * kEncTile, &mapQKV, &bars->kv_full[s], 0, 2 * H + h, 0, grp * G);
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      const uint32_t id_s = umma_idesc_bf16(128, Geo::NS, 0, 0);  // S: both operands K-major
      const uint32_t id_o = umma_idesc_bf16(128, 32, 0, 1);       // O: A = P K-major, B = V MN-major
      const uint64_t t64 = umma_sdesc_sw64(0, 16, 512);
      const uint64_t t128 = umma_sdesc_sw128(0, 16, 1024);
      auto d_at = [](uint64_t templ, uint32_t addr) { return templ + uint64_t((addr & 0x3ffffu) >> 4); };
      const uint32_t tiles0 = smem_u32(smem + EfSmem::qkv), p0 = smem_u32(smem + EfSmem::p);
      for (int n = 0; n <= my_items; ++n) {
        if (n < my_items) {  // S of item n
          const int gi = n & 1, k = n >> 1, s = n % EfSmem::kStages;
          mbar_wait(&bars->kv_full[s], (n / EfSmem::kStages) & 1);
          if (k >= 1) mbar_wait(&bars->s_free[gi], (k - 1) & 1);
          tc_fence_after();
          const uint32_t tb = tiles0 + uint32_t(s) * 3u * kEncTile;
          const uint64_t dq = d_at(t64, tb), dk = d_at(t64, tb + kEncTile);
          const uint32_t d = tmem_base + kEfColS + uint32_t(gi) * 128u;
          umma_bf16(d, dq, dk, id_s, 0);
          umma_bf16(d, dq + 2, dk + 2, id_s, 1);  // +32 B per K = 16
          umma_commit(&bars->s_full[gi]);
        }
        if (n >= 1) {  // O = P V of item n - 1 (its softmax ran under the S above)
          const int m = n - 1, gi = m & 1, k = m >> 1, s = m % EfSmem::kStages;
          mbar_wait(&bars->p_ready[gi], k & 1);  // also: the group has drained O of its previous item (program order)
          tc_fence_after();
          const uint64_t a0 = d_at(t128, p0 + uint32_t(gi) * kEncPTile);
          uint64_t b = d_at(t64, tiles0 + uint32_t(s) * 3u * kEncTile + 2 * kEncTile);
          const uint32_t d = tmem_base + kEfColO + uint32_t(gi) * 32u;
#pragma unroll
          for (int j = 0; j < Geo::KS; ++j, b += 64)  // V as MN-major B: +1024 B per 16 keys
            umma_bf16(d, a0 + uint64_t((uint32_t(j >> 2) * 16384u + uint32_t(j & 3) * 32u) >> 4), b, id_o, j != 0);
          umma_commit(&bars->o_full[gi]);
          umma_commit(&bars->p_free[gi]);
          umma_commit(&bars->kv_empty[s]);
        }
      }
    }
  } else {
    // softmax groups: warps 0..3 = group 0, 4..7 = group 1; row = TMEM lane
    const int gi = warp >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int g = r / LP, l = r - g * LP;               // sample slot and token of this row
    const int g_lo = (quarter * 32) / LP < G ? (quarter * 32) / LP : G - 1;  // (rows past G * LP: dead lanes)
    const bool two = g_lo + 1 < G && (quarter * 32 + 31) / LP > g_lo;
    const bool second = g > g_lo;
    const bool row_ok = g < G && l < L;
    const float sl2 = scale * 1.4426950408889634f;
    const uint32_t lane_off = uint32_t(quarter * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_off + kEfColS + uint32_t(gi) * 128u;
    const uint32_t o_addr = tmem_base + lane_off + kEfColO + uint32_t(gi) * 32u;
    const uint32_t p_base = smem_u32(smem + EfSmem::p) + uint32_t(gi) * kEncPTile;
    const int cnt = my_items > gi ? (my_items - gi + 1) / 2 : 0;
    const int D = H * hd;
    float inv_prev = 0.f;
#pragma unroll 1
    for (int k = 0; k <= cnt; ++k) {
      uint32_t sr[LP];
      if (k < cnt) {
        mbar_wait(&bars->s_full[gi], k & 1);
        tc_fence_after();
        tmem_ld_own<LP>(s_addr, g_lo, two, second, sr);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->s_free[gi]);
      }
      if (k >= 1) {  // O rows of the group's previous item: complete, because its P V was issued before this item's S
        const int item = int(blockIdx.x) + (2 * (k - 1) + gi) * int(gridDim.x);
        const int b = (item / H) * G + g, h = item % H;
        mbar_wait(&bars->o_full[gi], (k - 1) & 1);
        tc_fence_after();
        uint32_t o[32];
        tmem_ld32_nowait(o_addr, o);
        tmem_wait_ld();
        tc_fence_before();
        if (row_ok && b < B) {
          __nv_bfloat16* dst = out + (size_t(b) * L + l) * D + h * hd;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c * 8 >= hd) break;
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(o[8 * c]) * inv_prev, __uint_as_float(o[8 * c + 1]) * inv_prev);
            w.y = pack_bf16x2(__uint_as_float(o[8 * c + 2]) * inv_prev, __uint_as_float(o[8 * c + 3]) * inv_prev);
            w.z = pack_bf16x2(__uint_as_float(o[8 * c + 4]) * inv_prev, __uint_as_float(o[8 * c + 5]) * inv_prev);
            w.w = pack_bf16x2(__uint_as_float(o[8 * c + 6]) * inv_prev, __uint_as_float(o[8 * c + 7]) * inv_prev);
            *reinterpret_cast<uint4*>(dst + 8 * c) = w;
          }
        }
      }
      if (k < cnt) {
        const int item = int(blockIdx.x) + (2 * k + gi) * int(gridDim.x);
        const int b = (item / H) * G + g, h = item % H;
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < LP; ++i)
          if (i < L) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(sr[i]));
        const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        const float msl2 = mx * sl2;
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t pk[LP / 2];
#pragma unroll
        for (int i = 0; i < LP / 2; ++i) {
          float p0 = ex2_approx(__uint_as_float(sr[2 * i]) * sl2 - msl2);
          float p1 = ex2_approx(__uint_as_float(sr[2 * i + 1]) * sl2 - msl2);
          if (2 * i >= L) p0 = 0.f;
          if (2 * i + 1 >= L) p1 = 0.f;
          l4[i & 3] += p0 + p1;
          pk[i] = pack_bf16x2(p0, p1);
        }
        const float sum = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        if (k >= 1) mbar_wait(&bars->p_free[gi], (k - 1) & 1);  // the previous item's P V has read the tile
        if (g < G) {
          const int c0 = (g * LP) >> 3;  // first 16-byte chunk of this sample's keys
#pragma unroll
          for (int j = 0; j < LP / 8; ++j) {
            const int c = c0 + j;
            sts16(swz(p_base + uint32_t(c >> 3) * 16384u, r, c & 7), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p_ready[gi]);
        if (row_ok && b < B) lse[(size_t(b) * H + h) * L + l] = mx * scale + logf(sum);
        inv_prev = 1.0f / sum;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 512);
}

// =============================================================================================
// backward.  Per (group, head), with the forward's log-sum-exp:
//   S = Q K^T,  P = exp(S / sqrt(hd) - lse),  dP = dO V^T,  delta = rowsum(dO o O),  dS = P o (dP - delta)
//   dQ = dS K / sqrt(hd),  dK = dS^T Q / sqrt(hd),  dV = P^T dO
// Q, K, V, dO and O tiles by TMA (delta comes from the dO / O tiles in shared memory: no strided global row reads);
// S and dP share one TMEM region (a group pulls its rows into registers at once, which frees it for the next item);
// the bf16 P and dS tiles (block diagonal, one pair per group) feed dQ = dS K as K-major A and dK / dV as MN-major A.
// =============================================================================================
struct EbBars {
  uint64_t ld_full[2], ld_empty[2];    // [stage = group]
  uint64_t sd_full[2], sd_free[2];     // [group]: S and dP complete / pulled out of TMEM
  uint64_t pds_ready[2], pds_free[2];  // [group]: P and dS tiles written / the MMAs reading them are done
  uint64_t g_full[2];                  // [group]: dQ, dK, dV complete
  uint32_t tmem_base;
};
struct EbSmem {
  static constexpr uint32_t tiles = 0;                           // [2 stages][Q, K, V, dO, O]
  static constexpr uint32_t pds = 2 * 5 * kEncTile;              // [2 groups][P, dS]
  static constexpr uint32_t bars = pds + 2 * 2 * kEncPTile;
  static constexpr uint32_t total = bars + 1024;
};
static constexpr uint32_t kEbColS = 0, kEbColDP = 128, kEbColG = 256;  // per group at kEbColG + 96 g: dQ, dK, dV (32 each)

template <int LP>
__global__ void __launch_bounds__(kEncThreads, 1)
attn_bwd_enc_kernel(const __grid_constant__ CUtensorMap mapQKV, const __grid_constant__ CUtensorMap mapDO,
                    const __grid_constant__ CUtensorMap mapO, const float* __restrict__ lse,
                    __nv_bfloat16* __restrict__ dqkv, int B, int L, int H, int hd, float scale) {
  using Geo = EncGeo<LP>;
  constexpr int G = Geo::G;
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  EbBars* bars = reinterpret_cast<EbBars*>(smem + EbSmem::bars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ngroups = (B + G - 1) / G;
  const int nitems = ngroups * H;
  const int my_items = nitems > int(blockIdx.x) ? (nitems - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x) : 0;

  for (uint32_t i = threadIdx.x; i < EbSmem::bars / 16; i += kEncThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&mapQKV);
    tma_prefetch_desc(&mapDO);
    tma_prefetch_desc(&mapO);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->ld_full[i], 1); mbar_init(&bars->ld_empty[i], 5);  // the MMAs (commit) + the group's 4 warps (delta)
      mbar_init(&bars->sd_full[i], 1); mbar_init(&bars->sd_free[i], 4);
      mbar_init(&bars->pds_ready[i], 4); mbar_init(&bars->pds_free[i], 1);
      mbar_init(&bars->g_full[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&bars->tmem_base, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_wait();

  if (warp == 8) {
    if (lane == 0) {
      for (int n = 0; n < my_items; ++n) {
        const int item = int(blockIdx.x) + n * int(gridDim.x);
        const int grp = item / H, h = item % H, s = n & 1;
        mbar_wait(&bars->ld_empty[s], ((n >> 1) & 1) ^ 1);
        const uint32_t dst = smem_u32(smem + EbSmem::tiles + s * 5 * kEncTile);
        mbar_expect_tx(&bars->ld_full[s], 5 * Geo::kBox);
        tma_load_4d(dst, &mapQKV, &bars->ld_full[s], 0, h, 0, grp * G);
        tma_load_4d(dst + kEncTile, &mapQKV, &bars->ld_full[s], 0, H + h, 0, grp * G);
        tma_load_4d(dst + 2 * kEncTile, &mapQKV, &bars->ld_full[s], 0, 2 * H + h, 0, grp * G);
        tma_load_4d(dst + 3 * kEncTile, &mapDO, &bars->ld_full[s], 0, h, 0, grp * G);
        tma_load_4d(dst + 4 * kEncTile, &mapO, &bars->ld_full[s], 0, h, 0, grp * G);
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      const uint32_t id_s = umma_idesc_bf16(128, Geo::NS, 0, 0);  // S, dP: both operands K-major
      const uint32_t id_q = umma_idesc_bf16(128, 32, 0, 1);       // dQ: A K-major, B MN-major
      const uint32_t id_kv = umma_idesc_bf16(128, 32, 1, 1);      // dK, dV: both MN-major
      const uint64_t t64 = umma_sdesc_sw64(0, 16, 512);
      const uint64_t t128k = umma_sdesc_sw128(0, 16, 1024);        // P / dS tile as K-major A
      const uint64_t t128mn = umma_sdesc_sw128(0, 16384, 1024);    // ... as MN-major A (LBO = stride between the 64-key boxes)
      auto d_at = [](uint64_t templ, uint32_t addr) { return templ + uint64_t((addr & 0x3ffffu) >> 4); };
      const uint32_t tiles0 = smem_u32(smem + EbSmem::tiles), pds0 = smem_u32(smem + EbSmem::pds);
      auto issue_sdp = [&](int n) {
        const uint32_t tb = tiles0 + uint32_t(n & 1) * 5u * kEncTile;
        const uint64_t dq = d_at(t64, tb), dk = d_at(t64, tb + kEncTile), dv = d_at(t64, tb + 2 * kEncTile),
                       ddo = d_at(t64, tb + 3 * kEncTile);
        umma_bf16(tmem_base + kEbColS, dq, dk, id_s, 0);
        umma_bf16(tmem_base + kEbColS, dq + 2, dk + 2, id_s, 1);
        umma_bf16(tmem_base + kEbColDP, ddo, dv, id_s, 0);
        umma_bf16(tmem_base + kEbColDP, ddo + 2, dv + 2, id_s, 1);
        umma_commit(&bars->sd_full[n & 1]);
      };
      auto issue_grads = [&](int n) {
        const int gi = n & 1;
        const uint32_t tb = tiles0 + uint32_t(gi) * 5u * kEncTile;
        const uint32_t sp = pds0 + uint32_t(gi) * 2u * kEncPTile, sds = sp + kEncPTile;
        const uint32_t dcol = tmem_base + kEbColG + uint32_t(gi) * 96u;
        {  // dQ = dS . K   (B = K tile as MN-major: +1024 B per 16 keys)
          const uint64_t a0 = d_at(t128k, sds);
          uint64_t b = d_at(t64, tb + kEncTile);
#pragma unroll
          for (int j = 0; j < Geo::KS; ++j, b += 64)
            umma_bf16(dcol, a0 + uint64_t((uint32_t(j >> 2) * 16384u + uint32_t(j & 3) * 32u) >> 4), b, id_q, j != 0);
        }
        {  // dK = dS^T . Q,  dV = P^T . dO   (M = keys; K = query rows: A +2048 B, B +1024 B per 16 rows)
          uint64_t a = d_at(t128mn, sds), b = d_at(t64, tb);
#pragma unroll
          for (int s = 0; s < Geo::KS; ++s, a += 128, b += 64) umma_bf16(dcol + 32, a, b, id_kv, s != 0);
          a = d_at(t128mn, sp); b = d_at(t64, tb + 3 * kEncTile);
#pragma unroll
          for (int s = 0; s < Geo::KS; ++s, a += 128, b += 64) umma_bf16(dcol + 64, a, b, id_kv, s != 0);
        }
        umma_commit(&bars->g_full[gi]);
        umma_commit(&bars->pds_free[gi]);
        umma_commit(&bars->ld_empty[gi]);
      };
      if (my_items > 0) {
        mbar_wait(&bars->ld_full[0], 0);
        tc_fence_after();
        issue_sdp(0);
      }
      for (int n = 0; n < my_items; ++n) {
        if (n + 1 < my_items) {  // the next item's S / dP as soon as this item's rows are in registers
          mbar_wait(&bars->ld_full[(n + 1) & 1], ((n + 1) >> 1) & 1);
          mbar_wait(&bars->sd_free[n & 1], (n >> 1) & 1);
          tc_fence_after();
          issue_sdp(n + 1);
        }
        mbar_wait(&bars->pds_ready[n & 1], (n >> 1) & 1);  // also: the group has drained dQ / dK / dV of its previous item
        tc_fence_after();
        issue_grads(n);
      }
    }
  } else {
    const int gi = warp >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int g = r / LP, l = r - g * LP;
    const int g_lo = (quarter * 32) / LP < G ? (quarter * 32) / LP : G - 1;  // (rows past G * LP: dead lanes)
    const bool two = g_lo + 1 < G && (quarter * 32 + 31) / LP > g_lo;
    const bool second = g > g_lo;
    const bool row_ok = g < G && l < L;
    const float l2e = 1.4426950408889634f, sl2 = scale * l2e;
    const uint32_t lane_off = uint32_t(quarter * 32) << 16;
    const uint32_t g_addr = tmem_base + lane_off + kEbColG + uint32_t(gi) * 96u;
    const uint32_t p_base = smem_u32(smem + EbSmem::pds) + uint32_t(gi) * 2u * kEncPTile, ds_base = p_base + kEncPTile;
    const uint32_t tile_do = smem_u32(smem + EbSmem::tiles) + uint32_t(gi) * 5u * kEncTile + 3 * kEncTile, tile_o = tile_do + kEncTile;
    const int cnt = my_items > gi ? (my_items - gi + 1) / 2 : 0;
    const int D = H * hd;
#pragma unroll 1
    for (int k = 0; k <= cnt; ++k) {
      uint32_t sr[LP], dr[LP];
      float lse2 = INFINITY, delta = 0.f;
      if (k < cnt) {
        const int item = int(blockIdx.x) + (2 * k + gi) * int(gridDim.x);
        const int b = (item / H) * G + g, h = item % H;
        const bool ok = row_ok && b < B;
        if (ok) lse2 = __ldg(lse + (size_t(b) * H + h) * L + l) * l2e;  // rows that do not exist: +inf makes P and dS vanish
        mbar_wait(&bars->ld_full[gi], k & 1);
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // delta = rowsum(dO o O) from the staged tiles (padding columns are zero)
          if (c * 8 >= hd) break;
          const uint4 o4 = lds16(swz64(tile_o, r, c)), d4 = lds16(swz64(tile_do, r, c));
          delta += bf16_lo(o4.x) * bf16_lo(d4.x) + bf16_hi(o4.x) * bf16_hi(d4.x) + bf16_lo(o4.y) * bf16_lo(d4.y) + bf16_hi(o4.y) * bf16_hi(d4.y) +
                   bf16_lo(o4.z) * bf16_lo(d4.z) + bf16_hi(o4.z) * bf16_hi(d4.z) + bf16_lo(o4.w) * bf16_lo(d4.w) + bf16_hi(o4.w) * bf16_hi(d4.w);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->ld_empty[gi]);
        mbar_wait(&bars->sd_full[gi], k & 1);
        tc_fence_after();
        tmem_ld_own<LP>(tmem_base + lane_off + kEbColS, g_lo, two, second, sr);
        tmem_ld_own<LP>(tmem_base + lane_off + kEbColDP, g_lo, two, second, dr);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->sd_free[gi]);
      }
      if (k >= 1) {  // gradients of the group's previous item: complete (issued before this item's S / dP)
        const int item = int(blockIdx.x) + (2 * (k - 1) + gi) * int(gridDim.x);
        const int b = (item / H) * G + g, h = item % H;
        mbar_wait(&bars->g_full[gi], (k - 1) & 1);
        tc_fence_after();
        __nv_bfloat16* dst = dqkv + (size_t(b) * L + l) * 3 * D + h * hd;
        const bool ok = row_ok && b < B;
#pragma unroll 1
        for (int w = 0; w < 3; ++w) {  // dQ, dK, dV one after the other: 32 live registers next to the S / dP rows
          uint32_t o[32];
          tmem_ld32_nowait(g_addr + uint32_t(w) * 32u, o);
          tmem_wait_ld();
          const float mul = w == 2 ? 1.0f : scale;
          if (ok) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (c * 8 >= hd) break;
              uint4 v;
              v.x = pack_bf16x2(__uint_as_float(o[8 * c]) * mul, __uint_as_float(o[8 * c + 1]) * mul);
              v.y = pack_bf16x2(__uint_as_float(o[8 * c + 2]) * mul, __uint_as_float(o[8 * c + 3]) * mul);
              v.z = pack_bf16x2(__uint_as_float(o[8 * c + 4]) * mul, __uint_as_float(o[8 * c + 5]) * mul);
              v.w = pack_bf16x2(__uint_as_float(o[8 * c + 6]) * mul, __uint_as_float(o[8 * c + 7]) * mul);
              *reinterpret_cast<uint4*>(dst + w * D + 8 * c) = v;
            }
          }
        }
        tc_fence_before();
      }
      if (k < cnt) {
        uint32_t pk[LP / 2], dk[LP / 2];
#pragma unroll
        for (int i = 0; i < LP / 2; ++i) {
          float p0 = ex2_approx(__uint_as_float(sr[2 * i]) * sl2 - lse2);
          float p1 = ex2_approx(__uint_as_float(sr[2 * i + 1]) * sl2 - lse2);
          if (2 * i >= L) p0 = 0.f;
          if (2 * i + 1 >= L) p1 = 0.f;
          pk[i] = pack_bf16x2(p0, p1);
          dk[i] = pack_bf16x2(p0 * (__uint_as_float(dr[2 * i]) - delta), p1 * (__uint_as_float(dr[2 * i + 1]) - delta));
        }
        if (k >= 1) mbar_wait(&bars->pds_free[gi], (k - 1) & 1);  // the previous item's MMAs have read the tiles
        if (g < G) {
          const int c0 = (g * LP) >> 3;
#pragma unroll
          for (int j = 0; j < LP / 8; ++j) {
            const int c = c0 + j;
            const uint32_t off = uint32_t(c >> 3) * 16384u;
            sts16(swz(p_base + off, r, c & 7), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            sts16(swz(ds_base + off, r, c & 7), dk[4 * j], dk[4 * j + 1], dk[4 * j + 2], dk[4 * j + 3]);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->pds_ready[gi]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 512);
}

template <int LP>
static int launch_fwd_enc(const void* qkv, int B, int L, int H, int hd, void* out, float* lse, cudaStream_t st) {
  using Geo = EncGeo<LP>;
  CUtensorMap mq;
  int rc = make_tmap_enc(&mq, qkv, hd, 3 * H, L, B, LP, Geo::G);
  if (rc) return rc;
  static bool done[kMaxDevices] = {};
  if (!ensure_dyn_smem(attn_fwd_enc_kernel<LP>, EfSmem::total + 1024, done)) return MAE_ERR_LAUNCH;
  const int sms = num_sms();
  const int nitems = ((B + Geo::G - 1) / Geo::G) * H;
  const int grid = nitems < sms ? nitems : sms;
  mae_launch(attn_fwd_enc_kernel<LP>, dim3(grid), dim3(kEncThreads), size_t(EfSmem::total) + 1024, st, mq,
             reinterpret_cast<__nv_bfloat16*>(out), lse, B, L, H, hd, 1.0f / sqrtf(float(hd)));
  return MAE_CHECK_LAUNCH();
}

template <int LP>
static int launch_bwd_enc(const void* qkv, const void* out, const void* dout, const float* lse, int B, int L, int H, int hd,
                          void* dqkv, cudaStream_t st) {
  using Geo = EncGeo<LP>;
  CUtensorMap mq, md, mo;
  int rc = make_tmap_enc(&mq, qkv, hd, 3 * H, L, B, LP, Geo::G);
  if (rc) return rc;
  rc = make_tmap_enc(&md, dout, hd, H, L, B, LP, Geo::G);
  if (rc) return rc;
  rc = make_tmap_enc(&mo, out, hd, H, L, B, LP, Geo::G);
  if (rc) return rc;
  static bool done[kMaxDevices] = {};
  if (!ensure_dyn_smem(attn_bwd_enc_kernel<LP>, EbSmem::total + 1024, done)) return MAE_ERR_LAUNCH;
  const int sms = num_sms();
  const int nitems = ((B + Geo::G - 1) / Geo::G) * H;
  const int grid = nitems < sms ? nitems : sms;
  mae_launch(attn_bwd_enc_kernel<LP>, dim3(grid), dim3(kEncThreads), size_t(EbSmem::total) + 1024, st, mq, md, mo, lse,
             reinterpret_cast<__nv_bfloat16*>(dqkv), B, L, H, hd, 1.0f / sqrtf(float(hd)));
  return MAE_CHECK_LAUNCH();
}

// row pitch for L tokens: the smallest instantiated LP >= L (0: not covered by these kernels)
static int enc_pitch(int L) {
  const int opts[5] = {16, 32, 40, 48, 64};
  for (int i = 0; i < 5; ++i)
    if (opts[i] >= L) return opts[i];
  return 0;
}

bool attn_enc_covers(int L, int hd) { return L >= 1 && enc_pitch(L) != 0 && hd >= 8 && hd <= 32 && (hd & 7) == 0; }

int attn_fwd_enc(const void* qkv, int B, int L, int H, int hd, void* out, float* lse, cudaStream_t st) {
  switch (enc_pitch(L)) {
    case 16: return launch_fwd_enc<16>(qkv, B, L, H, hd, out, lse, st);
    case 32: return launch_fwd_enc<32>(qkv, B, L, H, hd, out, lse, st);
    case 40: return launch_fwd_enc<40>(qkv, B, L, H, hd, out, lse, st);
    case 48: return launch_fwd_enc<48>(qkv, B, L, H, hd, out, lse, st);
    case 64: return launch_fwd_enc<64>(qkv, B, L, H, hd, out, lse, st);
    default: return MAE_ERR_SHAPE;
  }
}

int attn_bwd_enc(const void* qkv, const void* out, const void* dout, const float* lse, int B, int L, int H, int hd,
                 void* dqkv, cudaStream_t st) {
  switch (enc_pitch(L)) {
    case 16: return launch_bwd_enc<16>(qkv, out, dout, lse, B, L, H, hd, dqkv, st);
    case 32: return launch_bwd_enc<32>(qkv, out, dout, lse, B, L, H, hd, dqkv, st);
    case 40: return launch_bwd_enc<40>(qkv, out, dout, lse, B, L, H, hd, dqkv, st);
    case 48: return launch_bwd_enc<48>(qkv, out, dout, lse, B, L, H, hd, dqkv, st);
    case 64: return launch_bwd_enc<64>(qkv, out, dout, lse, B, L, H, hd, dqkv, st);
    default: return MAE_ERR_SHAPE;
  }
}

}  // namespace mae
