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
// Operand tiles are [128 rows x 64 bytes], 64-byte swizzle, head dim 24 -> 32 and rows 36 -> 40 zero-padded, rows
// G*LP..127 zeroed once.  Backward: ONE 4-D TMA box per operand and item over (head dim, 3H slices, token, sample) =
// [32 x 1 x LP x G]: the box is larger than the tensor along the head dim and the token dim, so TMA zero-fills the K
// padding of the MMAs and the rows between samples; samples past the batch end are zero-filled the same way.  Forward:
// four loader warps copy the live 16-byte chunks with cp.async from a per-thread chunk table (TMA serves such a box at
// ~3.3 cycles per 48-byte row) and never touch the padding.  The same tile serves as K-major operand (S = Q K^T,
// dP = dO V^T) and as MN-major operand (O = P V, dQ = dS K, dK = dS^T Q, dV = P^T dO): no transposes.
//
// Persistent, one CTA per SM: two softmax groups of 4 warps (one TMEM lane quarter each) take the CTA's items in turn, so
// one group's softmax runs under the other group's MMAs; the MMAs come from separate issuer warps (forward: S | P V of
// the even items | P V of the odd items; backward: S, dP and dQ | dK | dV), each running warp-uniform and electing one
// lane per burst.  The forward relies on MMAs of ONE issuer completing in issue order only where the barriers say so:
// every accumulator is drained behind its own mbarrier.
#include <stdlib.h>

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

// Ampere-style 16-byte async copies (LDGSTS): TMA serves these tiles at ~3.3 cycles per 48-byte row (one request per
// (token, head) row: 1.2 k cycles per forward item, measured), loader warps with one 16-byte chunk per lane are faster
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {  // src_bytes 0: zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Chunk table of one loader thread: the 16-byte chunks (tensor slice t, live row, chunk j) it copies for every item, as
// offsets from the item's base address (global) and from the stage base (shared, 64-byte-swizzled tile t).  Rows >= L of a
// sample slot and the chunks past the head dim are never written: they stay zero from the kernel prologue.
template <int LP, int MAXC>
struct ChunkTable {
  uint32_t goff[MAXC], soff[MAXC];  // soff: bits 0..15 smem offset, bits 16.. sample slot (0xffffffff: no chunk)
  __device__ __forceinline__ void init(int lt, int nthreads, int nt, int L, int ch, uint32_t row_bytes, uint32_t slice_bytes, int dbg = 0) {
    constexpr int G = 128 / LP;
    const int per_t = G * L * ch, total = nt * per_t;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int idx = lt + nthreads * c;
      goff[c] = 0; soff[c] = 0xffffffffu;
      if (idx < total) {
        const int t = idx / per_t, rem = idx - t * per_t, rl = rem / ch, j = rem - rl * ch, g = rl / L, l = rl - g * L;
        const int row = g * LP + l;
        goff[c] = uint32_t(rl) * row_bytes + uint32_t(t) * slice_bytes + uint32_t(j) * 16u;
        soff[c] = (uint32_t(t) * (128u * 64u) + uint32_t(row) * 64u + (uint32_t(j ^ ((row >> 1) & 3)) << 4)) | (uint32_t(g) << 16);
        if ((dbg & 4) && j != 0) soff[c] = 0xffffffffu;  // timing experiment: one chunk per row (same requests, a third of the bytes)
        if (dbg & 8) goff[c] = uint32_t(idx) * 16u;      // timing experiment: the same bytes from contiguous addresses
      }
    }
  }
  __device__ __forceinline__ void issue(uint32_t stage, const char* base, int nvalid) const {
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (soff[c] != 0xffffffffu) {
        const bool ok = int(soff[c] >> 16) < nvalid;  // sample slots past the batch end are zero-filled
        cp_async16(stage + (soff[c] & 0xffffu), ok ? base + goff[c] : base, ok ? 16u : 0u);
      }
  }
};

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
  // a kernel instance serves MS < L <= LP (MS = the next smaller pitch): only columns >= MS need the `< L` predicate
  static constexpr int MS = LP == 16 ? 0 : LP == 32 ? 16 : LP == 40 ? 32 : LP == 48 ? 40 : 48;
  static constexpr int NG = 2;                              // forward: softmax groups
};

// (group of samples, head) of the k-th item of a softmax group, advanced without divisions
struct ItemIter {
  int grp, h, dq, dr, H;
  __device__ __forceinline__ ItemIter(int first, int step, int H_) : grp(first / H_), h(first % H_), dq(step / H_), dr(step % H_), H(H_) {}
  __device__ __forceinline__ void next() {
    grp += dq; h += dr;
    if (h >= H) { h -= H; ++grp; }
  }
};

// optional clock64 timeline of CTA 0 (MAE_ENC_TIMELINE = device address of 4 x 128 int64 slots, tools/enc_timeline.py)
#define ENC_TS(role, ev) do { if (tl != nullptr && blockIdx.x == 0 && (ev) < 128) tl[(role) * 128 + (ev)] = clock64(); } while (0)

// =============================================================================================
// forward
// =============================================================================================
template <int NG>
struct EfCfg {
  static constexpr int kStages = NG == 3 ? 5 : 6;                 // (NG = 3 was measured too: no faster, 128-register cap)
  static constexpr int kLoadAhead = 4;                            // cp.async groups in flight per loader thread (<= 4)
  static constexpr int kThreads = (4 * NG + 7) * 32;  // softmax groups + 4 loader warps + S issuer + two P V issuers
  static constexpr uint32_t qkv = 0;                              // [stages][Q, K, V]
  static constexpr uint32_t p = kStages * 3 * kEncTile;           // [NG groups] P tiles
  static constexpr uint32_t bars = p + NG * kEncPTile;
  static constexpr uint32_t total = bars + 1024;
  static constexpr uint32_t colS = 0, colO = NG * 128;            // TMEM: S NG x 128 columns, O NG x 32
};
struct EfBars {
  uint64_t kv_full[6], kv_empty[6];  // Q / K / V of an item landed / its MMAs have completed
  uint64_t s_full[3], s_free[3];     // [group]: S complete in TMEM / pulled into registers
  uint64_t p_ready[3];               // [group]: P tile written (and O of the previous item drained)
  uint64_t o_full[3];                // [group]: O complete in TMEM = P V has read the P tile
  uint32_t tmem_base;
};

template <int LP, int HDC>  // HDC = head_dim / 8 at compile time (0: runtime)
__global__ void __launch_bounds__(EfCfg<EncGeo<LP>::NG>::kThreads, 1)
attn_fwd_enc_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse,
                    int B, int L, int H, int hd, float scale, int dbg, long long* tl) {
  using Geo = EncGeo<LP>;
  constexpr int G = Geo::G, NG = Geo::NG, MS = Geo::MS;
  using Cfg = EfCfg<NG>;
  constexpr int kProd = 4 * NG, kMma = 4 * NG + 4, kPv = 4 * NG + 5;  // warps kProd..+3: loaders; kPv, kPv + 1: P V of the even / odd items
  if (threadIdx.x == 0) ENC_TS(3, 0);
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  EfBars* bars = reinterpret_cast<EfBars*>(smem + Cfg::bars);
  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const int ngroups = (B + G - 1) / G;
  const int nitems = ngroups * H;
  const int my_items = nitems > int(blockIdx.x) ? (nitems - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x) : 0;

  // zero the operand and P tiles once: rows past G * LP, and the off-diagonal blocks of P, are never written afterwards
  for (uint32_t i = threadIdx.x; i < Cfg::bars / 16; i += Cfg::kThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == kProd && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&bars->kv_full[i], 128); mbar_init(&bars->kv_empty[i], 1); }
    for (int i = 0; i < NG; ++i) {
      mbar_init(&bars->s_full[i], 1); mbar_init(&bars->s_free[i], 4);
      mbar_init(&bars->p_ready[i], 4);
      mbar_init(&bars->o_full[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == kMma) tmem_alloc(&bars->tmem_base, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) ENC_TS(3, 1);
  const uint32_t tmem_base = bars->tmem_base;
  ChunkTable<LP, 12> tab;  // loader warps: built before the grid dependency resolves (divisions; no global memory access)
  if (warp >= kProd && warp < kMma) {
    const int hdl = HDC ? HDC * 8 : hd;
    tab.init(int(threadIdx.x) - kProd * 32, 128, 3, L, hdl >> 3, uint32_t(3 * H * hdl) * 2u, uint32_t(H * hdl) * 2u, dbg);
  }
  pdl_wait();
  if (threadIdx.x == 0) ENC_TS(3, 2);

  if (warp >= kProd && warp < kMma) {
    // loader warps: Q / K / V chunks of the item by 16-byte async copies, kLoadAhead items in flight per thread; the
    // writer makes its completed copies visible to the async proxy (fence) before it arrives on the stage's barrier
    const int hdl = HDC ? HDC * 8 : hd;
    const uint32_t stage0 = smem_u32(smem + Cfg::qkv);
    const size_t grp_bytes = size_t(G) * L * 3 * H * hdl * 2;
    ItemIter it(int(blockIdx.x), int(gridDim.x), H);
    // event driven: issue up to kLoadAhead items ahead while stages are free, otherwise retire the oldest outstanding
    // group (its arrival must not wait for a stage that only frees once older items have been consumed)
    int issued = 0, retired = 0, s = 0, ph = 1, sd = 0;
    while (retired < my_items) {
      const int outstanding = issued - retired;
      bool go = issued < my_items && outstanding < Cfg::kLoadAhead;
      if (go) {
        if (outstanding == 0) mbar_wait(&bars->kv_empty[s], ph);
        else go = mbar_test_wait(&bars->kv_empty[s], ph);
      }
      if (go) {
        if (!((dbg & 2) && issued >= Cfg::kStages))
          tab.issue(stage0 + uint32_t(s) * 3u * kEncTile, reinterpret_cast<const char*>(qkv) + size_t(it.grp) * grp_bytes + size_t(it.h) * hdl * 2,
                    B - it.grp * G);
        cp_async_commit();
        it.next();
        ++issued;
        if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
      } else {
        switch (outstanding) {
          case 1: cp_async_wait<0>(); break;
          case 2: cp_async_wait<1>(); break;
          case 3: cp_async_wait<2>(); break;
          default: cp_async_wait<3>(); break;
        }
        fence_proxy_async_smem();
        mbar_arrive(&bars->kv_full[sd]);
        ++retired;
        if (++sd == Cfg::kStages) sd = 0;
      }
    }
  } else if (warp == kMma) {
    // One thread can issue a tcgen05.mma only every ~70 cycles (tools/probes/umma_issue_probe.cu), the tensor pipe takes a
    // small one every 40: the S MMAs and the P V MMAs of alternate items come from three different warps.
    {  // the whole warp runs the loop (uniform control flow, descriptors in uniform registers); one elected lane issues
      const uint32_t id_s = umma_idesc_bf16(128, Geo::NS, 0, 0);  // S: both operands K-major
      const uint64_t t64 = umma_sdesc_sw64(0, 16, 512);
      auto d_at = [](uint64_t templ, uint32_t addr) { return templ + uint64_t((addr & 0x3ffffu) >> 4); };
      const uint32_t tiles0 = smem_u32(smem + Cfg::qkv);
      for (int n = 0, st = 0, ph = 0, g = 0, k = 0; n < my_items; ++n) {
        mbar_wait(&bars->kv_full[st], ph);
        if (lane == 0) ENC_TS(0, 5 * n);
        if (k >= 1) mbar_wait(&bars->s_free[g], (k - 1) & 1);
        tc_fence_after();
        if (lane == 0) ENC_TS(0, 5 * n + 1);
        const uint32_t tb = tiles0 + uint32_t(st) * 3u * kEncTile;
        const uint64_t dq = d_at(t64, tb), dk = d_at(t64, tb + kEncTile);
        const uint32_t d = tmem_base + Cfg::colS + uint32_t(g) * 128u;
        if (elect_one_sync()) {
          umma_bf16(d, dq, dk, id_s, 0);
          umma_bf16(d, dq + 2, dk + 2, id_s, 1);  // +32 B per K = 16
          umma_commit(&bars->s_full[g]);
        }
        __syncwarp();
        if (lane == 0) ENC_TS(0, 5 * n + 2);
        if (++st == Cfg::kStages) { st = 0; ph ^= 1; }
        if (++g == NG) { g = 0; ++k; }
      }
    }
  } else if (warp >= kPv) {
    {
      const int par = warp - kPv;  // this issuer's items: n = par, par + 2, ...
      const uint32_t id_o = umma_idesc_bf16(128, 32, 0, 1);       // O: A = P K-major, B = V MN-major
      const uint64_t t64 = umma_sdesc_sw64(0, 16, 512);
      const uint64_t t128 = umma_sdesc_sw128(0, 16, 1024);
      auto d_at = [](uint64_t templ, uint32_t addr) { return templ + uint64_t((addr & 0x3ffffu) >> 4); };
      const uint32_t tiles0 = smem_u32(smem + Cfg::qkv), p0 = smem_u32(smem + Cfg::p);
      for (int n = par; n < my_items; n += 2) {
        const int st = n % Cfg::kStages, g = n % NG, k = n / NG;
        mbar_wait(&bars->kv_full[st], (n / Cfg::kStages) & 1);  // (complete long ago: makes the TMA writes visible to this thread's MMAs)
        mbar_wait(&bars->p_ready[g], k & 1);  // also: the group has drained O of its previous item (program order)
        tc_fence_after();
        if (par == 0 && lane == 0) ENC_TS(0, 5 * n + 3);
        const uint64_t a0 = d_at(t128, p0 + uint32_t(g) * kEncPTile);
        const uint64_t b0 = d_at(t64, tiles0 + uint32_t(st) * 3u * kEncTile + 2 * kEncTile);
        const uint32_t d = tmem_base + Cfg::colO + uint32_t(g) * 32u;
        if (elect_one_sync()) {
#pragma unroll
          for (int j = 0; j < Geo::KS; ++j)  // V as MN-major B: +1024 B per 16 keys
            umma_bf16(d, a0 + uint64_t((uint32_t(j >> 2) * 16384u + uint32_t(j & 3) * 32u) >> 4), b0 + uint64_t(64 * j), id_o, j != 0);
          umma_commit(&bars->o_full[g]);
          umma_commit(&bars->kv_empty[st]);
        }
        __syncwarp();
        if (par == 0 && lane == 0) ENC_TS(0, 5 * n + 4);
      }
    }
  } else {
    // softmax groups: warps 4 gi .. 4 gi + 3 = group gi, one TMEM lane quarter each; row = TMEM lane
    const int gi = warp >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int g = r / LP, l = r - g * LP;               // sample slot and token of this row
    const int g_lo = (quarter * 32) / LP < G ? (quarter * 32) / LP : G - 1;  // (rows past G * LP: dead lanes)
    const bool two = g_lo + 1 < G && (quarter * 32 + 31) / LP > g_lo;
    const bool second = g > g_lo;
    const bool row_ok = g < G && l < L;
    const float sl2 = scale * 1.4426950408889634f;
    const uint64_t sl2_2 = pk2(sl2, sl2);
    const uint32_t lane_off = uint32_t(quarter * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_off + Cfg::colS + uint32_t(gi) * 128u;
    const uint32_t o_addr = tmem_base + lane_off + Cfg::colO + uint32_t(gi) * 32u;
    const int c0 = (g * LP) >> 3;  // first 16-byte chunk of this sample's keys inside the P row
    const uint32_t p_row = smem_u32(smem + Cfg::p) + uint32_t(gi) * kEncPTile + uint32_t(r) * 128u;
    const int cnt = my_items > gi ? (my_items - gi + NG - 1) / NG : 0;
    const int D = HDC ? H * HDC * 8 : H * hd;
    const int hdv = HDC ? HDC * 8 : hd;
    ItemIter it(int(blockIdx.x) + gi * int(gridDim.x), NG * int(gridDim.x), H);
    float inv_prev = 0.f;
    __nv_bfloat16* dst_prev = nullptr;  // output row of the previous item (nullptr: nothing to store)
#pragma unroll 1
    for (int k = 0; k <= cnt; ++k) {
      uint32_t sr[LP];
      if (k < cnt) {
        mbar_wait(&bars->s_full[gi], k & 1);
        tc_fence_after();
        if (threadIdx.x == gi * 128) ENC_TS(1 + gi, 6 * k);
        tmem_ld_own<LP>(s_addr, g_lo, two, second, sr);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->s_free[gi]);
        if (threadIdx.x == gi * 128) ENC_TS(1 + gi, 6 * k + 1);
      }
      if (k >= 1) {  // O rows of the group's previous item: complete, because its P V was issued before this item's S
        mbar_wait(&bars->o_full[gi], (k - 1) & 1);
        tc_fence_after();
        uint32_t o[32];
        tmem_ld32_nowait(o_addr, o);
        tmem_wait_ld();
        tc_fence_before();
        if (dst_prev) {
          const uint64_t inv2 = pk2(inv_prev, inv_prev);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (HDC ? c >= HDC : c * 8 >= hd) break;
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float a, b;
              upk2(mul2(pk2u(o[8 * c + 2 * j], o[8 * c + 2 * j + 1]), inv2), a, b);
              w[j] = pack_bf16x2(a, b);
            }
            *reinterpret_cast<uint4*>(dst_prev + 8 * c) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      if (k < cnt) {
        if (threadIdx.x == gi * 128) ENC_TS(1 + gi, 6 * k + 2);
        const int b = it.grp * G + g, h = it.h;
        it.next();
        const bool ok = row_ok && b < B;
        if (dbg & 1) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->p_ready[gi]);
          inv_prev = __uint_as_float(sr[0]);
          dst_prev = ok ? out + (size_t(b) * L + l) * D + h * hdv : nullptr;
          continue;
        }
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < LP; ++i)
          if (i < MS || i < L) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(sr[i]));
        const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        const float msl2 = mx * sl2;
        const uint64_t nm2 = pk2(-msl2, -msl2);
        uint64_t acc[2] = {pk2(0.f, 0.f), pk2(0.f, 0.f)};
        uint32_t pk[LP / 2];
#pragma unroll
        for (int i = 0; i < LP / 2; ++i) {
          float a, b2;
          upk2(fma2(pk2u(sr[2 * i], sr[2 * i + 1]), sl2_2, nm2), a, b2);
          float p0 = ex2_approx(a), p1 = ex2_approx(b2);
          if (2 * i + 1 >= MS) {
            if (2 * i >= L) p0 = 0.f;
            if (2 * i + 1 >= L) p1 = 0.f;
          }
          acc[i & 1] = add2(acc[i & 1], pk2(p0, p1));
          pk[i] = pack_bf16x2(p0, p1);
        }
        float s0, s1;
        upk2(add2(acc[0], acc[1]), s0, s1);
        const float sum = s0 + s1;
        if (threadIdx.x == gi * 128) ENC_TS(1 + gi, 6 * k + 3);
        // (the P tile is free: o_full of the previous item, waited for above, is the completion of its P V)
        if (threadIdx.x == gi * 128) ENC_TS(1 + gi, 6 * k + 4);
        if (g < G) {
#pragma unroll
          for (int j = 0; j < LP / 8; ++j) {
            const int c = c0 + j;
            sts16(p_row + uint32_t(c >> 3) * 16384u + (uint32_t((c ^ r) & 7) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p_ready[gi]);
        if (threadIdx.x == gi * 128) ENC_TS(1 + gi, 6 * k + 5);
        if (ok) lse[(size_t(b) * H + h) * L + l] = (msl2 + lg2_approx(sum)) * 0.6931471805599453f;
        inv_prev = rcp_approx(sum);
        dst_prev = ok ? out + (size_t(b) * L + l) * D + h * hdv : nullptr;
      }
    }
  }
  if (threadIdx.x == 0) ENC_TS(3, 3);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) ENC_TS(3, 4);
  if (warp == kMma) tmem_dealloc(tmem_base, 512);
}

// =============================================================================================
// backward.  Per (group, head), with the forward's log-sum-exp:
//   S = Q K^T,  P = exp(S / sqrt(hd) - lse),  dP = dO V^T,  delta = rowsum(P o dP)  (= rowsum(dO o O) with O = P V),
//   dS = P o (dP - delta),  dQ = dS K / sqrt(hd),  dK = dS^T Q / sqrt(hd),  dV = P^T dO
// A softmax thread holds the whole S and dP rows of its sample, so delta comes from registers: O is never read.
// 1 / sqrt(hd) is folded into dS.  Q, K, V, dO tiles by TMA, three stages; S and dP share one TMEM region (a group
// pulls its rows into registers at once, which frees it for the next item); the bf16 P and dS tiles (block diagonal,
// one pair per group) feed dQ = dS K as K-major A and dK / dV as MN-major A.
// =============================================================================================
static constexpr int kEbThreads = 12 * 32;  // 2 softmax groups, TMA producer, three MMA issuers (S / dP + dQ, dK, dV)
static constexpr int kEbStages = 3;
struct EbBars {
  uint64_t ld_full[kEbStages], ld_empty[kEbStages];
  uint64_t sd_full[2], sd_free[2];     // [group]: S and dP complete / pulled out of TMEM
  uint64_t pds_ready[2];               // [group]: P and dS tiles written (and the previous item's gradients drained)
  uint64_t g_full[2];                  // [group]: dQ, dK, dV complete (three issuers) = the P / dS tiles have been read
  uint32_t tmem_base;
};
struct EbSmem {
  static constexpr uint32_t tiles = 0;                           // [3 stages][Q, K, V, dO]
  static constexpr uint32_t pds = kEbStages * 4 * kEncTile;      // [2 groups][P, dS]
  static constexpr uint32_t bars = pds + 2 * 2 * kEncPTile;
  static constexpr uint32_t total = bars + 1024;
};
static constexpr uint32_t kEbColS = 0, kEbColDP = 128, kEbColG = 256;  // per group at kEbColG + 96 g: dQ, dK, dV (32 each)

template <int LP, int HDC>
__global__ void __launch_bounds__(kEbThreads, 1)
attn_bwd_enc_kernel(const __grid_constant__ CUtensorMap mapQKV, const __grid_constant__ CUtensorMap mapDO,
                    const float* __restrict__ lse, __nv_bfloat16* __restrict__ dqkv, int B, int L, int H, int hd, float scale, int dbg) {
  using Geo = EncGeo<LP>;
  constexpr int G = Geo::G, MS = Geo::MS;
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  EbBars* bars = reinterpret_cast<EbBars*>(smem + EbSmem::bars);
  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const int ngroups = (B + G - 1) / G;
  const int nitems = ngroups * H;
  const int my_items = nitems > int(blockIdx.x) ? (nitems - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x) : 0;

  for (uint32_t i = threadIdx.x; i < EbSmem::bars / 16; i += kEbThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&mapQKV);
    tma_prefetch_desc(&mapDO);
    for (int i = 0; i < kEbStages; ++i) { mbar_init(&bars->ld_full[i], 1); mbar_init(&bars->ld_empty[i], 3); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->sd_full[i], 1); mbar_init(&bars->sd_free[i], 4);
      mbar_init(&bars->pds_ready[i], 4);
      mbar_init(&bars->g_full[i], 3);
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
    {  // whole warp, uniform control flow; one elected lane issues the TMA loads
      ItemIter it(int(blockIdx.x), int(gridDim.x), H);
      for (int n = 0, s = 0, ph = 1; n < my_items; ++n, it.next()) {
        mbar_wait(&bars->ld_empty[s], ph);
        const uint32_t dst = smem_u32(smem + EbSmem::tiles + s * 4 * kEncTile);
        if (elect_one_sync()) {
          if ((dbg & 2) && n >= kEbStages) {
            mbar_arrive(&bars->ld_full[s]);
          } else {
            mbar_expect_tx(&bars->ld_full[s], 4 * Geo::kBox);
            tma_load_4d(dst, &mapQKV, &bars->ld_full[s], 0, it.h, 0, it.grp * G);
            tma_load_4d(dst + kEncTile, &mapQKV, &bars->ld_full[s], 0, H + it.h, 0, it.grp * G);
            tma_load_4d(dst + 2 * kEncTile, &mapQKV, &bars->ld_full[s], 0, 2 * H + it.h, 0, it.grp * G);
            tma_load_4d(dst + 3 * kEncTile, &mapDO, &bars->ld_full[s], 0, it.h, 0, it.grp * G);
          }
        }
        __syncwarp();
        if (++s == kEbStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp >= 9) {
    // Three issuers (one thread issues a tcgen05.mma only every ~70 cycles, the tensor pipe takes a small one every 40):
    // warp 9: S / dP of the next item, then dQ;  warp 10: dK;  warp 11: dV.  Each commits to g_full and ld_empty.
    {  // whole warp, uniform control flow; one elected lane issues
      const int role = warp - 9;
      const uint32_t id_s = umma_idesc_bf16(128, Geo::NS, 0, 0);  // S, dP: both operands K-major
      const uint32_t id_q = umma_idesc_bf16(128, 32, 0, 1);       // dQ: A K-major, B MN-major
      const uint32_t id_kv = umma_idesc_bf16(128, 32, 1, 1);      // dK, dV: both MN-major
      const uint64_t t64 = umma_sdesc_sw64(0, 16, 512);
      const uint64_t t128k = umma_sdesc_sw128(0, 16, 1024);        // P / dS tile as K-major A
      const uint64_t t128mn = umma_sdesc_sw128(0, 16384, 1024);    // ... as MN-major A (LBO = stride between the 64-key boxes)
      auto d_at = [](uint64_t templ, uint32_t addr) { return templ + uint64_t((addr & 0x3ffffu) >> 4); };
      const uint32_t tiles0 = smem_u32(smem + EbSmem::tiles), pds0 = smem_u32(smem + EbSmem::pds);
      if (role == 0) {
        auto issue_sdp = [&](int st, int gi) {
          const uint32_t tb = tiles0 + uint32_t(st) * 4u * kEncTile;
          const uint64_t dq = d_at(t64, tb), dk = d_at(t64, tb + kEncTile), dv = d_at(t64, tb + 2 * kEncTile),
                         ddo = d_at(t64, tb + 3 * kEncTile);
          if (elect_one_sync()) {
            umma_bf16(tmem_base + kEbColS, dq, dk, id_s, 0);
            umma_bf16(tmem_base + kEbColS, dq + 2, dk + 2, id_s, 1);
            umma_bf16(tmem_base + kEbColDP, ddo, dv, id_s, 0);
            umma_bf16(tmem_base + kEbColDP, ddo + 2, dv + 2, id_s, 1);
            umma_commit(&bars->sd_full[gi]);
          }
          __syncwarp();
        };
        int st_next = 0, ph_next = 0;  // stage / phase of the next item whose S / dP will be issued
        if (my_items > 0) {
          mbar_wait(&bars->ld_full[0], 0);
          tc_fence_after();
          issue_sdp(0, 0);
          st_next = 1;
        }
        int st = 0;
        for (int n = 0; n < my_items; ++n) {
          const int gi = n & 1;
          if (n + 1 < my_items) {  // the next item's S / dP as soon as this item's rows are in registers
            mbar_wait(&bars->ld_full[st_next], ph_next);
            mbar_wait(&bars->sd_free[gi], (n >> 1) & 1);
            tc_fence_after();
            issue_sdp(st_next, gi ^ 1);
            if (++st_next == kEbStages) { st_next = 0; ph_next ^= 1; }
          }
          mbar_wait(&bars->pds_ready[gi], (n >> 1) & 1);  // also: the group has drained dQ / dK / dV of its previous item
          tc_fence_after();
          {  // dQ = dS . K   (B = K tile as MN-major: +1024 B per 16 keys)
            const uint64_t a0 = d_at(t128k, pds0 + uint32_t(gi) * 2u * kEncPTile + kEncPTile);
            const uint64_t b0 = d_at(t64, tiles0 + uint32_t(st) * 4u * kEncTile + kEncTile);
            const uint32_t dcol = tmem_base + kEbColG + uint32_t(gi) * 96u;
            if (elect_one_sync()) {
#pragma unroll
              for (int j = 0; j < Geo::KS; ++j)
                umma_bf16(dcol, a0 + uint64_t((uint32_t(j >> 2) * 16384u + uint32_t(j & 3) * 32u) >> 4), b0 + uint64_t(64 * j), id_q, j != 0);
              umma_commit(&bars->g_full[gi]);
              umma_commit(&bars->ld_empty[st]);
            }
            __syncwarp();
          }
          if (++st == kEbStages) st = 0;
        }
      } else {
        // dK = dS^T . Q  (role 1),  dV = P^T . dO  (role 2):  M = keys; K = query rows: A +2048 B, B +1024 B per 16 rows
        for (int n = 0, st = 0, ph = 0; n < my_items; ++n) {
          const int gi = n & 1;
          mbar_wait(&bars->ld_full[st], ph);  // (complete long ago: makes the TMA writes visible to this thread's MMAs)
          mbar_wait(&bars->pds_ready[gi], (n >> 1) & 1);
          tc_fence_after();
          const uint32_t tb = tiles0 + uint32_t(st) * 4u * kEncTile;
          const uint32_t sp = pds0 + uint32_t(gi) * 2u * kEncPTile;
          const uint64_t a = d_at(t128mn, role == 1 ? sp + kEncPTile : sp), b = d_at(t64, role == 1 ? tb : tb + 3 * kEncTile);
          const uint32_t dcol = tmem_base + kEbColG + uint32_t(gi) * 96u + (role == 1 ? 32u : 64u);
          if (elect_one_sync()) {
#pragma unroll
            for (int s = 0; s < Geo::KS; ++s) umma_bf16(dcol, a + uint64_t(128 * s), b + uint64_t(64 * s), id_kv, s != 0);
            umma_commit(&bars->g_full[gi]);
            umma_commit(&bars->ld_empty[st]);
          }
          __syncwarp();
          if (++st == kEbStages) { st = 0; ph ^= 1; }
        }
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
    const uint64_t sl2_2 = pk2(sl2, sl2), sc2 = pk2(scale, scale);
    const uint32_t lane_off = uint32_t(quarter * 32) << 16;
    const uint32_t g_addr = tmem_base + lane_off + kEbColG + uint32_t(gi) * 96u;
    const int c0 = (g * LP) >> 3;
    const uint32_t p_row = smem_u32(smem + EbSmem::pds) + uint32_t(gi) * 2u * kEncPTile + uint32_t(r) * 128u;
    const int cnt = my_items > gi ? (my_items - gi + 1) / 2 : 0;
    const int D = HDC ? H * HDC * 8 : H * hd;
    const int hdv = HDC ? HDC * 8 : hd;
    ItemIter it(int(blockIdx.x) + gi * int(gridDim.x), 2 * int(gridDim.x), H);
    __nv_bfloat16* dst_prev = nullptr;
    float lse2 = INFINITY;
    {
      const int b = it.grp * G + g;
      if (cnt > 0 && row_ok && b < B) lse2 = __ldg(lse + (size_t(b) * H + it.h) * L + l) * l2e;
    }
#pragma unroll 1
    for (int k = 0; k <= cnt; ++k) {
      uint32_t sr[LP], dr[LP];
      if (k < cnt) {
        mbar_wait(&bars->sd_full[gi], k & 1);
        tc_fence_after();
        tmem_ld_own<LP>(tmem_base + lane_off + kEbColS, g_lo, two, second, sr);
        tmem_ld_own<LP>(tmem_base + lane_off + kEbColDP, g_lo, two, second, dr);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->sd_free[gi]);
      }
      if (k >= 1) {  // gradients of the group's previous item: complete (issued before this item's S / dP)
        mbar_wait(&bars->g_full[gi], (k - 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int w = 0; w < 3; ++w) {  // dQ, dK, dV one after the other: 32 live registers next to the S / dP rows
          uint32_t o[32];
          tmem_ld32_nowait(g_addr + uint32_t(w) * 32u, o);
          tmem_wait_ld();
          if (dst_prev) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (HDC ? c >= HDC : c * 8 >= hd) break;
              uint4 v;
              v.x = pack_bf16x2(__uint_as_float(o[8 * c]), __uint_as_float(o[8 * c + 1]));
              v.y = pack_bf16x2(__uint_as_float(o[8 * c + 2]), __uint_as_float(o[8 * c + 3]));
              v.z = pack_bf16x2(__uint_as_float(o[8 * c + 4]), __uint_as_float(o[8 * c + 5]));
              v.w = pack_bf16x2(__uint_as_float(o[8 * c + 6]), __uint_as_float(o[8 * c + 7]));
              *reinterpret_cast<uint4*>(dst_prev + w * D + 8 * c) = v;
            }
          }
        }
        tc_fence_before();
      }
      if (k < cnt) {
        const int b = it.grp * G + g, h = it.h;
        it.next();
        const bool ok = row_ok && b < B;
        const uint64_t nl2 = pk2(-lse2, -lse2);
        // the next item's log-sum-exp: in flight during the arithmetic below
        float lse_next = INFINITY;
        {
          const int bn = it.grp * G + g;
          if (k + 1 < cnt && row_ok && bn < B) lse_next = __ldg(lse + (size_t(bn) * H + it.h) * L + l);
        }
        if (dbg & 1) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->pds_ready[gi]);
          dst_prev = ok ? dqkv + (size_t(b) * L + l) * 3 * D + h * hdv + (sr[0] == 0x12345u) + (dr[0] == 0x12345u) : nullptr;
          lse2 = lse_next * l2e + __uint_as_float(nl2 >> 40) * 0.f;
          continue;
        }
        uint64_t dacc[2] = {pk2(0.f, 0.f), pk2(0.f, 0.f)};
#pragma unroll
        for (int i = 0; i < LP / 2; ++i) {  // S -> P in place (fp32), delta accumulated on the way
          float a, b2;
          upk2(fma2(pk2u(sr[2 * i], sr[2 * i + 1]), sl2_2, nl2), a, b2);
          float p0 = ex2_approx(a), p1 = ex2_approx(b2);
          if (2 * i + 1 >= MS) {
            if (2 * i >= L) p0 = 0.f;
            if (2 * i + 1 >= L) p1 = 0.f;
          }
          sr[2 * i] = __float_as_uint(p0);
          sr[2 * i + 1] = __float_as_uint(p1);
          dacc[i & 1] = fma2(pk2(p0, p1), pk2u(dr[2 * i], dr[2 * i + 1]), dacc[i & 1]);
        }
        float d0, d1;
        upk2(add2(dacc[0], dacc[1]), d0, d1);
        const float nds = -(d0 + d1) * scale;  // - delta / sqrt(hd)
        const uint64_t nds2 = pk2(nds, nds);
        // (the P / dS tiles are free: g_full of the previous item, waited for above, is the completion of the MMAs reading them)
        if (g < G) {
#pragma unroll
          for (int j = 0; j < LP / 8; ++j) {
            const int c = c0 + j;
            const uint32_t a = p_row + uint32_t(c >> 3) * 16384u + (uint32_t((c ^ r) & 7) << 4);
            uint32_t pw[4], dk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {  // dS / sqrt(hd) = P o (dP / sqrt(hd) - delta / sqrt(hd))
              const int i = 4 * j + q;
              pw[q] = pack_bf16x2(__uint_as_float(sr[2 * i]), __uint_as_float(sr[2 * i + 1]));
              float x, y;
              upk2(mul2(pk2u(sr[2 * i], sr[2 * i + 1]), fma2(pk2u(dr[2 * i], dr[2 * i + 1]), sc2, nds2)), x, y);
              dk[q] = pack_bf16x2(x, y);
            }
            sts16(a, pw[0], pw[1], pw[2], pw[3]);
            sts16(a + kEncPTile, dk[0], dk[1], dk[2], dk[3]);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->pds_ready[gi]);
        dst_prev = ok ? dqkv + (size_t(b) * L + l) * 3 * D + h * hdv : nullptr;
        lse2 = lse_next * l2e;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 512);
}

// MAE_ENC_DBG (measurement only, results are wrong): 1 = softmax threads skip the arithmetic, 2 = no TMA after the first stages
static int enc_dbg() {
  static const int v = getenv("MAE_ENC_DBG") ? atoi(getenv("MAE_ENC_DBG")) : 0;
  return v;
}

static long long* enc_timeline() {
  static long long* v = getenv("MAE_ENC_TIMELINE") ? reinterpret_cast<long long*>(strtoull(getenv("MAE_ENC_TIMELINE"), nullptr, 0)) : nullptr;
  return v;
}

template <int LP, int HDC>
static int launch_fwd_enc(const void* qkv, int B, int L, int H, int hd, void* out, float* lse, cudaStream_t st) {
  using Geo = EncGeo<LP>;
  using Cfg = EfCfg<Geo::NG>;
  if (reinterpret_cast<uintptr_t>(qkv) & 15) return MAE_ERR_ALIGN;
  static bool done[kMaxDevices] = {};
  if (!ensure_dyn_smem(attn_fwd_enc_kernel<LP, HDC>, Cfg::total + 1024, done)) return MAE_ERR_LAUNCH;
  const int sms = num_sms();
  const int nitems = ((B + Geo::G - 1) / Geo::G) * H;
  const int grid = nitems < sms ? nitems : sms;
  mae_launch(attn_fwd_enc_kernel<LP, HDC>, dim3(grid), dim3(Cfg::kThreads), size_t(Cfg::total) + 1024, st,
             reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), lse, B, L, H, hd, 1.0f / sqrtf(float(hd)), enc_dbg(), enc_timeline());
  return MAE_CHECK_LAUNCH();
}

template <int LP, int HDC>
static int launch_bwd_enc(const void* qkv, const void* dout, const float* lse, int B, int L, int H, int hd, void* dqkv,
                          cudaStream_t st) {
  using Geo = EncGeo<LP>;
  CUtensorMap mq, md;
  int rc = make_tmap_enc(&mq, qkv, hd, 3 * H, L, B, LP, Geo::G);
  if (rc) return rc;
  rc = make_tmap_enc(&md, dout, hd, H, L, B, LP, Geo::G);
  if (rc) return rc;
  static bool done[kMaxDevices] = {};
  if (!ensure_dyn_smem(attn_bwd_enc_kernel<LP, HDC>, EbSmem::total + 1024, done)) return MAE_ERR_LAUNCH;
  const int sms = num_sms();
  const int nitems = ((B + Geo::G - 1) / Geo::G) * H;
  const int grid = nitems < sms ? nitems : sms;
  mae_launch(attn_bwd_enc_kernel<LP, HDC>, dim3(grid), dim3(kEbThreads), size_t(EbSmem::total) + 1024, st, mq, md, lse,
             reinterpret_cast<__nv_bfloat16*>(dqkv), B, L, H, hd, 1.0f / sqrtf(float(hd)), enc_dbg());
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

// head_dim 24 (configs/mae.yaml) gets compile-time chunk counts, every other head_dim the runtime loop bounds
#define MAE_ENC_CASE(FN, LPV, ...) \
  case LPV: return hd == 24 ? FN<LPV, 3>(__VA_ARGS__) : FN<LPV, 0>(__VA_ARGS__);

int attn_fwd_enc(const void* qkv, int B, int L, int H, int hd, void* out, float* lse, cudaStream_t st) {
  switch (enc_pitch(L)) {
    MAE_ENC_CASE(launch_fwd_enc, 16, qkv, B, L, H, hd, out, lse, st)
    MAE_ENC_CASE(launch_fwd_enc, 32, qkv, B, L, H, hd, out, lse, st)
    MAE_ENC_CASE(launch_fwd_enc, 40, qkv, B, L, H, hd, out, lse, st)
    MAE_ENC_CASE(launch_fwd_enc, 48, qkv, B, L, H, hd, out, lse, st)
    MAE_ENC_CASE(launch_fwd_enc, 64, qkv, B, L, H, hd, out, lse, st)
    default: return MAE_ERR_SHAPE;
  }
}

int attn_bwd_enc(const void* qkv, const void* out, const void* dout, const float* lse, int B, int L, int H, int hd,
                 void* dqkv, cudaStream_t st) {
  (void)out;  // delta = rowsum(P o dP) comes from registers: the forward output is not read
  switch (enc_pitch(L)) {
    MAE_ENC_CASE(launch_bwd_enc, 16, qkv, dout, lse, B, L, H, hd, dqkv, st)
    MAE_ENC_CASE(launch_bwd_enc, 32, qkv, dout, lse, B, L, H, hd, dqkv, st)
    MAE_ENC_CASE(launch_bwd_enc, 40, qkv, dout, lse, B, L, H, hd, dqkv, st)
    MAE_ENC_CASE(launch_bwd_enc, 48, qkv, dout, lse, B, L, H, hd, dqkv, st)
    MAE_ENC_CASE(launch_bwd_enc, 64, qkv, dout, lse, B, L, H, hd, dqkv, st)
    default: return MAE_ERR_SHAPE;
  }
}

}  // namespace mae
