// tcgen05 GEMMs for the MAE step (sm_100a).
//
//   gemm_tn     C[M,N]  = epi( A[M,K] * B[N,K]^T )      forward linears and dgrads (both operands K-major)
//   gemm_wgrad  O[P,Q] += sum_m  P[m,:]^T Q[m,:]        weight gradients (both operands MN-major)
//
// Shared structure: warp 0 = TMA producer (one lane), warp 1 = TMEM owner + MMA issuer (one lane), the
// remaining warps = epilogue (TMEM -> registers -> global).  gemm_tn runs 16 epilogue warps: a warp may only
// read the TMEM lane quarter (warp_id % 4), so the four warps of a quarter interleave the 16-column chunks.  Operand tiles are staged by TMA with the
// 128-byte swizzle in a 4-deep mbarrier ring; accumulators live in TMEM (fp32), double buffered for
// gemm_tn so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "gemm_common.cuh"

namespace mae {

static constexpr int kStages = 4;
static constexpr int kEpiGroups = 4;                               // epilogue warps per TMEM lane quarter
static constexpr int kGemmThreads = 64 + 4 * kEpiGroups * 32;       // gemm_tn: TMA + MMA + 16 epilogue warps
static constexpr int kWgradThreads = 192;                            // gemm_wgrad: TMA + MMA + 4 epilogue warps

// ---------------------------------------------------------------------------------------------
// host: TMA descriptor helper
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D bf16 row-major tensor [rows, cols] with leading dimension ld (elements); box = [box_rows, 64 cols]
int make_tmap_bf16(CUtensorMap* map, const void* ptr, int rows, int cols, int ld, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return MAE_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((size_t(ld) * 2) & 15) || box_rows > 256 || box_rows < 1)
    return MAE_ERR_ALIGN;
  cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  cuuint64_t strides[1] = {cuuint64_t(ld) * 2};
  cuuint32_t box[2] = {cuuint32_t(BK), cuuint32_t(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MAE_OK : MAE_ERR_DRIVER;
}

// ---------------------------------------------------------------------------------------------
// gemm_tn
// ---------------------------------------------------------------------------------------------
struct GemmTnParams {
  int M, N, K;
  int bn;       // N chunk per work item (multiple of 16, <= 256, divides N)
  int nchunks;  // N / bn
  int mtiles;   // ceil(M / 128)
  int nkb;      // ceil(K / 64)
  int nstages;  // TMA ring depth (2..4, whatever fits next to the epilogue staging tiles)
  int use_tma_store;  // epilogue tiles move through TMA (off: MAE_GEMM_NO_TMA_STORE=1, A/B experiments)
  // shared-memory carve-up (byte offsets from the 1024-aligned base; computed on the host, see plan_smem)
  uint32_t off_ring, off_hdr, off_tab;  // operand ring, barriers, GELU table (EPI 1 only)
  uint32_t stg_a0, stg_b0;              // staging tiles [0, stg_nA) start at stg_a0, the rest at stg_b0
  int stg_nA;
  uint32_t stg_tile;                    // bytes per staging tile
  int stg_tpw;                          // staging tiles per epilogue warp (1, or 2 for the dGELU TMA epilogue)
  const float* bias;          // [N] or null
  void* out;                  // bf16 [M,N] (EPI 0,1,3,4)  or  f32 [M,N] (EPI 2)
  __nv_bfloat16* pre_out;     // EPI 1: bf16 [M,N] GELU'(pre-activation), consumed by EPI 4
  const void* aux;            // EPI 2: f32 residual [M,N];  EPI 4: bf16 GELU' [M,N] saved by EPI 1
};

struct GemmSmemHdr {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint64_t auxbar[4 * 4][2];  // dGELU epilogue: one per (epilogue warp, staging tile): "saved GELU' tile has landed"
  uint32_t tmem_base;
};

// ---------------------------------------------------------------------------------------------
// Epilogue.  An epilogue warp owns TMEM lanes [32q, 32q+32) = 32 output rows, one row per thread.  To keep
// every global access a full 128-byte line, a warp works in units of 128 output bytes per row (64 bf16 or 32
// fp32 columns): the row-per-thread values go through a private padded staging tile in shared memory and
// leave (or, for aux inputs, arrive) with 8 lanes covering one row, i.e. 4 complete rows per instruction.
// ---------------------------------------------------------------------------------------------
// Staging tiles per epilogue warp.  Manual path: one 32 x 144 B tile.  TMA epilogues: 1024-byte aligned tiles (128-byte
// swizzle): plain bf16 / GELU one 5 KB tile (4 KB swizzled unit, or a ragged 32 x 144 B unit), dGELU two 4 KB tiles
// (the saved GELU' tile of the next item is prefetched by TMA while the current one is multiplied and stored).
// Kept as small as possible: dynamic smem comes out of the L1 carve-out.
__host__ __device__ constexpr int stg_tile_bytes(int epi, bool tma) {
  return !tma ? 32 * kStgRow : (epi == MAE_EPI_DGELU_BF16 ? 4 * 1024 : 5 * 1024);
}
__host__ __device__ constexpr int stg_tiles_per_warp(int epi, bool tma) { return (tma && epi == MAE_EPI_DGELU_BF16) ? 2 : 1; }

__device__ uint32_t g_gelu_tab[2 * kGeluNT];

// Coalesced move of a [32 rows x row_bytes] tile between global memory (row pitch gpitch bytes) and the staging
// tile.  row_bytes is a multiple of 16 and <= 128: (row_bytes/16) lanes cover one row, so one instruction moves
// 32/(row_bytes/16) >= 4 complete rows.  Rows >= rows_valid are skipped.  All global loads of a tile are issued
// before the first use (fixed 8-way unroll) so their latencies overlap instead of serialising.
template <bool TO_GLOBAL>
__device__ __forceinline__ void stg_copy(uint8_t* stg, uint8_t* gbase, size_t gpitch, int row_bytes, int rows_valid,
                                         int lane) {
  const int lpr = row_bytes >> 4;            // lanes per row (1..8)
  const int rpi = 32 / lpr;                  // rows per instruction (>= 4)
  const int rr = lane / lpr, ch = lane % lpr;
  const bool lane_ok = rr < rpi;
  if (TO_GLOBAL) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * rpi + rr;
      if (lane_ok && r < 32 && r < rows_valid)
        *reinterpret_cast<uint4*>(gbase + size_t(r) * gpitch + ch * 16) =
            *reinterpret_cast<const uint4*>(stg + r * kStgRow + ch * 16);
    }
  } else {
    uint4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * rpi + rr;
      v[i] = make_uint4(0, 0, 0, 0);
      if (lane_ok && r < 32 && r < rows_valid) v[i] = *reinterpret_cast<const uint4*>(gbase + size_t(r) * gpitch + ch * 16);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * rpi + rr;
      if (lane_ok && r < 32) *reinterpret_cast<uint4*>(stg + r * kStgRow + ch * 16) = v[i];
    }
  }
}

// fp32 variant with the residual added on the way out: out = staged + resid
__device__ __forceinline__ void stg_flush_add_f32(const uint8_t* stg, const uint8_t* rbase, uint8_t* obase,
                                                  size_t gpitch, int row_bytes, int rows_valid, int lane) {
  const int lpr = row_bytes >> 4, rpi = 32 / lpr;
  const int rr = lane / lpr, ch = lane % lpr;
  const bool lane_ok = rr < rpi;
  float4 x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * rpi + rr;
    x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane_ok && r < 32 && r < rows_valid) x[i] = *reinterpret_cast<const float4*>(rbase + size_t(r) * gpitch + ch * 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * rpi + rr;
    if (lane_ok && r < 32 && r < rows_valid) {
      const float4 a = *reinterpret_cast<const float4*>(stg + r * kStgRow + ch * 16);
      *reinterpret_cast<float4*>(obase + size_t(r) * gpitch + ch * 16) =
          make_float4(a.x + x[i].x, a.y + x[i].y, a.z + x[i].z, a.w + x[i].w);
    }
  }
}

// Manual-path epilogue (fp32 residual and dGELU): one work item (128 x bn accumulator tile) for one epilogue warp.  row0 = first global row of this warp's 32.
template <int EPI>
__device__ __forceinline__ void epilogue_item_manual(const GemmTnParams& p, uint32_t taddr, int row0, int col_base, int grp,
                                              uint8_t* stg, int lane) {
  constexpr bool F32 = (EPI == MAE_EPI_RESID_F32);
  constexpr int UC = F32 ? 32 : 64;          // columns per 128-byte unit
  constexpr int SB = F32 ? 64 : 32;          // bytes per 16-column sub-chunk
  constexpr int ES = F32 ? 4 : 2;            // output element size
  const int rows_valid = p.M - row0;         // may be <= 0 or > 32
  // Balanced split of the bn/16 sub-chunks over the kEpiGroups warps of this lane quarter (contiguous ranges,
  // e.g. bn=192: 3+3+3+3, bn=144: 3+2+2+2), each range processed in units of at most 128 bytes per row.
  const int nsc = p.bn >> 4;
  const int sc_begin = (nsc * grp) / kEpiGroups, sc_end = (nsc * (grp + 1)) / kEpiGroups;
  for (int sc = sc_begin; sc < sc_end; sc += UC / 16) {
    const int c0 = sc * 16;
    int nsub = sc_end - sc;
    if (nsub > UC / 16) nsub = UC / 16;
    const int row_bytes = nsub * SB;
    const int gcol = col_base + c0;
    const size_t gpitch = size_t(p.N) * ES;
    if constexpr (EPI == MAE_EPI_DGELU_BF16) {  // bring the saved GELU' tile in, coalesced
      uint8_t* abase = const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(p.aux)) + (size_t(row0) * p.N + gcol) * 2;
      stg_copy<false>(stg, abase, gpitch, row_bytes, rows_valid, lane);
      __syncwarp();
    }
    for (int sidx = 0; sidx < nsub; ++sidx) {
      uint32_t raw[16];
      tmem_ld16_issue(taddr + c0 + sidx * 16, raw);
      float bia[16];
      if (p.bias != nullptr) {
        const float4* bp = reinterpret_cast<const float4*>(p.bias + gcol + sidx * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t4 = __ldg(bp + i);
          bia[4 * i] = t4.x; bia[4 * i + 1] = t4.y; bia[4 * i + 2] = t4.z; bia[4 * i + 3] = t4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) bia[i] = 0.f;
      }
      uint8_t* my = stg + lane * kStgRow + sidx * SB;
      uint32_t pr[8];
      if constexpr (EPI == MAE_EPI_DGELU_BF16) {
        const uint4 a = *reinterpret_cast<const uint4*>(my), b = *reinterpret_cast<const uint4*>(my + 16);
        pr[0] = a.x; pr[1] = a.y; pr[2] = a.z; pr[3] = a.w; pr[4] = b.x; pr[5] = b.y; pr[6] = b.z; pr[7] = b.w;
      }
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]) + bia[i];
      if constexpr (F32) {
        float4* d = reinterpret_cast<float4*>(my);
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      } else {
        uint32_t o[8];
        if constexpr (EPI == MAE_EPI_BF16) {
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        } else {  // DGELU
          static_assert(EPI != MAE_EPI_GELU_BF16, "the GELU epilogue lives in epilogue_item");
#pragma unroll
          for (int i = 0; i < 8; ++i)  // pr = GELU'(pre) saved by the forward
            o[i] = pack_bf16x2(v[2 * i] * bf16_lo(pr[i]), v[2 * i + 1] * bf16_hi(pr[i]));
        }
        uint4* d = reinterpret_cast<uint4*>(my);
        d[0] = make_uint4(o[0], o[1], o[2], o[3]);
        d[1] = make_uint4(o[4], o[5], o[6], o[7]);
      }
    }
    __syncwarp();
    const size_t goff = (size_t(row0) * p.N + gcol) * ES;
    if constexpr (F32) {
      stg_flush_add_f32(stg, reinterpret_cast<const uint8_t*>(p.aux) + goff, reinterpret_cast<uint8_t*>(p.out) + goff,
                        gpitch, row_bytes, rows_valid, lane);
    } else {
      stg_copy<true>(stg, reinterpret_cast<uint8_t*>(p.out) + goff, gpitch, row_bytes, rows_valid, lane);
    }
    __syncwarp();
  }
}

// One work item (128 x bn accumulator tile) for one epilogue warp.  row0 = first global row of this warp's 32.
// bf16 outputs: the tile is split into 64-column units dealt round-robin to the kEpiGroups warps of the lane quarter
// (rotated by `rot` = per-CTA item counter so that e.g. 3 units keep all 4 warps busy on average).  A full unit is
// written row-per-thread into a 128-byte-swizzled smem tile and leaves through one TMA bulk tensor store (no smem
// read-back, no per-lane global stores: the LSU only sees the staging writes); ragged units (bn % 64) and the fp32
// residual epilogue use the manual coalesced path.
template <int EPI>
__device__ __forceinline__ void epilogue_item(const GemmTnParams& p, uint32_t taddr, int row0, int col_base, int grp,
                                              int rot, uint8_t* stg, int lane, const CUtensorMap* mo,
                                              const CUtensorMap* mo2, uint32_t tab_s) {
  constexpr bool F32 = (EPI == MAE_EPI_RESID_F32);
  constexpr int UC = F32 ? 32 : 64;          // columns per 128-byte unit
  constexpr int SB = F32 ? 64 : 32;          // bytes per 16-column sub-chunk
  constexpr int ES = F32 ? 4 : 2;            // output element size
  const int rows_valid = p.M - row0;         // may be <= 0 or > 32
  const uint32_t stg_s = smem_u32(stg);
  // Two ways of dealing the tile's 16-column sub-chunks to the kEpiGroups warps of this lane quarter:
  //  TMA epilogues (plain bf16, GELU): whole 64-column units, round-robin, rotated by the item counter;
  //  manual epilogues (fp32 residual, dGELU): balanced contiguous ranges (bn=192: 3+3+3+3), <= 128 B per row at a time.
  constexpr bool kTmaEpi = (EPI == MAE_EPI_BF16 || EPI == MAE_EPI_GELU_BF16);
  const bool rr = kTmaEpi && p.use_tma_store;
  const int nsc = p.bn >> 4;
  const int sc_begin = rr ? 0 : (nsc * grp) / kEpiGroups, sc_end = rr ? nsc : (nsc * (grp + 1)) / kEpiGroups;
  for (int sc = sc_begin, u = 0; sc < sc_end; sc += UC / 16, ++u) {
    if (rr && ((u + rot) & (kEpiGroups - 1)) != grp) continue;
    const int c0 = sc * 16;
    int nsub = sc_end - sc;
    if (nsub > UC / 16) nsub = UC / 16;
    const bool tma = rr && nsub == 4;  // warp-uniform
    const int row_bytes = nsub * SB;
    const int gcol = col_base + c0;
    const size_t gpitch = size_t(p.N) * ES;
    if (rr) {  // the previous TMA store of this warp must have finished READING the staging tile
      if (lane == 0) tma_store_wait_read();
      __syncwarp();
    }
    uint32_t prk[UC / 16][8];  // EPI_DGELU: the saved GELU' values of this thread's row
    if constexpr (EPI == MAE_EPI_DGELU_BF16) {  // bring the saved GELU' tile in, coalesced, then row-per-thread
      uint8_t* abase = const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(p.aux)) + (size_t(row0) * p.N + gcol) * 2;
      stg_copy<false>(stg, abase, gpitch, row_bytes, rows_valid, lane);
      __syncwarp();
#pragma unroll
      for (int sidx = 0; sidx < UC / 16; ++sidx)
        if (sidx < nsub) {
          const uint4 a = *reinterpret_cast<const uint4*>(stg + lane * kStgRow + sidx * SB);
          const uint4 b = *reinterpret_cast<const uint4*>(stg + lane * kStgRow + sidx * SB + 16);
          prk[sidx][0] = a.x; prk[sidx][1] = a.y; prk[sidx][2] = a.z; prk[sidx][3] = a.w;
          prk[sidx][4] = b.x; prk[sidx][5] = b.y; prk[sidx][6] = b.z; prk[sidx][7] = b.w;
        }
      __syncwarp();  // everyone has its inputs: the tile may now be overwritten (possibly in the swizzled layout)
    }
    uint32_t hkeep[UC / 16][8];  // EPI_GELU: second output (h) parked in registers while GELU' is flushed
#pragma unroll
    for (int sidx = 0; sidx < UC / 16; ++sidx) {
      if (sidx >= nsub) break;
      uint32_t raw[16];
      tmem_ld16_issue(taddr + c0 + sidx * 16, raw);
      float bia[16];
      if (p.bias != nullptr) {
        const float4* bp = reinterpret_cast<const float4*>(p.bias + gcol + sidx * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t4 = __ldg(bp + i);
          bia[4 * i] = t4.x; bia[4 * i + 1] = t4.y; bia[4 * i + 2] = t4.z; bia[4 * i + 3] = t4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) bia[i] = 0.f;
      }
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]) + bia[i];
      if constexpr (F32) {
        float4* d = reinterpret_cast<float4*>(stg + lane * kStgRow + sidx * SB);
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      } else {
        uint32_t o[8];
        if constexpr (EPI == MAE_EPI_BF16) {
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        } else if constexpr (EPI == MAE_EPI_GELU_BF16) {
          // autocast semantics: the linear's output is rounded to bf16 and GELU reads that bf16 value.  GELU and
          // GELU' of a bf16 input are table look-ups (exactly rounded); GELU' is what the backward needs, so it is
          // saved in place of the pre-activation and the dgrad epilogue is a plain multiply.
          uint32_t pre2[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) pre2[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
          gelu_table8(pre2, tab_s, hkeep[sidx], o);
        } else {  // DGELU
#pragma unroll
          for (int i = 0; i < 8; ++i)  // prk = GELU'(pre) saved by the forward
            o[i] = pack_bf16x2(v[2 * i] * bf16_lo(prk[sidx][i]), v[2 * i + 1] * bf16_hi(prk[sidx][i]));
        }
        if (tma) {
          sts128(swz128(stg_s, lane, 2 * sidx), o[0], o[1], o[2], o[3]);
          sts128(swz128(stg_s, lane, 2 * sidx + 1), o[4], o[5], o[6], o[7]);
        } else {
          uint4* d = reinterpret_cast<uint4*>(stg + lane * kStgRow + sidx * SB);
          d[0] = make_uint4(o[0], o[1], o[2], o[3]);
          d[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
    }
    const size_t goff = (size_t(row0) * p.N + gcol) * ES;
    if constexpr (F32) {
      __syncwarp();
      stg_flush_add_f32(stg, reinterpret_cast<const uint8_t*>(p.aux) + goff, reinterpret_cast<uint8_t*>(p.out) + goff,
                        gpitch, row_bytes, rows_valid, lane);
      __syncwarp();
    } else if (tma) {
      fence_proxy_async_smem();  // generic-proxy staging writes -> visible to the TMA (async proxy)
      __syncwarp();
      if constexpr (EPI == MAE_EPI_GELU_BF16) {
        if (lane == 0) { tma_store_2d(mo2, stg_s, gcol, row0); tma_store_wait_read(); }  // GELU' -> pre_out
        __syncwarp();
#pragma unroll
        for (int sidx = 0; sidx < 4; ++sidx) {
          sts128(swz128(stg_s, lane, 2 * sidx), hkeep[sidx][0], hkeep[sidx][1], hkeep[sidx][2], hkeep[sidx][3]);
          sts128(swz128(stg_s, lane, 2 * sidx + 1), hkeep[sidx][4], hkeep[sidx][5], hkeep[sidx][6], hkeep[sidx][7]);
        }
        fence_proxy_async_smem();
        __syncwarp();
      }
      if (lane == 0) tma_store_2d(mo, stg_s, gcol, row0);
    } else {
      __syncwarp();
      if constexpr (EPI == MAE_EPI_GELU_BF16) {
        stg_copy<true>(stg, reinterpret_cast<uint8_t*>(p.pre_out) + goff, gpitch, row_bytes, rows_valid, lane);
        __syncwarp();
        for (int sidx = 0; sidx < nsub; ++sidx) {
          uint4* d = reinterpret_cast<uint4*>(stg + lane * kStgRow + sidx * SB);
          d[0] = make_uint4(hkeep[sidx][0], hkeep[sidx][1], hkeep[sidx][2], hkeep[sidx][3]);
          d[1] = make_uint4(hkeep[sidx][4], hkeep[sidx][5], hkeep[sidx][6], hkeep[sidx][7]);
        }
        __syncwarp();
      }
      stg_copy<true>(stg, reinterpret_cast<uint8_t*>(p.out) + goff, gpitch, row_bytes, rows_valid, lane);
      __syncwarp();
    }
  }
}

// dGELU epilogue (fc2 dgrad), TMA both ways: the whole persistent loop of one epilogue warp.  Each item is split into
// 64-column units dealt to the 4 warps of the lane quarter, rotated by the item counter (at most one unit per warp and
// item since bn <= 256).  The saved GELU' tile [32 rows x 64 bf16] of the warp's NEXT unit is fetched by TMA into the
// other of its two staging tiles while the current one is multiplied with the accumulator (row per thread, in place,
// 128-byte swizzle = conflict free) and leaves through a TMA store: the LSU only sees shared-memory traffic.
__device__ __forceinline__ void epilogue_loop_dgelu_tma(const GemmTnParams& p, GemmSmemHdr* hdr, uint32_t tmem_base,
                                                        int q, int grp, int ew, uint8_t* tile0, uint8_t* tile1, int lane,
                                                        int total, const CUtensorMap* mo, const CUtensorMap* maux) {
  const int nunits = p.bn >> 6;
  uint64_t* bars = hdr->auxbar[ew];
  const uint32_t t_s[2] = {smem_u32(tile0), smem_u32(tile1)};
  uint32_t ph0 = 0, ph1 = 0;
  int acc = 0; uint32_t accphase = 0;
  int it = 0;
  int item = blockIdx.x;
  if (item < total && lane == 0) {
    const int u = (grp - 0) & (kEpiGroups - 1);
    if (u < nunits) {
      const int mt = item / p.nchunks, nc = item % p.nchunks;
      mbar_expect_tx(&bars[0], 4096);
      tma_load_2d(tile0, maux, &bars[0], nc * p.bn + u * 64, mt * BM + q * 32);
    }
  }
  for (; item < total; item += gridDim.x, ++it) {
    const int mt = item / p.nchunks, nc = item % p.nchunks;
    const int par = it & 1;
    {  // prefetch the next item's GELU' tile into the other staging tile
      const int nitem = item + gridDim.x;
      const int un = (grp - (it + 1)) & (kEpiGroups - 1);
      if (nitem < total && un < nunits && lane == 0) {
        tma_store_wait_read();  // the store issued from that tile (previous item) has finished reading it
        const int mtn = nitem / p.nchunks, ncn = nitem % p.nchunks;
        mbar_expect_tx(&bars[par ^ 1], 4096);
        tma_load_2d(par ? tile0 : tile1, maux, &bars[par ^ 1], ncn * p.bn + un * 64, mtn * BM + q * 32);
      }
    }
    mbar_wait(&hdr->tfull[acc], accphase);
    tc_fence_after();
    const int u = (grp - it) & (kEpiGroups - 1);
    if (u < nunits) {  // warp-uniform
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc) * 256u + uint32_t(u * 64);
      const uint32_t ts = t_s[par];
      mbar_wait(&bars[par], par ? ph1 : ph0);
      if (par) ph1 ^= 1; else ph0 ^= 1;
#pragma unroll
      for (int sidx = 0; sidx < 4; ++sidx) {
        uint32_t raw[16];
        tmem_ld16_issue(taddr + sidx * 16, raw);
        uint32_t g[8];
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(g[0]), "=r"(g[1]), "=r"(g[2]), "=r"(g[3]) : "r"(swz128(ts, lane, 2 * sidx)));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(g[4]), "=r"(g[5]), "=r"(g[6]), "=r"(g[7]) : "r"(swz128(ts, lane, 2 * sidx + 1)));
        tmem_ld_wait();
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          o[i] = pack_bf16x2(__uint_as_float(raw[2 * i]) * bf16_lo(g[i]), __uint_as_float(raw[2 * i + 1]) * bf16_hi(g[i]));
        sts128(swz128(ts, lane, 2 * sidx), o[0], o[1], o[2], o[3]);
        sts128(swz128(ts, lane, 2 * sidx + 1), o[4], o[5], o[6], o[7]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tma_store_2d(mo, ts, nc * p.bn + u * 64, mt * BM + q * 32);
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&hdr->tempty[acc]);
    acc ^= 1; if (acc == 0) accphase ^= 1;
  }
  if (lane == 0) tma_store_wait_all();
}

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_kernel(const GemmTnParams p, const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapO, const __grid_constant__ CUtensorMap mapO2) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128-byte swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_bytes = BM * BK * 2;             // 16 KB
  const uint32_t b_bytes = uint32_t(p.bn) * BK * 2;  // bn rows x 128 B
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const int nst = p.nstages;
  uint8_t* ring = smem + p.off_ring;
  GemmSmemHdr* hdr = reinterpret_cast<GemmSmemHdr*>(smem + p.off_hdr);
  const int warp = warp_id_uniform();  // (provably warp-uniform: role branches are uniform control flow)
  const int lane = threadIdx.x & 31;
  const int total = p.mtiles * p.nchunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    tma_prefetch_desc(&mapO);
    tma_prefetch_desc(&mapO2);
    for (int s = 0; s < nst; ++s) { mbar_init(&hdr->full[s], 1); mbar_init(&hdr->empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&hdr->tfull[a], 1); mbar_init(&hdr->tempty[a], 4 * kEpiGroups); }
    if constexpr (EPI == MAE_EPI_DGELU_BF16)
      for (int w = 0; w < 4 * kEpiGroups; ++w) { mbar_init(&hdr->auxbar[w][0], 1); mbar_init(&hdr->auxbar[w][1], 1); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&hdr->tmem_base, 512);
  if constexpr (EPI == MAE_EPI_GELU_BF16) {  // constant table (not produced by the previous kernel): stage it before pdl_wait
    uint32_t* tpos = reinterpret_cast<uint32_t*>(smem + p.off_tab);
    uint32_t* tneg = reinterpret_cast<uint32_t*>(smem + p.off_tab + kGeluNegOff);
    for (uint32_t i = threadIdx.x; i < kGeluNT; i += kGemmThreads) { tpos[i] = g_gelu_tab[i]; tneg[i] = g_gelu_tab[kGeluNT + i]; }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only below

  // The TMA and MMA warps run their loops with all 32 lanes (warp-uniform control flow keeps descriptors / coordinates in
  // uniform registers) and elect one lane per burst: under `if (lane == 0)` every UTCHMMA / UTMALDG sits in an
  // ELECT / R2UR.BROADCAST waterfall and costs the issuing thread ~70 cycles -- more than the tensor pipe needs for N <= 128.
  if (warp == 0) {
    {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const int mt = item / p.nchunks, nc = item % p.nchunks;
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&hdr->empty[stage], phase ^ 1);
          uint8_t* sa = ring + stage * stage_bytes;
          if (elect_one_sync()) {
            mbar_expect_tx(&hdr->full[stage], stage_bytes);
            tma_load_2d(sa, &mapA, &hdr->full[stage], kb * BK, mt * BM);
            tma_load_2d(sa + a_bytes, &mapB, &hdr->full[stage], kb * BK, nc * p.bn);
          }
          __syncwarp();
          if (++stage == nst) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc = umma_idesc_bf16(BM, p.bn, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t accphase = 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        mbar_wait(&hdr->tempty[acc], accphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc) * 256u;
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&hdr->full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + stage * stage_bytes);
          const uint32_t sb = sa + a_bytes;
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // K-major SW128: rows are 128 B, 8-row groups 1024 B apart (SBO); +32 B per K=16 step
              const uint64_t adesc = umma_sdesc_sw128(sa + k * 32, 16, 1024);
              const uint64_t bdesc = umma_sdesc_sw128(sb + k * 32, 16, 1024);
              umma_bf16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0);
            }
            umma_commit(&hdr->empty[stage]);
            if (kb == p.nkb - 1) umma_commit(&hdr->tfull[acc]);
          }
          __syncwarp();
          if (++stage == nst) { stage = 0; phase ^= 1; }
        }
        acc ^= 1; if (acc == 0) accphase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int grp = (warp - 2) >> 2;  // which of the quarter's kEpiGroups warps: takes units grp, grp+4, ...
    const int ew = warp - 2;
    auto tile_ptr = [&](int t) -> uint8_t* {
      return smem + (t < p.stg_nA ? p.stg_a0 + uint32_t(t) * p.stg_tile : p.stg_b0 + uint32_t(t - p.stg_nA) * p.stg_tile);
    };
    uint8_t* stg = tile_ptr(ew * p.stg_tpw);
    bool done = false;
    if constexpr (EPI == MAE_EPI_DGELU_BF16) {
      if (p.use_tma_store) {
        epilogue_loop_dgelu_tma(p, hdr, tmem_base, q, grp, ew, stg, tile_ptr(ew * p.stg_tpw + 1), lane, total, &mapO, &mapO2);
        done = true;
      }
    }
    if (!done) {
      uint32_t tab_s = 0;
      if constexpr (EPI == MAE_EPI_GELU_BF16) tab_s = smem_u32(smem + p.off_tab) - kGeluLo * 4;
      int acc = 0; uint32_t accphase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x, ++it) {
        const int mt = item / p.nchunks, nc = item % p.nchunks;
        mbar_wait(&hdr->tfull[acc], accphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc) * 256u;
        if constexpr (EPI == MAE_EPI_BF16 || EPI == MAE_EPI_GELU_BF16)
          epilogue_item<EPI>(p, taddr, mt * BM + q * 32, nc * p.bn, grp, it, stg, lane, &mapO, &mapO2, tab_s);
        else
          epilogue_item_manual<EPI>(p, taddr, mt * BM + q * 32, nc * p.bn, grp, stg, lane);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&hdr->tempty[acc]);
        acc ^= 1; if (acc == 0) accphase ^= 1;
      }
    }
    if (lane == 0) tma_store_wait_all();  // outstanding bulk stores must complete before the CTA's smem goes away
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static int pick_bn(int N) {
  for (int bn = 256; bn >= 16; bn -= 16)
    if (N % bn == 0) return bn;
  return 0;
}

template <int EPI>
static int launch_gemm_tn(const GemmTnParams& p, size_t smem, const CUtensorMap& ma, const CUtensorMap& mb,
                          const CUtensorMap& mo, const CUtensorMap& mo2, cudaStream_t st) {
  static bool attr_done[kMaxDevices] = {};
  if (!ensure_dyn_smem(gemm_tn_kernel<EPI>, 227 * 1024, attr_done)) return MAE_ERR_LAUNCH;
  const int total = p.mtiles * p.nchunks;
  const int grid = total < num_sms() ? total : num_sms();
  mae_launch(gemm_tn_kernel<EPI>, dim3(grid), dim3(kGemmThreads), smem, st, p, ma, mb, mo, mo2);
  return cudaGetLastError() == cudaSuccess ? MAE_OK : MAE_ERR_LAUNCH;
}

static constexpr size_t kSmemMax = 227 * 1024;
static inline uint32_t up1k(size_t x) { return uint32_t((x + 1023) & ~size_t(1023)); }

// Shared-memory plan of one gemm_tn launch: picks the ring depth and fills the offsets in `p`; returns the dynamic
// smem size to request (0 = does not fit).  Regions are 1024-byte aligned relative to the kernel's aligned base.
//   plain:  [ring][hdr][staging tiles]
//   GELU:   [gelu table +][ring][hdr][staging A ...]   ...128 KB above the + table: [gelu table -][staging B ...]
static size_t plan_smem(GemmTnParams& p, int epi) {
  const bool tma = p.use_tma_store != 0;
  p.stg_tile = uint32_t(stg_tile_bytes(epi, tma));
  p.stg_tpw = stg_tiles_per_warp(epi, tma);
  const int ntiles = 4 * kEpiGroups * p.stg_tpw;
  const size_t per_stage = BM * BK * 2 + size_t(p.bn) * BK * 2;
  const size_t hdr = up1k(sizeof(GemmSmemHdr));
  int want = 4;
  // Epilogues that touch global memory through the LSU (fp32 residual, manual dGELU) run faster with a bigger L1: dynamic
  // smem comes out of the same 228 KB, and with K <= 192 (3 k-blocks per item) two ring stages already cover the loop.
  if ((epi == MAE_EPI_DGELU_BF16 || epi == MAE_EPI_RESID_F32) && p.nkb <= 3) want = 2;
  if (const char* e = getenv("MAE_GEMM_STAGES")) { const int v = atoi(e); if (v >= 2 && v <= 4) want = v; }  // experiments
  p.off_tab = 0;
  if (epi != MAE_EPI_GELU_BF16) {
    const size_t fixed = hdr + size_t(ntiles) * p.stg_tile + 1024;
    int nst = want;
    while (nst > 2 && fixed + nst * per_stage > kSmemMax) --nst;
    if (fixed + nst * per_stage > kSmemMax) return 0;
    p.nstages = nst;
    p.off_ring = 0;
    p.off_hdr = uint32_t(nst * per_stage);
    p.stg_a0 = p.off_hdr + uint32_t(hdr);
    p.stg_nA = ntiles; p.stg_b0 = 0;
    return fixed + nst * per_stage;
  }
  // GELU: the two table halves are pinned 128 KB apart; ring, barriers and staging tiles fill the gaps around them
  p.off_ring = up1k(kGeluTabBytes);
  const size_t gapA_end = kGeluNegOff;  // start of the negative table
  int nst = want;
  while (nst > 2 && p.off_ring + nst * per_stage + hdr > gapA_end) --nst;
  if (p.off_ring + nst * per_stage + hdr > gapA_end) return 0;
  p.nstages = nst;
  p.off_hdr = p.off_ring + uint32_t(nst * per_stage);
  p.stg_a0 = p.off_hdr + uint32_t(hdr);
  int nA = int((gapA_end - p.stg_a0) / p.stg_tile);
  if (nA > ntiles) nA = ntiles;
  p.stg_nA = nA;
  p.stg_b0 = up1k(kGeluNegOff + kGeluTabBytes);
  const size_t total = size_t(p.stg_b0) + size_t(ntiles - nA) * p.stg_tile + 1024;
  return total <= kSmemMax ? total : 0;
}

// Host side of the GELU table (see gelu_table8): entry(x) = bf16(gelu(x)) | bf16(gelu'(x)) << 16 for the bf16 inputs
// x = +-[2^-20, 16], evaluated in double precision (gelu = x Phi(x), gelu' = Phi(x) + x phi(x)).
static uint32_t host_bf16_bits(double v) {
  float f = float(v);
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return u >> 16;
  u += 0x7fffu + ((u >> 16) & 1u);  // round to nearest even
  return u >> 16;
}
static int ensure_gelu_table() {
  static bool done[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return MAE_ERR_DRIVER;
  if (done[dev]) return MAE_OK;
  static uint32_t host[2 * kGeluNT];
  for (int sgn = 0; sgn < 2; ++sgn)
    for (uint32_t i = 0; i < kGeluNT; ++i) {
      const uint32_t bits = (kGeluLo + i) | (sgn ? 0x8000u : 0u);
      const uint32_t fb = bits << 16;
      float xf;
      memcpy(&xf, &fb, 4);
      const double x = xf;
      const double cdf = 0.5 * erfc(-x * 0.70710678118654752440);
      const double pdf = 0.39894228040143267794 * exp(-0.5 * x * x);
      host[sgn * kGeluNT + i] = host_bf16_bits(x * cdf) | (host_bf16_bits(cdf + x * pdf) << 16);
    }
  // one-time, synchronous: later launches on any stream see the table
  if (cudaMemcpyToSymbol(g_gelu_tab, host, sizeof(host)) != cudaSuccess) return MAE_ERR_DRIVER;
  done[dev] = true;
  return MAE_OK;
}

const uint32_t* gelu_table_device_ptr() {
  if (ensure_gelu_table() != MAE_OK) return nullptr;
  void* p = nullptr;
  if (cudaGetSymbolAddress(&p, g_gelu_tab) != cudaSuccess) return nullptr;
  return reinterpret_cast<const uint32_t*>(p);
}

}  // namespace mae

using namespace mae;

extern "C" int mae_gemm_tn(const void* A, const void* B, int M, int N, int K, int epi, const float* bias, void* out,
                           void* pre_out, const void* aux, cudaStream_t stream) {
  if (M <= 0) return MAE_OK;
  if (N <= 0 || K <= 0 || (K % 16) || (N % 16)) return MAE_ERR_SHAPE;
  GemmTnParams p;
  p.M = M; p.N = N; p.K = K;
  p.bn = pick_bn(N);
  if (!p.bn) return MAE_ERR_SHAPE;
  p.nchunks = N / p.bn;
  p.mtiles = (M + BM - 1) / BM;
  p.nkb = (K + BK - 1) / BK;
  p.bias = bias; p.out = out; p.pre_out = reinterpret_cast<__nv_bfloat16*>(pre_out); p.aux = aux;
  if (!A || !B || !out) return MAE_ERR_NULL;
  if (epi == MAE_EPI_GELU_BF16 && !pre_out) return MAE_ERR_NULL;
  if ((epi == MAE_EPI_RESID_F32 || epi == MAE_EPI_DGELU_BF16) && !aux) return MAE_ERR_NULL;
  CUtensorMap ma, mb, mo, mo2;
  int rc = make_tmap_bf16(&ma, A, M, K, K, BM);
  if (rc) return rc;
  rc = make_tmap_bf16(&mb, B, N, K, K, p.bn);
  if (rc) return rc;
  {
    static const bool no_tma = getenv("MAE_GEMM_NO_TMA_STORE") != nullptr;
    p.use_tma_store = ((epi == MAE_EPI_BF16 || epi == MAE_EPI_GELU_BF16 || (epi == MAE_EPI_DGELU_BF16 && p.bn % 64 == 0)) &&
                       !no_tma) ? 1 : 0;
  }
  const size_t smem = plan_smem(p, epi);
  if (!smem) return MAE_ERR_SHAPE;
  // epilogue tensor maps: [32 rows x 64 bf16] boxes, 128-byte swizzle (dummy = A's map when unused)
  mo = ma; mo2 = ma;
  if (p.use_tma_store) {
    rc = make_tmap_bf16(&mo, out, M, N, N, 32);
    if (rc) return rc;
    if (epi == MAE_EPI_GELU_BF16) rc = make_tmap_bf16(&mo2, pre_out, M, N, N, 32);
    if (epi == MAE_EPI_DGELU_BF16) rc = make_tmap_bf16(&mo2, aux, M, N, N, 32);
    if (rc) return rc;
  }
  if (epi == MAE_EPI_GELU_BF16) {
    rc = ensure_gelu_table();
    if (rc) return rc;
  }
  switch (epi) {
    case MAE_EPI_BF16: return launch_gemm_tn<MAE_EPI_BF16>(p, smem, ma, mb, mo, mo2, stream);
    case MAE_EPI_GELU_BF16: return launch_gemm_tn<MAE_EPI_GELU_BF16>(p, smem, ma, mb, mo, mo2, stream);
    case MAE_EPI_RESID_F32: return launch_gemm_tn<MAE_EPI_RESID_F32>(p, smem, ma, mb, mo, mo2, stream);
    case MAE_EPI_DGELU_BF16: return launch_gemm_tn<MAE_EPI_DGELU_BF16>(p, smem, ma, mb, mo, mo2, stream);
    default: return MAE_ERR_SHAPE;
  }
}

// ---------------------------------------------------------------------------------------------
// gemm_wgrad:  O[p, q] (+)= sum_m P[m, p] * Q[m, q]     (fp32 atomics into `out`)
//   P: bf16 [M, Pc] row-major, Q: bf16 [M, Qc] row-major, Qc % 16 == 0.
//   grid = (ceil(Pc/128), msplit, ceil(Qc/256));  each CTA reduces its own M range of one 128 x <=256 output
//   tile into one TMEM accumulator.
//   out element (p, q) lives at out[p * so_p + q * so_q].
// ---------------------------------------------------------------------------------------------
namespace mae {

static constexpr int WR = 64;  // m-rows per stage (4 MMAs of K=16)

struct WgradParams {
  int M, Pc, Qc;
  int nqb;          // ceil(Qc / 64) TMA boxes for Q per stage
  int rows_per_cta; // multiple of WR
  int pcc;          // P chunks (of 128 columns) per CTA: 1 or 2.  Two accumulators share every staged Q tile, which halves
                    // the Q re-reads and doubles the MMA work per ring stage (the ring is bound by the ~1.5 us turnaround
                    // of a stage, not by bytes)
  int nstages;      // ring depth (whatever fits: 3 or 4)
  float* out;
  long long so_p, so_q;
  float* colsum;    // optional: += column sums of P (colsum_of_q == 0) or of Q (== 1): the bias gradient
  int colsum_of_q;
};

// bf16 element (row r, col c) of a [64 rows x 64 cols] TMA box written with the 128-byte swizzle
__device__ __forceinline__ float swz_box_elem(const uint8_t* box, int r, int c) {
  const uint32_t off = uint32_t(r) * 128u + ((uint32_t(c >> 3) ^ uint32_t(r & 7)) << 4) + uint32_t(c & 7) * 2u;
  return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(box + off));
}

__global__ void __launch_bounds__(kWgradThreads, 1)
gemm_wgrad_kernel(const WgradParams p, const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapQ) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t box_bytes = WR * 128;  // one [64 rows x 64 cols] bf16 box = 8 KB
  const int npb = 2 * p.pcc;            // P boxes per stage
  const uint32_t p_bytes = uint32_t(npb) * box_bytes;
  const uint32_t q_bytes = uint32_t(p.nqb) * box_bytes;
  const uint32_t stage_bytes = p_bytes + q_bytes;
  const int nst = p.nstages;
  GemmSmemHdr* hdr = reinterpret_cast<GemmSmemHdr*>(smem + nst * stage_bytes);

  const int warp = warp_id_uniform();
  const int lane = threadIdx.x & 31;
  const int pc0 = blockIdx.x * p.pcc;                    // first 128-column P chunk of this CTA
  const int q0 = blockIdx.z * 256;                       // first Q column of this CTA's output tile
  const int Qn = (p.Qc - q0) < 256 ? (p.Qc - q0) : 256;  // its width (multiple of 16)
  const int nqb = (Qn + 63) >> 6;                        // TMA boxes of Q actually loaded per stage
  const uint32_t tx_bytes = p_bytes + uint32_t(nqb) * box_bytes;
  const int m_begin = blockIdx.y * p.rows_per_cta;
  int m_end = m_begin + p.rows_per_cta;
  if (m_end > p.M) m_end = p.M;
  const int nmb = m_end > m_begin ? (m_end - m_begin + WR - 1) / WR : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapP);
    tma_prefetch_desc(&mapQ);
    // a stage is released by the MMA commit and, when column sums are fused, by the 4 epilogue warps too
    for (int s = 0; s < nst; ++s) { mbar_init(&hdr->full[s], 1); mbar_init(&hdr->empty[s], p.colsum ? 5 : 1); }
    mbar_init(&hdr->tfull[0], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&hdr->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only below
  if (nmb > 0) {
    // (TMA / MMA warps: all lanes run the loop, one elected lane issues -- see gemm_tn_kernel)
    if (warp == 0) {
      {
        int stage = 0; uint32_t phase = 0;
        for (int mb = 0; mb < nmb; ++mb) {
          const int m0 = m_begin + mb * WR;
          mbar_wait(&hdr->empty[stage], phase ^ 1);
          uint8_t* sp = smem + stage * stage_bytes;
          if (elect_one_sync()) {
            mbar_expect_tx(&hdr->full[stage], tx_bytes);
            for (int j = 0; j < npb; ++j)  // columns past Pc are zero-filled by the TMA
              tma_load_2d(sp + j * box_bytes, &mapP, &hdr->full[stage], pc0 * 128 + j * 64, m0);
            for (int j = 0; j < nqb; ++j)
              tma_load_2d(sp + p_bytes + j * box_bytes, &mapQ, &hdr->full[stage], q0 + j * 64, m0);
          }
          __syncwarp();
          if (++stage == nst) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      {
        const uint32_t idesc = umma_idesc_bf16(BM, Qn, 1, 1);
        int stage = 0; uint32_t phase = 0;
        for (int mb = 0; mb < nmb; ++mb) {
          mbar_wait(&hdr->full[stage], phase);
          tc_fence_after();
          const uint32_t sp = smem_u32(smem + stage * stage_bytes);
          const uint32_t sq = sp + p_bytes;
          if (elect_one_sync()) {
            for (int c = 0; c < p.pcc; ++c) {
#pragma unroll
              for (int k = 0; k < WR / 16; ++k) {
                // MN-major SW128: 64 MN elements contiguous (128 B) per reduction row; 8-row groups 1024 B apart
                // (SBO); successive 64-element MN groups one TMA box apart (LBO); +2048 B per K=16 step.
                const uint64_t adesc = umma_sdesc_sw128(sp + uint32_t(2 * c) * box_bytes + k * 2048, box_bytes, 1024);
                const uint64_t bdesc = umma_sdesc_sw128(sq + k * 2048, box_bytes, 1024);
                umma_bf16(tmem_base + uint32_t(c) * 256u, adesc, bdesc, idesc, (mb | k) != 0);
              }
            }
            umma_commit(&hdr->empty[stage]);
            if (mb == nmb - 1) umma_commit(&hdr->tfull[0]);
          }
          __syncwarp();
          if (++stage == nst) { stage = 0; phase ^= 1; }
        }
      }
    } else {
      const int q = warp & 3;
      if (p.colsum != nullptr) {
        // The epilogue warps are idle during the main loop: they add up the columns of the staged dY tile
        // (bias gradient) straight from shared memory, so dY is read from HBM exactly once.
        // Thread t of 128: box (t>>6) + 2j, 16-byte column chunk cc = t&7, row group rg = (t>>3)&7 -> rows rg+8i.
        // Within a 128B-swizzled box the chunk of row r sits at (cc ^ (r&7)) = (cc ^ rg): constant per thread.
        const int t = (warp - 2) * 32 + lane;
        const bool of_q = p.colsum_of_q != 0;
        const bool active = of_q ? (blockIdx.x == 0) : (blockIdx.z == 0);  // each column summed by exactly one CTA row
        const int nbox = of_q ? nqb : npb;
        const uint32_t tile_off = of_q ? p_bytes : 0u;
        const int cc = t & 7, rg = (t >> 3) & 7;
        const uint32_t thr_off = uint32_t(rg) * 128u + (uint32_t(cc ^ rg) << 4);
        float acc[2][8];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[j][e] = 0.f;
        int stage = 0; uint32_t phase = 0;
        for (int mb = 0; mb < nmb; ++mb) {
          mbar_wait(&hdr->full[stage], phase);
          if (active) {
            const uint8_t* sp = smem + stage * stage_bytes + tile_off;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int box = (t >> 6) + 2 * j;
              if (box < nbox) {
                const uint8_t* bp = sp + box * box_bytes + thr_off;
#pragma unroll
                for (int i = 0; i < WR / 8; ++i) {
                  const uint4 v = *reinterpret_cast<const uint4*>(bp + i * 1024);
                  acc[j][0] += bf16_lo(v.x); acc[j][1] += bf16_hi(v.x); acc[j][2] += bf16_lo(v.y); acc[j][3] += bf16_hi(v.y);
                  acc[j][4] += bf16_lo(v.z); acc[j][5] += bf16_hi(v.z); acc[j][6] += bf16_lo(v.w); acc[j][7] += bf16_hi(v.w);
                }
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&hdr->empty[stage]);
          if (++stage == nst) { stage = 0; phase ^= 1; }
        }
        if (active) {
          const int col_base = of_q ? q0 : pc0 * 128;
          const int ncols = of_q ? p.Qc : p.Pc;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int box = (t >> 6) + 2 * j;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float v = acc[j][e];
              v += __shfl_xor_sync(0xffffffffu, v, 8);   // lanes differing in rg bits 0,1
              v += __shfl_xor_sync(0xffffffffu, v, 16);
              const int c = col_base + box * 64 + cc * 8 + e;
              if ((lane >> 3) == 0 && box < nbox && c < ncols) atomicAdd(p.colsum + c, v);
            }
          }
        }
      }
      mbar_wait(&hdr->tfull[0], 0);
      tc_fence_after();
      for (int c = 0; c < p.pcc; ++c) {
        const int prow = (pc0 + c) * 128 + q * 32 + lane;
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(c) * 256u;
        for (int c0 = 0; c0 < Qn; c0 += 16) {
          float v[16];
          tmem_ld16(taddr + c0, v);
          if (prow < p.Pc) {
            float* o = p.out + (long long)prow * p.so_p + (long long)(q0 + c0) * p.so_q;
            if (p.so_q == 1 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
              // 16 consecutive floats of one dW row: four 16-byte vector reductions instead of 16 scalar ones
#pragma unroll
              for (int i = 0; i < 16; i += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + i), "f"(v[i]), "f"(v[i + 1]),
                             "f"(v[i + 2]), "f"(v[i + 3])
                             : "memory");
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) atomicAdd(o + (long long)i * p.so_q, v[i]);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace mae

extern "C" int mae_gemm_wgrad(const void* P, int Pc, const void* Q, int Qc, int M, float* out, long long so_p,
                              long long so_q, float* colsum, int colsum_of_q, cudaStream_t stream) {
  if (M <= 0) return MAE_OK;
  if (!P || !Q || !out) return MAE_ERR_NULL;
  if (Pc <= 0 || Qc <= 0 || (Qc % 16) || (Pc % 8)) return MAE_ERR_SHAPE;
  WgradParams p;
  p.M = M; p.Pc = Pc; p.Qc = Qc;
  const int qchunks = (Qc + 255) / 256;
  p.nqb = Qc >= 256 ? 4 : (Qc + 63) / 64;  // smem is laid out for the widest chunk
  p.out = out; p.so_p = so_p; p.so_q = so_q;
  p.colsum = colsum; p.colsum_of_q = colsum_of_q;
  const int pchunks = (Pc + 127) / 128;
  {
    static const int force = getenv("MAE_WGRAD_PCC") ? atoi(getenv("MAE_WGRAD_PCC")) : 0;  // experiments
    p.pcc = (pchunks >= 6 && pchunks % 2 == 0) ? 2 : 1;  // measured: 768-wide P 106 -> 99 us; narrower P loses M-splits / stages
    if (force == 1 || force == 2) p.pcc = force;
  }
  const size_t stage = size_t(2 * p.pcc + p.nqb) * WR * 128;
  p.nstages = kStages;
  while (p.nstages > 2 && size_t(p.nstages) * stage + sizeof(GemmSmemHdr) + 1024 > 227 * 1024) --p.nstages;
  if (size_t(p.nstages) * stage + sizeof(GemmSmemHdr) + 1024 > 227 * 1024) return MAE_ERR_SHAPE;
  const int pgroups = (pchunks + p.pcc - 1) / p.pcc;
  int msplit = num_sms() / (pgroups * qchunks);
  if (msplit < 1) msplit = 1;
  int rows = (M + msplit - 1) / msplit;
  rows = ((rows + WR - 1) / WR) * WR;
  p.rows_per_cta = rows;
  msplit = (M + rows - 1) / rows;
  CUtensorMap mp, mq;
  int rc = make_tmap_bf16(&mp, P, M, Pc, Pc, WR);
  if (rc) return rc;
  rc = make_tmap_bf16(&mq, Q, M, Qc, Qc, WR);
  if (rc) return rc;
  const size_t smem = size_t(p.nstages) * stage + sizeof(GemmSmemHdr) + 1024;
  static bool attr_done[kMaxDevices] = {};
  if (!ensure_dyn_smem(gemm_wgrad_kernel, 227 * 1024, attr_done)) return MAE_ERR_LAUNCH;
  mae_launch(gemm_wgrad_kernel, dim3(pgroups, msplit, qchunks), dim3(kWgradThreads), smem, stream, p, mp, mq);
  return cudaGetLastError() == cudaSuccess ? MAE_OK : MAE_ERR_LAUNCH;
}
