// tcgen05 GEMMs for the MAE step (sm_100a).
//
//   gemm_tn     C[M,N]  = epi( A[M,K] * B[N,K]^T )      forward linears and dgrads (both operands K-major)
//   gemm_wgrad  O[P,Q] += sum_m  P[m,:]^T Q[m,:]        weight gradients (both operands MN-major)
//
// Shared structure: 6 warps.  warp 0 = TMA producer (one lane), warp 1 = TMEM owner + MMA issuer (one
// lane), warps 2..5 = epilogue (TMEM -> registers -> global).  Operand tiles are staged by TMA with the
// 128-byte swizzle in a 4-deep mbarrier ring; accumulators live in TMEM (fp32), double buffered for
// gemm_tn so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "mae_internal.h"
#include "ptx.cuh"

namespace mae {

static constexpr int kStages = 4;
static constexpr int kGemmThreads = 192;
static constexpr int BM = 128;  // UMMA M
static constexpr int BK = 64;   // one 128-byte swizzle span of bf16

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

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
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
  const float* bias;          // [N] or null
  void* out;                  // bf16 [M,N] (EPI 0,1,3,4)  or  f32 [M,N] (EPI 2)
  __nv_bfloat16* pre_out;     // EPI 1: bf16 [M,N] pre-activation
  const void* aux;            // EPI 2: f32 residual [M,N];  EPI 4: bf16 pre-activation [M,N]
};

struct GemmSmemHdr {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint32_t tmem_base;
};

template <int EPI>
__device__ __forceinline__ void epilogue16(const GemmTnParams& p, int row, int col, float (&v)[16]) {
  if (p.bias != nullptr) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += __ldg(p.bias + col + i);
  }
  const size_t off = size_t(row) * p.N + col;
  if constexpr (EPI == MAE_EPI_BF16) {
    uint4 o0, o1;
    o0.x = pack_bf16x2(v[0], v[1]);  o0.y = pack_bf16x2(v[2], v[3]);
    o0.z = pack_bf16x2(v[4], v[5]);  o0.w = pack_bf16x2(v[6], v[7]);
    o1.x = pack_bf16x2(v[8], v[9]);  o1.y = pack_bf16x2(v[10], v[11]);
    o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off);
    dst[0] = o0; dst[1] = o1;
  } else if constexpr (EPI == MAE_EPI_GELU_BF16) {
    // autocast semantics: the linear's output is rounded to bf16, GELU reads that bf16 value
    uint32_t pr[8], hr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      pr[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
      hr[i] = pack_bf16x2(gelu_erf(bf16_lo(pr[i])), gelu_erf(bf16_hi(pr[i])));
    }
    uint4* d0 = reinterpret_cast<uint4*>(p.pre_out + off);
    uint4* d1 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off);
    d0[0] = make_uint4(pr[0], pr[1], pr[2], pr[3]); d0[1] = make_uint4(pr[4], pr[5], pr[6], pr[7]);
    d1[0] = make_uint4(hr[0], hr[1], hr[2], hr[3]); d1[1] = make_uint4(hr[4], hr[5], hr[6], hr[7]);
  } else if constexpr (EPI == MAE_EPI_RESID_F32) {
    const float4* r = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.aux) + off);
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 x = r[i];
      x.x += v[4 * i]; x.y += v[4 * i + 1]; x.z += v[4 * i + 2]; x.w += v[4 * i + 3];
      dst[i] = x;
    }
  } else if constexpr (EPI == MAE_EPI_DGELU_BF16) {
    const uint4* pp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.aux) + off);
    uint4 a = pp[0], b = pp[1];
    const uint32_t pr[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      o[i] = pack_bf16x2(v[2 * i] * gelu_erf_grad(bf16_lo(pr[i])), v[2 * i + 1] * gelu_erf_grad(bf16_hi(pr[i])));
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off);
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]); dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
  }
}

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_kernel(const GemmTnParams p, const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128-byte swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_bytes = BM * BK * 2;             // 16 KB
  const uint32_t b_bytes = uint32_t(p.bn) * BK * 2;  // bn rows x 128 B
  const uint32_t stage_bytes = a_bytes + b_bytes;
  GemmSmemHdr* hdr = reinterpret_cast<GemmSmemHdr*>(smem + kStages * stage_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total = p.mtiles * p.nchunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    for (int s = 0; s < kStages; ++s) { mbar_init(&hdr->full[s], 1); mbar_init(&hdr->empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&hdr->tfull[a], 1); mbar_init(&hdr->tempty[a], 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&hdr->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const int mt = item / p.nchunks, nc = item % p.nchunks;
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&hdr->empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * stage_bytes;
          mbar_expect_tx(&hdr->full[stage], stage_bytes);
          tma_load_2d(sa, &mapA, &hdr->full[stage], kb * BK, mt * BM);
          tma_load_2d(sa + a_bytes, &mapB, &hdr->full[stage], kb * BK, nc * p.bn);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
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
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint32_t sb = sa + a_bytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major SW128: rows are 128 B, 8-row groups 1024 B apart (SBO); +32 B per K=16 step
            const uint64_t adesc = umma_sdesc_sw128(sa + k * 32, 16, 1024);
            const uint64_t bdesc = umma_sdesc_sw128(sb + k * 32, 16, 1024);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0);
          }
          umma_commit(&hdr->empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&hdr->tfull[acc]);
        acc ^= 1; if (acc == 0) accphase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0; uint32_t accphase = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      const int mt = item / p.nchunks, nc = item % p.nchunks;
      mbar_wait(&hdr->tfull[acc], accphase);
      tc_fence_after();
      const int row = mt * BM + q * 32 + lane;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc) * 256u;
      for (int c0 = 0; c0 < p.bn; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + c0, v);
        if (row < p.M) epilogue16<EPI>(p, row, nc * p.bn + c0, v);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&hdr->tempty[acc]);
      acc ^= 1; if (acc == 0) accphase ^= 1;
    }
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
static int launch_gemm_tn(const GemmTnParams& p, const CUtensorMap& ma, const CUtensorMap& mb, cudaStream_t st) {
  const size_t smem = size_t(kStages) * (BM * BK * 2 + p.bn * BK * 2) + sizeof(GemmSmemHdr) + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return MAE_ERR_LAUNCH;
    attr_done = true;
  }
  const int total = p.mtiles * p.nchunks;
  const int grid = total < num_sms() ? total : num_sms();
  gemm_tn_kernel<EPI><<<grid, kGemmThreads, smem, st>>>(p, ma, mb);
  return cudaGetLastError() == cudaSuccess ? MAE_OK : MAE_ERR_LAUNCH;
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
  CUtensorMap ma, mb;
  int rc = make_tmap_bf16(&ma, A, M, K, K, BM);
  if (rc) return rc;
  rc = make_tmap_bf16(&mb, B, N, K, K, p.bn);
  if (rc) return rc;
  switch (epi) {
    case MAE_EPI_BF16: return launch_gemm_tn<MAE_EPI_BF16>(p, ma, mb, stream);
    case MAE_EPI_GELU_BF16: return launch_gemm_tn<MAE_EPI_GELU_BF16>(p, ma, mb, stream);
    case MAE_EPI_RESID_F32: return launch_gemm_tn<MAE_EPI_RESID_F32>(p, ma, mb, stream);
    case MAE_EPI_DGELU_BF16: return launch_gemm_tn<MAE_EPI_DGELU_BF16>(p, ma, mb, stream);
    default: return MAE_ERR_SHAPE;
  }
}

// ---------------------------------------------------------------------------------------------
// gemm_wgrad:  O[p, q] (+)= sum_m P[m, p] * Q[m, q]     (fp32 atomics into `out`)
//   P: bf16 [M, Pc] row-major, Q: bf16 [M, Qc] row-major, Qc <= 256, Qc % 16 == 0.
//   grid = (ceil(Pc/128), msplit);  each CTA reduces its own M range into one TMEM accumulator.
//   out element (p, q) lives at out[p * so_p + q * so_q].
// ---------------------------------------------------------------------------------------------
namespace mae {

static constexpr int WR = 64;  // m-rows per stage (4 MMAs of K=16)

struct WgradParams {
  int M, Pc, Qc;
  int nqb;          // ceil(Qc / 64) TMA boxes for Q per stage
  int rows_per_cta; // multiple of WR
  float* out;
  long long so_p, so_q;
};

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_wgrad_kernel(const WgradParams p, const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapQ) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t box_bytes = WR * 128;  // one [64 rows x 64 cols] bf16 box = 8 KB
  const uint32_t p_bytes = 2 * box_bytes;
  const uint32_t q_bytes = uint32_t(p.nqb) * box_bytes;
  const uint32_t stage_bytes = p_bytes + q_bytes;
  GemmSmemHdr* hdr = reinterpret_cast<GemmSmemHdr*>(smem + kStages * stage_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int pc = blockIdx.x;
  const int m_begin = blockIdx.y * p.rows_per_cta;
  int m_end = m_begin + p.rows_per_cta;
  if (m_end > p.M) m_end = p.M;
  const int nmb = m_end > m_begin ? (m_end - m_begin + WR - 1) / WR : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapP);
    tma_prefetch_desc(&mapQ);
    for (int s = 0; s < kStages; ++s) { mbar_init(&hdr->full[s], 1); mbar_init(&hdr->empty[s], 1); }
    mbar_init(&hdr->tfull[0], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&hdr->tmem_base, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;
  if (nmb > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int mb = 0; mb < nmb; ++mb) {
          const int m0 = m_begin + mb * WR;
          mbar_wait(&hdr->empty[stage], phase ^ 1);
          uint8_t* sp = smem + stage * stage_bytes;
          mbar_expect_tx(&hdr->full[stage], stage_bytes);
          tma_load_2d(sp, &mapP, &hdr->full[stage], pc * 128, m0);
          tma_load_2d(sp + box_bytes, &mapP, &hdr->full[stage], pc * 128 + 64, m0);
          for (int j = 0; j < p.nqb; ++j)
            tma_load_2d(sp + p_bytes + j * box_bytes, &mapQ, &hdr->full[stage], j * 64, m0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = umma_idesc_bf16(BM, p.Qc, 1, 1);
        int stage = 0; uint32_t phase = 0;
        for (int mb = 0; mb < nmb; ++mb) {
          mbar_wait(&hdr->full[stage], phase);
          tc_fence_after();
          const uint32_t sp = smem_u32(smem + stage * stage_bytes);
          const uint32_t sq = sp + p_bytes;
#pragma unroll
          for (int k = 0; k < WR / 16; ++k) {
            // MN-major SW128: 64 MN elements contiguous (128 B) per reduction row; 8-row groups 1024 B apart
            // (SBO); successive 64-element MN groups one TMA box apart (LBO); +2048 B per K=16 step.
            const uint64_t adesc = umma_sdesc_sw128(sp + k * 2048, box_bytes, 1024);
            const uint64_t bdesc = umma_sdesc_sw128(sq + k * 2048, box_bytes, 1024);
            umma_bf16(tmem_base, adesc, bdesc, idesc, (mb | k) != 0);
          }
          umma_commit(&hdr->empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&hdr->tfull[0]);
      }
    } else {
      const int q = warp & 3;
      mbar_wait(&hdr->tfull[0], 0);
      tc_fence_after();
      const int prow = pc * 128 + q * 32 + lane;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16);
      for (int c0 = 0; c0 < p.Qc; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + c0, v);
        if (prow < p.Pc) {
          float* o = p.out + (long long)prow * p.so_p + (long long)c0 * p.so_q;
#pragma unroll
          for (int i = 0; i < 16; ++i) atomicAdd(o + (long long)i * p.so_q, v[i]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

}  // namespace mae

extern "C" int mae_gemm_wgrad(const void* P, int Pc, const void* Q, int Qc, int M, float* out, long long so_p,
                              long long so_q, cudaStream_t stream) {
  if (M <= 0) return MAE_OK;
  if (!P || !Q || !out) return MAE_ERR_NULL;
  if (Pc <= 0 || Qc <= 0 || Qc > 256 || (Qc % 16) || (Pc % 8)) return MAE_ERR_SHAPE;
  WgradParams p;
  p.M = M; p.Pc = Pc; p.Qc = Qc;
  p.nqb = (Qc + 63) / 64;
  p.out = out; p.so_p = so_p; p.so_q = so_q;
  const int pchunks = (Pc + 127) / 128;
  int msplit = num_sms() / pchunks;
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
  const size_t smem = size_t(kStages) * (2 + p.nqb) * WR * 128 + sizeof(GemmSmemHdr) + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return MAE_ERR_LAUNCH;
    attr_done = true;
  }
  gemm_wgrad_kernel<<<dim3(pchunks, msplit), kGemmThreads, smem, stream>>>(p, mp, mq);
  return cudaGetLastError() == cudaSuccess ? MAE_OK : MAE_ERR_LAUNCH;
}
