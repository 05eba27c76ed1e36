// Patch embedding of the KEPT tokens in one launch (sm_100a):
//
//   x[b, 0]     = cls_token + pos_embed[idx_keep[b, 0]]
//   x[b, j + 1] = bf16( patch(images[b], idx_keep[b, j+1] - 1) . W^T + bias ) + pos_embed[idx_keep[b, j+1]]
//
// i.e. timm PatchEmbed (Conv2d k = s = patch) -> cat(cls) -> + pos_embed -> get_at_index(idx_keep) of the reference
// (src/models/mae.py:54-55 via lightly MaskedVisionTransformerTIMM.encode), restricted to the tokens that survive the
// gather.  It replaces mae_gather_patches + mae_gemm_tn + mae_assemble_tokens and their two bf16 round trips
// through HBM (the im2col rows and the GEMM output): the producer warps read the 8 x 8 patch rows straight from the fp32
// images, convert them to bf16 in the conv-weight order (c, py, px) and write them as the 128-byte-swizzled K-major A tile
// of a tcgen05 GEMM (and, once, to `patches`, which the weight gradient needs); the weight (D x C p^2 bf16, 55 KB for
// configs/mae.yaml) is loaded once per CTA by TMA and stays in shared memory; the epilogue adds the bias, rounds to bf16
// exactly where autocast rounds the conv output, adds the position row in fp32 and writes the fp32 stream, class-token
// rows included.  A tiles and TMEM accumulators are double buffered: the warps gather tile t+1 while the tensor pipe
// works on tile t, and the epilogue of tile t overlaps the MMAs of tile t+1.
#include "gemm_common.cuh"

namespace mae {

static constexpr int kPeEpiWarps = 16;
static constexpr int kPeThreads = (2 + kPeEpiWarps) * 32;  // warp 0: weight TMA, warp 1: MMA, warps 2-17: gather + epilogue

struct PeHdr {
  uint64_t w_full;
  uint64_t a_ready[2], a_free[2];
  uint64_t acc_full[2], acc_free[2];
  uint32_t tmem_base;
};

struct PeParams {
  int B, C, img, patch, num_keep, D, pd, nkb, ntiles, grid;
  long long mrows;  // B * (num_keep - 1)
  const float* images;
  const int64_t* idx_keep;
  const float *bias, *cls, *pos;
  __nv_bfloat16* patches;
  float* x;
  uint32_t off_a, off_hdr;  // smem offsets: [W: nkb boxes of D x 128 B][A: 2 x nkb panels of 128 x 128 B][hdr]
};

__global__ void __launch_bounds__(kPeThreads, 1)
patch_embed_kernel(const PeParams p, const __grid_constant__ CUtensorMap mapW) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  PeHdr* hdr = reinterpret_cast<PeHdr*>(smem + p.off_hdr);
  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const int D = p.D, nkb = p.nkb;
  const uint32_t w_box = uint32_t(D) * 128u, a_buf = uint32_t(nkb) * 16384u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapW);
    mbar_init(&hdr->w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hdr->a_ready[i], kPeEpiWarps);
      mbar_init(&hdr->a_free[i], 1);
      mbar_init(&hdr->acc_full[i], 1);
      mbar_init(&hdr->acc_free[i], kPeEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&hdr->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;
  pdl_wait();
  const uint32_t w_s = smem_u32(smem), a_s = smem_u32(smem + p.off_a);

  if (warp == 0) {
    if (lane == 0) {  // the whole weight, once: nkb boxes of [D rows x 64 k] (k >= pd zero-filled)
      mbar_expect_tx(&hdr->w_full, uint32_t(nkb) * w_box);
      for (int kb = 0; kb < nkb; ++kb) tma_load_2d(smem + kb * w_box, &mapW, &hdr->w_full, kb * BK, 0);
    }
  } else if (warp == 1) {
    {  // all lanes run the loop, one elected lane issues (no ELECT / R2UR waterfall per UTCHMMA, see ptx.cuh)
      const uint32_t idesc = umma_idesc_bf16(BM, D, 0, 0);
      const int ks_n = p.pd >> 4;
      mbar_wait(&hdr->w_full, 0);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const uint32_t b = it & 1u, ph = (it >> 1) & 1u;
        mbar_wait(&hdr->acc_free[b], ph ^ 1u);
        mbar_wait(&hdr->a_ready[b], ph);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + b * 256u;
        if (elect_one_sync()) {
          for (int ks = 0; ks < ks_n; ++ks) {
            const int kb = ks >> 2, k = ks & 3;
            const uint64_t adesc = umma_sdesc_sw128(a_s + b * a_buf + kb * 16384 + k * 32, 16, 1024);
            const uint64_t bdesc = umma_sdesc_sw128(w_s + kb * w_box + k * 32, 16, 1024);
            umma_bf16(d_tmem, adesc, bdesc, idesc, ks != 0);
          }
          umma_commit(&hdr->a_free[b]);
          umma_commit(&hdr->acc_full[b]);
        }
        __syncwarp();
      }
    }
  } else {
    const int ew = warp - 2, q = warp & 3, grp = ew >> 2;
    const int nk = p.num_keep - 1, patch = p.patch, img = p.img;
    const int p4 = patch >> 2;               // float4 per patch row
    const int nf = p.pd >> 2;                // float4 per patch
    // this lane's float4 slots f = lane + 32 s of a patch in conv-weight order k = (c, py, px): fixed per lane
    int img_off[2], kk[2];
    bool fok[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int f = lane + 32 * s;
      fok[s] = f < nf;
      const int c = f / (patch * p4), py = (f / p4) % patch, px = (f % p4) * 4;
      img_off[s] = (c * img + py) * img + px;
      kk[s] = f * 4;
    }
    const size_t img_sz = size_t(p.C) * img * img;

    auto produce = [&](int tile, uint32_t it) {
      const uint32_t b = it & 1u, ph = (it >> 1) & 1u;
      mbar_wait(&hdr->a_free[b], ph ^ 1u);  // the MMAs that read this buffer two tiles ago have completed
      float4 v[8][2];
      long long grow[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long gr = (long long)tile * BM + ew * 8 + i;
        grow[i] = gr < p.mrows ? gr : -1;
        v[i][0] = v[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < p.mrows) {
          const int bb = int(uint32_t(gr) / uint32_t(nk)), j = int(uint32_t(gr) % uint32_t(nk)) + 1;
          const int pidx = int(__ldg(p.idx_keep + size_t(bb) * p.num_keep + j)) - 1;
          const int pr = pidx / p.grid, pc = pidx % p.grid;
          const float* src = p.images + size_t(bb) * img_sz + size_t(pr * patch) * img + pc * patch;
#pragma unroll
          for (int s = 0; s < 2; ++s)
            if (fok[s]) v[i][s] = *reinterpret_cast<const float4*>(src + img_off[s]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int lr = ew * 8 + i;
#pragma unroll
        for (int s = 0; s < 2; ++s)
          if (fok[s]) {
            const uint32_t lo = pack_bf16x2(v[i][s].x, v[i][s].y), hi = pack_bf16x2(v[i][s].z, v[i][s].w);
            const int k = kk[s];
            const uint32_t sa = a_s + b * a_buf + uint32_t(k >> 6) * 16384u + uint32_t(lr) * 128u +
                                ((uint32_t((k & 63) >> 3) ^ uint32_t(lr & 7)) << 4) + uint32_t(k & 7) * 2u;
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(sa), "r"(lo), "r"(hi) : "memory");
            if (grow[i] >= 0) *reinterpret_cast<uint2*>(p.patches + grow[i] * p.pd + k) = make_uint2(lo, hi);
          }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&hdr->a_ready[b]);
    };

    auto epilogue = [&](int tile, uint32_t it) {
      const uint32_t b = it & 1u, ph = (it >> 1) & 1u;
      const long long gr = (long long)tile * BM + q * 32 + lane;
      const bool ok = gr < p.mrows;
      int bb = 0, j = 0, tokidx = 0, tok0 = 0;
      if (ok) {
        bb = int(uint32_t(gr) / uint32_t(nk)); j = int(uint32_t(gr) % uint32_t(nk)) + 1;
        tokidx = int(__ldg(p.idx_keep + size_t(bb) * p.num_keep + j));
        if (j == 1) tok0 = int(__ldg(p.idx_keep + size_t(bb) * p.num_keep));
      }
      const float* prow = p.pos + size_t(tokidx) * D;
      float* xrow = p.x + (size_t(bb) * p.num_keep + j) * D;
      mbar_wait(&hdr->acc_full[b], ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + b * 256u;
      const int nsc = D >> 4;
      const int sc_begin = (nsc * grp) / 4, sc_end = (nsc * (grp + 1)) / 4;
      for (int sc = sc_begin; sc < sc_end; ++sc) {
        uint32_t raw[16];
        tmem_ld16_issue(taddr + sc * 16, raw);
        float bia[16], ps[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(p.bias + sc * 16) + i);
          bia[4 * i] = t4.x; bia[4 * i + 1] = t4.y; bia[4 * i + 2] = t4.z; bia[4 * i + 3] = t4.w;
          const float4 p4v = ok ? __ldg(reinterpret_cast<const float4*>(prow + sc * 16) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          ps[4 * i] = p4v.x; ps[4 * i + 1] = p4v.y; ps[4 * i + 2] = p4v.z; ps[4 * i + 3] = p4v.w;
        }
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)  // the conv output is a bf16 tensor under autocast; cat / + pos_embed run in fp32
              o[e] = bf16_round(__uint_as_float(raw[4 * i + e]) + bia[4 * i + e]) + ps[4 * i + e];
            *reinterpret_cast<float4*>(xrow + sc * 16 + 4 * i) = make_float4(o[0], o[1], o[2], o[3]);
          }
          if (j == 1) {  // the first kept patch of a sample also writes the sample's class-token row (its column range)
            const float* c0 = p.cls + sc * 16;
            const float* p0 = p.pos + size_t(tok0) * D + sc * 16;
            float* x0 = p.x + size_t(bb) * p.num_keep * D + sc * 16;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(c0) + i), c = __ldg(reinterpret_cast<const float4*>(p0) + i);
              reinterpret_cast<float4*>(x0)[i] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&hdr->acc_free[b]);
    };

    uint32_t it = 0;
    int tile = blockIdx.x;
    if (tile < p.ntiles) produce(tile, 0);
    for (; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int nxt = tile + gridDim.x;
      if (nxt < p.ntiles) produce(nxt, it + 1);  // gather the next tile while the tensor pipe works on this one
      epilogue(tile, it);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace mae

using namespace mae;

extern "C" int mae_patch_embed_tokens(const float* images, const int64_t* idx_keep, int B, int C, int img, int patch,
                                      int num_keep, int D, const void* w_bf16, const float* bias, const float* cls,
                                      const float* pos, void* patches_bf16, float* x, cudaStream_t stream) {
  if (B <= 0) return MAE_OK;
  if (!images || !idx_keep || !w_bf16 || !bias || !cls || !pos || !patches_bf16 || !x) return MAE_ERR_NULL;
  const int pd = C * patch * patch;
  if (patch <= 0 || (patch & 3) || img % patch || num_keep < 2 || (pd % 16) || pd > 256 || D < 16 || D > 256 || (D % 16) ||
      (long long)B * (num_keep - 1) >= (1ll << 31))
    return MAE_ERR_SHAPE;
  PeParams p;
  p.B = B; p.C = C; p.img = img; p.patch = patch; p.num_keep = num_keep; p.D = D; p.pd = pd; p.nkb = (pd + BK - 1) / BK;
  p.grid = img / patch;
  p.mrows = (long long)B * (num_keep - 1);
  p.ntiles = int((p.mrows + BM - 1) / BM);
  p.images = images; p.idx_keep = idx_keep; p.bias = bias; p.cls = cls; p.pos = pos;
  p.patches = reinterpret_cast<__nv_bfloat16*>(patches_bf16); p.x = x;
  const size_t w_bytes = (size_t(p.nkb) * D * 128 + 1023) & ~size_t(1023);
  p.off_a = uint32_t(w_bytes);
  p.off_hdr = uint32_t(w_bytes + 2 * size_t(p.nkb) * 16384);
  const size_t smem = p.off_hdr + sizeof(PeHdr) + 1024;
  if (smem > 227 * 1024) return MAE_ERR_SHAPE;
  CUtensorMap mw;
  int rc = make_tmap_bf16(&mw, w_bf16, D, pd, pd, D);  // W (D, pd): boxes of [D rows x 64 k]
  if (rc) return rc;
  static bool attr_done[kMaxDevices] = {};
  if (!ensure_dyn_smem(patch_embed_kernel, 227 * 1024, attr_done)) return MAE_ERR_LAUNCH;
  const int grid = p.ntiles < num_sms() ? p.ntiles : num_sms();
  mae_launch(patch_embed_kernel, dim3(grid), dim3(kPeThreads), smem, stream, p, mw);
  return MAE_CHECK_LAUNCH();
}
