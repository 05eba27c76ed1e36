// Device helpers shared by the tcgen05 GEMM kernels (gemm.cu, mlp_fused.cu): the bf16 GELU / GELU' table look-up, the
// 128-byte-swizzle staging helpers, TMA bulk stores and batched TMEM loads.
#pragma once
#include "mae_internal.h"
#include "ptx.cuh"

namespace mae {

static constexpr int BM = 128;  // UMMA M
static constexpr int BK = 64;   // one 128-byte swizzle span of bf16
static constexpr int kStgRow = 144;  // manual epilogue staging: 128 data bytes + 16 pad, conflict-free 16-byte accesses

// ---------------------------------------------------------------------------------------------
// erf-GELU by table (EPI 1).  The pre-activation is a bf16 value, so GELU and GELU' are pure functions of 16 bits:
// for 2^-20 <= |x| <= 16 both correctly rounded bf16 results come from a 2 x 3073-entry table (entry = gelu | gelu'<<16,
// built on the host in double precision from erf), staged in shared memory; outside that range the results are
// trivial (x/2 and 1/2; x and 1; -0 and 0) and are patched by a rarely taken warp-uniform branch.
// Shared-memory addressing: entry of the raw 16-bit pattern i (sign included) lives at tab_base + 4*(i - kGeluLo),
// i.e. the table of the negative inputs sits exactly 0x8000 entries (128 KB) above the positive one.
// ---------------------------------------------------------------------------------------------
static constexpr uint32_t kGeluLo = 107u << 7;                  // bf16 bits of 2^-20
static constexpr uint32_t kGeluHi = 131u << 7;                  // bf16 bits of 16.0
static constexpr uint32_t kGeluNT = kGeluHi - kGeluLo + 1;       // 3073 entries per sign
static constexpr uint32_t kGeluTabBytes = kGeluNT * 4;
static constexpr uint32_t kGeluNegOff = 0x8000u * 4;             // 131072

__device__ __forceinline__ uint32_t lds32(uint32_t saddr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
// out-of-table inputs (|x| < 2^-20, |x| > 16, inf, nan): entry = gelu | gelu' << 16
static __device__ __noinline__ uint32_t gelu_entry_slow(uint32_t bits16, uint32_t tab_s) {
  const uint32_t a = bits16 & 0x7fffu;
  if (a >= kGeluLo && a <= kGeluHi) return lds32(tab_s + bits16 * 4);
  if (a < kGeluLo) {  // gelu = x/2 (exact), gelu' = 1/2
    const float x = __uint_as_float(bits16 << 16);
    return uint32_t(__bfloat16_as_ushort(__float2bfloat16_rn(0.5f * x))) | (0x3f00u << 16);
  }
  if (a > 0x7f80u) return bits16 | (bits16 << 16);                    // nan propagates
  return (bits16 & 0x8000u) ? 0x8000u : (bits16 | (0x3f80u << 16));   // x <= -16: (-0, 0);  x >= 16: (x, 1)
}
// 8 packed bf16 pairs of pre-activations -> h (gelu) and d (gelu') pairs.  tab_s = shared address of the positive table
// minus 4*kGeluLo.
__device__ __forceinline__ void gelu_table8(const uint32_t (&pre)[8], uint32_t tab_s, uint32_t (&h)[8], uint32_t (&d)[8]) {
  constexpr uint32_t lo2 = kGeluLo | (kGeluLo << 16), hi2 = kGeluHi | (kGeluHi << 16);
  uint32_t bad = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t a2 = pre[i] & 0x7fff7fffu;
    const uint32_t c2 = __vminu2(__vmaxu2(a2, lo2), hi2);  // per-halfword clamp into the table
    bad |= c2 ^ a2;
    const uint32_t k2 = c2 | (pre[i] & 0x80008000u);
    const uint32_t e0 = lds32(tab_s + __byte_perm(k2, 0, 0x4410) * 4), e1 = lds32(tab_s + __byte_perm(k2, 0, 0x4432) * 4);
    h[i] = __byte_perm(e0, e1, 0x5410);
    d[i] = __byte_perm(e0, e1, 0x7632);
  }
  if (__any_sync(0xffffffffu, bad != 0)) {  // rare: some lane holds an input outside the table
#pragma unroll
    for (int i = 0; i < 8; ++i) {  // (fully unrolled: dynamic indexing would push the arrays to local memory)
      const uint32_t e0 = gelu_entry_slow(pre[i] & 0xffffu, tab_s), e1 = gelu_entry_slow(pre[i] >> 16, tab_s);
      h[i] = __byte_perm(e0, e1, 0x5410);
      d[i] = __byte_perm(e0, e1, 0x7632);
    }
  }
}

// TMA bulk tensor store of a [32 rows x 64 bf16] tile (128-byte swizzled in smem) -- SASS: UTMASTG
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// 16-byte chunk `c` (0..7) of row `t` inside a 128-byte-swizzled [32 x 128 B] tile
__device__ __forceinline__ uint32_t swz128(uint32_t base, int t, int c) { return base + t * 128 + ((c ^ (t & 7)) << 4); }
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Device address of the 2 x kGeluNT-entry GELU table of the current device (built on first use); nullptr on failure.
// gemm.cu owns the table; kernels in other translation units take the pointer as an argument.
const uint32_t* gelu_table_device_ptr();

}  // namespace mae
