// GPU side of the reference's STL-10 input transform (src/data.py:15-34): RandomResizedCrop -> horizontal flip ->
// ToTensor -> Normalize(0.5, 0.5), from uint8 images as torchvision's STL10 stores them (N, C, H, W).
//
// The crop is resized exactly like Pillow's 8-bit ImagingResample with the bilinear filter (what torchvision's
// RandomResizedCrop calls for PIL inputs): separable, horizontal pass then vertical pass, per-output-pixel triangle
// weights computed in double precision, normalised, converted to 22-bit fixed point, accumulated in int32 from 1 << 21
// and clipped to 8 bits between the passes.  Integer path: bit-exact against Pillow / torchvision (tests/golden).
//
// One CTA per (sample, channel): the coefficient tables of both axes are built once in shared memory (double-precision
// IEEE ops through the _rn intrinsics so nothing is contracted into an FMA: the host C code Pillow is compiled from
// does not fuse either), the crop's rows are staged in shared memory with 16-byte loads, the horizontal pass writes a
// uint8 [h x S] tile, the vertical pass converts through a 256-entry table and stores float4 (reversed when flipped).
// Four outputs per thread iteration.  HBM-bound: 1 byte in, 4 bytes out per element.
#include "mae_internal.h"
#include "ptx.cuh"

namespace mae {

static constexpr int kAugPrecision = 32 - 8 - 2;
static constexpr int kAugMaxTaps = 7;  // 2 * ceil(scale) + 1 with scale = crop / out <= 3

struct AugAxis {  // coefficient table of one axis, in shared memory
  int* lo;        // [S]  first source index
  int* cnt;       // [S]  number of taps
  int* kk;        // [S * kAugMaxTaps] fixed-point weights
};

// Pillow precompute_coeffs + normalize_coeffs_8bpc for output index xx (bilinear filter, box = [0, in_size)).
__device__ __forceinline__ void aug_coeffs(int in_size, int out_size, int xx, int& lo, int& cnt, int* kk) {
  const double scale = __ddiv_rn(double(in_size), double(out_size));
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = filterscale;  // bilinear support 1.0 * filterscale
  const double ss = __ddiv_rn(1.0, filterscale);
  const double center = __dmul_rn(__dadd_rn(double(xx), 0.5), scale);
  int x0 = __double2int_rz(__dadd_rn(__dsub_rn(center, support), 0.5));
  if (x0 < 0) x0 = 0;
  int x1 = __double2int_rz(__dadd_rn(__dadd_rn(center, support), 0.5));
  if (x1 > in_size) x1 = in_size;
  int n = x1 - x0;
  if (n > kAugMaxTaps) n = kAugMaxTaps;  // unreachable for crops <= 3x the output (checked by the host wrapper)
  double w[kAugMaxTaps];
  double ww = 0.0;
#pragma unroll
  for (int x = 0; x < kAugMaxTaps; ++x) {
    w[x] = 0.0;
    if (x < n) {
      double t = __dmul_rn(__dadd_rn(__dsub_rn(double(x + x0), center), 0.5), ss);
      if (t < 0.0) t = -t;
      w[x] = t < 1.0 ? __dsub_rn(1.0, t) : 0.0;
      ww = __dadd_rn(ww, w[x]);
    }
  }
#pragma unroll
  for (int x = 0; x < kAugMaxTaps; ++x) {
    double v = w[x];
    if (x < n && ww != 0.0) v = __ddiv_rn(v, ww);
    // (int)(0.5 + v * 2^22): the product is exact (power of two), weights are non-negative for this filter
    kk[x] = x < n ? __double2int_rz(__dadd_rn(0.5, __dmul_rn(v, double(1 << kAugPrecision)))) : 0;
  }
  lo = x0;
  cnt = n;
}

__device__ __forceinline__ int aug_clip8(int v) {
  v >>= kAugPrecision;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// TAPS3 = true: both axes up-sample (crop <= output), so every output pixel has at most 3 taps: fixed-length, fully
// unrolled tap loops (missing taps have weight 0 and a clamped source index).  Outputs are produced four at a time
// (packed uint8 words between the passes, float4 stores at the end).
template <bool TAPS3>
__device__ __forceinline__ void aug_resample(const AugAxis& ax, const AugAxis& ay, const uint8_t* src_s, int src_pitch,
                                             int cj, int ch, int cw, int S, uint8_t* tmp_s, const float* lut, bool flip,
                                             float* __restrict__ dst) {
  const int Q = S >> 2;
  // horizontal pass: tmp[r][xx] = clip8(2^21 + sum_k src[r][cj + lo + k] * kx[xx][k])
#pragma unroll 2
  for (int i = threadIdx.x; i < ch * Q; i += blockDim.x) {
    const int r = i / Q, xq = i - r * Q;
    const uint8_t* row = src_s + r * src_pitch + cj;
    uint32_t packed = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int xx = 4 * xq + e;
      const int lo = ax.lo[xx];
      int acc = 1 << (kAugPrecision - 1);
      if (TAPS3) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int c = lo + k < cw ? lo + k : cw - 1;
          acc += int(row[c]) * ax.kk[xx * kAugMaxTaps + k];
        }
      } else {
        const int n = ax.cnt[xx];
        for (int k = 0; k < n; ++k) acc += int(row[lo + k]) * ax.kk[xx * kAugMaxTaps + k];
      }
      packed |= uint32_t(aug_clip8(acc)) << (8 * e);
    }
    *reinterpret_cast<uint32_t*>(tmp_s + r * S + 4 * xq) = packed;
  }
  __syncthreads();
  // vertical pass + ToTensor + Normalize (256-entry table) + flip
#pragma unroll 2
  for (int i = threadIdx.x; i < S * Q; i += blockDim.x) {
    const int yy = i / Q, xq = i - yy * Q;
    const int lo = ay.lo[yy];
    int acc[4] = {1 << (kAugPrecision - 1), 1 << (kAugPrecision - 1), 1 << (kAugPrecision - 1), 1 << (kAugPrecision - 1)};
    if (TAPS3) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int rr = lo + k < ch ? lo + k : ch - 1;
        const uint32_t wv = *reinterpret_cast<const uint32_t*>(tmp_s + rr * S + 4 * xq);
        const int c = ay.kk[yy * kAugMaxTaps + k];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] += int((wv >> (8 * e)) & 0xffu) * c;
      }
    } else {
      const int n = ay.cnt[yy];
      for (int k = 0; k < n; ++k) {
        const uint32_t wv = *reinterpret_cast<const uint32_t*>(tmp_s + (lo + k) * S + 4 * xq);
        const int c = ay.kk[yy * kAugMaxTaps + k];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] += int((wv >> (8 * e)) & 0xffu) * c;
      }
    }
    float4 v = make_float4(lut[aug_clip8(acc[0])], lut[aug_clip8(acc[1])], lut[aug_clip8(acc[2])], lut[aug_clip8(acc[3])]);
    if (flip) {
      v = make_float4(v.w, v.z, v.y, v.x);
      *reinterpret_cast<float4*>(dst + yy * S + (S - 4 - 4 * xq)) = v;
    } else {
      *reinterpret_cast<float4*>(dst + yy * S + 4 * xq) = v;
    }
  }
}

__global__ void __launch_bounds__(256) augment_u8_kernel(const uint8_t* __restrict__ images,
                                                        const int* __restrict__ params, int C, int Hs, int Ws, int S,
                                                        float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t smem_aug[];
  __shared__ float lut[256];
  const int b = blockIdx.x / C, c = blockIdx.x % C;
  int ci = params[b * 5 + 0], cj = params[b * 5 + 1], ch = params[b * 5 + 2], cw = params[b * 5 + 3];
  const bool flip = params[b * 5 + 4] != 0;
  // defensive clamps: a malformed box must not read out of bounds (the host wrapper validates host-side params)
  ch = ch < 1 ? 1 : (ch > Hs ? Hs : ch);
  cw = cw < 1 ? 1 : (cw > Ws ? Ws : cw);
  ci = ci < 0 ? 0 : (ci > Hs - ch ? Hs - ch : ci);
  cj = cj < 0 ? 0 : (cj > Ws - cw ? Ws - cw : cj);

  int* tab = reinterpret_cast<int*>(smem_aug);
  AugAxis ax{tab, tab + S, tab + 2 * S};
  AugAxis ay{tab + (2 + kAugMaxTaps) * S, tab + (3 + kAugMaxTaps) * S, tab + (4 + kAugMaxTaps) * S};
  uint8_t* src_s = smem_aug + size_t(2) * (2 + kAugMaxTaps) * S * sizeof(int);  // rows [ci, ci+ch) of the plane, pitch Ws
  uint8_t* tmp_s = src_s + ((size_t(Hs) * Ws + 15) & ~size_t(15));               // [ch x S] after the horizontal pass

  // the crop's rows: one contiguous, 16-byte aligned span of the plane when Ws % 16 == 0
  const uint8_t* plane = images + (size_t(b) * C + c) * Hs * Ws + size_t(ci) * Ws;
  if ((Ws & 15) == 0) {
    const uint4* g4 = reinterpret_cast<const uint4*>(plane);
    uint4* s4 = reinterpret_cast<uint4*>(src_s);
    const int n4 = (ch * Ws) >> 4;
#pragma unroll 3
    for (int i = threadIdx.x; i < n4; i += blockDim.x) s4[i] = __ldg(g4 + i);
  } else {
    for (int i = threadIdx.x; i < ch * Ws; i += blockDim.x) src_s[i] = plane[i];
  }
  // ToTensor + Normalize of an 8-bit value: ((v / 255) - 0.5) / 0.5, IEEE fp32 (x / 0.5 == x * 2 exactly)
  if (threadIdx.x < 256) lut[threadIdx.x] = __fmul_rn(__fsub_rn(__fdiv_rn(float(threadIdx.x), 255.0f), 0.5f), 2.0f);
  for (int e = threadIdx.x; e < 2 * S; e += blockDim.x) {
    const bool is_y = e >= S;
    const int xx = is_y ? e - S : e;
    const AugAxis& a = is_y ? ay : ax;
    int lo, cnt, kk[kAugMaxTaps];
    aug_coeffs(is_y ? ch : cw, S, xx, lo, cnt, kk);
    a.lo[xx] = lo; a.cnt[xx] = cnt;
#pragma unroll
    for (int k = 0; k < kAugMaxTaps; ++k) a.kk[xx * kAugMaxTaps + k] = kk[k];
  }
  __syncthreads();
  float* dst = out + (size_t(b) * C + c) * S * S;
  if (ch <= S && cw <= S)
    aug_resample<true>(ax, ay, src_s, Ws, cj, ch, cw, S, tmp_s, lut, flip, dst);
  else
    aug_resample<false>(ax, ay, src_s, Ws, cj, ch, cw, S, tmp_s, lut, flip, dst);
}

// ---------------------------------------------------------------------------------------------
// Crop boxes and flips drawn on the device: torchvision RandomResizedCrop.get_params (10 attempts of area ~ U(scale),
// log-ratio ~ U(log ratio), w = round(sqrt(a r)), h = round(sqrt(a / r)), first fitting one wins, whole-image /
// centre-crop fallback) followed by RandomHorizontalFlip's coin, one thread per sample.  The uniforms come from a
// counter-based generator, so a call is a pure function of (seed, counter, sample): no generator state, no host work, no
// host <-> device traffic in the input pipeline.  u(k) = top 24 bits of splitmix64(splitmix64(splitmix64(seed + counter)
// + sample) + k) * 2^-24; k = 2a, 2a+1: area / ratio of attempt a; 20, 21: top / left; 22: flip.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(128) augment_params_kernel(int* __restrict__ params, int B, int H, int W, float s0, float s1,
                                                            float l0, float l1, float rmin, float rmax,
                                                            unsigned long long seed, unsigned long long counter, int train) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int ci = 0, cj = 0, ch = H, cw = W, flip = 0;
  if (train) {
    const unsigned long long key = splitmix64(splitmix64(seed + counter) + (unsigned long long)b);
    auto u = [&](int k) { return float(splitmix64(key + (unsigned long long)k) >> 40) * (1.0f / 16777216.0f); };
    bool found = false;
    for (int a = 0; a < 10 && !found; ++a) {
      // torch: empty(1).uniform_(from, to) = u * (to - from) + from in fp32; .item() -> double
      const float sf = __fadd_rn(__fmul_rn(u(2 * a), __fsub_rn(s1, s0)), s0);
      const float lr = __fadd_rn(__fmul_rn(u(2 * a + 1), __fsub_rn(l1, l0)), l0);
      const double target = double(H * W) * double(sf);
      const double ar = double(expf(lr));
      const int w = int(rint(sqrt(target * ar))), h = int(rint(sqrt(target / ar)));  // Python round(): half to even
      if (w > 0 && w <= W && h > 0 && h <= H) {
        ci = int(double(u(20)) * double(H - h + 1));
        cj = int(double(u(21)) * double(W - w + 1));
        ch = h; cw = w;
        found = true;
      }
    }
    if (!found) {
      const float in_ratio = float(W) / float(H);
      if (in_ratio < rmin) { cw = W; ch = int(rint(double(W) / double(rmin))); }
      else if (in_ratio > rmax) { ch = H; cw = int(rint(double(H) * double(rmax))); }
      else { cw = W; ch = H; }
      ch = ch > H ? H : ch; cw = cw > W ? W : cw;
      ci = (H - ch) / 2; cj = (W - cw) / 2;
    }
    flip = u(22) < 0.5f ? 1 : 0;
  }
  int* o = params + size_t(b) * 5;
  o[0] = ci; o[1] = cj; o[2] = ch; o[3] = cw; o[4] = flip;
}

}  // namespace mae

using namespace mae;

extern "C" int mae_augment_params(int* params, int B, int H, int W, float scale_lo, float scale_hi, float ratio_lo,
                                  float ratio_hi, unsigned long long seed, unsigned long long counter, int train,
                                  cudaStream_t stream) {
  if (B <= 0) return MAE_OK;
  if (!params) return MAE_ERR_NULL;
  if (H <= 0 || W <= 0 || !(scale_lo > 0.f) || scale_hi < scale_lo || !(ratio_lo > 0.f) || ratio_hi < ratio_lo) return MAE_ERR_SHAPE;
  mae_launch(augment_params_kernel, dim3((B + 127) / 128), dim3(128), 0, stream, params, B, H, W, scale_lo, scale_hi,
             logf(ratio_lo), logf(ratio_hi), ratio_lo, ratio_hi, seed, counter, train);
  return MAE_CHECK_LAUNCH();
}

extern "C" int mae_augment_u8(const uint8_t* images, const int* params, int B, int C, int Hs, int Ws, int S, float* out,
                              cudaStream_t stream) {
  if (B <= 0) return MAE_OK;
  if (!images || !params || !out) return MAE_ERR_NULL;
  if (C <= 0 || Hs <= 0 || Ws <= 0 || S <= 0 || S > 512 || (S & 3) || Hs > 3 * S || Ws > 3 * S) return MAE_ERR_SHAPE;
  const size_t smem = size_t(2) * (2 + kAugMaxTaps) * S * sizeof(int) + ((size_t(Hs) * Ws + 15) & ~size_t(15)) + size_t(Hs) * S;
  if (smem > 200 * 1024) return MAE_ERR_SHAPE;
  static bool attr_done[mae::kMaxDevices] = {};
  if (smem > 48 * 1024 && !mae::ensure_dyn_smem(augment_u8_kernel, 200 * 1024, attr_done)) return MAE_ERR_LAUNCH;
  mae_launch(augment_u8_kernel, dim3(B * C), dim3(256), smem, stream, images, params, C, Hs, Ws, S, out);
  return MAE_CHECK_LAUNCH();
}
