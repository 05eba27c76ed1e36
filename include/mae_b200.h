/* libmae_b200.so -- C ABI of the B200-native MAE pre-training hot path (sm_100a only).
 *
 * The reference (giolucasd/ssrl-vit-mae-jepa) has no FFI layer: its hot path is Python
 * (src/models/mae.py, src/training/mae.py) calling lightly 1.5.22 / timm 1.0.21 / ATen.  This header
 * is the boundary a maintainer would bind underneath those Python classes (see INTEGRATION.md for
 * the ctypes stub).  Each entry point names the reference call it replaces.
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers owned by the caller
 *  - no allocation, no synchronisation: every call only enqueues work on `stream` (re-entrant,
 *    stream-ordered, CUDA-graph capturable)
 *  - returns MAE_OK (0) or a negative MAE_ERR_* code; never throws; there is NO CPU fallback
 *  - bf16 tensors are passed as `void*` (raw __nv_bfloat16 storage); indices are int64 as in torch
 *  - all matrices are dense row-major
 */
#ifndef MAE_B200_H_
#define MAE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* mae_stream_t; /* == cudaStream_t */

enum {
  MAE_OK = 0,
  MAE_ERR_SHAPE = -1,  /* unsupported / inconsistent dimension */
  MAE_ERR_NULL = -2,   /* required pointer is null */
  MAE_ERR_ALIGN = -3,  /* pointer or leading dimension not 16-byte aligned */
  MAE_ERR_DRIVER = -4, /* driver entry point / tensor-map encode failed */
  MAE_ERR_LAUNCH = -5  /* cudaGetLastError() after launch */
};

/* epilogues of mae_gemm_tn */
enum {
  MAE_EPI_BF16 = 0,       /* out(bf16) = acc + bias                                  (qkv, embed, pred, dgrads) */
  MAE_EPI_GELU_BF16 = 1,  /* pre=bf16(acc+bias); out(bf16)=gelu_erf(pre); pre_out(bf16)=gelu_erf'(pre) (mlp.fc1) */
  MAE_EPI_RESID_F32 = 2,  /* out(f32) = aux(f32) + acc + bias                        (attn.proj, mlp.fc2)       */
  MAE_EPI_DGELU_BF16 = 4  /* out(bf16) = acc * aux(bf16 gelu' saved by EPI 1)         (fc2 dgrad)                */
};

int mae_version(void);
const char* mae_error_string(int code);

/* lightly.models.utils.random_token_mask (called at reference src/models/mae.py:79-83), with the
 * torch.rand draw hoisted to the caller: noise (B,L) f32.  Column 0 (class token) is treated as -1.
 * Stable ascending sort per row (ties -> lower index); idx_keep (B,num_keep), idx_mask (B,L-num_keep)
 * int64.  L <= 256.  Integer path: bit-exact. */
int mae_random_token_mask(const float* noise, int B, int L, int num_keep, int64_t* idx_keep, int64_t* idx_mask,
                          mae_stream_t stream);

/* The im2col half of timm PatchEmbed (Conv2d k=s=patch; reference src/models/mae.py:28-36,55) fused with
 * get_at_index(idx_keep): for kept tokens j=1..num_keep-1 writes patch idx_keep[b,j]-1 of image b in conv
 * weight order (c,py,px) as bf16 row b*(num_keep-1)+j-1 of patches (B*(num_keep-1), C*patch*patch). */
int mae_gather_patches(const float* images, const int64_t* idx_keep, int B, int C, int img, int patch, int num_keep,
                       void* patches_bf16, mae_stream_t stream);

/* Patch embedding of the kept tokens in ONE launch (timm PatchEmbed -> cat(cls_token) -> + pos_embed -> get_at_index,
 * reference src/models/mae.py:54-55; replaces mae_gather_patches + mae_gemm_tn + mae_assemble_tokens): the im2col rows are
 * gathered straight into the A tile of a tcgen05 GEMM (W (D, C*patch*patch) bf16 resident in shared memory) and the
 * epilogue writes  x (B,num_keep,D) f32 = bf16(patch . W^T + bias) + pos[idx_keep], class-token rows (cls + pos) included.
 * patches_bf16 (B*(num_keep-1), C*patch*patch) is written as by mae_gather_patches (the weight gradient reads it).
 * patch % 4 == 0, C*patch*patch <= 256 and % 16 == 0, D <= 256 and % 16 == 0. */
int mae_patch_embed_tokens(const float* images, const int64_t* idx_keep, int B, int C, int img, int patch, int num_keep,
                           int D, const void* w_bf16, const float* bias, const float* cls, const float* pos,
                           void* patches_bf16, float* x, mae_stream_t stream);

/* cat(cls_token, tokens) + pos_embed restricted to kept tokens (lightly MaskedVisionTransformerTIMM.
 * preprocess + get_at_index; reference src/models/mae.py:55): x (B,num_keep,D) f32. */
int mae_assemble_tokens(const void* tok_bf16, const float* cls, const float* pos, const int64_t* idx_keep, int B,
                        int num_keep, int D, float* x, mae_stream_t stream);

/* nn.LayerNorm(D, eps) forward over `rows` rows (timm Block.norm1/norm2, vit.norm, decoder_norm).  If
 * gather_idx != NULL, output row r=(b,j) reads source row b*L_src + gather_idx[b*n_sel+j] (get_at_index at
 * reference src/models/mae.py:72 commuted with the row-wise norm).  Writes xn bf16 (rows,D) and fp32
 * mean/rstd (rows). */
int mae_layernorm_fwd(const float* x, const float* gamma, const float* beta, const int64_t* gather_idx, int n_sel,
                      int L_src, int rows, int D, float eps, void* xn_bf16, float* mean, float* rstd,
                      mae_stream_t stream);

/* LayerNorm backward.  g bf16 (rows,D) is the gradient w.r.t. the normalised output; x/mean/rstd as in the
 * forward (same gather mapping; dx_out rows are written at the mapped row).  dx_out = (dx_in ? dx_in : 0) +
 * LN'(g); optional bf16 copy of dx_out; dgamma/dbeta (D) are ACCUMULATED with fp32 atomics. */
int mae_layernorm_bwd(const void* g_bf16, const float* x, const float* mean, const float* rstd, const float* gamma,
                      const float* dx_in, const int64_t* gather_idx, int n_sel, int L_src, int rows, int D,
                      float* dx_out, void* dx_out_bf16, float* dgamma, float* dbeta, mae_stream_t stream);

/* The two LayerNorms around the masked-rows-only tail of the last decoder block (get_at_index(idx_mask) at reference
 * src/models/mae.py:72 commuted backwards through decoder_norm, mlp and norm2 of the last block: the rows it drops are
 * never computed).
 * mae_layernorm_fwd_add: row r normalises  x[map(r)] + delta[r]  (delta bf16 (rows,D): the mlp.fc2 output before the
 *   residual add, x the fp32 stream gathered like mae_layernorm_fwd); the sum is also written to x_out (rows,D) f32.
 * mae_layernorm_bwd_scatter: mae_layernorm_bwd with the residual gradient dx_in indexed by the compact row r (rows,D)
 *   instead of the mapped row; dx_out rows are written at the mapped rows.  D <= 256. */
int mae_layernorm_fwd_add(const float* x, const void* delta_bf16, const float* gamma, const float* beta,
                          const int64_t* gather_idx, int n_sel, int L_src, int rows, int D, float eps, float* x_out,
                          void* xn_bf16, float* mean, float* rstd, mae_stream_t stream);
int mae_layernorm_bwd_scatter(const void* g_bf16, const float* x, const float* mean, const float* rstd,
                              const float* gamma, const float* dx_in_compact, const int64_t* gather_idx, int n_sel,
                              int L_src, int rows, int D, float* dx_out, void* dx_out_bf16, float* dgamma,
                              float* dbeta, mae_stream_t stream);

/* out = epi(A[M,K] * B[N,K]^T): every nn.Linear forward (B = weight (out,in)) and every dgrad (B = weight^T)
 * of timm Block / MAEDecoderTIMM (reference src/models/mae.py:55,59,71,73).  A, B bf16; fp32 accumulation in
 * TMEM via tcgen05.mma; K % 16 == 0, N % 16 == 0. */
int mae_gemm_tn(const void* A_bf16, const void* B_bf16, int M, int N, int K, int epi, const float* bias, void* out,
                void* pre_out_bf16, const void* aux, mae_stream_t stream);

/* Second half of a timm Block, forward, in ONE launch:  x_out = x + fc2(GELU(fc1(LayerNorm(x))))  (timm Block.norm2 /
 * Block.mlp as built at reference src/models/mae.py:28-36 and run at :54-55,71; replaces mae_layernorm_fwd +
 * mae_gemm_tn(EPI_GELU) + mae_gemm_tn(EPI_RESID) and their HBM round trips of xn / h / the residual).
 * x f32; if gather_idx != NULL, row r=(b,j) reads source row b*L_src + gather_idx[b*n_sel+j] and every output is
 * compact (rows,*) (the masked-rows-only tail of the last decoder block).  w1 bf16 (4D,D), w2 bf16 (D,4D), biases f32.
 * Saved for the backward: xn bf16 (rows,D), mean/rstd f32 (rows), h = GELU(pre) bf16 (rows,4D), dgelu = GELU'(pre) bf16
 * (rows,4D) (pre = bf16(fc1 output), exactly as EPI_GELU).  x_out f32 (rows,D).  64 <= D <= 192, D % 16 == 0. */
int mae_mlp_fwd(const float* x, const int64_t* gather_idx, int n_sel, int L_src, int rows, int D, const float* gamma,
                const float* beta, float eps, const void* w1_bf16, const float* b1, const void* w2_bf16, const float* b2,
                void* xn_bf16, float* mean, float* rstd, void* h_bf16, void* dgelu_bf16, float* x_out,
                mae_stream_t stream);

/* out[p*so_p + q*so_q] += sum_m P[m,p] * Q[m,q]: weight gradients (autograd of the Linears above).
 * P bf16 (M,Pc), Q bf16 (M,Qc), Qc % 16 == 0, Pc % 8 == 0.  fp32 atomics: zero `out` first.
 * If colsum != NULL the same launch also accumulates the bias gradient: colsum[c] += sum_m P[m,c]
 * (colsum_of_q == 0) or sum_m Q[m,c] (colsum_of_q != 0), read from the tiles already staged in shared memory. */
int mae_gemm_wgrad(const void* P_bf16, int Pc, const void* Q_bf16, int Qc, int M, float* out, long long so_p,
                   long long so_q, float* colsum, int colsum_of_q, mae_stream_t stream);

/* timm Attention core: F.scaled_dot_product_attention on qkv (B,L,3,H,hd) bf16 -> out (B,L,H*hd) bf16,
 * scale hd^-0.5, plus lse (B,H,L) f32 for the backward.  hd in {24,32,64}; L <= 208. */
int mae_attention_fwd(const void* qkv_bf16, int B, int L, int H, int hd, void* out_bf16, float* lse,
                      mae_stream_t stream);
int mae_attention_bwd(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse, int B,
                      int L, int H, int hd, void* dqkv_bf16, mae_stream_t stream);
/* The same op when only the first q_rows query tokens matter (the caller keeps the rows it needs first, e.g. the last
 * decoder block of the MAE, whose kept tokens are keys / values only).  Forward: out / lse rows >= q_rows are unspecified
 * (the decoder-sized tcgen05 kernel skips its second query tile for q_rows <= 128 and writes zeros to out rows >= 128).
 * Backward: the caller guarantees dout == 0 on rows >= q_rows (kernels are free to skip them: the tcgen05 kernel does and
 * ignores whatever is there; shapes served by other kernels run the plain backward); dqkv is complete (zero dQ there). */
int mae_attention_fwd_q(const void* qkv_bf16, int B, int L, int H, int hd, int q_rows, void* out_bf16, float* lse,
                        mae_stream_t stream);
int mae_attention_bwd_q(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse, int B,
                        int L, int H, int hd, int q_rows, void* dqkv_bf16, mae_stream_t stream);
/* Same op with a caller-provided scratch buffer (>= B*L*H*4 bytes, 16-byte aligned): kernels that want
 * delta = rowsum(dout o out) per (token, head) compute it there with a coalesced pre-pass instead of per-thread row
 * gathers inside the attention kernel (decoder-sized tcgen05 backward: ~0.9x the time).  ws may be NULL. */
int mae_attention_bwd_ws(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse, int B,
                         int L, int H, int hd, void* dqkv_bf16, void* ws, long long ws_bytes, mae_stream_t stream);

/* repeat_token(mask_token) + set_at_index(idx_keep, embed) + decoder_pos_embed add (reference
 * src/models/mae.py:61-69 and MAEDecoderTIMM.decode's pos add): x_dec (B,L,D) f32. */
int mae_scatter_decoder_tokens(const void* emb_bf16, const float* mask_token, const float* dec_pos,
                               const int64_t* idx_keep, const int64_t* idx_mask, int B, int L, int num_keep, int D,
                               float* x_dec, mae_stream_t stream);
/* Same tokens, stored as [mask tokens in idx_mask order | kept tokens in idx_keep order] per sample (row r < L - num_keep
 * holds token idx_mask[r], row L - num_keep + i holds token idx_keep[i]): the rows the prediction head reads come first,
 * so the last decoder block can use mae_attention_fwd_q / bwd_q with q_rows = L - num_keep. */
int mae_scatter_decoder_tokens_grouped(const void* emb_bf16, const float* mask_token, const float* dec_pos,
                               const int64_t* idx_keep, const int64_t* idx_mask, int B, int L, int num_keep, int D,
                               float* x_dec, mae_stream_t stream);

/* mae_scatter_decoder_tokens fused with the first decoder block's norm1 (timm Block.norm1 on the freshly assembled
 * stream): additionally writes xn bf16 (B*L,D) and the row statistics mean / rstd (B*L) like mae_layernorm_fwd, so the
 * stream is not read back by a separate launch.  D <= 256. */
int mae_scatter_decoder_tokens_ln(const void* emb_bf16, const float* mask_token, const float* dec_pos,
                                  const int64_t* idx_keep, const int64_t* idx_mask, int B, int L, int num_keep, int D,
                                  const float* gamma, const float* beta, float eps, float* x_dec, void* xn_bf16,
                                  float* mean, float* rstd, mae_stream_t stream);

/* patchify + get_at_index(idx_mask-1) + MSELoss (reference src/models/mae.py:90-92, src/training/mae.py:40,48):
 * pred bf16 (B,n_mask,C*patch*patch); target read straight from the images in (py,px,c) order.
 * loss (1) f32 = mean((pred-target)^2) (deterministic two-stage reduction through `partials`, >= 4096 floats);
 * optional dpred bf16 = 2*(pred-target)/numel; optional target f32 (B,n_mask,C*p*p).
 * norm_pix != 0 normalises each target patch by its own mean/var (opt-in; the reference does not). */
int mae_masked_mse(const void* pred_bf16, const float* images, const int64_t* idx_mask, int B, int C, int img,
                   int patch, int n_mask, int norm_pix, float* loss, float* partials, void* dpred_bf16,
                   float* target, mae_stream_t stream);

/* out[n] += sum_m x[m,n] (bias gradients). x bf16 (M,N). */
int mae_colsum_bf16(const void* x_bf16, int M, int N, float* out, mae_stream_t stream);

/* Row gather + cast: out bf16 (B*n_sel, D) <- src f32 row b*L_src + (idx ? idx[b*n_sel+j] : row_offset+j).
 * (autograd of set_at_index / cat at reference src/models/mae.py:64-69). */
int mae_gather_rows_bf16(const float* src, const int64_t* idx, int B, int n_sel, int L_src, int row_offset, int D,
                         void* out_bf16, mae_stream_t stream);

/* dst rows b*L_src + idx[b*n_sel+j] <- 0, in the f32 buffer and (optionally) its bf16 copy: the kept rows of the
 * decoder stream get no gradient from the prediction head (autograd of get_at_index(idx_mask) at reference
 * src/models/mae.py:72); replaces two whole-buffer memsets. */
int mae_zero_rows(float* dst, void* dst_bf16, const int64_t* idx, int B, int n_sel, int L_src, int D,
                  mae_stream_t stream);

/* out[d] += sum over the same row selection of src[.,d]  (mask_token / cls_token gradients). */
int mae_colsum_rows_f32(const float* src, const int64_t* idx, int B, int n_sel, int L_src, int row_offset, int D,
                        float* out, mae_stream_t stream);

/* dst(bf16)[i] = src(f32)[i] */
int mae_cast_bf16(const float* src, void* dst_bf16, long long n, mae_stream_t stream);

/* Multi-tensor transposing cast: for each entry t of table (n_entries x 4 int64: src_off, rows, cols, dst_off;
 * device memory) writes dst[dst_off + c*rows + r] = bf16(src[src_off + r*cols + c]).  max_rows / max_cols bound the
 * entries' shapes (they size the launch grid). */
int mae_transpose_cast_bf16(const float* src, void* dst_bf16, const long long* table, int n_entries, int max_rows,
                            int max_cols, mae_stream_t stream);

/* Lightning gradient_clip_val=1.0 ("norm") + torch.optim.AdamW step (reference
 * scripts/training/pretrain_mae.py:124-125, src/training/mae.py:59-65) over one flat fp32 buffer of n
 * elements.  Two launches: mae_grad_sumsq writes per-block partial sums; mae_clip_adamw reduces them in a
 * fixed order, computes coef = min(1, max_norm/(norm+1e-6)), applies decoupled weight decay + Adam and
 * refreshes the bf16 shadow copy of the parameters.  grad_norm_out (1) f32 receives the pre-clip norm. */
int mae_grad_sumsq(const float* g, long long n, float* partials, int max_partials, mae_stream_t stream);
int mae_clip_adamw(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, float max_norm, int step, const float* partials,
                   int n_partials, float* grad_norm_out, mae_stream_t stream);

/* Data-parallel gradient exchange over NVLink peer memory, fused with the first pass of the global-norm clip:
 *   out[offset + i] = (1/world) * sum_r peers[r][offset + i],  i < n;   partials[b] = sum of out^2 over block b's elements.
 * peer_ptrs_dev: DEVICE array of `world` pointers to the ranks' flat fp32 gradient buffers (symmetric memory: peer-mapped
 * over NVLink; entry `rank` is the local buffer).  Fixed summation order r = 0..world-1: bit-identical on every rank.
 * The caller orders it behind a cross-rank barrier (every rank's bucket is complete) and must not reuse the gradient
 * buffers before a second barrier.  n, offset % 4 == 0; world <= 16; writes `nblocks` partials (feed mae_clip_adamw).
 * Replaces ncclAllReduce(avg) + mae_grad_sumsq (reference: devices=1, scripts/training/pretrain_mae.py:116-118). */
int mae_allreduce_avg_sumsq(const void* peer_ptrs_dev, int world, long long offset, long long n, float* out,
                            float* partials, int nblocks, mae_stream_t stream);

/* The input transform that precedes the hot path (reference src/data.py:15-34), on the GPU: for sample b, crop
 * rows [i, i+h) x cols [j, j+w) of the uint8 image (B,C,Hs,Ws; torchvision STL10.data layout), resize to S x S exactly
 * like Pillow's 8-bit bilinear ImagingResample (what RandomResizedCrop does to a PIL image: two separable fixed-point
 * passes with a uint8 intermediate), flip horizontally if flip != 0, then ToTensor (/255) and Normalize((x-.5)/.5).
 * params (B,5) int32 = (i, j, h, w, flip) per sample, device memory (the random draws of RandomResizedCrop.get_params
 * and RandomHorizontalFlip stay with the caller, like the torch.rand of the mask).  The evaluation transform
 * (Resize(96)+CenterCrop(96) on 96x96 images = identity) is (0, 0, Hs, Ws, 0).  out (B,C,S,S) f32.  Integer path:
 * bit-exact.  S % 4 == 0; the source may be at most 3x the output size per axis. */
int mae_augment_u8(const uint8_t* images, const int* params, int B, int C, int Hs, int Ws, int S, float* out,
                   mae_stream_t stream);

/* Crop boxes + flips for mae_augment_u8 drawn ON THE DEVICE: torchvision RandomResizedCrop.get_params (reference
 * src/data.py:18-21: scale (0.8, 1.0), default ratio (3/4, 4/3)) + RandomHorizontalFlip's coin for B samples of an
 * H x W image -> params (B,5) int32 rows (top, left, height, width, flip).  Counter-based generator: the result is a pure
 * function of (seed, counter, sample index), see csrc/augment.cu; train == 0 gives the identity box (evaluation).
 * Keeps the host out of the per-step input pipeline (no RNG calls, no pinned staging copy). */
int mae_augment_params(int* params, int B, int H, int W, float scale_lo, float scale_hi, float ratio_lo, float ratio_hi,
                       unsigned long long seed, unsigned long long counter, int train, mae_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MAE_B200_H_ */
