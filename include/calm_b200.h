/* calm_b200 — C ABI of the B200 (sm_100a) kernel library behind the CALM-ViT drop-in modules.
 *
 * The reference (focegueda1998/CALM-ViT-DTE) has no FFI of its own: its hot path is stock PyTorch eager ops inside
 * CALM-ViT/Vi_Tools_CNN_less_V2.py and CALM-ViT/CALM_ViT_V2.py. Each entry point below replaces the aten op(s) that the
 * cited reference lines dispatch; the Python modules of the same names in calm-vit-dte_b200/ call these through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch caching allocator); the library allocates nothing,
 *     keeps no pointer across calls, never synchronises the device and only enqueues on `stream`;
 *   - return value 0 = success, <0 = error (calm_last_error() has the text); nothing throws or exits;
 *   - "bf16" = __nv_bfloat16 bits, "f32" = IEEE float; row-major with the feature axis innermost unless stated.
 */
#ifndef CALM_B200_H_
#define CALM_B200_H_
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define CALM_ABI_VERSION 1

enum { CALM_BF16 = 0, CALM_F32 = 1 };
enum { CALM_MAJOR_K = 0, CALM_MAJOR_MN = 1 };
enum { CALM_EPI_NONE = 0, CALM_EPI_GELU = 1, CALM_EPI_DGELU = 2 };
/* calm_gemm_args.flags — PER-CALL schedule selectors for tests and tuning (0 = the heuristics). The library keeps no global
 * mutable state: there is no process-wide switch that selects kernels. */
enum { CALM_GEMM_SIMT = 1 /* reference SIMT kernel (bring-up / bisecting) */, CALM_GEMM_NO_CLUSTER = 4, CALM_GEMM_FORCE_CLUSTER = 8,
       CALM_GEMM_PAIR_MULTICAST = 16 /* clusters use cta_group::1 + multicast B instead of cta_group::2 */,
       CALM_GEMM_DIRECT_EPILOGUE = 32 /* per-thread global loads/stores instead of the TMA-staged epilogue */ };

int32_t calm_abi_version(void);
const char* calm_last_error(void); /* thread-local, valid until the next failing call on this thread */

/* ------------------------------------------------------------------------------------------------------------------
 * GEMM  C[b](M,N) = epi(alpha * A[b](M,K) . B[b](N,K)^T + bias[n] + addend[b](m,n))        tcgen05 + TMA + TMEM
 * replaces: every sn(Linear) forward/backward contraction (Vi_Tools_CNN_less_V2.py:226-231,251-267,276-277,300,
 * 305-308,312), the mask logits bmm (:288-290) and linear_mask (:189-194,290), plus their autograd backward.
 *   a_major/b_major: CALM_MAJOR_K  -> element (r,k) at ptr[r*ld + k];  CALM_MAJOR_MN -> element (r,k) at ptr[k*ld + r]
 *   stride_* = batch strides in elements (0 = operand shared by all batch entries)
 *   reduce_batch=1: C = sum over b (the contraction runs over batch x K)
 *   splits>1: fp32 partial sums, partial s at c + s*stride_split (epilogue must be NONE; caller reduces)
 *   epilogue GELU : u = bf16(pre-activation); C <- gelu_erf(u), aux <- bf16(gelu_erf'(u))   (mlp.0 / linear_mask.0 forward)
 *   epilogue DGELU: C <- acc * aux      (their dgrad: aux is the derivative the forward epilogue saved)
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
  const void* a; const void* b; void* c;
  int32_t M, N, K, batch;
  int64_t lda, ldb, ldc;
  int64_t stride_a, stride_b, stride_c;
  int32_t a_major, b_major, c_dtype, epilogue;
  const float* bias;
  const void* addend; int32_t addend_dtype; int32_t bn_override /* tuning: force the N-tile width, 0 = automatic */; int64_t ld_addend, stride_addend;
  void* aux; int64_t ld_aux, stride_aux;
  int32_t reduce_batch, splits; int64_t stride_split;
  float alpha; int32_t flags /* CALM_GEMM_*, 0 = default */;
} calm_gemm_args;
int32_t calm_gemm(const calm_gemm_args* args, cudaStream_t stream);
int32_t calm_gemm_default_splits(int32_t M, int32_t N, int32_t K, int32_t batch, int32_t reduce_batch);

/* ------------------------------------------------------------------------------------------------------------------
 * Spectral norm, batched over a table of layers: replaces torch/nn/utils/spectral_norm.py:92-114 (one power iteration
 * v<-norm(W^T u), u<-norm(W v), sigma=u^T W v, W/sigma) for every sn(...) layer hit by a forward
 * (Vi_Tools_CNN_less_V2.py:137-204,380-384; CALM_ViT_V2.py:50-52,62-66) — 5 launches per scope (~40 layers) instead of ~12 per layer.
 * One table row per layer (device array of calm_sn_layer):
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
  const float* w;      /* weight_orig, (rows, cols) row-major fp32 (conv weights flattened to (out, in*kh*kw)) */
  float* u;            /* weight_u (rows)  - updated in place when training                                    */
  float* v;            /* weight_v (cols)  - updated in place when training                                    */
  const float* rowscale; /* optional LayerScale vector folded into the effective weight rows (or NULL)        */
  void* w_eff;         /* out: bf16 (rows, cols) = rowscale[r] * W/sigma  - or fp32 when eff_f32               */
  float* grad_w;       /* sn_backward: out fp32 (rows, cols) gradient wrt weight_orig                           */
  float* grad_rowscale;/* sn_backward: out fp32 (rows) gradient wrt rowscale (or NULL)                          */
  const float* g_eff;  /* sn_backward: in fp32 partial sums of dL/dW_eff: g_splits x (rows, cols), g_split_stride apart */
  float* tpart;        /* scratch: item_count x cols floats (partial W^T u per row chunk; partial dots in backward) */
  float* svec;         /* scratch: rows floats (W v)                                                            */
  float* sigma;        /* out: 1 float                                                                         */
  int64_t g_split_stride; /* elements between split partials of g_eff (layers stacked into one GEMM operand share it) */
  int32_t rows, cols, g_splits, eff_f32;
  int32_t item_count;  /* number of row-chunk items of this layer                                              */
  int32_t _pad;
} calm_sn_layer;
/* one CTA-sized piece of work: rows [row_begin, row_end) of layer `layer`; local_index = its rank within the layer */
typedef struct { int32_t layer, row_begin, row_end, local_index; } calm_sn_item;
int32_t calm_sn_forward(const calm_sn_layer* table_dev, int32_t n_layers, const calm_sn_item* items_dev, int32_t n_items,
                        int32_t training, float eps, cudaStream_t stream);
/* dL/dW_orig = rs*G/sigma - (<rs*G, W>/sigma^2) u v^T ; dL/drowscale[r] = sum_c G[r,c] * W[r,c]/sigma   */
int32_t calm_sn_backward(const calm_sn_layer* table_dev, int32_t n_layers, const calm_sn_item* items_dev, int32_t n_items,
                         cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Weight-only LayerNorm (eps 1e-6, no bias): ln_q / ln_kv / ln_2 / ln_final (Vi_Tools_CNN_less_V2.py:131-132,197,494)
 *   fwd: x f32 (rows, D) -> y bf16 (rows, D) [+ y_f32 optional], mean/rstd f32 (rows)
 *   bwd: dx f32 = LN'(dy) (+ dres) [+ the same values as bf16 in dx_bf16 when non-NULL: the operand form the next
 *        dgrad / wgrad GEMM reads], dw partials (nparts, D) to be summed by the caller-visible finalize
 * ------------------------------------------------------------------------------------------------------------------ */
int32_t calm_layernorm_fwd(const float* x, const float* w, void* y, int32_t y_dtype, float* mean, float* rstd,
                           int64_t rows, int32_t D, float eps, cudaStream_t stream);
/* first block: row tokenisation of the (B,3,S,S) image (Vi_Tools_CNN_less_V2.py:389-391) fused into its first LayerNorm:
 * tokens f32 (B, S, 3S) = img.permute(0,2,3,1).reshape(B, S, 3S) and y bf16 = LN(tokens) * w in one pass */
int32_t calm_layernorm_fwd_image(const float* img, const float* w, void* y_bf16, float* tokens, float* mean, float* rstd,
                                 int32_t B, int32_t S, float eps, cudaStream_t stream);
int32_t calm_layernorm_bwd(const void* dy, int32_t dy_dtype, const float* x, const float* w, const float* mean,
                           const float* rstd, const float* dres, float* dx, void* dx_bf16, float* dw_partial, int32_t nparts,
                           float* dw, int64_t rows, int32_t D, cudaStream_t stream);
int32_t calm_layernorm_bwd_parts(int64_t rows, int32_t D);

/* ------------------------------------------------------------------------------------------------------------------
 * RoPE with learned inv_freq (Vi_Tools_CNN_less_V2.py:80-95) fused with the per-head content|rope concat (:278-285).
 *   out[t, h, 0:dc]     = content[t, h, 0:dc]                      (dc = 0 for the non-reduce blocks)
 *   out[t, h, dc:dc+dr] = rope(ropein[t, h, 0:dr]) at position s = t % S
 * bwd also returns d inv_freq partials.
 * ------------------------------------------------------------------------------------------------------------------ */
int32_t calm_rope_fwd(const void* content, int64_t ld_content, const void* ropein, int64_t ld_rope, void* out, int64_t ld_out,
                      const float* inv_freq /* (dr/2) learned frequencies: cos / sin of position * inv_freq are formed in the kernel */,
                      int64_t tokens, int32_t S, int32_t heads, int32_t dc, int32_t dr, cudaStream_t stream);
int32_t calm_rope_bwd_scratch_floats(int32_t S, int32_t dr);
int32_t calm_rope_bwd(const void* dout, int64_t ld_dout, const void* out, int64_t ld_out, void* dcontent, int64_t ld_dcontent,
                      void* dropein, int64_t ld_drope, const float* inv_freq, float* dtheta_scratch /* calm_rope_bwd_scratch_floats */,
                      float* dinv_freq /* (dr/2) */, int64_t tokens, int32_t S, int32_t heads, int32_t dc, int32_t dr,
                      cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Axial attention with a learned additive bias shared by all heads:
 *   O = softmax(Q K^T / sqrt(hd) + bias[b]) V     replaces F.scaled_dot_product_attention (Vi_Tools_CNN_less_V2.py:293-298)
 * q/k/v/o: bf16, token-major (B*S, heads*hd) views with leading dims ld_* ; bias bf16 (B, S, S); lse f32 (B, heads, S).
 * bwd: dq/dk/dv bf16 (same layout), dbias bf16 (B,S,S) = sum_h dS_h, delta f32 (B, heads, S) scratch,
 *      ds_scratch: calm_attention_bwd_scratch_bytes(...) bytes, 16-byte aligned: the tcgen05 path stores every head's dS there
 *      (bf16, (B, heads, S, S)) and sums the heads with a second kernel (NULL forces the legacy kernels).
 * Two implementations behind these entry points, selected by the problem shape only: tcgen05/TMEM/TMA kernels
 * (attention_sm100.cu) when S <= 256, S % 16 == 0, head_dim <= 64, 16-byte aligned operands; warp-level mma.sync kernels
 * (attention.cu) otherwise (384^2 / 512^2 configs).
 * ------------------------------------------------------------------------------------------------------------------ */
int32_t calm_attention_fwd(const void* q, const void* k, const void* v, const void* bias, void* o, float* lse,
                           int64_t ld_q, int64_t ld_k, int64_t ld_v, int64_t ld_o, int32_t B, int32_t S, int32_t heads,
                           int32_t hd, cudaStream_t stream);
int64_t calm_attention_bwd_scratch_bytes(int32_t B, int32_t S, int32_t heads, int32_t hd);
int32_t calm_attention_bwd(const void* q, const void* k, const void* v, const void* bias, const void* o, const void* d_o,
                           const float* lse, float* delta, void* dq, void* dk, void* dv, void* dbias, void* ds_scratch, int64_t ld_q,
                           int64_t ld_k, int64_t ld_v, int64_t ld_o, int64_t ld_do, int64_t ld_dq, int64_t ld_dk,
                           int64_t ld_dv, int32_t B, int32_t S, int32_t heads, int32_t hd, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Latent bottleneck sampling (Vi_Tools_CNN_less_V2.py:232-244, 23-30): mv = [mu | rho] bf16 (rows, 2M)
 *   sigma = softplus(rho)+1e-6 ; z = mu + eps*sigma (eps = NULL in eval) ; zsum = zsum_prev + z ; KL partial sums
 * ------------------------------------------------------------------------------------------------------------------ */
int32_t calm_latent_fwd(const void* mv, const float* eps, const float* zsum_prev, float* zsum, void* zsum_bf16,
                        float* kl_partial /* (nblocks) */, int32_t nblocks, int64_t rows, int32_t Mh, cudaStream_t stream);
/* running KL total: kl_out = kl_prev (or 0) + scale * (sum part_q + sum part_kv), scale = -0.5 / (rows*M) (:24-26) */
int32_t calm_latent_kl(const float* part_q, const float* part_kv, int32_t nblocks, const float* kl_prev, float* kl_out,
                       float scale, cudaStream_t stream);
/* dz = dz_f32 (grad of the fp32 running sum, or NULL) + dz_bf16 (grad of its bf16 copy, or NULL);
 * dz_total (optional out, f32) receives that sum = the gradient flowing on to the previous block's running sum */
int32_t calm_latent_bwd(const void* mv, const float* eps, const float* dz, const void* dz_bf16, float kl_scale,
                        const float* dkl /* scalar on device or NULL */, void* dmv /* bf16 (rows, 2M) */,
                        float* dz_total, int64_t rows, int32_t Mh, cudaStream_t stream);
int32_t calm_latent_blocks(int64_t rows, int32_t Mh);

/* ------------------------------------------------------------------------------------------------------------------
 * Per-Block CNN residual on the (B, S, S, 3) token image (Vi_Tools_CNN_less_V2.py:378-385,400-403; CALM_ViT_V2.py:60-67,
 * 80-83): y = x + conv1x1(gelu(dwconv3x3(gelu(conv1x1(x))))) with 32 hidden channels, fused into one stencil kernel.
 * weights are the fp32 effective (already /sigma) tensors: w1 (32,3) b1 (32) w2 (32,9) b2 (32) w3 (3,32) b3 (3)
 * ------------------------------------------------------------------------------------------------------------------ */
int32_t calm_cnn_fwd(const float* x, float* y, const float* w1, const float* b1, const float* w2, const float* b2,
                     const float* w3, const float* b3, int32_t B, int32_t S, cudaStream_t stream);
/* grads: dx f32 (B,S,S,3); parameter-gradient partials gp (nblocks, CALM_CNN_NPARAM) then reduced into gparams */
#define CALM_CNN_NPARAM 547 /* 96 + 32 + 288 + 32 + 96 + 3 */
int32_t calm_cnn_bwd(const float* x, const float* dy, float* dx, void* dx_bf16 /* optional bf16 copy of dx, or NULL */,
                     const float* w1, const float* b1, const float* w2,
                     const float* b2, const float* w3, const float* b3, float* gpartial, int32_t nblocks, float* gparams,
                     int32_t B, int32_t S, cudaStream_t stream);
int32_t calm_cnn_bwd_blocks(int32_t B, int32_t S);

/* ------------------------------------------------------------------------------------------------------------------
 * Small memory-bound helpers
 * ------------------------------------------------------------------------------------------------------------------ */
/* row<->column re-tokenisation (Vi_Tools_CNN_less_V2.py:394-395,397-398): out[b,j,i,:] = in[b,i,j,:] (+ addend[b,j,i,:])
 * on (B,S,S,3) f32; the optional addend fuses the gradient accumulation of the un-transposed consumer in backward */
int32_t calm_token_transpose(const float* in, const float* addend, float* out, void* out_bf16 /* optional bf16 copy, or NULL */,
                             int32_t B, int32_t S, cudaStream_t stream);
/* first-block row tokenisation (:389-391): (B,3,S,S) NCHW f32 -> (B,S,S,3) ; and its inverse */
int32_t calm_nchw_to_tokens(const float* in, float* out, int32_t B, int32_t S, cudaStream_t stream);
/* column sums of a bf16 (rows, N) matrix -> f32 (N): linear_mask bias gradients */
int32_t calm_colsum(const void* x, int64_t ld, float* partial, int32_t nparts, float* out, int64_t rows, int32_t N, cudaStream_t stream);
int32_t calm_colsum_parts(int64_t rows, int32_t N);
/* out = a + b (+ c) elementwise f32 (U-Net skip adds, Vi_Tools_CNN_less_V2.py:513,516,520,522) */
int32_t calm_add3(const float* a, const float* b, const float* c, float* out, int64_t n, cudaStream_t stream);
/* f32 -> bf16 cast, and bf16 = bf16(a_f32 + b_f32) */
int32_t calm_cast_bf16(const float* in, void* out, int64_t n, cudaStream_t stream);
int32_t calm_cast_f32(const void* in_bf16, float* out, int64_t n, cudaStream_t stream);
/* mean over the sequence axis: x f32 (B,S,D) -> bf16/f32 (B,D) (CALM_ViT_V2.py:73-75), and its backward */
int32_t calm_seq_mean_fwd(const float* x, void* out_bf16, int32_t B, int32_t S, int32_t D, cudaStream_t stream);
int32_t calm_seq_mean_bwd(const void* dout_bf16, float* dx, int32_t B, int32_t S, int32_t D, cudaStream_t stream);
/* plain GELU(erf) forward on bf16 (head activation, CALM_ViT_V2.py:51) is covered by the GEMM epilogue. */

/* ------------------------------------------------------------------------------------------------------------------
 * Training-step glue of the per-rank loop: loss heads and the fused optimizer step.
 * ------------------------------------------------------------------------------------------------------------------ */
/* torch.nn.CrossEntropyLoss() on f32 logits (B,C) with probability targets (CutMix/MixUp soft labels,
 * distributed_trainer_cls.py:58-63,86) or, alternatively, int64 class labels (exactly one of target / labels non-NULL).
 * row_stats (B,4) f32 = {loss_b, logsumexp_b, sum_c t_bc, [argmax x_b == argmax t_b]} (saved for backward);
 * loss_out[0] = mean_b loss_b, loss_out[1] = dominant-class accuracy of the batch (:97-100, no host sync). */
int32_t calm_soft_ce_fwd(const float* logits, int64_t ld, const float* target, int64_t ld_t, const int64_t* labels,
                         float* row_stats, float* loss_out, int32_t B, int32_t C, cudaStream_t stream);
/* dlogits = (softmax(x) * sum_c t - t) * dloss / B ; dloss = device scalar (the scaled upstream gradient) or NULL (=1) */
int32_t calm_soft_ce_bwd(const float* logits, int64_t ld, const float* target, int64_t ld_t, const int64_t* labels,
                         const float* row_stats, const float* dloss, float* dlogits, int64_t ld_d, int32_t B, int32_t C,
                         cudaStream_t stream);
/* torch.nn.HuberLoss(delta)(img, x) + kl_weight * kl (distributed_trainer_reg.py:76-88) where img is the generated token
 * image (B,S,S,3) f32 read in place of its reshape/permute to NCHW and x the (B,3,S,S) f32 input image.
 * loss_out[0] = total, loss_out[1] = the Huber term; partial: calm_huber_parts(B,S) floats of scratch */
int32_t calm_huber_parts(int32_t B, int32_t S);
int32_t calm_huber_tokens_fwd(const float* tokens, const float* target_nchw, const float* kl /* device scalar or NULL */,
                              float kl_weight, float delta, float* partial, int32_t nparts, float* loss_out, int32_t B, int32_t S,
                              cudaStream_t stream);
/* dtokens = clamp(tokens - target, -delta, delta) * dloss / (3 B S^2) ; dkl[0] = kl_weight * dloss (dkl may be NULL) */
int32_t calm_huber_tokens_bwd(const float* tokens, const float* target_nchw, const float* dloss, float kl_weight, float delta,
                              float* dtokens, float* dkl, int32_t B, int32_t S, cudaStream_t stream);

/* Input side: torchvision.transforms.v2 MixUp / CutMix (distributed_trainer_cls.py:58-61, applied there by the
 * collate function on the CPU) on a batch that is already on the device. x, out f32 (B, channels, H, W), out != x.
 *   mode 0 (MixUp):  out[b] = lam * x[b] + (1 - lam) * x[b-1]                      (b-1 wraps: images.roll(1, 0))
 *   mode 1 (CutMix): out[b] = x[b] with rows [y1,y2) x columns [x1,x2) taken from x[b-1]
 *   soft (B, num_classes) f32 = lam_labels * onehot(labels[b]) + (1 - lam_labels) * onehot(labels[b-1])   (labels/soft may be NULL)
 * one_minus_* = the Python expression 1.0 - lam evaluated in double by the caller (torchvision's scalar), then rounded to f32.
 * The host draws lam / the box exactly like torchvision does (calm_trainer.MixBatch). */
int32_t calm_mix_batch(const float* x, const int64_t* labels, float* out, float* soft, int32_t B, int32_t channels, int32_t H,
                       int32_t W, int32_t num_classes, int32_t mode, float lam, float one_minus_lam, int32_t x1, int32_t y1, int32_t x2,
                       int32_t y2, float lam_labels, float one_minus_lam_labels, cudaStream_t stream);

/* GradScaler.unscale_ + clip_grad_norm_(max_norm) + GradScaler.step(AdamW) + GradScaler.update
 * (distributed_trainer_cls.py:88-96, distributed_trainer_reg.py:90-97) as three launches over all parameter tensors:
 * (1) per-chunk sum of squares of the unscaled gradients, (2) one CTA: total norm, inf/nan check, clip coefficient, step
 * count and bias corrections, loss-scale update, (3) AdamW on p / exp_avg / exp_avg_sq with the gradient multiplied by
 * clip/scale on the fly (torch's fused AdamW arithmetic order). A non-finite gradient skips the update and backs the
 * scale off, like GradScaler. The .grad tensors are read, not modified. All state lives in `state` on the device, so the
 * step needs no host synchronisation and can be captured in a CUDA graph (the host changes the learning rate by writing
 * state[CALM_OPT_LR]). */
#define CALM_OPT_CHUNK 8192   /* elements per CTA; chunk c covers elements [chunk_start[c]*CHUNK, +CHUNK) of tensor chunk_tensor[c] */
enum { CALM_OPT_SCALE = 0, CALM_OPT_GROWTH_TRACKER = 1, CALM_OPT_STEP = 2, CALM_OPT_LR = 3, CALM_OPT_GRAD_NORM = 4,
       CALM_OPT_FOUND_INF = 5, CALM_OPT_MULT = 6, CALM_OPT_BIAS1 = 7, CALM_OPT_BIAS2_SQRT = 8, CALM_OPT_STATE_FLOATS = 16 };
typedef struct {
  const void* params;        /* device array of n_tensors float*  (parameter data, updated in place)            */
  const void* grads;         /* device array of n_tensors const float* (their .grad, f32, contiguous)           */
  const void* grads_host;    /* optional pinned HOST copy of that array: when non-NULL it is uploaded into `grads` on
                                the stream first (autograd re-allocates .grad every step; inside a CUDA graph the copy
                                becomes a memcpy node that re-reads this buffer at every replay)                   */
  const int64_t* elem_off;   /* device (n_tensors + 1): prefix sum of numel = offsets into exp_avg / exp_avg_sq */
  const int32_t* chunk_tensor; /* device (n_chunks) */
  const int32_t* chunk_start;  /* device (n_chunks), in units of CALM_OPT_CHUNK elements                         */
  float* exp_avg;            /* flat f32 first moments  (elem_off[n_tensors] floats)                            */
  float* exp_avg_sq;         /* flat f32 second moments                                                         */
  float* partial;            /* scratch, n_chunks floats                                                        */
  float* state;              /* CALM_OPT_STATE_FLOATS floats, see the enum                                      */
  int32_t n_tensors, n_chunks;
  float beta1, beta2, eps, weight_decay;
  float max_norm;            /* <= 0: no clipping                                                               */
  float growth_factor, backoff_factor; int32_t growth_interval;
  int32_t use_scaler;        /* 0: gradients are not scaled, never skip                                         */
  int32_t _pad;
} calm_trainer_step_args;
int32_t calm_trainer_step(const calm_trainer_step_args* args, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CALM_B200_H_ */
