/*
 * tic_b200.h -- C-ABI of the B200-native ViT hot path for TouhouIC.
 *
 * The reference (fAKe2004/TouhouImageClassification) is pure Python and has no FFI of its own: every
 * kernel it runs is reached through PyTorch operators called by
 * transformers/models/vit/modeling_vit.py (the third-party module TIC/ViT/model.py:45 instantiates).
 * Each entry point below therefore cites the PyTorch/transformers call site it replaces
 * (SURVEY.md section 8a row numbers in brackets).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - no allocation inside: workspaces are caller-provided;
 *   - return 0 on success, non-zero on error; tic_last_error() returns the message (thread-local);
 *   - bf16 is the storage type __nv_bfloat16 (passed as void* / uint16_t*), row-major, 16-byte aligned.
 */
#ifndef TIC_B200_H_
#define TIC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TIC_API __attribute__((visibility("default")))
#else
#define TIC_API
#endif

/* ---- library ------------------------------------------------------------------------------- */
TIC_API int tic_abi_version(void);
TIC_API const char* tic_last_error(void);
/* Number of kernel launches this library has issued in this process (bench.py's gpu_launches). */
TIC_API int64_t tic_launch_count(void);
/* Optional per-launch timing: CUDA events on the launching stream around every kernel while enabled.
 * tic_prof_collect synchronises and writes "name\tlaunches\ttotal_ms\tflops\tbytes\n" lines (algorithmic
 * FLOPs / bytes per launch summed per kernel name) into a HOST buffer; returns the bytes written. */
TIC_API void tic_prof_enable(int on);
TIC_API int64_t tic_prof_collect(char* buf_host, int64_t buflen);

/* ---- GEMM (tcgen05 / TMEM / TMA) -------------------------------------------------------------
 * D[M,N] = A[M,K] * B[N,K]^T with bf16 operands and fp32 accumulation in tensor memory.
 * Replaces nn.Linear forward (modeling_vit.py:216-230 [a5], :262-268 [a7], :290-299 [a8],
 * :305-312 [a9]), the patch-embedding Conv2d as an im2col GEMM (:151-167 [a2]) and their autograd
 * backward GEMMs (dgrad / wgrad).
 *   a_mn_major = 0: A is [M,K] row-major (pitch lda);  1: A is stored [K,M] row-major (pitch lda).
 *   b_mn_major = 0: B is [N,K] row-major (pitch ldb);  1: B is stored [K,N] row-major (pitch ldb).
 * epilogue:
 *   0 TIC_EPI_BF16        out(bf16)  = acc + bias
 *   1 TIC_EPI_BF16_GELU   pre = bf16(acc + bias); out(bf16) = gelu_erf(pre); out2(bf16, optional) = gelu_erf'(pre),
 *                         the only thing the backward of this layer needs from the forward
 *   2 TIC_EPI_F32_RESID   out(f32)   = bf16(acc + bias) + aux(f32)[m,n]        (residual stream)
 *   3 TIC_EPI_BF16_DGELU  out(bf16)  = bf16(acc) * aux(bf16)[m,n], aux = out2 of TIC_EPI_BF16_GELU   (fc2 dgrad)
 *   4 TIC_EPI_F32         out(f32)   = acc + bias
 *   5 TIC_EPI_F32_ATOMIC  out(f32)  += acc   (split-K partial sums; `splits` > 1 allowed)
 *   6 TIC_EPI_F32_POSEMB  patch embedding: row m = img * P + p is written to out row
 *                         img * (P + 1) + 1 + p as bf16(acc + bias) + aux(f32)[1 + p, n]; aux_int = P
 * bias may be NULL. N must be a multiple of 8, pitches multiples of 8 elements.
 */
enum {
  TIC_EPI_BF16 = 0,
  TIC_EPI_BF16_GELU = 1,
  TIC_EPI_F32_RESID = 2,
  TIC_EPI_BF16_DGELU = 3,
  TIC_EPI_F32 = 4,
  TIC_EPI_F32_ATOMIC = 5,
  TIC_EPI_F32_POSEMB = 6,
  TIC_EPI_F32_GELU = 7 /* out(f32) = gelu_erf(acc + bias), exact erff (fp32 mode) */
};
TIC_API int tic_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major,
                          int M, int N, int K, int epilogue, void* out, int64_t ldo, void* out2, int64_t ldo2,
                          const float* bias, const void* aux, int64_t ldaux, int aux_int, int splits, void* stream);


/* Same GEMM with the TIC_EPI_BF16_DGELU epilogue that also ACCUMULATES the column sums of its bf16 output into
 * colsum_accum[N] (the fc1 bias gradient), saving a reduction pass over the [M, 4D] gradient. */
TIC_API int tic_gemm_bf16_colsum(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major,
                                 int M, int N, int K, int epilogue, void* out, int64_t ldo, const void* aux, int64_t ldaux,
                                 float* colsum_accum, void* stream);

/* ---- LayerNorm ---------------------------------------------------------------------------------
 * Replaces nn.LayerNorm (modeling_vit.py:325-326,333,340,455) [a4, a11]. fp32 statistics, eps from
 * ViTConfig.layer_norm_eps. Row pitches are in elements; y_bf16 / y_f32 / mean / rstd may be NULL.
 * Backward: dx = dres + LN'(dy) (dres may alias dx or be NULL); dgamma / dbeta are ACCUMULATED. dxsum (may be
 * NULL) ACCUMULATES the column sums of the bf16 dx it writes, i.e. the bias gradient of the Linear layer whose
 * output gradient that dx is (attention out-proj / fc2), so no separate reduction pass over dx is needed. */
TIC_API int tic_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, int rows,
                              int D, void* y_bf16, int64_t ldy, float* y_f32, int64_t ldyf, float* mean, float* rstd,
                              void* stream);
TIC_API int tic_layernorm_bwd(const void* dy_bf16, int64_t lddy, const float* x, int64_t ldx, const float* mean,
                              const float* rstd, const float* gamma, const float* dres, int64_t lddres, int rows, int D,
                              float* dx, int64_t lddx, void* dx_bf16, int64_t lddxb, float* dgamma, float* dbeta,
                              float* dxsum, void* stream);

/* ---- fused multi-head attention -----------------------------------------------------------------
 * Replaces F.scaled_dot_product_attention (modeling_vit.py:232-246; sdpa_attention.py:92-103) [a6].
 * q/k/v are token-major [B*N, ...] with row pitch ld (elements) and head h at column h*64; o is
 * [B*N, H*64] with pitch ldo. lse is [B, H, N] fp32 (NULL to skip). head_dim must be 64. */
TIC_API int tic_attention_fwd(const void* q, const void* k, const void* v, int64_t ld, void* o, int64_t ldo, float* lse,
                              int B, int N, int H, int head_dim, float scale, void* stream);
TIC_API int tic_attention_bwd(const void* q, const void* k, const void* v, int64_t ld, const void* o, int64_t ldo,
                              const void* dout, int64_t lddo, const float* lse, float* delta_scratch, void* dq, void* dk,
                              void* dv, int64_t lddqkv, int B, int N, int H, int head_dim, float scale, void* stream);
/* Number of floats delta_scratch must hold for a backward call of this shape: B*H*N for delta = rowsum(dO o O), plus,
 * for N > 256, the bf16 dQ partials (one per 128-key tile) that the fused long-sequence kernel writes and a reduction
 * pass sums. N <= 256 never touches the scratch (one kernel does everything). */
TIC_API int64_t tic_attention_bwd_scratch_floats(int B, int N, int H);

/* Same backward that also ACCUMULATES the column sums of dq | dk | dv into qkv_bias_grad (fp32 [3 * H * 64]): the bias
 * gradient of the fused QKV Linear (modeling_vit.py:216-230 [a5]). For N <= 256 the whole backward (delta, dQ, dK, dV,
 * bias gradient) is ONE kernel per call and delta_scratch is not touched. */
TIC_API int tic_attention_bwd_bias(const void* q, const void* k, const void* v, int64_t ld, const void* o, int64_t ldo,
                                   const void* dout, int64_t lddo, const float* lse, float* delta_scratch, void* dq,
                                   void* dk, void* dv, int64_t lddqkv, float* qkv_bias_grad, int B, int N, int H,
                                   int head_dim, float scale, void* stream);

/* Query-subset variants: only the first num_queries tokens of every image act as
 * queries -- their rows of o / dout are read, their rows of o / dq written, lse is [B, H, num_queries] -- while all N
 * tokens are keys. The engine uses num_queries = 1 in the LAST encoder layer: only the CLS row of that layer reaches
 * the final LayerNorm and the classifier (modeling_vit.py:455,641-642), so every other row of its attention output,
 * out-projection and MLP is dead work in the reference. qkv_bias_grad may be NULL. */
TIC_API int tic_attention_fwd_nq(const void* q, const void* k, const void* v, int64_t ld, void* o, int64_t ldo, float* lse,
                                 int B, int N, int num_queries, int H, int head_dim, float scale, void* stream);
TIC_API int tic_attention_bwd_nq(const void* q, const void* k, const void* v, int64_t ld, const void* o, int64_t ldo,
                                 const void* dout, int64_t lddo, const float* lse, float* delta_scratch, void* dq, void* dk,
                                 void* dv, int64_t lddqkv, float* qkv_bias_grad, int B, int N, int num_queries, int H,
                                 int head_dim, float scale, void* stream);

/* ---- classifier head and fused softmax cross-entropy ----------------------------------------------
 * tic_head_fwd: logits[B,C] = h[B,D] W[C,D]^T + b (classifier, modeling_vit.py:641-642 [a11]).
 * tic_softmax_xent: F.cross_entropy forward + backward in one launch, integer targets (finetune.py:61
 * [a13]) or soft targets from MixUp/CutMix (ntrain.py:48 [a15]); exactly one of hard/soft non-NULL.
 * loss = mean over the B rows; dlogits = (softmax * sum(y) - y) * grad_scale; correct = #argmax==hard. */
TIC_API int tic_head_fwd(const void* h_bf16, int64_t ldh, const void* w_bf16, const float* bias, int B, int D, int C,
                         int round_out_bf16, float* logits, void* stream);
TIC_API int tic_head_bwd(const float* dlogits, const void* h_bf16, int64_t ldh, const void* w_bf16, int B, int D, int C,
                         void* dh_bf16, int64_t lddh, float* dW_accum, float* db_accum, void* stream);
TIC_API int tic_softmax_xent(const float* logits, const int64_t* hard, const float* soft, int B, int C,
                             float grad_scale, int round_grad_bf16, float* loss, float* dlogits, int32_t* correct,
                             void* stream);

/* Post-processing of the serving path (TIC/utils/serve.py:107-109, web/runtime.py:117-118: torch.softmax(logits, 1) then
 * torch.max(probabilities, 1)) in one launch: confidence[b] = max_c softmax(logits[b])[c], index[b] = its class (first
 * index on ties); probs (optional, may be NULL) receives the whole fp32 distribution [B, C]. */
TIC_API int tic_softmax_top1(const float* logits, int B, int C, float* confidence, int32_t* index, float* probs,
                             void* stream);

/* ---- fused AdamW --------------------------------------------------------------------------------
 * Replaces torch.optim.AdamW.step as configured at ntrain.py:39-41 [a16] / finetune.py:314 over a flat
 * fp32 arena (n % 4 == 0); also writes the bf16 shadow (may be NULL). `step` is 1-based. Gradients are
 * multiplied by grad_scale first (1/world_size after a sum-allreduce, or 1). */
TIC_API int tic_adamw_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, int64_t n, float lr,
                           float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                           void* stream);

/* ---- elementwise glue ---------------------------------------------------------------------------- */
/* fp32 NCHW [B,3,S,S] -> bf16 patch rows [B*(S/16)^2, 768], K ordered (c, py, px) (modeling_vit.py:151,166 [a2]) */
TIC_API int tic_patchify_f32(const float* pixels, void* patches_bf16, int B, int S, void* stream);
/* CutMix / MixUp of a device batch (torchvision v2 RandomChoice([CutMix, MixUp]), ntrain.py:30-33,45-46 [a19]) fused
 * with the patchify: sample b is paired with sample b-1 (roll(1,0)). mode 0 = none, 1 = MixUp with weight lam,
 * 2 = CutMix pasting the box [y1,y2) x [x1,x2) from the rolled batch. mixed_out (fp32 [B,3,S,S], optional, must not
 * alias pixels) and patches_bf16 ([B*(S/16)^2, 768], optional) receive the result; the fp32 arithmetic is bit-exact
 * with the three torch ops of the reference. tic_mix_targets writes the soft labels
 * onehot(y[b-1]) * one_minus_lam + onehot(y[b]) * lam (lam = the box-adjusted lambda for CutMix). lam and
 * one_minus_lam are passed separately so the host can round each from its own double, as torch does. */
TIC_API int tic_mix_patchify_f32(const float* pixels, float* mixed_out, void* patches_bf16, int B, int S, int mode,
                                 float lam, float one_minus_lam, int x1, int y1, int x2, int y2, void* stream);
TIC_API int tic_mix_targets(const int64_t* labels, int B, int C, float lam, float one_minus_lam, float* soft_out,
                            void* stream);
TIC_API int tic_cast_f32_to_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);
TIC_API int tic_cast_bf16_to_f32(const void* src_bf16, float* dst, int64_t n, void* stream);
/* out[n] += sum_m dy[m, n] (bias gradients) */
TIC_API int tic_colsum_bf16(const void* dy_bf16, int64_t ld, int rows, int cols, float* out_accum, void* stream);

/* ---- fused augment + normalize + patchify ------------------------------------------------------------
 * Replaces the per-sample CPU torchvision pipeline of AugmentedDataset.setup (ntrain.py:104-112 [a18]) and the
 * im2col of the patch projection: RandomResizedCrop(size) -> HFlip -> ColorJitter(.2,.2,.2,.1, random order) ->
 * RandomGrayscale(.2) -> RandomErasing(.5, value 0) -> ToTensor -> Normalize(mean, std) -> bf16 patch rows.
 * tic_augment_sample_params (HOST code) fills, per sample, ints[16] = {top, left, h, w, flip, op0..op3, gray,
 * erase_i, erase_j, erase_h, erase_w, jitter_on, 0} and floats[4] = {brightness, contrast, saturation, hue} from a
 * counter-based RNG keyed by (seed, first_sample + b); recipe 0 = full, 1 = crop + flip + erase only
 * (ntrain.py:127-134). tic_augment_patchify: images uint8 NHWC [B,H,W,3] (device), ints/floats on the DEVICE,
 * patches bf16 [B*(size/16)^2, 768] with K ordered (c, py, px); pixels_out (optional, may be NULL) receives the
 * uint8 NHWC image after erasing. size must be a multiple of 16 and <= 224. Results are bit-exact against
 * oracle/augment_oracle.py. */
TIC_API int tic_augment_sample_params(int64_t seed, int64_t first_sample, int B, int H, int W, int size, int recipe,
                                      int32_t* ints_host, float* floats_host);
TIC_API int tic_augment_patchify(const void* images_u8, int B, int H, int W, const int32_t* ints, const float* floats,
                                 int size, const float* mean3_host, const float* std3_host, void* patches_bf16,
                                 void* pixels_out_u8, void* stream);
/* Same pipeline, same parameters, but yields what the reference's Dataset yields: the normalised fp32 tensor
 * [B, 3, size, size] (ntrain.py:104-112 ends in ToTensor + Normalize), i.e. the input of the per-batch CutMix / MixUp
 * (ntrain.py:45-46 -> tic_mix_patchify_f32). patches_bf16 is optional here (may be NULL). */
TIC_API int tic_augment_tensor(const void* images_u8, int B, int H, int W, const int32_t* ints, const float* floats,
                               int size, const float* mean3_host, const float* std3_host, float* tensor_out_f32,
                               void* patches_bf16, void* stream);

/* ---- whole-model engine ---------------------------------------------------------------------------
 * Stands where ViTForImageClassification.forward (modeling_vit.py:620-653) and its autograd backward
 * stand in the reference [a2-a12]. Parameters live in one fp32 arena whose element offsets are reported
 * by tic_vit_param_layout in HF named_parameters() order:
 *   cls_token, position_embeddings, patch.weight, patch.bias,
 *   per layer: q.w q.b k.w k.b v.w v.b attn_out.w attn_out.b fc1.w fc1.b fc2.w fc2.b ln_before.w ln_before.b
 *              ln_after.w ln_after.b,
 *   layernorm.w, layernorm.b, classifier.w, classifier.b               => 4 + 16*layers + 4 tensors.
 * The bf16 shadow and the gradient arena use the same offsets. */
typedef struct tic_vit_config {
  int32_t image_size;  /* 224 or 384 */
  int32_t patch_size;  /* 16 */
  int32_t hidden;      /* 768 / 1024 */
  int32_t layers;      /* 12 / 24 */
  int32_t heads;       /* 12 / 16 (head_dim 64) */
  int32_t mlp;         /* 3072 / 4096 */
  int32_t num_labels;  /* 120 */
  float ln_eps;        /* 1e-12 */
} tic_vit_config;

TIC_API int64_t tic_vit_param_arena_elems(const tic_vit_config* cfg);
/* Fills offsets[i], numels[i] for tensor i in HF order; returns the number of tensors, or -1 on error. */
TIC_API int tic_vit_param_layout(const tic_vit_config* cfg, int64_t* offsets, int64_t* numels, int max_tensors);
/* Element offset where the classifier (the only part trained when full_finetune=False, ntrain.py:35-37) begins. */
TIC_API int64_t tic_vit_head_offset(const tic_vit_config* cfg);
/* Element range [begin, end) of the gradient arena that backward stage `stage` completes (for bucketed allreduce). */
TIC_API int tic_vit_stage_grad_range(const tic_vit_config* cfg, int stage, int64_t* begin, int64_t* end);
TIC_API int64_t tic_vit_workspace_bytes(const tic_vit_config* cfg, int batch, int training);
/* Exactly one of pixels (fp32 [B,3,S,S]) / patches_bf16 ([B*P,768], e.g. from tic_augment_patchify) is non-NULL.
 * training != 0 keeps the activations the backward needs in `workspace`. logits: fp32 [B, num_labels]. */
TIC_API int tic_vit_forward(const tic_vit_config* cfg, const float* params_f32, const void* params_bf16,
                            const float* pixels, const void* patches_bf16, int batch, void* workspace,
                            int64_t workspace_bytes, int training, float* logits, void* stream);
/* Backward stages: 0 = classifier + final LayerNorm, 1..layers = encoder layers (last first), layers+1 = embeddings.
 * Runs stages [stage_begin, stage_end) and ACCUMULATES parameter gradients into grads_f32 (arena layout).
 * head_only != 0 computes only the classifier gradients (frozen backbone). */
TIC_API int tic_vit_backward(const tic_vit_config* cfg, const float* params_f32, const void* params_bf16, int batch,
                             void* workspace, int64_t workspace_bytes, const float* dlogits, float* grads_f32,
                             int stage_begin, int stage_end, int head_only, void* stream);

/* ---- fp32-accurate inference ("fp32 mode") ------------------------------------------------------------
 * The reference serves with no autocast, i.e. in plain fp32 (TIC/utils/serve.py:99-101 [a21], web/runtime.py:115-116
 * [a22]). Tensor cores have no fp32 operand type, so every Linear runs as a split-bf16 GEMM on the same tcgen05
 * kernel: x = hi + mid + lo (three bf16 terms = 24 mantissa bits), likewise w, and the six cross terms larger than
 * 2^-24 are one GEMM over a 6x longer reduction dimension with fp32 accumulation in tensor memory. LayerNorm, exact-erf
 * GELU, residual adds, softmax and the attention products are fp32 on the CUDA cores. Forward only.
 *   tic_vit_w6_elems            bf16 elements of the split weight buffer (6x the GEMM weights)
 *   tic_vit_prepare_w6          split the GEMM weights of the fp32 arena into it (once per weight update)
 *   tic_vit_workspace_bytes_f32 workspace for a batch
 *   tic_vit_forward_f32         pixels fp32 [B,3,S,S] -> logits fp32 [B, num_labels] */
TIC_API int64_t tic_vit_w6_elems(const tic_vit_config* cfg);
TIC_API int64_t tic_vit_workspace_bytes_f32(const tic_vit_config* cfg, int batch);
TIC_API int tic_vit_prepare_w6(const tic_vit_config* cfg, const float* params_f32, void* w6_bf16, void* stream);
TIC_API int tic_vit_forward_f32(const tic_vit_config* cfg, const float* params_f32, const void* w6_bf16,
                                const float* pixels, int batch, void* workspace, int64_t workspace_bytes, float* logits,
                                void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TIC_B200_H_ */
