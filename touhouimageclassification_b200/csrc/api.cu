// extern "C" surface of libtic_b200.so (declared in include/tic_b200.h). Thin argument adapters only.
#include "tic_b200.h"
#include "tic_internal.cuh"

using namespace tic;

namespace {
inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }
}  // namespace

extern "C" {

TIC_API int tic_abi_version(void) { return 1; }
TIC_API const char* tic_last_error(void) { return last_error(); }
TIC_API int64_t tic_launch_count(void) { return launch_count(); }
TIC_API void tic_prof_enable(int on) { prof_enable(on != 0); }
TIC_API int64_t tic_prof_collect(char* buf, int64_t buflen) { return prof_collect(buf, buflen); }

TIC_API int tic_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major,
                          int M, int N, int K, int epilogue, void* out, int64_t ldo, void* out2, int64_t ldo2,
                          const float* bias, const void* aux, int64_t ldaux, int aux_int, int splits, void* stream) {
  return gemm_bf16(A, lda, a_mn_major != 0, B, ldb, b_mn_major != 0, M, N, K, epilogue, out, ldo, out2, ldo2, bias,
                   aux, ldaux, aux_int, splits, S(stream));
}

TIC_API int tic_gemm_bf16_colsum(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major,
                                 int M, int N, int K, int epilogue, void* out, int64_t ldo, const void* aux, int64_t ldaux,
                                 float* colsum_accum, void* stream) {
  return gemm_bf16(A, lda, a_mn_major != 0, B, ldb, b_mn_major != 0, M, N, K, epilogue, out, ldo, nullptr, 0, nullptr, aux,
                   ldaux, 0, 1, S(stream), colsum_accum);
}

TIC_API int tic_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, int rows,
                              int D, void* y_bf16, int64_t ldy, float* y_f32, int64_t ldyf, float* mean, float* rstd,
                              void* stream) {
  return layernorm_fwd(x, ldx, gamma, beta, eps, rows, D, y_bf16, ldy, y_f32, ldyf, mean, rstd, S(stream));
}
TIC_API int tic_layernorm_bwd(const void* dy_bf16, int64_t lddy, const float* x, int64_t ldx, const float* mean,
                              const float* rstd, const float* gamma, const float* dres, int64_t lddres, int rows, int D,
                              float* dx, int64_t lddx, void* dx_bf16, int64_t lddxb, float* dgamma, float* dbeta,
                              float* dxsum, void* stream) {
  return layernorm_bwd(dy_bf16, lddy, x, ldx, mean, rstd, gamma, dres, lddres, rows, D, dx, lddx, dx_bf16, lddxb,
                       dgamma, dbeta, dxsum, S(stream));
}

TIC_API int tic_attention_fwd(const void* q, const void* k, const void* v, int64_t ld, void* o, int64_t ldo, float* lse,
                              int B, int N, int H, int head_dim, float scale, void* stream) {
  return attention_fwd_tc(q, k, v, ld, o, ldo, lse, B, N, H, head_dim, scale, S(stream));
}
TIC_API int tic_attention_bwd(const void* q, const void* k, const void* v, int64_t ld, const void* o, int64_t ldo,
                              const void* dout, int64_t lddo, const float* lse, float* delta_scratch, void* dq, void* dk,
                              void* dv, int64_t lddqkv, int B, int N, int H, int head_dim, float scale, void* stream) {
  return attention_bwd_tc(q, k, v, ld, o, ldo, dout, lddo, lse, delta_scratch, dq, dk, dv, lddqkv, B, N, H, head_dim, scale,
                       S(stream));
}

TIC_API int64_t tic_attention_bwd_scratch_floats(int B, int N, int H) { return attention_bwd_scratch_floats(B, N, H); }

TIC_API int tic_attention_bwd_bias(const void* q, const void* k, const void* v, int64_t ld, const void* o, int64_t ldo,
                                   const void* dout, int64_t lddo, const float* lse, float* delta_scratch, void* dq,
                                   void* dk, void* dv, int64_t lddqkv, float* qkv_bias_grad, int B, int N, int H,
                                   int head_dim, float scale, void* stream) {
  return attention_bwd_tc(q, k, v, ld, o, ldo, dout, lddo, lse, delta_scratch, dq, dk, dv, lddqkv, B, N, H, head_dim, scale,
                          S(stream), qkv_bias_grad);
}

TIC_API int tic_attention_fwd_nq(const void* q, const void* k, const void* v, int64_t ld, void* o, int64_t ldo, float* lse,
                                 int B, int N, int num_queries, int H, int head_dim, float scale, void* stream) {
  return attention_fwd_tc(q, k, v, ld, o, ldo, lse, B, N, H, head_dim, scale, S(stream), num_queries);
}
TIC_API int tic_attention_bwd_nq(const void* q, const void* k, const void* v, int64_t ld, const void* o, int64_t ldo,
                                 const void* dout, int64_t lddo, const float* lse, float* delta_scratch, void* dq, void* dk,
                                 void* dv, int64_t lddqkv, float* qkv_bias_grad, int B, int N, int num_queries, int H,
                                 int head_dim, float scale, void* stream) {
  return attention_bwd_tc(q, k, v, ld, o, ldo, dout, lddo, lse, delta_scratch, dq, dk, dv, lddqkv, B, N, H, head_dim, scale,
                          S(stream), qkv_bias_grad, 7, num_queries);
}

TIC_API int tic_head_fwd(const void* h_bf16, int64_t ldh, const void* w_bf16, const float* bias, int B, int D, int C,
                         int round_out_bf16, float* logits, void* stream) {
  return head_fwd(h_bf16, ldh, w_bf16, bias, B, D, C, round_out_bf16, logits, S(stream));
}
TIC_API int tic_head_bwd(const float* dlogits, const void* h_bf16, int64_t ldh, const void* w_bf16, int B, int D, int C,
                         void* dh_bf16, int64_t lddh, float* dW_accum, float* db_accum, void* stream) {
  return head_bwd(dlogits, h_bf16, ldh, w_bf16, B, D, C, dh_bf16, lddh, dW_accum, db_accum, S(stream));
}
TIC_API int tic_softmax_xent(const float* logits, const int64_t* hard, const float* soft, int B, int C,
                             float grad_scale, int round_grad_bf16, float* loss, float* dlogits, int32_t* correct,
                             void* stream) {
  return softmax_xent(logits, reinterpret_cast<const long long*>(hard), soft, B, C, grad_scale, round_grad_bf16, loss,
                      dlogits, correct, S(stream));
}

TIC_API int tic_softmax_top1(const float* logits, int B, int C, float* confidence, int32_t* index, float* probs, void* stream) {
  return softmax_top1(logits, B, C, confidence, index, probs, S(stream));
}

TIC_API int tic_adamw_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, int64_t n, float lr,
                           float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                           void* stream) {
  return adamw_step(p, g, m, v, shadow_bf16, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, S(stream));
}

TIC_API int tic_patchify_f32(const float* pixels, void* patches_bf16, int B, int Sz, void* stream) {
  return patchify_f32(pixels, patches_bf16, B, Sz, S(stream));
}
TIC_API int tic_mix_patchify_f32(const float* pixels, float* mixed_out, void* patches_bf16, int B, int Sz, int mode,
                                 float lam, float one_minus_lam, int x1, int y1, int x2, int y2, void* stream) {
  return mix_patchify_f32(pixels, mixed_out, patches_bf16, B, Sz, mode, lam, one_minus_lam, x1, y1, x2, y2, S(stream));
}
TIC_API int tic_mix_targets(const int64_t* labels, int B, int C, float lam, float one_minus_lam, float* soft_out,
                            void* stream) {
  return mix_targets(reinterpret_cast<const long long*>(labels), B, C, lam, one_minus_lam, soft_out, S(stream));
}
TIC_API int tic_cast_f32_to_bf16(const float* src, void* dst_bf16, int64_t n, void* stream) {
  return cast_f32_to_bf16(src, dst_bf16, n, S(stream));
}
TIC_API int tic_cast_bf16_to_f32(const void* src_bf16, float* dst, int64_t n, void* stream) {
  return cast_bf16_to_f32(src_bf16, dst, n, S(stream));
}
TIC_API int tic_colsum_bf16(const void* dy_bf16, int64_t ld, int rows, int cols, float* out_accum, void* stream) {
  return colsum_bf16(dy_bf16, ld, rows, cols, out_accum, S(stream));
}

TIC_API int tic_augment_sample_params(int64_t seed, int64_t first_sample, int B, int H, int W, int size, int recipe,
                                      int32_t* ints_host, float* floats_host) {
  return augment_sample_params(seed, first_sample, B, H, W, size, recipe, ints_host, floats_host);
}
TIC_API int tic_augment_patchify(const void* images_u8, int B, int H, int W, const int32_t* ints, const float* floats,
                                 int size, const float* mean3_host, const float* std3_host, void* patches_bf16,
                                 void* pixels_out_u8, void* stream) {
  return augment_patchify(images_u8, B, H, W, ints, floats, size, mean3_host, std3_host, patches_bf16, pixels_out_u8,
                          nullptr, S(stream));
}
TIC_API int tic_augment_tensor(const void* images_u8, int B, int H, int W, const int32_t* ints, const float* floats,
                               int size, const float* mean3_host, const float* std3_host, float* tensor_out_f32,
                               void* patches_bf16, void* stream) {
  if (tensor_out_f32 == nullptr) return set_error(kErrInvalidArg, "tic_augment_tensor: tensor_out_f32 is NULL");
  return augment_patchify(images_u8, B, H, W, ints, floats, size, mean3_host, std3_host, patches_bf16, nullptr,
                          tensor_out_f32, S(stream));
}

TIC_API int64_t tic_vit_param_arena_elems(const tic_vit_config* cfg) {
  if (vit_validate(cfg) != kOk) return -1;
  return vit_layout(cfg).total;
}

TIC_API int tic_vit_param_layout(const tic_vit_config* cfg, int64_t* offsets, int64_t* numels, int max_tensors) {
  if (vit_validate(cfg) != kOk) return -1;
  const VitLayout L = vit_layout(cfg);
  const int64_t D = cfg->hidden, F = cfg->mlp, C = cfg->num_labels;
  const int64_t G = cfg->image_size / 16, N = G * G + 1;
  const int count = 4 + 16 * cfg->layers + 4;
  if (max_tensors < count) {
    set_error(kErrInvalidArg, "tic_vit_param_layout: need room for %d tensors", count);
    return -1;
  }
  int i = 0;
  auto put = [&](int64_t off, int64_t n) { offsets[i] = off; numels[i] = n; ++i; };
  put(L.cls, D); put(L.pos, N * D); put(L.patch_w, D * 768); put(L.patch_b, D);
  for (int l = 0; l < cfg->layers; ++l) {
    const int64_t b = L.layer0 + l * L.layer_stride;
    put(b + L.qkv_w, D * D);             put(b + L.qkv_b, D);          // query
    put(b + L.qkv_w + D * D, D * D);     put(b + L.qkv_b + D, D);      // key
    put(b + L.qkv_w + 2 * D * D, D * D); put(b + L.qkv_b + 2 * D, D);  // value
    put(b + L.o_w, D * D);   put(b + L.o_b, D);
    put(b + L.fc1_w, F * D); put(b + L.fc1_b, F);
    put(b + L.fc2_w, D * F); put(b + L.fc2_b, D);
    put(b + L.ln1_w, D); put(b + L.ln1_b, D);
    put(b + L.ln2_w, D); put(b + L.ln2_b, D);
  }
  put(L.lnf_w, D); put(L.lnf_b, D); put(L.cls_w, C * D); put(L.cls_b, C);
  return i;
}

TIC_API int64_t tic_vit_head_offset(const tic_vit_config* cfg) {
  if (vit_validate(cfg) != kOk) return -1;
  return vit_layout(cfg).head_begin;
}

TIC_API int tic_vit_stage_grad_range(const tic_vit_config* cfg, int stage, int64_t* begin, int64_t* end) {
  int rc = vit_validate(cfg);
  if (rc != kOk) return rc;
  const VitLayout L = vit_layout(cfg);
  if (stage == 0) { *begin = L.lnf_w; *end = L.total; }
  else if (stage <= cfg->layers) {
    const int l = cfg->layers - stage;
    *begin = L.layer0 + l * L.layer_stride;
    *end = *begin + L.layer_stride;
  } else if (stage == cfg->layers + 1) { *begin = 0; *end = L.layer0; }
  else return set_error(kErrInvalidArg, "tic_vit_stage_grad_range: stage %d out of range", stage);
  return kOk;
}

TIC_API int64_t tic_vit_workspace_bytes(const tic_vit_config* cfg, int batch, int training) {
  if (vit_validate(cfg) != kOk || batch <= 0) return -1;
  return vit_workspace_bytes(cfg, batch, training);
}

TIC_API int tic_vit_forward(const tic_vit_config* cfg, const float* params_f32, const void* params_bf16,
                            const float* pixels, const void* patches_bf16, int batch, void* workspace,
                            int64_t workspace_bytes, int training, float* logits, void* stream) {
  return vit_forward(cfg, params_f32, params_bf16, pixels, patches_bf16, batch, workspace, workspace_bytes, training,
                     logits, S(stream));
}

TIC_API int tic_vit_backward(const tic_vit_config* cfg, const float* params_f32, const void* params_bf16, int batch,
                             void* workspace, int64_t workspace_bytes, const float* dlogits, float* grads_f32,
                             int stage_begin, int stage_end, int head_only, void* stream) {
  return vit_backward(cfg, params_f32, params_bf16, batch, workspace, workspace_bytes, dlogits, grads_f32, stage_begin,
                      stage_end, head_only, S(stream));
}

TIC_API int64_t tic_vit_w6_elems(const tic_vit_config* cfg) {
  if (vit_validate(cfg) != kOk) return -1;
  return vit_w6_elems(cfg);
}
TIC_API int64_t tic_vit_workspace_bytes_f32(const tic_vit_config* cfg, int batch) {
  if (vit_validate(cfg) != kOk || batch <= 0) return -1;
  return vit_workspace_bytes_f32(cfg, batch);
}
TIC_API int tic_vit_prepare_w6(const tic_vit_config* cfg, const float* params_f32, void* w6_bf16, void* stream) {
  return vit_prepare_w6(cfg, params_f32, w6_bf16, S(stream));
}
TIC_API int tic_vit_forward_f32(const tic_vit_config* cfg, const float* params_f32, const void* w6_bf16,
                                const float* pixels, int batch, void* workspace, int64_t workspace_bytes, float* logits,
                                void* stream) {
  return vit_forward_f32(cfg, params_f32, w6_bf16, pixels, batch, workspace, workspace_bytes, logits, S(stream));
}

}  // extern "C"
