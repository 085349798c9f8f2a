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

/* ---- GEMM (tcgen05 / TMEM / TMA) -------------------------------------------------------------
 * D[M,N] = A[M,K] * B[N,K]^T with bf16 operands and fp32 accumulation in tensor memory.
 * Replaces nn.Linear forward (modeling_vit.py:216-230 [a5], :262-268 [a7], :290-299 [a8],
 * :305-312 [a9]), the patch-embedding Conv2d as an im2col GEMM (:151-167 [a2]) and their autograd
 * backward GEMMs (dgrad / wgrad).
 *   a_mn_major = 0: A is [M,K] row-major (pitch lda);  1: A is stored [K,M] row-major (pitch lda).
 *   b_mn_major = 0: B is [N,K] row-major (pitch ldb);  1: B is stored [K,N] row-major (pitch ldb).
 * epilogue:
 *   0 TIC_EPI_BF16        out(bf16)  = acc + bias
 *   1 TIC_EPI_BF16_GELU   out2(bf16) = pre = bf16(acc + bias); out(bf16) = gelu_erf(pre)
 *   2 TIC_EPI_F32_RESID   out(f32)   = bf16(acc + bias) + aux(f32)[m,n]        (residual stream)
 *   3 TIC_EPI_BF16_DGELU  out(bf16)  = bf16(acc) * gelu_erf'(aux(bf16)[m,n])   (fc2 dgrad)
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
  TIC_EPI_F32_POSEMB = 6
};
TIC_API int tic_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major,
                          int M, int N, int K, int epilogue, void* out, int64_t ldo, void* out2, int64_t ldo2,
                          const float* bias, const void* aux, int64_t ldaux, int aux_int, int splits, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TIC_B200_H_ */
