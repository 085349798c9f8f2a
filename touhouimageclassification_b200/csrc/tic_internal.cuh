// Internal launcher prototypes shared between the kernel translation units, api.cu and engine.cu.
#pragma once
#include "tic_common.cuh"

struct tic_vit_config;

namespace tic {

const char* last_error();
long long launch_count();
void prof_enable(bool on);
long long prof_collect(char* buf, long long buflen);
// Per-device launch state (runtime.cu): several devices may be driven from one process.
int current_device();
int device_sm_count();                                                     // SM count of the current device (cached)
int ensure_dynamic_smem(const void* func, int bytes, const char* what);   // once per (device, kernel)
void tmap_cache_stats(long long* hits, long long* misses);
// RAII: when profiling is enabled, brackets one kernel launch with CUDA events on its stream.
struct ProfScope {
  ProfScope(const char* name, double flops, double bytes, cudaStream_t s);
  ~ProfScope();
  int idx_;
  cudaStream_t stream_;
};

// Programmatic dependent launch. The kernels that make up a forward / backward pass run back to back on one stream;
// launched this way, kernel k+1 may become resident and run its prologue (barrier init, TMEM allocation, descriptor
// prefetch, parameter loads) while kernel k drains, instead of paying launch latency + prologue after it -- the
// difference is a few microseconds per launch, i.e. most of a small-batch inference forward. Contract: a kernel
// launched through launch_pdl executes pdl_launch_dependents() first and pdl_wait() in EVERY CTA before it reads or
// writes global memory (pdl_wait returns once the preceding kernel has completed and its writes are visible; both are
// no-ops in a normal launch). Measured on B200 (ViT-L/16, same box): it takes 4 % off a batch-1 / batch-8 forward, where
// launch gaps are a large share, and ADDS 1-1.5 % to a batch-1024 forward or a batch-256 training step (next-kernel CTAs
// that are resident but blocked take registers and warp slots from the kernel still running). So it is opt-in per call
// sequence: launches inside a PdlScope(true) -- the engine opens one for small inference forwards -- carry the
// attribute, all others are plain. TIC_NO_PDL=1 in the environment disables it everywhere (development A/B).
bool pdl_enabled();
struct PdlScope {   // thread-local nesting counter: launches made by this thread while a scope(true) is alive use PDL
  explicit PdlScope(bool on);
  ~PdlScope();
  bool on_;
};
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// GEMM epilogues (values are part of the C-ABI: TIC_EPI_* in include/tic_b200.h)
enum Epilogue : int {
  kEpiBf16 = 0,         // out bf16 = acc (+bias)
  kEpiBf16Gelu = 1,     // pre = bf16(acc + bias); out bf16 = gelu(pre); out2 bf16 (optional) = gelu'(pre), kept for backward
  kEpiF32Resid = 2,     // out f32 = bf16(acc + bias) + aux_f32[m,n]
  kEpiBf16DGelu = 3,    // out bf16 = bf16(acc) * aux_bf16[m,n], aux = gelu'(pre) saved by kEpiBf16Gelu
  kEpiF32 = 4,          // out f32 = acc (+bias)
  kEpiF32Atomic = 5,    // out f32 += acc (split-K partial, red.global.add)
  kEpiF32PosEmbed = 6,  // patch embedding rows: see tic_b200.h
  kEpiF32Gelu = 7,      // out f32 = gelu_erf(acc + bias), exact erff (fp32 verification mode)
};

// gemm_tcgen05.cu
int gemm_bf16(const void* A, long long lda, bool a_mn, const void* B, long long ldb, bool b_mn, int M, int N, int K,
              int epilogue, void* out, long long ldo, void* out2, long long ldo2, const float* bias, const void* aux,
              long long ldaux, int aux_int, int splits, cudaStream_t stream, float* colsum = nullptr, bool exact = false);

// layernorm.cu
int layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, float eps, int rows, int D,
                  void* y_bf16, long long ldy, float* y_f32, long long ldyf, float* mean, float* rstd,
                  cudaStream_t stream);
int layernorm_bwd(const void* dy_bf16, long long lddy, const float* x, long long ldx, const float* mean,
                  const float* rstd, const float* gamma, const float* dres, long long lddres, int rows, int D,
                  float* dx, long long lddx, void* dx_bf16, long long lddxb, float* dgamma, float* dbeta,
                  float* dxsum, cudaStream_t stream);

// elementwise.cu
int patchify_f32(const float* x, void* out_bf16, int B, int S, cudaStream_t stream);
int mix_patchify_f32(const float* x, float* mixed, void* out_bf16, int B, int S, int mode, float lam, float one_minus_lam,
                     int x1, int y1, int x2, int y2, cudaStream_t stream);
int mix_targets(const long long* y, int B, int C, float lam, float one_minus_lam, float* soft, cudaStream_t stream);
int cls_rows(const float* cls, const float* pos, float* x, int B, int N, int D, cudaStream_t stream);
int embed_bwd(const float* dx, int B, int N, int D, float* dpos, float* dcls, void* dpatch_bf16, cudaStream_t stream);
int cast_f32_to_bf16(const float* src, void* dst, long long n, cudaStream_t stream);
int cast_bf16_to_f32(const void* src, float* dst, long long n, cudaStream_t stream);
int colsum_bf16(const void* dy, long long ld, int rows, int cols, float* out, cudaStream_t stream);

// xent.cu
int head_fwd(const void* h_bf16, long long ldh, const void* w_bf16, const float* bias, int B, int D, int C,
             int round_out, float* logits, cudaStream_t stream);
int head_bwd(const float* dlogits, const void* h_bf16, long long ldh, const void* w_bf16, int B, int D, int C,
             void* dh_bf16, long long lddh, float* dW, float* db, cudaStream_t stream);
int softmax_xent(const float* logits, const long long* hard, const float* soft, int B, int C, float grad_scale,
                 int round_grad, float* loss, float* dlogits, int* correct, cudaStream_t stream);

int softmax_top1(const float* logits, int B, int C, float* conf, int* idx, float* probs, cudaStream_t stream);

// adamw.cu
int adamw_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, long long n, float lr, float beta1,
               float beta2, float eps, float weight_decay, int step, float grad_scale, cudaStream_t stream);

// augment.cu
int augment_sample_params(long long seed, long long first_sample, int B, int H, int W, int size, int recipe,
                          int* ints_host, float* floats_host);
int augment_patchify(const void* images_u8, int B, int H, int W, const int* ints_dev, const float* floats_dev, int size,
                     const float* mean3_host, const float* std3_host, void* patches_bf16, void* pixels_out_u8,
                     float* tensor_out_f32, cudaStream_t stream);

// attention_tc.cu (tcgen05 / TMEM / TMA)
// Nq (0 = N): only the first Nq tokens of every image act as queries (N <= 224 / 256 paths only); lse is [B, H, Nq].
int attention_fwd_tc(const void* q, const void* k, const void* v, long long ld, void* o, long long ldo, float* lse,
                     int B, int N, int H, int head_dim, float scale, cudaStream_t stream, int Nq = 0);

// bias_grad (optional, fp32 [3 * H * 64] = q | k | v) ACCUMULATES the column sums of dq / dk / dv (QKV bias gradient).
// N <= 256 runs the single fused kernel of attention_bwd_fused.cu (delta is not used); longer sequences run the
// split dQ / dKdV kernels.
int attention_bwd_tc(const void* q, const void* k, const void* v, long long ld, const void* o, long long ldo,
                     const void* dout, long long lddo, const float* lse, float* delta, void* dq, void* dk, void* dv,
                     long long lddqkv, int B, int N, int H, int head_dim, float scale, cudaStream_t stream,
                     float* bias_grad = nullptr, int bias_mask = 7, int Nq = 0);
// attention_fwd_fused.cu (N <= 224: persistent kernel, both query tiles of a head per CTA, operands prefetched)
int attention_fwd_fused(const void* q, const void* k, const void* v, long long ld, void* o, long long ldo, float* lse, int B,
                        int N, int Nq, int H, float scale, cudaStream_t stream);
// attention_fwd_long.cu (N > 224: persistent, two query tiles per CTA ping-pong over a K / V ring, online softmax)
int attention_fwd_long(const void* q, const void* k, const void* v, long long ld, void* o, long long ldo, float* lse, int B,
                       int N, int Nq, int H, float scale, cudaStream_t stream);
// attention_bwd_long.cu (256 < N <= 640): one fused kernel, CTAs own (image, head) pairs and walk their 128-key tiles;
// dK / dV complete per tile, dQ accumulated over the tiles in a private fp32 slab per CTA (dq_scratch, L2-resident).
long long attention_bwd_scratch_floats(int B, int N, int H);
int attention_bwd_long_max_queries();
int attention_bwd_long(const void* q, const void* k, const void* v, long long ld, const void* o, long long ldo,
                       const void* dout, long long lddo, const float* lse, float* dq_scratch, void* dq, void* dk, void* dv,
                       long long ldg, float* bias_grad, int bias_mask, int B, int N, int Nq, int H, float scale,
                       cudaStream_t stream);
// attention_bwd_fused.cu
int attention_bwd_fused(const void* q, const void* k, const void* v, long long ld, const void* o, long long ldo,
                        const void* dout, long long lddo, const float* lse, void* dq, void* dk, void* dv, long long ldg,
                        float* bias_grad, int bias_mask, int B, int N, int Nq, int H, float scale, cudaStream_t stream);
int attention_delta(const void* o, long long ldo, const void* dout, long long lddo, float* delta, int B, int N, int H,
                    cudaStream_t stream, int Nq = 0);

// engine.cu
struct VitLayout {
  // element offsets into the parameter arena (fp32, bf16 shadow and gradient arenas share the layout)
  long long cls, pos, patch_w, patch_b;
  long long layer0, layer_stride;
  // per-layer offsets relative to the layer base
  long long qkv_w, qkv_b, o_w, o_b, fc1_w, fc1_b, fc2_w, fc2_b, ln1_w, ln1_b, ln2_w, ln2_b;
  long long lnf_w, lnf_b, head_begin, cls_w, cls_b, total;
};
int vit_validate(const tic_vit_config* c);
VitLayout vit_layout(const tic_vit_config* c);
long long vit_workspace_bytes(const tic_vit_config* c, int B, int training);
int vit_forward(const tic_vit_config* c, const float* P32, const void* P16, const float* pixels, const void* patches,
                int B, void* workspace, long long workspace_bytes, int training, float* logits, cudaStream_t st);
int vit_backward(const tic_vit_config* c, const float* P32, const void* P16, int B, void* workspace,
                 long long workspace_bytes, const float* dlogits, float* G, int stage_begin, int stage_end,
                 int head_only, cudaStream_t st);


// fp32_path.cu: fp32-accurate inference (split-bf16 GEMMs on tcgen05, fp32 everything else)
long long vit_w6_elems(const tic_vit_config* c);
long long vit_workspace_bytes_f32(const tic_vit_config* c, int B);
int vit_prepare_w6(const tic_vit_config* c, const float* P32, void* w6, cudaStream_t st);
int vit_forward_f32(const tic_vit_config* c, const float* P32, const void* w6, const float* pixels, int B,
                    void* workspace, long long workspace_bytes, float* logits, cudaStream_t st);

}  // namespace tic
