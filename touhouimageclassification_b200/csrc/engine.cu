// Native ViT engine: one C call runs the whole encoder forward (or backward) as a fixed sequence of
// kernel launches on the caller's stream -- no Python between kernels, no allocation, graph-capturable.
// Mirrors ViTForImageClassification.forward (modeling_vit.py:620-653) = ViTEmbeddings (:100-128) +
// Lyr x ViTLayer (:328-346) + final LayerNorm (:455) + classifier on the CLS row (:641-642) [a2-a12],
// and the autograd backward of the same graph.
//
// Memory model (all caller-provided):
//   parameter arena  fp32, private order (see vit_layout) with q/k/v contiguous so the QKV projection is
//                    one [3D, D] GEMM; the nn.Parameters on the Python side are views into it.
//   bf16 shadow      same layout, GEMM operands; refreshed by the fused AdamW or tic_cast_f32_to_bf16.
//   gradient arena   fp32, same layout; backward ACCUMULATES into it (caller zeroes it per step).
//   workspace        activations saved for backward + scratch, carved by vit_workspace().
#include "tic_b200.h"
#include "tic_internal.cuh"

#include <cstdlib>

namespace tic {

namespace {
inline long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }
}  // namespace

int vit_validate(const tic_vit_config* c) {
  if (c == nullptr) return set_error(kErrInvalidArg, "vit: null config");
  if (c->patch_size != 16) return set_error(kErrUnsupported, "vit: patch_size=%d (only 16)", c->patch_size);
  if (c->image_size <= 0 || c->image_size % 16) return set_error(kErrInvalidArg, "vit: bad image_size=%d", c->image_size);
  if (c->hidden % 128 || c->hidden <= 0) return set_error(kErrUnsupported, "vit: hidden=%d must be a multiple of 128", c->hidden);
  if (c->heads <= 0 || c->hidden != c->heads * 64) return set_error(kErrUnsupported, "vit: head_dim must be 64 (hidden=%d heads=%d)", c->hidden, c->heads);
  if (c->mlp % 8 || c->mlp <= 0) return set_error(kErrInvalidArg, "vit: bad mlp=%d", c->mlp);
  if (c->layers <= 0 || c->num_labels <= 0) return set_error(kErrInvalidArg, "vit: bad layers/num_labels");
  return kOk;
}

VitLayout vit_layout(const tic_vit_config* c) {
  VitLayout L;
  const long long D = c->hidden, F = c->mlp, C = c->num_labels;
  const long long G = c->image_size / 16, N = G * G + 1;
  long long off = 0;
  auto take = [&](long long n) { long long o = off; off += align_up(n, 64); return o; };
  L.cls = take(D);
  L.pos = take(N * D);
  L.patch_w = take(D * 768);
  L.patch_b = take(D);
  L.layer0 = off;
  long long lo = 0;
  auto ltake = [&](long long n) { long long o = lo; lo += align_up(n, 64); return o; };
  L.qkv_w = ltake(3 * D * D);
  L.qkv_b = ltake(3 * D);
  L.o_w = ltake(D * D);
  L.o_b = ltake(D);
  L.fc1_w = ltake(F * D);
  L.fc1_b = ltake(F);
  L.fc2_w = ltake(D * F);
  L.fc2_b = ltake(D);
  L.ln1_w = ltake(D);
  L.ln1_b = ltake(D);
  L.ln2_w = ltake(D);
  L.ln2_b = ltake(D);
  L.layer_stride = lo;
  off += lo * c->layers;
  L.lnf_w = take(D);
  L.lnf_b = take(D);
  L.head_begin = off;
  L.cls_w = take(C * D);
  L.cls_b = take(C);
  L.total = off;
  return L;
}

namespace {

struct Workspace {
  // saved for backward (training) / scratch (inference)
  long long patches, x, xmid, h1, qkv, ctx, lse, h2, pre, act, stats, hcls, logits_scratch;
  // backward scratch
  long long dx, dxb, dact, dh, dqkv, dctx, delta, dhcls, dpatch;
  long long total;
  // strides between per-layer copies (0 in inference mode: buffers are reused)
  long long s_x, s_tok_d_bf16, s_qkv, s_lse, s_f, s_stats;
};

Workspace carve(const tic_vit_config* c, int B, bool training) {
  Workspace w{};
  const long long D = c->hidden, F = c->mlp, H = c->heads, Lyr = c->layers;
  const long long G = c->image_size / 16, N = G * G + 1, P = N - 1, M = static_cast<long long>(B) * N;
  long long off = 0;
  auto take = [&](long long bytes) { long long o = off; off += align_up(bytes, 1024); return o; };
  const long long nl = training ? Lyr : 1;
  w.patches = take(static_cast<long long>(B) * P * 768 * 2);
  w.s_x = align_up(M * D * 4, 1024);
  w.x = take(w.s_x * (training ? Lyr + 1 : 1));
  w.xmid = take(w.s_x * nl);
  w.s_tok_d_bf16 = align_up(M * D * 2, 1024);
  w.h1 = take(w.s_tok_d_bf16 * nl);
  w.s_qkv = align_up(M * 3 * D * 2, 1024);
  w.qkv = take(w.s_qkv * nl);
  w.ctx = take(w.s_tok_d_bf16 * nl);
  w.s_lse = align_up(static_cast<long long>(B) * H * N * 4, 1024);
  w.lse = take(w.s_lse * nl);
  w.h2 = take(w.s_tok_d_bf16 * nl);
  w.s_f = align_up(M * F * 2, 1024);
  w.pre = training ? take(w.s_f * nl) : 0;
  w.act = take(w.s_f * nl);
  w.s_stats = align_up(M * 4, 1024);
  w.stats = take(w.s_stats * 4 * nl + static_cast<long long>(B) * 8 + 1024);  // ln1/ln2 mean+rstd per layer, then final LN
  w.hcls = take(static_cast<long long>(B) * D * 2);
  w.logits_scratch = take(static_cast<long long>(B) * c->num_labels * 4);
  if (training) {
    w.dx = take(M * D * 4);
    w.dxb = take(M * D * 2);
    w.dact = take(M * F * 2);
    w.dh = take(M * D * 2);
    w.dqkv = take(M * 3 * D * 2);
    w.dctx = take(M * D * 2);
    w.delta = take(attention_bwd_scratch_floats(B, static_cast<int>(N), static_cast<int>(H)) * 4);  // delta (+ dQ partials, N > 256)
    w.dhcls = take(static_cast<long long>(B) * D * 2);
    w.dpatch = take(static_cast<long long>(B) * P * D * 2);
  } else {
    w.s_x = w.s_tok_d_bf16 = w.s_qkv = w.s_lse = w.s_f = w.s_stats = 0;
  }
  w.total = off;
  return w;
}

// Pick a split-K factor for a wgrad GEMM so that tiles * splits fills whole waves of the 74 CTA pairs (256 x 256 output
// tiles, one pair per tile). Fewer splits win ties: every split adds one pass of fp32 red.add traffic over the output.
int pick_splits(int Mo, int No, int K) {
  const int tiles = ((Mo + 255) / 256) * ((No + 255) / 256);
  const int workers = 74;
  const int kblocks = (K + 63) / 64;
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= 32; ++s) {
    if (s > 1 && kblocks / s < 16) break;
    const long long t = static_cast<long long>(tiles) * s;
    const double eff = static_cast<double>(t) / (static_cast<double>((t + workers - 1) / workers) * workers);
    if (eff > best_eff + 0.03) { best_eff = eff; best = s; }
  }
  return best;
}

// The last encoder layer is evaluated on the CLS rows only (every attention kernel takes a query subset).
bool cls_only_last_layer(const tic_vit_config* c) {
  static const bool full = std::getenv("TIC_FULL_LAST_LAYER") != nullptr;  // development knob: A/B against the full layer
  (void)c;
  return !full;
}

#define TIC_TRY(expr)        \
  do {                       \
    int rc__ = (expr);       \
    if (rc__ != kOk) return rc__; \
  } while (0)

}  // namespace

long long vit_workspace_bytes(const tic_vit_config* c, int B, int training) {
  return carve(c, B, training != 0).total;
}

int vit_forward(const tic_vit_config* c, const float* P32, const void* P16v, const float* pixels,
                const void* patches_in, int B, void* workspace, long long workspace_bytes, int training,
                float* logits, cudaStream_t st) {
  TIC_TRY(vit_validate(c));
  if (B <= 0) return set_error(kErrInvalidArg, "vit_forward: empty batch");
  const Workspace w = carve(c, B, training != 0);
  if (workspace_bytes < w.total)
    return set_error(kErrInvalidArg, "vit_forward: workspace too small (%lld < %lld bytes)", workspace_bytes, w.total);
  if ((pixels == nullptr) == (patches_in == nullptr))
    return set_error(kErrInvalidArg, "vit_forward: give exactly one of pixels / patches");
  const VitLayout L = vit_layout(c);
  const __nv_bfloat16* P16 = reinterpret_cast<const __nv_bfloat16*>(P16v);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const int D = c->hidden, F = c->mlp, H = c->heads, C = c->num_labels, S = c->image_size;
  const int G = S / 16, N = G * G + 1, Pn = N - 1;
  const int M = B * N;
  const float scale = 0.125f;  // 1 / sqrt(head_dim = 64)
  // small inference forwards are dominated by launch gaps: overlap each kernel's prologue with its predecessor's tail
  PdlScope pdl(training == 0 && M <= 8192);

  // ---- embeddings: patchify -> projection GEMM (+bias, +pos) -> CLS rows
  const void* patches = patches_in;
  if (pixels != nullptr) {
    TIC_TRY(patchify_f32(pixels, ws + w.patches, B, S, st));
    patches = ws + w.patches;
  } else if (training) {
    // keep a copy for the patch-projection wgrad
    cudaError_t e = cudaMemcpyAsync(ws + w.patches, patches_in, static_cast<size_t>(B) * Pn * 768 * 2,
                                    cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return set_error(kErrCuda, "vit_forward: memcpy patches: %s", cudaGetErrorString(e));
    patches = ws + w.patches;
  }
  float* x0 = reinterpret_cast<float*>(ws + w.x);
  TIC_TRY(gemm_bf16(patches, 768, false, P16 + L.patch_w, 768, false, B * Pn, D, 768, kEpiF32PosEmbed, x0, D, nullptr, 0,
                    P32 + L.patch_b, P32 + L.pos, D, Pn, 1, st));
  TIC_TRY(cls_rows(P32 + L.cls, P32 + L.pos, x0, B, N, D, st));

  // ---- encoder
  for (int l = 0; l < c->layers; ++l) {
    const long long po = L.layer0 + static_cast<long long>(l) * L.layer_stride;
    const float* p32 = P32 + po;
    const __nv_bfloat16* p16 = P16 + po;
    float* x_in = reinterpret_cast<float*>(ws + w.x + w.s_x * l);
    float* x_out = reinterpret_cast<float*>(ws + w.x + w.s_x * (training ? l + 1 : 0));
    float* xmid = reinterpret_cast<float*>(ws + w.xmid + w.s_x * l);
    void* h1 = ws + w.h1 + w.s_tok_d_bf16 * l;
    __nv_bfloat16* qkv = reinterpret_cast<__nv_bfloat16*>(ws + w.qkv + w.s_qkv * l);
    void* ctx = ws + w.ctx + w.s_tok_d_bf16 * l;
    float* lse = reinterpret_cast<float*>(ws + w.lse + w.s_lse * l);
    void* h2 = ws + w.h2 + w.s_tok_d_bf16 * l;
    void* pre = training ? ws + w.pre + w.s_f * l : nullptr;
    void* act = ws + w.act + w.s_f * l;
    float* stats = reinterpret_cast<float*>(ws + w.stats + w.s_stats * 4 * l);
    const long long ss = w.s_stats / 4;
    float *mean1 = training ? stats : nullptr, *rstd1 = training ? stats + ss : nullptr;
    float *mean2 = training ? stats + 2 * ss : nullptr, *rstd2 = training ? stats + 3 * ss : nullptr;

    // The classifier reads only the CLS row of the last layer's output (modeling_vit.py:641), and inside a layer a
    // token's output depends on the other tokens only through the keys and values. So the last layer projects K and V
    // for every token but runs the query projection, the attention rows, the output projection and the MLP on the
    // B CLS rows alone (row pitch N*width in the same buffers).
    const bool cls_only = cls_only_last_layer(c) && l == c->layers - 1;
    const int rows = cls_only ? B : M;
    const long long rm = cls_only ? N : 1;  // row pitch multiplier of the per-token buffers

    TIC_TRY(layernorm_fwd(x_in, D, p32 + L.ln1_w, p32 + L.ln1_b, c->ln_eps, M, D, h1, D, nullptr, 0, mean1, rstd1, st));
    if (!cls_only) {
      TIC_TRY(gemm_bf16(h1, D, false, p16 + L.qkv_w, D, false, M, 3 * D, D, kEpiBf16, qkv, 3 * D, nullptr, 0,
                        p32 + L.qkv_b, nullptr, 0, 0, 1, st));
    } else {
      TIC_TRY(gemm_bf16(h1, D, false, p16 + L.qkv_w + static_cast<long long>(D) * D, D, false, M, 2 * D, D, kEpiBf16,
                        qkv + D, 3 * D, nullptr, 0, p32 + L.qkv_b + D, nullptr, 0, 0, 1, st));
      TIC_TRY(gemm_bf16(h1, rm * D, false, p16 + L.qkv_w, D, false, B, D, D, kEpiBf16, qkv, rm * 3 * D, nullptr, 0,
                        p32 + L.qkv_b, nullptr, 0, 0, 1, st));
    }
    TIC_TRY(attention_fwd_tc(qkv, qkv + D, qkv + 2 * D, 3 * D, ctx, D, training ? lse : nullptr, B, N, H, 64, scale, st,
                             cls_only ? 1 : 0));
    TIC_TRY(gemm_bf16(ctx, rm * D, false, p16 + L.o_w, D, false, rows, D, D, kEpiF32Resid, xmid, rm * D, nullptr, 0,
                      p32 + L.o_b, x_in, rm * D, 0, 1, st));
    TIC_TRY(layernorm_fwd(xmid, rm * D, p32 + L.ln2_w, p32 + L.ln2_b, c->ln_eps, rows, D, h2, rm * D, nullptr, 0, mean2,
                          rstd2, st));
    TIC_TRY(gemm_bf16(h2, rm * D, false, p16 + L.fc1_w, D, false, rows, F, D, kEpiBf16Gelu, act, rm * F, pre, rm * F,
                      p32 + L.fc1_b, nullptr, 0, 0, 1, st));
    TIC_TRY(gemm_bf16(act, rm * F, false, p16 + L.fc2_w, F, false, rows, D, F, kEpiF32Resid, x_out, rm * D, nullptr, 0,
                      p32 + L.fc2_b, xmid, rm * D, 0, 1, st));
  }

  // ---- final LayerNorm on the CLS rows only (the only rows the classifier reads) + head
  float* x_last = reinterpret_cast<float*>(ws + w.x + w.s_x * (training ? c->layers : 0));
  float* fstats = reinterpret_cast<float*>(ws + w.stats + w.s_stats * 4 * (training ? c->layers : 0));
  TIC_TRY(layernorm_fwd(x_last, static_cast<long long>(N) * D, P32 + L.lnf_w, P32 + L.lnf_b, c->ln_eps, B, D,
                        ws + w.hcls, D, nullptr, 0, fstats, fstats + B, st));
  TIC_TRY(head_fwd(ws + w.hcls, D, P16 + L.cls_w, P32 + L.cls_b, B, D, C, 1, logits, st));
  return kOk;
}

// stage 0 = classifier + final LayerNorm, stages 1..Lyr = encoder layers Lyr-1..0, stage Lyr+1 = embeddings.
int vit_backward(const tic_vit_config* c, const float* P32, const void* P16v, int B, void* workspace,
                 long long workspace_bytes, const float* dlogits, float* G, int stage_begin, int stage_end,
                 int head_only, cudaStream_t st) {
  TIC_TRY(vit_validate(c));
  const Workspace w = carve(c, B, true);
  if (workspace_bytes < w.total)
    return set_error(kErrInvalidArg, "vit_backward: workspace too small (%lld < %lld bytes)", workspace_bytes, w.total);
  const VitLayout L = vit_layout(c);
  const __nv_bfloat16* P16 = reinterpret_cast<const __nv_bfloat16*>(P16v);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const int D = c->hidden, F = c->mlp, H = c->heads, C = c->num_labels, Lyr = c->layers;
  const int Gd = c->image_size / 16, N = Gd * Gd + 1, Pn = N - 1;
  const int M = B * N;
  const float scale = 0.125f;
  float* dx = reinterpret_cast<float*>(ws + w.dx);
  __nv_bfloat16* dxb = reinterpret_cast<__nv_bfloat16*>(ws + w.dxb);
  void* dact = ws + w.dact;
  void* dh = ws + w.dh;
  __nv_bfloat16* dqkv = reinterpret_cast<__nv_bfloat16*>(ws + w.dqkv);
  void* dctx = ws + w.dctx;
  float* delta = reinterpret_cast<float*>(ws + w.delta);
  if (stage_begin < 0) stage_begin = 0;
  if (stage_end > Lyr + 2) stage_end = Lyr + 2;

  for (int stage = stage_begin; stage < stage_end; ++stage) {
    if (stage == 0) {
      const bool need_dh = !head_only;
      TIC_TRY(head_bwd(dlogits, ws + w.hcls, D, P16 + L.cls_w, B, D, C, need_dh ? ws + w.dhcls : nullptr, D,
                       G + L.cls_w, G + L.cls_b, st));
      if (head_only) continue;
      // only the CLS rows carry gradient here; the bf16 copy is read at CLS rows alone when the last layer is CLS-only
      cudaError_t e1 = cudaMemsetAsync(dx, 0, static_cast<size_t>(M) * D * 4, st);
      cudaError_t e2 = cls_only_last_layer(c) ? cudaSuccess : cudaMemsetAsync(dxb, 0, static_cast<size_t>(M) * D * 2, st);
      if (e1 != cudaSuccess || e2 != cudaSuccess) return set_error(kErrCuda, "vit_backward: memset failed");
      const float* x_last = reinterpret_cast<const float*>(ws + w.x + w.s_x * Lyr);
      const float* fstats = reinterpret_cast<const float*>(ws + w.stats + w.s_stats * 4 * Lyr);
      const long long rs = static_cast<long long>(N) * D;
      // dx of the last layer's output: its column sums are that layer's fc2 bias gradient
      float* fc2_b_last = G + L.layer0 + static_cast<long long>(Lyr - 1) * L.layer_stride + L.fc2_b;
      TIC_TRY(layernorm_bwd(ws + w.dhcls, D, x_last, rs, fstats, fstats + B, P32 + L.lnf_w, nullptr, 0, B, D, dx, rs,
                            dxb, rs, G + L.lnf_w, G + L.lnf_b, fc2_b_last, st));
    } else if (stage <= Lyr) {
      if (head_only) continue;
      const int l = Lyr - stage;
      const long long po = L.layer0 + static_cast<long long>(l) * L.layer_stride;
      const float* p32 = P32 + po;
      const __nv_bfloat16* p16 = P16 + po;
      float* g = G + po;
      const float* x_in = reinterpret_cast<const float*>(ws + w.x + w.s_x * l);
      const float* xmid = reinterpret_cast<const float*>(ws + w.xmid + w.s_x * l);
      const void* h1 = ws + w.h1 + w.s_tok_d_bf16 * l;
      const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(ws + w.qkv + w.s_qkv * l);
      const void* ctx = ws + w.ctx + w.s_tok_d_bf16 * l;
      const float* lse = reinterpret_cast<const float*>(ws + w.lse + w.s_lse * l);
      const void* h2 = ws + w.h2 + w.s_tok_d_bf16 * l;
      const void* pre = ws + w.pre + w.s_f * l;
      const void* act = ws + w.act + w.s_f * l;
      const float* stats = reinterpret_cast<const float*>(ws + w.stats + w.s_stats * 4 * l);
      const long long ss = w.s_stats / 4;
      const float *mean1 = stats, *rstd1 = stats + ss, *mean2 = stats + 2 * ss, *rstd2 = stats + 3 * ss;

      const bool cls_only = cls_only_last_layer(c) && l == Lyr - 1;  // see vit_forward: this layer ran on the CLS rows
      const int rows = cls_only ? B : M;
      const long long rm = cls_only ? N : 1;

      // fc2: x_out = xmid + act W2^T + b2        (dy = dxb, the bf16 copy of the residual-stream gradient)
      TIC_TRY(gemm_bf16(dxb, rm * D, false, p16 + L.fc2_w, F, true, rows, F, D, kEpiBf16DGelu, dact, rm * F, nullptr, 0,
                        nullptr, pre, rm * F, 0, 1, st, g + L.fc1_b));  // dact <- dpre = (dy W2) * gelu'(pre) (saved by the forward); fc1 bias grad = colsum(dpre)
      TIC_TRY(gemm_bf16(dxb, rm * D, true, act, rm * F, true, D, F, rows, kEpiF32Atomic, g + L.fc2_w, F, nullptr, 0,
                        nullptr, nullptr, 0, 0, pick_splits(D, F, rows), st));
      // fc1: pre = h2 W1^T + b1
      TIC_TRY(gemm_bf16(dact, rm * F, false, p16 + L.fc1_w, D, true, rows, D, F, kEpiBf16, dh, rm * D, nullptr, 0, nullptr,
                        nullptr, 0, 0, 1, st));
      TIC_TRY(gemm_bf16(dact, rm * F, true, h2, rm * D, true, F, D, rows, kEpiF32Atomic, g + L.fc1_w, D, nullptr, 0,
                        nullptr, nullptr, 0, 0, pick_splits(F, D, rows), st));
      // layernorm_after + residual
      TIC_TRY(layernorm_bwd(dh, rm * D, xmid, rm * D, mean2, rstd2, p32 + L.ln2_w, dx, rm * D, rows, D, dx, rm * D, dxb,
                            rm * D, g + L.ln2_w, g + L.ln2_b, g + L.o_b, st));  // + out-proj bias grad = colsum(dxb)
      // attention output projection: xmid = x_in + ctx Wo^T + bo
      // dctx = dxb Wo; its column sums are the VALUE bias gradient: sum_k dV[k,:] = sum_q (sum_k P[q,k]) dO[q,:] = sum_q dO[q,:]
      // because every softmax row sums to one -- accumulated for free in this GEMM's epilogue.
      TIC_TRY(gemm_bf16(dxb, rm * D, false, p16 + L.o_w, D, true, rows, D, D, kEpiBf16, dctx, rm * D, nullptr, 0, nullptr,
                        nullptr, 0, 0, 1, st, g + L.qkv_b + 2 * D));
      TIC_TRY(gemm_bf16(dxb, rm * D, true, ctx, rm * D, true, D, D, rows, kEpiF32Atomic, g + L.o_w, D, nullptr, 0, nullptr,
                        nullptr, 0, 0, pick_splits(D, D, rows), st));
      // attention core
      TIC_TRY(attention_bwd_tc(qkv, qkv + D, qkv + 2 * D, 3 * D, ctx, D, dctx, D, lse, delta, dqkv, dqkv + D, dqkv + 2 * D,
                            3 * D, B, N, H, 64, scale, st, g + L.qkv_b, 1, cls_only ? 1 : 0));
      // bias gradients of the fused QKV Linear: query = colsum(dq), accumulated by the attention backward (mask 1);
      // value = colsum(dctx), above; key = exactly 0 (sum_k dS[q,k] = 0: softmax is invariant to a shift of all keys,
      // SURVEY Appendix D.3 -- the reference computes ~1e-10 of rounding noise here), so it is left untouched.
      // fused QKV projection
      if (!cls_only) {
        TIC_TRY(gemm_bf16(dqkv, 3 * D, false, p16 + L.qkv_w, D, true, M, D, 3 * D, kEpiBf16, dh, D, nullptr, 0, nullptr,
                          nullptr, 0, 0, 1, st));
        TIC_TRY(gemm_bf16(dqkv, 3 * D, true, h1, D, true, 3 * D, D, M, kEpiF32Atomic, g + L.qkv_w, D, nullptr, 0, nullptr,
                          nullptr, 0, 0, pick_splits(3 * D, D, M), st));
      } else {
        // dq exists on the CLS rows only: every row gets its key/value part (K = 2D), then the CLS rows are recomputed
        // with all three parts (K = 3D); the weight gradient is taken per part over the rows that have it.
        const long long kv = static_cast<long long>(D) * D;
        TIC_TRY(gemm_bf16(dqkv + D, 3 * D, false, p16 + L.qkv_w + kv, D, true, M, D, 2 * D, kEpiBf16, dh, D, nullptr, 0,
                          nullptr, nullptr, 0, 0, 1, st));
        TIC_TRY(gemm_bf16(dqkv, rm * 3 * D, false, p16 + L.qkv_w, D, true, B, D, 3 * D, kEpiBf16, dh, rm * D, nullptr, 0,
                          nullptr, nullptr, 0, 0, 1, st));
        TIC_TRY(gemm_bf16(dqkv + D, 3 * D, true, h1, D, true, 2 * D, D, M, kEpiF32Atomic, g + L.qkv_w + kv, D, nullptr, 0,
                          nullptr, nullptr, 0, 0, pick_splits(2 * D, D, M), st));
        TIC_TRY(gemm_bf16(dqkv, rm * 3 * D, true, h1, rm * D, true, D, D, B, kEpiF32Atomic, g + L.qkv_w, D, nullptr, 0,
                          nullptr, nullptr, 0, 0, pick_splits(D, D, B), st));
      }
      // layernorm_before + residual
      // dx of layer l's input == gradient of layer (l-1)'s output: its column sums are that layer's fc2 bias gradient
      TIC_TRY(layernorm_bwd(dh, D, x_in, D, mean1, rstd1, p32 + L.ln1_w, dx, D, M, D, dx, D, l > 0 ? dxb : nullptr, D,
                            g + L.ln1_w, g + L.ln1_b, l > 0 ? g - L.layer_stride + L.fc2_b : nullptr, st));
    } else {
      if (head_only) continue;
      TIC_TRY(embed_bwd(dx, B, N, D, G + L.pos, G + L.cls, ws + w.dpatch, st));
      TIC_TRY(gemm_bf16(ws + w.dpatch, D, true, ws + w.patches, 768, true, D, 768, B * Pn, kEpiF32Atomic, G + L.patch_w,
                        768, nullptr, 0, nullptr, nullptr, 0, 0, pick_splits(D, 768, B * Pn), st));
      TIC_TRY(colsum_bf16(ws + w.dpatch, D, B * Pn, D, G + L.patch_b, st));
    }
  }
  return kOk;
}

}  // namespace tic
