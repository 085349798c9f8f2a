// fp32-accurate inference mode ("fp32 mode within 1e-4" of BASELINE.json north_star): the reference serves in plain
// fp32 with no autocast ($REF/TIC/utils/serve.py:99-101, $REF/web/runtime.py:115-116). Tensor cores have no fp32
// operand type, so every Linear runs as a SPLIT-bf16 GEMM on the same tcgen05 kernel:
//     x = x_hi + x_mid + x_lo (three bf16 terms, 24 mantissa bits),   w likewise,
//     x . w ~= hi.hi + hi.mid + mid.hi + hi.lo + lo.hi + mid.mid       (dropped terms are <= 2^-24 relative)
// The six cross terms are ONE GEMM over a 6x longer reduction dimension: A' = [hi|hi|mid|hi|lo|mid] (per row),
// B' = [hi|mid|hi|lo|hi|mid], fp32 accumulation in TMEM. LayerNorm, GELU (exact erf), residual adds, softmax and the
// attention matmuls stay in fp32 on the CUDA cores (attention is 3% of the FLOPs). Forward only: the reference never
// trains in fp32.
#include "tic_b200.h"
#include "tic_internal.cuh"

namespace tic {
namespace {

inline long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

// dst[r, 6K] <- split of src[r, K] (fp32). which = 0: A-side order [hi|hi|mid|hi|lo|mid]; 1: B-side [hi|mid|hi|lo|hi|mid]
__global__ void split3_kernel(const float* __restrict__ src, long long ld_src, long long rows, int K,
                              __nv_bfloat16* __restrict__ dst, int which) {
  const long long total = rows * (K / 4);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / (K / 4);
    const int c = static_cast<int>(i - r * (K / 4)) * 4;
    const float4 x = *reinterpret_cast<const float4*>(src + r * ld_src + c);
    const float xs[4] = {x.x, x.y, x.z, x.w};
    __nv_bfloat16 hi[4], mid[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      hi[j] = __float2bfloat16_rn(xs[j]);
      const float r1 = xs[j] - __bfloat162float(hi[j]);
      mid[j] = __float2bfloat16_rn(r1);
      lo[j] = __float2bfloat16_rn(r1 - __bfloat162float(mid[j]));
    }
    __nv_bfloat16* d = dst + r * (6LL * K) + c;
    const __nv_bfloat16* order_a[6] = {hi, hi, mid, hi, lo, mid};
    const __nv_bfloat16* order_b[6] = {hi, mid, hi, lo, hi, mid};
#pragma unroll
    for (int t = 0; t < 6; ++t) {
      const __nv_bfloat16* s = which == 0 ? order_a[t] : order_b[t];
      *reinterpret_cast<uint2*>(d + static_cast<long long>(t) * K) = *reinterpret_cast<const uint2*>(s);
    }
  }
}

// fp32 NCHW -> fp32 patch rows [B*G*G, 768], K ordered (c, py, px)
__global__ void patchify_f32_out_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int S) {
  const int G = S / 16;
  const long long total = static_cast<long long>(B) * G * G * 192;  // float4 chunks
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int chunk = static_cast<int>(i % 192);
    const long long row = i / 192;
    const int k = chunk * 4;
    const int c = k >> 8, py = (k & 255) >> 4, px = k & 15;
    const int b = static_cast<int>(row / (G * G));
    const int p = static_cast<int>(row - static_cast<long long>(b) * G * G);
    const int gy = p / G, gx = p - gy * G;
    reinterpret_cast<float4*>(out)[i] =
        __ldg(reinterpret_cast<const float4*>(x + ((static_cast<long long>(b) * 3 + c) * S + (gy * 16 + py)) * S + gx * 16 + px));
  }
}

// fp32 attention, one warp per query row: lane-per-key scores, warp softmax, lane-per-2-dims output.
constexpr int FA_WARPS = 8;
__global__ void __launch_bounds__(FA_WARPS * 32)
attn_fwd_f32_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, long long ld,
                    float* __restrict__ o, long long ldo, int N, int H, float scale) {
  extern __shared__ float fa_smem[];  // per warp: q[64] + p[Npad]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int Npad = (N + 31) / 32 * 32;
  float* sq = fa_smem + warp * (64 + Npad);
  float* sp = sq + 64;
  const long long tok0 = static_cast<long long>(b) * N;
  for (int row = blockIdx.x * FA_WARPS + warp; row < N; row += gridDim.x * FA_WARPS) {
    const float* qr = q + (tok0 + row) * ld + h * 64;
    sq[lane] = qr[lane];
    sq[lane + 32] = qr[lane + 32];
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < Npad; j += 32) {
      float s = -INFINITY;
      if (j < N) {
        const float4* kr = reinterpret_cast<const float4*>(k + (tok0 + j) * ld + h * 64);
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < 16; ++d) {
          const float4 kv = __ldg(kr + d);
          acc = fmaf(sq[4 * d + 0], kv.x, acc);
          acc = fmaf(sq[4 * d + 1], kv.y, acc);
          acc = fmaf(sq[4 * d + 2], kv.z, acc);
          acc = fmaf(sq[4 * d + 3], kv.w, acc);
        }
        s = acc * scale;
      }
      sp[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Npad; j += 32) {
      const float p = j < N ? expf(sp[j] - mx) : 0.f;
      sp[j] = p;
      sum += p;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < N; ++j) {
      const float p = sp[j];
      const float2 vv = __ldg(reinterpret_cast<const float2*>(v + (tok0 + j) * ld + h * 64) + lane);
      o0 = fmaf(p, vv.x, o0);
      o1 = fmaf(p, vv.y, o1);
    }
    const float inv = 1.0f / sum;
    reinterpret_cast<float2*>(o + (tok0 + row) * ldo + h * 64)[lane] = make_float2(o0 * inv, o1 * inv);
    __syncwarp();
  }
}

// logits[b, c] = h[b, :] . W[c, :] + bias[c], all fp32; one CTA per image, warp per class
__global__ void __launch_bounds__(128)
head_fwd_f32_kernel(const float* __restrict__ h, long long ldh, const float* __restrict__ w, const float* __restrict__ bias,
                    int D, int C, float* __restrict__ logits) {
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* hr = h + static_cast<long long>(b) * ldh;
  for (int c = warp; c < C; c += 4) {
    const float* wr = w + static_cast<long long>(c) * D;
    float s = 0.f;
    for (int i = lane; i < D; i += 32) s = fmaf(hr[i], __ldg(wr + i), s);
    s = warp_sum(s);
    if (lane == 0) logits[static_cast<long long>(b) * C + c] = s + bias[c];
  }
}

inline int grid_for(long long items) {
  long long g = (items + 255) / 256;
  if (g > 148LL * 16) g = 148LL * 16;
  return static_cast<int>(g < 1 ? 1 : g);
}

#define TIC_TRY(expr)             \
  do {                            \
    int rc__ = (expr);            \
    if (rc__ != kOk) return rc__; \
  } while (0)

struct W6Layout {
  long long patch, layer0, layer_stride, qkv, o, fc1, fc2, total;
};
W6Layout w6_layout(const tic_vit_config* c) {
  W6Layout L;
  const long long D = c->hidden, F = c->mlp;
  L.patch = 0;
  L.layer0 = D * 6 * 768;
  L.qkv = 0;
  L.o = L.qkv + 3 * D * 6 * D;
  L.fc1 = L.o + D * 6 * D;
  L.fc2 = L.fc1 + F * 6 * D;
  L.layer_stride = L.fc2 + D * 6 * F;
  L.total = L.layer0 + L.layer_stride * c->layers;
  return L;
}

struct F32Workspace {
  long long patches, patches6, x, xmid, h, h6, qkv, ctx, act, hcls, total;
};
F32Workspace carve_f32(const tic_vit_config* c, int B) {
  F32Workspace w{};
  const long long D = c->hidden, F = c->mlp;
  const long long G = c->image_size / 16, N = G * G + 1, P = N - 1, M = static_cast<long long>(B) * N;
  long long off = 0;
  auto take = [&](long long bytes) { long long o = off; off += align_up(bytes, 1024); return o; };
  w.patches = take(static_cast<long long>(B) * P * 768 * 4);
  w.patches6 = take(static_cast<long long>(B) * P * 6 * 768 * 2);
  w.x = take(M * D * 4);
  w.xmid = take(M * D * 4);
  w.h = take(M * F * 4);        // LayerNorm output [M, D] or GELU output [M, F]
  w.h6 = take(M * 6 * F * 2);   // split operand, up to [M, 6F]
  w.qkv = take(M * 3 * D * 4);
  w.ctx = take(M * D * 4);
  w.act = 0;
  w.hcls = take(static_cast<long long>(B) * D * 4);
  w.total = off;
  return w;
}

}  // namespace

long long vit_w6_elems(const tic_vit_config* c) { return w6_layout(c).total; }
long long vit_workspace_bytes_f32(const tic_vit_config* c, int B) { return carve_f32(c, B).total; }

int split3(const float* src, long long ld_src, long long rows, int K, void* dst_bf16, int which, cudaStream_t st) {
  if (K % 4 != 0) return set_error(kErrInvalidArg, "split3: K=%d must be a multiple of 4", K);
  if (rows <= 0) return kOk;
  ProfScope prof("split3_f32_to_bf16x6", 0.0, static_cast<double>(rows) * K * 16, st);
  split3_kernel<<<grid_for(rows * (K / 4)), 256, 0, st>>>(src, ld_src, rows, K, reinterpret_cast<__nv_bfloat16*>(dst_bf16), which);
  return check_launch("split3");
}

// Split every GEMM weight of the fp32 parameter arena into the [N, 6K] bf16 layout (once per weight update).
int vit_prepare_w6(const tic_vit_config* c, const float* P32, void* w6v, cudaStream_t st) {
  TIC_TRY(vit_validate(c));
  const VitLayout L = vit_layout(c);
  const W6Layout W = w6_layout(c);
  __nv_bfloat16* w6 = reinterpret_cast<__nv_bfloat16*>(w6v);
  const int D = c->hidden, F = c->mlp;
  TIC_TRY(split3(P32 + L.patch_w, 768, D, 768, w6 + W.patch, 1, st));
  for (int l = 0; l < c->layers; ++l) {
    const float* p = P32 + L.layer0 + static_cast<long long>(l) * L.layer_stride;
    __nv_bfloat16* w = w6 + W.layer0 + static_cast<long long>(l) * W.layer_stride;
    TIC_TRY(split3(p + L.qkv_w, D, 3LL * D, D, w + W.qkv, 1, st));
    TIC_TRY(split3(p + L.o_w, D, D, D, w + W.o, 1, st));
    TIC_TRY(split3(p + L.fc1_w, D, F, D, w + W.fc1, 1, st));
    TIC_TRY(split3(p + L.fc2_w, F, D, F, w + W.fc2, 1, st));
  }
  return kOk;
}

int attention_fwd_f32(const float* q, const float* k, const float* v, long long ld, float* o, long long ldo, int B, int N,
                      int H, float scale, cudaStream_t st) {
  if (B <= 0 || N <= 0) return kOk;
  const int Npad = (N + 31) / 32 * 32;
  const int smem = FA_WARPS * (64 + Npad) * 4;
  ProfScope prof("attention_fwd_f32", 4.0 * B * H * static_cast<double>(N) * N * 64, 0.0, st);
  dim3 grid((N + FA_WARPS - 1) / FA_WARPS, H, B);
  if (grid.x > 8) grid.x = 8;
  attn_fwd_f32_kernel<<<grid, FA_WARPS * 32, smem, st>>>(q, k, v, ld, o, ldo, N, H, scale);
  return check_launch("attention_fwd_f32");
}

int vit_forward_f32(const tic_vit_config* c, const float* P32, const void* w6v, const float* pixels, int B,
                    void* workspace, long long workspace_bytes, float* logits, cudaStream_t st) {
  TIC_TRY(vit_validate(c));
  if (B <= 0 || pixels == nullptr) return set_error(kErrInvalidArg, "vit_forward_f32: needs a non-empty pixel batch");
  const F32Workspace w = carve_f32(c, B);
  if (workspace_bytes < w.total)
    return set_error(kErrInvalidArg, "vit_forward_f32: workspace too small (%lld < %lld bytes)", workspace_bytes, w.total);
  const VitLayout L = vit_layout(c);
  const W6Layout W = w6_layout(c);
  const __nv_bfloat16* w6 = reinterpret_cast<const __nv_bfloat16*>(w6v);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const int D = c->hidden, F = c->mlp, H = c->heads, C = c->num_labels, S = c->image_size;
  const int G = S / 16, N = G * G + 1, Pn = N - 1, M = B * N;
  float* x = reinterpret_cast<float*>(ws + w.x);
  float* xmid = reinterpret_cast<float*>(ws + w.xmid);
  float* h = reinterpret_cast<float*>(ws + w.h);
  void* h6 = ws + w.h6;
  float* qkv = reinterpret_cast<float*>(ws + w.qkv);
  float* ctx = reinterpret_cast<float*>(ws + w.ctx);
  float* patches = reinterpret_cast<float*>(ws + w.patches);

  // embeddings
  {
    const long long total = static_cast<long long>(B) * G * G * 192;
    patchify_f32_out_f32_kernel<<<grid_for(total), 256, 0, st>>>(pixels, patches, B, S);
    TIC_TRY(check_launch("patchify_f32_out_f32"));
  }
  TIC_TRY(split3(patches, 768, static_cast<long long>(B) * Pn, 768, ws + w.patches6, 0, st));
  TIC_TRY(gemm_bf16(ws + w.patches6, 6 * 768, false, w6 + W.patch, 6 * 768, false, B * Pn, D, 6 * 768, kEpiF32PosEmbed, x, D,
                    nullptr, 0, P32 + L.patch_b, P32 + L.pos, D, Pn, 1, st, nullptr, true));
  TIC_TRY(cls_rows(P32 + L.cls, P32 + L.pos, x, B, N, D, st));

  for (int l = 0; l < c->layers; ++l) {
    const float* p32 = P32 + L.layer0 + static_cast<long long>(l) * L.layer_stride;
    const __nv_bfloat16* wl = w6 + W.layer0 + static_cast<long long>(l) * W.layer_stride;
    TIC_TRY(layernorm_fwd(x, D, p32 + L.ln1_w, p32 + L.ln1_b, c->ln_eps, M, D, nullptr, 0, h, D, nullptr, nullptr, st));
    TIC_TRY(split3(h, D, M, D, h6, 0, st));
    TIC_TRY(gemm_bf16(h6, 6 * D, false, wl + W.qkv, 6 * D, false, M, 3 * D, 6 * D, kEpiF32, qkv, 3 * D, nullptr, 0,
                      p32 + L.qkv_b, nullptr, 0, 0, 1, st));
    TIC_TRY(attention_fwd_f32(qkv, qkv + D, qkv + 2 * D, 3 * D, ctx, D, B, N, H, 0.125f, st));
    TIC_TRY(split3(ctx, D, M, D, h6, 0, st));
    TIC_TRY(gemm_bf16(h6, 6 * D, false, wl + W.o, 6 * D, false, M, D, 6 * D, kEpiF32Resid, xmid, D, nullptr, 0, p32 + L.o_b, x,
                      D, 0, 1, st, nullptr, true));
    TIC_TRY(layernorm_fwd(xmid, D, p32 + L.ln2_w, p32 + L.ln2_b, c->ln_eps, M, D, nullptr, 0, h, D, nullptr, nullptr, st));
    TIC_TRY(split3(h, D, M, D, h6, 0, st));
    TIC_TRY(gemm_bf16(h6, 6 * D, false, wl + W.fc1, 6 * D, false, M, F, 6 * D, kEpiF32Gelu, h, F, nullptr, 0, p32 + L.fc1_b,
                      nullptr, 0, 0, 1, st));
    TIC_TRY(split3(h, F, M, F, h6, 0, st));
    TIC_TRY(gemm_bf16(h6, 6LL * F, false, wl + W.fc2, 6LL * F, false, M, D, 6 * F, kEpiF32Resid, x, D, nullptr, 0, p32 + L.fc2_b,
                      xmid, D, 0, 1, st, nullptr, true));
  }
  float* hcls = reinterpret_cast<float*>(ws + w.hcls);
  TIC_TRY(layernorm_fwd(x, static_cast<long long>(N) * D, P32 + L.lnf_w, P32 + L.lnf_b, c->ln_eps, B, D, nullptr, 0, hcls, D,
                        nullptr, nullptr, st));
  head_fwd_f32_kernel<<<B, 128, 0, st>>>(hcls, D, P32 + L.cls_w, P32 + L.cls_b, D, C, logits);
  return check_launch("head_fwd_f32");
}

}  // namespace tic
