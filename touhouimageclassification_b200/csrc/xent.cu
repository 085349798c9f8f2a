// Classifier head (CLS row -> 120 logits) and fused softmax-cross-entropy forward + backward.
// Replaces nn.Linear classifier (modeling_vit.py:613,641-642 [a11]) and F.cross_entropy with integer
// targets (finetune.py:61 [a13]) or soft MixUp/CutMix targets (ntrain.py:48 [a15]).
// The head is tiny (B x D x 120): CUDA-core dot products, fp32 accumulate, bf16-rounded like autocast.
#include "tic_internal.cuh"

namespace tic {
namespace {

// logits[b, c] = round( sum_d h[b,d] * W[c,d] + bias[c] ); h, W bf16; CTA (b, g) computes classes [8 g, 8 g + 8) of image
// b, two per warp (one CTA per image walked its 120 classes in 30 dependent steps: 40 us, the longest kernel of a
// batch-1 forward).
constexpr int kHeadClassesPerCta = 8;
__global__ void __launch_bounds__(128)
head_fwd_kernel(const __nv_bfloat16* __restrict__ h, long long ldh, const __nv_bfloat16* __restrict__ w,
                const float* __restrict__ bias, int D, int C, int round_out, float* __restrict__ logits) {
  extern __shared__ uint4 hs[];  // D bf16
  const int b = blockIdx.x;
  const uint4* hr = reinterpret_cast<const uint4*>(h + static_cast<long long>(b) * ldh);
  for (int i = threadIdx.x; i < D / 8; i += blockDim.x) hs[i] = hr[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c_end = min(C, static_cast<int>(blockIdx.y + 1) * kHeadClassesPerCta);
  for (int c = blockIdx.y * kHeadClassesPerCta + warp; c < c_end; c += 4) {
    const uint4* wr = reinterpret_cast<const uint4*>(w + static_cast<long long>(c) * D);
    float s = 0.f;
    for (int i = lane; i < D / 8; i += 32) {
      const uint4 a = hs[i], v = __ldg(wr + i);
      s += bf16_lo(a.x) * bf16_lo(v.x) + bf16_hi(a.x) * bf16_hi(v.x);
      s += bf16_lo(a.y) * bf16_lo(v.y) + bf16_hi(a.y) * bf16_hi(v.y);
      s += bf16_lo(a.z) * bf16_lo(v.z) + bf16_hi(a.z) * bf16_hi(v.z);
      s += bf16_lo(a.w) * bf16_lo(v.w) + bf16_hi(a.w) * bf16_hi(v.w);
    }
    s = warp_sum(s);
    if (lane == 0) {
      s += bias[c];
      logits[static_cast<long long>(b) * C + c] = round_out ? round_bf16(s) : s;
    }
  }
}

// dh[b, d] = sum_c dlogits[b,c] * W[c,d]  (bf16 out, one CTA per image)
__global__ void __launch_bounds__(256)
head_bwd_dh_kernel(const float* __restrict__ dlogits, const __nv_bfloat16* __restrict__ w, int D, int C,
                   __nv_bfloat16* __restrict__ dh, long long lddh) {
  extern __shared__ float dl[];
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) dl[c] = dlogits[static_cast<long long>(b) * C + c];
  __syncthreads();
  for (int d2 = threadIdx.x; d2 < D / 2; d2 += blockDim.x) {
    float s0 = 0.f, s1 = 0.f;
    for (int c = 0; c < C; ++c) {
      const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(w + static_cast<long long>(c) * D) + d2);
      s0 += dl[c] * bf16_lo(v);
      s1 += dl[c] * bf16_hi(v);
    }
    reinterpret_cast<uint32_t*>(dh + static_cast<long long>(b) * lddh)[d2] = pack_bf16x2(s0, s1);
  }
}

// dW[c, d] += sum_b dlogits[b,c] * h[b,d];  db[c] += sum_b dlogits[b,c]   (one CTA per class)
__global__ void __launch_bounds__(256)
head_bwd_dw_kernel(const float* __restrict__ dlogits, const __nv_bfloat16* __restrict__ h, long long ldh, int B, int D,
                   int C, float* __restrict__ dW, float* __restrict__ db) {
  const int c = blockIdx.x;
  for (int d2 = threadIdx.x; d2 < D / 2; d2 += blockDim.x) {
    float s0 = 0.f, s1 = 0.f;
    for (int b = 0; b < B; ++b) {
      const float g = __ldg(dlogits + static_cast<long long>(b) * C + c);
      const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(h + static_cast<long long>(b) * ldh) + d2);
      s0 += g * bf16_lo(v);
      s1 += g * bf16_hi(v);
    }
    float2* dst = reinterpret_cast<float2*>(dW + static_cast<long long>(c) * D) + d2;
    float2 o = *dst;
    o.x += s0; o.y += s1;
    *dst = o;
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dlogits[static_cast<long long>(b) * C + c];
    db[c] += s;
  }
}

// One CTA of 32 warps (one row per warp at a time; a single CTA keeps the loss reduction order fixed, so the loss is
// bit-reproducible). Row-wise log-softmax in fp32; loss = mean_b( -sum_c y[b,c] * logp[b,c] ) * (1/B);
// dlogits[b,c] = (softmax[b,c] * sum_c y[b,c] - y[b,c]) * grad_scale   (grad_scale = upstream / global batch).
__global__ void __launch_bounds__(1024)
xent_kernel(const float* __restrict__ logits, const long long* __restrict__ hard, const float* __restrict__ soft,
            int B, int C, float grad_scale, int round_grad, float* __restrict__ loss, float* __restrict__ dlogits,
            int* __restrict__ correct) {
  __shared__ float wloss[32];
  __shared__ int wcorr[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float lsum = 0.f;
  int csum = 0;
  for (int b = warp; b < B; b += 32) {
    const float* lr = logits + static_cast<long long>(b) * C;
    float mx = -INFINITY;
    int amax = 0;
    for (int c = lane; c < C; c += 32) {
      const float v = lr[c];
      if (v > mx) { mx = v; amax = c; }
    }
    // argmax with first-index tie-break (torch.argmax semantics on ties are unspecified; first is what CPU gives)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oa = __shfl_xor_sync(0xffffffffu, amax, o);
      if (om > mx || (om == mx && oa < amax)) { mx = om; amax = oa; }
    }
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += __expf(lr[c] - mx);
    se = warp_sum(se);
    const float lse = mx + __logf(se);
    float rl = 0.f, ysum = 0.f;
    const long long t = hard ? hard[b] : -1;
    for (int c = lane; c < C; c += 32) {
      const float y = hard ? (c == t ? 1.f : 0.f) : soft[static_cast<long long>(b) * C + c];
      rl -= y * (lr[c] - lse);
      ysum += y;
    }
    rl = warp_sum(rl);
    ysum = warp_sum(ysum);
    if (dlogits) {
      for (int c = lane; c < C; c += 32) {
        const float y = hard ? (c == t ? 1.f : 0.f) : soft[static_cast<long long>(b) * C + c];
        const float g = (__expf(lr[c] - lse) * ysum - y) * grad_scale;
        dlogits[static_cast<long long>(b) * C + c] = round_grad ? round_bf16(g) : g;
      }
    }
    if (lane == 0) {
      lsum += rl;
      if (hard && amax == t) ++csum;
    }
  }
  if (lane == 0) { wloss[warp] = lsum; wcorr[warp] = csum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    int k = 0;
    for (int w = 0; w < 32; ++w) { s += wloss[w]; k += wcorr[w]; }
    *loss = s / static_cast<float>(B);
    if (correct) *correct = k;
  }
}

// serve.serve / ModelDaemon.predict post-processing (serve.py:107-109, web/runtime.py:117-118): softmax over the class
// logits, then the maximum probability and its index. One warp per row; probability of the arg-max = 1 / sum exp(l - max).
__global__ void __launch_bounds__(256)
softmax_top1_kernel(const float* __restrict__ logits, int B, int C, float* __restrict__ conf, int* __restrict__ idx,
                    float* __restrict__ probs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + warp;
  if (b >= B) return;
  const float* lr = logits + static_cast<long long>(b) * C;
  float mx = -INFINITY;
  int amax = 0;
  for (int c = lane; c < C; c += 32) {
    const float v = lr[c];
    if (v > mx) { mx = v; amax = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, amax, o);
    if (om > mx || (om == mx && oa < amax)) { mx = om; amax = oa; }
  }
  float se = 0.f;
  for (int c = lane; c < C; c += 32) se += expf(lr[c] - mx);
  se = warp_sum(se);
  const float inv = 1.0f / se;
  if (probs != nullptr)
    for (int c = lane; c < C; c += 32) probs[static_cast<long long>(b) * C + c] = expf(lr[c] - mx) * inv;
  if (lane == 0) { conf[b] = inv; idx[b] = amax; }
}

}  // namespace

int softmax_top1(const float* logits, int B, int C, float* conf, int* idx, float* probs, cudaStream_t stream) {
  if (B <= 0) return kOk;
  if (conf == nullptr || idx == nullptr) return set_error(kErrInvalidArg, "softmax_top1: conf / idx outputs are required");
  softmax_top1_kernel<<<(B + 7) / 8, 256, 0, stream>>>(logits, B, C, conf, idx, probs);
  return check_launch("softmax_top1");
}

int head_fwd(const void* h_bf16, long long ldh, const void* w_bf16, const float* bias, int B, int D, int C,
             int round_out, float* logits, cudaStream_t stream) {
  if (D % 8 != 0) return set_error(kErrInvalidArg, "head: D=%d must be a multiple of 8", D);
  if (B <= 0) return kOk;
  ProfScope prof("head_fwd", 2.0 * B * D * C, 2.0 * (static_cast<double>(B) * D + static_cast<double>(C) * D), stream);
  head_fwd_kernel<<<dim3(B, (C + kHeadClassesPerCta - 1) / kHeadClassesPerCta), 128, D * 2, stream>>>(reinterpret_cast<const __nv_bfloat16*>(h_bf16), ldh,
                                             reinterpret_cast<const __nv_bfloat16*>(w_bf16), bias, D, C, round_out,
                                             logits);
  return check_launch("head_fwd");
}

int head_bwd(const float* dlogits, const void* h_bf16, long long ldh, const void* w_bf16, int B, int D, int C,
             void* dh_bf16, long long lddh, float* dW, float* db, cudaStream_t stream) {
  if (B <= 0) return kOk;
  ProfScope prof("head_bwd", 4.0 * B * D * C, 2.0 * (static_cast<double>(B) * D + static_cast<double>(C) * D), stream);
  if (dh_bf16 != nullptr) {
    head_bwd_dh_kernel<<<B, 256, C * 4, stream>>>(dlogits, reinterpret_cast<const __nv_bfloat16*>(w_bf16), D, C,
                                                  reinterpret_cast<__nv_bfloat16*>(dh_bf16), lddh);
    int rc = check_launch("head_bwd_dh");
    if (rc) return rc;
  }
  head_bwd_dw_kernel<<<C, 256, 0, stream>>>(dlogits, reinterpret_cast<const __nv_bfloat16*>(h_bf16), ldh, B, D, C, dW, db);
  return check_launch("head_bwd_dw");
}

int softmax_xent(const float* logits, const long long* hard, const float* soft, int B, int C, float grad_scale,
                 int round_grad, float* loss, float* dlogits, int* correct, cudaStream_t stream) {
  if ((hard == nullptr) == (soft == nullptr))
    return set_error(kErrInvalidArg, "softmax_xent: exactly one of hard/soft targets must be given");
  if (B <= 0) return set_error(kErrInvalidArg, "softmax_xent: empty batch");
  ProfScope prof("softmax_xent", 0.0, static_cast<double>(B) * C * 8, stream);
  xent_kernel<<<1, 1024, 0, stream>>>(logits, hard, soft, B, C, grad_scale, round_grad, loss, dlogits, correct);
  return check_launch("softmax_xent");
}

}  // namespace tic
