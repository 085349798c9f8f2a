// Fused AdamW over the flat fp32 parameter arena: one launch updates p, m, v and refreshes the bf16
// shadow copy the tensor-core GEMMs read. Replaces torch.optim.AdamW.step as configured by
// ntrain.py:39-41 [a16] / finetune.py:314 (one parameter group, weight decay on every tensor, default
// betas (0.9, 0.999), eps 1e-8). Update order follows torch/optim/adamw.py (_single_tensor_adamw):
//   p *= 1 - lr*wd;  m = lerp(m, g, 1-b1);  v = b2*v + (1-b2)*g*g;
//   p -= (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
// 28 B/param of algorithmic HBM traffic (+2 B for the shadow), 128-bit accesses.
#include "tic_internal.cuh"

namespace tic {
namespace {

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             __nv_bfloat16* __restrict__ shadow, long long n4, float lr, float beta1, float beta2, float eps,
             float weight_decay, float bias_corr1, float bias_corr2_sqrt, float grad_scale) {
  const float decay = 1.0f - lr * weight_decay;
  const float step_size = lr / bias_corr1;
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pp = &pv.x; float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = gp[j] * grad_scale;
      pp[j] *= decay;
      mp[j] = mp[j] + (gj - mp[j]) * omb1;
      vp[j] = vp[j] * beta2 + gj * gj * omb2;
      const float denom = sqrtf(vp[j]) / bias_corr2_sqrt + eps;
      pp[j] -= step_size * (mp[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (shadow) {
      uint2 w;
      w.x = pack_bf16x2(pv.x, pv.y);
      w.y = pack_bf16x2(pv.z, pv.w);
      reinterpret_cast<uint2*>(shadow)[i] = w;
    }
  }
}

}  // namespace

int adamw_step(float* p, const float* g, float* m, float* v, void* shadow_bf16, long long n, float lr, float beta1,
               float beta2, float eps, float weight_decay, int step, float grad_scale, cudaStream_t stream) {
  if (n % 4 != 0) return set_error(kErrInvalidArg, "adamw: n=%lld must be a multiple of 4", n);
  if (step < 1) return set_error(kErrInvalidArg, "adamw: step must be >= 1");
  if (n == 0) return kOk;
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  ProfScope prof("adamw", 0.0, static_cast<double>(n) * (shadow_bf16 ? 30 : 28), stream);
  long long grid = (n / 4 + 255) / 256;
  if (grid > 148LL * 8) grid = 148LL * 8;
  adamw_kernel<<<static_cast<int>(grid), 256, 0, stream>>>(p, g, m, v, reinterpret_cast<__nv_bfloat16*>(shadow_bf16),
                                                          n / 4, lr, beta1, beta2, eps, weight_decay,
                                                          static_cast<float>(bc1), static_cast<float>(sqrt(bc2)),
                                                          grad_scale);
  return check_launch("adamw_step");
}

}  // namespace tic
