// Warp-per-row LayerNorm forward and backward with 128-bit HBM access.
// Replaces nn.LayerNorm at modeling_vit.py:325-326,333,340 (layernorm_before/after) and :416,455
// (final layernorm) [a4, a11]. eps comes from ViTConfig.layer_norm_eps (1e-12) and is NOT "fixed".
// Statistics and the affine transform are fp32 (autocast keeps LayerNorm in fp32);
// the output is written as bf16 (the operand dtype of the GEMM that consumes it) and/or fp32.
#include "tic_internal.cuh"

#include <cstdlib>

namespace tic {
namespace {

constexpr int LN_WARPS = 8;
constexpr int LN_CLC_STAGES = 4;  // ring of cluster-launch-control responses (backward)

// One warp per row; a lane owns float4 chunks lane, lane + 32, ... (VEC chunks, D = VEC * 128).
template <int VEC>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
              const float* __restrict__ beta, float eps, int rows, __nv_bfloat16* __restrict__ y_bf16,
              long long ldy, float* __restrict__ y_f32, long long ldyf, float* __restrict__ mean_out,
              float* __restrict__ rstd_out) {
  pdl_launch_dependents();
  pdl_wait();  // every thread, before the early exit below: x comes from the preceding kernel
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * LN_WARPS + warp;
  if (row >= rows) return;
  constexpr int D = VEC * 128;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * ldx);
  float4 v[VEC];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float var = warp_sum(q) * (1.0f / D);
  const float rstd = rsqrtf(var + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float4 g = __ldg(g4 + lane + 32 * i), b = __ldg(b4 + lane + 32 * i);
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + b.x;
    o.y = (v[i].y - mean) * rstd * g.y + b.y;
    o.z = (v[i].z - mean) * rstd * g.z + b.z;
    o.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (y_bf16) {
      uint2 w;
      w.x = pack_bf16x2(o.x, o.y);
      w.y = pack_bf16x2(o.z, o.w);
      reinterpret_cast<uint2*>(y_bf16 + static_cast<long long>(row) * ldy)[lane + 32 * i] = w;
    }
    if (y_f32) reinterpret_cast<float4*>(y_f32 + static_cast<long long>(row) * ldyf)[lane + 32 * i] = o;
  }
}

// Backward. Per row:  g = dy * gamma;  dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) [+ dres]
// dgamma += dy * xhat and dbeta += dy are accumulated over the rows a warp visits in a per-warp slice of
// shared memory (keeps the register count low enough for 3 CTAs / SM), reduced across the CTA's warps at
// the end, then one atomicAdd per column per CTA.
template <int VEC>
// One CTA (8 warps) per SM, whatever registers it takes (220 at D = 1024): asking for two CTAs per SM capped the kernel at
// 128 registers and ptxas spilled 200 bytes of the row held in registers -- the kernel keeps 10 KB of loads in flight per
// warp, so eight warps already cover the HBM latency. Same-box A/B at [50432, 1024]: 0.1755 -> 0.154 ms (4.7 -> 5.37 TB/s,
// 0.82 of the measured copy bandwidth).
__global__ void __launch_bounds__(LN_WARPS * 32, 1)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy, const float* __restrict__ x, long long ldx,
              const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const float* __restrict__ gamma,
              const float* dres, long long lddres, int rows, float* dx, long long lddx,
              __nv_bfloat16* __restrict__ dx_bf16, long long lddxb, float* __restrict__ dgamma,
              float* __restrict__ dbeta, float* __restrict__ dxsum, int dynamic, int chunk_rows) {
  constexpr int D = VEC * 128;
  pdl_launch_dependents();
  extern __shared__ float4 ln_smem[];
  float4* sg = ln_smem;                           // [LN_WARPS][VEC * 32] dgamma partials
  float4* sb = ln_smem + LN_WARPS * VEC * 32;     // [LN_WARPS][VEC * 32] dbeta partials
  float4* sc = ln_smem + 2 * LN_WARPS * VEC * 32; // [LN_WARPS][VEC * 32] column sums of the emitted bf16 dx
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4* my_g = sg + warp * VEC * 32;
  float4* my_b = sb + warp * VEC * 32;
  float4* my_c = sc + warp * VEC * 32;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    my_g[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
    my_b[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
    my_c[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  pdl_wait();  // the accumulators above needed nothing from the preceding kernel; dy / x / dres below do
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  // Work is handed out in chunks of chunk_rows rows. Static: chunks blockIdx.x, blockIdx.x + gridDim.x, ...
  // Dynamic: the grid holds one CTA per chunk and a running CTA takes over CTAs that have not started (cluster launch
  // control), so the per-CTA accumulators stay few while the rows spread over whatever SMs are free (a NCCL kernel
  // overlapping the backward may hold some). Warp 0 requests chunk n+1 before working on chunk n; the warps run
  // independently and meet only through the response ring.
  __shared__ __align__(16) uint4 clc_resp[LN_CLC_STAGES];
  __shared__ __align__(8) uint64_t clc_full[LN_CLC_STAGES], clc_empty[LN_CLC_STAGES];
  const int num_chunks = (rows + chunk_rows - 1) / chunk_rows;
  if (dynamic) {
    if (threadIdx.x == 0) {
      for (int i = 0; i < LN_CLC_STAGES; ++i) {
        mbar_init(&clc_full[i], 1);
        mbar_init(&clc_empty[i], LN_WARPS);
      }
      fence_mbar_init();
    }
    __syncthreads();
  }
  int qi = 0;
  for (int chunk = blockIdx.x; chunk >= 0 && chunk < num_chunks;) {
    if (dynamic && threadIdx.x == 0) {  // request the chunk after this one
      const int slot = qi % LN_CLC_STAGES;
      mbar_wait(&clc_empty[slot], ((qi / LN_CLC_STAGES) & 1) ^ 1);
      mbar_arrive_expect_tx(&clc_full[slot], 16);
      clc_try_cancel(&clc_resp[slot], &clc_full[slot]);
    }
    const int row_end = min(rows, (chunk + 1) * chunk_rows);
    for (int row = chunk * chunk_rows + warp; row < row_end; row += LN_WARPS) {
    const float mean = mean_in[row], rstd = rstd_in[row];
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * ldx);
    const uint2* dyr = reinterpret_cast<const uint2*>(dy + static_cast<long long>(row) * lddy);
    const float4* rr = dres ? reinterpret_cast<const float4*>(dres + static_cast<long long>(row) * lddres) : nullptr;
    // All of the row's HBM reads (x, dy and the residual gradient) are issued up front: 10 KB in flight per warp
    // instead of two dependent round trips. xhat and dy * gamma are recomputed in the second pass rather than kept.
    float4 xv[VEC], rv[VEC];
    uint2 dv[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      xv[i] = xr[lane + 32 * i];
      dv[i] = dyr[lane + 32 * i];
    }
    if (rr != nullptr) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) rv[i] = rr[lane + 32 * i];
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float4 gm = __ldg(g4 + lane + 32 * i);
      const float d0 = bf16_lo(dv[i].x), d1 = bf16_hi(dv[i].x), d2 = bf16_lo(dv[i].y), d3 = bf16_hi(dv[i].y);
      const float h0 = (xv[i].x - mean) * rstd, h1 = (xv[i].y - mean) * rstd;
      const float h2 = (xv[i].z - mean) * rstd, h3 = (xv[i].w - mean) * rstd;
      const float g0 = d0 * gm.x, g1 = d1 * gm.y, g2 = d2 * gm.z, g3 = d3 * gm.w;
      s1 += (g0 + g1) + (g2 + g3);
      s2 += (g0 * h0 + g1 * h1) + (g2 * h2 + g3 * h3);
      float4 ag = my_g[lane + 32 * i], ab = my_b[lane + 32 * i];
      ag.x += d0 * h0; ag.y += d1 * h1; ag.z += d2 * h2; ag.w += d3 * h3;
      ab.x += d0; ab.y += d1; ab.z += d2; ab.w += d3;
      my_g[lane + 32 * i] = ag;
      my_b[lane + 32 * i] = ab;
    }
    const float c1 = warp_sum(s1) * (1.0f / D), c2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float4 gm = __ldg(g4 + lane + 32 * i);
      const float d0 = bf16_lo(dv[i].x), d1 = bf16_hi(dv[i].x), d2 = bf16_lo(dv[i].y), d3 = bf16_hi(dv[i].y);
      const float h0 = (xv[i].x - mean) * rstd, h1 = (xv[i].y - mean) * rstd;
      const float h2 = (xv[i].z - mean) * rstd, h3 = (xv[i].w - mean) * rstd;
      float4 o;
      o.x = rstd * (d0 * gm.x - c1 - h0 * c2) + rv[i].x;
      o.y = rstd * (d1 * gm.y - c1 - h1 * c2) + rv[i].y;
      o.z = rstd * (d2 * gm.z - c1 - h2 * c2) + rv[i].z;
      o.w = rstd * (d3 * gm.w - c1 - h3 * c2) + rv[i].w;
      reinterpret_cast<float4*>(dx + static_cast<long long>(row) * lddx)[lane + 32 * i] = o;
      if (dx_bf16) {
        uint2 w;
        w.x = pack_bf16x2(o.x, o.y);
        w.y = pack_bf16x2(o.z, o.w);
        reinterpret_cast<uint2*>(dx_bf16 + static_cast<long long>(row) * lddxb)[lane + 32 * i] = w;
        if (dxsum != nullptr) {  // bias gradient of the Linear whose output gradient this dx is (sum of the bf16 values)
          float4 ac = my_c[lane + 32 * i];
          ac.x += bf16_lo(w.x); ac.y += bf16_hi(w.x); ac.z += bf16_lo(w.y); ac.w += bf16_hi(w.y);
          my_c[lane + 32 * i] = ac;
        }
      }
    }
    }
    if (dynamic) {
      const int slot = qi % LN_CLC_STAGES;
      mbar_wait(&clc_full[slot], (qi / LN_CLC_STAGES) & 1);
      chunk = clc_decode(&clc_resp[slot]);
      ++qi;
      __syncwarp();
      if (lane == 0) {
        fence_proxy_async();
        mbar_arrive(&clc_empty[slot]);
      }
    } else {
      chunk += gridDim.x;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < VEC * 32; c += LN_WARPS * 32) {
    float4 s = sg[c], t = sb[c], u = sc[c];
#pragma unroll
    for (int w = 1; w < LN_WARPS; ++w) {
      const float4 a = sg[w * VEC * 32 + c], b = sb[w * VEC * 32 + c], e = sc[w * VEC * 32 + c];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      t.x += b.x; t.y += b.y; t.z += b.z; t.w += b.w;
      u.x += e.x; u.y += e.y; u.z += e.z; u.w += e.w;
    }
    if (dxsum != nullptr) {
      atomicAdd(dxsum + 4 * c + 0, u.x); atomicAdd(dxsum + 4 * c + 1, u.y);
      atomicAdd(dxsum + 4 * c + 2, u.z); atomicAdd(dxsum + 4 * c + 3, u.w);
    }
    if (dgamma != nullptr) {
      atomicAdd(dgamma + 4 * c + 0, s.x); atomicAdd(dgamma + 4 * c + 1, s.y);
      atomicAdd(dgamma + 4 * c + 2, s.z); atomicAdd(dgamma + 4 * c + 3, s.w);
    }
    if (dbeta != nullptr) {
      atomicAdd(dbeta + 4 * c + 0, t.x); atomicAdd(dbeta + 4 * c + 1, t.y);
      atomicAdd(dbeta + 4 * c + 2, t.z); atomicAdd(dbeta + 4 * c + 3, t.w);
    }
  }
}

}  // namespace

int layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, float eps, int rows, int D,
                  void* y_bf16, long long ldy, float* y_f32, long long ldyf, float* mean, float* rstd,
                  cudaStream_t stream) {
  if (rows <= 0) return kOk;
  if (D % 128 != 0 || D > 1024 * 2) return set_error(kErrUnsupported, "layernorm: D=%d must be a multiple of 128", D);
  const int grid = (rows + LN_WARPS - 1) / LN_WARPS;
  ProfScope prof("layernorm_fwd", 0.0, static_cast<double>(rows) * D * (4 + (y_bf16 ? 2 : 0) + (y_f32 ? 4 : 0)), stream);
  auto* yb = reinterpret_cast<__nv_bfloat16*>(y_bf16);
#define TIC_LN_FWD(V)                                                                                              \
  case V:                                                                                                          \
    launch_pdl(ln_fwd_kernel<V>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, x, ldx, gamma, beta, eps, rows, yb, ldy, \
               y_f32, ldyf, mean, rstd);                                                                           \
    break;
  switch (D / 128) {
    TIC_LN_FWD(1) TIC_LN_FWD(2) TIC_LN_FWD(3) TIC_LN_FWD(4) TIC_LN_FWD(6) TIC_LN_FWD(8) TIC_LN_FWD(10) TIC_LN_FWD(12)
    TIC_LN_FWD(16)
    default:
      return set_error(kErrUnsupported, "layernorm: unsupported D=%d", D);
  }
#undef TIC_LN_FWD
  return check_launch("layernorm_fwd");
}

int layernorm_bwd(const void* dy_bf16, long long lddy, const float* x, long long ldx, const float* mean,
                  const float* rstd, const float* gamma, const float* dres, long long lddres, int rows, int D,
                  float* dx, long long lddx, void* dx_bf16, long long lddxb, float* dgamma, float* dbeta,
                  float* dxsum, cudaStream_t stream) {
  if (rows <= 0) return kOk;
  if (D % 128 != 0) return set_error(kErrUnsupported, "layernorm: D=%d must be a multiple of 128", D);
  if (dxsum != nullptr && dx_bf16 == nullptr)
    return set_error(kErrInvalidArg, "layernorm_bwd: dxsum needs the bf16 dx output");
  static const bool force_static = std::getenv("TIC_LN_STATIC") != nullptr;  // development A/B
  const int dynamic = force_static ? 0 : 1;
  // dynamic: 8 rows per warp between two hand-outs; a launch too small to fill the machine that way (the B CLS rows of
  // the final LayerNorm / the CLS-only last layer) gets one row per warp so that the rows still spread over the SMs
  const int chunk_rows = (dynamic && rows > 8 * LN_WARPS * 2 * device_sm_count()) ? 8 * LN_WARPS : LN_WARPS;
  int grid = (rows + chunk_rows - 1) / chunk_rows;  // dynamic: one CTA per chunk; running CTAs take over the rest
  const int max_grid = device_sm_count() * 2;
  if (!dynamic && grid > max_grid) grid = max_grid;
  ProfScope prof("layernorm_bwd", 0.0, static_cast<double>(rows) * D * (2 + 4 + (dres ? 4 : 0) + 4 + (dx_bf16 ? 2 : 0)), stream);
  auto* dyb = reinterpret_cast<const __nv_bfloat16*>(dy_bf16);
  auto* dxb = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
#define TIC_LN_BWD(V)                                                                                             \
  case V: {                                                                                                       \
    const int smem = 3 * LN_WARPS * V * 32 * 16;                                                                  \
    if (smem > 48 * 1024) {                                                                                       \
      if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(ln_bwd_kernel<V>), smem, "layernorm_bwd")) return rc; \
    }                                                                                                             \
    launch_pdl(ln_bwd_kernel<V>, dim3(grid), dim3(LN_WARPS * 32), smem, stream, dyb, lddy, x, ldx, mean, rstd, gamma, \
               dres, lddres, rows, dx, lddx, dxb, lddxb, dgamma, dbeta, dxsum, dynamic, chunk_rows);               \
    break;                                                                                                        \
  }
  switch (D / 128) {
    TIC_LN_BWD(1) TIC_LN_BWD(2) TIC_LN_BWD(3) TIC_LN_BWD(4) TIC_LN_BWD(6) TIC_LN_BWD(8)
    default:
      return set_error(kErrUnsupported, "layernorm_bwd: unsupported D=%d", D);
  }
#undef TIC_LN_BWD
  return check_launch("layernorm_bwd");
}

}  // namespace tic
