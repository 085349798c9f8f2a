// HBM-bound glue kernels around the GEMMs: patchify (im2col for the 16x16/stride-16 patch projection),
// CLS row assembly, embedding backward, bf16 shadow casts and the bias-gradient column sums.
// All global accesses are 128-bit; grids are sized in multiples of the SM count where the work is large.
#include "tic_internal.cuh"

namespace tic {
namespace {

// fp32 NCHW [B,3,S,S] -> bf16 patch matrix [B*G*G, 768], K ordered (c, py, px) so that it matches
// projection.weight.view(D, 768) (modeling_vit.py:151,166 [a2]); patch index = gy * G + gx.
__global__ void patchify_f32_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int S) {
  const int G = S / 16;
  const long long total = static_cast<long long>(B) * G * G * 96;  // 16-byte output chunks (8 bf16)
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int chunk = static_cast<int>(i % 96);
    const long long row = i / 96;
    const int k = chunk * 8;
    const int c = k >> 8, py = (k & 255) >> 4, px = k & 15;
    const int b = static_cast<int>(row / (G * G));
    const int p = static_cast<int>(row - static_cast<long long>(b) * G * G);
    const int gy = p / G, gx = p - gy * G;
    const float* src = x + ((static_cast<long long>(b) * 3 + c) * S + (gy * 16 + py)) * S + gx * 16 + px;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src));
    const float4 d = __ldg(reinterpret_cast<const float4*>(src) + 1);
    uint4 w;
    w.x = pack_bf16x2(a.x, a.y); w.y = pack_bf16x2(a.z, a.w);
    w.z = pack_bf16x2(d.x, d.y); w.w = pack_bf16x2(d.z, d.w);
    reinterpret_cast<uint4*>(out)[i] = w;
  }
}

// CutMix / MixUp (torchvision v2, ntrain.py:30-33,45-46 [a19]) fused with the patchify: sample b is paired with sample
// b-1 (roll(1, 0), _augment.py:260,327). mode 1 = MixUp: roll * (1 - lam) + x * lam, each product and the sum rounded
// to fp32 exactly as the three torch ops do (no FMA contraction); mode 2 = CutMix: pixels inside [y1,y2) x [x1,x2)
// come from the rolled batch. Writes the mixed fp32 image (optional) and / or its bf16 patch rows (optional).
__global__ void mix_patchify_kernel(const float* __restrict__ x, float* __restrict__ mixed, __nv_bfloat16* __restrict__ out,
                                    int B, int S, int mode, float lam, float one_minus_lam, int x1, int y1, int x2, int y2) {
  const int G = S / 16;
  const long long total = static_cast<long long>(B) * G * G * 96;  // 16-byte output chunks (8 bf16)
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int chunk = static_cast<int>(i % 96);
    const long long row = i / 96;
    const int k = chunk * 8;
    const int c = k >> 8, py = (k & 255) >> 4, px = k & 15;
    const int b = static_cast<int>(row / (G * G));
    const int p = static_cast<int>(row - static_cast<long long>(b) * G * G);
    const int gy = p / G, gx = p - gy * G;
    const int yy = gy * 16 + py, xx = gx * 16 + px;
    const int bp = b == 0 ? B - 1 : b - 1;
    const long long off = ((static_cast<long long>(b) * 3 + c) * S + yy) * S + xx;
    const long long offp = ((static_cast<long long>(bp) * 3 + c) * S + yy) * S + xx;
    float v[8], r[8];
    *reinterpret_cast<float4*>(v) = __ldg(reinterpret_cast<const float4*>(x + off));
    *reinterpret_cast<float4*>(v + 4) = __ldg(reinterpret_cast<const float4*>(x + off) + 1);
    if (mode != 0) {
      *reinterpret_cast<float4*>(r) = __ldg(reinterpret_cast<const float4*>(x + offp));
      *reinterpret_cast<float4*>(r + 4) = __ldg(reinterpret_cast<const float4*>(x + offp) + 1);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (mode == 1) v[j] = __fadd_rn(__fmul_rn(r[j], one_minus_lam), __fmul_rn(v[j], lam));
      else if (mode == 2 && yy >= y1 && yy < y2 && xx + j >= x1 && xx + j < x2) v[j] = r[j];
    }
    if (mixed != nullptr) {
      *reinterpret_cast<float4*>(mixed + off) = *reinterpret_cast<const float4*>(v);
      *(reinterpret_cast<float4*>(mixed + off) + 1) = *reinterpret_cast<const float4*>(v + 4);
    }
    if (out != nullptr) {
      uint4 w;
      w.x = pack_bf16x2(v[0], v[1]); w.y = pack_bf16x2(v[2], v[3]);
      w.z = pack_bf16x2(v[4], v[5]); w.w = pack_bf16x2(v[6], v[7]);
      reinterpret_cast<uint4*>(out)[i] = w;
    }
  }
}

// Soft targets of CutMix / MixUp: onehot(y[b-1]) * (1 - lam) + onehot(y[b]) * lam (_augment.py:214-219).
__global__ void mix_targets_kernel(const long long* __restrict__ y, int B, int C, float lam, float one_minus_lam,
                                   float* __restrict__ soft) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i - b * C;
  const int bp = b == 0 ? B - 1 : b - 1;
  const float prev = y[bp] == c ? 1.0f : 0.0f, cur = y[b] == c ? 1.0f : 0.0f;
  soft[i] = __fadd_rn(__fmul_rn(prev, one_minus_lam), __fmul_rn(cur, lam));
}

// x[b, 0, :] = cls + pos[0]   (modeling_vit.py:117-124 [a3]); the patch rows are written by the GEMM epilogue.
__global__ void cls_rows_kernel(const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ x,
                                int B, int N, int D) {
  const int b = blockIdx.x;
  float4* dst = reinterpret_cast<float4*>(x + static_cast<long long>(b) * N * D);
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(cls) + i), p = __ldg(reinterpret_cast<const float4*>(pos) + i);
    dst[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
  }
}

// Embedding backward: dpos[n,:] = sum_b dx[b,n,:]; dcls = dpos[0]; dpatch[b*P + p,:] = bf16(dx[b,1+p,:]).
__global__ void embed_bwd_kernel(const float* __restrict__ dx, int B, int N, int D, float* __restrict__ dpos,
                                 float* __restrict__ dcls, __nv_bfloat16* __restrict__ dpatch) {
  const int n = blockIdx.x;
  const int P = N - 1;
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < B; ++b) {
      const float4 v = reinterpret_cast<const float4*>(dx + (static_cast<long long>(b) * N + n) * D)[i];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      if (n > 0 && dpatch != nullptr) {
        uint2 w;
        w.x = pack_bf16x2(v.x, v.y);
        w.y = pack_bf16x2(v.z, v.w);
        reinterpret_cast<uint2*>(dpatch + (static_cast<long long>(b) * P + (n - 1)) * D)[i] = w;
      }
    }
    float4* dp = reinterpret_cast<float4*>(dpos + static_cast<long long>(n) * D) + i;
    float4 o = *dp;
    o.x += s.x; o.y += s.y; o.z += s.z; o.w += s.w;
    *dp = o;
    if (n == 0) {
      float4* dc = reinterpret_cast<float4*>(dcls) + i;
      float4 c = *dc;
      c.x += s.x; c.y += s.y; c.z += s.z; c.w += s.w;
      *dc = c;
    }
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n8) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(src)[2 * i], b = reinterpret_cast<const float4*>(src)[2 * i + 1];
    uint4 w;
    w.x = pack_bf16x2(a.x, a.y); w.y = pack_bf16x2(a.z, a.w);
    w.z = pack_bf16x2(b.x, b.y); w.w = pack_bf16x2(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = w;
  }
}

__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, long long n8) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 w = reinterpret_cast<const uint4*>(src)[i];
    reinterpret_cast<float4*>(dst)[2 * i] = make_float4(bf16_lo(w.x), bf16_hi(w.x), bf16_lo(w.y), bf16_hi(w.y));
    reinterpret_cast<float4*>(dst)[2 * i + 1] = make_float4(bf16_lo(w.z), bf16_hi(w.z), bf16_lo(w.w), bf16_hi(w.w));
  }
}

// Bias gradient: db[n] += sum_m dY[m, n] for bf16 dY. A warp reads 256 consecutive columns of one row
// (16 B per lane); the CTA's 8 warps take 8 rows per step over its row chunk.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ dy, long long ld, int rows, int cols, int rows_per_cta,
                   float* __restrict__ out) {
  __shared__ float red[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(rows, r0 + rows_per_cta);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (col < cols) {
    for (int r = r0 + warp; r < r1; r += 8) {
      const uint4 w = __ldg(reinterpret_cast<const uint4*>(dy + static_cast<long long>(r) * ld + col));
      acc[0] += bf16_lo(w.x); acc[1] += bf16_hi(w.x); acc[2] += bf16_lo(w.y); acc[3] += bf16_hi(w.y);
      acc[4] += bf16_lo(w.z); acc[5] += bf16_hi(w.z); acc[6] += bf16_lo(w.w); acc[7] += bf16_hi(w.w);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  const int c = threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w][c];
  if (blockIdx.x * 256 + c < cols) atomicAdd(out + blockIdx.x * 256 + c, s);
}

inline int ew_grid(long long work_items, int threads) {
  long long g = (work_items + threads - 1) / threads;
  const long long cap = 148LL * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

int patchify_f32(const float* x, void* out_bf16, int B, int S, cudaStream_t stream) {
  if (S % 16 != 0) return set_error(kErrInvalidArg, "patchify: image size %d is not a multiple of 16", S);
  const int G = S / 16;
  const long long total = static_cast<long long>(B) * G * G * 96;
  ProfScope prof("patchify_f32", 0.0, static_cast<double>(total) * 48, stream);
  patchify_f32_kernel<<<ew_grid(total, 256), 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out_bf16), B, S);
  return check_launch("patchify_f32");
}

int mix_patchify_f32(const float* x, float* mixed, void* out_bf16, int B, int S, int mode, float lam, float one_minus_lam,
                     int x1, int y1, int x2, int y2, cudaStream_t stream) {
  if (B <= 0) return kOk;
  if (S <= 0 || S % 16) return set_error(kErrInvalidArg, "mix_patchify: bad image size %d", S);
  if (mode < 0 || mode > 2) return set_error(kErrInvalidArg, "mix_patchify: mode %d (0 none, 1 mixup, 2 cutmix)", mode);
  if (mixed == x) return set_error(kErrInvalidArg, "mix_patchify: the mixed image cannot alias the input (rolled reads)");
  const long long total = static_cast<long long>(B) * (S / 16) * (S / 16) * 96;
  long long g = (total + 255) / 256;
  if (g > 148LL * 16) g = 148LL * 16;
  ProfScope prof("mix_patchify_f32", 0.0, static_cast<double>(B) * 3 * S * S * (mode ? 8.0 : 4.0) +
                     static_cast<double>(B) * 3 * S * S * ((mixed ? 4.0 : 0.0) + (out_bf16 ? 2.0 : 0.0)), stream);
  mix_patchify_kernel<<<static_cast<int>(g), 256, 0, stream>>>(x, mixed, reinterpret_cast<__nv_bfloat16*>(out_bf16), B, S, mode,
                                                              lam, one_minus_lam, x1, y1, x2, y2);
  return check_launch("mix_patchify_f32");
}

int mix_targets(const long long* y, int B, int C, float lam, float one_minus_lam, float* soft, cudaStream_t stream) {
  if (B <= 0 || C <= 0) return kOk;
  ProfScope prof("mix_targets", 0.0, static_cast<double>(B) * C * 4, stream);
  mix_targets_kernel<<<(B * C + 255) / 256, 256, 0, stream>>>(y, B, C, lam, one_minus_lam, soft);
  return check_launch("mix_targets");
}

int cls_rows(const float* cls, const float* pos, float* x, int B, int N, int D, cudaStream_t stream) {
  cls_rows_kernel<<<B, 256, 0, stream>>>(cls, pos, x, B, N, D);
  return check_launch("cls_rows");
}

int embed_bwd(const float* dx, int B, int N, int D, float* dpos, float* dcls, void* dpatch_bf16, cudaStream_t stream) {
  ProfScope prof("embed_bwd", 0.0, static_cast<double>(B) * N * D * 6, stream);
  embed_bwd_kernel<<<N, 256, 0, stream>>>(dx, B, N, D, dpos, dcls, reinterpret_cast<__nv_bfloat16*>(dpatch_bf16));
  return check_launch("embed_bwd");
}

int cast_f32_to_bf16(const float* src, void* dst, long long n, cudaStream_t stream) {
  if (n % 8 != 0) return set_error(kErrInvalidArg, "cast: n=%lld must be a multiple of 8", n);
  if (n == 0) return kOk;
  ProfScope prof("cast_f32_to_bf16", 0.0, static_cast<double>(n) * 6, stream);
  cast_f32_bf16_kernel<<<ew_grid(n / 8, 256), 256, 0, stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n / 8);
  return check_launch("cast_f32_to_bf16");
}

int cast_bf16_to_f32(const void* src, float* dst, long long n, cudaStream_t stream) {
  if (n % 8 != 0) return set_error(kErrInvalidArg, "cast: n=%lld must be a multiple of 8", n);
  if (n == 0) return kOk;
  cast_bf16_f32_kernel<<<ew_grid(n / 8, 256), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(src), dst, n / 8);
  return check_launch("cast_bf16_to_f32");
}

int colsum_bf16(const void* dy, long long ld, int rows, int cols, float* out, cudaStream_t stream) {
  if (cols % 8 != 0) return set_error(kErrInvalidArg, "colsum: cols=%d must be a multiple of 8", cols);
  ProfScope prof("colsum_bf16", 0.0, static_cast<double>(rows) * cols * 2, stream);
  const int gx = (cols + 255) / 256;
  int gy = (148 * 4 + gx - 1) / gx;
  int rows_per_cta = (rows + gy - 1) / gy;
  rows_per_cta = ((rows_per_cta + 7) / 8) * 8;
  gy = (rows + rows_per_cta - 1) / rows_per_cta;
  colsum_bf16_kernel<<<dim3(gx, gy), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(dy), ld, rows, cols,
                                                       rows_per_cta, out);
  return check_launch("colsum_bf16");
}

}  // namespace tic
