// Fused multi-head attention forward and backward for ViT sequence lengths (197 / 577 tokens, d = 64).
// Replaces F.scaled_dot_product_attention as called by ViTSelfAttention.forward
// (modeling_vit.py:232-246 via transformers/integrations/sdpa_attention.py:92-103 [a6]): non-causal,
// no mask, dropout 0, scale 1/sqrt(64), fp32 softmax statistics, bf16 probabilities for P.V.
//
// Round-1 implementation: flash-style tiles (64 queries x 64 keys per step), warp-level tensor-core MMA
// (mma.sync m16n8k16 bf16, fp32 accumulate), cp.async double-buffered K/V (or Q/dO) tiles in
// XOR-swizzled shared memory, online softmax in registers. q/k/v are read in place from the fused QKV
// GEMM output [B*N, 3*D]; the context is written token-major [B*N, D] (the layout the output projection
// GEMM consumes), so no head transposes ever touch HBM.
//   forward : grid (ceil(N/64), H, B); saves logsumexp per (b, h, n) for the backward pass
//   backward: delta = rowsum(dO * O); dQ kernel (one CTA per 64 queries, loops over keys);
//             dK/dV kernel (one CTA per 64 keys, loops over queries). No atomics, deterministic.
#include "tic_internal.cuh"

namespace tic {
namespace {

constexpr int HD = 64;          // head dim
constexpr int TILE = 64;        // rows per tile
constexpr int TILE_BYTES = TILE * HD * 2;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

TIC_DEVINL uint32_t tile_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

TIC_DEVINL void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
TIC_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
TIC_DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

TIC_DEVINL void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
TIC_DEVINL void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
TIC_DEVINL void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Copy a 64 x 64 bf16 tile (rows row0.., zero-filled past `nrows`) from a token-major matrix into swizzled smem.
TIC_DEVINL void load_tile(uint32_t smem_base, const __nv_bfloat16* gbase, long long ld, int row0, int nrows, int tid,
                          int nthreads) {
  for (int i = tid; i < TILE * 8; i += nthreads) {
    const int r = i >> 3, c = i & 7;
    const bool ok = row0 + r < nrows;
    const __nv_bfloat16* src = gbase + static_cast<long long>(ok ? row0 + r : 0) * ld + c * 8;
    cp_async16(smem_base + tile_off(r, c), src, ok);
  }
}

// A-operand fragments (16 rows x 64 k) of this warp's rows from a swizzled tile: a[kk][0..3].
TIC_DEVINL void load_a_frags(uint32_t tile, int warp_row0, int lane, uint32_t (&a)[4][4]) {
  const int mi = lane >> 3, r = lane & 7;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
    ldsm_x4(tile + tile_off(warp_row0 + (mi & 1) * 8 + r, 2 * kk + (mi >> 1)), a[kk][0], a[kk][1], a[kk][2], a[kk][3]);
}

// acc[j][*] (16 x 64, 8 n-tiles) = A(16 x 64) * T^T where T is a swizzled [64 rows (n)][64 (k)] tile.
// Only the first `n_valid` rows of T hold real data: 16-row groups past it are skipped (warp-uniform).
template <bool FULL>
TIC_DEVINL void mma_a_tT_impl(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t tile, int lane, int n_valid) {
  const int mi = lane >> 3, r = lane & 7;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      if (!FULL && 16 * jp >= n_valid) continue;
      uint32_t b0, b1, b2, b3;
      ldsm_x4(tile + tile_off(8 * (2 * jp + (mi >> 1)) + r, 2 * kk + (mi & 1)), b0, b1, b2, b3);
      mma16816(acc[2 * jp], a[kk], b0, b1);
      mma16816(acc[2 * jp + 1], a[kk], b2, b3);
    }
  }
}

TIC_DEVINL void mma_a_tT(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t tile, int lane, int n_valid) {
  if (n_valid >= TILE) mma_a_tT_impl<true>(acc, a, tile, lane, n_valid);   // branch-free fast path for full tiles
  else mma_a_tT_impl<false>(acc, a, tile, lane, n_valid);
}

// acc[j][*] (16 x 64) += P(16 x 64, given as fp32 C-fragments, rounded to bf16) * T, T = swizzled [64 (k)][64 (n)].
// Only the first `k_valid` rows of T (and columns of P) are non-zero: later 16-wide k-steps are skipped.
template <bool FULL>
TIC_DEVINL void mma_p_t_impl(float (&acc)[8][4], const float (&p)[8][4], uint32_t tile, int lane, int k_valid) {
  const int mi = lane >> 3, r = lane & 7;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    if (!FULL && 16 * kk >= k_valid) continue;
    uint32_t a[4];
    a[0] = pack_bf16x2(p[2 * kk][0], p[2 * kk][1]);
    a[1] = pack_bf16x2(p[2 * kk][2], p[2 * kk][3]);
    a[2] = pack_bf16x2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    a[3] = pack_bf16x2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(tile + tile_off(16 * kk + (mi & 1) * 8 + r, 2 * jp + (mi >> 1)), b0, b1, b2, b3);
      mma16816(acc[2 * jp], a, b0, b1);
      mma16816(acc[2 * jp + 1], a, b2, b3);
    }
  }
}

TIC_DEVINL void mma_p_t(float (&acc)[8][4], const float (&p)[8][4], uint32_t tile, int lane, int k_valid) {
  if (k_valid >= TILE) mma_p_t_impl<true>(acc, p, tile, lane, k_valid);
  else mma_p_t_impl<false>(acc, p, tile, lane, k_valid);
}

// Stage a warp's 16 x 64 fp32 C-fragments as bf16 into its rows of a swizzled tile, then store them
// to global with 16-byte accesses (rows >= nrows are skipped).
TIC_DEVINL void store_c_tile(uint32_t tile, uint8_t* tile_ptr, const float (&c)[8][4], float s0, float s1,
                             int warp_row0, int lane, __nv_bfloat16* gbase, long long ld, int row0, int nrows) {
  const int g = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    // element (row g, cols 8j + 2t, +1): chunk j, byte offset 4t inside the 16-byte chunk
    *reinterpret_cast<uint32_t*>(tile_ptr + tile_off(warp_row0 + g, j) + 4 * t) = pack_bf16x2(c[j][0] * s0, c[j][1] * s0);
    *reinterpret_cast<uint32_t*>(tile_ptr + tile_off(warp_row0 + g + 8, j) + 4 * t) =
        pack_bf16x2(c[j][2] * s1, c[j][3] * s1);
  }
  __syncwarp();
  (void)tile;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = lane + 32 * i;  // 16 rows x 8 chunks
    const int r = idx >> 3, ch = idx & 7;
    const int grow = row0 + warp_row0 + r;
    if (grow < nrows) {
      const uint4 v = *reinterpret_cast<const uint4*>(tile_ptr + tile_off(warp_row0 + r, ch));
      *reinterpret_cast<uint4*>(gbase + static_cast<long long>(grow) * ld + ch * 8) = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(128)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                const __nv_bfloat16* __restrict__ v, long long ld, __nv_bfloat16* __restrict__ o, long long ldo,
                float* __restrict__ lse, int N, int H, float scale) {
  __shared__ __align__(1024) uint8_t smem[5 * TILE_BYTES];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * TILE, h = blockIdx.y, b = blockIdx.z;
  const long long tok0 = static_cast<long long>(b) * N;
  const __nv_bfloat16* qg = q + tok0 * ld + h * HD;
  const __nv_bfloat16* kg = k + tok0 * ld + h * HD;
  const __nv_bfloat16* vg = v + tok0 * ld + h * HD;
  const uint32_t sQ = smem_u32(smem), sK = sQ + TILE_BYTES, sV = sQ + 3 * TILE_BYTES;

  load_tile(sQ, qg, ld, q0, N, tid, 128);
  load_tile(sK, kg, ld, 0, N, tid, 128);
  load_tile(sV, vg, ld, 0, N, tid, 128);
  cp_async_commit();

  const int nkv = (N + TILE - 1) / TILE;
  const float c2 = scale * LOG2E;
  const bool warp_active = q0 + warp * 16 < N;
  uint32_t qa[4][4];
  float oacc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) oacc[j][0] = oacc[j][1] = oacc[j][2] = oacc[j][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int t = lane & 3;

  for (int kb = 0; kb < nkv; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nkv) {
      load_tile(sK + (buf ^ 1) * TILE_BYTES, kg, ld, (kb + 1) * TILE, N, tid, 128);
      load_tile(sV + (buf ^ 1) * TILE_BYTES, vg, ld, (kb + 1) * TILE, N, tid, 128);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kb == 0) load_a_frags(sQ, warp * 16, lane, qa);
    const int kv_valid = min(TILE, N - kb * TILE);
    if (!warp_active) {  // all 16 query rows of this warp are padding: only take part in the tile pipeline
      __syncthreads();
      continue;
    }

    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
    mma_a_tT(s, qa, sK + buf * TILE_BYTES, lane, kv_valid);

    const int kbase = kb * TILE;
    const bool tail = kbase + TILE > N;
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float x = s[j][e] * c2;
        if (tail && kbase + 8 * j + 2 * t + (e & 1) >= N) x = -INFINITY;
        s[j][e] = x;
      }
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float a0 = exp2f(m0 - mx0), a1 = exp2f(m1 - mx1);  // first block: exp2(-inf) = 0
    m0 = mx0; m1 = mx1;
    float r0 = 0.f, r1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = exp2f(s[j][0] - m0); s[j][1] = exp2f(s[j][1] - m0);
      s[j][2] = exp2f(s[j][2] - m1); s[j][3] = exp2f(s[j][3] - m1);
      r0 += s[j][0] + s[j][1];
      r1 += s[j][2] + s[j][3];
      oacc[j][0] *= a0; oacc[j][1] *= a0; oacc[j][2] *= a1; oacc[j][3] *= a1;
    }
    l0 = l0 * a0 + r0;
    l1 = l1 * a1 + r1;
    mma_p_t(oacc, s, sV + buf * TILE_BYTES, lane, kv_valid);
    __syncthreads();
  }
  if (!warp_active) return;
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const int g = lane >> 2;
  if (lse != nullptr && t == 0) {
    const int r0i = q0 + warp * 16 + g, r1i = r0i + 8;
    float* lrow = lse + (static_cast<long long>(b) * H + h) * N;
    if (r0i < N) lrow[r0i] = (m0 + log2f(l0)) * LN2;
    if (r1i < N) lrow[r1i] = (m1 + log2f(l1)) * LN2;
  }
  store_c_tile(sQ, smem, oacc, 1.0f / l0, 1.0f / l1, warp * 16, lane, o + tok0 * ldo + h * HD, ldo, q0, N);
}

// ------------------------------------------------------------------------------------------------ backward
// delta[b,h,n] = sum_d dO[b,n,h,d] * O[b,n,h,d]; one warp per token, two lanes per head.
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, long long ldo, const __nv_bfloat16* __restrict__ dout,
                  long long lddo, float* __restrict__ delta, int B, int N, int H) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tok = blockIdx.x * 8LL + warp;
  if (tok >= static_cast<long long>(B) * N) return;
  const int b = static_cast<int>(tok / N), n = static_cast<int>(tok - static_cast<long long>(b) * N);
  for (int hbase = 0; hbase < H; hbase += 16) {  // warp-uniform trip count (full-mask shuffle below)
    const int hh = hbase + (lane >> 1);
    const bool valid = hh < H;
    float s = 0.f;
    if (valid) {
      const uint4* po = reinterpret_cast<const uint4*>(o + tok * ldo + hh * HD + (lane & 1) * 32);
      const uint4* pd = reinterpret_cast<const uint4*>(dout + tok * lddo + hh * HD + (lane & 1) * 32);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 a = __ldg(po + i), d = __ldg(pd + i);
        s += bf16_lo(a.x) * bf16_lo(d.x) + bf16_hi(a.x) * bf16_hi(d.x);
        s += bf16_lo(a.y) * bf16_lo(d.y) + bf16_hi(a.y) * bf16_hi(d.y);
        s += bf16_lo(a.z) * bf16_lo(d.z) + bf16_hi(a.z) * bf16_hi(d.z);
        s += bf16_lo(a.w) * bf16_lo(d.w) + bf16_hi(a.w) * bf16_hi(d.w);
      }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (valid && (lane & 1) == 0) delta[(static_cast<long long>(b) * H + hh) * N + n] = s;
  }
}

// dQ = scale * sum_keys dS K, dS = P o (dO V^T - delta); one CTA per 64 queries.
__global__ void __launch_bounds__(128)
attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                   const __nv_bfloat16* __restrict__ v, long long ld, const __nv_bfloat16* __restrict__ dout,
                   long long lddo, const float* __restrict__ lse, const float* __restrict__ delta,
                   __nv_bfloat16* __restrict__ dq, long long lddq, int N, int H, float scale) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * TILE, h = blockIdx.y, b = blockIdx.z;
  const long long tok0 = static_cast<long long>(b) * N;
  const __nv_bfloat16* qg = q + tok0 * ld + h * HD;
  const __nv_bfloat16* kg = k + tok0 * ld + h * HD;
  const __nv_bfloat16* vg = v + tok0 * ld + h * HD;
  const __nv_bfloat16* dog = dout + tok0 * lddo + h * HD;
  const uint32_t sQ = smem_u32(smem), sDO = sQ + TILE_BYTES, sK = sQ + 2 * TILE_BYTES, sV = sQ + 4 * TILE_BYTES;

  load_tile(sQ, qg, ld, q0, N, tid, 128);
  load_tile(sDO, dog, lddo, q0, N, tid, 128);
  load_tile(sK, kg, ld, 0, N, tid, 128);
  load_tile(sV, vg, ld, 0, N, tid, 128);
  cp_async_commit();

  const int g = lane >> 2, t = lane & 3;
  const int r0i = q0 + warp * 16 + g, r1i = r0i + 8;
  const float* lrow = lse + (static_cast<long long>(b) * H + h) * N;
  const float* drow = delta + (static_cast<long long>(b) * H + h) * N;
  const float L0 = r0i < N ? lrow[r0i] * LOG2E : INFINITY, L1 = r1i < N ? lrow[r1i] * LOG2E : INFINITY;
  const float D0 = r0i < N ? drow[r0i] : 0.f, D1 = r1i < N ? drow[r1i] : 0.f;
  const float c2 = scale * LOG2E;
  const int nkv = (N + TILE - 1) / TILE;
  const bool warp_active = q0 + warp * 16 < N;
  uint32_t qa[4][4], da[4][4];
  float acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;

  for (int kb = 0; kb < nkv; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nkv) {
      load_tile(sK + (buf ^ 1) * TILE_BYTES, kg, ld, (kb + 1) * TILE, N, tid, 128);
      load_tile(sV + (buf ^ 1) * TILE_BYTES, vg, ld, (kb + 1) * TILE, N, tid, 128);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kb == 0) {
      load_a_frags(sQ, warp * 16, lane, qa);
      load_a_frags(sDO, warp * 16, lane, da);
    }
    const int kv_valid = min(TILE, N - kb * TILE);
    if (!warp_active) {
      __syncthreads();
      continue;
    }
    float s[8][4], dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
    }
    mma_a_tT(s, qa, sK + buf * TILE_BYTES, lane, kv_valid);
    mma_a_tT(dp, da, sV + buf * TILE_BYTES, lane, kv_valid);
    const int kbase = kb * TILE;
    const bool tail = kbase + TILE > N;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float p = exp2f(s[j][e] * c2 - (e < 2 ? L0 : L1));
        if (tail && kbase + 8 * j + 2 * t + (e & 1) >= N) p = 0.f;
        s[j][e] = p * (dp[j][e] - (e < 2 ? D0 : D1));
      }
    }
    mma_p_t(acc, s, sK + buf * TILE_BYTES, lane, kv_valid);
    __syncthreads();
  }
  if (!warp_active) return;
  store_c_tile(sQ, smem, acc, scale, scale, warp * 16, lane, dq + tok0 * lddq + h * HD, lddq, q0, N);
}

// dK = scale * sum_q dS^T Q, dV = sum_q P^T dO; one CTA per 64 keys, everything computed transposed
// (keys as MMA rows) so that dK/dV accumulate in registers.
__global__ void __launch_bounds__(128)
attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                    const __nv_bfloat16* __restrict__ v, long long ld, const __nv_bfloat16* __restrict__ dout,
                    long long lddo, const float* __restrict__ lse, const float* __restrict__ delta,
                    __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, long long lddkv, int N, int H,
                    float scale) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k0 = blockIdx.x * TILE, h = blockIdx.y, b = blockIdx.z;
  const long long tok0 = static_cast<long long>(b) * N;
  const __nv_bfloat16* qg = q + tok0 * ld + h * HD;
  const __nv_bfloat16* kg = k + tok0 * ld + h * HD;
  const __nv_bfloat16* vg = v + tok0 * ld + h * HD;
  const __nv_bfloat16* dog = dout + tok0 * lddo + h * HD;
  const uint32_t sK = smem_u32(smem), sV = sK + TILE_BYTES, sQ = sK + 2 * TILE_BYTES, sDO = sK + 4 * TILE_BYTES;
  float* sL = reinterpret_cast<float*>(smem + 6 * TILE_BYTES);  // [2][64] lse * log2e (+inf for padded queries)
  float* sD = sL + 2 * TILE;                                     // [2][64] delta
  const float* lrow = lse + (static_cast<long long>(b) * H + h) * N;
  const float* drow = delta + (static_cast<long long>(b) * H + h) * N;

  load_tile(sK, kg, ld, k0, N, tid, 128);
  load_tile(sV, vg, ld, k0, N, tid, 128);
  load_tile(sQ, qg, ld, 0, N, tid, 128);
  load_tile(sDO, dog, lddo, 0, N, tid, 128);
  cp_async_commit();
  if (tid < TILE) {
    sL[tid] = tid < N ? lrow[tid] * LOG2E : INFINITY;
    sD[tid] = tid < N ? drow[tid] : 0.f;
  }

  const int t = lane & 3;
  const float c2 = scale * LOG2E;
  const int nq = (N + TILE - 1) / TILE;
  const bool warp_active = k0 + warp * 16 < N;
  uint32_t ka[4][4], va[4][4];
  float dkacc[8][4], dvacc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    dkacc[j][0] = dkacc[j][1] = dkacc[j][2] = dkacc[j][3] = 0.f;
    dvacc[j][0] = dvacc[j][1] = dvacc[j][2] = dvacc[j][3] = 0.f;
  }

  for (int qb = 0; qb < nq; ++qb) {
    const int buf = qb & 1;
    if (qb + 1 < nq) {
      load_tile(sQ + (buf ^ 1) * TILE_BYTES, qg, ld, (qb + 1) * TILE, N, tid, 128);
      load_tile(sDO + (buf ^ 1) * TILE_BYTES, dog, lddo, (qb + 1) * TILE, N, tid, 128);
      cp_async_commit();
      if (tid < TILE) {
        const int qi = (qb + 1) * TILE + tid;
        sL[(buf ^ 1) * TILE + tid] = qi < N ? lrow[qi] * LOG2E : INFINITY;
        sD[(buf ^ 1) * TILE + tid] = qi < N ? drow[qi] : 0.f;
      }
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (qb == 0) {
      load_a_frags(sK, warp * 16, lane, ka);
      load_a_frags(sV, warp * 16, lane, va);
    }
    const int q_valid = min(TILE, N - qb * TILE);
    if (!warp_active) {
      __syncthreads();
      continue;
    }
    float s[8][4], dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
    }
    mma_a_tT(s, ka, sQ + buf * TILE_BYTES, lane, q_valid);    // S^T[key, query]
    mma_a_tT(dp, va, sDO + buf * TILE_BYTES, lane, q_valid);  // dP^T[key, query]
    const float* Lb = sL + buf * TILE;
    const float* Db = sD + buf * TILE;
    float ds[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 Lq = *reinterpret_cast<const float2*>(Lb + 8 * j + 2 * t);
      const float2 Dq = *reinterpret_cast<const float2*>(Db + 8 * j + 2 * t);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p = exp2f(s[j][e] * c2 - ((e & 1) ? Lq.y : Lq.x));  // padded query: exp2(-inf) = 0
        s[j][e] = p;
        ds[j][e] = p * (dp[j][e] - ((e & 1) ? Dq.y : Dq.x));
      }
    }
    mma_p_t(dvacc, s, sDO + buf * TILE_BYTES, lane, q_valid);  // dV += P^T dO
    mma_p_t(dkacc, ds, sQ + buf * TILE_BYTES, lane, q_valid);  // dK += dS^T Q
    __syncthreads();
  }
  if (!warp_active) return;
  store_c_tile(sK, smem, dkacc, scale, scale, warp * 16, lane, dk + tok0 * lddkv + h * HD, lddkv, k0, N);
  store_c_tile(sV, smem + TILE_BYTES, dvacc, 1.f, 1.f, warp * 16, lane, dv + tok0 * lddkv + h * HD, lddkv, k0, N);
}

constexpr int DQ_SMEM = 6 * TILE_BYTES;
constexpr int DKV_SMEM = 6 * TILE_BYTES + 4 * TILE * 4;

}  // namespace

int attention_fwd(const void* q, const void* k, const void* v, long long ld, void* o, long long ldo, float* lse,
                  int B, int N, int H, int head_dim, float scale, cudaStream_t stream) {
  if (head_dim != HD) return set_error(kErrUnsupported, "attention: head_dim=%d (only 64 is supported)", head_dim);
  if (B <= 0 || N <= 0) return kOk;
  dim3 grid((N + TILE - 1) / TILE, H, B);
  ProfScope prof("attention_fwd", 4.0 * B * H * static_cast<double>(N) * N * HD, 8.0 * B * H * static_cast<double>(N) * HD, stream);
  attn_fwd_kernel<<<grid, 128, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(q),
                                            reinterpret_cast<const __nv_bfloat16*>(k),
                                            reinterpret_cast<const __nv_bfloat16*>(v), ld,
                                            reinterpret_cast<__nv_bfloat16*>(o), ldo, lse, N, H, scale);
  return check_launch("attention_fwd");
}

int attention_delta(const void* o, long long ldo, const void* dout, long long lddo, float* delta, int B, int N, int H,
                    cudaStream_t stream) {
  const long long toks = static_cast<long long>(B) * N;
  attn_delta_kernel<<<static_cast<int>((toks + 7) / 8), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(o), ldo,
                                                                          reinterpret_cast<const __nv_bfloat16*>(dout),
                                                                          lddo, delta, B, N, H);
  return check_launch("attention_delta");
}

int attention_bwd(const void* q, const void* k, const void* v, long long ld, const void* o, long long ldo,
                  const void* dout, long long lddo, const float* lse, float* delta, void* dq, void* dk, void* dv,
                  long long lddqkv, int B, int N, int H, int head_dim, float scale, cudaStream_t stream) {
  if (head_dim != HD) return set_error(kErrUnsupported, "attention: head_dim=%d (only 64 is supported)", head_dim);
  if (B <= 0 || N <= 0) return kOk;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e1 = cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM);
    cudaError_t e2 = cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DKV_SMEM);
    if (e1 != cudaSuccess || e2 != cudaSuccess) return set_error(kErrCuda, "attention_bwd: cudaFuncSetAttribute failed");
    attr_set = true;
  }
  auto* qb = reinterpret_cast<const __nv_bfloat16*>(q);
  auto* kb = reinterpret_cast<const __nv_bfloat16*>(k);
  auto* vb = reinterpret_cast<const __nv_bfloat16*>(v);
  auto* ob = reinterpret_cast<const __nv_bfloat16*>(o);
  auto* dob = reinterpret_cast<const __nv_bfloat16*>(dout);
  const long long toks = static_cast<long long>(B) * N;
  ProfScope prof("attention_bwd", 10.0 * B * H * static_cast<double>(N) * N * HD, 16.0 * B * H * static_cast<double>(N) * HD, stream);
  attn_delta_kernel<<<static_cast<int>((toks + 7) / 8), 256, 0, stream>>>(ob, ldo, dob, lddo, delta, B, N, H);
  int rc = check_launch("attention_delta");
  if (rc) return rc;
  dim3 grid((N + TILE - 1) / TILE, H, B);
  attn_bwd_dq_kernel<<<grid, 128, DQ_SMEM, stream>>>(qb, kb, vb, ld, dob, lddo, lse, delta,
                                                     reinterpret_cast<__nv_bfloat16*>(dq), lddqkv, N, H, scale);
  rc = check_launch("attention_bwd_dq");
  if (rc) return rc;
  attn_bwd_dkv_kernel<<<grid, 128, DKV_SMEM, stream>>>(qb, kb, vb, ld, dob, lddo, lse, delta,
                                                       reinterpret_cast<__nv_bfloat16*>(dk),
                                                       reinterpret_cast<__nv_bfloat16*>(dv), lddqkv, N, H, scale);
  return check_launch("attention_bwd_dkv");
}

}  // namespace tic
