// Fused multi-head attention FORWARD on the 5th-gen tensor cores (tcgen05.mma, S / P / O in tensor memory),
// Q / K / V tiles staged by TMA (cp.async.bulk.tensor.3d, 128-byte swizzle) straight out of the fused QKV GEMM
// output [B*N, 3*D]. Replaces F.scaled_dot_product_attention (modeling_vit.py:232-246 [a6]) for 197- and
// 577-token sequences (d = 64), non-causal, no mask, dropout 0.
//
// One CTA per (128-query tile, head, image); 2 CTAs per SM (256 TMEM columns, ~75 KB smem each).
//   warp 4 (one lane): TMA loads, S = Q K^T (SS MMA, M=128, N=KVB, K=64), O (+)= P V (TS MMA: P read from TMEM,
//                       V as an MN-major smem operand, M=128, N=64, K=KVB)
//   warps 0-3        : one query row per thread (TMEM lane == row): row max, exp2, row sum, P written back to TMEM
//                       as packed bf16 over the S columns it came from; online-softmax rescale of O between key
//                       blocks; final 1/l normalisation and bf16 store; logsumexp for the backward pass.
// TMEM columns: S [0, KVB) fp32, P [0, KVB/2) bf16x2 (aliases S), O [128, 192) fp32 (aliases the tail of S when
// the whole key range is a single block, which is the 197-token case: KVB = 224).
#include "tic_internal.cuh"

#include <cstdlib>

namespace tic {
namespace {

constexpr int AT_THREADS = 160;
constexpr int AT_QT = 128;      // queries per CTA (UMMA M)
constexpr int AT_HD = 64;       // head dim
constexpr int AT_O_COL = 128;   // TMEM column of the O accumulator
constexpr int AT_TMEM_COLS = 256;
constexpr float AT_LOG2E = 1.4426950408889634f;
constexpr float AT_LN2 = 0.6931471805599453f;

template <int KVB>
constexpr int at_smem_bytes() { return AT_QT * 128 + 2 * KVB * 128 + 1024 + 128; }

template <int KVB>
__global__ void __launch_bounds__(AT_THREADS)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, __nv_bfloat16* __restrict__ o, long long ldo,
                   float* __restrict__ lse, int N, int Nq, int H, float scale) {
  // N = keys per item; Nq = queries per item (the first Nq tokens; the grid covers ceil(Nq / AT_QT) query tiles)
  static_assert(KVB % 32 == 0 && KVB <= 256 && KVB / 2 <= AT_O_COL, "key block must be a multiple of 32, <= 256");
  extern __shared__ uint8_t at_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AT_QT * 128;
  uint8_t* sV = sK + KVB * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + KVB * 128);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_kv = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_p = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_QT, h = blockIdx.y, b = blockIdx.z;
  const int nblocks = (N + KVB - 1) / KVB;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_q);
      tma_prefetch_desc(&tm_k);
      tma_prefetch_desc(&tm_v);
      mbar_init(bar_q, 1);
      mbar_init(bar_kv, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 4);
      mbar_init(bar_o, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, AT_TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (elect_one()) {  // one issuing thread the compiler can keep on the uniform datapath
      constexpr uint32_t idesc_s = make_idesc_bf16(AT_QT, KVB, false, false);
      constexpr uint32_t idesc_o = make_idesc_bf16(AT_QT, AT_HD, false, true);
      mbar_arrive_expect_tx(bar_q, AT_QT * 128);
      tma_load_3d(sQ, &tm_q, bar_q, h * AT_HD, q0, b);
      const uint64_t dq = make_smem_desc_sw128(smem_u32(sQ), 0, 1024);
      const uint64_t dk = make_smem_desc_sw128(smem_u32(sK), 0, 1024);
      const uint64_t dv = make_smem_desc_sw128(smem_u32(sV), 8192, 1024);
      for (int j = 0; j < nblocks; ++j) {
        if (j > 0) {  // P V of the previous block has finished: K / V smem and the S / P columns are free again
          mbar_wait(bar_o, (j - 1) & 1);
          tc_fence_after();
        }
        mbar_arrive_expect_tx(bar_kv, 2 * KVB * 128);
        tma_load_3d(sK, &tm_k, bar_kv, h * AT_HD, j * KVB, b);
        tma_load_3d(sV, &tm_v, bar_kv, h * AT_HD, j * KVB, b);
        if (j == 0) mbar_wait(bar_q, 0);
        mbar_wait(bar_kv, j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < AT_HD / 16; ++k) umma_bf16_ss(tmem_base, dq + 2 * k, dk + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(bar_s);
        mbar_wait(bar_p, j & 1);
        tc_fence_after();
#pragma unroll 1
        for (int k = 0; k < KVB / 16; ++k)
          umma_bf16_ts(tmem_base + AT_O_COL, tmem_base + 8 * k, dv + 128 * k, idesc_o, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(bar_o);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps: one query row per thread
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const float c2 = scale * AT_LOG2E;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < nblocks; ++j) {
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      const int nvalid = min(KVB, N - j * KVB);
      // pass 1: row maximum of the raw scores over the valid keys (scaled into the log2 domain afterwards)
      float raw_max = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < KVB / 32; ++c) {
        if (c * 32 >= nvalid) break;
        uint32_t r[32];
        tmem_ld_32x32b_x32(lane_addr + c * 32, r);
        tmem_ld_wait();
        if (c * 32 + 32 <= nvalid) {
#pragma unroll
          for (int i = 0; i < 32; i += 2)
            raw_max = fmaxf(raw_max, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < nvalid) raw_max = fmaxf(raw_max, __uint_as_float(r[i]));
        }
      }
      const float mx = fmaxf(m, raw_max * c2);
      const float alpha = ex2_approx(m - mx);  // 0 on the first block (m = -inf)
      if (j > 0) {                        // rescale the running output accumulator
#pragma unroll 1
        for (int c = 0; c < AT_HD / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(lane_addr + AT_O_COL + c * 32, r);
          tmem_ld_wait();
          uint32_t w0[16], w1[16];
#pragma unroll
          for (int t = 0; t < 16; ++t) {
            w0[t] = __float_as_uint(__uint_as_float(r[t]) * alpha);
            w1[t] = __float_as_uint(__uint_as_float(r[16 + t]) * alpha);
          }
          tmem_st_32x32b_x16(lane_addr + AT_O_COL + c * 32, w0);
          tmem_st_32x32b_x16(lane_addr + AT_O_COL + c * 32 + 16, w1);
        }
      }
      // pass 2: p = exp2(s * c2 - mx), row sum, packed bf16 P over the S columns
      float sum = 0.f;
      const float neg_mx = -mx;
#pragma unroll 1
      for (int c = 0; c < KVB / 32; ++c) {
        uint32_t w[16];
        if (c * 32 >= nvalid) {  // block past the last key: P = 0
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = 0u;
        } else {
          uint32_t r[32];
          tmem_ld_32x32b_x32(lane_addr + c * 32, r);
          tmem_ld_wait();
          if (c * 32 + 32 <= nvalid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * i]), c2, neg_mx));
              const float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * i + 1]), c2, neg_mx));
              sum += p0 + p1;
              w[i] = pack_bf16x2(p0, p1);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * i]), c2, neg_mx));
              float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * i + 1]), c2, neg_mx));
              if (c * 32 + 2 * i >= nvalid) p0 = 0.f;
              if (c * 32 + 2 * i + 1 >= nvalid) p1 = 0.f;
              sum += p0 + p1;
              w[i] = pack_bf16x2(p0, p1);
            }
          }
        }
        tmem_st_32x32b_x16(lane_addr + c * 16, w);
      }
      tmem_st_wait();
      l = l * alpha + sum;
      m = mx;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
    }
    // ---- epilogue: O / l -> bf16, logsumexp
    mbar_wait(bar_o, (nblocks - 1) & 1);
    tc_fence_after();
    const int row = q0 + warp * 32 + lane;
    const float inv_l = 1.0f / l;
    uint32_t packed[32];
#pragma unroll
    for (int c = 0; c < AT_HD / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(lane_addr + AT_O_COL + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i)
        packed[c * 16 + i] = pack_bf16x2(__uint_as_float(r[2 * i]) * inv_l, __uint_as_float(r[2 * i + 1]) * inv_l);
    }
    if (row < Nq) {
      const long long tok = static_cast<long long>(b) * N + row;
      uint4* dst = reinterpret_cast<uint4*>(o + tok * ldo + h * AT_HD);
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
      if (lse != nullptr) lse[(static_cast<long long>(b) * H + h) * Nq + row] = (m + log2f(l)) * AT_LN2;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4) tmem_dealloc(tmem_base, AT_TMEM_COLS);
}

template <int KVB>
int launch_fwd(const void* q, const void* k, const void* v, long long ld, void* o, long long ldo, float* lse, int B,
               int N, int Nq, int H, float scale, cudaStream_t stream) {
  CUtensorMap tq, tk, tv;
  const uint64_t D = static_cast<uint64_t>(H) * AT_HD;
  int rc = encode_tmap_3d_bf16(&tq, q, D, Nq, B, ld, static_cast<uint64_t>(N) * ld, 64, AT_QT);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&tk, k, D, N, B, ld, static_cast<uint64_t>(N) * ld, 64, KVB);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&tv, v, D, N, B, ld, static_cast<uint64_t>(N) * ld, 64, KVB);
  if (rc) return rc;
  auto kern = attn_fwd_tc_kernel<KVB>;
  if (int rc2 = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), at_smem_bytes<KVB>(), "attention_fwd_tc")) return rc2;
  dim3 grid((Nq + AT_QT - 1) / AT_QT, H, B);
  kern<<<grid, AT_THREADS, at_smem_bytes<KVB>(), stream>>>(tq, tk, tv, reinterpret_cast<__nv_bfloat16*>(o), ldo, lse, N,
                                                          Nq, H, scale);
  return check_launch("attention_fwd_tc");
}


// ------------------------------------------------------------------------------------------------ backward
// One templated kernel for both halves of the backward pass; each CTA owns a 128-row tile and streams 64-wide
// column blocks through TMEM:
//   DKV = true : rows = keys   (R1 = K, R2 = V), columns = queries (C1 = Q, C2 = dO)
//                S^T = K Q^T, dP^T = V dO^T; P^T = exp2(S^T c - L[q]); dS^T = P^T o (dP^T - delta[q]);
//                dV += P^T dO, dK += dS^T Q          (P^T / dS^T stay in TMEM as the A operands of the TS MMAs)
//   DKV = false: rows = queries (R1 = Q, R2 = dO), columns = keys (C1 = K, C2 = V)
//                S = Q K^T, dP = dO V^T; dS = exp2(S c - L[row]) o (dP - delta[row]); dQ += dS K
// TMEM: S [0,64) -> P bf16 [0,32);  dP [64,128) -> dS bf16 [64,96);  acc1 [128,192);  acc2 [192,256).
constexpr int AB_CB = 64;  // column block
constexpr int AB_SMEM = 2 * AT_QT * 128 + 2 * 2 * AB_CB * 128 + 4 * AB_CB * 4 + 1024 + 128;

template <bool DKV>
__global__ void __launch_bounds__(AT_THREADS)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_r1, const __grid_constant__ CUtensorMap tm_r2,
                   const __grid_constant__ CUtensorMap tm_c1, const __grid_constant__ CUtensorMap tm_c2,
                   const float* __restrict__ lse, const float* __restrict__ delta, __nv_bfloat16* __restrict__ out1,
                   __nv_bfloat16* __restrict__ out2, long long ldout, int N, int Nq, int H, float scale) {
  // N = tokens (keys) per item, Nq = queries per item (the first Nq tokens). Rows / columns of this instantiation:
  constexpr bool kRowsAreKeys = DKV;
  const int Nr = kRowsAreKeys ? N : Nq, Nc = kRowsAreKeys ? Nq : N;
  extern __shared__ uint8_t at_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sR1 = smem;
  uint8_t* sR2 = sR1 + AT_QT * 128;
  uint8_t* sC1 = sR2 + AT_QT * 128;          // [2][64 rows][128 B]
  uint8_t* sC2 = sC1 + 2 * AB_CB * 128;      // [2][64 rows][128 B]
  float* sL = reinterpret_cast<float*>(sC2 + 2 * AB_CB * 128);  // [2][64]
  float* sD = sL + 2 * AB_CB;                                     // [2][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 2 * AB_CB);
  uint64_t* bar_rows = bars + 0;
  uint64_t* bar_ld = bars + 1;  // [2]
  uint64_t* bar_s = bars + 3;
  uint64_t* bar_p = bars + 4;
  uint64_t* bar_acc = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * AT_QT, h = blockIdx.y, b = blockIdx.z;
  const int nblk = (Nc + AB_CB - 1) / AB_CB;
  const float* lrow = lse + (static_cast<long long>(b) * H + h) * Nq;    // per query
  const float* drow = delta + (static_cast<long long>(b) * H + h) * Nq;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_r1); tma_prefetch_desc(&tm_r2); tma_prefetch_desc(&tm_c1); tma_prefetch_desc(&tm_c2);
      mbar_init(bar_rows, 1);
      mbar_init(&bar_ld[0], 1);
      mbar_init(&bar_ld[1], 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 4);
      mbar_init(bar_acc, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, AT_TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (elect_one()) {
      constexpr uint32_t idesc_ss = make_idesc_bf16(AT_QT, AB_CB, false, false);
      constexpr uint32_t idesc_ts = make_idesc_bf16(AT_QT, AT_HD, false, true);
      mbar_arrive_expect_tx(bar_rows, 2 * AT_QT * 128);
      tma_load_3d(sR1, &tm_r1, bar_rows, h * AT_HD, r0, b);
      tma_load_3d(sR2, &tm_r2, bar_rows, h * AT_HD, r0, b);
      mbar_arrive_expect_tx(&bar_ld[0], 2 * AB_CB * 128);
      tma_load_3d(sC1, &tm_c1, &bar_ld[0], h * AT_HD, 0, b);
      tma_load_3d(sC2, &tm_c2, &bar_ld[0], h * AT_HD, 0, b);
      const uint64_t d_r1 = make_smem_desc_sw128(smem_u32(sR1), 0, 1024);
      const uint64_t d_r2 = make_smem_desc_sw128(smem_u32(sR2), 0, 1024);
      for (int j = 0; j < nblk; ++j) {
        const int buf = j & 1;
        if (j > 0) {  // TS MMAs of block j-1 are done: its column buffers and the P / dS columns are free
          mbar_wait(bar_acc, (j - 1) & 1);
          tc_fence_after();
        }
        if (j + 1 < nblk) {
          const int nb = (j + 1) & 1;
          mbar_arrive_expect_tx(&bar_ld[nb], 2 * AB_CB * 128);
          tma_load_3d(sC1 + nb * AB_CB * 128, &tm_c1, &bar_ld[nb], h * AT_HD, (j + 1) * AB_CB, b);
          tma_load_3d(sC2 + nb * AB_CB * 128, &tm_c2, &bar_ld[nb], h * AT_HD, (j + 1) * AB_CB, b);
        }
        if (j == 0) mbar_wait(bar_rows, 0);
        mbar_wait(&bar_ld[buf], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t c1_addr = smem_u32(sC1 + buf * AB_CB * 128), c2_addr = smem_u32(sC2 + buf * AB_CB * 128);
        const uint64_t d_c1k = make_smem_desc_sw128(c1_addr, 0, 1024), d_c2k = make_smem_desc_sw128(c2_addr, 0, 1024);
#pragma unroll
        for (int k = 0; k < AT_HD / 16; ++k) umma_bf16_ss(tmem_base, d_r1 + 2 * k, d_c1k + 2 * k, idesc_ss, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < AT_HD / 16; ++k) umma_bf16_ss(tmem_base + 64, d_r2 + 2 * k, d_c2k + 2 * k, idesc_ss, k > 0 ? 1u : 0u);
        umma_commit(bar_s);
        mbar_wait(bar_p, j & 1);
        tc_fence_after();
        const uint64_t d_c1m = make_smem_desc_sw128(c1_addr, 8192, 1024), d_c2m = make_smem_desc_sw128(c2_addr, 8192, 1024);
        const uint32_t accf = j > 0 ? 1u : 0u;
        if constexpr (DKV) {
#pragma unroll
          for (int k = 0; k < AB_CB / 16; ++k)  // dV += P^T dO
            umma_bf16_ts(tmem_base + 128, tmem_base + 8 * k, d_c2m + 128 * k, idesc_ts, (k > 0) ? 1u : accf);
#pragma unroll
          for (int k = 0; k < AB_CB / 16; ++k)  // dK += dS^T Q
            umma_bf16_ts(tmem_base + 192, tmem_base + 64 + 8 * k, d_c1m + 128 * k, idesc_ts, (k > 0) ? 1u : accf);
        } else {
#pragma unroll
          for (int k = 0; k < AB_CB / 16; ++k)  // dQ += dS K
            umma_bf16_ts(tmem_base + 128, tmem_base + 64 + 8 * k, d_c1m + 128 * k, idesc_ts, (k > 0) ? 1u : accf);
        }
        umma_commit(bar_acc);
      }
    }
  } else {
    const int tid = threadIdx.x;  // 0..127
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const float c2 = scale * AT_LOG2E;
    const int row = r0 + warp * 32 + lane;
    float L_row = INFINITY, D_row = 0.f;
    if constexpr (!DKV) {
      if (row < Nr) { L_row = lrow[row] * AT_LOG2E; D_row = drow[row]; }
    }
    for (int j = 0; j < nblk; ++j) {
      const int buf = j & 1;
      if constexpr (DKV) {
        if (tid < AB_CB) {
          const int i = j * AB_CB + tid;
          sL[buf * AB_CB + tid] = i < Nc ? lrow[i] * AT_LOG2E : INFINITY;  // padded query: exp2(-inf) = 0
          sD[buf * AB_CB + tid] = i < Nc ? drow[i] : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      const int nvalid = min(AB_CB, Nc - j * AB_CB);
#pragma unroll 1
      for (int c = 0; c < AB_CB / 32; ++c) {
        uint32_t s[32], dp[32];
        tmem_ld_32x32b_x32(lane_addr + c * 32, s);
        tmem_ld_32x32b_x32(lane_addr + 64 + c * 32, dp);
        tmem_ld_wait();
        uint32_t pw[16], dw[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0, p1, ds0, ds1;
          if constexpr (DKV) {
            const float2 Lq = *reinterpret_cast<const float2*>(sL + buf * AB_CB + c * 32 + 2 * i);
            const float2 Dq = *reinterpret_cast<const float2*>(sD + buf * AB_CB + c * 32 + 2 * i);
            p0 = ex2_approx(fmaf(__uint_as_float(s[2 * i]), c2, -Lq.x));
            p1 = ex2_approx(fmaf(__uint_as_float(s[2 * i + 1]), c2, -Lq.y));
            ds0 = p0 * (__uint_as_float(dp[2 * i]) - Dq.x);
            ds1 = p1 * (__uint_as_float(dp[2 * i + 1]) - Dq.y);
          } else {
            p0 = ex2_approx(fmaf(__uint_as_float(s[2 * i]), c2, -L_row));
            p1 = ex2_approx(fmaf(__uint_as_float(s[2 * i + 1]), c2, -L_row));
            if (c * 32 + 2 * i >= nvalid) p0 = 0.f;       // padded key
            if (c * 32 + 2 * i + 1 >= nvalid) p1 = 0.f;
            ds0 = p0 * (__uint_as_float(dp[2 * i]) - D_row);
            ds1 = p1 * (__uint_as_float(dp[2 * i + 1]) - D_row);
          }
          pw[i] = pack_bf16x2(p0, p1);
          dw[i] = pack_bf16x2(ds0, ds1);
        }
        if constexpr (DKV) tmem_st_32x32b_x16(lane_addr + c * 16, pw);
        tmem_st_32x32b_x16(lane_addr + 64 + c * 16, dw);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
    }
    // ---- epilogue: accumulators -> bf16 rows
    mbar_wait(bar_acc, (nblk - 1) & 1);
    tc_fence_after();
    const long long tok = static_cast<long long>(b) * N + row;
#pragma unroll 1
    for (int a = 0; a < (DKV ? 2 : 1); ++a) {
      // DKV: a = 0 -> dV (acc1, unscaled) to out2; a = 1 -> dK (acc2, * scale) to out1.  DQ: dQ (acc1, * scale) to out1.
      const float f = (DKV && a == 0) ? 1.0f : scale;
      __nv_bfloat16* dst_base = (DKV && a == 0) ? out2 : out1;
      uint32_t packed[32];
#pragma unroll
      for (int c = 0; c < AT_HD / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(lane_addr + 128 + 64 * a + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          packed[c * 16 + i] = pack_bf16x2(__uint_as_float(r[2 * i]) * f, __uint_as_float(r[2 * i + 1]) * f);
      }
      if (row < Nr) {
        uint4* dst = reinterpret_cast<uint4*>(dst_base + tok * ldout + h * AT_HD);
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4) tmem_dealloc(tmem_base, AT_TMEM_COLS);
}

template <bool DKV>
int launch_bwd(const void* r1, const void* r2, long long ldr1, long long ldr2, const void* c1, const void* c2,
               long long ldc1, long long ldc2, const float* lse, const float* delta, void* out1, void* out2,
               long long ldout, int B, int N, int Nq, int H, float scale, cudaStream_t stream) {
  CUtensorMap t1, t2, t3, t4;
  const uint64_t D = static_cast<uint64_t>(H) * AT_HD;
  const int Nr = DKV ? N : Nq, Nc = DKV ? Nq : N;  // rows / columns of this half (keys x queries, or queries x keys)
  int rc = encode_tmap_3d_bf16(&t1, r1, D, Nr, B, ldr1, static_cast<uint64_t>(N) * ldr1, 64, AT_QT);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&t2, r2, D, Nr, B, ldr2, static_cast<uint64_t>(N) * ldr2, 64, AT_QT);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&t3, c1, D, Nc, B, ldc1, static_cast<uint64_t>(N) * ldc1, 64, AB_CB);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&t4, c2, D, Nc, B, ldc2, static_cast<uint64_t>(N) * ldc2, 64, AB_CB);
  if (rc) return rc;
  auto kern = attn_bwd_tc_kernel<DKV>;
  if (int rc2 = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), AB_SMEM, "attention_bwd_tc")) return rc2;
  dim3 grid((Nr + AT_QT - 1) / AT_QT, H, B);
  kern<<<grid, AT_THREADS, AB_SMEM, stream>>>(t1, t2, t3, t4, lse, delta, reinterpret_cast<__nv_bfloat16*>(out1),
                                             reinterpret_cast<__nv_bfloat16*>(out2), ldout, N, Nq, H, scale);
  return check_launch(DKV ? "attention_bwd_dkv_tc" : "attention_bwd_dq_tc");
}

// ------------------------------------------------------------------------------------------------ delta
// delta[b,h,n] = sum_d dO[b,n,h,d] * O[b,n,h,d]; one warp per token, two lanes per head.
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, long long ldo, const __nv_bfloat16* __restrict__ dout,
                  long long lddo, float* __restrict__ delta, int B, int N, int Nq, int H) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long qtok = blockIdx.x * 8LL + warp;  // index among the B * Nq query tokens
  if (qtok >= static_cast<long long>(B) * Nq) return;
  const int b = static_cast<int>(qtok / Nq), n = static_cast<int>(qtok - static_cast<long long>(b) * Nq);
  const long long tok = static_cast<long long>(b) * N + n;
  for (int hbase = 0; hbase < H; hbase += 16) {  // warp-uniform trip count (full-mask shuffle below)
    const int hh = hbase + (lane >> 1);
    const bool valid = hh < H;
    float s = 0.f;
    if (valid) {
      const uint4* po = reinterpret_cast<const uint4*>(o + tok * ldo + hh * AT_HD + (lane & 1) * 32);
      const uint4* pd = reinterpret_cast<const uint4*>(dout + tok * lddo + hh * AT_HD + (lane & 1) * 32);
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
    if (valid && (lane & 1) == 0) delta[(static_cast<long long>(b) * H + hh) * Nq + n] = s;
  }
}


}  // namespace

int attention_delta(const void* o, long long ldo, const void* dout, long long lddo, float* delta, int B, int N, int H,
                    cudaStream_t stream, int Nq) {
  if (Nq <= 0) Nq = N;
  const long long toks = static_cast<long long>(B) * Nq;
  attn_delta_kernel<<<static_cast<int>((toks + 7) / 8), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(o), ldo,
                                                                          reinterpret_cast<const __nv_bfloat16*>(dout),
                                                                          lddo, delta, B, N, Nq, H);
  return check_launch("attention_delta");
}

int attention_fwd_tc(const void* q, const void* k, const void* v, long long ld, void* o, long long ldo, float* lse,
                     int B, int N, int H, int head_dim, float scale, cudaStream_t stream, int Nq) {
  if (Nq <= 0) Nq = N;
  if (Nq > N) return set_error(kErrInvalidArg, "attention: Nq=%d > N=%d", Nq, N);
  if (head_dim != AT_HD) return set_error(kErrUnsupported, "attention: head_dim=%d (only 64 is supported)", head_dim);
  if (B <= 0 || N <= 0) return kOk;
  if ((ld % 8) || (reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(k) & 15) ||
      (reinterpret_cast<uintptr_t>(v) & 15))
    return set_error(kErrInvalidArg, "attention: q/k/v must be 16-byte aligned with a pitch that is a multiple of 8");
  ProfScope prof("attention_fwd", 4.0 * B * H * static_cast<double>(N) * N * AT_HD, 8.0 * B * H * static_cast<double>(N) * AT_HD, stream);
  if (N <= 224) {
    if ((ldo % 8) || (reinterpret_cast<uintptr_t>(o) & 15))
      return set_error(kErrInvalidArg, "attention: o must be 16-byte aligned with a pitch that is a multiple of 8");
    return attention_fwd_fused(q, k, v, ld, o, ldo, lse, B, N, Nq, H, scale, stream);  // persistent, one key block
  }
  if ((ldo % 8) || (reinterpret_cast<uintptr_t>(o) & 15))
    return set_error(kErrInvalidArg, "attention: o must be 16-byte aligned with a pitch that is a multiple of 8");
  static const bool split_kernels = std::getenv("TIC_ATTN_SPLIT") != nullptr;  // development A/B: the first-generation kernels
  if (split_kernels) return launch_fwd<128>(q, k, v, ld, o, ldo, lse, B, N, Nq, H, scale, stream);
  return attention_fwd_long(q, k, v, ld, o, ldo, lse, B, N, Nq, H, scale, stream);  // persistent, K / V ring, online softmax
}

int attention_bwd_tc(const void* q, const void* k, const void* v, long long ld, const void* o, long long ldo,
                     const void* dout, long long lddo, const float* lse, float* delta, void* dq, void* dk, void* dv,
                     long long lddqkv, int B, int N, int H, int head_dim, float scale, cudaStream_t stream,
                     float* bias_grad, int bias_mask, int Nq) {
  if (Nq <= 0) Nq = N;
  if (Nq > N) return set_error(kErrInvalidArg, "attention_bwd: Nq=%d > N=%d", Nq, N);
  if (head_dim != AT_HD) return set_error(kErrUnsupported, "attention: head_dim=%d (only 64 is supported)", head_dim);
  if (B <= 0 || N <= 0) return kOk;
  if ((ld % 8) || (lddo % 8) || (lddqkv % 8) || (ldo % 8))
    return set_error(kErrInvalidArg, "attention_bwd: pitches must be multiples of 8");
  if (N <= 256) {  // whole head in shared memory: one fused kernel (delta, dQ, dK, dV, QKV bias gradient)
    ProfScope prof("attention_bwd", 10.0 * B * H * static_cast<double>(N) * N * AT_HD, 16.0 * B * H * static_cast<double>(N) * AT_HD, stream);
    return attention_bwd_fused(q, k, v, ld, o, ldo, dout, lddo, lse, dq, dk, dv, lddqkv, bias_grad, bias_mask, B, N, Nq, H, scale, stream);
  }
  int rc;
  static const bool split_kernels = std::getenv("TIC_ATTN_SPLIT") != nullptr;  // development A/B: the first-generation kernels
  if (!split_kernels && N <= attention_bwd_long_max_queries()) {
    // one fused kernel (delta included: its loader warps compute it per head from O and dO); dQ is accumulated over the key
    // tiles in per-CTA fp32 slabs that sit behind the delta area of the scratch buffer (attention_bwd_scratch_floats)
    const long long delta_floats = (static_cast<long long>(B) * H * N + 63) / 64 * 64;
    ProfScope prof("attention_bwd", 10.0 * B * H * static_cast<double>(N) * N * AT_HD,
                   2.0 * B * H * AT_HD * (4.0 * N + 5.0 * Nq), stream);
    return attention_bwd_long(q, k, v, ld, o, ldo, dout, lddo, lse, delta + delta_floats, dq, dk, dv, lddqkv, bias_grad,
                              bias_mask, B, N, Nq, H, scale, stream);
  }
  {
    ProfScope prof("attention_delta", 0.0, 4.0 * B * static_cast<double>(Nq) * H * AT_HD, stream);
    rc = attention_delta(o, ldo, dout, lddo, delta, B, N, H, stream, Nq);
  }
  if (rc) return rc;
  ProfScope prof("attention_bwd", 10.0 * B * H * static_cast<double>(N) * N * AT_HD, 16.0 * B * H * static_cast<double>(N) * AT_HD, stream);
  // dQ: rows = queries (Q, dO), columns = keys (K, V)
  rc = launch_bwd<false>(q, dout, ld, lddo, k, v, ld, ld, lse, delta, dq, nullptr, lddqkv, B, N, Nq, H, scale, stream);
  if (rc) return rc;
  // dK / dV: rows = keys (K, V), columns = queries (Q, dO)
  rc = launch_bwd<true>(k, v, ld, ld, q, dout, ld, lddo, lse, delta, dk, dv, lddqkv, B, N, Nq, H, scale, stream);
  if (rc || bias_grad == nullptr) return rc;
  const int Dm = H * AT_HD;
  if (bias_mask & 1) {  // only the query rows of dq exist
    rc = Nq == N ? colsum_bf16(dq, lddqkv, B * N, Dm, bias_grad, stream)
                 : colsum_bf16(dq, static_cast<long long>(N) * lddqkv, B, Dm, bias_grad, stream);
  }
  if (rc) return rc;
  if (bias_mask & 2) rc = colsum_bf16(dk, lddqkv, B * N, Dm, bias_grad + Dm, stream);
  if (rc) return rc;
  if (bias_mask & 4) rc = colsum_bf16(dv, lddqkv, B * N, Dm, bias_grad + 2 * Dm, stream);
  return rc;
}

}  // namespace tic
