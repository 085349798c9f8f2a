// Fused multi-head attention FORWARD for sequences of up to 224 tokens (the 197-token ViT-*/16 224x224 case) as a
// persistent, warp-specialised kernel. Replaces F.scaled_dot_product_attention (modeling_vit.py:232-246 [a6]),
// non-causal, no mask, dropout 0, head_dim 64.
//
// One CTA per SM walks work units (image, head, 128-query tile). All keys of an item fit in ONE block, so there is no
// online rescaling: S = Q K^T (SS MMA, M = 128 queries, N = keys rounded up to 16), row max, P = exp2(S c - max),
// O = P V (TS MMA, P read from TMEM as packed bf16), O / rowsum -> bf16.
// The floor is the exp2 throughput of the SFUs (16 per clock per SM: 3328 clk per (image, head) at N = 197); the roles are
// arranged so that the SFUs wait as little as possible (measured: the exp2 pass itself runs at ~0.8 of the SFU rate, the
// kernel as a whole at 0.55 of the floor -- the rest is the load / max / commit phases of a unit, which are issue- and
// latency-bound):
//   warps 0-7   softmax engine. ALL eight work on the same unit: warp w owns TMEM lane quadrant w & 3 (one query row
//               per thread) and column half w >> 2 of the score tile, so every SM sub-partition has two warps feeding
//               its SFU. A thread loads its <= 112 scores into registers ONCE (row max and exp2 both run from
//               registers); the two halves of a row exchange their partial maxima through shared memory.
//   warps 8-11  drain: O of the PREVIOUS unit out of TMEM, 1/rowsum, bf16, shared-memory tile, TMA store, logsumexp --
//               under the engine's work on the current unit.
//   warp  12    one elected thread: TMA loads (operands of the next item prefetched into the other stage) and all MMAs.
// Two TMEM buffers of 256 columns alternate between consecutive units: S [0, nk), later P (packed bf16) in
// [0, nk / 2), and the O accumulator in [192, 256) over the dead tail of S. The score MMAs of unit u+1, the P V MMAs
// and the drain of unit u-1 all run under the engine's exp2 pass of unit u.
// (Measured and rejected: accumulating O(u) in the tail of the OTHER buffer so that S(u+2) need not wait for the drain
// of O(u) -- P V(u) must then wait until the engine has pulled S(u+1) into registers, which costs more than it saves.
// Also rejected: taking every 4th / 2nd exponential off the SFU with a cubic exp2 polynomial on the FMA pipe (the
// FlashAttention-4 trick): 0.109 -> 0.142 / 0.190 ms -- nine extra instructions per element cost more issue slots than
// the 8 SFU clocks they save.)
// Outputs leave through shared-memory tiles and TMA stores (rows past Nq are clipped by the tensor map): direct
// 16-byte stores at a 2 KB row pitch cost ~300 clk per instruction.
#include "tic_internal.cuh"

#include <type_traits>

namespace tic {
namespace {

constexpr int FF_ENGINE_WARPS = 8;
constexpr int FF_DRAIN_WARPS = 4;
constexpr int FF_ISSUE_WARP = FF_ENGINE_WARPS + FF_DRAIN_WARPS;
constexpr int FF_THREADS = 512;                             // 13 working warps + 3 idle (warps are allocated in fours)
// Register budget: compiled for 128 registers per thread, re-balanced at run time (setmaxnreg): the two engine
// warpgroups grow to 160 (a thread keeps its 112 scores in registers), the drain / issue warpgroups shrink to 96 --
// 8 x 32 x (160 + 96) = the whole register file.
constexpr int FF_REGS_ENGINE = 160, FF_REGS_OTHER = 96;
constexpr int FF_HD = 64;
constexpr int FF_KV = 224;                                  // key rows staged per item
constexpr int FF_Q_BYTES = 256 * 128;                       // 32 KB
constexpr int FF_KV_BYTES = FF_KV * 128;                    // 28 KB
constexpr int FF_STAGE_BYTES = FF_Q_BYTES + 2 * FF_KV_BYTES;  // 88 KB
constexpr int FF_OUT_BYTES = FF_DRAIN_WARPS * 4096;         // per drain warp: 32 rows x 128 B
constexpr int FF_STAT_BYTES = 2 /*buffers*/ * 5 /*max0 max1 sum0 sum1 mx*/ * 128 * 4;
constexpr int FF_SMEM_USED = 2 * FF_STAGE_BYTES + FF_OUT_BYTES + FF_STAT_BYTES + 256;
constexpr int FF_SMEM = FF_SMEM_USED + 1024;
constexpr uint32_t FF_O_COL = 192;
constexpr int FF_PAIRS = 7;                                 // 16-column register groups per engine warp (224 / 2 / 16)
constexpr float FF_LOG2E = 1.4426950408889634f;
constexpr float FF_LN2 = 0.6931471805599453f;

TIC_DEVINL void ff_st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
TIC_DEVINL void ff_tmem_ld_x8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
TIC_DEVINL void ff_tmem_st_x4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3])
               : "memory");
}
TIC_DEVINL float ff_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
TIC_DEVINL float ff_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// N16T = compile-time number of 16-key chunks (13: N in 193..208, which covers the 197-token ViT-*/16 224x224 sequence --
// every loop bound and column offset of the engine folds to a constant, which halves its instruction count) or 0 = any
// N <= 224 at run time (correct, but the engine's register arrays are then walked under run-time predicates: slow).
template <int N16T>
__global__ void __launch_bounds__(FF_THREADS, 1)
attn_fwd_fused_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                      float* __restrict__ lse, int N_rt, int Nq, int H, int num_items, float scale, long long* __restrict__ trace) {
#ifdef TIC_ATTN_TRACE  // development build only (-DTIC_ATTN_TRACE): clock64 stamps of CTA 0, units 8..11; the shipped library has no tracing code
#define FF_STAMP(u, slot) do { if (trace != nullptr && blockIdx.x == 0 && (u) >= 8 && (u) < 12) trace[((u) - 8) * 16 + (slot)] = clock64(); } while (0)
#else
#define FF_STAMP(u, slot) do { } while (0)
#endif
  // N = keys per item; Nq = queries per item (the first Nq tokens of each image: Nq = N normally, Nq = 1 when only the
  // CLS row of the last encoder layer is needed)
  const int N = N_rt;
  pdl_launch_dependents();
  extern __shared__ uint8_t ff_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ff_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sOut = smem + 2 * FF_STAGE_BYTES;
  float* sStat = reinterpret_cast<float*>(sOut + FF_OUT_BYTES);   // [2 buffers][5][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sStat) + FF_STAT_BYTES);
  // every per-unit barrier is indexed by the unit's parity u & 1 and completes once per unit of that parity
  uint64_t* bar_ld = bars + 0;     // [2 stages] operand stage loaded
  uint64_t* bar_s = bars + 2;      // [2] S of the unit is complete in TMEM
  uint64_t* bar_p = bars + 6;      // [2] P written (8 engine-warp arrivals)
  uint64_t* bar_o = bars + 8;      // [2] P V has completed (P consumed, O complete)
  uint64_t* bar_ofree = bars + 10; // [2] the drain warps have read O out of TMEM (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqt = (Nq + 127) >> 7;         // query tiles per item (1 or 2)
  const int n16 = N16T ? N16T : (N + 15) >> 4;  // 16-key column chunks (<= 14); padded keys are zero rows
  const int nk = n16 * 16;                 // keys rounded up to the UMMA N / K granularity
  const int h0_16 = (n16 + 1) >> 1;        // chunks of column half 0 (<= 7); half 1 takes the rest

  if (warp == FF_ISSUE_WARP) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_o);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bar_ld[i], 1);
        mbar_init(&bar_s[i], 1);
        mbar_init(&bar_p[i], FF_ENGINE_WARPS);
        mbar_init(&bar_o[i], 1);
        mbar_init(&bar_ofree[i], FF_DRAIN_WARPS);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // the set-up above is independent of the preceding kernel; its outputs are read (and buffers written) below
  auto par = [](int u) -> uint32_t { return static_cast<uint32_t>(u >> 1) & 1u; };  // phase parity of unit u's barriers

  // setmaxnreg sits INSIDE each side of the role dispatch: ptxas budgets registers for the code a setmaxnreg dominates
  if (warp < FF_ENGINE_WARPS) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FF_REGS_ENGINE));
    // ---------------------------------------------------------------------------------- softmax engine
    const int quad = warp & 3;
    const float c2 = scale * FF_LOG2E;
    auto engine = [&](auto half_tag) {
      constexpr int half = decltype(half_tag)::value;
      const int my16_begin = half ? h0_16 : 0;
      const int my_n16 = half ? n16 - h0_16 : h0_16;           // 16-column chunks of this warp (0..7)
      const int col0 = my16_begin * 16;                        // first score column of this warp
      // the only chunk that can hold padded keys (exact zeros, to be excluded) is the globally last one
      const int last_j = (n16 - 1) - my16_begin;               // its index in this warp (out of range for the other half)
      const int last_valid = N - (n16 - 1) * 16;               // valid keys in it (1..16)
      int u = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        for (int t = 0; t < nqt; ++t, ++u) {
          const int b = u & 1;
          const uint32_t lane_addr = tmem_base + b * 256 + (static_cast<uint32_t>(quad * 32) << 16);
          float* stat = sStat + b * 5 * 128;
          const int r = quad * 32 + lane;                       // row within the tile
          const bool warp_active = t * 128 + quad * 32 < Nq;    // warps whose 32 rows are all padding only keep the barriers moving
          mbar_wait(&bar_s[b], par(u));
          tc_fence_after();
          if (quad == 0 && lane == 0) FF_STAMP(u, 2 + 5 * half);
          uint32_t s[FF_PAIRS][16];
          if (warp_active) {
#pragma unroll
            for (int j = 0; j < FF_PAIRS; ++j)
              if (j < my_n16) tmem_ld_32x32b_x16(lane_addr + col0 + 16 * j, s[j]);
            tmem_ld_wait();
          }
          if (quad == 0 && lane == 0) FF_STAMP(u, 3 + 5 * half);
          if (warp_active) {
            float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
            for (int j = 0; j < FF_PAIRS; ++j) {
              if (j < my_n16) {
                if (j == last_j) {
#pragma unroll
                  for (int i = 0; i < 16; ++i)
                    if (i < last_valid) m0 = fmaxf(m0, __uint_as_float(s[j][i]));
                } else {
#pragma unroll
                  for (int i = 0; i < 16; i += 8) {
                    m0 = fmaxf(m0, fmaxf(__uint_as_float(s[j][i]), __uint_as_float(s[j][i + 1])));
                    m1 = fmaxf(m1, fmaxf(__uint_as_float(s[j][i + 2]), __uint_as_float(s[j][i + 3])));
                    m2 = fmaxf(m2, fmaxf(__uint_as_float(s[j][i + 4]), __uint_as_float(s[j][i + 5])));
                    m3 = fmaxf(m3, fmaxf(__uint_as_float(s[j][i + 6]), __uint_as_float(s[j][i + 7])));
                  }
                }
              }
            }
            // the two column halves of a row meet here: partial maxima through shared memory, one 64-thread barrier per
            // quadrant. Past it both warps of the quadrant hold their scores in registers, so P may overwrite any column.
            stat[half * 128 + r] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
            const float mx = fmaxf(stat[r], stat[128 + r]) * c2;
            if (quad == 0 && lane == 0) FF_STAMP(u, 4 + 5 * half);
            const float neg_mx = -mx;
            float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
            for (int j = 0; j < FF_PAIRS; ++j) {
              if (j < my_n16) {
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  float p0 = ex2_approx(fmaf(__uint_as_float(s[j][2 * i]), c2, neg_mx));
                  float p1 = ex2_approx(fmaf(__uint_as_float(s[j][2 * i + 1]), c2, neg_mx));
                  if (j == last_j) {
                    if (2 * i >= last_valid) p0 = 0.f;
                    if (2 * i + 1 >= last_valid) p1 = 0.f;
                  }
                  sum0 += p0;
                  sum1 += p1;
                  w[i] = pack_bf16x2(p0, p1);
                }
                tmem_st_32x32b_x8(lane_addr + ((col0 + 16 * j) >> 1), w);
              }
            }
            tmem_st_wait();
            if (quad == 0 && lane == 0) FF_STAMP(u, 5 + 5 * half);
            // the statistics slots of this parity were last read by the drain of unit u-2
            if (u >= 2) mbar_wait(&bar_ofree[b], par(u - 2));
            stat[(2 + half) * 128 + r] = sum0 + sum1;
            if (half == 0) stat[4 * 128 + r] = mx;
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_p[b]);
          if (quad == 0 && lane == 0) FF_STAMP(u, 6 + 5 * half);
        }
      }
    };
    if (warp < 4) engine(std::integral_constant<int, 0>{});
    else engine(std::integral_constant<int, 1>{});
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FF_REGS_OTHER));
    if (warp == FF_ISSUE_WARP) {
      if (elect_one()) {
        // ------------------------------------------------------------------------------ TMA + MMA issue thread
        const uint32_t idesc_s = make_idesc_bf16(128, nk, false, false);
        constexpr uint32_t idesc_o = make_idesc_bf16(128, FF_HD, false, true);
        auto issue_loads = [&](int item, int stage) {
          const int h = item % H, b = item / H;
          uint8_t* st = smem + stage * FF_STAGE_BYTES;
          mbar_arrive_expect_tx(&bar_ld[stage], FF_STAGE_BYTES);
          tma_load_3d(st + FF_Q_BYTES, &tm_k, &bar_ld[stage], h * FF_HD, 0, b);
          tma_load_3d(st, &tm_q, &bar_ld[stage], h * FF_HD, 0, b);
          tma_load_3d(st + FF_Q_BYTES + FF_KV_BYTES, &tm_v, &bar_ld[stage], h * FF_HD, 0, b);
        };
        // P V of unit u: P from the head of TMEM buffer u & 1, V from operand stage `stage`, O into the buffer's tail
        auto issue_pv = [&](int u, int stage) {
          const int b = u & 1;
          mbar_wait(&bar_p[b], par(u));
          tc_fence_after();
          FF_STAMP(u, 1);
          const uint64_t dv = make_smem_desc_sw128(smem_u32(smem + stage * FF_STAGE_BYTES) + FF_Q_BYTES + FF_KV_BYTES, 8192, 1024);
          for (int k = 0; k < n16; ++k)
            umma_bf16_ts(tmem_base + b * 256 + FF_O_COL, tmem_base + b * 256 + 8 * k, dv + 128 * k, idesc_o, k > 0 ? 1u : 0u);
          umma_commit(&bar_o[b]);
        };
        issue_loads(blockIdx.x, 0);
        int u = 0;
        int prev_stage = 0;
        for (int it = 0, item = blockIdx.x; item < num_items; ++it, item += gridDim.x) {
          const int stage = it & 1;
          const uint32_t aQ = smem_u32(smem + stage * FF_STAGE_BYTES), aK = aQ + FF_Q_BYTES;
          const uint64_t dk = make_smem_desc_sw128(aK, 0, 1024);
          const bool has_next = item + static_cast<int>(gridDim.x) < num_items;
          for (int t = 0; t < nqt; ++t, ++u) {
            const int b = u & 1;
            // buffer b is free for S(u) once O(u-2) has been drained out of its tail (P V(u-2) then has completed too)
            if (u >= 2) mbar_wait(&bar_ofree[b], par(u - 2));
            if (t == 0) mbar_wait(&bar_ld[stage], (it >> 1) & 1);
            tc_fence_after();
            const uint64_t dq = make_smem_desc_sw128(aQ + t * 16384, 0, 1024);
#pragma unroll
            for (int k = 0; k < FF_HD / 16; ++k) umma_bf16_ss(tmem_base + b * 256, dq + 2 * k, dk + 2 * k, idesc_s, k > 0 ? 1u : 0u);
            umma_commit(&bar_s[b]);
            FF_STAMP(u, 0);
            // Prefetch the next item into the other operand stage once every MMA that read it has completed. Two tiles
            // per item: those are units u-3 and u-2, and the drain of unit u-2 was awaited above. One tile per item: the
            // reader is unit u-1, whose P V is only issued below -- prefetch after it has completed.
            if (nqt == 2 && t == 1 && has_next) issue_loads(item + gridDim.x, stage ^ 1);
            if (u >= 1) issue_pv(u - 1, t == 0 ? prev_stage : stage);
            if (nqt == 1 && has_next) {
              if (u >= 1) mbar_wait(&bar_o[b ^ 1], par(u - 1));
              issue_loads(item + gridDim.x, stage ^ 1);
            }
          }
          prev_stage = stage;
        }
        if (u >= 1) issue_pv(u - 1, prev_stage);
      }
      __syncwarp();
    } else if (warp < FF_ISSUE_WARP) {
      // ---------------------------------------------------------------------------------- drain warps
      const int quad = warp & 3;
      const uint32_t out_tile = smem_u32(sOut) + quad * 4096;
      int u = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int h = item % H, bi = item / H;
        for (int t = 0; t < nqt; ++t, ++u) {
          const int b = u & 1;
          const uint32_t o_addr = tmem_base + b * 256 + FF_O_COL + (static_cast<uint32_t>(quad * 32) << 16);
          const float* stat = sStat + b * 5 * 128;
          const int r = quad * 32 + lane;
          const int row = t * 128 + r;                          // query row within the item
          const bool warp_active = t * 128 + quad * 32 < Nq;
          mbar_wait(&bar_p[b], par(u));    // the engine's row statistics are visible
          mbar_wait(&bar_o[b], par(u));    // P V has completed
          tc_fence_after();
          if (quad == 0 && lane == 0) FF_STAMP(u, 12);
          uint32_t o0[32], o1[32];
          float mx = 0.f, l = 1.f;
          if (warp_active) {
            tmem_ld_32x32b_x32(o_addr, o0);
            tmem_ld_32x32b_x32(o_addr + 32, o1);
            l = stat[2 * 128 + r] + stat[3 * 128 + r];
            mx = stat[4 * 128 + r];
            tmem_ld_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_ofree[b]);  // O(u) has left TMEM and the unit's statistics have been read
          if (quad == 0 && lane == 0) FF_STAMP(u, 13);
          if (warp_active) {
            const float inv_l = ff_rcp(l);
            if (lane == 0) tma_store_wait_read<0>();  // the previous unit's store has finished reading the staging tile
            __syncwarp();
            const uint32_t base = out_tile + lane * 128;
#pragma unroll
            for (int i = 0; i < 4; ++i) {   // 16-byte chunks 0-3 of the row (columns 0-31), XOR-swizzled by row
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                w[e] = pack_bf16x2(__uint_as_float(o0[8 * i + 2 * e]) * inv_l, __uint_as_float(o0[8 * i + 2 * e + 1]) * inv_l);
              ff_st_shared_v4(base + ((i ^ (lane & 7)) << 4), w[0], w[1], w[2], w[3]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {   // chunks 4-7 (columns 32-63)
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                w[e] = pack_bf16x2(__uint_as_float(o1[8 * i + 2 * e]) * inv_l, __uint_as_float(o1[8 * i + 2 * e + 1]) * inv_l);
              ff_st_shared_v4(base + (((4 + i) ^ (lane & 7)) << 4), w[0], w[1], w[2], w[3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d_addr(&tm_o, out_tile, h * FF_HD, t * 128 + quad * 32, bi);
              tma_store_commit();
            }
            if (lse != nullptr && row < Nq) lse[static_cast<long long>(item) * Nq + row] = (mx + ff_lg2(l)) * FF_LN2;
          }
          if (quad == 0 && lane == 0) FF_STAMP(u, 14);
        }
      }
      if (lane == 0) tma_store_wait_read<0>();  // the staging tile must outlive the last TMA store
    }
  }  // roles other than the engine warps
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == FF_ISSUE_WARP) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// q/k/v: [B*N, ...] with row pitch ld, head h at column h*64; o: [B*N, H*64] pitch ldo; lse: [B, H, Nq] or NULL.
// Only the first Nq tokens of every image act as queries (their rows of o / lse are written); all N are keys.
int attention_fwd_fused(const void* q, const void* k, const void* v, long long ld, void* o, long long ldo, float* lse, int B,
                        int N, int Nq, int H, float scale, cudaStream_t stream) {
  if (N > FF_KV) return set_error(kErrUnsupported, "attention_fwd_fused: N=%d > %d", N, FF_KV);
  if (Nq <= 0 || Nq > N) return set_error(kErrInvalidArg, "attention_fwd_fused: Nq=%d must be in [1, N=%d]", Nq, N);
  CUtensorMap tq, tk, tv, to;
  const uint64_t D = static_cast<uint64_t>(H) * FF_HD;
  int rc = encode_tmap_3d_bf16(&tq, q, D, Nq, B, ld, static_cast<uint64_t>(N) * ld, 64, 256);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&tk, k, D, N, B, ld, static_cast<uint64_t>(N) * ld, 64, FF_KV);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&tv, v, D, N, B, ld, static_cast<uint64_t>(N) * ld, 64, FF_KV);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&to, o, D, Nq, B, ldo, static_cast<uint64_t>(N) * ldo, 64, 32);  // one warp's 32-row tile
  if (rc) return rc;
  auto* kernel = (N + 15) / 16 == 13 ? attn_fwd_fused_kernel<13> : attn_fwd_fused_kernel<0>;
  if (int rc2 = ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), FF_SMEM, "attention_fwd_fused")) return rc2;
  const int num_sms = device_sm_count();
  const int items = B * H;
  dim3 grid(items < num_sms ? items : num_sms);
  long long* trace = nullptr;
#ifdef TIC_ATTN_TRACE
  cudaMallocManaged(&trace, 64 * sizeof(long long));
  for (int i = 0; i < 64; ++i) trace[i] = 0;
#endif
  launch_pdl(kernel, grid, dim3(FF_THREADS), FF_SMEM, stream, tq, tk, tv, to, lse, N, Nq, H, items, scale, trace);
#ifdef TIC_ATTN_TRACE
  cudaDeviceSynchronize();
  {
    static const char* names[16] = {"S_issued", "PV_issue", "h0_S_ready", "h0_loaded", "h0_max", "h0_p2_done", "h0_P_arrive",
                                    "h1_S_ready", "h1_loaded", "h1_max", "h1_p2_done", "h1_P_arrive", "drain_o_ready",
                                    "drain_ofree", "drain_done", ""};
    const long long t0 = trace[0];
    for (int u = 0; u < 4; ++u) {
      fprintf(stderr, "[ff trace] unit %d:", 8 + u);
      for (int i = 0; i < 15; ++i) fprintf(stderr, " %s=%lld", names[i], trace[u * 16 + i] - t0);
      fprintf(stderr, "\n");
    }
    cudaFree(trace);
  }
#endif
  return check_launch("attention_fwd_fused");
}

}  // namespace tic
