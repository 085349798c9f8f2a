// Fused multi-head attention FORWARD for sequences that need more than one key block (N > 224; the 577-token
// ViT-L/16 384x384 case) as a persistent, warp-specialised kernel with online softmax. Replaces
// F.scaled_dot_product_attention (modeling_vit.py:232-246 [a6]), non-causal, no mask, dropout 0, head_dim 64.
//
// One CTA per SM walks the work items; an item is (image, head, PAIR of 128-query tiles). Both tiles see the same
// K / V blocks (128 keys each), which stream through a 3-stage TMA ring, and each tile has its own softmax
// warpgroup, so the exp2 pass of one tile runs while the tensor core computes the other tile's products (the two
// groups share each SM sub-partition's MUFU, which is what bounds this kernel):
//   warps 0-3   softmax group 0: one query row per thread (TMEM lane == row) of the item's first tile
//   warps 4-7   softmax group 1: the second tile. When the pair has only one tile (the last tile of an odd count,
//               or the CLS-only last layer) the two groups split that tile's KEY blocks instead (even / odd), each
//               with its own running (reference, sum, O), and group 0 merges the two partial results at the end.
//   warp  8     one elected thread: TMA producer (Q tiles of the NEXT item are prefetched, K / V ring)
//   warp  9     one elected thread: every tcgen05.mma, in the order
//                   S0(0) S1(0) | PV0(0) S0(1) PV1(0) S1(1) | PV0(1) S0(2) ...
//               S_t(j) = Q_t K_j^T (SS, M=128, N=keys of the block), O_t += P_t(j) V_j (TS, P read from TMEM as packed
//               bf16 over the S columns it came from). The tensor pipe executes in issue order, so S_t(j+1) may
//               overwrite the columns P_t(j) was read from, and a softmax group that sees S_t(j) complete knows that
//               PV_t(j-1) has completed as well -- it may rescale O_t in place without another barrier.
// Online softmax with a LAZY reference maximum: a row keeps exponentiating against the reference it already has while
// the running maximum stays within 2^8 of it (p <= 256: harmless in bf16 / fp32) and rescales O and l only when a warp
// sees a larger jump, which after the first block is rare. The result O / l and logsumexp = ref + log2(l) do not depend
// on the reference chosen.
// TMEM (512 columns): S0 / P0 [0,128) | S1 / P1 [128,256) | O0 [256,320) | O1 [320,384).
#include "tic_internal.cuh"

namespace tic {
namespace {

constexpr int FL_THREADS = 320;
constexpr int FL_HD = 64;
constexpr int FL_QT = 128;  // queries per tile (UMMA M)
constexpr int FL_KB = 128;  // keys per block
constexpr int FL_KV_STAGES = 3;
constexpr int FL_TILE_BYTES = FL_QT * 128;             // one Q tile / K block / V block: 16 KB
constexpr int FL_Q_STAGE_BYTES = 2 * FL_TILE_BYTES;    // both query tiles of an item
constexpr int FL_KV_STAGE_BYTES = 2 * FL_TILE_BYTES;   // K block + V block
constexpr int FL_OUT_BYTES = 8 * 4096;                 // per softmax warp: 32 rows x 128 B
constexpr int FL_SMEM_USED = 2 * FL_Q_STAGE_BYTES + FL_KV_STAGES * FL_KV_STAGE_BYTES + FL_OUT_BYTES + 256 + 1024;
constexpr int FL_SMEM = FL_SMEM_USED + 1024;
constexpr uint32_t FL_COL_S = 0, FL_COL_O = 256;
constexpr float FL_LOG2E = 1.4426950408889634f;
constexpr float FL_LN2 = 0.6931471805599453f;
constexpr float FL_RESCALE_THRESHOLD = 8.0f;  // log2 domain

TIC_DEVINL void fl_st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__global__ void __launch_bounds__(FL_THREADS, 1)
attn_fwd_long_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                     float* __restrict__ lse, int N, int Nq, int H, int num_items, float scale) {
  // N = keys per image; Nq = queries per image (the first Nq tokens; Nq = 1 for the CLS-only last layer)
  pdl_launch_dependents();
  extern __shared__ uint8_t fl_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fl_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                     // [2 stages][2 tiles][128 rows][128 B]
  uint8_t* sKV = sQ + 2 * FL_Q_STAGE_BYTES;               // [3 stages][K block | V block]
  uint8_t* sOut = sKV + FL_KV_STAGES * FL_KV_STAGE_BYTES;  // [8 warps][32 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + FL_OUT_BYTES);
  uint64_t* q_full = bars + 0;    // [2]
  uint64_t* q_empty = bars + 2;   // [2]
  uint64_t* k_full = bars + 4;    // [3]
  uint64_t* v_full = bars + 7;    // [3]
  uint64_t* kv_empty = bars + 10;  // [3]
  uint64_t* s_full = bars + 13;   // [2] S of tile t is in TMEM
  uint64_t* p_full = bars + 15;   // [2] P of tile t written (4 warp arrivals)
  uint64_t* o_full = bars + 17;   // [2] the last P V of tile t has completed
  uint64_t* o_empty = bars + 19;  // [2] the tile's warps have read O out of TMEM (4 warp arrivals)
  uint64_t* merge_full = bars + 21;  // group 1 has published its partial (reference, sum) of a split tile (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
  float* sM = reinterpret_cast<float*>(bars + 32);  // [2][128] group 1's reference / sum per row of a split tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqt = (Nq + FL_QT - 1) / FL_QT;   // query tiles per (image, head)
  const int npairs = (nqt + 1) >> 1;          // items per (image, head)
  const int nb = (N + FL_KB - 1) / FL_KB;     // key blocks

  if (warp == 9) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&q_full[i], 1);
        mbar_init(&q_empty[i], 1);
        mbar_init(&s_full[i], 1);
        mbar_init(&p_full[i], 4);
        mbar_init(&o_full[i], 1);
        mbar_init(&o_empty[i], 4);
      }
      for (int i = 0; i < FL_KV_STAGES; ++i) {
        mbar_init(&k_full[i], 1);
        mbar_init(&v_full[i], 1);
        mbar_init(&kv_empty[i], 1);
      }
      mbar_init(merge_full, 4);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_o);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // the set-up above is independent of the preceding kernel; its outputs are read (and buffers written) below

  if (warp == 8) {
    if (elect_one()) {
      // -------------------------------------------------------------------------------------- TMA producer
      int kvc = 0;  // K / V blocks loaded so far (ring position)
      for (int it = 0, item = blockIdx.x; item < num_items; ++it, item += gridDim.x) {
        const int bh = item / npairs, pair = item - bh * npairs;
        const int h = bh % H, b = bh / H;
        const int qs = it & 1;
        const bool has1 = 2 * pair + 1 < nqt;
        mbar_wait(&q_empty[qs], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[qs], has1 ? 2 * FL_TILE_BYTES : FL_TILE_BYTES);
        tma_load_3d(sQ + qs * FL_Q_STAGE_BYTES, &tm_q, &q_full[qs], h * FL_HD, 2 * pair * FL_QT, b);
        if (has1) tma_load_3d(sQ + qs * FL_Q_STAGE_BYTES + FL_TILE_BYTES, &tm_q, &q_full[qs], h * FL_HD, (2 * pair + 1) * FL_QT, b);
        for (int j = 0; j < nb; ++j, ++kvc) {
          const int slot = kvc % FL_KV_STAGES;
          mbar_wait(&kv_empty[slot], ((kvc / FL_KV_STAGES) & 1) ^ 1);
          uint8_t* st = sKV + slot * FL_KV_STAGE_BYTES;
          mbar_arrive_expect_tx(&k_full[slot], FL_TILE_BYTES);
          tma_load_3d(st, &tm_k, &k_full[slot], h * FL_HD, j * FL_KB, b);
          mbar_arrive_expect_tx(&v_full[slot], FL_TILE_BYTES);
          tma_load_3d(st + FL_TILE_BYTES, &tm_v, &v_full[slot], h * FL_HD, j * FL_KB, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    if (elect_one()) {
      // -------------------------------------------------------------------------------------- MMA issue thread
      constexpr uint32_t idesc_o = make_idesc_bf16(FL_QT, FL_HD, false, true);
      int kvc = 0;
      uint32_t np0 = 0, np1 = 0;      // P hand-overs consumed per softmax group (barrier phase = count & 1)
      uint32_t items0 = 0, items1 = 0;  // items in which each group was active so far
      for (int it = 0, item = blockIdx.x; item < num_items; ++it, item += gridDim.x) {
        const int bh = item / npairs, pair = item - bh * npairs;
        const int qs = it & 1;
        const bool has1 = 2 * pair + 1 < nqt;
        const uint32_t aQ = smem_u32(sQ + qs * FL_Q_STAGE_BYTES);
        const uint64_t dq0 = make_smem_desc_sw128(aQ, 0, 1024), dq1 = make_smem_desc_sw128(aQ + FL_TILE_BYTES, 0, 1024);
        auto keys_of = [&](int j) { return (min(FL_KB, N - j * FL_KB) + 15) & ~15; };  // padded keys are zero rows
        auto slot_of = [&](int j) { return (kvc + j) % FL_KV_STAGES; };
        auto phase_of = [&](int j) { return static_cast<uint32_t>(((kvc + j) / FL_KV_STAGES) & 1); };
        // S of key block j -> score region g, queries from descriptor dq
        auto issue_s = [&](int g, uint64_t dq, int j) {
          const uint32_t idesc_s = make_idesc_bf16(FL_QT, keys_of(j), false, false);
          const uint64_t dk = make_smem_desc_sw128(smem_u32(sKV + slot_of(j) * FL_KV_STAGE_BYTES), 0, 1024);
          mbar_wait(&k_full[slot_of(j)], phase_of(j));
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < FL_HD / 16; ++k) umma_bf16_ss(tmem_base + FL_COL_S + g * 128, dq + 2 * k, dk + 2 * k, idesc_s, k > 0 ? 1u : 0u);
          umma_commit(&s_full[g]);
        };
        // O of group g (+)= P(region g) V_j, once the group has written P and (first product of the item) read out its
        // previous O
        auto issue_pv = [&](int g, int j, bool first) {
          mbar_wait(&v_full[slot_of(j)], phase_of(j));
          if (g == 0) { mbar_wait(&p_full[0], np0 & 1); ++np0; }
          else        { mbar_wait(&p_full[1], np1 & 1); ++np1; }
          const uint32_t prev = g == 0 ? items0 : items1;
          if (first && prev > 0) mbar_wait(&o_empty[g], (prev - 1) & 1);
          tc_fence_after();
          const uint64_t dv = make_smem_desc_sw128(smem_u32(sKV + slot_of(j) * FL_KV_STAGE_BYTES + FL_TILE_BYTES), 8192, 1024);
          const int ksteps = keys_of(j) >> 4;
          for (int k = 0; k < ksteps; ++k)
            umma_bf16_ts(tmem_base + FL_COL_O + g * 64, tmem_base + FL_COL_S + g * 128 + 8 * k, dv + 128 * k, idesc_o,
                         (!first || k > 0) ? 1u : 0u);
        };
        mbar_wait(&q_full[qs], (it >> 1) & 1);
        if (has1) {
          // two query tiles, every key block goes to both groups
          issue_s(0, dq0, 0);
          issue_s(1, dq1, 0);
          for (int j = 0; j < nb; ++j) {
            issue_pv(0, j, j == 0);
            if (j == nb - 1) umma_commit(&o_full[0]);
            if (j + 1 < nb) issue_s(0, dq0, j + 1);
            issue_pv(1, j, j == 0);
            if (j == nb - 1) umma_commit(&o_full[1]);
            umma_commit(&kv_empty[slot_of(j)]);  // every product that reads K_j / V_j has been issued
            if (j + 1 < nb) issue_s(1, dq1, j + 1);
          }
        } else {
          // one query tile: the two groups split its KEY blocks (even / odd) and group 0 merges the two partial results
          issue_s(0, dq0, 0);
          if (nb > 1) issue_s(1, dq0, 1);
          for (int j = 0; j < nb; ++j) {
            const int g = j & 1;
            issue_pv(g, j, j < 2);
            if (j + 2 >= nb) umma_commit(&o_full[g]);
            umma_commit(&kv_empty[slot_of(j)]);
            if (j + 2 < nb) issue_s(g, dq0, j + 2);
          }
        }
        umma_commit(&q_empty[qs]);  // the last S products of this item have been issued
        kvc += nb;
        ++items0;
        if (has1 || nb > 1) ++items1;
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------------------------- softmax warps
    const int g = warp >> 2, quad = warp & 3;  // softmax group, TMEM lane quadrant
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t s_addr = lane_addr + FL_COL_S + g * 128, o_addr = lane_addr + FL_COL_O + g * 64;
    const float c2 = scale * FL_LOG2E;
    const uint32_t out_tile = smem_u32(sOut) + warp * 4096;
    uint32_t ns = 0, no = 0, nm = 0;  // S blocks / O tiles / merges seen by this group (barrier phases)
    for (int it = 0, item = blockIdx.x; item < num_items; ++it, item += gridDim.x) {
      const int bh = item / npairs, pair = item - bh * npairs;
      const int h = bh % H, b = bh / H;
      const bool lone = 2 * pair + 1 >= nqt;     // single query tile: the groups split the key blocks
      if (lone && g == 1 && nb < 2) continue;    // nothing for group 1 to do
      const int tile = lone ? 2 * pair : 2 * pair + g;
      const int row0 = tile * FL_QT + quad * 32;
      const int row = row0 + lane;
      const bool warp_active = row0 < Nq;  // a warp whose 32 rows are all padding only keeps the barriers moving
      float mref = 0.f, l = 0.f;
      // p = exp2(s * c2 - mref) over the block's columns from TMEM, packed bf16 P over the S columns (zeros past the last
      // key); returns the row sum
      auto exp_pass = [&](int nvalid) -> float {
        float sum = 0.f;
        const float neg_ref = -mref;
        const int nk = (nvalid + 15) & ~15;
#pragma unroll 1
        for (int c = 0; c < (nk + 31) >> 5; ++c) {
          uint32_t r[32], w[16];
          tmem_ld_32x32b_x32(s_addr + c * 32, r);
          tmem_ld_wait();
          if (c * 32 + 32 <= nvalid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * i]), c2, neg_ref));
              const float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * i + 1]), c2, neg_ref));
              sum += p0 + p1;
              w[i] = pack_bf16x2(p0, p1);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * i]), c2, neg_ref));
              float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * i + 1]), c2, neg_ref));
              if (c * 32 + 2 * i >= nvalid) p0 = 0.f;
              if (c * 32 + 2 * i + 1 >= nvalid) p1 = 0.f;
              sum += p0 + p1;
              w[i] = pack_bf16x2(p0, p1);
            }
          }
          tmem_st_32x32b_x16(s_addr + c * 16, w);
        }
        return sum;
      };
      // The running output of this group is rescaled in place (P V of the group's previous block has completed: it was
      // issued before this block's S). The TMEM accesses are warp-wide, the decision is per row (alpha = 1 for the rows
      // that keep their reference), so a row's result never depends on which other rows share its warp.
      auto rescale = [&](bool jump, float nref) {
        const float alpha = jump ? ex2_approx(mref - nref) : 1.0f;
        l *= alpha;
        if (jump) mref = nref;
#pragma unroll 1
        for (int c = 0; c < FL_HD / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(o_addr + c * 32, r);
          tmem_ld_wait();
          uint32_t w0[16], w1[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            w0[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            w1[i] = __float_as_uint(__uint_as_float(r[16 + i]) * alpha);
          }
          tmem_st_32x32b_x16(o_addr + c * 32, w0);
          tmem_st_32x32b_x16(o_addr + c * 32 + 16, w1);
        }
      };
      const int jstep = lone ? 2 : 1;
      for (int j = lone ? g : 0; j < nb; j += jstep) {
        const bool first = j < jstep;
        mbar_wait(&s_full[g], ns & 1);
        ++ns;
        tc_fence_after();
        if (warp_active) {
          const int nvalid = min(FL_KB, N - j * FL_KB);
          if (first || nvalid < FL_KB) {
            // two passes: the first block has no reference yet, a partial block needs its padded keys masked
            const int nchunk = (nvalid + 31) >> 5;
            float raw_max = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < nchunk; ++c) {
              uint32_t r[32];
              tmem_ld_32x32b_x32(s_addr + c * 32, r);
              tmem_ld_wait();
              if (c * 32 + 32 <= nvalid) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) raw_max = fmaxf(raw_max, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (c * 32 + i < nvalid) raw_max = fmaxf(raw_max, __uint_as_float(r[i]));
              }
            }
            const float mx = raw_max * c2;
            if (first) {
              mref = mx;
            } else if (const bool jump = mx > mref + FL_RESCALE_THRESHOLD; __any_sync(0xffffffffu, jump)) {
              rescale(jump, mx);
            }
            l += exp_pass(nvalid);
          } else {
            // one pass against the reference the row already has: exponentials go to registers while the block maximum is
            // tracked on the side; P is committed to TMEM only if no row of the warp jumped past the threshold
            uint32_t w[64];
            float sum = 0.f, xmax = -INFINITY;
            const float neg_ref = -mref;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t r[32];
              tmem_ld_32x32b_x32(s_addr + c * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float x0 = fmaf(__uint_as_float(r[2 * i]), c2, neg_ref);
                const float x1 = fmaf(__uint_as_float(r[2 * i + 1]), c2, neg_ref);
                xmax = fmaxf(xmax, fmaxf(x0, x1));
                const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
                sum += p0 + p1;
                w[c * 16 + i] = pack_bf16x2(p0, p1);
              }
            }
            const bool jump = xmax > FL_RESCALE_THRESHOLD;
            if (__any_sync(0xffffffffu, jump)) {
              rescale(jump, mref + xmax);   // S is still intact in TMEM: redo the block against the new reference
              l += exp_pass(nvalid);
            } else {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                uint32_t wc[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) wc[i] = w[c * 16 + i];
                tmem_st_32x32b_x16(s_addr + c * 16, wc);
              }
              l += sum;
            }
          }
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
      }
      // ---- O / l -> bf16 through this warp's staging tile and a TMA store; logsumexp for the backward pass
      mbar_wait(&o_full[g], no & 1);
      ++no;
      tc_fence_after();
      if (lone && g == 1) {
        // hand this group's partial (reference, sum; O1 stays in TMEM) to group 0, which owns the same TMEM lanes
        sM[quad * 32 + lane] = mref;
        sM[128 + quad * 32 + lane] = l;
        __syncwarp();
        if (lane == 0) mbar_arrive(merge_full);
        continue;
      }
      uint32_t packed[32];
      float lse_row = 0.f;
      if (lone && nb > 1) {
        mbar_wait(merge_full, nm & 1);
        ++nm;
        tc_fence_after();
        if (warp_active) {
          const float m1 = sM[quad * 32 + lane], l1 = sM[128 + quad * 32 + lane];
          const float m = fmaxf(mref, m1);
          const float a0 = ex2_approx(mref - m), a1 = ex2_approx(m1 - m);
          const float lt = l * a0 + l1 * a1;
          const float f0 = a0 / lt, f1 = a1 / lt;
          lse_row = (m + log2f(lt)) * FL_LN2;
#pragma unroll
          for (int c = 0; c < FL_HD / 32; ++c) {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32b_x32(o_addr + c * 32, r0);
            tmem_ld_32x32b_x32(o_addr + 64 + c * 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              packed[c * 16 + i] = pack_bf16x2(fmaf(__uint_as_float(r0[2 * i]), f0, __uint_as_float(r1[2 * i]) * f1),
                                               fmaf(__uint_as_float(r0[2 * i + 1]), f0, __uint_as_float(r1[2 * i + 1]) * f1));
          }
        }
      } else if (warp_active) {
        const float inv_l = 1.0f / l;
        lse_row = (mref + log2f(l)) * FL_LN2;
#pragma unroll
        for (int c = 0; c < FL_HD / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(o_addr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i)
            packed[c * 16 + i] = pack_bf16x2(__uint_as_float(r[2 * i]) * inv_l, __uint_as_float(r[2 * i + 1]) * inv_l);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&o_empty[g]);  // the next tile of this group may start accumulating
        if (lone && nb > 1) mbar_arrive(&o_empty[1]);  // ... and group 1's partial has been merged
      }
      if (warp_active) {
        if (lane == 0) tma_store_wait_read<0>();  // the previous store has finished reading the staging tile
        __syncwarp();
        const uint32_t base = out_tile + lane * 128;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          fl_st_shared_v4(base + ((i ^ (lane & 7)) << 4), packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d_addr(&tm_o, out_tile, h * FL_HD, row0, b);
          tma_store_commit();
        }
        if (lse != nullptr && row < Nq) lse[static_cast<long long>(bh) * Nq + row] = lse_row;
      }
    }
    if (lane == 0) tma_store_wait_read<0>();  // the staging tile must outlive the last TMA store
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 9) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// q/k/v: [B*N, ...] with row pitch ld, head h at column h*64; o: [B*N, H*64] pitch ldo; lse: [B, H, Nq] or NULL.
// Only the first Nq tokens of every image act as queries (their rows of o / lse are written); all N are keys.
int attention_fwd_long(const void* q, const void* k, const void* v, long long ld, void* o, long long ldo, float* lse, int B,
                       int N, int Nq, int H, float scale, cudaStream_t stream) {
  if (Nq <= 0 || Nq > N) return set_error(kErrInvalidArg, "attention_fwd_long: Nq=%d must be in [1, N=%d]", Nq, N);
  CUtensorMap tq, tk, tv, to;
  const uint64_t D = static_cast<uint64_t>(H) * FL_HD;
  int rc = encode_tmap_3d_bf16(&tq, q, D, Nq, B, ld, static_cast<uint64_t>(N) * ld, 64, FL_QT);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&tk, k, D, N, B, ld, static_cast<uint64_t>(N) * ld, 64, FL_KB);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&tv, v, D, N, B, ld, static_cast<uint64_t>(N) * ld, 64, FL_KB);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&to, o, D, Nq, B, ldo, static_cast<uint64_t>(N) * ldo, 64, 32);  // one warp's 32-row tile
  if (rc) return rc;
  if (int rc2 = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_fwd_long_kernel), FL_SMEM, "attention_fwd_long")) return rc2;
  const int nqt = (Nq + FL_QT - 1) / FL_QT;
  const int items = B * H * ((nqt + 1) / 2);
  const int num_sms = device_sm_count();
  dim3 grid(items < num_sms ? items : num_sms);
  launch_pdl(attn_fwd_long_kernel, grid, dim3(FL_THREADS), FL_SMEM, stream, tq, tk, tv, to, lse, N, Nq, H, items, scale);
  return check_launch("attention_fwd_long");
}

}  // namespace tic
