// Fused multi-head attention BACKWARD for sequences of 257..640 tokens (the 577-token ViT-L/16 384x384 case): ONE
// persistent kernel computes dQ, dK, dV (and the QKV bias gradient) of F.scaled_dot_product_attention
// (modeling_vit.py:232-246 [a6]) with every score tile computed exactly once -- 5 matmuls per tile, one exp2 pass --
// instead of the 7 matmuls / two passes of separate dQ and dK/dV kernels.
//
// Same roles, same tensor-memory map and same per-block arithmetic as attention_bwd_fused.cu (N <= 256); what changes
// is what is resident. A CTA owns an (image, head) and walks its 128-key tiles; a work item is (image, head, ONE key
// tile): that tile's K and V rows stay in shared memory (double buffered across items), the queries stream past it --
// Q and dO in 64-query blocks through a 4-stage TMA ring fed by a producer warp -- and
//        S^T  = K  Q^T,  dP^T = V dO^T            (SS, M=128 keys, N=64 queries)            -> TMEM S[buf], dP[buf]
//        dV  += P^T  dO,  dK += dS^T Q            (TS, A = P^T / dS^T bf16 in TMEM)          accumulate over ALL queries
//        dQp  = dS   K                            (SS, A = dS^T staged in smem, M=128 queries, K=128 keys)
// dK / dV of the tile are complete inside the item. dQ needs the sum over the head's key tiles, and tensor memory cannot
// hold dQ of a whole 577-token head next to the score buffers (5 x 64 + 128 + 256 columns). The drain warps therefore
// accumulate the per-tile products dQp in fp32 in a scratch slab PRIVATE to this CTA (640 x 64 floats = 160 KB, reused
// for every head the CTA processes, so all 148 slabs stay in L2 and never travel to HBM): first tile writes, middle tiles
// read-add-write, the last tile adds, scales, rounds to bf16 and stores dq through TMA (and takes the query-bias column
// sums). Each slab element is only ever touched by one thread, in program order: no atomics, deterministic.
// (The ring depth matters: a Q / dO block is reloaded when its products have completed and is needed again two blocks
// later, so with 3 stages the block period was pinned to the TMA load latency.)
// logsumexp is read and delta = rowsum(dO o O) is COMPUTED (from the head's rows of O and dO) once per head by two loader
// warps, one head ahead of the compute warps: there is no separate delta kernel on this path.
// TMEM (512 columns): S[2] 0-127 | dP[2] 128-255 | dV 256-319 | dK 320-383 | dQp[2 query tiles in flight] 384-511.
#include "tic_internal.cuh"

#ifdef TIC_EXP_NO_MMA  // development experiment: the elementwise / barrier chain alone (results are garbage)
#define umma_bf16_ss(...) ((void)0)
#define umma_bf16_ts(...) ((void)0)
#endif

namespace tic {
namespace {

constexpr int BL_THREADS = 512;  // 8 compute warps | MMA warp | 4 drain warps | TMA producer warp | 2 statistics loader warps
constexpr int BL_REGS_COMPUTE = 160, BL_REGS_OTHER = 96;  // see attention_bwd_fused.cu
constexpr int BL_HD = 64;
constexpr int BL_NQ_MAX = 640;   // queries per image this kernel has logsumexp / delta slots for
constexpr float BL_LOG2E = 1.4426950408889634f;
constexpr int BL_TILE_BYTES = 128 * 128;          // K or V tile: 16 KB
constexpr int BL_KV_BYTES = 2 * BL_TILE_BYTES;    // K | V of one item
constexpr int BL_BLOCK_BYTES = 64 * 128;          // Q or dO block: 8 KB
constexpr int BL_RING_STAGES = 4;
constexpr int BL_RING_BYTES = 2 * BL_BLOCK_BYTES;  // Q | dO
constexpr int BL_STAGE_BYTES = 2 * 128 * 128;      // one dS^T tile: 2 query chunks x 128 key rows x 128 B
constexpr int BL_OUT_BYTES = 4 * 2 * 2048;         // per drain warp: two 32-row x 64-byte tiles
constexpr int BL_SMEM_USED = 2 * BL_KV_BYTES + BL_RING_STAGES * BL_RING_BYTES + 2 * BL_STAGE_BYTES + BL_OUT_BYTES +
                             2 * 2 * BL_NQ_MAX * 4 + 256;
constexpr int BL_SMEM = BL_SMEM_USED + 1024;
static_assert(BL_SMEM <= 232448, "attention_bwd_long: shared memory budget");
constexpr uint32_t BL_COL_DP = 128, BL_COL_DV = 256, BL_COL_DK = 320, BL_COL_DQ = 384;

TIC_DEVINL void bl_st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
TIC_DEVINL float4 bl_ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// 32 lanes each hold v[0..31] (one row, 32 columns): returns, in lane i, the sum over the 32 rows of column i.
TIC_DEVINL float bl_warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = upper ? v[i] : v[i + s];
      const float keep = upper ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

__global__ void __launch_bounds__(BL_THREADS, 1)
attn_bwd_long_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                     const __grid_constant__ CUtensorMap tm_dq, const __grid_constant__ CUtensorMap tm_dk,
                     const __grid_constant__ CUtensorMap tm_dv, const float* __restrict__ lse,
                     const __nv_bfloat16* __restrict__ o_rows, long long ldo, const __nv_bfloat16* __restrict__ do_rows,
                     long long lddo, float* __restrict__ dq_scratch, float* __restrict__ bias_grad,
                     int bias_mask, int N, int Nq, int H, int num_heads, float scale, long long* __restrict__ trace) {
  // trace (development builds with -DTIC_ATTN_TRACE only, else NULL): clock64 stamps of CTA 0, item 3
#ifdef TIC_ATTN_TRACE
#define BL_STAMP(slot) do { if (trace != nullptr && blockIdx.x == 0 && it == 3) trace[slot] = clock64(); } while (0)
#else
#define BL_STAMP(slot) do { } while (0)
#endif
  // N = keys per image; Nq = queries per image (the first Nq tokens; lse / delta are [B, H, Nq])
  pdl_launch_dependents();
  extern __shared__ uint8_t bl_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(bl_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sKV = smem;                                         // [2][K tile | V tile]
  uint8_t* sRing = sKV + 2 * BL_KV_BYTES;                      // [3][Q block | dO block]
  uint8_t* sStage = sRing + BL_RING_STAGES * BL_RING_BYTES;    // [2 query tiles][2 chunks of 64 queries][128 key rows][128 B]
  uint8_t* sOut = sStage + 2 * BL_STAGE_BYTES;                 // [4 drain warps][2][32 rows][64 B], 64-byte swizzle
  float* sL = reinterpret_cast<float*>(sOut + BL_OUT_BYTES);   // [2 items][BL_NQ_MAX] logsumexp * log2(e), +inf past Nq
  float* sD = sL + 2 * BL_NQ_MAX;                              // [2 items][BL_NQ_MAX] delta
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 2 * BL_NQ_MAX);
  uint64_t* k_full = bars + 0;       // [2]
  uint64_t* v_full = bars + 2;       // [2]
  uint64_t* kv_empty = bars + 4;     // [2] every MMA of the item that used this K / V buffer has completed
  uint64_t* bar_s = bars + 6;        // [2] score tiles of a block are in TMEM
  uint64_t* bar_p = bars + 8;        // [2] P^T / dS^T of a block written (8 warp arrivals)
  uint64_t* bar_dq = bars + 10;      // [2] the dQ product of a query tile has completed
  uint64_t* bar_free_q = bars + 12;  // [2] the drain warps have read that dQ buffer out of TMEM (4 warp arrivals)
  uint64_t* bar_acc = bars + 14;     // dV / dK of the item are complete
  uint64_t* bar_free_vk = bars + 15;  // the drain warps have read dV / dK out of TMEM (4 warp arrivals)
  uint64_t* ld_full = bars + 16;     // [2] logsumexp / delta of an item are in shared memory (2 loader-warp arrivals)
  uint64_t* ld_empty = bars + 18;    // [2] the compute warps have finished the item that used them (8 warp arrivals)
  uint64_t* ring_full = bars + 20;                    // [BL_RING_STAGES]
  uint64_t* ring_empty = ring_full + BL_RING_STAGES;  // [BL_RING_STAGES]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ring_empty + BL_RING_STAGES);
  static_assert((20 + 2 * BL_RING_STAGES) * 8 + 4 <= 256, "attention_bwd_long: barrier region");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkt = (N + 127) >> 7;   // key tiles per (image, head) == items per (image, head)
  const int J = (Nq + 63) >> 6;     // query blocks of 64 per item
  const int nqt = (J + 1) >> 1;     // query tiles of 128 per item
  const int w_last = ((Nq - (J - 1) * 64) + 15) & ~15;  // width of the last query block (multiple of 16)
  // This CTA owns the (image, head) pairs blockIdx.x, blockIdx.x + gridDim.x, ...; item `it` = (it / nkt)-th of them, key
  // tile it % nkt.
  const int my_heads = (num_heads - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int num_its = my_heads * nkt;

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_do);
      tma_prefetch_desc(&tm_dq); tma_prefetch_desc(&tm_dk); tma_prefetch_desc(&tm_dv);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&k_full[i], 1);
        mbar_init(&v_full[i], 1);
        mbar_init(&kv_empty[i], 1);
        mbar_init(&bar_s[i], 1);
        mbar_init(&bar_p[i], 8);
        mbar_init(&bar_dq[i], 1);
        mbar_init(&bar_free_q[i], 4);
        mbar_init(&ld_full[i], 2);
        mbar_init(&ld_empty[i], 8);
      }
      for (int i = 0; i < BL_RING_STAGES; ++i) {
        mbar_init(&ring_full[i], 1);
        mbar_init(&ring_empty[i], 1);
      }
      mbar_init(bar_acc, 1);
      mbar_init(bar_free_vk, 4);
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

  // setmaxnreg sits INSIDE each side of the role dispatch: ptxas budgets registers for the code a setmaxnreg dominates
  // (placed before the dispatch, every role was compiled under the kernel-wide figure and the compute loop spilled)
  if (warp < 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(BL_REGS_COMPUTE));
    // -------------------------------------------------------------------------------------- compute warps
    const int quad = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float c2 = scale * BL_LOG2E;
    const int r = quad * 32 + lane;  // key row within the 128-row tile
    const uint32_t stage_row = smem_u32(sStage) + r * 128;
    const int sw = r & 7;
    const uint32_t aL = smem_u32(sL), aD = smem_u32(sD);
    uint32_t ph_s = 0;  // bit b = parity of the next completion of bar_s[b]
    int g0 = 0, tiles0 = 0;
    for (int it = 0; it < num_its; ++it) {
      const int kt = it % nkt, hd = it / nkt;
      const int sb = hd & 1;   // the statistics (logsumexp, delta) belong to the head: loaded once, used by its nkt items
      const uint32_t aLi = aL + sb * BL_NQ_MAX * 4, aDi = aD + sb * BL_NQ_MAX * 4;
      if (kt == 0) mbar_wait(&ld_full[sb], (hd >> 1) & 1);
      const bool quad_active = kt * 128 + quad * 32 < N;
      const bool row_valid = kt * 128 + r < N;
      for (int j = 0; j < J; ++j) {
        const int buf = (g0 + j) & 1;
        const int w = j == J - 1 ? w_last : 64;
        const int tile = tiles0 + (j >> 1);
        if (threadIdx.x == 0) BL_STAMP(3 * j);
        mbar_wait(&bar_s[buf], (ph_s >> buf) & 1);
        if (threadIdx.x == 0) BL_STAMP(3 * j + 1);
        ph_s ^= 1u << buf;
        tc_fence_after();
#ifdef TIC_EXP_NO_ELEMENTWISE  // development experiment: the tensor / barrier chain alone (results are garbage)
        if (false) {
#else
        if (quad_active && half * 32 < w) {
#endif
          uint32_t s[32], dp[32];
          tmem_ld_32x32b_x32(lane_addr + buf * 64 + half * 32, s);
          tmem_ld_32x32b_x32(lane_addr + BL_COL_DP + buf * 64 + half * 32, dp);
          tmem_ld_wait();
          const uint32_t L4 = aLi + (j * 64 + half * 32) * 4, D4 = aDi + (j * 64 + half * 32) * 4;
          uint32_t pw[16], dw[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 Lq = bl_ld_shared_f4(L4 + 16 * i), Dq = bl_ld_shared_f4(D4 + 16 * i);
            const float p0 = ex2_approx(fmaf(__uint_as_float(s[4 * i + 0]), c2, -Lq.x));
            const float p1 = ex2_approx(fmaf(__uint_as_float(s[4 * i + 1]), c2, -Lq.y));
            const float p2 = ex2_approx(fmaf(__uint_as_float(s[4 * i + 2]), c2, -Lq.z));
            const float p3 = ex2_approx(fmaf(__uint_as_float(s[4 * i + 3]), c2, -Lq.w));
            pw[2 * i] = pack_bf16x2(p0, p1);
            pw[2 * i + 1] = pack_bf16x2(p2, p3);
            dw[2 * i] = pack_bf16x2(p0 * (__uint_as_float(dp[4 * i + 0]) - Dq.x), p1 * (__uint_as_float(dp[4 * i + 1]) - Dq.y));
            dw[2 * i + 1] = pack_bf16x2(p2 * (__uint_as_float(dp[4 * i + 2]) - Dq.z), p3 * (__uint_as_float(dp[4 * i + 3]) - Dq.w));
          }
          tmem_st_32x32b_x16(lane_addr + buf * 64 + half * 32, pw);
          tmem_st_32x32b_x16(lane_addr + BL_COL_DP + buf * 64 + half * 32, dw);
          // dS^T row -> staging tile of this query tile (chunk = 64-query block), zero for key rows past N so that the
          // dQ product never sees a non-finite value against the zero-filled K rows
          const uint32_t dst = stage_row + (tile & 1) * BL_STAGE_BYTES + (j & 1) * 16384;
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            const uint32_t a = dst + (((half * 4 + pc) ^ sw) << 4);
            if (row_valid) bl_st_shared_v4(a, dw[4 * pc], dw[4 * pc + 1], dw[4 * pc + 2], dw[4 * pc + 3]);
            else bl_st_shared_v4(a, 0u, 0u, 0u, 0u);
          }
          tmem_st_wait();
          fence_proxy_async();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_p[buf]);
        if (threadIdx.x == 0) BL_STAMP(3 * j + 2);
      }
      if (kt == nkt - 1 && lane == 0) mbar_arrive(&ld_empty[sb]);  // this warp no longer reads the head's logsumexp / delta
      g0 += J;
      tiles0 += nqt;
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(BL_REGS_OTHER));
    if (warp == 13) {
      if (elect_one()) {
        // ------------------------------------------------------------------------------------ TMA producer
        int rc = 0;  // ring position (blocks loaded so far)
        for (int it = 0; it < num_its; ++it) {
          const int kt = it % nkt, bh = static_cast<int>(blockIdx.x) + (it / nkt) * static_cast<int>(gridDim.x);
          const int h = bh % H, b = bh / H;
          const int kvb = it & 1;
          uint8_t* kv = sKV + kvb * BL_KV_BYTES;
          mbar_wait(&kv_empty[kvb], ((it >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&k_full[kvb], BL_TILE_BYTES);
          tma_load_3d(kv, &tm_k, &k_full[kvb], h * BL_HD, kt * 128, b);
          mbar_arrive_expect_tx(&v_full[kvb], BL_TILE_BYTES);
          tma_load_3d(kv + BL_TILE_BYTES, &tm_v, &v_full[kvb], h * BL_HD, kt * 128, b);
          for (int j = 0; j < J; ++j, ++rc) {
            const int slot = rc % BL_RING_STAGES;
            mbar_wait(&ring_empty[slot], ((rc / BL_RING_STAGES) & 1) ^ 1);
            uint8_t* st = sRing + slot * BL_RING_BYTES;
            mbar_arrive_expect_tx(&ring_full[slot], BL_RING_BYTES);
            tma_load_3d(st, &tm_q, &ring_full[slot], h * BL_HD, j * 64, b);
            tma_load_3d(st + BL_BLOCK_BYTES, &tm_do, &ring_full[slot], h * BL_HD, j * 64, b);
          }
        }
      }
      __syncwarp();
    } else if (warp >= 14) {
      // -------------------------------------------------------------------------------------- statistics loaders
      // Per head, one head ahead of the compute warps: logsumexp (log2 domain) from global memory and
      // delta[q] = sum_d dO[q, d] * O[q, d], computed here from the head's rows of O and dO (eight lanes per 128-byte row,
      // four rows in flight per lane group) -- no separate delta kernel, no delta round trip through HBM.
      const int t = threadIdx.x - 14 * 32;  // 0..63
      const int grp = t >> 3, sub = t & 7;  // 8 row groups x 8 lanes
      const int num_hd = num_its / nkt;
      for (int hd = 0; hd < num_hd; ++hd) {
        const int bh = static_cast<int>(blockIdx.x) + hd * static_cast<int>(gridDim.x);
        const int h = bh % H, b = bh / H;
        const int sb = hd & 1;
        mbar_wait(&ld_empty[sb], ((hd >> 1) & 1) ^ 1);
        const float* lrow = lse + static_cast<long long>(bh) * Nq;
        for (int qi = t; qi < J * 64; qi += 64)
          sL[sb * BL_NQ_MAX + qi] = qi < Nq ? __ldg(lrow + qi) * BL_LOG2E : INFINITY;  // exp2(-inf) = 0 for padded queries
        const __nv_bfloat16* obase = o_rows + static_cast<long long>(b) * N * ldo + h * BL_HD + sub * 8;
        const __nv_bfloat16* gbase = do_rows + static_cast<long long>(b) * N * lddo + h * BL_HD + sub * 8;
        for (int base = 0; base < J * 64; base += 32) {
          uint4 ov[4], gv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int row = base + 8 * u + grp;
            ov[u] = gv[u] = make_uint4(0u, 0u, 0u, 0u);
            if (row < Nq) {
              ov[u] = __ldg(reinterpret_cast<const uint4*>(obase + static_cast<long long>(row) * ldo));
              gv[u] = __ldg(reinterpret_cast<const uint4*>(gbase + static_cast<long long>(row) * lddo));
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float d0 = bf16_lo(ov[u].x) * bf16_lo(gv[u].x), d1 = bf16_hi(ov[u].x) * bf16_hi(gv[u].x);
            d0 = fmaf(bf16_lo(ov[u].y), bf16_lo(gv[u].y), d0); d1 = fmaf(bf16_hi(ov[u].y), bf16_hi(gv[u].y), d1);
            d0 = fmaf(bf16_lo(ov[u].z), bf16_lo(gv[u].z), d0); d1 = fmaf(bf16_hi(ov[u].z), bf16_hi(gv[u].z), d1);
            d0 = fmaf(bf16_lo(ov[u].w), bf16_lo(gv[u].w), d0); d1 = fmaf(bf16_hi(ov[u].w), bf16_hi(gv[u].w), d1);
            float d = d0 + d1;
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            d += __shfl_xor_sync(0xffffffffu, d, 4);
            if (sub == 0) sD[sb * BL_NQ_MAX + base + 8 * u + grp] = d;   // rows past Nq: exact zeros
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ld_full[sb]);  // release semantics order the warp's stores before the arrival
      }
    } else if (warp == 8) {
      if (elect_one()) {  // one issuing thread on the uniform datapath
        // ------------------------------------------------------------------------------------ MMA issue loop
        constexpr uint32_t idesc_ts = make_idesc_bf16(128, BL_HD, false, true);
        constexpr uint32_t idesc_dq = make_idesc_bf16(128, BL_HD, true, true);
        const uint32_t aRing = smem_u32(sRing), aS = smem_u32(sStage);
        uint32_t ph_p = 0;  // bit b = parity of the next completion of bar_p[b]
        int rc0 = 0;        // ring position of the item's first block
        int g0 = 0;         // blocks processed before this item (score buffer = (g0 + j) & 1)
        int tiles = 0;      // dQ products issued so far (dQ buffer / staging tile = tiles & 1)
        for (int it = 0; it < num_its; ++it) {
          const int kt = it % nkt;
          const int kvb = it & 1;
          const uint32_t aK = smem_u32(sKV + kvb * BL_KV_BYTES), aV = aK + BL_TILE_BYTES;
          const uint64_t dK_ = make_smem_desc_sw128(aK, 0, 1024), dV_ = make_smem_desc_sw128(aV, 0, 1024);
          const uint64_t dK_mn = make_smem_desc_sw128(aK, 8192, 1024);
          auto issue_scores = [&](int j) {
            const int buf = (g0 + j) & 1, r = rc0 + j, slot = r % BL_RING_STAGES;
            const int w = j == J - 1 ? w_last : 64;
            const uint32_t idesc = make_idesc_bf16(128, w, false, false);
            const uint32_t aQ = aRing + slot * BL_RING_BYTES, aDO = aQ + BL_BLOCK_BYTES;
            const uint64_t dQ_ = make_smem_desc_sw128(aQ, 0, 1024), dO_ = make_smem_desc_sw128(aDO, 0, 1024);
            mbar_wait(&ring_full[slot], (r / BL_RING_STAGES) & 1);
            if (j == 0) mbar_wait(&k_full[kvb], (it >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < BL_HD / 16; ++k) umma_bf16_ss(tmem_base + buf * 64, dK_ + 2 * k, dQ_ + 2 * k, idesc, k > 0 ? 1u : 0u);
            if (j == 0) { mbar_wait(&v_full[kvb], (it >> 1) & 1); tc_fence_after(); }
#pragma unroll
            for (int k = 0; k < BL_HD / 16; ++k)
              umma_bf16_ss(tmem_base + BL_COL_DP + buf * 64, dV_ + 2 * k, dO_ + 2 * k, idesc, k > 0 ? 1u : 0u);
            umma_commit(&bar_s[buf]);
          };
          issue_scores(0);
          if (J > 1) issue_scores(1);
          for (int j = 0; j < J; ++j) {
            const int buf = (g0 + j) & 1, r = rc0 + j, slot = r % BL_RING_STAGES;
            BL_STAMP(64 + 4 * j);
            mbar_wait(&bar_p[buf], (ph_p >> buf) & 1);
            BL_STAMP(65 + 4 * j);
            ph_p ^= 1u << buf;
            tc_fence_after();
            if (j == 0 && it > 0) {  // the previous item's dV / dK have left TMEM
              mbar_wait(bar_free_vk, (it - 1) & 1);
              tc_fence_after();
            }
            const int ksteps = (j == J - 1 ? w_last : 64) >> 4;
            const uint32_t aQ = aRing + slot * BL_RING_BYTES, aDO = aQ + BL_BLOCK_BYTES;
            const uint64_t dO_mn = make_smem_desc_sw128(aDO, 8192, 1024);
            const uint64_t dQ_mn = make_smem_desc_sw128(aQ, 8192, 1024);
            for (int k = 0; k < ksteps; ++k) {  // dV += P^T dO
              const uint32_t a = tmem_base + buf * 64 + (k >> 1) * 32 + (k & 1) * 8;
              umma_bf16_ts(tmem_base + BL_COL_DV, a, dO_mn + 128 * k, idesc_ts, (j > 0 || k > 0) ? 1u : 0u);
            }
            for (int k = 0; k < ksteps; ++k) {  // dK += dS^T Q
              const uint32_t a = tmem_base + BL_COL_DP + buf * 64 + (k >> 1) * 32 + (k & 1) * 8;
              umma_bf16_ts(tmem_base + BL_COL_DK, a, dQ_mn + 128 * k, idesc_ts, (j > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&ring_empty[slot]);  // this block of Q / dO has been consumed once these complete
            // the P^T / dS^T columns of this buffer have been consumed (in issue order): refill it with the scores of
            // block j+2 before the dQ product, which the compute warps do not wait for
            BL_STAMP(66 + 4 * j);
            if (j + 2 < J) issue_scores(j + 2);
            BL_STAMP(67 + 4 * j);
            if ((j & 1) || j == J - 1) {  // dQ partial of this query tile: dS K over this item's key tile
              const int dqb = tiles & 1;
              if (tiles >= 2) {  // the drain warps have read the product issued two tiles ago out of this buffer
                mbar_wait(&bar_free_q[dqb], ((tiles >> 1) - 1) & 1);
                tc_fence_after();
              }
              const int kvalid = min(128, N - kt * 128);
              const int ks = (kvalid + 15) >> 4;
              const uint64_t dS_mn = make_smem_desc_sw128(aS + dqb * BL_STAGE_BYTES, 16384, 1024);
              for (int k = 0; k < ks; ++k)
                umma_bf16_ss(tmem_base + BL_COL_DQ + dqb * 64, dS_mn + 128 * k, dK_mn + 128 * k, idesc_dq, k > 0 ? 1u : 0u);
              umma_commit(&bar_dq[dqb]);
              ++tiles;
            }
            if (j == J - 1) {
              umma_commit(bar_acc);           // dV / dK of this key tile are complete
              umma_commit(&kv_empty[kvb]);    // and nothing reads this K / V buffer any more
            }
          }
          rc0 += J;
          g0 += J;
        }
      }
      __syncwarp();
    } else if (warp < 13) {
      // -------------------------------------------------------------------------------------- drain warps
      const int quad = warp & 3;  // warps 9, 10, 11, 12 -> TMEM lane quadrants 1, 2, 3, 0
      const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
      const uint32_t stage = smem_u32(sOut) + (warp - 9) * 4096;  // two 2 KB tiles: [32 rows][64 B], 64-byte swizzle
      const int r = quad * 32 + lane;
      // this CTA's private fp32 dQ slab: [query tile][quadrant][half][8 column groups][32 lanes][4 floats] -- every warp
      // access is 512 contiguous bytes, and an element is only ever touched by the thread that owns its (row, columns)
      float* slab = dq_scratch + static_cast<long long>(blockIdx.x) * (BL_NQ_MAX * BL_HD);
      int tiles = 0;
      for (int it = 0; it < num_its; ++it) {
        const int kt = it % nkt, bh = static_cast<int>(blockIdx.x) + (it / nkt) * static_cast<int>(gridDim.x);
        const int h = bh % H, b = bh / H;
        // 32 rows x 32 packed bf16 columns -> staging slot (64-byte swizzle); optional column sums into a bias gradient
        auto stage_packed = [&](int slot, const uint32_t (&pk)[16], bool valid, float* bias_dst, int half) {
          const uint32_t base = stage + slot * 2048 + lane * 64;
          const int x = (lane >> 1) & 3;
#pragma unroll
          for (int pc = 0; pc < 4; ++pc)
            bl_st_shared_v4(base + ((pc ^ x) << 4), pk[4 * pc], pk[4 * pc + 1], pk[4 * pc + 2], pk[4 * pc + 3]);
          if (bias_dst != nullptr) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              v[2 * i] = valid ? bf16_lo(pk[i]) : 0.f;
              v[2 * i + 1] = valid ? bf16_hi(pk[i]) : 0.f;
            }
            const float cs = bl_warp_colsum32(v, lane);
            atomicAdd(bias_dst + h * BL_HD + half * 32 + lane, cs);
          }
        };
        // One half tile (32 rows x 32 fp32 accumulator columns of this warp's lane quadrant): TMEM -> scaled, packed bf16
        auto stage_half = [&](int slot, uint32_t col, float f, bool valid, float* bias_dst, int half) {
          uint32_t rr[32];
          tmem_ld_32x32b_x32(lane_addr + col, rr);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(__uint_as_float(rr[2 * i]) * f, __uint_as_float(rr[2 * i + 1]) * f);
          stage_packed(slot, pk, valid, bias_dst, half);
        };
        const bool first_kt = kt == 0, last_kt = kt == nkt - 1;
        for (int qt = 0; qt < nqt; ++qt, ++tiles) {  // dQ contribution of each query tile as its product completes
          const int dqb = tiles & 1;
          const bool active = qt * 128 + quad * 32 < Nq;
          float4 acc_next[4];
          if (active && !first_kt) {  // this thread's slab values of the tile's first quarter: in flight under the wait
            const float4* sp0 = reinterpret_cast<const float4*>(slab + ((qt * 4 + quad) * 4) * 512) + lane;
#pragma unroll
            for (int i = 0; i < 4; ++i) acc_next[i] = __ldcg(sp0 + i * 32);
          }
          if (warp == 12 && lane == 0) BL_STAMP(112 + 2 * qt);
          mbar_wait(&bar_dq[dqb], (tiles >> 1) & 1);
          if (warp == 12 && lane == 0) BL_STAMP(113 + 2 * qt);
          tc_fence_after();
          if (last_kt) {
            if (lane == 0) tma_store_wait_read<0>();  // this warp's staging slots are free again
            __syncwarp();
          }
          if (active) {
            float* qdst = (last_kt && (bias_mask & 1)) ? bias_grad : nullptr;
            const bool row_ok = qt * 128 + r < Nq;
            // 16 accumulator columns at a time (the drain warps live within 88 registers); the slab loads of a quarter are
            // issued one quarter ahead (the first before the product is even complete), so their L2 latency is hidden
#pragma unroll 1
            for (int qr = 0; qr < 4; ++qr) {
              float4* sp = reinterpret_cast<float4*>(slab + ((qt * 4 + quad) * 4 + qr) * 512) + lane;
              float4 acc[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[i] = acc_next[i];
              if (!first_kt && qr < 3) {
#pragma unroll
                for (int i = 0; i < 4; ++i) acc_next[i] = __ldcg(sp + 128 + i * 32);
              }
              uint32_t rr[16];
              tmem_ld_32x32b_x16(lane_addr + BL_COL_DQ + dqb * 64 + qr * 16, rr);
              tmem_ld_wait();
              if (!first_kt) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  rr[4 * i + 0] = __float_as_uint(__uint_as_float(rr[4 * i + 0]) + acc[i].x);
                  rr[4 * i + 1] = __float_as_uint(__uint_as_float(rr[4 * i + 1]) + acc[i].y);
                  rr[4 * i + 2] = __float_as_uint(__uint_as_float(rr[4 * i + 2]) + acc[i].z);
                  rr[4 * i + 3] = __float_as_uint(__uint_as_float(rr[4 * i + 3]) + acc[i].w);
                }
              }
              if (!last_kt) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  __stcg(sp + i * 32, make_float4(__uint_as_float(rr[4 * i]), __uint_as_float(rr[4 * i + 1]),
                                                  __uint_as_float(rr[4 * i + 2]), __uint_as_float(rr[4 * i + 3])));
              } else {
                // scale, round to bf16, stage: quarter qr = 16-byte chunks (qr & 1) * 2 + {0, 1} of slot qr >> 1
                const uint32_t p0 = pack_bf16x2(__uint_as_float(rr[0]) * scale, __uint_as_float(rr[1]) * scale);
                const uint32_t p1 = pack_bf16x2(__uint_as_float(rr[2]) * scale, __uint_as_float(rr[3]) * scale);
                const uint32_t p2 = pack_bf16x2(__uint_as_float(rr[4]) * scale, __uint_as_float(rr[5]) * scale);
                const uint32_t p3 = pack_bf16x2(__uint_as_float(rr[6]) * scale, __uint_as_float(rr[7]) * scale);
                const uint32_t p4 = pack_bf16x2(__uint_as_float(rr[8]) * scale, __uint_as_float(rr[9]) * scale);
                const uint32_t p5 = pack_bf16x2(__uint_as_float(rr[10]) * scale, __uint_as_float(rr[11]) * scale);
                const uint32_t p6 = pack_bf16x2(__uint_as_float(rr[12]) * scale, __uint_as_float(rr[13]) * scale);
                const uint32_t p7 = pack_bf16x2(__uint_as_float(rr[14]) * scale, __uint_as_float(rr[15]) * scale);
                const uint32_t base = stage + (qr >> 1) * 2048 + lane * 64;
                const int x = (lane >> 1) & 3, pc = (qr & 1) * 2;
                bl_st_shared_v4(base + ((pc ^ x) << 4), p0, p1, p2, p3);
                bl_st_shared_v4(base + (((pc + 1) ^ x) << 4), p4, p5, p6, p7);
                if (qdst != nullptr) {
                  float v[16];
                  const uint32_t pk[8] = {p0, p1, p2, p3, p4, p5, p6, p7};
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    v[2 * i] = row_ok ? bf16_lo(pk[i]) : 0.f;
                    v[2 * i + 1] = row_ok ? bf16_hi(pk[i]) : 0.f;
                  }
                  // recursive halving over 16 columns, then the two half-warps are added: lane i < 16 holds column i
#pragma unroll
                  for (int sft = 8; sft >= 1; sft >>= 1) {
                    const bool upper = (lane & sft) != 0;
#pragma unroll
                    for (int i = 0; i < sft; ++i) {
                      const float send = upper ? v[i] : v[i + sft];
                      const float keep = upper ? v[i + sft] : v[i];
                      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
                    }
                  }
                  const float cs = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
                  if (lane < 16) atomicAdd(qdst + h * BL_HD + qr * 16 + lane, cs);
                }
              }
            }
            if (last_kt) fence_proxy_async();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&bar_free_q[dqb]);
            if (active && last_kt) {
              tma_store_3d_addr(&tm_dq, stage, h * BL_HD, qt * 128 + quad * 32, b);
              tma_store_3d_addr(&tm_dq, stage + 2048, h * BL_HD + 32, qt * 128 + quad * 32, b);
              tma_store_commit();
            }
          }
        }
        {  // dV, then dK of the item's key tile (two staging slots: one tensor at a time)
          const bool quad_active = kt * 128 + quad * 32 < N;
          const int row0 = kt * 128 + quad * 32;
          const bool valid = kt * 128 + r < N;
          float* vdst = (bias_mask & 4) ? bias_grad + 2 * H * BL_HD : nullptr;
          float* kdst = (bias_mask & 2) ? bias_grad + H * BL_HD : nullptr;
          mbar_wait(bar_acc, it & 1);
          tc_fence_after();
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
          if (quad_active) {
            stage_half(0, BL_COL_DV, 1.0f, valid, vdst, 0);
            stage_half(1, BL_COL_DV + 32, 1.0f, valid, vdst, 1);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d_addr(&tm_dv, stage, h * BL_HD, row0, b);
              tma_store_3d_addr(&tm_dv, stage + 2048, h * BL_HD + 32, row0, b);
              tma_store_commit();
              tma_store_wait_read<0>();
            }
            __syncwarp();
            stage_half(0, BL_COL_DK, scale, valid, kdst, 0);
            stage_half(1, BL_COL_DK + 32, scale, valid, kdst, 1);
            fence_proxy_async();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bar_free_vk);  // the MMA thread may overwrite dV / dK
            if (quad_active) {
              tma_store_3d_addr(&tm_dk, stage, h * BL_HD, row0, b);
              tma_store_3d_addr(&tm_dk, stage + 2048, h * BL_HD + 32, row0, b);
              tma_store_commit();
            }
          }
        }
      }
      if (lane == 0) tma_store_wait_read<0>();  // the staging tiles must outlive the last TMA stores
    }
  }  // roles other than the compute warps
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// Scratch (in floats) the backward needs behind `delta`: delta [B, H, N] and, for N > 256, one private fp32 dQ slab
// (BL_NQ_MAX x 64) per SM of the current device.
long long attention_bwd_scratch_floats(int B, int N, int H) {
  const long long delta = (static_cast<long long>(B) * H * N + 63) / 64 * 64;
  if (N <= 256) return delta;
  return delta + static_cast<long long>(device_sm_count()) * BL_NQ_MAX * BL_HD;
}

// q/k/v: [B*N, ...] pitch ld, head h at column h*64; o / dout: [B*N, H*64] pitch ldo / lddo; dq/dk/dv pitch ldg. dq_scratch: fp32,
// (SM count) x BL_NQ_MAX x 64. Only the first Nq tokens of every image are queries.
// bias_grad (optional): fp32 [3*H*64] = q | k | v, ACCUMULATES the column sums of dq (bit 0) / dk (bit 1) / dv (bit 2).
int attention_bwd_long(const void* q, const void* k, const void* v, long long ld, const void* o, long long ldo,
                       const void* dout, long long lddo, const float* lse, float* dq_scratch, void* dq, void* dk, void* dv,
                       long long ldg, float* bias_grad, int bias_mask, int B, int N, int Nq, int H, float scale,
                       cudaStream_t stream) {
  if (Nq <= 0 || Nq > N) return set_error(kErrInvalidArg, "attention_bwd_long: Nq=%d must be in [1, N=%d]", Nq, N);
  if (Nq > BL_NQ_MAX) return set_error(kErrUnsupported, "attention_bwd_long: Nq=%d > %d", Nq, BL_NQ_MAX);
  CUtensorMap tq, tk, tv, tdo, tdq, tdk, tdv;
  const uint64_t D = static_cast<uint64_t>(H) * BL_HD;
  int rc = encode_tmap_3d_bf16(&tq, q, D, Nq, B, ld, static_cast<uint64_t>(N) * ld, 64, 64);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&tk, k, D, N, B, ld, static_cast<uint64_t>(N) * ld, 64, 128);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&tv, v, D, N, B, ld, static_cast<uint64_t>(N) * ld, 64, 128);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&tdo, dout, D, Nq, B, lddo, static_cast<uint64_t>(N) * lddo, 64, 64);
  if (rc) return rc;
  // outputs: 32-column x 32-row boxes (one drain warp's half tile), 64-byte swizzle
  rc = encode_tmap_3d_bf16_sw(&tdq, dq, D, Nq, B, ldg, static_cast<uint64_t>(N) * ldg, 32, 32, 64);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16_sw(&tdk, dk, D, N, B, ldg, static_cast<uint64_t>(N) * ldg, 32, 32, 64);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16_sw(&tdv, dv, D, N, B, ldg, static_cast<uint64_t>(N) * ldg, 32, 32, 64);
  if (rc) return rc;
  if (int rc2 = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_bwd_long_kernel), BL_SMEM, "attention_bwd_long")) return rc2;
  const int heads = B * H;
  const int num_sms = device_sm_count();
  if (bias_grad == nullptr) bias_mask = 0;
  dim3 grid(heads < num_sms ? heads : num_sms);  // persistent: a CTA owns (image, head) pairs and walks their key tiles
  long long* trace = nullptr;
#ifdef TIC_ATTN_TRACE
  cudaMallocManaged(&trace, 128 * sizeof(long long));
  for (int i = 0; i < 128; ++i) trace[i] = 0;
#endif
  launch_pdl(attn_bwd_long_kernel, grid, dim3(BL_THREADS), BL_SMEM, stream, tq, tk, tv, tdo, tdq, tdk, tdv, lse,
             reinterpret_cast<const __nv_bfloat16*>(o), ldo, reinterpret_cast<const __nv_bfloat16*>(dout), lddo, dq_scratch,
             bias_grad, bias_mask, N, Nq, H, heads, scale, trace);
#ifdef TIC_ATTN_TRACE
  cudaDeviceSynchronize();
  {
    const long long t0 = trace[0];
    fprintf(stderr, "[bl trace] compute warp 0 (3j: before wait S, +1: S ready, +2: P arrived):");
    for (int i = 0; i < 64; ++i) if (trace[i]) fprintf(stderr, " c%d=%lld", i, trace[i] - t0);
    fprintf(stderr, "\n[bl trace] mma thread (4j: before wait P, +1: P ready, +2: dV/dK issued, +3: scores(j+2) issued):");
    for (int i = 64; i < 112; ++i) if (trace[i]) fprintf(stderr, " m%d=%lld", i - 64, trace[i] - t0);
    fprintf(stderr, "\n[bl trace] drain warp 12 (2qt: before wait dQ, +1: dQ ready):");
    for (int i = 112; i < 128; ++i) if (trace[i]) fprintf(stderr, " d%d=%lld", i - 112, trace[i] - t0);
    fprintf(stderr, "\n");
    cudaFree(trace);
  }
#endif
  return check_launch("attention_bwd_long");
}

int attention_bwd_long_max_queries() { return BL_NQ_MAX; }

}  // namespace tic
