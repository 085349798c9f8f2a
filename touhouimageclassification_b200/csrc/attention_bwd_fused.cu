// Fused multi-head attention BACKWARD for sequences of up to 256 tokens (the 197-token ViT-*/16 224x224 case):
// ONE kernel per layer computes delta, dQ, dK and dV of F.scaled_dot_product_attention (modeling_vit.py:232-246 [a6]),
// with every score tile computed exactly once (5 matmuls per tile instead of the 7 of the split dQ / dKdV kernels in
// attention_tc.cu, and one exp2 / elementwise pass instead of two).
//
// One CTA per (image, head); all of Q, K, V, dO of that head (<= 256 rows x 64) sit in shared memory (TMA, 128-byte
// swizzle, rows past N zero-filled). Work is tiled as key tiles of 128 rows (UMMA M = TMEM lanes) x query blocks of 64:
//   warp 8 (one lane)  TMA loads, then the MMA issue loop
//        S^T  = K  Q^T      (SS, M=128 keys, N=64 queries, K=64)      -> TMEM S[buf]
//        dP^T = V  dO^T     (SS)                                      -> TMEM dP[buf]
//        dV  += P^T  dO     (TS: A = P^T  bf16 in TMEM, B = dO MN-major smem, K = 64 queries)
//        dK  += dS^T Q      (TS: A = dS^T bf16 in TMEM)
//        dQ  += dS   K      (SS: A = dS^T tile staged in smem by the compute warps, read as an MN-major operand
//                            with M = 128 queries, K = 128 keys; issued once per pair of query blocks)
//   warps 0-7          one key row per thread, two warps per TMEM lane quadrant (32 query columns each):
//        P^T = exp2(S^T c - L[q]),  dS^T = P^T o (dP^T - delta[q]);  packed bf16 back into TMEM (in place) and into
//        the smem staging tile.
//   warps 9-12         drain warps, one per lane quadrant: dV / dK after each key tile and dQ at the end leave TMEM
//        through registers -> per-warp smem tiles -> TMA stores, off the compute warps' critical path; the MMA thread
//        only waits for the (fast) TMEM reads before it overwrites an accumulator.
// S / dP are double buffered in TMEM so the score MMAs of block j+1 run under the elementwise pass of block j.
// delta[q] = rowsum(dO o O) comes from shared memory: O is TMA-loaded into the (still unused) second staging tile and
// each compute thread dots its own row of O and dO; no separate delta kernel, no global loads on the critical path.
// TMEM (512 columns): S[2] 0-127 | dP[2] 128-255 | dV 256-319 | dK 320-383 | dQ[2 query tiles] 384-511.
#include "tic_internal.cuh"


namespace tic {
namespace {

constexpr int FB_THREADS = 512;  // 8 compute warps + 1 TMA / MMA warp + 4 drain warps (one per TMEM lane quadrant) + 3 idle
// Register budget: warps are allocated in groups of four, so 13 warps cost 16 x 32 x regs. The kernel is compiled for
// 128 registers per thread and re-balances at run time (setmaxnreg): the two compute warpgroups grow to 152, the MMA /
// drain warpgroups shrink to 104 -- 8 x 32 x (152 + 104) = the whole register file (the two sides must balance: asking for
// more than the shrinking warps release blocks forever). The split is chosen by what ptxas then spills: 168 / 88 left 36
// bytes of spills in the MMA thread's per-block issue loop (on the critical P -> MMA path), 152 / 104 leaves 8; same-box
// A/B 0.314 -> 0.310 ms. Sixteen compute warps (16 query columns each, 88 registers)
// measured slower, 0.312 vs 0.297 ms: the block period is set by the score -> P -> dV/dK -> next-score dependency
// chain through the tensor pipe and the mbarrier hops, not by the elementwise pass. Signalling the S^T and dP^T
// halves of a block separately (four barriers per block instead of two) also measured slower (0.319 ms).
constexpr int FB_REGS_COMPUTE = 152, FB_REGS_OTHER = 104;
constexpr int FB_ROWS = 256;     // rows staged per operand
constexpr int FB_HD = 64;
constexpr float FB_LOG2E = 1.4426950408889634f;
constexpr int FB_OPER_BYTES = FB_ROWS * 128;   // 32 KB per operand
constexpr int FB_STAGE_BYTES = 2 * 128 * 128;  // one dS^T tile: 2 query chunks x 128 key rows x 128 B
constexpr int FB_OUT_BYTES = 4 * 4 * 2048;       // per drain warp: four 32-row x 64-byte output tiles for the TMA stores
constexpr int FB_SMEM_USED = 4 * FB_OPER_BYTES + 2 * FB_STAGE_BYTES + FB_OUT_BYTES + 2 * FB_ROWS * 4 + 128;
constexpr int FB_SMEM = 232448;                  // everything the SM has; the slack (960 B) absorbs the 1024-byte alignment
static_assert(FB_SMEM_USED <= FB_SMEM, "attention_bwd_fused: shared memory budget");
constexpr uint32_t FB_COL_DP = 128, FB_COL_DV = 256, FB_COL_DK = 320, FB_COL_DQ = 384;

TIC_DEVINL void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

TIC_DEVINL uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
TIC_DEVINL float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// 32 lanes each hold v[0..31] (one row, 32 columns): returns, in lane i, the sum over the 32 rows of column i.
// Recursive halving: 16 + 8 + 4 + 2 + 1 shuffles.
TIC_DEVINL float warp_colsum32(float (&v)[32], int lane) {
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

__global__ void __launch_bounds__(FB_THREADS, 1)
attn_bwd_fused_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                      const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_dq,
                      const __grid_constant__ CUtensorMap tm_dk, const __grid_constant__ CUtensorMap tm_dv,
                      const float* __restrict__ lse, float* __restrict__ bias_grad, int bias_mask, int N, int Nq, int H,
                      int num_items, float scale, long long* __restrict__ trace) {
  // N = keys per item; Nq = queries per item (the first Nq tokens of each image; Nq = 1 for the CLS-only last layer)
  // trace (dev tool, normally NULL): clock64 stamps of CTA 0 -- [0..63] compute warp 0, [64..127] the MMA thread
#ifdef TIC_ATTN_TRACE  // development build only (-DTIC_ATTN_TRACE): the shipped library has no tracing code
#define FB_STAMP(slot) do { if (trace != nullptr && blockIdx.x == 0 && it == 3) trace[slot] = clock64(); } while (0)
#else
#define FB_STAMP(slot) do { } while (0)
#endif
  pdl_launch_dependents();
  extern __shared__ uint8_t fb_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fb_smem_raw) + 1023) & ~uintptr_t(1023));
  if (smem + FB_SMEM_USED > fb_smem_raw + FB_SMEM) {  // never observed: the dynamic window starts 1024-byte aligned
    if (threadIdx.x == 0) printf("tic: attention_bwd_fused: shared memory window is misaligned\n");
    __trap();
  }
  uint8_t* sQ = smem;
  uint8_t* sDO = sQ + FB_OPER_BYTES;
  uint8_t* sK = sDO + FB_OPER_BYTES;
  uint8_t* sV = sK + FB_OPER_BYTES;
  uint8_t* sStage = sV + FB_OPER_BYTES;  // [2 query tiles][2 chunks of 64 queries][128 key rows][128 B]
  uint8_t* sOut = sStage + 2 * FB_STAGE_BYTES;  // [8 warps][2][32 rows][64 B], 64-byte swizzle (TMA store sources)
  float* sL = reinterpret_cast<float*>(sOut + FB_OUT_BYTES);  // [256] logsumexp * log2(e), +inf past N
  float* sD = sL + FB_ROWS;                                           // [256] delta = rowsum(dO o O)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + FB_ROWS);
  uint64_t* bar_load = bars + 0;  // [2] K + Q  |  dO + O + V
  uint64_t* bar_s = bars + 2;    // [2] score tiles of a block are in TMEM
  uint64_t* bar_p = bars + 4;    // [2] P^T / dS^T of a block written (8 warp arrivals)
  uint64_t* bar_acc = bars + 6;      // every MMA of a key tile has completed
  uint64_t* bar_free_vk = bars + 7;  // the drain warps have read dV / dK of a key tile out of TMEM (4 warp arrivals)
  uint64_t* bar_free_q = bars + 8;   // ... and dQ of an item
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  uint8_t* sO = sStage + FB_STAGE_BYTES;  // O rows live in the second staging tile until delta has been computed

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkt = (N + 127) >> 7;             // key tiles of 128
  const int nqb = (Nq + 63) >> 6;             // query blocks of 64
  const int w_last = ((Nq - (nqb - 1) * 64) + 15) & ~15;  // width of the last query block (multiple of 16)
  const int J = nkt * nqb;

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_do);
      tma_prefetch_desc(&tm_o); tma_prefetch_desc(&tm_dq); tma_prefetch_desc(&tm_dk); tma_prefetch_desc(&tm_dv);
      mbar_init(&bar_load[0], 1);
      mbar_init(&bar_load[1], 1);
      mbar_init(&bar_s[0], 1);
      mbar_init(&bar_s[1], 1);
      mbar_init(&bar_p[0], 8);
      mbar_init(&bar_p[1], 8);
      mbar_init(bar_acc, 1);
      mbar_init(bar_free_vk, 4);
      mbar_init(bar_free_q, 4);
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

  // Persistent: this CTA walks the (image, head) items blockIdx.x, blockIdx.x + gridDim.x, ... The loads of item i+1
  // are issued as soon as the last MMA of item i has completed, i.e. under the final accumulator drain of item i.
  // setmaxnreg sits INSIDE each side of the role dispatch: ptxas budgets registers for the code a setmaxnreg dominates
  // (placed before the dispatch, every role was compiled under the kernel-wide figure and the compute loop spilled)
  if (warp < 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FB_REGS_COMPUTE));
    // ------------------------------------------------------------------------------------ compute warps
    const int quad = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float c2 = scale * FB_LOG2E;
    const int r = quad * 32 + lane;  // row within the 128-row tile
    const uint32_t stage_row = smem_u32(sStage) + r * 128;
    const int sw = r & 7;
    const int t = threadIdx.x;       // 0..255: the query whose delta / logsumexp this thread prepares
    const uint32_t aL = smem_u32(sL), aD = smem_u32(sD);
    const uint32_t ro = smem_u32(sO) + t * 128, rd = smem_u32(sDO) + t * 128;
    uint32_t ph_s = 0;  // bit b = parity of the next completion of bar_s[b]
    auto load_lse = [&](int item) -> float {  // per-query logsumexp, +inf past N: exp2(-inf) = 0
      return (item < num_items && t < Nq) ? __ldg(lse + static_cast<long long>(item) * Nq + t) : INFINITY;
    };
    float L_next = load_lse(blockIdx.x);
    for (int it = 0, item = blockIdx.x; item < num_items; ++it, item += gridDim.x) {
      const int h = item % H, b = item / H;
#define FB_CSTAMP(slot) do { if (threadIdx.x == 0) FB_STAMP(slot); } while (0)
      FB_CSTAMP(0);
      {
        // delta[q] = sum_d dO[q, d] * O[q, d], one query row per thread, both rows read from swizzled shared memory
        mbar_wait(&bar_load[1], it & 1);
        FB_CSTAMP(1);
        // four independent partial sums: one 64-long FMA chain kept this pass (which sits on the path between two heads)
        // waiting on its own latency
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t off = static_cast<uint32_t>((i ^ (t & 7)) << 4);
          const uint4 a = ld_shared_v4(ro + off), g = ld_shared_v4(rd + off);
          d0 = fmaf(bf16_lo(a.x), bf16_lo(g.x), d0); d1 = fmaf(bf16_hi(a.x), bf16_hi(g.x), d1);
          d2 = fmaf(bf16_lo(a.y), bf16_lo(g.y), d2); d3 = fmaf(bf16_hi(a.y), bf16_hi(g.y), d3);
          d0 = fmaf(bf16_lo(a.z), bf16_lo(g.z), d0); d1 = fmaf(bf16_hi(a.z), bf16_hi(g.z), d1);
          d2 = fmaf(bf16_lo(a.w), bf16_lo(g.w), d2); d3 = fmaf(bf16_hi(a.w), bf16_hi(g.w), d3);
        }
        const float dsum = (d0 + d1) + (d2 + d3);
        sL[t] = L_next * FB_LOG2E;  // log2 domain (L_next was loaded under the previous item's final drain)
        sD[t] = dsum;    // rows past N are zero-filled by TMA: delta = 0
        asm volatile("bar.sync 1, 256;" ::: "memory");
        FB_CSTAMP(2);
      }
      for (int j = 0; j < J; ++j) {
        const int kt = j >= nqb ? 1 : 0, qb = j - kt * nqb, buf = j & 1;
        const int w = qb == nqb - 1 ? w_last : 64;
        const bool quad_active = kt * 128 + quad * 32 < N;
        const bool row_valid = kt * 128 + r < N;
        mbar_wait(&bar_s[buf], (ph_s >> buf) & 1);
        FB_CSTAMP(4 + 3 * j);
        ph_s ^= 1u << buf;
        tc_fence_after();
        if (quad_active && half * 32 < w) {
          uint32_t s[32], dp[32];
          tmem_ld_32x32b_x32(lane_addr + buf * 64 + half * 32, s);
          tmem_ld_32x32b_x32(lane_addr + FB_COL_DP + buf * 64 + half * 32, dp);
          tmem_ld_wait();
          const uint32_t L4 = aL + (qb * 64 + half * 32) * 4, D4 = aD + (qb * 64 + half * 32) * 4;
          uint32_t pw[16], dw[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 Lq = ld_shared_f4(L4 + 16 * i), Dq = ld_shared_f4(D4 + 16 * i);
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
          tmem_st_32x32b_x16(lane_addr + FB_COL_DP + buf * 64 + half * 32, dw);
          // dS^T row -> staging tile of this query tile (chunk = 64-query block), zero for key rows past N so that the
          // dQ product never sees a non-finite value against the zero-filled K rows
          const uint32_t dst = stage_row + (nqb <= 2 ? (kt & 1) : (qb >> 1)) * FB_STAGE_BYTES + (qb & 1) * 16384;
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            const uint32_t a = dst + (((half * 4 + pc) ^ sw) << 4);
            if (row_valid) st_shared_v4(a, dw[4 * pc], dw[4 * pc + 1], dw[4 * pc + 2], dw[4 * pc + 3]);
            else st_shared_v4(a, 0u, 0u, 0u, 0u);
          }
          tmem_st_wait();
          fence_proxy_async();
        }
        FB_CSTAMP(5 + 3 * j);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_p[buf]);
        FB_CSTAMP(6 + 3 * j);

        if (j == J - 1) L_next = load_lse(item + gridDim.x);  // in flight across the item boundary
      }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FB_REGS_OTHER));
    if (warp == 8) {
      if (elect_one()) {  // one issuing thread on the uniform datapath (a lane == 0 test costs ~45 clk per MMA)
        // ---------------------------------------------------------------------------------- TMA + MMA issue loop
        constexpr uint32_t idesc_ts = make_idesc_bf16(128, FB_HD, false, true);
        constexpr uint32_t idesc_dq = make_idesc_bf16(128, FB_HD, true, true);
        const uint32_t aQ = smem_u32(sQ), aDO = smem_u32(sDO), aK = smem_u32(sK), aV = smem_u32(sV), aS = smem_u32(sStage);
        auto issue_loads = [&](int item) {
          const int h = item % H, b = item / H;
          mbar_arrive_expect_tx(&bar_load[0], 2 * FB_OPER_BYTES);
          tma_load_3d(sK, &tm_k, &bar_load[0], h * FB_HD, 0, b);
          tma_load_3d(sQ, &tm_q, &bar_load[0], h * FB_HD, 0, b);
          mbar_arrive_expect_tx(&bar_load[1], 3 * FB_OPER_BYTES);
          tma_load_3d(sDO, &tm_do, &bar_load[1], h * FB_HD, 0, b);
          tma_load_3d(sO, &tm_o, &bar_load[1], h * FB_HD, 0, b);
          tma_load_3d(sV, &tm_v, &bar_load[1], h * FB_HD, 0, b);
        };
        uint32_t ph_p = 0, use_acc = 0, use_vk = 0;  // ph_p: bit b = parity of the next completion of bar_p[b]
        issue_loads(blockIdx.x);
        for (int it = 0, item = blockIdx.x; item < num_items; ++it, item += gridDim.x) {
          // The operand buffers are single (the SM's shared memory holds one head), so the next item's loads can only be
          // issued when this item's last MMA has completed -- but its addresses are known now: pull the five boxes into L2
          // so that those loads pay L2 latency, not DRAM latency, on the critical path between two items.
          if (item + static_cast<int>(gridDim.x) < num_items) {
            const int nx = item + gridDim.x, nh = nx % H, nb = nx / H;
            tma_prefetch_l2_3d(&tm_k, nh * FB_HD, 0, nb);
            tma_prefetch_l2_3d(&tm_q, nh * FB_HD, 0, nb);
            tma_prefetch_l2_3d(&tm_do, nh * FB_HD, 0, nb);
            tma_prefetch_l2_3d(&tm_o, nh * FB_HD, 0, nb);
            tma_prefetch_l2_3d(&tm_v, nh * FB_HD, 0, nb);
          }
          auto issue_scores = [&](int j) {
            const int kt = j >= nqb ? 1 : 0, qb = j - kt * nqb, buf = j & 1;
            const int w = qb == nqb - 1 ? w_last : 64;
            const uint32_t idesc = make_idesc_bf16(128, w, false, false);
            const uint64_t dK_ = make_smem_desc_sw128(aK + kt * 16384, 0, 1024), dQ_ = make_smem_desc_sw128(aQ + qb * 8192, 0, 1024);
            const uint64_t dV_ = make_smem_desc_sw128(aV + kt * 16384, 0, 1024), dO_ = make_smem_desc_sw128(aDO + qb * 8192, 0, 1024);
            if (j == 0) { mbar_wait(&bar_load[0], it & 1); tc_fence_after(); }
#pragma unroll
            for (int k = 0; k < FB_HD / 16; ++k) umma_bf16_ss(tmem_base + buf * 64, dK_ + 2 * k, dQ_ + 2 * k, idesc, k > 0 ? 1u : 0u);
            if (j == 0) { mbar_wait(&bar_load[1], it & 1); tc_fence_after(); }
#pragma unroll
            for (int k = 0; k < FB_HD / 16; ++k)
              umma_bf16_ss(tmem_base + FB_COL_DP + buf * 64, dV_ + 2 * k, dO_ + 2 * k, idesc, k > 0 ? 1u : 0u);
            umma_commit(&bar_s[buf]);
          };
          FB_STAMP(64);
          issue_scores(0);
          FB_STAMP(65);
          if (J > 1) issue_scores(1);
          FB_STAMP(66);
          for (int j = 0; j < J; ++j) {
            const int kt = j >= nqb ? 1 : 0, qb = j - kt * nqb, buf = j & 1;
            mbar_wait(&bar_p[buf], (ph_p >> buf) & 1);
            FB_STAMP(70 + 2 * j);
            ph_p ^= 1u << buf;
            tc_fence_after();
            if (qb == 0 && (it > 0 || kt > 0)) {  // the previous key tile's dV / dK have left TMEM
              mbar_wait(bar_free_vk, use_vk & 1);
              ++use_vk;
              tc_fence_after();
            }
            const int ksteps = (qb == nqb - 1 ? w_last : 64) >> 4;
            const uint64_t dO_mn = make_smem_desc_sw128(aDO + qb * 8192, 8192, 1024);
            const uint64_t dQ_mn = make_smem_desc_sw128(aQ + qb * 8192, 8192, 1024);
            for (int k = 0; k < ksteps; ++k) {  // dV += P^T dO
              const uint32_t a = tmem_base + buf * 64 + (k >> 1) * 32 + (k & 1) * 8;
              umma_bf16_ts(tmem_base + FB_COL_DV, a, dO_mn + 128 * k, idesc_ts, (qb > 0 || k > 0) ? 1u : 0u);
            }
            for (int k = 0; k < ksteps; ++k) {  // dK += dS^T Q
              const uint32_t a = tmem_base + FB_COL_DP + buf * 64 + (k >> 1) * 32 + (k & 1) * 8;
              umma_bf16_ts(tmem_base + FB_COL_DK, a, dQ_mn + 128 * k, idesc_ts, (qb > 0 || k > 0) ? 1u : 0u);
            }
            // dV / dK of a key tile that is not the last one are complete here: let the drain warps start before the score
            // and dQ products queued behind them
            const bool tile_end = qb == nqb - 1, early = tile_end && kt < nkt - 1;
            if (early) umma_commit(bar_acc);
            // the P^T / dS^T columns of this buffer have been consumed (in issue order): refill it with the scores of
            // block j+2 before the dQ product, which the compute warps do not wait for
            if (j + 2 < J) issue_scores(j + 2);
            if ((qb & 1) || qb == nqb - 1) {    // dQ[query tile] += dS K over this key tile
              const int qt = qb >> 1;
              if (kt == 0 && qt == 0 && it > 0) {  // the previous item's dQ has left TMEM
                mbar_wait(bar_free_q, (it - 1) & 1);
                tc_fence_after();
              }
              const int kvalid = min(128, N - kt * 128);
              const int ks = (kvalid + 15) >> 4;
              // staging tile of (key tile, query tile): with at most two query blocks there is one query tile and the two key
              // tiles alternate between the two staging tiles (a tile is never rewritten while its dQ product may be reading it)
              const int st = nqb <= 2 ? (kt & 1) : qt;
              const uint64_t dS_mn = make_smem_desc_sw128(aS + st * FB_STAGE_BYTES, 16384, 1024);
              const uint64_t dK_mn = make_smem_desc_sw128(aK + kt * 16384, 8192, 1024);
              for (int k = 0; k < ks; ++k)
                umma_bf16_ss(tmem_base + FB_COL_DQ + qt * 64, dS_mn + 128 * k, dK_mn + 128 * k, idesc_dq, (kt > 0 || k > 0) ? 1u : 0u);
            }
            FB_STAMP(71 + 2 * j);
            if (tile_end) {
              if (!early) umma_commit(bar_acc);
              if (kt == nkt - 1 && item + static_cast<int>(gridDim.x) < num_items) {
                // every MMA of this item has read its operands: refill the operand buffers for the next item
                mbar_wait(bar_acc, use_acc & 1);
                FB_STAMP(100);
                issue_loads(item + gridDim.x);
                FB_STAMP(101);
              }
              ++use_acc;
            }
          }
        }
      }
      __syncwarp();
    }
    if (warp > 8 && warp < 13) {
      // ------------------------------------------------------------------------------------ drain warps
      const int quad = warp & 3;  // warps 9, 10, 11, 12 -> TMEM lane quadrants 1, 2, 3, 0
      const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
      const uint32_t stage = smem_u32(sOut) + (warp - 9) * 8192;  // four 2 KB tiles: [32 rows][64 B], 64-byte swizzle
      const int r = quad * 32 + lane;
      uint32_t use_acc = 0;
      for (int it = 0, item = blockIdx.x; item < num_items; ++it, item += gridDim.x) {
        const int h = item % H, b = item / H;
        // One half tile (32 rows x 32 fp32 accumulator columns of this warp's lane quadrant): TMEM -> scaled, packed bf16
        // -> staging slot (64-byte swizzle); column sums of the bf16 rows accumulate into the QKV bias gradient when asked
        // for. Done one half tile at a time so that the drain warps live within their registers.
        auto stage_half = [&](int slot, uint32_t col, float f, bool valid, float* bias_dst, int half) {
          uint32_t rr[32];
          tmem_ld_32x32b_x32(lane_addr + col, rr);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(__uint_as_float(rr[2 * i]) * f, __uint_as_float(rr[2 * i + 1]) * f);
          const uint32_t base = stage + slot * 2048 + lane * 64;
          const int x = (lane >> 1) & 3;
#pragma unroll
          for (int pc = 0; pc < 4; ++pc)
            st_shared_v4(base + ((pc ^ x) << 4), pk[4 * pc], pk[4 * pc + 1], pk[4 * pc + 2], pk[4 * pc + 3]);
          if (bias_dst != nullptr) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              v[2 * i] = valid ? bf16_lo(pk[i]) : 0.f;
              v[2 * i + 1] = valid ? bf16_hi(pk[i]) : 0.f;
            }
            const float cs = warp_colsum32(v, lane);
            atomicAdd(bias_dst + h * FB_HD + half * 32 + lane, cs);
          }
        };
        for (int kt = 0; kt < nkt; ++kt) {
          const bool quad_active = kt * 128 + quad * 32 < N;
          const int row0 = kt * 128 + quad * 32;
          const bool valid = kt * 128 + r < N;
          mbar_wait(bar_acc, use_acc & 1);
          ++use_acc;
          tc_fence_after();
          if (lane == 0) tma_store_wait_read<0>();  // this warp's staging slots are free again (stores issued long ago)
          __syncwarp();
          if (quad_active) {
            float* vdst = (bias_mask & 4) ? bias_grad + 2 * H * FB_HD : nullptr;
            float* kdst = (bias_mask & 2) ? bias_grad + H * FB_HD : nullptr;
            stage_half(0, FB_COL_DV, 1.0f, valid, vdst, 0);
            stage_half(1, FB_COL_DV + 32, 1.0f, valid, vdst, 1);
            stage_half(2, FB_COL_DK, scale, valid, kdst, 0);
            stage_half(3, FB_COL_DK + 32, scale, valid, kdst, 1);
            fence_proxy_async();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bar_free_vk);  // the MMA thread may overwrite dV / dK
            if (quad_active) {
              tma_store_3d_addr(&tm_dv, stage, h * FB_HD, row0, b);
              tma_store_3d_addr(&tm_dv, stage + 2048, h * FB_HD + 32, row0, b);
              tma_store_3d_addr(&tm_dk, stage + 4096, h * FB_HD, row0, b);
              tma_store_3d_addr(&tm_dk, stage + 6144, h * FB_HD + 32, row0, b);
              tma_store_commit();
            }
          }
          if (kt == nkt - 1) {  // dQ of both query tiles (all of this item's MMAs have completed)
            const bool q0 = quad * 32 < Nq, q1 = 128 + quad * 32 < Nq;
            float* qdst = (bias_mask & 1) ? bias_grad : nullptr;
            if (lane == 0) tma_store_wait_read<0>();  // the dV / dK stores above have finished reading the slots
            __syncwarp();
            if (q0) {
              stage_half(0, FB_COL_DQ, scale, r < Nq, qdst, 0);
              stage_half(1, FB_COL_DQ + 32, scale, r < Nq, qdst, 1);
            }
            if (q1) {
              stage_half(2, FB_COL_DQ + 64, scale, 128 + r < Nq, qdst, 0);
              stage_half(3, FB_COL_DQ + 96, scale, 128 + r < Nq, qdst, 1);
            }
            if (q0) fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(bar_free_q);
              if (q0) {
                tma_store_3d_addr(&tm_dq, stage, h * FB_HD, quad * 32, b);
                tma_store_3d_addr(&tm_dq, stage + 2048, h * FB_HD + 32, quad * 32, b);
              }
              if (q1) {
                tma_store_3d_addr(&tm_dq, stage + 4096, h * FB_HD, 128 + quad * 32, b);
                tma_store_3d_addr(&tm_dq, stage + 6144, h * FB_HD + 32, 128 + quad * 32, b);
              }
              if (q0) tma_store_commit();
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

// q/k/v: [B*N, ...] pitch ld, head h at column h*64; o / dout: [B*N, H*64]; dq/dk/dv pitch ldg. Only the first Nq tokens
// of every image are queries (their rows of o / dout are read, their rows of dq written; lse is [B, H, Nq]).
// bias_grad (optional):
// fp32 [3*H*64] laid out q | k | v, ACCUMULATES the column sums of dq (bias_mask bit 0) / dk (bit 1) / dv (bit 2).
int attention_bwd_fused(const void* q, const void* k, const void* v, long long ld, const void* o, long long ldo,
                        const void* dout, long long lddo, const float* lse, void* dq, void* dk, void* dv, long long ldg,
                        float* bias_grad, int bias_mask, int B, int N, int Nq, int H, float scale, cudaStream_t stream) {
  if (Nq <= 0 || Nq > N) return set_error(kErrInvalidArg, "attention_bwd_fused: Nq=%d must be in [1, N=%d]", Nq, N);
  if (N > FB_ROWS) return set_error(kErrUnsupported, "attention_bwd_fused: N=%d > %d", N, FB_ROWS);
  CUtensorMap tq, tk, tv, tdo, to, tdq, tdk, tdv;
  const uint64_t D = static_cast<uint64_t>(H) * FB_HD;
  int rc = encode_tmap_3d_bf16(&tq, q, D, Nq, B, ld, static_cast<uint64_t>(N) * ld, 64, FB_ROWS);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&tk, k, D, N, B, ld, static_cast<uint64_t>(N) * ld, 64, FB_ROWS);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&tv, v, D, N, B, ld, static_cast<uint64_t>(N) * ld, 64, FB_ROWS);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&tdo, dout, D, Nq, B, lddo, static_cast<uint64_t>(N) * lddo, 64, FB_ROWS);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&to, o, D, Nq, B, ldo, static_cast<uint64_t>(N) * ldo, 64, FB_ROWS);
  if (rc) return rc;
  // outputs: 32-column x 32-row boxes (one compute warp's tile), 64-byte swizzle
  rc = encode_tmap_3d_bf16_sw(&tdq, dq, D, Nq, B, ldg, static_cast<uint64_t>(N) * ldg, 32, 32, 64);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16_sw(&tdk, dk, D, N, B, ldg, static_cast<uint64_t>(N) * ldg, 32, 32, 64);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16_sw(&tdv, dv, D, N, B, ldg, static_cast<uint64_t>(N) * ldg, 32, 32, 64);
  if (rc) return rc;
  if (int rc2 = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_bwd_fused_kernel), FB_SMEM, "attention_bwd_fused")) return rc2;
  const int num_sms = device_sm_count();
  const int items = B * H;
  if (bias_grad == nullptr) bias_mask = 0;
  dim3 grid(items < num_sms ? items : num_sms);  // persistent: one CTA per SM walks the (image, head) items
  long long* trace = nullptr;
#ifdef TIC_ATTN_TRACE
  cudaMallocManaged(&trace, 128 * sizeof(long long));
  for (int i = 0; i < 128; ++i) trace[i] = 0;
#endif
  launch_pdl(attn_bwd_fused_kernel, grid, dim3(FB_THREADS), FB_SMEM, stream, tq, tk, tv, tdo, to, tdq, tdk, tdv, lse, bias_grad,
             bias_mask, N, Nq, H, items, scale, trace);
#ifdef TIC_ATTN_TRACE
  cudaDeviceSynchronize();
  {
    const long long t0 = trace[0];
    fprintf(stderr, "[fb trace] compute warp 0 (clk since item start):");
    for (int i = 0; i < 64; ++i) if (trace[i]) fprintf(stderr, " c%d=%lld", i, trace[i] - t0);
    fprintf(stderr, "\n[fb trace] mma thread:");
    for (int i = 64; i < 128; ++i) if (trace[i]) fprintf(stderr, " m%d=%lld", i - 64, trace[i] - t0);
    fprintf(stderr, "\n");
    cudaFree(trace);
  }
#endif
  return check_launch("attention_bwd_fused");
}

}  // namespace tic
