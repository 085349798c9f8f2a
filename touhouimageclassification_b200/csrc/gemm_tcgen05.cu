// bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged in shared memory by TMA with 128-byte swizzle, fused epilogues.
//
// Replaces, on the hot path, every nn.Linear the reference reaches through
// transformers/models/vit/modeling_vit.py (q/k/v :216-230, attention out :262-268,
// intermediate :290-299, output :305-312, patch projection :151-167) and their autograd
// backward GEMMs (cuBLASLt in the reference's stack).
//
// One CTA pair (cluster of 2, tcgen05 cta_group::2) per two SMs computes 256 x 256 tiles; 10 warps per CTA (18 for the
// two GELU epilogues). Few-tile problems run single CTAs with 128 x 128 tiles instead (prefer_small_tiles below):
//   warps 0-7  epilogue   (TMEM -> registers -> fused math -> global), 2 warps per TMEM lane quadrant
//   warp  8    TMA producer (one elected lane) + tile scheduler of the pair (leader CTA)
//   warp  9    MMA issuer  (one elected lane, leader CTA) + TMEM allocator
// mbarrier pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), tile hand-out full/empty;
// two 256-column accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
// Tiles are handed out by the hardware (cluster launch control): the grid holds one pair per tile, the pairs that run
// cancel not-yet-started ones and take their tiles, so the kernel is persistent without a fixed tile -> SM mapping.
//
// Operand majors (template): K-major = reduction dimension contiguous in memory,
// MN-major = the M (or N) dimension contiguous. The three GEMMs of a Linear layer map to
//   forward  Y = X W^T   : A = X  (K-major)   B = W  (K-major)
//   dgrad    dX = dY W   : A = dY (K-major)   B = W  (MN-major; no transposed weight copy needed)
//   wgrad    dW = dY^T X : A = dY (MN-major)  B = X  (MN-major), split over the token dimension
#include "tic_internal.cuh"

#include <cstdlib>

#include <type_traits>

namespace tic {

namespace {

constexpr int BM = 128;   // accumulator rows per CTA (one TMEM lane per row)
constexpr int BN_WIDE = 256;  // N tile of the CTA-pair configuration (256 x 256 per pair)
constexpr int BN_SMALL = 128; // N tile of the single-CTA configuration for few-tile problems (128 x 128 per CTA)
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int ACC_STAGES = 2;
constexpr int CLC_STAGES = 4;  // ring of cluster-launch-control responses (tile hand-outs in flight)
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
// Epilogue warps per CTA: 8 (two per TMEM lane quadrant) everywhere except the two GELU epilogues (forward: value +
// derivative, ~24 instructions per element; fc2 dgrad x GELU' + column sums), whose per-chunk dependency chains need
// more warps per SM sub-partition to fill the issue slots inside one tile's MMA time: they run 16 (four per quadrant, two
// 32-column chunks each) and give up one operand stage. Measured at M = 50432 (B200, same box): forward 12 -> 16 warps
// 1176 -> 1191 TFLOP/s, dgrad 8 -> 16 warps 1180 -> 1262; the residual epilogue gets SLOWER with more warps (8: 1360,
// 12: 1334, 16: 1256 TFLOP/s at K = 4096 -- its auxiliary-operand prefetch registers and the sixth stage matter more).
template <int EPI, int BN = BN_WIDE>
constexpr int epi_warps() { return ((EPI == kEpiBf16Gelu || EPI == kEpiBf16DGelu) && BN == BN_WIDE) ? 16 : 8; }
constexpr int EPI_STAGE_BYTES = 32 * 32 * 4;  // per epilogue warp: 32 rows x 32 fp32 columns

// NCTA = 1: one CTA computes a 128 x 256 tile.  NCTA = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) computes a
// 256 x 256 tile with one UMMA M=256: each CTA stages its own 128 A rows and HALF of the B tile (128 of the 256
// N rows), which halves the L2 -> smem operand traffic per FLOP and frees smem for a deeper ring.
template <int NCTA, int NEPI = 8, int BN = BN_WIDE>
struct Cfg {
  static constexpr int B_ROWS = BN / NCTA;                 // B rows (N) staged per CTA
  static constexpr int B_STAGE_BYTES = B_ROWS * BK * 2;    // 32 KB / 16 KB
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = NCTA == 2 ? (NEPI > 8 ? 5 : 6) : (BN == BN_SMALL ? 6 : 4);
  static constexpr int TILE_M = BM * NCTA;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NEPI * EPI_STAGE_BYTES + 1024 /*align*/ + 512 /*barriers*/;
};


struct GemmParams {
  int M, N, K;
  int splits;
  void* out;
  long long ldo;
  void* out2;
  long long ldo2;
  const float* bias;
  const void* aux;
  long long ldaux;
  int aux_int;  // kEpiF32PosEmbed: patches per image (P)
  float* colsum;  // kEpiBf16 / kEpiBf16DGelu: optional, accumulates column sums of the bf16 output (a bias gradient)
  int exact;      // fp32 verification mode: kEpiF32Resid / kEpiF32PosEmbed add without the bf16 rounding of autocast
  int dynamic;    // 1: one cluster per tile in the grid, running clusters take over not-yet-started ones (cluster launch control)
};

template <bool A_MN, bool B_MN, int EPI, int NCTA, int BN>
__global__ void __launch_bounds__((epi_warps<EPI, BN>() + 2) * 32, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const GemmParams p) {
  pdl_launch_dependents();  // programmatic dependent launch: the next kernel's prologue may overlap this kernel
  constexpr int NUM_EPI_WARPS = epi_warps<EPI, BN>();
  constexpr int TMEM_COLS = ACC_STAGES * BN;
  constexpr int kTmaWarp = NUM_EPI_WARPS, kMmaWarp = NUM_EPI_WARPS + 1;
  constexpr int kParts = NUM_EPI_WARPS / 4;  // epilogue warps per TMEM lane quadrant
  using C = Cfg<NCTA, NUM_EPI_WARPS, BN>;
  constexpr int STAGES = C::STAGES;
  constexpr int B_STAGE_BYTES = C::B_STAGE_BYTES;
  constexpr int STAGE_BYTES = C::STAGE_BYTES;
  const uint32_t cta_rank = NCTA == 2 ? cluster_ctarank() : 0u;   // 0 = leader of the pair
  const int worker = NCTA == 2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int num_workers = NCTA == 2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms are 1024 B: align the operand ring.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint8_t* smem_stage = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + NUM_EPI_WARPS * EPI_STAGE_BYTES);
  uint64_t* full_bar = bars;                      // [STAGES]   TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;            // [STAGES]   MMA -> TMA
  uint64_t* tmem_full_bar = bars + 2 * STAGES;    // [ACC_STAGES] MMA -> epilogue
  uint64_t* tmem_empty_bar = tmem_full_bar + ACC_STAGES;  // [ACC_STAGES] epilogue -> MMA
  uint64_t* clc_full = tmem_empty_bar + ACC_STAGES;       // [CLC_STAGES] hardware -> every role, per CTA
  uint64_t* clc_empty = clc_full + CLC_STAGES;            // [CLC_STAGES] every role (both CTAs) -> the leader's scheduler
  uint4* clc_resp = reinterpret_cast<uint4*>(bars + 32);  // [CLC_STAGES] 16-byte responses (256 B into the region)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(clc_resp + CLC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_blocks = (p.M + C::TILE_M - 1) / C::TILE_M;
  const int n_blocks = (p.N + BN - 1) / BN;
  const int k_blocks_total = (p.K + BK - 1) / BK;
  const int k_blocks_per_split = (k_blocks_total + p.splits - 1) / p.splits;
  const int tiles_per_split = m_blocks * n_blocks;
  const int total_tiles = tiles_per_split * p.splits;

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if constexpr (NCTA == 2) cluster_sync_all();  // both CTAs are resident before the paired TMEM allocation
  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&full_bar[i], NCTA);   // the leader's barrier collects one producer arrival per CTA
        mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < ACC_STAGES; ++i) {
        mbar_init(&tmem_full_bar[i], 1);
        mbar_init(&tmem_empty_bar[i], NUM_EPI_WARPS * NCTA);
      }
      for (int i = 0; i < CLC_STAGES; ++i) {
        mbar_init(&clc_full[i], 1);
        // readers of a response: the leader's MMA thread, the peer's TMA thread, every epilogue warp of the cluster
        mbar_init(&clc_empty[i], NUM_EPI_WARPS * NCTA + 1 + (NCTA - 1));
      }
      fence_mbar_init();
    }
    __syncwarp();
    if constexpr (NCTA == 2) tmem_alloc_2cta(tmem_base_slot, TMEM_COLS);
    else tmem_alloc(tmem_base_slot, TMEM_COLS);
  }
  tc_fence_before();
  if constexpr (NCTA == 2) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  pdl_wait();  // everything above is independent of the preceding kernel; operands and outputs are touched below

  // ---- tile sequence of this worker (a CTA or a CTA pair).
  // Static: tiles worker, worker + num_workers, ... of a grid sized to the machine.
  // Dynamic (p.dynamic): the grid holds one cluster per tile; a running cluster starts with its own tile and then asks
  // the hardware (cluster launch control) to cancel a cluster that has not started and processes that one's tile, so
  // the tiles spread over whatever SMs are actually free (e.g. while a NCCL kernel holds some) instead of waiting for
  // a fixed owner. The leader's TMA thread issues request i before loading tile i; every role reads response i after
  // tile i and releases the slot.
  auto next_tile = [&](int tile, int& qi, bool release) -> int {
    if (!p.dynamic) {
      const int t = tile + num_workers;
      return t < total_tiles ? t : -1;
    }
    const int slot = qi % CLC_STAGES;
    const uint32_t ph = static_cast<uint32_t>(qi / CLC_STAGES) & 1u;
    ++qi;
    mbar_wait(&clc_full[slot], ph);
    const int x = clc_decode(&clc_resp[slot]);
    if (release) {
      fence_proxy_async();  // this read happens before the next asynchronous write of the slot
      if constexpr (NCTA == 2) mbar_arrive_cluster(&clc_empty[slot], 0);
      else mbar_arrive(&clc_empty[slot]);
    }
    return x < 0 ? -1 : x / NCTA;
  };
  auto request_tile = [&](int q) {  // leader's TMA thread only
    const int slot = q % CLC_STAGES;
    const uint32_t ph = static_cast<uint32_t>(q / CLC_STAGES) & 1u;
    mbar_wait(&clc_empty[slot], ph ^ 1);
    mbar_arrive_expect_tx(&clc_full[slot], 16);
    if constexpr (NCTA == 2) {
      mbar_arrive_expect_tx_cluster(&clc_full[slot], 16, 1);
      clc_try_cancel_multicast(&clc_resp[slot], &clc_full[slot]);
    } else {
      clc_try_cancel(&clc_resp[slot], &clc_full[slot]);
    }
  };

  auto next_tile_warp = [&](int tile, int& qi) -> int {  // whole-warp form: every lane reads, lane 0 releases
    const int t = next_tile(tile, qi, false);
    if (p.dynamic) {
      __syncwarp();
      if (lane == 0) {
        const int slot = (qi - 1) % CLC_STAGES;
        fence_proxy_async();
        if constexpr (NCTA == 2) mbar_arrive_cluster(&clc_empty[slot], 0);
        else mbar_arrive(&clc_empty[slot]);
      }
    }
    return t;
  };

  if (warp == kTmaWarp) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      auto load = [&](void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
        if constexpr (NCTA == 2) tma_load_2d_2cta(dst, map, bar, c0, c1);
        else tma_load_2d(dst, map, bar, c0, c1);
      };
      int qi = 0;
      for (int tile = worker; tile >= 0; tile = next_tile(tile, qi, cta_rank != 0)) {
        if (p.dynamic && cta_rank == 0) request_tile(qi);
        const int split = tile / tiles_per_split;
        const int rem = tile - split * tiles_per_split;
        const int m_idx = (rem / n_blocks) * C::TILE_M + static_cast<int>(cta_rank) * BM;        // this CTA's A rows
        const int n_idx = (rem % n_blocks) * BN + static_cast<int>(cta_rank) * C::B_ROWS;        // this CTA's B rows
        const int kb_begin = split * k_blocks_per_split;
        const int kb_end = min(k_blocks_total, kb_begin + k_blocks_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES * NCTA);
          else mbar_arrive_cluster(&full_bar[stage], 0);
          uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * B_STAGE_BYTES;
          const int k_idx = kb * BK;
          if constexpr (!A_MN) {
            load(sa, &tmap_a, &full_bar[stage], k_idx, m_idx);  // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)  // boxes {64 m, 64 k}
              load(sa + j * (64 * BK * 2), &tmap_a, &full_bar[stage], m_idx + 64 * j, k_idx);
          }
          if constexpr (!B_MN) {
            load(sb, &tmap_b, &full_bar[stage], k_idx, n_idx);  // box {64 k, B_ROWS rows}
          } else {
#pragma unroll
            for (int j = 0; j < C::B_ROWS / 64; ++j)  // boxes {64 n, 64 k}
              load(sb + j * (64 * BK * 2), &tmap_b, &full_bar[stage], n_idx + 64 * j, k_idx);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    if (cta_rank == 0 && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(C::TILE_M, BN, A_MN, B_MN);
      // K-major SW128: 8-row atoms 1024 B apart (SBO); LBO unused.
      // MN-major SW128: atoms of 64 (MN) x 8 (K); next 8 k-rows at SBO = 1024 B, next 64-wide MN chunk at
      // LBO = 64 * BK * 2 = 8192 B (one TMA box).
      constexpr uint32_t a_lbo = A_MN ? 64 * BK * 2 : 0, b_lbo = B_MN ? 64 * BK * 2 : 0;
      constexpr uint32_t a_kstep = A_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;  // desc.lo units of 16 B
      constexpr uint32_t b_kstep = B_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc_stage = 0;
      uint32_t acc_phase = 0;
      int qi = 0;
      for (int tile = worker; tile >= 0; tile = next_tile(tile, qi, true)) {
        const int split = tile / tiles_per_split;
        const int kb_begin = split * k_blocks_per_split;
        const int kb_end = min(k_blocks_total, kb_begin + k_blocks_per_split);
        mbar_wait(&tmem_empty_bar[acc_stage], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc_stage * BN;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_a + stage * A_STAGE_BYTES), a_lbo, 1024);
          const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_b + stage * B_STAGE_BYTES), b_lbo, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint32_t acc = (kb > kb_begin || k > 0) ? 1u : 0u;
            if constexpr (NCTA == 2) umma_bf16_ss_2cta(tmem_d, a_desc + k * a_kstep, b_desc + k * b_kstep, idesc, acc);
            else umma_bf16_ss(tmem_d, a_desc + k * a_kstep, b_desc + k * b_kstep, idesc, acc);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs have read it
          if constexpr (NCTA == 2) umma_commit_2cta(&empty_bar[stage], 3);
          else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if constexpr (NCTA == 2) umma_commit_2cta(&tmem_full_bar[acc_stage], 3);
        else umma_commit(&tmem_full_bar[acc_stage]);
        if (++acc_stage == ACC_STAGES) { acc_stage = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    // Phase A: TMEM -> registers (one accumulator row per thread) -> this warp's 32x32 fp32 staging tile in
    //          shared memory (16-byte chunks XOR-swizzled by row, conflict-free both ways).
    // Phase B: read the tile back transposed so that a warp instruction covers 4 rows x 128 contiguous bytes;
    //          all epilogue math and every global access (bias, residual, pre-activation, outputs) happens
    //          here with fully coalesced 128-bit (fp32) / 64-bit (bf16) accesses.
    const int quad = warp & 3;   // TMEM lane quadrant this warp may access
    const int part = warp >> 2;  // this warp takes the 32-column chunks part, part + kParts, ... of the accumulator
    uint8_t* stg = smem_stage + warp * EPI_STAGE_BYTES;
    const int sub_row = lane >> 3;  // phase B: row within a group of 4
    const int ch = lane & 7;        // phase B: 16-byte chunk (4 fp32 columns)
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    int qi = 0;
    for (int tile = worker; tile >= 0; tile = next_tile_warp(tile, qi)) {
      const int split = tile / tiles_per_split;
      const int rem = tile - split * tiles_per_split;
      const int m_idx = (rem / n_blocks) * C::TILE_M + static_cast<int>(cta_rank) * BM;  // this CTA's accumulator rows
      const int n_idx = (rem % n_blocks) * BN;
      const int row_base = m_idx + quad * 32;
      // The auxiliary operand of a 32-column chunk (residual / GELU' / pos-emb rows): coalesced loads, one chunk AHEAD of
      // the chunk being processed (the first one before the accumulator is even complete), so that their DRAM latency
      // hides under the previous chunk instead of stalling every chunk -- the K = 1024 residual GEMMs were bound by
      // exactly that stall (tensor pipe 49 % active under ncu).
      auto load_aux = [&](int c, float4 (&auxf)[8], uint2 (&auxh)[8]) {
        const int col = n_idx + c * 32 + ch * 4;
        if constexpr (EPI == kEpiF32Resid || EPI == kEpiF32PosEmbed || EPI == kEpiBf16DGelu) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = row_base + 4 * i + sub_row;
            if (row < p.M && col < p.N) {
              if constexpr (EPI == kEpiF32Resid) {
                auxf[i] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.aux) +
                                                           static_cast<long long>(row) * p.ldaux + col);
              } else if constexpr (EPI == kEpiF32PosEmbed) {
                const int pidx = row % p.aux_int;
                auxf[i] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.aux) +
                                                                static_cast<long long>(1 + pidx) * p.ldaux + col));
              } else {
                auxh[i] = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.aux) +
                                                               static_cast<long long>(row) * p.ldaux + col));
              }
            }
          }
        }
      };
      auto do_chunk = [&](int c, float4 (&auxf)[8], uint2 (&auxh)[8]) {
        const int col0 = n_idx + c * 32;
        const int col = col0 + ch * 4;
        (void)col0;
        {
          uint32_t r[32];
          const uint32_t taddr = tmem_base + acc_stage * BN + c * 32 + (static_cast<uint32_t>(quad * 32) << 16);
          tmem_ld_32x32b_x32(taddr, r);
          tmem_ld_wait();
          const uint32_t stg_wr = smem_u32(stg) + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_wr + ((j ^ (lane & 7)) << 4)), "r"(r[4 * j]),
                         "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                         : "memory");
        }
        __syncwarp();
        if (col < p.N) {
          float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
          float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
          if constexpr (EPI != kEpiF32Atomic && EPI != kEpiBf16DGelu) {
            if (p.bias != nullptr) bias = __ldg(reinterpret_cast<const float4*>(p.bias + col));
          }
          // Row pointers advance by 4 rows per step (no per-row 64-bit multiplies in the loop).
          constexpr int kOutElem = (EPI == kEpiBf16 || EPI == kEpiBf16Gelu || EPI == kEpiBf16DGelu) ? 2 : 4;
          const int row0 = row_base + sub_row;
          uint8_t* po = reinterpret_cast<uint8_t*>(p.out) + (static_cast<long long>(row0) * p.ldo + col) * kOutElem;
          const long long po_step = 4 * p.ldo * kOutElem;
          uint8_t* po2 = nullptr;
          long long po2_step = 0;
          if constexpr (EPI == kEpiBf16Gelu) {
            if (p.out2 != nullptr) {
              po2 = reinterpret_cast<uint8_t*>(p.out2) + (static_cast<long long>(row0) * p.ldo2 + col) * 2;
              po2_step = 8 * p.ldo2;
            }
          }
          const uint32_t stg_rd = smem_u32(stg) + sub_row * 128;
          // Branch-free over the 8 rows (only the stores are predicated) so that their dependency chains interleave:
          // a per-row `continue` serialised the rows and left the epilogue latency-bound.
          auto rows = [&](auto save_grad_tag) {
            constexpr bool kSaveGrad = decltype(save_grad_tag)::value;
            // all eight staged rows are fetched before the math starts: the shared-memory latency is paid once per chunk
            // instead of once per row (ncu: the first FADD of every row sat on the short scoreboard)
            float4 av[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 4 * i + sub_row;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(av[i].x), "=f"(av[i].y), "=f"(av[i].z), "=f"(av[i].w)
                           : "r"(stg_rd + i * 512 + ((ch ^ (rr & 7)) << 4)));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = row0 + 4 * i;
              const bool ok = row < p.M;
              uint8_t* o = po + i * po_step;
              float4 a = av[i];
              a.x += bias.x; a.y += bias.y; a.z += bias.z; a.w += bias.w;
              if constexpr (EPI == kEpiBf16) {
                const uint2 w = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
                if (ok) *reinterpret_cast<uint2*>(o) = w;
                if (p.colsum != nullptr && ok) {
                  csum.x += bf16_lo(w.x); csum.y += bf16_hi(w.x); csum.z += bf16_lo(w.y); csum.w += bf16_hi(w.y);
                }
              } else if constexpr (EPI == kEpiBf16Gelu) {
                // The reference evaluates GELU on the bf16-rounded fc1 output (autocast). What the
                // backward needs from this layer is only GELU'(pre), so that is what out2 keeps (bf16), not pre itself.
                const uint32_t p0 = pack_bf16x2(a.x, a.y), p1 = pack_bf16x2(a.z, a.w);
                float y0, y1, y2, y3;
                if constexpr (kSaveGrad) {
                  float g0, g1, g2, g3;
                  gelu_and_grad_fast(bf16_lo(p0), y0, g0); gelu_and_grad_fast(bf16_hi(p0), y1, g1);
                  gelu_and_grad_fast(bf16_lo(p1), y2, g2); gelu_and_grad_fast(bf16_hi(p1), y3, g3);
                  if (ok) *reinterpret_cast<uint2*>(po2 + i * po2_step) = make_uint2(pack_bf16x2(g0, g1), pack_bf16x2(g2, g3));
                } else {
                  y0 = gelu_fast(bf16_lo(p0)); y1 = gelu_fast(bf16_hi(p0)); y2 = gelu_fast(bf16_lo(p1)); y3 = gelu_fast(bf16_hi(p1));
                }
                if (ok) *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
              } else if constexpr (EPI == kEpiBf16DGelu) {
                // aux = GELU'(pre) saved by the forward epilogue; autograd's product is rounded to bf16 once more.
                const uint2 gp = ok ? auxh[i] : make_uint2(0u, 0u);
                const uint32_t d0 = pack_bf16x2(a.x, a.y), d1 = pack_bf16x2(a.z, a.w);
                const uint2 w = make_uint2(pack_bf16x2(bf16_lo(d0) * bf16_lo(gp.x), bf16_hi(d0) * bf16_hi(gp.x)),
                                           pack_bf16x2(bf16_lo(d1) * bf16_lo(gp.y), bf16_hi(d1) * bf16_hi(gp.y)));
                if (ok) *reinterpret_cast<uint2*>(o) = w;
                csum.x += bf16_lo(w.x); csum.y += bf16_hi(w.x); csum.z += bf16_lo(w.y); csum.w += bf16_hi(w.y);
              } else if constexpr (EPI == kEpiF32Resid) {
                // bf16 GEMM output added to the fp32 residual stream (Appendix B).
                const float4 x = auxf[i];
                if (!p.exact) { a.x = round_bf16(a.x); a.y = round_bf16(a.y); a.z = round_bf16(a.z); a.w = round_bf16(a.w); }
                if (ok) *reinterpret_cast<float4*>(o) = make_float4(a.x + x.x, a.y + x.y, a.z + x.z, a.w + x.w);
              } else if constexpr (EPI == kEpiF32Gelu) {
                if (ok) *reinterpret_cast<float4*>(o) = make_float4(gelu_erf(a.x), gelu_erf(a.y), gelu_erf(a.z), gelu_erf(a.w));
              } else if constexpr (EPI == kEpiF32) {
                if (ok) *reinterpret_cast<float4*>(o) = a;
              } else if constexpr (EPI == kEpiF32Atomic) {
                if (ok)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w)
                               : "memory");
              } else if constexpr (EPI == kEpiF32PosEmbed) {
                if (ok) {
                  const int P = p.aux_int;
                  const int img = row / P, pidx = row - img * P;
                  const float4 x = auxf[i];
                  if (!p.exact) { a.x = round_bf16(a.x); a.y = round_bf16(a.y); a.z = round_bf16(a.z); a.w = round_bf16(a.w); }
                  *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) +
                                             (static_cast<long long>(img) * (P + 1) + 1 + pidx) * p.ldo + col) =
                      make_float4(a.x + x.x, a.y + x.y, a.z + x.z, a.w + x.w);
                }
              }
            }
          };
          if (po2 != nullptr) rows(std::true_type{});
          else rows(std::false_type{});
          if constexpr (EPI == kEpiBf16DGelu || EPI == kEpiBf16) {
            if (p.colsum != nullptr) {  // lanes with the same 4-column group (lane & 7) hold partial sums over 8 rows each
#pragma unroll
              for (int o = 8; o < 32; o <<= 1) {
                csum.x += __shfl_xor_sync(0xffffffffu, csum.x, o);
                csum.y += __shfl_xor_sync(0xffffffffu, csum.y, o);
                csum.z += __shfl_xor_sync(0xffffffffu, csum.z, o);
                csum.w += __shfl_xor_sync(0xffffffffu, csum.w, o);
              }
              if (lane < 8) {
                atomicAdd(p.colsum + col + 0, csum.x); atomicAdd(p.colsum + col + 1, csum.y);
                atomicAdd(p.colsum + col + 2, csum.z); atomicAdd(p.colsum + col + 3, csum.w);
              }
            }
          }
        }
        __syncwarp();
      };
      float4 fa[8], fb[8];
      uint2 ha[8], hb[8];
      load_aux(part, fa, ha);
      mbar_wait(&tmem_full_bar[acc_stage], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = part; c < BN / 32; c += 2 * kParts) {
        const int c1 = c + kParts, c2 = c + 2 * kParts;
        if (c1 < BN / 32) load_aux(c1, fb, hb);
        do_chunk(c, fa, ha);
        if (c1 < BN / 32) {
          if (c2 < BN / 32) load_aux(c2, fa, ha);
          do_chunk(c1, fb, hb);
        }
      }
      // All TMEM reads of this warp have completed (wait::ld above): hand the accumulator stage back.
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (NCTA == 2) mbar_arrive_cluster(&tmem_empty_bar[acc_stage], 0);  // the leader owns the MMA pipeline
        else mbar_arrive(&tmem_empty_bar[acc_stage]);
      }
      if (++acc_stage == ACC_STAGES) { acc_stage = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if constexpr (NCTA == 2) cluster_sync_all();  // remote arrivals and the peer's smem reads have all landed
  else __syncthreads();
  tc_fence_after();
  if (warp == kMmaWarp) {
    if constexpr (NCTA == 2) tmem_dealloc_2cta(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// Few-tile problems (small-batch inference: M = 197 ... a few thousand rows) leave most of the 74 CTA pairs idle with
// 256 x 256 tiles -- a ViT-L out-projection at batch 8 has 28 tiles. They run the single-CTA configuration instead:
// 128 x 128 tiles, one per SM, 6-stage ring. It moves twice the operand bytes per FLOP through L2, so it only wins
// while the wide configuration cannot fill the machine: pick by estimated waves x work per wave.
inline bool prefer_small_tiles(int M, int N) {
  const int sms = device_sm_count();
  const long long wide = static_cast<long long>((M + 255) / 256) * ((N + BN_WIDE - 1) / BN_WIDE);
  const long long small = static_cast<long long>((M + BM - 1) / BM) * ((N + BN_SMALL - 1) / BN_SMALL);
  const long long pairs = sms / 2 > 0 ? sms / 2 : 1;
  const double cost_wide = 2.0 * static_cast<double>((wide + pairs - 1) / pairs);
  const double cost_small = 1.3 * static_cast<double>((small + sms - 1) / sms);
  return cost_small < cost_wide;
}

template <bool A_MN, bool B_MN, int EPI, int kNcta = 2, int BN = BN_WIDE>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  using C = Cfg<kNcta, epi_warps<EPI, BN>(), BN>;
  auto kern = gemm_bf16_tcgen05_kernel<A_MN, B_MN, EPI, kNcta, BN>;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kern), C::SMEM_BYTES, "gemm")) return rc;
  const int m_blocks = (p.M + C::TILE_M - 1) / C::TILE_M, n_blocks = (p.N + BN - 1) / BN;
  const int total = m_blocks * n_blocks * p.splits;
  const int max_workers = device_sm_count() / kNcta;
  // dynamic: one cluster per tile, the running ones take over the rest (see the kernel); TIC_GEMM_STATIC=1 keeps the
  // fixed persistent grid (development A/B)
  static const bool force_static = std::getenv("TIC_GEMM_STATIC") != nullptr;
  GemmParams pl = p;
  pl.dynamic = force_static ? 0 : 1;
  const int workers = (pl.dynamic || total < max_workers) ? total : max_workers;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(workers * kNcta);
  cfg.blockDim = dim3((epi_warps<EPI, BN>() + 2) * 32);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kNcta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, pl);
  if (e != cudaSuccess) return set_error(kErrCuda, "gemm: cudaLaunchKernelEx: %s", cudaGetErrorString(e));
  return check_launch("gemm_bf16_tcgen05");
}

}  // namespace

// A: K-major => [M, K] row-major with pitch lda; MN-major => [K, M] row-major with pitch lda.
// B: K-major => [N, K] row-major with pitch ldb; MN-major => [K, N] row-major with pitch ldb.
int gemm_bf16(const void* A, long long lda, bool a_mn, const void* B, long long ldb, bool b_mn, int M, int N, int K,
              int epilogue, void* out, long long ldo, void* out2, long long ldo2, const float* bias, const void* aux,
              long long ldaux, int aux_int, int splits, cudaStream_t stream, float* colsum, bool exact) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error(kErrInvalidArg, "gemm: empty problem %dx%dx%d", M, N, K);
  if (N % 8 != 0) return set_error(kErrInvalidArg, "gemm: N=%d must be a multiple of 8", N);
  if ((lda % 8) || (ldb % 8)) return set_error(kErrInvalidArg, "gemm: operand pitches must be multiples of 8 elements");
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 15))
    return set_error(kErrInvalidArg, "gemm: pointers must be 16-byte aligned");
  const int k_blocks = (K + BK - 1) / BK;
  if (splits < 1) splits = 1;
  if (splits > k_blocks) splits = k_blocks;
  {  // make every split non-empty
    const int per = (k_blocks + splits - 1) / splits;
    splits = (k_blocks + per - 1) / per;
  }
  if (splits > 1 && epilogue != kEpiF32Atomic)
    return set_error(kErrInvalidArg, "gemm: split-K requires the atomic fp32 epilogue");

  CUtensorMap ta, tb;
  int rc;
  if (!a_mn) rc = encode_tmap_2d_bf16(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, BM);
  else       rc = encode_tmap_2d_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK);
  if (rc) return rc;
  static_assert(Cfg<2, 8, BN_WIDE>::B_ROWS == 128 && Cfg<1, 8, BN_SMALL>::B_ROWS == 128, "both configurations stage 128 B rows per CTA");
  if (!b_mn) rc = encode_tmap_2d_bf16(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, 128);
  else       rc = encode_tmap_2d_bf16(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK);
  if (rc) return rc;

  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.splits = splits;
  p.out = out; p.ldo = ldo; p.out2 = out2; p.ldo2 = ldo2;
  p.bias = bias; p.aux = aux; p.ldaux = ldaux; p.aux_int = aux_int; p.colsum = colsum; p.exact = exact ? 1 : 0;

  static const char* kNames[2][2][8] = {
      {{"gemm_fwd_bf16", "gemm_fwd_gelu", "gemm_fwd_resid", "gemm_fwd_x", "gemm_fwd_f32", "gemm_fwd_x", "gemm_fwd_posemb", "gemm_fwd_gelu_f32"},
       {"gemm_dgrad_bf16", "gemm_dgrad_x", "gemm_dgrad_x", "gemm_dgrad_dgelu", "gemm_dgrad_f32", "gemm_dgrad_x", "gemm_dgrad_x", "gemm_dgrad_x"}},
      {{"gemm_x", "gemm_x", "gemm_x", "gemm_x", "gemm_x", "gemm_x", "gemm_x", "gemm_x"},
       {"gemm_wgrad_x", "gemm_wgrad_x", "gemm_wgrad_x", "gemm_wgrad_x", "gemm_wgrad_f32", "gemm_wgrad_atomic", "gemm_wgrad_x", "gemm_wgrad_x"}}};
  ProfScope prof(epilogue >= 0 && epilogue < 8 ? kNames[a_mn][b_mn][epilogue] : "gemm_x", 2.0 * M * N * K,
                 2.0 * (static_cast<double>(M) * K + static_cast<double>(N) * K) +
                     static_cast<double>(M) * N * ((epilogue == kEpiBf16 || epilogue == kEpiBf16DGelu) ? 2 : 4),
                 stream);
#define TIC_GEMM_CASE(AMN, BMN, E) \
  if (a_mn == AMN && b_mn == BMN && epilogue == E) return launch<AMN, BMN, E>(ta, tb, p, stream);
#define TIC_GEMM_CASE_SMALL(E) \
  if (!a_mn && !b_mn && epilogue == E && small) return launch<false, false, E, 1, BN_SMALL>(ta, tb, p, stream);
  // forward of a few-tile problem (K-major x K-major, single-CTA 128 x 128 tiles)
  const bool small = splits == 1 && prefer_small_tiles(M, N);
  TIC_GEMM_CASE_SMALL(kEpiBf16)
  TIC_GEMM_CASE_SMALL(kEpiBf16Gelu)
  TIC_GEMM_CASE_SMALL(kEpiF32Resid)
  TIC_GEMM_CASE_SMALL(kEpiF32PosEmbed)
#undef TIC_GEMM_CASE_SMALL
  // forward (K-major x K-major)
  TIC_GEMM_CASE(false, false, kEpiBf16)
  TIC_GEMM_CASE(false, false, kEpiBf16Gelu)
  TIC_GEMM_CASE(false, false, kEpiF32Resid)
  TIC_GEMM_CASE(false, false, kEpiF32)
  TIC_GEMM_CASE(false, false, kEpiF32PosEmbed)
  TIC_GEMM_CASE(false, false, kEpiF32Gelu)
  // dgrad (K-major x MN-major)
  TIC_GEMM_CASE(false, true, kEpiBf16)
  TIC_GEMM_CASE(false, true, kEpiBf16DGelu)
  TIC_GEMM_CASE(false, true, kEpiF32)
  // wgrad (MN-major x MN-major)
  TIC_GEMM_CASE(true, true, kEpiF32)
  TIC_GEMM_CASE(true, true, kEpiF32Atomic)
#undef TIC_GEMM_CASE
  return set_error(kErrUnsupported, "gemm: no kernel for a_mn=%d b_mn=%d epilogue=%d", (int)a_mn, (int)b_mn, epilogue);
}

}  // namespace tic
